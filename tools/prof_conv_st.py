"""The staged convolution kernels alone at the benchmark shapes, for `ncu --set full -k regex:conv_.*_st`:
conv2 forward (512 and 256 rows), conv2 / conv1 weight gradient (256 rows), conv2 input gradient (256 rows)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from agilerl_b200 import _lib  # noqa: E402


def main():
    bench.BUFFER = 16384
    agents, mem, nmem = bench.build_rank("cuda:0", 1, 0)
    eng = agents[0].engine
    lib = _lib.load()
    desc = eng.layout.desc
    L0, L1 = desc.enc[0], desc.enc[1]
    B = bench.B
    s = _lib.stream_ptr(torch.device("cuda:0"))
    params = eng.actor.params
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    frames = nmem._fields[("obs",)]
    a1 = torch.randn(2 * B, L1.in_c, L1.in_h, L1.in_w, device="cuda").relu_()
    out2 = torch.empty(2 * B * L1.out_c * L1.out_h * L1.out_w, dtype=torch.float32, device="cuda")
    g2 = torch.randn(B, L1.out_c, L1.out_h, L1.out_w, device="cuda")
    g1 = torch.randn(B, L0.out_c, L0.out_h, L0.out_w, device="cuda")
    gin = torch.empty(B, L1.in_c, L1.in_h, L1.in_w, device="cuda")
    grads = torch.empty_like(params)
    reps = int(os.environ.get("REPS", "3"))
    for it in range(reps):
        idx = torch.randint(0, bench.BUFFER, (B,), device="cuda")
        _lib.check(lib.b2rl_encoder_layer_forward(ctypes.byref(desc), 1, params.data_ptr(), a1.data_ptr(), None, 2 * B,
                                                  out2.data_ptr(), ws.data_ptr(), ws.numel(), 0, s))
        _lib.check(lib.b2rl_encoder_layer_forward(ctypes.byref(desc), 1, params.data_ptr(), a1.data_ptr(), None, B,
                                                  out2.data_ptr(), ws.data_ptr(), ws.numel(), 0, s))
        _lib.check(lib.b2rl_encoder_layer_wgrad(ctypes.byref(desc), 1, a1.data_ptr(), None, B, g2.data_ptr(), grads.data_ptr(),
                                                ws.data_ptr(), ws.numel(), s))
        _lib.check(lib.b2rl_encoder_layer_wgrad(ctypes.byref(desc), 0, frames.data_ptr(), idx.data_ptr(), B, g1.data_ptr(),
                                                grads.data_ptr(), ws.data_ptr(), ws.numel(), s))
        _lib.check(lib.b2rl_encoder_layer_dgrad(ctypes.byref(desc), 1, params.data_ptr(), g2.data_ptr(), B, gin.data_ptr(),
                                                ws.data_ptr(), ws.numel(), s))
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
