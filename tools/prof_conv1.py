"""The roofline kernel alone (conv1 forward of B=256 ring rows, one weight set) for `ncu --set full`."""
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
    L = desc.enc[0]
    B = bench.B
    out = torch.empty(B * L.out_c * L.out_h * L.out_w, dtype=torch.float32, device="cuda")
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    frames = nmem._fields[("obs",)]
    s = _lib.stream_ptr(torch.device("cuda:0"))
    for it in range(4):
        idx = torch.randint(0, bench.BUFFER, (B,), device="cuda")
        _lib.check(lib.b2rl_encoder_layer_forward(ctypes.byref(desc), 0, eng.actor.params.data_ptr(), frames.data_ptr(),
                                                  idx.data_ptr(), B, out.data_ptr(), ws.data_ptr(), ws.numel(), int(it > 0), s))
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
