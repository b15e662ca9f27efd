"""Per-role timeline of one CTA of the tensor-core forward convolution (clock64 stamps recorded by the
kernel when B2RL_TC_DBG=<cta> is set):  B2RL_TC_DBG=300 python tools/conv_timeline.py"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from agilerl_b200 import _lib  # noqa: E402


def main():
    bench.BUFFER = 8192
    dev = "cuda:0"
    agents, mem, nmem = bench.build_rank(dev, 1, 0)
    eng = agents[0].engine
    lib = _lib.load()
    desc = eng.layout.desc
    L = desc.enc[0]
    B = bench.B
    out = torch.empty(B * L.out_c * L.out_h * L.out_w, dtype=torch.float32, device=dev)
    ws = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
    frames = nmem._fields[("obs",)]
    stream = _lib.stream_ptr(torch.device(dev))
    for it in range(5):
        idx = torch.randint(0, bench.BUFFER, (B,), device=dev)
        _lib.check(lib.b2rl_encoder_layer_forward(ctypes.byref(desc), 0, eng.actor.params.data_ptr(), frames.data_ptr(),
                                                  idx.data_ptr(), B, out.data_ptr(), ws.data_ptr(), ws.numel(), stream))
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 128)()
    _lib.check(lib.b2rl_debug_read(buf, 128))
    t = list(buf)
    t0 = t[0]
    rel = lambda x: x - t0 if x else None
    print("producer warp 0 (cycles since CTA entry):")
    print(f"  setup done {rel(t[1])}, prologue gather issued {rel(t[2])}")
    for kb in range(8):
        print(f"  kb{kb}: stage free {rel(t[3 + 3 * kb])}  converted+stored {rel(t[4 + 3 * kb])}  arrived {rel(t[5 + 3 * kb])}")
    print(f"  all MMAs retired {rel(t[60])}, epilogue done {rel(t[61])}")
    print("MMA lane:")
    for kb in range(8):
        print(f"  kb{kb}: weights landed {rel(t[64 + 3 * kb])}  im2col full {rel(t[65 + 3 * kb])}  MMAs+commit issued {rel(t[66 + 3 * kb])}")


if __name__ == "__main__":
    main()
