"""Small driver for ncu captures: runs the first conv layer forward (dominant kernel) and a few
fused learn steps on a reduced ring, so a --set full capture stays short."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "conv1"
    bench.BUFFER = 8192
    dev = "cuda:0"
    engines, mem, nmem = bench.build_rank(dev, 1, 0)
    support = torch.linspace(bench.V_MIN, bench.V_MAX, bench.N_ATOMS).to(dev)
    if mode == "conv1":
        for _ in range(3):
            print(bench.conv1_roofline(engines[0].engine, nmem, dev))
    else:
        for _ in range(4):
            bench.fused_population_step(engines, mem, nmem, support)
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
