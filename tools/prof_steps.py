"""A few fused population steps on a small replay (for ncu launch lists / full captures).
usage: [B2RL_BENCH_POP=n] [B2RL_GRAPH=0] python tools/prof_steps.py [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    pop = int(os.environ.get("B2RL_BENCH_POP", "1"))
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    bench.BUFFER = 16384
    agents, mem, nmem = bench.build_rank("cuda:0", pop, 0)
    for _ in range(steps):
        bench.fused_population_step(agents, mem, nmem)
        for a in agents:
            a.synchronize()
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
