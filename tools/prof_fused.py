"""cProfile of the fused population step on the host (are the small-population shards launch-bound?)."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    pop = int(os.environ.get("B2RL_BENCH_POP", "1"))
    bench.BUFFER = 8192
    agents, mem, nmem = bench.build_rank("cuda:0", pop, 0)
    for _ in range(20):
        bench.fused_population_step(agents, mem, nmem)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        bench.fused_population_step(agents, mem, nmem)
    t_enqueue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_total = time.perf_counter() - t0
    print(f"pop={pop}: host enqueue {t_enqueue / 200 / pop * 1e6:.0f} us per agent-step, "
          f"wall {t_total / 200 / pop * 1e6:.0f} us per agent-step")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(200):
        bench.fused_population_step(agents, mem, nmem)
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)


if __name__ == "__main__":
    main()
