"""cProfile of the public-API (e2e) population step: where the host time goes."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    bench.BUFFER = 8192
    dev = "cuda:0"
    agents, mem, nmem = bench.build_rank(dev, 2, 0)
    mem.device_rng = False
    g = torch.Generator().manual_seed(1)
    host_tr = {
        "obs": torch.randint(0, 256, (bench.NUM_ENVS, *bench.OBS), dtype=torch.uint8, generator=g).pin_memory(),
        "action": torch.randint(0, bench.N_ACT, (bench.NUM_ENVS,), generator=g).float().pin_memory(),
        "next_obs": torch.randint(0, 256, (bench.NUM_ENVS, *bench.OBS), dtype=torch.uint8, generator=g).pin_memory(),
        "reward": torch.randn(bench.NUM_ENVS, generator=g).pin_memory(),
        "done": (torch.rand(bench.NUM_ENVS, generator=g) < 0.01).float().pin_memory(),
    }
    for _ in range(5):
        bench.api_population_step(agents, mem, nmem, None, host_tr)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(50):
        bench.api_population_step(agents, mem, nmem, None, host_tr)
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(40)


if __name__ == "__main__":
    main()
