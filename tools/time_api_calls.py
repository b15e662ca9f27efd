"""Host wall time of every call of the public-API learn step (what bench.py's e2e times), per agent-step."""
import collections
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from agilerl_b200.components import Transition  # noqa: E402


def main():
    bench.BUFFER = 16384
    dev = "cuda:0"
    agents, mem, nmem = bench.build_rank(dev, 2, 0)
    mem.device_rng = False
    g = torch.Generator().manual_seed(1)
    E = bench.NUM_ENVS
    host_tr = {
        "obs": torch.randint(0, 256, (E, *bench.OBS), dtype=torch.uint8, generator=g).pin_memory(),
        "action": torch.randint(0, bench.N_ACT, (E,), generator=g).float().pin_memory(),
        "next_obs": torch.randint(0, 256, (E, *bench.OBS), dtype=torch.uint8, generator=g).pin_memory(),
        "reward": torch.randn(E, generator=g).pin_memory(),
        "done": (torch.rand(E, generator=g) < 0.01).float().pin_memory(),
    }
    acc = collections.OrderedDict()

    gpu_ev = []

    def timed(name, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        out = fn()
        acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
        e1.record()
        gpu_ev.append((name, e0, e1))
        return out

    def step(record):
        for agent in agents:
            t = timed if record else (lambda n, f: f())
            td = t("Transition", lambda: Transition(obs=host_tr["obs"], action=host_tr["action"], reward=host_tr["reward"],
                                                    next_obs=host_tr["next_obs"], done=host_tr["done"],
                                                    batch_size=[E]).to_tensordict())
            one = t("nmem.add", lambda: nmem.add(td))
            if one is not None:
                t("mem.add", lambda: mem.add(one))
            exp = t("mem.sample", lambda: mem.sample(agent.batch_size, agent.beta))
            nexp = t("nmem.sample_from_indices", lambda: nmem.sample_from_indices(exp["idxs"].squeeze(1)))
            exp["weights"] = exp["weights"].squeeze(1)
            loss, idxs, pri = t("agent.learn", lambda: agent.learn(exp, n_experiences=nexp, per=True))
            t("mem.update_priorities", lambda: mem.update_priorities(idxs, pri))

    for _ in range(10):
        step(False)
    torch.cuda.synchronize()
    n = 100
    t0 = time.perf_counter()
    for _ in range(n):
        step(True)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    per = n * len(agents)
    print(f"wall {wall / per * 1e6:.0f} us per agent-step")
    for k, v in acc.items():
        print(f"  {k:28s} {v / per * 1e6:8.1f} us")
    e = agents[0].engine
    gacc = collections.OrderedDict()
    for name, e0, e1 in gpu_ev:
        gacc[name] = gacc.get(name, 0.0) + e0.elapsed_time(e1)
    print("stream time between the events around each call (ms -> us per agent-step):")
    for k, v in gacc.items():
        print(f"  {k:28s} {v / per * 1e3:8.1f} us")
    print("api plans:", len(e._api_plans), "replays:", sum(p.seen for p in e._api_plans.values()))


if __name__ == "__main__":
    main()
