"""Numbers behind tests/test_maddpg_gpu.py::test_learn_matches_reference_golden, printed as one JSON object (run on the
GPU box): loss differences per learn call, the worst gradient tensor of the first call relative to its largest
element, and the parameter differences after the three calls (per network kind: max, 99.9th percentile)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_maddpg_gpu as T  # noqa: E402
from conftest import load_golden  # noqa: E402

g = load_golden("maddpg_vector.npz")
ids, agent = T._agent(g)
out = {"loss_abs_diff": [], "grad_rel": {}, "param": {}}
for st in range(int(g["steps"])):
    losses = agent.learn(T._batch(g, st, ids))
    out["loss_abs_diff"].append(max(abs(losses[a][j] - float(g[f"s{st}_{n}/{a}"])) for a in ids
                                    for j, n in enumerate(("actor_loss", "critic_loss"))))
    if st == 0:
        for group, nets, opts in (("critic", agent.critics, agent.critic_optimizers), ("actor", agent.actors, agent.actor_optimizers)):
            worst = (0.0, "")
            for a in ids:
                for key, e in nets[a].layout.entries.items():
                    ref = torch.from_numpy(g[f"s0_grad/{group}/{a}/{key}"].copy())
                    got = opts[a].grads[e.offset:e.offset + ref.numel()].view(ref.shape).cpu()
                    rel = float((got - ref).abs().max()) / max(float(ref.abs().max()), 1e-12)
                    worst = max(worst, (rel, f"{a}/{key}"))
            out["grad_rel"][group] = worst
for tag, nets in (("actor1", agent.actors), ("actor_target1", agent.actor_targets), ("critic1", agent.critics),
                  ("critic_target1", agent.critic_targets)):
    d = torch.cat([(nets[a].state_dict()[k].cpu() - ref).abs().reshape(-1) for a in ids for k, ref in T._sd(g, f"{tag}/{a}").items()])
    out["param"][tag] = {"max": float(d.max()), "p999": float(d.quantile(0.999)), "median": float(d.median()), "n": d.numel()}
print(json.dumps(out))
