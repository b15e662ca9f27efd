"""Three MADDPG learn calls at BASELINE configs[4]'s shapes (4 agents x 18-dim obs, batch 64) — the command profiled
under ncu for profiles/r2_maddpg_launches.*:
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file out.csv python tools/maddpg_one_call.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from agilerl_b200.algorithms import MADDPG  # noqa: E402
from agilerl_b200.compat import spaces  # noqa: E402

ids = [f"agent_{i}" for i in range(4)]
agent = MADDPG([spaces.Box(-np.inf, np.inf, (18,), np.float32) for _ in ids], [spaces.Box(-1.0, 1.0, (5,), np.float32) for _ in ids],
               agent_ids=ids, batch_size=64, lr_actor=1e-4, lr_critic=1e-3, tau=1e-3)
g = torch.Generator(device="cuda").manual_seed(0)
B = 64
exp = ({a: torch.randn(B, 18, device="cuda", generator=g) for a in ids}, {a: torch.rand(B, 5, device="cuda", generator=g) * 2 - 1 for a in ids},
       {a: torch.randn(B, 1, device="cuda", generator=g) for a in ids}, {a: torch.randn(B, 18, device="cuda", generator=g) for a in ids},
       {a: (torch.rand(B, 1, device="cuda", generator=g) < 0.1).float() for a in ids})
agent.use_graph = os.environ.get("GRAPH", "0") == "1"      # eager by default: ncu lists the kernels either way
for _ in range(3):
    out = agent.learn_device(exp)
torch.cuda.synchronize()
print("losses", out.tolist())
