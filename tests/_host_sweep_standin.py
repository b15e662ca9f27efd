"""Helper of tests/test_reference_driver_cpu.py (own process): forty mutations of every kind over Rainbow DQN / DQN members on
image and vector observations, with the C entry points replaced by the stand-ins of tests/_refsuite_plugin.py.  After each:
target mirrors the evaluation network, the layer table is consistent, acting works, and the member survives a cross-rank
move (export_state -> pickle -> from_state) and a checkpoint round trip with its mutated architecture and hyper-parameters."""
import json
import os
import pickle
import sys
import tempfile
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["B2RL_GRAPH"] = "0"
import numpy as np  # noqa: E402
import torch  # noqa: E402

import _refsuite_plugin  # noqa: E402,F401  (stand-ins for the C entry points)
from agilerl_b200 import _lib
from agilerl_b200.algorithms import DQN, RainbowDQN
from agilerl_b200.algorithms.core.registry import HyperparameterConfig, RLParameter
from agilerl_b200.compat import spaces
from agilerl_b200.hpo import Mutations
hp = HyperparameterConfig(lr=RLParameter(min=1e-5, max=1e-2), batch_size=RLParameter(min=8, max=64, dtype=int), learn_step=RLParameter(min=1, max=10, dtype=int))
NET = {"encoder_config": {"channel_size": [8, 16], "kernel_size": [4, 3], "stride_size": [2, 1]}, "head_config": {"hidden_size": [32]}, "latent_dim": 16}
cases = [(RainbowDQN, spaces.Box(0, 255, (3, 20, 20), np.uint8), dict(net_config=dict(NET), v_min=-10.0, v_max=10.0)),
         (RainbowDQN, spaces.Box(-1, 1, (6,), np.float32), dict(v_min=-10.0, v_max=10.0)),
         (DQN, spaces.Box(0, 255, (3, 20, 20), np.uint8), dict(net_config=dict(NET))), (DQN, spaces.Box(-1, 1, (6,), np.float32), {})]
results = []
for cls, osp, kw in cases:
    a = cls(osp, spaces.Discrete(4), batch_size=8, hp_config=hp, **kw)
    m = Mutations(0.1, 0.4, 0.3, 0.2, 0.2, 0.2, rand_seed=5, device="cuda")
    seen = {}
    for it in range(40):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            [a] = m.mutation([a.clone()])
        seen[a.mut] = seen.get(a.mut, 0) + 1
        sd, sdt = a.actor.state_dict(), a.actor_target.state_dict()
        assert list(sd) == list(sdt) and all(sd[k].shape == sdt[k].shape for k in sd), a.mut
        d = a.actor.layout.desc
        for layers, n in ((d.enc, d.n_enc), (d.val, d.n_val), (d.adv, d.n_adv)):
            for i in range(1, n):
                if layers[i].kind == 1 and layers[i - 1].kind == 1: assert layers[i].in_c == layers[i - 1].out_c, (a.mut, i)
        meta, tensors = a.export_state(); b = cls.from_state(pickle.loads(pickle.dumps(meta)), [t.clone() for t in tensors], a)
        assert torch.equal(b.actor.buffers.params, a.actor.buffers.params) and b.batch_size == a.batch_size and b.lr == a.lr, a.mut
        p = os.path.join(tempfile.mkdtemp(), "c.pt"); a.save_checkpoint(p)
        c = cls(osp, spaces.Discrete(4), batch_size=8, hp_config=hp, **kw); c.load_checkpoint(p)
        assert torch.equal(c.actor.buffers.params, a.actor.buffers.params) and c.lr == a.lr and c.batch_size == a.batch_size and c.learn_step == a.learn_step, a.mut
        obs = np.zeros((2, *osp.shape), dtype=osp.dtype)
        act = a.get_action(obs)
        assert act.shape == (2,)
    results.append({"cls": cls.__name__, "obs": list(osp.shape), "seen": seen})
print("RESULT " + json.dumps(results))
