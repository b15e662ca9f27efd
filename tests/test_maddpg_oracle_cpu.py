"""CPU: the MADDPG / multi-agent replay oracle against the golden vectors recorded from the
unmodified reference (tests/golden/make_golden.py::gen_maddpg / gen_ma_replay): three consecutive ``MADDPG.learn``
calls (bit-exact losses and parameters, NaN reward / done handling) and seeded ``MultiAgentReplayBuffer.sample``
batches out of a wrapped buffer (bit-exact leaves)."""
import random

import numpy as np
import torch

from conftest import load_golden

FIELDS = ("obs", "action", "reward", "next_obs", "done")


def _sd(g, tag):
    return {k[len(tag) + 1:]: torch.from_numpy(g[k].copy()) for k in g.files if k.startswith(tag + "/")}


def _oracle(g):
    from oracle import maddpg as om
    ids = [str(a) for a in g["agent_ids"]]
    a_hidden, c_hidden = [int(h) for h in g["a_hidden"]], [int(h) for h in g["c_hidden"]]
    a_specs = {a: om.actor_specs(int(o), int(d), head_hidden=a_hidden) for a, o, d in zip(ids, g["obs_dims"], g["act_dims"])}
    c_head = om.critic_head_spec(int(g["act_dims"].sum()), head_hidden=c_hidden)
    orc = om.OracleMADDPG(ids, a_specs, c_head, {a: _sd(g, f"actor0/{a}") for a in ids},
                          {a: _sd(g, f"actor_target0/{a}") for a in ids}, {a: _sd(g, f"critic0/{a}") for a in ids},
                          {a: _sd(g, f"critic_target0/{a}") for a in ids}, gamma=float(g["gamma"]), tau=float(g["tau"]),
                          lr_actor=float(g["lr_actor"]), lr_critic=float(g["lr_critic"]))
    return ids, orc


def test_oracle_reproduces_reference_maddpg_learn_calls():
    g = load_golden("maddpg_vector.npz")
    ids, orc = _oracle(g)
    torch.set_num_threads(1)
    saw_nan = False
    for st in range(int(g["steps"])):
        exp = tuple({a: torch.from_numpy(g[f"s{st}_{f}/{a}"].copy()) for a in ids} for f in FIELDS)
        saw_nan |= any(torch.isnan(exp[2][a]).any() or torch.isnan(exp[4][a]).any() for a in ids)
        losses = orc.learn(exp)
        for a in ids:
            assert losses[a][0] == float(g[f"s{st}_actor_loss/{a}"]), (st, a)
            assert losses[a][1] == float(g[f"s{st}_critic_loss/{a}"]), (st, a)
    assert saw_nan, "the fixture must exercise the NaN reward / done path (maddpg.py:683-694)"
    for gname, nets in (("actor1", orc.actors), ("actor_target1", orc.actor_targets), ("critic1", orc.critics),
                        ("critic_target1", orc.critic_targets)):
        for a in ids:
            for k, v in _sd(g, f"{gname}/{a}").items():
                assert torch.equal(nets[a][k].data, v), (gname, a, k)


def test_oracle_replay_reproduces_reference_samples():
    from oracle import maddpg as om
    g = load_golden("ma_replay.npz")
    ids, fields = [str(a) for a in g["agent_ids"]], [str(f) for f in g["fields"]]
    orc = om.OracleMAReplay(int(g["cap"]), fields, ids)
    for t in range(int(g["n_steps"])):
        args = [{a: g[f"t{t}_{f}/{a}"] for a in ids} for f in fields]
        orc.save_to_memory(*args, is_vectorised=bool(int(g[f"t{t}_vect"])))
    assert len(orc) == int(g["final_len"]) == int(g["cap"]) and orc.counter == int(g["final_counter"])
    for c in range(int(g["n_samples"])):
        random.seed(int(g[f"sample{c}_seed"]))
        batch = orc.sample(int(g[f"sample{c}_B"]))
        for f, d in zip(fields, batch):
            for a in ids:
                assert d[a].dtype == torch.float32
                np.testing.assert_array_equal(d[a].numpy(), g[f"sample{c}_{f}/{a}"], err_msg=f"{c} {f} {a}")
