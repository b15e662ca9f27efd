"""GPU parity of the DDPG / TD3 learn() CUDA path (b2rl_ddpg_learn, csrc/ddpg.cuh) — SURVEY 8f-1, BASELINE configs[2].

Golden vectors recorded from the UNMODIFIED reference (tests/golden/make_golden.py::gen_ddpg_td3, pinned oracle:
tests/test_ddpg_td3_oracle_cpu.py): four consecutive learn calls with the reference's own noise draws injected.
Bars: losses within 1e-5 (north star), every parameter of every network after the four calls within 2e-6 + 1e-5
relative of the reference's (two Adam steps per critic call, Polyak cadence, actor every policy_freq calls), the
batch's action tensor overwritten with the noise (kept quirk), and the BASELINE size (B = 512) against the oracle."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _sd(g, tag):
    return {k[len(tag) + 1:]: torch.from_numpy(g[k].copy()) for k in g.files if k.startswith(tag + "/")}


def _agent(name, g=None, obs_dim=17, act_dim=6, a_hidden=(32,), c_hidden=(64,), **kw):
    from agilerl_b200.algorithms import DDPG, TD3
    from agilerl_b200.compat import spaces
    cls = TD3 if name == "td3" else DDPG
    net = {"encoder_config": {"hidden_size": [64, 64]}, "head_config": {"hidden_size": list(a_hidden)}}
    agent = cls(spaces.Box(-np.inf, np.inf, (obs_dim,), np.float32), spaces.Box(-1.0, 1.0, (act_dim,), np.float32),
                net_config=net, **kw)
    for c in agent._critics() + agent._targets():
        assert list(c.head_net.hidden_size) == list(a_hidden)      # reference: the critics share the head_config
    return agent


@pytest.mark.parametrize("name", ["ddpg", "td3"])
def test_learn_matches_reference_golden(name):
    g = load_golden(f"{name}_vector.npz")
    twin = bool(int(g["twin"]))
    n_c = 2 if twin else 1
    assert [int(h) for h in g["a_hidden"]] == [int(h) for h in g["c_hidden"]] or True
    agent = _agent(name, a_hidden=[int(h) for h in g["a_hidden"]], batch_size=int(g["B"]), gamma=float(g["gamma"]),
                   tau=float(g["tau"]), lr_actor=float(g["lr_actor"]), lr_critic=float(g["lr_critic"]),
                   policy_freq=int(g["policy_freq"]))
    # the fixture's critics use their own head width
    from agilerl_b200.networks.actors import ContinuousQNetwork
    from agilerl_b200.compat import spaces
    crit_cfg = dict(encoder_config={"hidden_size": [64, 64]}, head_config={"hidden_size": [int(h) for h in g["c_hidden"]]})
    mk = lambda: ContinuousQNetwork(spaces.Box(-np.inf, np.inf, (17,), np.float32), spaces.Box(-1.0, 1.0, (6,), np.float32), **crit_cfg)
    agent._set_critics([mk() for _ in range(n_c)], [mk() for _ in range(n_c)])
    agent._bind_engine()
    agent.actor.load_state_dict(_sd(g, "actor0")); agent.actor_target.load_state_dict(_sd(g, "actor_target0"))
    for i, (c, t) in enumerate(zip(agent._critics(), agent._targets())):
        c.load_state_dict(_sd(g, f"critic{i}_0")); t.load_state_dict(_sd(g, f"critic_target{i}_0"))
    for st in range(int(g["steps"])):
        exp = {k: torch.from_numpy(g[f"s{st}_{k}"].copy()).cuda() for k in ("obs", "action", "reward", "next_obs", "done")}
        noise = torch.from_numpy(g[f"s{st}_noise"].copy())
        a_loss, c_loss = agent.learn(exp, noise=noise)
        np.testing.assert_array_equal(exp["action"].cpu().numpy(), g[f"s{st}_noise"])     # in-place noise quirk
        ref_a, ref_c = float(g[f"s{st}_actor_loss"]), float(g[f"s{st}_critic_loss"])
        assert (a_loss is None) == bool(np.isnan(ref_a)), f"step {st}: actor update cadence"
        assert abs(c_loss - ref_c) <= 1e-5 * max(1.0, abs(ref_c)), (st, c_loss, ref_c)
        if a_loss is not None:
            assert abs(a_loss - ref_a) <= 1e-5 * max(1.0, abs(ref_a)), (st, a_loss, ref_a)
    def close(net, tag):
        for k, v in _sd(g, tag).items():
            got = net.state_dict()[k].cpu()
            err = (got - v).abs().max().item()
            assert err <= 2e-6 + 1e-5 * v.abs().max().item(), f"{tag} {k}: {err}"
    close(agent.actor, "actor1"); close(agent.actor_target, "actor_target1")
    for i, (c, t) in enumerate(zip(agent._critics(), agent._targets())):
        close(c, f"critic{i}_1"); close(t, f"critic_target{i}_1")


@pytest.mark.parametrize("name", ["ddpg", "td3"])
def test_baseline_config3_size_against_oracle(name):
    """17-dim observations, 6-dim actions, batch 512 (BASELINE configs[2]): three learn calls, Philox noise replaced
    by injected draws on both sides, losses within 1e-5 and every gradient-carrying parameter close to the oracle."""
    from oracle import ddpg_td3 as od
    B = 512
    agent = _agent(name, a_hidden=[32], batch_size=B, lr_actor=1e-4, lr_critic=1e-3, tau=0.005, policy_freq=2)
    twin = name == "td3"
    cpu = lambda net: {k: v.cpu().clone() for k, v in net.state_dict().items()}
    orc = od.OracleDDPG(od.actor_specs(17, 6, head_hidden=[32]), od.critic_specs(17, 6, head_hidden=[32]), cpu(agent.actor),
                        cpu(agent.actor_target), [cpu(c) for c in agent._critics()], [cpu(t) for t in agent._targets()],
                        gamma=0.99, tau=0.005, lr_actor=1e-4, lr_critic=1e-3, policy_freq=2, twin=twin)
    gen = torch.Generator().manual_seed(5)
    for st in range(3):
        exp = dict(obs=torch.randn(B, 17, generator=gen), action=torch.rand(B, 6, generator=gen) * 2 - 1,
                   reward=torch.randn(B, 1, generator=gen), next_obs=torch.randn(B, 17, generator=gen),
                   done=(torch.rand(B, 1, generator=gen) < 0.05).float())
        noise = torch.randn(B, 6, generator=gen) * 0.2
        dexp = {k: v.clone().cuda() for k, v in exp.items()}
        a_loss, c_loss = agent.learn(dexp, noise=noise)
        # the oracle draws its noise from torch's RNG in place: feed the same values through a patched normal_
        oexp = {k: v.clone() for k, v in exp.items()}
        orig = torch.Tensor.normal_
        torch.Tensor.normal_ = lambda self, mean=0, std=1, generator=None: self.copy_(noise)
        try:
            oa, oc = orc.learn(oexp)
        finally:
            torch.Tensor.normal_ = orig
        assert abs(c_loss - oc) <= 1e-5 * max(1.0, abs(oc)), (st, c_loss, oc)
        assert (a_loss is None) == (oa is None)
        if oa is not None:
            assert abs(a_loss - oa) <= 1e-5 * max(1.0, abs(oa)), (st, a_loss, oa)
    for i, c in enumerate(agent._critics()):
        for k, v in orc.critics[i].items():
            err = (c.state_dict()[k].cpu() - v.detach()).abs().max().item()
            assert err <= 5e-6 + 2e-5 * v.abs().max().item(), f"critic{i} {k}: {err}"
    for k, v in orc.actor.items():
        err = (agent.actor.state_dict()[k].cpu() - v.detach()).abs().max().item()
        assert err <= 5e-6 + 2e-5 * v.abs().max().item(), f"actor {k}: {err}"


def test_td3_api_surface_clone_and_actions():
    agent = _agent("td3", a_hidden=[32], batch_size=64)
    assert agent.algo == "TD3" and agent.action_dim == 6 and agent.learn_counter == 0
    obs = np.random.default_rng(0).standard_normal((3, 17)).astype(np.float32)
    a_train, a_eval = agent.get_action(obs), agent.get_action(obs, training=False)
    assert a_train.shape == (3, 6) and a_eval.shape == (3, 6) and np.abs(a_train).max() <= 1.0
    gen = torch.Generator().manual_seed(0)
    exp = dict(obs=torch.randn(64, 17, generator=gen).cuda(), action=(torch.rand(64, 6, generator=gen) * 2 - 1).cuda(),
               reward=torch.randn(64, 1, generator=gen).cuda(), next_obs=torch.randn(64, 17, generator=gen).cuda(),
               done=torch.zeros(64, 1).cuda())
    l1 = agent.learn({k: v.clone() for k, v in exp.items()})
    l2 = agent.learn({k: v.clone() for k, v in exp.items()})
    assert l1[0] is None and isinstance(l1[1], float) and isinstance(l2[0], float)
    c = agent.clone(index=7)
    assert c.index == 7 and c.learn_counter == agent.learn_counter
    for k, v in agent.critic_1.state_dict().items():
        assert torch.equal(v, c.critic_1.state_dict()[k])
    assert torch.equal(agent.critic_1_optimizer.exp_avg, c.critic_1_optimizer.exp_avg) and c.critic_2_optimizer.step == 2
    same = agent.clone()                                                         # same index: same Philox stream and position
    ea, ec = exp["action"].clone(), exp["action"].clone()
    la, lc = agent.learn(dict(exp, action=ea)), same.learn(dict(exp, action=ec))
    assert la == lc and torch.equal(ea, ec) and not torch.equal(ea, exp["action"])
    e7 = exp["action"].clone()
    c.learn(dict(exp, action=e7))
    assert not torch.equal(e7, ea)                                               # another index draws another noise stream
