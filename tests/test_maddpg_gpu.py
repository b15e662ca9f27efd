"""GPU parity of the multi-agent path (BASELINE configs[4]): ``MultiAgentReplayBuffer`` over HBM rings,
``MADDPG.learn`` as ``b2rl_maddpg_learn`` (csrc/maddpg.cuh) and the on-device Gaussian parameter mutation
(``b2rl_gaussian_mutate``).

Golden vectors recorded from the UNMODIFIED reference (tests/golden/make_golden.py::gen_maddpg / gen_ma_replay, pinned
oracle: tests/test_maddpg_oracle_cpu.py).  Bars: sampled leaves bit-exact; losses within 1e-5 (north star) on each of three
consecutive learn calls (a NaN reward / done rides in the second batch); every gradient tensor of the first call within
2e-5 of its largest element; parameters after the three calls within Adam's reach of the reference's."""
import random

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

FIELDS = ("obs", "action", "reward", "next_obs", "done")


def _sd(g, tag):
    return {k[len(tag) + 1:]: torch.from_numpy(g[k].copy()) for k in g.files if k.startswith(tag + "/")}


def _agent(g, **kw):
    from agilerl_b200.algorithms import MADDPG
    from agilerl_b200.compat import spaces
    ids = [str(a) for a in g["agent_ids"]]
    obs = [spaces.Box(-1.0, 1.0, (int(d),), np.float32) for d in g["obs_dims"]]
    act = [spaces.Box(-1.0, 1.0, (int(d),), np.float32) for d in g["act_dims"]]
    net = {"head_config": {"hidden_size": [int(h) for h in g["a_hidden"]]}}
    assert [int(h) for h in g["a_hidden"]] == [int(h) for h in g["c_hidden"]]
    agent = MADDPG(obs, act, agent_ids=ids, net_config=net, batch_size=int(g["B"]), gamma=float(g["gamma"]), tau=float(g["tau"]),
                   lr_actor=float(g["lr_actor"]), lr_critic=float(g["lr_critic"]), **kw)
    for a in ids:
        agent.actors[a].load_state_dict(_sd(g, f"actor0/{a}")); agent.actor_targets[a].load_state_dict(_sd(g, f"actor_target0/{a}"))
        agent.critics[a].load_state_dict(_sd(g, f"critic0/{a}")); agent.critic_targets[a].load_state_dict(_sd(g, f"critic_target0/{a}"))
    return ids, agent


def _batch(g, st, ids, device="cuda"):
    return tuple({a: torch.from_numpy(g[f"s{st}_{f}/{a}"].copy()).to(device) for a in ids} for f in FIELDS)


def test_state_dict_keys_are_the_references():
    g = load_golden("maddpg_vector.npz")
    ids, agent = _agent(g)
    a0 = ids[0]
    assert list(agent.actors[a0].state_dict()) == list(_sd(g, f"actor0/{a0}"))
    assert list(agent.critics[a0].state_dict()) == list(_sd(g, f"critic0/{a0}"))
    for k, v in _sd(g, f"critic0/{a0}").items():
        assert tuple(agent.critics[a0].state_dict()[k].shape) == tuple(v.shape), k


def test_learn_matches_reference_golden():
    g = load_golden("maddpg_vector.npz")
    ids, agent = _agent(g)
    for st in range(int(g["steps"])):
        losses = agent.learn(_batch(g, st, ids))
        for a in ids:
            for j, name in enumerate(("actor_loss", "critic_loss")):
                ref = float(g[f"s{st}_{name}/{a}"])
                assert abs(losses[a][j] - ref) <= 1e-5 * max(1.0, abs(ref)), (st, a, name, losses[a][j], ref)
        if st == 0:          # gradients of the first call, every tensor of every network
            for a in ids:
                for group, nets, opts in (("critic", agent.critics, agent.critic_optimizers), ("actor", agent.actors, agent.actor_optimizers)):
                    lay, grads = nets[a].layout, opts[a].grads
                    for key, e in lay.entries.items():
                        if e.buf != "param":
                            continue
                        ref_g = torch.from_numpy(g[f"s0_grad/{group}/{a}/{key}"].copy())
                        got = grads[e.offset:e.offset + ref_g.numel()].view(ref_g.shape).cpu()
                        tol = 2e-5 * max(float(ref_g.abs().max()), 1e-6)
                        assert float((got - ref_g).abs().max()) <= tol, (group, a, key, float((got - ref_g).abs().max()), tol)
    # parameters after three Adam steps.  Adam's step is lr * m / (sqrt(v) + eps): an element whose gradient nearly
    # cancels over the batch can move by a visible fraction of lr differently in two fp32 implementations, everything
    # else agrees to rounding.  Bar per network: no element further than 3 lr, 99.5 % of the elements within
    # 0.2 % of lr + 1e-5 relative (the DDPG / TD3 bar scaled to this learning rate).
    lr = {"actor": float(g["lr_actor"]), "critic": float(g["lr_critic"])}
    for tag, nets, kind in (("actor1", agent.actors, "actor"), ("actor_target1", agent.actor_targets, "actor"),
                            ("critic1", agent.critics, "critic"), ("critic_target1", agent.critic_targets, "critic")):
        for a in ids:
            sd = nets[a].state_dict()
            ref_sd = _sd(g, f"{tag}/{a}")
            d = torch.cat([(sd[k].cpu() - ref).abs().reshape(-1) for k, ref in ref_sd.items()])
            r = torch.cat([ref.abs().reshape(-1) for ref in ref_sd.values()])
            tight = d <= 2e-3 * lr[kind] + 1e-5 * r
            assert float(d.max()) <= 3 * lr[kind], (tag, a, float(d.max()))
            assert float(tight.float().mean()) >= 0.995, (tag, a, float(tight.float().mean()), float(d.max()))


@pytest.mark.parametrize("mode", ["graph+streams", "eager+streams", "graph+serial"])
def test_graph_replay_and_concurrent_agents_are_bit_identical_to_the_serial_eager_call(mode):
    """The captured learn call (one graph launch, bias corrections from the device step state) and the per-agent side
    streams change scheduling only: losses, every parameter, target and Adam moment equal the serial eager call's bit
    for bit over three steps (Adam's bias corrections differ at every step)."""
    g = load_golden("maddpg_vector.npz")
    ids, ref = _agent(g)
    ref.use_graph, ref.concurrent_agents = False, False
    ids, alt = _agent(g)
    alt.use_graph, alt.concurrent_agents = mode.startswith("graph"), mode.endswith("streams")
    for st in range(int(g["steps"])):
        l_ref, l_alt = ref.learn(_batch(g, st, ids)), alt.learn(_batch(g, st, ids))
        for a in ids:
            same = [x == y or (np.isnan(x) and np.isnan(y)) for x, y in zip(l_ref[a], l_alt[a])]
            assert all(same), (mode, st, a, l_ref[a], l_alt[a])
    if alt.use_graph:
        assert alt._plans[int(g["B"])].graph is not None
    for a in ids:
        for x, y in ((ref.actors, alt.actors), (ref.actor_targets, alt.actor_targets), (ref.critics, alt.critics),
                     (ref.critic_targets, alt.critic_targets)):
            assert torch.equal(x[a].buffers.params, y[a].buffers.params), (mode, a)
        assert torch.equal(ref.actor_optimizers[a].exp_avg_sq, alt.actor_optimizers[a].exp_avg_sq)
        assert torch.equal(ref.critic_optimizers[a].exp_avg, alt.critic_optimizers[a].exp_avg)


def test_replay_gathers_straight_into_the_captured_batch_buffers():
    from agilerl_b200.components import MultiAgentReplayBuffer
    g = load_golden("maddpg_vector.npz")
    ids, a1 = _agent(g)
    a2 = a1.clone()
    a2.use_graph = False
    B = 16
    buf = MultiAgentReplayBuffer(64, list(FIELDS), ids, device="cuda")
    buf.save_to_memory(*tuple({a: g[f"s0_{f}/{a}"] for a in ids} for f in FIELDS), is_vectorised=True)
    for _ in range(2):
        batch = buf.sample_device(B, out=a1.batch_buffers(B))
        assert batch[0].packed.data_ptr() == a1.batch_buffers(B)[0].data_ptr()
        plain = tuple({a: d[a].clone() for a in ids} for d in batch)
        l1, l2 = a1.learn(batch), a2.learn(plain)
        assert l1 == l2
    for a in ids:
        assert torch.equal(a1.critics[a].buffers.params, a2.critics[a].buffers.params)
        assert torch.equal(a1.actor_targets[a].buffers.params, a2.actor_targets[a].buffers.params)


def test_overlapped_population_learn_equals_member_by_member():
    """training.population.multi_agent_population_learn: every member on its own stream (draw, gather, graph launch)
    == the members one after another, bit for bit."""
    from agilerl_b200.components import MultiAgentReplayBuffer
    from agilerl_b200.training.population import multi_agent_population_learn
    g = load_golden("maddpg_vector.npz")
    ids, base = _agent(g)
    B = 16

    def run(overlap):
        torch.manual_seed(0)
        pop = [base.clone(index=k) for k in range(3)]
        for k, m in enumerate(pop):                      # distinct members
            for a in ids:
                m.actors[a].buffers.params.mul_(1.0 + 0.01 * k)
        buf = MultiAgentReplayBuffer(64, list(FIELDS), ids, device="cuda")
        buf.save_to_memory(*tuple({a: g[f"s0_{f}/{a}"] for a in ids} for f in FIELDS), is_vectorised=True)
        out = None
        for _ in range(3):
            out = multi_agent_population_learn(pop, buf, B, overlap=overlap)
        torch.cuda.synchronize()
        return pop, [o.clone() for o in out]
    p1, l1 = run(True)
    p2, l2 = run(False)
    for m1, m2, x, y in zip(p1, p2, l1, l2):
        assert torch.equal(x, y)
        for a in ids:
            assert torch.equal(m1.actors[a].buffers.params, m2.actors[a].buffers.params)
            assert torch.equal(m1.critic_targets[a].buffers.params, m2.critic_targets[a].buffers.params)


def test_packed_replay_batch_equals_dict_batch_bit_for_bit():
    """learn() fed by MultiAgentReplayBuffer.sample (packed [B, sum] matrices) == learn() fed by the reference-shaped
    dicts of per-agent tensors."""
    from agilerl_b200.components import MultiAgentReplayBuffer
    g = load_golden("maddpg_vector.npz")
    ids, a1 = _agent(g)
    a2 = a1.clone()
    B = int(g["B"])
    buf = MultiAgentReplayBuffer(256, list(FIELDS), ids, device="cuda")
    host = tuple({a: g[f"s0_{f}/{a}"] for a in ids} for f in FIELDS)
    buf.save_to_memory(*host, is_vectorised=True)
    random.seed(5)
    packed = buf.sample(B)
    plain = tuple({a: d[a].clone() for a in ids} for d in packed)
    l1, l2 = a1.learn(packed), a2.learn(plain)
    assert l1 == l2
    for a in ids:
        assert torch.equal(a1.actors[a].buffers.params, a2.actors[a].buffers.params)
        assert torch.equal(a1.critics[a].buffers.params, a2.critics[a].buffers.params)
        assert torch.equal(a1.critic_targets[a].buffers.params, a2.critic_targets[a].buffers.params)


def test_learn_at_bench_batch_matches_oracle():
    """BASELINE configs[4] shapes (4 agents x 18-dim observations) at B = 256 against the pinned oracle."""
    from oracle import maddpg as om
    g = load_golden("maddpg_vector.npz")
    ids, agent = _agent(g)
    B = 256
    a_hidden = [int(h) for h in g["a_hidden"]]
    a_specs = {a: om.actor_specs(int(o), int(d), head_hidden=a_hidden) for a, o, d in zip(ids, g["obs_dims"], g["act_dims"])}
    orc = om.OracleMADDPG(ids, a_specs, om.critic_head_spec(int(g["act_dims"].sum()), head_hidden=a_hidden),
                          {a: _sd(g, f"actor0/{a}") for a in ids}, {a: _sd(g, f"actor_target0/{a}") for a in ids},
                          {a: _sd(g, f"critic0/{a}") for a in ids}, {a: _sd(g, f"critic_target0/{a}") for a in ids},
                          gamma=float(g["gamma"]), tau=float(g["tau"]), lr_actor=float(g["lr_actor"]), lr_critic=float(g["lr_critic"]))
    gen = torch.Generator().manual_seed(3)
    exp = ({a: torch.randn(B, int(o), generator=gen) for a, o in zip(ids, g["obs_dims"])},
           {a: torch.rand(B, int(d), generator=gen) * 2 - 1 for a, d in zip(ids, g["act_dims"])},
           {a: torch.randn(B, 1, generator=gen) for a in ids},
           {a: torch.randn(B, int(o), generator=gen) for a, o in zip(ids, g["obs_dims"])},
           {a: (torch.rand(B, 1, generator=gen) < 0.2).float() for a in ids})
    ref = orc.learn(tuple({a: v.clone() for a, v in d.items()} for d in exp))
    got = agent.learn(tuple({a: v.cuda() for a, v in d.items()} for d in exp))
    for a in ids:
        for j in range(2):
            assert abs(got[a][j] - ref[a][j]) <= 1e-5 * max(1.0, abs(ref[a][j])), (a, j, got[a][j], ref[a][j])


def test_replay_matches_reference_golden():
    from agilerl_b200.components import MultiAgentReplayBuffer
    g = load_golden("ma_replay.npz")
    ids, fields = [str(a) for a in g["agent_ids"]], [str(f) for f in g["fields"]]
    buf = MultiAgentReplayBuffer(int(g["cap"]), fields, ids, device="cuda")
    for t in range(int(g["n_steps"])):
        args = [{a: g[f"t{t}_{f}/{a}"] for a in ids} for f in fields]
        buf.save_to_memory(*args, is_vectorised=bool(int(g[f"t{t}_vect"])))
    assert len(buf) == int(g["final_len"]) and buf.counter == int(g["final_counter"])
    for c in range(int(g["n_samples"])):
        random.seed(int(g[f"sample{c}_seed"]))
        batch = buf.sample(int(g[f"sample{c}_B"]))
        assert isinstance(batch, tuple) and len(batch) == len(fields)
        for f, d in zip(fields, batch):
            for a in ids:
                assert d[a].dtype == torch.float32 and d[a].is_cuda
                np.testing.assert_array_equal(d[a].cpu().numpy(), g[f"sample{c}_{f}/{a}"], err_msg=f"{c} {f} {a}")
    with pytest.raises(ValueError):
        buf.sample(int(g["cap"]) + 1)
    dev_batch = buf.sample_device(16)
    assert dev_batch[0].packed.shape == (16, int(g["obs_dims"].sum())) and torch.isfinite(dev_batch[0].packed).all()


def _restate_mutation(W, rows, cols, r, z, sd):
    """mutation.py:771-822 on a CPU copy with the noise given per slot (cur + |k cur| z / z), index_put_ last-writer."""
    rows, cols = torch.as_tensor(rows), torch.as_tensor(cols)
    cur = W[rows, cols]
    new = cur.clone()
    ms, mr, mn = r < 0.05, (r >= 0.05) & (r < 0.1), r >= 0.1
    new[ms] = cur[ms] + (10 * cur[ms]).abs() * z[ms]
    new[mr] = z[mr]
    new[mn] = cur[mn] + (sd * cur[mn]).abs() * z[mn]
    out = W.clone()
    out[rows, cols] = new.clamp(-1000000, 1000000)
    return out


def test_device_gaussian_mutation_matches_index_put_restatement():
    from agilerl_b200.hpo import Mutations
    g = load_golden("maddpg_vector.npz")
    ids, agent = _agent(g)
    net = agent.actors[ids[0]]
    before = {k: v.cpu().clone() for k, v in net.state_dict().items()}
    keys = [k for k, e in net.layout.entries.items() if "norm" not in k and len(e.shape) == 2]
    gz = torch.Generator().manual_seed(9)
    normals = {k: torch.randn(int(np.ceil(0.1 * np.prod(net.layout.entries[k].shape))), generator=gz) for k in keys}
    m = Mutations(0, 0, 0.5, 1, 0, 0, mutation_sd=0.1, rand_seed=3, device="cuda")
    m._gaussian_parameter_mutation_device(net, normals=normals)
    rng = np.random.default_rng(3)                              # the same decisions, restated on the host
    expect = dict(before)
    for key in rng.choice(keys, int(rng.integers(1, len(keys) + 1)), replace=False):
        key = str(key)
        W = expect[key]
        n_mut = int(np.ceil(0.1 * W.shape[0] * W.shape[1]))
        rows, cols = rng.integers(0, W.shape[0], size=n_mut), rng.integers(0, W.shape[1], size=n_mut)
        r = torch.tensor(rng.uniform(0, 1, size=n_mut), dtype=W.dtype)
        expect[key] = _restate_mutation(W, rows, cols, r, normals[key], 0.1)
    after = net.state_dict()
    changed = 0
    for k in before:
        np.testing.assert_allclose(after[k].cpu().numpy(), expect[k].numpy(), rtol=1e-6, atol=1e-7, err_msg=k)
        changed += int((after[k].cpu() != before[k]).sum())
    assert changed > 0


def test_parameter_mutation_on_device_keeps_targets_in_step_and_learns():
    from agilerl_b200.hpo import Mutations
    g = load_golden("maddpg_vector.npz")
    ids, agent = _agent(g)
    clone = agent.clone(index=7)
    assert clone.index == 7
    for a in ids:
        assert torch.equal(clone.actors[a].buffers.params, agent.actors[a].buffers.params)
        assert torch.equal(clone.critic_targets[a].buffers.params, agent.critic_targets[a].buffers.params)
    m = Mutations(0, 0, 0.5, 1, 0, 0, rand_seed=1, device="cuda")
    m.device_parameter_mutation = True
    before = {a: clone.actors[a].buffers.params.clone() for a in ids}
    [mutated] = m.mutation([clone])
    assert mutated.mut == "param"
    for a in ids:
        p = mutated.actors[a].buffers.params
        frac = float((p != before[a]).float().mean())
        assert 0.0 < frac < 0.12 and torch.isfinite(p).all(), (a, frac)      # <= 10 % of each chosen matrix
        assert torch.equal(mutated.actor_targets[a].buffers.params, p)        # shared networks reloaded (mutation.py:557-570)
        assert torch.equal(agent.actors[a].buffers.params, before[a])         # the parent is untouched
    losses = mutated.learn(_batch(g, 0, ids))
    assert all(np.isfinite(v) for pair in losses.values() for v in pair)
    acts, raw = mutated.get_action({a: np.zeros((1, int(d)), np.float32) for a, d in zip(ids, g["obs_dims"])})
    assert set(acts) == set(ids) and all(v.shape == (1, int(d)) for v, d in zip(acts.values(), g["act_dims"]))


def test_tournament_selects_clones_and_moved_members_learn_identically():
    """TournamentSelection over a MADDPG population (hpo/tournament.py:41-119) and the cross-rank move used by the sharded
    tournament (export_state -> from_state): the rebuilt member continues bit-identically to the original."""
    import pickle
    from agilerl_b200.hpo import TournamentSelection
    g = load_golden("maddpg_vector.npz")
    ids, base = _agent(g)
    pop = [base.clone(index=k) for k in range(4)]
    for k, m in enumerate(pop):
        m.fitness = [float(k)]
        m.learn(_batch(g, 0, ids))
    np.random.seed(0)
    elite, new_pop = TournamentSelection(2, True, 4, 1).select(pop)
    assert elite.index == 3 and len(new_pop) == 4 and all(type(m) is type(base) for m in new_pop)
    assert torch.equal(new_pop[0].critics[ids[0]].buffers.params, pop[3].critics[ids[0]].buffers.params)     # elitism
    src = pop[1]
    meta, tensors = src.export_state()
    moved = type(src).from_state(pickle.loads(pickle.dumps(meta)), [t.clone() for t in tensors], src)
    l1, l2 = src.learn(_batch(g, 2, ids)), moved.learn(_batch(g, 2, ids))
    assert l1 == l2
    for a in ids:
        assert torch.equal(src.actors[a].buffers.params, moved.actors[a].buffers.params)
        assert torch.equal(src.critic_optimizers[a].exp_avg_sq, moved.critic_optimizers[a].exp_avg_sq)
