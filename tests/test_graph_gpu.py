"""CUDA-graph replay of the fused gradient step == the eager launch sequence, bit for bit.

Two identical worlds (same weights, same replay contents, same Philox seeds); one runs
``rainbow_fused_step(graph=False)``, the other the default graph path (step 1 eager, step 2 captures, later steps
replay).  beta and lr change between steps (what train_off_policy.py:346-351 and an lr mutation do), the replay
grows between steps (len(memory) is a per-step scalar), so every field of b2rl_step_state is exercised."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

OBS, A, B = (3, 20, 20), 4, 32


def _world(seed, n_agents=1):
    from agilerl_b200.compat import TensorDict
    from agilerl_b200.components import MultiStepReplayBuffer, PrioritizedReplayBuffer
    from agilerl_b200.components.replay_buffer import ReplayBuffer
    from agilerl_b200.engine import LearnEngine, NetBuffers
    from agilerl_b200.networks.init import init_state_dict
    from agilerl_b200.networks.spec import FlatLayout, rainbow_spec
    torch.manual_seed(seed)
    layout = FlatLayout(rainbow_spec(OBS, A, channel_size=(8, 16), kernel_size=(4, 3), stride_size=(2, 1), latent_dim=16,
                                     hidden_size=(32,), obs_low=0.0, obs_high=255.0, obs_u8=True))
    engines = []
    for a in range(n_agents):
        sd = init_state_dict(layout)
        actor, target = NetBuffers(layout, "cuda"), NetBuffers(layout, "cuda")
        actor.load_state_dict(sd, strict=False); target.load_state_dict(sd, strict=False)
        eng = LearnEngine(layout, actor, target)
        eng.philox_seed = 1000 + a
        eng.reset_noise(actor); eng.reset_noise(target)
        engines.append(eng)
    mem, nmem = PrioritizedReplayBuffer(512, 0.6), MultiStepReplayBuffer(512, 3, 0.99)
    mem.device_rng = True
    g = torch.Generator().manual_seed(seed + 1)

    def grow(n):
        small = dict(action=torch.randint(0, A, (n,), generator=g).float(), reward=torch.randn(n, generator=g),
                     done=(torch.rand(n, generator=g) < 0.1).float())
        td = TensorDict(dict(small, obs=torch.randint(0, 256, (n, *OBS), dtype=torch.uint8, generator=g),
                             next_obs=torch.randint(0, 256, (n, *OBS), dtype=torch.uint8, generator=g)), batch_size=[n])
        ReplayBuffer.add(nmem, td.to("cuda"))
        nmem.done_key = "done"
        mem.add(TensorDict(small, batch_size=[n]).to("cuda"))
    grow(200)
    pri = torch.rand(200, generator=g) + 0.1
    mem.update_priorities(torch.arange(200), pri.numpy())
    return engines, mem, nmem, grow


def _snapshot(engines, mem):
    torch.cuda.synchronize()
    out = [mem.sum_tree._t.clone(), mem.min_tree._t.clone(), mem._max_priority_dev.clone()]
    for e in engines:
        out += [e.actor.params.clone(), e.target.params.clone(), e.actor.eps.clone(), e.target.eps.clone(),
                e.exp_avg.clone(), e.exp_avg_sq.clone(), e.grads.clone()]
    return out


def _run(graph, overlap, n_agents=1, steps=7):
    engines, mem, nmem, grow = _world(3, n_agents)
    support = torch.linspace(-10.0, 10.0, 51).cuda()
    trace = []
    hi = torch.cuda.Stream()
    for step in range(steps):
        hp = dict(v_min=-10.0, v_max=10.0, delta_z=20.0 / 50, lr=1e-3 if step < 4 else 5e-4, tau=1e-3, prior_eps=1e-6)
        beta = 0.4 + 0.05 * step
        hi.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(hi):
            for eng in engines:
                loss, idx, pri = eng.rainbow_fused_step(mem, nmem, B=B, beta=beta, support=support, hp=hp,
                                                        gamma_n=0.99 ** 3, overlap=overlap, graph=graph)
                trace.append((loss.clone(), idx.clone(), pri.clone()))
        torch.cuda.current_stream().wait_stream(hi)
        for eng in engines:
            eng.join()
        torch.cuda.synchronize()
        if step in (1, 4):
            grow(16)                                     # len(memory) changes under the captured graphs
    plans = sum(len(e._plans) for e in engines)
    return trace, _snapshot(engines, mem), plans, engines


@pytest.mark.parametrize("overlap,n_agents", [(False, 1), (True, 1), (True, 3)])
def test_graph_replay_is_bit_identical_to_eager(overlap, n_agents):
    t_e, s_e, _, _ = _run(False, overlap, n_agents)
    t_g, s_g, plans, engines = _run(True, overlap, n_agents)
    assert plans == n_agents and all(p.front and p.tail for e in engines for p in e._plans.values())
    assert all(p.kernels >= 20 for e in engines for p in e._plans.values())
    for k, ((le, ie, pe), (lg, ig, pg)) in enumerate(zip(t_e, t_g)):
        assert torch.equal(ie, ig), f"sampled indices differ at call {k}"
        assert torch.equal(le, lg) and torch.equal(pe, pg), f"loss / priorities differ at call {k}"
    for k, (a, b) in enumerate(zip(s_e, s_g)):
        assert torch.equal(a, b), f"state tensor {k} differs between eager and graph replay"


def test_graph_path_counts_its_kernels_and_matches_population_helper():
    """population_learn (what bench.py times) goes through the graph path; b2rl_launch_count advances by the
    graphs' kernel-node count per replay."""
    from agilerl_b200 import _lib
    engines, mem, nmem, _ = _world(5, 1)
    eng = engines[0]
    support = torch.linspace(-10.0, 10.0, 51).cuda()
    hp = dict(v_min=-10.0, v_max=10.0, delta_z=20.0 / 50, lr=1e-3, tau=1e-3, prior_eps=1e-6)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            eng.rainbow_fused_step(mem, nmem, B=B, beta=0.4, support=support, hp=hp, gamma_n=0.99 ** 3)
        lib = _lib.load()
        n0 = lib.b2rl_launch_count()
        eng.rainbow_fused_step(mem, nmem, B=B, beta=0.4, support=support, hp=hp, gamma_n=0.99 ** 3)
        n1 = lib.b2rl_launch_count()
    torch.cuda.synchronize()
    plan = next(iter(eng._plans.values()))
    assert n1 - n0 == plan.kernels and plan.kernels > 0


def test_api_learn_graph_replay_matches_eager_bit_for_bit():
    """``agent.learn`` replays two captured graphs when it is handed the same device buffers again (the steady
    state of a training loop under torch's caching allocator): losses, priorities, parameters, moments and noise are
    identical to the eager launch sequence; a call on other buffers falls back to eager and stays correct."""
    import agilerl_b200.engine as E
    from agilerl_b200.algorithms import RainbowDQN
    from agilerl_b200.compat import spaces

    def run(graph):
        E._GRAPH = graph
        try:
            torch.manual_seed(0)
            net = {"encoder_config": {"channel_size": [8, 16], "kernel_size": [4, 3], "stride_size": [2, 1]},
                   "head_config": {"hidden_size": [32]}, "latent_dim": 16}
            agent = RainbowDQN(spaces.Box(0, 255, OBS, np.uint8), spaces.Discrete(A), index=2, net_config=net, batch_size=B,
                               v_min=-10.0, v_max=10.0, lr=1e-3)
            g = torch.Generator().manual_seed(9)
            exp = dict(obs=torch.empty((B, *OBS), dtype=torch.uint8, device="cuda"), next_obs=torch.empty((B, *OBS), dtype=torch.uint8, device="cuda"),
                       action=torch.empty(B, 1, device="cuda"), reward=torch.empty(B, 1, device="cuda"), done=torch.empty(B, 1, device="cuda"),
                       weights=torch.empty(B, device="cuda"), idxs=torch.arange(B, device="cuda"))
            other = {k: v.clone() for k, v in exp.items()}
            trace = []
            for step in range(7):
                tgt = other if step == 5 else exp                      # one call on different buffers (eager fallback)
                tgt["obs"].copy_(torch.randint(0, 256, (B, *OBS), dtype=torch.uint8, generator=g))
                tgt["next_obs"].copy_(torch.randint(0, 256, (B, *OBS), dtype=torch.uint8, generator=g))
                tgt["action"].copy_(torch.randint(0, A, (B, 1), generator=g).float())
                tgt["reward"].copy_(torch.randn(B, 1, generator=g))
                tgt["done"].copy_((torch.rand(B, 1, generator=g) < 0.1).float())
                tgt["weights"].copy_(torch.rand(B, generator=g) + 0.5)
                if step == 3:
                    agent.lr = 5e-4                                    # an lr mutation between steps
                loss, idxs, pri = agent.learn(tgt, n_experiences=tgt, per=True)
                trace.append((loss, pri.copy()))
            agent.synchronize()
            torch.cuda.synchronize()
            e = agent.engine
            replays = sum(p.seen for p in e._api_plans.values())
            return trace, [e.actor.params.clone(), e.target.params.clone(), e.actor.eps.clone(), e.target.eps.clone(),
                           e.exp_avg.clone(), e.exp_avg_sq.clone()], replays
        finally:
            E._GRAPH = True

    t_e, s_e, r_e = run(False)
    t_g, s_g, r_g = run(True)
    assert r_e == 0 and r_g >= 4, (r_e, r_g)
    for k, ((le, pe), (lg, pg)) in enumerate(zip(t_e, t_g)):
        assert le == lg and np.array_equal(pe, pg), f"loss / priorities differ at step {k}"
    for k, (a, b) in enumerate(zip(s_e, s_g)):
        assert torch.equal(a, b), f"state tensor {k} differs"
