"""GPU: the reference-shaped Python API end to end — RainbowDQN / DQN through Sampler + buffers,
clone, every mutation kind, tournament selection, and the driver loop (BASELINE configs[0]:
DQN pop=4 plumbing on a CartPole-like env)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class VecEnv:
    """Synthetic vector env: random observations, short episodes (the reference's test DummyEnv
    idea, tests/test_train/test_train.py:46-69)."""

    def __init__(self, obs_shape, n_act, num_envs=2, image=False, seed=0):
        self.num_envs, self.obs_shape, self.n_act, self.image = num_envs, obs_shape, n_act, image
        self.rng = np.random.default_rng(seed)
        self.t = 0

    def _obs(self):
        if self.image:
            return self.rng.integers(0, 256, (self.num_envs, *self.obs_shape), dtype=np.uint8)
        return self.rng.standard_normal((self.num_envs, *self.obs_shape)).astype(np.float32)

    def reset(self):
        self.t = 0
        return self._obs(), {}

    def step(self, action):
        assert np.asarray(action).shape == (self.num_envs,)
        self.t += 1
        done = np.array([self.t % 7 == 0] * self.num_envs)
        return self._obs(), self.rng.standard_normal(self.num_envs), done, np.zeros(self.num_envs, bool), {}


def _spaces(image=True):
    from agilerl_b200.compat import spaces
    if image:
        return spaces.Box(0, 255, (3, 20, 20), np.uint8), spaces.Discrete(4)
    return spaces.Box(-1, 1, (4,), np.float32), spaces.Discrete(2)


NET = {"encoder_config": {"channel_size": [8, 16], "kernel_size": [4, 3], "stride_size": [2, 1]},
       "head_config": {"hidden_size": [32]}, "latent_dim": 16}


def _fill(agent, mem, nmem, env, steps=40):
    from agilerl_b200.components import Transition
    obs, _ = env.reset()
    for _ in range(steps):
        a = agent.get_action(obs)
        nobs, r, d, t, _ = env.step(a)
        one = nmem.add(Transition(obs=obs, action=a, reward=r, next_obs=nobs, done=d, batch_size=[env.num_envs]).to_tensordict())
        if one is not None:
            mem.add(one)
        obs = nobs


def test_rainbow_api_learn_matches_oracle_driver_and_canonical_shapes(monkeypatch):
    from agilerl_b200.algorithms import RainbowDQN
    from agilerl_b200.components import MultiStepReplayBuffer, PrioritizedReplayBuffer, Sampler
    from oracle import learn as olearn, nets as onets
    obs_space, act_space = _spaces()
    ospec = onets.rainbow_spec((3, 20, 20), 4, (8, 16), (4, 3), (2, 1), 16, (32,))
    for driver_shapes in (True, False):
        torch.manual_seed(0)          # fresh agent (fresh Adam state) per case, like the oracle side
        agent = RainbowDQN(obs_space, act_space, net_config=dict(NET), batch_size=16, v_min=-10.0, v_max=10.0, lr=1e-3)
        assert agent.algo == "Rainbow DQN" and agent.action_dim == 4
        mem, nmem = PrioritizedReplayBuffer(128, 0.6), MultiStepReplayBuffer(128, 3, 0.99)
        _fill(agent, mem, nmem, VecEnv((3, 20, 20), 4, image=True))
        s, ns = Sampler(memory=mem), Sampler(memory=nmem)
        assert s.per and ns.n_step
        exp = s.sample(16, agent.beta)
        nexp = ns.sample(exp["idxs"] if driver_shapes else exp["idxs"].squeeze(1))
        if not driver_shapes:
            exp["weights"] = exp["weights"].squeeze(1)
        oa = olearn.OracleAgent(ospec, {k: v.cpu() for k, v in agent.actor.state_dict().items()},
                                {k: v.cpu() for k, v in agent.actor_target.state_dict().items()}, batch_size=16, lr=1e-3)
        oexp = {k: v.cpu() for k, v in exp.items()}
        onexp = {k: v.cpu() for k, v in nexp.items()}
        z = (torch.randn(agent.engine.noise_count), torch.randn(agent.engine.noise_count))
        oloss, _, opri = oa.learn_rainbow(oexp, onexp, per=True, noise_normals=z)
        loss, idxs, pri = agent.learn(exp, n_experiences=nexp, per=True, noise_normals=z)
        assert isinstance(loss, float) and isinstance(pri, np.ndarray) and pri.shape == (16,)
        assert idxs is exp["idxs"]
        scale = max(1.0, abs(oloss))
        assert abs(loss - oloss) <= 1e-5 * scale, (driver_shapes, loss, oloss)
        np.testing.assert_allclose(pri, opri, rtol=1e-5, atol=1e-5 * scale)
        mem.update_priorities(idxs, pri)
        for k in oa.pkeys:     # parameters after the step (where Adam's first step has a clear sign)
            mask = oa.last_grads[k].abs() > 1e-5
            if mask.any():
                diff = (agent.actor.state_dict()[k].cpu() - oa.actor[k].detach())[mask].abs().max().item()
                assert diff <= 2e-5, (driver_shapes, k, diff)


def test_learn_variants_and_return_types():
    from agilerl_b200.algorithms import RainbowDQN
    from agilerl_b200.components import MultiStepReplayBuffer, ReplayBuffer, Sampler
    obs_space, act_space = _spaces()
    agent = RainbowDQN(obs_space, act_space, net_config=dict(NET), batch_size=8, v_min=-10.0, v_max=10.0,
                       combined_reward=True)
    mem, nmem = ReplayBuffer(64), MultiStepReplayBuffer(64, 3, 0.99)
    _fill(agent, mem, nmem, VecEnv((3, 20, 20), 4, image=True), steps=20)
    before = agent.actor.state_dict()
    exp = Sampler(memory=mem).sample(8, return_idx=True)
    nexp = Sampler(memory=nmem).sample(exp["idxs"])
    loss, idxs, pri = agent.learn(exp, n_experiences=nexp)          # n-step, no PER, combined
    assert loss > 0 and pri is None and idxs is not None
    loss2, idxs2, pri2 = agent.learn(Sampler(memory=mem).sample(8))  # 1-step, no PER
    assert loss2 > 0 and idxs2 is None and pri2 is None
    after = agent.actor.state_dict()
    assert any(not torch.equal(before[k], after[k]) for k in before if "epsilon" not in k)
    # soft update formula (test_dqn_rainbow.py:572-618)
    p, t = agent.actor.buffers.params.clone(), agent.actor_target.buffers.params.clone()
    agent.soft_update()
    assert torch.allclose(agent.actor_target.buffers.params, agent.tau * p + (1 - agent.tau) * t)


def test_clone_and_every_mutation_kind_then_learn():
    from agilerl_b200.algorithms import RainbowDQN
    from agilerl_b200.algorithms.core.registry import HyperparameterConfig, RLParameter
    from agilerl_b200.components import ReplayBuffer, Sampler
    from agilerl_b200.hpo import Mutations
    obs_space, act_space = _spaces()
    hp = HyperparameterConfig(lr=RLParameter(min=1e-5, max=1e-2), batch_size=RLParameter(min=8, max=32, dtype=int))
    agent = RainbowDQN(obs_space, act_space, index=3, hp_config=hp, net_config=dict(NET), batch_size=8, v_min=-10.0,
                       v_max=10.0)
    mem = ReplayBuffer(64)
    from agilerl_b200.components import MultiStepReplayBuffer
    _fill(agent, mem, MultiStepReplayBuffer(64, 1, 0.99), VecEnv((3, 20, 20), 4, image=True), steps=20)
    agent.learn(Sampler(memory=mem).sample(agent.batch_size))
    agent.fitness, agent.scores = [1.0, 2.0], [3.0]
    c = agent.clone(index=9)
    assert c.index == 9 and c.fitness == agent.fitness and c.batch_size == agent.batch_size
    for k, v in agent.actor.state_dict().items():
        assert torch.equal(v, c.actor.state_dict()[k]), k
    assert torch.equal(agent.engine.exp_avg, c.engine.exp_avg) and c.engine.step == agent.engine.step
    c.learn(Sampler(memory=mem).sample(c.batch_size))
    assert not torch.equal(agent.actor.buffers.params, c.actor.buffers.params)     # deep copy

    muts = Mutations(0, 1, 0.5, 0, 0, 0, rand_seed=1)
    seen = set()
    for fn in (muts.architecture_mutate,) * 6 + (muts.parameter_mutation, muts.activation_mutation,
                                                 muts.rl_hyperparam_mutation, muts.rl_hyperparam_mutation, muts.no_mutation):
        old = {k: v.clone() for k, v in c.actor.state_dict().items()}
        c = fn(c)
        seen.add(c.mut)
        new = c.actor.state_dict()
        assert set(c.actor_target.state_dict()) == set(new)
        for k in new:                                   # target mirrors the evaluation net's shapes
            assert c.actor_target.state_dict()[k].shape == new[k].shape, k
        if fn == muts.architecture_mutate:              # untouched overlapping slices are preserved
            for k in new:
                if k in old and "epsilon" not in k and "norm" not in k:
                    sl = tuple(slice(0, min(a, b)) for a, b in zip(old[k].shape, new[k].shape))
                    assert torch.equal(old[k][sl], new[k][sl]), (c.mut, k)
        loss, *_ = c.learn(Sampler(memory=mem).sample(c.batch_size))
        assert np.isfinite(loss), c.mut
        assert c.get_action(np.zeros((2, 3, 20, 20), np.uint8)).shape == (2,)
    assert "param" in seen and "act" in seen and "None" in seen and len(seen) >= 5, seen


def test_tournament_known_answer_and_population():
    """tests/test_hpo/test_tournament.py:74-125 — elite = best mean fitness, index preserved."""
    from agilerl_b200.hpo import TournamentSelection
    from agilerl_b200.utils.utils import create_population
    obs_space, act_space = _spaces(image=False)
    pop = create_population("DQN", obs_space, act_space, None, {"BATCH_SIZE": 8}, population_size=5)
    for i, a in enumerate(pop):
        a.fitness = [i + 1.0, i + 2.0, i + 3.0]
    ts = TournamentSelection(3, True, 5, 3)
    np.random.seed(0)
    elite, new_pop = ts.select(pop)
    assert elite.index == 4 and elite.fitness == [5.0, 6.0, 7.0]
    assert len(new_pop) == 5 and new_pop[0].index == 4
    assert [a.index for a in new_pop[1:]] == [5, 6, 7, 8]
    from oracle.tournament import select_positions
    np.random.seed(0)
    _, sel = select_positions([a.fitness for a in pop], [a.index for a in pop], 3, True, 5, 3)
    assert [p for p, _ in sel] == [next(i for i, a in enumerate(pop) if a.fitness == n.fitness) for n in new_pop]


def test_driver_config0_dqn_pop4_vector_env():
    """BASELINE configs[0] plumbing: DQN pop=4 through train_off_policy with tournament + mutation."""
    from agilerl_b200.components import ReplayBuffer
    from agilerl_b200.hpo import Mutations, TournamentSelection
    from agilerl_b200.training import train_off_policy
    from agilerl_b200.utils.utils import create_population
    obs_space, act_space = _spaces(image=False)
    pop = create_population("DQN", obs_space, act_space, None, {"BATCH_SIZE": 16, "LEARN_STEP": 2, "DOUBLE": True},
                            population_size=4)
    env = VecEnv((4,), 2, num_envs=2)
    pop, fits = train_off_policy(env, "synthetic", "DQN", pop, ReplayBuffer(512), max_steps=120, evo_steps=40,
                                 eval_steps=10, eval_loop=1, tournament=TournamentSelection(2, True, 4, 1),
                                 mutation=Mutations(0.4, 0.2, 0.2, 0.2, 0.2, 0.2, rand_seed=0), verbose=False)
    assert len(pop) == 4 and len(fits) >= 2 and all(len(f) == 4 for f in fits)
    assert all(a.steps[-1] >= 120 for a in pop)


def test_driver_rainbow_per_nstep_fused_and_api_paths():
    from agilerl_b200.components import MultiStepReplayBuffer, PrioritizedReplayBuffer
    from agilerl_b200.training import train_off_policy
    from agilerl_b200.utils.utils import create_population
    obs_space, act_space = _spaces()
    for fused in (False, True):
        pop = create_population("Rainbow DQN", obs_space, act_space, dict(NET),
                                {"BATCH_SIZE": 8, "LEARN_STEP": 2, "V_MIN": -10.0, "V_MAX": 10.0}, population_size=2)
        env = VecEnv((3, 20, 20), 4, num_envs=2, image=True)
        pop, fits = train_off_policy(env, "synthetic", "Rainbow DQN", pop, PrioritizedReplayBuffer(256, 0.6),
                                     max_steps=60, evo_steps=30, eval_steps=5, n_step=True, per=True,
                                     n_step_memory=MultiStepReplayBuffer(256, 3, 0.99), verbose=False, fused=fused)
        assert len(pop) == 2 and all(a.beta > 0.4 for a in pop)


def test_checkpoint_round_trip(tmp_path):
    from agilerl_b200.algorithms import RainbowDQN
    obs_space, act_space = _spaces()
    a = RainbowDQN(obs_space, act_space, net_config=dict(NET), batch_size=8, v_min=-10.0, v_max=10.0)
    a.fitness = [1.5]
    p = str(tmp_path / "agent.pt")
    a.save_checkpoint(p)
    b = RainbowDQN(obs_space, act_space, net_config=dict(NET), batch_size=8, v_min=-10.0, v_max=10.0)
    b.load_checkpoint(p)
    for k, v in a.actor.state_dict().items():
        assert torch.equal(v, b.actor.state_dict()[k]), k
    assert b.fitness == [1.5]


def test_overlapped_learn_tail_is_bit_identical_to_sequential():
    """learn_from_buffers(overlap=True) leaves backward + optimiser on the agent's own stream; the
    tree write-back stays ahead of the next agent's sampling on the caller's stream.  Parameters of
    every agent and the priority trees must equal the sequential run bit for bit."""
    from agilerl_b200.components import MultiStepReplayBuffer, PrioritizedReplayBuffer
    from agilerl_b200.utils.utils import create_population
    obs_space, act_space = _spaces()

    def run(overlap):
        torch.manual_seed(0)
        np.random.seed(0)
        pop = create_population("Rainbow DQN", obs_space, act_space, dict(NET),
                                {"BATCH_SIZE": 16, "LEARN_STEP": 1, "V_MIN": -10.0, "V_MAX": 10.0, "N_STEP": 3},
                                population_size=3)
        for i, a in enumerate(pop):
            a.engine.philox_seed, a.engine.philox_offset = 1000 + i, 0
        mem, nmem = PrioritizedReplayBuffer(256, 0.6), MultiStepReplayBuffer(256, 3, 0.99)
        _fill(pop[0], mem, nmem, VecEnv((3, 20, 20), 4, num_envs=2, image=True, seed=3), steps=60)
        from agilerl_b200.training import population_learn
        losses = []
        for _ in range(4):
            if overlap:      # high-priority forward chain + per-agent backward streams
                losses += population_learn(pop, mem, nmem, overlap=True, join=False)
            else:
                losses += [a.learn_from_buffers(mem, nmem) for a in pop]
        for a in pop:
            a.synchronize()
        torch.cuda.synchronize()
        params = [a.actor.buffers.params.clone() for a in pop] + [a.actor_target.buffers.params.clone() for a in pop]
        return params, torch.stack([l.reshape(()) for l in losses]), mem.sum_tree._t.clone(), mem.min_tree._t.clone()

    p0, l0, s0, m0 = run(False)
    p1, l1, s1, m1 = run(True)
    import os
    if os.environ.get("B2RL_DISABLE_TC") == "1":
        # the FFMA fallback's conv input gradient accumulates with fp32 atomics: run-to-run bit noise
        assert torch.allclose(l0, l1, rtol=1e-4, atol=1e-5)
        for a, b in zip(p0, p1):
            assert torch.allclose(a, b, rtol=1e-3, atol=1e-5)
        return
    assert torch.equal(l0, l1)
    assert torch.equal(s0, s1) and torch.equal(m0, m1)
    for a, b in zip(p0, p1):
        assert torch.equal(a, b)


def test_clone_keeps_hyperparameter_config_and_rl_hp_mutation_acts_on_clones(tmp_path):
    """the cloned agent must carry the HyperparameterConfig, otherwise every
    rl_hyperparam_mutation after the first tournament is a silent no-op; the sharded move (export_state /
    from_state) and checkpoints carry it too."""
    from agilerl_b200.algorithms import RainbowDQN
    from agilerl_b200.algorithms.core.registry import HyperparameterConfig, RLParameter
    from agilerl_b200.hpo import Mutations
    obs_space, act_space = _spaces()
    hp = HyperparameterConfig(lr=RLParameter(min=1e-5, max=1e-2), batch_size=RLParameter(min=8, max=32, dtype=int))
    agent = RainbowDQN(obs_space, act_space, index=3, hp_config=hp, net_config=dict(NET), batch_size=16, v_min=-10.0,
                       v_max=10.0)
    c = agent.clone(index=4).clone(index=5)                       # two generations of tournament clones
    assert bool(c.registry.hp_config) and sorted(c.registry.hp_config.names()) == ["batch_size", "lr"]
    assert c.registry.hp_config is not agent.registry.hp_config   # deep copy, like copy_attributes
    mut = Mutations(0, 0, 0, 0, 0, 1.0, rand_seed=1)
    seen = set()
    for _ in range(8):
        before = (c.lr, c.batch_size)
        c = mut.rl_hyperparam_mutation(c)
        assert c.mut in ("lr", "batch_size")
        assert (c.lr, c.batch_size) != before
        seen.add(c.mut)
        c = c.clone()
    assert seen == {"lr", "batch_size"}
    assert c.optimizer.lr == c.lr
    moved = type(c).from_state(*c.export_state(), like=agent)
    assert sorted(moved.registry.hp_config.names()) == ["batch_size", "lr"] and moved.lr == c.lr
    # checkpoint round trip restores mutated hyper-parameters, the derived support and the architecture
    c.actor.encoder.add_channel(hidden_layer=0, numb_new_channels=8)
    c.actor_target = type(c.actor)(**c.actor.init_dict)
    c.actor_target.load_state_dict(c.actor.state_dict())
    c.reinit_optimizers()
    c.beta, c.tau, c.v_min, c.v_max = 0.7, 5e-3, -5.0, 5.0
    c._after_hyperparameter_restore()
    c.engine.exp_avg.fill_(0.5); c.engine.step = 11
    path = str(tmp_path / "agent.pt")
    c.save_checkpoint(path)
    fresh = RainbowDQN(obs_space, act_space, index=0, net_config=dict(NET), batch_size=16, v_min=-10.0, v_max=10.0)
    fresh.load_checkpoint(path)
    assert (fresh.lr, fresh.batch_size, fresh.beta, fresh.tau, fresh.v_min, fresh.v_max) == \
           (c.lr, c.batch_size, 0.7, 5e-3, -5.0, 5.0)
    assert torch.equal(fresh.support.cpu(), torch.linspace(-5.0, 5.0, 51)) and fresh.delta_z == 10.0 / 50
    assert list(fresh.actor.encoder.channel_size) == list(c.actor.encoder.channel_size)
    assert fresh.engine.step == 11 and float(fresh.engine.exp_avg.mean()) == 0.5 and fresh.optimizer.lr == c.lr
    assert sorted(fresh.registry.hp_config.names()) == ["batch_size", "lr"]
    for k, v in c.actor.state_dict().items():
        assert torch.equal(v, fresh.actor.state_dict()[k]), k
    with pytest.raises(ValueError):                               # optimiser state of another architecture is not dropped silently
        other = RainbowDQN(obs_space, act_space, net_config=dict(NET), batch_size=16, v_min=-10.0, v_max=10.0)
        other.optimizer.load_state_dict(c.optimizer.state_dict(), strict=True)
