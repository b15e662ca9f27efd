"""CPU, world_size 2 over gloo: the population-sharded tournament (one fitness all-gather, same
plan on every rank, point-to-point move of winners) and the no-collective learn sharding."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeAgent:
    """Host-only agent exposing what the sharded tournament needs (fitness, index, clone,
    export_state / from_state) with CPU tensors standing in for the flat HBM buffers."""
    device = "cpu"

    def __init__(self, index, weights=None, fitness=None):
        self.index = index
        self.fitness = list(fitness or [])
        self.w = torch.full((16,), float(index)) if weights is None else weights

    def clone(self, index=None, wrap=True):
        return FakeAgent(self.index if index is None else index, self.w.clone(), self.fitness)

    def export_state(self):
        return {"index": self.index, "fitness": self.fitness}, [self.w]

    @classmethod
    def from_state(cls, meta, tensors, like):
        return cls(meta["index"], tensors[0].clone(), meta["fitness"])


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from agilerl_b200.hpo.tournament import TournamentSelection
    n_local, pop_size = 2, 2 * world
    fitness = {0: 1.0, 1: 9.0, 2: 5.0, 3: 3.0}
    results = []
    ts = TournamentSelection(2, True, pop_size, 1, seed=7)
    pop = [FakeAgent(rank * n_local + i, fitness=[fitness[rank * n_local + i]]) for i in range(n_local)]
    for gen in range(3):
        elite, new_pop = ts.select(pop)
        elite_pos, slots = ts.last_plan
        results.append({"plan": slots, "elite_pos": elite_pos,
                        "local": [(a.index, float(a.w[0]), a.fitness) for a in new_pop]})
        pop = new_pop
        for a in pop:                              # new generation's fitness depends on lineage only
            a.fitness = a.fitness + [float(a.w[0]) + gen]
    out[rank] = results
    dist.destroy_process_group()


def test_sharded_tournament_world2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    for g0, g1 in zip(r0, r1):
        assert g0["plan"] == g1["plan"] and g0["elite_pos"] == g1["elite_pos"]     # same plan everywhere
    # generation 0: fitness [1, 9, 5, 3] -> elite is global position 1 (index 1, on rank 0)
    assert r0[0]["elite_pos"] == 1
    plan = r0[0]["plan"]
    assert plan[0] == (1, 1)                                       # elitism keeps the index
    assert [ni for _, ni in plan[1:]] == [4, 5, 6]                 # fresh indices max_id+1...
    # every slot carries its parent's weights (w == parent's original index), wherever it lived
    new = r0[0]["local"] + r1[0]["local"]
    for (parent, new_index), (idx, w0, fit) in zip(plan, new):
        assert idx == new_index and w0 == float(parent)
    # the plan matches the single-process oracle arithmetic given the same RandomState
    from agilerl_b200.hpo.tournament import TournamentSelection
    ts = TournamentSelection(2, True, 4, 1, seed=7)
    _, ref = ts.plan(np.array([1.0, 9.0, 5.0, 3.0]), np.array([0, 1, 2, 3]))
    assert ref == plan


def test_plan_is_rank_invariant_and_ranks_by_mean_fitness():
    from agilerl_b200.hpo.tournament import TournamentSelection
    ts = TournamentSelection(3, False, 8, 3, seed=1)
    fit = np.array([3.0, 1.0, 8.0, 2.0, 7.0, 5.0, 4.0, 6.0])
    idx = np.arange(10, 18)
    a = ts.plan(fit, idx)
    b = TournamentSelection(3, False, 8, 3, seed=1).plan(fit, idx)
    assert a == b and a[0] == 2 and len(a[1]) == 8
    assert [n for _, n in a[1]] == list(range(18, 26))
    # winners are never the worst-ranked of their draws: with tournament_size 3 the worst agent
    # (position 1) can only win if drawn three times
    assert sum(1 for p, _ in a[1] if p == 1) <= 1


def _share_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from agilerl_b200.compat import TensorDict
    from agilerl_b200.training.population import share_transitions
    E = 3
    g = torch.Generator().manual_seed(100 + rank)
    td = TensorDict({"obs": torch.randint(0, 256, (E, 4, 6, 6), dtype=torch.uint8, generator=g),
                     "action": torch.randint(0, 5, (E,), generator=g).float(), "reward": torch.randn(E, generator=g),
                     "next_obs": torch.randint(0, 256, (E, 4, 6, 6), dtype=torch.uint8, generator=g),
                     "done": (torch.rand(E, generator=g) < 0.5).float()}, batch_size=[E])
    allr = share_transitions(td)
    out[rank] = {k: v.clone() for k, v in allr.items()} | {"_own": {k: v.clone() for k, v in td.items()},
                                                         "_bs": tuple(allr.batch_size)}
    dist.destroy_process_group()


def test_cross_rank_experience_sharing_world2_gloo():
    """share_transitions: one all-gather per environment step; every rank ends up with the transitions of ALL ranks
    concatenated along the environment dimension in rank order (the reference's single shared buffer,
    train_off_policy.py:327-345), dtypes and shapes preserved."""
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_share_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert r0["_bs"] == (6,) and r1["_bs"] == (6,)
    for k in ("obs", "action", "reward", "next_obs", "done"):
        want = torch.cat([r0["_own"][k], r1["_own"][k]], dim=0)
        assert r0[k].dtype == want.dtype and torch.equal(r0[k], want) and torch.equal(r1[k], want), k


def _maddpg_worker(rank, world, port, out):
    """Real ``MADDPG`` members (host side only: the C entry points are never reached — construction, clone, export_state /
    from_state and the tournament's collectives are pure host + tensor-copy work) through the sharded tournament."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from agilerl_b200 import _lib
    _lib.as_device = lambda d: torch.device("cpu")                 # test-only: flat buffers as CPU tensors
    _lib.load = lambda require_cuda=False: None
    from agilerl_b200.algorithms import MADDPG
    from agilerl_b200.compat import spaces
    from agilerl_b200.hpo.tournament import TournamentSelection
    ids = ["speaker_0", "listener_0"]
    obs = [spaces.Box(-1.0, 1.0, (3,), np.float32), spaces.Box(-1.0, 1.0, (11,), np.float32)]
    act = [spaces.Box(-1.0, 1.0, (3,), np.float32), spaces.Box(-1.0, 1.0, (5,), np.float32)]
    n_local, pop_size = 2, 2 * world
    fitness = {0: 1.0, 1: 9.0, 2: 5.0, 3: 3.0}
    pop = []
    for i in range(n_local):
        gi = rank * n_local + i
        torch.manual_seed(100 + gi)
        m = MADDPG(obs, act, agent_ids=ids, index=gi, batch_size=8, lr_actor=1e-4 * (gi + 1))
        m.fitness = [fitness[gi]]
        for o in m._all_opts:
            o.step = 10 + gi
            o.exp_avg.fill_(float(gi))
        pop.append(m)
    marks = {m.index: float(m.actors["listener_0"].buffers.params.sum()) for m in pop}
    ts = TournamentSelection(2, True, pop_size, 1, seed=7)
    elite, new_pop = ts.select(pop)
    out[rank] = {"plan": ts.last_plan, "marks": marks,
                 "local": [(m.index, float(m.actors["listener_0"].buffers.params.sum()), m.lr_actor, m._all_opts[0].step,
                            float(m.critic_optimizers["speaker_0"].exp_avg[0]), list(m.fitness), m.agent_ids) for m in new_pop]}
    dist.destroy_process_group()


def test_sharded_tournament_moves_maddpg_members_world2_gloo():
    """BASELINE configs[4] shards the MADDPG population two members per GPU: a winner living on another rank travels as
    its pickled description + 8 flat tensors per agent (parameters, targets, Adam moments) and arrives with its weights,
    hyper-parameters, optimiser step and fitness history."""
    world = 2
    port = 33500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_maddpg_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert r0["plan"] == r1["plan"]
    elite_pos, slots = r0["plan"]
    assert elite_pos == 1 and slots[0] == (1, 1)
    marks = {**r0["marks"], **r1["marks"]}
    new = r0["local"] + r1["local"]
    crossed = 0
    for slot, ((parent, new_index), (idx, mark, lr_actor, step, m0, fit, ids)) in enumerate(zip(slots, new)):
        assert idx == new_index and mark == marks[parent]                       # the parent's weights, wherever it lived
        assert lr_actor == pytest.approx(1e-4 * (parent + 1)) and step == 10 + parent and m0 == float(parent)
        assert fit == [{0: 1.0, 1: 9.0, 2: 5.0, 3: 3.0}[parent]] and ids == ["speaker_0", "listener_0"]
        crossed += int(slot // 2 != parent // 2)
    assert crossed >= 1                                                          # at least one member changed rank


def _td3_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from agilerl_b200 import _lib
    _lib.as_device = lambda d: torch.device("cpu")                 # test-only: flat buffers as CPU tensors
    _lib.load = lambda require_cuda=False: None
    from agilerl_b200.algorithms import TD3
    from agilerl_b200.compat import spaces
    from agilerl_b200.hpo.tournament import TournamentSelection
    osp, asp = spaces.Box(-np.inf, np.inf, (17,), np.float32), spaces.Box(-1.0, 1.0, (6,), np.float32)
    n_local, fitness = 2, {0: 1.0, 1: 9.0, 2: 5.0, 3: 3.0}
    pop = []
    for i in range(n_local):
        gi = rank * n_local + i
        torch.manual_seed(200 + gi)
        m = TD3(osp, asp, index=gi, batch_size=8, lr_critic=1e-3 * (gi + 1),
                net_config={"head_config": {"hidden_size": [16 * (gi + 1)]}})         # members differ in architecture too
        m.fitness, m.learn_counter = [fitness[gi]], 10 + gi
        m.critic_2_optimizer.step = 20 + gi
        m.critic_2_optimizer.exp_avg_sq.fill_(float(gi))
        pop.append(m)
    marks = {m.index: (float(m.critic_2.buffers.params.sum()), list(m.critic_2.head_net.hidden_size)) for m in pop}
    ts = TournamentSelection(2, True, 2 * world, 1, seed=7)
    elite, new_pop = ts.select(pop)
    out[rank] = {"plan": ts.last_plan, "marks": marks,
                 "local": [(m.index, float(m.critic_2.buffers.params.sum()), list(m.critic_2.head_net.hidden_size), m.lr_critic,
                            m.critic_2_optimizer.step, float(m.critic_2_optimizer.exp_avg_sq[0]), m.learn_counter, list(m.fitness))
                           for m in new_pop]}
    dist.destroy_process_group()


def test_sharded_tournament_moves_td3_members_world2_gloo():
    """BASELINE configs[2] shards the TD3 population over the GPUs: a winner on another rank travels with its six networks
    (their own, possibly mutated, architectures), all three optimisers' moments / step counts, the policy_freq phase and its
    hyper-parameters."""
    world = 2
    port = 35500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_td3_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert r0["plan"] == r1["plan"]
    _, slots = r0["plan"]
    marks = {**r0["marks"], **r1["marks"]}
    crossed = 0
    for slot, ((parent, new_index), (idx, mark, hidden, lr, step, v0, lc, fit)) in enumerate(zip(slots, r0["local"] + r1["local"])):
        assert idx == new_index and (mark, hidden) == marks[parent]
        assert lr == pytest.approx(1e-3 * (parent + 1)) and step == 20 + parent and v0 == float(parent) and lc == 10 + parent
        assert fit == [{0: 1.0, 1: 9.0, 2: 5.0, 3: 3.0}[parent]]
        crossed += int(slot // 2 != parent // 2)
    assert crossed >= 1
