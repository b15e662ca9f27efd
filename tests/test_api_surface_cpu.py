"""CPU, build container only: static drop-in check against the UNMODIFIED reference (imported
through oracle.refshim): constructor parameters and public methods of every replaced class must be
present with the same names (and order for constructors)."""
import inspect

import pytest

from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.available(), reason="/root/reference not present (GPU box)")


@pytest.fixture(scope="module")
def ref():
    refshim.install()
    import agilerl.algorithms.dqn as r_dqn
    import agilerl.algorithms.dqn_rainbow as r_rb
    import agilerl.components.replay_buffer as r_buf
    import agilerl.components.sampler as r_samp
    import agilerl.components.segment_tree as r_tree
    import agilerl.hpo.mutation as r_mut
    import agilerl.hpo.tournament as r_tour
    import agilerl.training.train_off_policy as r_train
    return dict(dqn=r_dqn, rb=r_rb, buf=r_buf, samp=r_samp, tree=r_tree, mut=r_mut, tour=r_tour, train=r_train)


def _ctor(cls):
    return [p for p in inspect.signature(cls.__init__).parameters if p != "self"]


def _public(cls):
    return {n for n, m in inspect.getmembers(cls) if not n.startswith("_") and (inspect.isfunction(m) or isinstance(m, property))}


def test_constructors_and_methods_match(ref):
    import agilerl_b200.algorithms as a
    import agilerl_b200.components as c
    import agilerl_b200.hpo as h
    pairs = [
        (ref["buf"].ReplayBuffer, c.ReplayBuffer), (ref["buf"].MultiStepReplayBuffer, c.MultiStepReplayBuffer),
        (ref["buf"].PrioritizedReplayBuffer, c.PrioritizedReplayBuffer), (ref["rb"].RainbowDQN, a.RainbowDQN),
        (ref["dqn"].DQN, a.DQN), (ref["tour"].TournamentSelection, h.TournamentSelection),
        (ref["mut"].Mutations, h.Mutations), (ref["tree"].SumSegmentTree, c.SumSegmentTree),
        (ref["tree"].MinSegmentTree, c.MinSegmentTree),
    ]
    for r, m in pairs:
        rc, mc = _ctor(r), _ctor(m)
        assert mc[:len(rc)] == rc or set(rc) <= set(mc), f"{r.__name__}: ctor {rc} vs {mc}"
    need = {
        c.ReplayBuffer: {"add", "sample", "clear", "storage", "size", "is_full"},
        c.MultiStepReplayBuffer: {"add", "sample_from_indices"},
        c.PrioritizedReplayBuffer: {"add", "sample", "update_priorities"},
        a.RainbowDQN: {"learn", "get_action", "test", "soft_update", "clone", "save_checkpoint", "load_checkpoint"},
        a.DQN: {"learn", "get_action", "test", "soft_update", "clone"},
        h.TournamentSelection: {"select"}, h.Mutations: {"mutation", "no_mutation", "architecture_mutate",
                                                         "parameter_mutation", "activation_mutation",
                                                         "rl_hyperparam_mutation"},
    }
    for cls, names in need.items():
        assert names <= _public(cls) | set(dir(cls)), f"{cls.__name__} lacks {names - set(dir(cls))}"


def test_reference_defaults_match(ref):
    import agilerl_b200.algorithms as a
    for r, m in ((ref["rb"].RainbowDQN, a.RainbowDQN), (ref["dqn"].DQN, a.DQN)):
        rs, ms = inspect.signature(r.__init__).parameters, inspect.signature(m.__init__).parameters
        for name, p in rs.items():
            if name in ("self", "device") or p.default is inspect.Parameter.empty:
                continue
            assert ms[name].default == p.default, f"{r.__name__}.{name}: {ms[name].default} != {p.default}"


def test_driver_signature_matches(ref):
    from agilerl_b200.training import train_off_policy as mine
    rs = list(inspect.signature(ref["train"].train_off_policy).parameters)
    ms = list(inspect.signature(mine).parameters)
    assert ms[:len(rs)] == rs, (rs, ms)


def test_multi_agent_surface_matches(ref):
    """MADDPG / MultiAgentReplayBuffer: same constructor parameters in the same order with the same
    defaults (``device`` aside), same public methods of the replay and the learner's hot-path methods."""
    import agilerl.algorithms.maddpg as r_ma
    import agilerl.components.multi_agent_replay_buffer as r_mb
    import agilerl_b200.algorithms as a
    import agilerl_b200.components as c
    assert _ctor(a.MADDPG) == _ctor(r_ma.MADDPG)
    assert _ctor(c.MultiAgentReplayBuffer) == _ctor(r_mb.MultiAgentReplayBuffer)
    rs, ms = inspect.signature(r_ma.MADDPG.__init__).parameters, inspect.signature(a.MADDPG.__init__).parameters
    for name, p in rs.items():
        if name in ("self", "device") or p.default is inspect.Parameter.empty:
            continue
        assert ms[name].default == p.default, f"MADDPG.{name}: {ms[name].default} != {p.default}"
    for name in ("save_to_memory", "save_to_memory_single_env", "save_to_memory_vect_envs", "sample"):
        r_sig = list(inspect.signature(getattr(r_mb.MultiAgentReplayBuffer, name)).parameters)
        m_sig = list(inspect.signature(getattr(c.MultiAgentReplayBuffer, name)).parameters)
        assert m_sig[:len(r_sig)] == r_sig, (name, r_sig, m_sig)
    assert {"learn", "get_action", "action_noise", "reset_action_noise", "soft_update", "test", "clone"} <= set(dir(a.MADDPG))
