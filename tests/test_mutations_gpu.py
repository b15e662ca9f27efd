"""GPU: ``Mutations`` against golden vectors recorded from the UNMODIFIED reference (tests/golden/make_golden.py::
gen_mutations, agilerl/hpo/mutation.py): everything the reference's SEEDED generators decide is reproduced —
the mutation drawn for each population member, the Gaussian parameter mutation of a RainbowQNetwork state_dict
(bit for bit, including the 2-D NoisyLinear epsilon buffers the reference also walks), the RL hyper-parameter
mutation sequence and the activation picks.  Not pinnable: architecture mutations — the reference orders its
mutation-method list through ``list(set(...))`` (modules/base.py:570-571: hash-randomised per process) and each
module draws from its own unseeded generator."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu

NET = {"encoder_config": {"channel_size": [8, 16], "kernel_size": [4, 3], "stride_size": [2, 1]},
       "head_config": {"hidden_size": [32]}, "latent_dim": 16}


def _agent(**kw):
    from agilerl_b200.algorithms import RainbowDQN
    from agilerl_b200.compat import spaces
    return RainbowDQN(spaces.Box(0, 255, (3, 20, 20), np.uint8), spaces.Discrete(4), net_config=dict(NET), v_min=-10.0,
                      v_max=10.0, **kw)


def test_mutation_choice_per_member_matches_reference():
    from agilerl_b200.hpo import Mutations
    g = load_golden("mutations.npz")
    for c in range(3):
        p = g[f"choice{c}_probs"]
        m = Mutations(no_mutation=p[0], architecture=p[1], new_layer_prob=0.5, parameters=p[2], activation=p[3], rl_hp=p[4],
                      rand_seed=int(g[f"choice{c}_seed"]))
        names = [f.__name__ for f in m.rng.choice(m.mut_options, 12, p=m.mut_proba)]
        pre = [f.__name__ for f in m.rng.choice(m.pretraining_mut_options, 12, p=m.pretraining_mut_proba)]
        assert names == list(g[f"choice{c}_names"]) and pre == list(g[f"choice{c}_pre"])


@pytest.mark.parametrize("c", [0, 1])
def test_gaussian_parameter_mutation_bit_exact(c):
    from agilerl_b200.hpo import Mutations
    g = load_golden("mutations.npz")
    seed = int(g[f"gauss{c}_seed"])
    agent = _agent(batch_size=8)
    before = {k[len(f"gauss{c}_before/"):]: torch.from_numpy(g[k].copy()) for k in g.files if k.startswith(f"gauss{c}_before/")}
    after = {k[len(f"gauss{c}_after/"):]: g[k] for k in g.files if k.startswith(f"gauss{c}_after/")}
    agent.actor.load_state_dict(before)
    m = Mutations(0, 0, 0.5, 1, 0, 0, mutation_sd=0.1, rand_seed=seed)
    torch.manual_seed(500 + seed)
    m._gaussian_parameter_mutation(agent.actor)
    sd = agent.actor.state_dict()
    changed = 0
    for k, v in after.items():
        np.testing.assert_array_equal(sd[k].cpu().numpy(), v, err_msg=k)
        changed += int((v != before[k].numpy()).sum())
    assert changed > 100


def test_rl_hyperparameter_mutation_sequence_matches_reference():
    from agilerl_b200.algorithms.core.registry import HyperparameterConfig, RLParameter
    from agilerl_b200.hpo import Mutations
    g = load_golden("mutations.npz")
    hp = HyperparameterConfig(lr=RLParameter(min=1e-5, max=1e-2), batch_size=RLParameter(min=8, max=64, dtype=int),
                              learn_step=RLParameter(min=1, max=16, dtype=int, grow_factor=1.5, shrink_factor=0.75))
    agent = _agent(hp_config=hp, batch_size=16, lr=1e-3, learn_step=4)
    m = Mutations(0, 0, 0.5, 0, 0, 1, rand_seed=5)
    torch.manual_seed(900)
    for i in range(len(g["rlhp_mut"])):
        agent = m.rl_hyperparam_mutation(agent)
        assert agent.mut == str(g["rlhp_mut"][i]), i
        assert (float(agent.lr), int(agent.batch_size), int(agent.learn_step)) == tuple(g["rlhp_vals"][i]), i
    # the activation picks continue from the same agent in the fixture
    m = Mutations(0, 0, 0.5, 0, 1, 0, activation_selection=["ReLU", "ELU", "GELU"], rand_seed=9)
    for i in range(len(g["act_seq"])):
        agent = m.activation_mutation(agent)
        assert agent.mut == "act" and agent.actor.activation == str(g["act_seq"][i]), i
        assert agent.actor_target.activation == agent.actor.activation
