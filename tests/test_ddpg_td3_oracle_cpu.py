"""CPU: the DDPG / TD3 oracle (SURVEY §8(f) rank 1 groundwork) against the golden vectors recorded from
the unmodified reference (tests/golden/make_golden.py::gen_ddpg_td3): four consecutive learn calls,
bit-exact losses and parameters.  No CUDA path exists for these learners yet — this pins the oracle the
next round builds against."""
import numpy as np
import pytest
import torch

from conftest import load_golden


def _sd(g, tag):
    return {k[len(tag) + 1:]: torch.from_numpy(g[k].copy()) for k in g.files if k.startswith(tag + "/")}


@pytest.mark.parametrize("name", ["ddpg", "td3"])
def test_oracle_reproduces_reference_learn_calls(name):
    from oracle import ddpg_td3 as od
    g = load_golden(f"{name}_vector.npz")
    twin = bool(int(g["twin"]))
    n_c = 2 if twin else 1
    a_specs = od.actor_specs(17, 6, head_hidden=[int(h) for h in g["a_hidden"]])
    c_specs = od.critic_specs(17, 6, head_hidden=[int(h) for h in g["c_hidden"]])
    orc = od.OracleDDPG(a_specs, c_specs, _sd(g, "actor0"), _sd(g, "actor_target0"),
                        [_sd(g, f"critic{i}_0") for i in range(n_c)], [_sd(g, f"critic_target{i}_0") for i in range(n_c)],
                        gamma=float(g["gamma"]), tau=float(g["tau"]), lr_actor=float(g["lr_actor"]),
                        lr_critic=float(g["lr_critic"]), policy_freq=int(g["policy_freq"]), twin=twin)
    torch.set_num_threads(1)
    for st in range(int(g["steps"])):
        exp = {k: torch.from_numpy(g[f"s{st}_{k}"].copy()) for k in ("obs", "action", "reward", "next_obs", "done")}
        torch.manual_seed(int(g[f"s{st}_seed"]))
        a_loss, c_loss = orc.learn(exp)
        # quirk: the batch's action tensor was overwritten in place with the target-policy noise
        np.testing.assert_array_equal(exp["action"].numpy(), g[f"s{st}_noise"])
        ref_a = float(g[f"s{st}_actor_loss"])
        assert (a_loss is None) == bool(np.isnan(ref_a)), f"step {st}: actor update cadence"
        if a_loss is not None:
            assert a_loss == ref_a
        assert c_loss == float(g[f"s{st}_critic_loss"])
    for k, v in _sd(g, "actor1").items():
        assert torch.equal(orc.actor[k].data, v), f"actor {k}"
    for k, v in _sd(g, "actor_target1").items():
        assert torch.equal(orc.actor_target[k], v), f"actor_target {k}"
    for i in range(n_c):
        for k, v in _sd(g, f"critic{i}_1").items():
            assert torch.equal(orc.critics[i][k].data, v), f"critic{i} {k}"
        for k, v in _sd(g, f"critic_target{i}_1").items():
            assert torch.equal(orc.critic_targets[i][k], v), f"critic_target{i} {k}"


def test_actor_and_targets_move_every_policy_freq_calls():
    from oracle import ddpg_td3 as od
    g = load_golden("td3_vector.npz")
    losses = [float(g[f"s{st}_actor_loss"]) for st in range(int(g["steps"]))]
    assert [np.isnan(x) for x in losses] == [True, False, True, False]        # policy_freq = 2 (td3.py:520)
