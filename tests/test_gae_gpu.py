"""GPU parity of the PPO return / advantage scan (b2rl_gae_scan) — bit-exact against the golden vectors
recorded from the unmodified reference's RolloutBuffer and against the oracle at larger sizes."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["gae", "mc"])
def test_golden_rollout_bit_exact(mode):
    from agilerl_b200.components.rollout import compute_returns_and_advantages
    g = load_golden("gae_rollout.npz")
    adv, ret = compute_returns_and_advantages(torch.from_numpy(g[f"{mode}_rewards"]).cuda(),
                                              torch.from_numpy(g[f"{mode}_dones"].astype(bool)).cuda(),
                                              torch.from_numpy(g[f"{mode}_values"]).cuda(), g[f"{mode}_last_value"],
                                              g[f"{mode}_last_done"], float(g["gamma"]), float(g["gae_lambda"]),
                                              use_gae=(mode == "gae"))
    np.testing.assert_array_equal(adv.cpu().numpy(), g[f"{mode}_advantages"])
    np.testing.assert_array_equal(ret.cpu().numpy(), g[f"{mode}_returns"])


@pytest.mark.parametrize("T,E", [(1, 1), (2, 3), (128, 256), (2048, 256), (17, 1000)])
@pytest.mark.parametrize("use_gae", [True, False])
def test_sizes_against_oracle_bit_exact(T, E, use_gae):
    """Single step, ragged env counts, BASELINE config 4's 256 envs at a long horizon: identical bits."""
    from oracle import gae
    from agilerl_b200.components.rollout import compute_returns_and_advantages
    rng = np.random.default_rng(T * 1000 + E + int(use_gae))
    R = rng.standard_normal((T, E)).astype(np.float32)
    V = rng.standard_normal((T, E)).astype(np.float32)
    D = rng.random((T, E)) < 0.03
    lv = rng.standard_normal(E).astype(np.float32)
    ld = (rng.random(E) < 0.03).astype(np.float32)
    oa, orr = gae.compute_returns_and_advantages(R, D, V, lv, ld, 0.99, 0.95, use_gae)
    adv, ret = compute_returns_and_advantages(torch.from_numpy(R).cuda(), torch.from_numpy(D).cuda(),
                                              torch.from_numpy(V).cuda(), torch.from_numpy(lv).cuda(), ld, 0.99, 0.95, use_gae)
    np.testing.assert_array_equal(adv.cpu().numpy(), oa)
    np.testing.assert_array_equal(ret.cpu().numpy(), orr)


def test_requires_cuda_tensors():
    from agilerl_b200 import _lib
    from agilerl_b200.components.rollout import compute_returns_and_advantages
    with pytest.raises(_lib.B2RLError):
        compute_returns_and_advantages(torch.zeros(2, 2), torch.zeros(2, 2, dtype=torch.bool), torch.zeros(2, 2),
                                       np.zeros(2, np.float32), np.zeros(2, np.float32))
