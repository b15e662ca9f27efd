"""CPU: the GAE / Monte-Carlo return oracle (SURVEY §8(f) rank 2 groundwork) against the golden vectors
recorded from the unmodified reference's RolloutBuffer (rollout_buffer.py:413-481), bit-exact, plus the
properties a device scan will be checked with at full size."""
import numpy as np
import pytest

from conftest import load_golden


@pytest.mark.parametrize("mode", ["gae", "mc"])
def test_oracle_matches_reference_rollout_buffer(mode):
    from oracle import gae
    g = load_golden("gae_rollout.npz")
    adv, ret = gae.compute_returns_and_advantages(g[f"{mode}_rewards"], g[f"{mode}_dones"].astype(bool), g[f"{mode}_values"],
                                                  g[f"{mode}_last_value"], g[f"{mode}_last_done"], float(g["gamma"]),
                                                  float(g["gae_lambda"]), use_gae=(mode == "gae"))
    assert adv.dtype == np.float32 and ret.dtype == np.float32
    np.testing.assert_array_equal(adv, g[f"{mode}_advantages"])
    np.testing.assert_array_equal(ret, g[f"{mode}_returns"])


def test_gae_properties():
    from oracle import gae
    rng = np.random.default_rng(0)
    T, E = 40, 5
    R = rng.standard_normal((T, E)).astype(np.float32)
    V = rng.standard_normal((T, E)).astype(np.float32)
    D = np.zeros((T, E), dtype=bool)
    lv = rng.standard_normal(E).astype(np.float32)
    ld = np.zeros(E, dtype=np.float32)
    # lambda = 1, no terminations: advantages telescope to the discounted return minus the value
    adv, ret = gae.compute_returns_and_advantages(R, D, V, lv, ld, 0.9, 1.0, True)
    disc = np.zeros((T, E))
    run = lv.astype(float)
    for t in reversed(range(T)):
        run = R[t] + 0.9 * run
        disc[t] = run
    np.testing.assert_allclose(adv, disc - V, rtol=2e-6, atol=2e-6)
    np.testing.assert_array_equal(ret, adv + V)
    # a done at t+1 cuts the recurrence: rows <= t do not depend on anything after t
    D2 = D.copy(); D2[20] = True
    a1, _ = gae.compute_returns_and_advantages(R, D2, V, lv, ld, 0.9, 0.95, True)
    R3 = R.copy(); R3[25:] += 7.0
    a2, _ = gae.compute_returns_and_advantages(R3, D2, V, lv, ld, 0.9, 0.95, True)
    np.testing.assert_array_equal(a1[:20], a2[:20])
