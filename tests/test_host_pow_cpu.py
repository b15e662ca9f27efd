"""CPU: b2rl_host_priority_pow (the host half of PrioritizedReplayBuffer.update_priorities) is bit-identical to the
reference's per-element CPython arithmetic ``max(p.item(), 1e-5) ** alpha`` (replay_buffer.py:425, :322) and tracks
``max_priority`` (:329) the same way."""
import ctypes

import numpy as np


def test_host_pow_matches_cpython_pow_bit_for_bit():
    from agilerl_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    p = np.concatenate([np.abs(rng.standard_normal(50_000)).astype(np.float32),
                        np.array([0.0, 1e-7, 1e-5, 1.0000001e-5, 1.0, 3.5e4, 1e-30, 65504.0], np.float32),
                        (10.0 ** rng.uniform(-8, 6, 50_000)).astype(np.float32)])
    out = np.empty(p.size)
    for alpha in (0.6, 0.5, 1.0, 0.37, 0.0):
        mx = ctypes.c_double(1.0)
        assert lib.b2rl_host_priority_pow(p.ctypes.data, p.size, alpha, 1e-5, out.ctypes.data, ctypes.byref(mx)) == 0
        ref = np.array([max(float(x), 1e-5) ** alpha for x in p.tolist()])
        np.testing.assert_array_equal(out, ref)
        assert mx.value == max(1.0, max(max(float(x), 1e-5) for x in p.tolist()))
