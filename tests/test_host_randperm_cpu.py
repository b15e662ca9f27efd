"""CPU: b2rl_host_randperm_prefix (the index draw of ReplayBuffer.sample, replay_buffer.py:126) returns exactly
``torch.randperm(n)[:B]`` AND leaves torch's global CPU generator in exactly the state ``torch.randperm(n)`` leaves it in —
so a seeded run samples the same transitions and every later draw (noise, epsilon, the next sample) is unchanged."""
import warnings

import numpy as np
import pytest
import torch


def _fast(lib, n, B):
    from agilerl_b200.components.replay_buffer import _randperm_prefix_fast
    return _randperm_prefix_fast(lib, n, B)


def test_prefix_and_generator_state_match_torch_randperm():
    from agilerl_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    cases = [(int(n), int(rng.integers(0, n + 1))) for n in rng.integers(1, 6000, size=200)]
    cases += [(1, 0), (1, 1), (2, 2), (2, 1), (624, 624), (625, 3), (100_000, 256), (1_000_000, 512), (1_000_000, 0)]
    for k, (n, B) in enumerate(cases):
        seed, pre = int(rng.integers(0, 2 ** 31)), int(rng.integers(0, 1500))      # `pre`: any position inside a twist block
        torch.manual_seed(seed); torch.rand(pre)
        want = torch.randperm(n)[:B]; want_state = torch.get_rng_state(); want_next = torch.rand(5)
        torch.manual_seed(seed); torch.rand(pre)
        got = _fast(lib, n, B); got_state = torch.get_rng_state(); got_next = torch.rand(5)
        assert torch.equal(want, got), (n, B, seed, pre)
        assert torch.equal(want_state, got_state) and torch.equal(want_next, got_next), (n, B, seed, pre)


def test_consecutive_samples_stay_in_step_with_the_reference_stream():
    from agilerl_b200 import _lib
    from agilerl_b200.components import replay_buffer as rb
    lib = _lib.load()
    rb._RANDPERM_FAST = None
    torch.manual_seed(11)
    want = [torch.randperm(50_000)[:64] for _ in range(20)] + [torch.rand(4)]
    torch.manual_seed(11)
    got = [rb.randperm_prefix(lib, 50_000, 64) for _ in range(20)] + [torch.rand(4)]
    assert rb._RANDPERM_FAST is True                       # the once-per-process check passed and left the stream untouched
    assert all(torch.equal(a, b) for a, b in zip(want, got))


def test_wrapper_keeps_torch_randperm_where_the_helper_does_not_apply():
    from agilerl_b200 import _lib
    from agilerl_b200.components import replay_buffer as rb
    lib = _lib.load()

    class NoHelper:                                      # a binding without the entry point (the CPU stand-ins of the driver tests)
        pass
    for args in [(NoHelper(), 50_000, 64), (lib, 100, 8), (lib, 50_000, 50_000), (lib, 5000, 6000)]:
        torch.manual_seed(3); want = torch.randperm(args[1])[:args[2]]; s = torch.get_rng_state()
        torch.manual_seed(3); got = rb.randperm_prefix(*args)
        assert torch.equal(want, got) and torch.equal(s, torch.get_rng_state()), args[1:]
    # a helper that disagrees with torch is dropped for the rest of the process, with a warning, and the call still answers
    class Wrong:
        def b2rl_host_randperm_prefix(self, st, nbytes, n, B, out):
            return 0                                       # leaves `out` and the state untouched: cannot match
    rb._RANDPERM_FAST = None
    torch.manual_seed(4); want = torch.randperm(50_000)[:16]
    torch.manual_seed(4)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        got = rb.randperm_prefix(Wrong(), 50_000, 16)
    assert rb._RANDPERM_FAST is False and torch.equal(want, got) and any("torch.randperm" in str(x.message) for x in w)
    rb._RANDPERM_FAST = None


def test_bad_arguments_are_refused():
    from agilerl_b200 import _lib
    lib = _lib.load()
    st = torch.get_rng_state()
    out = torch.empty(8, dtype=torch.int64)
    assert lib.b2rl_host_randperm_prefix(st.data_ptr(), 100, 10, 4, out.data_ptr()) != 0          # state blob too short
    assert lib.b2rl_host_randperm_prefix(st.data_ptr(), st.numel(), 4, 8, out.data_ptr()) != 0     # B > n
    assert lib.b2rl_host_randperm_prefix(st.data_ptr(), st.numel(), 2 ** 31, 8, out.data_ptr()) != 0   # torch's 64-bit branch
    assert torch.equal(st, torch.get_rng_state())
