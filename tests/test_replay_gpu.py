"""GPU parity: HBM replay buffers (ring write, n-step roll, PER) through the reference-shaped
Python API vs golden vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _build_from_golden(g):
    from agilerl_b200.components import MultiStepReplayBuffer, PrioritizedReplayBuffer, Transition
    E = int(g["E"])
    mem = PrioritizedReplayBuffer(int(g["max_size"]), alpha=float(g["alpha"]), device="cuda")
    nmem = MultiStepReplayBuffer(int(g["max_size"]), n_step=int(g["n_step"]), gamma=float(g["gamma"]), device="cuda")
    T = g["raw_obs"].shape[0]
    for t in range(T):
        td = Transition(obs=g["raw_obs"][t], action=g["raw_action"][t], reward=g["raw_reward"][t],
                        next_obs=g["raw_next_obs"][t], done=g["raw_done"][t], batch_size=[E]).to_tensordict()
        one = nmem.add(td)
        if one is not None:
            mem.add(one)
    return mem, nmem


def test_golden_ingest_nstep_and_per_storage():
    g = load_golden("replay_per_nstep.npz")
    mem, nmem = _build_from_golden(g)
    for k in ("obs", "action", "reward", "next_obs", "done"):
        assert np.array_equal(mem.storage[k].cpu().numpy(), g[f"per_{k}"]), k
        assert np.array_equal(nmem.storage[k].cpu().numpy(), g[f"nstep_{k}"]), k
        assert mem.storage[k].dtype == torch.from_numpy(g[f"per_{k}"]).dtype
    assert (mem._cursor, mem._size, mem.tree_ptr) == (int(g["cursor"]), int(g["size"]), int(g["tree_ptr"]))
    assert nmem.done_key == "done"


def test_golden_priority_updates_and_sampling_bit_exact(monkeypatch):
    g = load_golden("replay_per_nstep.npz")
    mem, nmem = _build_from_golden(g)
    for r in range(3):
        # idxs [B,1] like sample() returns, priorities as numpy f32 like learn() returns
        mem.update_priorities(torch.from_numpy(g["upd_idx"][r]).unsqueeze(1), g["upd_pri"][r])
        assert np.array_equal(np.array(mem.sum_tree.tree), g["sum_after"][r])
        assert np.array_equal(np.array(mem.min_tree.tree), g["min_after"][r])
    assert mem.max_priority == float(g["max_priority"])
    u = torch.from_numpy(g["uniforms"])
    monkeypatch.setattr(torch, "rand", lambda *a, **k: u.clone())
    batch = mem.sample(len(u), float(g["beta"]))
    assert batch["idxs"].shape == (len(u), 1) and batch["idxs"].dtype == torch.int64
    assert batch["weights"].shape == (len(u), 1) and batch["weights"].dtype == torch.float32
    assert np.array_equal(batch["idxs"].cpu().numpy(), g["sample_idxs"])
    np.testing.assert_allclose(batch["weights"].cpu().numpy(), g["sample_weights"], rtol=1e-6)
    for k in ("obs", "action", "reward", "next_obs", "done"):
        assert np.array_equal(batch[k].cpu().numpy(), g[f"batch_{k}"]), k
    nb = nmem.sample_from_indices(batch["idxs"].squeeze(1))
    for k in ("obs", "action", "reward", "next_obs", "done"):
        assert np.array_equal(nb[k].cpu().numpy(), g[f"nbatch_{k}"]), k
    # driver shape (quirk Q2): [B,1] idxs give [B,1,...] fields
    nb2 = nmem.sample_from_indices(batch["idxs"])
    assert nb2["reward"].shape == (len(u), 1, 1) and nb2["obs"].shape[:2] == (len(u), 1)


def test_golden_ring_wrap_layout():
    """tests/test_components/test_replay_buffer.py:144-177."""
    from agilerl_b200.components import ReplayBuffer
    from agilerl_b200.compat import TensorDict
    g = load_golden("replay_ring.npz")
    rb = ReplayBuffer(max_size=3, device="cuda")
    rb.add(TensorDict({"obs": torch.tensor([[1.0, 2.0], [3.0, 4.0]]), "reward": torch.tensor([1.0, 2.0])}, batch_size=[2]))
    rb.add(TensorDict({"obs": torch.tensor([[5.0, 6.0], [7.0, 8.0]]), "reward": torch.tensor([3.0, 4.0])}, batch_size=[2]))
    assert np.array_equal(rb.storage["obs"].cpu().numpy(), g["obs"])
    assert np.array_equal(rb.storage["reward"].cpu().numpy(), g["reward"])
    assert rb.storage["reward"].shape == (3, 1)      # 1-D leaves reshaped to (n, 1)
    assert (rb._cursor, rb._size, rb.counter) == (int(g["cursor"]), int(g["size"]), int(g["counter"]))
    assert rb.is_full and len(rb) == 3
    s = rb.sample(2, return_idx=True)
    assert s["idxs"].shape == (2,) and s["obs"].shape == (2, 2)
    assert len(set(s["idxs"].tolist())) == 2          # without replacement
    rb.clear()
    assert len(rb) == 0 and rb.storage is None and not rb.initialized


def test_uniform_sample_of_a_large_buffer_keeps_the_reference_stream():
    """ReplayBuffer.sample (replay_buffer.py:114-131) on a buffer large enough for the host prefix helper
    (b2rl_host_randperm_prefix): same rows as ``storage[torch.randperm(size)[:B]]`` and the same generator state afterwards."""
    from agilerl_b200.components import ReplayBuffer, replay_buffer as rbm
    from agilerl_b200.compat import TensorDict
    n, B = 20_000, 64
    rb = ReplayBuffer(max_size=n, device="cuda")
    rows = torch.arange(n, dtype=torch.float32)
    rb.add(TensorDict({"obs": torch.stack([rows, -rows], dim=1), "reward": rows * 0.5}, batch_size=[n]))
    torch.manual_seed(123)
    want = [torch.randperm(n)[:B] for _ in range(3)]
    want_next = torch.rand(4)
    torch.manual_seed(123)
    got = [rb.sample(B, return_idx=True) for _ in range(3)]
    got_next = torch.rand(4)
    for w, g in zip(want, got):
        assert torch.equal(g["idxs"].cpu(), w)
        assert torch.equal(g["obs"].cpu(), torch.stack([w.float(), -w.float()], dim=1))
        assert torch.equal(g["reward"].cpu().reshape(-1), w.float() * 0.5)
    assert torch.equal(want_next, got_next)
    if rbm._RANDPERM_FAST is not True:                 # either path must give the rows above; say which one ran
        pytest.skip("b2rl_host_randperm_prefix fell back to torch.randperm on this box")


def test_nstep_known_answers():
    """test_replay_buffer.py:499-594 — r1 + g r2 + g^2 r3, and early termination."""
    from agilerl_b200.components import MultiStepReplayBuffer
    from agilerl_b200.compat import TensorDict

    def mk(i, done):
        return TensorDict({"state": torch.tensor([[i, i + 1, i + 2]]), "action": torch.tensor([[i]]),
                           "reward": torch.tensor([[float(i + 1)]]), "next_state": torch.tensor([[i + 3, i + 4, i + 5]]),
                           "done": torch.tensor([[done]])}, batch_size=[1])
    buf = MultiStepReplayBuffer(max_size=1000, n_step=3, gamma=0.9, device="cuda")
    buf.ns_key = "next_state"
    for i in range(3):
        buf.n_step_buffer.append(mk(i, False).to("cuda"))
    out = buf._get_n_step_info()
    assert torch.isclose(out["reward"].cpu(), torch.tensor(1.0 + 0.9 * 2.0 + 0.9 * 0.9 * 3.0)).all()
    assert torch.equal(out["next_state"].cpu(), torch.tensor([[5, 6, 7]]))
    buf = MultiStepReplayBuffer(max_size=1000, n_step=3, gamma=0.9, device="cuda")
    buf.ns_key = "next_state"
    for i, d in enumerate([False, True, False]):
        buf.n_step_buffer.append(mk(i, d).to("cuda"))
    out = buf._get_n_step_info()
    assert torch.isclose(out["reward"].cpu(), torch.tensor(1.0 + 0.9 * 2.0)).all()
    assert torch.equal(out["next_state"].cpu(), torch.tensor([[4, 5, 6]]))
    assert bool(out["done"].cpu().all())


def test_per_leaf_values_and_floor():
    """test_replay_buffer.py:650-690, 801-836, 1037-1051: leaf == p**alpha exactly, 1e-5 floor."""
    from agilerl_b200.components import PrioritizedReplayBuffer
    from agilerl_b200.compat import TensorDict
    buf = PrioritizedReplayBuffer(max_size=10, alpha=0.6, device="cuda")
    for i in range(3):
        buf.add(TensorDict({"state": torch.tensor([[i]]), "action": torch.tensor([[0]]), "reward": torch.tensor([[1.0]])},
                           batch_size=[1]))
    assert buf.tree_ptr == 3 and buf.sum_tree[0] == 1.0 ** 0.6
    buf.update_priorities(torch.tensor([0, 1]), torch.tensor([1e-10, 4.0]))
    assert buf.sum_tree[0] == 1e-5 ** 0.6
    assert buf.sum_tree[1] == 4.0 ** 0.6 and buf.min_tree[1] == 4.0 ** 0.6
    assert buf.max_priority == 4.0
    buf._update_priority(2, 2.5)
    assert buf.sum_tree[2] == 2.5 ** 0.6
    with pytest.raises(AssertionError):
        buf._update_priority(10, 1.0)
    w = buf._calculate_weights(torch.tensor([1]), 0.4)
    p = (4.0 ** 0.6) / buf.sum_tree.sum(); pmin = buf.min_tree.min() / buf.sum_tree.sum()
    assert torch.isclose(w.cpu()[0], torch.tensor(((p * 3) ** -0.4) / ((pmin * 3) ** -0.4), dtype=torch.float32), rtol=1e-5)


def test_full_size_gather_round_trip():
    """BASELINE config-2 shapes: 84x84x4 uint8 frames, B=256 — gather(ring_write(x)) == x[idx]."""
    from agilerl_b200.components import ReplayBuffer
    from agilerl_b200.compat import TensorDict
    N, B = 4096, 256
    g = torch.Generator().manual_seed(0)
    frames = torch.randint(0, 256, (N, 4, 84, 84), dtype=torch.uint8, generator=g)
    rew = torch.randn(N, generator=g)
    rb = ReplayBuffer(max_size=N, device="cuda")
    for s in range(0, N, 1000):   # ragged adds, last one wraps nothing but is short
        e = min(N, s + 1000)
        rb.add(TensorDict({"obs": frames[s:e], "reward": rew[s:e]}, batch_size=[e - s]))
    rb.add(TensorDict({"obs": frames[:7], "reward": rew[:7]}, batch_size=[7]))   # wraps to slot 0..6
    idx = torch.randint(0, N, (B,), generator=g)
    out = rb._gather(idx)
    assert torch.equal(out["obs"].cpu(), frames[idx])
    assert torch.equal(out["reward"].cpu(), rew[idx].unsqueeze(1))


def test_device_uniform_sampling_is_distinct_uniform_and_deterministic():
    """b2rl_sample_uniform_distinct (the HBM-resident counterpart of randperm(size)[:B], replay_buffer.py:125): B
    distinct in-range indices, identical for identical (seed, offset), different across offsets, and uniform over
    the range (chi-square over 16 bins of 200 x 512 draws)."""
    import ctypes
    from agilerl_b200 import _lib
    lib = _lib.load()
    s = _lib.stream_ptr(torch.device("cuda:0"))

    def draw(N, B, off):
        out = torch.empty(B, dtype=torch.int64, device="cuda")
        _lib.check(lib.b2rl_sample_uniform_distinct(77, off, N, B, out.data_ptr(), s))
        return out.cpu().numpy()
    for N, B in [(1_000_000, 512), (600, 512), (1024, 1024), (7, 7), (5, 1)]:
        a = draw(N, B, 0)
        assert len(set(a.tolist())) == B and a.min() >= 0 and a.max() < N
        assert np.array_equal(a, draw(N, B, 0))
        if N > B:
            assert not np.array_equal(a, draw(N, B, 64 * B))
    counts = np.zeros(16)
    for k in range(200):
        counts += np.bincount(draw(1_000_000, 512, 64 * 512 * k) * 16 // 1_000_000, minlength=16)
    exp = 200 * 512 / 16
    chi2 = float(((counts - exp) ** 2 / exp).sum())
    assert chi2 < 45.0, chi2            # 15 degrees of freedom: P(chi2 > 45) < 1e-4


@pytest.mark.gpu
def test_max_priority_host_device_ownership():
    """add() after a device-side priority update must not read the device scalar back, and an assigned
    max_priority must not be overridden by a stale device maximum."""
    from agilerl_b200.components import PrioritizedReplayBuffer, Transition
    buf = PrioritizedReplayBuffer(64, alpha=0.6, device="cuda")

    def tr(n):
        return Transition(obs=torch.zeros(n, 3), action=torch.zeros(n), reward=torch.zeros(n), next_obs=torch.zeros(n, 3),
                          done=torch.zeros(n), batch_size=[n]).to_tensordict()

    buf.add(tr(8))
    idx = torch.arange(4, device="cuda")
    buf.update_priorities_device(idx, torch.tensor([0.5, 7.0, 2.0, 3.0], device="cuda"))
    assert buf._dev_dirty
    buf.add(tr(2))                                   # leaves 8, 9 = pow(max(1.0, 7.0), alpha) without a host read
    assert buf._dev_dirty                            # ... and without reconciling
    torch.cuda.synchronize()
    leaves = torch.tensor(buf.sum_tree.tree[buf._cap + 8:buf._cap + 10], dtype=torch.float64)
    assert torch.allclose(leaves, torch.full((2,), 7.0 ** 0.6, dtype=torch.float64), rtol=1e-15)
    assert buf.max_priority == 7.0 and not buf._dev_dirty
    buf.max_priority = 2.0                           # user override: the old device maximum must not come back
    buf.update_priorities_device(idx[:1], torch.tensor([1.5], device="cuda"))
    assert buf.max_priority == 2.0
    buf.add(tr(1))
    torch.cuda.synchronize()
    assert abs(buf.sum_tree.tree[buf._cap + 10] - 2.0 ** 0.6) < 1e-15
