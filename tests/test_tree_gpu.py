"""GPU parity: device fp64 trees + PER sampling vs the oracle and the reference's golden vectors
(bit-exact), called through the C ABI / the reference-shaped Python objects."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _trees(cap):
    from agilerl_b200.components.segment_tree import MinSegmentTree, SumSegmentTree
    return SumSegmentTree(cap), MinSegmentTree(cap)


def _set_both(st, mt, idx, val):
    from agilerl_b200 import _lib
    lib = _lib.load()
    i = torch.as_tensor(np.asarray(idx), dtype=torch.int64).cuda()
    v = torch.as_tensor(np.asarray(val), dtype=torch.float64).cuda()
    _lib.check(lib.b2rl_tree_set(st.data_ptr, mt.data_ptr, st.capacity, i.data_ptr(), v.data_ptr(), i.numel(),
                                 _lib.stream_ptr()))
    torch.cuda.synchronize()


# ---- the reference's own known-answer tests (tests/test_components/test_segment_tree.py) --------
def test_kat_tree_set():                      # :37-48
    t, _ = _trees(4)
    t[2] = 1.0
    t[3] = 3.0
    assert np.isclose(t.sum(), 4.0)
    assert np.isclose(t.sum(0, 2), 0.0)
    assert np.isclose(t.sum(0, 3), 1.0)
    assert np.isclose(t.sum(2, 3), 1.0)
    assert np.isclose(t.sum(2, -1), 1.0)
    assert np.isclose(t.sum(2, 4), 4.0)


def test_kat_tree_set_overlap():              # :51-61
    t, _ = _trees(4)
    t[2] = 1.0
    t[2] = 3.0
    assert np.isclose(t.sum(), 3.0)
    assert np.isclose(t.sum(2, 3), 3.0)
    assert np.isclose(t.sum(1, 2), 0.0)


def test_kat_prefixsum_idx():                 # :64-91
    t, _ = _trees(4)
    t[2] = 1.0
    t[3] = 3.0
    assert [t.retrieve(x) for x in (0.0, 0.5, 0.99, 1.01, 3.0, 4.0)] == [2, 2, 2, 3, 3, 3]
    t, _ = _trees(4)
    t[0] = 0.5; t[1] = 1.0; t[2] = 1.0; t[3] = 3.0
    assert [t.retrieve(x) for x in (0.0, 0.55, 0.99, 1.51, 3.0, 5.5)] == [0, 1, 1, 2, 3, 3]


def test_kat_min_tree():                      # :94-126
    _, t = _trees(4)
    t[0] = 1.0; t[2] = 0.5; t[3] = 3.0
    assert np.isclose(t.min(), 0.5) and np.isclose(t.min(0, 2), 1.0) and np.isclose(t.min(3, 4), 3.0)
    t[2] = 0.7
    assert np.isclose(t.min(), 0.7) and np.isclose(t.min(0, 3), 0.7)
    t[2] = 4.0
    assert np.isclose(t.min(), 1.0) and np.isclose(t.min(2, 4), 3.0) and np.isclose(t.min(2, 3), 4.0)


def test_invalid_capacity():
    from agilerl_b200.components.segment_tree import SumSegmentTree
    with pytest.raises(AssertionError):
        SumSegmentTree(6)
    with pytest.raises(AssertionError):
        SumSegmentTree(0)


# ---- golden vectors generated from the reference's segment_tree.py ------------------------------
def test_golden_tree_ops_bit_exact():
    g = load_golden("tree_ops.npz")
    cap = int(g["cap"])
    st, mt = _trees(cap)
    idx, val = g["idx"], g["val"]
    for k in range(len(g["snap_sum"])):      # batches of 50 sequential ops incl. duplicate indices
        _set_both(st, mt, idx[50 * k:50 * (k + 1)], val[50 * k:50 * (k + 1)])
        assert np.array_equal(np.array(st.tree), g["snap_sum"][k])
        assert np.array_equal(np.array(mt.tree), g["snap_min"][k])
    from agilerl_b200 import _lib
    ub = torch.from_numpy(g["ubs"]).cuda()
    out = torch.empty(len(g["ubs"]), dtype=torch.int64, device="cuda")
    _lib.check(_lib.load().b2rl_tree_retrieve(st.data_ptr, cap, ub.data_ptr(), ub.numel(), out.data_ptr(),
                                              _lib.stream_ptr()))
    assert np.array_equal(out.cpu().numpy(), g["retrieve"])
    for (a, b), rs, rm in zip(g["ranges"], g["range_sum"], g["range_min"]):
        assert st.sum(int(a), int(b)) == rs
        assert mt.min(int(a), int(b)) == rm


# ---- BASELINE-size tree against the oracle ------------------------------------------------------
@pytest.mark.parametrize("cap,live", [(131072, 100000), (1024, 1000), (2, 2), (1, 1)])
def test_full_size_updates_and_sampling_bit_exact(cap, live):
    from oracle.segtree import CSegTree, load_lib
    import ctypes
    from agilerl_b200 import _lib
    rng = np.random.default_rng(cap)
    st, mt = _trees(cap)
    os_, om = CSegTree(cap, "sum"), CSegTree(cap, "min")
    olib = load_lib()
    # bulk fill via the range path (PER.add of many rows), then random batched updates
    lib = _lib.load()
    p0 = 1.0 ** 0.6
    _lib.check(lib.b2rl_tree_set_range(st.data_ptr, mt.data_ptr, cap, 0, live, live, p0, _lib.stream_ptr()))
    for i in range(live):
        os_[i] = p0; om[i] = p0
    B = min(256, live)
    for it in range(20):
        idx = rng.integers(0, live, B)
        pri = (np.abs(rng.standard_normal(B)) + 1e-6).astype(np.float32)
        pa = np.array([max(float(p), 1e-5) ** 0.6 for p in pri])
        _set_both(st, mt, idx, pa)
        for i, v in zip(idx, pa):
            os_[int(i)] = float(v); om[int(i)] = float(v)
    assert np.array_equal(st._t.cpu().numpy()[1:], os_.tree[1:])
    assert np.array_equal(mt._t.cpu().numpy()[1:], om.tree[1:])
    # stratified sampling: identical indices, weights to fp32 rounding
    u = torch.rand(B)
    oidx = np.empty(B, dtype=np.int64); ow = np.empty(B, dtype=np.float32)
    un = u.numpy()
    olib.oper_sample(os_.tree.ctypes.data, cap, un.ctypes.data, B, oidx.ctypes.data)
    olib.oper_weights(os_.tree.ctypes.data, om.tree.ctypes.data, cap, oidx.ctypes.data, B, 0.4, live, ow.ctypes.data)
    ud = u.cuda()
    idx_d = torch.empty(B, dtype=torch.int64, device="cuda"); w_d = torch.empty(B, device="cuda")
    _lib.check(lib.b2rl_per_sample(st.data_ptr, mt.data_ptr, cap, ud.data_ptr(), B, 0.4, live, idx_d.data_ptr(),
                                   w_d.data_ptr(), _lib.stream_ptr()))
    assert np.array_equal(idx_d.cpu().numpy(), oidx)
    np.testing.assert_allclose(w_d.cpu().numpy(), ow, rtol=1e-6, atol=0)


def test_tree_invariants_full_size_device_pow():
    """Fused-path leaf update (device pow): leaves within 1 ulp of glibc pow, every internal node
    exactly op(children), max_priority folded."""
    from agilerl_b200 import _lib
    lib = _lib.load()
    cap, live, B = 131072, 100000, 256
    st, mt = _trees(cap)
    _lib.check(lib.b2rl_tree_set_range(st.data_ptr, mt.data_ptr, cap, 0, live, live, 1.0, _lib.stream_ptr()))
    rng = np.random.default_rng(0)
    idx = torch.from_numpy(rng.integers(0, live, B)).cuda()
    pri = torch.from_numpy((np.abs(rng.standard_normal(B)) * 3).astype(np.float32)).cuda()
    mp = torch.ones(1, dtype=torch.float64, device="cuda")
    _lib.check(lib.b2rl_tree_set_from_priorities(st.data_ptr, mt.data_ptr, cap, idx.data_ptr(), pri.data_ptr(), B,
                                                 0.6, 1e-5, mp.data_ptr(), _lib.stream_ptr()))
    s = st._t.cpu().numpy(); m = mt._t.cpu().numpy()
    n = np.arange(1, cap)
    assert np.array_equal(s[n], s[2 * n] + s[2 * n + 1])
    assert np.array_equal(m[n], np.minimum(m[2 * n], m[2 * n + 1]))
    pri_h = np.maximum(pri.cpu().numpy().astype(np.float64), 1e-5)
    assert mp.item() == max(1.0, pri_h.max())
    last = {int(i): k for k, i in enumerate(idx.cpu().numpy())}
    for i, k in last.items():
        ref = float(pri_h[k]) ** 0.6
        assert abs(s[cap + i] - ref) <= 2 * np.spacing(ref)


@pytest.mark.parametrize("cap", [2, 64, 4096, 131072])
@pytest.mark.parametrize("n", [1, 2, 31, 33, 255, 256, 257, 511, 512, 513, 1500])
def test_batched_update_duplicates_and_sizes_bit_exact(cap, n):
    """Batch sizes around the warp / fast-path limits (<= 512: partner-map kernel, above: level-by-level
    kernel), with heavy index collisions (small trees): leaves follow the last writer, every ancestor is
    bit-identical to n sequential __setitem__ calls of the reference (segment_tree.py:86-108)."""
    from oracle.segtree import CSegTree
    rng = np.random.default_rng(cap * 1000 + n)
    st, mt = _trees(cap)
    os_, om = CSegTree(cap, "sum"), CSegTree(cap, "min")
    for rnd in range(3):
        idx = rng.integers(0, cap, n)
        if rnd == 1:
            idx[: n // 2] = idx[0]                      # one leaf hit many times
        val = np.abs(rng.standard_normal(n)) + 1e-9
        _set_both(st, mt, idx, val)
        for i, v in zip(idx, val):
            os_[int(i)] = float(v); om[int(i)] = float(v)
        assert np.array_equal(st._t.cpu().numpy()[1:], os_.tree[1:])
        assert np.array_equal(mt._t.cpu().numpy()[1:], om.tree[1:])
