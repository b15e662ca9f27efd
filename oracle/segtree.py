"""ORACLE (test infrastructure, never imported by the product path).

Two restatements of the reference's segment trees (agilerl/components/segment_tree.py):

* ``PySegTree`` — list-backed, pure Python, same cost profile as the reference (used by the
  ``cpu_baseline`` timing so the baseline pays what the reference pays);
* ``CSegTree``  — the C restatement in ``oracle/csrc/segtree.c`` through ctypes (fast enough to
  check 100k-leaf trees and thousands of updates bit-for-bit).

Both are pinned against the reference's *own* file, loaded by path, in
``tests/test_oracle_tree.py`` and against the golden vectors in ``tests/golden/tree_*.npz``.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle_segtree.so")
_lib = None


def load_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            from . import build as _b
            _b.build()
        lib = ctypes.CDLL(_LIB_PATH)
        d, i64, p = ctypes.c_double, ctypes.c_int64, ctypes.c_void_p
        lib.ost_init.argtypes = [p, i64, ctypes.c_int]
        lib.ost_set.argtypes = [p, i64, ctypes.c_int, i64, d]
        lib.ost_get.argtypes = [p, i64, i64]; lib.ost_get.restype = d
        lib.ost_operate.argtypes = [p, i64, ctypes.c_int, i64, i64]; lib.ost_operate.restype = d
        lib.ost_retrieve.argtypes = [p, i64, d]; lib.ost_retrieve.restype = i64
        lib.oper_update.argtypes = [p, p, i64, p, p, i64, d, p]
        lib.oper_sample.argtypes = [p, i64, p, i64, p]
        lib.oper_weights.argtypes = [p, p, i64, p, i64, d, i64, p]
        _lib = lib
    return _lib


class PySegTree:
    """segment_tree.py:5-108 restated (list of Python floats, recursion for range queries)."""

    def __init__(self, capacity: int, kind: str):
        assert capacity > 0 and capacity & (capacity - 1) == 0
        self.capacity, self.kind = capacity, kind
        self.op = (lambda a, b: a + b) if kind == "sum" else min
        self.tree = [0.0 if kind == "sum" else float("inf")] * (2 * capacity)

    def __setitem__(self, idx, val):
        i = idx + self.capacity
        t = self.tree
        t[i] = val
        i //= 2
        while i >= 1:
            t[i] = self.op(t[2 * i], t[2 * i + 1])
            i //= 2

    def __getitem__(self, idx):
        assert 0 <= idx < self.capacity
        return self.tree[self.capacity + idx]

    def _q(self, s, e, node, ns, ne):
        if s == ns and e == ne:
            return self.tree[node]
        mid = (ns + ne) // 2
        if e <= mid:
            return self._q(s, e, 2 * node, ns, mid)
        if mid + 1 <= s:
            return self._q(s, e, 2 * node + 1, mid + 1, ne)
        return self.op(self._q(s, mid, 2 * node, ns, mid), self._q(mid + 1, e, 2 * node + 1, mid + 1, ne))

    def operate(self, start=0, end=0):
        if end <= 0:
            end += self.capacity
        end -= 1
        return self._q(start, end, 1, 0, self.capacity - 1)

    sum = operate
    min = operate

    def retrieve(self, ub):
        assert 0 <= ub <= self.operate() + 1e-5
        idx, t = 1, self.tree
        while idx < self.capacity:
            left = 2 * idx
            if t[left] > ub:
                idx = left
            else:
                ub -= t[left]
                idx = left + 1
        return idx - self.capacity


class CSegTree:
    """Same tree in a numpy float64 array driven by the C restatement."""

    def __init__(self, capacity: int, kind: str):
        assert capacity > 0 and capacity & (capacity - 1) == 0
        self.capacity, self.kind = capacity, kind
        self.opc = 0 if kind == "sum" else 1
        self.tree = np.empty(2 * capacity, dtype=np.float64)
        self.lib = load_lib()
        self.lib.ost_init(self.tree.ctypes.data, capacity, self.opc)

    def __setitem__(self, idx, val):
        self.lib.ost_set(self.tree.ctypes.data, self.capacity, self.opc, int(idx), float(val))

    def __getitem__(self, idx):
        assert 0 <= idx < self.capacity
        return float(self.tree[self.capacity + idx])

    def operate(self, start=0, end=0):
        return self.lib.ost_operate(self.tree.ctypes.data, self.capacity, self.opc, start, end)

    sum = operate
    min = operate

    def retrieve(self, ub):
        return int(self.lib.ost_retrieve(self.tree.ctypes.data, self.capacity, float(ub)))
