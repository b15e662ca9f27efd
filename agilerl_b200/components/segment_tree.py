"""Device-resident segment trees with the reference's object interface.

Mirror of ``agilerl/components/segment_tree.py`` (``SegmentTree`` :5-108, ``SumSegmentTree``
:111-156, ``MinSegmentTree`` :159-182): same constructor asserts, ``tree[i]``, ``tree[i] = v``,
``sum()/min()`` range queries, ``retrieve(upperbound)``, ``capacity`` and ``tree``.  The array
heap (2*capacity fp64) lives in HBM and is updated by ``b2rl_tree_set`` (csrc/tree.cu); the hot
path never goes through these per-element accessors — ``PrioritizedReplayBuffer`` drives the
batched kernels directly — they exist so code and tests written against the reference objects
keep working.
"""
from __future__ import annotations

import operator
from collections.abc import Callable

import torch

from .. import _lib


class SegmentTree:
    def __init__(self, capacity: int, operation: Callable, init_value: float, device="cuda",
                 _storage: torch.Tensor | None = None) -> None:
        assert capacity > 0, "capacity must be positive and a power of 2."
        assert capacity & (capacity - 1) == 0, "capacity must be positive and a power of 2."
        if operation not in (operator.add, min):
            raise NotImplementedError("device segment trees implement operator.add and min only")
        self.capacity = capacity
        self.operation = operation
        self.device = _lib.as_device(device)
        self._is_sum = operation is operator.add
        if _storage is None:
            _storage = torch.full((2 * capacity,), float(init_value), dtype=torch.float64, device=self.device)
        self._t = _storage

    # -- raw views -------------------------------------------------------------------------
    @property
    def tree(self) -> list[float]:
        """Host copy of the array heap (the reference attribute is a Python list)."""
        return self._t.cpu().tolist()

    @property
    def data_ptr(self) -> int:
        return self._t.data_ptr()

    def root(self) -> float:
        return float(self._t[1].item())

    # -- reference API ---------------------------------------------------------------------
    def _operate_helper(self, tree, start, end, node, node_start, node_end) -> float:
        if start == node_start and end == node_end:
            return tree[node]
        mid = (node_start + node_end) // 2
        if end <= mid:
            return self._operate_helper(tree, start, end, 2 * node, node_start, mid)
        if mid + 1 <= start:
            return self._operate_helper(tree, start, end, 2 * node + 1, mid + 1, node_end)
        return self.operation(
            self._operate_helper(tree, start, mid, 2 * node, node_start, mid),
            self._operate_helper(tree, mid + 1, end, 2 * node + 1, mid + 1, node_end),
        )

    def operate(self, start: int = 0, end: int = 0) -> float:
        if end <= 0:
            end += self.capacity
        end -= 1
        if start == 0 and end == self.capacity - 1:
            return self.root()          # full range == root; one 8-byte read
        # partial ranges are off the hot path: walk a host snapshot in the reference's order so
        # the fp64 combination order (and hence the bits) match
        return self._operate_helper(self.tree, start, end, 1, 0, self.capacity - 1)

    def __setitem__(self, idx: int, val: float) -> None:
        idx_t = torch.tensor([int(idx)], dtype=torch.int64, device=self.device)
        val_t = torch.tensor([float(val)], dtype=torch.float64, device=self.device)
        lib = _lib.load()
        s, m = (self._t.data_ptr(), None) if self._is_sum else (None, self._t.data_ptr())
        _lib.check(lib.b2rl_tree_set(s, m, self.capacity, idx_t.data_ptr(), val_t.data_ptr(), 1,
                                     _lib.stream_ptr(self.device)))

    def __getitem__(self, idx: int) -> float:
        idx = int(idx)
        assert 0 <= idx < self.capacity
        return float(self._t[self.capacity + idx].item())


class SumSegmentTree(SegmentTree):
    def __init__(self, capacity: int, device="cuda", _storage=None) -> None:
        super().__init__(capacity=capacity, operation=operator.add, init_value=0.0, device=device,
                         _storage=_storage)

    def sum(self, start: int = 0, end: int = 0) -> float:
        return super().operate(start, end)

    def retrieve(self, upperbound: float) -> int:
        assert 0 <= upperbound <= self.sum() + 1e-5, f"upperbound: {upperbound}"
        ub = torch.tensor([float(upperbound)], dtype=torch.float64, device=self.device)
        out = torch.empty(1, dtype=torch.int64, device=self.device)
        _lib.check(_lib.load().b2rl_tree_retrieve(self._t.data_ptr(), self.capacity, ub.data_ptr(), 1,
                                                  out.data_ptr(), _lib.stream_ptr(self.device)))
        return int(out.item())


class MinSegmentTree(SegmentTree):
    def __init__(self, capacity: int, device="cuda", _storage=None) -> None:
        super().__init__(capacity=capacity, operation=min, init_value=float("inf"), device=device,
                         _storage=_storage)

    def min(self, start: int = 0, end: int = 0) -> float:
        return super().operate(start, end)
