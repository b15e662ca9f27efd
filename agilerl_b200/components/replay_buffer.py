"""HBM-resident replay buffers with the reference's class surface.

Drop-in for ``agilerl/components/replay_buffer.py``:

* ``ReplayBuffer``            (:12-138)  circular SoA storage, ``add`` / ``sample`` / ``clear``
* ``MultiStepReplayBuffer``   (:141-258) n-step return roll at ingest, ``sample_from_indices``
* ``PrioritizedReplayBuffer`` (:261-428) PER with fp64 sum/min trees

Same constructor arguments, attributes (``max_size, device, dtype, counter, initialized, _cursor,
_size, _storage, n_step_buffer, reward_key/done_key/ns_key, alpha, max_priority, tree_ptr,
sum_tree, min_tree``) and return shapes (``weights [B,1]`` f32, ``idxs [B,1]`` int64).  All data
movement and tree arithmetic run in libb2rl.so (csrc/replay.cu, csrc/tree.cu) on the buffer's CUDA
device; there is no CPU storage path.

Host work that stays on the host by design: the cursor/size bookkeeping (integers), the RNG draw
(``torch.randperm`` / ``torch.rand`` from torch's global CPU generator, so seeded runs consume the
same stream as the reference) and ``priority ** alpha`` in CPython doubles — the reference computes
exactly that on the host (replay_buffer.py:322) and glibc ``pow`` is what makes the leaves
bit-identical; the trees themselves are updated on the GPU.
"""
from __future__ import annotations

import ctypes
import warnings
from collections import deque
from typing import Any

import numpy as np
import torch

from .. import _lib
from ..compat import TensorDict, TensorDictBase, is_tensor_collection
from .segment_tree import MinSegmentTree, SumSegmentTree

DataType = Any


def _leaf_items(td, prefix=()):
    """Flatten a (possibly nested) tensor collection into ((path), tensor) pairs."""
    for k, v in td.items():
        if is_tensor_collection(v) or isinstance(v, dict):
            yield from _leaf_items(v, (*prefix, k))
        else:
            yield (*prefix, k), v


def _unflatten(leaves: dict, batch_size) -> TensorDict:
    root: dict = {}
    for path, v in leaves.items():
        d = root
        for k in path[:-1]:
            d = d.setdefault(k, {})
        d[path[-1]] = v

    def build(d):
        return TensorDict({k: (build(v) if isinstance(v, dict) else v) for k, v in d.items()},
                          batch_size=batch_size)

    return build(root)


def _shallow(td):
    """New collection(s) around the same leaf tensors."""
    if hasattr(td, "_map"):                      # compat stand-in
        return td._map(lambda v: _shallow(v) if is_tensor_collection(v) else v)
    return td.clone(False)                       # real tensordict: non-recursive clone shares the leaves


def _numel(shape) -> int:
    n = 1
    for d in shape:
        n *= int(d)
    return n


def _as_collection(data) -> Any:
    if is_tensor_collection(data):
        return data
    if isinstance(data, dict):
        return TensorDict(data)
    if hasattr(data, "to_tensordict"):
        return data.to_tensordict()
    raise TypeError(f"Cannot store data of type {type(data)} in a replay buffer")


import os as _os
_NOSYNC = _os.environ.get("B2RL_RING_NOSYNC") == "1"     # diagnostics only


class _PinnedRing:
    """A few pinned host blocks used round-robin as the source of asynchronous H2D copies.  A slot is handed
    out again only after the copy that last read it has completed (its event), so callers may overwrite their
    own host arrays as soon as ``add`` / ``update_priorities`` return — the blocking ``.to(device)`` semantics
    of the reference without blocking on the stream."""

    def __init__(self, slots: int = 8):
        self.slots = slots
        self.bufs: list = [None] * slots
        self.events: list = [None] * slots
        self.i = 0

    def take(self, nbytes: int) -> tuple[torch.Tensor, int]:
        k = self.i
        self.i = (k + 1) % self.slots
        if self.events[k] is not None and not _NOSYNC:
            self.events[k].synchronize()
        buf = self.bufs[k]
        if buf is None or buf.numel() < nbytes:
            buf = self.bufs[k] = torch.empty(max(nbytes, 4096), dtype=torch.uint8).pin_memory()
        return buf, k

    def sent(self, k: int, device) -> None:
        ev = self.events[k]
        if ev is None:
            ev = self.events[k] = torch.cuda.Event()
        ev.record()


_RANDPERM_FAST: bool | None = None          # None: not checked yet in this process


def _randperm_prefix_fast(lib, n: int, B: int) -> torch.Tensor:
    st = torch.get_rng_state()
    out = torch.empty(B, dtype=torch.int64)
    _lib.check(lib.b2rl_host_randperm_prefix(st.data_ptr(), st.numel(), n, B, out.data_ptr()))
    torch.set_rng_state(st)
    return out


def randperm_prefix(lib, n: int, batch_size: int) -> torch.Tensor:
    """``torch.randperm(n)[:batch_size]`` (replay_buffer.py:126) — the same indices AND the same state of torch's global CPU
    generator afterwards — without shuffling all n indices when only a short prefix is used: ``b2rl_host_randperm_prefix``
    replays the first ``batch_size`` iterations of torch's Fisher-Yates loop on the serialised mt19937 state and skips the
    draws of the rest (1 M transitions, B = 512: 8 ms -> 0.4 ms of host time per ``sample()``).  Checked once per process
    against ``torch.randperm`` itself (indices and generator state, on a copy of the live state); any difference — another
    torch build, another generator — switches the process back to ``torch.randperm`` for good."""
    global _RANDPERM_FAST
    B = min(int(batch_size), int(n))
    fn = getattr(lib, "b2rl_host_randperm_prefix", None)
    if fn is None or _RANDPERM_FAST is False or B < 0 or n < 4096 or 16 * B > n or n >= 0xFFFFFFFF // 20:
        return torch.randperm(n)[:batch_size]
    if _RANDPERM_FAST is None:
        live = torch.get_rng_state()
        try:
            want = torch.randperm(5000)[:37]
            want_state = torch.get_rng_state()
            torch.set_rng_state(live)
            got = _randperm_prefix_fast(lib, 5000, 37)
            _RANDPERM_FAST = bool(torch.equal(want, got) and torch.equal(want_state, torch.get_rng_state()))
        except Exception:
            _RANDPERM_FAST = False
        finally:
            torch.set_rng_state(live)
        if not _RANDPERM_FAST:
            warnings.warn("b2rl_host_randperm_prefix does not reproduce torch.randperm on this torch build: "
                          "ReplayBuffer.sample keeps torch.randperm", stacklevel=2)
            return torch.randperm(n)[:batch_size]
    return _randperm_prefix_fast(lib, n, B)


class ReplayBuffer:
    """Circular replay buffer resident in HBM (replay_buffer.py:12-138)."""

    def __init__(self, max_size: int, device: str | torch.device = "cuda",
                 dtype: torch.dtype = torch.float32) -> None:
        self.max_size = max_size
        self.device = device
        self._dev = _lib.as_device(device)      # raises: no CPU fallback
        self.dtype = dtype
        self.counter = 0
        self.initialized = False
        self._cursor = 0
        self._size = 0
        self._storage: TensorDict | None = None
        self._fields: dict[tuple, torch.Tensor] = {}
        self._row_bytes: dict[tuple, int] = {}
        self._lib = _lib.load()
        self._ring = _PinnedRing()
        self._stage_plans: dict = {}

    # -- host -> device staging -----------------------------------------------------------------
    def _to_device(self, data):
        """Every host leaf of a transition through ONE pinned staging block and ONE H2D copy (the reference does
        ``data.to(device)``: a blocking copy per leaf).  Leaves already on the device pass through."""
        leaves = list(_leaf_items(data))
        host = [(path, v) for path, v in leaves if isinstance(v, torch.Tensor) and v.device.type == "cpu"]
        if not host or len(host) != len(leaves):
            return data.to(self._dev)
        key = tuple((path, tuple(v.shape), v.dtype) for path, v in host)
        plan = self._stage_plans.get(key)
        if plan is None:
            off, plan_l = 0, []
            for path, v in host:
                nb = v.numel() * v.element_size()
                plan_l.append((path, off, nb, v.dtype, tuple(v.shape)))
                off = (off + nb + 255) & ~255
            plan = self._stage_plans[key] = (plan_l, max(off, 256))
        plan_l, total = plan
        stage, slot = self._ring.take(total)
        base = stage.data_ptr()
        for (path, off, nb, dt, shape), (_, v) in zip(plan_l, host):
            # plain memcpy: a torch CPU copy of a frame stack is large enough to wake torch's whole intra-op thread
            # pool, whose workers then spin on every core of the box for their block time (measured: 5x slower loop)
            if not v.is_contiguous():
                v = v.contiguous()
            ctypes.memmove(base + off, v.data_ptr(), nb)
        dev = torch.empty(total, dtype=torch.uint8, device=self._dev)
        dev.copy_(stage[:total], non_blocking=True)
        self._ring.sent(slot, self._dev)
        out = {path: dev[off:off + nb].view(dt).view(shape) for path, off, nb, dt, shape in plan_l}
        bs = tuple(data.batch_size) if hasattr(data, "batch_size") else ()
        return _unflatten(out, bs)

    # -- properties (replay_buffer.py:39-58) ---------------------------------------------------
    @property
    def storage(self) -> TensorDict | None:
        return self._storage

    @property
    def size(self) -> int:
        return self._size

    @size.setter
    def size(self, value: int) -> None:
        self._size = value

    @property
    def is_full(self) -> bool:
        return len(self) == self.max_size

    def __len__(self) -> int:
        return self._size

    # -- ingest ---------------------------------------------------------------------------------
    def _prepare(self, data) -> tuple[dict[tuple, torch.Tensor], int]:
        data = _as_collection(data)
        if any(isinstance(v, torch.Tensor) and v.device.type == "cpu" for _, v in _leaf_items(data)):
            data = self._to_device(data)
        leaves = {}
        n = None
        for path, v in _leaf_items(data):
            if not isinstance(v, torch.Tensor):
                v = torch.as_tensor(v)
            if v.device != self._dev:
                v = v.to(self._dev, non_blocking=True)
            if n is None:
                n = v.shape[0]
            if v.ndim == 1:                       # :85-94 scalar leaves become (n, 1)
                v = v.reshape(v.shape[0], 1)
            leaves[path] = v if v.is_contiguous() else v.contiguous()
        if n is None:
            raise ValueError("empty transition")
        return leaves, n

    def _init(self, leaves: dict[tuple, torch.Tensor]) -> None:
        """:60-70 — storage shape/dtype follow the first batch (uint8 frames stay uint8)."""
        self._fields = {
            path: torch.zeros((self.max_size, *v.shape[1:]), dtype=v.dtype, device=self._dev)
            for path, v in leaves.items()
        }
        self._storage = _unflatten(self._fields, (self.max_size,))
        self._row_bytes = {path: (t.numel() // self.max_size) * t.element_size() for path, t in self._fields.items()}
        # tables for the multi-field C entry points (b2rl_ring_write_multi / b2rl_gather_rows_multi)
        self._field_order = list(self._fields)
        nf = len(self._field_order)
        self._field_ptrs = (ctypes.c_void_p * nf)(*[self._fields[p].data_ptr() for p in self._field_order])
        self._field_row_bytes = (ctypes.c_int64 * nf)(*[self._row_bytes[p] for p in self._field_order])
        self.initialized = True

    def add(self, data: DataType) -> None:
        """:72-112 — ring write with wrap-around split (b2rl_ring_write, one call per field)."""
        leaves, n = self._prepare(data)
        self._add_leaves(leaves, n)

    def _add_leaves(self, leaves: dict, n: int) -> None:
        if self._storage is None:
            self._init(leaves)
        if n > self.max_size:
            raise ValueError("cannot add more transitions than max_size in one call")
        nf = len(leaves)
        srcs = []
        for path, v in leaves.items():
            dst = self._fields[path]
            if v.dtype != dst.dtype:
                v = v.to(dst.dtype)
            assert v.numel() * v.element_size() == self._row_bytes[path] * n, f"shape mismatch for {path}"
            srcs.append(v)
        if nf <= 8 and list(leaves) == self._field_order:       # every field of the transition in one launch
            arr = ctypes.c_void_p * nf
            _lib.check(self._lib.b2rl_ring_write_multi(nf, self._field_ptrs, arr(*[v.data_ptr() for v in srcs]),
                                                       self._field_row_bytes, self._cursor, n, self.max_size,
                                                       _lib.stream_ptr(self._dev)))
        else:
            stream = _lib.stream_ptr(self._dev)
            for (path, _), v in zip(leaves.items(), srcs):
                _lib.check(self._lib.b2rl_ring_write(self._fields[path].data_ptr(), v.data_ptr(), self._row_bytes[path],
                                                     self._cursor, n, self.max_size, stream))
        self._keep_add = srcs
        self._cursor = (self._cursor + n) % self.max_size
        self._size = min(self._size + n, self.max_size)
        self.counter += n

    # -- gather ---------------------------------------------------------------------------------
    def _gather(self, idx: torch.Tensor) -> TensorDict:
        """storage[idx] -> fresh tensors of shape idx.shape + feature shape (b2rl_gather_rows)."""
        idx_dev = idx
        if idx.device != self._dev or idx.dtype != torch.int64:
            idx_dev = idx.to(self._dev, dtype=torch.int64, non_blocking=True)
        if not idx_dev.is_contiguous():
            idx_dev = idx_dev.contiguous()
        flat = idx_dev.reshape(-1)
        nrows = flat.numel()
        stream = _lib.stream_ptr(self._dev)
        out = {}
        dsts = [torch.empty((nrows, *src.shape[1:]), dtype=src.dtype, device=self._dev) for src in self._fields.values()]
        nf = len(dsts)
        if nf <= 8:
            arr = ctypes.c_void_p * nf
            _lib.check(self._lib.b2rl_gather_rows_multi(nf, arr(*[d.data_ptr() for d in dsts]), self._field_ptrs,
                                                        self._field_row_bytes, flat.data_ptr(), nrows, stream))
        else:
            for (path, src), dst in zip(self._fields.items(), dsts):
                _lib.check(self._lib.b2rl_gather_rows(dst.data_ptr(), src.data_ptr(), flat.data_ptr(),
                                                      self._row_bytes[path], nrows, stream))
        for (path, src), dst in zip(self._fields.items(), dsts):
            out[path] = dst.reshape(*idx_dev.shape, *src.shape[1:])
        return _unflatten(out, tuple(idx_dev.shape))

    def sample(self, batch_size: int, return_idx: bool = False) -> TensorDict:
        """:114-131 — uniform WITHOUT replacement (randperm from torch's CPU generator, Q12)."""
        indices = randperm_prefix(self._lib, self.size, batch_size)
        samples = self._gather(indices)
        if return_idx:
            samples["idxs"] = indices.to(self._dev)
        return samples

    def sample_device(self, batch_size: int, return_idx: bool = False) -> TensorDict:
        """``sample`` for the HBM-resident loop: the B distinct uniform indices are drawn on device (Philox,
        b2rl_sample_uniform_distinct) instead of ``torch.randperm(size)`` on the host — no host round trip, not the
        reference's RNG stream."""
        idx = torch.empty(batch_size, dtype=torch.int64, device=self._dev)
        off = getattr(self, "_uniform_offset", 0)
        _lib.check(self._lib.b2rl_sample_uniform_distinct(getattr(self, "_uniform_seed", 0x5A11), off, self._size, batch_size,
                                                          idx.data_ptr(), _lib.stream_ptr(self._dev)))
        self._uniform_offset = off + 64 * batch_size
        samples = self._gather(idx)
        if return_idx:
            samples["idxs"] = idx
        return samples

    def clear(self) -> None:
        """:133-138."""
        self._size = 0
        self._cursor = 0
        self._storage = None
        self._fields = {}
        self._row_bytes = {}
        self.initialized = False


class MultiStepReplayBuffer(ReplayBuffer):
    """n-step returns rolled at ingest (replay_buffer.py:141-258)."""

    def __init__(self, max_size: int, n_step: int = 3, gamma: float = 0.99,
                 device: str | torch.device = "cuda", dtype: torch.dtype = torch.float32) -> None:
        super().__init__(max_size, device, dtype)
        self.n_step = n_step
        self.gamma = gamma
        self.n_step_buffer: deque = deque(maxlen=n_step)
        self._window_ptrs: deque = deque(maxlen=n_step)
        self.reward_key = "reward"
        self.done_key = None
        self.ns_key = "next_obs"

    def add(self, data: DataType):
        """:173-194 — returns the oldest transition of the window (or None while filling)."""
        data = self._to_device(_as_collection(data))
        self.n_step_buffer.append(data)
        self._window_ptrs.append(self._leaf_ptrs(data))
        if len(self.n_step_buffer) < self.n_step:
            return None
        if not self._ingest_fused():
            n_step_data = self._get_n_step_info()
            super().add(n_step_data)
        return self.n_step_buffer[0]

    # -- fused ingest: fold + select + ring write in one launch -------------------------------------
    def _leaf_ptrs(self, data):
        """(path -> (data_ptr, shape, dtype)) of a window entry, taken once when it enters the window."""
        out = {}
        for path, v in _leaf_items(data):
            if not (isinstance(v, torch.Tensor) and v.device == self._dev and v.is_contiguous()):
                return None
            out[path] = (v.data_ptr(), tuple(v.shape), v.dtype)
        return out

    def _ingest_fused(self) -> bool:
        """One ``b2rl_nstep_ingest`` launch instead of fold + two select-copies + a multi-field ring write with their
        temporaries.  Needs an initialised storage whose fields the window entries match leaf for leaf (flat float32
        reward / done of one element per env); anything else takes the general path."""
        if self._storage is None or self.n_step < 2 or self.done_key is None:
            return False
        win = list(self._window_ptrs)
        order = self._field_order
        nf, n = len(order), self.n_step
        if nf > 8 or any(w is None or list(w) != order for w in win):
            return False
        E = win[0][order[0]][1][0]
        rk, dk, nk = (self.reward_key,), (self.done_key,), (self.ns_key,)
        if rk not in win[0] or dk not in win[0] or nk not in win[0] or E > self.max_size:
            return False
        for w in win:
            for path in order:
                ptr, shape, dt = w[path]
                dst = self._fields[path]
                if dt != dst.dtype or shape[0] != E or _numel(shape[1:]) != _numel(dst.shape[1:]):
                    return False
            if w[rk][2] != torch.float32 or w[dk][2] != torch.float32 or _numel(w[rk][1]) != E or _numel(w[dk][1]) != E:
                return False
        src = (ctypes.c_void_p * (nf * n))(*[win[k][path][0] for path in order for k in range(n)])
        role = (ctypes.c_int32 * nf)(*[2 if path == rk else (1 if path in (dk, nk) else 0) for path in order])
        rew = (ctypes.c_void_p * n)(*[w[rk][0] for w in win])
        don = (ctypes.c_void_p * n)(*[w[dk][0] for w in win])
        _lib.check(self._lib.b2rl_nstep_ingest(nf, self._field_ptrs, src, self._field_row_bytes, role, rew, don, n, E,
                                               float(self.gamma), self._cursor, self.max_size, _lib.stream_ptr(self._dev)))
        self._cursor = (self._cursor + E) % self.max_size
        self._size = min(self._size + E, self.max_size)
        self.counter += E
        return True

    def sample_from_indices(self, idxs: torch.Tensor) -> TensorDict:
        """:196-204 — ``storage[idxs]`` (keeps the idxs shape, e.g. [B,1] -> fields [B,1,...])."""
        if not isinstance(idxs, torch.Tensor):
            idxs = torch.as_tensor(np.asarray(idxs))
        return self._gather(idxs)

    def _get_n_step_info(self) -> TensorDict:
        """:206-258 on device: b2rl_nstep_fold (reward) + b2rl_select_copy (next_obs, done)."""
        window = list(self.n_step_buffer)
        first = window[0]
        if not self.initialized:
            assert self.reward_key in first, (
                f"Reward key not found in transition. Expected key: {self.reward_key}")
            assert self.ns_key in first, (
                f"Next observation key not found in transition. Expected key: {self.ns_key}")
            done_key = None
            expected_keys = ["done", "termination", "terminated"]
            for key in expected_keys:
                if key in first:
                    done_key = key
                    break
            assert done_key is not None, (
                f"No done/termination key found in transition. Expected keys: {expected_keys}")
            self.done_key = done_key

        n = len(window)
        if n == 1:
            return first.clone()
        # entries of `out` are replaced below, never written in place: share the untouched tensors
        out = _shallow(first)
        stream = _lib.stream_ptr(self._dev)
        rewards = [w[self.reward_key].to(torch.float32).contiguous() for w in window]
        dones = [w[self.done_key].to(torch.float32).contiguous() for w in window]
        num_envs = rewards[0].numel()
        reward_out = torch.empty_like(rewards[0])
        last = torch.empty(1, dtype=torch.int32, device=self._dev)
        arr_t = ctypes.c_void_p * n
        _lib.check(self._lib.b2rl_nstep_fold(arr_t(*[r.data_ptr() for r in rewards]),
                                             arr_t(*[d.data_ptr() for d in dones]), n, num_envs,
                                             float(self.gamma), reward_out.data_ptr(), last.data_ptr(), stream))
        # carry next_obs / done of the step the fold stopped at; `last` stays on device
        for key in (self.ns_key, self.done_key):
            leaves0 = dict(_leaf_items({key: first[key]}))
            for path in leaves0:
                srcs = []
                for w in window:
                    v = w[path[0]]
                    for k in path[1:]:
                        v = v[k]
                    srcs.append(v.contiguous())
                dst = torch.empty_like(srcs[0])
                # window[0] can only be selected when n_step == 1; slot 0 is a valid dummy
                _lib.check(self._lib.b2rl_select_copy(dst.data_ptr(), arr_t(*[s.data_ptr() for s in srcs]), n,
                                                      last.data_ptr(), dst.numel() * dst.element_size(), stream))
                if len(path) == 1:
                    out[key] = dst
                else:
                    tgt = out[path[0]]
                    for k in path[1:-1]:
                        tgt = tgt[k]
                    tgt[path[-1]] = dst
                self._keepalive = (srcs, rewards, dones)
        out[self.reward_key] = reward_out.reshape(first[self.reward_key].shape).to(first[self.reward_key].dtype)
        return out


class PrioritizedReplayBuffer(ReplayBuffer):
    """Prioritized replay with fp64 sum/min trees in HBM (replay_buffer.py:261-428)."""

    def __init__(self, max_size: int, alpha: float = 0.6, device: str | torch.device = "cuda",
                 dtype: torch.dtype = torch.float32) -> None:
        super().__init__(max_size, device, dtype)
        self.alpha = alpha
        self._max_priority = 1.0
        self._max_priority_dev = torch.ones(1, dtype=torch.float64, device=self._dev)
        self._dev_dirty = False
        self.tree_ptr = 0
        tree_capacity = 1
        while tree_capacity < max_size:
            tree_capacity *= 2
        self.sum_tree = SumSegmentTree(tree_capacity, device=self._dev)
        self.min_tree = MinSegmentTree(tree_capacity, device=self._dev)
        self._cap = tree_capacity
        # production sampling may draw uniforms on device (Philox) instead of torch's CPU stream
        self.device_rng = False
        self._philox_seed = 0x5EED
        self._philox_offset = 0

    @property
    def max_priority(self) -> float:
        """Running max of raw priorities (replay_buffer.py:329).  The fused device path folds its
        maxima into a device scalar; reading the attribute reconciles the two (one 8-byte D2H)."""
        if self._dev_dirty:
            self._max_priority = max(self._max_priority, float(self._max_priority_dev.item()))
            self._dev_dirty = False
        return self._max_priority

    @max_priority.setter
    def max_priority(self, value: float) -> None:
        # an assigned value is THE maximum: the device scalar stops being authoritative (it is re-seeded from the host
        # value before the next device-side fold, see _own_max_on_device) and cannot resurrect an older maximum
        self._max_priority = float(value)
        self._dev_dirty = False

    def _own_max_on_device(self) -> None:
        """Call before enqueuing a device-side fold into ``_max_priority_dev`` (update_priorities_device, the captured
        step): seeds the device scalar with the host value unless the device copy is already the authoritative one."""
        if not self._dev_dirty:
            self._max_priority_dev.fill_(self._max_priority)
            self._dev_dirty = True

    def add(self, data: DataType) -> None:
        """:296-309 — ring write, then the n new leaves get max_priority**alpha.  While the device path owns the running
        maximum (after update_priorities_device) the leaf value is formed on the device from max(host, device) — no host
        read of the device scalar per env step; the Python attribute reconciles lazily when it is read."""
        leaves, n = self._prepare(data)
        self._add_leaves(leaves, n)
        if self._dev_dirty:
            _lib.check(self._lib.b2rl_tree_set_range_devmax(self.sum_tree.data_ptr, self.min_tree.data_ptr, self._cap,
                                                            self.tree_ptr, n, self.max_size, float(self._max_priority),
                                                            self._max_priority_dev.data_ptr(), float(self.alpha),
                                                            _lib.stream_ptr(self._dev)))
        else:
            p_alpha = float(self._max_priority) ** self.alpha
            _lib.check(self._lib.b2rl_tree_set_range(self.sum_tree.data_ptr, self.min_tree.data_ptr, self._cap,
                                                     self.tree_ptr, n, self.max_size, p_alpha,
                                                     _lib.stream_ptr(self._dev)))
        self.tree_ptr = (self.tree_ptr + n) % self.max_size

    def _update_priority(self, idx: int, priority: float) -> None:
        """:311-329."""
        assert 0 <= idx < self.max_size
        priority_alpha = priority ** self.alpha
        self.sum_tree[idx] = priority_alpha
        self.min_tree[idx] = priority_alpha
        self.max_priority = max(self.max_priority, priority)

    def _uniforms(self, batch_size: int) -> torch.Tensor:
        # B x torch.rand(1).item() (replay_buffer.py:377) == torch.rand(B) on the CPU generator; drawn straight
        # into a pinned block so the H2D copy is asynchronous
        stage, slot = self._ring.take(batch_size * 4)
        host = stage[:batch_size * 4].view(torch.float32)
        u = torch.rand(batch_size)
        ctypes.memmove(stage.data_ptr(), u.data_ptr(), batch_size * 4)
        dev = torch.empty(batch_size, dtype=torch.float32, device=self._dev)
        dev.copy_(host, non_blocking=True)
        self._ring.sent(slot, self._dev)
        return dev

    def _sample(self, batch_size: int, beta: float | None):
        idx = torch.empty(batch_size, dtype=torch.int64, device=self._dev)
        w = torch.empty(batch_size, dtype=torch.float32, device=self._dev) if beta is not None else None
        stream = _lib.stream_ptr(self._dev)
        if self.device_rng:
            _lib.check(self._lib.b2rl_per_sample_philox(
                self.sum_tree.data_ptr, self.min_tree.data_ptr, self._cap, self._philox_seed, self._philox_offset,
                batch_size, 0.0 if beta is None else float(beta), self._size, idx.data_ptr(),
                None if w is None else w.data_ptr(), stream))
            self._philox_offset += batch_size
        else:
            u = self._uniforms(batch_size)
            _lib.check(self._lib.b2rl_per_sample(
                self.sum_tree.data_ptr, self.min_tree.data_ptr, self._cap, u.data_ptr(), batch_size,
                0.0 if beta is None else float(beta), self._size, idx.data_ptr(),
                None if w is None else w.data_ptr(), stream))
        return idx, w

    def _sample_proportional(self, batch_size: int) -> torch.Tensor:
        """:357-381 — stratified proportional sampling (indices int64 [B])."""
        return self._sample(batch_size, None)[0]

    def _calculate_weights(self, indices: torch.Tensor, beta: float) -> torch.Tensor:
        """:383-409 for arbitrary indices (off the fused path; fp64 on device, cast to f32)."""
        idx = indices.to(self._dev, dtype=torch.int64).reshape(-1)
        st, mt = self.sum_tree._t, self.min_tree._t
        total = st[1]
        p_min = mt[1] / total
        max_w = (p_min * self._size) ** -beta
        p = st[self._cap + idx] / total
        return (((p * self._size) ** -beta) / max_w).to(torch.float32)

    def sample(self, batch_size: int, beta: float = 0.4) -> TensorDict:
        """:331-355 — one kernel for tree descent + IS weights, one gather per field."""
        indices, weights = self._sample(batch_size, beta)
        samples = self._gather(indices)
        samples["weights"] = weights.unsqueeze(1)
        samples["idxs"] = indices.unsqueeze(1)
        return samples

    def update_priorities(self, indices, priorities) -> None:
        """:411-428 — floor at 1e-5, leaf = p**alpha in host doubles with the C library's ``pow`` (what CPython's
        ``**`` evaluates: bit-identical leaves, tests/test_host_pow_cpu.py), one asynchronous H2D from pinned
        memory, batched last-writer-wins tree update on device (b2rl_tree_set)."""
        if isinstance(priorities, torch.Tensor):
            pri = priorities.detach().reshape(-1).to(device="cpu", dtype=torch.float32).numpy()
        else:
            pri = np.ascontiguousarray(np.asarray(priorities).reshape(-1), dtype=np.float32)
        if isinstance(indices, torch.Tensor):
            idx_dev = indices.detach().reshape(-1)
            if idx_dev.device != self._dev or idx_dev.dtype != torch.int64:
                idx_dev = idx_dev.to(self._dev, dtype=torch.int64)
        else:
            idx_dev = torch.as_tensor(np.asarray(indices).reshape(-1), dtype=torch.int64).to(self._dev)
        n = min(idx_dev.numel(), pri.size)
        if n == 0:
            return
        stage, slot = self._ring.take(n * 8)
        mx = ctypes.c_double(self.max_priority)
        _lib.check(self._lib.b2rl_host_priority_pow(pri.ctypes.data, n, float(self.alpha), 1e-5, stage.data_ptr(),
                                                    ctypes.byref(mx)))
        self.max_priority = mx.value
        pa_dev = torch.empty(n, dtype=torch.float64, device=self._dev)
        pa_dev.copy_(stage[:n * 8].view(torch.float64), non_blocking=True)
        self._ring.sent(slot, self._dev)
        _lib.check(self._lib.b2rl_tree_set(self.sum_tree.data_ptr, self.min_tree.data_ptr, self._cap,
                                           idx_dev.data_ptr(), pa_dev.data_ptr(), n, _lib.stream_ptr(self._dev)))
        self._keep = (idx_dev, pa_dev)

    # -- fused HBM-resident path (no host round trips) -------------------------------------------
    def sample_fused(self, batch_size: int, beta: float, n_step_memory: "MultiStepReplayBuffer",
                     uniforms: torch.Tensor | None = None):
        """One kernel: tree descent + IS weights + gather of the sampled slots' n-step
        action/reward/done.  Returns device tensors (idx int64[B], weights, action, reward, done
        f32[B]); frames stay in the ring and are read through ``idx`` by the encoder."""
        f = n_step_memory._fields
        dk = n_step_memory.done_key or "done"
        ring = [f[("action",)], f[(n_step_memory.reward_key,)], f[(dk,)]]
        for t in ring:
            assert t.dtype == torch.float32 and t[0].numel() == 1, "fused path needs scalar f32 action/reward/done"
        B = batch_size
        idx = torch.empty(B, dtype=torch.int64, device=self._dev)
        out = [torch.empty(B, dtype=torch.float32, device=self._dev) for _ in range(4)]
        if uniforms is None and not self.device_rng:
            uniforms = self._uniforms(B)
        _lib.check(self._lib.b2rl_per_sample_fused(
            self.sum_tree.data_ptr, self.min_tree.data_ptr, self._cap,
            None if uniforms is None else uniforms.data_ptr(), self._philox_seed, self._philox_offset, B, float(beta),
            self._size, ring[0].data_ptr(), ring[1].data_ptr(), ring[2].data_ptr(), idx.data_ptr(),
            out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(), _lib.stream_ptr(self._dev)))
        if uniforms is None:
            self._philox_offset += B
        self._keep_u = uniforms
        return idx, out[0], out[1], out[2], out[3]

    def update_priorities_device(self, indices: torch.Tensor, priorities: torch.Tensor) -> None:
        """Device-only variant of ``update_priorities``: floor, ``p**alpha`` (CUDA fp64 pow, <= 1 ulp
        from glibc — see DESIGN.md), last-writer-wins tree update and max-priority fold, no sync."""
        idx = indices.reshape(-1)
        pri = priorities.reshape(-1)
        assert idx.is_cuda and pri.is_cuda and idx.dtype == torch.int64 and pri.dtype == torch.float32
        self._own_max_on_device()
        _lib.check(self._lib.b2rl_tree_set_from_priorities(
            self.sum_tree.data_ptr, self.min_tree.data_ptr, self._cap, idx.data_ptr(), pri.data_ptr(), idx.numel(),
            float(self.alpha), 1e-5, self._max_priority_dev.data_ptr(), _lib.stream_ptr(self._dev)))

