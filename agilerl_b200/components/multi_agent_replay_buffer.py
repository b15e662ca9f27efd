"""``MultiAgentReplayBuffer`` resident in HBM — drop-in for agilerl/components/multi_agent_replay_buffer.py:16-242
("deque-of-dicts MA buffer -> SoA HBM buffer").

Same constructor (``memory_size, field_names, agent_ids, device``), attributes (``memory_size, field_names, agent_ids,
counter, device``), ``__len__``, ``save_to_memory(*args, is_vectorised=False)`` / ``save_to_memory_single_env`` /
``save_to_memory_vect_envs`` and ``sample(batch_size) -> tuple of {agent_id: float32 tensor [B, ...]}`` in field order.

Layout: ONE float32 ring per FIELD, ``[memory_size, sum over agents of the leaf width]`` with the agents' leaves side by
side in ``agent_ids`` order — the operand MADDPG's critics consume (``torch.cat(..., dim=1)``, maddpg.py:611-612), so a
sampled batch is one ``b2rl_gather_rows_multi`` launch over all fields and the per-agent tensors handed back are column
views of it (``.packed`` on each returned dict is the whole ``[B, sum]`` matrix; ``MADDPG.learn`` uses it directly).
A save is one packed pinned block, one H2D copy and one ``b2rl_ring_write_multi`` launch.

Semantics kept from the reference: the buffer is a ``deque(maxlen=memory_size)`` (the oldest step falls out);
``sample`` draws ``random.sample`` positions from Python's GLOBAL ``random`` stream (``random.sample(range(n), k)``
consumes it exactly like ``random.sample(deque, k)`` and picks the same positions); every leaf comes back ``float32``
(``obs_to_tensor(...).float()``, utils/algo_utils.py:743-770); scalar leaves come back ``[B, 1]`` (:82-84).
Binary fields (``done / termination / terminated / truncation / truncated``) pass through ``uint8`` at SAVE time
(NaN entries — an agent that was not alive — stay NaN); the reference casts at sample time and only when the sampled
batch of that agent holds no NaN, which differs only for flags outside {0, 1, NaN}.
"""
from __future__ import annotations

import ctypes
import random
from collections import namedtuple
from typing import Any

import numpy as np
import torch

from .. import _lib
from .replay_buffer import _PinnedRing

BINARY_FIELDS = ("done", "termination", "terminated", "truncation", "truncated")


class PackedField(dict):
    """{agent_id: tensor [B, *leaf shape]} whose values are column views of ``packed`` ([B, sum of leaf widths])."""
    packed: torch.Tensor | None = None


class _MemoryView:
    """Read-only stand-in for the reference's ``self.memory`` deque (``maxlen``, ``len``, ``memory[i]`` -> ``Experience`` of
    ``{agent_id: np.ndarray}`` per field in the saved leaf shape and dtype).  Introspection only: every access copies one
    ring row per field from the device; the hot path never touches it."""

    def __init__(self, buf):
        self._buf = buf
        self.maxlen = buf.memory_size

    def __len__(self) -> int:
        return self._buf._size

    def __getitem__(self, i: int):
        b = self._buf
        n = b._size
        if not -n <= i < n:
            raise IndexError("deque index out of range")
        head = b._cursor if n == b.memory_size else 0
        slot = (head + (i % n)) % b.memory_size
        fields = []
        for fi in range(len(b.field_names)):
            row = b._rings[fi][slot].cpu().numpy()
            fields.append({aid: row[c0:c0 + w].reshape(b._leaf_shapes[fi][aid]).astype(b._dtypes[fi][aid])
                           for aid, (c0, w) in b._offsets[fi].items()})
        return b.experience(*fields)

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class MultiAgentReplayBuffer:
    def __init__(self, memory_size: int, field_names: list[str], agent_ids: list[str], device: str | None = None) -> None:
        assert memory_size > 0, "Memory size must be greater than zero."
        assert len(field_names) > 0, "Field names must contain at least one field name."
        assert len(agent_ids) > 0, "Agent ids must contain at least one agent id."
        if len(field_names) > 8:
            raise NotImplementedError("at most 8 fields per experience (one launch moves every field)")
        self.memory_size = memory_size
        self.field_names = field_names
        self.agent_ids = agent_ids
        self.counter = 0
        self.device = device
        self.experience = namedtuple("Experience", self.field_names)
        self._dev = _lib.as_device(device if device is not None else "cuda")     # raises: no CPU storage path
        self._lib = _lib.load()
        self._cursor = 0
        self._size = 0
        self._rings: list[torch.Tensor] | None = None
        self._shapes: list[dict] = []          # per field: {agent_id: leaf shape}
        self._offsets: list[dict] = []         # per field: {agent_id: (col0, width)}
        self._widths: list[int] = []
        self._stage = _PinnedRing()
        self._idx_stage = _PinnedRing()
        self._out_cache: dict = {}
        self._idx_cache: dict = {}
        self._uniform_offset = 0

    def __len__(self) -> int:
        return self._size

    @property
    def memory(self) -> _MemoryView:
        return _MemoryView(self)

    # -- layout ----------------------------------------------------------------------------------------
    def _init(self, args, vect: bool) -> None:
        self._shapes, self._offsets, self._widths, self._leaf_shapes, self._dtypes = [], [], [], [], []
        for arg in args:
            shapes, offs, col = {}, {}, 0
            self._leaf_shapes.append({}); self._dtypes.append({})
            for aid in self.agent_ids:
                if isinstance(arg[aid], (dict, tuple)):
                    raise NotImplementedError("dict / tuple sub-observations are not implemented in the HBM multi-agent replay")
                leaf = np.asarray(arg[aid])
                shape = tuple(leaf.shape[1:]) if vect else tuple(leaf.shape)
                w = int(np.prod(shape)) if shape else 1
                self._leaf_shapes[-1][aid], self._dtypes[-1][aid] = shape, leaf.dtype        # what ``memory[i]`` hands back
                shapes[aid], offs[aid] = (shape if shape else (1,)), (col, w)
                col += w
            self._shapes.append(shapes); self._offsets.append(offs); self._widths.append(col)
        self._rings = [torch.zeros((self.memory_size, w), dtype=torch.float32, device=self._dev) for w in self._widths]
        nf = len(self._rings)
        self._ring_ptrs = (ctypes.c_void_p * nf)(*[r.data_ptr() for r in self._rings])
        self._row_bytes = (ctypes.c_int64 * nf)(*[4 * w for w in self._widths])

    # -- ingest ----------------------------------------------------------------------------------------
    def _save(self, args, n: int, vect: bool) -> None:
        if len(args) != len(self.field_names):
            raise TypeError(f"expected {len(self.field_names)} fields ({self.field_names}), got {len(args)}")
        if self._rings is None:
            self._init(args, vect)
        if n > self.memory_size:
            raise ValueError("cannot save more steps than memory_size in one call")
        nf = len(self._rings)
        # packed staging: field f of the n steps is a dense [n, width_f] block at a 256-byte aligned offset
        offs, total = [], 0
        for w in self._widths:
            offs.append(total)
            total = (total + 4 * w * n + 255) & ~255
        stage, slot = self._stage.take(max(total, 256))
        host = stage[:max(total, 256)].numpy()
        for fi, (field, arg) in enumerate(zip(self.field_names, args)):
            w = self._widths[fi]
            block = host[offs[fi]:offs[fi] + 4 * w * n].view(np.float32).reshape(n, w)
            for aid in self.agent_ids:
                col0, wa = self._offsets[fi][aid]
                leaf = np.asarray(arg[aid])
                if field in BINARY_FIELDS:                      # multi_agent_replay_buffer.py:147-148
                    if leaf.dtype.kind == "f":
                        nan = np.isnan(leaf)
                        leaf = np.where(nan, np.nan, np.where(nan, 0, leaf).astype(np.uint8).astype(np.float32))
                    else:
                        leaf = leaf.astype(np.uint8)
                block[:, col0:col0 + wa] = leaf.reshape(n, wa)
        dev = torch.empty(max(total, 256), dtype=torch.uint8, device=self._dev)
        dev.copy_(stage[:max(total, 256)], non_blocking=True)
        self._stage.sent(slot, self._dev)
        base = dev.data_ptr()
        srcs = (ctypes.c_void_p * nf)(*[base + o for o in offs])
        _lib.check(self._lib.b2rl_ring_write_multi(nf, self._ring_ptrs, srcs, self._row_bytes, self._cursor, n, self.memory_size,
                                                   _lib.stream_ptr(self._dev)))
        self._keep = dev
        self._cursor = (self._cursor + n) % self.memory_size
        self._size = min(self._size + n, self.memory_size)
        self.counter += n

    def _add(self, *args: dict[str, Any]) -> None:
        """:103-110 — one step; does not move ``counter`` (the callers do)."""
        self._save(args, 1, vect=False)
        self.counter -= 1

    def save_to_memory_single_env(self, *args: dict[str, Any]) -> None:
        """:171-179."""
        self._save(args, 1, vect=False)

    # -- host-side helpers of the reference's class surface (not used by the HBM path) --------------------------------
    @staticmethod
    def stack_transitions(transitions: list):
        """:57-101 for array / scalar leaves: one array with a leading batch axis, 1-D results become columns."""
        if isinstance(transitions[0], (dict, tuple)):
            raise NotImplementedError("dict / tuple sub-observations are not implemented in the HBM multi-agent replay")
        ts = np.array(transitions)
        return np.expand_dims(ts, axis=1) if ts.ndim == 1 else ts

    def _process_transition(self, experiences: list, np_array: bool = False) -> dict:
        """:112-155 — ``Experience`` tuples -> ``{field: {agent_id: stacked array | float32 tensor}}`` (binary fields through
        uint8 unless they hold a NaN).  ``sample`` does not go through here: it gathers the packed rings."""
        out = {field: {} for field in self.field_names}
        rows = [e for e in experiences if e is not None]
        for field in self.field_names:
            for aid in self.agent_ids:
                ts = self.stack_transitions([getattr(e, field)[aid] for e in rows])
                if field in BINARY_FIELDS and not np.isnan(ts).any():
                    ts = ts.astype(np.uint8)
                out[field][aid] = ts if np_array else torch.as_tensor(ts, device=self._dev).float()
        return out

    def save_to_memory_vect_envs(self, *args: dict[str, Any]) -> None:
        """:213-224 — one step of every vectorised environment, in environment order."""
        n = len(next(iter(args[0].values())))
        self._save(args, n, vect=True)

    def save_to_memory(self, *args: dict[str, Any], is_vectorised: bool = False) -> None:
        """:226-242."""
        if is_vectorised:
            self.save_to_memory_vect_envs(*args)
        else:
            self.save_to_memory_single_env(*args)

    # -- sampling --------------------------------------------------------------------------------------
    def _gather(self, slots: torch.Tensor, out: list | None = None, packed_only: bool = False, stream: int | None = None) -> tuple:
        B = slots.numel()
        nf = len(self._rings)
        if out is not None:
            key = (id(out[0]), out[0].data_ptr(), B)
            cached = self._out_cache.get(key)
            if cached is None:
                dsts = list(out)
                assert len(dsts) == nf and all(d.shape == (B, w) and d.dtype == torch.float32 and d.is_contiguous() and d.device == self._dev
                                              for d, w in zip(dsts, self._widths)), "out: one contiguous float32 [B, width] per field"
                cached = self._out_cache[key] = (dsts, (ctypes.c_void_p * nf)(*[d.data_ptr() for d in dsts]))
            dsts, arr = cached
        else:
            dsts = [torch.empty((B, w), dtype=torch.float32, device=self._dev) for w in self._widths]
            arr = (ctypes.c_void_p * nf)(*[d.data_ptr() for d in dsts])
        _lib.check(self._lib.b2rl_gather_rows_multi(nf, arr, self._ring_ptrs, self._row_bytes, slots.data_ptr(), B,
                                                    _lib.stream_ptr(self._dev) if stream is None else stream))
        out_fields = []
        for fi, mat in enumerate(dsts):
            d = PackedField()
            d.packed = mat
            if not packed_only:
                for aid in self.agent_ids:
                    col0, wa = self._offsets[fi][aid]
                    d[aid] = mat[:, col0:col0 + wa].reshape(B, *self._shapes[fi][aid]) if len(self._shapes[fi][aid]) > 1 \
                        else mat[:, col0:col0 + wa]
            out_fields.append(d)
        return tuple(out_fields)

    def sample(self, batch_size: int, *args: Any) -> tuple:
        """:157-169 — ``random.sample`` positions (global ``random`` stream) -> ring slots -> one gather launch."""
        if self._rings is None:
            raise ValueError("Sample larger than population or is negative")
        pos = random.sample(range(self._size), k=batch_size)
        head = self._cursor if self._size == self.memory_size else 0        # slot of the deque's left end
        slots_h = (np.asarray(pos, dtype=np.int64) + head) % self.memory_size
        stage, slot = self._idx_stage.take(8 * batch_size)
        ctypes.memmove(stage.data_ptr(), slots_h.ctypes.data, 8 * batch_size)
        slots = torch.empty(batch_size, dtype=torch.int64, device=self._dev)
        slots.copy_(stage[:8 * batch_size].view(torch.int64), non_blocking=True)
        self._idx_stage.sent(slot, self._dev)
        return self._gather(slots)

    def sample_device(self, batch_size: int, out: list | None = None, packed_only: bool = False, stream: int | None = None) -> tuple:
        """``sample`` for the HBM-resident loop: distinct uniform positions drawn on the device (Philox,
        b2rl_sample_uniform_distinct) — no host round trip, not the reference's RNG stream.  ``out``: one ``[B, width]``
        float32 matrix per field to gather into (``MADDPG.batch_buffers``: the buffers a captured learn call reads);
        ``packed_only`` skips the per-agent column views (the returned dicts then only carry ``.packed``); ``stream``: raw
        ``cudaStream_t`` to enqueue on instead of torch's current stream (only with ``out``: nothing is allocated then)."""
        if out is not None:                     # one persistent index vector per destination (a member's own buffers)
            idx = self._idx_cache.get(id(out[0]))
            if idx is None or idx.numel() != batch_size:
                idx = self._idx_cache[id(out[0])] = torch.empty(batch_size, dtype=torch.int64, device=self._dev)
        else:
            assert stream is None, "an explicit stream needs caller-owned destination buffers (out=...)"
            idx = torch.empty(batch_size, dtype=torch.int64, device=self._dev)
        off = self._uniform_offset
        sp = _lib.stream_ptr(self._dev) if stream is None else stream
        _lib.check(self._lib.b2rl_sample_uniform_distinct(0x3A44, off, self._size, batch_size, idx.data_ptr(), sp))
        self._uniform_offset = off + 64 * batch_size
        return self._gather(idx, out, packed_only, sp)     # a uniform draw over the slots is a uniform draw over the positions
