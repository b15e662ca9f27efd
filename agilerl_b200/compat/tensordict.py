"""Minimal stand-in for the parts of ``tensordict`` the off-policy path touches.

The reference stores transitions in a ``tensordict.TensorDict``
(agilerl/components/replay_buffer.py:4, data.py:8).  ``tensordict`` is not installed in the
build image, so this module provides a small, dependency-free ``TensorDict`` with the semantics
the path relies on: per-leaf indexing with batch-size bookkeeping, ``to``/``clone``/``items``,
``unsqueeze``/``expand`` and ``torch.zeros_like``.  When the real package is importable it is
used instead (see ``agilerl_b200.compat.__init__``).
"""
from __future__ import annotations

from collections.abc import Mapping
from typing import Any

import numpy as np
import torch


def _as_tensor(v: Any, device=None) -> Any:
    if isinstance(v, (TensorDict, torch.Tensor)):
        return v if device is None else v.to(device)
    if isinstance(v, Mapping):
        return TensorDict(v, device=device)
    if isinstance(v, np.ndarray):
        t = torch.from_numpy(v)
    else:
        t = torch.as_tensor(v)
    return t if device is None else t.to(device)


def _infer_batch(values) -> torch.Size:
    """Largest common leading shape of all leaves (tensordict's auto batch-size rule)."""
    shapes = []
    for v in values:
        shapes.append(tuple(v.batch_size) if isinstance(v, TensorDict) else tuple(v.shape))
    if not shapes:
        return torch.Size([])
    common = []
    for dims in zip(*shapes):
        if all(d == dims[0] for d in dims):
            common.append(dims[0])
        else:
            break
    return torch.Size(common)


class TensorDictBase:
    """Marker base (``tensordict.TensorDictBase``)."""


class TensorDict(TensorDictBase):
    def __init__(self, source: Mapping | None = None, batch_size=None, device=None, **kwargs):
        source = {} if source is None else dict(source)
        source.update(kwargs)
        self._data: dict[str, Any] = {k: _as_tensor(v, device) for k, v in source.items()}
        if batch_size is None:
            # the real TensorDict defaults to batch_size=[] when none is given
            self._batch_size = torch.Size([])
        else:
            self._batch_size = torch.Size(
                [batch_size] if isinstance(batch_size, int) else list(batch_size)
            )
        self.device = device

    # ---- dict protocol --------------------------------------------------------------------
    def keys(self):
        return self._data.keys()

    def values(self):
        return self._data.values()

    def items(self):
        return self._data.items()

    def __contains__(self, key) -> bool:
        return key in self._data

    def __iter__(self):
        return iter(self._data)

    def __len__(self) -> int:
        return self._batch_size[0] if len(self._batch_size) else 0

    def get(self, key, default=None):
        return self._data.get(key, default)

    def pop(self, key, *default):
        return self._data.pop(key, *default)

    def to_dict(self) -> dict:
        return {k: (v.to_dict() if isinstance(v, TensorDict) else v) for k, v in self.items()}

    # ---- batch bookkeeping ----------------------------------------------------------------
    @property
    def batch_size(self) -> torch.Size:
        return self._batch_size

    @batch_size.setter
    def batch_size(self, value) -> None:
        self._batch_size = torch.Size([value] if isinstance(value, int) else list(value))

    @property
    def shape(self) -> torch.Size:
        return self._batch_size

    def auto_batch_size_(self) -> "TensorDict":
        self._batch_size = _infer_batch(self._data.values())
        return self

    # ---- indexing -------------------------------------------------------------------------
    def __getitem__(self, index):
        if isinstance(index, str):
            return self._data[index]
        out = {k: v[index] for k, v in self._data.items()}
        probe = torch.empty(tuple(self._batch_size), device="meta")[index] if len(
            self._batch_size
        ) else None
        bs = probe.shape if probe is not None else torch.Size([])
        td = TensorDict(out, batch_size=bs)
        td.device = self.device
        return td

    def __setitem__(self, index, value) -> None:
        if isinstance(index, str):
            self._data[index] = _as_tensor(value)
            return
        if isinstance(value, (TensorDict, Mapping)):
            for k in value.keys():              # like tensordict: only the entries the source carries are written
                if k not in self._data and isinstance(index, int) and isinstance(value[k], torch.Tensor):
                    # a key the destination lacks is created (zero-filled) over the full batch shape
                    v = value[k]
                    feat = tuple(v.shape[len(self._batch_size) - 1:])
                    self._data[k] = torch.zeros((*self._batch_size, *feat), dtype=v.dtype, device=v.device)
                self._data[k][index] = value[k]
        else:
            for k in self._data:
                self._data[k][index] = value

    def set(self, key, value):
        self[key] = value
        return self

    # ---- tensor-like ops ------------------------------------------------------------------
    def _map(self, fn, batch_size=None) -> "TensorDict":
        td = TensorDict({k: fn(v) for k, v in self._data.items()},
                        batch_size=self._batch_size if batch_size is None else batch_size)
        td.device = self.device
        return td

    @property
    def dtype(self):
        """The leaves' common dtype, ``None`` when they differ (tensordict's semantics)."""
        kinds = {v.dtype for v in self._data.values() if v.dtype is not None}
        return kinds.pop() if len(kinds) == 1 else None

    def to(self, *args, **kwargs) -> "TensorDict":
        dtype = kwargs.get("dtype")
        device = kwargs.get("device")
        for a in args:
            if isinstance(a, torch.dtype):
                dtype = a
            elif a is not None:
                device = a
        def conv(v):
            if isinstance(v, TensorDict):
                return v.to(*args, **kwargs)
            if dtype is not None:
                v = v.to(dtype=dtype)
            if device is not None:
                v = v.to(device=device)
            return v
        td = self._map(conv)
        if device is not None:
            td.device = device
        return td

    def clone(self, recurse: bool = True) -> "TensorDict":
        return self._map(lambda v: v.clone())

    def detach(self) -> "TensorDict":
        return self._map(lambda v: v.detach())

    def cpu(self) -> "TensorDict":
        return self.to("cpu")

    def unsqueeze(self, dim: int) -> "TensorDict":
        assert dim == 0, "stand-in TensorDict only supports unsqueeze(0)"
        return self._map(lambda v: v.unsqueeze(0), batch_size=(1, *self._batch_size))

    def expand(self, *shape) -> "TensorDict":
        if len(shape) == 1 and not isinstance(shape[0], int):
            shape = tuple(shape[0])
        nb = len(self._batch_size)
        def ex(v):
            feat = tuple(v.shape[nb:]) if not isinstance(v, TensorDict) else ()
            if isinstance(v, TensorDict):
                return v.expand(*shape)
            return v.expand(*shape, *feat)
        return self._map(ex, batch_size=shape)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func is torch.zeros_like:
            td = args[0]
            return td._map(lambda v: torch.zeros(v.shape, dtype=v.dtype, device=v.device)
                           if not isinstance(v, TensorDict) else torch.zeros_like(v))
        if func is torch.stack:
            tds, dim = args[0], (args[1] if len(args) > 1 else kwargs.get("dim", 0))
            assert dim == 0
            return TensorDict({k: torch.stack([t[k] for t in tds]) for k in tds[0].keys()},
                              batch_size=(len(tds), *tds[0].batch_size))
        if func is torch.cat:
            tds, dim = args[0], (args[1] if len(args) > 1 else kwargs.get("dim", 0))
            assert dim == 0
            n = sum(t.batch_size[0] for t in tds)
            return TensorDict({k: torch.cat([t[k] for t in tds]) for k in tds[0].keys()},
                              batch_size=(n, *tds[0].batch_size[1:]))
        return NotImplemented

    def __repr__(self) -> str:
        fields = ", ".join(
            f"{k}: {tuple(v.shape)} {getattr(v, 'dtype', '')}" for k, v in self._data.items()
        )
        return f"TensorDict({{{fields}}}, batch_size={list(self._batch_size)})"


def is_tensor_collection(x: Any) -> bool:
    return isinstance(x, TensorDictBase)


def tensorclass(cls):
    """Tiny ``@tensorclass``: a dataclass-like holder with ``to_tensordict()`` and batch size.

    Mirrors what ``Transition`` needs (agilerl/components/data.py:68-93): keyword construction,
    ``__post_init__`` conversion hook, ``batch_size=[n]`` and ``to_tensordict()``.
    """
    fields = list(getattr(cls, "__annotations__", {}).keys())

    def __init__(self, *args, batch_size=None, device=None, **kwargs):
        for name, val in zip(fields, args):
            kwargs[name] = val
        for name in fields:
            setattr(self, name, kwargs.get(name))
        self.batch_size = torch.Size([] if batch_size is None else list(batch_size))
        self.device = device
        if hasattr(self, "__post_init__"):
            self.__post_init__()

    def to_tensordict(self) -> TensorDict:
        td = TensorDict({name: getattr(self, name) for name in fields},
                        batch_size=self.batch_size)
        return td

    cls.__init__ = __init__
    cls.to_tensordict = to_tensordict
    return cls
