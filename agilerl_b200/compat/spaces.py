"""Minimal ``gymnasium.spaces`` stand-ins (Box / Discrete / MultiDiscrete / MultiBinary / Dict / Tuple).

The reference describes observations and actions with gymnasium spaces
(agilerl/algorithms/dqn_rainbow.py:79-80, agilerl/typing.py:12).  ``gymnasium`` is absent from the
build image; these classes carry exactly the attributes the off-policy path reads
(``shape``, ``dtype``, ``low``, ``high``, ``n``) plus ``flatdim``.  When the real package is
importable it is used instead (see ``agilerl_b200.compat.__init__``).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None, seed=None):
        self._shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self._rng = np.random.default_rng(seed)

    @property
    def shape(self):
        return self._shape

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def contains(self, x) -> bool:  # pragma: no cover - trivial
        return True

    def __contains__(self, x) -> bool:
        return self.contains(x)


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        super().__init__(shape, dtype, seed)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self._shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self._shape).copy()

    def sample(self):
        if np.issubdtype(self.dtype, np.integer):
            return self._rng.integers(self.low, self.high, endpoint=True, size=self._shape).astype(self.dtype)
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi, size=self._shape).astype(self.dtype)

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self._shape}, {self.dtype})"

    def __eq__(self, other):
        return (isinstance(other, Box) and self._shape == other._shape and self.dtype == other.dtype
                and np.array_equal(self.low, other.low) and np.array_equal(self.high, other.high))

    __hash__ = None


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        super().__init__((), np.int64, seed)
        self.n = int(n)
        self.start = int(start)

    def sample(self):
        return int(self.start + self._rng.integers(self.n))

    def __repr__(self):
        return f"Discrete({self.n})"

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n and self.start == other.start

    __hash__ = None


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64, seed=None):
        self.nvec = np.asarray(nvec, dtype=dtype)
        super().__init__(self.nvec.shape, dtype, seed)

    def sample(self):
        return (self._rng.random(self.nvec.shape) * self.nvec).astype(self.dtype)


class MultiBinary(Space):
    def __init__(self, n, seed=None):
        self.n = n
        shape = (n,) if np.isscalar(n) else tuple(n)
        super().__init__(shape, np.int8, seed)

    def sample(self):
        return self._rng.integers(0, 2, size=self._shape).astype(self.dtype)


class Dict(Space):
    def __init__(self, spaces=None, seed=None, **kwargs):
        super().__init__(None, None, seed)
        self.spaces = OrderedDict(spaces or {})
        self.spaces.update(kwargs)

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()

    def values(self):
        return self.spaces.values()

    def items(self):
        return self.spaces.items()

    def __eq__(self, other):                 # gymnasium: equal when the sub-spaces are
        return isinstance(other, Dict) and list(self.spaces.items()) == list(other.spaces.items())

    __hash__ = None

    def get(self, k, default=None):          # gymnasium's Dict is a Mapping
        return self.spaces.get(k, default)

    def __contains__(self, k):
        return k in self.spaces

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)


class Tuple(Space):
    def __init__(self, spaces, seed=None):
        super().__init__(None, None, seed)
        self.spaces = tuple(spaces)

    def __eq__(self, other):
        return isinstance(other, Tuple) and self.spaces == other.spaces

    __hash__ = None

    def __getitem__(self, i):
        return self.spaces[i]

    def __len__(self):
        return len(self.spaces)

    def __iter__(self):
        return iter(self.spaces)


def flatdim(space) -> int:
    if isinstance(space, Box):
        return int(np.prod(space.shape))
    if isinstance(space, Discrete):
        return int(space.n)
    if isinstance(space, MultiDiscrete):
        return int(np.sum(space.nvec))
    if isinstance(space, MultiBinary):
        return int(np.prod(space.shape))
    if isinstance(space, Dict):
        return sum(flatdim(s) for s in space.spaces.values())
    if isinstance(space, Tuple):
        return sum(flatdim(s) for s in space.spaces)
    raise NotImplementedError(type(space))
