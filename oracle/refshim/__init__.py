"""TEST INFRASTRUCTURE ONLY — import the *unmodified* reference package from /root/reference.

The reference (``agilerl`` 2.6.1) is not installed in the build image and several of its
third-party dependencies are absent (tensordict, gymnasium, accelerate, pettingzoo, fastrand,
h5py, minari, ...).  ``install()`` makes ``import agilerl.…`` work *from the read-only source tree*
by

  * registering an empty ``agilerl`` package whose ``__path__`` is ``/root/reference/agilerl``
    (the real ``agilerl/__init__.py`` calls ``importlib.metadata.metadata("agilerl")`` and fails
    when the distribution is not installed, agilerl/__init__.py:8,19);
  * providing *functional* stand-ins for the two dependencies the off-policy path actually
    executes — ``gymnasium.spaces`` and ``tensordict`` (the stand-ins of ``agilerl_b200.compat``);
  * auto-stubbing every other missing third-party import with inert placeholder classes
    (they are only needed so that module-level ``import``/type-alias lines succeed).

Used by ``tests/golden/make_golden.py`` to generate golden vectors from the real reference code
and by the oracle-vs-reference tests (skipped when /root/reference does not exist, e.g. on the
GPU box).  Nothing in the product package imports this module.
"""
from __future__ import annotations

import abc
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("B2RL_REFERENCE_ROOT", "/root/reference")

_OPTIONAL = [
    "gymnasium", "tensordict", "accelerate", "pettingzoo", "fastrand", "h5py", "minari",
    "flatten_dict", "matplotlib", "deepspeed", "vllm", "peft", "liger_kernel", "datasets",
    "hydra", "omegaconf", "supersuit", "jax", "redis", "pygame", "seaborn", "termcolor",
    "ucimlrepo", "gem",
]


class _Meta(abc.ABCMeta):
    def __getattr__(cls, n):
        if n.startswith("__"):
            raise AttributeError(n)
        c = _Meta(n, (_Dummy,), {})
        setattr(cls, n, c)
        return c

    def __or__(cls, o):
        return cls

    def __ror__(cls, o):
        return cls

    def __getitem__(cls, k):
        return cls

    def __iter__(cls):
        return iter(())

    def __instancecheck__(cls, inst):
        return False


class _Dummy(metaclass=_Meta):
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return _Dummy()

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Dummy()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = _Meta(name, (_Dummy,), {})
        setattr(self, name, cls)
        return cls


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, missing):
        self.missing = set(missing)

    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in self.missing and fullname not in sys.modules:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "agilerl"))


def install() -> None:
    """Idempotently make ``import agilerl.<sub>`` resolve to the read-only reference tree."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    missing = []
    for name in _OPTIONAL:
        try:
            importlib.import_module(name)
        except Exception:  # noqa: BLE001
            missing.append(name)

    repo_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if repo_root not in sys.path:
        sys.path.insert(0, repo_root)

    if "tensordict" in missing:
        from agilerl_b200.compat import tensordict as td_mod

        td = _StubModule("tensordict")
        td.__path__ = []
        for k in ("TensorDict", "TensorDictBase", "is_tensor_collection", "tensorclass"):
            setattr(td, k, getattr(td_mod, k))
        sys.modules["tensordict"] = td
    if "gymnasium" in missing:
        from agilerl_b200.compat import spaces as sp_mod

        gym = _StubModule("gymnasium")
        gym.__path__ = []
        sp = _StubModule("gymnasium.spaces")
        sp.__path__ = []
        for k in ("Space", "Box", "Discrete", "MultiDiscrete", "MultiBinary", "Dict", "Tuple", "flatdim"):
            setattr(sp, k, getattr(sp_mod, k))
        gym.spaces = sp
        sys.modules["gymnasium"] = gym
        sys.modules["gymnasium.spaces"] = sp

    sys.meta_path.insert(0, _Finder(missing))

    pkg = types.ModuleType("agilerl")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "agilerl")]
    pkg.HAS_LLM_DEPENDENCIES = False
    pkg.HAS_LIGER_KERNEL = False
    pkg.HAS_DEEPSPEED = False
    pkg.HAS_VLLM = False
    sys.modules["agilerl"] = pkg
    _installed = True
