"""Evolvable-module contract (host side) — mirror of agilerl/modules/base.py.

In the reference an ``EvolvableModule`` is an ``nn.Module`` that can mutate its own architecture
and rebuild itself while preserving weights (``@mutation`` registry :27-54, ``MutationContext``
:57-158, ``preserve_parameters`` :471-502, ``clone`` :713-736).  Here a module is an
*architecture description*: the weights live in its owning network's flat HBM buffers
(``networks/base.py``), so a mutation edits the description and the owning network re-lays the
buffers out and copies the overlapping slices (same ``preserve_parameters`` rule).  The public
surface that ``Mutations`` / ``clone`` / checkpoints use is kept: ``mutation_methods``,
``layer_mutation_methods``, ``node_mutation_methods``, ``sample_mutation_method``,
``last_mutation_attr``, ``init_dict``, ``net_config``, ``activation``, ``change_activation``,
``clone``, ``rng``, ``disable_mutations``.
"""
from __future__ import annotations

import copy
import inspect
from enum import Enum
from functools import wraps
from typing import Any, Callable

import numpy as np


class MutationType(Enum):
    LAYER = "layer"
    NODE = "node"
    ACTIVATION = "activation"


def mutation(mutation_type: MutationType, **recreate_kwargs) -> Callable:
    """Mark a method as an architecture mutation (modules/base.py:27-54)."""

    def decorator(func):
        @wraps(func)
        def wrapper(self, *args, **kwargs):
            self._mutation_depth += 1
            self.last_mutation_attr = func.__name__
            try:
                out = func(self, *args, **kwargs)
            finally:
                self._mutation_depth -= 1
            if self._mutation_depth == 0:
                # the outermost mutation finished: let the owner rebuild (MutationContext.__exit__)
                self._recreate_kwargs = dict(recreate_kwargs)
                for hook in list(self._mutation_hooks):
                    hook()
            return out

        wrapper._mutation_type = mutation_type
        wrapper._recreate_kwargs = recreate_kwargs
        return wrapper

    return decorator


class EvolvableModule:
    def __init__(self, device: str = "cuda", random_seed: int | None = None) -> None:
        self.device = device
        self.random_seed = random_seed
        self.rng = np.random.default_rng(random_seed)
        self._mutation_depth = 0
        self._mutation_hooks: list[Callable] = []
        self._disabled: set[MutationType] = set()
        self._disabled_names: set[str] = set()
        self.last_mutation_attr: str | None = None
        self.last_mutation = None
        self._recreate_kwargs: dict = {}

    # -- registry ------------------------------------------------------------------------------
    def _own_mutation_methods(self) -> dict[str, Callable]:
        out = {}
        for name, member in inspect.getmembers(type(self), predicate=inspect.isfunction):
            mt = getattr(member, "_mutation_type", None)
            if mt is None or mt in self._disabled or name in self._disabled_names:
                continue
            out[name] = getattr(self, name)
        return out

    def get_mutation_methods(self) -> dict[str, Callable]:
        return self._own_mutation_methods()

    @property
    def mutation_methods(self) -> list[str]:
        return list(self.get_mutation_methods().keys())

    @property
    def layer_mutation_methods(self) -> list[str]:
        return [n for n, m in self.get_mutation_methods().items() if m._mutation_type == MutationType.LAYER]

    @property
    def node_mutation_methods(self) -> list[str]:
        return [n for n, m in self.get_mutation_methods().items() if m._mutation_type == MutationType.NODE]

    def disable_mutations(self, mut_type: MutationType | None = None) -> None:
        """networks/base.py:268-270 disables LAYER mutations of encoders."""
        if mut_type is None:
            self._disabled.update(MutationType)
        else:
            self._disabled.add(mut_type)

    def filter_mutation_methods(self, remove: str) -> None:
        for name in list(self.get_mutation_methods()):
            if remove in name:
                self._disabled_names.add(name)

    def register_mutation_hook(self, hook: Callable) -> None:
        self._mutation_hooks.append(hook)

    def sample_mutation_method(self, new_layer_prob: float, rng: np.random.Generator | None = None) -> Callable:
        """modules/base.py sample: LAYER mutation with probability new_layer_prob, else NODE."""
        rng = self.rng if rng is None else rng
        methods = self.get_mutation_methods()
        layer = [n for n in methods if methods[n]._mutation_type == MutationType.LAYER]
        node = [n for n in methods if methods[n]._mutation_type == MutationType.NODE]
        if not layer and not node:
            raise ValueError("module has no mutation methods")
        if layer and (not node or rng.uniform(0, 1) < new_layer_prob):
            name = rng.choice(layer)
        else:
            name = rng.choice(node)
        return methods[str(name)]

    # -- construction --------------------------------------------------------------------------
    @property
    def init_dict(self) -> dict[str, Any]:
        """Constructor kwargs read back off the instance (modules/base.py:378-392)."""
        params = inspect.signature(type(self).__init__).parameters
        return {k: copy.deepcopy(getattr(self, k)) for k in params if k != "self" and hasattr(self, k)}

    @property
    def net_config(self) -> dict[str, Any]:
        cfg = self.init_dict
        for k in ("num_inputs", "num_outputs", "device", "name", "input_shape", "random_seed"):
            cfg.pop(k, None)
        return cfg

    def clone(self):
        c = type(self)(**self.init_dict)
        c._disabled = set(self._disabled)
        c._disabled_names = set(self._disabled_names)
        c.rng = copy.deepcopy(self.rng)
        c.last_mutation_attr = self.last_mutation_attr
        return c

    def change_activation(self, activation: str, output: bool = False) -> None:
        if output:
            self.output_activation = activation
        self.activation = activation
        for hook in list(self._mutation_hooks):
            hook()
