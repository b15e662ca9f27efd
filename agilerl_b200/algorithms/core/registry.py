"""Hyper-parameter / network registries — mirror of agilerl/algorithms/core/registry.py:108-320
(``RLParameter`` :108-186, ``HyperparameterConfig`` :189-241, ``NetworkGroup``, ``OptimizerConfig``,
``MutationRegistry``)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

import numpy as np
import torch


@dataclass
class RLParameter:
    min: float
    max: float
    shrink_factor: float = 0.8
    grow_factor: float = 1.2
    dtype: type = float
    value: Any = field(default=None, init=False)

    def mutate(self):
        """registry.py:135-186: coin flip torch.rand(1) < 0.5 -> shrink else grow, clamp, cast."""
        assert self.value is not None, "Hyperparameter value is not set"
        if torch.rand(1).item() < 0.5:
            new_value = self.value * self.shrink_factor if self.value * self.shrink_factor > self.min else self.min
        else:
            new_value = self.value * self.grow_factor if self.value * self.grow_factor < self.max else self.max
        new_value = min(max(new_value, self.min), self.max)
        self.value = self.dtype(new_value)
        return self.value


class HyperparameterConfig:
    def __init__(self, **kwargs: RLParameter) -> None:
        self.config = kwargs
        for key, value in kwargs.items():
            if not isinstance(value, RLParameter):
                raise TypeError("Expected RLParameter object for hyperparameter configuration.")
            setattr(self, key, value)

    def __bool__(self) -> bool:
        return bool(self.config)

    def __eq__(self, other) -> bool:
        return set(self.names()) == set(other.names())

    def __iter__(self):
        return iter(self.config)

    def __getitem__(self, key: str) -> RLParameter:
        return self.config[key]

    def names(self) -> list[str]:
        return list(self.config.keys())

    def items(self):
        return self.config.items()

    def sample(self) -> tuple[str, RLParameter]:
        key = int(torch.randperm(len(self.config))[0])           # registry.py:234-241
        return list(self.config.keys())[key], list(self.config.values())[key]


@dataclass
class NetworkGroup:
    eval_network: str
    shared_networks: list | None = None
    policy: bool = False

    def __post_init__(self):
        if isinstance(self.shared_networks, str):
            self.shared_networks = [self.shared_networks]


@dataclass
class OptimizerConfig:
    name: str
    networks: list
    lr: str
    optimizer_cls: str = "Adam"
    optimizer_kwargs: dict = field(default_factory=dict)


class MutationRegistry:
    def __init__(self, hp_config: HyperparameterConfig | None = None) -> None:
        self.hp_config = hp_config if hp_config is not None else HyperparameterConfig()
        self.groups: list[NetworkGroup] = []
        self.optimizers: list[OptimizerConfig] = []
        self.hooks: list = []

    def register_group(self, group: NetworkGroup) -> None:
        self.groups.append(group)

    def register_optimizer(self, cfg: OptimizerConfig) -> None:
        self.optimizers.append(cfg)

    def policy(self, return_group: bool = False):
        for g in self.groups:
            if g.policy:
                return g if return_group else g.eval_network
        return None

    def all_registered(self) -> list[str]:
        out = []
        for g in self.groups:
            out.append(g.eval_network)
            out.extend(g.shared_networks or [])
        return out

    def __eq__(self, other) -> bool:
        return (self.hp_config == other.hp_config and [g.eval_network for g in self.groups] ==
                [g.eval_network for g in other.groups])
