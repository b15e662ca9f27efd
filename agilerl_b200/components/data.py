"""``Transition`` — mirror of agilerl/components/data.py:68-93: a keyword container that casts
action / reward / done to float32 tensors (observations are left untouched, so uint8 frames stay
uint8) and converts to a TensorDict with ``to_tensordict()``."""
from __future__ import annotations

from numbers import Number

import numpy as np
import torch

from ..compat import TensorDict, tensorclass


def to_torch_tensor(data, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """data.py:47-65."""
    if isinstance(data, (np.ndarray, Number, bool)):
        return torch.tensor(data, dtype=dtype)
    if isinstance(data, torch.Tensor):
        return data.to(dtype=dtype)
    return torch.tensor(data, dtype=dtype)


def to_tensordict(data, dtype: torch.dtype = torch.float32) -> TensorDict:
    """data.py:15-44 (dict / tuple observations)."""
    ok = (torch.Tensor, np.ndarray, Number)
    if isinstance(data, tuple):
        assert all(isinstance(el, ok) for el in data), "Expected all elements of the tuple to be torch.Tensor or np.ndarray."
        data = TensorDict({f"tuple_obs_{i}": el for i, el in enumerate(data)})
    elif isinstance(data, dict):
        assert all(isinstance(el, ok) for el in data.values()), "Expected all values of the dict to be torch.Tensor or np.ndarray."
        data = TensorDict(data)
    return data.to(dtype=dtype)


@tensorclass
class Transition:
    obs: object
    action: object
    next_obs: object
    reward: object
    done: object

    def __post_init__(self) -> None:
        if isinstance(self.obs, (dict, tuple)):
            self.obs = to_tensordict(self.obs)
        if isinstance(self.next_obs, (dict, tuple)):
            self.next_obs = to_tensordict(self.next_obs)
        self.action = to_torch_tensor(self.action)
        self.done = to_torch_tensor(self.done)
        self.reward = to_torch_tensor(self.reward)
        if self.done.ndim == 0:
            self.done = self.done.unsqueeze(-1)
        if self.reward.ndim == 0:
            self.reward = self.reward.unsqueeze(-1)


class ReplayDataset:
    """agilerl/components/data.py:96-118 — the accelerate DataLoader bridge of the reference's distributed replay.
    The population is sharded one process per GPU here (no accelerate), so the class only exists for the
    ``from agilerl.components.data import ReplayDataset`` line of the unchanged driver."""

    def __init__(self, buffer, batch_size: int = 256) -> None:
        raise NotImplementedError("accelerate-sharded replay is replaced by one-agent-per-GPU sharding (DESIGN.md section 6)")
