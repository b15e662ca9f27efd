"""ORACLE (test infrastructure, never imported by the product path).

CPU restatement of the reference's replay buffers over a plain dict of torch tensors.

Follows (reference file:line):
  * ReplayBuffer.add / _init / sample / clear          agilerl/components/replay_buffer.py:60-138
  * MultiStepReplayBuffer.add / _get_n_step_info       agilerl/components/replay_buffer.py:173-258
  * PrioritizedReplayBuffer.add / _update_priority /
    sample / _sample_proportional / _calculate_weights /
    update_priorities                                  agilerl/components/replay_buffer.py:296-428
"""
from __future__ import annotations

from collections import deque

import torch

from .segtree import CSegTree, PySegTree


def _rows(data: dict) -> int:
    return next(iter(data.values())).shape[0]


class OracleReplay:
    """replay_buffer.py:12-138."""

    def __init__(self, max_size: int):
        self.max_size = max_size
        self.storage: dict[str, torch.Tensor] | None = None
        self.cursor = 0
        self.size = 0
        self.counter = 0

    def add(self, data: dict) -> None:
        data = {k: (v.reshape(-1, 1) if v.ndim == 1 else v) for k, v in data.items()}  # :85-94
        n = _rows(data)
        if self.storage is None:                                                      # :60-70
            self.storage = {k: torch.zeros((self.max_size, *v.shape[1:]), dtype=v.dtype)
                            for k, v in data.items()}
        start, end = self.cursor, self.cursor + n
        if end > self.max_size:                                                       # :100-107
            k = self.max_size - start
            for key, v in data.items():
                self.storage[key][start:] = v[:k]
                self.storage[key][: n - k] = v[k:]
        else:
            for key, v in data.items():
                self.storage[key][start:end] = v
        self.cursor = end % self.max_size
        self.size = min(self.size + n, self.max_size)
        self.counter += n

    def gather(self, idx: torch.Tensor) -> dict:
        return {k: v[idx] for k, v in self.storage.items()}

    def sample(self, batch_size: int, return_idx: bool = False) -> dict:
        idx = torch.randperm(self.size)[:batch_size]                                 # :125
        out = self.gather(idx)
        if return_idx:
            out["idxs"] = idx
        return out


def n_step_roll(window: list[dict], gamma: float, reward_key="reward", done_key="done",
                ns_key="next_obs") -> dict:
    """replay_buffer.py:206-258: fold a window of n per-env transition batches into one n-step
    transition (quirk Q3: reward added before done is examined, break on ANY env done, first
    transition's own done ignored, gamma**(i+1) a Python double)."""
    first = {k: v.clone() for k, v in window[0].items()}
    R = first[reward_key].clone()
    for i, tr in enumerate(window[1:]):
        R += tr[reward_key] * (gamma ** (i + 1))
        first[ns_key] = tr[ns_key].clone()
        first[done_key] = tr[done_key].clone()
        if tr[done_key].bool().any():
            break
    first[reward_key] = R
    return first


class OracleNStep(OracleReplay):
    """replay_buffer.py:141-258."""

    def __init__(self, max_size: int, n_step: int = 3, gamma: float = 0.99):
        super().__init__(max_size)
        self.n_step, self.gamma = n_step, gamma
        self.window: deque = deque(maxlen=n_step)

    def add(self, data: dict):
        self.window.append(data)
        if len(self.window) < self.n_step:
            return None
        done_key = next(k for k in ("done", "termination", "terminated") if k in data)  # :224-234
        super().add(n_step_roll(list(self.window), self.gamma, done_key=done_key))
        return self.window[0]


class OraclePER(OracleReplay):
    """replay_buffer.py:261-428.  ``tree_cls`` = CSegTree (fast) or PySegTree (reference cost)."""

    def __init__(self, max_size: int, alpha: float = 0.6, tree_cls=CSegTree):
        super().__init__(max_size)
        self.alpha = alpha
        self.max_priority = 1.0
        self.tree_ptr = 0
        cap = 1
        while cap < max_size:
            cap *= 2
        self.sum_tree = tree_cls(cap, "sum")
        self.min_tree = tree_cls(cap, "min")

    def add(self, data: dict) -> None:
        super().add(data)
        for _ in range(_rows(data)):                                                 # :306-309
            self.update_priority(self.tree_ptr, self.max_priority)
            self.tree_ptr = (self.tree_ptr + 1) % self.max_size

    def update_priority(self, idx: int, priority: float) -> None:                    # :311-329
        assert 0 <= idx < self.max_size
        pa = priority ** self.alpha
        self.sum_tree[idx] = pa
        self.min_tree[idx] = pa
        self.max_priority = max(self.max_priority, priority)

    def sample_proportional(self, batch_size: int, uniforms=None) -> torch.Tensor:    # :357-381
        idx = torch.zeros(batch_size, dtype=torch.int64)
        total = self.sum_tree.sum()
        segment = total / batch_size
        for i in range(batch_size):
            a, b = segment * i, segment * (i + 1)
            u = torch.rand(1).item() if uniforms is None else float(uniforms[i])
            idx[i] = self.sum_tree.retrieve(u * (b - a) + a)
        return idx

    def calculate_weights(self, idx: torch.Tensor, beta: float) -> torch.Tensor:      # :383-409
        w = torch.zeros(len(idx))
        p_min = self.min_tree.min() / self.sum_tree.sum()
        max_w = (p_min * self.size) ** -beta
        for i, j in enumerate(idx):
            p = self.sum_tree[int(j)] / self.sum_tree.sum()
            w[i] = ((p * self.size) ** -beta) / max_w
        return w

    def sample(self, batch_size: int, beta: float = 0.4, uniforms=None) -> dict:       # :331-355
        idx = self.sample_proportional(batch_size, uniforms)
        out = {k: v.clone() for k, v in self.gather(idx).items()}
        out["weights"] = self.calculate_weights(idx, beta).unsqueeze(1)
        out["idxs"] = idx.unsqueeze(1)
        return out

    def update_priorities(self, idx, priorities) -> None:                             # :411-428
        for i, p in zip(idx, priorities):
            self.update_priority(int(i.item() if hasattr(i, "item") else i),
                                 max(float(p.item() if hasattr(p, "item") else p), 1e-5))
