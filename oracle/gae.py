"""ORACLE (test infrastructure, never imported by the product path) — SURVEY §8(f) rank 2.

NumPy restatement of ``RolloutBuffer.compute_returns_and_advantages``
(agilerl/components/rollout_buffer.py:413-481): the reverse recurrence PPO runs on the host after every
rollout.  The dtype walk is part of the contract: rewards / values / dones arrive as float32, but
``1.0 - dones.astype(float)`` and ``last_value.astype(float)`` are float64, so every ``delta`` and the
carried ``last_gae_lambda`` are float64 and each row is rounded to float32 only when it is stored;
``returns = advantages(f32) + values(f32)``.  Monte-Carlo mode likewise carries float64.
Pinned bit-exactly against the unmodified reference by tests/golden/make_golden.py::gen_gae.
"""
from __future__ import annotations

import numpy as np


def compute_returns_and_advantages(rewards: np.ndarray, dones: np.ndarray, values: np.ndarray, last_value: np.ndarray,
                                   last_done: np.ndarray, gamma: float, gae_lambda: float, use_gae: bool = True):
    """rewards / dones / values: float32 [T, E]; last_value / last_done: [E].  Returns (advantages, returns) f32 [T, E]."""
    T, E = rewards.shape
    last_value = np.asarray(last_value).reshape(E)
    last_done = np.asarray(last_done).reshape(E)
    advantages = np.zeros((T, E), dtype=np.float32)
    returns = np.zeros((T, E), dtype=np.float32)
    if use_gae:
        last_gae_lambda = np.zeros(E, dtype=np.float32)
        for t in reversed(range(T)):
            if t == T - 1:
                next_non_terminal = 1.0 - last_done.astype(float)
                next_values = last_value.astype(float)
            else:
                next_non_terminal = 1.0 - dones[t + 1].astype(float)
                next_values = values[t + 1]
            delta = rewards[t] + gamma * next_values * next_non_terminal - values[t]
            advantages[t] = last_gae_lambda = delta + gamma * gae_lambda * next_non_terminal * last_gae_lambda
        returns = advantages + values
    else:
        last_returns = last_value.astype(float) * (1.0 - last_done.astype(float))
        for t in reversed(range(T)):
            returns[t] = last_returns = rewards[t] + gamma * last_returns * (1.0 - dones[t].astype(float))
        advantages = returns - values
    return advantages, returns


def normalize_advantages(advantages: np.ndarray) -> np.ndarray:
    """PPO's global advantage normalisation, agilerl/algorithms/ppo.py:831-834 (flat path) and :935-944
    (recurrent path): ``(a - a.mean()) / (a.std() + 1e-8)`` with torch's float32 reductions and Bessel-corrected
    ``std`` — the literal torch expression on CPU."""
    import torch
    flat = torch.from_numpy(np.ascontiguousarray(advantages, dtype=np.float32)).reshape(-1)
    out = (flat - flat.mean()) / (flat.std() + 1e-8)
    return out.reshape(advantages.shape).numpy()
