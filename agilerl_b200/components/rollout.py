"""Device version of ``RolloutBuffer.compute_returns_and_advantages``
(agilerl/components/rollout_buffer.py:413-481) — SURVEY §8(f) rank 2.

The reference copies rewards / dones / values to the host, runs a NumPy loop over the T steps and copies
advantages / returns back.  Here the rollout stays in HBM: one launch of ``b2rl_gae_scan`` (one thread per
environment, float64 carry, float32 stores — bit-identical to the NumPy loop, oracle/gae.py).  Only this
method of the reference's RolloutBuffer is mirrored so far; the rest of the on-policy path is out of scope
for this round.
"""
from __future__ import annotations

import torch

from .. import _lib


def _prep(rewards, dones, values, last_value, last_done):
    _lib.require_cuda_tensor(rewards, "rewards")
    dev = rewards.device
    T, E = int(rewards.shape[0]), int(rewards.shape[1])
    r = rewards.to(torch.float32).reshape(T, E).contiguous()
    v = values.to(device=dev, dtype=torch.float32).reshape(T, E).contiguous()
    d = (dones.to(device=dev).reshape(T, E) != 0).to(torch.uint8).contiguous()      # bool / 0-1 floats -> one byte per flag
    lv = torch.as_tensor(last_value).to(device=dev, dtype=torch.float64).reshape(E).contiguous()   # .astype(float)
    ld = torch.as_tensor(last_done).to(device=dev, dtype=torch.float32).reshape(E).contiguous()
    return dev, T, E, r, d, v, lv, ld


def normalize_advantages(advantages: torch.Tensor) -> torch.Tensor:
    """``(a - a.mean()) / (a.std() + 1e-8)`` over the whole rollout (ppo.py:831-834, :935-944) in one launch."""
    _lib.require_cuda_tensor(advantages, "advantages")
    a = advantages.to(torch.float32).contiguous()
    out = torch.empty_like(a)
    _lib.check(_lib.load(require_cuda=True).b2rl_advantage_normalize(a.data_ptr(), a.numel(), out.data_ptr(),
                                                                    _lib.stream_ptr(a.device)))
    return out


def compute_returns_and_normalized_advantages(rewards, dones, values, last_value, last_done, gamma: float = 0.99,
                                              gae_lambda: float = 0.95, use_gae: bool = True):
    """Rollout post-processing of one PPO learn call without leaving the device: the return / advantage
    recurrence (rollout_buffer.py:413-481) and the global advantage normalisation (ppo.py:831-834) — ONE launch
    for up to 1024 environments.  Returns (advantages, returns, normalized_advantages), float32 [T, E]."""
    dev, T, E, r, d, v, lv, ld = _prep(rewards, dones, values, last_value, last_done)
    adv, ret, nrm = (torch.empty((T, E), dtype=torch.float32, device=dev) for _ in range(3))
    _lib.check(_lib.load(require_cuda=True).b2rl_gae_scan_normalize(
        r.data_ptr(), d.data_ptr(), v.data_ptr(), lv.data_ptr(), ld.data_ptr(), T, E, float(gamma), float(gae_lambda),
        int(bool(use_gae)), adv.data_ptr(), ret.data_ptr(), nrm.data_ptr(), _lib.stream_ptr(dev)))
    return adv, ret, nrm


def compute_returns_and_advantages(rewards: torch.Tensor, dones: torch.Tensor, values: torch.Tensor,
                                   last_value, last_done, gamma: float = 0.99, gae_lambda: float = 0.95,
                                   use_gae: bool = True) -> tuple[torch.Tensor, torch.Tensor]:
    """``rewards`` / ``values`` float32 [T, E], ``dones`` bool [T, E] (what ``RolloutBuffer.buffer`` holds),
    ``last_value`` / ``last_done`` [E] (tensor or array, as the reference accepts).  Returns
    (advantages, returns), float32 [T, E] on the rollout's device.  CUDA only."""
    dev, T, E, r, d, v, lv, ld = _prep(rewards, dones, values, last_value, last_done)
    adv = torch.empty((T, E), dtype=torch.float32, device=dev)
    ret = torch.empty((T, E), dtype=torch.float32, device=dev)
    lib = _lib.load(require_cuda=True)
    _lib.check(lib.b2rl_gae_scan(r.data_ptr(), d.data_ptr(), v.data_ptr(), lv.data_ptr(), ld.data_ptr(), T, E,
                                 float(gamma), float(gae_lambda), int(bool(use_gae)), adv.data_ptr(), ret.data_ptr(),
                                 _lib.stream_ptr(dev)))
    return adv, ret
