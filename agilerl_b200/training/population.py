"""Population-level fused learn step.

The reference learns the population agent by agent against ONE shared replay
(train_off_policy.py:399-412: sample -> learn -> update_priorities per agent), so agent i+1's sampler
must see the priorities agent i just wrote.  That ordering is kept exactly: sample, forward, loss and
the priority write-back of all agents form one chain on a single (high-priority) stream.  What does
not feed that chain — each agent's backward, optimiser step, Polyak update and noise reset — is left
running on the agent's own stream underneath the following agents' forwards.
"""
from __future__ import annotations

import torch

_HI: dict = {}


def _hi_priority_stream(device) -> "torch.cuda.Stream":
    key = torch.device(device)
    if key not in _HI:
        _HI[key] = torch.cuda.Stream(device=key, priority=-1)
    return _HI[key]


def population_learn(pop, memory, n_step_memory, overlap: bool = True, join: bool = True) -> list:
    """One ``learn_from_buffers`` step for every agent of ``pop`` (all on one device) against the shared
    HBM-resident buffers.  Returns the per-agent losses as device tensors (no host sync).

    ``join=True`` makes the caller's stream wait for every overlapped tail before returning control —
    required before anything writes the buffers (env ingest) or reads parameters through torch
    (evaluation, tournament, mutation).  Loops that call this back to back pass ``join=False`` and
    join once at the end (``agent.synchronize()``)."""
    if not pop:
        return []
    if not overlap:
        return [agent.learn_from_buffers(memory, n_step_memory) for agent in pop]
    device = pop[0]._dev
    cur = torch.cuda.current_stream(device)
    hi = _hi_priority_stream(device)
    hi.wait_stream(cur)
    with torch.cuda.stream(hi):
        # with few local agents the GPU still has room for each backward's weight gradients on a side stream
        # (a lone agent still gains: its priority write-back runs beside its own backward)
        side = 3 if len(pop) <= 2 else 1
        losses = [agent.learn_from_buffers(memory, n_step_memory, overlap=True, side_streams=side) for agent in pop]
    cur.wait_stream(hi)
    for loss in losses:
        loss.record_stream(cur)
    if join:
        for agent in pop:
            agent.synchronize()
    return losses
