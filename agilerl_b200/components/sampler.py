"""``Sampler`` — mirror of agilerl/components/sampler.py:53-202 for in-memory buffers: picks
``sample_per`` / ``sample_n_step`` / ``sample_standard`` by ``isinstance`` on the buffer class,
exactly like the reference (:71-72, :92-113).  The DataLoader ("distributed") branch belongs to
the accelerate path, which this build replaces with one-agent-per-GPU sharding."""
from __future__ import annotations

import warnings
from typing import Any

from .multi_agent_replay_buffer import MultiAgentReplayBuffer
from .replay_buffer import MultiStepReplayBuffer, PrioritizedReplayBuffer, ReplayBuffer


class Sampler:
    def __init__(self, memory=None, dataset=None, dataloader=None) -> None:
        assert (memory is not None) or ((dataset is not None) and (dataloader is not None)), (
            "Sampler needs to be initialized with either 'memory' or ('dataset' AND 'dataloader')."
        )
        if memory is None:
            raise NotImplementedError(
                "DataLoader-based distributed sampling is replaced by per-GPU buffers in agilerl_b200")
        self.distributed = False
        self.per = isinstance(memory, PrioritizedReplayBuffer)
        self.n_step = isinstance(memory, MultiStepReplayBuffer)
        self.memory = memory
        self.dataset = dataset
        self.dataloader = dataloader
        if self.per:
            if not isinstance(self.memory, PrioritizedReplayBuffer):                           # sampler.py:92-97 (kept check)
                warnings.warn("Memory is not an agilerl PrioritizedReplayBuffer.", stacklevel=2)
            self.sample = self.sample_per
        elif self.n_step:
            if not isinstance(self.memory, MultiStepReplayBuffer):                             # sampler.py:99-104
                warnings.warn("Memory is not an agilerl MultiStepReplayBuffer.", stacklevel=2)
            self.sample = self.sample_n_step
        else:
            if not isinstance(self.memory, (ReplayBuffer, MultiAgentReplayBuffer)):          # sampler.py:106-111
                warnings.warn("Memory is not an agilerl ReplayBuffer or MultiAgentReplayBuffer.", stacklevel=2)
            self.sample = self.sample_standard

    def sample_standard(self, batch_size: int, return_idx: bool = False):
        return self.memory.sample(batch_size, return_idx)

    def sample_per(self, batch_size: int, beta: float):
        return self.memory.sample(batch_size, beta)

    def sample_n_step(self, idxs: Any):
        return self.memory.sample_from_indices(idxs)
