from .multi_agent_replay_buffer import MultiAgentReplayBuffer
from .replay_buffer import MultiStepReplayBuffer, PrioritizedReplayBuffer, ReplayBuffer
from .sampler import Sampler
from .data import Transition
from .segment_tree import MinSegmentTree, SegmentTree, SumSegmentTree

__all__ = ["ReplayBuffer", "MultiStepReplayBuffer", "PrioritizedReplayBuffer", "MultiAgentReplayBuffer", "Sampler", "Transition",
           "SegmentTree", "SumSegmentTree", "MinSegmentTree"]
