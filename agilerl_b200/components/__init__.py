from .replay_buffer import MultiStepReplayBuffer, PrioritizedReplayBuffer, ReplayBuffer
from .sampler import Sampler
from .data import Transition
from .segment_tree import MinSegmentTree, SegmentTree, SumSegmentTree

__all__ = ["ReplayBuffer", "MultiStepReplayBuffer", "PrioritizedReplayBuffer", "Sampler", "Transition",
           "SegmentTree", "SumSegmentTree", "MinSegmentTree"]
