"""agilerl_b200 — B200-native (sm_100a) implementation of AgileRL's population-parallel off-policy
``learn()`` hot path behind the reference's Python API.  CUDA only; no CPU fallback."""
__version__ = "0.1.0"
