from .dqn import DQN
from .dqn_rainbow import RainbowDQN

__all__ = ["DQN", "RainbowDQN"]
