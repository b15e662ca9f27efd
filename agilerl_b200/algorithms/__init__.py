from .dqn import DQN
from .dqn_rainbow import RainbowDQN
from .maddpg import MADDPG
from .td3 import DDPG, TD3

__all__ = ["DQN", "RainbowDQN", "DDPG", "TD3", "MADDPG"]
