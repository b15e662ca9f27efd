from .population import multi_agent_population_learn, population_learn
from .train_multi_agent_off_policy import train_multi_agent_off_policy
from .train_off_policy import train_off_policy

__all__ = ["train_off_policy", "train_multi_agent_off_policy", "population_learn", "multi_agent_population_learn"]
