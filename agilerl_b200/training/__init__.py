from .population import population_learn
from .train_off_policy import train_off_policy

__all__ = ["train_off_policy", "population_learn"]
