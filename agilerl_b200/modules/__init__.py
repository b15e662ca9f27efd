from .base import EvolvableModule, MutationType, mutation
from .cnn import EvolvableCNN
from .mlp import EvolvableMLP

__all__ = ["EvolvableModule", "MutationType", "mutation", "EvolvableCNN", "EvolvableMLP"]
