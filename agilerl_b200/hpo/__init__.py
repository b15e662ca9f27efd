from .mutation import Mutations
from .tournament import TournamentSelection

__all__ = ["Mutations", "TournamentSelection"]
