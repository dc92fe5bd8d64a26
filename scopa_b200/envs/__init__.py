"""Scopa game environments (drop-in for the reference's `envs` package): 1v1 Miniscopa (the solvers' game) and
the 2v2 team variant.  The 40-card game is outside the accelerated path."""
from .mini_scopa_game import Card, MiniDeck, MiniScopaEnv, MiniScopaGame, Player
from .team_mini_scopa_game import TeamMiniScopaEnv, TeamMiniScopaGame

__all__ = ["MiniScopaGame", "MiniScopaEnv", "MiniDeck", "Card", "Player", "TeamMiniScopaGame", "TeamMiniScopaEnv"]
