"""Scopa game environments (drop-in for the reference's `envs` package): 1v1 Miniscopa (the solvers' game), the 2v2
team variant and the 40-card game, all evaluated on the GPU."""
from .full_scopa_game import FullDeck, FullScopaEnv, FullScopaGame
from .mini_scopa_game import Card, MiniDeck, MiniScopaEnv, MiniScopaGame, Player
from .team_mini_scopa_game import TeamMiniScopaEnv, TeamMiniScopaGame

__all__ = ["MiniScopaGame", "MiniScopaEnv", "MiniDeck", "Card", "Player", "TeamMiniScopaGame", "TeamMiniScopaEnv",
           "FullScopaGame", "FullScopaEnv", "FullDeck"]
