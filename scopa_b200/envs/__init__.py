"""Scopa game environments (drop-in for the reference's `envs` package, Miniscopa only:
the 40-card and team variants are outside the accelerated path)."""
from .mini_scopa_game import Card, MiniDeck, MiniScopaEnv, MiniScopaGame, Player

__all__ = ["MiniScopaGame", "MiniScopaEnv", "MiniDeck", "Card", "Player"]
