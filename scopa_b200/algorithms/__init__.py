"""CFR algorithms (drop-in for the reference's `algorithms` package), CUDA-backed."""
from .vanilla_cfr import CFRTrainer, InfoNode, LearnedCFRPolicy, RandomPolicy
from .mc_cfr import MCCFRTrainer, ScopaLearnedPolicy

__all__ = ["CFRTrainer", "InfoNode", "LearnedCFRPolicy", "RandomPolicy", "MCCFRTrainer", "ScopaLearnedPolicy"]
