"""Drop-in for the reference's src/algorithms/deep_cfr package (SDCFR)."""
from .deep_cfr import AdvantageNetwork, DeepCFR, RandomPolicy, StrategyBuffer
from .nets import FlexibleNet, MLPBlock, masked_softmax, positive_regret_policy

__all__ = ["AdvantageNetwork", "DeepCFR", "RandomPolicy", "StrategyBuffer", "FlexibleNet", "MLPBlock",
           "masked_softmax", "positive_regret_policy"]
