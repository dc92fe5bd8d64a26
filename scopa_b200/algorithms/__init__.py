"""CUDA-backed solvers under the reference's package name and exports
(/root/reference/src/algorithms/__init__.py): vanilla CFR and the sampled (MC) CFR trainer with their
tabular policies.  The SDCFR classes live in the `deep_cfr` sub-package, as upstream."""
from . import mc_cfr as _mc, vanilla_cfr as _cfr

CFRTrainer, InfoNode = _cfr.CFRTrainer, _cfr.InfoNode
LearnedCFRPolicy, RandomPolicy = _cfr.LearnedCFRPolicy, _cfr.RandomPolicy
MCCFRTrainer, ScopaLearnedPolicy = _mc.MCCFRTrainer, _mc.ScopaLearnedPolicy

__all__ = ["CFRTrainer", "InfoNode", "LearnedCFRPolicy", "RandomPolicy", "MCCFRTrainer", "ScopaLearnedPolicy"]
