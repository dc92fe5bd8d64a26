"""scopa_b200 -- B200-native (sm_100a CUDA) implementation of the Miniscopa hot path of
rug-marl-group2/scopa: the game env and the CFR / MCCFR / SDCFR traversal loops, behind the
reference's Python class API.  See DESIGN.md."""
from . import _lib  # noqa: F401
from ._lib import MsError  # noqa: F401

__version__ = "0.1.0"
