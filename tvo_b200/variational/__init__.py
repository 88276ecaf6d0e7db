from .TVOVariationalStates import TVOVariationalStates
from .fullem import FullEM, FullEMSingleCauseModels
from .RandomSampledVarStates import RandomSampledVarStates
from .evo import EVOVariationalStates
from .tvs import TVSVariationalStates
from ._packed import PackedStates

__all__ = ["TVOVariationalStates", "FullEM", "FullEMSingleCauseModels", "RandomSampledVarStates", "EVOVariationalStates", "TVSVariationalStates",
           "PackedStates"]
