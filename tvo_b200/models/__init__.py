from .bsc import BSC
from .noisyor import NoisyOR

__all__ = ["BSC", "NoisyOR"]
