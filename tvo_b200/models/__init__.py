from .bsc import BSC
from .noisyor import NoisyOR
from .sssc import SSSC

__all__ = ["BSC", "NoisyOR", "SSSC"]
