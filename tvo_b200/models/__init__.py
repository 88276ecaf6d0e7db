from .bsc import BSC
from .noisyor import NoisyOR
from .sssc import SSSC
from .mixtures import GMM, PMM

__all__ = ["BSC", "NoisyOR", "SSSC", "GMM", "PMM"]
