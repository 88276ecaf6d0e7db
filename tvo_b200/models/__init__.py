from .bsc import BSC
from .noisyor import NoisyOR
from .sssc import SSSC
from .mixtures import GMM, PMM
from .tvae import GaussianTVAE, BernoulliTVAE

TVAE = GaussianTVAE  # tvo/models/__init__.py:8

__all__ = ["BSC", "NoisyOR", "SSSC", "GMM", "PMM", "GaussianTVAE", "BernoulliTVAE", "TVAE"]
