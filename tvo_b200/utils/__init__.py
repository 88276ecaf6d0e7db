from ._utils import get
from .param_init import init_W_data_mean, init_sigma2_default, init_pies_default

__all__ = ["get", "init_W_data_mean", "init_sigma2_default", "init_pies_default"]
