"""Parameter initialisers (mirror of tvo/utils/param_init.py:11-70), torch on the device."""
import math

import torch as to

import tvo_b200


def init_W_data_mean(data: to.Tensor, H: int, std_factor: float = 0.25, dtype=to.float64, device=None) -> to.Tensor:
    """W = mean(data) + N(0, (std_factor * mean std)^2), shape (D, H)."""
    device = tvo_b200.get_device() if device is None else device
    data_nanmean = to.nanmean(data.to(to.float64), dim=0) if hasattr(to, "nanmean") else data.mean(0)
    var = to.nanmean((data.to(to.float64) - data_nanmean) ** 2, dim=0)
    std = to.sqrt(var).mean()
    D = data.shape[1]
    W = data_nanmean[:, None].repeat(1, H) + std_factor * std * to.randn(D, H, dtype=to.float64, device=data.device)
    return W.to(dtype=dtype, device=device)


def init_sigma2_default(data: to.Tensor, dtype=to.float64, device=None) -> to.Tensor:
    device = tvo_b200.get_device() if device is None else device
    d = data.to(to.float64)
    m = to.nanmean(d, dim=0)
    var = to.nanmean((d - m) ** 2, dim=0)
    return var.mean().reshape(1).to(dtype=dtype, device=device)


def init_pies_default(H: int, crowdedness: float = 2.0, dtype=to.float64, device=None) -> to.Tensor:
    device = tvo_b200.get_device() if device is None else device
    return to.full((H,), crowdedness / H if crowdedness / H < 1.0 else 1.0 / math.e, dtype=dtype, device=device)
