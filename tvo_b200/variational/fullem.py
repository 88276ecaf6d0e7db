"""Exhaustive E-step (mirror of tvo/variational/fullem.py:14-58)."""
import torch as to
from torch import Tensor

import tvo_b200
from tvo_b200 import _lib
from tvo_b200.utils.model_protocols import Optimized
from .TVOVariationalStates import TVOVariationalStates
from ._packed import PackedStates, n_words, unpack


def state_matrix(H: int, device: to.device = None) -> to.Tensor:
    """All 2^H states, row i = binary of i, most significant bit first (fullem.py:14-28); uint8 (2^H, H)."""
    return unpack(_state_matrix_packed(H, device), H)


def _state_matrix_packed(H: int, device=None) -> to.Tensor:
    device = tvo_b200.get_device() if device is None else device
    out = to.empty((2 ** H, n_words(H)), dtype=to.int64, device=device)
    _lib.call("tvo_state_matrix", _lib.ptr(out), H, _lib.stream())
    return out


class FullEM(TVOVariationalStates):
    def __init__(self, N: int, H: int, precision: to.dtype, K_init=None):
        conf = dict(N=N, S=None, S_new=None, H=H, precision=precision)
        self.config = conf
        dev = tvo_b200.get_device()
        self.lpj = to.empty((N, 2 ** H), dtype=precision, device=dev)
        self.precision = precision
        self._H = H
        self._N = N
        self._Kp = _state_matrix_packed(H, dev)[None]  # (1, 2^H, W64): one set shared by all datapoints
        self._nsubs_dev = to.zeros((), dtype=to.int64, device=dev)
        self._nsubs_read = 0
        self.seed, self.n_offset, self._step = 0, 0, 0

    @property
    def K(self) -> PackedStates:
        return PackedStates(self._Kp, self._H, N=self._N)

    @K.setter
    def K(self, value):
        raise AttributeError("FullEM.K is the fixed full state matrix")

    def update(self, idx: Tensor, batch: Tensor, model) -> int:
        self.update_device(idx, batch, model)
        return 0

    def update_device(self, idx: Tensor, batch: Tensor, model) -> None:
        idx = idx.to(device=self.lpj.device, dtype=to.int64).contiguous()
        if hasattr(model, "_tvo_lpj_rows"):
            model._tvo_lpj_rows(self, idx, batch)
        else:
            lpj_fn = model.log_pseudo_joint if isinstance(model, Optimized) else model.log_joint
            self.lpj[idx] = lpj_fn(batch, self.K[idx]).to(self.lpj.dtype)
