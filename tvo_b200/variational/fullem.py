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
        supports = getattr(model, "_supports_fast", None)
        if hasattr(model, "_tvo_lpj_rows") and (supports is None or supports(self)):
            model._tvo_lpj_rows(self, idx, batch)
        else:
            lpj_fn = model.log_pseudo_joint if isinstance(model, Optimized) else model.log_joint
            self.lpj[idx] = lpj_fn(batch, self.K[idx]).to(self.lpj.dtype)


class SingleCauseK:
    """The state set of FullEMSingleCauseModels: K_n = eye(H) for every datapoint (fullem.py:79).  Behaves like the
    reference's expanded (N,H,H) tensor under indexing / .to / .sum, but carries the `single_cause` marker the
    tvo_b200 mixture models use to take their dense kernels instead of materialising anything."""
    single_cause = True

    def __init__(self, N: int, H: int, precision: to.dtype, device):
        self.N, self.H, self.precision, self.device = N, H, precision, device
        self._eye = to.eye(H, dtype=precision, device=device)

    @property
    def shape(self):
        return (self.N, self.H, self.H)

    def dense(self) -> Tensor:
        return self._eye[None, :, :].expand(self.N, -1, -1)

    def __getitem__(self, idx):
        if isinstance(idx, Tensor) and idx.dim() == 1 or isinstance(idx, slice):
            n = idx.numel() if isinstance(idx, Tensor) else len(range(*idx.indices(self.N)))
            return SingleCauseK(n, self.H, self.precision, self.device)
        return self.dense()[idx]

    def to(self, *args, **kwargs):
        return self.dense().to(*args, **kwargs)

    def sum(self, *args, **kwargs):
        return self.dense().sum(*args, **kwargs)


class FullEMSingleCauseModels(FullEM):
    """fullem.py:61-82: full EM for single-cause models (GMM, PMM): the H one-hot states for every datapoint."""
    single_cause = True

    def __init__(self, N: int, H: int, precision: to.dtype):
        conf = dict(N=N, S=None, S_new=None, H=H, precision=precision)
        for c in ("N", "H", "precision"):
            assert c in conf and conf[c] is not None
        self.config = conf
        dev = tvo_b200.get_device()
        self.lpj = to.empty((N, H), dtype=precision, device=dev)
        self.precision = precision
        self._H, self._N = H, N
        self._K = SingleCauseK(N, H, precision, dev)
        self._nsubs_dev = to.zeros((), dtype=to.int64, device=dev)
        self._nsubs_read = 0
        self.seed, self.n_offset, self._step = 0, 0, 0

    @property
    def K(self) -> SingleCauseK:
        return self._K

    @K.setter
    def K(self, value):
        raise AttributeError("FullEMSingleCauseModels.K is the fixed identity state set")

    def update_device(self, idx: Tensor, batch: Tensor, model) -> None:
        idx = idx.to(device=self.lpj.device, dtype=to.int64).contiguous()
        if hasattr(model, "_tvo_lpj_rows") and getattr(model, "KIND", None) is not None:
            model._tvo_lpj_rows(self, idx, batch)
        else:  # third-party single-cause model: the reference's call with the dense identity states
            lpj_fn = model.log_pseudo_joint if isinstance(model, Optimized) else model.log_joint
            self.lpj[idx] = lpj_fn(batch, self.K[idx].dense()).to(self.lpj.dtype)
