"""Base class of the variational-state containers (mirror of
tvo/variational/TVOVariationalStates.py:18-56).

Storage differs from the reference by design: K lives bit-packed in HBM (int64 (N,S,W64) holding
uint64 patterns) and `K` is a property returning a PackedStates view with the reference's uint8
(N,S,H) semantics.  `lpj` is (N,S) in `precision` on tvo_b200.get_device().  A device-side int64
counter accumulates the number of substitutions so that E-steps never force a host sync."""
from abc import ABC, abstractmethod
from typing import Any, Dict

import torch as to
from torch import Tensor

import tvo_b200
from tvo_b200.utils import get
from ._packed import PackedStates, pack
from ._utils import generate_unique_states


class TVOVariationalStates(ABC):
    def __init__(self, conf: Dict[str, Any], K_init: Tensor = None):
        required_keys = ("N", "H", "S", "S_new", "precision")
        for c in required_keys:
            assert c in conf and conf[c] is not None
        self.config = conf
        N, H, S, _, precision = get(conf, *required_keys)
        dev = tvo_b200.get_device()
        if conf.get("K_init_file") is not None:  # TVOVariationalStates.py:30-38 (.h5 needs h5py; .npz always works)
            from tvo_b200.utils.io import load_initial_states
            K_init = load_initial_states(conf["K_init_file"], N, S, H)
        if K_init is not None:
            assert tuple(K_init.shape) == (N, S, H)
            K_u8 = K_init.clone().to(dtype=to.uint8, device=dev)
            self._Kp = pack(K_u8)
        else:
            # S unique Bernoulli(1/H) states, identical for every datapoint (TVOVariationalStates.py:43)
            row = pack(generate_unique_states(S, H, device=dev)[None])  # (1,S,W64)
            self._Kp = row.repeat(N, 1, 1)
        self._H = H
        self.lpj = to.empty((N, S), dtype=precision, device=dev)
        self.precision = precision
        self._nsubs_dev = to.zeros((), dtype=to.int64, device=dev)
        self._nsubs_read = 0
        self.seed = int(conf.get("seed") if conf.get("seed") is not None else to.initial_seed()) & (2 ** 64 - 1)
        self.n_offset = int(conf.get("n_offset") or 0)  # global index of local datapoint 0 (rank shard offset)
        self._step = 0

    # reference-facing uint8 view ------------------------------------------------------------
    @property
    def K(self) -> PackedStates:
        return PackedStates(self._Kp, self._H)

    @K.setter
    def K(self, value):
        if isinstance(value, PackedStates):
            self._Kp, self._H = value.packed.clone(), value.H
        else:
            value = value.to(device=tvo_b200.get_device(), dtype=to.uint8)
            self._Kp, self._H = pack(value.contiguous()), value.shape[-1]

    def _take_nsubs(self) -> int:
        """Host read (one sync) of the substitutions accumulated since the last read."""
        total = int(self._nsubs_dev.item())
        delta, self._nsubs_read = total - self._nsubs_read, total
        return delta

    @abstractmethod
    def update(self, idx: Tensor, batch: Tensor, model) -> int:
        """Generate new states, keep the S best per datapoint; returns the number of substitutions."""
        pass  # pragma: no cover

    def update_device(self, idx: Tensor, batch: Tensor, model) -> None:
        """Same as update() but without the host sync; read the count with _take_nsubs()."""
        raise NotImplementedError
