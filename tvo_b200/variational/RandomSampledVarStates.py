"""Random-sampling E-step (mirror of tvo/variational/RandomSampledVarStates.py:13-64)."""
from typing import Optional

import torch as to
from torch import Tensor

from tvo_b200 import _lib
from tvo_b200.utils.model_protocols import Optimized
from .TVOVariationalStates import TVOVariationalStates
from ._packed import unpack
from ._utils import select_topS_packed


class RandomSampledVarStates(TVOVariationalStates):
    def __init__(self, N: int, H: int, S: int, precision: to.dtype, S_new: int, sparsity: float = 0.5,
                 K_init_file: str = None, seed: Optional[int] = None, K_init: Tensor = None):
        conf = dict(N=N, H=H, S=S, precision=precision, S_new=S_new, sparsity=sparsity, K_init_file=K_init_file,
                    seed=seed)
        super().__init__(conf, K_init=K_init)

    def update(self, idx: Tensor, batch: Tensor, model) -> int:
        self.update_device(idx, batch, model)
        return self._take_nsubs()

    def update_device(self, idx: Tensor, batch: Tensor, model) -> None:
        """lpj[idx] = lpj_fn(batch, K[idx]); S_new Bernoulli(sparsity) states; NO dedup (quirk Q4);
        top-S merge (RandomSampledVarStates.py:46-64)."""
        idx = idx.to(device=self.lpj.device, dtype=to.int64).contiguous()
        B, S_new, H = idx.numel(), self.config["S_new"], self._H
        W = self._Kp.shape[2]
        new_K = to.empty((B, S_new, W), dtype=to.int64, device=self.lpj.device)
        _lib.call("tvo_random_states", _lib.ptr(new_K), _lib.ptr(idx), B, S_new, H, float(self.config["sparsity"]),
                  self.seed, self._step & 0xFFFFFFFF, self.n_offset & 0xFFFFFFFF, _lib.stream())
        self._step += 1
        if hasattr(model, "_tvo_lpj_packed"):
            model._tvo_lpj_rows(self, idx, batch)
            new_lpj = model._tvo_lpj_packed(batch, new_K, H)
        else:
            lpj_fn = model.log_pseudo_joint if isinstance(model, Optimized) else model.log_joint
            self.lpj[idx] = lpj_fn(batch, unpack(self._Kp[idx], H)).to(self.lpj.dtype)
            new_lpj = lpj_fn(batch, unpack(new_K, H)).to(self.lpj.dtype)
        select_topS_packed(new_K, new_lpj.contiguous(), self._Kp, self.lpj, idx, H, self._nsubs_dev)
