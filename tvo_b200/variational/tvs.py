"""Truncated Variational Sampling E-step (mirror of tvo/variational/tvs.py:12-81).

Same constructor, config keys and `update(idx, batch, model) -> int` as the reference.  Per batch:
re-score K (tvs.py:58), S_new_prior states with bits ~ Bernoulli(pies) (:60-63), S_new_marg states with
bits ~ Bernoulli(p(s_h=1|y_n)) where the marginals are the posterior mean of the states in K (:65-74),
dedup against K (:80) and top-S merge (:82).  The marginals come from the packed K
(tvo_state_marginals, no (B,S,H) float temporary) and the draws from the counter-based Philox stream
(seed, step, global datapoint index) instead of torch's global generator.
"""
from typing import Optional

import torch as to
from torch import Tensor

from tvo_b200 import _lib
from tvo_b200.utils.model_protocols import Optimized
from .TVOVariationalStates import TVOVariationalStates
from ._packed import unpack
from ._utils import mark_redundant_packed, select_topS_packed, state_marginals

GROUP_PRIOR, GROUP_MARG = 0, 1


class TVSVariationalStates(TVOVariationalStates):
    def __init__(self, N: int, H: int, S: int, precision: to.dtype, S_new_prior: int, S_new_marg: int,
                 K_init_file: str = None, seed: Optional[int] = None, K_init: Tensor = None):
        conf = dict(N=N, H=H, S=S, S_new_prior=S_new_prior, S_new_marg=S_new_marg, S_new=S_new_prior + S_new_marg,
                    precision=precision, K_init_file=K_init_file, seed=seed)
        super().__init__(conf, K_init=K_init)

    def update(self, idx: Tensor, batch: Tensor, model) -> int:
        self.update_device(idx, batch, model)
        return self._take_nsubs()

    def update_device(self, idx: Tensor, batch: Tensor, model) -> None:
        dev = self.lpj.device
        idx = idx.to(device=dev, dtype=to.int64).contiguous()
        B, H, W = idx.numel(), self._H, self._Kp.shape[2]
        Sp, Sm = self.config["S_new_prior"], self.config["S_new_marg"]
        fast = hasattr(model, "_tvo_lpj_packed")
        lpj_fn = None if fast else (model.log_pseudo_joint if isinstance(model, Optimized) else model.log_joint)
        if fast:
            model._tvo_lpj_rows(self, idx, batch)                                      # tvs.py:58
        else:
            self.lpj[idx] = lpj_fn(batch, unpack(self._Kp[idx], H)).to(self.lpj.dtype)
        new_K = to.empty((B, Sp + Sm, W), dtype=to.int64, device=dev)
        sfx = _lib.sfx(self.lpj.dtype)
        pies = model.theta["pies"].to(device=dev, dtype=self.lpj.dtype).expand(H).contiguous()
        step, noff = self._step & 0xFFFFFFFF, self.n_offset & 0xFFFFFFFF
        _lib.call(f"tvo_sample_states_probs_{sfx}", _lib.ptr(new_K), (Sp + Sm) * W, 0, _lib.ptr(pies), 0, _lib.ptr(idx),
                  B, Sp, H, self.seed, step, GROUP_PRIOR, noff, _lib.stream())       # tvs.py:60-63
        marg = state_marginals(self._Kp, self.lpj, idx, H)                             # tvs.py:65-70
        _lib.call(f"tvo_sample_states_probs_{sfx}", _lib.ptr(new_K), (Sp + Sm) * W, Sp, _lib.ptr(marg), H, _lib.ptr(idx),
                  B, Sm, H, self.seed, step, GROUP_MARG, noff, _lib.stream())        # tvs.py:71-74
        self._step += 1
        if fast:
            new_lpj = model._tvo_lpj_packed(batch, new_K, H)
        else:
            new_lpj = lpj_fn(batch, unpack(new_K, H)).to(self.lpj.dtype)
        new_lpj = new_lpj.contiguous()
        mark_redundant_packed(new_K, new_lpj, self._Kp, idx, H)                        # tvs.py:80
        select_topS_packed(new_K, new_lpj, self._Kp, self.lpj, idx, H, self._nsubs_dev)  # tvs.py:82
