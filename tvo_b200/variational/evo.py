"""Evolutionary variational E-step (mirror of tvo/variational/evo.py:20-115).

Same constructor, config keys and `update(idx, batch, model) -> int` as the reference.  What runs:

 * model is a tvo_b200 CUDA model (BSC / NoisyOR / SSSC) and the configuration allows it
   -> ONE fused kernel per batch (re-score K, G generations of select+mutate+score in shared
      memory, dedup, top-S merge); see csrc/bsc.cu:estep_evo_bsc_kernel.
 * otherwise (third-party model, batch-global fitparents, D > 256)
   -> the same device functions as stand-alone kernels around the model's own lpj function:
      tvo_evo_generation -> lpj_fn -> tvo_mark_redundant -> tvo_select_topS.

Candidates come from a counter-based Philox stream keyed by (seed, step, global datapoint index,
generation) instead of torch's global RNG (evo.py:263,325,363,414,440): results do not depend on
batch composition or rank count.  `fitparents_fixed` is an additional parent_selection with the
per-datapoint normaliser (the reference's `batch_fitparents` normalises by the whole batch, SURVEY Q3;
that behaviour is kept under its own name).
"""
import ctypes as C
from typing import Optional

import torch as to
from torch import Tensor

from tvo_b200 import _lib
from tvo_b200.utils import get
from tvo_b200.utils.model_protocols import Optimized
from .TVOVariationalStates import TVOVariationalStates
from ._packed import unpack
from ._utils import mark_redundant_packed, select_topS_packed


def get_n_new_states(mutation: str, n_parents: int, n_children: int, n_gen: int) -> int:
    """evo.py:214-218"""
    if mutation[:5] == "cross":
        return n_parents * (n_parents - 1) * n_gen
    return n_parents * n_children * n_gen


def check_EA(parent_selection: str, mutation: str):
    """evo.py:221-245 (ValueError for unknown operator names)."""
    if parent_selection not in _lib.SEL:
        raise ValueError(f'Parent selection "{parent_selection}" not supported. Valid options: {list(_lib.SEL)}')
    if mutation not in _lib.MUT:
        raise ValueError(f'Mutation operator "{mutation}" not supported. Valid options: {list(_lib.MUT)}')


class EVOVariationalStates(TVOVariationalStates):
    def __init__(
        self,
        N: int,
        H: int,
        S: int,
        precision: to.dtype,
        parent_selection: str,
        mutation: str,
        n_parents: int,
        n_generations: int,
        n_children: int = None,
        crossover: bool = False,
        bitflip_frequency: float = None,
        K_init_file: str = None,
        seed: Optional[int] = None,
        K_init: Tensor = None,
    ):
        assert not crossover or n_children is None, "Exactly one of n_children and crossover may be provided."
        if crossover:
            mutation = f"cross_{mutation}"
            n_children = n_parents - 1
        assert n_children is not None
        check_EA(parent_selection, mutation)
        S_new = get_n_new_states(mutation, n_parents, n_children, n_generations)
        conf = dict(
            N=N, H=H, S=S, S_new=S_new, precision=precision, parent_selection=parent_selection, mutation=mutation,
            n_parents=n_parents, n_children=n_children, n_generations=n_generations, p_bf=bitflip_frequency,
            K_init_file=K_init_file, seed=seed,
        )
        super().__init__(conf, K_init=K_init)
        self._fit_norm = to.zeros(2, dtype=to.float64, device=self.lpj.device)

    # ------------------------------------------------------------------------------------------
    def evo_config(self) -> _lib.EvoConfig:
        ps, mut, P, Cn, G = get(self.config, "parent_selection", "mutation", "n_parents", "n_children", "n_generations")
        p_bf = self.config.get("p_bf")
        if "sparseflip" in mut and p_bf is None:
            raise ValueError("bitflip_frequency is required by the sparseflip mutation")
        return _lib.EvoConfig(_lib.SEL[ps], _lib.MUT[mut], P, Cn, G, 0, float(p_bf or 0.0), self.seed,
                              self._step & 0xFFFFFFFF, self.n_offset & 0xFFFFFFFF)

    def update(self, idx: Tensor, batch: Tensor, model) -> int:
        self.update_device(idx, batch, model)
        return self._take_nsubs()

    def update_device(self, idx: Tensor, batch: Tensor, model) -> None:
        idx = idx.to(device=self.lpj.device, dtype=to.int64).contiguous()
        cfg = self.evo_config()
        self._step += 1
        if hasattr(model, "_tvo_estep_evo"):
            if cfg.parent_selection == _lib.SEL["batch_fitparents"] and cfg.n_generations == 1 \
                    and self.lpj.dtype == getattr(model, "precision", self.lpj.dtype):
                # the reference's fitness selection normalises by (min, sum) of the WHOLE batch's re-scored lpj
                # (evo.py:404-406): re-score, reduce (deterministic two-pass), then the fused kernel skips its own
                # re-score and selects with that normaliser.  Later generations would need the same reduction over
                # the previous generation's children of the whole batch: those run through the stand-alone kernels.
                model._tvo_lpj_rows(self, idx, batch)
                _lib.call(f"tvo_fit_norm_{_lib.sfx(self.lpj.dtype)}", _lib.ptr(self.lpj), self.lpj.shape[1], _lib.ptr(idx),
                          idx.numel(), self.lpj.shape[1], _lib.ptr(self._fit_norm), _lib.stream())
                cfg.flags |= _lib.EVO_LPJ_CURRENT
                cfg.fit_norm = self._fit_norm.data_ptr()
            if model._tvo_estep_evo(self, idx, batch, cfg):
                return
            lpj_rows = lambda: model._tvo_lpj_rows(self, idx, batch)  # noqa: E731
            lpj_children = lambda ch: model._tvo_lpj_packed(batch, ch, self._H)  # noqa: E731
            sparsity = model._tvo_sparsity()
        else:
            lpj_fn = model.log_pseudo_joint if isinstance(model, Optimized) else model.log_joint
            H = self._H

            def lpj_rows():
                self.lpj[idx] = lpj_fn(batch, unpack(self._Kp[idx], H)).to(self.lpj.dtype)

            def lpj_children(ch):
                return lpj_fn(batch, unpack(ch, H)).to(self.lpj.dtype)

            sparsity = None
            if "sparseflip" in self.config["mutation"]:
                sparsity = model.theta["pies"].mean().to(dtype=to.float64, device=self.lpj.device).reshape(1)
        self._evolve_unfused(idx, cfg, lpj_rows, lpj_children, sparsity)

    def _evolve_unfused(self, idx, cfg, lpj_rows, lpj_children, sparsity):
        """evolve_states (evo.py:118-211) + update_states_for_batch with stand-alone kernels."""
        B = idx.numel()
        N, S, W = self._Kp.shape
        H, G, S_new = self._H, cfg.n_generations, self.config["S_new"]
        Sgen = S_new // G
        sfx = _lib.sfx(self.lpj.dtype)
        dev = self.lpj.device
        lpj_rows()  # lpj[idx] = lpj_fn(batch, K[idx])  (evo.py:95)
        children = to.empty((B, S_new, W), dtype=to.int64, device=dev)
        new_lpj = to.empty((B, S_new), dtype=self.lpj.dtype, device=dev)
        gen_children = to.empty((B, Sgen, W), dtype=to.int64, device=dev)
        batch_fit = cfg.parent_selection == _lib.SEL["batch_fitparents"]
        for g in range(G):
            if g == 0:
                cand, cstride, clpj, lstride, indexed, S_c = self._Kp, S * W, self.lpj, S, 1, S
            else:
                cand, cstride = children[:, (g - 1) * Sgen:], S_new * W
                clpj, lstride, indexed, S_c = new_lpj[:, (g - 1) * Sgen:], S_new, 0, Sgen
            if batch_fit:
                _lib.call(f"tvo_fit_norm_{sfx}", _lib.ptr(clpj), lstride, _lib.ptr(idx) if indexed else None, B, S_c,
                          _lib.ptr(self._fit_norm), _lib.stream())
            _lib.call(f"tvo_evo_generation_{sfx}", _lib.ptr(cand), cstride, _lib.ptr(clpj), lstride, _lib.ptr(idx),
                      indexed, S_c, _lib.ptr(gen_children), C.byref(cfg), g, _lib.ptr(sparsity),
                      _lib.ptr(self._fit_norm) if batch_fit else None, B, H, _lib.stream())
            children[:, g * Sgen:(g + 1) * Sgen] = gen_children
            new_lpj[:, g * Sgen:(g + 1) * Sgen] = lpj_children(gen_children)
        mark_redundant_packed(children, new_lpj, self._Kp, idx, H)
        select_topS_packed(children, new_lpj, self._Kp, self.lpj, idx, H, self._nsubs_dev)
