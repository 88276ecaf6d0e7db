"""Single-cause mixture models (mirrors of tvo/models/gmm.py:20-221 and tvo/models/pmm.py:20-198) on sm_100a
kernels, for use with FullEMSingleCauseModels (tvo/variational/fullem.py:61-82).

Same constructors, `theta` keys (GMM: pies, W, sigma2; PMM: pies, W), `log_pseudo_joint`, `log_joint`,
`update_param_batch`, `update_param_epoch`, `free_energy`, `generate_data`, `data_estimator`.  The reference's
accumulators `my_Wp`, `my_Wq`, `my_pies`, `my_sigma2`, `my_N` are views of ONE packed fp64 buffer
[Wp^T | sum_n q | sigma2 | N | F] that is all-reduced once per EM iteration (the reference issues 4-5,
gmm.py:137-141).

With the single-cause state set (K_n = identity) every call is one kernel of csrc/mixture.cu: the dense
(N x D) x (D x H) log-joint contraction, and the fused posterior-weight / sufficient-statistic pass followed by
one DGEMM.  Arbitrary state tensors (the reference formula is written for any K) take the torch expression of
the reference on the device — a compatibility path, not the hot path.
"""
import math
from typing import Tuple, Union

import torch as to
from torch import Tensor
from torch.distributions.one_hot_categorical import OneHotCategorical

import tvo_b200
from tvo_b200 import _lib
from tvo_b200.utils.model_protocols import Optimized, Reconstructor, Sampler
from tvo_b200.utils.parallel import all_reduce, broadcast, pprint
from tvo_b200.utils.sanity import fix_theta
from tvo_b200.variational._utils import mean_posterior


def _is_single_cause(states) -> bool:
    return getattr(states, "single_cause", False)


class _SingleCauseMixture(Optimized, Sampler, Reconstructor):
    KIND = -1  # 0 GMM, 1 PMM (csrc/mixture.cu)

    def _init_common(self, H: int, D: int, precision: to.dtype, device):
        self._precision = precision
        self._stats = to.zeros(H * D + H + 3, dtype=to.float64, device=device)
        self._config = dict(H=H, D=D, precision=precision, device=device)
        self._shape = (D, H)
        self._ws = None
        self._scratch = None
        self._last_F = None
        self.scratch_datapoints = 131072

    # accumulators under the reference's names (views into the packed buffer) ---------------------
    @property
    def my_Wp(self) -> Tensor:
        D, H = self._shape
        return self._stats[: H * D].view(H, D).t()

    @property
    def my_Wq(self) -> Tensor:
        D, H = self._shape
        return self._stats[H * D: H * D + H]

    my_pies = my_Wq  # the same sum_n <s>_n (gmm.py:126,128)

    @property
    def my_sigma2(self) -> Tensor:
        D, H = self._shape
        return self._stats[H * D + H: H * D + H + 1]

    @property
    def my_N(self) -> Tensor:
        D, H = self._shape
        return self._stats[H * D + H + 1: H * D + H + 2]

    @property
    def shape(self) -> Tuple[int, ...]:
        return self.theta["W"].shape

    # kernels ---------------------------------------------------------------------------------------
    def _prepare(self):
        D, H = self._shape
        th, prec, dev = self._theta, self._precision, self._stats.device
        W = th["W"].to(device=dev, dtype=prec).contiguous()
        pies = th["pies"].to(device=dev, dtype=prec).contiguous()
        s2 = th["sigma2"].to(device=dev, dtype=prec).contiguous() if "sigma2" in th else None
        nbytes = _lib.lib().tvo_mixture_ws_bytes(H, D, int(prec == to.float64))
        if self._ws is None or self._ws.numel() != nbytes or self._ws.device != dev:
            self._ws = to.empty(nbytes, dtype=to.uint8, device=dev)
        _lib.call(f"tvo_mixture_prepare_{_lib.sfx(prec)}", self.KIND, _lib.ptr(W), _lib.ptr(pies), _lib.ptr(s2), H, D,
                  _lib.ptr(self._ws), _lib.stream())
        return self._ws

    def _batch(self, data: Tensor) -> Tensor:
        data = data.to(device=self._stats.device, dtype=self._precision)
        return data if data.dim() == 2 and data.stride(-1) == 1 else data.contiguous()

    def _lpj_all_causes(self, batch: Tensor, out: Tensor, idx: Tensor = None) -> None:
        """out[idx or arange] = lpj of all H causes for every datapoint of the batch."""
        D, H = self._shape
        Y = self._batch(batch)
        assert out.dtype == self._precision and out.stride(-1) == 1 and out.shape[1] == H
        _lib.call(f"tvo_mixture_lpj_{_lib.sfx(self._precision)}", self.KIND, _lib.ptr(Y), Y.stride(0), _lib.ptr(idx),
                  _lib.ptr(self._prepare()), _lib.ptr(out), out.stride(0), Y.shape[0], H, D, _lib.stream())

    def _supports_fast(self, states) -> bool:
        """The dense kernels evaluate all H single-cause states at once: only FullEMSingleCauseModels qualifies."""
        return _is_single_cause(states)

    def _tvo_lpj_rows(self, states, idx: Tensor, batch: Tensor) -> None:
        """states.lpj[idx] = lpj(batch, identity) in place (FullEMSingleCauseModels.update)."""
        assert _is_single_cause(states)
        self._lpj_all_causes(batch, states.lpj, idx.to(device=states.lpj.device, dtype=to.int64).contiguous())

    def _free_energy_device(self, idx: Tensor, batch: Tensor, states, out: Tensor) -> None:
        D, H = self._shape
        Y = self._batch(batch)
        idx = idx.to(device=Y.device, dtype=to.int64).contiguous()
        lpj = states.lpj
        assert lpj.dtype == self._precision and lpj.is_contiguous() and lpj.shape[1] == H
        _lib.call(f"tvo_mixture_free_energy_{_lib.sfx(self._precision)}", self.KIND, _lib.ptr(Y), Y.stride(0), _lib.ptr(idx),
                  _lib.ptr(lpj), lpj.shape[1], _lib.ptr(self._prepare()), idx.numel(), H, D, _lib.ptr(out), _lib.stream())

    # reference API -----------------------------------------------------------------------------------
    def log_pseudo_joint(self, data: Tensor, states) -> Tensor:  # type: ignore
        """gmm.py:91-102 / pmm.py:76-91.  `states`: SingleCauseK (fast path) or any (N,S,H) tensor."""
        if _is_single_cause(states):
            out = to.empty((data.shape[0], self._shape[1]), dtype=self._precision, device=self._stats.device)
            self._lpj_all_causes(data, out)
            return out
        return self._lpj_generic(data.to(self._stats.device), states.to(self._stats.device))

    def free_energy(self, idx: Tensor, batch: Tensor, states) -> float:
        """model_protocols.py:179-193"""
        if _is_single_cause(states):
            out = to.zeros((), dtype=to.float64, device=self._stats.device)
            self._free_energy_device(idx, batch, states, out)
            return out.item()
        return super().free_energy(idx, batch, states)

    def update_param_batch(self, idx: Tensor, batch: Tensor, states) -> None:
        """gmm.py:111-131 / pmm.py:99-113 -> tvo_mixture_mstep (F rides along in the packed statistics)."""
        if not _is_single_cause(states):
            return self._update_param_batch_generic(idx, batch, states)
        D, H = self._shape
        Y = self._batch(batch)
        idx = idx.to(device=Y.device, dtype=to.int64).contiguous()
        lpj = states.lpj
        assert lpj.dtype == self._precision and lpj.is_contiguous()
        n = min(max(idx.numel(), 1), int(self.scratch_datapoints))
        nbytes = _lib.lib().tvo_mixture_scratch_bytes(n, H, D, int(self._precision == to.float64))
        if self._scratch is None or self._scratch.numel() < nbytes:
            self._scratch = to.empty(nbytes, dtype=to.uint8, device=Y.device)
        _lib.call(f"tvo_mixture_mstep_{_lib.sfx(self._precision)}", self.KIND, _lib.ptr(Y), Y.stride(0), _lib.ptr(idx),
                  _lib.ptr(lpj), lpj.shape[1], _lib.ptr(self._prepare()), _lib.ptr(self._stats), _lib.ptr(self._scratch),
                  self._scratch.numel(), idx.numel(), H, D, _lib.stream())
        return None

    def _update_param_batch_generic(self, idx, batch, states) -> None:
        lpj = states.lpj[idx]
        K = states.K[idx]
        Kf = K.to(dtype=lpj.dtype)
        s_pjc = mean_posterior(Kf, lpj).to(to.float64)
        batch = batch.to(device=lpj.device)
        D, H = self._shape
        self._stats[: H * D].view(H, D).add_(s_pjc.t() @ batch.to(to.float64))
        self.my_Wq.add_(s_pjc.sum(dim=0))
        if "sigma2" in self._theta:
            Wbar = Kf @ self.theta["W"].t()
            self.my_sigma2.add_(mean_posterior(((batch[:, None, :] - Wbar) ** 2).sum(dim=2), lpj).sum().to(to.float64))
        self.my_N.add_(float(K.shape[0]))
        self._stats[-1] += super().free_energy(idx, batch, states)

    def update_param_epoch(self) -> None:
        """gmm.py:133-180 / pmm.py:115-153: one packed all-reduce, then the reference's closed forms."""
        theta, policy, prec = self.theta, self.policy, self._precision
        D, H = self._shape
        all_reduce(self._stats)
        self._last_F = self._stats[-1].clone()
        N = self.my_N
        self._last_N = N.clone()
        Wold_noisy = theta["W"] + 0.1 * to.randn_like(theta["W"])
        broadcast(Wold_noisy)
        theta_new = {"W": (self.my_Wp / self.my_Wq[None, :]).to(prec), "pies": (self.my_pies / N).to(prec)}
        policy["W"][0] = Wold_noisy
        policy["pies"][0] = theta["pies"]
        if "sigma2" in theta:
            theta_new["sigma2"] = (self.my_sigma2 / N / D).to(prec)
            policy["sigma2"][0] = theta["sigma2"]
        fix_theta(theta_new, policy)
        for key in theta:
            theta[key][:] = theta_new[key]  # in place: users keep their references to theta tensors
        self._stats.zero_()

    def data_estimator(self, idx: Tensor, batch: Tensor, states) -> Tensor:
        """gmm.py:209-221 / pmm.py:186-198: <W s>_q = W <s>_q"""
        lpj = states.lpj[idx]
        if _is_single_cause(states):
            q = to.softmax(lpj.to(to.float64), dim=1).to(self._precision)  # <s>_q for K = identity
        else:
            q = mean_posterior(states.K[idx].to(dtype=self._precision), lpj)
        return q @ self.theta["W"].t()


class GMM(_SingleCauseMixture):
    KIND = 0

    def __init__(self, H: int, D: int, W_init: Tensor = None, sigma2_init: Tensor = None, pies_init: Tensor = None,
                 precision: to.dtype = to.float64):
        assert precision in (to.float32, to.float64), "precision must be one of torch.float{32,64}"
        device = tvo_b200.get_device()
        if W_init is not None:
            assert W_init.shape == (D, H)
            W_init = W_init.to(dtype=precision, device=device).clone()
        else:
            W_init = to.rand((D, H), dtype=precision, device=device)
            broadcast(W_init)
        if pies_init is not None:
            assert pies_init.shape == (H,)
            pies_init = pies_init.to(dtype=precision, device=device).clone()
        else:
            pies_init = to.full((H,), 1.0 / H, dtype=precision, device=device)
        if sigma2_init is not None:
            assert sigma2_init.shape == (1,)
            sigma2_init = sigma2_init.to(dtype=precision, device=device).clone()
        else:
            sigma2_init = to.tensor([1.0], dtype=precision, device=device)
        self._theta = {"pies": pies_init, "W": W_init, "sigma2": sigma2_init}
        eps, inf = 1.0e-5, math.inf
        self.policy = {
            "W": [None, to.full_like(W_init, -inf), to.full_like(W_init, inf)],
            "pies": [None, to.full_like(pies_init, eps), to.full_like(pies_init, 1.0 - eps)],
            "sigma2": [None, to.full_like(sigma2_init, eps), to.full_like(sigma2_init, inf)],
        }
        self._init_common(H, D, precision, device)

    def _lpj_generic(self, data: Tensor, states: Tensor) -> Tensor:  # gmm.py:91-102
        Kf = states.to(dtype=self.theta["W"].dtype)
        Wbar = Kf @ self.theta["W"].t()
        return ((Wbar - data[:, None, :]) ** 2).sum(dim=2) * (-1 / 2 / self.theta["sigma2"]) + Kf @ to.log(self.theta["pies"])

    def log_joint(self, data: Tensor, states, lpj: Tensor = None) -> Tensor:
        """gmm.py:104-109"""
        if lpj is None:
            lpj = self.log_pseudo_joint(data, states)
        return lpj - self.shape[0] / 2 * to.log(2 * math.pi * self.theta["sigma2"])

    def generate_data(self, N: int = None, hidden_state: Tensor = None) -> Union[Tensor, Tuple[Tensor, Tensor]]:
        """gmm.py:186-207"""
        precision, device = self.precision, self._stats.device
        D, H = self.shape
        if hidden_state is None:
            assert N is not None
            pies = self.theta["pies"]
            hidden_state = OneHotCategorical(probs=pies).sample([N]) == 1
            must_return_hidden_state = True
        else:
            N = hidden_state.shape[0] if N is None else N
            assert hidden_state.shape == (N, H), f"hidden_state has shape {hidden_state.shape}, expected ({N},{H})"
            must_return_hidden_state = False
        Wbar = hidden_state.to(device=device, dtype=precision) @ self.theta["W"].t()
        Y = Wbar + to.sqrt(self.theta["sigma2"]) * to.randn((N, D), dtype=precision, device=device)
        return (Y, hidden_state) if must_return_hidden_state else Y


class PMM(_SingleCauseMixture):
    KIND = 1

    def __init__(self, H: int, D: int, W_init: Tensor = None, pies_init: Tensor = None, precision: to.dtype = to.float64):
        assert precision in (to.float32, to.float64), "precision must be one of torch.float{32,64}"
        device = tvo_b200.get_device()
        if W_init is not None:
            assert W_init.shape == (D, H)
            W_init = W_init.to(dtype=precision, device=device).clone()
        else:
            W_init = to.rand((D, H), dtype=precision, device=device)
            broadcast(W_init)
        if pies_init is not None:
            assert pies_init.shape == (H,)
            pies_init = pies_init.to(dtype=precision, device=device).clone()
        else:
            pies_init = to.full((H,), 1.0 / H, dtype=precision, device=device)
        self._theta = {"pies": pies_init, "W": W_init}
        eps, inf = 1.0e-5, math.inf
        self.policy = {
            "W": [None, to.full_like(W_init, eps), to.full_like(W_init, inf)],
            "pies": [None, to.full_like(pies_init, eps), to.full_like(pies_init, 1.0 - eps)],
        }
        self._init_common(H, D, precision, device)

    def _lpj_generic(self, data: Tensor, states: Tensor) -> Tensor:  # pmm.py:76-91
        Kf = states.to(dtype=self.theta["W"].dtype)
        Wbar = Kf @ self.theta["W"].t() + to.finfo(to.float32).tiny
        return (data[:, None, :] * to.log(Wbar)).sum(dim=2) - Wbar.sum(dim=2) + Kf @ to.log(self.theta["pies"])

    def log_joint(self, data: Tensor, states, lpj: Tensor = None) -> Tensor:
        """pmm.py:93-97"""
        if lpj is None:
            lpj = self.log_pseudo_joint(data, states)
        return lpj - to.lgamma(data.to(lpj.device) + 1).sum(dim=1)[:, None]

    def generate_data(self, N: int = None, hidden_state: Tensor = None) -> Union[Tensor, Tuple[Tensor, Tensor]]:
        """pmm.py:159-184"""
        precision, device = self.precision, self._stats.device
        D, H = self.shape
        if hidden_state is None:
            assert N is not None
            hidden_state = OneHotCategorical(probs=self.theta["pies"]).sample([N]) == 1
            must_return_hidden_state = True
        else:
            N = hidden_state.shape[0] if N is None else N
            assert hidden_state.shape == (N, H), f"hidden_state has shape {hidden_state.shape}, expected ({N},{H})"
            must_return_hidden_state = False
        Wbar = hidden_state.to(device=device, dtype=precision) @ self.theta["W"].t()
        Y = to.poisson(Wbar)
        return (Y, hidden_state) if must_return_hidden_state else Y
