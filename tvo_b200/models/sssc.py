"""Spike-and-Slab Sparse Coding (mirror of tvo/models/sssc.py:19-599) on sm_100a kernels.

Same constructor (W_init, sigma2_init, mus_init, Psi_init, pies_init, reformulated_lpj, use_storage,
reformulated_psi_update, precision), `theta` keys (W, sigma2, mus, Psi, pies), `log_pseudo_joint`,
`log_joint`, `update_param_batch`, `update_param_epoch`, `free_energy`, `generate_data`,
`data_estimator`.  `reformulated_lpj` / `use_storage` are accepted for compatibility: the kernel always
evaluates the (consistent) reformulated form (the straightforward form only differs by the state-
independent -D/2 log sigma2, quirk Q7) and needs no hash-keyed storage.

Sufficient statistics live in ONE packed fp64 buffer (one all-reduce per EM iteration instead of the
reference's 8-9, sssc.py:483-514); the reference's `_my_sum_*` names are views of it."""
import ctypes as C
from math import pi as MATH_PI
from typing import Any, Dict, Tuple, Union

import torch as to

import tvo_b200
from tvo_b200 import _lib
from tvo_b200.utils.model_protocols import Optimized, Reconstructor, Sampler
from tvo_b200.utils.parallel import all_reduce, broadcast, pprint
from tvo_b200.variational._packed import as_packed
from tvo_b200.variational._utils import mean_posterior


class SSSC(Sampler, Optimized, Reconstructor):
    def __init__(self, H: int, D: int, W_init: to.Tensor = None, sigma2_init: to.Tensor = None,
                 mus_init: to.Tensor = None, Psi_init: to.Tensor = None, pies_init: to.Tensor = None,
                 reformulated_lpj: bool = True, use_storage: bool = True, reformulated_psi_update: bool = False,
                 precision: to.dtype = to.float32):
        assert precision in (to.float32, to.float64), "precision must be one of torch.float{32,64}"
        device = tvo_b200.get_device()
        self._precision = precision
        self._shape = (D, H)
        self._reformulated_lpj = reformulated_lpj
        self._use_storage = use_storage
        self._reformulated_psi_update = reformulated_psi_update

        def init(t, shape, default):
            if t is not None:
                assert tuple(t.shape) == shape
                return t.to(dtype=precision, device=device).clone()
            v = default()
            return v

        def rand_W():
            w = to.rand((D, H), dtype=precision, device=device)
            broadcast(w)
            return w

        def rand_mus():
            m = to.normal(mean=to.zeros(H, dtype=precision, device=device), std=to.ones(H, dtype=precision, device=device))
            broadcast(m)
            return m

        self._theta: Dict[str, to.Tensor] = {}
        self._theta["W"] = init(W_init, (D, H), rand_W)
        self._theta["sigma2"] = init(sigma2_init, (1,), lambda: to.tensor([1.0], dtype=precision, device=device))
        self._theta["mus"] = init(mus_init, (H,), rand_mus)
        self._theta["Psi"] = init(Psi_init, (H, H), lambda: to.eye(H, dtype=precision, device=device))
        self._theta["pies"] = init(pies_init, (H,), lambda: 0.1 + 0.5 * to.rand(H, dtype=precision, device=device))
        self._stats = to.zeros(H * D + 4 * H * H + 2 * H + D + 2, dtype=to.float64, device=device)
        self._overflow = to.zeros(1, dtype=to.int32, device=device)
        self._eps_eyeH = to.eye(H, dtype=precision, device=device) * 1e-6
        self._config = dict(shape=self._shape, reformulated_lpj=reformulated_lpj,
                            reformulated_psi_update=reformulated_psi_update, use_storage=use_storage,
                            precision=precision, device=device)
        self.scratch_datapoints = 65536
        self._scratch = None
        self._ws = None
        self._last_F = None

    # --- packed statistics under the reference's names ---------------------------------------------
    def _seg(self, name):
        D, H = self._shape
        HD, HH = H * D, H * H
        o = {"y_szT": (0, HD), "sz_szT": (HD, HH), "szszT": (HD + HH, HH), "ssT": (HD + 2 * HH, HH),
             "ssz": (HD + 3 * HH, HH), "s": (HD + 4 * HH, H), "sz": (HD + 4 * HH + H, H),
             "yy": (HD + 4 * HH + 2 * H, D), "N": (HD + 4 * HH + 2 * H + D, 1)}[name]
        return self._stats[o[0]: o[0] + o[1]]

    @property
    def _my_sum_y_szT(self):
        D, H = self._shape
        return self._seg("y_szT").view(H, D).t()

    @property
    def _my_sum_xpt_sz_xpt_szT(self):
        H = self._shape[1]
        return self._seg("sz_szT").view(H, H)

    @property
    def _my_sum_xpt_szszT(self):
        H = self._shape[1]
        return self._seg("szszT").view(H, H)

    @property
    def _my_sum_xpt_ssT(self):
        H = self._shape[1]
        U = self._seg("ssT").view(H, H)  # the kernel accumulates the upper triangle
        return U + U.t() - to.diag(to.diagonal(U))

    @property
    def _my_sum_xpt_ssz(self):
        H = self._shape[1]
        return self._seg("ssz").view(H, H) if self._reformulated_psi_update else None

    @property
    def _my_sum_xpt_s(self):
        return self._seg("s")

    @property
    def _my_sum_xpt_sz(self):
        return self._seg("sz")

    @property
    def _my_sum_diag_yyT(self):
        return self._seg("yy")

    @property
    def _my_N(self):
        return self._seg("N")

    # --- kernels -----------------------------------------------------------------------------------
    def _prepare(self):
        D, H = self._shape
        prec = self._precision
        th = {k: self._theta[k].to(dtype=prec, device=self._stats.device).contiguous() for k in self._theta}
        nbytes = _lib.lib().tvo_sssc_theta_ws_bytes(H, D, int(prec == to.float64))
        if self._ws is None or self._ws.numel() != nbytes:
            self._ws = to.empty(nbytes, dtype=to.uint8, device=self._stats.device)
        _lib.call(f"tvo_sssc_prepare_{_lib.sfx(prec)}", _lib.ptr(th["W"]), _lib.ptr(th["sigma2"]), _lib.ptr(th["mus"]),
                  _lib.ptr(th["Psi"]), _lib.ptr(th["pies"]), H, D, _lib.ptr(self._ws), _lib.stream())
        return self._ws

    def _get_scratch(self, B):
        D, H = self._shape
        n = min(max(int(B), 1), int(self.scratch_datapoints))
        nbytes = _lib.lib().tvo_sssc_scratch_bytes(n, H, D, int(self._precision == to.float64))
        if self._scratch is None or self._scratch.numel() < nbytes:
            self._scratch = to.empty(nbytes, dtype=to.uint8, device=self._stats.device)
        return self._scratch

    def _batch(self, data):
        data = data.to(device=self._stats.device, dtype=self._precision)
        return data if data.dim() == 2 and data.stride(-1) == 1 else data.contiguous()

    def _check_overflow(self):
        if int(self._overflow.item()):
            self._overflow.zero_()
            raise RuntimeError("SSSC: a state with more active units than the kernel's per-state limit (32) was "
                               "evaluated; its lpj was set to finfo.min")

    def _tvo_sparsity(self):
        return self._theta["pies"].mean().to(to.float64).reshape(1)

    def _tvo_lpj_packed(self, batch, packed, H, k_broadcast=False):
        D = self._shape[0]
        Y = self._batch(batch)
        B, S, W = Y.shape[0], packed.shape[1], packed.shape[2]
        packed = packed.contiguous()
        out = to.empty((B, S), dtype=self._precision, device=Y.device)
        sc = self._get_scratch(B)
        _lib.call(f"tvo_sssc_lpj_{_lib.sfx(self._precision)}", _lib.ptr(Y), Y.stride(0), _lib.ptr(packed),
                  0 if k_broadcast else S * W, None, _lib.ptr(self._prepare()), _lib.ptr(out), S, _lib.ptr(sc), sc.numel(),
                  _lib.ptr(self._overflow), B, S, H, D, _lib.stream())
        return out

    def _tvo_lpj_rows(self, states, idx, batch):
        D = self._shape[0]
        Y = self._batch(batch)
        Kp = states._Kp
        S, W = Kp.shape[1], Kp.shape[2]
        assert states.lpj.dtype == self._precision and states.lpj.is_contiguous()
        sc = self._get_scratch(idx.numel())
        _lib.call(f"tvo_sssc_lpj_{_lib.sfx(self._precision)}", _lib.ptr(Y), Y.stride(0), _lib.ptr(Kp),
                  0 if Kp.shape[0] == 1 and states.lpj.shape[0] != 1 else S * W, _lib.ptr(idx), _lib.ptr(self._prepare()),
                  _lib.ptr(states.lpj), states.lpj.shape[1], _lib.ptr(sc), sc.numel(), _lib.ptr(self._overflow),
                  idx.numel(), S, states._H, D, _lib.stream())

    def _tvo_estep_evo(self, states, idx, batch, cfg) -> bool:
        D, H = self._shape
        if states.lpj.dtype != self._precision:
            return False
        if cfg.parent_selection == _lib.SEL["batch_fitparents"] and not (cfg.fit_norm and cfg.flags & _lib.EVO_LPJ_CURRENT):
            return False  # batch-global normaliser: fused for one generation (EVOVariationalStates.update_device)
        Y = self._batch(batch)
        sparsity = self._tvo_sparsity() if cfg.mutation in (_lib.MUT["sparseflip"], _lib.MUT["cross_sparseflip"]) else None
        sc = self._get_scratch(idx.numel())
        _lib.call(f"tvo_estep_evo_sssc_{_lib.sfx(self._precision)}", _lib.ptr(Y), Y.stride(0), _lib.ptr(idx),
                  _lib.ptr(states._Kp), _lib.ptr(states.lpj), _lib.ptr(self._prepare()), C.byref(cfg), _lib.ptr(sparsity),
                  _lib.ptr(sc), sc.numel(), _lib.ptr(self._overflow), idx.numel(), states._Kp.shape[1], H, D,
                  _lib.ptr(states._nsubs_dev), _lib.stream())
        return True

    # --- reference API -----------------------------------------------------------------------------
    def log_pseudo_joint(self, data: to.Tensor, states) -> to.Tensor:
        """sssc.py:339-346 (reformulated form; NaN/inf lpj -> finfo.min)."""
        packed, H, bc = as_packed(states)
        return self._tvo_lpj_packed(data, packed, H, k_broadcast=bc)

    def log_joint(self, data: to.Tensor, states, lpj=None) -> to.Tensor:
        """sssc.py:348-364"""
        if isinstance(states, to.Tensor):
            assert states.dtype == to.uint8
        notnan = to.logical_not(to.isnan(data))
        if lpj is None:
            lpj = self.log_pseudo_joint(data, states)
        pies = self.theta["pies"].clamp(1e-2, 1.0 - 1e-2)
        Dn = to.sum(notnan, dim=1).to(lpj.device)
        log2pi = to.log(to.tensor([2.0 * MATH_PI], dtype=self._precision, device=lpj.device))
        logjoints = lpj + to.log(1.0 - pies).sum() - Dn.unsqueeze(1) / 2.0 * (log2pi + to.log(self.theta["sigma2"]))
        assert logjoints.shape == lpj.shape
        assert not to.isnan(logjoints).any() and not to.isinf(logjoints).any()
        return logjoints

    def _free_energy_device(self, idx, batch, states, out) -> None:
        idx = idx.to(device=out.device, dtype=to.int64).contiguous()
        lpj = states.lpj
        _lib.call(f"tvo_logsumexp_sum_{_lib.sfx(lpj.dtype)}", _lib.ptr(lpj), lpj.shape[1], _lib.ptr(idx), idx.numel(),
                  lpj.shape[1], _lib.ptr(out), _lib.stream())
        data = batch.to(out.device)
        Dn = to.sum(to.logical_not(to.isnan(data)), dim=1).to(to.float64) if data.is_floating_point() else None
        pies = self.theta["pies"].clamp(1e-2, 1.0 - 1e-2).to(to.float64)
        c = to.log(1.0 - pies).sum() * idx.numel()
        c = c - Dn.sum() / 2.0 * (to.log(to.tensor(2.0 * MATH_PI, dtype=to.float64, device=out.device))
                                 + to.log(self.theta["sigma2"].to(to.float64)).squeeze())
        out += c

    def free_energy(self, idx, batch, states) -> float:
        out = to.zeros((), dtype=to.float64, device=self._stats.device)
        self._free_energy_device(idx, batch, states, out)
        return out.item()

    def update_param_batch(self, idx: to.Tensor, batch: to.Tensor, states, **kwargs: Dict[str, Any]) -> None:
        """sssc.py:366-452 -> tvo_sssc_mstep"""
        D, H = self._shape
        Y = self._batch(batch)
        idx = idx.to(device=Y.device, dtype=to.int64).contiguous()
        Kp, lpj = states._Kp, states.lpj
        S, W = Kp.shape[1], Kp.shape[2]
        assert lpj.dtype == self._precision and lpj.is_contiguous()
        sc = self._get_scratch(idx.numel())
        _lib.call(f"tvo_sssc_mstep_{_lib.sfx(self._precision)}", _lib.ptr(Y), Y.stride(0), _lib.ptr(idx), _lib.ptr(Kp),
                  0 if Kp.shape[0] == 1 and lpj.shape[0] != 1 else S * W, _lib.ptr(lpj), lpj.shape[1],
                  _lib.ptr(self._prepare()), _lib.ptr(self._stats), _lib.ptr(sc), sc.numel(), _lib.ptr(self._overflow),
                  idx.numel(), S, H, D, int(self._reformulated_psi_update) | (2 if kwargs.get("_nan_aware") else 0),
                  _lib.stream())
        return None

    def _invert_my_sum_xpt_ssT(self, ssT) -> to.Tensor:
        eps_eyeH = self._eps_eyeH.to(ssT.dtype)
        try:
            return to.linalg.inv(ssT)
        except Exception:
            try:
                pprint("Psi update: Addd diag(eps) before computing inverse")
                return to.linalg.inv(ssT + eps_eyeH)
            except Exception:
                pprint("Psi update: Added diag(eps) and computed pseudo-inverse")
                return to.linalg.pinv(ssT + eps_eyeH)

    def update_param_epoch(self) -> None:
        """sssc.py:467-540: one packed all-reduce, then the reference's closed-form updates in fp64."""
        self._check_overflow()
        theta, prec = self.theta, self._precision
        D, H = self.shape
        dtype_eps, eps = to.finfo(prec).eps, 1e-5
        all_reduce(self._stats)
        self._last_F = self._stats[-1].clone()
        N = self._my_N
        self._last_N = N.clone()
        y_szT, szszT, sz_szT = self._my_sum_y_szT, self._my_sum_xpt_szszT, self._my_sum_xpt_sz_xpt_szT
        s, sz, ssT = self._my_sum_xpt_s, self._my_sum_xpt_sz, self._my_sum_xpt_ssT
        eps_eye = self._eps_eyeH.to(to.float64)
        Inv_ssT = self._invert_my_sum_xpt_ssT(ssT)
        try:
            W = y_szT @ to.linalg.inv(szszT)
        except Exception:
            try:
                noise = eps * to.randn(H, dtype=to.float64, device=szszT.device)
                W = y_szT @ to.linalg.pinv(szszT + noise.unsqueeze(1) * noise.unsqueeze(0))
                pprint("W update: Used noisy pseudo-inverse")
            except Exception:
                W = theta["W"].to(to.float64) + eps * to.randn_like(theta["W"].to(to.float64))
                pprint("W update: Failed to compute W^(new). Pertubed current W with AWGN.")
        pies = s / N
        mus = sz / (s + dtype_eps)
        if self._reformulated_psi_update:
            _Psi = to.outer(mus, mus) * ssT + szszT - 2.0 * mus.unsqueeze(1) * self._my_sum_xpt_ssz
            Psi = _Psi * Inv_ssT + eps_eye
        else:
            Psi = (szszT - ssT * to.outer(mus, mus)) * Inv_ssT + eps_eye
        sigma2 = (self._my_sum_diag_yyT.sum() - to.trace(sz_szT @ (W.t() @ W))) / N / D + eps
        new = {"W": W, "sigma2": sigma2.reshape(1), "mus": mus, "Psi": Psi, "pies": pies}
        flat = to.cat([new[k].reshape(-1) for k in theta]).to(prec)
        broadcast(flat)
        o = 0
        for k in theta:
            n = theta[k].numel()
            theta[k][:] = flat[o:o + n].view_as(theta[k])
            o += n
        self._stats.zero_()

    @property
    def shape(self) -> Tuple[int, ...]:
        return self._shape

    def generate_data(self, N: int = None, hidden_state: to.Tensor = None) -> Union[to.Tensor, Tuple[to.Tensor, to.Tensor]]:
        """sssc.py:150-173"""
        precision, device = self.precision, self._stats.device
        D, H = self.shape
        if hidden_state is None:
            assert N is not None
            hidden_state = to.rand((N, H), dtype=precision, device=device) < self.theta["pies"]
            must_return_hidden_state = True
        else:
            if N is None:
                N = hidden_state.shape[0]
            assert hidden_state.shape == (N, H)
            must_return_hidden_state = False
        Z = to.distributions.multivariate_normal.MultivariateNormal(loc=self.theta["mus"],
                                                                    covariance_matrix=self.theta["Psi"])
        Wbar = to.einsum("dh,nh->nd", (self.theta["W"], hidden_state.to(device).to(precision) * Z.sample((N,))))
        Y = Wbar + to.sqrt(self.theta["sigma2"]) * to.randn((N, D), dtype=precision, device=device)
        return (Y, hidden_state) if must_return_hidden_state else Y

    def data_estimator(self, idx: to.Tensor, batch: to.Tensor, states) -> to.Tensor:
        """sssc.py:542-599: W <kappa>_q.  <kappa>_q is taken from a scratch M-step pass (statistics discarded).
        Datapoints with missing values (NaN) are handled like the reference (sssc.py:563-597): kappa of a state is
        computed from the observed dimensions only (the kernel's restricted path), not from zero-filled data."""
        D, H = self._shape
        saved = self._stats.clone()
        Y = self._batch(batch)
        idx = idx.to(device=Y.device, dtype=to.int64).contiguous()
        self.update_param_batch(idx, Y, states, _nan_aware=bool(to.isnan(Y).any()))
        self._stats.copy_(saved)
        sc = self._scratch
        B = idx.numel()
        cap = min(B, int(self.scratch_datapoints))
        assert B <= cap, "data_estimator: batch larger than scratch_datapoints"
        a_bytes = ((cap * H * (8 if self._precision == to.float64 else 4)) + 15) // 16 * 16
        X = sc[a_bytes: a_bytes + cap * H * 8].view(to.float64).view(cap, H)[:B]
        return (X @ self.theta["W"].to(to.float64).t()).to(self._precision)
