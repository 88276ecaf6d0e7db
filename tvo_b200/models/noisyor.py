"""NoisyOR (mirror of tvo/models/noisyor.py:16-191) on sm_100a kernels.

Same constructor, `theta` keys (pies, W), `log_pseudo_joint` (uint8 data and states required,
noisyor.py:71), `log_joint`, `update_param_batch`, `update_param_epoch`, `free_energy`,
`generate_data`; accumulators `new_pi`, `Btilde`, `Ctilde`, `_train_datapoints` are views of ONE packed
fp64 buffer [BtildeT | CtildeT | new_pi | N | F] that is all-reduced once per EM iteration (the
reference issues four, noisyor.py:141-149)."""
import ctypes as C
from typing import Dict, Optional, Tuple, Union

import torch as to
from torch import Tensor

import tvo_b200
from tvo_b200 import _lib
from tvo_b200.utils.model_protocols import Optimized, Sampler
from tvo_b200.utils.parallel import all_reduce, broadcast
from tvo_b200.variational._packed import as_packed


class NoisyOR(Optimized, Sampler):
    eps = 1e-7

    def __init__(self, H: int, D: int, W_init: Tensor = None, pi_init: Tensor = None, precision: to.dtype = to.float64):
        assert precision in (to.float32, to.float64), "precision must be one of torch.float{32,64}"
        self._precision = precision
        device = tvo_b200.get_device()
        if W_init is not None:
            assert W_init.shape == (D, H)
        else:
            W_init = to.rand(D, H, device=device)
            broadcast(W_init)
        if pi_init is not None:
            assert pi_init.shape == (H,)
            assert (pi_init <= 1.0).all() and (pi_init >= 0).all()
        else:
            pi_init = to.full((H,), 1.0 / H, device=device, dtype=precision)
        self._theta = {"pies": pi_init.to(device=device, dtype=precision).clone(),
                       "W": W_init.to(device=device, dtype=precision).clone()}
        self._stats = to.zeros(2 * H * D + H + 2, dtype=to.float64, device=device)
        self._config = dict(H=H, D=D, precision=precision, device=device)
        self._shape = (D, H)
        self._ws = None
        self._last_F = None

    # accumulators under the reference's names -----------------------------------------------------
    @property
    def Btilde(self) -> Tensor:
        D, H = self._shape
        return self._stats[: H * D].view(H, D).t()

    @property
    def Ctilde(self) -> Tensor:
        D, H = self._shape
        return self._stats[H * D: 2 * H * D].view(H, D).t()

    @property
    def new_pi(self) -> Tensor:
        D, H = self._shape
        return self._stats[2 * H * D: 2 * H * D + H]

    @property
    def _train_datapoints(self) -> Tensor:
        D, H = self._shape
        return self._stats[2 * H * D + H: 2 * H * D + H + 1]

    # kernels ---------------------------------------------------------------------------------------
    def _prepare(self):
        D, H = self._shape
        prec = self._precision
        W, pies = (self._theta[k].to(dtype=prec, device=self._stats.device).contiguous() for k in ("W", "pies"))
        nbytes = _lib.lib().tvo_noisyor_theta_ws_bytes(H, D, int(prec == to.float64))
        if self._ws is None or self._ws.numel() != nbytes or self._ws.device != W.device:
            self._ws = to.empty(nbytes, dtype=to.uint8, device=W.device)
        _lib.call(f"tvo_noisyor_prepare_{_lib.sfx(prec)}", _lib.ptr(W), _lib.ptr(pies), H, D, _lib.ptr(self._ws),
                  _lib.stream())
        return self._ws

    def _batch(self, data: Tensor) -> Tensor:
        assert data.dtype == to.uint8, "NoisyOR data must be uint8 (noisyor.py:71)"
        data = data.to(device=self._stats.device)
        return data if data.dim() == 2 and data.stride(-1) == 1 else data.contiguous()

    def _tvo_sparsity(self) -> Tensor:
        return self._theta["pies"].mean().to(to.float64).reshape(1)

    def _tvo_lpj_packed(self, batch: Tensor, packed: Tensor, H: int, k_broadcast: bool = False) -> Tensor:
        D = self._shape[0]
        Y = self._batch(batch)
        B, S, W = Y.shape[0], packed.shape[1], packed.shape[2]
        packed = packed.contiguous()
        out = to.empty((B, S), dtype=self._precision, device=Y.device)
        _lib.call(f"tvo_noisyor_lpj_{_lib.sfx(self._precision)}", _lib.ptr(Y), Y.stride(0), _lib.ptr(packed),
                  0 if k_broadcast else S * W, None, _lib.ptr(self._prepare()), _lib.ptr(out), S, B, S, H, D,
                  _lib.stream())
        return out

    def _tvo_lpj_rows(self, states, idx: Tensor, batch: Tensor) -> None:
        D = self._shape[0]
        Y = self._batch(batch)
        Kp = states._Kp
        S, W = Kp.shape[1], Kp.shape[2]
        assert states.lpj.dtype == self._precision and states.lpj.is_contiguous()
        _lib.call(f"tvo_noisyor_lpj_{_lib.sfx(self._precision)}", _lib.ptr(Y), Y.stride(0), _lib.ptr(Kp),
                  0 if Kp.shape[0] == 1 and states.lpj.shape[0] != 1 else S * W, _lib.ptr(idx),
                  _lib.ptr(self._prepare()), _lib.ptr(states.lpj), states.lpj.shape[1], idx.numel(), S, states._H, D,
                  _lib.stream())

    def _tvo_estep_evo(self, states, idx: Tensor, batch: Tensor, cfg) -> bool:
        D, H = self._shape
        if states.lpj.dtype != self._precision:
            return False
        if cfg.parent_selection == _lib.SEL["batch_fitparents"] and not (cfg.fit_norm and cfg.flags & _lib.EVO_LPJ_CURRENT):
            return False  # batch-global normaliser: fused for one generation (EVOVariationalStates.update_device)
        Y = self._batch(batch)
        sparsity = self._tvo_sparsity() if cfg.mutation in (_lib.MUT["sparseflip"], _lib.MUT["cross_sparseflip"]) else None
        _lib.call(f"tvo_estep_evo_noisyor_{_lib.sfx(self._precision)}", _lib.ptr(Y), Y.stride(0), _lib.ptr(idx),
                  _lib.ptr(states._Kp), _lib.ptr(states.lpj), _lib.ptr(self._prepare()), C.byref(cfg), _lib.ptr(sparsity),
                  idx.numel(), states._Kp.shape[1], H, D, _lib.ptr(states._nsubs_dev), _lib.stream())
        return True

    # reference API ---------------------------------------------------------------------------------
    def log_pseudo_joint(self, data: Tensor, states) -> Tensor:  # type: ignore
        """noisyor.py:67-106 (the reference's assert on uint8 inputs is kept)."""
        assert data.dtype == to.uint8
        if isinstance(states, Tensor):
            assert states.dtype == to.uint8
        packed, H, bc = as_packed(states)
        lpj = self._tvo_lpj_packed(data, packed, H, k_broadcast=bc)
        return lpj

    def log_joint(self, data, states, lpj=None):
        """noisyor.py:157-162"""
        if lpj is None:
            lpj = self.log_pseudo_joint(data, states)
        return to.sum(to.log(1 - self.theta["pies"])) + lpj

    def _free_energy_device(self, idx: Tensor, batch: Tensor, states, out: Tensor) -> None:
        idx = idx.to(device=out.device, dtype=to.int64).contiguous()
        lpj = states.lpj
        _lib.call(f"tvo_logsumexp_sum_{_lib.sfx(lpj.dtype)}", _lib.ptr(lpj), lpj.shape[1], _lib.ptr(idx), idx.numel(),
                  lpj.shape[1], _lib.ptr(out), _lib.stream())
        out += idx.numel() * to.log(1 - self.theta["pies"]).sum().to(to.float64)

    def free_energy(self, idx: Tensor, batch: Tensor, states) -> float:
        out = to.zeros((), dtype=to.float64, device=self._stats.device)
        self._free_energy_device(idx, batch, states, out)
        return out.item()

    def update_param_batch(self, idx: Tensor, batch: Tensor, states, mstep_factors: Dict[str, Tensor] = None
                           ) -> Optional[float]:
        """noisyor.py:108-138 -> tvo_noisyor_mstep"""
        D, H = self._shape
        Y = self._batch(batch)
        idx = idx.to(device=Y.device, dtype=to.int64).contiguous()
        Kp, lpj = states._Kp, states.lpj
        S, W = Kp.shape[1], Kp.shape[2]
        assert lpj.dtype == self._precision and lpj.is_contiguous()
        _lib.call(f"tvo_noisyor_mstep_{_lib.sfx(self._precision)}", _lib.ptr(Y), Y.stride(0), _lib.ptr(idx), _lib.ptr(Kp),
                  0 if Kp.shape[0] == 1 and lpj.shape[0] != 1 else S * W, _lib.ptr(lpj), lpj.shape[1],
                  _lib.ptr(self._prepare()), _lib.ptr(self._stats), idx.numel(), S, H, D, _lib.stream())
        return None

    def update_param_epoch(self) -> None:
        """noisyor.py:140-155, one packed all-reduce."""
        all_reduce(self._stats)
        self._last_F = self._stats[-1].clone()
        N = self._train_datapoints
        self._last_N = N.clone()
        prec, eps = self._precision, self.eps
        pies = to.clamp((self.new_pi / N).to(prec), eps, 1 - eps)
        Wn = to.clamp((1 + self.Btilde / (self.Ctilde + eps)).to(prec), eps, 1 - eps)
        flat = to.cat([pies.reshape(-1), Wn.reshape(-1)])
        broadcast(flat)
        self.theta["pies"][:] = flat[: pies.numel()]
        self.theta["W"][:] = flat[pies.numel():].view_as(self.theta["W"])
        self._stats.zero_()

    def generate_data(self, N: int = None, hidden_state: Tensor = None) -> Union[Tensor, Tuple[Tensor, Tensor]]:
        """noisyor.py:164-191"""
        W, pies = self.theta["W"], self.theta["pies"]
        if hidden_state is None:
            hidden_state = to.rand((N, W.shape[1]), dtype=pies.dtype, device=pies.device) < pies
            must_return_hidden_state = True
        else:
            if N is not None:
                assert hidden_state.shape == (N, W.shape[1])
            must_return_hidden_state = False
        py = 1 - to.prod(1 - W[None, :, :] * hidden_state.to(W.device).type_as(W)[:, None, :], dim=2)
        Y = (to.rand_like(py) < py).byte()
        return (Y, hidden_state) if must_return_hidden_state else Y
