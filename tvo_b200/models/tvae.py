"""Truncated Variational Autoencoder (mirror of tvo/models/tvae.py:85-701): `GaussianTVAE`, `BernoulliTVAE`.

Same constructors, theta keys (`W_i`, `b_i`, `pies`, `sigma2`), `log_joint` / `free_energy` /
`update_param_batch` / `update_param_epoch` / `generate_data` / `data_estimator` / `forward` as the reference.

What differs underneath: the variational states are bit-packed, and the reference's two dense products with the
uint8 states cast to float — `states @ W_0 + b_0` (tvae.py:434-459) and `states @ log(pi/(1-pi))` (tvae.py:337) —
are replaced by kernels that walk the set bits (`tvo_tvae_layer0`, `tvo_tvae_layer0_grad` for the weight
gradient, `tvo_tvae_lpj` for the log-pseudo-joint from the decoder output), so no (N, S, H0) float tensor is ever
built.  The hidden layers, autograd, Adam and the cyclic learning-rate schedule stay torch (plain library GEMMs on
(N*S, H_l) activations).  The E-step runs through the stand-alone candidate / dedup / top-S kernels with this
model's lpj in the middle (`EVOVariationalStates._evolve_unfused`).

`external_model` decoders and gradient-trained priors take the dense route (states unpacked on the device), like
any third-party model.
"""
from math import pi as MATH_PI
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

import torch as to
import torch.optim as opt
from torch import Tensor

import tvo_b200
from tvo_b200 import _lib
from tvo_b200.utils.parallel import all_reduce, broadcast
from tvo_b200.variational._packed import PackedStates, as_packed, unpack
from tvo_b200.variational._utils import state_marginals


def get_world_size() -> int:
    """Ranks taking part in the run (1 unless the run policy is 'mpi' with an initialised process group)."""
    import torch.distributed as dist
    return dist.get_world_size() if tvo_b200.get_run_policy() == "mpi" and dist.is_initialized() else 1


def _get_net_shape(net_shape: Sequence[int] = None, W_init: Sequence[Tensor] = None):
    """tvae.py:18-26: `shape` is given observables-first (D, ..., H0); stored hidden-first."""
    if net_shape is not None:
        return tuple(reversed(net_shape))
    assert W_init is not None, "Must pass one of `net_shape` and `W_init` to __init__ of the TVAE model"
    return tuple(w.shape[0] for w in W_init) + (W_init[-1].shape[1],)


class _Layer0(to.autograd.Function):
    """pre-activation of the first decoder layer from PACKED states: b_0 + sum_{h in s} W_0[h, :]."""

    @staticmethod
    def forward(ctx, packed: Tensor, H0: int, W0: Tensor, b0: Tensor):
        rows = packed.reshape(-1, packed.shape[-1])
        W0c, b0c = W0.contiguous(), b0.contiguous()
        out = to.empty((rows.shape[0], W0.shape[1]), dtype=W0.dtype, device=W0.device)
        _lib.call(f"tvo_tvae_layer0_{_lib.sfx(W0.dtype)}", _lib.ptr(rows), rows.shape[0], H0, _lib.ptr(W0c),
                  _lib.ptr(b0c), W0.shape[1], _lib.ptr(out), _lib.stream())
        ctx.save_for_backward(rows)
        ctx.H0, ctx.wshape = H0, W0.shape
        return out

    @staticmethod
    def backward(ctx, g: Tensor):
        (rows,) = ctx.saved_tensors
        g = g.contiguous()
        gW0 = gb0 = None
        if ctx.needs_input_grad[2]:
            gW0 = to.zeros(ctx.wshape, dtype=g.dtype, device=g.device)
            _lib.call(f"tvo_tvae_layer0_grad_{_lib.sfx(g.dtype)}", _lib.ptr(rows), rows.shape[0], ctx.H0, _lib.ptr(g),
                      g.shape[1], _lib.ptr(gW0), _lib.stream())
        if ctx.needs_input_grad[3]:
            gb0 = g.sum(dim=0)
        return None, None, gW0, gb0


class _TVAE:
    """Shared machinery (tvae.py:85-183)."""

    _kind = 0  # tvo_tvae_lpj: 0 Gaussian, 1 Bernoulli

    def _setup(self, shape, precision, min_lr, max_lr, cycliclr_step_size_up, pi_init, W_init, b_init,
               analytical_pi_updates, activation, external_model, optimizer, extra_gd: List[Tensor]):
        dev = tvo_b200.get_device()
        assert (
            (shape is not None and W_init is None and b_init is None and external_model is None)
            or (shape is None and W_init is not None and b_init is not None and external_model is None)
            or (shape is None and W_init is None and b_init is None and external_model is not None)
        ), "Must exclusively specify one one `shape`, (`W_init`, `b_init`), `external_model`"
        self._precision = precision
        self._activation = activation
        self._external_model = external_model
        if external_model is not None:
            assert hasattr(external_model, "H0"), "for externally defined models, H0 has to be provided manually"
            assert hasattr(external_model, "shape"), "for externally defined models, shape has to be provided manually"
            assert activation is None, "Must specify activation as part of external_model"
            H0 = external_model.H0
            self._net_shape = external_model.shape
            self.W = self.b = None
            gd_parameters = list(external_model.parameters())
        else:
            self._net_shape = _get_net_shape(shape, W_init)
            H0 = self._net_shape[0]
            n_layers = len(self._net_shape) - 1
            if W_init is None:  # Xavier (tvae.py:37-40)
                Ws = [to.nn.init.xavier_normal_(to.empty(self._net_shape[ln], self._net_shape[ln + 1]))
                      for ln in range(n_layers)]
            else:
                assert len(W_init) == n_layers, f"Shape is {self._net_shape} but {len(W_init)} weights passed"
                expected = [(self._net_shape[ln], self._net_shape[ln + 1]) for ln in range(n_layers)]
                got = [tuple(w.shape) for w in W_init]
                assert got == expected, f"Input W shapes: {got}\nExpected W shapes {expected}"
                Ws = [w.clone() for w in W_init]
            if b_init is None:
                bs = [to.zeros(s) for s in self._net_shape[1:]]
            else:
                assert all(b.shape == (self._net_shape[ln + 1],) for ln, b in enumerate(b_init))
                bs = [b.clone() for b in b_init]
            Ws = [w.detach().to(device=dev, dtype=precision) for w in Ws]
            for w in Ws:
                broadcast(w)
            self.W = [w.requires_grad_(True) for w in Ws]
            self.b = [b.detach().to(device=dev, dtype=precision).requires_grad_(True) for b in bs]
            self._theta.update({f"W_{i}": W for i, W in enumerate(self.W)})
            self._theta.update({f"b_{i}": b for i, b in enumerate(self.b)})
            gd_parameters = self.W + self.b
            self._activation = to.nn.ReLU() if activation is None else activation
            assert callable(self._activation)
        if pi_init is None:
            pi = to.full((H0,), 1 / H0)
        else:
            assert pi_init.shape == (H0,)
            pi = pi_init.clone()
        self._theta["pies"] = pi.detach().to(device=dev, dtype=precision).requires_grad_(not analytical_pi_updates)
        self._min_lr, self._max_lr, self._step_size_up = min_lr, max_lr, cycliclr_step_size_up
        gd_parameters = gd_parameters + extra_gd
        self._analytical_pi_updates = bool(analytical_pi_updates)
        if analytical_pi_updates:
            self._new_pi = to.zeros(H0, dtype=precision, device=dev)
        else:
            gd_parameters.append(self._theta["pies"])
        self._optimizer = opt.Adam(gd_parameters, lr=min_lr) if optimizer is None else optimizer
        # triangular cyclic schedule without momentum cycling (tvae.py:303-309; tvo/utils/CyclicLR.py)
        self._scheduler = opt.lr_scheduler.CyclicLR(self._optimizer, base_lr=min_lr, max_lr=max_lr,
                                                    step_size_up=cycliclr_step_size_up, cycle_momentum=False)
        self._train_datapoints = to.tensor([0], dtype=to.int, device=dev)

    # ------------------------------------------------------------------------------------ protocol
    @property
    def theta(self) -> Dict[str, Tensor]:
        return self._theta

    @property
    def config(self):
        return self._config

    @property
    def precision(self) -> to.dtype:
        return self._precision

    @property
    def shape(self) -> Tuple[int, ...]:
        """(D, H0): the model as a Bayes net; the network in between is ignored (tvae.py:128-134)."""
        return (self._net_shape[-1], self._net_shape[0])

    @property
    def net_shape(self) -> Tuple[int, ...]:
        """(D, Hn, ..., H0) (tvae.py:136-139)."""
        return tuple(reversed(self._net_shape))

    # NOTE: no log_pseudo_joint / init_storage / init_epoch / init_batch / sorted_by_lpj here.  The reference's TVAE is
    # `Trainable`, not `Optimized` (tvae.py:86); `Optimized` is a runtime-checkable Protocol, so defining those names
    # would make isinstance(model, Optimized) true and the Trainer / generic states paths would store pseudo-joints
    # in states.lpj where free_energy expects log-joints.

    # ------------------------------------------------------------------------------------ decoder
    def _finish(self, x: Tensor) -> Tensor:  # output non-linearity
        return x

    def forward(self, x: Tensor) -> Tensor:
        """Decoder applied to DENSE states (..., H0) (tvae.py:434-467, 672-701)."""
        assert x.shape[-1] == self._net_shape[0], "Incompatible shape in forward input"
        out = x.to(dtype=self.precision)
        if self._external_model is not None:
            return self._external_model.forward(out).to(dtype=self.precision)
        for W, b in zip(self.W[:-1], self.b[:-1]):
            out = self._activation(out @ W + b)
        return self._finish(out @ self.W[-1] + self.b[-1])

    def _uses_packed_path(self) -> bool:
        return self._external_model is None

    def _forward_packed(self, packed: Tensor) -> Tensor:
        """Decoder applied to PACKED states (B, S, W64) -> (B, S, D); differentiable w.r.t. W_i, b_i."""
        B, S = packed.shape[0], packed.shape[1]
        H0 = self._net_shape[0]
        out = _Layer0.apply(packed.contiguous(), H0, self.W[0], self.b[0])
        if len(self.W) == 1:
            return self._finish(out).view(B, S, -1)
        out = self._activation(out)
        for W, b in zip(self.W[1:-1], self.b[1:-1]):
            out = self._activation(out @ W + b)
        return self._finish(out @ self.W[-1] + self.b[-1]).view(B, S, -1)

    def _prior_logits(self) -> Tensor:
        pi = self._theta["pies"].detach()
        return to.log(pi / (1.0 - pi)).contiguous()

    def _lpj_from_out(self, data: Tensor, out: Tensor, packed: Tensor, st_stride_b: int, lpj: Tensor,
                      lpj_stride_b: int, B: int, S: int) -> None:
        """tvo_tvae_lpj: lpj[b, s] from the decoder output (no grad)."""
        D, H0 = self.shape
        Y = data.to(dtype=self.precision)
        Y = Y if Y.stride(-1) == 1 else Y.contiguous()
        sig = self._theta["sigma2"].detach() if "sigma2" in self._theta else None
        outc, prior = out.contiguous(), self._prior_logits()  # named: they must outlive the pointer extraction
        _lib.call(f"tvo_tvae_lpj_{_lib.sfx(self.precision)}", self._kind, _lib.ptr(Y), Y.stride(0), _lib.ptr(outc),
                  _lib.ptr(packed), st_stride_b, _lib.ptr(prior), _lib.ptr(sig), _lib.ptr(lpj), lpj_stride_b, B, S, H0,
                  D, _lib.stream())

    # hooks of EVOVariationalStates (generic path with this model's packed lpj) --------------------
    def _tvo_estep_evo(self, states, idx, batch, cfg) -> bool:
        return False

    def _tvo_sparsity(self) -> Tensor:
        return self._theta["pies"].detach().mean().to(to.float64).reshape(1)

    def _tvo_lpj_packed(self, batch: Tensor, packed: Tensor, H: int, k_broadcast: bool = False) -> Tensor:
        """What the variational states store for this model: the reference's TVAE is `Trainable` but not `Optimized`,
        so its E-step scores states with `log_joint` (evo.py:81-86) and `states.lpj` holds actual log-joints, which
        `free_energy` reads back (model_protocols.py:78-92)."""
        with to.no_grad():
            return self.log_joint(batch, None, self._lpj_packed(batch, packed, H, k_broadcast))

    def _lpj_packed(self, batch: Tensor, packed: Tensor, H: int, k_broadcast: bool = False) -> Tensor:
        B = batch.shape[0]
        S, W = packed.shape[1], packed.shape[2]
        if not self._uses_packed_path():
            dense = unpack(packed, H)
            if k_broadcast:
                dense = dense.expand(B, S, H)
            with to.no_grad():
                return self._lpj_and_mlpout(batch, dense)[0]
        packed = packed.contiguous()
        lpj = to.empty((B, S), dtype=self.precision, device=packed.device)
        with to.no_grad():
            out = self._forward_packed(packed)  # (1 or B, S, D)
            if k_broadcast:
                out = out.expand(B, S, out.shape[-1])
        self._lpj_from_out(batch, out, packed, 0 if k_broadcast else S * W, lpj, S, B, S)
        self._check_finite(lpj)
        return lpj

    def _tvo_lpj_rows(self, states, idx: Tensor, batch: Tensor) -> None:
        rows = states._Kp if states._Kp.shape[0] == 1 and states.lpj.shape[0] != 1 else states._Kp[idx]
        bc = rows.shape[0] == 1 and idx.numel() != 1
        states.lpj[idx] = self._tvo_lpj_packed(batch, rows, states._H, k_broadcast=bc).to(states.lpj.dtype)

    @staticmethod
    def _check_finite(lpj: Tensor) -> None:
        assert not to.isnan(lpj).any() and not to.isinf(lpj).any()  # tvae.py:340

    # ------------------------------------------------------------------------------------ reference API
    def _lpj_and_mlpout(self, data: Tensor, states) -> Tuple[Tensor, Tensor]:
        """tvae.py:326-341 / 584-604 with autograd; `states` uint8 (N, S, H0) or PackedStates."""
        raise NotImplementedError

    def _s2_packed(self, packed: Tensor, bc: bool, N: int) -> Tensor:
        """states @ log(pi/(1-pi)) (tvae.py:337) from the bits; differentiable only through the dense route."""
        S, W = packed.shape[1], packed.shape[2]
        H0 = self._net_shape[0]
        zero_y = to.zeros((N, 1), dtype=self.precision, device=packed.device)
        zero_o = to.zeros((N * S, 1), dtype=self.precision, device=packed.device)
        s2 = to.empty((N, S), dtype=self.precision, device=packed.device)
        one = to.ones(1, dtype=self.precision, device=packed.device)
        packed, prior = packed.contiguous(), self._prior_logits()
        _lib.call(f"tvo_tvae_lpj_{_lib.sfx(self.precision)}", 0, _lib.ptr(zero_y), 1, _lib.ptr(zero_o), _lib.ptr(packed),
                  0 if bc else S * W, _lib.ptr(prior), _lib.ptr(one), _lib.ptr(s2), S, N, S, H0, 1, _lib.stream())
        return s2

    def _log_pseudo_joint(self, data: Tensor, states) -> Tensor:
        packed, H, bc = as_packed(states)
        return self._lpj_packed(data, packed, H, k_broadcast=bc)

    def _free_energy_from_logjoints(self, logjoints: Tensor) -> Tensor:
        Fn = to.logsumexp(logjoints, dim=1)
        assert Fn.shape == (logjoints.shape[0],)
        assert not to.isnan(Fn).any() and not to.isinf(Fn).any()
        return Fn.sum()

    def free_energy(self, idx: Tensor, batch: Tensor, states) -> float:
        """model_protocols.py:78-92 under no_grad (tvae.py:108-110): F from the stored log-joints."""
        with to.no_grad():
            return to.logsumexp(states.lpj[idx], dim=1).sum(dim=0).item()

    def update_param_batch(self, idx: Tensor, batch: Tensor, states) -> float:
        """tvae.py:118-126: one optimizer step on -F/B, then the analytical accumulators."""
        if to.isnan(batch).any():
            raise RuntimeError("There are NaNs in this batch")
        F, mlp_out = self._optimize_nn_params(idx, batch, states)
        with to.no_grad():
            self._accumulate_param_updates(idx, batch, states, mlp_out)
        return F

    def _batch_states(self, states, idx: Tensor):
        idx = idx.to(device=states.lpj.device, dtype=to.int64)
        Kp = states._Kp
        if Kp.shape[0] == 1 and states.lpj.shape[0] != 1:  # FullEM: one broadcast state set
            return PackedStates(Kp, states._H, N=idx.numel())
        return PackedStates(Kp[idx], states._H)

    def _optimize_nn_params(self, idx: Tensor, data: Tensor, states) -> Tuple[float, Tensor]:
        """tvae.py:141-166: F and the decoder output BEFORE the weight update."""
        Kb = self._batch_states(states, idx)
        lpj, mlp_out = self._lpj_and_mlpout(data, Kb)
        F = self._free_energy_from_logjoints(self.log_joint(data, Kb, lpj))
        loss = -F / data.shape[0]
        loss.backward()
        if get_world_size() > 1:  # mpi_average_grads (parallel.py:276-290)
            with to.no_grad():
                for p in self._optimizer.param_groups[0]["params"]:
                    if p.grad is not None:
                        all_reduce(p.grad)
                        p.grad /= get_world_size()
        self._optimizer.step()
        self._scheduler.step()
        self._optimizer.zero_grad()
        self._clamp_gd_parameters()
        return F.item(), mlp_out.detach()

    def _clamp_gd_parameters(self) -> None:
        with to.no_grad():
            pi = self._theta["pies"]
            if pi.requires_grad:
                to.clamp(pi, 1e-5, 1 - 1e-5, out=pi)

    def _accumulate_param_updates(self, idx, data, states, mlp_out) -> None:
        if not self._theta["pies"].requires_grad:  # pi_h = 1/N sum_n <K_nsh>_q  (tvae.py:421-423)
            idx = idx.to(device=states.lpj.device, dtype=to.int64).contiguous()
            self._new_pi.add_(state_marginals(states._Kp, states.lpj, idx, states._H).to(self.precision).sum(dim=0))
        self._train_datapoints.add_(data.shape[0])

    def data_estimator(self, idx: Tensor, batch: Tensor, states) -> Tensor:
        """<<y>_{p(y|s)}>_q (tvae.py:168-182)."""
        idx = idx.to(device=states.lpj.device, dtype=to.int64)
        lpj = states.lpj[idx]
        with to.no_grad():
            if self._uses_packed_path():
                Kb = self._batch_states(states, idx)
                means = self._forward_packed(Kb.packed)
                if Kb.broadcast:
                    means = means.expand(idx.numel(), -1, -1)
            else:
                means = self.forward(states.K[idx])
            q = to.softmax(lpj.to(self.precision), dim=1)
            return (q.unsqueeze(2) * means).sum(dim=1)

    def _hidden(self, N, hidden_state):
        H = self.shape[-1]
        if hidden_state is None:
            pies = self._theta["pies"].detach()
            return to.rand((N, H), dtype=pies.dtype, device=pies.device) < pies, True
        if N is not None:
            assert hidden_state.shape == (N, H), f"hidden_state has shape {hidden_state.shape}, expected ({N}, {H})"
        return hidden_state, False


class GaussianTVAE(_TVAE):
    """tvae.py:185-467."""

    _kind = 0

    def __init__(self, shape: Sequence[int] = None, precision: to.dtype = to.float64, min_lr: float = 0.001,
                 max_lr: float = 0.01, cycliclr_step_size_up=400, pi_init: Tensor = None,
                 W_init: Sequence[Tensor] = None, b_init: Sequence[Tensor] = None, sigma2_init: float = None,
                 analytical_sigma_updates: bool = True, analytical_pi_updates: bool = True,
                 clamp_sigma_updates: bool = False, activation: Callable = None,
                 external_model: Optional[to.nn.Module] = None, optimizer: Optional[opt.Optimizer] = None):
        dev = tvo_b200.get_device()
        self._theta: Dict[str, Tensor] = {}
        self._clamp_sigma = clamp_sigma_updates
        sigma2 = to.tensor([0.01] if sigma2_init is None else [sigma2_init])
        sigma2 = sigma2.to(device=dev, dtype=precision).requires_grad_(not analytical_sigma_updates)
        self._analytical_sigma_updates = bool(analytical_sigma_updates)
        self._setup(shape, precision, min_lr, max_lr, cycliclr_step_size_up, pi_init, W_init, b_init,
                    analytical_pi_updates, activation, external_model, optimizer,
                    extra_gd=[] if analytical_sigma_updates else [sigma2])
        pies = self._theta.pop("pies")  # keep the reference's key order: W_i, b_i, pies, sigma2
        self._theta["pies"] = pies
        self._theta["sigma2"] = sigma2
        if analytical_sigma_updates:
            self._new_sigma2 = to.zeros(1, dtype=precision, device=dev)
        self._config = dict(
            net_shape=self._net_shape, activation=self._activation, external_model=self._external_model,
            precision=self.precision, min_lr=self._min_lr, max_lr=self._max_lr, step_size_up=self._step_size_up,
            analytical_sigma_updates=self._analytical_sigma_updates,
            analytical_pi_updates=self._analytical_pi_updates, clamp_sigma_updates=self._clamp_sigma, device=dev,
        )

    def _uses_packed_path(self) -> bool:
        return self._external_model is None and not self._theta["pies"].requires_grad

    def _lpj_and_mlpout(self, data: Tensor, states) -> Tuple[Tensor, Tensor]:
        N = data.shape[0]
        pi, sigma2 = self._theta["pies"], self._theta["sigma2"]
        if self._uses_packed_path():
            packed, H, bc = as_packed(states)
            assert bc or packed.shape[0] == N, "Shape mismatch between data and states"
            mlp_out = self._forward_packed(packed)
            if bc:
                mlp_out = mlp_out.expand(N, -1, -1)
            s2 = self._s2_packed(packed, bc, N)
        else:
            dense = states[:] if isinstance(states, PackedStates) else states
            assert dense.shape[0] == N, "Shape mismatch between data and states"
            dense = dense.to(dtype=self.precision)
            mlp_out = self.forward(dense)
            s2 = dense @ to.log(pi / (1.0 - pi))
        # nansum ignores missing data (tvae.py:335-336)
        s1 = to.nansum((data.to(self.precision).unsqueeze(1) - mlp_out).pow(2), dim=2).div(2 * sigma2)
        lpj = s2 - s1
        assert lpj.shape == (N, mlp_out.shape[1])
        self._check_finite(lpj.detach())
        return lpj, mlp_out

    def log_joint(self, data, states, lpj=None):
        """tvae.py:343-351"""
        pi, sigma2 = self._theta["pies"], self._theta["sigma2"]
        D = (data.shape[1] - to.isnan(data).sum(dim=1)).unsqueeze(1)
        if lpj is None:
            lpj = self._log_pseudo_joint(data, states)
        return lpj - D / 2 * to.log(2 * MATH_PI * sigma2) + to.log(1 - pi).sum()

    def _clamp_gd_parameters(self) -> None:
        super()._clamp_gd_parameters()
        with to.no_grad():
            sigma2 = self._theta["sigma2"]
            if sigma2.requires_grad:
                to.clamp(sigma2, 1e-5, out=sigma2)

    def _accumulate_param_updates(self, idx, data, states, mlp_out) -> None:
        if not self._theta["sigma2"].requires_grad:
            # sigma2 = 1/(DN) sum_{n,d} <(y_nd - a_d)^2>_q  (tvae.py:425-431), summed over d before the posterior mean
            i = idx.to(device=states.lpj.device, dtype=to.int64)
            q = to.softmax(states.lpj[i].to(self.precision), dim=1)
            sq = (data.to(self.precision).unsqueeze(1) - mlp_out).pow(2).sum(dim=2)
            self._new_sigma2.add_((q * sq).sum())
        super()._accumulate_param_updates(idx, data, states, mlp_out)

    def update_param_epoch(self) -> None:
        """tvae.py:353-391"""
        pi, sigma2 = self._theta["pies"], self._theta["sigma2"]
        if get_world_size() > 1:
            with to.no_grad():
                for p in self._theta.values():
                    if p.requires_grad:
                        broadcast(p)
        if pi.requires_grad and sigma2.requires_grad:
            return
        D = self._net_shape[-1]
        all_reduce(self._train_datapoints)
        N = self._train_datapoints.item()
        with to.no_grad():
            if not pi.requires_grad:
                all_reduce(self._new_pi)
                pi[:] = self._new_pi / N
                to.clamp(pi, 1e-5, 1 - 1e-5, out=pi)
                self._new_pi.zero_()
            if not sigma2.requires_grad:
                all_reduce(self._new_sigma2)
                lo = (sigma2 - sigma2.abs() * 0.01).item()  # at most 1 % change per epoch when clamped
                hi = (sigma2 + sigma2.abs() * 0.01).item()
                sigma2[:] = self._new_sigma2 / (N * D)
                if self._clamp_sigma:
                    to.clamp(sigma2, lo, hi, out=sigma2)
                self._new_sigma2.zero_()
        self._train_datapoints[:] = 0

    def generate_data(self, N: int = None, hidden_state: Tensor = None) -> Union[Tensor, Tuple[Tensor, Tensor]]:
        """tvae.py:393-410"""
        hidden_state, ret = self._hidden(N, hidden_state)
        with to.no_grad():
            mlp_out = self.forward(hidden_state)
        Y = to.distributions.Normal(loc=mlp_out, scale=to.sqrt(self._theta["sigma2"].detach())).sample()
        return (Y, hidden_state) if ret else Y


class BernoulliTVAE(_TVAE):
    """tvae.py:469-701: sigmoid output layer, binary cross-entropy likelihood."""

    _kind = 1

    def __init__(self, shape: Sequence[int] = None, precision: to.dtype = to.float64, min_lr: float = 0.001,
                 max_lr: float = 0.01, cycliclr_step_size_up=400, pi_init: Tensor = None,
                 W_init: Sequence[Tensor] = None, b_init: Sequence[Tensor] = None, analytical_pi_updates: bool = True,
                 activation: Callable = None, external_model: Optional[to.nn.Module] = None,
                 optimizer: Optional[opt.Optimizer] = None):
        self._theta: Dict[str, Tensor] = {}
        self._setup(shape, precision, min_lr, max_lr, cycliclr_step_size_up, pi_init, W_init, b_init,
                    analytical_pi_updates, activation, external_model, optimizer, extra_gd=[])
        self._config = dict(
            net_shape=self._net_shape, activation=self._activation, external_model=self._external_model,
            precision=self.precision, min_lr=self._min_lr, max_lr=self._max_lr, step_size_up=self._step_size_up,
            analytical_pi_updates=self._analytical_pi_updates, device=tvo_b200.get_device(),
        )

    def _finish(self, x: Tensor) -> Tensor:
        return to.sigmoid(x)

    def _uses_packed_path(self) -> bool:
        return self._external_model is None and not self._theta["pies"].requires_grad

    def _lpj_and_mlpout(self, data: Tensor, states) -> Tuple[Tensor, Tensor]:
        N, D = data.shape
        pi = self._theta["pies"]
        if self._uses_packed_path():
            packed, H, bc = as_packed(states)
            assert bc or packed.shape[0] == N, "Shape mismatch between data and states"
            mlp_out = self._forward_packed(packed)
            if bc:
                mlp_out = mlp_out.expand(N, -1, -1)
            s2 = self._s2_packed(packed, bc, N)
        else:
            dense = states[:] if isinstance(states, PackedStates) else states
            assert dense.shape[0] == N, "Shape mismatch between data and states"
            dense = dense.to(dtype=self.precision)
            mlp_out = self.forward(dense)
            s2 = dense @ to.log(pi / (1.0 - pi))
        S = mlp_out.shape[1]
        target = data.to(self.precision).unsqueeze(1).expand(N, S, D)
        s1 = to.nansum(to.nn.functional.binary_cross_entropy(mlp_out, target, reduction="none"), dim=2)
        lpj = s2 - s1
        assert lpj.shape == (N, S)
        self._check_finite(lpj.detach())
        return lpj, mlp_out

    def log_joint(self, data, states, lpj=None):
        """tvae.py:606-613"""
        if lpj is None:
            lpj = self._log_pseudo_joint(data, states)
        return lpj + to.log(1 - self._theta["pies"]).sum()

    def update_param_epoch(self) -> None:
        """tvae.py:615-636"""
        pi = self._theta["pies"]
        if get_world_size() > 1:
            with to.no_grad():
                for p in self._theta.values():
                    if p.requires_grad:
                        broadcast(p)
        if pi.requires_grad:
            return
        all_reduce(self._train_datapoints)
        N = self._train_datapoints.item()
        with to.no_grad():
            all_reduce(self._new_pi)
            pi[:] = self._new_pi / N
            to.clamp(pi, 1e-5, 1 - 1e-5, out=pi)
            self._new_pi.zero_()
        self._train_datapoints[:] = 0

    def generate_data(self, N: int = None, hidden_state: Tensor = None) -> Union[Tensor, Tuple[Tensor, Tensor]]:
        """tvae.py:638-656"""
        hidden_state, ret = self._hidden(N, hidden_state)
        with to.no_grad():
            mlp_out = self.forward(hidden_state)
        Y = to.distributions.Bernoulli(mlp_out).sample()
        return (Y, hidden_state) if ret else Y
