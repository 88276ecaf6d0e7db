"""Binary Sparse Coding (mirror of tvo/models/bsc.py:23-252) on sm_100a kernels.

Public surface identical to the reference: constructor arguments, `theta` keys (pies, W, sigma2,
mutated in place, user-rebindable), `log_pseudo_joint`, `log_joint`, `update_param_batch`,
`update_param_epoch`, `free_energy`, `generate_data`, `data_estimator`, accumulators
`my_Wp/my_Wq/my_pies/my_sigma2/my_N`.

Beneath it:
 * theta is digested on the device by tvo_bsc_prepare (padded W^T, log-odds, constants) at every
   call — nothing derived from theta is cached across calls;
 * log_pseudo_joint -> tvo_bsc_lpj (sparse column walk over bit-packed states);
 * EVO E-step -> tvo_estep_evo_bsc (one fused kernel);
 * update_param_batch -> tvo_bsc_mstep (softmax weights on chip, sparse accumulation);
 * all sufficient statistics + N + F live in ONE fp64 buffer `_stats` = [WpT | Wq | pies | sigma2 |
   N | F], all-reduced once per EM iteration, followed by one packed broadcast of the new theta
   (the reference issues 5 all_reduce + 4 broadcasts, bsc.py:164-175, sanity.py:40).
"""
import ctypes as C
import math
from typing import Tuple, Union

import torch as to
from torch import Tensor

import tvo_b200
from tvo_b200 import _lib
from tvo_b200.utils.model_protocols import Optimized, Reconstructor, Sampler
from tvo_b200.utils.parallel import all_reduce, broadcast, pprint
from tvo_b200.utils.sanity import fix_theta
from tvo_b200.variational._packed import PackedStates, as_packed
from tvo_b200.variational._utils import mean_posterior, state_marginals


_SIDE = {}  # device -> (side stream, pinned 2-double buffer) for the factorisation flags of update_param_epoch


class BSC(Optimized, Sampler, Reconstructor):
    def __init__(
        self,
        H: int,
        D: int,
        W_init: Tensor = None,
        sigma2_init: Tensor = None,
        pies_init: Tensor = None,
        individual_priors: bool = True,
        precision: to.dtype = to.float64,
    ):
        assert precision in (to.float32, to.float64), "precision must be one of torch.float{32,64}"
        self._precision = precision
        device = tvo_b200.get_device()
        if W_init is not None:
            assert W_init.shape == (D, H)
            W_init = W_init.to(dtype=precision, device=device).clone()
        else:
            W_init = to.rand((D, H), dtype=precision, device=device)
            broadcast(W_init)
        if pies_init is not None:
            assert pies_init.shape == (H,) if individual_priors else pies_init.shape == (1,)
            pies_init = pies_init.to(dtype=precision, device=device).clone()
        else:
            pies_init = (to.full((H,), 1.0 / H, dtype=precision, device=device) if individual_priors
                         else to.tensor([1.0 / H], dtype=precision, device=device))
        if sigma2_init is not None:
            assert sigma2_init.shape == (1,)
            sigma2_init = sigma2_init.to(dtype=precision, device=device).clone()
        else:
            sigma2_init = to.tensor([1.0], dtype=precision, device=device)
        self._theta = {"pies": pies_init, "W": W_init, "sigma2": sigma2_init}
        eps, inf = 1.0e-5, math.inf
        self.policy = {
            "W": [None, to.full_like(self._theta["W"], -inf), to.full_like(self._theta["W"], inf)],
            "pies": [None, to.full_like(self._theta["pies"], eps), to.full_like(self._theta["pies"], 1.0 - eps)],
            "sigma2": [None, to.full_like(self._theta["sigma2"], eps), to.full_like(self._theta["sigma2"], inf)],
        }
        self._config = dict(H=H, D=D, individual_priors=individual_priors, precision=precision, device=device)
        self._shape = (D, H)
        self._stats = to.zeros(H * D + H * H + H + 3, dtype=to.float64, device=device)
        self._ws = None
        self._fresh = None  # (theta versions, batch ptr/version): lpj in `states` was computed from these
        self._ws_token = None  # (theta versions, stream) the workspace was digested from
        self._chol_work = None  # factor / flags of the hand-written epoch solve (H <= 256)
        self._chol_sol = None
        self._last_F = None
        self.use_gram = True  # Gram-form fused E-step (False: direct column-walk form, D <= 256)
        self.scratch_datapoints = 524288  # chunk size of the Gram E-step / M-step GEMMs (scratch ~ chunk*(H*8+12) B = 1.1 GB at
        #                                   H = 256; 131072 cost 2.5 % of the E-step and 13 % of the M-step in launch tails)
        self._scratch = None
        self.W_solver = "gpu"  # "gpu": Cholesky on the device with lstsq as fallback; "gpu_lstsq": torch.linalg.lstsq
        #                        on the device (as the reference under TVO_GPU); "cpu": rank-revealing gelsy on the host
        #                        (as the reference's CPU path); see update_param_epoch
        self._wq_full = None

    # --- accumulators under the reference's names (views into the packed buffer) ----------------
    @property
    def my_Wp(self) -> Tensor:
        D, H = self._shape
        return self._stats[: H * D].view(H, D).t()

    @property
    def my_Wq(self) -> Tensor:
        """Symmetric sum_n <s s^T>_n (bsc.py:146-147).  The kernel accumulates the upper triangle only."""
        D, H = self._shape
        U = self._stats[H * D: H * D + H * H].view(H, H)
        return U + U.t() - to.diag(to.diagonal(U))

    def _get_scratch(self, B: int) -> Tensor:
        D, H = self._shape
        n = min(max(int(B), 1), int(self.scratch_datapoints), max(1024, (2 << 30) // (H * 8)))  # at most ~2 GB of scratch
        # chunk rows (a = W^T y / E and the fp64 copy of Y for fp32 models) + the E-step's sparseflip table / overflow list
        f64 = int(self._precision == to.float64)
        nbytes = max(_lib.lib().tvo_bsc_estep_scratch_bytes(n, H, f64), _lib.lib().tvo_bsc_mstep_scratch_bytes(n, H, D, f64))
        if self._scratch is None or self._scratch.numel() < nbytes or self._scratch.device != self._stats.device:
            self._scratch = to.empty(nbytes, dtype=to.uint8, device=self._stats.device)
        return self._scratch

    @property
    def my_pies(self) -> Tensor:
        D, H = self._shape
        o = H * D + H * H
        return self._stats[o: o + H]

    @property
    def my_sigma2(self) -> Tensor:
        D, H = self._shape
        o = H * D + H * H + H
        return self._stats[o: o + 1]

    @property
    def my_N(self) -> Tensor:
        D, H = self._shape
        o = H * D + H * H + H + 1
        return self._stats[o: o + 1]

    @property
    def shape(self) -> Tuple[int, ...]:
        return self.theta["W"].shape

    # --- theta digestion -------------------------------------------------------------------------
    def _theta_token(self):
        return tuple((k, t.data_ptr(), t._version) for k, t in self._theta.items())

    def _prepare(self):
        D, H = self._shape
        th = self._theta
        prec = self._precision
        W, pies, s2 = (th[k].to(dtype=prec, device=self._stats.device).contiguous() for k in ("W", "pies", "sigma2"))
        nbytes = _lib.lib().tvo_bsc_theta_ws_bytes(H, D, int(prec == to.float64))
        if self._ws is None or self._ws.numel() != nbytes or self._ws.device != W.device:
            self._ws = to.empty(nbytes, dtype=to.uint8, device=W.device)
            self._ws_token = None
        # theta unchanged since the last digestion on this stream (in-place torch writes bump the tensors' versions,
        # update_param_epoch drops the token): the E-step and the M-step of one batch share one digestion
        token = (self._theta_token(), _lib.stream().value)  # c_void_p objects compare by identity
        if self._ws_token == token:
            return self._ws
        self._ws_token = token
        _lib.call(f"tvo_bsc_prepare_{_lib.sfx(prec)}", _lib.ptr(W), _lib.ptr(pies), pies.numel(), _lib.ptr(s2), H, D,
                  _lib.ptr(self._ws), _lib.stream())
        return self._ws

    def _batch(self, data: Tensor) -> Tensor:
        data = data.to(device=self._stats.device, dtype=self._precision)
        return data if data.stride(-1) == 1 and data.dim() == 2 else data.contiguous()

    # --- hooks used by the tvo_b200 variational states -------------------------------------------
    def _tvo_sparsity(self) -> Tensor:
        """mean(pies) (evo.py:109) as a device double: computed by the theta digestion (BscThetaHdr::sparsity, byte 32
        of the workspace) — no extra reduction kernel per E-step."""
        return self._prepare()[32:40].view(to.float64)

    def _tvo_lpj_packed(self, batch: Tensor, packed: Tensor, H: int, k_broadcast: bool = False, ws=None) -> Tensor:
        """lpj (B,S) for packed states (B or 1, S, W64)."""
        D = self._shape[0]
        Y = self._batch(batch)
        B, S, W = Y.shape[0], packed.shape[1], packed.shape[2]
        packed = packed.contiguous()
        out = to.empty((B, S), dtype=self._precision, device=Y.device)
        ws = self._prepare() if ws is None else ws
        _lib.call(f"tvo_bsc_lpj_{_lib.sfx(self._precision)}", _lib.ptr(Y), Y.stride(0), _lib.ptr(packed),
                  0 if k_broadcast else S * W, None, _lib.ptr(ws), _lib.ptr(out), S, B, S, H, D, _lib.stream())
        return out

    def _tvo_lpj_rows(self, states, idx: Tensor, batch: Tensor, ws=None) -> None:
        """states.lpj[idx] = lpj(batch, states.K[idx]) in place, rows addressed on the device."""
        D = self._shape[0]
        Y = self._batch(batch)
        Kp = states._Kp
        S, W = Kp.shape[1], Kp.shape[2]
        assert states.lpj.dtype == self._precision and states.lpj.is_contiguous()
        ws = self._prepare() if ws is None else ws
        _lib.call(f"tvo_bsc_lpj_{_lib.sfx(self._precision)}", _lib.ptr(Y), Y.stride(0), _lib.ptr(Kp),
                  0 if Kp.shape[0] == 1 and states.lpj.shape[0] != 1 else S * W, _lib.ptr(idx), _lib.ptr(ws),
                  _lib.ptr(states.lpj), states.lpj.shape[1], idx.numel(), S, states._H, D, _lib.stream())
        self._fresh = (self._theta_token(), batch.data_ptr(), batch._version, id(states))

    def _tvo_estep_evo(self, states, idx: Tensor, batch: Tensor, cfg) -> bool:
        """Fused EVO E-step; returns False when the configuration needs the unfused kernels."""
        D, H = self._shape
        if states.lpj.dtype != self._precision:
            return False
        if cfg.parent_selection == _lib.SEL["batch_fitparents"] and not (cfg.fit_norm and cfg.flags & _lib.EVO_LPJ_CURRENT):
            return False  # batch-global normaliser: fused for one generation (EVOVariationalStates.update_device)
        if (D > 256 or cfg.parent_selection == _lib.SEL["batch_fitparents"]) and not self.use_gram:
            return False
        Y = self._batch(batch)
        ws = self._prepare()
        S = states._Kp.shape[1]
        sparsity = self._tvo_sparsity() if cfg.mutation in (_lib.MUT["sparseflip"], _lib.MUT["cross_sparseflip"]) else None
        scratch = self._get_scratch(idx.numel()) if self.use_gram else None
        _lib.call(f"tvo_estep_evo_bsc_{_lib.sfx(self._precision)}", _lib.ptr(Y), Y.stride(0), _lib.ptr(idx),
                  _lib.ptr(states._Kp), _lib.ptr(states.lpj), _lib.ptr(ws), C.byref(cfg), _lib.ptr(sparsity),
                  _lib.ptr(scratch), scratch.numel() if scratch is not None else 0, idx.numel(), S, H, D,
                  _lib.ptr(states._nsubs_dev), _lib.stream())
        self._fresh = (self._theta_token(), batch.data_ptr(), batch._version, id(states))
        return True

    # --- reference API ---------------------------------------------------------------------------
    def log_pseudo_joint(self, data: Tensor, states) -> Tensor:  # type: ignore
        """bsc.py:100-119.  `states`: uint8 (N,S,H) tensor or PackedStates."""
        packed, H, bc = as_packed(states)
        return self._tvo_lpj_packed(data, packed, H, k_broadcast=bc)

    def log_joint(self, data: Tensor, states, lpj: Tensor = None) -> Tensor:
        """bsc.py:121-132"""
        if lpj is None:
            lpj = self.log_pseudo_joint(data, states)
        data = data.to(lpj.device)
        Dn = to.sum(to.logical_not(to.isnan(data)), dim=1)
        H = self.shape[1]
        th = self.theta
        priorterm = (to.log(1 - th["pies"]).sum() if self.config["individual_priors"]
                     else H * to.log(1 - th["pies"]))
        return lpj + priorterm - Dn.unsqueeze(1) / 2 * to.log(2 * math.pi * th["sigma2"])

    def _free_energy_device(self, idx: Tensor, batch: Tensor, states, out: Tensor, ws=None) -> None:
        """out (device double scalar) += sum_b logsumexp_s log_joint  (no host sync)."""
        D = self._shape[0]
        Y = self._batch(batch)
        idx = idx.to(device=Y.device, dtype=to.int64).contiguous()
        ws = self._prepare() if ws is None else ws
        lpj = states.lpj
        assert lpj.is_contiguous() and lpj.dtype == self._precision
        _lib.call(f"tvo_bsc_free_energy_{_lib.sfx(self._precision)}", _lib.ptr(Y), Y.stride(0), _lib.ptr(lpj),
                  lpj.shape[1], _lib.ptr(idx), _lib.ptr(ws), idx.numel(), lpj.shape[1], D, _lib.ptr(out), _lib.stream())

    def free_energy(self, idx: Tensor, batch: Tensor, states) -> float:
        """model_protocols.py:179-193"""
        out = to.zeros((), dtype=to.float64, device=self._stats.device)
        self._free_energy_device(idx, batch, states, out)
        return out.item()

    def update_param_batch(self, idx: Tensor, batch: Tensor, states) -> None:
        """bsc.py:134-158 -> tvo_bsc_mstep.  The batch free energy is accumulated into the packed
        statistics as a by-product (read by the tvo_b200 Trainer instead of a per-batch free_energy)."""
        D, H = self._shape
        Y = self._batch(batch)
        idx = idx.to(device=Y.device, dtype=to.int64).contiguous()
        Kp, lpj = states._Kp, states.lpj
        S, W = Kp.shape[1], Kp.shape[2]
        assert lpj.dtype == self._precision and lpj.is_contiguous()
        fresh = self._fresh == (self._theta_token(), batch.data_ptr(), batch._version, id(states))
        ws = self._prepare()
        scratch = self._get_scratch(idx.numel())
        _lib.call(f"tvo_bsc_mstep_{_lib.sfx(self._precision)}", _lib.ptr(Y), Y.stride(0), _lib.ptr(idx), _lib.ptr(Kp),
                  0 if Kp.shape[0] == 1 and lpj.shape[0] != 1 else S * W, _lib.ptr(lpj), lpj.shape[1], _lib.ptr(ws),
                  _lib.ptr(self._stats), _lib.ptr(scratch), scratch.numel(), idx.numel(), S, H, D, 0 if fresh else 1,
                  _lib.stream())
        return None

    def update_param_epoch(self) -> None:
        """bsc.py:160-204.  One all-reduce of the packed statistics, then on every rank: the solve of
        my_Wq W^T = my_Wp^T and ONE kernel (tvo_bsc_epoch_finish) that turns the solution and the statistics into the
        new theta in place, fix_theta included.

        W_solver "gpu" (default): my_Wq is symmetric positive semi-definite, so the system is solved by a Cholesky
        factorisation — for H <= 256 by the hand-written kernels of csrc/bsc_chol.cuh (tvo_bsc_epoch_solve: 2 launches,
        deterministic, so every rank derives bit-identical theta from the all-reduced statistics and NO broadcast
        follows), for larger H by potrf + potrs through torch; torch.linalg.lstsq (the QR the reference calls under
        TVO_GPU) takes over if the factorisation fails, its pivots signal a condition number beyond ~1e12 (a unit
        that was never active) or the solution is not finite, and if that raises the reference's fallback applies
        (W <- W + 0.1 randn, bsc.py:174,180-181).  "gpu_lstsq": always lstsq on the device; "cpu": the host's
        rank-revealing gelsy (the reference's CPU path).  After any library solve a packed broadcast of theta from
        rank 0 follows when there is more than one rank (cuSOLVER / LAPACK are not guaranteed bit-reproducible across
        processes; the reference broadcasts W, pies, sigma2 separately, sanity.py:40)."""
        theta = self.theta
        D, H = self.shape
        all_reduce(self._stats)
        self._last_F = self._stats[-1].clone()
        self._last_N = self.my_N.clone()
        self._fresh = None  # theta changes under the lpj stored in the states
        self._ws_token = None
        prec = self._precision
        Wold_noisy = theta["W"] + 0.1 * to.randn_like(theta["W"])  # always drawn, as in the reference (bsc.py:174)
        WpT = self._stats[: H * D].view(H, D)
        direct = self._stats.is_cuda and all(theta[k].is_contiguous() and theta[k].dtype == prec and
                                             theta[k].device == self._stats.device for k in theta)
        sol = None
        if self.W_solver == "gpu" and direct:
            dev = self._stats.device
            if dev not in _SIDE:
                _SIDE[dev] = (to.cuda.Stream(device=dev), to.empty(2, dtype=to.float64).pin_memory())
            side, flags_host = _SIDE[dev]
            nwork = _lib.lib().tvo_bsc_epoch_solve_work_doubles(H)
            if nwork > 0:
                # H <= 256: blocked Cholesky + substitution kernels straight from the packed statistics (csrc/bsc_chol.cuh),
                # then the finishing kernel GUARDED by the factorisation flags on the device: the whole update is queued
                # before the host reads the flags (one small device-to-host copy per EM iteration, nothing behind it)
                if self._chol_work is None or self._chol_work.numel() != nwork + 2 or self._chol_work.device != dev:
                    self._chol_work = to.empty(nwork + 2, dtype=to.float64, device=dev)
                    self._chol_sol = to.empty((H, D), dtype=to.float64, device=dev)
                flags_dev = self._chol_work[nwork:]
                sol = self._chol_sol
                _lib.call("tvo_bsc_epoch_solve", _lib.ptr(self._stats), _lib.ptr(self._chol_work), _lib.ptr(sol),
                          _lib.ptr(flags_dev), H, D, _lib.stream())
                _lib.call(f"tvo_bsc_epoch_finish_{_lib.sfx(prec)}", _lib.ptr(self._stats), _lib.ptr(sol),
                          _lib.ptr(Wold_noisy.contiguous()), _lib.ptr(theta["W"]), _lib.ptr(theta["pies"]),
                          theta["pies"].numel(), _lib.ptr(theta["sigma2"]), H, D, _lib.ptr(flags_dev), _lib.stream())
                flags_host.copy_(flags_dev, non_blocking=True)
                self._prepare()  # digest the new theta now (G = W^T W, Wg, priors): runs while the host waits for the flags
                to.cuda.current_stream().synchronize()
                flags = flags_host.tolist()
                self._ws_token = self._ws_token if flags[0] == 0 and flags[1] > 1e-6 else None
                if flags[0] == 0 and flags[1] > 1e-6:  # cond(my_Wq) ~ (max/min pivot)^2 < 1e12
                    # deterministic solve on bit-identical all-reduced statistics: theta agrees on every rank without
                    # a broadcast ("theta update then runs identically on every rank")
                    self._stats.zero_()
                    return
                sol = None  # theta untouched: the reference's lstsq route below
            else:
                if self._wq_full is None or self._wq_full.device != dev:
                    self._wq_full = to.empty((H, H), dtype=to.float64, device=dev)
                _lib.call("tvo_bsc_symmetrize", _lib.ptr(self._stats[H * D: H * D + H * H]), _lib.ptr(self._wq_full), H,
                          _lib.stream())
                L, info = to.linalg.cholesky_ex(self._wq_full)
                dg = to.diagonal(L)
                flags_dev = to.stack([info.to(to.float64), dg.min() / dg.max()])
                # one host read per EM iteration, taken on a side stream while the triangular solves already run on
                # the launching stream (the solution is dropped if the factorisation turns out unusable)
                factored = to.cuda.Event()
                factored.record()
                sol = to.cholesky_solve(WpT, L)
                with to.cuda.stream(side):
                    side.wait_event(factored)
                    flags_host.copy_(flags_dev, non_blocking=True)
                    side.synchronize()
                flags = flags_host.tolist()
                if not (flags[0] == 0 and flags[1] > 1e-6):
                    sol = None
        if sol is None:
            try:
                Wq = self.my_Wq
                sol = (to.linalg.lstsq(Wq.cpu(), WpT.cpu())[0].to(WpT.device) if self.W_solver == "cpu"
                       else to.linalg.lstsq(Wq, WpT)[0])
            except RuntimeError:
                pprint("Inversion error. Will not update W but add some noise instead.")
                broadcast(Wold_noisy)
                sol = None
        if direct:
            sol = None if sol is None else sol.contiguous()
            _lib.call(f"tvo_bsc_epoch_finish_{_lib.sfx(prec)}", _lib.ptr(self._stats), _lib.ptr(sol),
                      _lib.ptr(Wold_noisy.contiguous()), _lib.ptr(theta["W"]), _lib.ptr(theta["pies"]),
                      theta["pies"].numel(), _lib.ptr(theta["sigma2"]), H, D, None, _lib.stream())
        else:  # user-rebound theta tensors (other dtype / device / strides): the reference's torch expressions
            N = self.my_N
            policy = self.policy
            theta_new = {"W": Wold_noisy if sol is None else sol.t().to(prec),
                         "pies": ((self.my_pies / N) if self.config["individual_priors"]
                                  else self.my_pies.sum(dim=0, keepdim=True) / N / H).to(prec),
                         "sigma2": (self.my_sigma2 / N / D).to(prec)}
            policy["W"][0], policy["pies"][0], policy["sigma2"][0] = Wold_noisy, theta["pies"], theta["sigma2"]
            fix_theta(theta_new, policy)
            for key in theta:
                theta[key][:] = theta_new[key].to(device=theta[key].device, dtype=theta[key].dtype)
        # library solvers (cuSOLVER / LAPACK) are not guaranteed bit-reproducible across processes: broadcast after those
        if tvo_b200.get_run_policy() == "mpi" and to.distributed.is_initialized() and to.distributed.get_world_size() > 1:
            flat = to.cat([theta[k].reshape(-1) for k in theta])
            broadcast(flat)
            o = 0
            for key in theta:
                n = theta[key].numel()
                theta[key][:] = flat[o:o + n].view_as(theta[key])
                o += n
        self._stats.zero_()

    def generate_data(self, N: int = None, hidden_state: Tensor = None) -> Union[Tensor, Tuple[Tensor, Tensor]]:
        """bsc.py:210-239"""
        precision, device = self.precision, self._stats.device
        D, H = self.shape
        if hidden_state is None:
            assert N is not None
            hidden_state = to.rand((N, H), dtype=precision, device=device) < self.theta["pies"]
            must_return_hidden_state = True
        else:
            if N is None:
                N = hidden_state.shape[0]
            assert hidden_state.shape == (N, H), f"hidden_state has shape {hidden_state.shape}, expected ({N},{H})"
            must_return_hidden_state = False
        Wbar = hidden_state.to(device=device, dtype=precision) @ self.theta["W"].t()
        Y = Wbar + to.sqrt(self.theta["sigma2"]) * to.randn((N, D), dtype=precision, device=device)
        return (Y, hidden_state) if must_return_hidden_state else Y

    def data_estimator(self, idx: Tensor, batch: Tensor, states) -> Tensor:
        """bsc.py:241-252: <W s>_q = W <s>_q, with <s>_q taken from the packed K (tvo_state_marginals)."""
        lpj = states.lpj
        idx = idx.to(device=lpj.device, dtype=to.int64).contiguous()
        marg = state_marginals(states._Kp, lpj, idx, states._H)
        return marg.to(self.precision) @ self.theta["W"].t()
