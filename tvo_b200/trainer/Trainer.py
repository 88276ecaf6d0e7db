"""Trainer (mirror of tvo/trainer/Trainer.py:14-327): same constructor, `em_step`, `e_step`,
`eval_free_energies`, same result dictionaries.

Fast path (tvo_b200 CUDA model + tvo_b200 variational states): the batch loop only enqueues
kernels — E-step (`states.update_device`), M-step accumulation (`model.update_param_batch`, which
also accumulates the batch free energy) — and the epoch ends with ONE packed all-reduce inside
`model.update_param_epoch()` plus one host read of (F, subs).  The reference syncs twice per batch
(`.item()` in update_states_for_batch and free_energy, Trainer.py:239,251) and accumulates F in
float32 (Trainer.py:230, quirk Q8); here F is accumulated in fp64.

Generic path (any model following the reference's protocols): the reference's loop, verbatim in
behaviour.
"""
from typing import Any, Callable, Dict, Sequence, Union

import torch as to

import tvo_b200
from tvo_b200.utils.data import TVODataLoader
from tvo_b200.utils.model_protocols import Optimized, Reconstructor, Trainable
from tvo_b200.utils.parallel import all_reduce
from tvo_b200.variational import TVOVariationalStates


def _fast(model, states) -> bool:
    """tvo_b200 CUDA model + tvo_b200 states, and the model's device paths cover this kind of states (the mixtures'
    dense kernels only do for FullEMSingleCauseModels; GMM / PMM with EVO / TVS / FullEM states take the generic
    path like any third-party model, as the reference allows: gmm.py:91-102)."""
    if not (hasattr(model, "_free_energy_device") and hasattr(states, "_nsubs_dev")):
        return False
    supports = getattr(model, "_supports_fast", None)
    return True if supports is None else bool(supports(states))


class Trainer:
    def __init__(
        self,
        model: Trainable,
        train_data: Union[TVODataLoader, to.Tensor] = None,
        train_states: TVOVariationalStates = None,
        test_data: Union[TVODataLoader, to.Tensor] = None,
        test_states: TVOVariationalStates = None,
        rollback_if_F_decreases: Sequence[str] = [],
        will_reconstruct: bool = False,
        eval_F_at_epoch_end: bool = False,
        data_transform: Callable[[to.Tensor], to.Tensor] = None,
    ):
        for data, states in ((train_data, train_states), (test_data, test_states)):
            assert (data is not None) == (states is not None), \
                "Please provide both dataset and variational states, or neither"
        train_data = TVODataLoader(train_data) if isinstance(train_data, to.Tensor) else train_data
        test_data = TVODataLoader(test_data) if isinstance(test_data, to.Tensor) else test_data
        self.can_train = train_data is not None and train_states is not None
        self.can_test = test_data is not None and test_states is not None
        if not self.can_train and not self.can_test:  # pragma: no cover
            raise RuntimeError("Please provide at least one pair of dataset and variational states")
        _d, _s = (train_data, train_states) if self.can_train else (test_data, test_states)
        if isinstance(model, Optimized):
            model.init_storage(_s.config["S"], _s.config["S_new"], _d.batch_size)
        self.model = model
        self.train_data, self.train_states = train_data, train_states
        self.test_data, self.test_states = test_data, test_states
        self.will_reconstruct = will_reconstruct
        self.eval_F_at_epoch_end = eval_F_at_epoch_end
        if train_data is not None:
            n = to.tensor(len(train_data.dataset), device=tvo_b200.get_device())
            all_reduce(n)
            self.N_train = n.item()
            if self.will_reconstruct:
                self.train_reconstruction = train_data.dataset.tensors[1].clone()
        if test_data is not None:
            n = to.tensor(len(test_data.dataset), device=tvo_b200.get_device())
            all_reduce(n)
            self.N_test = n.item()
            if self.will_reconstruct:
                self.test_reconstruction = test_data.dataset.tensors[1].clone()
        self._to_rollback = rollback_if_F_decreases
        self.data_transform = data_transform if data_transform is not None else lambda x: x

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _do_e_step(data, states, model, N, data_transform, reconstruction=None):
        """Trainer.py:88-116"""
        if reconstruction is not None and not isinstance(model, Reconstructor):
            raise NotImplementedError(f"reconstruction not implemented for model {type(model).__name__}")
        dev = tvo_b200.get_device()
        if isinstance(model, Optimized):
            model.init_epoch()
        if _fast(model, states):
            F = to.zeros((), dtype=to.float64, device=dev)
            for idx, batch in data:
                batch = data_transform(batch)
                model.init_batch()
                states.update_device(idx, batch, model)
                model._free_energy_device(idx, batch, states, F)
                if reconstruction is not None:
                    reconstruction[idx] = model.data_estimator(idx, batch, states).to(reconstruction.dtype)
            fs = to.stack([F, states._nsubs_dev.to(to.float64)])
            all_reduce(fs)
            fs = fs.tolist()
            local = states._take_nsubs()  # keep the host-side cursor in step  # noqa: F841
            return fs[0] / N, _global_subs(states, fs[1]) / N, reconstruction
        F = to.tensor(0.0, dtype=to.float64, device=dev)
        subs = to.tensor(0, device=dev)
        for idx, batch in data:
            batch = data_transform(batch)
            if isinstance(model, Optimized):
                model.init_batch()
            subs += states.update(idx, batch, model)
            F += model.free_energy(idx, batch, states)
            if reconstruction is not None:
                reconstruction[idx] = model.data_estimator(idx, batch, states)
        all_reduce(F)
        all_reduce(subs)
        return F.item() / N, subs.item() / N, reconstruction

    def e_step(self, compute_reconstruction: bool = False) -> Dict[str, Any]:
        """Trainer.py:118-164"""
        ret = {}
        train_rec = self.train_reconstruction if (compute_reconstruction and hasattr(self, "train_reconstruction")) else None
        test_rec = self.test_reconstruction if (compute_reconstruction and hasattr(self, "test_reconstruction")) else None
        if self.can_train:
            ret["train_F"], ret["train_subs"], rec = self._do_e_step(
                self.train_data, self.train_states, self.model, self.N_train, self.data_transform, train_rec)
            if rec is not None:
                ret["train_rec"] = rec
        if self.can_test:
            ret["test_F"], ret["test_subs"], rec = self._do_e_step(
                self.test_data, self.test_states, self.model, self.N_test, self.data_transform, test_rec)
            if rec is not None:
                ret["test_rec"] = rec
        return ret

    def em_step(self, compute_reconstruction: bool = False) -> Dict[str, Any]:
        """Trainer.py:166-217.  F reported for epoch X uses the K sets of epoch X and theta of X-1."""
        ret = {}
        if self.can_train:
            F, subs, reco = self._train_epoch(compute_reconstruction)
            ret["train_F"] = F / self.N_train
            ret["train_subs"] = subs / self.N_train
            if reco is not None:
                ret["train_rec"] = reco
        if self.can_test:
            test_rec = self.test_reconstruction if (compute_reconstruction and hasattr(self, "test_reconstruction")) else None
            res = self._do_e_step(self.test_data, self.test_states, self.model, self.N_test, self.data_transform, test_rec)
            ret["test_F"], ret["test_subs"], _ = res
            if test_rec is not None:
                ret["test_rec"] = test_rec
        return ret

    def _train_epoch(self, compute_reconstruction: bool):
        """Trainer.py:219-255; returns GLOBAL (F, subs) as python numbers."""
        model, data, states = self.model, self.train_data, self.train_states
        rec = self.train_reconstruction if (compute_reconstruction and hasattr(self, "train_reconstruction")) else None
        dev = tvo_b200.get_device()
        fast = _fast(model, states) and hasattr(model, "_stats")
        F = to.zeros((), dtype=to.float64, device=dev)
        subs = to.tensor(0, device=dev)
        if isinstance(model, Optimized):
            model.init_epoch()
        for idx, batch in data:
            batch = self.data_transform(batch)
            if isinstance(model, Optimized):
                model.init_batch()
            with to.no_grad():
                if fast:
                    states.update_device(idx, batch, model)
                else:
                    subs += states.update(idx, batch, model)
                if rec is not None:
                    assert isinstance(model, Reconstructor)
                    rec[idx] = model.data_estimator(idx, batch, states).to(rec.dtype)
            if batch.is_floating_point() and (rec is not None) and to.isnan(batch).any():
                mask = to.isnan(batch)
                batch[mask] = rec[idx][mask]
                rec[idx] = batch
            batch_F = model.update_param_batch(idx, batch, states)
            if not self.eval_F_at_epoch_end and not fast:
                if batch_F is None:
                    batch_F = model.free_energy(idx, batch, states)
                F += batch_F
        self._update_parameters_with_rollback()
        if fast:
            # F rode along in the packed statistics and is already reduced over ranks
            Fg = 0.0 if self.eval_F_at_epoch_end else float(model._last_F.item())
            s = states._nsubs_dev.to(to.float64).reshape(1).clone()
            all_reduce(s)
            states._take_nsubs()
            return Fg, _global_subs(states, s.item()), rec
        all_reduce(F)
        all_reduce(subs)
        return F.item(), subs.item(), rec

    def eval_free_energies(self) -> Dict[str, Any]:
        """Trainer.py:257-299: re-score K under the current theta and report F; no training."""
        m = self.model
        lpj_fn = m.log_pseudo_joint if isinstance(m, Optimized) else m.log_joint
        ret = {}
        for name, data, states, N in (("train", self.train_data, self.train_states, getattr(self, "N_train", None)),
                                      ("test", self.test_data, self.test_states, getattr(self, "N_test", None))):
            if data is None or states is None:
                continue
            F = to.zeros((), dtype=to.float64, device=tvo_b200.get_device())
            if isinstance(m, Optimized):
                m.init_epoch()
            for idx, batch in data:
                batch = self.data_transform(batch)
                if isinstance(m, Optimized):
                    m.init_batch()
                if _fast(m, states):
                    idx = idx.to(states.lpj.device)
                    m._tvo_lpj_rows(states, idx, batch)
                    m._free_energy_device(idx, batch, states, F)
                else:
                    states.lpj[idx] = lpj_fn(batch, states.K[idx])
                    F += m.free_energy(idx, batch, states)
            all_reduce(F)
            ret[f"{name}_F"] = F.item() / N
            ret[f"{name}_subs"] = 0
        return ret

    def _update_parameters_with_rollback(self) -> None:
        """Trainer.py:301-327"""
        if len(self._to_rollback) == 0:
            self.model.update_param_epoch()
            return
        m = self.model
        lpj_fn = m.log_pseudo_joint if isinstance(m, Optimized) else m.log_joint
        all_data = self.train_data.dataset.tensors[1]
        states = self.train_states
        ar = to.arange(all_data.shape[0], device=states.lpj.device)
        old_params = {p: m.theta[p].clone() for p in self._to_rollback}
        old_F = to.tensor(m.free_energy(idx=ar, batch=all_data, states=states), dtype=to.float64,
                          device=tvo_b200.get_device())
        all_reduce(old_F)
        old_lpj = states.lpj.clone()
        m.update_param_epoch()
        states.lpj[:] = lpj_fn(all_data, states.K)
        new_F = to.tensor(m.free_energy(idx=ar, batch=all_data, states=states), dtype=to.float64,
                          device=tvo_b200.get_device())
        all_reduce(new_F)
        if new_F < old_F:
            for p in self._to_rollback:
                m.theta[p][:] = old_params[p]
            states.lpj[:] = old_lpj


def _global_subs(states, total_now: float) -> float:
    """Substitutions of THIS epoch over all ranks: the device counters are cumulative."""
    prev = getattr(states, "_global_subs_prev", 0.0)
    states._global_subs_prev = total_now
    return total_now - prev
