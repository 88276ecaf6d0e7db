"""Trainable / Optimized / Sampler / Reconstructor protocols (mirror of
tvo/utils/model_protocols.py:15-266).  Third-party models written against the reference's
protocols work unchanged with the tvo_b200 variational states (generic path)."""
from abc import abstractmethod
from typing import Any, Dict, Optional, Tuple, Union, TYPE_CHECKING

import torch as to
from typing_extensions import Protocol, runtime_checkable

if TYPE_CHECKING:
    from tvo_b200.variational.TVOVariationalStates import TVOVariationalStates


@runtime_checkable
class Trainable(Protocol):
    _theta: Dict[str, to.Tensor]
    _config: Dict[str, Any] = {}
    _optimizer: Optional[to.optim.Optimizer] = None

    @abstractmethod
    def log_joint(self, data: to.Tensor, states: to.Tensor) -> to.Tensor:
        ...

    def update_param_batch(self, idx: to.Tensor, batch: to.Tensor, states: "TVOVariationalStates") -> Optional[float]:
        """Default: one Adam step on -F/B (model_protocols.py:37-66)."""
        if self._optimizer is None:
            for t in self._theta.values():
                t.requires_grad_(True)
            self._optimizer = to.optim.Adam(self._theta.values())
        log_joints = self.log_joint(batch, states.K[idx])
        F = to.logsumexp(log_joints, dim=1).sum(dim=0)
        loss = -F / batch.shape[0]
        loss.backward()
        self._optimizer.step()
        self._optimizer.zero_grad()
        return F.item()

    def update_param_epoch(self) -> None:
        return

    def free_energy(self, idx: to.Tensor, batch: to.Tensor, states: "TVOVariationalStates") -> float:
        log_joints = states.lpj[idx]
        return to.logsumexp(log_joints, dim=1).sum(dim=0).item()

    @property
    def shape(self) -> Tuple[int, ...]:
        if hasattr(self, "_shape"):
            return getattr(self, "_shape")
        th = list(self.theta.values())
        return (th[-1].shape[-1], th[0].shape[0])

    @property
    def config(self) -> Dict[str, Any]:
        return self._config

    @property
    def theta(self) -> Dict[str, to.Tensor]:
        return self._theta

    @property
    def precision(self) -> to.dtype:
        if hasattr(self, "_precision"):
            return getattr(self, "_precision")
        prec = None
        for dt in (p.dtype for p in self.theta.values() if p.dtype.is_floating_point):
            assert prec is None or dt == prec
            prec = dt
        return prec


@runtime_checkable
class Optimized(Trainable, Protocol):
    @abstractmethod
    def log_joint(self, data: to.Tensor, states: to.Tensor, lpj: to.Tensor = None) -> to.Tensor:
        raise NotImplementedError

    @abstractmethod
    def log_pseudo_joint(self, data: to.Tensor, states: to.Tensor) -> to.Tensor:
        ...

    def free_energy(self, idx: to.Tensor, batch: to.Tensor, states: "TVOVariationalStates") -> float:
        with to.no_grad():
            log_joints = self.log_joint(batch, states.K[idx], states.lpj[idx])
        return to.logsumexp(log_joints, dim=1).sum(dim=0).item()

    def init_storage(self, S: int, Snew: int, batch_size: int) -> None:
        pass

    def init_epoch(self) -> None:
        pass

    def init_batch(self) -> None:
        pass

    @property
    def sorted_by_lpj(self) -> Dict[str, to.Tensor]:
        return {}


@runtime_checkable
class Sampler(Protocol):
    @abstractmethod
    def generate_data(self, N: int = None, hidden_state: to.Tensor = None
                      ) -> Union[to.Tensor, Tuple[to.Tensor, to.Tensor]]:
        ...


@runtime_checkable
class Reconstructor(Protocol):
    @abstractmethod
    def data_estimator(self, idx: to.Tensor, batch: to.Tensor, states: "TVOVariationalStates") -> to.Tensor:
        ...
