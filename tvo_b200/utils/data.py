"""Batch iterator over device-resident tensors with the interface of the reference's TVODataLoader
(tvo/utils/data.py:14-58): `TVODataLoader(*data, batch_size=..., shuffle=...)` yields `(idx, batch, ...)`.

The reference builds a torch DataLoader over a TensorDataset, i.e. it collates every batch from single rows on the
host.  Here the data already live in HBM next to K and lpj, so a batch is one indexing operation on the device (a
zero-copy slice when the order is contiguous) and nothing is collated.  What the Trainer and user code read from a
loader is kept: iteration, `len()`, `.batch_size`, `.precision`, `.dataset.tensors` (index tensor first) and
`len(.dataset)`.  The Trainer accepts the reference's own loader class as well (it only iterates and reads
those attributes).
"""
from typing import Iterator, Optional, Sequence, Tuple

import torch as to
import torch.distributed as dist

import tvo_b200
from tvo_b200.utils.parallel import broadcast


class _ResidentTensors:
    """`.tensors` / `len()` of the loader's data, index tensor first (what `loader.dataset` exposes)."""

    def __init__(self, tensors: Tuple[to.Tensor, ...]):
        self.tensors = tensors

    def __len__(self) -> int:
        return self.tensors[0].shape[0]

    def __getitem__(self, i):
        return tuple(t[i] for t in self.tensors)


class TVODataLoader:
    def __init__(self, *data: to.Tensor, batch_size: Optional[int] = 1, shuffle: Optional[bool] = False,
                 sampler: Optional[Sequence[int]] = None, drop_last: bool = False, **unused):
        N = data[0].shape[0]
        assert all(d.shape[0] == N for d in data), "Dimension mismatch in data sets."
        if data[0].dtype is not to.uint8:
            self.precision = data[0].dtype
        self.dataset = _ResidentTensors((to.arange(N, device=data[0].device),) + tuple(data))
        self.batch_size = N if batch_size is None else int(batch_size)
        self.shuffle, self.sampler, self.drop_last = bool(shuffle), sampler, drop_last
        # Data-parallel runs: every rank must walk the same number of batches because the M-step ends in a
        # collective (data.py:39-58).  Shards differ by at most one row (parallel.py:148-166); the short ranks
        # revisit rows to reach the longest shard's length, in shuffled order like the reference.
        self._n_samples = N
        if tvo_b200.get_run_policy() == "mpi" and dist.is_initialized() and sampler is None:
            n = to.tensor(N, device=data[0].device if data[0].is_cuda else "cpu")
            broadcast(n, src=dist.get_world_size() - 1)
            self._n_samples = int(n)
            self.shuffle = True

    def _order(self) -> Optional[to.Tensor]:
        """Row order of one pass, or None for the identity (contiguous slices)."""
        N = len(self.dataset)
        dev = self.dataset.tensors[0].device
        if self.sampler is not None:
            return to.as_tensor(list(iter(self.sampler)), dtype=to.int64, device=dev)
        if not self.shuffle and self._n_samples == N:
            return None
        order = to.randperm(N, device=dev) if self.shuffle else to.arange(N, device=dev)
        if self._n_samples > N:
            # the revisited rows come from the FRONT of the order and go to its end: a row meets its own copy in one
            # batch only if the whole pass is a single batch (the E-step kernels update K[idx] in place, one warp per
            # batch row, so two rows of a batch should not share an idx)
            extra = order[to.arange(self._n_samples - N, device=dev) % N]
            order = to.cat([order, extra])
        return order[: self._n_samples]

    def __iter__(self) -> Iterator[Tuple[to.Tensor, ...]]:
        order = self._order()
        n = self._n_samples if order is None else order.numel()
        bs = self.batch_size
        for b0 in range(0, n, bs):
            b1 = min(n, b0 + bs)
            if self.drop_last and b1 - b0 < bs:
                return
            if order is None:
                yield tuple(t[b0:b1] for t in self.dataset.tensors)
            else:
                sel = order[b0:b1]
                yield tuple(t[sel] for t in self.dataset.tensors)

    def __len__(self) -> int:
        n = self._n_samples if self.sampler is None else len(self.sampler)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size


class ShufflingSampler:
    """Index sampler of the reference's interface (tvo/utils/data.py:61-89): `n_samples` shuffled indices of the
    dataset, rows revisited when n_samples exceeds its length.  Usable as `TVODataLoader(..., sampler=...)`."""

    def __init__(self, dataset, n_samples: int = None):
        self._len = len(dataset)
        self.n_samples = self._len if n_samples is None else int(n_samples)

    def __iter__(self):
        order = to.randperm(self._len)
        if self.n_samples > self._len:
            order = to.cat([order, order[to.arange(self.n_samples - self._len) % self._len]])
        return iter(order[: self.n_samples].tolist())

    def __len__(self) -> int:
        return self.n_samples
