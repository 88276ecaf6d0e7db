"""TVODataLoader (mirror of tvo/utils/data.py:14-89): yields (idx, batch...) tuples."""
import numpy as np
import torch as to
import torch.distributed as dist
from torch.utils.data import DataLoader, Dataset, Sampler, TensorDataset

import tvo_b200
from tvo_b200.utils.parallel import broadcast


class TVODataLoader(DataLoader):
    def __init__(self, *data: to.Tensor, **kwargs):
        N = data[0].shape[0]
        assert all(d.shape[0] == N for d in data), "Dimension mismatch in data sets."
        if data[0].dtype is not to.uint8:
            self.precision = data[0].dtype
        dataset = TensorDataset(to.arange(N, device=data[0].device), *data)
        if tvo_b200.get_run_policy() == "mpi" and dist.is_initialized() and "sampler" not in kwargs:
            # all ranks must iterate the same number of batches (data.py:39-58)
            n_samples = to.tensor(N, device=data[0].device if data[0].is_cuda else "cpu")
            broadcast(n_samples, src=dist.get_world_size() - 1)
            kwargs["sampler"] = ShufflingSampler(dataset, int(n_samples))
            kwargs["shuffle"] = None
        super().__init__(dataset, **kwargs)


class ShufflingSampler(Sampler):
    def __init__(self, dataset: Dataset, n_samples: int = None):
        self._ds_len = len(dataset)
        self.n_samples = n_samples if n_samples is not None else self._ds_len

    def __iter__(self):
        idxs = np.arange(self._ds_len)
        np.random.shuffle(idxs)
        if self.n_samples > self._ds_len:
            n_extra = self.n_samples - self._ds_len
            extra = np.random.choice(idxs, size=n_extra, replace=n_extra > idxs.size)
            idxs = np.concatenate((idxs, extra))
        else:
            idxs = idxs[: self.n_samples]
        return iter(idxs)

    def __len__(self):
        return self.n_samples
