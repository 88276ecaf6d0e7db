"""On-disk formats either side of the hot path.

The reference reads datasets and initial states from HDF5 (parallel.py:292-306, TVOVariationalStates.py:30-38)
and logs K / lpj / theta through H5Logger (_experiments.py:179-238).  h5py is not in this image, so the same
arrays travel as `.npz` (or `.npy` for a bare dataset) under the reference's dataset names; when h5py is
importable the `.h5` / `.hdf5` spellings work too.  States are stored in the reference's uint8 (N,S,H) layout
and converted to / from the packed uint64 layout on the device.
"""
import os
from typing import Dict, Optional

import numpy as np
import torch as to

import tvo_b200
from tvo_b200.utils.parallel import gather_from_processes, my_shard


def _is_h5(path: str) -> bool:
    return os.path.splitext(path)[1].lower() in (".h5", ".hdf5")


def read_arrays(path: str, *names: str) -> Dict[str, np.ndarray]:
    """Named arrays of an .npz / .npy / .h5 file.  A bare .npy answers to any single requested name."""
    if _is_h5(path):
        try:
            import h5py
        except ImportError as e:  # pragma: no cover
            raise RuntimeError(f"{path}: reading HDF5 needs h5py, which is not installed; use .npz") from e
        with h5py.File(path, "r") as f:
            if names:
                return {n: f[n][...] for n in names if n in f}
            out: Dict[str, np.ndarray] = {}
            f.visititems(lambda n, o: out.__setitem__(n, o[...]) if isinstance(o, h5py.Dataset) else None)
            return out
    if path.endswith(".npy"):
        assert len(names) <= 1, ".npy holds one array"
        return {(names[0] if names else "data"): np.load(path, allow_pickle=False)}
    with np.load(path, allow_pickle=False) as z:
        return {n: z[n] for n in (names or z.files) if n in z.files}


def load_shard(path: str, name: str = "data", dtype: Optional[to.dtype] = None) -> to.Tensor:
    """This rank's contiguous chunk of dataset `name` on its device (the reference's
    scatter_to_processes layout, parallel.py:148-189, read directly instead of scattered from rank 0)."""
    arrs = read_arrays(path, name)
    if name not in arrs:
        raise KeyError(f"{path} has no dataset '{name}'")
    a = arrs[name]
    lo, hi = my_shard(a.shape[0])
    t = to.from_numpy(np.ascontiguousarray(a[lo:hi])).to(tvo_b200.get_device())
    return t if dtype is None else t.to(dtype)


def load_initial_states(path: str, N: int, S: int, H: int) -> to.Tensor:
    """uint8 (N_local,S,H) initial states from `initial_states` (or `states`), TVOVariationalStates.py:30-38."""
    arrs = read_arrays(path, "initial_states", "states")
    key = "initial_states" if "initial_states" in arrs else "states"
    if key not in arrs:
        raise KeyError(f"{path} has neither 'initial_states' nor 'states'")
    a = arrs[key]
    lo, hi = my_shard(a.shape[0])
    K = to.from_numpy(np.ascontiguousarray(a[lo:hi])).to(dtype=to.uint8)
    assert tuple(K.shape) == (N, S, H), f"{path}:{key} has shape {tuple(K.shape)}, expected {(N, S, H)}"
    return K


def _checkpoint_path(path: str) -> str:
    """The file save_checkpoint writes for `path` (and load_checkpoint reads): .h5 / .hdf5 as given, anything else is
    an .npz archive (the suffix is appended when missing)."""
    return path if _is_h5(path) or path.endswith(".npz") else path + ".npz"


def save_checkpoint(path: str, model=None, states=None, extra: Optional[Dict[str, np.ndarray]] = None) -> None:
    """theta, K (uint8, gathered over ranks) and lpj under the names the reference's H5Logger writes
    (_experiments.py:200-207, H5Logger.py:94-123): 'theta/<name>' (an HDF5 group; the same slash-separated key in an
    .npz), 'train_states', 'train_lpj'.  Rank 0 writes.  The tvo.exp layer itself (ExpConfig, Training / Testing,
    H5Logger) is out of scope; this is the file format only."""
    out: Dict[str, np.ndarray] = dict(extra or {})
    if model is not None:
        for k, v in model.theta.items():
            out["theta/" + k] = v.detach().cpu().numpy()
    if states is not None:
        K = gather_from_processes(states.K.unpack())
        lpj = gather_from_processes(states.lpj)
        if K is not None:
            out["train_states"] = K.cpu().numpy()
            out["train_lpj"] = lpj.cpu().numpy()
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_rank() != 0:
        return
    if _is_h5(path):
        try:
            import h5py
        except ImportError as e:  # pragma: no cover
            raise RuntimeError(f"{path}: writing HDF5 needs h5py, which is not installed; use .npz") from e
        with h5py.File(path, "w") as f:
            for k, v in out.items():
                f.create_dataset(k, data=v)
        return
    tmp = path + ".tmp.npz"
    np.savez(tmp, **out)
    os.replace(tmp, _checkpoint_path(path))


def load_checkpoint(path: str, model=None, states=None) -> Dict[str, np.ndarray]:
    """Inverse of save_checkpoint: theta written in place (model.theta tensors keep their identity), K packed
    onto the device, lpj restored.  Returns every array of the file."""
    arrs = read_arrays(_checkpoint_path(path))
    if model is not None:
        for k, v in model.theta.items():
            src = arrs.get("theta/" + k, arrs.get(k))  # the reference's layout, or the flat names of older files
            if src is not None:
                v.copy_(to.from_numpy(np.asarray(src)).to(device=v.device, dtype=v.dtype))
    if states is not None and "train_states" in arrs:
        lo, hi = my_shard(arrs["train_states"].shape[0])
        states.K = to.from_numpy(np.ascontiguousarray(arrs["train_states"][lo:hi]))
        if "train_lpj" in arrs:
            states.lpj.copy_(to.from_numpy(arrs["train_lpj"][lo:hi]).to(device=states.lpj.device, dtype=states.lpj.dtype))
    return arrs
