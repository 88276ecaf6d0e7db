"""Device / run-policy singletons (mirror of tvo/utils/global_settings.py).

Differences from the reference, both forced by the product being GPU-only:
 * the default device is ``cuda:<TVO_GPU | LOCAL_RANK | 0>`` whenever CUDA is available (the
   reference defaults to CPU and needs TVO_GPU); without CUDA it is ``cpu`` so that host logic can
   be imported and unit-tested, but any compute call then fails loudly (no CPU fallback);
 * policy ``'mpi'`` means "data-parallel over torch.distributed" (NCCL on GPUs, gloo in CPU tests);
   it is selected by TVO_MPI / OMPI_COMM_WORLD_SIZE as in the reference, or by WORLD_SIZE > 1
   (torchrun).
"""
import os

import torch as to


def _choose_device() -> to.device:
    if not to.cuda.is_available():
        return to.device("cpu")
    n = os.environ.get("TVO_GPU", os.environ.get("LOCAL_RANK", "0"))
    return to.device(f"cuda:{int(n)}")


class _GlobalDevice:
    _device: to.device = _choose_device()

    @classmethod
    def get_device(cls) -> to.device:
        return cls._device

    @classmethod
    def set_device(cls, dev: to.device):
        cls._device = dev


def get_device() -> to.device:
    return _GlobalDevice.get_device()


def _set_device(dev: to.device):
    _GlobalDevice.set_device(to.device(dev))


def _choose_run_policy() -> str:
    if ("TVO_MPI" in os.environ and os.environ["TVO_MPI"] != "0") or "OMPI_COMM_WORLD_SIZE" in os.environ:
        return "mpi"
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return "mpi"
    return "seq"


class _GlobalPolicy:
    _policy: str = _choose_run_policy()


def get_run_policy() -> str:
    return _GlobalPolicy._policy


def _set_run_policy(policy: str):
    assert policy in ("seq", "mpi")
    _GlobalPolicy._policy = policy
