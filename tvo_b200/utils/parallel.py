"""Data-parallel plumbing (mirror of tvo/utils/parallel.py, NCCL instead of MPI).

The reference scatters the dataset from rank 0 over `torch.distributed` backend "mpi"
(parallel.py:120-189) and all-reduces every sufficient statistic separately (e.g. bsc.py:164-168).
Here each rank owns a contiguous shard that stays resident in its GPU's HBM for the whole run,
theta is replicated, and ONE all-reduce per EM iteration carries the packed fp64 statistics buffer
(models keep all accumulators + N + F in one tensor).  Backend: NCCL over NVLink on GPUs, gloo
for the CPU tests of this host logic.
"""
import math
import os
from typing import Iterable, Tuple, Union

import torch
import torch.distributed as dist
from torch import Tensor

import tvo_b200
from tvo_b200.utils import global_settings


def pprint(obj: object = "", end: str = "\n"):
    """Print on rank 0 only (parallel.py:17-26)."""
    if tvo_b200.get_run_policy() == "mpi" and dist.is_initialized() and dist.get_rank() != 0:
        return
    print(obj, end=end)


def init_processes(multi_node: bool = False, backend: str = None):
    """Join the process group (parallel.py:29-62).  One process per GPU; rank -> cuda:LOCAL_RANK."""
    if dist.is_initialized():
        return
    use_cuda = torch.cuda.is_available()
    backend = backend or ("nccl" if use_cuda else "gloo")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if use_cuda:
        local_rank = int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0"))) % torch.cuda.device_count()
        torch.cuda.set_device(local_rank)
        tvo_b200._set_device(torch.device(f"cuda:{local_rank}"))
        dist.init_process_group(backend, device_id=torch.device(f"cuda:{local_rank}"))
    else:
        dist.init_process_group(backend)
    global_settings._set_run_policy("mpi")


def shard_bounds(total_length: int, comm_size: int, comm_rank: int) -> Tuple[int, int]:
    """[start, stop) of rank's contiguous chunk, the reference's chunking (parallel.py:148-166):
    with no_dummy = ceil(N/size)*size - N, the first `no_dummy` ranks hold one row fewer."""
    assert total_length / comm_size >= 1, "number of data points must be >= number of processes"
    ceiled = math.ceil(total_length / comm_size)
    no_dummy = ceiled * comm_size - total_length
    start = 0
    for r in range(comm_rank):
        start += ceiled - 1 if r < no_dummy else ceiled
    length = ceiled - 1 if comm_rank < no_dummy else ceiled
    return start, start + length


def my_shard(total_length: int) -> Tuple[int, int]:
    if tvo_b200.get_run_policy() == "seq" or not dist.is_initialized():
        return 0, total_length
    return shard_bounds(total_length, dist.get_world_size(), dist.get_rank())


def scatter_to_processes(*tensors: Tensor, src: int = 0) -> Union[Tensor, Iterable[Tensor]]:
    """Rank `src` holds the full tensors (others pass None); every rank gets its chunk on its device
    (parallel.py:120-189).  Implemented as shape/dtype broadcast + per-rank send of the chunk."""
    out = []
    if tvo_b200.get_run_policy() == "seq" or not dist.is_initialized():
        out = list(tensors)
    else:
        size, rank = dist.get_world_size(), dist.get_rank()
        dev = tvo_b200.get_device()
        for data in tensors:
            meta = [None]
            if rank == src:
                meta = [(tuple(data.shape), data.dtype)]
            dist.broadcast_object_list(meta, src=src)
            shape, dtype = meta[0]
            lo, hi = shard_bounds(shape[0], size, rank)
            mine = torch.empty((hi - lo,) + tuple(shape[1:]), dtype=dtype, device=dev)
            if rank == src:
                reqs = []
                for r in range(size):
                    a, b = shard_bounds(shape[0], size, r)
                    chunk = data[a:b].to(dev).contiguous()
                    if r == src:
                        mine.copy_(chunk)
                    else:
                        reqs.append(dist.isend(chunk, dst=r))
                for q in reqs:
                    q.wait()
            else:
                dist.recv(mine, src=src)
            out.append(mine)
    return out[0] if len(out) == 1 else out


def gather_from_processes(*my_tensors: Tensor, dst: int = 0) -> Union[Tensor, Iterable[Tensor]]:
    """Inverse of scatter_to_processes (parallel.py:192-255); only rank `dst` gets the data."""
    out = []
    if tvo_b200.get_run_policy() == "seq" or not dist.is_initialized():
        out = list(my_tensors)
    else:
        size, rank = dist.get_world_size(), dist.get_rank()
        for mine in my_tensors:
            n = torch.tensor([mine.shape[0]], device=mine.device, dtype=torch.int64)
            lens = [torch.zeros_like(n) for _ in range(size)]
            dist.all_gather(lens, n)
            if rank == dst:
                parts = []
                for r in range(size):
                    if r == dst:
                        parts.append(mine)
                    else:
                        buf = torch.empty((int(lens[r].item()),) + tuple(mine.shape[1:]), dtype=mine.dtype,
                                          device=mine.device)
                        dist.recv(buf, src=r)
                        parts.append(buf)
                out.append(torch.cat(parts))
            else:
                dist.send(mine.contiguous(), dst=dst)
                out.append(None)
    return out[0] if len(out) == 1 else out


def all_reduce(tensor: Tensor, op=dist.ReduceOp.SUM):
    """dist.all_reduce under policy 'mpi', no-op otherwise (parallel.py:258-261)."""
    if tvo_b200.get_run_policy() == "mpi" and dist.is_initialized():
        dist.all_reduce(tensor, op)


def broadcast(tensor: Tensor, src: int = 0):
    if tvo_b200.get_run_policy() == "mpi" and dist.is_initialized():
        dist.broadcast(tensor, src)


def barrier():
    if tvo_b200.get_run_policy() == "mpi" and dist.is_initialized():
        dist.barrier()
