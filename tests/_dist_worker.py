"""Worker for the world_size-2 gloo tests (host-side data-parallel logic, CPU only)."""
import os
import sys

import numpy as np
import torch as to

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out = sys.argv[1]
    import tvo_b200
    from tvo_b200.utils import parallel
    from tvo_b200.models import BSC
    from oracle import models as om

    parallel.init_processes(backend="gloo")
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    assert tvo_b200.get_run_policy() == "mpi"

    # scatter / gather with the reference's uneven chunking (N=7 over 2 ranks -> 3 + 4)
    full = to.arange(7 * 3, dtype=to.float64).reshape(7, 3) if rank == 0 else None
    mine = parallel.scatter_to_processes(full)
    lo, hi = parallel.shard_bounds(7, world, rank)
    assert mine.shape[0] == hi - lo and to.equal(mine, to.arange(7 * 3, dtype=to.float64).reshape(7, 3)[lo:hi])
    back = parallel.gather_from_processes(mine)
    if rank == 0:
        assert to.equal(back, full)

    # packed statistics: every rank accumulates its shard, ONE all-reduce, identical theta everywhere
    rng = np.random.default_rng(0)
    N, D, H, S = 40, 5, 6, 7
    W = rng.standard_normal((D, H))
    pies = rng.uniform(0.1, 0.5, H)
    Y = rng.standard_normal((N, D))
    K = (rng.random((N, S, H)) < 0.4).astype(np.uint8)
    o = om.BSC(H, D, W, [0.8], pies)
    lpj = o.log_pseudo_joint(Y, K)
    F_full = o.free_energy(Y, K, lpj)
    lo, hi = parallel.shard_bounds(N, world, rank)
    osh = om.BSC(H, D, W, [0.8], pies)
    osh.update_param_batch(Y[lo:hi], K[lo:hi], lpj[lo:hi])
    m = BSC(H, D, to.tensor(W), to.tensor([0.8], dtype=to.float64), to.tensor(pies), precision=to.float64)
    m.W_solver = "cpu"
    st = m._stats
    st[: H * D] = to.tensor(osh.my_Wp.T.reshape(-1))
    st[H * D: H * D + H * H] = to.tensor(np.triu(osh.my_Wq).reshape(-1))  # kernel convention: upper triangle
    st[H * D + H * H: H * D + H * H + H] = to.tensor(osh.my_pies)
    st[-3] = float(osh.my_sigma2[0])
    st[-2] = hi - lo
    st[-1] = osh.free_energy(Y[lo:hi], K[lo:hi], lpj[lo:hi])
    m.update_param_epoch()
    o.update_param_batch(Y, K, lpj)
    o.update_param_epoch()
    assert np.allclose(m.theta["W"].numpy(), o.theta["W"], rtol=1e-9, atol=1e-11)
    assert np.allclose(m.theta["pies"].numpy(), o.theta["pies"], rtol=1e-12)
    assert np.allclose(m.theta["sigma2"].numpy(), o.theta["sigma2"], rtol=1e-12)
    assert abs(float(m._last_F) - F_full) < 1e-9 * abs(F_full)
    assert int(m._last_N.item()) == N
    # theta is bit-identical on both ranks (no broadcast needed)
    Wcat = [to.zeros_like(m.theta["W"]) for _ in range(world)]
    dist.all_gather(Wcat, m.theta["W"])
    assert to.equal(Wcat[0], Wcat[1])
    dist.barrier()
    open(f"{out}.{rank}", "w").write("ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
