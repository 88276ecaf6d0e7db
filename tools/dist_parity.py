"""Multi-GPU parity: N datapoints sharded over the ranks (NCCL) vs the same N on one GPU.
Run:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_parity.py
Because the candidate stream is keyed by the GLOBAL datapoint index, the selected K sets must be
bit-identical; F, subs and theta agree to fp64 reduction-order noise."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch as to, torch.distributed as dist
import tvo_b200
from tvo_b200.utils import parallel, global_settings
from tvo_b200.models import BSC
from tvo_b200.trainer import Trainer
from tvo_b200.utils.data import TVODataLoader
from tvo_b200.variational import EVOVariationalStates

parallel.init_processes()
rank, world = dist.get_rank(), dist.get_world_size()
dev = tvo_b200.get_device()
N, H, D, S = 4000, 64, 36, 16  # divisible by the world size: uneven shards are padded by the
# reference-style sampler (data.py:39-58), which re-visits one datapoint and so changes the statistics
g = to.Generator(device="cpu").manual_seed(5)
Wg = to.randn(D, H, dtype=to.float64, generator=g)
Y = ((to.rand(N, H, generator=g) < 3.0 / H).double() @ Wg.t() + to.randn(N, D, dtype=to.float64, generator=g)).to(dev)
W0 = (Wg + 0.3 * to.randn(D, H, dtype=to.float64, generator=g)).to(dev)
np.random.seed(3)
from tvo_b200.variational._utils import generate_unique_states
K0 = generate_unique_states(S, H, device=dev)[None].repeat(N, 1, 1)


def run(lo, hi, policy):
    global_settings._set_run_policy(policy)
    m = BSC(H, D, W0.clone(), to.tensor([1.0], dtype=to.float64), to.full((H,), 3.0 / H, dtype=to.float64), precision=to.float64)
    st = EVOVariationalStates(hi - lo, H, S, to.float64, "randparents", "sparseflip", 4, 2, 2, bitflip_frequency=1.0 / H,
                              seed=11, K_init=K0[lo:hi])
    st.n_offset = lo
    tr = Trainer(m, TVODataLoader(Y[lo:hi], batch_size=hi - lo), st)
    out = [tr.em_step() for _ in range(3)]
    return m, st, out


lo, hi = parallel.shard_bounds(N, world, rank)
m_d, st_d, out_d = run(lo, hi, "mpi")
if rank == 0:
    m_s, st_s, out_s = run(0, N, "seq")
global_settings._set_run_policy("mpi")
Ks = parallel.gather_from_processes(st_d._Kp)
# theta must be BIT-identical on every rank WITHOUT a broadcast (deterministic hand-written solve on the all-reduced
# statistics, csrc/bsc_chol.cuh): gather every rank's theta and compare
flat = to.cat([m_d.theta[k].reshape(-1) for k in sorted(m_d.theta)]).contiguous()
allflat = [to.empty_like(flat) for _ in range(world)]
dist.all_gather(allflat, flat)
assert all(to.equal(allflat[0], f) for f in allflat), "theta differs between the ranks"
if rank == 0:
    assert to.equal(Ks, st_s._Kp), "sharded K differs from single-GPU K"
    for a, b in zip(out_d, out_s):
        assert abs(a["train_F"] - b["train_F"]) <= 1e-10 * abs(b["train_F"]), (a, b)
        assert abs(a["train_subs"] - b["train_subs"]) < 1e-12, (a, b)
    for k in m_s.theta:
        assert to.allclose(m_d.theta[k], m_s.theta[k], rtol=1e-9, atol=1e-11), k
    print(f"DIST PARITY OK world={world}: K bit-identical, F/subs/theta match; F trajectory", [round(o['train_F'], 6) for o in out_d])
dist.barrier()
dist.destroy_process_group()
