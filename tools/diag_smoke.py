import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch as to
import tvo_b200
from tvo_b200 import _lib
from tvo_b200.models import BSC
from tvo_b200.variational import EVOVariationalStates
from oracle import evo as oevo, models as om, states as ost
tvo_b200._set_device(to.device("cuda:0"))
np.random.seed(0)
N, D, H, S = 64, 20, 70, 12
rng = np.random.default_rng(0)
W = rng.standard_normal((D, H)); pies = np.full(H, 3.0 / H)
Y = ((rng.random((N, H)) < pies) @ W.T + rng.standard_normal((N, D)))
model = BSC(H, D, to.tensor(W), to.tensor([1.0]), to.tensor(pies), precision=to.float64)
states = EVOVariationalStates(N, H, S, to.float64, "randparents", "sparseflip", n_parents=4, n_generations=2, n_children=2, bitflip_frequency=1.0 / H, seed=123)
K0 = states.K.unpack().cpu().numpy()
print("popcounts of init states:", K0[0].sum(-1))
sp_dev = float(model._tvo_sparsity().item()); sp_np = float(pies.mean())
print("sparsity dev %r np %r equal %s" % (sp_dev, sp_np, sp_dev == sp_np))
# generation 0 children via the unfused kernel
cfg = states.evo_config()
idx = to.arange(N, device="cuda:0"); W64 = 2
out = to.empty((N, 8, W64), dtype=to.int64, device="cuda:0")
sp = to.tensor([sp_np], dtype=to.float64, device="cuda:0")
_lib.call("tvo_evo_generation_f64", _lib.ptr(states._Kp), S * W64, None, S, _lib.ptr(idx), 1, S, _lib.ptr(out), C.byref(cfg), 0, _lib.ptr(sp), None, N, H, _lib.stream())
got = ost.unpack_states(out.cpu().numpy().view(np.uint64), H)
pidx = oevo.randparents_idx(np.arange(N), S, 4, 123, 0, 0)
want = oevo.sparseflip(oevo.take_parents(K0, pidx), 2, sp_np, 1.0 / H, np.arange(N), 123, 0, 0)
bad = np.argwhere((got != want).any(-1))
print("gen0 children mismatches (n, j):", bad[:10].tolist(), "of", N * 8)
for n, j in bad[:3]:
    par = K0[n, pidx[n, j // 2]]
    print(" n", n, "j", j, "parent popcount", par.sum(), "parent bits", np.flatnonzero(par), "got", np.flatnonzero(got[n, j]), "want", np.flatnonzero(want[n, j]))
    a = par.sum(); print("  p0,p1", oevo.sparseflip_probs(a, H, sp_np, 1.0 / H))

# ---- fused vs unfused vs oracle on the full E-step
import copy
Kp0 = states._Kp.clone()
def run(fused):
    st = EVOVariationalStates(N, H, S, to.float64, "randparents", "sparseflip", n_parents=4, n_generations=2, n_children=2, bitflip_frequency=1.0 / H, seed=123)
    st._Kp.copy_(Kp0)
    Yd = to.tensor(Y, device="cuda:0")
    if fused:
        ns = st.update(idx, Yd, model)
    else:
        cfg = st.evo_config(); st._step += 1
        st._evolve_unfused(idx, cfg, lambda: model._tvo_lpj_rows(st, idx, Yd), lambda ch: model._tvo_lpj_packed(Yd, ch, H), model._tvo_sparsity())
        ns = st._take_nsubs()
    return st.K.unpack().cpu().numpy(), st.lpj.cpu().numpy(), ns
Kf, Lf, nf = run(True)
Ku, Lu, nu = run(False)
omod = om.BSC(H, D, W, [1.0], pies)
K, lpj = K0.copy(), omod.log_pseudo_joint(Y, K0)
new_s, new_l = oevo.evolve_states(lpj, K, lambda s: omod.log_pseudo_joint(Y, s), 4, 2, 2, "randparents", "sparseflip", sp_dev, 1.0 / H, np.arange(N), seed=123, step=0)
raw_l = new_l.copy()
nw = ost.update_states_for_batch(new_s, new_l, np.arange(N), K, lpj)
print("nsubs fused %d unfused %d oracle %d" % (nf, nu, nw))
print("fused==unfused K", np.array_equal(Kf, Ku), " fused==oracle K", np.array_equal(Kf, K), " unfused==oracle K", np.array_equal(Ku, K))
bad = [n for n in range(N) if not np.array_equal(Kf[n], K[n])]
print("datapoints differing (fused vs oracle):", bad[:10], len(bad))
for n in bad[:2]:
    print(" n", n, "\n  lpj fused ", Lf[n], "\n  lpj oracle", lpj[n])
    print("  new_l (oracle, after dedup)", new_l[n])
    print("  children popcounts", new_s[n].sum(-1))
