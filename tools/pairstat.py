import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch as to
import tvo_b200
from tvo_b200.models import BSC
from tvo_b200.variational import EVOVariationalStates
tvo_b200._set_device(to.device("cuda:0"))
N, H, D, S = 50000, 256, 144, 64
g = to.Generator(device="cuda:0").manual_seed(1)
Wgen = to.randn(D, H, dtype=to.float64, device="cuda:0", generator=g)
s = (to.rand(N, H, device="cuda:0", generator=g) < 5.0 / H).double()
Y = s @ Wgen.t() + to.randn(N, D, dtype=to.float64, device="cuda:0", generator=g)
m = BSC(H, D, Wgen + 0.5 * to.randn(D, H, dtype=to.float64, device="cuda:0", generator=g), to.tensor([1.0]), to.full((H,), 1.0 / H), precision=to.float64)
np.random.seed(0)
st = EVOVariationalStates(N, H, S, to.float64, "randparents", "sparseflip", 16, 2, 2, bitflip_frequency=1.0 / H, seed=7)
idx = to.arange(N, device="cuda:0")
for it in range(41):
    st.update_device(idx, Y, m); m.update_param_batch(idx, Y, st); m.update_param_epoch()
    if it in (0, 3, 12, 40):
        K = st.K.unpack()[:300].cpu().numpy().astype(np.int64)  # (300,S,H)
        inst = dist = uni = 0
        for n in range(300):
            pc = K[n].sum(1)
            inst += int((pc * (pc + 1) // 2).sum())
            A = K[n].T @ K[n]
            dist += int((np.triu(A) > 0).sum())
            uni += int((K[n].sum(0) > 0).sum())
        print(f"it {it}: pair instances/dp {inst/300:.0f}  distinct pairs/dp {dist/300:.0f}  union bits/dp {uni/300:.1f}  mean |s| {K.sum(2).mean():.2f} F/N {float(m._last_F)/N:.2f}")
