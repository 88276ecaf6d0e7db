"""Quick device timing of the fused E-step + M-step at the BASELINE shape (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch as to
import tvo_b200
from tvo_b200.models import BSC
from tvo_b200.variational import EVOVariationalStates
tvo_b200._set_device(to.device("cuda:0"))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
H, D, S = 256, 144, 64
prec = to.float64 if (len(sys.argv) < 3 or sys.argv[2] == "f64") else to.float32
g = to.Generator(device="cuda:0").manual_seed(1)
Wgen = to.randn(D, H, dtype=to.float64, device="cuda:0", generator=g)
s = (to.rand(N, H, device="cuda:0", generator=g) < 5.0 / H).double()
Y = (s @ Wgen.t() + to.randn(N, D, dtype=to.float64, device="cuda:0", generator=g)).to(prec)
m = BSC(H, D, Wgen + 0.5 * to.randn(D, H, dtype=to.float64, device="cuda:0", generator=g), to.tensor([1.0]), to.full((H,), 1.0 / H), precision=prec)
np.random.seed(0)
st = EVOVariationalStates(N, H, S, prec, "randparents", "sparseflip", 16, 2, 2, bitflip_frequency=1.0 / H, seed=7)
idx = to.arange(N, device="cuda:0")
ev = [to.cuda.Event(enable_timing=True) for _ in range(4)]
for it in range(6):
    ev[0].record(); st.update_device(idx, Y, m); ev[1].record(); m.update_param_batch(idx, Y, st); ev[2].record(); m.update_param_epoch(); ev[3].record()
    to.cuda.synchronize()
    print("it %d: E-step %.2f ms (%.1f M dp/s)  M-step %.2f ms  epoch %.2f ms  F/N %.4f  subs/N %.2f" % (
        it, ev[0].elapsed_time(ev[1]), N / ev[0].elapsed_time(ev[1]) / 1e3, ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]),
        float(m._last_F) / N, st._take_nsubs() / N))
