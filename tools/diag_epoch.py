#!/usr/bin/env python
"""Where the time of BSC.update_param_epoch goes at H=1024 (diagnostic; run under gpurun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as to

dev = to.device("cuda:0")
H, D = 1024, 256
g = to.Generator(device=dev).manual_seed(0)


def timed(name, fn, reps=5):
    ts = []
    for _ in range(reps):
        to.cuda.synchronize()
        t = time.perf_counter()
        out = fn()
        to.cuda.synchronize()
        ts.append((time.perf_counter() - t) * 1e3)
    print(f"{name:50s} " + " ".join(f"{x:8.2f}" for x in ts), flush=True)
    return out


E = (to.rand(200000, H, device=dev, generator=g) < 8.0 / H).double()
Wq = E.t() @ E
Wp = to.randn(H, D, dtype=to.float64, device=dev, generator=g)
timed("lstsq full rank", lambda: to.linalg.lstsq(Wq, Wp)[0])
Wq0 = Wq.clone()
Wq0[5, :] = 0
Wq0[:, 5] = 0
timed("lstsq one zero row/col", lambda: to.linalg.lstsq(Wq0, Wp)[0])
timed("cholesky_ex + cholesky_solve", lambda: to.cholesky_solve(Wp, to.linalg.cholesky_ex(Wq)[0]))
timed("solve (LU)", lambda: to.linalg.solve(Wq, Wp))
timed("randn_like", lambda: to.randn_like(Wp))


def spread(name, fn, reps=200):
    ts = []
    for _ in range(reps):
        to.cuda.synchronize()
        t = time.perf_counter()
        fn()
        to.cuda.synchronize()
        ts.append((time.perf_counter() - t) * 1e3)
    ts.sort()
    print(f"{name:40s} median {ts[len(ts)//2]:.2f}  p90 {ts[int(.9*len(ts))]:.2f}  p99 {ts[int(.99*len(ts))]:.2f}  max {ts[-1]:.2f}", flush=True)


spread("lstsq x200", lambda: to.linalg.lstsq(Wq, Wp)[0])
spread("cholesky x200", lambda: to.cholesky_solve(Wp, to.linalg.cholesky_ex(Wq)[0]))
Wq2 = Wq[:256, :256].contiguous(); Wp2 = Wp[:256, :144].contiguous()
spread("lstsq 256 x200", lambda: to.linalg.lstsq(Wq2, Wp2)[0])
spread("cholesky 256 x200", lambda: to.cholesky_solve(Wp2, to.linalg.cholesky_ex(Wq2)[0]))
print(to.__config__.show().count("MAGMA"), "magma mentions;", to.backends.cuda.preferred_linalg_library())
