#!/usr/bin/env python
"""Per-statement wall time of BSC.update_param_epoch inside a C5 EM loop (diagnostic; run under gpurun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch as to
import tvo_b200
from tvo_b200.models import BSC
from tvo_b200.variational import EVOVariationalStates

dev = to.device("cuda:0")
tvo_b200._set_device(dev)
g = to.Generator(device=dev).manual_seed(3)
np.random.seed(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
H, D, S, P, C, G = 1024, 256, 128, 16, 4, 2
Wgen = to.randn(D, H, dtype=to.float64, device=dev, generator=g)
Y = to.empty(N, D, dtype=to.float64, device=dev)
for b0 in range(0, N, 65536):
    b1 = min(N, b0 + 65536)
    s = (to.rand(b1 - b0, H, device=dev, generator=g) < 8.0 / H).double()
    Y[b0:b1] = s @ Wgen.t() + to.randn(b1 - b0, D, dtype=to.float64, device=dev, generator=g)
model = BSC(H, D, Wgen + 0.5 * to.randn(D, H, dtype=to.float64, device=dev, generator=g), to.tensor([1.0]),
            to.full((H,), 1.0 / H), precision=to.float64)
states = EVOVariationalStates(N, H, S, to.float64, "randparents", "sparseflip", P, G, C, bitflip_frequency=1.0 / H, seed=5)
idx = to.arange(N, device=dev)

import torch.linalg
orig_lstsq = to.linalg.lstsq
marks = {}


def lstsq_timed(*a, **k):
    to.cuda.synchronize()
    t = time.perf_counter()
    r = orig_lstsq(*a, **k)
    to.cuda.synchronize()
    marks["lstsq"] = (time.perf_counter() - t) * 1e3
    return r


to.linalg.lstsq = lstsq_timed
for it in range(12):
    to.cuda.synchronize(); t0 = time.perf_counter()
    states.update_device(idx, Y, model)
    to.cuda.synchronize(); t1 = time.perf_counter()
    model.update_param_batch(idx, Y, states)
    to.cuda.synchronize(); t2 = time.perf_counter()
    model.update_param_epoch()
    to.cuda.synchronize(); t3 = time.perf_counter()
    print(f"it {it}: E {1e3*(t1-t0):7.2f}  M {1e3*(t2-t1):7.2f}  epoch {1e3*(t3-t2):7.2f}  (lstsq {marks.get('lstsq', 0):7.2f})", flush=True)
