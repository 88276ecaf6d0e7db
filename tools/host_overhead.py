"""Host-side cost of one EM iteration (Python + ctypes + launch calls) at a shard so small that the GPU is never the
limit: what bounds the EM iteration when the datapoints are sharded thinly (strong scaling).
usage: python tools/host_overhead.py [N] [iters] [--profile]"""
import cProfile
import os
import pstats
import sys
import time

import torch as to

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import workloads  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
N = int(args[0]) if args else 2048
iters = int(args[1]) if len(args) > 1 else 300
dev = to.device("cuda:0")
wl = workloads.make_c2(dev, N)
model, states, Y = wl["model"], wl["states"], wl["Y"]
idx = to.arange(N, device=dev)


def em(t):
    t0 = time.perf_counter()
    states.update_device(idx, Y, model)
    t1 = time.perf_counter()
    model.update_param_batch(idx, Y, states)
    t2 = time.perf_counter()
    model.update_param_epoch()
    t3 = time.perf_counter()
    t[0] += t1 - t0
    t[1] += t2 - t1
    t[2] += t3 - t2


t = [0.0, 0.0, 0.0]
for _ in range(20):
    em(t)
to.cuda.synchronize()
t = [0.0, 0.0, 0.0]
for _ in range(iters):
    em(t)
to.cuda.synchronize()
print(f"N={N}: host us per call: update_device {t[0] / iters * 1e6:.1f}  update_param_batch {t[1] / iters * 1e6:.1f}  "
      f"update_param_epoch (incl. its sync) {t[2] / iters * 1e6:.1f}")
if "--profile" in sys.argv:
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(iters):
        em(t)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(22)
