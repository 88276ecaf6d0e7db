"""Kernel timeline of one EM iteration (CUPTI through torch.profiler): start offset, duration and the idle gap before
every kernel, so the fixed per-iteration cost at small shards (strong scaling) can be read.
usage: python tools/timeline.py [N]"""
import os
import sys

import torch as to

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import workloads  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
dev = to.device("cuda:0")
wl = workloads.make_c2(dev, N)
model, states, Y = wl["model"], wl["states"], wl["Y"]
idx = to.arange(N, device=dev)


def em():
    states.update_device(idx, Y, model)
    model.update_param_batch(idx, Y, states)
    model.update_param_epoch()


for _ in range(5):
    em()
to.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        em()
    to.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == to.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
last_end = t0
half = len(ev) // 2
print(f"{len(ev)} device activities in 2 iterations; second iteration:")
tot = 0.0
for e in ev[half:]:
    s, d = e.time_range.start, e.time_range.end - e.time_range.start
    gap = s - last_end
    print(f"{(s - ev[half].time_range.start):9.1f} us  gap {gap:7.1f}  dur {d:8.1f}  {e.name[:90]}")
    last_end = max(last_end, e.time_range.end)
    tot += d
print(f"busy {tot:.1f} us of {ev[-1].time_range.end - ev[half].time_range.start:.1f} us")
