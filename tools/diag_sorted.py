import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import numpy as np, torch as to
import tvo_b200, workloads
dev = to.device("cuda:0"); tvo_b200._set_device(dev)
N = 200000
wl = workloads.make_c2(dev, N, seed=1234, states_seed=2024)
m, st, Y = wl["model"], wl["states"], wl["Y"]
idx = to.arange(N, device=dev)
for it in range(16):
    st.update_device(idx, Y, m); m.update_param_batch(idx, Y, st); m.update_param_epoch()
    l = m._tvo_lpj_packed(Y, st._Kp, 256)
    srt = (l[:, 1:] >= l[:, :-1]).all(dim=1).float().mean().item()
    inv = (l[:, 1:] < l[:, :-1]).float().sum(dim=1).mean().item()
    # survivors: children beating worst old are not observable here; report nsubs instead
    print(it, f"rows still sorted after the theta update: {srt:.3f}, mean adjacent inversions {inv:.1f}, nsubs/dp {st._take_nsubs()/N:.1f}, |s| {workloads.mean_active_units(st._Kp):.2f}", flush=True)
