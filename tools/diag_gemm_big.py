"""Own GEMMs (a = Y W, WpT += E^T Y) against the library GEMM path at sizes where every CTA walks many tiles / slabs
(the unit tests stop at 4099 datapoints = one tile per CTA): stored lpj vs a direct re-score, K and M-step statistics
between the two paths.  usage: python tools/diag_gemm_big.py"""
import sys, numpy as np, torch as to
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import tvo_b200, workloads
from tvo_b200 import _lib
dev = to.device("cuda:0"); tvo_b200._set_device(dev)
for N in (9473, 70000, 300000, 1000000):
    res = {}
    for variant in (0, 0x10000):
        _lib.call("tvo_bsc_estep_variant", variant)
        wl = workloads.make_c2(dev, N)
        m, st, Y = wl["model"], wl["states"], wl["Y"]
        idx = to.arange(N, device=dev)
        st.update(idx, Y, m)
        full = m._tvo_lpj_packed(Y, st._Kp, 256)
        bad = ((full - st.lpj).abs() > 1e-9 * full.abs().clamp_min(1.0)).any(dim=1)
        m._stats.zero_()
        m.update_param_batch(idx, Y, st)
        res[variant] = (st._Kp.clone(), int(bad.sum()), bad.nonzero().flatten()[:8].tolist(), m._stats.clone())
    _lib.call("tvo_bsc_estep_variant", 0)
    print(N, "own gemm: bad rows", res[0][1], res[0][2], "| library gemm: bad rows", res[0x10000][1], "| K equal", bool(to.equal(res[0][0], res[0x10000][0])),
          "| M-step statistics own vs library GEMM: max rel diff",
          float(((res[0][3] - res[0x10000][3]).abs() / res[0x10000][3].abs().clamp_min(1e-3)).max()))
