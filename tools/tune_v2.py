#!/usr/bin/env python
"""Time the fused BSC E-step at the headline shape for several settings of tvo_bsc_estep_variant (tuning aid).
usage (GPU box): python tools/tune_v2.py [N] [variant ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch as to
import tvo_b200
from tvo_b200 import _lib
from tvo_b200.models import BSC
from tvo_b200.variational import EVOVariationalStates

N = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
PREC = to.float32 if os.environ.get("PREC") == "f32" else to.float64
variants = [int(v, 0) for v in sys.argv[2:]] or [0]
H, D, S = 256, 144, 64
dev = to.device("cuda:0")
tvo_b200._set_device(dev)
g0 = to.Generator(device=dev).manual_seed(1234)
Wgen = to.randn(D, H, dtype=to.float64, device=dev, generator=g0)
W0 = Wgen + 0.5 * to.randn(D, H, dtype=to.float64, device=dev, generator=g0)
s = (to.rand(N, H, device=dev, generator=g0) < 5.0 / H).double()
Y = (s @ Wgen.t() + to.randn(N, D, dtype=to.float64, device=dev, generator=g0)).to(PREC)
del s
idx = to.arange(N, device=dev)
for v in variants:
    np.random.seed(0)
    _lib.call("tvo_bsc_estep_variant", v)
    model = BSC(H, D, W0.to(PREC), to.tensor([1.0], dtype=PREC), to.full((H,), 1.0 / H, dtype=PREC), precision=PREC)
    if os.environ.get("CHUNK"):
        model.scratch_datapoints = int(os.environ["CHUNK"])
    st = EVOVariationalStates(N, H, S, PREC, "randparents", "sparseflip", n_parents=16, n_generations=2, n_children=2,
                              bitflip_frequency=1.0 / H, seed=2024)
    for _ in range(6):  # same warm-up trajectory as bench.py (EM iterations from the 1/H init)
        st.update_device(idx, Y, model); model.update_param_batch(idx, Y, st); model.update_param_epoch()
    to.cuda.synchronize()
    _lib.kernel_timing_enable(True)
    ev = [to.cuda.Event(enable_timing=True) for _ in range(4)]
    tot = totm = tote = 0.0
    for _ in range(6):
        ev[0].record(); st.update_device(idx, Y, model); ev[1].record()
        model.update_param_batch(idx, Y, st); ev[2].record(); model.update_param_epoch(); ev[3].record()
        to.cuda.synchronize(); tot += ev[0].elapsed_time(ev[1]); totm += ev[1].elapsed_time(ev[2]); tote += ev[2].elapsed_time(ev[3])
    k_ms, k_n = _lib.kernel_timing_read(_lib.TIMED_ESTEP_EVO_BSC)
    g_ms, g_n = _lib.kernel_timing_read(_lib.TIMED_BSC_GEMM_YW)
    m_ms, _ = _lib.kernel_timing_read(_lib.TIMED_BSC_MSTEP)
    ge_ms, _ = _lib.kernel_timing_read(_lib.TIMED_BSC_GEMM_ETY)
    print(f"   M-step {totm / 6:.3f} ms (kernel {m_ms / 6:.3f}, own E^T Y GEMM {ge_ms / 6:.3f}), epoch update {tote / 6:.3f} ms", flush=True)
    _lib.kernel_timing_enable(False)
    print(f"variant {v:#x}: E-step {tot / 6:.3f} ms ({N / (tot / 6) / 1e3:.1f} M dp/s), kernel {k_ms / 6:.3f} ms/step, "
          f"own GEMM {g_ms / 6:.3f} ms/step ({g_n} launches), nsubs/dp {st._take_nsubs() / N / 6:.2f}, "
          f"F/N {float(model._last_F) / N:.9f}", flush=True)
_lib.call("tvo_bsc_estep_variant", 0)
