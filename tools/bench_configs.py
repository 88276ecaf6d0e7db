#!/usr/bin/env python
"""Device timing of the EM iteration at the other BASELINE.json configurations:
C3 SSSC D64 H256 S32, C4 NoisyOR D100 H64 S64 crossover fp32, C5 BSC H1024 D256 S128 — plus the headline shape in
fp64 / the stated fp32 mode (c2, c2f32), the single-cause mixtures (gmm, pmm) and a GaussianTVAE (tvae).
Development / reporting aid (bench.py stays the measurement of record for the headline workload).

    python tools/bench_configs.py c2|c2f32|c3|c4|c5|gmm|pmm|tvae [N] [iters]
prints one JSON line with E-step / M-step / epoch times (CUDA events, median of the timed iterations);
TVO_STATE_STATS=1 adds the active-unit histogram of the final K.
"""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch as to

import tvo_b200
from tvo_b200.models import BSC, SSSC, NoisyOR
from tvo_b200.variational import EVOVariationalStates

cfg = sys.argv[1]
dev = to.device("cuda:0")
tvo_b200._set_device(dev)
g = to.Generator(device=dev).manual_seed(3)
np.random.seed(0)


def chunks(N, step=65536):
    for b0 in range(0, N, step):
        yield b0, min(N, b0 + step)


if cfg in ("c2", "c2f32"):
    # the headline shape through this tool, fp64 and the stated fp32 mode (bench.py is the fp64 measurement of record)
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    H, D, S, P, C, G = 256, 144, 64, 16, 2, 2
    prec = to.float64 if cfg == "c2" else to.float32
    Wgen = to.randn(D, H, dtype=to.float64, device=dev, generator=g)
    Y = to.empty(N, D, dtype=prec, device=dev)
    for b0, b1 in chunks(N):
        s = (to.rand(b1 - b0, H, device=dev, generator=g) < 5.0 / H).double()
        Y[b0:b1] = (s @ Wgen.t() + to.randn(b1 - b0, D, dtype=to.float64, device=dev, generator=g)).to(prec)
    model = BSC(H, D, (Wgen + 0.5 * to.randn(D, H, dtype=to.float64, device=dev, generator=g)).to(prec),
                to.tensor([1.0], dtype=prec), to.full((H,), 1.0 / H, dtype=prec), precision=prec)
    states = EVOVariationalStates(N, H, S, prec, "randparents", "sparseflip", P, G, C, bitflip_frequency=1.0 / H, seed=5)
    name = f"C2 BSC H256 D144 S64 EVO(randparents,sparseflip,P16,C2,G2) {'fp64' if cfg == 'c2' else 'fp32'}"
elif cfg == "c5":
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    H, D, S, P, C, G = 1024, 256, 128, 16, 4, 2
    Wgen = to.randn(D, H, dtype=to.float64, device=dev, generator=g)
    Y = to.empty(N, D, dtype=to.float64, device=dev)
    for b0, b1 in chunks(N):
        s = (to.rand(b1 - b0, H, device=dev, generator=g) < 8.0 / H).double()
        Y[b0:b1] = s @ Wgen.t() + to.randn(b1 - b0, D, dtype=to.float64, device=dev, generator=g)
    model = BSC(H, D, Wgen + 0.5 * to.randn(D, H, dtype=to.float64, device=dev, generator=g), to.tensor([1.0]),
                to.full((H,), 1.0 / H), precision=to.float64)
    states = EVOVariationalStates(N, H, S, to.float64, "randparents", "sparseflip", P, G, C, bitflip_frequency=1.0 / H, seed=5)
    name = "C5 BSC H1024 D256 S128 EVO(randparents,sparseflip,P16,C4,G2) fp64"
elif cfg == "c3":
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
    H, D, S, P, C, G = 256, 64, 32, 8, 4, 1
    Wgen = to.randn(D, H, dtype=to.float64, device=dev, generator=g)
    Y = to.empty(N, D, dtype=to.float64, device=dev)
    for b0, b1 in chunks(N):
        s = (to.rand(b1 - b0, H, device=dev, generator=g) < 4.0 / H).double()
        z = to.randn(b1 - b0, H, dtype=to.float64, device=dev, generator=g)
        Y[b0:b1] = (s * z) @ Wgen.t() + (0.1 ** 0.5) * to.randn(b1 - b0, D, dtype=to.float64, device=dev, generator=g)
    model = SSSC(H, D, Wgen + 0.1 * to.randn(D, H, dtype=to.float64, device=dev, generator=g), to.tensor([0.1]),
                 to.zeros(H), to.eye(H), to.full((H,), 4.0 / H), precision=to.float64)
    states = EVOVariationalStates(N, H, S, to.float64, "randparents", "sparseflip", P, G, C, bitflip_frequency=1.0 / H, seed=5)
    name = "C3 SSSC D64 H256 S32 EVO(randparents,sparseflip,P8,C4,G1) fp64"
elif cfg == "c4":
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 500_000
    H, D, S, P, G = 64, 100, 64, 8, 1
    Wb = to.full((D, 20), 0.1, device=dev)
    for i in range(10):  # 20 bars on a 10x10 grid
        Wb.view(10, 10, 20)[i, :, i] = 0.8
        Wb.view(10, 10, 20)[:, i, 10 + i] = 0.8
    Y = to.empty(N, D, dtype=to.uint8, device=dev)
    for b0, b1 in chunks(N):
        s = (to.rand(b1 - b0, 20, device=dev, generator=g) < 0.1).float()
        pr = 1.0 - to.prod(1.0 - Wb[None] * s[:, None, :], dim=2)
        Y[b0:b1] = (to.rand(b1 - b0, D, device=dev, generator=g) < pr).to(to.uint8)
    model = NoisyOR(H, D, to.rand(D, H, device=dev, generator=g), to.full((H,), 1.0 / H), precision=to.float32)
    states = EVOVariationalStates(N, H, S, to.float32, "fitparents_fixed", "randflip", P, G, crossover=True, seed=5)
    name = "C4 NoisyOR D100 H64 S64 EVO(fitparents_fixed,cross_randflip,P8,G1) fp32"
elif cfg in ("gmm", "pmm"):
    from tvo_b200.models import GMM, PMM
    from tvo_b200.variational import FullEMSingleCauseModels
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    H, D = 256, 144
    if cfg == "gmm":
        Wgen = 3 * to.randn(D, H, dtype=to.float64, device=dev, generator=g)
        Y = to.empty(N, D, dtype=to.float64, device=dev)
        for b0, b1 in chunks(N):
            c = to.randint(0, H, (b1 - b0,), device=dev, generator=g)
            Y[b0:b1] = Wgen.t()[c] + to.randn(b1 - b0, D, dtype=to.float64, device=dev, generator=g)
        model = GMM(H, D, Wgen + to.randn(D, H, dtype=to.float64, device=dev, generator=g), to.tensor([2.0]),
                    to.full((H,), 1.0 / H), precision=to.float64)
    else:
        Wgen = 1 + 9 * to.rand(D, H, dtype=to.float64, device=dev, generator=g)
        Y = to.empty(N, D, dtype=to.float64, device=dev)
        for b0, b1 in chunks(N):
            c = to.randint(0, H, (b1 - b0,), device=dev, generator=g)
            Y[b0:b1] = to.poisson(Wgen.t()[c])
        model = PMM(H, D, Wgen * (0.5 + to.rand(D, H, dtype=to.float64, device=dev, generator=g)),
                    to.full((H,), 1.0 / H), precision=to.float64)
    states = FullEMSingleCauseModels(N, H, to.float64)
    name = f"{cfg.upper()} full EM, single cause, D144 C256 fp64"
elif cfg == "tvae":
    # GaussianTVAE with the headline's latent / observable sizes and one hidden layer; mini-batches like the
    # reference's training loop (the M-step is one Adam step per batch)
    from tvo_b200.models import GaussianTVAE
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
    H, H1, D, S, P, C, G = 256, 512, 144, 64, 16, 2, 2
    prec = to.float32
    Wgen = to.randn(D, H, device=dev, generator=g)
    Y = to.empty(N, D, dtype=prec, device=dev)
    for b0, b1 in chunks(N):
        s = (to.rand(b1 - b0, H, device=dev, generator=g) < 5.0 / H).float()
        Y[b0:b1] = s @ Wgen.t() + to.randn(b1 - b0, D, device=dev, generator=g)
    to.manual_seed(0)
    model = GaussianTVAE(shape=(D, H1, H), precision=prec, sigma2_init=1.0)
    states = EVOVariationalStates(N, H, S, prec, "randparents", "sparseflip", P, G, C, bitflip_frequency=1.0 / H, seed=5)
    name = "GaussianTVAE D144 H1=512 H0=256 S64 EVO(randparents,sparseflip,P16,C2,G2) fp32, batches of 8192"
else:
    raise SystemExit("usage: bench_configs.py c2|c2f32|c3|c4|c5|gmm|pmm|tvae [N] [iters]")

iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
idx = to.arange(N, device=dev)
te, tm, tp, Fs = [], [], [], []
for it in range(iters + 2):
    ev = [to.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    if cfg == "tvae":  # mini-batches: E-step and the Adam step alternate per batch (Trainer.py:232-254)
        evs = []
        for b0 in range(0, N, 8192):
            ib = idx[b0:b0 + 8192]
            e3 = [to.cuda.Event(enable_timing=True) for _ in range(3)]
            e3[0].record()
            states.update_device(ib, Y[ib], model)
            e3[1].record()
            model.update_param_batch(ib, Y[ib], states)
            e3[2].record()
            evs.append(e3)
        ev[1].record()
        ev[2].record()
    else:
        states.update_device(idx, Y, model)
        ev[1].record()
        model.update_param_batch(idx, Y, states)
        ev[2].record()
    model.update_param_epoch()
    ev[3].record()
    to.cuda.synchronize()
    if it >= 2 and cfg == "tvae":
        te.append(sum(e[0].elapsed_time(e[1]) for e in evs))
        tm.append(sum(e[1].elapsed_time(e[2]) for e in evs))
        tp.append(ev[2].elapsed_time(ev[3]))
    elif it >= 2:
        te.append(ev[0].elapsed_time(ev[1]))
        tm.append(ev[1].elapsed_time(ev[2]))
        tp.append(ev[2].elapsed_time(ev[3]))
    F = getattr(model, "_last_F", None)
    if cfg == "tvae":
        F = to.logsumexp(states.lpj, dim=1).sum()
    Fs.append(float(F) / N if F is not None else float("nan"))
e, m, p = statistics.median(te), statistics.median(tm), statistics.median(tp)
print(json.dumps({"config": name, "N": N, "estep_ms": e, "mstep_ms": m, "epoch_update_ms": p,
                  "estep_datapoints_per_s": N / e * 1e3, "em_iter_datapoints_per_s": N / (e + m + p) * 1e3,
                  "epoch_update_ms_all": [round(x, 2) for x in tp], "estep_ms_all": [round(x, 2) for x in te],
                  "F_per_datapoint_trajectory": [round(f, 4) for f in Fs], "subs_per_datapoint_last": states._take_nsubs() / N / (iters + 2)}))
if os.environ.get("TVO_STATE_STATS"):  # sparsity of the final K (diagnostic)
    from tvo_b200.variational._packed import unpack
    pc = unpack(states._Kp[:20000], H).sum(dim=2)
    hist = to.bincount(pc.flatten().long(), minlength=12)[:12].float()
    print(json.dumps({"hist_active_units_0_11": [round(float(x), 5) for x in hist / hist.sum()],
                      "mean_active_units": float(pc.float().mean()), "max_active_units": int(pc.max()),
                      "frac_datapoints_with_state_over_7": float((pc.max(dim=1).values > 7).float().mean())}))
