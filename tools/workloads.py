"""Synthetic workloads of the BASELINE.json configurations, shared by bench.py and tools/bench_configs.py.

Every builder returns a dict {model, states, Y, name, H, D, S, alg_bytes_per_dp} with the data, the model (random
init of the configuration's architecture) and the variational states resident on `dev`.  `alg_bytes_per_dp` is the
packed minimum E-step traffic of SURVEY 8(d):  D*w_y + 2*S*ceil(H/64)*8 + S*w.
"""
import numpy as np
import torch as to


def _chunks(N, step=65536):
    for b0 in range(0, N, step):
        yield b0, min(N, b0 + step)


def _alg_bytes(D, wy, S, H, w):
    return D * wy + 2 * S * ((H + 63) // 64) * 8 + S * w


def make_bsc(dev, N, H, D, S, P, C, G, pi_gen, prec=to.float64, seed=3, states_seed=5, n_offset=0, data_seed=None,
             at_generating_parameters=False):
    """C2 / C5: W_gen ~ N(0,1), s ~ Bern(pi_gen), Y = W s + N(0,1); init W = W_gen + 0.5 N(0,1), pi = 1/H, sigma2 = 1;
    EVO randparents + sparseflip, p_bf = 1/H.  `seed` fixes theta (identical on all ranks), `data_seed` the shard."""
    from tvo_b200.models import BSC
    from tvo_b200.variational import EVOVariationalStates
    g = to.Generator(device=dev).manual_seed(seed)
    Wgen = to.randn(D, H, dtype=to.float64, device=dev, generator=g)
    W0 = Wgen + 0.5 * to.randn(D, H, dtype=to.float64, device=dev, generator=g)
    gd = to.Generator(device=dev).manual_seed(seed + 1000 if data_seed is None else data_seed)
    Y = to.empty(N, D, dtype=prec, device=dev)
    for b0, b1 in _chunks(N, 262144):
        s = (to.rand(b1 - b0, H, device=dev, generator=gd) < pi_gen).double()
        Y[b0:b1] = (s @ Wgen.t() + to.randn(b1 - b0, D, dtype=to.float64, device=dev, generator=gd)).to(prec)
    del s
    np.random.seed(0)  # K init: the reference's scheme, S unique Bernoulli(1/H) states repeated for every datapoint
    if at_generating_parameters:  # a converged model: theta = the generating parameters, states as dense as the data
        W0, pi0 = Wgen, pi_gen
    else:
        pi0 = 1.0 / H
    model = BSC(H, D, W0.to(prec), to.tensor([1.0], dtype=prec), to.full((H,), pi0, dtype=prec), precision=prec)
    states = EVOVariationalStates(N, H, S, prec, "randparents", "sparseflip", n_parents=P, n_generations=G, n_children=C,
                                  bitflip_frequency=1.0 / H, seed=states_seed)
    states.n_offset = n_offset
    w = 8 if prec == to.float64 else 4
    name = (f"BSC H{H} D{D} S{S} EVO(randparents,sparseflip,P{P},C{C},G{G}) {'fp64' if w == 8 else 'fp32'}"
            + (" at the generating parameters (pi = 5/H)" if at_generating_parameters else ""))
    return dict(model=model, states=states, Y=Y, name=name, H=H, D=D, S=S, alg_bytes_per_dp=_alg_bytes(D, w, S, H, w))


def make_c2(dev, N, **kw):
    return make_bsc(dev, N, 256, 144, 64, 16, 2, 2, 5.0 / 256, **kw)


def make_c2_converged(dev, N, **kw):
    """configs[1] with theta at the generating parameters: after a few E-steps K holds states as dense as the data
    (mean |s| ~ 4-5 instead of ~1.1 early in training from the pi = 1/H initialisation)."""
    return make_bsc(dev, N, 256, 144, 64, 16, 2, 2, 5.0 / 256, at_generating_parameters=True, **kw)


def make_c2_fp32(dev, N, **kw):
    """configs[1] in the stated fp32 mode: a = Y W runs on tcgen05 (kind::tf32, 3xTF32 split, TMEM accumulators)."""
    return make_bsc(dev, N, 256, 144, 64, 16, 2, 2, 5.0 / 256, prec=to.float32, **kw)


def make_c5(dev, N, **kw):
    return make_bsc(dev, N, 1024, 256, 128, 16, 4, 2, 8.0 / 1024, **kw)


def make_c3(dev, N, seed=3):
    """C3: SSSC D=64 H=256 S=32, pi_gen = 4/H, mu = 0, Psi = I, sigma2 = 0.1; EVO randparents + sparseflip P=8 C=4 G=1."""
    from tvo_b200.models import SSSC
    from tvo_b200.variational import EVOVariationalStates
    H, D, S, P, C, G = 256, 64, 32, 8, 4, 1
    g = to.Generator(device=dev).manual_seed(seed)
    Wgen = to.randn(D, H, dtype=to.float64, device=dev, generator=g)
    Y = to.empty(N, D, dtype=to.float64, device=dev)
    for b0, b1 in _chunks(N):
        s = (to.rand(b1 - b0, H, device=dev, generator=g) < 4.0 / H).double()
        z = to.randn(b1 - b0, H, dtype=to.float64, device=dev, generator=g)
        Y[b0:b1] = (s * z) @ Wgen.t() + (0.1 ** 0.5) * to.randn(b1 - b0, D, dtype=to.float64, device=dev, generator=g)
    np.random.seed(0)
    model = SSSC(H, D, Wgen + 0.1 * to.randn(D, H, dtype=to.float64, device=dev, generator=g), to.tensor([0.1]),
                 to.zeros(H), to.eye(H), to.full((H,), 4.0 / H), precision=to.float64)
    states = EVOVariationalStates(N, H, S, to.float64, "randparents", "sparseflip", P, G, C, bitflip_frequency=1.0 / H, seed=5)
    return dict(model=model, states=states, Y=Y, name="SSSC D64 H256 S32 EVO(randparents,sparseflip,P8,C4,G1) fp64",
                H=H, D=D, S=S, alg_bytes_per_dp=_alg_bytes(D, 8, S, H, 8))


def make_c4(dev, N, parent_selection="batch_fitparents", seed=3):
    """C4: NoisyOR D=100 (20 bars on 10x10, bar 0.8 / background 0.1, pi_gen = 2/20) learnt with H=64, S=64, fp32;
    EVO fitness selection (the reference's `batch_fitparents`), crossover + randflip, P=8, G=1 => S'=56."""
    from tvo_b200.models import NoisyOR
    from tvo_b200.variational import EVOVariationalStates
    H, D, S, P, G = 64, 100, 64, 8, 1
    g = to.Generator(device=dev).manual_seed(seed)
    Wb = to.full((D, 20), 0.1, device=dev)
    for i in range(10):
        Wb.view(10, 10, 20)[i, :, i] = 0.8
        Wb.view(10, 10, 20)[:, i, 10 + i] = 0.8
    Y = to.empty(N, D, dtype=to.uint8, device=dev)
    for b0, b1 in _chunks(N):
        s = (to.rand(b1 - b0, 20, device=dev, generator=g) < 0.1).float()
        pr = 1.0 - to.prod(1.0 - Wb[None] * s[:, None, :], dim=2)
        Y[b0:b1] = (to.rand(b1 - b0, D, device=dev, generator=g) < pr).to(to.uint8)
    np.random.seed(0)
    model = NoisyOR(H, D, to.rand(D, H, device=dev, generator=g), to.full((H,), 1.0 / H), precision=to.float32)
    states = EVOVariationalStates(N, H, S, to.float32, parent_selection, "randflip", P, G, crossover=True, seed=5)
    return dict(model=model, states=states, Y=Y,
                name=f"NoisyOR D100 H64 S64 EVO({parent_selection},cross_randflip,P8,G1) fp32",
                H=H, D=D, S=S, alg_bytes_per_dp=_alg_bytes(D, 1, S, H, 4))


def mean_active_units(Kp, max_rows=65536):
    """Mean number of active units per stored state, counted on the packed words of (a prefix of) K."""
    lut = to.tensor([bin(i).count("1") for i in range(256)], dtype=to.int64, device=Kp.device)
    n = min(max_rows, Kp.shape[0])
    total = 0
    for b0 in range(0, n, 8192):
        sub = Kp[b0:min(n, b0 + 8192)].contiguous().view(to.uint8)
        total += int(lut[sub.long()].sum().item())
    return total / (n * Kp.shape[1])
