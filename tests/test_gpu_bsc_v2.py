"""GPU parity tests of the list-based fused BSC E-step (csrc/bsc_estep_v2.cuh): against the CPU oracle, against the
generic kernel (tvo_bsc_estep_variant(1)) and through its overflow hand-back (dense states, NaN rows, many flips).

Bars as everywhere: K / nsubs bit-exact, lpj 1e-10 relative (fp64)."""
import numpy as np
import pytest
import torch as to

pytestmark = pytest.mark.gpu

cuda = to.cuda.is_available()
if cuda:
    import tvo_b200
    from tvo_b200 import _lib
    from tvo_b200.models import BSC
    from tvo_b200.variational import EVOVariationalStates

from oracle import evo as oevo
from oracle import models as om
from oracle import states as ost

DEV = "cuda:0"


def dev(a, dtype=None):
    t = to.as_tensor(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def rand_K(rng, N, S, H, p):
    K = np.zeros((N, S, H), np.uint8)
    for n in range(N):
        seen = {}
        while len(seen) < S:
            s = (rng.random(H) < p).astype(np.uint8)
            seen.setdefault(s.tobytes(), s)
        K[n] = list(seen.values())
    return K


def oracle_estep(o, Y, K, P, Cn, G, p_bf, n_glob, seed, step, sparsity):
    lpj = o.log_pseudo_joint(Y, K)
    new_s, new_l = oevo.evolve_states(lpj, K, lambda s: o.log_pseudo_joint(Y, s), P, Cn, G, "randparents", "sparseflip",
                                      sparsity, p_bf, n_glob, seed, step)
    Kc, lc = K.copy(), lpj.copy()
    nsubs = ost.update_states_for_batch(new_s, new_l, np.arange(len(K)), Kc, lc)
    return Kc, lc, nsubs


def make(rng, N, H, D, S, p_state, pies_val, dense_rows=(), nan_rows=()):
    W = rng.standard_normal((D, H))
    pies = np.full(H, pies_val)
    Y = (rng.random((N, H)) < pies) @ W.T + rng.standard_normal((N, D))
    K = rand_K(rng, N, S, H, p_state)
    for n in dense_rows:  # one state with more than 15 active units: the list cannot hold it
        K[n, n % S] = 0
        K[n, n % S, rng.choice(H, 17 + (n % 5), replace=False)] = 1
    for n in nan_rows:
        Y[n, n % D] = np.nan
    return W, pies, Y, K


@pytest.mark.parametrize("H,D,S,P,Cn,G,pbf_mult", [
    (256, 144, 64, 16, 2, 2, 1.0),   # BASELINE configs[1]
    (64, 20, 10, 4, 3, 2, 2.0),      # one word, smallest H of the fast path
    (100, 33, 33, 7, 2, 3, 1.5),     # two words, S not a multiple of 32, three generations
    (130, 30, 100, 20, 3, 2, 1.5),   # three words, S' = 120
    (200, 16, 128, 32, 4, 1, 1.0),   # S = S' = 128 (EPL 4)
    (256, 40, 17, 5, 7, 2, 6.0),     # many flips per child: more than 8 added units happen -> overflow hand-back
    (257, 24, 40, 8, 3, 2, 1.5),     # two-byte positions (H > 256), five words (odd), d_h gathered from the a row
    (600, 32, 50, 10, 2, 2, 3.0),    # ten words
    (1024, 64, 128, 16, 4, 2, 1.0),  # BASELINE configs[4] (D reduced for the oracle)
    (1000, 20, 24, 6, 5, 1, 8.0),    # many flips per child at H > 256
])
def test_v2_matches_oracle_and_generic(H, D, S, P, Cn, G, pbf_mult):
    rng = np.random.default_rng(H * 7 + S)
    N = 40
    dense_rows, nan_rows = (3, 17), (5, 17, 39)
    W, pies, Y, K = make(rng, N, H, D, S, min(0.3, 3.0 / H), min(0.3, 3.0 / H), dense_rows, nan_rows)
    p_bf = pbf_mult / H
    tvo_b200._set_device(to.device(DEV))
    outs = {}
    for variant in (0, 1):
        _lib.call("tvo_bsc_estep_variant", variant)
        try:
            m = BSC(H, D, dev(W), to.tensor([1.2], dtype=to.float64), dev(pies), precision=to.float64)
            st = EVOVariationalStates(N, H, S, to.float64, "randparents", "sparseflip", P, G, Cn, bitflip_frequency=p_bf,
                                      seed=31, K_init=dev(K))
            st.n_offset = 1000
            res = []
            for step in range(3):
                nsubs = st.update(to.arange(N, device=DEV), dev(Y), m)
                res.append((st.K.unpack().cpu().numpy(), st.lpj.cpu().numpy().copy(), nsubs))
            outs[variant] = (res, float(m._tvo_sparsity().item()))
        finally:
            _lib.call("tvo_bsc_estep_variant", 0)
    # list-based == generic, step by step (K, order, nsubs exactly; lpj to rounding)
    for (Ka, la, na), (Kb, lb, nb) in zip(outs[0][0], outs[1][0]):
        assert np.array_equal(Ka, Kb)
        assert na == nb
        ok = ~np.isnan(la)
        assert np.array_equal(np.isnan(la), np.isnan(lb))
        assert np.allclose(la[ok], lb[ok], rtol=1e-12, atol=1e-10)
    # == oracle on the rows without NaN (the oracle's nansum rows are covered by test_gpu_bsc_path)
    good = np.array([n for n in range(N) if n not in nan_rows])
    o = om.BSC(H, D, W, [1.2], pies)
    Kw = K[good]
    for step in range(3):
        Kw, Lw, _ = oracle_estep(o, Y[good], Kw, P, Cn, G, p_bf, 1000 + good, 31, step, outs[0][1])
        Kg, lg, _ = outs[0][0][step]
        assert np.array_equal(Kg[good], Kw), f"K differs from the oracle at step {step}"
        assert np.allclose(lg[good], Lw, rtol=1e-10, atol=1e-10)


def test_v2_incremental_lpj_equals_full_rescore():
    """Children are scored as parent + flipped terms; the stored lpj must equal a full re-score of the stored
    state (to 1e-12)."""
    rng = np.random.default_rng(4)
    N, H, D, S = 4000, 256, 144, 64
    W, pies, Y, K = make(rng, 64, H, D, S, 5.0 / H, 5.0 / H)
    Y = np.tile(Y, (N // 64 + 1, 1))[:N] + 0.01 * rng.standard_normal((N, D))
    K = np.tile(K, (N // 64 + 1, 1, 1))[:N]
    tvo_b200._set_device(to.device(DEV))
    m = BSC(H, D, dev(W), to.tensor([1.0], dtype=to.float64), dev(pies), precision=to.float64)
    st = EVOVariationalStates(N, H, S, to.float64, "randparents", "sparseflip", 16, 2, 2, bitflip_frequency=1.0 / H,
                              seed=5, K_init=dev(K))
    idx = to.arange(N, device=DEV)
    Yd = dev(Y)
    for it in range(4):
        nsubs = st.update(idx, Yd, m)
        assert nsubs > 0
        full = m._tvo_lpj_packed(Yd, st._Kp, H)
        rel = ((full - st.lpj).abs() / full.abs().clamp_min(1.0)).max().item()
        assert rel < 1e-12, rel
        assert bool((st.lpj[:, 1:] >= st.lpj[:, :-1]).all())
        rows = to.cat([idx[:, None, None].expand(-1, S, 1), st._Kp], dim=2).reshape(N * S, -1)
        assert to.unique(rows, dim=0).shape[0] == N * S  # no duplicate state inside any K_n


@pytest.mark.parametrize("N,H,D", [(1, 256, 144), (15, 256, 144), (17, 256, 144), (64, 256, 144), (200, 256, 144),
                                   (4099, 256, 144), (333, 512, 48), (130, 1024, 256), (70, 768, 16)])
def test_own_gemms_match_library(N, H, D):
    """The hand-written TMA + FP64 tensor-core GEMMs (csrc/bsc_gemm.cuh: a = Y W with ||y||^2 for the E-step,
    WpT += E^T Y for the M-step) against the library GEMM path (tvo_bsc_estep_variant bit 16) on the same inputs,
    including partial 64-row tiles and partial 16-datapoint chunks: identical K, lpj / statistics to 1e-12."""
    rng = np.random.default_rng(N)
    S = 16  # H > 256: a = Y W in column panels of 256 (the M-step GEMM of those shapes is the library's either way)
    W, pies, Y, K = make(rng, N, H, D, S, 4.0 / H, 4.0 / H)
    tvo_b200._set_device(to.device(DEV))
    out = {}
    for variant in (0, 0x10000):
        _lib.call("tvo_bsc_estep_variant", variant)
        try:
            m = BSC(H, D, dev(W), to.tensor([0.9], dtype=to.float64), dev(pies), precision=to.float64)
            st = EVOVariationalStates(N, H, S, to.float64, "randparents", "sparseflip", 4, 2, 2, bitflip_frequency=1.0 / H,
                                      seed=8, K_init=dev(K))
            idx = to.arange(N, device=DEV)
            _lib.reset_launch_count()
            nsubs = st.update(idx, dev(Y), m)
            m.update_param_batch(idx, dev(Y), st)
            out[variant] = (st.K.unpack().cpu().numpy(), st.lpj.cpu().numpy().copy(), nsubs, m._stats.cpu().numpy().copy())
        finally:
            _lib.call("tvo_bsc_estep_variant", 0)
    a, b = out[0], out[0x10000]
    assert np.array_equal(a[0], b[0]) and a[2] == b[2]
    assert np.allclose(a[1], b[1], rtol=1e-12, atol=1e-10)
    assert np.allclose(a[3], b[3], rtol=1e-12, atol=1e-12 * np.abs(b[3]).max())
    # and against a plain fp64 product of the same operands
    lp = to.as_tensor(a[1], device=DEV)
    q = to.softmax(lp, dim=1)
    E = (q[:, :, None] * to.as_tensor(a[0], device=DEV).double()).sum(dim=1)
    WpT = (E.t() @ dev(Y)).cpu().numpy()
    # <s> is accumulated in 2^-56 fixed point: absolute accuracy S 2^-56 ~ 1e-15 per unit and datapoint (N = 1 with a
    # dominant empty state has |WpT| ~ 1e-37 altogether)
    assert np.allclose(a[3][: H * D].reshape(H, D), WpT, rtol=1e-10, atol=1e-10 * np.abs(WpT).max() + 1e-14 * N * np.abs(Y).max())


@pytest.mark.parametrize("N,H,D", [(1, 256, 144), (100, 256, 144), (129, 256, 144), (4099, 256, 144), (300, 128, 64), (77, 64, 40)])
def test_tf32x3_tcgen05_gemm_matches_library_fp32(N, H, D):
    """fp32 mode: a = Y W on the 5th-generation tensor cores (csrc/bsc_gemm_tf32.cuh: tcgen05.mma kind::tf32 with the
    hi / lo split, accumulators in TMEM, TMA-fed) against the library SGEMM path on the same inputs: the E-step's lpj
    agree to fp32 accuracy (1e-5 of their scale), K sets agree on all but near-ties, and the statistics to 1e-4."""
    rng = np.random.default_rng(N + H)
    S = 16
    W, pies, Y, K = make(rng, N, H, D, S, 4.0 / H, 4.0 / H)
    W, Y = W.astype(np.float32), Y.astype(np.float32)
    tvo_b200._set_device(to.device(DEV))
    out = {}
    for variant in (0, 0x10000):
        _lib.call("tvo_bsc_estep_variant", variant)
        try:
            m = BSC(H, D, dev(W), to.tensor([0.9], dtype=to.float32), dev(pies, to.float32), precision=to.float32)
            st = EVOVariationalStates(N, H, S, to.float32, "randparents", "sparseflip", 4, 2, 2, bitflip_frequency=1.0 / H,
                                      seed=8, K_init=dev(K))
            idx = to.arange(N, device=DEV)
            _lib.kernel_timing_enable(True)
            nsubs = st.update(idx, dev(Y), m)
            own_launches = _lib.kernel_timing_read(_lib.TIMED_BSC_GEMM_YW)[1]
            _lib.kernel_timing_enable(False)
            assert own_launches == (1 if variant == 0 else 0)  # the tcgen05 kernel really is the one that ran
            m.update_param_batch(idx, dev(Y), st)
            out[variant] = (st.K.unpack().cpu().numpy(), st.lpj.double().cpu().numpy(), nsubs, m._stats.cpu().numpy().copy())
        finally:
            _lib.call("tvo_bsc_estep_variant", 0)
    a, b = out[0], out[0x10000]
    same = [set(map(bytes, a[0][n])) == set(map(bytes, b[0][n])) for n in range(N)]
    assert sum(same) >= N - max(1, N // 200), f"{N - sum(same)} of {N} K sets differ"
    ok = np.array(same)
    scale = np.abs(b[1]).max()
    assert np.allclose(np.sort(a[1][ok], 1), np.sort(b[1][ok], 1), rtol=1e-5, atol=1e-5 * scale)
    assert np.allclose(a[3], b[3], rtol=1e-4, atol=1e-4 * np.abs(b[3]).max())


@pytest.mark.parametrize("H,D,S,P,Cn,G", [(256, 144, 64, 16, 2, 2), (600, 32, 50, 10, 2, 2), (1024, 64, 128, 16, 4, 2)])
def test_v2_fp32_mode_matches_generic_kernel(H, D, S, P, Cn, G):
    """fp32 mode of the list-based kernel (one- and two-byte positions) against the generic fused kernel on the same
    inputs: the candidates are bit-identical by construction of the stream, the two kernels round lpj differently
    (fp64 accumulation cast to fp32 vs fp32 arithmetic), so the K sets agree on all but near-ties and lpj to 1e-5."""
    rng = np.random.default_rng(H + S)
    N = 300
    W, pies, Y, K = make(rng, N, H, D, S, min(0.3, 3.0 / H), min(0.3, 3.0 / H), dense_rows=(7,), nan_rows=(11,))
    W, Y = W.astype(np.float32), Y.astype(np.float32)
    tvo_b200._set_device(to.device(DEV))
    out = {}
    for variant in (0, 1):
        _lib.call("tvo_bsc_estep_variant", variant)
        try:
            m = BSC(H, D, dev(W), to.tensor([1.1], dtype=to.float32), dev(pies, to.float32), precision=to.float32)
            st = EVOVariationalStates(N, H, S, to.float32, "randparents", "sparseflip", P, G, Cn, bitflip_frequency=1.5 / H,
                                      seed=4, K_init=dev(K))
            nsubs = st.update(to.arange(N, device=DEV), dev(Y), m)
            out[variant] = (st.K.unpack().cpu().numpy(), st.lpj.double().cpu().numpy(), nsubs)
        finally:
            _lib.call("tvo_bsc_estep_variant", 0)
    a, b = out[0], out[1]
    same = np.array([set(map(bytes, a[0][n])) == set(map(bytes, b[0][n])) for n in range(N)])
    assert same.sum() >= N - 3, f"{N - same.sum()} of {N} K sets differ"
    fin = same & ~np.isnan(b[1]).any(axis=1)
    scale = np.abs(b[1][fin]).max()
    assert np.allclose(np.sort(a[1][fin], 1), np.sort(b[1][fin], 1), rtol=1e-5, atol=1e-5 * scale)
    assert bool((np.diff(a[1][fin], axis=1) >= 0).all())


def test_own_gemms_match_library_many_tiles_per_cta():
    """The hand-written GEMMs where every CTA walks MANY tiles / slabs (N = 300 000: 4688 row tiles of a = Y W over 148
    persistent CTAs; the small cases above give each CTA at most one tile): K identical to the library-GEMM path, the
    stored lpj equal to a direct re-score, the M-step statistics equal.  A producer / consumer race in the mbarrier
    pipeline shows up here as a few stale rows (one such variant was caught by exactly this comparison)."""
    N, H, D, S = 300_000, 256, 144, 64
    tvo_b200._set_device(to.device(DEV))
    g = to.Generator(device=DEV).manual_seed(9)
    Wg = to.randn(D, H, dtype=to.float64, device=DEV, generator=g)
    Y = (to.rand(N, H, device=DEV, generator=g) < 5.0 / H).double() @ Wg.t() + to.randn(N, D, dtype=to.float64, device=DEV, generator=g)
    W0 = Wg + 0.5 * to.randn(D, H, dtype=to.float64, device=DEV, generator=g)
    idx = to.arange(N, device=DEV)
    out = {}
    for variant in (0, 0x10000):
        _lib.call("tvo_bsc_estep_variant", variant)
        try:
            np.random.seed(0)
            m = BSC(H, D, W0, to.tensor([1.0], dtype=to.float64), to.full((H,), 1.0 / H, dtype=to.float64), precision=to.float64)
            st = EVOVariationalStates(N, H, S, to.float64, "randparents", "sparseflip", 16, 2, 2, bitflip_frequency=1.0 / H, seed=3)
            for _ in range(2):
                nsubs = st.update(idx, Y, m)
            full = m._tvo_lpj_packed(Y, st._Kp, H)
            assert to.allclose(full, st.lpj, rtol=1e-12, atol=1e-9)
            m.update_param_batch(idx, Y, st)
            out[variant] = (st._Kp.clone(), nsubs, m._stats.clone())
        finally:
            _lib.call("tvo_bsc_estep_variant", 0)
    a, b = out[0], out[0x10000]
    assert to.equal(a[0], b[0]) and a[1] == b[1]
    assert to.allclose(a[2], b[2], rtol=1e-10, atol=1e-10 * float(b[2].abs().max()))
