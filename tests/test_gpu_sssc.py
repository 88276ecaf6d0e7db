"""GPU parity tests of the SSSC path (through the C ABI) against the oracle and the reference fixtures."""
import math

import numpy as np
import pytest
import torch as to

from conftest import load_golden

pytestmark = pytest.mark.gpu

if to.cuda.is_available():
    import tvo_b200
    from tvo_b200.models import SSSC
    from tvo_b200.trainer import Trainer
    from tvo_b200.utils.data import TVODataLoader
    from tvo_b200.variational import EVOVariationalStates, FullEM

if to.cuda.is_available():
    from tvo_b200.variational import _packed as pk
from oracle import evo as oevo
from oracle import models as om
from oracle import states as ost

DEV = "cuda:0"


def dev(a, dtype=None):
    t = to.as_tensor(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def make(g, dtype=to.float64, reform=False):
    D, H = g["W"].shape
    return SSSC(H, D, dev(g["W"]), dev(g["sigma2"]), dev(g["mus"]), dev(g["Psi"]), dev(g["pies"]),
                reformulated_psi_update=reform, precision=dtype)


@pytest.mark.parametrize("tag,reform", [("f64", False), ("f64_psi", True)])
def test_sssc_golden(tag, reform):
    g = load_golden(f"sssc_{tag}.npz")
    D, H = g["W"].shape
    tvo_b200._set_device(to.device(DEV))
    m = make(g, reform=reform)
    Y, K = dev(g["Y"]), dev(g["K"])
    lpj = m.log_pseudo_joint(Y, K)
    assert np.allclose(lpj.cpu().numpy(), g["lpj"], rtol=1e-10, atol=1e-10)
    assert np.allclose(m.log_joint(Y, K, lpj).cpu().numpy(), g["log_joint"], rtol=1e-10, atol=1e-10)
    N, S = g["K"].shape[:2]
    st = EVOVariationalStates(N, H, S, to.float64, "randparents", "randflip", 2, 1, 1, K_init=K)
    st.lpj[:] = dev(g["lpj"])
    assert math.isclose(m.free_energy(to.arange(N), Y, st), float(g["F"]), rel_tol=1e-10)
    m.update_param_batch(to.arange(N), Y, st)
    names = {"sum_y_szT": "_my_sum_y_szT", "sum_xpt_sz_xpt_szT": "_my_sum_xpt_sz_xpt_szT", "sum_xpt_szszT": "_my_sum_xpt_szszT",
             "sum_xpt_s": "_my_sum_xpt_s", "sum_xpt_sz": "_my_sum_xpt_sz", "sum_xpt_ssT": "_my_sum_xpt_ssT",
             "sum_diag_yyT": "_my_sum_diag_yyT"}
    if reform:
        names["sum_xpt_ssz"] = "_my_sum_xpt_ssz"
    for gname, attr in names.items():
        assert np.allclose(getattr(m, attr).cpu().numpy(), g[gname], rtol=1e-9, atol=1e-9), gname
    assert math.isclose(float(m._stats[-1]), float(g["F"]), rel_tol=1e-10)
    N2 = g["K2"].shape[0]
    st2 = EVOVariationalStates(N2, H, S, to.float64, "randparents", "randflip", 2, 1, 1, K_init=dev(g["K2"]))
    st2.lpj[:] = dev(g["lpj2"])
    m.update_param_batch(to.arange(N2), dev(g["Y2"]), st2)
    m.update_param_epoch()
    for k in ("W", "pies", "mus", "Psi", "sigma2"):
        assert np.allclose(m.theta[k].cpu().numpy(), g[k + "_new"], rtol=1e-7, atol=1e-7), k


def test_sssc_appendix_a():
    W = np.array([[1.0, -0.5, 0.25, 2.0], [0.5, 1.5, -1.0, 0.0], [-0.25, 0.75, 0.5, -1.5]])
    Y = np.array([[1.0, 0.5, -0.25], [0.0, 2.0, 1.25]])
    K = np.array([[[0, 0, 0, 0], [1, 0, 0, 0], [1, 0, 1, 1]], [[0, 1, 0, 0], [0, 1, 1, 0], [1, 1, 1, 1]]], np.uint8)
    Psi = np.array([[1, 0.25, 0, 0], [0.25, 2, 0, 0.5], [0, 0, 0.5, 0], [0, 0.5, 0, 1.5]], float)
    tvo_b200._set_device(to.device(DEV))
    m = SSSC(4, 3, dev(W), dev(np.array([0.5])), dev(np.array([0.5, -0.25, 1.0, 0.0])), dev(Psi),
             dev(np.array([0.125, 0.25, 0.5, 0.375])), precision=to.float64)
    lpj = m.log_pseudo_joint(dev(Y), dev(K)).cpu().numpy()
    want = np.array([[-1.3125, -2.680354534587943, -5.344007104422202], [-3.435717196959697, -4.09862850737411, -7.485431308356349]])
    assert np.allclose(lpj, want, rtol=1e-12)


@pytest.mark.parametrize("dtype,tol", [(to.float64, 1e-9), (to.float32, 2e-4)])
def test_sssc_lpj_random_nan_and_zero_states(dtype, tol):
    rng = np.random.default_rng(4)
    N, D, H, S = 12, 20, 40, 9
    npd = np.float64 if dtype == to.float64 else np.float32
    W = rng.standard_normal((D, H))
    A = rng.standard_normal((H, H)) * 0.2
    Psi = A @ A.T + np.eye(H)
    mus = rng.standard_normal(H) * 0.5
    pies = rng.uniform(0.005, 0.3, H)  # includes values below the 0.01 clamp
    Y = rng.standard_normal((N, D)) * 2
    Y[3, 2] = np.nan
    Y[7, 5:9] = np.nan
    K = (rng.random((N, S, H)) < 0.1).astype(np.uint8)
    K[:, 0] = 0
    tvo_b200._set_device(to.device(DEV))
    m = SSSC(H, D, dev(W), dev(np.array([0.3])), dev(mus), dev(Psi), dev(pies), precision=dtype)
    o = om.SSSC(H, D, W.astype(npd), np.array([0.3], npd), mus.astype(npd), Psi.astype(npd), pies.astype(npd), precision=npd)
    want = o.log_pseudo_joint(Y.astype(npd), K)
    got = m.log_pseudo_joint(dev(Y, dtype), dev(K)).cpu().numpy()
    assert np.allclose(got, want, rtol=tol, atol=tol * np.abs(want).max())
    lj = m.log_joint(dev(Y, dtype), dev(K)).cpu().numpy()
    assert np.allclose(lj, o.log_joint(Y.astype(npd), K, want), rtol=tol, atol=tol * np.abs(want).max())


def test_sssc_estep_bit_exact_and_training():
    rng = np.random.default_rng(8)
    np.random.seed(8)
    N, D, H, S, P, Cn, G = 40, 16, 24, 8, 3, 2, 2
    Wg = rng.standard_normal((D, H))
    Psi = np.eye(H)
    mus = np.zeros(H) + 1.0
    pies = np.full(H, 3.0 / H)
    s = rng.random((N, H)) < pies
    Y = (s * (mus + rng.standard_normal((N, H)))) @ Wg.T + 0.3 * rng.standard_normal((N, D))
    K = np.zeros((N, S, H), np.uint8)
    for n in range(N):
        seen = {}
        while len(seen) < S:
            v = (rng.random(H) < 3.0 / H).astype(np.uint8)
            seen.setdefault(v.tobytes(), v)
        K[n] = list(seen.values())
    tvo_b200._set_device(to.device(DEV))
    m = SSSC(H, D, dev(Wg + 0.3 * rng.standard_normal((D, H))), dev(np.array([0.5])), dev(mus), dev(Psi), dev(pies),
             precision=to.float64)
    o = om.SSSC(H, D, m.theta["W"].cpu().numpy(), [0.5], mus, Psi, pies)
    st = EVOVariationalStates(N, H, S, to.float64, "randparents", "sparseflip", P, G, Cn, bitflip_frequency=1.0 / H, seed=21,
                              K_init=dev(K))
    nsubs = st.update(to.arange(N, device=DEV), dev(Y), m)
    lpj = o.log_pseudo_joint(Y, K)
    new_s, new_l = oevo.evolve_states(lpj, K, lambda x: o.log_pseudo_joint(Y, x), P, Cn, G, "randparents", "sparseflip",
                                      float(m._tvo_sparsity().item()), 1.0 / H, np.arange(N), 21, 0)
    Kw = K.copy()
    nw = ost.update_states_for_batch(new_s, new_l, np.arange(N), Kw, lpj)
    assert np.array_equal(st.K.unpack().cpu().numpy(), Kw)
    assert nsubs == nw
    assert np.allclose(st.lpj.cpu().numpy(), lpj, rtol=1e-9, atol=1e-9)
    # a few EM iterations through the Trainer: F must increase (test/sssc_test.py:146-159)
    tr = Trainer(m, TVODataLoader(dev(Y), batch_size=16), st)
    Fs = [tr.em_step()["train_F"] for _ in range(4)]
    assert np.all(np.isfinite(Fs)) and Fs[-1] > Fs[0]


def test_sssc_fullem_train_increases_F():
    # test/sssc_test.py:146-159 (FullEM, default-initialised model)
    to.manual_seed(0)
    tvo_b200._set_device(to.device(DEV))
    N, D, H = 5, 4, 2
    for prec in (to.float32, to.float64):
        m = SSSC(H=H, D=D, precision=prec)
        data = to.rand((N, D), dtype=prec, device=DEV)
        st = FullEM(N, H, prec)
        st.lpj[:] = m.log_pseudo_joint(data, st.K)
        F0 = m.free_energy(to.arange(N), data, st)
        m.update_param_batch(to.arange(N), data, st)
        m.update_param_epoch()
        st.lpj[:] = m.log_pseudo_joint(data, st.K)
        assert m.free_energy(to.arange(N), data, st) > F0


def test_sssc_dense_states_all_solver_levels_and_overflow():
    """States with 0..6 active units run in sub-warp groups, 7..16 with the full warp in shared memory,
    17..32 in the warp's global scratch row, more than 32 in a pooled scratch slot sized for |s| = H (no input
    raises): lpj and the M-step statistics against the oracle for a K that mixes all levels."""
    rng = np.random.default_rng(21)
    N, D, H, S = 6, 40, 48, 12
    W = rng.standard_normal((D, H))
    A = rng.standard_normal((H, H)) * 0.1
    Psi = A @ A.T + np.eye(H)
    mus = rng.standard_normal(H) * 0.3
    pies = rng.uniform(0.02, 0.3, H)
    Y = rng.standard_normal((N, D)) * 2
    counts = [0, 1, 2, 5, 8, 9, 12, 16, 17, 24, 32, 3]
    K = np.zeros((N, S, H), np.uint8)
    for n in range(N):
        for s, c in enumerate(counts):
            K[n, s, rng.permutation(H)[:c]] = 1
    tvo_b200._set_device(to.device(DEV))
    m = SSSC(H, D, dev(W), dev(np.array([0.5])), dev(mus), dev(Psi), dev(pies), precision=to.float64)
    o = om.SSSC(H, D, W, np.array([0.5]), mus, Psi, pies)
    want = o.log_pseudo_joint(Y, K)
    got = m.log_pseudo_joint(dev(Y), dev(K)).cpu().numpy()
    assert np.allclose(got, want, rtol=1e-9, atol=1e-9 * np.abs(want).max())
    # M-step statistics with posterior weights that give every level a non-negligible share
    lpj = rng.standard_normal((N, S))
    st = EVOVariationalStates(N, H, S, to.float64, "randparents", "randflip", 2, 1, 1, K_init=dev(K))
    st.lpj[:] = dev(lpj)
    m.update_param_batch(to.arange(N), dev(Y), st)
    o.update_param_batch(Y, K, lpj)
    for attr in ("_my_sum_xpt_s", "_my_sum_xpt_sz", "_my_sum_xpt_ssT", "_my_sum_xpt_szszT", "_my_sum_xpt_sz_xpt_szT",
                 "_my_sum_y_szT", "_my_sum_diag_yyT"):
        a, b = getattr(m, attr).cpu().numpy(), getattr(o, attr[len("_my_"):])
        assert np.allclose(a, b, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(b).max())), attr
    m._check_overflow()  # nothing overflowed so far
    # states beyond 32 active units (the reference evaluates any |s|, sssc.py:287-335): the slow tier in global memory
    K2 = K.copy()
    K2[2, 4, :40] = 1
    K2[5, 0, :] = 1   # all 48 units active
    want2 = o.log_pseudo_joint(Y, K2)
    got2 = m.log_pseudo_joint(dev(Y), dev(K2)).cpu().numpy()
    assert np.allclose(got2, want2, rtol=1e-9, atol=1e-9 * np.abs(want2).max())
    m._check_overflow()
    m2 = SSSC(H, D, dev(W), dev(np.array([0.5])), dev(mus), dev(Psi), dev(pies), precision=to.float64)
    o2 = om.SSSC(H, D, W, np.array([0.5]), mus, Psi, pies)
    st2 = EVOVariationalStates(N, H, S, to.float64, "randparents", "randflip", 2, 1, 1, K_init=dev(K2))
    st2.lpj[:] = dev(lpj)
    m2.update_param_batch(to.arange(N), dev(Y), st2)
    o2.update_param_batch(Y, K2, lpj)
    for attr in ("_my_sum_xpt_s", "_my_sum_xpt_sz", "_my_sum_xpt_ssT", "_my_sum_xpt_szszT", "_my_sum_xpt_sz_xpt_szT",
                 "_my_sum_y_szT", "_my_sum_diag_yyT"):
        a, b = getattr(m2, attr).cpu().numpy(), getattr(o2, attr[len("_my_"):])
        assert np.allclose(a, b, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(b).max())), attr
    m2._check_overflow()


def test_sssc_c3_shape_properties():
    """BASELINE configs[2] shape (D=64, H=256, S=32), 20k datapoints: the fused E-step runs as ONE wide CTA per SM
    (16 warps, several datapoints per warp) and the M-step likewise.  Size-independent properties: the stored lpj
    equal a stand-alone re-score of the final K, rows are sorted ascending with pairwise distinct states, F does
    not decrease, and the M-step statistics are additive over the two halves of the data."""
    tvo_b200._set_device(to.device(DEV))
    g = to.Generator(device=DEV).manual_seed(11)
    N, H, D, S = 20_000, 256, 64, 32
    Wg = to.randn(D, H, dtype=to.float64, device=DEV, generator=g)
    s = (to.rand(N, H, device=DEV, generator=g) < 4.0 / H).double()
    z = to.randn(N, H, dtype=to.float64, device=DEV, generator=g)
    Y = (s * z) @ Wg.t() + (0.1 ** 0.5) * to.randn(N, D, dtype=to.float64, device=DEV, generator=g)
    mk = lambda: SSSC(H, D, Wg + 0.1 * to.randn(D, H, dtype=to.float64, device=DEV, generator=to.Generator(device=DEV).manual_seed(5)),  # noqa: E731
                      to.tensor([0.1]), to.zeros(H), to.eye(H), to.full((H,), 4.0 / H), precision=to.float64)
    m = mk()
    np.random.seed(0)
    st = EVOVariationalStates(N, H, S, to.float64, "randparents", "sparseflip", 8, 1, 4, bitflip_frequency=1.0 / H, seed=5)
    idx = to.arange(N, device=DEV)
    st.lpj[:] = m.log_pseudo_joint(Y, st.K)
    F0 = m.free_energy(idx, Y, st)
    for _ in range(2):
        st.update(idx, Y, m)
    again = m.log_pseudo_joint(Y, st.K)
    assert to.allclose(st.lpj, again, rtol=1e-9, atol=1e-9 * float(again.abs().max()))
    assert bool((st.lpj[:, 1:] >= st.lpj[:, :-1]).all())
    Kp = st._Kp[:2000]
    same = (Kp[:, :, None, :] == Kp[:, None, :, :]).all(dim=3)
    assert int(same.sum()) == Kp.shape[0] * S  # only the diagonal: no duplicate states in a row
    assert m.free_energy(idx, Y, st) >= F0
    m.update_param_batch(idx, Y, st)
    full = m._stats.clone()
    m2 = mk()
    m2.update_param_batch(idx[: N // 2], Y[: N // 2], st)
    m2.update_param_batch(idx[N // 2:], Y[N // 2:], st)
    assert to.allclose(full, m2._stats, rtol=1e-9, atol=1e-9 * float(full.abs().max()))
    m._check_overflow()


@pytest.mark.parametrize("N", [20_000, 200_000], ids=["20k", "configs2_full_size"])
def test_sssc_c3_shape_oracle_subset(N):
    """BASELINE configs[2] shape (SSSC D=64 H=256 S=32, EVO randparents + sparseflip P=8 C=4 G=1, fp64) inside a
    20k-datapoint launch and at its full size N = 200k, so that the wide-CTA tier (16 warps, 4 lanes per state) is the one running: E-step K
    bit-exact and lpj to 1e-10 against the oracle on 10 of the datapoints, two consecutive E-steps; M-step statistics
    of those datapoints against `om.SSSC` (same kernel tier: it is picked by the shape, not the batch size) and
    additivity of the 20k launch over (subset, complement)."""
    tvo_b200._set_device(to.device(DEV))
    rng = np.random.default_rng(31)
    H, D, S, P, Cn, G = 256, 64, 32, 8, 4, 1
    Wg = rng.standard_normal((D, H))
    s = rng.random((N, H)) < 4.0 / H
    Y = (s * rng.standard_normal((N, H))) @ Wg.T + (0.1 ** 0.5) * rng.standard_normal((N, D))
    W0 = Wg + 0.1 * rng.standard_normal((D, H))
    A = 0.05 * rng.standard_normal((H, H))
    Psi = np.eye(H) + A @ A.T          # a full covariance: the state-dependent blocks are not diagonal
    mus = 0.2 * rng.standard_normal(H)
    pies = np.full(H, 4.0 / H)
    m = SSSC(H, D, dev(W0), dev(np.array([0.1])), dev(mus), dev(Psi), dev(pies), precision=to.float64)
    o = om.SSSC(H, D, W0, np.array([0.1]), mus, Psi, pies)
    np.random.seed(0)
    st = EVOVariationalStates(N, H, S, to.float64, "randparents", "sparseflip", P, G, Cn, bitflip_frequency=1.0 / H, seed=77)
    sub = np.array([0, 1, 2, 777, 5000, 9999, 10000, 15001, N - 2, N - 1])
    idx = to.arange(N, device=DEV)
    Yd = dev(Y)
    Kw = st.K.unpack()[dev(sub)].cpu().numpy() if N <= 20_000 else pk.unpack(st._Kp[dev(sub)], H).cpu().numpy()
    sparsity = float(m._tvo_sparsity().item())
    for step in range(2):
        lpj = o.log_pseudo_joint(Y[sub], Kw)
        new_s, new_l = oevo.evolve_states(lpj, Kw, lambda x: o.log_pseudo_joint(Y[sub], x), P, Cn, G, "randparents",
                                          "sparseflip", sparsity, 1.0 / H, sub, 77, step)
        Kw = Kw.copy()
        ost.update_states_for_batch(new_s, new_l, np.arange(len(sub)), Kw, lpj)
        nsubs = st.update(idx, Yd, m)
        assert 0 < nsubs <= N * S
        got = pk.unpack(st._Kp[dev(sub)], H).cpu().numpy()
        assert np.array_equal(got, Kw), f"K differs from the oracle at step {step}"
        assert np.allclose(st.lpj[dev(sub)].cpu().numpy(), lpj, rtol=1e-10, atol=1e-10 * np.abs(lpj).max())
    # M-step: the subset against the oracle, then additivity of the full launch
    names = ("_my_sum_xpt_s", "_my_sum_xpt_sz", "_my_sum_xpt_ssT", "_my_sum_xpt_szszT", "_my_sum_xpt_sz_xpt_szT",
             "_my_sum_y_szT", "_my_sum_diag_yyT")
    lp = st.lpj.cpu().numpy()
    m.update_param_batch(dev(sub), Yd[dev(sub)], st)
    o.update_param_batch(Y[sub], Kw, lp[sub])
    for attr in names:
        a, b = getattr(m, attr).cpu().numpy(), getattr(o, attr[len("_my_"):])
        assert np.allclose(a, b, rtol=1e-10, atol=1e-10 * max(1.0, np.abs(b).max())), attr
    part = m._stats.clone()
    comp = np.setdiff1d(np.arange(N), sub)
    m.update_param_batch(dev(comp), Yd[dev(comp)], st)   # accumulates on top of the subset
    both = m._stats.clone()
    m._stats.zero_()
    m.update_param_batch(idx, Yd, st)
    assert to.allclose(m._stats, both, rtol=1e-10, atol=1e-10 * float(both.abs().max()))
    assert float((both - part).abs().max()) > 0
    m._check_overflow()


def test_sssc_data_estimator_golden_with_missing_values():
    """SSSC.data_estimator against the unmodified reference (tests/golden/sssc_estimator.npz, tools/make_golden.py):
    complete data and data with NaNs — the reference computes kappa from the observed dimensions only
    (sssc.py:563-597); a zero fill would give different reconstructions."""
    g = load_golden("sssc_estimator.npz")
    D, H = g["W"].shape
    tvo_b200._set_device(to.device(DEV))
    m = SSSC(H, D, dev(g["W"]), dev(g["sigma2"]), dev(g["mus"]), dev(g["Psi"]), dev(g["pies"]), precision=to.float64)
    N, S = g["K"].shape[:2]
    for tag, Y in (("full", g["Y"]), ("nan", g["Ynan"])):
        st = EVOVariationalStates(N, H, S, to.float64, "randparents", "randflip", 2, 1, 1, K_init=dev(g["K"]))
        lpj = m.log_pseudo_joint(dev(Y), st.K)
        assert np.allclose(lpj.cpu().numpy(), g[f"lpj_{tag}"], rtol=1e-9, atol=1e-9)
        st.lpj[:] = lpj
        before = m._stats.clone()
        est = m.data_estimator(to.arange(N, device=DEV), dev(Y), st)
        assert np.allclose(est.cpu().numpy(), g[f"est_{tag}"], rtol=1e-9, atol=1e-9), tag
        assert to.equal(m._stats, before)  # the scratch pass leaves the statistics alone
    assert not np.allclose(g["est_full"], g["est_nan"])
