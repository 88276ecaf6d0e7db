"""GPU tests of the hand-written W solve of BSC.update_param_epoch (csrc/bsc_chol.cuh, tvo_bsc_epoch_solve):
blocked Cholesky + substitution against torch's fp64 solve on the same packed statistics, the factorisation flags,
bit-reproducibility, and the model-level route (golden theta is covered by test_gpu_bsc_path)."""
import numpy as np
import pytest
import torch as to

pytestmark = pytest.mark.gpu

if to.cuda.is_available():
    import tvo_b200
    from tvo_b200 import _lib
    from tvo_b200.models import BSC

DEV = "cuda:0"


def packed_stats(rng, H, D, n=None, scale=1.0):
    """[WpT (H*D) | WqU (H*H, upper incl. diagonal) | pies (H) | sigma2 | N | F] from n random sparse states."""
    n = n or 6 * H + 10
    s = (rng.random((n, H)) < max(2.0 / H, 0.05)).astype(np.float64)
    s[np.arange(H) % n, np.arange(H)] = 1.0  # every unit active at least once: positive definite with the jitter below
    q = rng.random(n) * scale
    Wq = (s * q[:, None]).T @ s + 1e-3 * scale * np.eye(H)
    WpT = rng.standard_normal((H, D))
    stats = np.concatenate([WpT.ravel(), np.triu(Wq).ravel(), s.sum(0), [1.0, float(n), 0.0]])
    return stats, Wq, WpT


def run_solve(stats, H, D):
    st = to.as_tensor(stats, device=DEV)
    nwork = _lib.lib().tvo_bsc_epoch_solve_work_doubles(H)
    assert nwork > 0
    work = to.full((nwork + 2,), float("nan"), dtype=to.float64, device=DEV)  # the kernels must not read unwritten work
    sol = to.empty((H, D), dtype=to.float64, device=DEV)
    _lib.call("tvo_bsc_epoch_solve", _lib.ptr(st), _lib.ptr(work), _lib.ptr(sol), _lib.ptr(work[nwork:]), H, D, _lib.stream())
    to.cuda.synchronize()
    return sol.cpu().numpy(), work[nwork:].cpu().numpy(), work[:nwork].cpu().numpy()


@pytest.mark.parametrize("single_cta", [0, 1], ids=["cluster", "one_cta"])
@pytest.mark.parametrize("H,D", [(1, 1), (8, 16), (31, 5), (32, 144), (33, 20), (64, 40), (65, 9), (100, 33), (128, 64),
                                 (200, 150), (255, 7), (256, 144)])
def test_epoch_solve_matches_fp64_solve(H, D, single_cta):
    """Both factorisation kernels (the 8-CTA cluster for H > 64, the one-CTA kernel) against numpy's fp64 Cholesky /
    solve of the same matrix."""
    tvo_b200._set_device(to.device(DEV))
    _lib.call("tvo_bsc_epoch_solve_variant", single_cta)
    rng = np.random.default_rng(100 * H + D)
    stats, Wq, WpT = packed_stats(rng, H, D)
    sol, flags, work = run_solve(stats, H, D)
    ref = np.linalg.solve(Wq, WpT)
    assert flags[0] == 0.0
    Lref = np.linalg.cholesky(Wq)
    assert abs(flags[1] - np.diag(Lref).min() / np.diag(Lref).max()) < 1e-12
    Hp = (H + 31) // 32 * 32
    L = np.tril(work[: Hp * Hp].reshape(Hp, Hp))[:H, :H]
    assert np.allclose(L, Lref, rtol=1e-11, atol=1e-13)
    # backward error of the solve at the level of an fp64 Cholesky
    resid = np.abs(Wq @ sol - WpT).max() / (np.abs(Wq).max() * np.abs(sol).max())
    assert resid < 1e-13, resid
    assert np.allclose(sol, ref, rtol=1e-8, atol=1e-10 * np.abs(ref).max())
    sol2, _, _ = run_solve(stats, H, D)
    _lib.call("tvo_bsc_epoch_solve_variant", 0)
    assert np.array_equal(sol, sol2)  # deterministic: what makes the theta broadcast unnecessary


def test_epoch_solve_flags_a_unit_that_was_never_active():
    tvo_b200._set_device(to.device(DEV))
    rng = np.random.default_rng(5)
    H, D = 64, 12
    stats, Wq, WpT = packed_stats(rng, H, D)
    U = stats[H * D: H * D + H * H].reshape(H, H)
    U[17, :] = 0.0
    U[:, 17] = 0.0
    _, flags, _ = run_solve(stats, H, D)
    assert flags[0] != 0.0 or not (flags[1] > 1e-6)
    stats2, _, _ = packed_stats(rng, H, D)
    stats2[3] = np.nan  # non-finite statistics reach the solution: reported, the caller falls back
    _, flags2, _ = run_solve(stats2, H, D)
    assert flags2[0] != 0.0


def test_update_param_epoch_uses_the_own_solve_and_matches_lstsq():
    tvo_b200._set_device(to.device(DEV))
    rng = np.random.default_rng(11)
    H, D = 48, 20
    outs = {}
    for solver in ("gpu", "gpu_lstsq"):
        m = BSC(H, D, to.as_tensor(rng.standard_normal((D, H)), device=DEV), to.tensor([1.0], dtype=to.float64),
                to.full((H,), 0.1, dtype=to.float64), precision=to.float64)
        m.W_solver = solver
        stats, Wq, WpT = packed_stats(np.random.default_rng(3), H, D)
        m._stats.copy_(to.as_tensor(stats, device=DEV))
        _lib.reset_launch_count()
        m.update_param_epoch()
        to.cuda.synchronize()
        outs[solver] = (m.theta["W"].cpu().numpy().copy(), _lib.launch_count())
    assert np.allclose(outs["gpu"][0], outs["gpu_lstsq"][0], rtol=1e-8, atol=1e-10)
    assert np.allclose(outs["gpu"][0], np.linalg.solve(Wq, WpT).T, rtol=1e-8, atol=1e-10)
    assert outs["gpu"][1] >= 3  # factor + solve + finish launched from libtvo_b200.so
