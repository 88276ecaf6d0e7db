"""GPU parity tests of the TVS E-step (tvo/variational/tvs.py) and the state-marginal kernel through the C ABI."""
import numpy as np
import pytest
import torch as to

pytestmark = pytest.mark.gpu

if to.cuda.is_available():
    import tvo_b200
    from tvo_b200 import _lib
    from tvo_b200.models import BSC
    from tvo_b200.variational import TVSVariationalStates, EVOVariationalStates
    from tvo_b200.variational._utils import state_marginals, mean_posterior

from oracle import evo as oevo
from oracle import models as om
from oracle import states as ost

DEV = "cuda:0"


def dev(a, dtype=None):
    t = to.as_tensor(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def packed(K):
    return dev(ost.pack_states(K).view(np.int64))


@pytest.mark.parametrize("dtype", [to.float64, to.float32])
@pytest.mark.parametrize("N,S,H", [(7, 5, 3), (40, 16, 70), (33, 64, 256), (5, 9, 130)])
def test_state_marginals_vs_oracle_and_mean_posterior(dtype, N, S, H):
    rng = np.random.default_rng(N + S + H)
    npd = np.float64 if dtype == to.float64 else np.float32
    K = (rng.random((N, S, H)) < 0.2).astype(np.uint8)
    K[:, :, 0] = 1  # a unit that is active in every state -> marginal exactly ~1
    if H > 1:
        K[:, :, 1] = 0
    lpj = (5 * rng.standard_normal((N, S))).astype(npd)
    want = oevo.state_marginals(K, lpj)
    tvo_b200._set_device(to.device(DEV))
    idx = rng.permutation(N)[: max(1, N // 2)]
    got = state_marginals(packed(K), dev(lpj), dev(idx.astype(np.int64)), H).cpu().numpy()
    tol = 1e-12 if dtype == to.float64 else 1e-6
    assert np.allclose(got, want[idx], rtol=tol, atol=tol)
    assert np.all(got[:, 0] > 1 - 1e-6) and (H == 1 or np.all(got[:, 1] == 0))
    # the same quantity through the reference's generic mean_posterior (_utils.py:194-230)
    ref = mean_posterior(dev(K[idx].astype(npd)), dev(lpj[idx])).cpu().numpy()
    assert np.allclose(got, ref, rtol=tol, atol=tol)
    # idx = None: all rows in order
    got_all = state_marginals(packed(K), dev(lpj), None, H).cpu().numpy()
    assert np.allclose(got_all, want, rtol=tol, atol=tol)


@pytest.mark.parametrize("H,S_new", [(3, 4), (64, 5), (70, 33), (256, 8), (300, 3)])
def test_sample_states_probs_bit_exact(H, S_new):
    rng = np.random.default_rng(H * S_new)
    N, S_off, S_tot = 11, 2, S_new + 3
    probs = rng.random((N, H))
    probs[:, 0] = 0.0
    probs[:, -1] = 1.0
    idx = rng.permutation(50)[:N]
    W = ost.n_words(H)
    out = to.zeros((N, S_tot, W), dtype=to.int64, device=DEV)
    probs_d, idx_d = dev(probs), dev(idx.astype(np.int64))  # keep the device buffers alive across the call
    _lib.call("tvo_sample_states_probs_f64", _lib.ptr(out), S_tot * W, S_off, _lib.ptr(probs_d), H,
              _lib.ptr(idx_d), N, S_new, H, 1234, 7, 1, 100, _lib.stream())
    got = ost.unpack_states(out.cpu().numpy().view(np.uint64), H)
    want = oevo.sample_states_probs(probs, S_new, H, 100 + idx, 1234, 7, 1)
    assert np.array_equal(got[:, S_off:S_off + S_new], want)
    assert not got[:, :S_off].any() and not got[:, S_off + S_new:].any()      # untouched slots
    assert not got[:, :, 0].any() and got[:, S_off:S_off + S_new, -1].all()    # p = 0 never, p = 1 always
    # broadcast row (stride 0) in float32
    row = rng.random(H).astype(np.float32)
    out2 = to.zeros((N, S_new, W), dtype=to.int64, device=DEV)
    row_d = dev(row)
    _lib.call("tvo_sample_states_probs_f32", _lib.ptr(out2), S_new * W, 0, _lib.ptr(row_d), 0, None, N, S_new, H,
              99, 0, 0, 5, _lib.stream())
    got2 = ost.unpack_states(out2.cpu().numpy().view(np.uint64), H)
    assert np.array_equal(got2, oevo.sample_states_probs(row.astype(np.float64), S_new, H, 5 + np.arange(N), 99, 0, 0))


def test_sample_states_probs_rate():
    H, N, S_new = 40, 4000, 8
    probs = np.linspace(0.02, 0.98, H)
    out = to.zeros((N, S_new, 1), dtype=to.int64, device=DEV)
    probs_d = dev(probs)
    _lib.call("tvo_sample_states_probs_f64", _lib.ptr(out), S_new, 0, _lib.ptr(probs_d), 0, None, N, S_new, H, 5, 0, 0, 0,
              _lib.stream())
    freq = ost.unpack_states(out.cpu().numpy().view(np.uint64), H).reshape(-1, H).mean(axis=0)
    assert np.all(np.abs(freq - probs) < 5 * np.sqrt(probs * (1 - probs) / (N * S_new)) + 1e-3)


@pytest.mark.parametrize("dtype", [to.float64, to.float32])
def test_tvs_update_bit_exact_vs_oracle(dtype):
    rng = np.random.default_rng(5)
    N, H, D, S, Sp, Sm = 30, 20, 12, 10, 6, 7
    npd = np.float64 if dtype == to.float64 else np.float32
    W = rng.standard_normal((D, H)).astype(npd)
    pies = rng.uniform(0.05, 0.3, H).astype(npd)
    sigma2 = np.array([0.8], npd)
    Y = ((rng.random((N, H)) < pies) @ W.T + rng.standard_normal((N, D))).astype(npd)
    K0 = (rng.random((N, S, H)) < 0.15).astype(np.uint8)
    for n in range(N):  # distinct states per datapoint
        seen = set()
        for s in range(S):
            while K0[n, s].tobytes() in seen:
                K0[n, s] = (rng.random(H) < 0.15)
            seen.add(K0[n, s].tobytes())
    tvo_b200._set_device(to.device(DEV))
    m = BSC(H, D, dev(W), dev(sigma2), dev(pies), precision=dtype)
    st = TVSVariationalStates(N, H, S, dtype, Sp, Sm, seed=42, K_init=dev(K0))
    st.n_offset = 1000
    assert st.config["S_new"] == Sp + Sm and st.config["S_new_prior"] == Sp and st.config["S_new_marg"] == Sm
    o = om.BSC(H, D, W, sigma2, pies, precision=npd)
    Kw = K0.copy()
    total = 0
    idx = np.arange(N)
    for step in range(3):
        lw = o.log_pseudo_joint(Y, Kw).astype(npd)
        # the marginals are taken from the device's own lpj in fp32 (rounding of lpj moves them by ~1e-7)
        nsubs = st.update(to.arange(N, device=DEV), dev(Y), m)
        new = oevo.tvs_new_states(Kw, lw, pies.astype(np.float64), Sp, Sm, 1000 + idx, 42, step)
        nl = o.log_pseudo_joint(Y, new).astype(npd)
        ost.set_redundant_lpj_to_low(new, nl, Kw)
        Kc, lc = Kw.copy(), lw.copy()
        want_n = ost.update_states_for_batch(new, nl, idx, Kc, lc)
        got_K = st.K.unpack().cpu().numpy()
        if dtype == to.float64:
            assert np.array_equal(got_K, Kc), f"K differs at step {step}"
            assert nsubs == want_n
            assert np.allclose(st.lpj.cpu().numpy(), lc, rtol=1e-10, atol=1e-10)
            Kw = Kc
        else:
            same = [set(map(bytes, got_K[n])) == set(map(bytes, Kc[n])) for n in range(N)]
            assert sum(same) >= N - 3
            Kw = got_K
        total += nsubs
    assert total > 0


def test_tvs_with_third_party_model_and_trainer_protocol():
    """TVS through the duck-typed protocol: a model that only implements log_joint (tvs.py:49-55)."""
    tvo_b200._set_device(to.device(DEV))

    class CountModel:  # prefers states with many active units
        theta = {"pies": to.full((6,), 0.5, device=DEV)}

        def log_joint(self, data, states):
            return states.to(to.float64).sum(dim=2)

    N, H, S = 12, 6, 4
    st = TVSVariationalStates(N, H, S, to.float64, 5, 5, seed=1)
    before = None
    for _ in range(6):
        st.update(to.arange(N, device=DEV), to.zeros(N, 3, device=DEV), CountModel())
        tot = st.lpj.sum().item()
        assert before is None or tot >= before
        before = tot
    K = st.K.unpack().cpu().numpy()
    assert all(len(set(map(bytes, K[n]))) == S for n in range(N))  # dedup keeps the sets duplicate-free
    assert np.allclose(st.lpj.cpu().numpy(), K.sum(axis=2))


def test_bsc_data_estimator_matches_reference_formula():
    rng = np.random.default_rng(2)
    N, H, D, S = 9, 12, 7, 6
    W = rng.standard_normal((D, H))
    K = (rng.random((N, S, H)) < 0.3).astype(np.uint8)
    lpj = rng.standard_normal((N, S))
    tvo_b200._set_device(to.device(DEV))
    m = BSC(H, D, dev(W), dev(np.array([1.0])), dev(np.full(H, 0.2)), precision=to.float64)
    st = EVOVariationalStates(N, H, S, to.float64, "randparents", "randflip", 2, 1, 1, K_init=dev(K))
    st.lpj[:] = dev(lpj)
    idx = to.tensor([4, 0, 7], device=DEV)
    got = m.data_estimator(idx, None, st).cpu().numpy()
    q = ost.lpj2pjc(lpj)
    want = np.einsum("ns,nsd->nd", q, K.astype(np.float64) @ W.T)[[4, 0, 7]]   # bsc.py:249-252
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)
