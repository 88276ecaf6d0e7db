#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (tvlearn/tvo at /root/reference).

Runs in the build container only (the reference does not travel to the GPU box).  The reference
is imported through oracle.ref_loader.import_reference() (h5py/munch stubbed, Cython dedup built
by `make -C oracle ref`).  Every array stored is an input or an output of a reference function;
nothing here calls the oracle or the product.

    python tools/make_golden.py          # rewrites tests/golden/
"""
import os
import sys

import numpy as np
import torch as to

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_loader import import_reference  # noqa: E402

tvo = import_reference()
from tvo.models import BSC, NoisyOR, SSSC, GMM, PMM  # noqa: E402
from tvo.variational import FullEM, EVOVariationalStates, FullEMSingleCauseModels  # noqa: E402
from tvo.variational import evo as ref_evo  # noqa: E402
from tvo.variational._utils import (  # noqa: E402
    set_redundant_lpj_to_low, update_states_for_batch, lpj2pjc, mean_posterior)
from tvo.trainer import Trainer  # noqa: E402
from tvo.utils.data import TVODataLoader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


class _States:
    """Minimal stand-in for TVOVariationalStates (the models only read .K and .lpj)."""

    def __init__(self, K, lpj):
        self.K, self.lpj = K, lpj


def _np(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, to.Tensor) else np.asarray(v)) for k, v in d.items()}


def rand_states(rng, N, S, H, p=0.3, zero_row=False):
    K = (rng.random((N, S, H)) < p).astype(np.uint8)
    if zero_row:
        K[:, 0] = 0
    return K


def golden_bsc():
    rng = np.random.default_rng(1)
    for tag, indiv, prec in (("ind_f64", True, to.float64), ("single_f64", False, to.float64),
                             ("ind_f32", True, to.float32)):
        N, D, H, S = 7, 5, 9, 6
        W = rng.standard_normal((D, H))
        pies = rng.uniform(0.05, 0.6, H if indiv else 1)
        sigma2 = np.array([0.7])
        Y = rng.standard_normal((N, D)) * 1.5
        K = rand_states(rng, N, S, H)
        m = BSC(H, D, to.tensor(W), to.tensor(sigma2), to.tensor(pies), individual_priors=indiv, precision=prec)
        Yt, Kt = to.tensor(Y, dtype=prec), to.tensor(K)
        lpj = m.log_pseudo_joint(Yt, Kt)
        lj = m.log_joint(Yt, Kt, lpj)
        st = _States(Kt, lpj.clone())
        F = m.free_energy(to.arange(N), Yt, st)
        Ynan = Y.copy()
        Ynan[0, 1] = np.nan
        Ynan[3, :2] = np.nan
        lpj_nan = m.log_pseudo_joint(to.tensor(Ynan, dtype=prec), Kt)
        lj_nan = m.log_joint(to.tensor(Ynan, dtype=prec), Kt, lpj_nan)
        m.update_param_batch(to.arange(N), Yt, st)
        acc = dict(my_Wp=m.my_Wp.clone(), my_Wq=m.my_Wq.clone(), my_pies=m.my_pies.clone(),
                   my_sigma2=m.my_sigma2.clone(), my_N=m.my_N.clone())
        # second batch so that Wq is better conditioned, then the epoch update
        Y2 = rng.standard_normal((40, D)) * 1.5
        K2 = rand_states(rng, 40, S, H)
        Y2t, K2t = to.tensor(Y2, dtype=prec), to.tensor(K2)
        lpj2 = m.log_pseudo_joint(Y2t, K2t)
        m.update_param_batch(to.arange(40), Y2t, _States(K2t, lpj2))
        m.update_param_epoch()
        np.savez(os.path.join(OUT, f"bsc_{tag}.npz"), **_np(dict(
            W=W, pies=pies, sigma2=sigma2, Y=Y, K=K, lpj=lpj, log_joint=lj, F=F, Ynan=Ynan,
            lpj_nan=lpj_nan, log_joint_nan=lj_nan, Y2=Y2, K2=K2, lpj2=lpj2,
            W_new=m.theta["W"], pies_new=m.theta["pies"], sigma2_new=m.theta["sigma2"], **acc)))


def golden_noisyor():
    rng = np.random.default_rng(2)
    for tag, prec in (("f64", to.float64), ("f32", to.float32)):
        N, D, H, S = 6, 7, 6, 8
        W = rng.uniform(0.05, 0.95, (D, H))
        pies = rng.uniform(0.05, 0.6, H)
        Y = (rng.random((N, D)) < 0.4).astype(np.uint8)
        Y[1] = 0  # an all-zero datapoint: lpj of the zero state must be 0, not -1e30
        K = rand_states(rng, N, S, H, zero_row=True)
        m = NoisyOR(H, D, to.tensor(W), to.tensor(pies), precision=prec)
        Yt, Kt = to.tensor(Y), to.tensor(K)
        lpj = m.log_pseudo_joint(Yt, Kt)
        st = _States(Kt, lpj.clone())
        F = m.free_energy(to.arange(N), Yt, st)
        m.update_param_batch(to.arange(N), Yt, st)
        acc = dict(new_pi=m.new_pi.clone(), Btilde=m.Btilde.clone(), Ctilde=m.Ctilde.clone())
        m.update_param_epoch()
        np.savez(os.path.join(OUT, f"noisyor_{tag}.npz"), **_np(dict(
            W=W, pies=pies, Y=Y, K=K, lpj=lpj, F=F, W_new=m.theta["W"], pies_new=m.theta["pies"], **acc)))


def golden_sssc():
    rng = np.random.default_rng(3)
    for tag, reform_psi in (("f64", False), ("f64_psi", True)):
        N, D, H, S = 5, 6, 5, 7
        prec = to.float64
        W = rng.standard_normal((D, H))
        A = rng.standard_normal((H, H)) * 0.3
        Psi = A @ A.T + np.eye(H)
        mus = rng.standard_normal(H)
        pies = rng.uniform(0.1, 0.6, H)
        sigma2 = np.array([0.4])
        Y = rng.standard_normal((N, D)) * 1.5
        K = rand_states(rng, N, S, H, p=0.4, zero_row=True)
        m = SSSC(H, D, to.tensor(W), to.tensor(sigma2), to.tensor(mus), to.tensor(Psi), to.tensor(pies),
                 reformulated_psi_update=reform_psi, precision=prec)
        Yt, Kt = to.tensor(Y), to.tensor(K)
        lpj = m.log_pseudo_joint(Yt, Kt)
        lj = m.log_joint(Yt, Kt, lpj)
        st = _States(Kt, lpj.clone())
        F = m.free_energy(to.arange(N), Yt, st)
        m2 = SSSC(H, D, to.tensor(W), to.tensor(sigma2), to.tensor(mus), to.tensor(Psi), to.tensor(pies),
                  reformulated_lpj=False, precision=prec)
        lpj_straight = m2.log_pseudo_joint(Yt, Kt)
        m.update_param_batch(to.arange(N), Yt, st)
        acc = dict(sum_y_szT=m._my_sum_y_szT.clone(), sum_xpt_sz_xpt_szT=m._my_sum_xpt_sz_xpt_szT.clone(),
                   sum_xpt_szszT=m._my_sum_xpt_szszT.clone(), sum_xpt_s=m._my_sum_xpt_s.clone(),
                   sum_xpt_sz=m._my_sum_xpt_sz.clone(), sum_xpt_ssT=m._my_sum_xpt_ssT.clone(),
                   sum_diag_yyT=m._my_sum_diag_yyT.clone())
        if reform_psi:
            acc["sum_xpt_ssz"] = m._my_sum_xpt_ssz.clone()
        # more data for a well-conditioned epoch update
        Y2 = rng.standard_normal((60, D)) * 1.5
        K2 = rand_states(rng, 60, S, H, p=0.4)
        Y2t, K2t = to.tensor(Y2), to.tensor(K2)
        lpj2 = m.log_pseudo_joint(Y2t, K2t)
        m.update_param_batch(to.arange(60), Y2t, _States(K2t, lpj2))
        m.update_param_epoch()
        np.savez(os.path.join(OUT, f"sssc_{tag}.npz"), **_np(dict(
            W=W, Psi=Psi, mus=mus, pies=pies, sigma2=sigma2, Y=Y, K=K, lpj=lpj, log_joint=lj, F=F,
            lpj_straight=lpj_straight, Y2=Y2, K2=K2, lpj2=lpj2,
            W_new=m.theta["W"], pies_new=m.theta["pies"], mus_new=m.theta["mus"],
            Psi_new=m.theta["Psi"], sigma2_new=m.theta["sigma2"], **acc)))


def golden_sssc_estimator():
    """SSSC.data_estimator (sssc.py:542-599) on complete data and on data with missing values (NaN): the reference
    restricts W_s and the datapoint to the observed dimensions (sssc.py:563-597)."""
    rng = np.random.default_rng(33)
    N, D, H, S = 6, 8, 6, 7
    W = rng.standard_normal((D, H))
    A = rng.standard_normal((H, H)) * 0.3
    Psi = A @ A.T + np.eye(H)
    mus = rng.standard_normal(H)
    pies = rng.uniform(0.1, 0.6, H)
    sigma2 = np.array([0.4])
    Y = rng.standard_normal((N, D)) * 1.5
    Ynan = Y.copy()
    Ynan[0, 2] = np.nan
    Ynan[3, [0, 5, 7]] = np.nan
    Ynan[5, 1] = np.nan
    K = rand_states(rng, N, S, H, p=0.4, zero_row=True)
    out = {}
    for tag, data in (("full", Y), ("nan", Ynan)):
        m = SSSC(H, D, to.tensor(W), to.tensor(sigma2), to.tensor(mus), to.tensor(Psi), to.tensor(pies), precision=to.float64)
        Yt, Kt = to.tensor(data), to.tensor(K)
        lpj = m.log_pseudo_joint(Yt, Kt)
        est = m.data_estimator(to.arange(N), Yt, _States(Kt, lpj.clone()))
        out[f"lpj_{tag}"], out[f"est_{tag}"] = lpj, est
    np.savez(os.path.join(OUT, "sssc_estimator.npz"), **_np(dict(
        W=W, Psi=Psi, mus=mus, pies=pies, sigma2=sigma2, Y=Y, Ynan=Ynan, K=K, **out)))


def golden_mixtures():
    """GMM / PMM with FullEMSingleCauseModels (gmm.py:91-180, pmm.py:76-143, fullem.py:61-82): lpj, log_joint,
    free energy, accumulators after one batch, theta after the epoch update, F after a second E-step."""
    rng = np.random.default_rng(12)
    for name, prec in (("gmm_f64", to.float64), ("gmm_f32", to.float32), ("pmm_f64", to.float64)):
        N, D, H = 23, 7, 5
        to.manual_seed(3)
        if name.startswith("gmm"):
            W = rng.standard_normal((D, H)) * 2
            Y = (W[:, rng.integers(0, H, N)].T + 0.7 * rng.standard_normal((N, D)))
            m = GMM(H, D, to.tensor(W + 0.3 * rng.standard_normal((D, H))), to.tensor([0.9]),
                    to.tensor(rng.dirichlet(np.ones(H))), precision=prec)
        else:
            W = rng.uniform(0.5, 9.0, (D, H))
            Y = rng.poisson(W[:, rng.integers(0, H, N)].T).astype(np.float64)
            m = PMM(H, D, to.tensor(W * rng.uniform(0.7, 1.3, (D, H))), to.tensor(rng.dirichlet(np.ones(H))),
                    precision=prec)
        theta0 = {k: v.clone() for k, v in m.theta.items()}
        st = FullEMSingleCauseModels(N, H, prec)
        Yt = to.tensor(Y, dtype=prec)
        idx = to.arange(N)
        st.update(idx, Yt, m)
        lpj = st.lpj.clone()
        lj = m.log_joint(Yt, st.K, lpj)
        F = m.free_energy(idx, Yt, st)
        m.update_param_batch(idx, Yt, st)
        acc = {"my_Wp": m.my_Wp.clone(), "my_Wq": m.my_Wq.clone(), "my_pies": m.my_pies.clone(), "my_N": m.my_N.clone()}
        if hasattr(m, "my_sigma2"):
            acc["my_sigma2"] = m.my_sigma2.clone()
        m.update_param_epoch()
        st.update(idx, Yt, m)
        F2 = m.free_energy(idx, Yt, st)
        out = dict(Y=Y, lpj=lpj, log_joint=lj, F=F, F2=F2, lpj2=st.lpj.clone(), **acc)
        out.update({f"theta0_{k}": v for k, v in theta0.items()})
        out.update({f"theta1_{k}": v for k, v in m.theta.items()})
        np.savez(os.path.join(OUT, f"{name}.npz"), **_np(out))


def golden_dedup_merge():
    rng = np.random.default_rng(4)
    N, S, Snew, H, Ntot = 12, 6, 9, 4, 20  # H=4 -> many collisions
    all_K = rand_states(rng, Ntot, S, H, p=0.4)
    all_lpj = rng.standard_normal((Ntot, S))  # tie-free
    idx = rng.permutation(Ntot)[:N]
    new_K = rand_states(rng, N, Snew, H, p=0.4)
    new_lpj = rng.standard_normal((N, Snew)) + 0.5
    nl = to.tensor(new_lpj.copy())
    set_redundant_lpj_to_low(to.tensor(new_K), nl, to.tensor(all_K[idx]))
    aK, al = to.tensor(all_K.copy()), to.tensor(all_lpj.copy())
    nsubs = update_states_for_batch(to.tensor(new_K), nl.clone(), to.tensor(idx), aK, al)
    q = lpj2pjc(to.tensor(all_lpj))
    g = rng.standard_normal((Ntot, S, 3))
    mp = mean_posterior(to.tensor(g), to.tensor(all_lpj))
    np.savez(os.path.join(OUT, "dedup_merge.npz"), **_np(dict(
        all_K=all_K, all_lpj=all_lpj, idx=idx, new_K=new_K, new_lpj=new_lpj, marked_lpj=nl,
        K_after=aK, lpj_after=al, nsubs=nsubs, pjc=q, g=g, mean_post=mp)))


def golden_evo_stats():
    """Empirical behaviour of the reference's stochastic operators (torch / numpy global RNG),
    used for distributional tests of the counter-based restatement."""
    to.manual_seed(11)
    np.random.seed(11)
    H, P, C, reps = 12, 4, 2, 20000
    rng = np.random.default_rng(5)
    parents = (rng.random((1, P, H)) < 0.3).astype(np.uint8)
    parents[0, 0] = 0
    parents[0, 0, 3] = 1
    sparsity, p_bf = 2.0 / H, 1.5 / H
    par = to.tensor(parents).repeat(reps, 1, 1)
    ch = ref_evo.batch_sparseflip(par, C, sparsity, p_bf).numpy()  # (reps, P*C, H)
    flips = (ch != np.repeat(parents, C, axis=1)).mean(axis=0)  # (P*C, H) flip frequency
    ch_r = ref_evo.batch_randflip(par, C).numpy()
    rflips = (ch_r != np.repeat(parents, C, axis=1)).mean(axis=0)
    pair_same = float(((ch_r[:, 0::C] != np.repeat(parents, C, axis=1)[:, 0::C]).argmax(-1) ==
                       (ch_r[:, 1::C] != np.repeat(parents, C, axis=1)[:, 1::C]).argmax(-1)).mean())
    cr = ref_evo.batch_cross(par).numpy()  # (reps, P(P-1), H)
    # cut distribution of pair 0: first position where child0 switches from parent0 to parent1
    # is only identifiable where parents differ; store mean child instead
    cross_mean = cr.mean(axis=0)
    # parent selection
    S = 6
    cand = to.tensor((rng.random((1, S, H)) < 0.5).astype(np.uint8))
    lpj1 = to.tensor(rng.standard_normal((1, S)) * 2 - 3)
    hist_fit = {}
    for B in (1, 4):
        lp = lpj1.repeat(B, 1)
        cd = cand.repeat(B, 1, 1)
        cnt = np.zeros(S)
        for _ in range(reps // 4):
            par_sel = ref_evo.batch_fitparents(cd, 3, lp).numpy()  # (B,3,H)
            for b in range(B):
                for p in range(3):
                    j = int(np.where((cand[0].numpy() == par_sel[b, p]).all(-1))[0][0])
                    cnt[j] += 1
        hist_fit[B] = cnt / cnt.sum()
    cnt = np.zeros((S, S))
    for _ in range(reps // 4):
        ps = ref_evo.batch_randparents(cand, 3).numpy()[0]
        js = [int(np.where((cand[0].numpy() == ps[p]).all(-1))[0][0]) for p in range(3)]
        assert len(set(js)) == 3
        for p, j in enumerate(js):
            cnt[p, j] += 1
    np.savez(os.path.join(OUT, "evo_stats.npz"), parents=parents, sparsity=sparsity, p_bf=p_bf, C=C, reps=reps,
             sparseflip_freq=flips, randflip_freq=rflips, randflip_pair_same=pair_same, cross_mean=cross_mean,
             cand=cand.numpy(), cand_lpj=lpj1.numpy(), fit_hist_B1=hist_fit[1], fit_hist_B4=hist_fit[4],
             rand_hist=cnt[:3] / cnt[:3].sum(axis=1, keepdims=True))


def golden_trainer():
    """Deterministic end-to-end trajectory: BSC + FullEM (no RNG in the E-step), 4 EM epochs."""
    rng = np.random.default_rng(6)
    N, D, H = 40, 6, 4
    Wgen = rng.standard_normal((D, H)) * 2
    s = (rng.random((N, H)) < 0.4)
    Y = s @ Wgen.T + 0.3 * rng.standard_normal((N, D))
    W0 = Wgen + 0.5 * rng.standard_normal((D, H))
    pies0 = np.full(H, 0.3)
    sigma20 = np.array([1.0])
    m = BSC(H, D, to.tensor(W0), to.tensor(sigma20), to.tensor(pies0), precision=to.float64)
    st = FullEM(N, H, to.float64)
    data = TVODataLoader(to.tensor(Y), batch_size=16)
    tr = Trainer(m, data, st)
    Fs, Ws, ps, s2 = [], [], [], []
    for _ in range(4):
        # F is accumulated by the reference Trainer in float32 (Trainer.py:230); recompute in fp64
        tr.em_step()
        F = 0.0
        for idx, batch in data:
            st.lpj[idx] = m.log_pseudo_joint(batch, st.K[idx])
            F += m.free_energy(idx, batch, st)
        Fs.append(F)
        Ws.append(m.theta["W"].clone().numpy())
        ps.append(m.theta["pies"].clone().numpy())
        s2.append(m.theta["sigma2"].clone().numpy())
    np.savez(os.path.join(OUT, "trainer_bsc_fullem.npz"), Y=Y, W0=W0, pies0=pies0, sigma20=sigma20,
             F_after=np.array(Fs), W=np.array(Ws), pies=np.array(ps), sigma2=np.array(s2))


def golden_tvae():
    """GaussianTVAE / BernoulliTVAE (tvo/models/tvae.py): lpj, log_joint, F, and theta after three
    update_param_batch steps (Adam + the reference's CyclicLR) followed by update_param_epoch."""
    from tvo.models import GaussianTVAE, BernoulliTVAE
    rng = np.random.default_rng(11)
    for kind, prec in (("gauss", to.float64), ("gauss", to.float32), ("bern", to.float64)):
        N, S, H0, H1, D = 6, 7, 9, 5, 4
        tag = f"tvae_{kind}_{'f64' if prec == to.float64 else 'f32'}"
        W = [rng.standard_normal((H0, H1)) * 0.7, rng.standard_normal((H1, D)) * 0.7]
        b = [rng.standard_normal(H1) * 0.3, rng.standard_normal(D) * 0.3]
        pies = rng.uniform(0.1, 0.5, H0)
        K = rand_states(rng, N, S, H0, p=0.3, zero_row=True)
        kw = dict(W_init=[to.tensor(w) for w in W], b_init=[to.tensor(x) for x in b], pi_init=to.tensor(pies),
                  precision=prec, min_lr=0.002, max_lr=0.02, cycliclr_step_size_up=2)
        if kind == "gauss":
            Y = rng.standard_normal((N, D)) * 1.3
            m = GaussianTVAE(sigma2_init=0.6, **kw)
        else:
            Y = (rng.random((N, D)) < 0.4).astype(np.float64)
            m = BernoulliTVAE(**kw)
        Yt, Kt = to.tensor(Y, dtype=prec), to.tensor(K)
        out = dict(W_0=W[0], W_1=W[1], b_0=b[0], b_1=b[1], pies=pies, Y=Y, K=K)
        if kind == "gauss":
            out["sigma2"] = np.array([0.6])
            Ynan = Yt.clone()
            Ynan[1, 2] = float("nan")
            out["lpj_nan"] = m._log_pseudo_joint(Ynan, Kt)
            out["log_joint_nan"] = m.log_joint(Ynan, Kt).detach()
        lpj = m._log_pseudo_joint(Yt, Kt)
        out["lpj"] = lpj
        out["log_joint"] = m.log_joint(Yt, Kt).detach()
        out["mlp_out"] = m.forward(Kt).detach()
        st = _States(Kt, m.log_joint(Yt, Kt).detach().clone())
        out["F"] = np.array(m.free_energy(to.arange(N), Yt, st))
        Fs, lrs = [], []
        for step in range(3):
            st.lpj[:] = m.log_joint(Yt, Kt).detach()
            Fs.append(m.update_param_batch(to.arange(N), Yt, st))
            lrs.append(m._optimizer.param_groups[0]["lr"])
        m.update_param_epoch()
        out["F_steps"] = np.array(Fs)
        out["lr_steps"] = np.array(lrs)
        for k, v in m.theta.items():
            out[f"{k}_after"] = v.detach()
        out["data_estimator"] = m.data_estimator(to.arange(N), Yt, st).detach()
        np.savez(os.path.join(OUT, f"{tag}.npz"), **_np(out))


if __name__ == "__main__":
    to.set_default_dtype(to.float32)
    todo = {"bsc": golden_bsc, "noisyor": golden_noisyor, "sssc": golden_sssc, "sssc_estimator": golden_sssc_estimator, "dedup_merge": golden_dedup_merge,
            "evo_stats": golden_evo_stats, "trainer": golden_trainer, "mixtures": golden_mixtures, "tvae": golden_tvae}
    for name in (sys.argv[1:] or list(todo)):  # python tools/make_golden.py [name ...]
        todo[name]()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
