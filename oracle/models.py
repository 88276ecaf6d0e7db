"""Model maths (log-pseudo-joint, log-joint, free energy, M-step).  TEST INFRASTRUCTURE.

numpy restatement of
  tvo/models/bsc.py:100-204        BSC
  tvo/models/noisyor.py:67-162     NoisyOR
  tvo/models/sssc.py:175-540       SSSC
  tvo/models/tvae.py:85-701        GaussianTVAE / BernoulliTVAE (decoder, lpj, Adam + CyclicLR step)
  tvo/utils/model_protocols.py:179-193   Optimized.free_energy
  tvo/utils/sanity.py:14-86        fix_theta (clamps / non-finite replacement)

Each class keeps ``theta`` (dict of numpy arrays, updated in place) and the reference's
accumulators under the reference's names.  ``K`` is uint8 (N,S,H) as in the reference.
"""
import math

import numpy as np

from .states import lpj2pjc, mean_posterior, logsumexp


def _fix(new, old, lo, hi):
    """sanity.py:43-86: replace non-finite entries by `old`, then clamp to [lo, hi]."""
    new = np.array(new, copy=True)
    bad = ~np.isfinite(new)
    if bad.any():
        new[bad] = np.broadcast_to(old, new.shape)[bad]
    if lo is not None:
        new = np.maximum(new, lo)
    if hi is not None:
        new = np.minimum(new, hi)
    return new


class BSC:
    def __init__(self, H, D, W_init, sigma2_init, pies_init, individual_priors=True, precision=np.float64):
        self.H, self.D, self.individual_priors, self.precision = H, D, individual_priors, precision
        self.theta = {
            "pies": np.array(pies_init, dtype=precision),
            "W": np.array(W_init, dtype=precision),
            "sigma2": np.array(sigma2_init, dtype=precision).reshape(1),
        }
        self.my_Wp = np.zeros((D, H), precision)
        self.my_Wq = np.zeros((H, H), precision)
        self.my_pies = np.zeros(H, precision)
        self.my_sigma2 = np.zeros(1, precision)
        self.my_N = 0

    def log_pseudo_joint(self, data, states):  # bsc.py:100-119
        th = self.theta
        Kf = states.astype(th["W"].dtype)
        Wbar = Kf @ th["W"].T
        lr = np.log(th["pies"] / (1 - th["pies"]))
        prior = Kf @ lr if self.individual_priors else lr * Kf.sum(axis=2)
        return np.nansum((Wbar - data[:, None, :]) ** 2, axis=2) * (-1 / 2 / th["sigma2"]) + prior

    def log_joint(self, data, states, lpj=None):  # bsc.py:121-132
        th = self.theta
        if lpj is None:
            lpj = self.log_pseudo_joint(data, states)
        Dn = np.sum(~np.isnan(data), axis=1)
        prior = np.log(1 - th["pies"]).sum() if self.individual_priors else self.H * np.log(1 - th["pies"])
        return lpj + prior - Dn[:, None] / 2 * np.log(2 * math.pi * th["sigma2"])

    def free_energy(self, data, states, lpj):  # model_protocols.py:179-193
        return float(logsumexp(self.log_joint(data, states, lpj), axis=1).sum())

    def update_param_batch(self, batch, K, lpj):  # bsc.py:134-158
        Kf = K.astype(lpj.dtype)
        Wbar = Kf @ self.theta["W"].T
        s_pjc = mean_posterior(Kf, lpj)
        q = lpj2pjc(lpj)
        self.my_pies += s_pjc.sum(axis=0)
        self.my_Wp += batch.T @ s_pjc
        H = Kf.shape[2]
        self.my_Wq += (Kf * q[:, :, None]).reshape(-1, H).T @ Kf.reshape(-1, H)  # einsum("ijk,ijl->kl") as a GEMM
        self.my_sigma2 += mean_posterior(((batch[:, None, :] - Wbar) ** 2).sum(axis=2), lpj).sum()
        self.my_N += batch.shape[0]

    def update_param_epoch(self):  # bsc.py:160-204 (the always-drawn randn of :174 only matters
        th = self.theta            # as the non-finite replacement for W and is omitted)
        N, D, H = self.my_N, self.D, self.H
        W_new = np.linalg.lstsq(self.my_Wq, self.my_Wp.T, rcond=None)[0].T
        pies_new = self.my_pies / N if self.individual_priors else self.my_pies.sum(keepdims=True) / N / H
        sigma2_new = self.my_sigma2 / N / D
        eps = 1.0e-5
        th["W"][:] = _fix(W_new, th["W"], None, None)
        th["pies"][:] = _fix(pies_new, th["pies"], eps, 1.0 - eps)
        th["sigma2"][:] = _fix(sigma2_new, th["sigma2"], eps, None)
        self.my_Wp[:] = 0
        self.my_Wq[:] = 0
        self.my_pies[:] = 0
        self.my_sigma2[:] = 0
        self.my_N = 0


class NoisyOR:
    eps = 1e-7

    def __init__(self, H, D, W_init, pi_init, precision=np.float64):
        self.H, self.D, self.precision = H, D, precision
        self.theta = {"pies": np.array(pi_init, dtype=precision), "W": np.array(W_init, dtype=precision)}
        self.new_pi = np.zeros(H, precision)
        self.Btilde = np.zeros((D, H), precision)
        self.Ctilde = np.zeros((D, H), precision)
        self.n_train = 0

    def log_pseudo_joint(self, data, states):  # noisyor.py:67-106
        K, Y = states, data
        assert K.dtype == np.uint8 and Y.dtype == np.uint8
        pi, W = self.theta["pies"], self.theta["W"]
        one = self.precision(1.0)
        logPriors = K.astype(pi.dtype) @ np.log(pi / (1 - pi))
        zero = (K == 0).all(axis=2)  # (N,S)
        Kw = K.copy()
        Kw[zero] = 1
        prods = (one - W[None, None] * Kw.astype(W.dtype)[:, :, None, :]).prod(axis=-1)  # (N,S,D)
        prods = np.clip(prods, self.precision(self.eps), self.precision(1 - self.eps))
        f1 = np.log(one / prods - one)
        f1 = np.where(Y[:, None, :] == 1, f1, 0.0).astype(self.precision)
        logPy = f1.sum(axis=-1) + np.log(prods).sum(axis=2)
        lpj = logPriors + logPy
        any_y = Y.any(axis=1)
        lpj[zero] = (-1e30 * np.broadcast_to(any_y[:, None], zero.shape)[zero]).astype(lpj.dtype)
        assert not np.isnan(lpj).any() and not np.isinf(lpj).any()
        return lpj

    def log_joint(self, data, states, lpj=None):  # noisyor.py:157-162
        if lpj is None:
            lpj = self.log_pseudo_joint(data, states)
        return np.log(1 - self.theta["pies"]).sum() + lpj

    def free_energy(self, data, states, lpj):
        return float(logsumexp(self.log_joint(data, states, lpj), axis=1).sum())

    def update_param_batch(self, batch, K, lpj):  # noisyor.py:108-138
        Kf = K.astype(lpj.dtype)
        W = self.theta["W"]
        one = lpj.dtype.type(1.0)
        self.new_pi += mean_posterior(Kf, lpj).sum(axis=0)
        Ws = one - W[None, None, :, :] * Kf[:, :, None, :]  # (N,S,D,H)
        Ws_prod = Ws.prod(axis=3, keepdims=True)
        B = Kf[:, :, None, :] / (Ws * (one - Ws_prod) + lpj.dtype.type(self.eps))
        self.Btilde += (mean_posterior(B, lpj) * (batch.astype(lpj.dtype) - 1)[:, :, None]).sum(axis=0)
        C = B * Ws_prod / Ws
        self.Ctilde += mean_posterior(C, lpj).sum(axis=0)
        self.n_train += batch.shape[0]

    def update_param_epoch(self):  # noisyor.py:140-155
        th, eps = self.theta, self.precision(self.eps)
        th["pies"][:] = np.clip(self.new_pi / self.n_train, eps, 1 - eps)
        th["W"][:] = np.clip(1 + self.Btilde / (self.Ctilde + eps), eps, 1 - eps)
        self.new_pi[:] = 0
        self.Btilde[:] = 0
        self.Ctilde[:] = 0
        self.n_train = 0


class SSSC:
    def __init__(self, H, D, W_init, sigma2_init, mus_init, Psi_init, pies_init,
                 reformulated_psi_update=False, precision=np.float64):
        self.H, self.D, self.precision = H, D, precision
        self.reformulated_psi_update = reformulated_psi_update
        p = precision
        self.theta = {
            "W": np.array(W_init, dtype=p), "sigma2": np.array(sigma2_init, dtype=p).reshape(1),
            "mus": np.array(mus_init, dtype=p), "Psi": np.array(Psi_init, dtype=p),
            "pies": np.array(pies_init, dtype=p),
        }
        self._zero()

    def _zero(self):
        D, H, p = self.D, self.H, self.precision
        self.sum_y_szT = np.zeros((D, H), p)
        self.sum_xpt_sz_xpt_szT = np.zeros((H, H), p)
        self.sum_xpt_szszT = np.zeros((H, H), p)
        self.sum_xpt_s = np.zeros(H, p)
        self.sum_xpt_sz = np.zeros(H, p)
        self.sum_xpt_ssT = np.zeros((H, H), p)
        self.sum_xpt_ssz = np.zeros((H, H), p)
        self.sum_diag_yyT = np.zeros(D, p)
        self.my_N = 0

    def _terms(self, state, notnan):  # sssc.py:216-248
        th = self.theta
        W_s = th["W"][notnan][:, state]
        Psi_s = th["Psi"][state][:, state]
        mus_s = th["mus"][state]
        Inv_Lambda_s = W_s.T @ W_s / th["sigma2"] + np.linalg.inv(Psi_s)
        Lambda_s = np.linalg.inv(Inv_Lambda_s)
        return W_s, mus_s, Psi_s, Inv_Lambda_s, Lambda_s, Lambda_s @ W_s.T / th["sigma2"]

    def log_pseudo_joint(self, data, states):  # sssc.py:266-346 (default, reformulated form)
        th = self.theta
        sigma2 = th["sigma2"]
        pies = np.clip(th["pies"], 1e-2, 1 - 1e-2)
        Kb = states.astype(bool)
        notnan = ~np.isnan(data)
        lpj = states.astype(self.precision) @ np.log(pies / (1.0 - pies))
        N, S = lpj.shape
        for n in range(N):
            y = data[n][notnan[n]]
            Dn = int(notnan[n].sum())
            for s in range(S):
                W_s, mus_s, Psi_s, ILam, Lam, LWs = self._terms(Kb[n, s], notnan[n])
                Inv_C = np.eye(Dn, dtype=self.precision) / sigma2 - W_s @ LWs / sigma2
                ld = np.linalg.slogdet(ILam)[1] + np.linalg.slogdet(Psi_s)[1]
                r = y - W_s @ mus_s
                lpj[n, s] -= 0.5 * (ld + (r * (Inv_C @ r)).sum())
        mn = np.finfo(self.precision).min
        lpj[np.isnan(lpj)] = mn
        lpj[np.isinf(lpj)] = mn
        return lpj

    def log_joint(self, data, states, lpj=None):  # sssc.py:348-364
        if lpj is None:
            lpj = self.log_pseudo_joint(data, states)
        pies = np.clip(self.theta["pies"], 1e-2, 1 - 1e-2)
        Dn = np.sum(~np.isnan(data), axis=1)
        return lpj + np.log(1.0 - pies).sum() - Dn[:, None] / 2.0 * (
            np.log(2.0 * math.pi) + np.log(self.theta["sigma2"]))

    def free_energy(self, data, states, lpj):
        return float(logsumexp(self.log_joint(data, states, lpj), axis=1).sum())

    def update_param_batch(self, batch, K, lpj):  # sssc.py:366-452
        Kb = K.astype(bool)
        Kf = K.astype(lpj.dtype)
        N, S, H = K.shape
        notnan = np.ones(batch.shape[1], bool)
        kap = np.zeros((N, S, H), self.precision)
        LpKK = np.zeros((N, S, H, H), self.precision)
        for n in range(N):
            for s in range(S):
                st = Kb[n, s]
                if st.sum() == 0:
                    continue
                W_s, mus_s, _, _, Lam, LWs = self._terms(st, notnan)
                r = batch[n] - W_s @ mus_s
                k = mus_s + LWs @ r
                kap[n, s][st] = k
                LpKK[n, s][np.outer(st, st)] = (Lam + np.outer(k, k)).flatten()
        xs = mean_posterior(Kf, lpj)
        xssT = mean_posterior(Kf[:, :, :, None] * Kf[:, :, None, :], lpj)
        xsz = mean_posterior(kap, lpj)
        xszszT = mean_posterior(LpKK, lpj)
        self.sum_xpt_s += xs.sum(0)
        self.sum_xpt_ssT += xssT.sum(0)
        self.sum_xpt_sz += xsz.sum(0)
        self.sum_xpt_sz_xpt_szT += xsz.T @ xsz
        self.sum_xpt_szszT += xszszT.sum(0)
        self.sum_diag_yyT += (batch ** 2).sum(0)
        self.sum_y_szT += batch.T @ xsz
        self.sum_xpt_ssz += (xs[:, :, None] * xsz[:, None, :]).sum(0)
        self.my_N += N

    def update_param_epoch(self):  # sssc.py:467-540
        th, H, D, N = self.theta, self.H, self.D, self.my_N
        dtype_eps, eps = np.finfo(self.precision).eps, 1e-5
        eps_eye = np.eye(H, dtype=self.precision) * 1e-6
        Inv_ssT = np.linalg.inv(self.sum_xpt_ssT)
        th["W"][:] = self.sum_y_szT @ np.linalg.inv(self.sum_xpt_szszT)
        th["pies"][:] = self.sum_xpt_s / N
        th["mus"][:] = self.sum_xpt_sz / (self.sum_xpt_s + dtype_eps)
        mus = th["mus"]
        if self.reformulated_psi_update:
            _Psi = (np.outer(mus, mus) * self.sum_xpt_ssT + self.sum_xpt_szszT
                    - 2.0 * mus[:, None] * self.sum_xpt_ssz)
            th["Psi"][:] = _Psi * Inv_ssT + eps_eye
        else:
            th["Psi"][:] = (self.sum_xpt_szszT - self.sum_xpt_ssT * np.outer(mus, mus)) * Inv_ssT + eps_eye
        W = th["W"]
        th["sigma2"][:] = (self.sum_diag_yyT.sum() - np.trace(self.sum_xpt_sz_xpt_szT @ (W.T @ W))) / N / D + eps
        self._zero()


class _Mixture:
    """Shared bookkeeping of the single-cause mixture models (gmm.py:82-88, pmm.py:70-75)."""

    def _zero(self):
        D, H, p = self.D, self.H, self.precision
        self.my_Wp = np.zeros((D, H), p)
        self.my_Wq = np.zeros(H, p)
        self.my_pies = np.zeros(H, p)
        self.my_sigma2 = np.zeros(1, p)
        self.my_N = 0

    def free_energy(self, data, states, lpj):  # model_protocols.py:179-193
        return float(logsumexp(self.log_joint(data, states, lpj), axis=1).sum())


class GMM(_Mixture):
    """tvo/models/gmm.py:20-221.  ``states`` is float/uint8 (N,S,H); single-cause use passes eye(H) per datapoint."""

    def __init__(self, H, D, W_init, sigma2_init, pies_init, precision=np.float64):
        self.H, self.D, self.precision = H, D, precision
        self.theta = {"pies": np.array(pies_init, dtype=precision), "W": np.array(W_init, dtype=precision),
                      "sigma2": np.array(sigma2_init, dtype=precision).reshape(1)}
        self._zero()

    def log_pseudo_joint(self, data, states):  # gmm.py:91-102
        th = self.theta
        Kf = states.astype(th["W"].dtype)
        Wbar = Kf @ th["W"].T
        return ((Wbar - data[:, None, :]) ** 2).sum(axis=2) * (-1 / 2 / th["sigma2"]) + Kf @ np.log(th["pies"])

    def log_joint(self, data, states, lpj=None):  # gmm.py:104-109
        if lpj is None:
            lpj = self.log_pseudo_joint(data, states)
        return lpj - self.D / 2 * np.log(2 * math.pi * self.theta["sigma2"])

    def update_param_batch(self, batch, K, lpj):  # gmm.py:111-131
        Kf = K.astype(lpj.dtype)
        Wbar = Kf @ self.theta["W"].T
        s_pjc = mean_posterior(Kf, lpj)
        self.my_pies += s_pjc.sum(axis=0)
        self.my_Wp += batch.T @ s_pjc
        self.my_Wq += s_pjc.sum(axis=0)
        self.my_sigma2 += mean_posterior(((batch[:, None, :] - Wbar) ** 2).sum(axis=2), lpj).sum()
        self.my_N += batch.shape[0]

    def update_param_epoch(self):  # gmm.py:133-180
        th, N, D = self.theta, self.my_N, self.D
        eps = 1.0e-5
        with np.errstate(all="ignore"):
            W_new = self.my_Wp / self.my_Wq[None, :]
        th["W"] = _fix(W_new, th["W"], None, None)
        th["pies"] = _fix(self.my_pies / N, th["pies"], eps, 1.0 - eps)
        th["sigma2"] = _fix(self.my_sigma2 / N / D, th["sigma2"], eps, None)
        self._zero()


class PMM(_Mixture):
    """tvo/models/pmm.py:20-198 (Poisson mixture)."""

    def __init__(self, H, D, W_init, pies_init, precision=np.float64):
        self.H, self.D, self.precision = H, D, precision
        self.theta = {"pies": np.array(pies_init, dtype=precision), "W": np.array(W_init, dtype=precision)}
        self._zero()

    def log_pseudo_joint(self, data, states):  # pmm.py:76-91
        th = self.theta
        Kf = states.astype(th["W"].dtype)
        Wbar = Kf @ th["W"].T + np.finfo(np.float32).tiny
        return (data[:, None, :] * np.log(Wbar)).sum(axis=2) - Wbar.sum(axis=2) + Kf @ np.log(th["pies"])

    def log_joint(self, data, states, lpj=None):  # pmm.py:93-97
        from scipy.special import gammaln
        if lpj is None:
            lpj = self.log_pseudo_joint(data, states)
        return lpj - gammaln(data + 1).sum(axis=1)[:, None]

    def update_param_batch(self, batch, K, lpj):  # pmm.py:99-113
        Kf = K.astype(lpj.dtype)
        s_pjc = mean_posterior(Kf, lpj)
        self.my_pies += s_pjc.sum(axis=0)
        self.my_Wp += batch.T @ s_pjc
        self.my_Wq += s_pjc.sum(axis=0)
        self.my_N += batch.shape[0]

    def update_param_epoch(self):  # pmm.py:115-153
        th, N = self.theta, self.my_N
        eps = 1.0e-5
        with np.errstate(all="ignore"):
            W_new = self.my_Wp / self.my_Wq[None, :]
        th["W"] = _fix(W_new, th["W"], eps, None)
        th["pies"] = _fix(self.my_pies / N, th["pies"], eps, 1.0 - eps)
        self._zero()


# ------------------------------------------------------------------------------------ TVAE
def tvae_layer0(K, W0, b0):
    """states @ W_0 + b_0 for uint8 states (..., H0) (tvae.py:457-461), as the sum of the active rows."""
    K = np.asarray(K)
    return K.reshape(-1, K.shape[-1]).astype(W0.dtype) @ W0 + b0


def tvae_layer0_grad(K, g):
    """d/dW_0 of a loss with gradient g (rows, H1) w.r.t. the first pre-activation: states^T @ g."""
    K = np.asarray(K)
    return K.reshape(-1, K.shape[-1]).astype(g.dtype).T @ g


class _TVAE:
    """numpy restatement of tvo/models/tvae.py (ReLU decoder given by W_init / b_init, analytical pi and
    sigma2 updates, Adam + triangular CyclicLR as in tvae.py:297-309 and tvo/utils/CyclicLR.py:194-214).
    theta keys as in the reference: W_i, b_i, pies[, sigma2]."""

    kind = "gauss"

    def __init__(self, W_init, b_init, pi_init, sigma2_init=None, precision=np.float64, min_lr=0.001, max_lr=0.01,
                 cycliclr_step_size_up=400):
        self.precision = precision
        self.W = [np.array(w, dtype=precision) for w in W_init]
        self.b = [np.array(x, dtype=precision) for x in b_init]
        self.theta = {f"W_{i}": w for i, w in enumerate(self.W)}
        self.theta.update({f"b_{i}": x for i, x in enumerate(self.b)})
        self.theta["pies"] = np.array(pi_init, dtype=precision)
        if self.kind == "gauss":
            # the reference builds to.tensor([sigma2_init]) in the default dtype (float32) before casting
            # (tvae.py:80-82): a Python float initial value passes through float32
            self.theta["sigma2"] = np.array([np.float32(0.01 if sigma2_init is None else sigma2_init)], dtype=precision)
            self.new_sigma2 = np.zeros(1, dtype=precision)
        self.new_pi = np.zeros(self.W[0].shape[0], dtype=precision)
        self.n_train = 0
        self.min_lr, self.max_lr, self.step_up = min_lr, max_lr, cycliclr_step_size_up
        self.lr, self.it = min_lr, 0
        self.m = [np.zeros_like(p) for p in self.W + self.b]
        self.v = [np.zeros_like(p) for p in self.W + self.b]
        self.t = 0

    # decoder ------------------------------------------------------------------------------------
    def _forward_all(self, K):
        """(pre-activations z_l, activations a_l) with a_0 = states (rows, H0)."""
        a = [np.asarray(K).reshape(-1, np.asarray(K).shape[-1]).astype(self.precision)]
        z = []
        for l, (W, b) in enumerate(zip(self.W, self.b)):
            z.append(a[-1] @ W + b)
            a.append(np.maximum(z[-1], 0) if l < len(self.W) - 1 else self._finish(z[-1]))
        return z, a

    def _finish(self, z):
        return z

    def forward(self, K):  # tvae.py:434-467 / 672-701
        K = np.asarray(K)
        return self._forward_all(K)[1][-1].reshape(K.shape[:-1] + (self.W[-1].shape[1],))

    def _s1(self, data, out):
        raise NotImplementedError

    def log_pseudo_joint(self, data, K):  # tvae.py:326-341 / 584-604
        pi = self.theta["pies"]
        out = self.forward(K)
        s2 = np.asarray(K).astype(self.precision) @ np.log(pi / (1.0 - pi))
        return s2 - self._s1(np.asarray(data, dtype=self.precision), out)

    def free_energy(self, data, K, lpj=None):  # tvae.py:108-116
        return float(logsumexp(self.log_joint(data, K, lpj), axis=1).sum())

    # training -----------------------------------------------------------------------------------
    def _dlpj_dz(self, data, out):
        """d lpj_ns / d z_nsd of the output pre-activation (NaN observations carry no gradient)."""
        raise NotImplementedError

    def update_param_batch(self, data, K):  # tvae.py:118-166, 412-432
        data = np.asarray(data, dtype=self.precision)
        K = np.asarray(K)
        N, S, _ = K.shape
        z, a = self._forward_all(K)
        out = a[-1].reshape(N, S, -1)
        lj = self.log_joint(data, K, self.log_pseudo_joint(data, K))
        F = float(logsumexp(lj, axis=1).sum())
        q = lpj2pjc(lj)  # dF/d lpj_ns
        delta = (-(q / N)[:, :, None] * self._dlpj_dz(data, out)).reshape(N * S, -1)
        gW, gb = [None] * len(self.W), [None] * len(self.W)
        for l in reversed(range(len(self.W))):
            gW[l] = a[l].T @ delta
            gb[l] = delta.sum(axis=0)
            if l > 0:
                delta = (delta @ self.W[l].T) * (z[l - 1] > 0)
        # Adam (torch defaults) with the lr of the current cycle position, then the scheduler step
        self.t += 1
        b1, b2, eps = 0.9, 0.999, 1e-8
        for i, (p, g) in enumerate(zip(self.W + self.b, gW + gb)):
            self.m[i] = b1 * self.m[i] + (1 - b1) * g
            self.v[i] = b2 * self.v[i] + (1 - b2) * g * g
            step = self.lr / (1 - b1 ** self.t)
            p -= step * self.m[i] / (np.sqrt(self.v[i]) / math.sqrt(1 - b2 ** self.t) + eps)
        self.it += 1
        cycle = math.floor(1 + self.it / (2 * self.step_up))
        x = 1.0 + self.it / (2 * self.step_up) - cycle
        scale = x / 0.5 if x <= 0.5 else (x - 1) / (0.5 - 1)
        self.lr = self.min_lr + (self.max_lr - self.min_lr) * scale
        # analytical accumulators use the decoder output from BEFORE the step and the stored lpj (= lj here)
        self.new_pi += mean_posterior(K.astype(self.precision), lj).sum(axis=0)
        if self.kind == "gauss":
            sq = ((data[:, None, :] - out) ** 2).sum(axis=2)
            self.new_sigma2 += (q * sq).sum()
        self.n_train += N
        return F

    def update_param_epoch(self):  # tvae.py:353-391 / 615-636
        self.theta["pies"][:] = np.clip(self.new_pi / self.n_train, 1e-5, 1 - 1e-5)
        self.new_pi[:] = 0
        if self.kind == "gauss":
            self.theta["sigma2"][:] = self.new_sigma2 / (self.n_train * self.W[-1].shape[1])
            self.new_sigma2[:] = 0
        self.n_train = 0


class GaussianTVAE(_TVAE):
    kind = "gauss"

    def _s1(self, data, out):
        return np.nansum((data[:, None, :] - out) ** 2, axis=2) / (2 * self.theta["sigma2"])

    def log_joint(self, data, K, lpj=None):  # tvae.py:343-351
        data = np.asarray(data, dtype=self.precision)
        if lpj is None:
            lpj = self.log_pseudo_joint(data, K)
        Dn = (data.shape[1] - np.isnan(data).sum(axis=1))[:, None]
        pi, s2 = self.theta["pies"], self.theta["sigma2"]
        return lpj - Dn / 2 * np.log(2 * math.pi * s2) + np.log(1 - pi).sum()

    def _dlpj_dz(self, data, out):
        d = (data[:, None, :] - out) / self.theta["sigma2"]
        return np.where(np.isnan(d), 0.0, d)


class BernoulliTVAE(_TVAE):
    kind = "bern"

    def _finish(self, z):
        return 1.0 / (1.0 + np.exp(-z))

    def _s1(self, data, out):
        with np.errstate(divide="ignore"):
            bce = -(data[:, None, :] * np.maximum(np.log(out), -100) + (1 - data[:, None, :]) * np.maximum(np.log(1 - out), -100))
        return np.nansum(bce, axis=2)

    def log_joint(self, data, K, lpj=None):  # tvae.py:606-613
        if lpj is None:
            lpj = self.log_pseudo_joint(data, K)
        return lpj + np.log(1 - self.theta["pies"]).sum()

    def _dlpj_dz(self, data, out):
        d = data[:, None, :] - out
        return np.where(np.isnan(d), 0.0, d)
