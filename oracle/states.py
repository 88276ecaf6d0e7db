"""State layout helpers + dedup + top-S merge.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates
  tvo/variational/_set_redundant_lpj_to_low_CPU.pyx:13-30   (set_redundant_lpj_to_low)
  tvo/variational/_utils.py:124-179                         (update_states_for_batch)
  tvo/variational/_utils.py:95-121                          (generate_unique_states)
  tvo/variational/fullem.py:14-28                           (state_matrix)
  tvo/variational/_utils.py:182-230                         (lpj2pjc, mean_posterior)
"""
import ctypes
import os

import numpy as np

# value the reference's CPU path writes for redundant children: `cdef float low_lpj = -1e20`
# (_set_redundant_lpj_to_low_CPU.pyx:17) -> float32(-1e20) widened = -1.0000000200408773e20
LOW_LPJ = float(np.float32(-1e20))


# ----------------------------------------------------------------------------- packing
def n_words(H):
    return (int(H) + 63) // 64


def pack_states(K):
    """uint8 (..., H) -> uint64 (..., ceil(H/64)); bit (h % 64) of word (h // 64) = K[..., h]."""
    K = np.asarray(K, dtype=np.uint8)
    H = K.shape[-1]
    W = n_words(H)
    pad = W * 64 - H
    if pad:
        K = np.concatenate([K, np.zeros(K.shape[:-1] + (pad,), np.uint8)], axis=-1)
    bits = K.reshape(K.shape[:-1] + (W, 64)).astype(np.uint64)
    weights = np.uint64(1) << np.arange(64, dtype=np.uint64)
    return (bits * weights).sum(axis=-1, dtype=np.uint64)


def unpack_states(P, H):
    P = np.asarray(P, dtype=np.uint64)
    bits = (P[..., :, None] >> np.arange(64, dtype=np.uint64)) & np.uint64(1)
    return bits.reshape(P.shape[:-1] + (-1,))[..., :H].astype(np.uint8)


def state_matrix(H):
    """All 2^H states, row i = binary of i, most significant bit first (fullem.py:14-28)."""
    i = np.arange(2 ** H, dtype=np.int64)[:, None]
    sh = np.arange(H - 1, -1, -1, dtype=np.int64)[None, :]
    return ((i >> sh) & 1).astype(np.uint8)


def generate_unique_states(n_states, H, crowdedness=1.0, rng=None):
    """S distinct Bernoulli(crowdedness/H) states (_utils.py:95-121).  Host-side init helper."""
    rng = np.random if rng is None else rng
    assert n_states <= 2 ** H
    n_samples = max(n_states // 2, 1)
    s_set = {tuple(s) for s in rng.binomial(1, p=crowdedness / H, size=(n_samples, H))}
    while len(s_set) < n_states:
        s_set.update({tuple(s) for s in rng.binomial(1, p=crowdedness / H, size=(n_samples, H))})
    while len(s_set) > n_states:
        s_set.pop()
    return np.array(tuple(s for s in s_set), dtype=np.uint8)


# ----------------------------------------------------------------------------- dedup
_clib = None


def _load_clib():
    global _clib
    if _clib is None:
        here = os.path.dirname(os.path.abspath(__file__))
        path = os.path.join(here, "_build", "liboracle.so")
        if os.path.exists(path):
            lib = ctypes.CDLL(path)
            for name in ("oracle_mark_redundant_f64", "oracle_mark_redundant_f32"):
                getattr(lib, name).restype = None
            _clib = lib
        else:
            _clib = False
    return _clib


def set_redundant_lpj_to_low(new_states, new_lpj, old_states, low=None):
    """In-place; _set_redundant_lpj_to_low_CPU.pyx:13-30.

    Child s is marked if it equals a LATER child (so the last copy survives), else if it equals
    any old state.  Uses the C restatement (oracle/csrc/dedup.c) when built, else numpy.
    """
    new_states = np.ascontiguousarray(new_states, dtype=np.uint8)
    old_states = np.ascontiguousarray(old_states, dtype=np.uint8)
    N, Snew, H = new_states.shape
    S = old_states.shape[1]
    lib = _load_clib()
    if lib and low is None and new_lpj.flags.c_contiguous and new_lpj.dtype in (np.float64, np.float32):
        fn = lib.oracle_mark_redundant_f64 if new_lpj.dtype == np.float64 else lib.oracle_mark_redundant_f32
        fn(new_states.ctypes.data_as(ctypes.c_void_p), new_lpj.ctypes.data_as(ctypes.c_void_p),
           old_states.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(N), ctypes.c_size_t(Snew),
           ctypes.c_size_t(S), ctypes.c_size_t(H))
        return new_lpj
    low = LOW_LPJ if low is None else low
    for n in range(N):
        eq_new = (new_states[n][:, None, :] == new_states[n][None, :, :]).all(-1)  # (Snew,Snew)
        later = np.triu(eq_new, k=1).any(axis=1)
        eq_old = (new_states[n][:, None, :] == old_states[n][None, :, :]).all(-1).any(axis=1)
        new_lpj[n, later | eq_old] = low
    return new_lpj


# ----------------------------------------------------------------------------- top-S merge
def _sort_key(lpj):
    """Total order used for selection: larger lpj first, NaN above everything (torch.topk
    semantics), ties -> lower concatenated index first (BASELINE north_star)."""
    k = np.array(lpj, dtype=np.float64, copy=True)
    k[np.isnan(k)] = np.inf
    return k


def update_states_for_batch(new_states, new_lpj, idx, all_states, all_lpj):
    """In-place top-S merge (_utils.py:124-179).  Returns nsubs.

    K_n ends up in ASCENDING lpj order (best state last), exactly as `flip(topk(...sorted))`
    leaves it (:165).  Ties (unspecified in torch.topk) are broken by lower concatenated index:
    an equal-lpj candidate with lower index ranks higher, i.e. sits later in the ascending row.
    """
    idx = np.asarray(idx, dtype=np.int64)
    S = all_states.shape[1]
    old_states = all_states[idx]
    old_lpj = all_lpj[idx]
    conc_states = np.concatenate([old_states, new_states], axis=1)
    conc_lpj = np.concatenate([old_lpj, new_lpj.astype(all_lpj.dtype)], axis=1)
    key = _sort_key(conc_lpj)
    # stable argsort of -key: descending lpj, ties by ascending index
    order = np.argsort(-key, axis=1, kind="stable")[:, :S]
    sorted_idx = order[:, ::-1]
    b = np.arange(len(idx))[:, None]
    all_states[idx] = conc_states[b, sorted_idx]
    all_lpj[idx] = conc_lpj[b, sorted_idx]
    return int((sorted_idx >= S).sum())


# ----------------------------------------------------------------------------- posterior
def lpj2pjc(lpj):
    """softmax over S with max shift (_utils.py:182-191)."""
    shft = 0.0 - lpj.max(axis=1, keepdims=True)
    tmp = np.exp(lpj + shft)
    return tmp / tmp.sum(axis=1, keepdims=True)


def mean_posterior(g, lpj):
    """<g>_q = sum_s q_ns g_ns...  (_utils.py:194-230)."""
    q = lpj2pjc(lpj)
    out = np.einsum("ns...,ns->n...", g, q)
    assert not np.isnan(out).any() and not np.isinf(out).any()
    return out


def logsumexp(a, axis=1):
    m = a.max(axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    return (m + np.log(np.exp(a - m).sum(axis=axis, keepdims=True))).squeeze(axis)
