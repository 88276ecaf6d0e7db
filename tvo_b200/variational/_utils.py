"""Drop-in mirrors of tvo/variational/_utils.py backed by CUDA kernels.

    set_redundant_lpj_to_low   _utils.py:88-92  (+ the Cython routine .pyx:13-30)
    update_states_for_batch    _utils.py:124-179
    generate_unique_states     _utils.py:95-121 (host-side init helper, numpy RNG as the reference)
    lpj2pjc / mean_posterior   _utils.py:182-230
"""
from typing import Dict

import numpy as np
import torch as to

import tvo_b200
from tvo_b200 import _lib
from ._packed import PackedStates, as_packed, pack, n_words


def generate_unique_states(n_states: int, H: int, crowdedness: float = 1.0, device: to.device = None) -> to.Tensor:
    """n_states distinct Bernoulli(crowdedness/H) vectors, uint8 (n_states, H)."""
    if device is None:
        device = tvo_b200.get_device()
    assert n_states <= 2 ** H, "n_states must be smaller than 2**H"
    n_samples = max(n_states // 2, 1)
    s_set = {tuple(s) for s in np.random.binomial(1, p=crowdedness / H, size=(n_samples, H))}
    while len(s_set) < n_states:
        s_set.update({tuple(s) for s in np.random.binomial(1, p=crowdedness / H, size=(n_samples, H))})
    while len(s_set) > n_states:
        s_set.pop()
    return to.from_numpy(np.array(tuple(s for s in s_set), dtype=int)).to(dtype=to.uint8, device=device)


def _rows3(packed: to.Tensor):
    assert packed.dim() == 3
    return packed if packed.is_contiguous() else packed.contiguous()


def mark_redundant_packed(new_packed: to.Tensor, new_lpj: to.Tensor, K_packed: to.Tensor, idx, H: int,
                          k_broadcast: bool = False):
    """In place on new_lpj (B,S_new).  Old states are rows idx of K_packed (N,S,W64)."""
    B, S_new, W = new_packed.shape
    S = K_packed.shape[1]
    assert new_lpj.is_contiguous() and new_packed.is_contiguous() and K_packed.is_contiguous()
    _lib.call(f"tvo_mark_redundant_{_lib.sfx(new_lpj.dtype)}", _lib.ptr(new_packed), _lib.ptr(new_lpj),
              _lib.ptr(K_packed), 0 if k_broadcast else S * W, _lib.ptr(idx), B, S_new, S, H, _lib.stream())


def select_topS_packed(new_packed, new_lpj, K_packed, lpj, idx, H, nsubs_dev):
    """In place on K_packed / lpj rows idx; nsubs_dev (int64 device scalar) += substitutions."""
    B, S_new, W = new_packed.shape
    S = K_packed.shape[1]
    assert new_lpj.dtype == lpj.dtype and lpj.is_contiguous() and K_packed.is_contiguous()
    _lib.call(f"tvo_select_topS_{_lib.sfx(lpj.dtype)}", _lib.ptr(new_packed), _lib.ptr(new_lpj), _lib.ptr(K_packed),
              _lib.ptr(lpj), _lib.ptr(idx), B, S_new, S, H, _lib.ptr(nsubs_dev), _lib.stream())


def state_marginals(K_packed: to.Tensor, lpj: to.Tensor, idx, H: int) -> to.Tensor:
    """(B,H) posterior mean of the states, sum_s q_ns K[n,s,:], rows idx (None = all) — mean_posterior of the
    states themselves (tvs.py:66-70, bsc.py:249-252) without the (B,S,H) float temporary."""
    N, S, W = K_packed.shape
    bc = N == 1 and lpj.shape[0] != 1
    B = idx.numel() if idx is not None else lpj.shape[0]
    assert lpj.is_contiguous() and K_packed.is_contiguous()
    out = to.empty((B, H), dtype=lpj.dtype, device=lpj.device)
    _lib.call(f"tvo_state_marginals_{_lib.sfx(lpj.dtype)}", _lib.ptr(K_packed), 0 if bc else S * W, _lib.ptr(lpj),
              lpj.shape[1], _lib.ptr(idx) if idx is not None else None, _lib.ptr(out), B, S, H, _lib.stream())
    return out


def set_redundant_lpj_to_low(new_states, new_lpj: to.Tensor, old_states):
    """Reference signature (_utils.py:88): uint8 (B,S',H), lpj (B,S') modified in place, uint8 (B,S,H)."""
    npk, H, _ = as_packed(new_states)
    opk, H2, bc = as_packed(old_states, H)
    assert H == H2
    lp = new_lpj if new_lpj.is_contiguous() else new_lpj.contiguous()
    mark_redundant_packed(_rows3(npk), lp, _rows3(opk), None, H, bc)
    if lp is not new_lpj:
        new_lpj.copy_(lp)


def update_states_for_batch(new_states, new_lpj: to.Tensor, idx: to.Tensor, all_states, all_lpj: to.Tensor,
                            sort_by_lpj: Dict[str, to.Tensor] = {}) -> int:
    """Reference signature (_utils.py:124).  `all_states` may be a PackedStates (updated in place, no
    conversion) or a uint8 (N,S,H) tensor (packed, merged, unpacked back in place)."""
    assert not sort_by_lpj, "sort_by_lpj is empty for all in-scope models (model_protocols.py:216-229)"
    npk, H, _ = as_packed(new_states)
    idx = idx.to(device=all_lpj.device, dtype=to.int64).contiguous()
    nsubs = to.zeros((), dtype=to.int64, device=all_lpj.device)
    new_lpj = new_lpj.to(dtype=all_lpj.dtype).contiguous()
    if isinstance(all_states, PackedStates):
        select_topS_packed(_rows3(npk), new_lpj, all_states.packed, all_lpj, idx, H, nsubs)
    else:
        assert all_states.dtype == to.uint8
        Kp = pack(all_states)
        lp = all_lpj if all_lpj.is_contiguous() else all_lpj.contiguous()
        select_topS_packed(_rows3(npk), new_lpj, Kp, lp, idx, H, nsubs)
        from ._packed import unpack
        all_states.copy_(unpack(Kp, H))
        if lp is not all_lpj:
            all_lpj.copy_(lp)
    return int(nsubs.item())


def lpj2pjc(lpj: to.Tensor) -> to.Tensor:
    """softmax over S with max shift (_utils.py:182-191), kernel lpj2pjc."""
    shape = lpj.shape
    l2 = lpj.reshape(shape[0], -1) if lpj.dim() > 2 else lpj
    l2 = l2.contiguous()
    out = to.empty_like(l2)
    _lib.call(f"tvo_lpj2pjc_{_lib.sfx(l2.dtype)}", _lib.ptr(l2), _lib.ptr(out), l2.shape[0], l2.shape[1], _lib.stream())
    return out.reshape(shape)


def mean_posterior(g: to.Tensor, lpj: to.Tensor) -> to.Tensor:
    """<g>_q with q = softmax(lpj) (_utils.py:216-230), kernel mean_posterior (q never stored)."""
    N, S = lpj.shape
    assert g.shape[:2] == (N, S)
    g2 = g.to(lpj.dtype).reshape(N, S, -1).contiguous()
    out = to.empty((N, g2.shape[2]), dtype=lpj.dtype, device=lpj.device)
    _lib.call(f"tvo_mean_posterior_{_lib.sfx(lpj.dtype)}", _lib.ptr(g2), _lib.ptr(lpj.contiguous()), _lib.ptr(out), N, S,
              g2.shape[2], _lib.stream())
    out = out.reshape((N,) + tuple(g.shape[2:]))
    assert not to.isnan(out).any() and not to.isinf(out).any()
    return out
