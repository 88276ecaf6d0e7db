"""Evolutionary candidate generation.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates tvo/variational/evo.py with the reference's sampling semantics but a counter-based
stream (oracle/philox.py) in place of torch's / numpy's global generators:

  batch_randparents      evo.py:436-451   P distinct candidates in random order (P smallest random keys)
  batch_fitparents       evo.py:397-433   fitness-proportional, with replacement, batch-global
                                          normalisation and the S-1-count index (quirk Q3)
  batch_randflip         evo.py:248-279   C distinct bit positions per parent, one flip per child
  batch_sparseflip       evo.py:282-331   per-bit independent flips with p0 / p1
  batch_cross            evo.py:335-380   one-point crossover over combinations(range(P), 2)
  batch_cross_randflip   evo.py:383-387
  batch_cross_sparseflip evo.py:390-394
  evolve_states          evo.py:118-211
  get_n_new_states       evo.py:214-218

All functions take/return uint8 (N, ., H) arrays like the reference.  ``n_glob`` is the (N,) array
of global datapoint indices that addresses the stream.
"""
from itertools import combinations

import numpy as np

from . import philox as px
from .states import pack_states, unpack_states, n_words, set_redundant_lpj_to_low


def get_n_new_states(mutation, n_parents, n_children, n_gen):
    if mutation[:5] == "cross":
        return n_parents * (n_parents - 1) * n_gen
    return n_parents * n_children * n_gen


# ----------------------------------------------------------------------------- selection
def randparents_idx(n_glob, S_c, P, seed, step, gen):
    """(N,P) indices of P distinct candidates in random order (evo.py:436-451: first P of a randperm).

    Candidate i gets the key (r_i with its low `bits` bits replaced by i), r_i = u32 stream element i
    of sub-stream 0, bits = bit_length(S_c - 1); the parents are the candidates with the P smallest
    keys in ascending key order.  Keys are unique, so no tie rule is needed."""
    n_glob = np.asarray(n_glob, dtype=np.uint32)
    tag = px.make_tag(step, gen, px.PUR_SELECT)
    bits = int(S_c - 1).bit_length()
    i = np.arange(S_c, dtype=np.uint32)[None, :]
    r = px.stream_u32(seed, n_glob[:, None], np.uint32(0), tag, i).astype(np.uint64)
    key = ((r >> np.uint64(bits)) << np.uint64(bits)) | i.astype(np.uint64)
    return np.argsort(key, axis=1, kind="stable")[:, :P].astype(np.int64)


def fit_norm(lpj):
    """(m, T) of evo.py:404-406: m = min(min(lpj), 0), T = sum(lpj - 2m) over the WHOLE batch.

    `to.tensor([min, 0.0])` (:404) is a float32 tensor, so the reference rounds m through
    float32 even for fp64 lpj; restated as such."""
    lpj = np.asarray(lpj, dtype=np.float64)
    m = float(np.float32(min(float(lpj.min()), 0.0)))
    T = float((lpj - 2.0 * m).sum())
    return m, T


def fitparents_idx(lpj, n_glob, P, seed, step, gen, norm=None, corrected=False):
    """(N,P) indices.  Reference-compatible mode reproduces evo.py:404-425 including the
    batch-global normaliser and the `S-1-count` index that wraps -1 to the last candidate.
    ``corrected=True`` is the documented alternative: per-datapoint normaliser and the
    inverse-CDF index (first j with x < cum_p[j])."""
    lpj = np.asarray(lpj, dtype=np.float64)
    N, S_c = lpj.shape
    n_glob = np.asarray(n_glob, dtype=np.uint32)
    tag = px.make_tag(step, gen, px.PUR_SELECT)
    if corrected:
        m = np.minimum(lpj.min(axis=1, keepdims=True), 0.0).astype(np.float32).astype(np.float64)
        f = lpj - 2.0 * m
        T = np.zeros((N, 1))
        for j in range(S_c):  # sequential sum, same order as the kernel
            T[:, 0] = T[:, 0] + f[:, j]
    else:
        m, T = fit_norm(lpj) if norm is None else norm
        f = lpj - 2.0 * m
    fn = f / T
    cum = np.zeros_like(fn)
    acc = np.zeros(N)
    for j in range(S_c):  # sequential cumsum (torch.cumsum order)
        acc = acc + fn[:, j]
        cum[:, j] = acc
    i = np.arange(P, dtype=np.uint32)[None, :]
    x = px.u64_to_unit_double(px.stream_u64(seed, n_glob[:, None], np.uint32(0), tag, i))  # (N,P)
    count = (x[:, :, None] < cum[:, None, :]).sum(axis=-1)
    if corrected:
        return np.minimum(S_c - count, S_c - 1)
    chosen = S_c - 1 - count
    chosen[chosen < 0] += S_c
    return chosen


def take_parents(candidates, idx):
    return candidates[np.arange(candidates.shape[0])[:, None], idx]


# ----------------------------------------------------------------------------- mutation
def _distinct_positions(seed, n, slot, tag, H, C):
    """C distinct positions in [0,H) by virtual Fisher-Yates (scalar)."""
    swapped = {}
    out = []
    for i in range(C):
        r = int(px.stream_u32(seed, np.uint32(n), np.uint32(slot), tag, np.uint32(i)))
        k = i + ((r * (H - i)) >> 32)
        vk = swapped.get(k, k)
        vi = swapped.get(i, i)
        out.append(vk)
        swapped[k] = vi
    return out


def randflip(parents, C, n_glob, seed, step, gen):
    """(N,P,H) -> (N,P*C,H); child p*C+c = parent p with bit pos[p][c] flipped."""
    N, P, H = parents.shape
    tag = px.make_tag(step, gen, px.PUR_RANDFLIP)
    children = np.repeat(parents, C, axis=1).copy()
    for n in range(N):
        for p in range(P):
            pos = _distinct_positions(seed, n_glob[n], p, tag, H, C)
            for c, h in enumerate(pos):
                children[n, p * C + c, h] ^= 1
    return children


def sparseflip_probs(a, H, sparsity, p_bf):
    """p0, p1 of evo.py:310-316 in fp64, operation for operation."""
    eps = 1e-100
    Hf = float(H)
    a = np.asarray(a, dtype=np.float64)
    crowd = float(sparsity) * H
    with np.errstate(all="ignore"):
        alpha = (Hf - a) * ((Hf * p_bf) - (crowd - a)) / ((crowd - a + Hf * p_bf) * a + eps)
        p0 = (Hf * p_bf) / (Hf + (alpha - 1.0) * a) + eps
        p1 = alpha * p0
    return p0, p1


def _pow_int(q, n):
    """q**n by binary exponentiation with a fixed operation order (bit-reproducible on the GPU)."""
    result, base, e = 1.0, float(q), int(n)
    while e > 0:
        if e & 1:
            result = result * base
        e >>= 1
        if e:
            base = base * base
    return result


def binomial_inverse(u, n, p):
    """k ~ Binomial(n, p) by CDF inversion of the uniform u in [0,1); only +, *, / in a fixed order.

    p <= 0 (or NaN) -> 0, p >= 1 -> n, p > 1/2 -> n - Binomial(n, 1-p) with the same u."""
    n = int(n)
    p = float(p)
    if n == 0 or not (p > 0.0):
        return 0
    if p >= 1.0:
        return n
    if p > 0.5:
        return n - binomial_inverse(u, n, 1.0 - p)
    q = 1.0 - p
    pmf = _pow_int(q, n)
    if pmf == 0.0:  # (1-p)^n underflowed (n > 1074): take the mean
        return int(n * p)
    r = p / q
    c, k = pmf, 0
    while u >= c and k < n:
        t = (pmf * float(n - k)) * r
        pmf = t * INV_TAB[k + 1] if k + 1 <= 64 else t / float(k + 1)
        c = c + pmf
        k += 1
    return k


INV_TAB = [0.0] + [1.0 / k for k in range(1, 65)]  # 1/(k+1) factors of the pmf recurrence


def _choose_positions(k, cand, H, seed, n_glob, slot, tag, t):
    """Absolute positions of k of the candidates ``cand`` (ascending), uniform without replacement, and the
    advanced attempt counter.  For k > n/2 the n-k candidates that are NOT flipped are drawn instead.
    u32 stream elements 4 + t of the child's sub-stream are consumed one per attempt (t runs on across the
    two classes of a child):

      dense class (2 n >= H): rejection — pos = randint(H) is accepted when it is a candidate that
          has not been drawn yet (after 64 + 64*draws attempts the sparse rule finishes the job);
      sparse class: draw picks the i-th REMAINING candidate, i = randint(#remaining)."""
    n = len(cand)
    if k <= 0:
        return [], t
    if k >= n:
        return list(cand), t
    complement = 2 * k > n
    draws = n - k if complement else k
    remaining = list(cand)
    picked = []

    def u32():
        nonlocal t
        r = int(px.stream_u32(seed, np.uint32(n_glob), np.uint32(slot), tag, np.uint32(4 + t)))
        t += 1
        return r

    if 2 * n >= H:
        cap = t + 64 + 64 * draws
        while len(picked) < draws and t < cap:
            pos = (u32() * H) >> 32
            if pos in remaining:
                remaining.remove(pos)
                picked.append(pos)
    while len(picked) < draws:
        picked.append(remaining.pop((u32() * len(remaining)) >> 32))
    return (sorted(remaining) if complement else picked), t


def sparseflip(parents, C, sparsity, p_bf, n_glob, seed, step, gen):
    """(N,P,H) -> (N,P*C,H); every bit flips independently with p1 if set else p0 (evo.py:282-331).

    Sampled exactly but cheaply: per child the NUMBER of flips among the zero bits is
    Binomial(H-a, p0) and among the one bits Binomial(a, p1); given the counts, the flipped positions
    are uniform subsets.  This is the same joint law as H independent `rand < p` draws (evo.py:325)
    while using ~2 Philox blocks per child instead of H uniforms."""
    N, P, H = parents.shape
    tag = px.make_tag(step, gen, px.PUR_FLIP0)
    a = parents.sum(axis=2)
    p0, p1 = sparseflip_probs(a, H, sparsity, p_bf)
    children = np.repeat(parents, C, axis=1).copy()
    for n in range(N):
        for j in range(P * C):
            p = j // C
            par = parents[n, p]
            zeros, ones = np.flatnonzero(par == 0), np.flatnonzero(par == 1)
            t = 0  # attempt counter of the child's sub-stream (slot j), shared by the two classes
            for cls, cand, prob in ((0, zeros, p0[n, p]), (1, ones, p1[n, p])):
                # count uniform: u64 element 0 (zero bits) / 1 (one bits) of the child's block 0
                u = float(px.u64_to_unit_double(px.stream_u64(seed, np.uint32(n_glob[n]), np.uint32(j), tag, np.uint32(cls))))
                k = binomial_inverse(u, len(cand), prob)
                flips, t = _choose_positions(k, cand, H, seed, n_glob[n], j, tag, t)
                for h in flips:
                    children[n, j, h] ^= 1
    return children


def cross(parents, n_glob, seed, step, gen):
    """(N,P,H) -> (N,P(P-1),H); child 2k = p1[:c] | p2[c:], child 2k+1 = p2[:c] | p1[c:]."""
    N, P, H = parents.shape
    pairs = list(combinations(range(P), 2))
    tag = px.make_tag(step, gen, px.PUR_CROSS)
    n_glob = np.asarray(n_glob, dtype=np.uint32)
    children = np.empty((N, 2 * len(pairs), H), dtype=np.uint8)
    for k, (a, b) in enumerate(pairs):
        r = px.stream_u32(seed, n_glob, np.uint32(0), tag, np.uint32(k))
        cut = 1 + px.randint_u32(r, H - 1)  # in [1, H-1]
        low = np.arange(H)[None, :] < cut[:, None]
        children[:, 2 * k] = np.where(low, parents[:, a], parents[:, b])
        children[:, 2 * k + 1] = np.where(low, parents[:, b], parents[:, a])
    return children


# ----------------------------------------------------------------------------- reference-cost mode
# The functions below restate the reference's operators WITH the reference's own way of drawing
# random numbers (one fp64 uniform per bit / per candidate from a global generator, evo.py:263,325,
# 363,414,440).  They are what bench.py times as the CPU baseline: same arithmetic, same memory
# traffic and the same number of random draws as the reference's torch-CPU path.  Their output is
# distributionally identical to the counter-based functions above but not stream-identical, so they
# are never used for parity checks.
def randparents_idx_np(rng, N, S_c, P):
    return np.stack([rng.permutation(S_c)[:P] for _ in range(N)])  # evo.py:440-442 (python loop per datapoint)


def fitparents_idx_np(rng, lpj, P):
    m = float(np.float32(min(float(lpj.min()), 0.0)))
    f = lpj - 2.0 * m
    cum = np.cumsum(f / f.sum(), axis=-1)
    x = rng.random((lpj.shape[0], P))
    chosen = lpj.shape[1] - 1 - (x[:, :, None] < cum[:, None, :]).sum(axis=-1)
    chosen[chosen < 0] += lpj.shape[1]
    return chosen


def randflip_np(rng, parents, C):
    N, P, H = parents.shape
    ind = np.argpartition(rng.random((N, P, H)), H - C, axis=2)[:, :, H - C:].reshape(N, P * C)  # topk(rand) :263
    children = np.repeat(parents, C, axis=1).copy()
    n_idx, s_idx = np.arange(N)[:, None], np.arange(P * C)[None, :]
    children[n_idx, s_idx, ind] = 1 - children[n_idx, s_idx, ind]
    return children


def sparseflip_np(rng, parents, C, sparsity, p_bf):
    N, P, H = parents.shape
    p0, p1 = sparseflip_probs(parents.sum(axis=2), H, sparsity, p_bf)
    children = np.repeat(parents, C, axis=1).copy()
    p = np.where(children.astype(bool), np.repeat(p1, C, axis=1)[:, :, None], np.repeat(p0, C, axis=1)[:, :, None])
    flips = rng.random((N, P * C, H)) < p  # evo.py:325
    children[flips] = 1 - children[flips]
    return children


def cross_np(rng, parents):
    N, P, H = parents.shape
    pairs = list(combinations(range(P), 2))
    cuts = rng.integers(1, H, size=(N, len(pairs)))  # evo.py:363
    children = np.empty((N, 2 * len(pairs), H), dtype=np.uint8)
    for k, (a, b) in enumerate(pairs):
        low = np.arange(H)[None, :] < cuts[:, k][:, None]
        children[:, 2 * k] = np.where(low, parents[:, a], parents[:, b])
        children[:, 2 * k + 1] = np.where(low, parents[:, b], parents[:, a])
    return children


def evolve_states_np(rng, lpj, states, lpj_fn, n_parents, n_children, n_generations, parent_selection, mutation,
                     sparsity, p_bf):
    """evolve_states (evo.py:118-211) drawing from a numpy Generator like the reference draws from torch."""
    N, S, H = states.shape
    S_new = get_n_new_states(mutation, n_parents, n_children, n_generations)
    per_gen = S_new // n_generations
    new_states = np.empty((N, S_new, H), dtype=np.uint8)
    new_lpj = np.empty((N, S_new), dtype=lpj.dtype)
    for g in range(n_generations):
        cand, cand_lpj = (states, lpj) if g == 0 else (new_states[:, (g - 1) * per_gen:g * per_gen],
                                                       new_lpj[:, (g - 1) * per_gen:g * per_gen])
        pidx = (randparents_idx_np(rng, N, cand.shape[1], n_parents) if parent_selection == "randparents"
                else fitparents_idx_np(rng, cand_lpj, n_parents))
        parents = take_parents(cand, pidx)
        if mutation == "randflip":
            ch = randflip_np(rng, parents, n_children)
        elif mutation == "sparseflip":
            ch = sparseflip_np(rng, parents, n_children, sparsity, p_bf)
        elif mutation == "cross":
            ch = cross_np(rng, parents)
        elif mutation == "cross_randflip":
            ch = randflip_np(rng, cross_np(rng, parents), 1)
        elif mutation == "cross_sparseflip":
            ch = sparseflip_np(rng, cross_np(rng, parents), 1, sparsity, p_bf)
        else:
            raise ValueError(mutation)
        sl = slice(g * per_gen, (g + 1) * per_gen)
        new_states[:, sl] = ch
        new_lpj[:, sl] = lpj_fn(ch)
    set_redundant_lpj_to_low(new_states, new_lpj, states)
    return new_states, new_lpj


def mutate(parents, mutation, n_children, sparsity, p_bf, n_glob, seed, step, gen):
    if mutation == "randflip":
        return randflip(parents, n_children, n_glob, seed, step, gen)
    if mutation == "sparseflip":
        return sparseflip(parents, n_children, sparsity, p_bf, n_glob, seed, step, gen)
    if mutation == "cross":
        return cross(parents, n_glob, seed, step, gen)
    if mutation == "cross_randflip":
        return randflip(cross(parents, n_glob, seed, step, gen), 1, n_glob, seed, step, gen)
    if mutation == "cross_sparseflip":
        return sparseflip(cross(parents, n_glob, seed, step, gen), 1, sparsity, p_bf, n_glob, seed, step, gen)
    raise ValueError(f'Mutation operator "{mutation}" not supported.')


# ----------------------------------------------------------------------------- generations
def evolve_states(lpj, states, lpj_fn, n_parents, n_children, n_generations, parent_selection,
                  mutation, sparsity, p_bf, n_glob, seed, step, fit_corrected=False, dedup=True):
    """evo.py:118-211.  Returns (new_states (N,S',H) uint8, new_lpj (N,S'))."""
    N, S, H = states.shape
    S_new = get_n_new_states(mutation, n_parents, n_children, n_generations)
    per_gen = S_new // n_generations
    new_states = np.empty((N, S_new, H), dtype=np.uint8)
    new_lpj = np.empty((N, S_new), dtype=lpj.dtype)
    for g in range(n_generations):
        if g == 0:
            cand, cand_lpj = states, lpj
        else:
            sl = slice((g - 1) * per_gen, g * per_gen)
            cand, cand_lpj = new_states[:, sl], new_lpj[:, sl]
        if parent_selection == "randparents":
            pidx = randparents_idx(n_glob, cand.shape[1], n_parents, seed, step, g)
        elif parent_selection == "batch_fitparents":
            pidx = fitparents_idx(cand_lpj, n_glob, n_parents, seed, step, g, corrected=fit_corrected)
        else:
            raise ValueError(f'Parent selection "{parent_selection}" not supported.')
        parents = take_parents(cand, pidx)
        sl = slice(g * per_gen, (g + 1) * per_gen)
        new_states[:, sl] = mutate(parents, mutation, n_children, sparsity, p_bf, n_glob, seed, step, g)
        new_lpj[:, sl] = lpj_fn(new_states[:, sl])
    if dedup:
        set_redundant_lpj_to_low(new_states, new_lpj, states)
    return new_states, new_lpj


def random_states(N, S_new, H, sparsity, n_glob, seed, step):
    """RandomSampledVarStates.py:57-59: every bit ~ Bernoulli(sparsity), no dedup."""
    W = n_words(H)
    tag = px.make_tag(step, 0, px.PUR_SAMPLE)
    out = np.zeros((N, S_new, W), dtype=np.uint64)
    for n in range(N):
        for j in range(S_new):
            for w in range(W):
                valid = (1 << min(64, H - 64 * w)) - 1
                out[n, j, w] = np.uint64(
                    px.bernoulli_mask64(seed, np.uint32(n_glob[n]), np.uint32((j << 8) | w), tag, sparsity, valid))
    return unpack_states(out, H)


# ----------------------------------------------------------------------------- TVS (tvs.py:47-81)
def state_marginals(K, lpj):
    """(N,H) approximate marginals p(s_h=1|y_n) = mean_posterior of the states themselves (tvs.py:66-70):
    sum_s softmax(lpj_n)_s K[n,s,h], evaluated as (sum_s e_s K) / (sum_s e_s) with e_s = exp(lpj - max)."""
    lpj = np.asarray(lpj, dtype=np.float64)
    e = np.exp(lpj - lpj.max(axis=1, keepdims=True))
    return (e[:, :, None] * K.astype(np.float64)).sum(axis=1) / e.sum(axis=1, keepdims=True)


def sample_states_probs(probs, S_new, H, n_glob, seed, step, group):
    """(N,S_new,H) uint8: bit h of state (n, j) = [u < probs[n, h]] with the 24-bit uniform
    u = (r >> 8) * 2^-24, r = u32 stream element h of sub-stream (n_glob[n], slot j, tag (step, group, SAMPLE))
    (tvs.py:60-63 / :71-74 draw float32 uniforms).  ``probs`` is (N,H) or one row (H,) for all datapoints."""
    n_glob = np.asarray(n_glob, dtype=np.uint32)
    N = len(n_glob)
    probs = np.broadcast_to(np.asarray(probs, dtype=np.float64), (N, H))
    tag = px.make_tag(step, group, px.PUR_SAMPLE)
    j = np.arange(S_new, dtype=np.uint32)[None, :, None]
    h = np.arange(H, dtype=np.uint32)[None, None, :]
    r = px.stream_u32(seed, n_glob[:, None, None], j, tag, h)
    u = (r >> np.uint32(8)).astype(np.float64) * 2.0 ** -24
    return (u < probs[:, None, :]).astype(np.uint8)


def tvs_new_states(K, lpj, pies, S_new_prior, S_new_marg, n_glob, seed, step):
    """tvs.py:60-76: prior samples (group 0) followed by marginal samples (group 1)."""
    H = K.shape[2]
    prior = sample_states_probs(np.broadcast_to(np.asarray(pies, dtype=np.float64), (H,)), S_new_prior, H, n_glob, seed,
                                step, 0)
    marg = sample_states_probs(state_marginals(K, lpj), S_new_marg, H, n_glob, seed, step, 1)
    return np.concatenate([prior, marg], axis=1)
