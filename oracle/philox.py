"""Counter-based random stream shared by the oracle and the CUDA kernels.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Philox4x32 with ROUNDS = 7 rounds (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
SC'11; Random123 v1.14 ``philox.h``: "Philox4x32-7", the smallest round count that passes BigCrush with margin;
the library default of 10 only adds safety margin).  It replaces the reference's global torch /
numpy generators (tvo/variational/evo.py:263,325,363,414,440;
RandomSampledVarStates.py:57) with an addressable stream so that a datapoint's
candidates depend only on (seed, E-step counter, global datapoint index,
generation) and never on batch composition, rank count or launch geometry.

Addressing (all uint32):
    key     = (seed_lo, seed_hi)
    counter = (block, slot, n, tag)
        n    : global datapoint index
        tag  : (step << 9) | (generation << 3) | purpose
        slot : operator-specific sub-stream (parent / child / (child<<8 | word))
        block: running 128-bit block index inside the sub-stream
u32 element i of a sub-stream  = philox(block=i>>2)[i&3]
u64 element i of a sub-stream  = philox(block=i>>1)[2*(i&1)] | philox(...)[2*(i&1)+1] << 32
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)

# purposes (3 bits)
PUR_SELECT = 0      # parent selection
PUR_RANDFLIP = 1    # one-distinct-bit-per-child mutation
PUR_FLIP0 = 2       # sparseflip: per-bit uniforms for bits that are 0 in the parent
PUR_FLIP1 = 3       # sparseflip: per-bit uniforms for bits that are 1 in the parent
PUR_CROSS = 4       # crossover cut points
PUR_SAMPLE = 5      # RandomSampledVarStates Bernoulli(sparsity) states


def make_tag(step, gen, purpose):
    return np.uint32(((int(step) & 0x7FFFFF) << 9) | ((int(gen) & 63) << 3) | (int(purpose) & 7))


ROUNDS = 7  # part of the stream specification (csrc/common.cuh: kPhiloxRounds)


def philox4x32_10(c0, c1, c2, c3, k0, k1, rounds=None):
    """Vectorised Philox4x32-R (R = ROUNDS unless given).  All inputs broadcast; returns 4 uint32 arrays.
    (The name is historical: rounds=10 gives Random123's philox4x32_10, pinned by its known-answer vectors.)"""
    c0, c1, c2, c3 = np.broadcast_arrays(*(np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)))
    c0, c1, c2, c3 = c0.copy(), c1.copy(), c2.copy(), c3.copy()
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(ROUNDS if rounds is None else rounds):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK32).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def _split_seed(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.uint32(seed & 0xFFFFFFFF), np.uint32(seed >> 32)


def stream_u32(seed, n, slot, tag, i):
    """u32 element ``i`` of sub-stream (n, slot, tag).  All arguments broadcast."""
    k0, k1 = _split_seed(seed)
    i = np.asarray(i, dtype=np.uint32)
    x = philox4x32_10(i >> np.uint32(2), slot, n, tag, k0, k1)
    lane = np.broadcast_to(i & np.uint32(3), x[0].shape)
    return np.choose(lane, x).astype(np.uint32)


def stream_u64(seed, n, slot, tag, i):
    """u64 element ``i`` of sub-stream (n, slot, tag)."""
    k0, k1 = _split_seed(seed)
    i = np.asarray(i, dtype=np.uint32)
    x = philox4x32_10(i >> np.uint32(1), slot, n, tag, k0, k1)
    odd = np.broadcast_to((i & np.uint32(1)).astype(bool), x[0].shape)
    lo = np.where(odd, x[2], x[0]).astype(np.uint64)
    hi = np.where(odd, x[3], x[1]).astype(np.uint64)
    return (hi << np.uint64(32)) | lo


def u64_to_unit_double(w):
    """53-bit uniform in [0,1) from a u64 word: (w >> 11) * 2^-53 (exact)."""
    return (np.asarray(w, dtype=np.uint64) >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def randint_u32(r, span):
    """Multiply-shift map of a u32 onto [0, span): (r * span) >> 32."""
    return ((np.asarray(r, dtype=np.uint64) * np.asarray(span, dtype=np.uint64)) >> np.uint64(32)).astype(np.int64)


def prob_to_fixed64(p):
    """Threshold T such that  U < p  <=>  U64 < T  for U64 = floor(U * 2^64).

    Returns (T as python int in [0, 2^64), all_ones flag).  p<=0 or NaN -> never, p>=1 -> always.
    """
    p = float(p)
    if not (p > 0.0):
        return 0, False
    if p >= 1.0:
        return 0, True
    return int(np.floor(np.ldexp(p, 64))), False


def bernoulli_mask64(seed, n, slot, tag, p, undecided):
    """64 independent Bernoulli(p) bits, restricted to the lanes set in ``undecided``.

    Lane b's uniform U_b is the binary fraction whose j-th most significant bit is bit b of u64
    stream element j; the result bit is (U_b < p).  Bits are revealed MSB first and the loop
    stops as soon as every requested lane is decided, so the expected number of stream words
    is about log2(popcount(undecided)) + 2 — the value of each lane does not depend on where
    the loop stops.  Scalar (per word) python ints; used by the small-case oracle only.
    """
    und = int(undecided) & 0xFFFFFFFFFFFFFFFF
    T, all_ones = prob_to_fixed64(p)
    if all_ones:
        return und
    res = 0
    j = 0
    while und and j < 64:
        b = 63 - j
        if (T & ((1 << (b + 1)) - 1)) == 0:
            break  # no remaining 1-bits in T: nobody still undecided can fall below it
        r = int(stream_u64(seed, n, slot, tag, j))
        if (T >> b) & 1:
            res |= und & ~r
            und &= r
        else:
            und &= ~r
        j += 1
    return res & 0xFFFFFFFFFFFFFFFF
