// tvo_b200 — warp-cooperative building blocks of the truncated E-step.
//
// One warp owns one datapoint at a time.  All per-datapoint state (y, the S old states, the S'
// children, their lpj) lives in a private slice of shared memory described by WarpBuf; states are
// bit-packed uint64 rows.  The same device functions back the stand-alone kernels (C-ABI items
// evo_generation / mark_redundant / select_topS) and the fused E-step kernel.
//
// Semantics follow the reference (cited per function) and the stream spec in oracle/philox.py.
#pragma once
#include "common.cuh"

namespace tvo {

struct EstepDims {
  int S;       // kept states per datapoint
  int Snew;    // children per E-step (all generations)
  int Sgen;    // children per generation
  int P;       // parents
  int W64;     // words per state
  int Dpad;    // padded observables (0 when y is not staged)
  int Scmax;   // max candidates a selection sees = max(S, Sgen)
  int need_fit;  // cum_p scratch needed
  int need_fmask;     // separate flip-mask buffer (sparseflip in place on crossed children); otherwise the masks
                      // are built in the children's own slots
  int RS;             // row stride (words) of the children / parents / flip-mask rows in shared memory; 0 = W64.
                      // W64 | 1 makes "lane j reads word w of row j" conflict-free (stride W64 = 16 words is a 32-way
                      // bank conflict); used by the large-H*S variant of the fused BSC kernel
  int old_in_global;  // the S old states are NOT staged in shared memory: WarpBuf::oldK points at the K row in
                      // global memory (large H*S working sets) and the top-S merge goes through a staging row
};

template <typename T>
struct WarpBuf {
  uint64_t *oldK;     // S*W64
  uint64_t *newK;     // Snew*W64
  uint64_t *parents;  // P*W64
  uint64_t *keys;     // S+Snew
  uint64_t *fmask;    // Sgen*W64 (sparseflip flip masks)
  double *dscr;       // Scmax (fitparents cum_p) or 0
  T *lpj;             // S+Snew
  uint32_t *hash;     // S+Snew
  int32_t *iscr;      // max(Scmax, 2*Sgen)
  int rs;             // row stride of newK / parents / fmask (words); oldK rows are always W64 apart
};
__host__ __device__ inline int row_stride(const EstepDims &d) { return d.RS ? d.RS : d.W64; }

// keys (select) and fmask (mutation) are never live at the same time: one region
__host__ __device__ inline size_t keys_region_bytes(const EstepDims &d) {
  const size_t k = (size_t)(d.S + d.Snew) * 8, f = d.need_fmask ? (size_t)d.Sgen * row_stride(d) * 8 : 0;
  return align16(k > f ? k : f);
}

// the parents (selection / mutation) and the dedup hashes + survivor list (dedup / top-S) are never live at the
// same time either
__host__ __device__ inline size_t parents_hash_bytes(const EstepDims &d) {
  const size_t p = (size_t)d.P * row_stride(d) * 8, h = (size_t)(((d.S + 3) & ~3) + ((d.Snew + 3) & ~3)) * 4;  // hash_len
  return align16(p > h ? p : h);
}

template <typename T>
__host__ __device__ inline size_t warp_buf_bytes(const EstepDims &d) {
  size_t b = 0;
  b += (d.old_in_global ? 0 : align16((size_t)d.S * d.W64 * 8)) + align16((size_t)d.Snew * row_stride(d) * 8) +
       parents_hash_bytes(d);
  b += keys_region_bytes(d);
  b += align16(d.need_fit ? (size_t)d.Scmax * 8 : 0);
  b += align16((size_t)(d.S + d.Snew) * sizeof(T));
  const int isz = d.Scmax > 2 * d.Sgen ? d.Scmax : 2 * d.Sgen;
  b += align16((size_t)isz * 4);
  return b;
}

template <typename T>
__device__ inline WarpBuf<T> carve_warp_buf(unsigned char *p, const EstepDims &d) {
  WarpBuf<T> w;
  w.oldK = d.old_in_global ? nullptr : (uint64_t *)p;  // old_in_global: the kernel points it at the K row
  p += d.old_in_global ? 0 : align16((size_t)d.S * d.W64 * 8);
  w.rs = row_stride(d);
  w.newK = (uint64_t *)p;
  p += align16((size_t)d.Snew * row_stride(d) * 8);
  w.parents = (uint64_t *)p;
  w.hash = (uint32_t *)p;
  p += parents_hash_bytes(d);
  w.keys = (uint64_t *)p;
  w.fmask = (uint64_t *)p;
  p += keys_region_bytes(d);
  w.dscr = (double *)p;
  p += align16(d.need_fit ? (size_t)d.Scmax * 8 : 0);
  w.lpj = (T *)p;
  p += align16((size_t)(d.S + d.Snew) * sizeof(T));
  w.iscr = (int32_t *)p;
  return w;
}

// ------------------------------------------------------------------------------------ selection
// batch_randparents (evo.py:436-451): P distinct candidates in random order.  Candidate i gets the key
// (r_i with its low `bits` bits replaced by i), r_i = u32 stream element i, bits = bit_length(S_c-1);
// the parents are the candidates with the P smallest keys, in ascending key order (keys are unique, so
// this is a uniform random P-permutation up to 2^-(32-bits) ties in the random part).  One register
// per candidate, sorted by a warp bitonic network — no serial Fisher-Yates chain on one lane.
__host__ __device__ inline int index_bits(int S_c) {
  int b = 0;
  while ((1 << b) < S_c) ++b;
  return b;
}

// UNROLL: the stage loops fully unrolled (list-based BSC kernel: half of the rolled network's instructions are loop
// control); the generic kernels keep them rolled for code size.
template <int EPL, bool UNROLL = false>
__device__ __forceinline__ void randparents_sort(int S_c, int P, int bits, const Stream &st, int32_t *chosen, int lane) {
  constexpr int MP = 32 * EPL;
  uint32_t key[EPL];
  const uint32_t himask = ~((1u << bits) - 1u);
  if (EPL >= 4) {
#pragma unroll
    for (int q = 0; q < EPL / 4; ++q) {
      const int e0 = lane * EPL + 4 * q;
      uint4 r = make_uint4(0, 0, 0, 0);
      if (e0 < S_c) r = st.block(0, (uint32_t)(e0 >> 2));
      key[4 * q + 0] = r.x;
      if (EPL > 1) key[(4 * q + 1) % EPL] = r.y;
      if (EPL > 2) key[(4 * q + 2) % EPL] = r.z;
      if (EPL > 3) key[(4 * q + 3) % EPL] = r.w;
    }
  } else {
    const int e0 = lane * EPL;
    uint4 r = make_uint4(0, 0, 0, 0);
    if (e0 < S_c) r = st.block(0, (uint32_t)(e0 >> 2));
    const int o = e0 & 3;  // EPL 1: any of 0..3; EPL 2: 0 or 2
    key[0] = o == 0 ? r.x : o == 1 ? r.y : o == 2 ? r.z : r.w;
    if (EPL == 2) key[EPL - 1] = o == 0 ? r.y : r.w;
  }
#pragma unroll
  for (int r = 0; r < EPL; ++r) {
    const int e = lane * EPL + r;
    key[r] = e < S_c ? ((key[r] & himask) | (uint32_t)e) : 0xFFFFFFFFu;
  }
#pragma unroll(UNROLL ? 16 : 1)
  for (int k = 2; k <= MP; k <<= 1) {
#pragma unroll(UNROLL ? 16 : 1)
    for (int j = k >> 1; j >= EPL; j >>= 1) {  // partner in another lane
      const int lm = j / EPL;
      // j and k are multiples of EPL here, so the direction is a property of the lane, not of the element
      const bool take_min = ((lane & lm) == 0) == ((lane & (k / EPL)) == 0);
#pragma unroll
      for (int r = 0; r < EPL; ++r) {
        const uint32_t o = __shfl_xor_sync(FULL, key[r], lm);
        key[r] = ((o < key[r]) == take_min) ? o : key[r];
      }
    }
#pragma unroll
    for (int jj = EPL >> 1; jj > 0; jj >>= 1) {  // partner in the same lane
      if (jj < k) {
#pragma unroll
        for (int r = 0; r < EPL; ++r) {
          if ((r & jj) == 0) {
            const int e = lane * EPL + r;
            const bool asc = (e & k) == 0;
            const uint32_t lo = min(key[r], key[r ^ jj]), hi = max(key[r], key[r ^ jj]);
            key[r] = asc ? lo : hi;
            key[r ^ jj] = asc ? hi : lo;
          }
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < EPL; ++r) {
    const int e = lane * EPL + r;
    if (e < P) chosen[e] = (int32_t)(key[r] & ~himask);
  }
}

// cs / ps: row strides (words) of the candidate and parent rows
__device__ inline void select_randparents(const uint64_t *cand, int S_c, int P, int W64, const Stream &st,
                                          int32_t *scr, uint64_t *parents, int lane, int cs, int ps) {
  const int bits = index_bits(S_c);
  if (S_c <= 256) {
    if (S_c <= 32)
      randparents_sort<1>(S_c, P, bits, st, scr, lane);
    else if (S_c <= 64)
      randparents_sort<2>(S_c, P, bits, st, scr, lane);
    else if (S_c <= 128)
      randparents_sort<4>(S_c, P, bits, st, scr, lane);
    else
      randparents_sort<8>(S_c, P, bits, st, scr, lane);
    __syncwarp();
    for (int t = lane; t < P * W64; t += 32) {
      const int r = t / W64, w = t - r * W64;
      parents[r * ps + w] = cand[scr[r] * cs + w];
    }
  } else {  // large candidate sets: rank every key by counting (same result, O(S_c^2 / 32) per datapoint)
    const uint32_t himask = ~((1u << bits) - 1u);
    for (int i = lane; i < S_c; i += 32) scr[i] = (int32_t)((st.u32(0, (uint32_t)i) & himask) | (uint32_t)i);
    __syncwarp();
    for (int i0 = 0; i0 < S_c; i0 += 32) {
      const int i = i0 + lane;
      const uint32_t ki = i < S_c ? (uint32_t)scr[i] : 0xFFFFFFFFu;
      int rank = 0;
      for (int j = 0; j < S_c; ++j) rank += ((uint32_t)scr[j] < ki) ? 1 : 0;
      if (i < S_c && rank < P)
        for (int w = 0; w < W64; ++w) parents[rank * ps + w] = cand[i * cs + w];
    }
  }
  __syncwarp();
}

// batch_fitparents (evo.py:397-433).  f = lpj - 2m, p = f/T, cum_p = sequential cumsum, x ~ U[0,1)
// from u64 stream element i; reference mode: chosen = S_c-1-#{x<cum_p} with -1 wrapping to S_c-1
// (quirk Q3; m and T are batch-global and come from fit_norm);  fixed mode: m, T per datapoint and
// chosen = min(S_c - #{x<cum_p}, S_c-1).
template <typename T>
__device__ inline void select_fitparents(const uint64_t *cand, const T *cand_lpj, int S_c, int P, int W64,
                                         const Stream &st, bool fixed, double m_glob, double T_glob, double *cum,
                                         int32_t *chosen, uint64_t *parents, int lane, int cs, int ps) {
  if (lane == 0) {
    double m = m_glob, Tn = T_glob;
    if (fixed) {
      double mn = 0.0;
      for (int j = 0; j < S_c; ++j) mn = fmin(mn, (double)cand_lpj[j]);
      m = (double)__double2float_rn(mn);
      Tn = 0.0;
      for (int j = 0; j < S_c; ++j) Tn = __dadd_rn(Tn, __dsub_rn((double)cand_lpj[j], __dmul_rn(2.0, m)));
    }
    double acc = 0.0;
    for (int j = 0; j < S_c; ++j) {
      const double f = __dsub_rn((double)cand_lpj[j], __dmul_rn(2.0, m));
      acc = __dadd_rn(acc, __ddiv_rn(f, Tn));
      cum[j] = acc;
    }
  }
  __syncwarp();
  for (int i = lane; i < P; i += 32) {
    const double x = unit_double(st.u64(0, (uint32_t)i));
    int count = 0;
    for (int j = 0; j < S_c; ++j) count += (x < cum[j]) ? 1 : 0;
    int c;
    if (fixed) {
      c = S_c - count;
      if (c > S_c - 1) c = S_c - 1;
    } else {
      c = S_c - 1 - count;
      if (c < 0) c += S_c;
    }
    chosen[i] = c;
  }
  __syncwarp();
  for (int t = lane; t < P * W64; t += 32) {
    const int r = t / W64, w = t - r * W64;
    parents[r * ps + w] = cand[chosen[r] * cs + w];
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------ mutation
// batch_randflip (evo.py:248-279): per parent, C DISTINCT bit positions (the reference takes the
// top-C of H uniforms = a uniform C-subset in random order) by a virtual Fisher-Yates shuffle over
// [0,H); child p*C+c = parent p with bit pos[c] flipped.  `map` holds 2*C ints per parent.
__device__ inline void mutate_randflip(const uint64_t *parents, int P, int C, int W64, int H, const Stream &st,
                                       int32_t *map, uint64_t *children, int lane, int rs) {
  for (int t = lane; t < P * C * W64; t += 32) {
    const int r = t / W64, w = t - r * W64;
    children[r * rs + w] = parents[(r / C) * rs + w];
  }
  __syncwarp();
  for (int p = lane; p < P; p += 32) {
    int32_t *mk = map + 2 * p * C, *mv = mk + C;
    int nmap = 0;
    uint4 r = make_uint4(0, 0, 0, 0);
    for (int i = 0; i < C; ++i) {
      if ((i & 3) == 0) r = st.block((uint32_t)p, i >> 2);
      const uint32_t ri = (i & 3) == 0 ? r.x : (i & 3) == 1 ? r.y : (i & 3) == 2 ? r.z : r.w;
      const int k = i + (int)randint_u32(ri, (uint32_t)(H - i));
      int vk = k, vi = i, kslot = -1;
      for (int t = 0; t < nmap; ++t) {
        if (mk[t] == k) {
          vk = mv[t];
          kslot = t;
        }
        if (mk[t] == i) vi = mv[t];
      }
      if (kslot < 0) {
        kslot = nmap++;
        mk[kslot] = k;
      }
      mv[kslot] = vi;
      children[(p * C + i) * rs + (vk >> 6)] ^= 1ull << (vk & 63);
    }
  }
  __syncwarp();
}

// p0 / p1 of batch_sparseflip (evo.py:310-316), fp64, operation for operation, no FMA contraction.
__device__ __forceinline__ void sparseflip_probs(double a, double Hf, double crowd, double p_bf, double &p0,
                                                 double &p1) {
  const double eps = 1e-100;
  const double Hp = __dmul_rn(Hf, p_bf);
  const double ca = __dsub_rn(crowd, a);
  const double num = __dmul_rn(__dsub_rn(Hf, a), __dsub_rn(Hp, ca));
  const double den = __dadd_rn(__dmul_rn(__dadd_rn(ca, Hp), a), eps);
  const double alpha = __ddiv_rn(num, den);
  p0 = __dadd_rn(__ddiv_rn(Hp, __dadd_rn(Hf, __dmul_rn(__dsub_rn(alpha, 1.0), a))), eps);
  p1 = __dmul_rn(alpha, p0);
}

// q^n by binary exponentiation, fixed operation order (oracle/evo.py:_pow_int)
__device__ __forceinline__ double pow_int(double q, int n) {
  double result = 1.0, base = q;
  int e = n;
  while (e > 0) {
    if (e & 1) result = __dmul_rn(result, base);
    e >>= 1;
    if (e) base = __dmul_rn(base, base);
  }
  return result;
}

// 1/k for k = 1..64 (IEEE division at compile time == oracle.evo.INV_TAB)
__device__ __constant__ double kInvTab[65] = {
    0.0,      1.0 / 1,  1.0 / 2,  1.0 / 3,  1.0 / 4,  1.0 / 5,  1.0 / 6,  1.0 / 7,  1.0 / 8,  1.0 / 9,  1.0 / 10,
    1.0 / 11, 1.0 / 12, 1.0 / 13, 1.0 / 14, 1.0 / 15, 1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19, 1.0 / 20, 1.0 / 21,
    1.0 / 22, 1.0 / 23, 1.0 / 24, 1.0 / 25, 1.0 / 26, 1.0 / 27, 1.0 / 28, 1.0 / 29, 1.0 / 30, 1.0 / 31, 1.0 / 32,
    1.0 / 33, 1.0 / 34, 1.0 / 35, 1.0 / 36, 1.0 / 37, 1.0 / 38, 1.0 / 39, 1.0 / 40, 1.0 / 41, 1.0 / 42, 1.0 / 43,
    1.0 / 44, 1.0 / 45, 1.0 / 46, 1.0 / 47, 1.0 / 48, 1.0 / 49, 1.0 / 50, 1.0 / 51, 1.0 / 52, 1.0 / 53, 1.0 / 54,
    1.0 / 55, 1.0 / 56, 1.0 / 57, 1.0 / 58, 1.0 / 59, 1.0 / 60, 1.0 / 61, 1.0 / 62, 1.0 / 63, 1.0 / 64};

// k ~ Binomial(n, p) by CDF inversion of u (oracle/evo.py:binomial_inverse), +,*,/ only, no FMA.
// Split into the part that depends on (n, p) only — FlipClass: pmf(0) = q^n, r = p/q and the special
// cases — and the short CDF walk that consumes u.  For sparseflip (n, p) are functions of the parent's
// popcount a alone, so the fused E-step reads FlipClass from a table built once per launch
// (flip_table_kernel) instead of redoing 2 divisions and a ~log2(H)-step power per child and class.
struct FlipClass {
  double pmf0, r;
  int32_t kconst;  // >= 0: the result regardless of u;  -1: walk the CDF
  int32_t flip;    // p > 1/2: sampled as n - Binomial(n, 1-p)
};
struct FlipEntry {
  FlipClass c[2];  // [0]: the H-a zero bits with p0, [1]: the a one bits with p1
};

__device__ __forceinline__ FlipClass flip_class_make(int n, double p) {
  FlipClass e;
  e.pmf0 = 0.0;
  e.r = 0.0;
  e.flip = 0;
  e.kconst = -1;
  if (n == 0 || !(p > 0.0)) {
    e.kconst = 0;
    return e;
  }
  if (p >= 1.0) {
    e.kconst = n;
    return e;
  }
  const bool flip = p > 0.5;
  if (flip) p = __dsub_rn(1.0, p);
  const double q = __dsub_rn(1.0, p);
  const double pmf = pow_int(q, n);
  if (pmf == 0.0) {
    const int k = (int)__dmul_rn((double)n, p);
    e.kconst = flip ? n - k : k;
    return e;
  }
  e.pmf0 = pmf;
  e.r = __ddiv_rn(p, q);
  e.flip = flip ? 1 : 0;
  return e;
}

__device__ __forceinline__ int flip_class_sample(const FlipClass &e, double u, int n) {
  if (e.kconst >= 0) return e.kconst;
  double pmf = e.pmf0, c = pmf;
  int k = 0;
  while (u >= c && k < n) {
    const double t = __dmul_rn(__dmul_rn(pmf, (double)(n - k)), e.r);
    pmf = k + 1 <= 64 ? __dmul_rn(t, kInvTab[k + 1]) : __ddiv_rn(t, (double)(k + 1));
    c = __dadd_rn(c, pmf);
    ++k;
  }
  return e.flip ? n - k : k;
}

__device__ __forceinline__ int binomial_inverse(double u, int n, double p) {
  return flip_class_sample(flip_class_make(n, p), u, n);
}

// position of the (i+1)-th set bit of a 64-bit mask (i < popcount), branch-light binary search
__device__ __forceinline__ int nth_set_bit64(uint64_t m, int i) {
  uint32_t x = (uint32_t)m;
  int base = 0;
  int c = __popc(x);
  if (i >= c) {
    i -= c;
    x = (uint32_t)(m >> 32);
    base = 32;
  }
  c = __popc(x & 0xFFFFu);
  if (i >= c) { i -= c; x >>= 16; base += 16; }
  c = __popc(x & 0xFFu);
  if (i >= c) { i -= c; x >>= 8; base += 8; }
  c = __popc(x & 0xFu);
  if (i >= c) { i -= c; x >>= 4; base += 4; }
  c = __popc(x & 0x3u);
  if (i >= c) { i -= c; x >>= 2; base += 2; }
  c = (int)(x & 1u);
  if (i >= c) base += 1;
  return base;
}

// position (0..H-1) of the idx-th set bit of the multi-word mask  m_w = cand_w(src) & ~fm_w
__device__ __forceinline__ int nth_remaining(const uint64_t *src, const uint64_t *fm, int W64, int H, int cls, int idx) {
  for (int w = 0; w < W64; ++w) {
    const int nb = H - 64 * w;
    const uint64_t valid = nb >= 64 ? ~0ull : ((1ull << nb) - 1);
    const uint64_t m = (cls ? src[w] : (~src[w] & valid)) & ~fm[w];
    const int c = __popcll(m);
    if (idx < c) return (w << 6) + nth_set_bit64(m, idx);
    idx -= c;
  }
  return 0;  // unreachable for idx < #remaining
}

// batch_sparseflip (evo.py:282-331): child j = parent j/C with every bit flipped independently with
// probability p1 if the bit is set in the parent, p0 otherwise.  Sampled exactly as: number of flips
// among the zero bits ~ Binomial(H-a, p0), among the one bits ~ Binomial(a, p1); positions = uniform
// subset drawn without replacement (dense class, 2n >= H: rejection over all H positions; sparse class:
// draw = the t-th remaining candidate; one u32 per attempt).  One LANE per child, one Philox block for
// each count plus one per 4 attempts, instead of H uniforms per child.
// `fmask` (nchild x W64 words) collects the flips; dst = src ^ fmask, so src == dst (C == 1) is fine.
__device__ inline void mutate_sparseflip(const uint64_t *src, int P, int C, int W64, int H, double sparsity,
                                         double p_bf, const Stream &st0, const Stream &st1, uint64_t *fmask,
                                         uint64_t *dst, int lane, int rs, const FlipEntry *__restrict__ tab = nullptr) {
  const double Hf = (double)H;
  const double crowd = __dmul_rn(sparsity, Hf);
  const int nchild = P * C;
  for (int j0 = 0; j0 < nchild; j0 += 32) {
    const int j = j0 + lane;
    if (j < nchild) {
      const uint64_t *par = src + (size_t)(j / C) * rs;
      uint64_t *fm = fmask + (size_t)j * rs;
      int a = 0;
      for (int w = 0; w < W64; ++w) {
        a += __popcll(par[w]);
        fm[w] = 0;
      }
      double p0 = 0.0, p1 = 0.0;
      if (!tab) sparseflip_probs((double)a, Hf, crowd, p_bf, p0, p1);
      // ONE sub-stream per child (slot j of st0): block 0 holds the two count uniforms (u64 element 0 for
      // the zero bits, element 1 for the one bits), u32 elements 4, 5, ... feed the position attempts of
      // both classes in turn — typically two Philox blocks per child in total
      uint4 b0 = make_uint4(0, 0, 0, 0), r = make_uint4(0, 0, 0, 0);
      bool have_b0 = false;
      uint32_t t = 0;
#pragma unroll 1
      for (int cls = 0; cls < 2; ++cls) {
        const int n = cls ? a : H - a;
        FlipClass fc;
        if (tab) {
          const FlipClass *tp = &tab[a].c[cls];
          fc.pmf0 = __ldg(&tp->pmf0);
          fc.r = __ldg(&tp->r);
          fc.kconst = __ldg(&tp->kconst);
          fc.flip = __ldg(&tp->flip);
        } else {
          fc = flip_class_make(n, cls ? p1 : p0);
        }
        const Stream &st = st0;
        int k = fc.kconst;
        if (k < 0) {
          if (!have_b0) {
            b0 = st.block((uint32_t)j, 0);
            have_b0 = true;
          }
          const uint64_t uw = cls ? ((uint64_t)b0.w << 32 | b0.z) : ((uint64_t)b0.y << 32 | b0.x);
          k = flip_class_sample(fc, unit_double(uw), n);
        }
        if (k == 0) continue;
        const bool complement = 2 * k > n;  // draw the smaller of {flipped, not flipped}
        const int draws = k >= n ? 0 : (complement ? n - k : k);
        int got = 0;
        if (2 * n >= H) {  // dense class: rejection sampling over all H positions
          const uint32_t cap = t + 64u + 64u * (uint32_t)draws;
          while (got < draws && t < cap) {
            if ((t & 3) == 0) r = st.block((uint32_t)j, 1 + (t >> 2));
            const uint32_t ri = (t & 3) == 0 ? r.x : (t & 3) == 1 ? r.y : (t & 3) == 2 ? r.z : r.w;
            ++t;
            const int pos = (int)randint_u32(ri, (uint32_t)H);
            const uint64_t bit = 1ull << (pos & 63);
            const uint64_t pw = par[pos >> 6], fw = fm[pos >> 6];
            if (((pw & bit) != 0) == (cls != 0) && !(fw & bit)) {
              fm[pos >> 6] = fw | bit;
              ++got;
            }
          }
        }
        for (; got < draws; ++got) {  // sparse class: the idx-th remaining candidate
          if ((t & 3) == 0) r = st.block((uint32_t)j, 1 + (t >> 2));
          const uint32_t ri = (t & 3) == 0 ? r.x : (t & 3) == 1 ? r.y : (t & 3) == 2 ? r.z : r.w;
          ++t;
          const int pos = nth_remaining(par, fm, W64, H, cls, (int)randint_u32(ri, (uint32_t)(n - got)));
          fm[pos >> 6] |= 1ull << (pos & 63);
        }
        if (k >= n || complement) {  // flip every candidate of the class that was NOT drawn
          for (int w = 0; w < W64; ++w) {
            const int nb = H - 64 * w;
            const uint64_t valid = nb >= 64 ? ~0ull : ((1ull << nb) - 1);
            const uint64_t cand = cls ? par[w] : (~par[w] & valid);
            fm[w] = (fm[w] & ~cand) | (cand & ~fm[w]);
          }
        }
      }
      for (int w = 0; w < W64; ++w) dst[(size_t)j * rs + w] = par[w] ^ fm[w];
    }
    __syncwarp();
  }
}

// batch_cross (evo.py:335-380): pairs in combinations(range(P),2) order, cut ~ U{1..H-1} from u32
// stream element k; child 2k = p1[:cut] | p2[cut:], child 2k+1 = p2[:cut] | p1[cut:].
__device__ inline void mutate_cross(const uint64_t *parents, int P, int W64, int H, const Stream &st,
                                    uint64_t *children, int lane, int rs) {
  const int npairs = P * (P - 1) / 2;
  for (int t = lane; t < npairs * W64; t += 32) {
    const int k = t / W64, w = t % W64;
    int a = 0, rem = k;  // k-th pair (a,b), a<b, lexicographic
    while (rem >= P - 1 - a) {
      rem -= P - 1 - a;
      ++a;
    }
    const int b = a + 1 + rem;
    const int cut = 1 + (int)randint_u32(st.u32(0, (uint32_t)k), (uint32_t)(H - 1));
    const int lo = cut - 64 * w;
    const uint64_t L = lo <= 0 ? 0ull : (lo >= 64 ? ~0ull : ((1ull << lo) - 1));
    const uint64_t pa = parents[a * rs + w], pb = parents[b * rs + w];
    children[(2 * k) * rs + w] = (pa & L) | (pb & ~L);
    children[(2 * k + 1) * rs + w] = (pb & L) | (pa & ~L);
  }
  __syncwarp();
}

struct EvoParams {
  int sel, mut, P, C, G;
  double p_bf;
  uint2 key;
  uint32_t step, n_offset;
  const FlipEntry *flip_tab;  // sparseflip: per-popcount table (flip_table_kernel) or nullptr = compute inline
  const double *fit_norm;     // batch_fitparents: device {m, T} of the batch (tvo_fit_norm); generation 0 only
  int lpj_current;            // the lpj rows already hold lpj_fn(batch, K) under the current theta: no re-score
};

// shared argument check of the fused E-steps for the reference's batch-global fitness selection
#define TVO_REQUIRE_FITPARENTS(ep, who)                                                                              \
  TVO_REQUIRE((ep).sel != TVO_SEL_FITPARENTS || ((ep).fit_norm && (ep).lpj_current && (ep).G == 1),                  \
              "%s: batch-global fitparents runs fused for ONE generation and needs cfg->fit_norm (tvo_fit_norm of "   \
              "the re-scored batch) with TVO_EVO_LPJ_CURRENT; otherwise use the stand-alone kernels", who)

// tab[a], a = 0..H: the two FlipClass entries of a parent with a active units
static __global__ void flip_table_kernel(FlipEntry *tab, const double *__restrict__ sparsity_p, int H, double p_bf) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a > H) return;
  const double Hf = (double)H;
  double p0, p1;
  sparseflip_probs((double)a, Hf, __dmul_rn(*sparsity_p, Hf), p_bf, p0, p1);
  tab[a].c[0] = flip_class_make(H - a, p0);
  tab[a].c[1] = flip_class_make(a, p1);
}
__host__ inline size_t flip_table_bytes(int H) { return (((size_t)(H + 1) * sizeof(FlipEntry)) + 255) & ~size_t(255); }

__host__ inline int children_per_gen(int mut, int P, int C) { return mut >= TVO_MUT_CROSS ? P * (P - 1) : P * C; }

// One generation: select P parents from `cand`, mutate into `children` (Sgen rows).
template <typename T>
__device__ inline void evo_generation(const WarpBuf<T> &wb, const uint64_t *cand, const T *cand_lpj, int S_c,
                                      uint64_t *children, const EvoParams &ep, int gen, uint32_t n_glob, int W64,
                                      int H, double sparsity, double m_glob, double T_glob, int lane, int cs = 0) {
  if (cs == 0) cs = W64;  // row stride of `cand`; the children are written wb.rs words apart
  const int rs = wb.rs;
  Stream st;
  st.key = ep.key;
  st.n = n_glob;
  st.tag = make_tag(ep.step, gen, PUR_SELECT);
  if (ep.sel == TVO_SEL_RANDPARENTS)
    select_randparents(cand, S_c, ep.P, W64, st, wb.iscr, wb.parents, lane, cs, rs);
  else
    select_fitparents<T>(cand, cand_lpj, S_c, ep.P, W64, st, ep.sel == TVO_SEL_FITPARENTS_FIXED, m_glob, T_glob,
                         wb.dscr, wb.iscr, wb.parents, lane, cs, rs);
  Stream s0 = st, s1 = st;
  s0.tag = make_tag(ep.step, gen, PUR_FLIP0);
  s1.tag = make_tag(ep.step, gen, PUR_FLIP1);
  switch (ep.mut) {
    case TVO_MUT_RANDFLIP:
      st.tag = make_tag(ep.step, gen, PUR_RANDFLIP);
      mutate_randflip(wb.parents, ep.P, ep.C, W64, H, st, wb.iscr, children, lane, rs);
      break;
    case TVO_MUT_SPARSEFLIP:
      mutate_sparseflip(wb.parents, ep.P, ep.C, W64, H, sparsity, ep.p_bf, s0, s1, /*fmask=*/children, children, lane,
                        rs, ep.flip_tab);  // the masks are built in the children's own slots
      break;
    default: {
      st.tag = make_tag(ep.step, gen, PUR_CROSS);
      mutate_cross(wb.parents, ep.P, W64, H, st, children, lane, rs);
      const int nc = ep.P * (ep.P - 1);
      if (ep.mut == TVO_MUT_CROSS_RANDFLIP) {
        // randflip with C = 1 on every crossed child (evo.py:383-387): flip bit randint(H)
        st.tag = make_tag(ep.step, gen, PUR_RANDFLIP);
        for (int j = lane; j < nc; j += 32) {
          const int k = (int)randint_u32(st.u32((uint32_t)j, 0), (uint32_t)H);
          children[j * rs + (k >> 6)] ^= 1ull << (k & 63);
        }
        __syncwarp();
      } else if (ep.mut == TVO_MUT_CROSS_SPARSEFLIP) {
        mutate_sparseflip(children, nc, 1, W64, H, sparsity, ep.p_bf, s0, s1, wb.fmask, children, lane, rs,
                          ep.flip_tab);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ dedup
// set_redundant_lpj_to_low (_set_redundant_lpj_to_low_CPU.pyx:13-30): child c is redundant if it
// equals a LATER child, else if it equals an old state.  Marks wb.lpj[S+c] = (float)-1e20.
static __device__ __noinline__ bool same_state(const uint64_t *x, const uint64_t *y, int W64) {
  bool eq = true;
  for (int w = 0; w < W64; ++w) eq &= (x[w] == y[w]);
  return eq;
}

// Child c is redundant if an old state or a LATER child has the same bits.
//
// Hash scan with one deferred full compare.  The 32-bit hashes (forced odd, so the zero padding never
// matches) live in two 16-byte-aligned runs: S old hashes padded to a multiple of 4, then Snew child
// hashes padded likewise.  Each lane owns CPL children (c = c0 + 32 t + lane) and remembers ANY
// matching candidate:
//   * old states: broadcast LDS.128 of 4 hashes, 2 instructions per (hash, child) pair;
//   * later children of a LATER 32-group: same scan, no index condition needed;
//   * later children of the SAME 32-group: one __match_any_sync on the hash register.
// One full state compare per child confirms the match; only a genuine hash collision takes the
// exhaustive rescan.
__host__ __device__ inline int hash_old_pad(int S) { return (S + 3) & ~3; }
__host__ __device__ inline int hash_len(int S, int Snew) { return hash_old_pad(S) + ((Snew + 3) & ~3); }

#define TVO_CMP4(h4, base)                      \
  _Pragma("unroll") for (int t = 0; t < CPL; ++t) { \
    if ((h4).x == hc[t]) hit[t] = (base);       \
    if ((h4).y == hc[t]) hit[t] = (base) + 1;   \
    if ((h4).z == hc[t]) hit[t] = (base) + 2;   \
    if ((h4).w == hc[t]) hit[t] = (base) + 3;   \
  }

template <typename T, int CPL>
__device__ __forceinline__ void mark_redundant_cpl(const WarpBuf<T> &wb, int S, int Snew, int W64, int lane) {
  const int Sp = hash_old_pad(S), Snp = (Snew + 3) & ~3;
  const uint32_t *ho = wb.hash, *hn = wb.hash + Sp;
#pragma unroll 1
  for (int c0 = 0; c0 < Snew; c0 += 32 * CPL) {
    uint32_t hc[CPL];
    int hit[CPL];  // concatenated index (old: j, child: S + j) of a candidate with the same hash
#pragma unroll
    for (int t = 0; t < CPL; ++t) {
      const int c = c0 + 32 * t + lane;
      hc[t] = c < Snew ? hn[c] : 0u;
      hit[t] = -1;
    }
#pragma unroll 2
    for (int j = 0; j < Sp; j += 4) {
      const uint4 h4 = *reinterpret_cast<const uint4 *>(ho + j);
      TVO_CMP4(h4, j)
    }
#pragma unroll
    for (int g = 1; g < CPL; ++g) {  // children of later 32-groups of this sweep
      const int jb = c0 + 32 * g, je = min(jb + 32, Snp);
      for (int j = jb; j < je; j += 4) {
        const uint4 h4 = *reinterpret_cast<const uint4 *>(hn + j);
#pragma unroll
        for (int t = 0; t < CPL; ++t) {
          if (t < g) {
            if (h4.x == hc[t]) hit[t] = S + j;
            if (h4.y == hc[t]) hit[t] = S + j + 1;
            if (h4.z == hc[t]) hit[t] = S + j + 2;
            if (h4.w == hc[t]) hit[t] = S + j + 3;
          }
        }
      }
    }
    for (int j = c0 + 32 * CPL; j < Snp; j += 4) {  // children of later sweeps (Snew > 32*CPL only)
      const uint4 h4 = *reinterpret_cast<const uint4 *>(hn + j);
      TVO_CMP4(h4, S + j)
    }
#pragma unroll
    for (int t = 0; t < CPL; ++t) {  // later children of the same 32-group
      const unsigned later = __match_any_sync(FULL, hc[t]) & ~((2u << lane) - 1u);
      if (later && hc[t]) hit[t] = S + c0 + 32 * t + __ffs(later) - 1;
    }
#pragma unroll
    for (int t = 0; t < CPL; ++t) {
      const int c = c0 + 32 * t + lane;
      if (c < Snew && hit[t] >= 0) {
        const uint64_t *sc = wb.newK + (size_t)c * wb.rs;
        const int jj = hit[t];
        bool dup = same_state(jj < S ? wb.oldK + (size_t)jj * W64 : wb.newK + (size_t)(jj - S) * wb.rs, sc, W64);
        if (!dup) {  // hash collision: exhaustive rescan (practically never)
          for (int j2 = 0; j2 < S && !dup; ++j2)
            if (ho[j2] == hc[t]) dup = same_state(wb.oldK + (size_t)j2 * W64, sc, W64);
          for (int j2 = c + 1; j2 < Snew && !dup; ++j2)
            if (hn[j2] == hc[t]) dup = same_state(wb.newK + (size_t)j2 * wb.rs, sc, W64);
        }
        if (dup) wb.lpj[S + c] = LowLpj<T>::v();
      }
    }
  }
}
#undef TVO_CMP4

template <typename T>
__device__ inline void mark_redundant(const WarpBuf<T> &wb, int S, int Snew, int W64, int lane) {
  const int Sp = hash_old_pad(S), Snp = (Snew + 3) & ~3;
  for (int i = lane; i < Sp; i += 32) wb.hash[i] = i < S ? (hash_words(wb.oldK + (size_t)i * W64, W64) | 1u) : 0u;
  for (int i = lane; i < Snp; i += 32)
    wb.hash[Sp + i] = i < Snew ? (hash_words(wb.newK + (size_t)i * wb.rs, W64) | 1u) : 0u;
  __syncwarp();
  if (Snew <= 32)
    mark_redundant_cpl<T, 1>(wb, S, Snew, W64, lane);
  else if (Snew <= 64)
    mark_redundant_cpl<T, 2>(wb, S, Snew, W64, lane);
  else
    mark_redundant_cpl<T, 4>(wb, S, Snew, W64, lane);
  __syncwarp();
}

// ------------------------------------------------------------------------------------ top-S merge
// update_states_for_batch (_utils.py:124-179): keep the S best of old+new, stored ascending (best
// last); ties -> lower concatenated index ranks higher.  Writes the datapoint's K and lpj rows in
// global memory and returns the number of selected children (warp-uniform).
// "x ranks before y": larger key first, ties -> lower concatenated index first
__device__ __forceinline__ bool ranks_before(uint64_t kx, int ix, uint64_t ky, int iy) {
  // (kx, ~ix) > (ky, ~iy) as one 96-bit unsigned compare: Y - X borrows.  The low word ~iy - ~ix equals
  // ix - iy and borrows iff ix < iy, so the indices enter the chain un-negated (both are >= 0).
  uint32_t b;
  asm("{\n\t.reg .u32 t;\n\t"
      "sub.cc.u32 t, %1, %2;\n\t"
      "subc.cc.u32 t, %3, %4;\n\t"
      "subc.cc.u32 t, %5, %6;\n\t"
      "subc.u32 %0, 0, 0;\n\t}"
      : "=r"(b)
      : "r"((uint32_t)ix), "r"((uint32_t)iy), "r"((uint32_t)ky), "r"((uint32_t)kx), "r"((uint32_t)(ky >> 32)),
        "r"((uint32_t)(kx >> 32)));
  return b != 0;
}

// Warp bitonic network over (key, src) pairs, best-first; EPL elements per lane, element e =
// lane*EPL + r; distances < EPL are register-local, larger ones use shuffles.  Runs the stages
// k = k_first, 2 k_first, ..., 32*EPL:  k_first = 2 is a full sort, k_first = 32*EPL only merges a
// bitonic sequence.  Stage loops are kept rolled (code size: the unrolled network is 1.3k instructions
// per EPL and the fused kernels live or die by the instruction cache).
template <int EPL>
__device__ __forceinline__ void bitonic_network(uint64_t (&key)[EPL], int (&src)[EPL], int lane, int k_first) {
  constexpr int MP = 32 * EPL;
#pragma unroll 1
  for (int k = k_first; k <= MP; k <<= 1) {
#pragma unroll 1
    for (int j = k >> 1; j >= EPL; j >>= 1) {  // partner in another lane
      const int lm = j / EPL;
      // j and k are multiples of EPL here, so the direction is a property of the lane, not of the element
      const bool want_better = ((lane & lm) == 0) == ((lane & (k / EPL)) == 0);
#pragma unroll
      for (int r = 0; r < EPL; ++r) {
        const uint64_t ok = __shfl_xor_sync(FULL, key[r], lm);
        const int oi = __shfl_xor_sync(FULL, src[r], lm);
        if (ranks_before(ok, oi, key[r], src[r]) == want_better) {
          key[r] = ok;
          src[r] = oi;
        }
      }
    }
#pragma unroll
    for (int jj = EPL >> 1; jj > 0; jj >>= 1) {  // partner in the same lane (static register indices)
      if (jj < k) {
#pragma unroll
        for (int r = 0; r < EPL; ++r) {
          if ((r & jj) == 0) {
            const int e = lane * EPL + r;
            const bool up = (e & k) == 0;  // best-first block
            if (ranks_before(key[r ^ jj], src[r ^ jj], key[r], src[r]) == up) {
              const uint64_t tk = key[r];
              key[r] = key[r ^ jj];
              key[r ^ jj] = tk;
              const int ti = src[r];
              src[r] = src[r ^ jj];
              src[r ^ jj] = ti;
            }
          }
        }
      }
    }
  }
}

constexpr int kPadSrc = 0x7fffffff;  // padding element: key 0 (below every real key), ranks last

// Sort the ns surviving children (indices in `slist`, keys in wb.keys) best-first, in place.
template <typename T, int EPL>
__device__ __forceinline__ void sort_survivors(const WarpBuf<T> &wb, int32_t *slist, int ns, int lane) {
  uint64_t key[EPL];
  int src[EPL];
#pragma unroll
  for (int r = 0; r < EPL; ++r) {
    const int e = lane * EPL + r;
    src[r] = e < ns ? slist[e] : kPadSrc;
    key[r] = e < ns ? wb.keys[src[r]] : 0ull;
  }
  __syncwarp();
  bitonic_network<EPL>(key, src, lane, 2);
#pragma unroll
  for (int r = 0; r < EPL; ++r) {
    const int e = lane * EPL + r;
    if (e < ns) slist[e] = src[r];
  }
  __syncwarp();
}

// Top-S of old + new, best-first in registers, written out ascending.  Three cheaper pieces instead of
// one sort of all M = S + Snew candidates:
//   1. sort the S old states (network of 32*EPL >= S elements);
//   2. only children that beat the worst old state can enter the top S: compact them by ballot (late
//      in training that is a handful of the Snew children) and sort them;
//   3. C[e] = better(old[e], survivor[MP-1-e]) is bitonic and holds the top MP: one merge stage.
// Returns false (nothing written) when more than 256 children survive: the caller ranks by counting.
template <typename T, int EPL>
__device__ __forceinline__ bool select_merge(const WarpBuf<T> &wb, int S, int Snew, T *lpj_row, int lane, int &nsub) {
  constexpr int MP = 32 * EPL;
  uint64_t key[EPL];
  int src[EPL];
#pragma unroll
  for (int r = 0; r < EPL; ++r) {
    const int e = lane * EPL + r;
    key[r] = e < S ? wb.keys[e] : 0ull;
    src[r] = e < S ? e : kPadSrc;
  }
  bitonic_network<EPL>(key, src, lane, 2);
  uint64_t wkey = 0;
#pragma unroll
  for (int r = 0; r < EPL; ++r)
    if (((S - 1) % EPL) == r) wkey = key[r];
  wkey = __shfl_sync(FULL, wkey, (S - 1) / EPL);
  int32_t *slist = (int32_t *)wb.hash;  // free after mark_redundant; holds >= S + Snew ints
  int ns = 0;
  for (int c0 = 0; c0 < Snew; c0 += 32) {
    const int c = c0 + lane;
    const bool keep = c < Snew && wb.keys[S + c] > wkey;  // ties go to the old state (lower index)
    const unsigned m = __ballot_sync(FULL, keep);
    if (keep) slist[ns + __popc(m & ((1u << lane) - 1u))] = S + c;
    ns += __popc(m);
  }
  __syncwarp();
  if (ns > 256) return false;
  if (ns > 0) {
    if (ns <= 32)
      sort_survivors<T, 1>(wb, slist, ns, lane);
    else if (ns <= 64)
      sort_survivors<T, 2>(wb, slist, ns, lane);
    else if (ns <= 128)
      sort_survivors<T, 4>(wb, slist, ns, lane);
    else
      sort_survivors<T, 8>(wb, slist, ns, lane);
#pragma unroll
    for (int r = 0; r < EPL; ++r) {
      const int bi = MP - 1 - (lane * EPL + r);
      if (bi < ns) {
        const int bs = slist[bi];
        const uint64_t bk = wb.keys[bs];
        if (ranks_before(bk, bs, key[r], src[r])) {
          key[r] = bk;
          src[r] = bs;
        }
      }
    }
    bitonic_network<EPL>(key, src, lane, MP);
  }
  nsub = 0;
#pragma unroll
  for (int r = 0; r < EPL; ++r) {
    const int e = lane * EPL + r;
    if (e < S) {
      const int pos = S - 1 - e;
      lpj_row[pos] = wb.lpj[src[r]];
      wb.iscr[pos] = src[r];
      nsub += (src[r] >= S) ? 1 : 0;
    }
  }
  return true;
}

// `stage` (S*W64 words, private to the warp) is required when wb.oldK aliases K_row (old states not staged in
// shared memory): the selected rows are assembled there first, then copied over the K row.
template <typename T, bool BATCHED = false>
__device__ inline int select_topS(const WarpBuf<T> &wb, int S, int Snew, int W64, uint64_t *K_row, T *lpj_row,
                                  int lane, uint64_t *stage = nullptr) {
  const int M = S + Snew;
  for (int i = lane; i < M; i += 32) wb.keys[i] = sort_key((double)wb.lpj[i]);
  __syncwarp();
  int nsub = 0;
  bool done = false;
  if (S <= 32)
    done = select_merge<T, 1>(wb, S, Snew, lpj_row, lane, nsub);
  else if (S <= 64)
    done = select_merge<T, 2>(wb, S, Snew, lpj_row, lane, nsub);
  else if (S <= 128)
    done = select_merge<T, 4>(wb, S, Snew, lpj_row, lane, nsub);
  else if (S <= 256)
    done = select_merge<T, 8>(wb, S, Snew, lpj_row, lane, nsub);
  if (!done) {
    for (int i0 = 0; i0 < M; i0 += 32) {  // rank by counting
      const int i = i0 + lane;
      const uint64_t ki = i < M ? wb.keys[i] : 0;
      int rank = 0;
      for (int j = 0; j < M; ++j) rank += ranks_before(wb.keys[j], j, ki, i) ? 1 : 0;
      if (i < M && rank < S) {
        const int pos = S - 1 - rank;
        lpj_row[pos] = wb.lpj[i];
        wb.iscr[pos] = i;  // iscr has >= S entries: Scmax >= S
        nsub += (i >= S) ? 1 : 0;
      }
    }
  }
  __syncwarp();
  // Gather the selected rows.  Loads are issued in batches of kCopyBatch before the stores: source (old states
  // in global memory for large H*S) and destination may alias as far as the compiler knows, and a load -> store
  // chain per word leaves one L2 round trip in flight per warp (BATCHED: the large-H*S variant of the fused kernel).
  constexpr int kCopyBatch = BATCHED ? 4 : 1;
  uint64_t *dst = stage ? stage : K_row;
  const int nw = S * W64;
  const int wsh = (W64 & (W64 - 1)) == 0 ? __ffs(W64) - 1 : -1;  // t / W64 as a shift when W64 is a power of two
  for (int t0 = lane; t0 < nw; t0 += 32 * kCopyBatch) {
    uint64_t v[kCopyBatch];
#pragma unroll
    for (int u = 0; u < kCopyBatch; ++u) {
      const int t = t0 + 32 * u;
      if (t < nw) {
        const int r = wsh >= 0 ? (t >> wsh) : t / W64, src = wb.iscr[r];
        const uint64_t *row = src < S ? wb.oldK + (size_t)src * W64 : wb.newK + (size_t)(src - S) * wb.rs;
        v[u] = row[t - r * W64];
      }
    }
#pragma unroll
    for (int u = 0; u < kCopyBatch; ++u)
      if (t0 + 32 * u < nw) dst[t0 + 32 * u] = v[u];
  }
  __syncwarp();
  if (stage) {
    for (int t0 = lane; t0 < nw; t0 += 32 * kCopyBatch) {
      uint64_t v[kCopyBatch];
#pragma unroll
      for (int u = 0; u < kCopyBatch; ++u)
        if (t0 + 32 * u < nw) v[u] = stage[t0 + 32 * u];
#pragma unroll
      for (int u = 0; u < kCopyBatch; ++u)
        if (t0 + 32 * u < nw) K_row[t0 + 32 * u] = v[u];
    }
    __syncwarp();
  }
  return warp_sum(nsub);
}

}  // namespace tvo
