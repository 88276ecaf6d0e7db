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
};

template <typename T>
struct WarpBuf {
  uint64_t *oldK;     // S*W64
  uint64_t *newK;     // Snew*W64
  uint64_t *parents;  // P*W64
  uint64_t *keys;     // S+Snew
  uint64_t *thr;      // 2*Sgen   (T0 | T1)
  double *dscr;       // Scmax (fitparents cum_p) or 0
  T *lpj;             // S+Snew
  uint32_t *hash;     // S+Snew
  int32_t *iscr;      // max(Scmax, 2*Sgen)
  uint8_t *flags;     // Sgen
};

template <typename T>
__host__ __device__ inline size_t warp_buf_bytes(const EstepDims &d) {
  size_t b = 0;
  b += align16((size_t)d.S * d.W64 * 8) + align16((size_t)d.Snew * d.W64 * 8) + align16((size_t)d.P * d.W64 * 8);
  b += align16((size_t)(d.S + d.Snew) * 8);
  b += align16((size_t)2 * d.Sgen * 8);
  b += align16(d.need_fit ? (size_t)d.Scmax * 8 : 0);
  b += align16((size_t)(d.S + d.Snew) * sizeof(T));
  b += align16((size_t)(d.S + d.Snew) * 4);
  const int isz = d.Scmax > 2 * d.Sgen ? d.Scmax : 2 * d.Sgen;
  b += align16((size_t)isz * 4);
  b += align16((size_t)d.Sgen);
  return b;
}

template <typename T>
__device__ inline WarpBuf<T> carve_warp_buf(unsigned char *p, const EstepDims &d) {
  WarpBuf<T> w;
  w.oldK = (uint64_t *)p;
  p += align16((size_t)d.S * d.W64 * 8);
  w.newK = (uint64_t *)p;
  p += align16((size_t)d.Snew * d.W64 * 8);
  w.parents = (uint64_t *)p;
  p += align16((size_t)d.P * d.W64 * 8);
  w.keys = (uint64_t *)p;
  p += align16((size_t)(d.S + d.Snew) * 8);
  w.thr = (uint64_t *)p;
  p += align16((size_t)2 * d.Sgen * 8);
  w.dscr = (double *)p;
  p += align16(d.need_fit ? (size_t)d.Scmax * 8 : 0);
  w.lpj = (T *)p;
  p += align16((size_t)(d.S + d.Snew) * sizeof(T));
  w.hash = (uint32_t *)p;
  p += align16((size_t)(d.S + d.Snew) * 4);
  const int isz = d.Scmax > 2 * d.Sgen ? d.Scmax : 2 * d.Sgen;
  w.iscr = (int32_t *)p;
  p += align16((size_t)isz * 4);
  w.flags = (uint8_t *)p;
  return w;
}

// ------------------------------------------------------------------------------------ selection
// batch_randparents (evo.py:436-451): first P entries of a uniform random permutation of the S_c
// candidates, by a partial Fisher-Yates shuffle driven by u32 stream elements 0..P-1.
__device__ inline void select_randparents(const uint64_t *cand, int S_c, int P, int W64, const Stream &st,
                                          int32_t *perm, uint64_t *parents, int lane) {
  for (int i = lane; i < S_c; i += 32) perm[i] = i;
  __syncwarp();
  if (lane == 0) {
    uint4 r = make_uint4(0, 0, 0, 0);
    for (int i = 0; i < P; ++i) {
      if ((i & 3) == 0) r = st.block(0, i >> 2);
      const uint32_t ri = (i & 3) == 0 ? r.x : (i & 3) == 1 ? r.y : (i & 3) == 2 ? r.z : r.w;
      const int k = i + (int)randint_u32(ri, (uint32_t)(S_c - i));
      const int32_t t = perm[i];
      perm[i] = perm[k];
      perm[k] = t;
    }
  }
  __syncwarp();
  for (int t = lane; t < P * W64; t += 32) parents[t] = cand[perm[t / W64] * W64 + t % W64];
  __syncwarp();
}

// batch_fitparents (evo.py:397-433).  f = lpj - 2m, p = f/T, cum_p = sequential cumsum, x ~ U[0,1)
// from u64 stream element i; reference mode: chosen = S_c-1-#{x<cum_p} with -1 wrapping to S_c-1
// (quirk Q3; m and T are batch-global and come from fit_norm);  fixed mode: m, T per datapoint and
// chosen = min(S_c - #{x<cum_p}, S_c-1).
template <typename T>
__device__ inline void select_fitparents(const uint64_t *cand, const T *cand_lpj, int S_c, int P, int W64,
                                         const Stream &st, bool fixed, double m_glob, double T_glob, double *cum,
                                         int32_t *chosen, uint64_t *parents, int lane) {
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
  for (int t = lane; t < P * W64; t += 32) parents[t] = cand[chosen[t / W64] * W64 + t % W64];
  __syncwarp();
}

// ------------------------------------------------------------------------------------ mutation
// batch_randflip (evo.py:248-279): per parent, C DISTINCT bit positions (the reference takes the
// top-C of H uniforms = a uniform C-subset in random order) by a virtual Fisher-Yates shuffle over
// [0,H); child p*C+c = parent p with bit pos[c] flipped.  `map` holds 2*C ints per parent.
__device__ inline void mutate_randflip(const uint64_t *parents, int P, int C, int W64, int H, const Stream &st,
                                       int32_t *map, uint64_t *children, int lane) {
  for (int t = lane; t < P * C * W64; t += 32) children[t] = parents[(t / W64 / C) * W64 + t % W64];
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
      children[(p * C + i) * W64 + (vk >> 6)] ^= 1ull << (vk & 63);
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

// batch_sparseflip (evo.py:282-331): child j = parent j/C with every bit flipped independently,
// probability p1 if the bit is set in the parent, p0 otherwise.  May run in place (src == dst, C == 1).
__device__ inline void mutate_sparseflip(const uint64_t *src, int P, int C, int W64, int H, double sparsity,
                                         double p_bf, const Stream &st0, const Stream &st1, uint64_t *thr,
                                         uint8_t *flags, uint64_t *dst, int lane) {
  const double Hf = (double)H;
  const double crowd = __dmul_rn(sparsity, Hf);
  for (int p = lane; p < P; p += 32) {
    int a = 0;
    for (int w = 0; w < W64; ++w) a += __popcll(src[p * W64 + w]);
    double p0, p1;
    sparseflip_probs((double)a, Hf, crowd, p_bf, p0, p1);
    uint64_t T0, T1;
    uint32_t a0, a1;
    prob_to_fixed64(p0, T0, a0);
    prob_to_fixed64(p1, T1, a1);
    thr[2 * p] = T0;
    thr[2 * p + 1] = T1;
    flags[p] = (uint8_t)(a0 | (a1 << 1));
  }
  __syncwarp();
  // tasks (child j, word w); descending so that in-place C==1 and out-of-place C>1 both read
  // their source word before any task overwrites it (each task only touches word (j,w)).
  const int ntask = P * C * W64;
  for (int t0 = 0; t0 < ntask; t0 += 32) {
    const int t = t0 + lane;
    uint64_t out = 0;
    if (t < ntask) {
      const int j = t / W64, w = t % W64, p = j / C;
      const uint64_t s = src[p * W64 + w];
      const int nb = H - 64 * w;
      const uint64_t valid = nb >= 64 ? ~0ull : ((1ull << nb) - 1);
      const uint32_t slot = ((uint32_t)j << 8) | (uint32_t)w;
      const uint8_t f = flags[p];
      const uint64_t m0 = bernoulli_mask64(st0, slot, thr[2 * p], f & 1u, ~s & valid);
      const uint64_t m1 = s ? bernoulli_mask64(st1, slot, thr[2 * p + 1], (f >> 1) & 1u, s) : 0ull;
      out = s ^ (m0 | m1);
    }
    __syncwarp();
    if (t < ntask) dst[t] = out;
  }
  __syncwarp();
}

// batch_cross (evo.py:335-380): pairs in combinations(range(P),2) order, cut ~ U{1..H-1} from u32
// stream element k; child 2k = p1[:cut] | p2[cut:], child 2k+1 = p2[:cut] | p1[cut:].
__device__ inline void mutate_cross(const uint64_t *parents, int P, int W64, int H, const Stream &st,
                                    uint64_t *children, int lane) {
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
    const uint64_t pa = parents[a * W64 + w], pb = parents[b * W64 + w];
    children[(2 * k) * W64 + w] = (pa & L) | (pb & ~L);
    children[(2 * k + 1) * W64 + w] = (pb & L) | (pa & ~L);
  }
  __syncwarp();
}

struct EvoParams {
  int sel, mut, P, C, G;
  double p_bf;
  uint2 key;
  uint32_t step, n_offset;
};

__host__ inline int children_per_gen(int mut, int P, int C) { return mut >= TVO_MUT_CROSS ? P * (P - 1) : P * C; }

// One generation: select P parents from `cand`, mutate into `children` (Sgen rows).
template <typename T>
__device__ inline void evo_generation(const WarpBuf<T> &wb, const uint64_t *cand, const T *cand_lpj, int S_c,
                                      uint64_t *children, const EvoParams &ep, int gen, uint32_t n_glob, int W64,
                                      int H, double sparsity, double m_glob, double T_glob, int lane) {
  Stream st;
  st.key = ep.key;
  st.n = n_glob;
  st.tag = make_tag(ep.step, gen, PUR_SELECT);
  if (ep.sel == TVO_SEL_RANDPARENTS)
    select_randparents(cand, S_c, ep.P, W64, st, wb.iscr, wb.parents, lane);
  else
    select_fitparents<T>(cand, cand_lpj, S_c, ep.P, W64, st, ep.sel == TVO_SEL_FITPARENTS_FIXED, m_glob, T_glob,
                         wb.dscr, wb.iscr, wb.parents, lane);
  Stream s0 = st, s1 = st;
  s0.tag = make_tag(ep.step, gen, PUR_FLIP0);
  s1.tag = make_tag(ep.step, gen, PUR_FLIP1);
  switch (ep.mut) {
    case TVO_MUT_RANDFLIP:
      st.tag = make_tag(ep.step, gen, PUR_RANDFLIP);
      mutate_randflip(wb.parents, ep.P, ep.C, W64, H, st, wb.iscr, children, lane);
      break;
    case TVO_MUT_SPARSEFLIP:
      mutate_sparseflip(wb.parents, ep.P, ep.C, W64, H, sparsity, ep.p_bf, s0, s1, wb.thr, wb.flags, children, lane);
      break;
    default: {
      st.tag = make_tag(ep.step, gen, PUR_CROSS);
      mutate_cross(wb.parents, ep.P, W64, H, st, children, lane);
      const int nc = ep.P * (ep.P - 1);
      if (ep.mut == TVO_MUT_CROSS_RANDFLIP) {
        // randflip with C = 1 on every crossed child (evo.py:383-387): flip bit randint(H)
        st.tag = make_tag(ep.step, gen, PUR_RANDFLIP);
        for (int j = lane; j < nc; j += 32) {
          const int k = (int)randint_u32(st.u32((uint32_t)j, 0), (uint32_t)H);
          children[j * W64 + (k >> 6)] ^= 1ull << (k & 63);
        }
        __syncwarp();
      } else if (ep.mut == TVO_MUT_CROSS_SPARSEFLIP) {
        mutate_sparseflip(children, nc, 1, W64, H, sparsity, ep.p_bf, s0, s1, wb.thr, wb.flags, children, lane);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ dedup
// set_redundant_lpj_to_low (_set_redundant_lpj_to_low_CPU.pyx:13-30): child c is redundant if it
// equals a LATER child, else if it equals an old state.  Marks wb.lpj[S+c] = (float)-1e20.
template <typename T>
__device__ inline void mark_redundant(const WarpBuf<T> &wb, int S, int Snew, int W64, int lane) {
  const int M = S + Snew;
  for (int i = lane; i < M; i += 32)
    wb.hash[i] = hash_words(i < S ? wb.oldK + (size_t)i * W64 : wb.newK + (size_t)(i - S) * W64, W64);
  __syncwarp();
  for (int c = lane; c < Snew; c += 32) {
    const uint32_t hc = wb.hash[S + c];
    const uint64_t *sc = wb.newK + (size_t)c * W64;
    bool dup = false;
    for (int j = c + 1; j < Snew && !dup; ++j) {
      if (wb.hash[S + j] == hc) {
        const uint64_t *sj = wb.newK + (size_t)j * W64;
        bool eq = true;
        for (int w = 0; w < W64; ++w) eq &= (sj[w] == sc[w]);
        dup = eq;
      }
    }
    for (int j = 0; j < S && !dup; ++j) {
      if (wb.hash[j] == hc) {
        const uint64_t *sj = wb.oldK + (size_t)j * W64;
        bool eq = true;
        for (int w = 0; w < W64; ++w) eq &= (sj[w] == sc[w]);
        dup = eq;
      }
    }
    if (dup) wb.lpj[S + c] = LowLpj<T>::v();
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------ top-S merge
// update_states_for_batch (_utils.py:124-179): keep the S best of old+new, stored ascending (best
// last); ties -> lower concatenated index ranks higher.  Writes the datapoint's K and lpj rows in
// global memory and returns the number of selected children (warp-uniform).
template <typename T>
__device__ inline int select_topS(const WarpBuf<T> &wb, int S, int Snew, int W64, uint64_t *K_row, T *lpj_row,
                                  int lane) {
  const int M = S + Snew;
  for (int i = lane; i < M; i += 32) wb.keys[i] = sort_key((double)wb.lpj[i]);
  __syncwarp();
  int nsub = 0;
  for (int i0 = 0; i0 < M; i0 += 32) {
    const int i = i0 + lane;
    const uint64_t ki = i < M ? wb.keys[i] : 0;
    int rank = 0;
    for (int j = 0; j < M; ++j) {
      const uint64_t kj = wb.keys[j];
      rank += (kj > ki || (kj == ki && j < i)) ? 1 : 0;
    }
    const bool sel = i < M && rank < S;
    if (sel) {
      const int pos = S - 1 - rank;
      lpj_row[pos] = wb.lpj[i];
      wb.iscr[pos] = i;  // source row of output position (iscr has >= S entries: Scmax >= S)
      nsub += (i >= S) ? 1 : 0;
    }
  }
  __syncwarp();
  for (int t = lane; t < S * W64; t += 32) {
    const int src = wb.iscr[t / W64];
    const uint64_t *row = src < S ? wb.oldK + (size_t)src * W64 : wb.newK + (size_t)(src - S) * W64;
    K_row[t] = row[t % W64];
  }
  __syncwarp();
  return warp_sum(nsub);
}

}  // namespace tvo
