// tvo_b200 — BSC fast path: Gram-form fused EVO E-step and the M-step accumulation.
//
// E-step.  ||y - W s||^2 = ||y||^2 - 2 a.s + s^T G s  with  a = W^T y  (one DGEMM per chunk of
// datapoints, cuBLAS: a plain library GEMM) and G = W^T W (theta workspace).  For a sparse bit-packed
// state this is |s| reads of `a` from shared memory and |s|(|s|+1)/2 gathers from G — one LANE per
// state — instead of |s|*D column adds by a whole warp (bsc.cu, ~48% of that kernel's instructions).
// Datapoints whose y contains NaN take the direct form (nansum semantics, bsc.py:115).
//
// M-step (bsc.py:134-158).  One warp per datapoint: softmax weights on chip, E = <s>_q scattered into
// shared memory, Wq via one RED per (state, bit pair), sigma2 from lpj (or recomputed).  E is written
// per chunk so that  WpT += E^T Y  is one DGEMM (the union of active bits per datapoint is ~100 of 256,
// so the outer product is effectively dense).
#include <cublas_v2.h>

#include "bsc_device.cuh"
#include "bsc_estep_v2.cuh"
#include "bsc_gemm.cuh"
#include "bsc_gemm_tf32.cuh"

namespace tvo {

constexpr int kMaxSmemFast = 227 * 1024;
constexpr int kMaxWarpsLarge = 12;  // warps per CTA of the one-CTA-per-SM E-step variant (large H*S)
static unsigned g_v2_sync_mask = 0x41u;  // which phase barriers of the list-based kernel are on (tuning)
static int g_library_gemm = 0;  // tvo_bsc_estep_variant bit 16: cuBLAS instead of the hand-written GEMMs (tests)
static int g_disable_v2 = 0;         // tvo_bsc_estep_variant(1): generic kernel only (tests compare the two)

int evo_params_from_cfg(const tvo_evo_config *cfg, int S, int H, EvoParams &ep, EstepDims &dims, const char *who);
template <typename T>
int estep_evo_bsc_direct(const T *Y, int64_t ldY, const int64_t *idx, uint64_t *K, T *lpj, const void *ws,
                         const tvo_evo_config *cfg, const double *sparsity, int64_t B, int S, int H, int D,
                         int64_t *nsubs, void *stream, const int32_t *bsel = nullptr, const int32_t *bsel_count = nullptr);

// ------------------------------------------------------------------------------------ cuBLAS
static cublasHandle_t g_handles[64];
static int cublas_for_stream(cudaStream_t st, cublasHandle_t &h) {
  int dev = 0;
  TVO_CUDA(cudaGetDevice(&dev));
  TVO_REQUIRE(dev >= 0 && dev < 64, "cublas: device index %d out of range", dev);
  if (!g_handles[dev]) {
    cublasStatus_t s = cublasCreate(&g_handles[dev]);
    if (s != CUBLAS_STATUS_SUCCESS) return set_error(TVO_ECUDA, "cublasCreate failed (%d)", (int)s);
  }
  h = g_handles[dev];
  cublasStatus_t s = cublasSetStream(h, st);
  if (s != CUBLAS_STATUS_SUCCESS) return set_error(TVO_ECUDA, "cublasSetStream failed (%d)", (int)s);
  return TVO_OK;
}
#define TVO_CUBLAS(expr)                                                                     \
  do {                                                                                       \
    cublasStatus_t s_ = (expr);                                                              \
    if (s_ != CUBLAS_STATUS_SUCCESS) return set_error(TVO_ECUDA, "%s failed (%d)", #expr, (int)s_); \
    count_launch();                                                                          \
  } while (0)

// A (Bc x H, row-major) = Y (Bc x D, ld ldY) * W (D x H), W read from the padded W^T in the workspace
static int gemm_A(cublasHandle_t h, const double *Wt, int Dpad, const double *Y, int64_t ldY, double *A, int Bc, int H, int D) {
  const double one = 1.0, zero = 0.0;
  TVO_CUBLAS(cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, H, Bc, D, &one, Wt, Dpad, Y, (int)ldY, &zero, A, H));
  return TVO_OK;
}
static int gemm_A(cublasHandle_t h, const float *Wt, int Dpad, const float *Y, int64_t ldY, float *A, int Bc, int H, int D) {
  const float one = 1.f, zero = 0.f;
  TVO_CUBLAS(cublasSgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, H, Bc, D, &one, Wt, Dpad, Y, (int)ldY, &zero, A, H));
  return TVO_OK;
}

// Per-device staging rows of the large-H*S E-step variant (one S x W64 row per resident warp, ~20 MB): part of the
// per-device context, grown on demand and kept.  (A cudaMallocAsync / cudaFreeAsync pair per call hands the memory
// back to the driver at every synchronisation with the default release threshold — measured as sporadic host-side
// stalls of tens of ms per E-step.)  Calls on one device are serialised by the boundary's threading contract.
// slot 0: the E-step's staging rows; slot 1: the slab partials of the M-step GEMM (22 MB at the headline shape)
static void *g_stage[3][64];  // slot 2: the overflow list of the list-based E-step kernel
static size_t g_stage_bytes[3][64];
static int stage_buffer(size_t bytes, uint64_t **out, int slot = 0) {
  int dev = 0;
  TVO_CUDA(cudaGetDevice(&dev));
  TVO_REQUIRE(dev >= 0 && dev < 64, "stage buffer: device index %d out of range", dev);
  if (g_stage_bytes[slot][dev] < bytes) {
    if (g_stage[slot][dev]) TVO_CUDA(cudaFree(g_stage[slot][dev]));  // synchronises: no kernel still uses the old rows
    g_stage[slot][dev] = nullptr;
    g_stage_bytes[slot][dev] = 0;
    TVO_CUDA(cudaMalloc(&g_stage[slot][dev], bytes));
    g_stage_bytes[slot][dev] = bytes;
  }
  *out = (uint64_t *)g_stage[slot][dev];
  return TVO_OK;
}

// ------------------------------------------------------------------------------------ Gram lpj
// lpj of ONE state, evaluated by ONE lane.  a: shared memory (H), G: global (H x H), prior: global.
// Bit-walk form: used when the state has more active units than the lane's position list holds.
template <typename T>
__device__ __noinline__ T bsc_gram_lpj_walk(const uint64_t *st, int W64, int H, const T *a, const double *__restrict__ G,
                                            const T *__restrict__ prior, double yy, double coef) {
  double f = 0.0, pr = 0.0;
  for (int w1 = 0; w1 < W64; ++w1) {
    uint64_t m1 = st[w1];
    while (m1) {
      const int h1 = (w1 << 6) + __ffsll((long long)m1) - 1;
      m1 &= m1 - 1;
      const double *Grow = G + (size_t)h1 * H;
      f += __ldg(Grow + h1) - 2.0 * (double)__ldg(a + h1);
      pr += (double)__ldg(prior + h1);
      double g2 = 0.0;
      uint64_t m2 = m1;
      int w2 = w1;
      while (true) {
        while (m2) {
          const int h2 = (w2 << 6) + __ffsll((long long)m2) - 1;
          m2 &= m2 - 1;
          g2 += __ldg(Grow + h2);
        }
        if (++w2 >= W64) break;
        m2 = st[w2];
      }
      f += 2.0 * g2;
    }
  }
  return (T)((yy + f) * coef + pr);
}

// Active positions of one state (ascending) into the lane-strided list plist[j * 32]; returns |s| (entries beyond
// `cap` are counted, not stored).  The words are fetched four at a time — four loads in flight instead of a
// load -> test -> load chain, which matters when the state sits in global memory (old states of large H*S shapes,
// the M-step) — and each group is walked with ONE "next set bit of the group" loop, so a warp iterates
// max-over-lanes(bits in the group) times (H <= 256: one group).
// GROUPED = false: one "next set bit of the whole state" loop, words loaded as they are reached (states in shared
// memory; fewer instructions, iterations = max-over-lanes |s|).
template <bool GROUPED>
__device__ __forceinline__ int position_list(const uint64_t *st, int W64, uint16_t *plist, int cap) {
  int n = 0;
  if (!GROUPED) {
    int w = 0;
    uint64_t m = st[0];
    while (true) {
      while (m == 0 && ++w < W64) m = st[w];
      if (w >= W64) break;
      const uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
      const int pos = (w << 6) + (lo ? __ffs((int)lo) - 1 : 32 + __ffs((int)hi) - 1);
      m &= m - 1;
      if (n < cap) plist[n * 32] = (uint16_t)pos;
      ++n;
    }
    return n;
  }
  if (W64 <= 32) {
    // pass 1: which words are non-empty (four loads in flight); pass 2: ONE loop over the set bits of the state that
    // re-reads a word (now an L1 / shared-memory hit) only when the current one is exhausted — empty words cost
    // nothing and the warp iterates max-over-lanes |s| times
    uint32_t nz = 0;
    for (int w0 = 0; w0 < W64; w0 += 4) {
      const uint64_t d0 = st[w0], d1 = w0 + 1 < W64 ? st[w0 + 1] : 0ull, d2 = w0 + 2 < W64 ? st[w0 + 2] : 0ull,
                     d3 = w0 + 3 < W64 ? st[w0 + 3] : 0ull;
      nz |= ((d0 ? 1u : 0u) | (d1 ? 2u : 0u) | (d2 ? 4u : 0u) | (d3 ? 8u : 0u)) << w0;
    }
    int w = 0;
    uint64_t m = 0;
    while (true) {
      if (m == 0) {
        if (nz == 0) break;
        w = __ffs((int)nz) - 1;
        nz &= nz - 1;
        m = st[w];
      }
      const uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
      const int pos = (w << 6) + (lo ? __ffs((int)lo) - 1 : 32 + __ffs((int)hi) - 1);
      m &= m - 1;
      if (n < cap) plist[n * 32] = (uint16_t)pos;
      ++n;
    }
    return n;
  }
  for (int w0 = 0; w0 < W64; w0 += 4) {
    const uint64_t d0 = st[w0], d1 = w0 + 1 < W64 ? st[w0 + 1] : 0ull, d2 = w0 + 2 < W64 ? st[w0 + 2] : 0ull,
                   d3 = w0 + 3 < W64 ? st[w0 + 3] : 0ull;
    uint64_t m = d0;
    int u = 0;
    while (true) {
      while (m == 0 && ++u < 4) m = u == 1 ? d1 : (u == 2 ? d2 : d3);
      if (u >= 4) break;
      const uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
      const int pos = ((w0 + u) << 6) + (lo ? __ffs((int)lo) - 1 : 32 + __ffs((int)hi) - 1);
      m &= m - 1;
      if (n < cap) plist[n * 32] = (uint16_t)pos;
      ++n;
    }
  }
  return n;
}

// Same sums in the same order, but the active positions are first written to a lane-strided list in
// shared memory (plist[j * 32], the lane's base already applied), so the |s|(|s|-1)/2 pair loop is
// LDS + address + LDG + DADD instead of a 64-bit find-first-set / clear-lowest-bit walk per pair.
// LAT: latency-oriented variant for the shapes that run at a few warps per SM (grouped position list, four
// gathers in flight); the 32-warps-per-SM shapes are issue-bound and keep the shorter code.
template <typename T, bool LAT>
__device__ __forceinline__ T bsc_gram_lpj(const uint64_t *st, int W64, int H, const T *a, const double *__restrict__ G,
                                          const T *__restrict__ prior, double yy, double coef, uint16_t *plist, int cap) {
  // states are sparse (|s| ~ 1-5 of H): walk the set bits with one loop over "next set bit of the multi-word
  // mask" (empty words cost one test) instead of a bit loop per half word
  const int n = position_list<LAT>(st, W64, plist, cap);
  if (n > cap) return bsc_gram_lpj_walk<T>(st, W64, H, a, G, prior, yy, coef);
  double f = 0.0, pr = 0.0;
  if constexpr (LAT) {
    // latency-oriented order: all diagonal / a / prior terms first (four rows of loads in flight), then the
    // |s|(|s|-1)/2 off-diagonal gathers as ONE flattened pair list in batches of eight — about 1 + |s|^2/16 L2
    // round trips per state instead of one or two per active unit.  (A state is always scored by this same
    // function inside one kernel, so the different rounding order against the other variant is never observable.)
    int i = 0;
    for (; i + 4 <= n; i += 4) {
      const int h0 = plist[i * 32], h1 = plist[(i + 1) * 32], h2 = plist[(i + 2) * 32], h3 = plist[(i + 3) * 32];
      const double g0 = __ldg(G + (size_t)h0 * H + h0), g1 = __ldg(G + (size_t)h1 * H + h1),
                   g2 = __ldg(G + (size_t)h2 * H + h2), g3 = __ldg(G + (size_t)h3 * H + h3);
      const double a0 = (double)__ldg(a + h0), a1 = (double)__ldg(a + h1), a2 = (double)__ldg(a + h2),
                   a3 = (double)__ldg(a + h3);
      const double p0 = (double)__ldg(prior + h0), p1 = (double)__ldg(prior + h1), p2 = (double)__ldg(prior + h2),
                   p3 = (double)__ldg(prior + h3);
      f += g0 - 2.0 * a0;
      f += g1 - 2.0 * a1;
      f += g2 - 2.0 * a2;
      f += g3 - 2.0 * a3;
      pr += p0;
      pr += p1;
      pr += p2;
      pr += p3;
    }
    for (; i < n; ++i) {
      const int h0 = plist[i * 32];
      f += __ldg(G + (size_t)h0 * H + h0) - 2.0 * (double)__ldg(a + h0);
      pr += (double)__ldg(prior + h0);
    }
    double g = 0.0;
    int pi = 0, pj = 1;  // next pair (pi < pj), row-major over the strict upper triangle
    while (pi < n - 1) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        v[u] = 0.0;
        if (pi < n - 1) {
          v[u] = __ldg(G + (size_t)plist[pi * 32] * H + plist[pj * 32]);
          if (++pj == n) {
            ++pi;
            pj = pi + 1;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) g += v[u];
    }
    return (T)((yy + (f + 2.0 * g)) * coef + pr);
  } else {
    for (int i = 0; i < n; ++i) {
      const int h1 = plist[i * 32];
      const double *Grow = G + (size_t)h1 * H;
      f += __ldg(Grow + h1) - 2.0 * (double)__ldg(a + h1);
      pr += (double)__ldg(prior + h1);
      double g2 = 0.0;
      for (int j = i + 1; j < n; ++j) g2 += __ldg(Grow + plist[j * 32]);
      f += 2.0 * g2;
    }
    return (T)((yy + f) * coef + pr);
  }
}

// direct form for one state, any D, warp-cooperative, y from global memory (NaN rows only)
template <typename T>
__device__ inline T bsc_direct_lpj_slow(const uint64_t *st, int W64, const BscTheta<T> &th, const T *__restrict__ y,
                                        int D, int lane, T coef) {
  T ss = T(0), pr = T(0);
  for (int c = 0; c < th.Dpad; c += 32) {
    T acc = T(0);
    for (int w = 0; w < W64; ++w) {
      uint64_t word = st[w];
      while (word) {
        const int h = (w << 6) + __ffsll((long long)word) - 1;
        word &= word - 1;
        acc += __ldg(th.Wt + (size_t)h * th.Dpad + c + lane);
        if (c == 0) pr += __ldg(th.prior + h);
      }
    }
    const int d = c + lane;
    const T yv = d < D ? y[d] : T(0);
    const T t = acc - yv;
    if (yv == yv) ss += t * t;
  }
  return warp_sum(ss) * coef + pr;
}

template <typename T, bool LAT>
__device__ __forceinline__ void score_states(const uint64_t *states, int count, int W64, int H, int D, const BscTheta<T> &th,
                                             const T *a_sm, const T *__restrict__ y, double yy, bool has_nan, T *out,
                                             uint16_t *plist, int pcap, int lane, int rs) {
  if (!has_nan) {
    for (int s = lane; s < count; s += 32)
      out[s] = bsc_gram_lpj<T, LAT>(states + (size_t)s * rs, W64, H, a_sm, th.G, th.prior, yy, th.hdr->coef, plist + lane, pcap);
  } else {
    for (int s = 0; s < count; ++s) {
      const T v = bsc_direct_lpj_slow<T>(states + (size_t)s * rs, W64, th, y, D, lane, (T)th.hdr->coef);
      if (lane == 0) out[s] = v;
    }
  }
  __syncwarp();
}

// EVOVariationalStates.update (evo.py:80-115) for BSC, Gram form.  Same structure and same random
// stream as estep_evo_bsc_kernel (bsc.cu); `A` holds a = W^T y for the B datapoints of this launch.
// MINB = 4: the per-warp working set allows 4 CTAs (32 warps) per SM, registers are capped at 64 for it, the old
//           states are staged in shared memory;
// MINB = 1: large H*S shapes are shared-memory-limited to a few warps per SM: no register cap, and the old states
//           stay in global memory (dims.old_in_global; compile-time here so that the staged variant keeps plain
//           shared-memory loads for wb.oldK).
template <typename T, int MINB>
__global__ void __launch_bounds__(MINB == 1 ? 32 * kMaxWarpsLarge : 256, MINB)
estep_evo_bsc_gram_kernel(const T *__restrict__ Y, int64_t ldY, const T *__restrict__ A, const int64_t *__restrict__ idx,
                          int64_t row0, uint64_t *__restrict__ K, T *__restrict__ lpj, const void *ws, EvoParams ep,
                          const double *__restrict__ sparsity_p, EstepDims dims, int64_t B, int H, int D,
                          unsigned long long *nsubs, uint64_t *stage_rows, const int32_t *__restrict__ bsel,
                          const int32_t *__restrict__ bsel_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const BscTheta<T> th = bsc_theta_view<T>(ws, H, D);
  // bsel: process only the listed datapoints of this launch (the overflow list of the list-based kernel)
  if (bsel) B = *bsel_count;
  const size_t wbytes = warp_buf_bytes<T>(dims);
  unsigned char *base = smem_raw + (size_t)wib * wbytes;
  WarpBuf<T> wb = carve_warp_buf<T>(base, dims);
  // large H*S shapes: the old states stay in global memory (L1/L2), the merge goes through this warp's staging row
  constexpr bool kOldInGlobal = MINB == 1;
  uint64_t *stage = kOldInGlobal ? stage_rows + ((size_t)blockIdx.x * wpb + wib) * dims.S * dims.W64 : nullptr;
  const int S = dims.S, Snew = dims.Snew, Sgen = dims.Sgen, W64 = dims.W64;
  const double sparsity = sparsity_p ? *sparsity_p : 0.0;
  unsigned long long my_subs = 0;
  // position lists of the Gram scorer live in the keys/fmask region (free while states are scored)
  uint16_t *plist = (uint16_t *)wb.keys;
  const int pcap = H <= 65536 ? (int)(keys_region_bytes(dims) / 64) : 0;

  // All warps of a CTA walk the phases of the E-step in lockstep (PHASE_SYNC): the fused program is
  // ~160 KB of SASS (~50 KB executed for one configuration), far beyond the instruction cache, and 24
  // independent warps at 24 different places thrash it (ncu: 35% of stalls were instruction fetch; without
  // the barriers the kernel is 15% slower).  With the barriers the 8 warps of a CTA
  // share the cache lines of the phase they are all in.
#define PHASE_SYNC() __syncthreads()
  for (int64_t b0 = (int64_t)blockIdx.x * wpb; b0 < B; b0 += (int64_t)gridDim.x * wpb) {
    const bool active = b0 + wib < B;
    const int64_t b = active ? (bsel ? (int64_t)bsel[b0 + wib] : b0 + wib) : 0;
    const int64_t row = active ? (idx ? idx[b] : row0 + b) : 0;
    const uint32_t n_glob = ep.n_offset + (uint32_t)row;
    uint64_t *Krow = K + row * (int64_t)S * W64;
    T *lrow = lpj + row * (int64_t)S;
    const T *y = Y + b * ldY;
    double yy = 0.0;
    bool has_nan = false;
    if (active) {
      for (int d = lane; d < D; d += 32) {
        const double v = (double)y[d];
        has_nan |= (v != v);
        yy += v * v;
      }
      yy = warp_sum(yy);
      has_nan = __any_sync(FULL, has_nan);
      if (kOldInGlobal) {
        wb.oldK = Krow;
      } else {
        for (int t = lane; t < S * W64; t += 32) wb.oldK[t] = Krow[t];
      }
      __syncwarp();
    }
    PHASE_SYNC();
    const T *a_sm = A + b * (int64_t)H;  // a = W^T y row, read through L1/L2 (written by the DGEMM)
    if (active) {  // evo.py:95, unless the rows are current already (batch_fitparents)
      if (ep.lpj_current) {
        for (int i = lane; i < S; i += 32) wb.lpj[i] = lrow[i];
        __syncwarp();
      } else {
        score_states<T, kOldInGlobal>(wb.oldK, S, W64, H, D, th, a_sm, y, yy, has_nan, wb.lpj, plist, pcap, lane, W64);
      }
    }
    for (int g = 0; g < ep.G; ++g) {
      const uint64_t *cand = g == 0 ? wb.oldK : wb.newK + (size_t)(g - 1) * Sgen * wb.rs;
      const T *cand_lpj = g == 0 ? wb.lpj : wb.lpj + S + (g - 1) * Sgen;
      uint64_t *children = wb.newK + (size_t)g * Sgen * wb.rs;
      PHASE_SYNC();
      if (active)
        evo_generation<T>(wb, cand, cand_lpj, g == 0 ? S : Sgen, children, ep, g, n_glob, W64, H, sparsity,
                          ep.fit_norm ? ep.fit_norm[0] : 0.0, ep.fit_norm ? ep.fit_norm[1] : 1.0, lane, g == 0 ? W64 : wb.rs);
      PHASE_SYNC();
      if (active) score_states<T, kOldInGlobal>(children, Sgen, W64, H, D, th, a_sm, y, yy, has_nan, wb.lpj + S + g * Sgen, plist, pcap,
                                  lane, wb.rs);
    }
    PHASE_SYNC();
    if (active) mark_redundant<T>(wb, S, Snew, W64, lane);
    PHASE_SYNC();
    if (active) my_subs += (unsigned long long)select_topS<T, kOldInGlobal>(wb, S, Snew, W64, Krow, lrow, lane, stage);
  }
#undef PHASE_SYNC
  if (lane == 0 && my_subs) atomicAdd(nsubs, my_subs);
}

// ------------------------------------------------------------------------------------ M-step
constexpr int kMstepListCap = 16;  // positions per lane in the M-step's list (denser states walk their bits)
// stats layout: [WpT (H*D) | WqU (H*H, UPPER triangle incl. diagonal of my_Wq) | pies (H) | sigma2 | N | F]
template <typename T>
__global__ void __launch_bounds__(256)
bsc_mstep_kernel(const T *__restrict__ Y, int64_t ldY, const int64_t *__restrict__ idx, int64_t row0,
                 const uint64_t *__restrict__ K, int64_t k_stride_n, const T *__restrict__ lpj, int64_t lpj_stride_n,
                 const void *ws, double *__restrict__ stats, double *__restrict__ E_out, double *__restrict__ Yd_out,
                 int64_t B, int S, int H, int D, int exact_resid, size_t warp_bytes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const BscTheta<T> th = bsc_theta_view<T>(ws, H, D);
  const int W64 = n_words(H);
  unsigned char *p = smem_raw + (size_t)wib * warp_bytes;
  double *q = (double *)p;
  p += align16((size_t)S * 8);
  double *Esm = (double *)p;
  p += align16((size_t)H * 8);
  double *pies_w = (double *)p;  // per-warp running sum of E over this warp's datapoints
  p += align16((size_t)H * 8);
  uint16_t *plist = (uint16_t *)p;  // kMstepListCap x 32 lane-strided positions

  double *WqU = stats + (size_t)H * D, *pies = WqU + (size_t)H * H;
  double *sig = pies + H, *Ncnt = sig + 1, *Facc = Ncnt + 1;
  const double inv_coef = 1.0 / th.hdr->coef;  // -2 sigma2
  const int lcap = H <= 65536 ? kMstepListCap : 0;  // uint16 positions
  double acc_sig = 0.0, acc_F = 0.0, acc_N = 0.0;
  for (int h = lane; h < H; h += 32) pies_w[h] = 0.0;

  for (int64_t b = (int64_t)blockIdx.x * wpb + wib; b < B; b += (int64_t)gridDim.x * wpb) {
    const int64_t row = idx ? idx[b] : row0 + b;
    const uint64_t *Kn = K + row * k_stride_n;
    const T *l = lpj + row * lpj_stride_n;
    const T *y = Y + b * ldY;
    __syncwarp();
    for (int h = lane; h < H; h += 32) Esm[h] = 0.0;
    // softmax weights (lpj2pjc, _utils.py:182-191)
    double mx = -INFINITY;
    for (int s = lane; s < S; s += 32) mx = fmax(mx, (double)l[s]);
    mx = warp_max(mx);
    double se = 0.0;
    for (int s = lane; s < S; s += 32) {
      const double e = exp((double)l[s] - mx);
      q[s] = e;
      se += e;
    }
    se = warp_sum(se);
    int dn = 0;
    for (int d = lane; d < D; d += 32) {
      const T v = y[d];
      dn += (v == v) ? 1 : 0;
      if (Yd_out) Yd_out[b * (int64_t)D + d] = (double)v;
    }
    dn = warp_sum(dn);
    acc_F += mx + log(se) + th.hdr->prior_sum - 0.5 * (double)dn * th.hdr->log_2pi_sigma2;
    acc_N += 1.0;
    __syncwarp();
    // one lane per state: E_h += q_s, WqU[h,h'] += q_s (h <= h'), sigma2 term from lpj.  The active
    // positions go to a lane-strided list first (states are sparse; same walk as the E-step scorer).
    double part = 0.0;
    const double inv_se = 1.0 / se;
    for (int s = lane; s < S; s += 32) {
      const double qs = q[s] * inv_se;
      q[s] = qs;
      const uint64_t *st = Kn + (size_t)s * W64;  // lane s reads its own 8*W64 contiguous bytes: coalesced, no staging
      double pr = 0.0;
      const int n = position_list<true>(st, W64, plist + lane, lcap);
      if (n <= lcap) {
        for (int i = 0; i < n; ++i) {
          const int h1 = plist[i * 32 + lane];
          pr += (double)__ldg(th.prior + h1);
          atomicAdd(Esm + h1, qs);
          double *Wrow = WqU + (size_t)h1 * H;
          for (int j = i; j < n; ++j) atomicAdd(Wrow + plist[j * 32 + lane], qs);  // j == i: the diagonal
        }
      } else {  // dense state: walk the bits in place
        for (int w1 = 0; w1 < W64; ++w1) {
          uint64_t m1 = st[w1];
          while (m1) {
            const int h1 = (w1 << 6) + __ffsll((long long)m1) - 1;
            pr += (double)__ldg(th.prior + h1);
            atomicAdd(Esm + h1, qs);
            double *Wrow = WqU + (size_t)h1 * H;
            uint64_t m2 = m1;  // includes h1 itself: the diagonal
            int w2 = w1;
            while (true) {
              while (m2) {
                const int h2 = (w2 << 6) + __ffsll((long long)m2) - 1;
                m2 &= m2 - 1;
                atomicAdd(Wrow + h2, qs);
              }
              if (++w2 >= W64) break;
              m2 = st[w2];
            }
            m1 &= m1 - 1;
          }
        }
      }
      if (!exact_resid) part += qs * (((double)l[s] - pr) * inv_coef);
    }
    if (!exact_resid) acc_sig += warp_sum(part);
    __syncwarp();
    for (int h = lane; h < H; h += 32) {
      const double e = Esm[h];
      E_out[b * (int64_t)H + h] = e;
      pies_w[h] += e;
    }
    if (exact_resid) {  // <||y - W s||^2>_q recomputed from Y and theta, plain sum (bsc.py:148-150)
      double part2 = 0.0;
      for (int s = 0; s < S; ++s) {
        T ss = T(0);
        for (int c = 0; c < th.Dpad; c += 32) {
          T a = T(0);
          for (int w = 0; w < W64; ++w) {
            uint64_t word = Kn[(size_t)s * W64 + w];
            while (word) {
              const int h = (w << 6) + __ffsll((long long)word) - 1;
              word &= word - 1;
              a += __ldg(th.Wt + (size_t)h * th.Dpad + c + lane);
            }
          }
          const int d = c + lane;
          const T t = (d < D ? y[d] : T(0)) - a;
          ss += t * t;
        }
        part2 += q[s] * (double)warp_sum(ss);
      }
      acc_sig += part2;
    }
  }
  __syncwarp();
  for (int h = lane; h < H; h += 32) {
    const double v = pies_w[h];
    if (v != 0.0) atomicAdd(pies + h, v);
  }
  if (lane == 0 && acc_N != 0.0) {
    atomicAdd(sig, acc_sig);
    atomicAdd(Ncnt, acc_N);
    atomicAdd(Facc, acc_F);
  }
}

// ------------------------------------------------------------------------------------ M-step, list walk
// BSC.update_param_batch (bsc.py:134-158) for H <= 256 when lpj was just produced by the E-step (residual from lpj).
// One warp per datapoint, one LANE per state: the state's 32-byte bit row is read with 128-bit loads (the next
// datapoint's rows and lpj are already in flight: the global K walk was what the previous kernel waited on), its
// active units are walked once (non-empty half words first, as in the list-based E-step), and every unit adds the
// state's posterior weight to
//   * <s>_h in shared memory as a 64-bit FIXED-POINT integer (2^-56 units; q <= 1, S <= 128 states): shared fp64
//     atomicAdd is a compare-and-swap spin loop on sm_100a, the integer add is one native ATOMS.ADD.64 and is
//     order-independent (the sum is deterministic); the truncation is < S 2^-56 absolute per unit;
//   * my_Wq's upper triangle in global memory, one fp64 RED per unit pair of the state.
constexpr double kEFix = 72057594037927936.0;  // 2^56

template <typename T, int EPL>  // EPL: states per lane = ceil(S / 32) (1, 2 or 4)
__global__ void __launch_bounds__(256, EPL <= 2 ? 3 : 2)
bsc_mstep_list_kernel(const T *__restrict__ Y, int64_t ldY, const int64_t *__restrict__ idx, int64_t row0,
                      const uint64_t *__restrict__ K, int64_t k_stride_n, const T *__restrict__ lpj, int64_t lpj_stride_n,
                      const void *ws, double *__restrict__ stats, double *__restrict__ E_out, double *__restrict__ Yd_out,
                      int64_t B, int S, int H, int D, int /*exact_resid: never*/, size_t warp_bytes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const BscTheta<T> th = bsc_theta_view<T>(ws, H, D);
  const int W64 = n_words(H), HW = 2 * W64;
  unsigned char *p = smem_raw + (size_t)wib * warp_bytes;
  unsigned long long *Efx = (unsigned long long *)p;  // H fixed-point <s>_h of the current datapoint
  double *pies_w = (double *)(p + (size_t)H * 8);     // H: running sum of <s> over this warp's datapoints
  unsigned char *plist = p + (size_t)2 * H * 8;       // 32 lanes x 16 positions of the state being walked
  double *WqU = stats + (size_t)H * D, *pies = WqU + (size_t)H * H;
  double *sig = pies + H, *Ncnt = sig + 1, *Facc = Ncnt + 1;
  const double inv_coef = 1.0 / th.hdr->coef;  // -2 sigma2
  double acc_sig = 0.0, acc_F = 0.0, acc_N = 0.0;
  for (int h = lane; h < H; h += 32) {
    pies_w[h] = 0.0;
    Efx[h] = 0ull;
  }
  unsigned char *pl = plist + lane * 16;

  // software pipeline: registers of the NEXT datapoint (bit rows of this lane's states, their lpj, this lane's y
  // values) and the row index of the one after it — every global read is issued a full iteration before its use
  uint32_t nx_w[EPL][8];
  T nx_l[EPL];
  auto fetch = [&](int64_t bb, int64_t row) {
    const uint64_t *Kn = K + row * k_stride_n;
    const T *l = lpj + row * lpj_stride_n;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const int s = lane + 32 * e;
#pragma unroll
      for (int k = 0; k < 8; ++k) nx_w[e][k] = 0u;
      nx_l[e] = T(0);
      if (s < S) {
        const uint32_t *st = (const uint32_t *)(Kn + (size_t)s * W64);
        if ((W64 & 1) == 0) {
#pragma unroll
          for (int q4 = 0; q4 < 2; ++q4) {
            if (q4 < W64 / 2) {
              const uint4 v = *reinterpret_cast<const uint4 *>(st + 4 * q4);
              nx_w[e][4 * q4 + 0] = v.x;
              nx_w[e][4 * q4 + 1] = v.y;
              nx_w[e][4 * q4 + 2] = v.z;
              nx_w[e][4 * q4 + 3] = v.w;
            }
          }
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < HW) nx_w[e][k] = st[k];
        }
        nx_l[e] = l[s];
      }
    }
    // y is only counted for NaNs here (and widened for the fp32 mode's GEMM): an L2 prefetch instead of registers
    const T *yn = Y + bb * ldY;
    for (int d = lane * (int)(128 / sizeof(T)); d < D; d += 32 * (int)(128 / sizeof(T)))
      asm volatile("prefetch.global.L2 [%0];" ::"l"(yn + d));
  };
  const int64_t stride = (int64_t)gridDim.x * wpb;
  int64_t b = (int64_t)blockIdx.x * wpb + wib;
  int64_t row_nx = 0;  // row index of datapoint b + stride
  if (b < B) {
    fetch(b, idx ? idx[b] : row0 + b);
    if (b + stride < B) row_nx = idx ? idx[b + stride] : row0 + b + stride;
  }
  uint32_t *wl = (uint32_t *)(plist + 512) + lane * 8;  // this lane's 8 half words of the state being walked
  for (; b < B; b += stride) {
    uint32_t w[EPL][8];
    double lv[EPL];
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
#pragma unroll
      for (int k = 0; k < 8; ++k) w[e][k] = nx_w[e][k];
      lv[e] = (double)nx_l[e];
    }
    if (b + stride < B) {
      const int64_t row = row_nx;
      if (b + 2 * stride < B) row_nx = idx ? idx[b + 2 * stride] : row0 + b + 2 * stride;
      fetch(b + stride, row);
    }
    const T *y = Y + b * ldY;
    // softmax weights (lpj2pjc, _utils.py:182-191)
    double mx = -INFINITY;
#pragma unroll
    for (int e = 0; e < EPL; ++e)
      if (lane + 32 * e < S) mx = fmax(mx, lv[e]);
    mx = warp_max(mx);
    double ex[EPL], se = 0.0;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      ex[e] = lane + 32 * e < S ? exp(lv[e] - mx) : 0.0;
      se += ex[e];
    }
    se = warp_sum(se);
    int dn = 0;
    for (int d = lane; d < D; d += 32) {
      const T v = y[d];
      dn += (v == v) ? 1 : 0;
      if (Yd_out) Yd_out[b * (int64_t)D + d] = (double)v;
    }
    const double inv_se = 1.0 / se;
    double part = 0.0;
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      if (lane + 32 * e < S) {
        const double qs = ex[e] * inv_se;
        const unsigned long long qfx = (unsigned long long)(qs * kEFix);
        uint32_t nz = 0;
        int ntot = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          nz |= (w[e][k] ? 1u : 0u) << k;
          ntot += __popc(w[e][k]);
        }
        *reinterpret_cast<uint4 *>(wl) = make_uint4(w[e][0], w[e][1], w[e][2], w[e][3]);  // refills index the row
        *reinterpret_cast<uint4 *>(wl + 4) = make_uint4(w[e][4], w[e][5], w[e][6], w[e][7]);
        double pr = 0.0;
        uint32_t m = 0;
        int hw = 0;
#pragma unroll 1
        for (int n = 0; n < ntot; ++n) {
          if (m == 0) {
            hw = __ffs((int)nz) - 1;
            nz &= nz - 1;
            m = wl[hw];
          }
          const int h = (hw << 5) + __ffs((int)m) - 1;
          m &= m - 1;
          pr += (double)__ldg(th.prior + h);
          atomicAdd(Efx + h, qfx);
          // (the diagonal of my_Wq is <s_h s_h> = <s_h>: it is added once per warp from pies_w at the end instead of
          //  one RED per state and unit — all datapoints hit the same few diagonal entries, and same-address REDs
          //  serialise in L2)
          if (ntot <= 16) {
#pragma unroll 1
            for (int i = 0; i < n; ++i) atomicAdd(WqU + (size_t)pl[i] * H + h, qs);  // earlier units are smaller: upper triangle
            pl[n] = (unsigned char)h;
          } else {  // dense state: the earlier units straight from the bit row
            for (int k2 = 0; k2 <= hw; ++k2) {
              uint32_t m2 = wl[k2];
              if (k2 == hw) m2 &= (1u << (h & 31)) - 1u;
              while (m2) {
                const int h2 = (k2 << 5) + __ffs((int)m2) - 1;
                m2 &= m2 - 1;
                atomicAdd(WqU + (size_t)h2 * H + h, qs);
              }
            }
          }
        }
        part += qs * ((lv[e] - pr) * inv_coef);
      }
    }
    acc_sig += warp_sum(part);
    dn = warp_sum(dn);
    acc_F += mx + log(se) + th.hdr->prior_sum - 0.5 * (double)dn * th.hdr->log_2pi_sigma2;
    acc_N += 1.0;
    __syncwarp();
    for (int h = lane; h < H; h += 32) {
      const double ev = (double)Efx[h] * (1.0 / kEFix);
      Efx[h] = 0ull;
      E_out[b * (int64_t)H + h] = ev;
      pies_w[h] += ev;
    }
    __syncwarp();
  }
  for (int h = lane; h < H; h += 32) {
    const double v = pies_w[h];
    if (v != 0.0) {
      atomicAdd(pies + h, v);
      atomicAdd(WqU + (size_t)h * H + h, v);  // diag(my_Wq) = sum_n <s_h>_n
    }
  }
  if (lane == 0 && acc_N != 0.0) {
    atomicAdd(sig, acc_sig);
    atomicAdd(Ncnt, acc_N);
    atomicAdd(Facc, acc_F);
  }
}

template <typename T>
static size_t bsc_mstep_warp_bytes(int S, int H) {
  return align16((size_t)S * 8) + 2 * align16((size_t)H * 8) + (size_t)kMstepListCap * 32 * 2;
}

// ------------------------------------------------------------------------------------ host side
template <typename T>
static int64_t estep_scratch_per_dp(int H) {
  return (int64_t)H * sizeof(T);
}
template <typename T>
static int64_t mstep_scratch_per_dp(int H, int D) {
  return (int64_t)H * 8 + (sizeof(T) == 8 ? 0 : (int64_t)D * 8);
}

template <typename T>
int estep_evo_bsc(const T *Y, int64_t ldY, const int64_t *idx, uint64_t *K, T *lpj, const void *ws,
                  const tvo_evo_config *cfg, const double *sparsity, void *scratch, int64_t scratch_bytes, int64_t B,
                  int S, int H, int D, int64_t *nsubs, void *stream) {
  TVO_REQUIRE(Y && K && lpj && ws && cfg && nsubs, "estep_evo_bsc: null pointer");
  const int64_t per = estep_scratch_per_dp<T>(H);
  // scratch layout: [sparseflip table (flip_table_bytes(H), only when it fits)] [overflow counter (16 B) + overflow
  // list of the list-based kernel (4 B per datapoint of a chunk)] [||y||^2 (8 B per datapoint)] [a = W^T y rows]
  const bool sparse = cfg->mutation == TVO_MUT_SPARSEFLIP || cfg->mutation == TVO_MUT_CROSS_SPARSEFLIP;
  const int64_t tab_bytes = sparse && scratch && sparsity ? (int64_t)flip_table_bytes(H) : 0;
  const bool use_tab = tab_bytes > 0 && scratch_bytes >= tab_bytes + 48 + (per + 12) * min((int64_t)1024, B);
  int64_t cap = !scratch ? 0 : use_tab ? (scratch_bytes - tab_bytes - 48) / (per + 12) : scratch_bytes / per;
  if (cap > B) cap = B;
  const int64_t ovf_bytes = use_tab ? (int64_t)align16((size_t)(16 + 4 * cap)) : 0;
  const int64_t yy_bytes = use_tab ? (int64_t)align16((size_t)(8 * cap)) : 0;
  int32_t *ovf_count = use_tab ? (int32_t *)((unsigned char *)scratch + tab_bytes) : nullptr;
  int32_t *ovf_list = use_tab ? ovf_count + 4 : nullptr;
  double *yy_buf = use_tab ? (double *)((unsigned char *)scratch + tab_bytes + ovf_bytes) : nullptr;
  unsigned char *scratch_a = (unsigned char *)scratch + (use_tab ? tab_bytes + ovf_bytes + yy_bytes : 0);
  if (cap < 1024 && cap < B)  // no (useful) scratch: direct column-walk form
    return estep_evo_bsc_direct<T>(Y, ldY, idx, K, lpj, ws, cfg, sparsity, B, S, H, D, nsubs, stream);
  EvoParams ep;
  EstepDims dims;
  int rc = evo_params_from_cfg(cfg, S, H, ep, dims, "estep_evo_bsc");
  if (rc) return rc;
  TVO_REQUIRE_FITPARENTS(ep, "estep_evo_bsc");
  TVO_REQUIRE(!(ep.mut == TVO_MUT_SPARSEFLIP || ep.mut == TVO_MUT_CROSS_SPARSEFLIP) || sparsity,
              "estep_evo_bsc: sparseflip needs the device scalar `sparsity`");
  TVO_REQUIRE(B >= 0 && D > 0 && ldY >= D && ldY < (1ll << 31), "estep_evo_bsc: bad sizes");
  if (B == 0) return TVO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  cublasHandle_t hb;
  rc = cublas_for_stream(st, hb);
  if (rc) return rc;
  const BscTheta<T> th = bsc_theta_view<T>(ws, H, D);
  size_t wbytes = warp_buf_bytes<T>(dims);
  // 8 warps per CTA, 4 CTAs per SM at the headline shape (6.9 kB per warp).  Working sets that do not allow that
  // (large H*S) leave the old states in global memory instead of staging them: more resident warps.
  if (wbytes * 8 * 4 > (size_t)kMaxSmemFast) {
    dims.old_in_global = 1;
    dims.RS = dims.W64 | 1;  // conflict-free per-lane row access (see EstepDims::RS); at the headline shape the
                             // extra 1.1 kB per warp would cost the fourth CTA per SM (measured: 6 % slower)
    wbytes = warp_buf_bytes<T>(dims);
  }
  auto kern = dims.old_in_global ? estep_evo_bsc_gram_kernel<T, 1> : estep_evo_bsc_gram_kernel<T, 4>;
  // the one-CTA-per-SM variant takes as many warps as its shared memory holds (up to 12: it is latency-bound and
  // every resident warp counts; 10 at H=1024 S=128)
  int wpb = dims.old_in_global ? (int)min((size_t)kMaxWarpsLarge, (size_t)kMaxSmemFast / wbytes) : 8;
  if (wpb < 1) wpb = 1;
  while (wpb > 1 && wbytes * wpb > (size_t)kMaxSmemFast) wpb >>= 1;
  TVO_REQUIRE(wbytes * wpb <= (size_t)kMaxSmemFast, "estep_evo_bsc: per-datapoint working set %zu B exceeds shared memory",
              wbytes);
  const size_t smem = wbytes * wpb;
  TVO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  TVO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpb * 32, smem));
  occ = max(occ, 1);
  uint64_t *stage_rows = nullptr;
  if (dims.old_in_global) {
    rc = stage_buffer((size_t)sm_count() * occ * wpb * S * dims.W64 * 8, &stage_rows);
    if (rc) return rc;
  }
  T *A = (T *)scratch_a;
  if (use_tab) {
    FlipEntry *tab = (FlipEntry *)scratch;
    flip_table_kernel<<<(H + 1 + 127) / 128, 128, 0, st>>>(tab, sparsity, H, ep.p_bf);
    TVO_AFTER_LAUNCH("flip_table");
    ep.flip_tab = tab;
  }
  // list-based kernel (bsc_estep_v2.cuh) for sparse states at 64 <= H <= 1024, randparents + sparseflip; the generic
  // kernel then takes the datapoints it handed back (overflow list) — or everything when the shape is not covered
  V2Dims v2{S, dims.Snew, dims.Sgen, ep.P, ep.C, ep.G, dims.W64, H, 0, 0, 0, 0, 0, 0};
  v2_finish(v2);
  const size_t v2_smem = (size_t)v2.warp_bytes * 8;
  const bool use_v2 = use_tab && v2_supported(ep, v2) && v2_smem <= (size_t)kMaxSmemFast && !g_disable_v2;
  const int eS = S <= 32 ? 1 : S <= 64 ? 2 : 4, eG = dims.Sgen <= 32 ? 1 : dims.Sgen <= 64 ? 2 : 4,
            eN = dims.Snew <= 32 ? 1 : dims.Snew <= 64 ? 2 : 4;
  // H > 256 (two-byte positions, d_h gathered): one catch-all shape class
  auto kern2 = H > 256                         ? estep_evo_bsc_v2_kernel<T, 4, 4, 4, unsigned short>
               : eS == 1 && eG == 1 && eN == 1 ? estep_evo_bsc_v2_kernel<T, 1, 1, 1>
               : eS <= 2 && eG == 1 && eN <= 2 ? estep_evo_bsc_v2_kernel<T, 2, 1, 2>
                                               : estep_evo_bsc_v2_kernel<T, 4, 4, 4>;
  int occ2 = 1;
  if (use_v2) {
    TVO_CUDA(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v2_smem));
    TVO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, kern2, 256, v2_smem));
    occ2 = max(occ2, 1);
  }
  // a = Y W: the hand-written TMA + FP64 tensor-core kernel (bsc_gemm.cuh; also ||y||^2, so that the list-based
  // kernel does not read Y), the library GEMM for shapes it does not cover and for the fp32 mode
  const bool own_gemm = yy_buf && !g_library_gemm &&
                        (sizeof(T) == 8 ? gemm_yw_supported(H, D, ldY, Y) : gemm_yw_tf32_supported(H, D, ldY, Y));
  // Overflow datapoints of the list-based kernel.  D <= 256: ONE list for the whole call (a per-device buffer of B
  // entries) and one launch of the direct-form kernel at the end — it needs no a = Y W rows, so nothing of a chunk has
  // to survive (a launch per chunk cost ~60 us each for the few datapoints it finds: 4 % of the E-step).  Larger D:
  // the Gram-form kernel per chunk.
  const bool ovf_deferred = use_v2 && D <= 256 && B < (1ll << 31);
  if (ovf_deferred) {
    uint64_t *buf = nullptr;
    rc = stage_buffer((size_t)16 + (size_t)B * 4, &buf, 2);
    if (rc) return rc;
    ovf_count = (int32_t *)buf;
    ovf_list = ovf_count + 4;
    TVO_CUDA(cudaMemsetAsync(ovf_count, 0, 16, st));
  }
  for (int64_t c0 = 0; c0 < B; c0 += cap) {
    const int Bc = (int)min(cap, B - c0);
    if (own_gemm) {
      void *tokg;
      kernel_timing_begin(TVO_TIMED_BSC_GEMM_YW, st, &tokg);
      cudaError_t ge = sizeof(T) == 8
                           ? launch_gemm_yw((const double *)(Y + c0 * ldY), ldY, th.Wg, (double *)A, yy_buf, Bc, H, D, sm_count(), st)
                           : launch_gemm_yw_tf32((const float *)(Y + c0 * ldY), ldY, (const unsigned char *)th.Wg, (float *)A,
                                                 yy_buf, Bc, H, D, sm_count(), st);
      kernel_timing_end(tokg, st);
      if (ge != cudaSuccess) return set_error(TVO_ECUDA, "launch bsc_gemm_yw: %s", cudaGetErrorString(ge));
      count_launch();
    } else {
      rc = gemm_A(hb, th.Wt, th.Dpad, Y + c0 * ldY, ldY, A, Bc, H, D);
      if (rc) return rc;
    }
    const int grid = (int)max((int64_t)1, min((int64_t)(Bc + wpb - 1) / wpb, (int64_t)sm_count() * occ));
    void *tok;
    if (use_v2) {
      if (!ovf_deferred) TVO_CUDA(cudaMemsetAsync(ovf_count, 0, 16, st));
      const int grid2 = (int)max((int64_t)1, min((int64_t)(Bc + 7) / 8, (int64_t)sm_count() * occ2));
      kernel_timing_begin(TVO_TIMED_ESTEP_EVO_BSC, st, &tok);
      kern2<<<grid2, 256, v2_smem, st>>>(Y + c0 * ldY, ldY, A, idx ? idx + c0 : nullptr, c0, K, lpj, ws, ep, v2, Bc, D,
                                         (unsigned long long *)nsubs, ovf_count, ovf_list, g_v2_sync_mask,
                                         own_gemm ? yy_buf : nullptr, ovf_deferred ? c0 : 0);
      kernel_timing_end(tok, st);
      TVO_AFTER_LAUNCH("estep_evo_bsc_v2");
      if (ovf_deferred) continue;
      // overflow datapoints: the generic kernel reads their count on the device (an empty list costs one idle launch)
      kern<<<min(grid, sm_count()), wpb * 32, smem, st>>>(Y + c0 * ldY, ldY, A, idx ? idx + c0 : nullptr, c0, K, lpj, ws, ep,
                                                          sparsity, dims, Bc, H, D, (unsigned long long *)nsubs, stage_rows,
                                                          ovf_list, ovf_count);
      TVO_AFTER_LAUNCH("estep_evo_bsc_gram(overflow)");
      continue;
    }
    kernel_timing_begin(TVO_TIMED_ESTEP_EVO_BSC, st, &tok);
    kern<<<grid, wpb * 32, smem, st>>>(Y + c0 * ldY, ldY, A, idx ? idx + c0 : nullptr, c0, K, lpj, ws, ep, sparsity, dims,
                                       Bc, H, D, (unsigned long long *)nsubs, stage_rows, nullptr, nullptr);
    kernel_timing_end(tok, st);
    TVO_AFTER_LAUNCH("estep_evo_bsc_gram");
  }
  if (ovf_deferred)
    return estep_evo_bsc_direct<T>(Y, ldY, idx, K, lpj, ws, cfg, sparsity, B, S, H, D, nsubs, stream, ovf_list, ovf_count);
  return TVO_OK;
}

template <typename T>
int bsc_mstep(const T *Y, int64_t ldY, const int64_t *idx, const uint64_t *K, int64_t k_stride_n, const T *lpj,
              int64_t lpj_stride_n, const void *ws, double *stats, void *scratch, int64_t scratch_bytes, int64_t B, int S,
              int H, int D, int exact, void *stream) {
  TVO_REQUIRE(Y && K && lpj && ws && stats && scratch, "bsc_mstep: null pointer");
  TVO_REQUIRE(B >= 0 && S > 0 && H > 0 && D > 0 && ldY >= D && ldY < (1ll << 31), "bsc_mstep: bad sizes");
  const int64_t per = mstep_scratch_per_dp<T>(H, D);
  const int64_t cap = scratch_bytes / per;
  TVO_REQUIRE(cap >= 1, "bsc_mstep: scratch of %lld B holds no datapoint (need %lld B each)", (long long)scratch_bytes,
              (long long)per);
  if (B == 0) return TVO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  cublasHandle_t hb;
  int rc = cublas_for_stream(st, hb);
  if (rc) return rc;
  // list-walk kernel: lpj fresh from the E-step (residual from lpj), H <= 256, S <= 128
  const bool use_list = !exact && H <= 256 && S <= 128 && !g_disable_v2;
  size_t wbytes = use_list ? (size_t)2 * H * 8 + 512 + 1024 : bsc_mstep_warp_bytes<T>(S, H);
  int wpb = 8;
  while (wpb > 1 && wbytes * wpb > (size_t)kMaxSmemFast) wpb >>= 1;
  TVO_REQUIRE(wbytes * wpb <= (size_t)kMaxSmemFast, "bsc_mstep: working set %zu B exceeds shared memory", wbytes);
  const size_t smem = wbytes * wpb;
  auto kern = !use_list ? bsc_mstep_kernel<T>
              : S <= 32 ? bsc_mstep_list_kernel<T, 1>
              : S <= 64 ? bsc_mstep_list_kernel<T, 2>
                        : bsc_mstep_list_kernel<T, 4>;
  TVO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  TVO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpb * 32, smem));
  occ = max(occ, 1);
  const double one = 1.0;
  for (int64_t c0 = 0; c0 < B; c0 += cap) {
    const int Bc = (int)min(cap, B - c0);
    double *E = (double *)scratch;
    double *Yd = sizeof(T) == 8 ? nullptr : E + (size_t)Bc * H;
    const int grid = (int)max((int64_t)1, min((int64_t)(Bc + wpb - 1) / wpb, (int64_t)sm_count() * occ));
    void *tok;
    kernel_timing_begin(TVO_TIMED_BSC_MSTEP, st, &tok);
    kern<<<grid, wpb * 32, smem, st>>>(Y + c0 * ldY, ldY, idx ? idx + c0 : nullptr, c0, K, k_stride_n, lpj, lpj_stride_n, ws,
                                       stats, E, Yd, Bc, S, H, D, exact, wbytes);
    kernel_timing_end(tok, st);
    TVO_AFTER_LAUNCH("bsc_mstep");
    // WpT (H x D, row-major) += E^T (H x Bc) * Y (Bc x D)     (bsc.py:145,153): the hand-written TMA + FP64
    // tensor-core kernel (bsc_gemm.cuh, deterministic slab reduction), the library GEMM for shapes it does not cover
    const double *Yp = sizeof(T) == 8 ? (const double *)(Y + c0 * ldY) : Yd;
    const int ldy = sizeof(T) == 8 ? (int)ldY : D;
    if (!g_library_gemm && gemm_ety_supported(H, D, ldy, Yp, E)) {
      uint64_t *part = nullptr;
      rc = stage_buffer(gemm_ety_part_bytes(H, D, sm_count()), &part, 1);
      if (rc) return rc;
      void *tokg;
      kernel_timing_begin(TVO_TIMED_BSC_GEMM_ETY, st, &tokg);
      cudaError_t ge = launch_gemm_ety(E, Yp, ldy, (double *)part, stats, D, Bc, H, D, sm_count(), st);
      kernel_timing_end(tokg, st);
      if (ge != cudaSuccess) return set_error(TVO_ECUDA, "launch bsc_gemm_ety: %s", cudaGetErrorString(ge));
      count_launch(2);
    } else {
      TVO_CUBLAS(cublasDgemm(hb, CUBLAS_OP_N, CUBLAS_OP_T, D, H, Bc, &one, Yp, ldy, E, H, &one, stats, D));
    }
  }
  return TVO_OK;
}

}  // namespace tvo

using namespace tvo;

extern "C" {
int64_t tvo_bsc_stats_len(int32_t H, int32_t D) { return (int64_t)H * D + (int64_t)H * H + H + 3; }
int64_t tvo_bsc_estep_scratch_bytes(int64_t B, int32_t H, int32_t is_f64) {
  return B * ((int64_t)H * (is_f64 ? 8 : 4) + 12) + 48 + (int64_t)flip_table_bytes(H);
}
int tvo_bsc_estep_variant(int32_t generic_only) {
  g_disable_v2 = (generic_only & 1) ? 1 : 0;
  g_library_gemm = (generic_only & 0x10000) ? 1 : 0;
  if (generic_only & 0xFFFE) g_v2_sync_mask = ((unsigned)generic_only & 0xFFFFu) >> 1;  // tuning: phase-barrier mask
  return TVO_OK;
}
int64_t tvo_bsc_mstep_scratch_bytes(int64_t B, int32_t H, int32_t D, int32_t is_f64) {
  return B * ((int64_t)H * 8 + (is_f64 ? 0 : (int64_t)D * 8));
}
int tvo_estep_evo_bsc_f64(const double *Y, int64_t ldY, const int64_t *idx, uint64_t *K, double *lpj, const void *ws,
                          const tvo_evo_config *cfg, const double *sparsity, void *scratch, int64_t scratch_bytes,
                          int64_t B, int32_t S, int32_t H, int32_t D, int64_t *nsubs, void *stream) {
  return estep_evo_bsc<double>(Y, ldY, idx, K, lpj, ws, cfg, sparsity, scratch, scratch_bytes, B, S, H, D, nsubs, stream);
}
int tvo_estep_evo_bsc_f32(const float *Y, int64_t ldY, const int64_t *idx, uint64_t *K, float *lpj, const void *ws,
                          const tvo_evo_config *cfg, const double *sparsity, void *scratch, int64_t scratch_bytes,
                          int64_t B, int32_t S, int32_t H, int32_t D, int64_t *nsubs, void *stream) {
  return estep_evo_bsc<float>(Y, ldY, idx, K, lpj, ws, cfg, sparsity, scratch, scratch_bytes, B, S, H, D, nsubs, stream);
}
int tvo_bsc_mstep_f64(const double *Y, int64_t ldY, const int64_t *idx, const uint64_t *K, int64_t k_stride_n,
                      const double *lpj, int64_t lpj_stride_n, const void *ws, double *stats, void *scratch,
                      int64_t scratch_bytes, int64_t B, int32_t S, int32_t H, int32_t D, int32_t exact_residual,
                      void *stream) {
  return bsc_mstep<double>(Y, ldY, idx, K, k_stride_n, lpj, lpj_stride_n, ws, stats, scratch, scratch_bytes, B, S, H, D,
                           exact_residual, stream);
}
int tvo_bsc_mstep_f32(const float *Y, int64_t ldY, const int64_t *idx, const uint64_t *K, int64_t k_stride_n,
                      const float *lpj, int64_t lpj_stride_n, const void *ws, double *stats, void *scratch,
                      int64_t scratch_bytes, int64_t B, int32_t S, int32_t H, int32_t D, int32_t exact_residual,
                      void *stream) {
  return bsc_mstep<float>(Y, ldY, idx, K, k_stride_n, lpj, lpj_stride_n, ws, stats, scratch, scratch_bytes, B, S, H, D,
                          exact_residual, stream);
}
}
