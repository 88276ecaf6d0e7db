// tvo_b200 — model-independent kernels: state packing, candidate generation, dedup, top-S merge,
// posterior weights.  These are the stand-alone C-ABI entry points; the fused E-step kernels call
// the same device functions (estep_device.cuh).
#include "estep_device.cuh"

namespace tvo {

constexpr int kMaxSmemStates = 227 * 1024;

// ------------------------------------------------------------------------------------ pack / unpack
// uint8 (n_states, H) -> uint64 (n_states, W64).  One warp per 64-bit word: lane l reads bytes
// h = 64w + l and 64w + 32 + l (two coalesced 32 B segments) and the word is built by two ballots.
__global__ void __launch_bounds__(256) pack_states_kernel(const uint8_t *__restrict__ in, uint64_t *__restrict__ out,
                                                          int64_t n_states, int H) {
  const int lane = threadIdx.x & 31;
  const int W64 = n_words(H);
  const int64_t nwords = n_states * W64;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t wi = warp; wi < nwords; wi += nwarps) {
    const int64_t s = wi / W64;
    const int w = (int)(wi % W64);
    const int h0 = 64 * w + lane, h1 = h0 + 32;
    const uint8_t *row = in + s * H;
    const uint32_t lo = __ballot_sync(FULL, h0 < H && row[h0] != 0);
    const uint32_t hi = __ballot_sync(FULL, h1 < H && row[h1] != 0);
    if (lane == 0) out[wi] = (uint64_t)hi << 32 | lo;
  }
}

__global__ void __launch_bounds__(256) unpack_states_kernel(const uint64_t *__restrict__ in, uint8_t *__restrict__ out,
                                                            int64_t n_states, int H) {
  const int W64 = n_words(H);
  const int64_t total = n_states * H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / H;
    const int h = (int)(i % H);
    out[i] = (uint8_t)((in[s * W64 + (h >> 6)] >> (h & 63)) & 1ull);
  }
}

// fullem.py:14-28: row i = binary representation of i, most significant bit first -> s_h = bit (H-1-h) of i
__global__ void state_matrix_kernel(uint64_t *out, int H) {
  const int W64 = n_words(H);
  const int64_t n = 1ll << H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    for (int w = 0; w < W64; ++w) {
      uint64_t word = 0;
      for (int b = 0; b < 64; ++b) {
        const int h = 64 * w + b;
        if (h < H && ((i >> (H - 1 - h)) & 1)) word |= 1ull << b;
      }
      out[i * W64 + w] = word;
    }
  }
}

// ------------------------------------------------------------------------------------ config
int evo_params_from_cfg(const tvo_evo_config *cfg, int S, int H, EvoParams &ep, EstepDims &dims, const char *who) {
  TVO_REQUIRE(cfg->parent_selection >= 0 && cfg->parent_selection <= 2, "%s: unknown parent_selection %d", who,
              cfg->parent_selection);
  TVO_REQUIRE(cfg->mutation >= 0 && cfg->mutation <= 4, "%s: unknown mutation %d", who, cfg->mutation);
  ep.sel = cfg->parent_selection;
  ep.mut = cfg->mutation;
  ep.P = cfg->n_parents;
  ep.C = ep.mut >= TVO_MUT_CROSS ? ep.P - 1 : cfg->n_children;
  ep.G = cfg->n_generations;
  ep.p_bf = cfg->p_bf;
  ep.key = make_uint2((uint32_t)(cfg->seed & 0xffffffffu), (uint32_t)(cfg->seed >> 32));
  ep.step = cfg->step;
  ep.n_offset = cfg->n_offset;
  ep.flip_tab = nullptr;
  ep.fit_norm = cfg->fit_norm;
  ep.lpj_current = (cfg->flags & TVO_EVO_LPJ_CURRENT) ? 1 : 0;
  TVO_REQUIRE(S > 0 && H > 0, "%s: S=%d H=%d must be positive", who, S, H);
  TVO_REQUIRE(ep.P >= 1 && ep.G >= 1 && ep.G < 64, "%s: need n_parents >= 1 and 1 <= n_generations < 64", who);
  TVO_REQUIRE(ep.mut >= TVO_MUT_CROSS ? ep.P >= 2 : ep.C >= 1, "%s: need n_children >= 1 (n_parents >= 2 for cross)", who);
  TVO_REQUIRE(S >= ep.P, "%s: S=%d < n_parents=%d (evo.py:143 pre-condition K >= n_parents)", who, S, ep.P);
  TVO_REQUIRE(ep.mut >= TVO_MUT_CROSS || H >= ep.C, "%s: H=%d < n_children=%d (evo.py:143)", who, H, ep.C);
  TVO_REQUIRE(ep.mut < TVO_MUT_CROSS || H >= 2, "%s: crossover needs H >= 2", who);
  const int Sgen = children_per_gen(ep.mut, ep.P, ep.C);
  TVO_REQUIRE(ep.G == 1 || Sgen >= ep.P, "%s: children per generation %d < n_parents %d", who, Sgen, ep.P);
  TVO_REQUIRE(Sgen < (1 << 24) && n_words(H) <= 256, "%s: S_gen / H too large for the stream slot encoding", who);
  dims.S = S;
  dims.Sgen = Sgen;
  dims.Snew = Sgen * ep.G;
  dims.P = ep.P;
  dims.W64 = n_words(H);
  dims.Dpad = 0;
  dims.Scmax = S > Sgen ? S : Sgen;
  dims.need_fit = ep.sel != TVO_SEL_RANDPARENTS;
  dims.need_fmask = ep.mut == TVO_MUT_CROSS_SPARSEFLIP;
  dims.old_in_global = 0;
  dims.RS = 0;
  return TVO_OK;
}

// ------------------------------------------------------------------------------------ fit_norm
// {m, T} of evo.py:404-406 over the whole (B, S_c) block: m = float32(min(min lpj, 0)),
// T = sum(lpj - 2m).  Deterministic two-pass reduction (fixed partition, fixed order).
template <typename T>
__global__ void __launch_bounds__(256) fit_norm_partial_kernel(const T *__restrict__ lpj, int64_t stride_n,
                                                               const int64_t *__restrict__ idx, int64_t B, int S_c,
                                                               double *part /* 2*gridDim */) {
  __shared__ double smn[256], ssum[256];
  double mn = 0.0, sum = 0.0;
  const int64_t total = B * S_c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / S_c;
    const int s = (int)(i % S_c);
    const double v = (double)lpj[(idx ? idx[b] : b) * stride_n + s];
    mn = fmin(mn, v);
    sum += v;
  }
  smn[threadIdx.x] = mn;
  ssum[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      smn[threadIdx.x] = fmin(smn[threadIdx.x], smn[threadIdx.x + o]);
      ssum[threadIdx.x] += ssum[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part[2 * blockIdx.x] = smn[0];
    part[2 * blockIdx.x + 1] = ssum[0];
  }
}
__global__ void fit_norm_final_kernel(const double *part, int nblocks, int64_t count, double *out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double mn = 0.0, sum = 0.0;
    for (int i = 0; i < nblocks; ++i) {
      mn = fmin(mn, part[2 * i]);
      sum += part[2 * i + 1];
    }
    const double m = (double)__double2float_rn(mn);
    out[0] = m;
    out[1] = sum - 2.0 * m * (double)count;
  }
}

// ------------------------------------------------------------------------------------ generation kernel
template <typename T>
__global__ void __launch_bounds__(256)
evo_generation_kernel(const uint64_t *__restrict__ cand, int64_t cand_stride_n, const T *__restrict__ cand_lpj,
                      int64_t cand_lpj_stride_n, const int64_t *__restrict__ idx, int cand_indexed, int S_c,
                      uint64_t *__restrict__ children, EvoParams ep, int gen, const double *__restrict__ sparsity_p,
                      const double *__restrict__ fit_norm, EstepDims dims, int64_t B, int H) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const size_t wbytes = warp_buf_bytes<T>(dims);
  const WarpBuf<T> wb = carve_warp_buf<T>(smem_raw + (size_t)wib * wbytes, dims);
  const int W64 = dims.W64, Sgen = dims.Sgen;
  const double sparsity = sparsity_p ? *sparsity_p : 0.0;
  const double m_glob = fit_norm ? fit_norm[0] : 0.0, T_glob = fit_norm ? fit_norm[1] : 1.0;
  for (int64_t b = (int64_t)blockIdx.x * wpb + wib; b < B; b += (int64_t)gridDim.x * wpb) {
    const int64_t row = idx ? idx[b] : b;
    const int64_t crow = cand_indexed ? row : b;
    const uint32_t n_glob = ep.n_offset + (uint32_t)row;
    __syncwarp();
    // candidates staged in oldK (sized S >= S_c or Scmax), their lpj in wb.lpj
    for (int t = lane; t < S_c * W64; t += 32) wb.oldK[t] = cand[crow * cand_stride_n + t];
    if (cand_lpj)
      for (int s = lane; s < S_c; s += 32) wb.lpj[s] = cand_lpj[crow * cand_lpj_stride_n + s];
    __syncwarp();
    evo_generation<T>(wb, wb.oldK, wb.lpj, S_c, wb.newK, ep, gen, n_glob, W64, H, sparsity, m_glob, T_glob, lane);
    uint64_t *out = children + b * (int64_t)Sgen * W64;
    for (int t = lane; t < Sgen * W64; t += 32) out[t] = wb.newK[t];
  }
}

// RandomSampledVarStates.py:57-59: every bit ~ Bernoulli(sparsity).  One thread per (b, j, w) word.
__global__ void __launch_bounds__(256) random_states_kernel(uint64_t *__restrict__ out, const int64_t *__restrict__ idx,
                                                            int64_t B, int S_new, int H, double sparsity, uint2 key,
                                                            uint32_t step, uint32_t n_offset) {
  const int W64 = n_words(H);
  const int64_t total = B * S_new * W64;
  uint64_t T;
  uint32_t always;
  prob_to_fixed64(sparsity, T, always);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W64);
    const int j = (int)((i / W64) % S_new);
    const int64_t b = i / ((int64_t)W64 * S_new);
    Stream st;
    st.key = key;
    st.n = n_offset + (uint32_t)(idx ? idx[b] : b);
    st.tag = make_tag(step, 0, PUR_SAMPLE);
    const int nb = H - 64 * w;
    const uint64_t valid = nb >= 64 ? ~0ull : ((1ull << nb) - 1);
    out[i] = bernoulli_mask64(st, ((uint32_t)j << 8) | (uint32_t)w, T, always, valid);
  }
}

// ------------------------------------------------------------------------------------ dedup / merge kernels
template <typename T>
__global__ void __launch_bounds__(256)
mark_redundant_kernel(const uint64_t *__restrict__ new_states, T *__restrict__ new_lpj, const uint64_t *__restrict__ K,
                      int64_t k_stride_n, const int64_t *__restrict__ idx, EstepDims dims, int64_t B) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const size_t wbytes = warp_buf_bytes<T>(dims);
  const WarpBuf<T> wb = carve_warp_buf<T>(smem_raw + (size_t)wib * wbytes, dims);
  const int S = dims.S, Snew = dims.Snew, W64 = dims.W64;
  for (int64_t b = (int64_t)blockIdx.x * wpb + wib; b < B; b += (int64_t)gridDim.x * wpb) {
    const int64_t row = idx ? idx[b] : b;
    __syncwarp();
    for (int t = lane; t < S * W64; t += 32) wb.oldK[t] = K[row * k_stride_n + t];
    for (int t = lane; t < Snew * W64; t += 32) wb.newK[t] = new_states[b * (int64_t)Snew * W64 + t];
    for (int s = lane; s < Snew; s += 32) wb.lpj[S + s] = new_lpj[b * (int64_t)Snew + s];
    __syncwarp();
    mark_redundant<T>(wb, S, Snew, W64, lane);
    for (int s = lane; s < Snew; s += 32) new_lpj[b * (int64_t)Snew + s] = wb.lpj[S + s];
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
select_topS_kernel(const uint64_t *__restrict__ new_states, const T *__restrict__ new_lpj, uint64_t *__restrict__ K,
                   T *__restrict__ lpj, const int64_t *__restrict__ idx, EstepDims dims, int64_t B,
                   unsigned long long *nsubs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const size_t wbytes = warp_buf_bytes<T>(dims);
  const WarpBuf<T> wb = carve_warp_buf<T>(smem_raw + (size_t)wib * wbytes, dims);
  const int S = dims.S, Snew = dims.Snew, W64 = dims.W64;
  unsigned long long my_subs = 0;
  for (int64_t b = (int64_t)blockIdx.x * wpb + wib; b < B; b += (int64_t)gridDim.x * wpb) {
    const int64_t row = idx ? idx[b] : b;
    uint64_t *Krow = K + row * (int64_t)S * W64;
    T *lrow = lpj + row * (int64_t)S;
    __syncwarp();
    for (int t = lane; t < S * W64; t += 32) wb.oldK[t] = Krow[t];
    for (int t = lane; t < Snew * W64; t += 32) wb.newK[t] = new_states[b * (int64_t)Snew * W64 + t];
    for (int s = lane; s < S; s += 32) wb.lpj[s] = lrow[s];
    for (int s = lane; s < Snew; s += 32) wb.lpj[S + s] = new_lpj[b * (int64_t)Snew + s];
    __syncwarp();
    my_subs += (unsigned long long)select_topS<T>(wb, S, Snew, W64, Krow, lrow, lane);
  }
  if (lane == 0 && my_subs) atomicAdd(nsubs, my_subs);
}

// ------------------------------------------------------------------------------------ posterior weights
// lpj2pjc (_utils.py:182-191): one warp per row.
template <typename T>
__global__ void __launch_bounds__(256) lpj2pjc_kernel(const T *__restrict__ lpj, T *__restrict__ out, int64_t N, int S) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t n = warp; n < N; n += nwarps) {
    const T *l = lpj + n * S;
    T mx = -INFINITY;
    for (int s = lane; s < S; s += 32) mx = l[s] > mx ? l[s] : mx;
    mx = warp_max(mx);
    T se = T(0);
    for (int s = lane; s < S; s += 32) se += exp(l[s] - mx);
    se = warp_sum(se);
    for (int s = lane; s < S; s += 32) out[n * S + s] = exp(l[s] - mx) / se;
  }
}

// mean_posterior (_utils.py:194-230): out[n,g] = sum_s q_ns g[n,s,g].  One block per datapoint,
// q in shared memory, threads over g (coalesced reads of g[n,s,:]).
template <typename T>
__global__ void __launch_bounds__(256)
mean_posterior_kernel(const T *__restrict__ g, const T *__restrict__ lpj, T *__restrict__ out, int64_t N, int S,
                      int64_t G) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *q = (T *)smem_raw;
  __shared__ T red[256];
  for (int64_t n = blockIdx.x; n < N; n += gridDim.x) {
    const T *l = lpj + n * S;
    T mx = -INFINITY;
    for (int s = threadIdx.x; s < S; s += blockDim.x) mx = l[s] > mx ? l[s] : mx;
    red[threadIdx.x] = mx;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] = red[threadIdx.x + o] > red[threadIdx.x] ? red[threadIdx.x + o] : red[threadIdx.x];
      __syncthreads();
    }
    mx = red[0];
    __syncthreads();
    T se = T(0);
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
      const T e = exp(l[s] - mx);
      q[s] = e;
      se += e;
    }
    red[threadIdx.x] = se;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    se = red[0];
    __syncthreads();
    for (int64_t j = threadIdx.x; j < G; j += blockDim.x) {
      T acc = T(0);
      const T *gp = g + n * S * G + j;
      for (int s = 0; s < S; ++s) acc += (q[s] / se) * gp[(int64_t)s * G];
      out[n * G + j] = acc;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------ host wrappers

// ------------------------------------------------------------------------------------ TVS (tvs.py:47-81)
// Approximate marginals p(s_h = 1 | y_n) = sum_s q_ns s_nsh (mean_posterior of the states themselves,
// tvs.py:66-70; also W <s>_q of BSC.data_estimator, bsc.py:241-252) straight from the packed K: one warp
// per datapoint, softmax weights in shared memory, lanes over h.
template <typename T>
__global__ void __launch_bounds__(256)
state_marginals_kernel(const uint64_t *__restrict__ K, int64_t k_stride_n, const T *__restrict__ lpj,
                       int64_t lpj_stride_n, const int64_t *__restrict__ idx, T *__restrict__ out, int64_t B, int S,
                       int H) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int W64 = n_words(H);
  double *q = (double *)smem_raw + (size_t)wib * S;
  for (int64_t b = (int64_t)blockIdx.x * wpb + wib; b < B; b += (int64_t)gridDim.x * wpb) {
    const int64_t row = idx ? idx[b] : b;
    const T *l = lpj + row * lpj_stride_n;
    const uint64_t *Kn = K + row * k_stride_n;
    __syncwarp();
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
    __syncwarp();
    for (int h = lane; h < H; h += 32) {
      double acc = 0.0;
      const int w = h >> 6, sh = h & 63;
      for (int s = 0; s < S; ++s)
        if ((Kn[(size_t)s * W64 + w] >> sh) & 1ull) acc += q[s];
      out[b * (int64_t)H + h] = (T)(acc / se);
    }
  }
}

// Bit h of new state (b, j) = [u < p_h], u = (r >> 8) * 2^-24 with r = u32 stream element h of sub-stream
// (n, slot j, tag (step, group, PUR_SAMPLE)) — a 24-bit uniform like torch.rand's float32 draw
// (tvs.py:60-74).  probs: (B, H) with row stride probs_stride_n, or one broadcast row (stride 0).
// One thread per (b, j, word).
template <typename T>
__global__ void __launch_bounds__(256)
sample_states_probs_kernel(uint64_t *__restrict__ out, int64_t out_stride_n, int S_off, const T *__restrict__ probs,
                           int64_t probs_stride_n, const int64_t *__restrict__ idx, int64_t B, int S_new, int H, uint2 key,
                           uint32_t step, uint32_t group, uint32_t n_offset) {
  const int W64 = n_words(H);
  const int64_t total = B * S_new * W64;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W64);
    const int j = (int)((i / W64) % S_new);
    const int64_t b = i / ((int64_t)W64 * S_new);
    Stream st;
    st.key = key;
    st.n = n_offset + (uint32_t)(idx ? idx[b] : b);
    st.tag = make_tag(step, group, PUR_SAMPLE);
    const T *p = probs + b * probs_stride_n;
    uint64_t word = 0;
    const int nb = min(64, H - 64 * w);
    for (int c = 0; c < nb; c += 4) {
      const uint4 r = st.block((uint32_t)j, (uint32_t)((64 * w + c) >> 2));
      const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int h = 64 * w + c + t;
        if (c + t < nb) {
          const double u = (double)(rr[t] >> 8) * 0x1.0p-24;
          if (u < (double)p[h]) word |= 1ull << (c + t);
        }
      }
    }
    out[b * out_stride_n + (int64_t)(S_off + j) * W64 + w] = word;
  }
}

static int grid_for(int64_t items, int per_block, int blocks_per_sm) {
  const int64_t want = (items + per_block - 1) / per_block;
  return (int)max((int64_t)1, min(want, (int64_t)sm_count() * blocks_per_sm));
}

template <typename T, typename Kern>
static int warp_kernel_geometry(Kern kern, const EstepDims &dims, int64_t B, const char *who, int &grid, int &wpb,
                                size_t &smem) {
  const size_t wbytes = warp_buf_bytes<T>(dims);
  wpb = 8;
  while (wpb > 1 && wbytes * wpb > (size_t)kMaxSmemStates) wpb >>= 1;
  TVO_REQUIRE(wbytes * wpb <= (size_t)kMaxSmemStates, "%s: per-datapoint working set %zu B exceeds shared memory", who,
              wbytes);
  smem = wbytes * wpb;
  TVO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  TVO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpb * 32, smem));
  grid = (int)max((int64_t)1, min((B + wpb - 1) / wpb, (int64_t)sm_count() * max(occ, 1)));
  return TVO_OK;
}

template <typename T>
int evo_generation_host(const uint64_t *cand, int64_t cand_stride_n, const T *cand_lpj, int64_t cand_lpj_stride_n,
                        const int64_t *idx, int cand_indexed, int S_c, uint64_t *children, const tvo_evo_config *cfg,
                        int gen, const double *sparsity, const double *fit_norm, int64_t B, int H, void *stream) {
  TVO_REQUIRE(cand && children && cfg, "evo_generation: null pointer");
  EvoParams ep;
  EstepDims dims;
  int rc = evo_params_from_cfg(cfg, S_c, H, ep, dims, "evo_generation");
  if (rc) return rc;
  TVO_REQUIRE(gen >= 0 && gen < ep.G, "evo_generation: gen=%d out of range", gen);
  TVO_REQUIRE(ep.sel == TVO_SEL_RANDPARENTS || cand_lpj, "evo_generation: fitparents needs cand_lpj");
  TVO_REQUIRE(ep.sel != TVO_SEL_FITPARENTS || fit_norm, "evo_generation: batch-global fitparents needs fit_norm");
  TVO_REQUIRE(!(ep.mut == TVO_MUT_SPARSEFLIP || ep.mut == TVO_MUT_CROSS_SPARSEFLIP) || sparsity,
              "evo_generation: sparseflip needs the device scalar `sparsity`");
  if (B == 0) return TVO_OK;
  // the stand-alone kernel stages S_c candidates in oldK and one generation in newK
  dims.S = S_c;
  dims.Snew = dims.Sgen;
  dims.Scmax = S_c > dims.Sgen ? S_c : dims.Sgen;
  int grid, wpb;
  size_t smem;
  auto kern = evo_generation_kernel<T>;
  rc = warp_kernel_geometry<T>(kern, dims, B, "evo_generation", grid, wpb, smem);
  if (rc) return rc;
  kern<<<grid, wpb * 32, smem, (cudaStream_t)stream>>>(cand, cand_stride_n, cand_lpj, cand_lpj_stride_n, idx, cand_indexed,
                                                       S_c, children, ep, gen, sparsity, fit_norm, dims, B, H);
  TVO_AFTER_LAUNCH("evo_generation");
  return TVO_OK;
}

template <typename T>
int fit_norm_host(const T *lpj, int64_t stride_n, const int64_t *idx, int64_t B, int S_c, double *out, void *stream) {
  TVO_REQUIRE(lpj && out, "fit_norm: null pointer");
  TVO_REQUIRE(B > 0 && S_c > 0, "fit_norm: empty block");
  static double *part = nullptr;  // 2 * 1024 doubles of scratch per process/device
  static int part_dev = -1;
  int dev = 0;
  TVO_CUDA(cudaGetDevice(&dev));
  if (!part || part_dev != dev) {
    TVO_CUDA(cudaMalloc(&part, 2 * 1024 * sizeof(double)));
    part_dev = dev;
  }
  const int grid = (int)min((int64_t)1024, (B * S_c + 255) / 256);
  fit_norm_partial_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(lpj, stride_n, idx, B, S_c, part);
  TVO_AFTER_LAUNCH("fit_norm_partial");
  fit_norm_final_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(part, grid, B * S_c, out);
  TVO_AFTER_LAUNCH("fit_norm_final");
  return TVO_OK;
}

template <typename T>
int mark_redundant_host(const uint64_t *new_states, T *new_lpj, const uint64_t *K, int64_t k_stride_n,
                        const int64_t *idx, int64_t B, int S_new, int S, int H, void *stream) {
  TVO_REQUIRE(new_states && new_lpj && K, "mark_redundant: null pointer");
  TVO_REQUIRE(B >= 0 && S_new > 0 && S > 0 && H > 0, "mark_redundant: bad sizes");
  if (B == 0) return TVO_OK;
  EstepDims dims{S, S_new, 1, 1, n_words(H), 0, S, 0};
  int grid, wpb;
  size_t smem;
  auto kern = mark_redundant_kernel<T>;
  int rc = warp_kernel_geometry<T>(kern, dims, B, "mark_redundant", grid, wpb, smem);
  if (rc) return rc;
  kern<<<grid, wpb * 32, smem, (cudaStream_t)stream>>>(new_states, new_lpj, K, k_stride_n, idx, dims, B);
  TVO_AFTER_LAUNCH("mark_redundant");
  return TVO_OK;
}

template <typename T>
int select_topS_host(const uint64_t *new_states, const T *new_lpj, uint64_t *K, T *lpj, const int64_t *idx, int64_t B,
                     int S_new, int S, int H, int64_t *nsubs, void *stream) {
  TVO_REQUIRE(new_states && new_lpj && K && lpj && nsubs, "select_topS: null pointer");
  TVO_REQUIRE(B >= 0 && S_new > 0 && S > 0 && H > 0, "select_topS: bad sizes");
  if (B == 0) return TVO_OK;
  EstepDims dims{S, S_new, 1, 1, n_words(H), 0, S, 0};
  int grid, wpb;
  size_t smem;
  auto kern = select_topS_kernel<T>;
  int rc = warp_kernel_geometry<T>(kern, dims, B, "select_topS", grid, wpb, smem);
  if (rc) return rc;
  kern<<<grid, wpb * 32, smem, (cudaStream_t)stream>>>(new_states, new_lpj, K, lpj, idx, dims, B,
                                                       (unsigned long long *)nsubs);
  TVO_AFTER_LAUNCH("select_topS");
  return TVO_OK;
}

template <typename T>
int lpj2pjc_host(const T *lpj, T *pjc, int64_t N, int S, void *stream) {
  TVO_REQUIRE(lpj && pjc, "lpj2pjc: null pointer");
  TVO_REQUIRE(N >= 0 && S > 0, "lpj2pjc: bad sizes");
  if (N == 0) return TVO_OK;
  lpj2pjc_kernel<T><<<grid_for(N, 8, 8), 256, 0, (cudaStream_t)stream>>>(lpj, pjc, N, S);
  TVO_AFTER_LAUNCH("lpj2pjc");
  return TVO_OK;
}

template <typename T>
int mean_posterior_host(const T *g, const T *lpj, T *out, int64_t N, int S, int64_t G, void *stream) {
  TVO_REQUIRE(g && lpj && out, "mean_posterior: null pointer");
  TVO_REQUIRE(N >= 0 && S > 0 && G > 0, "mean_posterior: bad sizes");
  TVO_REQUIRE((size_t)S * sizeof(T) <= 200 * 1024, "mean_posterior: S=%d too large", S);
  if (N == 0) return TVO_OK;
  const size_t smem = (size_t)S * sizeof(T);
  auto kern = mean_posterior_kernel<T>;
  if (smem > 40 * 1024) TVO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid_for(N, 1, 8), 256, smem, (cudaStream_t)stream>>>(g, lpj, out, N, S, G);
  TVO_AFTER_LAUNCH("mean_posterior");
  return TVO_OK;
}


template <typename T>
int state_marginals_host(const uint64_t *K, int64_t k_stride_n, const T *lpj, int64_t lpj_stride_n, const int64_t *idx,
                         T *out, int64_t B, int S, int H, void *stream) {
  TVO_REQUIRE(K && lpj && out, "state_marginals: null pointer");
  TVO_REQUIRE(B >= 0 && S > 0 && H > 0, "state_marginals: bad sizes");
  if (B == 0) return TVO_OK;
  const size_t smem = (size_t)8 * S * 8;
  TVO_REQUIRE(smem <= 200 * 1024, "state_marginals: S=%d too large", S);
  auto kern = state_marginals_kernel<T>;
  if (smem > 40 * 1024) TVO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid_for(B, 8, 8), 256, smem, (cudaStream_t)stream>>>(K, k_stride_n, lpj, lpj_stride_n, idx, out, B, S, H);
  TVO_AFTER_LAUNCH("state_marginals");
  return TVO_OK;
}

template <typename T>
int sample_states_probs_host(uint64_t *out, int64_t out_stride_n, int S_off, const T *probs, int64_t probs_stride_n,
                             const int64_t *idx, int64_t B, int S_new, int H, uint64_t seed, uint32_t step, uint32_t group,
                             uint32_t n_offset, void *stream) {
  TVO_REQUIRE(out && probs, "sample_states_probs: null pointer");
  TVO_REQUIRE(B >= 0 && S_new >= 0 && S_off >= 0 && H > 0 && S_new < (1 << 24) && group < 64, "sample_states_probs: bad sizes");
  if (B == 0 || S_new == 0) return TVO_OK;
  const uint2 key = make_uint2((uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
  sample_states_probs_kernel<T><<<grid_for(B * S_new * n_words(H), 256, 8), 256, 0, (cudaStream_t)stream>>>(
      out, out_stride_n, S_off, probs, probs_stride_n, idx, B, S_new, H, key, step, group, n_offset);
  TVO_AFTER_LAUNCH("sample_states_probs");
  return TVO_OK;
}

}  // namespace tvo

using namespace tvo;

extern "C" {
int tvo_pack_states(const uint8_t *K_u8, uint64_t *K_packed, int64_t n_states, int32_t H, void *stream) {
  TVO_REQUIRE(K_u8 && K_packed, "pack_states: null pointer");
  TVO_REQUIRE(n_states >= 0 && H > 0, "pack_states: bad sizes");
  if (n_states == 0) return TVO_OK;
  pack_states_kernel<<<grid_for(n_states * n_words(H), 8, 16), 256, 0, (cudaStream_t)stream>>>(K_u8, K_packed, n_states, H);
  TVO_AFTER_LAUNCH("pack_states");
  return TVO_OK;
}
int tvo_unpack_states(const uint64_t *K_packed, uint8_t *K_u8, int64_t n_states, int32_t H, void *stream) {
  TVO_REQUIRE(K_u8 && K_packed, "unpack_states: null pointer");
  TVO_REQUIRE(n_states >= 0 && H > 0, "unpack_states: bad sizes");
  if (n_states == 0) return TVO_OK;
  unpack_states_kernel<<<grid_for(n_states * H, 256, 16), 256, 0, (cudaStream_t)stream>>>(K_packed, K_u8, n_states, H);
  TVO_AFTER_LAUNCH("unpack_states");
  return TVO_OK;
}
int tvo_state_matrix(uint64_t *K_packed, int32_t H, void *stream) {
  TVO_REQUIRE(K_packed, "state_matrix: null pointer");
  TVO_REQUIRE(H > 0 && H <= 30, "state_matrix: H=%d out of range (1..30)", H);
  state_matrix_kernel<<<grid_for(1ll << H, 256, 8), 256, 0, (cudaStream_t)stream>>>(K_packed, H);
  TVO_AFTER_LAUNCH("state_matrix");
  return TVO_OK;
}
int tvo_evo_generation_f64(const uint64_t *cand, int64_t cand_stride_n, const double *cand_lpj, int64_t cand_lpj_stride_n,
                           const int64_t *idx, int32_t cand_indexed, int32_t S_c, uint64_t *children,
                           const tvo_evo_config *cfg, int32_t gen, const double *sparsity, const double *fit_norm,
                           int64_t B, int32_t H, void *stream) {
  return evo_generation_host<double>(cand, cand_stride_n, cand_lpj, cand_lpj_stride_n, idx, cand_indexed, S_c, children,
                                     cfg, gen, sparsity, fit_norm, B, H, stream);
}
int tvo_evo_generation_f32(const uint64_t *cand, int64_t cand_stride_n, const float *cand_lpj, int64_t cand_lpj_stride_n,
                           const int64_t *idx, int32_t cand_indexed, int32_t S_c, uint64_t *children,
                           const tvo_evo_config *cfg, int32_t gen, const double *sparsity, const double *fit_norm,
                           int64_t B, int32_t H, void *stream) {
  return evo_generation_host<float>(cand, cand_stride_n, cand_lpj, cand_lpj_stride_n, idx, cand_indexed, S_c, children,
                                    cfg, gen, sparsity, fit_norm, B, H, stream);
}
int tvo_fit_norm_f64(const double *lpj, int64_t stride_n, const int64_t *idx, int64_t B, int32_t S_c, double *out,
                     void *stream) {
  return fit_norm_host<double>(lpj, stride_n, idx, B, S_c, out, stream);
}
int tvo_fit_norm_f32(const float *lpj, int64_t stride_n, const int64_t *idx, int64_t B, int32_t S_c, double *out,
                     void *stream) {
  return fit_norm_host<float>(lpj, stride_n, idx, B, S_c, out, stream);
}
int tvo_random_states(uint64_t *out, const int64_t *idx, int64_t B, int32_t S_new, int32_t H, double sparsity,
                      uint64_t seed, uint32_t step, uint32_t n_offset, void *stream) {
  TVO_REQUIRE(out, "random_states: null pointer");
  TVO_REQUIRE(B >= 0 && S_new > 0 && H > 0 && S_new < (1 << 24) && n_words(H) <= 256, "random_states: bad sizes");
  if (B == 0) return TVO_OK;
  const uint2 key = make_uint2((uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
  random_states_kernel<<<grid_for(B * S_new * n_words(H), 256, 8), 256, 0, (cudaStream_t)stream>>>(
      out, idx, B, S_new, H, sparsity, key, step, n_offset);
  TVO_AFTER_LAUNCH("random_states");
  return TVO_OK;
}
int tvo_mark_redundant_f64(const uint64_t *new_states, double *new_lpj, const uint64_t *K, int64_t k_stride_n,
                           const int64_t *idx, int64_t B, int32_t S_new, int32_t S, int32_t H, void *stream) {
  return mark_redundant_host<double>(new_states, new_lpj, K, k_stride_n, idx, B, S_new, S, H, stream);
}
int tvo_mark_redundant_f32(const uint64_t *new_states, float *new_lpj, const uint64_t *K, int64_t k_stride_n,
                           const int64_t *idx, int64_t B, int32_t S_new, int32_t S, int32_t H, void *stream) {
  return mark_redundant_host<float>(new_states, new_lpj, K, k_stride_n, idx, B, S_new, S, H, stream);
}
int tvo_select_topS_f64(const uint64_t *new_states, const double *new_lpj, uint64_t *K, double *lpj, const int64_t *idx,
                        int64_t B, int32_t S_new, int32_t S, int32_t H, int64_t *nsubs, void *stream) {
  return select_topS_host<double>(new_states, new_lpj, K, lpj, idx, B, S_new, S, H, nsubs, stream);
}
int tvo_select_topS_f32(const uint64_t *new_states, const float *new_lpj, uint64_t *K, float *lpj, const int64_t *idx,
                        int64_t B, int32_t S_new, int32_t S, int32_t H, int64_t *nsubs, void *stream) {
  return select_topS_host<float>(new_states, new_lpj, K, lpj, idx, B, S_new, S, H, nsubs, stream);
}
int tvo_lpj2pjc_f64(const double *lpj, double *pjc, int64_t N, int32_t S, void *stream) {
  return lpj2pjc_host<double>(lpj, pjc, N, S, stream);
}
int tvo_lpj2pjc_f32(const float *lpj, float *pjc, int64_t N, int32_t S, void *stream) {
  return lpj2pjc_host<float>(lpj, pjc, N, S, stream);
}
int tvo_mean_posterior_f64(const double *g, const double *lpj, double *out, int64_t N, int32_t S, int64_t G, void *stream) {
  return mean_posterior_host<double>(g, lpj, out, N, S, G, stream);
}
int tvo_mean_posterior_f32(const float *g, const float *lpj, float *out, int64_t N, int32_t S, int64_t G, void *stream) {
  return mean_posterior_host<float>(g, lpj, out, N, S, G, stream);
}
int tvo_state_marginals_f64(const uint64_t *K, int64_t k_stride_n, const double *lpj, int64_t lpj_stride_n,
                            const int64_t *idx, double *out, int64_t B, int32_t S, int32_t H, void *stream) {
  return state_marginals_host<double>(K, k_stride_n, lpj, lpj_stride_n, idx, out, B, S, H, stream);
}
int tvo_state_marginals_f32(const uint64_t *K, int64_t k_stride_n, const float *lpj, int64_t lpj_stride_n,
                            const int64_t *idx, float *out, int64_t B, int32_t S, int32_t H, void *stream) {
  return state_marginals_host<float>(K, k_stride_n, lpj, lpj_stride_n, idx, out, B, S, H, stream);
}
int tvo_sample_states_probs_f64(uint64_t *out, int64_t out_stride_n, int32_t S_off, const double *probs,
                                int64_t probs_stride_n, const int64_t *idx, int64_t B, int32_t S_new, int32_t H,
                                uint64_t seed, uint32_t step, uint32_t group, uint32_t n_offset, void *stream) {
  return sample_states_probs_host<double>(out, out_stride_n, S_off, probs, probs_stride_n, idx, B, S_new, H, seed, step,
                                          group, n_offset, stream);
}
int tvo_sample_states_probs_f32(uint64_t *out, int64_t out_stride_n, int32_t S_off, const float *probs,
                                int64_t probs_stride_n, const int64_t *idx, int64_t B, int32_t S_new, int32_t H,
                                uint64_t seed, uint32_t step, uint32_t group, uint32_t n_offset, void *stream) {
  return sample_states_probs_host<float>(out, out_stride_n, S_off, probs, probs_stride_n, idx, B, S_new, H, seed, step,
                                         group, n_offset, stream);
}
}
