// tvo_b200 — Binary Sparse Coding kernels (tvo/models/bsc.py) and the fused EVO E-step.
//
// Design notes:
//  * states are sparse bit vectors, so  Wbar_s = sum_{h in s} W[:,h]  is evaluated by walking the
//    set bits of the packed state and adding columns of a padded W^T (H x Dpad), lanes over d.
//    This does |s|*D adds instead of the reference's dense 2*H*D flop GEMM (bsc.py:105).
//  * one warp per datapoint; y lives in registers, the datapoint's K / children / lpj in shared
//    memory; nothing but y, K and lpj ever touches HBM.
#include "estep_device.cuh"

namespace tvo {

struct BscThetaHdr {
  double coef;            // -1/2/sigma2           (bsc.py:116)
  double prior_sum;       // sum_h log(1-pi_h)     (bsc.py:127-131)
  double log_2pi_sigma2;  // log(2 pi sigma2)      (bsc.py:132)
  double sigma2;
  double pad[4];
};

__host__ __device__ inline int bsc_dpad(int D) { return D <= 256 ? pad32(D) : ((D + 255) & ~255); }

template <typename T>
__host__ __device__ inline size_t bsc_prior_off() {
  return sizeof(BscThetaHdr);
}
template <typename T>
__host__ __device__ inline size_t bsc_wt_off(int H) {
  return sizeof(BscThetaHdr) + align16((size_t)H * sizeof(T));
}
template <typename T>
__host__ __device__ inline size_t bsc_ws_bytes(int H, int D) {
  return bsc_wt_off<T>(H) + (size_t)H * bsc_dpad(D) * sizeof(T);
}

template <typename T>
struct BscTheta {
  const BscThetaHdr *hdr;
  const T *prior;  // H
  const T *Wt;     // H x Dpad, zero padded
  int Dpad;
};
template <typename T>
__host__ __device__ inline BscTheta<T> bsc_theta_view(const void *ws, int H, int D) {
  BscTheta<T> t;
  const unsigned char *p = (const unsigned char *)ws;
  t.hdr = (const BscThetaHdr *)p;
  t.prior = (const T *)(p + bsc_prior_off<T>());
  t.Wt = (const T *)(p + bsc_wt_off<T>(H));
  t.Dpad = bsc_dpad(D);
  return t;
}

// ------------------------------------------------------------------------------------ prepare
template <typename T>
__global__ void bsc_prepare_kernel(const T *__restrict__ W, const T *__restrict__ pies, int n_pies,
                                   const T *__restrict__ sigma2, int H, int D, void *ws) {
  unsigned char *p = (unsigned char *)ws;
  BscThetaHdr *hdr = (BscThetaHdr *)p;
  T *prior = (T *)(p + bsc_prior_off<T>());
  T *Wt = (T *)(p + bsc_wt_off<T>(H));
  const int Dpad = bsc_dpad(D);
  const size_t total = (size_t)H * Dpad;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int h = (int)(i / Dpad), d = (int)(i % Dpad);
    Wt[i] = d < D ? W[(size_t)d * H + h] : T(0);
  }
  for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < H; h += gridDim.x * blockDim.x) {
    const T pi = pies[n_pies == 1 ? 0 : h];
    prior[h] = log(pi / (T(1) - pi));
  }
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    double s = 0.0;
    if (n_pies == 1) {
      s = (double)H * (double)log(T(1) - pies[0]);
    } else {
      for (int h = threadIdx.x; h < H; h += 32) s += (double)log(T(1) - pies[h]);
      s = warp_sum(s);
    }
    if (threadIdx.x == 0) {
      const T s2 = sigma2[0];
      hdr->coef = (double)(T(-0.5) / s2);
      hdr->prior_sum = s;
      hdr->log_2pi_sigma2 = (double)log(T(2.0 * 3.14159265358979323846) * s2);
      hdr->sigma2 = (double)s2;
    }
  }
}

// ------------------------------------------------------------------------------------ lpj core
// y of one datapoint in registers: yreg[i] = y[lane + 32 i] (0 for NaN / padding), nanmask bit i set
// where y is NaN (those terms are skipped: to.nansum, bsc.py:115).
template <typename T, int DPL>
struct YReg {
  T v[DPL];
  uint32_t nanmask;
  __device__ __forceinline__ void load(const T *__restrict__ y, int D, int lane) {
    nanmask = 0;
#pragma unroll
    for (int i = 0; i < DPL; ++i) {
      const int d = lane + 32 * i;
      T x = d < D ? y[d] : T(0);
      if (x != x) {
        nanmask |= 1u << i;
        x = T(0);
      }
      v[i] = x;
    }
  }
};

// sum_d (Wbar_d - y_d)^2 and sum_{h in s} prior_h for ONE state; the warp cooperates, lanes over d.
template <typename T, int DPL>
__device__ __forceinline__ void bsc_state_terms(const uint64_t *st, int W64, const BscTheta<T> &th,
                                                const YReg<T, DPL> &y, int lane, T &sumsq, T &prior) {
  T acc[DPL];
#pragma unroll
  for (int i = 0; i < DPL; ++i) acc[i] = T(0);
  T pr = T(0);
  for (int w = 0; w < W64; ++w) {
    uint64_t word = st[w];
    while (word) {
      const int h = (w << 6) + __ffsll((long long)word) - 1;
      word &= word - 1;
      const T *col = th.Wt + (size_t)h * th.Dpad + lane;
#pragma unroll
      for (int i = 0; i < DPL; ++i) acc[i] += __ldg(col + 32 * i);
      pr += __ldg(th.prior + h);
    }
  }
  T ss = T(0);
#pragma unroll
  for (int i = 0; i < DPL; ++i) {
    const T t = acc[i] - y.v[i];
    if (!((y.nanmask >> i) & 1u)) ss += t * t;
  }
  sumsq = warp_sum(ss);
  prior = pr;
}

// generic-D variant (D > 256): d is walked in chunks of 256, y read from shared memory
template <typename T>
__device__ inline void bsc_state_terms_big(const uint64_t *st, int W64, const BscTheta<T> &th, const T *ysm, int lane,
                                           bool skip_nan, T &sumsq, T &prior) {
  T ss = T(0), pr = T(0);
  for (int c = 0; c < th.Dpad; c += 256) {
    T acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = T(0);
    for (int w = 0; w < W64; ++w) {
      uint64_t word = st[w];
      while (word) {
        const int h = (w << 6) + __ffsll((long long)word) - 1;
        word &= word - 1;
        const T *col = th.Wt + (size_t)h * th.Dpad + c + lane;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += __ldg(col + 32 * i);
        if (c == 0) pr += __ldg(th.prior + h);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const T yv = ysm[c + lane + 32 * i];
      const T t = acc[i] - yv;
      if (!(skip_nan && yv != yv)) ss += t * t;
    }
  }
  sumsq = warp_sum(ss);
  prior = pr;
}

// ------------------------------------------------------------------------------------ lpj kernel
// log_pseudo_joint (bsc.py:100-119).  Work item = (datapoint, chunk of CH states); one warp each.
template <typename T, int DPL>
__global__ void __launch_bounds__(256)
bsc_lpj_kernel(const T *__restrict__ Y, int64_t ldY, const uint64_t *__restrict__ K, int64_t k_stride_n,
               const int64_t *__restrict__ idx, const void *ws, T *__restrict__ lpj_out, int64_t lpj_stride_n,
               int64_t B, int S, int H, int D, int CH) {
  extern __shared__ unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const BscTheta<T> th = bsc_theta_view<T>(ws, H, D);
  const int W64 = n_words(H);
  const int nchunk = (S + CH - 1) / CH;
  const int64_t ntask = B * nchunk;
  const T coef = (T)th.hdr->coef;
  T *ysm = (T *)smem_raw + (size_t)wib * (DPL == 0 ? th.Dpad : 0);
  for (int64_t task = (int64_t)blockIdx.x * wpb + wib; task < ntask; task += (int64_t)gridDim.x * wpb) {
    const int64_t b = task / nchunk;
    const int c = (int)(task % nchunk);
    const int64_t row = idx ? idx[b] : b;
    const uint64_t *Kn = K + row * k_stride_n;
    T *out = lpj_out + row * lpj_stride_n;
    const int s1 = min(S, (c + 1) * CH);
    if constexpr (DPL > 0) {
      YReg<T, DPL> y;
      y.load(Y + b * ldY, D, lane);
      for (int s = c * CH; s < s1; ++s) {
        T ss, pr;
        bsc_state_terms<T, DPL>(Kn + (size_t)s * W64, W64, th, y, lane, ss, pr);
        if (lane == 0) out[s] = ss * coef + pr;
      }
    } else {
      __syncwarp();
      for (int d = lane; d < th.Dpad; d += 32) ysm[d] = d < D ? Y[b * ldY + d] : T(0);
      __syncwarp();
      for (int s = c * CH; s < s1; ++s) {
        T ss, pr;
        bsc_state_terms_big<T>(Kn + (size_t)s * W64, W64, th, ysm, lane, true, ss, pr);
        if (lane == 0) out[s] = ss * coef + pr;
      }
    }
  }
}

// ------------------------------------------------------------------------------------ fused E-step
// EVOVariationalStates.update (evo.py:80-115) for BSC in ONE kernel: re-score K (:95), G generations
// of select + mutate + score (:191-207) with the children in shared memory, dedup against the
// original K (:209), top-S merge and nsubs (_utils.py:124-179).
template <typename T, int DPL>
__global__ void __launch_bounds__(256)
estep_evo_bsc_kernel(const T *__restrict__ Y, int64_t ldY, const int64_t *__restrict__ idx, uint64_t *__restrict__ K,
                     T *__restrict__ lpj, const void *ws, EvoParams ep, const double *__restrict__ sparsity_p,
                     EstepDims dims, int64_t B, int H, int D, unsigned long long *nsubs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const BscTheta<T> th = bsc_theta_view<T>(ws, H, D);
  const size_t wbytes = warp_buf_bytes<T>(dims);
  const WarpBuf<T> wb = carve_warp_buf<T>(smem_raw + (size_t)wib * wbytes, dims);
  const int S = dims.S, Snew = dims.Snew, Sgen = dims.Sgen, W64 = dims.W64;
  const T coef = (T)th.hdr->coef;
  const double sparsity = sparsity_p ? *sparsity_p : 0.0;
  unsigned long long my_subs = 0;

  for (int64_t b = (int64_t)blockIdx.x * wpb + wib; b < B; b += (int64_t)gridDim.x * wpb) {
    const int64_t row = idx ? idx[b] : b;
    const uint32_t n_glob = ep.n_offset + (uint32_t)row;
    uint64_t *Krow = K + row * (int64_t)S * W64;
    T *lrow = lpj + row * (int64_t)S;
    YReg<T, DPL> y;
    y.load(Y + b * ldY, D, lane);
    __syncwarp();
    for (int t = lane; t < S * W64; t += 32) wb.oldK[t] = Krow[t];
    __syncwarp();
    // lpj[idx] = lpj_fn(batch, K[idx])   (evo.py:95)
    for (int s = 0; s < S; ++s) {
      T ss, pr;
      bsc_state_terms<T, DPL>(wb.oldK + (size_t)s * W64, W64, th, y, lane, ss, pr);
      if (lane == 0) wb.lpj[s] = ss * coef + pr;
    }
    __syncwarp();
    for (int g = 0; g < ep.G; ++g) {
      const uint64_t *cand = g == 0 ? wb.oldK : wb.newK + (size_t)(g - 1) * Sgen * W64;
      const T *cand_lpj = g == 0 ? wb.lpj : wb.lpj + S + (g - 1) * Sgen;
      uint64_t *children = wb.newK + (size_t)g * Sgen * W64;
      evo_generation<T>(wb, cand, cand_lpj, g == 0 ? S : Sgen, children, ep, g, n_glob, W64, H, sparsity, 0.0, 1.0,
                        lane);
      for (int j = 0; j < Sgen; ++j) {
        T ss, pr;
        bsc_state_terms<T, DPL>(children + (size_t)j * W64, W64, th, y, lane, ss, pr);
        if (lane == 0) wb.lpj[S + g * Sgen + j] = ss * coef + pr;
      }
      __syncwarp();
    }
    mark_redundant<T>(wb, S, Snew, W64, lane);
    my_subs += (unsigned long long)select_topS<T>(wb, S, Snew, W64, Krow, lrow, lane);
    __syncwarp();
  }
  if (lane == 0 && my_subs) atomicAdd(nsubs, my_subs);
}

// ------------------------------------------------------------------------------------ free energy
// F += sum_b [ logsumexp_s lpj[b,s] + sum_h log(1-pi_h) - D_b/2 log(2 pi sigma2) ]
// (model_protocols.py:179-193 with bsc.py:121-132).  One warp per datapoint.
template <typename T>
__global__ void __launch_bounds__(256)
bsc_free_energy_kernel(const T *__restrict__ Y, int64_t ldY, const T *__restrict__ lpj, int64_t lpj_stride_n,
                       const int64_t *__restrict__ idx, const void *ws, int64_t B, int S, int D, double *F_out) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const BscThetaHdr *hdr = (const BscThetaHdr *)ws;
  double acc = 0.0;
  for (int64_t b = (int64_t)blockIdx.x * wpb + wib; b < B; b += (int64_t)gridDim.x * wpb) {
    const int64_t row = idx ? idx[b] : b;
    const T *l = lpj + row * lpj_stride_n;
    double mx = -INFINITY;
    for (int s = lane; s < S; s += 32) mx = fmax(mx, (double)l[s]);
    mx = warp_max(mx);
    if (!(mx > -INFINITY && mx < INFINITY)) mx = mx != mx ? mx : 0.0;
    double se = 0.0;
    for (int s = lane; s < S; s += 32) se += exp((double)l[s] - mx);
    se = warp_sum(se);
    int dn = 0;
    if (Y) {
      for (int d = lane; d < D; d += 32) {
        const T v = Y[b * ldY + d];
        dn += (v == v) ? 1 : 0;
      }
      dn = warp_sum(dn);
    } else {
      dn = D;
    }
    acc += mx + log(se) + hdr->prior_sum - 0.5 * (double)dn * hdr->log_2pi_sigma2;
  }
  __shared__ double part[8];
  if (lane == 0) part[wib] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < wpb; ++w) s += part[w];
    if (s != 0.0 || s != s) atomicAdd(F_out, s);
  }
}

// ------------------------------------------------------------------------------------ M-step
// BSC.update_param_batch (bsc.py:134-158), one warp per datapoint, everything sparse:
//   q_s = softmax(lpj)                     registers / shared only, never in HBM
//   U   = union of the bits of the datapoint's S states (|U| << H)
//   E_h = sum_s q_s s_h  (h in U)          -> pies[h] += E_h ; WpT[h,:] += E_h * y
//   Wq[h,h'] += sum_{s: s_h s_h'} q_s      via per-bit membership masks over s
//   sigma2 += sum_s q_s ||y - W s||^2      derived from lpj (fast) or recomputed (exact)
// stats layout: [WpT (H*D) | Wq (H*H) | pies (H) | sigma2 | N | F]
template <typename T>
__global__ void __launch_bounds__(256)
bsc_mstep_kernel(const T *__restrict__ Y, int64_t ldY, const int64_t *__restrict__ idx, const uint64_t *__restrict__ K,
                 int64_t k_stride_n, const T *__restrict__ lpj, int64_t lpj_stride_n, const void *ws,
                 double *__restrict__ stats, int64_t B, int S, int H, int D, int exact_resid, size_t warp_bytes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const BscTheta<T> th = bsc_theta_view<T>(ws, H, D);
  const int W64 = n_words(H), SW = (S + 63) >> 6;
  unsigned char *p = smem_raw + (size_t)wib * warp_bytes;
  uint64_t *Ksm = (uint64_t *)p;
  p += align16((size_t)S * W64 * 8);
  uint64_t *colm = (uint64_t *)p;  // per union bit: membership mask over s (SW words)
  p += align16((size_t)H * SW * 8);
  double *q = (double *)p;
  p += align16((size_t)S * 8);
  double *Esm = (double *)p;
  p += align16((size_t)H * 8);
  int32_t *hlist = (int32_t *)p;
  p += align16((size_t)H * 4);
  T *ysm = (T *)p;  // Dpad (exact mode only)

  double *WpT = stats, *Wq = stats + (size_t)H * D, *pies = Wq + (size_t)H * H;
  double *sig = pies + H, *Ncnt = sig + 1, *Facc = Ncnt + 1;
  const double coef = th.hdr->coef;
  double acc_sig = 0.0, acc_F = 0.0, acc_N = 0.0;

  for (int64_t b = (int64_t)blockIdx.x * wpb + wib; b < B; b += (int64_t)gridDim.x * wpb) {
    const int64_t row = idx ? idx[b] : b;
    const uint64_t *Kn = K + row * k_stride_n;
    const T *l = lpj + row * lpj_stride_n;
    const T *y = Y + b * ldY;
    __syncwarp();
    for (int t = lane; t < S * W64; t += 32) Ksm[t] = Kn[t];
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
    __syncwarp();
    for (int s = lane; s < S; s += 32) q[s] = q[s] / se;
    int dn = 0;
    for (int d = lane; d < D; d += 32) {
      const T v = y[d];
      dn += (v == v) ? 1 : 0;
    }
    dn = warp_sum(dn);
    acc_F += mx + log(se) + th.hdr->prior_sum - 0.5 * (double)dn * th.hdr->log_2pi_sigma2;
    acc_N += 1.0;
    __syncwarp();
    // union of active bits -> hlist[0..nU)
    int nU = 0;
    for (int w = 0; w < W64; ++w) {
      uint64_t u = 0;
      for (int s = lane; s < S; s += 32) u |= Ksm[(size_t)s * W64 + w];
      uint32_t lo = __reduce_or_sync(FULL, (uint32_t)u), hi = __reduce_or_sync(FULL, (uint32_t)(u >> 32));
      uint64_t U = (uint64_t)hi << 32 | lo;
      // lane t takes the t-th set bit (t < popc); two rounds cover 64 bits
      const int cnt = __popcll(U);
      for (int base = 0; base < cnt; base += 32) {
        const int t = base + lane;
        if (t < cnt) {
          // position of the t-th set bit of U
          uint64_t m = U;
          for (int k = 0; k < t; ++k) m &= m - 1;
          hlist[nU + t] = (w << 6) + __ffsll((long long)m) - 1;
        }
      }
      nU += cnt;
    }
    __syncwarp();
    // E_h and membership masks
    for (int t = lane; t < nU; t += 32) {
      const int h = hlist[t];
      const int w = h >> 6, bit = h & 63;
      double e = 0.0;
      for (int sw = 0; sw < SW; ++sw) {
        uint64_t mask = 0;
        const int s1 = min(S, (sw + 1) * 64);
        for (int s = sw * 64; s < s1; ++s) {
          const uint64_t on = (Ksm[(size_t)s * W64 + w] >> bit) & 1ull;
          mask |= on << (s & 63);
          if (on) e += q[s];
        }
        colm[(size_t)t * SW + sw] = mask;
      }
      Esm[t] = e;
      atomicAdd(pies + h, e);
    }
    __syncwarp();
    // WpT[h, :] += E_h * y   (bsc.py:145,153; stored transposed so the RED is coalesced over d)
    for (int t = 0; t < nU; ++t) {
      const double e = Esm[t];
      if (e == 0.0) continue;
      double *dst = WpT + (size_t)hlist[t] * D;
      for (int d = lane; d < D; d += 32) atomicAdd(dst + d, e * (double)y[d]);
    }
    // Wq[h, h'] += sum_{s in C_h & C_h'} q_s   (bsc.py:146-147)
    for (int t = lane; t < nU; t += 32) {
      const int h = hlist[t];
      for (int t2 = 0; t2 < nU; ++t2) {
        double v = 0.0;
        bool any = false;
        for (int sw = 0; sw < SW; ++sw) {
          uint64_t m = colm[(size_t)t * SW + sw] & colm[(size_t)t2 * SW + sw];
          any |= (m != 0);
          while (m) {
            const int s = (sw << 6) + __ffsll((long long)m) - 1;
            m &= m - 1;
            v += q[s];
          }
        }
        if (any) atomicAdd(Wq + (size_t)h * H + hlist[t2], v);
      }
    }
    // sigma2 += <||y - W s||^2>_q   (bsc.py:148-150)
    if (!exact_resid) {
      double part = 0.0;
      for (int s = lane; s < S; s += 32) {
        double pr = 0.0;
        for (int w = 0; w < W64; ++w) {
          uint64_t word = Ksm[(size_t)s * W64 + w];
          while (word) {
            const int h = (w << 6) + __ffsll((long long)word) - 1;
            word &= word - 1;
            pr += (double)__ldg(th.prior + h);
          }
        }
        part += q[s] * (((double)l[s] - pr) / coef);
      }
      acc_sig += warp_sum(part);
    } else {
      __syncwarp();
      for (int d = lane; d < th.Dpad; d += 32) ysm[d] = d < D ? y[d] : T(0);
      __syncwarp();
      double part = 0.0;
      for (int s = 0; s < S; ++s) {
        // plain sum, no NaN skipping (bsc.py:149 uses to.sum)
        T ss = T(0);
        for (int c = 0; c < th.Dpad; c += 32) {
          T a = T(0);
          for (int w = 0; w < W64; ++w) {
            uint64_t word = Ksm[(size_t)s * W64 + w];
            while (word) {
              const int h = (w << 6) + __ffsll((long long)word) - 1;
              word &= word - 1;
              a += __ldg(th.Wt + (size_t)h * th.Dpad + c + lane);
            }
          }
          const T t = ysm[c + lane] - a;
          ss += t * t;
        }
        part += q[s] * (double)warp_sum(ss);
      }
      acc_sig += part;
    }
  }
  if (lane == 0) {
    if (acc_N != 0.0) {
      atomicAdd(sig, acc_sig);
      atomicAdd(Ncnt, acc_N);
      atomicAdd(Facc, acc_F);
    }
  }
}

template <typename T>
size_t bsc_mstep_warp_bytes(int S, int H, int D) {
  const int W64 = n_words(H), SW = (S + 63) >> 6;
  return align16((size_t)S * W64 * 8) + align16((size_t)H * SW * 8) + align16((size_t)S * 8) + align16((size_t)H * 8) +
         align16((size_t)H * 4) + align16((size_t)bsc_dpad(D) * sizeof(T));
}

// ------------------------------------------------------------------------------------ host side
constexpr int kMaxSmem = 227 * 1024;

template <typename T>
int bsc_prepare(const T *W, const T *pies, int n_pies, const T *sigma2, int H, int D, void *ws, void *stream) {
  TVO_REQUIRE(W && pies && sigma2 && ws, "bsc_prepare: null pointer");
  TVO_REQUIRE(H > 0 && D > 0, "bsc_prepare: H=%d D=%d must be positive", H, D);
  TVO_REQUIRE(n_pies == 1 || n_pies == H, "bsc_prepare: n_pies=%d must be 1 or H=%d", n_pies, H);
  const size_t total = (size_t)H * bsc_dpad(D);
  const int blocks = (int)min((size_t)4 * sm_count(), (total + 255) / 256);
  bsc_prepare_kernel<T><<<max(blocks, 1), 256, 0, (cudaStream_t)stream>>>(W, pies, n_pies, sigma2, H, D, ws);
  TVO_AFTER_LAUNCH("bsc_prepare");
  return TVO_OK;
}

template <typename T, int DPL>
int launch_bsc_lpj(const T *Y, int64_t ldY, const uint64_t *K, int64_t k_stride_n, const int64_t *idx, const void *ws,
                   T *lpj_out, int64_t lpj_stride_n, int64_t B, int S, int H, int D, cudaStream_t stream) {
  const int CH = S <= 64 ? S : 64;
  const int64_t ntask = B * ((S + CH - 1) / CH);
  const int wpb = 8;
  const size_t smem = DPL == 0 ? (size_t)wpb * bsc_dpad(D) * sizeof(T) : 0;
  const int64_t want = (ntask + wpb - 1) / wpb;
  const int grid = (int)max((int64_t)1, min(want, (int64_t)sm_count() * 8));
  auto kern = bsc_lpj_kernel<T, DPL>;
  if (smem > 48 * 1024) TVO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, wpb * 32, smem, stream>>>(Y, ldY, K, k_stride_n, idx, ws, lpj_out, lpj_stride_n, B, S, H, D, CH);
  TVO_AFTER_LAUNCH("bsc_lpj");
  return TVO_OK;
}

template <typename T>
int bsc_lpj(const T *Y, int64_t ldY, const uint64_t *K, int64_t k_stride_n, const int64_t *idx, const void *ws,
            T *lpj_out, int64_t lpj_stride_n, int64_t B, int S, int H, int D, void *stream) {
  TVO_REQUIRE(Y && K && ws && lpj_out, "bsc_lpj: null pointer");
  TVO_REQUIRE(B >= 0 && S > 0 && H > 0 && D > 0, "bsc_lpj: bad sizes B=%lld S=%d H=%d D=%d", (long long)B, S, H, D);
  if (B == 0) return TVO_OK;
  cudaStream_t st = (cudaStream_t)stream;
#define TVO_DPL_CASE(n) \
  case n:               \
    return launch_bsc_lpj<T, n>(Y, ldY, K, k_stride_n, idx, ws, lpj_out, lpj_stride_n, B, S, H, D, st);
  if (D > 256) return launch_bsc_lpj<T, 0>(Y, ldY, K, k_stride_n, idx, ws, lpj_out, lpj_stride_n, B, S, H, D, st);
  switch (bsc_dpad(D) / 32) {
    TVO_DPL_CASE(1) TVO_DPL_CASE(2) TVO_DPL_CASE(3) TVO_DPL_CASE(4) TVO_DPL_CASE(5) TVO_DPL_CASE(6) TVO_DPL_CASE(7)
    TVO_DPL_CASE(8)
  }
#undef TVO_DPL_CASE
  return set_error(TVO_EINVAL, "bsc_lpj: unsupported D=%d", D);
}

int evo_params_from_cfg(const tvo_evo_config *cfg, int S, int H, EvoParams &ep, EstepDims &dims, const char *who);

template <typename T, int DPL>
int launch_estep_evo_bsc(const T *Y, int64_t ldY, const int64_t *idx, uint64_t *K, T *lpj, const void *ws,
                         const EvoParams &ep, const double *sparsity, const EstepDims &dims, int64_t B, int H, int D,
                         int64_t *nsubs, cudaStream_t stream) {
  const size_t wbytes = warp_buf_bytes<T>(dims);
  int wpb = 8;
  while (wpb > 1 && wbytes * wpb > (size_t)kMaxSmem) wpb >>= 1;
  TVO_REQUIRE(wbytes * wpb <= (size_t)kMaxSmem, "estep_evo_bsc: per-datapoint working set %zu B exceeds shared memory",
              wbytes);
  const size_t smem = wbytes * wpb;
  auto kern = estep_evo_bsc_kernel<T, DPL>;
  TVO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  TVO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpb * 32, smem));
  occ = max(occ, 1);
  const int64_t want = (B + wpb - 1) / wpb;
  const int grid = (int)max((int64_t)1, min(want, (int64_t)sm_count() * occ));
  kern<<<grid, wpb * 32, smem, stream>>>(Y, ldY, idx, K, lpj, ws, ep, sparsity, dims, B, H, D,
                                         (unsigned long long *)nsubs);
  TVO_AFTER_LAUNCH("estep_evo_bsc");
  return TVO_OK;
}

template <typename T>
int estep_evo_bsc(const T *Y, int64_t ldY, const int64_t *idx, uint64_t *K, T *lpj, const void *ws,
                  const tvo_evo_config *cfg, const double *sparsity, int64_t B, int S, int H, int D, int64_t *nsubs,
                  void *stream) {
  TVO_REQUIRE(Y && K && lpj && ws && cfg && nsubs, "estep_evo_bsc: null pointer");
  TVO_REQUIRE(B >= 0 && D > 0 && D <= 256, "estep_evo_bsc: need 0 < D <= 256 (got %d); use the unfused path", D);
  EvoParams ep;
  EstepDims dims;
  int rc = evo_params_from_cfg(cfg, S, H, ep, dims, "estep_evo_bsc");
  if (rc) return rc;
  TVO_REQUIRE(ep.sel != TVO_SEL_FITPARENTS,
              "estep_evo_bsc: batch-global fitparents needs a grid-wide reduction per generation; use the unfused "
              "kernels (or TVO_SEL_FITPARENTS_FIXED)");
  TVO_REQUIRE(!(ep.mut == TVO_MUT_SPARSEFLIP || ep.mut == TVO_MUT_CROSS_SPARSEFLIP) || sparsity,
              "estep_evo_bsc: sparseflip needs the device scalar `sparsity`");
  if (B == 0) return TVO_OK;
  dims.Dpad = 0;
  cudaStream_t st = (cudaStream_t)stream;
#define TVO_DPL_CASE(n) \
  case n:               \
    return launch_estep_evo_bsc<T, n>(Y, ldY, idx, K, lpj, ws, ep, sparsity, dims, B, H, D, nsubs, st);
  switch (bsc_dpad(D) / 32) {
    TVO_DPL_CASE(1) TVO_DPL_CASE(2) TVO_DPL_CASE(3) TVO_DPL_CASE(4) TVO_DPL_CASE(5) TVO_DPL_CASE(6) TVO_DPL_CASE(7)
    TVO_DPL_CASE(8)
  }
#undef TVO_DPL_CASE
  return set_error(TVO_EINVAL, "estep_evo_bsc: unsupported D=%d", D);
}

template <typename T>
int bsc_free_energy(const T *Y, int64_t ldY, const T *lpj, int64_t lpj_stride_n, const int64_t *idx, const void *ws,
                    int64_t B, int S, int D, double *F_out, void *stream) {
  TVO_REQUIRE(lpj && ws && F_out, "bsc_free_energy: null pointer");
  TVO_REQUIRE(B >= 0 && S > 0 && D > 0, "bsc_free_energy: bad sizes");
  if (B == 0) return TVO_OK;
  const int grid = (int)max((int64_t)1, min((B + 7) / 8, (int64_t)sm_count() * 8));
  bsc_free_energy_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(Y, ldY, lpj, lpj_stride_n, idx, ws, B, S, D, F_out);
  TVO_AFTER_LAUNCH("bsc_free_energy");
  return TVO_OK;
}

template <typename T>
int bsc_mstep(const T *Y, int64_t ldY, const int64_t *idx, const uint64_t *K, int64_t k_stride_n, const T *lpj,
              int64_t lpj_stride_n, const void *ws, double *stats, int64_t B, int S, int H, int D, int exact,
              void *stream) {
  TVO_REQUIRE(Y && K && lpj && ws && stats, "bsc_mstep: null pointer");
  TVO_REQUIRE(B >= 0 && S > 0 && H > 0 && D > 0, "bsc_mstep: bad sizes");
  if (B == 0) return TVO_OK;
  const size_t wbytes = bsc_mstep_warp_bytes<T>(S, H, D);
  int wpb = 8;
  while (wpb > 1 && wbytes * wpb > (size_t)kMaxSmem) wpb >>= 1;
  TVO_REQUIRE(wbytes * wpb <= (size_t)kMaxSmem, "bsc_mstep: working set %zu B exceeds shared memory", wbytes);
  const size_t smem = wbytes * wpb;
  auto kern = bsc_mstep_kernel<T>;
  TVO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  TVO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpb * 32, smem));
  occ = max(occ, 1);
  const int grid = (int)max((int64_t)1, min((B + wpb - 1) / wpb, (int64_t)sm_count() * occ));
  kern<<<grid, wpb * 32, smem, (cudaStream_t)stream>>>(Y, ldY, idx, K, k_stride_n, lpj, lpj_stride_n, ws, stats, B, S, H,
                                                       D, exact, wbytes);
  TVO_AFTER_LAUNCH("bsc_mstep");
  return TVO_OK;
}

}  // namespace tvo

using namespace tvo;

extern "C" {
int64_t tvo_bsc_theta_ws_bytes(int32_t H, int32_t D, int32_t is_f64) {
  return (int64_t)(is_f64 ? bsc_ws_bytes<double>(H, D) : bsc_ws_bytes<float>(H, D));
}
int64_t tvo_bsc_stats_len(int32_t H, int32_t D) { return (int64_t)H * D + (int64_t)H * H + H + 3; }

int tvo_bsc_prepare_f64(const double *W, const double *pies, int32_t n_pies, const double *sigma2, int32_t H, int32_t D,
                        void *ws, void *stream) {
  return bsc_prepare<double>(W, pies, n_pies, sigma2, H, D, ws, stream);
}
int tvo_bsc_prepare_f32(const float *W, const float *pies, int32_t n_pies, const float *sigma2, int32_t H, int32_t D,
                        void *ws, void *stream) {
  return bsc_prepare<float>(W, pies, n_pies, sigma2, H, D, ws, stream);
}
int tvo_bsc_lpj_f64(const double *Y, int64_t ldY, const uint64_t *K, int64_t k_stride_n, const int64_t *idx,
                    const void *ws, double *lpj_out, int64_t lpj_stride_n, int64_t B, int32_t S, int32_t H, int32_t D,
                    void *stream) {
  return bsc_lpj<double>(Y, ldY, K, k_stride_n, idx, ws, lpj_out, lpj_stride_n, B, S, H, D, stream);
}
int tvo_bsc_lpj_f32(const float *Y, int64_t ldY, const uint64_t *K, int64_t k_stride_n, const int64_t *idx,
                    const void *ws, float *lpj_out, int64_t lpj_stride_n, int64_t B, int32_t S, int32_t H, int32_t D,
                    void *stream) {
  return bsc_lpj<float>(Y, ldY, K, k_stride_n, idx, ws, lpj_out, lpj_stride_n, B, S, H, D, stream);
}
int tvo_estep_evo_bsc_f64(const double *Y, int64_t ldY, const int64_t *idx, uint64_t *K, double *lpj, const void *ws,
                          const tvo_evo_config *cfg, const double *sparsity, double *fit_norm, int64_t B, int32_t S,
                          int32_t H, int32_t D, int64_t *nsubs, void *stream) {
  (void)fit_norm;
  return estep_evo_bsc<double>(Y, ldY, idx, K, lpj, ws, cfg, sparsity, B, S, H, D, nsubs, stream);
}
int tvo_estep_evo_bsc_f32(const float *Y, int64_t ldY, const int64_t *idx, uint64_t *K, float *lpj, const void *ws,
                          const tvo_evo_config *cfg, const double *sparsity, double *fit_norm, int64_t B, int32_t S,
                          int32_t H, int32_t D, int64_t *nsubs, void *stream) {
  (void)fit_norm;
  return estep_evo_bsc<float>(Y, ldY, idx, K, lpj, ws, cfg, sparsity, B, S, H, D, nsubs, stream);
}
int tvo_bsc_free_energy_f64(const double *Y, int64_t ldY, const double *lpj, int64_t lpj_stride_n, const int64_t *idx,
                            const void *ws, int64_t B, int32_t S, int32_t D, double *F_out, void *stream) {
  return bsc_free_energy<double>(Y, ldY, lpj, lpj_stride_n, idx, ws, B, S, D, F_out, stream);
}
int tvo_bsc_free_energy_f32(const float *Y, int64_t ldY, const float *lpj, int64_t lpj_stride_n, const int64_t *idx,
                            const void *ws, int64_t B, int32_t S, int32_t D, double *F_out, void *stream) {
  return bsc_free_energy<float>(Y, ldY, lpj, lpj_stride_n, idx, ws, B, S, D, F_out, stream);
}
int tvo_bsc_mstep_f64(const double *Y, int64_t ldY, const int64_t *idx, const uint64_t *K, int64_t k_stride_n,
                      const double *lpj, int64_t lpj_stride_n, const void *ws, double *stats, int64_t B, int32_t S,
                      int32_t H, int32_t D, int32_t exact_residual, void *stream) {
  return bsc_mstep<double>(Y, ldY, idx, K, k_stride_n, lpj, lpj_stride_n, ws, stats, B, S, H, D, exact_residual, stream);
}
int tvo_bsc_mstep_f32(const float *Y, int64_t ldY, const int64_t *idx, const uint64_t *K, int64_t k_stride_n,
                      const float *lpj, int64_t lpj_stride_n, const void *ws, double *stats, int64_t B, int32_t S,
                      int32_t H, int32_t D, int32_t exact_residual, void *stream) {
  return bsc_mstep<float>(Y, ldY, idx, K, k_stride_n, lpj, lpj_stride_n, ws, stats, B, S, H, D, exact_residual, stream);
}
}
