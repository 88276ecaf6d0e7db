// tvo_b200 — BSC theta workspace layout and per-state device functions shared by bsc.cu / bsc_fast.cu.
#pragma once
#include "estep_device.cuh"

namespace tvo {

struct BscThetaHdr {
  double coef;            // -1/2/sigma2           (bsc.py:116)
  double prior_sum;       // sum_h log(1-pi_h)     (bsc.py:127-131)
  double log_2pi_sigma2;  // log(2 pi sigma2)      (bsc.py:132)
  double sigma2;
  double sparsity;        // mean_h pi_h in the model's precision: `sparsity` of the sparseflip mutation (evo.py:109)
  double pad[3];
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
__host__ __device__ inline size_t bsc_gram_off(int H, int D) {
  return align16(bsc_wt_off<T>(H) + (size_t)H * bsc_dpad(D) * sizeof(T));
}
// G2 = 2 coef G (H x H fp64) and gp_h = coef G_hh + prior_h (H fp64): the list-based fused E-step (bsc_estep_v2.cuh)
// evaluates  lpj = coef ||y||^2 + sum_{h in s} (gp_h - 2 coef a_h) + sum_{h<h' in s} G2_hh'
template <typename T>
__host__ __device__ inline size_t bsc_gram2_off(int H, int D) {
  return bsc_gram_off<T>(H, D) + (size_t)H * H * sizeof(double);
}
template <typename T>
__host__ __device__ inline size_t bsc_gp_off(int H, int D) {
  return bsc_gram2_off<T>(H, D) + (size_t)H * H * sizeof(double);
}
// Wg: W in the operand layout of the hand-written FP64 tensor-core GEMM a = Y W (bsc_gemm.cuh): [k/4][h][k%4] doubles,
// k zero-padded to a multiple of 16, so that a K-chunk is one contiguous bulk copy and a warp's B fragments are
// contiguous 256-byte shared-memory reads.
__host__ __device__ inline int bsc_kpad16(int D) { return (D + 15) & ~15; }
template <typename T>
__host__ __device__ inline size_t bsc_wg_off(int H, int D) {
  return bsc_gp_off<T>(H, D) + align16((size_t)H * sizeof(double));
}
// fp32 model: the pre-split (TF32 hi / lo), pre-swizzled images of W for the tcgen05 GEMM (bsc_gemm_tf32.cuh):
// per K block of 32 a hi and a lo tile of H rows x 128 bytes.  The region replaces Wg (fp64 model only).
__host__ __device__ inline size_t bsc_wtf_bytes(int H, int D) { return (size_t)((D + 31) / 32) * 2 * H * 128; }
template <typename T>
__host__ __device__ inline size_t bsc_ws_bytes(int H, int D) {
  const size_t o = (bsc_wg_off<T>(H, D) + 1023) & ~(size_t)1023;
  return o + (sizeof(T) == 8 ? (size_t)bsc_kpad16(D) * H * sizeof(double) : bsc_wtf_bytes(H, D));
}

template <typename T>
struct BscTheta {
  const BscThetaHdr *hdr;
  const T *prior;  // H
  const T *Wt;     // H x Dpad, zero padded
  const double *G; // H x H Gram matrix W^T W (fp64)
  const double *G2;  // 2 coef G
  const double *gp;  // coef G_hh + prior_h
  const double *Wg;  // [kpad16/4][H][4]: W for the tensor-core GEMM
  int Dpad;
};
template <typename T>
__host__ __device__ inline BscTheta<T> bsc_theta_view(const void *ws, int H, int D) {
  BscTheta<T> t;
  const unsigned char *p = (const unsigned char *)ws;
  t.hdr = (const BscThetaHdr *)p;
  t.prior = (const T *)(p + bsc_prior_off<T>());
  t.Wt = (const T *)(p + bsc_wt_off<T>(H));
  t.G = (const double *)(p + bsc_gram_off<T>(H, D));
  t.G2 = (const double *)(p + bsc_gram2_off<T>(H, D));
  t.gp = (const double *)(p + bsc_gp_off<T>(H, D));
  t.Wg = (const double *)(p + ((bsc_wg_off<T>(H, D) + 1023) & ~(size_t)1023));  // fp32 model: the TF32 images
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
  if (sizeof(T) == 8) {
    double *Wg = (double *)(p + ((bsc_wg_off<T>(H, D) + 1023) & ~(size_t)1023));
    const size_t nwg = (size_t)bsc_kpad16(D) * H;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwg; i += (size_t)gridDim.x * blockDim.x) {
      const int kk = (int)(i & 3), h = (int)((i >> 2) % H), k4 = (int)((i >> 2) / H);
      const int d = 4 * k4 + kk;
      Wg[i] = d < D ? (double)W[(size_t)d * H + h] : 0.0;
    }
  }
  // G = W^T W in fp64 (read straight from W so that it does not depend on Wt being complete)
  double *G = (double *)(p + bsc_gram_off<T>(H, D));
  double *G2 = (double *)(p + bsc_gram2_off<T>(H, D));
  double *gp = (double *)(p + bsc_gp_off<T>(H, D));
  const double coef_d = (double)(T(-0.5) / sigma2[0]);  // == hdr->coef
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)H * H; i += (size_t)gridDim.x * blockDim.x) {
    const int h1 = (int)(i / H), h2 = (int)(i % H);
    double acc = 0.0;
    for (int d = 0; d < D; ++d) acc += (double)W[(size_t)d * H + h1] * (double)W[(size_t)d * H + h2];
    G[i] = acc;
    G2[i] = 2.0 * coef_d * acc;
    if (h1 == h2) {
      const T pi = pies[n_pies == 1 ? 0 : h1];
      gp[h1] = coef_d * acc + (double)log(pi / (T(1) - pi));  // the same expression as prior[h1] above
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    double s = 0.0, pm = 0.0;
    if (n_pies == 1) {
      s = (double)H * (double)log(T(1) - pies[0]);
      pm = (double)pies[0];
    } else {
      for (int h = threadIdx.x; h < H; h += 32) {
        s += (double)log(T(1) - pies[h]);
        pm += (double)pies[h];
      }
      s = warp_sum(s);
      pm = (double)(T)(warp_sum(pm) / (double)H);
    }
    if (threadIdx.x == 0) {
      hdr->sparsity = pm;
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

}  // namespace tvo
