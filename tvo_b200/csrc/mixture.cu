// tvo_b200 — single-cause mixture models under full EM:
//   GMM  tvo/models/gmm.py:91-131   lpj[n,c] = -1/(2 sigma2) sum_d (W[d,c] - y[n,d])^2 + log pi_c
//   PMM  tvo/models/pmm.py:76-113   lpj[n,c] = sum_d y[n,d] log(W[d,c] + tiny) - sum_d (W[d,c] + tiny) + log pi_c
// evaluated for ALL C = H causes of every datapoint (FullEMSingleCauseModels, fullem.py:61-82: K_n = identity),
// i.e. a dense (N x D) x (D x C) contraction.  The reference materialises Wbar (N,C,D) and (N,D,H) outer
// products; here
//   * the lpj is one tiled kernel (64 datapoints x 64 causes per CTA, operands staged in shared memory, 4x4
//     register tile per thread).  GMM accumulates (w - y)^2 directly — the ||y||^2 - 2 y.w + ||w||^2 expansion
//     that would make it a library GEMM cancels catastrophically for tight clusters — so it is a hand-written
//     contraction on the FP64 pipe;
//   * the M-step kernel (one warp per datapoint) turns the lpj row into posterior weights on chip, accumulates
//     sum q (pies, Wq), the sigma2 term, N and the free energy, and writes q per chunk so that
//     Wp += Y^T q is one DGEMM.
#include <cublas_v2.h>

#include "common.cuh"

namespace tvo {

constexpr double kTinyF32 = 1.1754943508222875e-38;  // to.finfo(to.float32).tiny (pmm.py:82)

struct MixHdr {
  double coef;        // GMM: -1/(2 sigma2);  PMM: 1
  double lj_const;    // GMM: -D/2 log(2 pi sigma2) (gmm.py:109);  PMM: 0 (per-datapoint lgamma term instead)
  double pad[6];
};
// workspace: hdr | bias (H, T) | Weff (D x H row-major, T): GMM W, PMM log(W + tiny)
template <typename T>
__host__ __device__ inline size_t mix_off_w(int H) { return sizeof(MixHdr) + align16((size_t)H * sizeof(T)); }
template <typename T>
__host__ __device__ inline size_t mix_ws_bytes(int H, int D) { return mix_off_w<T>(H) + (size_t)D * H * sizeof(T); }

template <typename T>
__global__ void mixture_prepare_kernel(int kind, const T *__restrict__ W, const T *__restrict__ pies,
                                       const T *__restrict__ sigma2, int H, int D, void *ws) {
  unsigned char *p = (unsigned char *)ws;
  MixHdr *hdr = (MixHdr *)p;
  T *bias = (T *)(p + sizeof(MixHdr));
  T *Weff = (T *)(p + mix_off_w<T>(H));
  const size_t gt = (size_t)gridDim.x * blockDim.x, g0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = g0; i < (size_t)D * H; i += gt) Weff[i] = kind == 0 ? W[i] : (T)log(W[i] + (T)kTinyF32);
  for (size_t c = g0; c < (size_t)H; c += gt) {
    T b = (T)log(pies[c]);
    if (kind == 1) {
      T sw = T(0);
      for (int d = 0; d < D; ++d) sw += W[(size_t)d * H + c] + (T)kTinyF32;  // sum_d Wbar (pmm.py:87)
      b -= sw;
    }
    bias[c] = b;
  }
  if (g0 == 0) {
    if (kind == 0) {
      hdr->coef = (double)(T(-1) / T(2) / sigma2[0]);
      hdr->lj_const = -0.5 * (double)D * log(2.0 * 3.14159265358979323846 * (double)sigma2[0]);
    } else {
      hdr->coef = 1.0;
      hdr->lj_const = 0.0;
    }
  }
}

// ------------------------------------------------------------------------------------ lpj
constexpr int kMixTile = 64, kMixK = 16;
template <typename T, int KIND>
__global__ void __launch_bounds__(256)
mixture_lpj_kernel(const T *__restrict__ Y, int64_t ldY, const int64_t *__restrict__ idx, const void *ws,
                   T *__restrict__ lpj_out, int64_t lpj_stride_n, int64_t B, int H, int D) {
  __shared__ T Ys[kMixK][kMixTile + 1];
  __shared__ T Ws[kMixK][kMixTile];
  const unsigned char *p = (const unsigned char *)ws;
  const MixHdr *hdr = (const MixHdr *)p;
  const T *bias = (const T *)(p + sizeof(MixHdr));
  const T *Weff = (const T *)(p + mix_off_w<T>(H));
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t n0 = (int64_t)blockIdx.y * kMixTile;
  const int c0 = blockIdx.x * kMixTile;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
  for (int k0 = 0; k0 < D; k0 += kMixK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = tid + 256 * r;
      const int nl = e / kMixK, kl = e % kMixK;
      const int64_t n = n0 + nl;
      Ys[kl][nl] = (n < B && k0 + kl < D) ? Y[n * ldY + k0 + kl] : T(0);
      const int kw = e / kMixTile, cl = e % kMixTile;
      Ws[kw][cl] = (k0 + kw < D && c0 + cl < H) ? Weff[(size_t)(k0 + kw) * H + c0 + cl] : T(0);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kMixK; ++kk) {
      T y[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) y[i] = Ys[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (KIND == 0) {
            const T t = w[j] - y[i];
            acc[i][j] += t * t;
          } else {
            acc[i][j] += y[i] * w[j];
          }
        }
    }
    __syncthreads();
  }
  const T coef = (T)hdr->coef;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t n = n0 + ty + 16 * i;
    if (n >= B) continue;
    const int64_t row = idx ? idx[n] : n;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx + 16 * j;
      if (c < H) lpj_out[row * lpj_stride_n + c] = (KIND == 0 ? acc[i][j] * coef : acc[i][j]) + bias[c];
    }
  }
}

// ------------------------------------------------------------------------------------ M-step / free energy
// stats: [WpT (H*D) | Wq = pies sums (H) | sigma2 | N | F].  One warp per datapoint.
template <typename T, int KIND>
__global__ void __launch_bounds__(256)
mixture_mstep_kernel(const T *__restrict__ Y, int64_t ldY, const int64_t *__restrict__ idx, int64_t row0,
                     const T *__restrict__ lpj, int64_t lpj_stride_n, const void *ws, double *__restrict__ stats,
                     double *__restrict__ Q_out, double *__restrict__ Yd_out, double *__restrict__ F_only, int64_t B, int H,
                     int D) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const unsigned char *p = (const unsigned char *)ws;
  const MixHdr *hdr = (const MixHdr *)p;
  const T *bias = (const T *)(p + sizeof(MixHdr));
  double *qsum = (double *)smem_raw + (size_t)wib * H;  // per-warp running sum_n q[n, :]
  const bool accumulate = F_only == nullptr;
  const double inv_coef = 1.0 / hdr->coef;
  double acc_sig = 0.0, acc_F = 0.0, acc_N = 0.0;
  if (accumulate)
    for (int c = lane; c < H; c += 32) qsum[c] = 0.0;
  for (int64_t b = (int64_t)blockIdx.x * wpb + wib; b < B; b += (int64_t)gridDim.x * wpb) {
    const int64_t row = idx ? idx[b] : row0 + b;
    const T *l = lpj + row * lpj_stride_n;
    const T *y = Y + b * ldY;
    double mx = -INFINITY;
    for (int c = lane; c < H; c += 32) mx = fmax(mx, (double)l[c]);
    mx = warp_max(mx);
    double se = 0.0;
    for (int c = lane; c < H; c += 32) se += exp((double)l[c] - mx);
    se = warp_sum(se);
    double cn = hdr->lj_const;
    if (KIND == 1) {  // - sum_d lgamma(y_d + 1)   (pmm.py:97)
      double t = 0.0;
      for (int d = lane; d < D; d += 32) t += lgamma((double)y[d] + 1.0);
      cn = -warp_sum(t);
    }
    acc_F += mx + log(se) + cn;
    if (!accumulate) continue;
    acc_N += 1.0;
    const double inv_se = 1.0 / se;
    double part = 0.0;
    for (int c = lane; c < H; c += 32) {
      const double lc = (double)l[c];
      const double q = exp(lc - mx) * inv_se;
      Q_out[b * (int64_t)H + c] = q;
      qsum[c] += q;
      if (KIND == 0) part += q * ((lc - (double)bias[c]) * inv_coef);  // q * ||y - W_c||^2 (gmm.py:124)
    }
    if (KIND == 0) acc_sig += warp_sum(part);
    if (Yd_out)
      for (int d = lane; d < D; d += 32) Yd_out[b * (int64_t)D + d] = (double)y[d];
  }
  if (accumulate) {
    double *Wq = stats + (size_t)H * D;
    __syncwarp();
    for (int c = lane; c < H; c += 32)
      if (qsum[c] != 0.0) atomicAdd(Wq + c, qsum[c]);
    if (lane == 0 && acc_N != 0.0) {
      atomicAdd(Wq + H, acc_sig);
      atomicAdd(Wq + H + 1, acc_N);
      atomicAdd(Wq + H + 2, acc_F);
    }
  } else if (lane == 0 && acc_F != 0.0) {
    atomicAdd(F_only, acc_F);
  }
}

// ------------------------------------------------------------------------------------ host side
static cublasHandle_t g_mix_handles[64];
static int mix_cublas(cudaStream_t st, cublasHandle_t &h) {
  int dev = 0;
  TVO_CUDA(cudaGetDevice(&dev));
  TVO_REQUIRE(dev >= 0 && dev < 64, "cublas: device index %d out of range", dev);
  if (!g_mix_handles[dev] && cublasCreate(&g_mix_handles[dev]) != CUBLAS_STATUS_SUCCESS)
    return set_error(TVO_ECUDA, "cublasCreate failed");
  h = g_mix_handles[dev];
  if (cublasSetStream(h, st) != CUBLAS_STATUS_SUCCESS) return set_error(TVO_ECUDA, "cublasSetStream failed");
  return TVO_OK;
}

template <typename T>
int mixture_prepare(int kind, const T *W, const T *pies, const T *sigma2, int H, int D, void *ws, void *stream) {
  TVO_REQUIRE(W && pies && ws && (kind == 1 || sigma2), "mixture_prepare: null pointer");
  TVO_REQUIRE((kind == 0 || kind == 1) && H > 0 && D > 0, "mixture_prepare: bad arguments (kind=%d H=%d D=%d)", kind, H, D);
  const size_t total = (size_t)D * H;
  mixture_prepare_kernel<T><<<(int)max((size_t)1, min((size_t)4 * sm_count(), (total + 255) / 256)), 256, 0,
                              (cudaStream_t)stream>>>(kind, W, pies, sigma2, H, D, ws);
  TVO_AFTER_LAUNCH("mixture_prepare");
  return TVO_OK;
}

template <typename T>
int mixture_lpj(int kind, const T *Y, int64_t ldY, const int64_t *idx, const void *ws, T *lpj_out, int64_t lpj_stride_n,
                int64_t B, int H, int D, void *stream) {
  TVO_REQUIRE(Y && ws && lpj_out, "mixture_lpj: null pointer");
  TVO_REQUIRE((kind == 0 || kind == 1) && B >= 0 && H > 0 && D > 0 && ldY >= D && lpj_stride_n >= H,
              "mixture_lpj: bad arguments");
  if (B == 0) return TVO_OK;
  const int64_t ny = (B + kMixTile - 1) / kMixTile;
  TVO_REQUIRE(ny <= 65535 * 64, "mixture_lpj: batch too large for one call");
  for (int64_t y0 = 0; y0 < ny; y0 += 65535) {  // gridDim.y limit
    const int64_t gy = min((int64_t)65535, ny - y0), off = y0 * kMixTile;
    const dim3 grid((H + kMixTile - 1) / kMixTile, (unsigned)gy);
    if (kind == 0)
      mixture_lpj_kernel<T, 0><<<grid, 256, 0, (cudaStream_t)stream>>>(Y + off * ldY, ldY, idx ? idx + off : nullptr, ws,
                                                                       idx ? lpj_out : lpj_out + off * lpj_stride_n,
                                                                       lpj_stride_n, min(B - off, gy * kMixTile), H, D);
    else
      mixture_lpj_kernel<T, 1><<<grid, 256, 0, (cudaStream_t)stream>>>(Y + off * ldY, ldY, idx ? idx + off : nullptr, ws,
                                                                       idx ? lpj_out : lpj_out + off * lpj_stride_n,
                                                                       lpj_stride_n, min(B - off, gy * kMixTile), H, D);
    TVO_AFTER_LAUNCH("mixture_lpj");
  }
  return TVO_OK;
}

template <typename T>
int mixture_mstep(int kind, const T *Y, int64_t ldY, const int64_t *idx, const T *lpj, int64_t lpj_stride_n, const void *ws,
                  double *stats, void *scratch, int64_t scratch_bytes, double *F_only, int64_t B, int H, int D,
                  void *stream) {
  TVO_REQUIRE(Y && lpj && ws && (F_only || (stats && scratch)), "mixture_mstep: null pointer");
  TVO_REQUIRE((kind == 0 || kind == 1) && B >= 0 && H > 0 && D > 0 && ldY >= D && ldY < (1ll << 31), "mixture_mstep: bad sizes");
  TVO_REQUIRE((size_t)H * 8 * 8 <= 200 * 1024, "mixture_mstep: H=%d too large", H);
  if (B == 0) return TVO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool f64 = sizeof(T) == 8;
  const int64_t per = (int64_t)H * 8 + (f64 ? 0 : (int64_t)D * 8);
  const int64_t cap = F_only ? B : scratch_bytes / per;
  TVO_REQUIRE(cap >= 1, "mixture_mstep: scratch of %lld B holds no datapoint (need %lld B each)", (long long)scratch_bytes,
              (long long)per);
  cublasHandle_t hb = nullptr;
  if (!F_only) {
    int rc = mix_cublas(st, hb);
    if (rc) return rc;
  }
  const size_t smem = (size_t)8 * H * 8;
  auto kern = kind == 0 ? mixture_mstep_kernel<T, 0> : mixture_mstep_kernel<T, 1>;
  if (smem > 40 * 1024) TVO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const double one = 1.0;
  for (int64_t c0 = 0; c0 < B; c0 += cap) {
    const int Bc = (int)min(cap, B - c0);
    double *Q = (double *)scratch;
    double *Yd = (f64 || F_only) ? nullptr : Q + (size_t)Bc * H;
    const int grid = (int)max((int64_t)1, min((int64_t)(Bc + 7) / 8, (int64_t)sm_count() * 4));
    kern<<<grid, 256, smem, st>>>(Y + c0 * ldY, ldY, idx ? idx + c0 : nullptr, c0, lpj, lpj_stride_n, ws, stats, Q, Yd, F_only,
                                  Bc, H, D);
    TVO_AFTER_LAUNCH("mixture_mstep");
    if (F_only) continue;
    // WpT (H x D, row-major) += Q^T (H x Bc) * Y (Bc x D)     (gmm.py:122,127 / pmm.py:107,111)
    const double *Yp = f64 ? (const double *)(Y + c0 * ldY) : Yd;
    const int ldy = f64 ? (int)ldY : D;
    cublasStatus_t s = cublasDgemm(hb, CUBLAS_OP_N, CUBLAS_OP_T, D, H, Bc, &one, Yp, ldy, Q, H, &one, stats, D);
    if (s != CUBLAS_STATUS_SUCCESS) return set_error(TVO_ECUDA, "cublasDgemm failed (%d)", (int)s);
    count_launch();
  }
  return TVO_OK;
}

}  // namespace tvo

using namespace tvo;

extern "C" {
int64_t tvo_mixture_ws_bytes(int32_t H, int32_t D, int32_t is_f64) {
  return (int64_t)(is_f64 ? mix_ws_bytes<double>(H, D) : mix_ws_bytes<float>(H, D));
}
int64_t tvo_mixture_stats_len(int32_t H, int32_t D) { return (int64_t)H * D + H + 3; }
int64_t tvo_mixture_scratch_bytes(int64_t B, int32_t H, int32_t D, int32_t is_f64) {
  return B * ((int64_t)H * 8 + (is_f64 ? 0 : (int64_t)D * 8));
}
#define TVO_MIX_WRAP(sfx, T)                                                                                              \
  int tvo_mixture_prepare_##sfx(int32_t kind, const T *W, const T *pies, const T *sigma2, int32_t H, int32_t D, void *ws, \
                                void *stream) {                                                                           \
    return mixture_prepare<T>(kind, W, pies, sigma2, H, D, ws, stream);                                                   \
  }                                                                                                                       \
  int tvo_mixture_lpj_##sfx(int32_t kind, const T *Y, int64_t ldY, const int64_t *idx, const void *ws, T *lpj_out,        \
                            int64_t lpj_stride_n, int64_t B, int32_t H, int32_t D, void *stream) {                        \
    return mixture_lpj<T>(kind, Y, ldY, idx, ws, lpj_out, lpj_stride_n, B, H, D, stream);                                 \
  }                                                                                                                       \
  int tvo_mixture_mstep_##sfx(int32_t kind, const T *Y, int64_t ldY, const int64_t *idx, const T *lpj,                    \
                              int64_t lpj_stride_n, const void *ws, double *stats, void *scratch, int64_t scratch_bytes,  \
                              int64_t B, int32_t H, int32_t D, void *stream) {                                            \
    return mixture_mstep<T>(kind, Y, ldY, idx, lpj, lpj_stride_n, ws, stats, scratch, scratch_bytes, nullptr, B, H, D,    \
                            stream);                                                                                      \
  }                                                                                                                       \
  int tvo_mixture_free_energy_##sfx(int32_t kind, const T *Y, int64_t ldY, const int64_t *idx, const T *lpj,              \
                                    int64_t lpj_stride_n, const void *ws, int64_t B, int32_t H, int32_t D, double *F_out, \
                                    void *stream) {                                                                       \
    TVO_REQUIRE(F_out, "mixture_free_energy: null pointer");                                                              \
    return mixture_mstep<T>(kind, Y, ldY, idx, lpj, lpj_stride_n, ws, nullptr, nullptr, 0, F_out, B, H, D, stream);       \
  }
TVO_MIX_WRAP(f64, double)
TVO_MIX_WRAP(f32, float)
}
