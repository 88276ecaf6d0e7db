// tvo_b200 — the parts of the TVAE decoder that touch the bit-packed states
// (tvo/models/tvae.py:326-341 Gaussian, 584-604 Bernoulli, 434-467 / 672-701 forward).
//
// The reference casts the uint8 states (N, S, H0) to the model precision and runs a dense MLP on them.
// With packed states the FIRST layer is a sum of rows of W_0 over the set bits (no (N, S, H0) float tensor
// ever exists: 65 GB at N = 1M, S = 64, H0 = 256 in fp32), its weight gradient is the matching scatter, and
// the LAST step (log-pseudo-joint from the decoder output) reads the prior term straight from the bits.
// Everything in between (activations, the dense hidden layers, autograd, Adam) stays torch: plain library
// GEMMs on (N*S, H_l) activations.
#include "common.cuh"

namespace tvo {

constexpr int kTvaeCols = 4;  // columns per lane and pass

// grid for a grid-stride kernel: enough CTAs for `items` at `per_block` each, at most `blocks_per_sm` waves-worth
static int grid_for(int64_t items, int per_block, int blocks_per_sm) {
  const int64_t need = (items + per_block - 1) / per_block;
  const int64_t cap = (int64_t)sm_count() * blocks_per_sm;
  return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

// out[r, j] = b0[j] + sum_{h in state r} W0[h, j], ascending h.  One warp per state, lanes over j.
template <typename T>
__global__ void __launch_bounds__(256)
tvae_layer0_kernel(const uint64_t *__restrict__ states, int64_t n_rows, int W64, const T *__restrict__ W0,
                   const T *__restrict__ b0, int H1, T *__restrict__ out) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * wpb + wib; r < n_rows; r += (int64_t)gridDim.x * wpb) {
    const uint64_t *st = states + r * W64;
    for (int j0 = 0; j0 < H1; j0 += 32 * kTvaeCols) {
      T acc[kTvaeCols];
#pragma unroll
      for (int c = 0; c < kTvaeCols; ++c) {
        const int j = j0 + 32 * c + lane;
        acc[c] = (b0 && j < H1) ? __ldg(b0 + j) : T(0);
      }
      for (int w = 0; w < W64; ++w) {
        uint64_t m = st[w];  // warp-uniform: no divergence in the bit walk
        while (m) {
          const int h = (w << 6) + __ffsll((long long)m) - 1;
          m &= m - 1;
          const T *row = W0 + (size_t)h * H1 + j0 + lane;
#pragma unroll
          for (int c = 0; c < kTvaeCols; ++c)
            if (j0 + 32 * c + lane < H1) acc[c] += __ldg(row + 32 * c);
        }
      }
#pragma unroll
      for (int c = 0; c < kTvaeCols; ++c) {
        const int j = j0 + 32 * c + lane;
        if (j < H1) out[r * (int64_t)H1 + j] = acc[c];
      }
    }
  }
}

// gW0[h, j] += sum_{r : h in state r} g[r, j]  (the transpose of the above; d/dW_0 of any loss whose gradient
// w.r.t. the first pre-activation is g).  The scatter goes to global memory as fp32 / fp64 REDs, one per (state,
// active unit, column): the states of a batch share few units, so there is little to combine on chip, and TVAE
// mini-batches are small (tens of datapoints, i.e. a few thousand states per call).
template <typename T>
__global__ void __launch_bounds__(256)
tvae_layer0_grad_kernel(const uint64_t *__restrict__ states, int64_t n_rows, int W64, const T *__restrict__ g, int H1,
                        T *__restrict__ gW0) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * wpb + wib; r < n_rows; r += (int64_t)gridDim.x * wpb) {
    const uint64_t *st = states + r * W64;
    for (int j0 = 0; j0 < H1; j0 += 32 * kTvaeCols) {
      T gv[kTvaeCols];
#pragma unroll
      for (int c = 0; c < kTvaeCols; ++c) {
        const int j = j0 + 32 * c + lane;
        gv[c] = j < H1 ? g[r * (int64_t)H1 + j] : T(0);
      }
      for (int w = 0; w < W64; ++w) {
        uint64_t m = st[w];
        while (m) {
          const int h = (w << 6) + __ffsll((long long)m) - 1;
          m &= m - 1;
          T *row = gW0 + (size_t)h * H1 + j0 + lane;
#pragma unroll
          for (int c = 0; c < kTvaeCols; ++c)
            if (j0 + 32 * c + lane < H1 && gv[c] != T(0)) atomicAdd(row + 32 * c, gv[c]);
        }
      }
    }
  }
}

// lpj[b, s] = sum_{h in state (b,s)} prior[h] - s1,  prior[h] = log(pi_h / (1 - pi_h)) and
//   KIND 0 (Gaussian, tvae.py:335-338):  s1 = nansum_d (y_bd - out_bsd)^2 / (2 sigma2)
//   KIND 1 (Bernoulli, tvae.py:594-601): s1 = nansum_d BCE(out_bsd, y_bd), log terms clamped at -100 like
//                                        torch.nn.functional.binary_cross_entropy
// One warp per state, lanes over d.  states / lpj rows are addressed as (b, s) with explicit strides.
template <typename T, int KIND>
__global__ void __launch_bounds__(256)
tvae_lpj_kernel(const T *__restrict__ Y, int64_t ldY, const T *__restrict__ out, const uint64_t *__restrict__ states,
                int64_t st_stride_b, const T *__restrict__ prior, const T *__restrict__ sigma2, T *__restrict__ lpj,
                int64_t lpj_stride_b, int64_t B, int S, int W64, int D) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int64_t n_rows = B * S;
  for (int64_t r = (int64_t)blockIdx.x * wpb + wib; r < n_rows; r += (int64_t)gridDim.x * wpb) {
    const int64_t b = r / S;
    const int s = (int)(r - b * S);
    const T *y = Y + b * ldY;
    const T *o = out + r * (int64_t)D;
    T acc = T(0);
    for (int d = lane; d < D; d += 32) {
      const T yv = y[d], ov = o[d];
      T t;
      if (KIND == 0) {
        const T df = yv - ov;
        t = df * df;
      } else {
        const T l1 = max(log(ov), T(-100)), l0 = max(log(T(1) - ov), T(-100));
        t = -(yv * l1 + (T(1) - yv) * l0);
      }
      if (t == t) acc += t;  // nansum
    }
    acc = warp_sum(acc);
    const uint64_t *st = states + b * st_stride_b + (int64_t)s * W64;
    T pr = T(0);
    for (int w = lane; w < W64; w += 32) {
      uint64_t m = st[w];
      while (m) {
        const int h = (w << 6) + __ffsll((long long)m) - 1;
        m &= m - 1;
        pr += __ldg(prior + h);
      }
    }
    pr = warp_sum(pr);
    if (lane == 0) {
      const T s1 = KIND == 0 ? acc / (T(2) * sigma2[0]) : acc;
      lpj[b * lpj_stride_b + s] = pr - s1;
    }
  }
}

template <typename T>
static int tvae_layer0(const uint64_t *states, int64_t n_rows, int H0, const T *W0, const T *b0, int H1, T *out,
                       void *stream) {
  TVO_REQUIRE(states && W0 && out, "tvae_layer0: null pointer");
  TVO_REQUIRE(n_rows >= 0 && H0 > 0 && H1 > 0, "tvae_layer0: bad sizes");
  if (n_rows == 0) return TVO_OK;
  tvae_layer0_kernel<T><<<grid_for(n_rows, 8, 8), 256, 0, (cudaStream_t)stream>>>(states, n_rows, n_words(H0), W0, b0, H1,
                                                                                  out);
  TVO_AFTER_LAUNCH("tvae_layer0");
  return TVO_OK;
}

template <typename T>
static int tvae_layer0_grad(const uint64_t *states, int64_t n_rows, int H0, const T *g, int H1, T *gW0, void *stream) {
  TVO_REQUIRE(states && g && gW0, "tvae_layer0_grad: null pointer");
  TVO_REQUIRE(n_rows >= 0 && H0 > 0 && H1 > 0, "tvae_layer0_grad: bad sizes");
  if (n_rows == 0) return TVO_OK;
  tvae_layer0_grad_kernel<T><<<grid_for(n_rows, 8, 8), 256, 0, (cudaStream_t)stream>>>(states, n_rows, n_words(H0), g, H1,
                                                                                       gW0);
  TVO_AFTER_LAUNCH("tvae_layer0_grad");
  return TVO_OK;
}

template <typename T>
static int tvae_lpj(int kind, const T *Y, int64_t ldY, const T *out, const uint64_t *states, int64_t st_stride_b,
                    const T *prior, const T *sigma2, T *lpj, int64_t lpj_stride_b, int64_t B, int S, int H0, int D,
                    void *stream) {
  TVO_REQUIRE(Y && out && states && prior && lpj, "tvae_lpj: null pointer");
  TVO_REQUIRE(kind == 0 || kind == 1, "tvae_lpj: kind must be 0 (Gaussian) or 1 (Bernoulli), got %d", kind);
  TVO_REQUIRE(kind == 1 || sigma2, "tvae_lpj: the Gaussian form needs sigma2");
  TVO_REQUIRE(B >= 0 && S > 0 && H0 > 0 && D > 0 && ldY >= D, "tvae_lpj: bad sizes");
  if (B == 0) return TVO_OK;
  const int grid = grid_for(B * S, 8, 8);
  if (kind == 0)
    tvae_lpj_kernel<T, 0><<<grid, 256, 0, (cudaStream_t)stream>>>(Y, ldY, out, states, st_stride_b, prior, sigma2, lpj,
                                                                  lpj_stride_b, B, S, n_words(H0), D);
  else
    tvae_lpj_kernel<T, 1><<<grid, 256, 0, (cudaStream_t)stream>>>(Y, ldY, out, states, st_stride_b, prior, sigma2, lpj,
                                                                  lpj_stride_b, B, S, n_words(H0), D);
  TVO_AFTER_LAUNCH("tvae_lpj");
  return TVO_OK;
}

}  // namespace tvo

using namespace tvo;

extern "C" {
int tvo_tvae_layer0_f64(const uint64_t *states, int64_t n_states, int32_t H0, const double *W0, const double *b0,
                        int32_t H1, double *out, void *stream) {
  return tvae_layer0<double>(states, n_states, H0, W0, b0, H1, out, stream);
}
int tvo_tvae_layer0_f32(const uint64_t *states, int64_t n_states, int32_t H0, const float *W0, const float *b0,
                        int32_t H1, float *out, void *stream) {
  return tvae_layer0<float>(states, n_states, H0, W0, b0, H1, out, stream);
}
int tvo_tvae_layer0_grad_f64(const uint64_t *states, int64_t n_states, int32_t H0, const double *g, int32_t H1,
                             double *gW0, void *stream) {
  return tvae_layer0_grad<double>(states, n_states, H0, g, H1, gW0, stream);
}
int tvo_tvae_layer0_grad_f32(const uint64_t *states, int64_t n_states, int32_t H0, const float *g, int32_t H1,
                             float *gW0, void *stream) {
  return tvae_layer0_grad<float>(states, n_states, H0, g, H1, gW0, stream);
}
int tvo_tvae_lpj_f64(int32_t kind, const double *Y, int64_t ldY, const double *out, const uint64_t *states,
                     int64_t st_stride_b, const double *prior, const double *sigma2, double *lpj, int64_t lpj_stride_b,
                     int64_t B, int32_t S, int32_t H0, int32_t D, void *stream) {
  return tvae_lpj<double>(kind, Y, ldY, out, states, st_stride_b, prior, sigma2, lpj, lpj_stride_b, B, S, H0, D, stream);
}
int tvo_tvae_lpj_f32(int32_t kind, const float *Y, int64_t ldY, const float *out, const uint64_t *states,
                     int64_t st_stride_b, const float *prior, const float *sigma2, float *lpj, int64_t lpj_stride_b,
                     int64_t B, int32_t S, int32_t H0, int32_t D, void *stream) {
  return tvae_lpj<float>(kind, Y, ldY, out, states, st_stride_b, prior, sigma2, lpj, lpj_stride_b, B, S, H0, D, stream);
}
}
