// tvo_b200 — NoisyOR kernels (tvo/models/noisyor.py:67-162).
//
// lpj (noisyor.py:67-106):  P_d = clamp(prod_{h in s}(1 - W_dh), eps, 1-eps),
//     lpj = sum_{h in s} log(pi_h/(1-pi_h)) + sum_d [ y_d log(1/P_d - 1) + log P_d ],
//     all-zero state: 0 if y is all zero, else -1e30.
// The product only runs over the SET bits of the packed state (1 - W*0 == 1 exactly), lanes over d.
// M-step (noisyor.py:108-138): Btilde / Ctilde only receive contributions from active h, so the
// (N,S,D,H) temporaries of the reference never exist; one CTA owns private [h][d] accumulators in
// shared memory (threads over d, no atomics) and flushes them once at the end.
#include "estep_fused.cuh"

namespace tvo {

constexpr double kNorEps = 1e-7;  // NoisyOR.eps (noisyor.py:17)

struct NorThetaHdr {
  double prior_sum;  // sum_h log(1 - pi_h)   (noisyor.py:162)
  double pad[7];
};
__host__ __device__ inline int nor_dpad(int D) { return pad32(D); }
template <typename T>
__host__ __device__ inline size_t nor_w_off(int H) {
  return sizeof(NorThetaHdr) + align16((size_t)H * sizeof(T));
}
template <typename T>
__host__ __device__ inline size_t nor_ws_bytes(int H, int D) {
  return nor_w_off<T>(H) + (size_t)H * nor_dpad(D) * sizeof(T);
}
template <typename T>
struct NorTheta {
  const NorThetaHdr *hdr;
  const T *prior;  // H
  const T *W1mT;   // H x Dpad: 1 - W_dh, padded with 1
  int Dpad;
};
template <typename T>
__host__ __device__ inline NorTheta<T> nor_theta_view(const void *ws, int H, int D) {
  NorTheta<T> t;
  const unsigned char *p = (const unsigned char *)ws;
  t.hdr = (const NorThetaHdr *)p;
  t.prior = (const T *)(p + sizeof(NorThetaHdr));
  t.W1mT = (const T *)(p + nor_w_off<T>(H));
  t.Dpad = nor_dpad(D);
  return t;
}

template <typename T>
__global__ void nor_prepare_kernel(const T *__restrict__ W, const T *__restrict__ pies, int H, int D, void *ws) {
  unsigned char *p = (unsigned char *)ws;
  NorThetaHdr *hdr = (NorThetaHdr *)p;
  T *prior = (T *)(p + sizeof(NorThetaHdr));
  T *W1mT = (T *)(p + nor_w_off<T>(H));
  const int Dpad = nor_dpad(D);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)H * Dpad; i += (size_t)gridDim.x * blockDim.x) {
    const int h = (int)(i / Dpad), d = (int)(i % Dpad);
    W1mT[i] = d < D ? T(1) - W[(size_t)d * H + h] : T(1);
  }
  for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < H; h += gridDim.x * blockDim.x) {
    const T pi = pies[h];
    prior[h] = log(pi / (T(1) - pi));
  }
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    double s = 0.0;
    for (int h = threadIdx.x; h < H; h += 32) s += (double)log(T(1) - pies[h]);
    s = warp_sum(s);
    if (threadIdx.x == 0) hdr->prior_sum = s;
  }
}

// eight consecutive values, 32-byte aligned
__device__ __forceinline__ void ldg8(const float *p, float (&v)[8]) {
  const float4 t = __ldg(reinterpret_cast<const float4 *>(p)), u = __ldg(reinterpret_cast<const float4 *>(p) + 1);
  v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w, v[4] = u.x, v[5] = u.y, v[6] = u.z, v[7] = u.w;
}
__device__ __forceinline__ void ldg8(const double *p, double (&v)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double2 t = __ldg(reinterpret_cast<const double2 *>(p) + i);
    v[2 * i] = t.x, v[2 * i + 1] = t.y;
  }
}

// lpj of TWO states per pass: lanes 0-15 score state s, lanes 16-31 state s + 1.  Within a group of 128
// observables half-lane hl owns the eight CONSECUTIVE columns 8 hl .. 8 hl + 7, so one active unit costs two (fp32)
// vector loads and eight multiplies per lane for two states, and the reduction over d is four shuffles.
// (One warp per state with one column per lane and chunk spent ~225 warp instructions per state, 74 % of the fused
// E-step; the kernel issues at 75 % of peak, so instructions are what it is made of.)
// ybits: bit 8 g + q of half-lane hl = y[128 g + 8 hl + q] != 0 (same in both halves).
// sum_d log t_d is taken as one log per four columns: every term is clamped to [1e-7, 1 - 1e-7], so a product of
// four stays above 1e-28 (eight would underflow fp32).
template <typename T>
__device__ inline void nor_score_pair(const uint64_t *states, int s, int count, int W64, const NorTheta<T> &th,
                                      const uint32_t *ybits, bool any_y, int D, T *out, int lane) {
  const int half = lane >> 4, hl = lane & 15;
  const int ms = s + half;
  const bool valid = ms < count;
  const uint64_t *st = states + (size_t)(valid ? ms : s) * W64;
  const T lo = (T)kNorEps, hi = T(1) - (T)kNorEps;
  uint64_t any = 0;
  T acc = T(0), pr = T(0);
  for (int c0 = 0, g = 0; c0 < th.Dpad; c0 += 128, ++g) {
    T P[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) P[q] = T(1);
    const bool colok = c0 + 8 * hl < th.Dpad;  // Dpad is a multiple of 32: the eight columns are inside the row
    for (int w = 0; w < W64; ++w) {
      uint64_t word = st[w];
      if (c0 == 0) any |= word;
      while (word) {  // per-lane trip count: the halves walk different states
        const int h = (w << 6) + __ffsll((long long)word) - 1;
        word &= word - 1;
        if (colok) {
          T v[8];
          ldg8(th.W1mT + (size_t)h * th.Dpad + c0 + 8 * hl, v);
#pragma unroll
          for (int q = 0; q < 8; ++q) P[q] *= v[q];
        }
        if (c0 == 0) pr += __ldg(th.prior + h);
      }
    }
    // static indices only: the four words stay in registers
    const uint32_t yw = g < 4 ? ybits[0] : (g < 8 ? ybits[1] : (g < 12 ? ybits[2] : ybits[3]));
    const uint32_t yb = (yw >> (8 * (g & 3))) & 0xFFu;
    T prod0 = T(1), prod1 = T(1);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (c0 + 8 * hl + q < D) {
        const T Pc = P[q] < lo ? lo : (P[q] > hi ? hi : P[q]);  // to.clamp (noisyor.py:89)
        // y log(1/P - 1) + log P  ==  y ? log(1 - P) : log P   (noisyor.py:91-97)
        const T t = ((yb >> q) & 1u) ? T(1) - Pc : Pc;
        if (q < 4)
          prod0 *= t;
        else
          prod1 *= t;
      }
    }
    acc += log(prod0) + log(prod1);
  }
  __syncwarp();
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);  // stays inside the 16-lane half
  if (hl == 0 && valid) out[ms] = any ? pr + acc : (any_y ? T(-1e30) : T(0));  // all-zero state: noisyor.py:83-86,102
}

// model adapter for the generic fused E-step kernel
template <typename T>
struct NorModel {
  struct Params {
    const uint8_t *Y;
    int64_t ldY;
    const void *ws;
    int H, D;
  };
  static size_t smem_bytes(const Params &) { return 0; }
  struct Ctx {
    uint32_t ybits[4];  // D <= 2048: 16 groups of 128 observables x 8 columns per half-lane
    bool any_y;
    __device__ void load(const Params &p, int64_t b, unsigned char *, int lane) {
      const uint8_t *y = p.Y + b * p.ldY;
      const int hl = lane & 15;
      ybits[0] = ybits[1] = ybits[2] = ybits[3] = 0;
      for (int c0 = 0, g = 0; c0 < p.D; c0 += 128, ++g)
        for (int q = 0; q < 8; ++q) {
          const int d = c0 + 8 * hl + q;
          if (d < p.D && y[d]) {
            const uint32_t bit = 1u << (8 * (g & 3) + q);
            if (g < 4)
              ybits[0] |= bit;
            else if (g < 8)
              ybits[1] |= bit;
            else if (g < 12)
              ybits[2] |= bit;
            else
              ybits[3] |= bit;
          }
        }
      any_y = __any_sync(FULL, (ybits[0] | ybits[1] | ybits[2] | ybits[3]) != 0);
    }
    __device__ void score(const Params &p, const uint64_t *states, int count, int W64, T *out, int lane) {
      const NorTheta<T> th = nor_theta_view<T>(p.ws, p.H, p.D);
      for (int s = 0; s < count; s += 2) nor_score_pair<T>(states, s, count, W64, th, ybits, any_y, p.D, out, lane);
      __syncwarp();
    }
  };
};

// log_pseudo_joint for explicit state sets; work item = (datapoint, chunk of CH states)
template <typename T>
__global__ void __launch_bounds__(256)
nor_lpj_kernel(const uint8_t *__restrict__ Y, int64_t ldY, const uint64_t *__restrict__ K, int64_t k_stride_n,
               const int64_t *__restrict__ idx, const void *ws, T *__restrict__ lpj_out, int64_t lpj_stride_n, int64_t B,
               int S, int H, int D, int CH) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int W64 = n_words(H);
  const int nchunk = (S + CH - 1) / CH;
  typename NorModel<T>::Params p{Y, ldY, ws, H, D};
  typename NorModel<T>::Ctx ctx;
  for (int64_t task = (int64_t)blockIdx.x * wpb + wib; task < B * nchunk; task += (int64_t)gridDim.x * wpb) {
    const int64_t b = task / nchunk;
    const int c = (int)(task % nchunk);
    const int64_t row = idx ? idx[b] : b;
    ctx.load(p, b, nullptr, lane);
    const int s0 = c * CH, n = min(S, s0 + CH) - s0;
    ctx.score(p, K + row * k_stride_n + (size_t)s0 * W64, n, W64, lpj_out + row * lpj_stride_n + s0, lane);
  }
}

// F_out += sum_b logsumexp_s lpj[b, s]   (model_protocols.py:191-193; the model constant is added by the caller)
template <typename T>
__global__ void __launch_bounds__(256)
logsumexp_sum_kernel(const T *__restrict__ lpj, int64_t stride_n, const int64_t *__restrict__ idx, int64_t B, int S,
                     double *out) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  double acc = 0.0;
  for (int64_t b = (int64_t)blockIdx.x * wpb + wib; b < B; b += (int64_t)gridDim.x * wpb) {
    const T *l = lpj + (idx ? idx[b] : b) * stride_n;
    double mx = -INFINITY;
    for (int s = lane; s < S; s += 32) mx = fmax(mx, (double)l[s]);
    mx = warp_max(mx);
    if (!(mx > -INFINITY && mx < INFINITY)) mx = mx != mx ? mx : 0.0;
    double se = 0.0;
    for (int s = lane; s < S; s += 32) se += exp((double)l[s] - mx);
    acc += mx + log(warp_sum(se));
  }
  __shared__ double part[8];
  if (lane == 0) part[wib] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < wpb; ++w) s += part[w];
    atomicAdd(out, s);
  }
}

// ------------------------------------------------------------------------------------ M-step
// stats: [BtT (H*D) | CtT (H*D) | new_pi (H) | N | F]; the CTA accumulates the h-tile [h0, h0+Ht).
template <typename T>
__global__ void __launch_bounds__(256)
nor_mstep_kernel(const uint8_t *__restrict__ Y, int64_t ldY, const int64_t *__restrict__ idx,
                 const uint64_t *__restrict__ K, int64_t k_stride_n, const T *__restrict__ lpj, int64_t lpj_stride_n,
                 const void *ws, double *__restrict__ stats, int64_t B, int S, int H, int D, int Ht) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const NorTheta<T> th = nor_theta_view<T>(ws, H, D);
  const int W64 = n_words(H), Dpad = th.Dpad;
  const int h0 = blockIdx.y * Ht, h1 = min(H, h0 + Ht);
  double *Bt = (double *)smem_raw;              // Ht x Dpad
  double *Ct = Bt + (size_t)Ht * Dpad;          // Ht x Dpad
  double *pi_acc = Ct + (size_t)Ht * Dpad;      // Ht
  double *q = pi_acc + Ht;                      // S
  uint64_t *Ksm = (uint64_t *)(q + S);          // S x W64
  __shared__ double red[2];
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
  for (int i = tid; i < 2 * Ht * Dpad + Ht; i += nt) Bt[i] = 0.0;
  double accF = 0.0, accN = 0.0;
  const T eps = (T)kNorEps;
  __syncthreads();
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const int64_t row = idx ? idx[b] : b;
    const T *l = lpj + row * lpj_stride_n;
    for (int t = tid; t < S * W64; t += nt) Ksm[t] = K[row * k_stride_n + t];
    if (tid < 32) {  // softmax weights by warp 0 (lpj2pjc, _utils.py:182-191)
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
      if (lane == 0) {
        red[0] = se;
        red[1] = mx + log(se);
      }
    }
    __syncthreads();
    const double se = red[0];
    if (tid == 0 && blockIdx.y == 0) {
      accF += red[1] + th.hdr->prior_sum;
      accN += 1.0;
    }
    // threads over d: Btilde, Ctilde (noisyor.py:121-132)
    for (int d = tid; d < D; d += nt) {
      const T ym1 = (T)Y[b * ldY + d] - T(1);
      for (int s = 0; s < S; ++s) {
        const double qs = q[s] / se;
        const uint64_t *st = Ksm + (size_t)s * W64;
        T Q = T(1);
        for (int w = 0; w < W64; ++w) {
          uint64_t word = st[w];
          while (word) {
            const int h = (w << 6) + __ffsll((long long)word) - 1;
            word &= word - 1;
            Q *= __ldg(th.W1mT + (size_t)h * Dpad + d);
          }
        }
        for (int w = 0; w < W64; ++w) {
          uint64_t word = st[w];
          while (word) {
            const int h = (w << 6) + __ffsll((long long)word) - 1;
            word &= word - 1;
            if (h < h0 || h >= h1) continue;
            const T Ws = __ldg(th.W1mT + (size_t)h * Dpad + d);
            const T Bv = T(1) / (Ws * (T(1) - Q) + eps);  // K / (Ws (1 - prod) + eps)
            const T Cv = Bv * Q / Ws;
            Bt[(size_t)(h - h0) * Dpad + d] += qs * (double)(Bv * ym1);
            Ct[(size_t)(h - h0) * Dpad + d] += qs * (double)Cv;
          }
        }
      }
    }
    // threads over h: new_pi += <s_h>  (noisyor.py:117)
    for (int h = h0 + tid; h < h1; h += nt) {
      double e = 0.0;
      for (int s = 0; s < S; ++s)
        if ((Ksm[(size_t)s * W64 + (h >> 6)] >> (h & 63)) & 1ull) e += q[s] / se;
      pi_acc[h - h0] += e;
    }
    __syncthreads();
  }
  const size_t HD = (size_t)H * D;
  for (int i = tid; i < (h1 - h0) * D; i += nt) {
    const int hh = i / D, d = i % D;
    const double bv = Bt[(size_t)hh * Dpad + d], cv = Ct[(size_t)hh * Dpad + d];
    if (bv != 0.0) atomicAdd(stats + (size_t)(h0 + hh) * D + d, bv);
    if (cv != 0.0) atomicAdd(stats + HD + (size_t)(h0 + hh) * D + d, cv);
  }
  for (int h = h0 + tid; h < h1; h += nt)
    if (pi_acc[h - h0] != 0.0) atomicAdd(stats + 2 * HD + h, pi_acc[h - h0]);
  if (tid == 0 && accN != 0.0) {
    atomicAdd(stats + 2 * HD + H, accN);
    atomicAdd(stats + 2 * HD + H + 1, accF);
  }
}

// M-step, register-stationary form (Dpad in {32,64,128,256}).  A CTA of 512 threads owns the h-tile
// [h0, h0 + Ht), Ht = 8 * (512 / Dpad): thread (hg, d) keeps the Btilde / Ctilde accumulators of its 8
// consecutive h and its column d in REGISTERS together with 1 - W_dh and its reciprocal, so the
// datapoint loop has no atomics and no accumulator traffic at all.  Per datapoint:
//   phase 1: Q_sd = prod_{h in s} (1 - W_dh) for all (s, d) into shared memory (threads over (s, d));
//   phase 2: every thread walks the S states; the byte of the state covering its 8 h is warp-uniform,
//            so warps whose h-group is inactive in a state skip it in 4 instructions.
constexpr int kNorHpt = 8;
// a / b of the accumulation terms: IEEE division in fp64; in the stated fp32 mode the 2-ulp MUFU.RCP form
// (the terms carry 1e-4 tolerance and there are 2 S |s| D of them per datapoint)
__device__ __forceinline__ double nor_div(double a, double b) { return a / b; }
__device__ __forceinline__ float nor_div(float a, float b) { return __fdividef(a, b); }
template <typename T>
__global__ void __launch_bounds__(512, 2)
nor_mstep_reg_kernel(const uint8_t *__restrict__ Y, int64_t ldY, const int64_t *__restrict__ idx,
                     const uint64_t *__restrict__ K, int64_t k_stride_n, const T *__restrict__ lpj, int64_t lpj_stride_n,
                     const T *__restrict__ W1mT, const double *__restrict__ prior_sum, double *__restrict__ stats,
                     int64_t B, int S, int H, int D, int W64, int Dpad, int q_off, int k_off) {
  // W64, Dpad and the shared-memory offsets are kernel parameters (constant bank) on purpose: with 64 registers
  // the compiler re-derived them from H, D and S inside the loops (19% of the executed instructions)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int HG = 512 / Dpad, Ht = kNorHpt * HG;
  const int d = tid % Dpad, hg = tid / Dpad;
  const int h0 = blockIdx.y * Ht;          // first h of the tile (multiple of 8)
  const int hb = h0 + hg * kNorHpt;        // first h of this thread
  T *Qsm = (T *)smem_raw;                            // S x Dpad
  double *q = (double *)(smem_raw + q_off);          // S
  uint64_t *Ksm = (uint64_t *)(smem_raw + k_off);    // S x W64
  __shared__ double red[2];
  double accB[kNorHpt], accC[kNorHpt];  // the only long-lived per-thread state (32 registers)
#pragma unroll
  for (int j = 0; j < kNorHpt; ++j) {
    accB[j] = 0.0;
    accC[j] = 0.0;
  }
  const T *Wcol = W1mT + (size_t)min(hb, H - 1) * Dpad + d;  // 1 - W_dh of this thread's h (L1-resident)
  double pi_acc = 0.0, accF = 0.0, accN = 0.0;  // pi: thread tid < Ht owns h0 + tid
  const T eps = (T)kNorEps;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const int64_t row = idx ? idx[b] : b;
    const T *l = lpj + row * lpj_stride_n;
    __syncthreads();  // previous datapoint fully consumed
    for (int t = tid; t < S * W64; t += 512) Ksm[t] = K[row * k_stride_n + t];
    const T ym1 = (d < D ? (T)Y[b * ldY + d] : T(1)) - T(1);  // issued early: consumed in phase 2
    __syncthreads();
    if (tid < 32) {  // softmax weights by warp 0 (lpj2pjc, _utils.py:182-191) while the other warps start phase 1
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
      for (int s = lane; s < S; s += 32) q[s] = q[s] / se;  // q_s, normalised once
      if (lane == 0) red[1] = mx + log(se);
    }
    // phase 1: Q_sd over ALL active h of the state (noisyor.py:123-124), ascending h
    for (int s = hg; s < S; s += HG) {
      const uint64_t *st = Ksm + s * W64;
      T Q = T(1);
      for (int w = 0; w < W64; ++w) {
        uint64_t word = st[w];
        while (word) {
          const int h = (w << 6) + __ffsll((long long)word) - 1;
          word &= word - 1;
          Q *= __ldg(W1mT + (size_t)h * Dpad + d);
        }
      }
      Qsm[s * Dpad + d] = Q;
    }
    __syncthreads();
    if (tid == 0 && blockIdx.y == 0) {
      accF += red[1] + *prior_sum;
      accN += 1.0;
    }
    if (tid < Ht && h0 + tid < H) {  // new_pi += <s_h>  (noisyor.py:117)
      const int h = h0 + tid;
      double e = 0.0;
      for (int s = 0; s < S; ++s)
        if ((Ksm[s * W64 + (h >> 6)] >> (h & 63)) & 1ull) e += q[s];
      pi_acc += e;
    }
    // phase 2: Btilde, Ctilde (noisyor.py:121-132) for this thread's 8 h and column d
    if (hb < H) {
      const int sh = hb & 63, wsel = hb >> 6;
      // a warp shares its h-group (Dpad is a multiple of 32): 32 states are tested at once, one per lane, and
      // only the states with an active unit in the group are visited
      for (int s0 = 0; s0 < S; s0 += 32) {
        const int sl = s0 + lane;
        const uint32_t ml = sl < S ? (uint32_t)(Ksm[sl * W64 + wsel] >> sh) & 0xFFu : 0u;
        unsigned act = __ballot_sync(FULL, ml != 0u);
        while (act) {
          const int src = __ffs((int)act) - 1;
          act &= act - 1;
          const uint32_t m = __shfl_sync(FULL, ml, src);
          const int s = s0 + src;
          const double qs = q[s];
          const T Q = Qsm[s * Dpad + d];
          const T omq = T(1) - Q;
#pragma unroll
          for (int j = 0; j < kNorHpt; ++j) {
            if ((m >> j) & 1u) {
              const T Ws = __ldg(Wcol + j * Dpad);
              const T Bv = nor_div(T(1), Ws * omq + eps);
              const T Cv = nor_div(Bv * Q, Ws);
              accB[j] += qs * (double)(Bv * ym1);
              accC[j] += qs * (double)Cv;
            }
          }
        }
      }
    }
  }
  const size_t HD = (size_t)H * D;
  if (d < D) {
#pragma unroll
    for (int j = 0; j < kNorHpt; ++j) {
      const int h = hb + j;
      if (h < H) {
        if (accB[j] != 0.0) atomicAdd(stats + (size_t)h * D + d, accB[j]);
        if (accC[j] != 0.0) atomicAdd(stats + HD + (size_t)h * D + d, accC[j]);
      }
    }
  }
  if (tid < Ht && h0 + tid < H && pi_acc != 0.0) atomicAdd(stats + 2 * HD + h0 + tid, pi_acc);
  if (tid == 0 && accN != 0.0) {
    atomicAdd(stats + 2 * HD + H, accN);
    atomicAdd(stats + 2 * HD + H + 1, accF);
  }
}

// M-step, slot form (fp32 mode).  The (state, active unit) pairs of a datapoint are what the work is made of
// (S |s| of them, |s| ~ 1-3 of H); the register-stationary kernel above visits them with 8 predicated h per thread
// and one h-tile per CTA (ncu at C4: 69.6 k warp instructions per datapoint, barrier the top stall).  Here a CTA of
// 512 threads holds Btilde / Ctilde for ALL h in shared memory ([h][d], fp32) and is split into slots of Dpad / V
// threads (thread = slot, columns d .. d + V - 1): slot k owns the units h = k (mod slots), so no two warps ever
// touch the same accumulator and the updates are plain shared-memory read-modify-writes.
// Nothing else is shared: every warp computes the 64 softmax weights of the datapoint itself (registers, handed out
// by shuffle), reads the states straight from global memory (one word per lane, L1) and rebuilds Q_sd for the states
// it visits from 1 - W (|s| vector loads; Q_sd == 1 - W_dh for the single-unit states that dominate) — so the
// datapoint loop has NO barrier and the warps drift apart freely (the first slot version staged Q in shared memory
// behind three barriers per datapoint: ncu showed 6.2 warps per issue slot waiting there).  Every kNorFlush
// datapoints the fp32 partial sums are added to the fp64 statistics in global memory (REDs) and cleared, which bounds
// the fp32 accumulation error (a few hundred terms per flush).
// V = 4 (Dpad >= 128): a slot is ONE warp — 4 consecutive columns per lane, vector loads; V = 1 for narrower D.
constexpr int kNorFlush = 256;  // datapoints per CTA between flushes (1024 measured the same: the wait at the flush barrier is the slowest slot, not the flush)
constexpr int kNorQR = 4;  // softmax weights per lane: S <= 128
constexpr int kNorCountSample = 8192;  // datapoints whose states are counted to balance the slots

// counts[h] += number of states (of the first B datapoints of this call) that contain unit h: the work a slot gets per
// unit is exactly this number, so the slot kernel deals the units out by it.
__global__ void __launch_bounds__(256)
nor_unit_counts_kernel(const uint64_t *__restrict__ K, int64_t k_stride_n, const int64_t *__restrict__ idx, int64_t B,
                       int S, int W64, int H, unsigned *__restrict__ counts) {
  extern __shared__ unsigned hist[];
  for (int h = threadIdx.x; h < H; h += blockDim.x) hist[h] = 0;
  __syncthreads();
  const int64_t nwords = B * S * W64;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nwords; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t / ((int64_t)S * W64);
    const int r = (int)(t - b * (int64_t)S * W64);
    const int64_t row = idx ? idx[b] : b;
    uint64_t m = K[row * k_stride_n + r];
    const int w = r % W64;
    while (m) {
      const int h = (w << 6) + __ffsll((long long)m) - 1;
      m &= m - 1;
      atomicAdd(&hist[h], 1u);
    }
  }
  __syncthreads();
  for (int h = threadIdx.x; h < H; h += blockDim.x)
    if (hist[h]) atomicAdd(counts + h, hist[h]);
}
template <int V>
__global__ void __launch_bounds__(512, 2)
nor_mstep_slot_kernel(const uint8_t *__restrict__ Y, int64_t ldY, const int64_t *__restrict__ idx,
                      const uint64_t *__restrict__ K, int64_t k_stride_n, const float *__restrict__ lpj,
                      int64_t lpj_stride_n, const float *__restrict__ W1mT, const double *__restrict__ prior_sum,
                      const unsigned *__restrict__ counts, double *__restrict__ stats, int64_t B, int S, int H, int D,
                      int W64, int Dpad) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int tps = Dpad / V;        // threads per slot (a multiple of 32)
  const int nslots = 512 / tps;    // power of two, <= 16
  const int d = (tid % tps) * V, slot = tid / tps;
  const bool pi_owner = (tid % tps) < 32;  // first warp of the slot keeps its units' share of new_pi
  float *accB = (float *)smem_raw;                       // H x Dpad
  float *accC = accB + (size_t)H * Dpad;                 // H x Dpad
  float *pi_sm = accC + (size_t)H * Dpad;                // H
  uint64_t *masks = (uint64_t *)(pi_sm + ((H + 3) & ~3));  // nslots x W64: the units owned by each slot
  unsigned char *slot_of = (unsigned char *)(masks + (size_t)nslots * W64);  // H
  // Which slot owns which unit.  No barrier means a slot's warp is as slow as its units are popular, and learned
  // units are anything but equally popular: h mod nslots left the warps of the rare units idle at the flush barrier.
  // Longest-processing-time-first on the number of states that contain each unit (counted over a sample of the
  // datapoints by nor_unit_counts_kernel — the per-unit work is exactly that number; the prior pi_h predicted it
  // poorly): units by descending count, each to the least loaded slot.  Any assignment gives the same sums (one owner per
  // accumulator, datapoints in the same order), so this only moves work around.
  for (int h = tid; h < H; h += 512) pi_sm[h] = (float)__ldg(counts + h);  // staged: the greedy pass is one thread
  __syncthreads();
  if (tid == 0) {
    float load[16];
    for (int k = 0; k < nslots; ++k) load[k] = 0.f;
    for (int h = 0; h < H; ++h) slot_of[h] = 255;
    if (H <= 256) {
      for (int it = 0; it < H; ++it) {
        int best = -1;
        float bp = -INFINITY;
        for (int h = 0; h < H; ++h) {
          const float ph = pi_sm[h];
          if (slot_of[h] == 255 && ph > bp) {
            bp = ph;
            best = h;
          }
        }
        int ks = 0;
        for (int k = 1; k < nslots; ++k)
          if (load[k] < load[ks]) ks = k;
        slot_of[best] = (unsigned char)ks;
        load[ks] += bp + 1.f;
      }
    } else {
      for (int h = 0; h < H; ++h) slot_of[h] = (unsigned char)(h % nslots);
    }
  }
  __syncthreads();
  for (int t = tid; t < 2 * H * Dpad + H; t += 512) accB[t] = 0.f;
  for (int t = tid; t < nslots * W64; t += 512) {
    const int k = t / W64, w = t - k * W64;
    uint64_t mk = 0;
    for (int j = 0; j < 64 && (w << 6) + j < H; ++j)
      if (slot_of[(w << 6) + j] == k) mk |= 1ull << j;
    masks[t] = mk;
  }
  __syncthreads();
  double accF = 0.0, accN = 0.0;
  const float eps = (float)kNorEps;
  const size_t HD = (size_t)H * D;
  int since_flush = 0;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const int64_t row = idx ? idx[b] : b;
    const float *l = lpj + row * lpj_stride_n;
    const uint64_t *Kn = K + row * k_stride_n;
    // softmax weights (lpj2pjc, _utils.py:182-191): q of state 32 r + lane in qreg[r]
    float qreg[kNorQR], mx = -INFINITY;
#pragma unroll
    for (int r = 0; r < kNorQR; ++r) {
      qreg[r] = 32 * r + lane < S ? l[32 * r + lane] : -INFINITY;
      mx = fmaxf(mx, qreg[r]);
    }
    mx = warp_max(mx);
    float se = 0.f;
#pragma unroll
    for (int r = 0; r < kNorQR; ++r) {
      qreg[r] = 32 * r + lane < S ? expf(qreg[r] - mx) : 0.f;
      se += qreg[r];
    }
    se = warp_sum(se);
    const float inv = 1.f / se;
#pragma unroll
    for (int r = 0; r < kNorQR; ++r) qreg[r] *= inv;
    if (tid == 0) {
      accF += (double)mx + log((double)se) + *prior_sum;
      accN += 1.0;
    }
    float ym1[V];
#pragma unroll
    for (int v = 0; v < V; ++v) ym1[v] = (d + v < D ? (float)Y[b * ldY + d + v] : 1.f) - 1.f;
    // Btilde, Ctilde (noisyor.py:121-132) and new_pi (:117): this slot's units of every state, columns d .. d + V - 1
    for (int w = 0; w < W64; ++w) {
      const uint64_t slotmask = masks[slot * W64 + w];
#pragma unroll
      for (int r = 0; r < kNorQR; ++r) {  // 32 states tested at once, one per lane (the slot is warp-uniform)
        const int s0 = 32 * r, sl = s0 + lane;
        if (s0 >= S) break;
        const uint64_t word = sl < S ? __ldg(Kn + (size_t)sl * W64 + w) : 0ull;
        const uint64_t ml = word & slotmask;
        unsigned act = __ballot_sync(FULL, ml != 0ull);
        while (act) {
          const int src = __ffs((int)act) - 1;
          act &= act - 1;
          uint64_t m = __shfl_sync(FULL, ml, src);
          const uint64_t full_w = __shfl_sync(FULL, word, src);
          const int s = s0 + src;
          const float qs = __shfl_sync(FULL, qreg[r], src);
          // Q_sd over ALL active units of the state, ascending h (noisyor.py:123-124)
          float Q[V], omq[V];
#pragma unroll
          for (int v = 0; v < V; ++v) Q[v] = 1.f;
          for (int w2 = 0; w2 < W64; ++w2) {
            uint64_t f = w2 == w ? full_w : __ldg(Kn + (size_t)s * W64 + w2);
            while (f) {
              const int h2 = (w2 << 6) + __ffsll((long long)f) - 1;
              f &= f - 1;
              const float *wr = W1mT + (size_t)h2 * Dpad + d;
              if (V == 4) {
                const float4 t = __ldg(reinterpret_cast<const float4 *>(wr));
                Q[0] *= t.x, Q[V > 1 ? 1 : 0] *= t.y, Q[V > 2 ? 2 : 0] *= t.z, Q[V > 3 ? 3 : 0] *= t.w;
              } else {
                Q[0] *= __ldg(wr);
              }
            }
          }
#pragma unroll
          for (int v = 0; v < V; ++v) omq[v] = 1.f - Q[v];
          while (m) {
            const int h = (w << 6) + __ffsll((long long)m) - 1;
            m &= m - 1;
            float Ws[V], ab[V], ac[V];
            float *pb = accB + h * Dpad + d, *pc = accC + h * Dpad + d;
            if (V == 4) {
              const float4 t = __ldg(reinterpret_cast<const float4 *>(W1mT + (size_t)h * Dpad + d));
              Ws[0] = t.x, Ws[V > 1 ? 1 : 0] = t.y, Ws[V > 2 ? 2 : 0] = t.z, Ws[V > 3 ? 3 : 0] = t.w;
              const float4 u = *reinterpret_cast<const float4 *>(pb), x = *reinterpret_cast<const float4 *>(pc);
              ab[0] = u.x, ab[V > 1 ? 1 : 0] = u.y, ab[V > 2 ? 2 : 0] = u.z, ab[V > 3 ? 3 : 0] = u.w;
              ac[0] = x.x, ac[V > 1 ? 1 : 0] = x.y, ac[V > 2 ? 2 : 0] = x.z, ac[V > 3 ? 3 : 0] = x.w;
            } else {
              Ws[0] = __ldg(W1mT + (size_t)h * Dpad + d);
              ab[0] = *pb;
              ac[0] = *pc;
            }
#pragma unroll
            for (int v = 0; v < V; ++v) {
              const float Bv = __fdividef(1.f, Ws[v] * omq[v] + eps);
              const float Cv = __fdividef(Bv * Q[v], Ws[v]);
              ab[v] += qs * (Bv * ym1[v]);
              ac[v] += qs * Cv;
            }
            if (V == 4) {
              *reinterpret_cast<float4 *>(pb) = make_float4(ab[0], ab[V > 1 ? 1 : 0], ab[V > 2 ? 2 : 0], ab[V > 3 ? 3 : 0]);
              *reinterpret_cast<float4 *>(pc) = make_float4(ac[0], ac[V > 1 ? 1 : 0], ac[V > 2 ? 2 : 0], ac[V > 3 ? 3 : 0]);
            } else {
              *pb = ab[0];
              *pc = ac[0];
            }
            if (pi_owner && lane == 0) pi_sm[h] += qs;  // new_pi += <s_h>: this slot is the only writer of unit h
          }
        }
      }
    }
    if (++since_flush == kNorFlush || b + gridDim.x >= B) {  // every CTA iteration count is uniform within the CTA
      since_flush = 0;
      __syncthreads();
      for (int t = tid; t < H * Dpad; t += 512) {
        const int h = t / Dpad, dd = t - h * Dpad;
        const float vb = accB[t], vc = accC[t];
        if (dd < D) {
          if (vb != 0.f) atomicAdd(stats + (size_t)h * D + dd, (double)vb);
          if (vc != 0.f) atomicAdd(stats + HD + (size_t)h * D + dd, (double)vc);
        }
        accB[t] = 0.f;
        accC[t] = 0.f;
      }
      for (int h = tid; h < H; h += 512) {
        if (pi_sm[h] != 0.f) atomicAdd(stats + 2 * HD + h, (double)pi_sm[h]);
        pi_sm[h] = 0.f;
      }
      __syncthreads();
    }
  }
  if (tid == 0 && accN != 0.0) {
    atomicAdd(stats + 2 * HD + H, accN);
    atomicAdd(stats + 2 * HD + H + 1, accF);
  }
}

// per-device buffer of unit counts (part of the per-device context, grown on demand and kept)
static unsigned *g_nor_counts[64];
static int g_nor_counts_len[64];
static int nor_counts_buffer(int H, unsigned **out) {
  int dev = 0;
  TVO_CUDA(cudaGetDevice(&dev));
  TVO_REQUIRE(dev >= 0 && dev < 64, "noisyor counts buffer: device index %d out of range", dev);
  if (g_nor_counts_len[dev] < H) {
    if (g_nor_counts[dev]) TVO_CUDA(cudaFree(g_nor_counts[dev]));
    g_nor_counts[dev] = nullptr;
    g_nor_counts_len[dev] = 0;
    TVO_CUDA(cudaMalloc((void **)&g_nor_counts[dev], (size_t)H * sizeof(unsigned)));
    g_nor_counts_len[dev] = H;
  }
  *out = g_nor_counts[dev];
  return TVO_OK;
}

template <typename T>
static int nor_mstep_slot(const uint8_t *, int64_t, const int64_t *, const uint64_t *, int64_t, const T *, int64_t,
                          const void *, double *, int64_t, int, int, int, void *, bool &done) {
  done = false;  // fp64 mode keeps the register-stationary kernel (fp64 accumulators, IEEE division)
  return TVO_OK;
}
template <>
int nor_mstep_slot<float>(const uint8_t *Y, int64_t ldY, const int64_t *idx, const uint64_t *K, int64_t k_stride_n,
                          const float *lpj, int64_t lpj_stride_n, const void *ws, double *stats, int64_t B, int S, int H,
                          int D, void *stream, bool &done) {
  done = false;
  const int Dpad = nor_dpad(D), W64 = n_words(H);
  if (!(Dpad <= 256 && 512 % Dpad == 0 && S <= 32 * kNorQR)) return TVO_OK;
  const int nslots_h = 512 / (Dpad >= 128 ? Dpad / 4 : Dpad);
  const size_t smem = ((size_t)2 * H * Dpad + ((H + 3) & ~3)) * 4 + (size_t)nslots_h * W64 * 8 + ((H + 15) & ~15);
  if (smem > 110 * 1024) return TVO_OK;  // two CTAs per SM
  auto kern = Dpad >= 128 ? nor_mstep_slot_kernel<4> : nor_mstep_slot_kernel<1>;
  TVO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  TVO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 512, smem));
  const int gx = (int)max((int64_t)1, min(B, (int64_t)sm_count() * max(occ, 1)));
  const NorTheta<float> th = nor_theta_view<float>(ws, H, D);
  unsigned *counts = nullptr;
  {
    int rc = nor_counts_buffer(H, &counts);
    if (rc) return rc;
  }
  TVO_CUDA(cudaMemsetAsync(counts, 0, (size_t)H * sizeof(unsigned), (cudaStream_t)stream));
  const int64_t Bs = min(B, (int64_t)kNorCountSample);
  nor_unit_counts_kernel<<<(int)min((Bs * S * W64 + 255) / 256, (int64_t)sm_count() * 4), 256, (size_t)H * 4,
                           (cudaStream_t)stream>>>(K, k_stride_n, idx, Bs, S, W64, H, counts);
  TVO_AFTER_LAUNCH("noisyor_unit_counts");
  kern<<<gx, 512, smem, (cudaStream_t)stream>>>(Y, ldY, idx, K, k_stride_n, lpj, lpj_stride_n, th.W1mT, &th.hdr->prior_sum,
                                                counts, stats, B, S, H, D, W64, Dpad);
  TVO_AFTER_LAUNCH("noisyor_mstep_slot");
  done = true;
  return TVO_OK;
}

// ------------------------------------------------------------------------------------ host side
template <typename T>
int nor_prepare(const T *W, const T *pies, int H, int D, void *ws, void *stream) {
  TVO_REQUIRE(W && pies && ws, "noisyor_prepare: null pointer");
  TVO_REQUIRE(H > 0 && D > 0 && D <= 2048, "noisyor_prepare: need H > 0 and 0 < D <= 2048 (got H=%d D=%d)", H, D);
  const size_t total = (size_t)H * nor_dpad(D);
  nor_prepare_kernel<T><<<(int)max((size_t)1, min((size_t)4 * sm_count(), (total + 255) / 256)), 256, 0,
                          (cudaStream_t)stream>>>(W, pies, H, D, ws);
  TVO_AFTER_LAUNCH("noisyor_prepare");
  return TVO_OK;
}

template <typename T>
int nor_lpj(const uint8_t *Y, int64_t ldY, const uint64_t *K, int64_t k_stride_n, const int64_t *idx, const void *ws,
            T *lpj_out, int64_t lpj_stride_n, int64_t B, int S, int H, int D, void *stream) {
  TVO_REQUIRE(Y && K && ws && lpj_out, "noisyor_lpj: null pointer");
  TVO_REQUIRE(B >= 0 && S > 0 && H > 0 && D > 0 && D <= 2048, "noisyor_lpj: bad sizes");
  if (B == 0) return TVO_OK;
  const int CH = S <= 64 ? S : 64;
  const int64_t ntask = B * ((S + CH - 1) / CH);
  const int grid = (int)max((int64_t)1, min((ntask + 7) / 8, (int64_t)sm_count() * 8));
  nor_lpj_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(Y, ldY, K, k_stride_n, idx, ws, lpj_out, lpj_stride_n, B, S, H,
                                                            D, CH);
  TVO_AFTER_LAUNCH("noisyor_lpj");
  return TVO_OK;
}

template <typename T>
int nor_estep(const uint8_t *Y, int64_t ldY, const int64_t *idx, uint64_t *K, T *lpj, const void *ws,
              const tvo_evo_config *cfg, const double *sparsity, int64_t B, int S, int H, int D, int64_t *nsubs,
              void *stream) {
  TVO_REQUIRE(Y && K && lpj && ws && cfg && nsubs, "estep_evo_noisyor: null pointer");
  TVO_REQUIRE(B >= 0 && D > 0 && D <= 2048, "estep_evo_noisyor: need 0 < D <= 2048");
  typename NorModel<T>::Params mp{Y, ldY, ws, H, D};
  return launch_estep_evo_fused<T, NorModel<T>>(mp, idx, 0, K, lpj, cfg, sparsity, B, S, H, nsubs, (cudaStream_t)stream,
                                                "estep_evo_noisyor");
}

template <typename T>
int nor_mstep(const uint8_t *Y, int64_t ldY, const int64_t *idx, const uint64_t *K, int64_t k_stride_n, const T *lpj,
              int64_t lpj_stride_n, const void *ws, double *stats, int64_t B, int S, int H, int D, void *stream) {
  TVO_REQUIRE(Y && K && lpj && ws && stats, "noisyor_mstep: null pointer");
  TVO_REQUIRE(B >= 0 && S > 0 && H > 0 && D > 0, "noisyor_mstep: bad sizes");
  if (B == 0) return TVO_OK;
  const int Dpad = nor_dpad(D), W64 = n_words(H);
  {
    bool done = false;  // fp32 mode: shared-memory accumulators, one unit class per thread slot
    int rc = nor_mstep_slot<T>(Y, ldY, idx, K, k_stride_n, lpj, lpj_stride_n, ws, stats, B, S, H, D, stream, done);
    if (rc || done) return rc;
  }
  if (Dpad <= 256 && 512 % Dpad == 0) {  // register-stationary accumulators
    const size_t smem2 = align16((size_t)S * Dpad * sizeof(T)) + (size_t)S * 8 + (size_t)S * W64 * 8;
    if (smem2 <= 100 * 1024) {
      const int Ht = kNorHpt * (512 / Dpad), ntile = (H + Ht - 1) / Ht;
      auto kern2 = nor_mstep_reg_kernel<T>;
      TVO_CUDA(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      int occ2 = 1;
      TVO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, kern2, 512, smem2));
      const int gx2 = (int)max((int64_t)1, min(B, (int64_t)max(1, sm_count() * max(occ2, 1) / ntile)));
      const NorTheta<T> th2 = nor_theta_view<T>(ws, H, D);
      const int q_off = (int)align16((size_t)S * Dpad * sizeof(T)), k_off = q_off + S * 8;
      kern2<<<dim3(gx2, ntile), 512, smem2, (cudaStream_t)stream>>>(Y, ldY, idx, K, k_stride_n, lpj, lpj_stride_n, th2.W1mT,
                                                                    &th2.hdr->prior_sum, stats, B, S, H, D, W64, Dpad, q_off,
                                                                    k_off);
      TVO_AFTER_LAUNCH("noisyor_mstep_reg");
      return TVO_OK;
    }
  }
  const size_t fixed = (size_t)S * 8 + (size_t)S * W64 * 8 + 64;
  const size_t budget = 200 * 1024;
  TVO_REQUIRE(fixed + (size_t)(2 * Dpad + 1) * 8 <= budget, "noisyor_mstep: S/D too large for shared memory");
  int Ht = (int)min((size_t)H, (budget - fixed) / ((size_t)(2 * Dpad + 1) * 8));
  const int ntile = (H + Ht - 1) / Ht;
  Ht = (H + ntile - 1) / ntile;
  const size_t smem = (size_t)(2 * Ht * Dpad + Ht) * 8 + fixed;
  auto kern = nor_mstep_kernel<T>;
  TVO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = min(256, max(64, Dpad));
  int occ = 1;
  TVO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
  const int gx = (int)max((int64_t)1, min(B, (int64_t)sm_count() * max(occ, 1)));
  kern<<<dim3(gx, ntile), threads, smem, (cudaStream_t)stream>>>(Y, ldY, idx, K, k_stride_n, lpj, lpj_stride_n, ws, stats, B,
                                                                 S, H, D, Ht);
  TVO_AFTER_LAUNCH("noisyor_mstep");
  return TVO_OK;
}

template <typename T>
int logsumexp_sum(const T *lpj, int64_t stride_n, const int64_t *idx, int64_t B, int S, double *out, void *stream) {
  TVO_REQUIRE(lpj && out, "logsumexp_sum: null pointer");
  TVO_REQUIRE(B >= 0 && S > 0, "logsumexp_sum: bad sizes");
  if (B == 0) return TVO_OK;
  const int grid = (int)max((int64_t)1, min((B + 7) / 8, (int64_t)sm_count() * 8));
  logsumexp_sum_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(lpj, stride_n, idx, B, S, out);
  TVO_AFTER_LAUNCH("logsumexp_sum");
  return TVO_OK;
}

}  // namespace tvo

using namespace tvo;

extern "C" {
int64_t tvo_noisyor_theta_ws_bytes(int32_t H, int32_t D, int32_t is_f64) {
  return (int64_t)(is_f64 ? nor_ws_bytes<double>(H, D) : nor_ws_bytes<float>(H, D));
}
int64_t tvo_noisyor_stats_len(int32_t H, int32_t D) { return 2 * (int64_t)H * D + H + 2; }
int tvo_noisyor_prepare_f64(const double *W, const double *pies, int32_t H, int32_t D, void *ws, void *stream) {
  return nor_prepare<double>(W, pies, H, D, ws, stream);
}
int tvo_noisyor_prepare_f32(const float *W, const float *pies, int32_t H, int32_t D, void *ws, void *stream) {
  return nor_prepare<float>(W, pies, H, D, ws, stream);
}
int tvo_noisyor_lpj_f64(const uint8_t *Y, int64_t ldY, const uint64_t *K, int64_t k_stride_n, const int64_t *idx,
                        const void *ws, double *lpj_out, int64_t lpj_stride_n, int64_t B, int32_t S, int32_t H, int32_t D,
                        void *stream) {
  return nor_lpj<double>(Y, ldY, K, k_stride_n, idx, ws, lpj_out, lpj_stride_n, B, S, H, D, stream);
}
int tvo_noisyor_lpj_f32(const uint8_t *Y, int64_t ldY, const uint64_t *K, int64_t k_stride_n, const int64_t *idx,
                        const void *ws, float *lpj_out, int64_t lpj_stride_n, int64_t B, int32_t S, int32_t H, int32_t D,
                        void *stream) {
  return nor_lpj<float>(Y, ldY, K, k_stride_n, idx, ws, lpj_out, lpj_stride_n, B, S, H, D, stream);
}
int tvo_estep_evo_noisyor_f64(const uint8_t *Y, int64_t ldY, const int64_t *idx, uint64_t *K, double *lpj, const void *ws,
                              const tvo_evo_config *cfg, const double *sparsity, int64_t B, int32_t S, int32_t H,
                              int32_t D, int64_t *nsubs, void *stream) {
  return nor_estep<double>(Y, ldY, idx, K, lpj, ws, cfg, sparsity, B, S, H, D, nsubs, stream);
}
int tvo_estep_evo_noisyor_f32(const uint8_t *Y, int64_t ldY, const int64_t *idx, uint64_t *K, float *lpj, const void *ws,
                              const tvo_evo_config *cfg, const double *sparsity, int64_t B, int32_t S, int32_t H,
                              int32_t D, int64_t *nsubs, void *stream) {
  return nor_estep<float>(Y, ldY, idx, K, lpj, ws, cfg, sparsity, B, S, H, D, nsubs, stream);
}
int tvo_noisyor_mstep_f64(const uint8_t *Y, int64_t ldY, const int64_t *idx, const uint64_t *K, int64_t k_stride_n,
                          const double *lpj, int64_t lpj_stride_n, const void *ws, double *stats, int64_t B, int32_t S,
                          int32_t H, int32_t D, void *stream) {
  return nor_mstep<double>(Y, ldY, idx, K, k_stride_n, lpj, lpj_stride_n, ws, stats, B, S, H, D, stream);
}
int tvo_noisyor_mstep_f32(const uint8_t *Y, int64_t ldY, const int64_t *idx, const uint64_t *K, int64_t k_stride_n,
                          const float *lpj, int64_t lpj_stride_n, const void *ws, double *stats, int64_t B, int32_t S,
                          int32_t H, int32_t D, void *stream) {
  return nor_mstep<float>(Y, ldY, idx, K, k_stride_n, lpj, lpj_stride_n, ws, stats, B, S, H, D, stream);
}
int tvo_logsumexp_sum_f64(const double *lpj, int64_t stride_n, const int64_t *idx, int64_t B, int32_t S, double *out,
                          void *stream) {
  return logsumexp_sum<double>(lpj, stride_n, idx, B, S, out, stream);
}
int tvo_logsumexp_sum_f32(const float *lpj, int64_t stride_n, const int64_t *idx, int64_t B, int32_t S, double *out,
                          void *stream) {
  return logsumexp_sum<float>(lpj, stride_n, idx, B, S, out, stream);
}
}
