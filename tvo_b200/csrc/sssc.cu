// tvo_b200 — Spike-and-Slab Sparse Coding kernels (tvo/models/sssc.py:175-540).
//
// For a state s with active set A (|A| = a) the reference (sssc.py:216-337) forms
//   Lambda^-1 = W_A^T W_A / s2 + Psi_A^-1,   C_s^-1 = I/s2 - W_A Lambda W_A^T / s2^2,
//   lpj = sum_A log(pi/(1-pi)) - 1/2 [ logdet(Lambda^-1) + logdet(Psi_A) + r^T C_s^-1 r ],  r = y - W_A mu_A
// with explicit a x a inverses in a python loop per (n, s).  Here, per state and warp-cooperatively in
// shared memory, with G = W^T W, a = W^T y (one GEMM per chunk) and Psi_A = R R^T (Cholesky):
//   M = I + R^T G_AA R / s2 = C C^T            (SPD; det M = det(Lambda^-1) det(Psi_A))
//   b = a_A - G_AA mu_A = W_A^T r,  ||r||^2 = yy - 2 mu_A.a_A + mu_A^T G_AA mu_A
//   r^T C_s^-1 r = ||r||^2/s2 - ||C^-1 R^T b||^2 / s2^2,   logdet terms = 2 sum log C_ii
//   Lambda = U U^T with U = R C^-T,  kappa = mu_A + Lambda b / s2              (M-step, sssc.py:418-423)
// No explicit inverse is formed.  Rows of y with NaN take a slow path with G_AA, a_A restricted to the
// observed dimensions (sssc.py:283,298).  All arithmetic is fp64; lpj is stored in the model precision.
#include <cublas_v2.h>

#include "estep_fused.cuh"
#include "bsc_gemm.cuh"

namespace tvo {

struct SsscThetaHdr {
  double sigma2;
  double prior_sum;  // sum_h log(1 - clamp(pi_h))        (sssc.py:356-359)
  double log_2pi_sigma2;  // log(2 pi) + log(sigma2)      (sssc.py:360)
  double pad[5];
};
__host__ __device__ inline int sssc_dpad(int D) { return pad32(D); }
// workspace: hdr | prior (H dbl) | mus (H dbl) | Psi (H*H dbl) | G (H*H dbl) | Wt (H*Dpad, T)
__host__ __device__ inline size_t sssc_off_prior() { return sizeof(SsscThetaHdr); }
__host__ __device__ inline size_t sssc_off_mus(int H) { return sssc_off_prior() + (size_t)H * 8; }
__host__ __device__ inline size_t sssc_off_psi(int H) { return sssc_off_mus(H) + (size_t)H * 8; }
__host__ __device__ inline size_t sssc_off_g(int H) { return sssc_off_psi(H) + (size_t)H * H * 8; }
__host__ __device__ inline size_t sssc_off_wt(int H) { return align16(sssc_off_g(H) + (size_t)H * H * 8); }
// fp64 model: W once more as Wg[k/4][h][k%4] (k zero-padded to a multiple of 16), the operand layout of the
// hand-written FP64 tensor-core GEMM a = Y W (bsc_gemm.cuh)
template <typename T>
__host__ __device__ inline size_t sssc_off_wg(int H, int D) {
  return align16(sssc_off_wt(H) + (size_t)H * sssc_dpad(D) * sizeof(T));
}
template <typename T>
__host__ __device__ inline size_t sssc_ws_bytes(int H, int D) {
  return sssc_off_wg<T>(H, D) + (sizeof(T) == 8 ? (size_t)((D + 15) & ~15) * H * 8 : 0);
}
template <typename T>
struct SsscTheta {
  const SsscThetaHdr *hdr;
  const double *prior, *mus, *Psi, *G;
  const T *Wt;
  const double *Wg;  // fp64 model only
  int Dpad;
};
template <typename T>
__host__ __device__ inline SsscTheta<T> sssc_theta_view(const void *ws, int H, int D) {
  const unsigned char *p = (const unsigned char *)ws;
  SsscTheta<T> t;
  t.hdr = (const SsscThetaHdr *)p;
  t.prior = (const double *)(p + sssc_off_prior());
  t.mus = (const double *)(p + sssc_off_mus(H));
  t.Psi = (const double *)(p + sssc_off_psi(H));
  t.G = (const double *)(p + sssc_off_g(H));
  t.Wt = (const T *)(p + sssc_off_wt(H));
  t.Wg = (const double *)(p + sssc_off_wg<T>(H, D));
  t.Dpad = sssc_dpad(D);
  return t;
}

template <typename T>
__global__ void sssc_prepare_kernel(const T *__restrict__ W, const T *__restrict__ sigma2, const T *__restrict__ mus,
                                    const T *__restrict__ Psi, const T *__restrict__ pies, int H, int D, void *ws) {
  unsigned char *p = (unsigned char *)ws;
  SsscThetaHdr *hdr = (SsscThetaHdr *)p;
  double *prior = (double *)(p + sssc_off_prior());
  double *mu = (double *)(p + sssc_off_mus(H));
  double *Ps = (double *)(p + sssc_off_psi(H));
  double *G = (double *)(p + sssc_off_g(H));
  T *Wt = (T *)(p + sssc_off_wt(H));
  const int Dpad = sssc_dpad(D);
  const size_t gt = (size_t)gridDim.x * blockDim.x, g0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (size_t i = g0; i < (size_t)H * Dpad; i += gt) {
    const int h = (int)(i / Dpad), d = (int)(i % Dpad);
    Wt[i] = d < D ? W[(size_t)d * H + h] : T(0);
  }
  if (sizeof(T) == 8) {
    double *Wg = (double *)(p + sssc_off_wg<T>(H, D));
    for (size_t i = g0; i < (size_t)((D + 15) & ~15) * H; i += gt) {
      const int kk = (int)(i & 3), h = (int)((i >> 2) % H), d = 4 * (int)((i >> 2) / H) + kk;
      Wg[i] = d < D ? (double)W[(size_t)d * H + h] : 0.0;
    }
  }
  for (size_t i = g0; i < (size_t)H * H; i += gt) {
    Ps[i] = (double)Psi[i];
    const int h1 = (int)(i / H), h2 = (int)(i % H);
    double acc = 0.0;
    for (int d = 0; d < D; ++d) acc += (double)W[(size_t)d * H + h1] * (double)W[(size_t)d * H + h2];
    G[i] = acc;
  }
  for (size_t h = g0; h < (size_t)H; h += gt) {
    T pi = pies[h];
    pi = pi < T(1e-2) ? T(1e-2) : (pi > T(1.0 - 1e-2) ? T(1.0 - 1e-2) : pi);  // clamp (sssc.py:276)
    prior[h] = (double)log(pi / (T(1) - pi));
    mu[h] = (double)mus[h];
  }
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    double s = 0.0;
    for (int h = threadIdx.x; h < H; h += 32) {
      T pi = pies[h];
      pi = pi < T(1e-2) ? T(1e-2) : (pi > T(1.0 - 1e-2) ? T(1.0 - 1e-2) : pi);
      s += (double)log(T(1) - pi);
    }
    s = warp_sum(s);
    if (threadIdx.x == 0) {
      hdr->sigma2 = (double)sigma2[0];
      hdr->prior_sum = s;
      hdr->log_2pi_sigma2 = log(2.0 * 3.14159265358979323846) + (double)log(sigma2[0]);
    }
  }
}

// ------------------------------------------------------------------------------------ per-state solve
// One state is solved cooperatively by a GROUP of GS lanes.  States are sparse (|A| ~ 1-4), so a whole
// warp per state leaves most lanes idle: the scorers below run 32/kSsscGS states of a datapoint
// concurrently in sub-warp groups (scratch for up to kSsscAS active units each), then the rare denser
// states with the full warp (up to kSsscAM units in the same shared memory, up to kSsscAB in a global
// scratch row of the warp).  States beyond kSsscAB set the overflow flag (finfo.min, sssc.py:343-345).
constexpr int kSsscGS = 4;   // lanes per sub-warp group
constexpr int kSsscNG = 32 / kSsscGS;
constexpr int kSsscAS = 6;   // active units handled by a sub-warp group (stored states at C3: |A| histogram 0: 3 %, 1: 81 %, 2: 11 %, 3: 3.6 %, 4: 1.2 %, >= 7: 0.02 %; children carry a flip or two more: 4 sent too many of them to the full-warp path)
constexpr int kSsscAM = 16;  // ... by the full warp in shared memory (fits inside the eight group scratches)
constexpr int kSsscAB = 32;  // ... by the full warp in its global scratch row

struct SsscScratch {
  double *R;    // amax*amax : Psi_A -> Cholesky factor R (lower)
  double *Gm;   // amax*amax : G_AA  -> M -> Cholesky factor C (lower)
  double *X;    // amax*amax : G R, then C^-1, then U = R C^-T; finally Lambda + kappa kappa^T (M-step)
  double *vec;  // 6*amax    : mu_A | a_A | b | v / z | kappa | tmp
  int *act;     // amax
};
__host__ __device__ inline size_t sssc_scratch_bytes(int amax) {
  return (size_t)3 * amax * amax * 8 + (size_t)6 * amax * 8 + align16((size_t)amax * 4);
}
// per-warp shared scratch: kSsscNG group scratches, re-carved as ONE kSsscAM scratch for the full warp
__host__ __device__ inline size_t sssc_warp_scratch_bytes() {
  const size_t g = (size_t)kSsscNG * sssc_scratch_bytes(kSsscAS), m = sssc_scratch_bytes(kSsscAM);
  return align16(g > m ? g : m);
}
__host__ __device__ inline size_t sssc_pool_slot_bytes(int H) { return (sssc_scratch_bytes(H) + 255) & ~(size_t)255; }
__device__ inline SsscScratch sssc_carve(unsigned char *p, int amax) {
  SsscScratch s;
  s.R = (double *)p;
  s.Gm = s.R + (size_t)amax * amax;
  s.X = s.Gm + (size_t)amax * amax;
  s.vec = s.X + (size_t)amax * amax;
  s.act = (int *)(s.vec + 6 * amax);
  return s;
}

template <int GS>
__device__ __forceinline__ double group_sum(double v, unsigned gmask) {
#pragma unroll
  for (int o = GS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
  return v;
}

// in-place Cholesky (lower) of the a x a SPD matrix A (row stride ld); returns false on a bad pivot
template <int GS>
__device__ inline bool chol_lower(double *A, int a, int ld, int gl, unsigned gmask) {
  bool ok = true;
  for (int j = 0; j < a; ++j) {
    __syncwarp(gmask);
    double djj = A[j * ld + j];
    for (int k = 0; k < j; ++k) djj -= A[j * ld + k] * A[j * ld + k];
    ok = ok && (djj > 0.0);
    const double ljj = sqrt(djj);
    for (int i = j + 1 + gl; i < a; i += GS) {
      double v = A[i * ld + j];
      for (int k = 0; k < j; ++k) v -= A[i * ld + k] * A[j * ld + k];
      A[i * ld + j] = v / ljj;
    }
    __syncwarp(gmask);
    if (gl == 0) A[j * ld + j] = ljj;
  }
  __syncwarp(gmask);
  return ok;
}

__device__ __forceinline__ int state_popcount(const uint64_t *st, int W64) {
  int a = 0;
  for (int w = 0; w < W64; ++w) a += __popcll(st[w]);
  return a;
}

// lpj of one state, cooperatively by the GS lanes of `gmask` (gl = lane within the group).  The state has
// a <= amax active units (checked by the caller).  a_vec: shared (H) = W^T y; yy = ||y||^2.  If
// want_moments, sc.vec[4*amax ..] receives kappa (a) and sc.X receives Lambda + kappa kappa^T (a x a).
// y_nan_row != nullptr selects the restricted (NaN-aware) path.  Returns lpj (NaN if not evaluable).
template <typename T, int GS>
__device__ inline double sssc_state(const uint64_t *st, int W64, int H, int D, const SsscTheta<T> &th,
                                    const double *a_vec, double yy, const T *y_nan_row, const SsscScratch &sc, int amax,
                                    bool want_moments, int a, int gl, unsigned gmask) {
  const double s2 = th.hdr->sigma2;
  __syncwarp(gmask);
  if (gl == 0) {
    int k = 0;
    for (int w = 0; w < W64; ++w) {
      uint64_t m = st[w];
      while (m) {
        sc.act[k++] = (w << 6) + __ffsll((long long)m) - 1;
        m &= m - 1;
      }
    }
  }
  __syncwarp(gmask);
  double *mu = sc.vec, *aA = sc.vec + amax, *bv = sc.vec + 2 * amax, *vz = sc.vec + 3 * amax, *kap = sc.vec + 4 * amax;
  double prior = 0.0;
  for (int i = 0; i < a; ++i) prior += th.prior[sc.act[i]];
  if (a == 0) {
    double yyn = yy;
    if (y_nan_row) {
      double t = 0.0;
      for (int d = gl; d < D; d += GS) {
        const double v = (double)y_nan_row[d];
        if (v == v) t += v * v;
      }
      yyn = group_sum<GS>(t, gmask);
    }
    return -0.5 * (yyn / s2);
  }
  const int ld = amax;
  // Psi_A, G_AA, mu_A, a_A
  for (int t = gl; t < a * a; t += GS) {
    const int i = t / a, j = t % a;
    sc.R[i * ld + j] = th.Psi[(size_t)sc.act[i] * H + sc.act[j]];
    if (!y_nan_row) sc.Gm[i * ld + j] = th.G[(size_t)sc.act[i] * H + sc.act[j]];
  }
  for (int i = gl; i < a; i += GS) {
    mu[i] = th.mus[sc.act[i]];
    if (!y_nan_row) aA[i] = a_vec[sc.act[i]];
  }
  if (y_nan_row) {  // restricted to the observed dimensions (sssc.py:283-298)
    for (int t = gl; t < a * a + a; t += GS) {
      const int i = t < a * a ? t / a : t - a * a, j = t < a * a ? t % a : -1;
      const T *wi = th.Wt + (size_t)sc.act[i] * th.Dpad;
      const T *wj = j >= 0 ? th.Wt + (size_t)sc.act[j] * th.Dpad : nullptr;
      double acc = 0.0;
      for (int d = 0; d < D; ++d) {
        const double yv = (double)y_nan_row[d];
        if (yv == yv) acc += (double)wi[d] * (j >= 0 ? (double)wj[d] : yv);
      }
      if (j >= 0) sc.Gm[i * ld + j] = acc; else aA[i] = acc;
    }
    double t = 0.0;
    for (int d = gl; d < D; d += GS) {
      const double v = (double)y_nan_row[d];
      if (v == v) t += v * v;
    }
    yy = group_sum<GS>(t, gmask);
  }
  __syncwarp(gmask);
  // b = a_A - G mu,  rr = yy - 2 mu.a_A + mu^T G mu
  for (int i = gl; i < a; i += GS) {
    double gm = 0.0;
    for (int k = 0; k < a; ++k) gm += sc.Gm[i * ld + k] * mu[k];
    bv[i] = aA[i] - gm;
    vz[i] = mu[i] * (gm - 2.0 * aA[i]);
  }
  __syncwarp(gmask);
  double rr = yy;
  for (int i = 0; i < a; ++i) rr += vz[i];
  // Psi_A = R R^T
  bool ok = chol_lower<GS>(sc.R, a, ld, gl, gmask);
  // X = G R  (R lower: k >= j)
  for (int t = gl; t < a * a; t += GS) {
    const int i = t / a, j = t % a;
    double acc = 0.0;
    for (int k = j; k < a; ++k) acc += sc.Gm[i * ld + k] * sc.R[k * ld + j];
    sc.X[i * ld + j] = acc;
  }
  __syncwarp(gmask);
  // M = I + R^T X / s2  (into Gm)
  for (int t = gl; t < a * a; t += GS) {
    const int i = t / a, j = t % a;
    double acc = 0.0;
    for (int k = i; k < a; ++k) acc += sc.R[k * ld + i] * sc.X[k * ld + j];
    sc.Gm[i * ld + j] = acc / s2 + (i == j ? 1.0 : 0.0);
  }
  __syncwarp(gmask);
  ok = chol_lower<GS>(sc.Gm, a, ld, gl, gmask) && ok;  // M = C C^T
  double logdet = 0.0;
  for (int i = 0; i < a; ++i) logdet += log(sc.Gm[i * ld + i]);
  logdet *= 2.0;
  // v = R^T b ; z = C^-1 v (forward substitution)
  for (int i = gl; i < a; i += GS) {
    double acc = 0.0;
    for (int k = i; k < a; ++k) acc += sc.R[k * ld + i] * bv[k];
    vz[i] = acc;
  }
  __syncwarp(gmask);
  if (gl == 0) {
    for (int i = 0; i < a; ++i) {
      double v = vz[i];
      for (int k = 0; k < i; ++k) v -= sc.Gm[i * ld + k] * vz[k];
      vz[i] = v / sc.Gm[i * ld + i];
    }
  }
  __syncwarp(gmask);
  double zz = 0.0;
  for (int i = 0; i < a; ++i) zz += vz[i] * vz[i];
  const double quad = rr / s2 - zz / (s2 * s2);
  double lpj = prior - 0.5 * (logdet + quad);
  if (!ok) lpj = __longlong_as_double(0x7ff8000000000000ll);
  if (want_moments) {
    // Cinv = C^-1 (lower) column by column into X ; U = R Cinv^T ; Lambda = U U^T
    for (int j = gl; j < a; j += GS) {
      for (int i = 0; i < a; ++i) {
        double v = (i == j) ? 1.0 : 0.0;
        for (int k = j; k < i; ++k) v -= sc.Gm[i * ld + k] * sc.X[k * ld + j];
        sc.X[i * ld + j] = i < j ? 0.0 : v / sc.Gm[i * ld + i];
      }
    }
    __syncwarp(gmask);
    // t = C^-T z  -> kappa = mu + R t / s2 ;  U = R Cinv^T written over Gm (C no longer needed)
    for (int i = gl; i < a; i += GS) {
      double acc = 0.0;
      for (int k = i; k < a; ++k) acc += sc.X[k * ld + i] * vz[k];
      kap[i] = acc;  // t_i
    }
    __syncwarp(gmask);
    for (int t = gl; t < a * a; t += GS) {
      const int i = t / a, j = t % a;
      double acc = 0.0;
      for (int k = 0; k <= i; ++k) acc += sc.R[i * ld + k] * sc.X[j * ld + k];  // Cinv^T[k][j] = Cinv[j][k]
      sc.Gm[i * ld + j] = acc;  // U
    }
    for (int i = gl; i < a; i += GS) {
      double acc = 0.0;
      for (int k = 0; k <= i; ++k) acc += sc.R[i * ld + k] * kap[k];
      bv[i] = mu[i] + acc / s2;  // kappa_i (bv reused; b no longer needed)
    }
    __syncwarp(gmask);
    for (int i = gl; i < a; i += GS) kap[i] = bv[i];
    for (int t = gl; t < a * a; t += GS) {
      const int i = t / a, j = t % a;
      double acc = 0.0;
      for (int k = 0; k < a; ++k) acc += sc.Gm[i * ld + k] * sc.Gm[j * ld + k];
      sc.X[i * ld + j] = acc + bv[i] * bv[j];  // Lambda + kappa kappa^T  (sssc.py:421-423)
    }
    __syncwarp(gmask);
  }
  return lpj;
}

template <typename T>
__device__ __forceinline__ T sssc_finish(double lpj) {
  // lpj[isnan] = lpj[isinf] = finfo.min  (sssc.py:343-345)
  const double mn = sizeof(T) == 8 ? -1.7976931348623157e308 : -3.4028234663852886e38;
  if (lpj != lpj || lpj == INFINITY || lpj == -INFINITY) return (T)mn;
  if (sizeof(T) == 4 && (lpj < mn)) return (T)mn;
  return (T)lpj;
}

template <typename T>
struct SsscModel {
  struct Params {
    const T *Y;
    int64_t ldY;
    const T *A;  // B x H : W^T y per datapoint of this launch
    const void *ws;
    int H, D;
    unsigned char *big;  // global scratch, one sssc_scratch_bytes(kSsscAB) row per warp of the grid (or nullptr)
    int *overflow;       // set when a state cannot be evaluated (more than kSsscAB active units and no pool)
    // states with more than kSsscAB active units (the reference evaluates any |s|, sssc.py:287-335): a small pool
    // of scratch slots sized for |s| = H in global memory, taken with a spin lock by the warp that needs one
    unsigned char *pool;
    int *pool_locks;
    int pool_slots;
  };
  // fp64: the a = W^T y row is read where the GEMM left it (global scratch, through L1/L2) — 2 kB per warp less
  // shared memory, i.e. two more resident warps per SM at H = 256; fp32 keeps a converted fp64 copy in shared memory
  __host__ __device__ static size_t a_row_bytes(const Params &p) { return sizeof(T) == 8 ? 0 : (size_t)p.H * 8; }
  __host__ __device__ static size_t smem_bytes(const Params &p) { return a_row_bytes(p) + sssc_warp_scratch_bytes(); }
  struct Ctx {
    double yy;
    bool has_nan;
    const T *y;
    const double *a_sm;
    unsigned char *scr;  // the warp's shared scratch
    __device__ void load(const Params &p, int64_t b, unsigned char *smem, int lane) {
      y = p.Y + b * p.ldY;
      scr = smem + a_row_bytes(p);
      double t = 0.0;
      bool nan = false;
      for (int d = lane; d < p.D; d += 32) {
        const double v = (double)y[d];
        nan |= (v != v);
        t += v * v;
      }
      yy = warp_sum(t);
      has_nan = __any_sync(FULL, nan);
      if (sizeof(T) == 8) {
        a_sm = reinterpret_cast<const double *>(p.A) + b * (int64_t)p.H;
      } else {
        double *a_w = (double *)smem;
        for (int h = lane; h < p.H; h += 32) a_w[h] = (double)p.A[b * (int64_t)p.H + h];
        a_sm = a_w;
      }
    }
    // Evaluate `count` states.  `sink(s, value, n_active, scratch, amax, gl, GS-lane-count, gmask)` is called by
    // every lane of the group that solved state s (want_moments: scratch holds kappa and Lambda + kappa kappa^T).
    template <bool MOMENTS, typename Sink>
    __device__ __forceinline__ void for_each_state(const Params &p, const uint64_t *states, int count, int W64, int lane,
                                                   Sink &&sink) {
      const SsscTheta<T> th = sssc_theta_view<T>(p.ws, p.H, p.D);
      const int grp = lane / kSsscGS, gl = lane % kSsscGS;
      const unsigned gmask = ((1u << kSsscGS) - 1u) << (grp * kSsscGS);
      const SsscScratch scg = sssc_carve(scr + (size_t)grp * sssc_scratch_bytes(kSsscAS), kSsscAS);
      const T *ynan = has_nan ? y : nullptr;
      bool dense_seen = false;
      __syncwarp();
      for (int s0 = 0; s0 < count; s0 += kSsscNG) {  // sparse states: one per sub-warp group
        const int s = s0 + grp;
        if (s < count) {
          const uint64_t *st = states + (size_t)s * W64;
          const int a = state_popcount(st, W64);
          if (a <= kSsscAS) {
            const double v = sssc_state<T, kSsscGS>(st, W64, p.H, p.D, th, a_sm, yy, ynan, scg, kSsscAS, MOMENTS, a, gl, gmask);
            sink(s, v, a, scg, kSsscAS, gl, kSsscGS, gmask);
          } else {
            dense_seen = true;
          }
        }
      }
      __syncwarp();
      if (!__any_sync(FULL, dense_seen)) return;
      const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
      for (int s = 0; s < count; ++s) {  // denser states: the full warp, one at a time
        const uint64_t *st = states + (size_t)s * W64;
        const int a = state_popcount(st, W64);
        if (a <= kSsscAS) continue;
        double v = __longlong_as_double(0x7ff8000000000000ll);  // NaN -> finfo.min (sssc.py:343-345)
        if (a <= kSsscAM) {
          const SsscScratch sc = sssc_carve(scr, kSsscAM);
          v = sssc_state<T, 32>(st, W64, p.H, p.D, th, a_sm, yy, ynan, sc, kSsscAM, MOMENTS, a, lane, FULL);
          sink(s, v, a, sc, kSsscAM, lane, 32, FULL);
        } else if (a <= kSsscAB && p.big) {
          const SsscScratch sc = sssc_carve(p.big + (size_t)wid * sssc_scratch_bytes(kSsscAB), kSsscAB);
          v = sssc_state<T, 32>(st, W64, p.H, p.D, th, a_sm, yy, ynan, sc, kSsscAB, MOMENTS, a, lane, FULL);
          sink(s, v, a, sc, kSsscAB, lane, 32, FULL);
        } else if (p.pool) {  // the slow tier: any number of active units, scratch sized for |s| = H
          int slot = -1;
          if (lane == 0) {
            int i = wid % p.pool_slots;
            while (atomicCAS(p.pool_locks + i, 0, 1) != 0) i = i + 1 == p.pool_slots ? 0 : i + 1;
            slot = i;
          }
          slot = __shfl_sync(FULL, slot, 0);
          const SsscScratch sc = sssc_carve(p.pool + (size_t)slot * sssc_pool_slot_bytes(p.H), p.H);
          v = sssc_state<T, 32>(st, W64, p.H, p.D, th, a_sm, yy, ynan, sc, p.H, MOMENTS, a, lane, FULL);
          sink(s, v, a, sc, p.H, lane, 32, FULL);
          __syncwarp();
          __threadfence();
          if (lane == 0) atomicExch(p.pool_locks + slot, 0);
        } else {
          if (lane == 0) *p.overflow = 1;
          sink(s, v, -1, scg, kSsscAS, lane, 32, FULL);
        }
        __syncwarp();
      }
    }
    __device__ void score(const Params &p, const uint64_t *states, int count, int W64, T *out, int lane) {
      for_each_state<false>(p, states, count, W64, lane,
                            [&](int s, double v, int, const SsscScratch &, int, int gl, int, unsigned) {
                              if (gl == 0) out[s] = sssc_finish<T>(v);
                            });
      __syncwarp();
    }
  };
};

// log_pseudo_joint for explicit state sets: one warp per (datapoint, chunk of states)
template <typename T>
__global__ void __launch_bounds__(256)
sssc_lpj_kernel(typename SsscModel<T>::Params p, const uint64_t *__restrict__ K, int64_t k_stride_n,
                const int64_t *__restrict__ idx, T *__restrict__ lpj_out, int64_t lpj_stride_n, int64_t B, int S, int CH,
                size_t msm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int W64 = n_words(p.H);
  const int nchunk = (S + CH - 1) / CH;
  typename SsscModel<T>::Ctx ctx;
  for (int64_t task = (int64_t)blockIdx.x * wpb + wib; task < B * nchunk; task += (int64_t)gridDim.x * wpb) {
    const int64_t b = task / nchunk;
    const int c = (int)(task % nchunk);
    const int64_t row = idx ? idx[b] : b;
    __syncwarp();
    ctx.load(p, b, smem_raw + (size_t)wib * msm, lane);
    __syncwarp();
    const int s0 = c * CH, n = min(S, s0 + CH) - s0;
    ctx.score(p, K + row * k_stride_n + (size_t)s0 * W64, n, W64, lpj_out + row * lpj_stride_n + s0, lane);
  }
}

// ------------------------------------------------------------------------------------ M-step
// stats: [y_szT^T (H*D) | sz_szT (H*H) | szszT (H*H) | ssT (H*H, UPPER) | ssz (H*H) | s (H) | sz (H) |
//         diag_yyT (D) | N | F]
// The kernel accumulates szszT, ssT, s, sz, diag_yyT, N, F and writes X = <kappa>_q (B x H) and
// E = <s>_q (B x H); the three outer-product sums are GEMMs on X, E, Y done by the host wrapper.
template <typename T>
__global__ void __launch_bounds__(512)  // up to 16 warps: one CTA per SM at H = 256 (17 kB per warp), see the launcher
sssc_mstep_kernel(typename SsscModel<T>::Params p, const int64_t *__restrict__ idx, int64_t row0,
                  const uint64_t *__restrict__ K, int64_t k_stride_n, const T *__restrict__ lpj, int64_t lpj_stride_n,
                  double *__restrict__ stats, double *__restrict__ X_out, double *__restrict__ E_out,
                  double *__restrict__ Yd_out, int64_t B, int S, size_t wbytes, int nan_aware) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int H = p.H, D = p.D, W64 = n_words(H);
  const SsscTheta<T> th = sssc_theta_view<T>(p.ws, H, D);
  unsigned char *base = smem_raw + (size_t)wib * wbytes;
  typename SsscModel<T>::Ctx ctx;
  double *xk = (double *)(base + SsscModel<T>::smem_bytes(p));  // <kappa> (H)
  double *xe = xk + H;                                           // <s> (H)
  double *q = xe + H;                                            // S
  uint64_t *Ksm = (uint64_t *)(q + S);                           // S*W64
  const size_t HD = (size_t)H * D, HH = (size_t)H * H;
  double *szszT = stats + HD + HH, *ssT = szszT + HH, *s_acc = ssT + 2 * HH, *sz_acc = s_acc + H;
  double *yy_acc = sz_acc + H, *Ncnt = yy_acc + D, *Facc = Ncnt + 1;
  double accF = 0.0, accN = 0.0;
  for (int64_t b = (int64_t)blockIdx.x * wpb + wib; b < B; b += (int64_t)gridDim.x * wpb) {
    const int64_t row = idx ? idx[b] : row0 + b;
    const T *l = lpj + row * lpj_stride_n;
    __syncwarp();
    ctx.load(p, b, base, lane);
    for (int t = lane; t < S * W64; t += 32) Ksm[t] = K[row * k_stride_n + t];
    for (int h = lane; h < H; h += 32) {
      xk[h] = 0.0;
      xe[h] = 0.0;
    }
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
      const double v = (double)ctx.y[d];
      dn += (v == v) ? 1 : 0;
      atomicAdd(yy_acc + d, v * v);  // sum_n y^2 (sssc.py:445)
      if (Yd_out) Yd_out[b * (int64_t)D + d] = v;
    }
    dn = warp_sum(dn);
    accF += mx + log(se) + th.hdr->prior_sum - 0.5 * (double)dn * th.hdr->log_2pi_sigma2;
    accN += 1.0;
    __syncwarp();
    // M-step moments with the batch's own y (NaNs were filled by the Trainer): the plain path.  nan_aware
    // (data_estimator on incomplete data, sssc.py:563-597): kappa from the observed dimensions only.
    const bool saved_nan = ctx.has_nan;
    if (!nan_aware) ctx.has_nan = false;
    ctx.template for_each_state<true>(
        p, Ksm, S, W64, lane, [&](int s, double, int na, const SsscScratch &sc, int amax, int gl, int gs, unsigned gmask) {
          const double qs = q[s] / se;
          if (na > 0 && qs != 0.0) {
            const double *kap = sc.vec + 4 * amax;
            for (int i = gl; i < na; i += gs) {
              const int h = sc.act[i];
              atomicAdd(xe + h, qs);  // groups of the warp work on different states of the same datapoint
              atomicAdd(xk + h, qs * kap[i]);
            }
            for (int t = gl; t < na * na; t += gs) {
              const int i = t / na, j = t % na;
              const int hi = sc.act[i], hj = sc.act[j];
              atomicAdd(szszT + (size_t)hi * H + hj, qs * sc.X[i * amax + j]);  // <Lambda + kappa kappa^T>
              if (j >= i) atomicAdd(ssT + (size_t)hi * H + hj, qs);             // <s s^T>, upper triangle
            }
          }
          __syncwarp(gmask);
        });
    ctx.has_nan = saved_nan;
    __syncwarp();
    for (int h = lane; h < H; h += 32) {
      const double e = xe[h], k = xk[h];
      X_out[b * (int64_t)H + h] = k;
      E_out[b * (int64_t)H + h] = e;
      if (e != 0.0) {
        atomicAdd(s_acc + h, e);
        atomicAdd(sz_acc + h, k);
      }
    }
  }
  if (lane == 0 && accN != 0.0) {
    atomicAdd(Ncnt, accN);
    atomicAdd(Facc, accF);
  }
}

// ------------------------------------------------------------------------------------ host side
static cublasHandle_t g_sssc_handles[64];
static int sssc_cublas(cudaStream_t st, cublasHandle_t &h) {
  int dev = 0;
  TVO_CUDA(cudaGetDevice(&dev));
  TVO_REQUIRE(dev >= 0 && dev < 64, "cublas: device index %d out of range", dev);
  if (!g_sssc_handles[dev] && cublasCreate(&g_sssc_handles[dev]) != CUBLAS_STATUS_SUCCESS)
    return set_error(TVO_ECUDA, "cublasCreate failed");
  h = g_sssc_handles[dev];
  if (cublasSetStream(h, st) != CUBLAS_STATUS_SUCCESS) return set_error(TVO_ECUDA, "cublasSetStream failed");
  return TVO_OK;
}
#define TVO_CUBLAS(expr)                                                                     \
  do {                                                                                       \
    cublasStatus_t s_ = (expr);                                                              \
    if (s_ != CUBLAS_STATUS_SUCCESS) return set_error(TVO_ECUDA, "%s failed (%d)", #expr, (int)s_); \
    count_launch();                                                                          \
  } while (0)

// a = Y W: the hand-written TMA + FP64 tensor-core kernel of bsc_gemm.cuh where its shape class applies (H in
// {64, 128, 256}, D a multiple of 16), the library GEMM otherwise and for the fp32 model
static int sssc_gemm_A(cublasHandle_t h, const double *Wt, int Dpad, const double *Y, int64_t ldY, double *A, int Bc, int H, int D,
                       const double *Wg = nullptr, cudaStream_t st = nullptr) {
  if (Wg && gemm_yw_supported(H, D, ldY, Y) && ((uintptr_t)A % 16) == 0) {
    cudaError_t e = launch_gemm_yw(Y, ldY, Wg, A, nullptr, Bc, H, D, sm_count(), st);
    if (e != cudaSuccess) return set_error(TVO_ECUDA, "launch gemm_yw (sssc): %s", cudaGetErrorString(e));
    count_launch();
    return TVO_OK;
  }
  const double one = 1.0, zero = 0.0;
  TVO_CUBLAS(cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, H, Bc, D, &one, Wt, Dpad, Y, (int)ldY, &zero, A, H));
  return TVO_OK;
}
// slab partials of the M-step GEMMs (per-device context, grown on demand and kept)
static double *g_sssc_part[64];
static size_t g_sssc_part_bytes[64];
static int sssc_part_buffer(size_t bytes, double **out) {
  int dev = 0;
  TVO_CUDA(cudaGetDevice(&dev));
  TVO_REQUIRE(dev >= 0 && dev < 64, "sssc part buffer: device index %d out of range", dev);
  if (g_sssc_part_bytes[dev] < bytes) {
    if (g_sssc_part[dev]) TVO_CUDA(cudaFree(g_sssc_part[dev]));
    g_sssc_part[dev] = nullptr;
    g_sssc_part_bytes[dev] = 0;
    TVO_CUDA(cudaMalloc((void **)&g_sssc_part[dev], bytes));
    g_sssc_part_bytes[dev] = bytes;
  }
  *out = g_sssc_part[dev];
  return TVO_OK;
}
// C (H x N, row stride ldC) += E^T (H x Bc) Z (Bc x N, row stride ldZ): own kernel in column blocks it covers
// (N = D for y_szT; 128-column blocks of the H x H products), library GEMM otherwise
static int sssc_gemm_EtZ(cublasHandle_t hb, cudaStream_t st, const double *E, const double *Z, int64_t ldZ, double *C, int ldC,
                         int Bc, int H, int N) {
  const int nb = gemm_ety_supported(H, N, ldZ, Z, E) ? N : (N % 128 == 0 && gemm_ety_supported(H, 128, ldZ, Z, E) ? 128 : 0);
  if (nb) {
    double *part = nullptr;
    int rc = sssc_part_buffer(gemm_ety_part_bytes(H, nb, sm_count()), &part);
    if (rc) return rc;
    for (int c0 = 0; c0 < N; c0 += nb) {
      cudaError_t e = launch_gemm_ety(E, Z + c0, ldZ, part, C + c0, ldC, Bc, H, nb, sm_count(), st);
      if (e != cudaSuccess) return set_error(TVO_ECUDA, "launch gemm_ety (sssc): %s", cudaGetErrorString(e));
      count_launch(2);
    }
    return TVO_OK;
  }
  const double one = 1.0;
  TVO_CUBLAS(cublasDgemm(hb, CUBLAS_OP_N, CUBLAS_OP_T, N, H, Bc, &one, Z, (int)ldZ, E, H, &one, C, ldC));
  return TVO_OK;
}
static int sssc_gemm_A(cublasHandle_t h, const float *Wt, int Dpad, const float *Y, int64_t ldY, float *A, int Bc, int H, int D,
                       const double * = nullptr, cudaStream_t = nullptr) {
  const float one = 1.f, zero = 0.f;
  TVO_CUBLAS(cublasSgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, H, Bc, D, &one, Wt, Dpad, Y, (int)ldY, &zero, A, H));
  return TVO_OK;
}

// Global scratch rows for states with kSsscAM < |A| <= kSsscAB: one row per warp of the launch, stream-ordered.
static int sssc_big_alloc(cudaStream_t st, int64_t warps, unsigned char **out) {
  *out = nullptr;
  static bool pool_set[64];
  int dev = 0;
  TVO_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !pool_set[dev]) {  // keep freed blocks in the pool between calls
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    pool_set[dev] = true;
  }
  TVO_CUDA(cudaMallocAsync((void **)out, (size_t)warps * sssc_scratch_bytes(kSsscAB), st));
  return TVO_OK;
}

// Pool of scratch slots for states with more than kSsscAB active units: part of the per-device context, grown on
// demand and kept; the locks are zero between kernels (every taker releases its slot).
static unsigned char *g_sssc_pool[64];
static size_t g_sssc_pool_bytes[64];
static int g_sssc_pool_slots[64];
template <typename P>
static int sssc_pool_attach(P &p, int H) {
  int dev = 0;
  TVO_CUDA(cudaGetDevice(&dev));
  TVO_REQUIRE(dev >= 0 && dev < 64, "sssc pool: device index %d out of range", dev);
  const size_t slot = sssc_pool_slot_bytes(H);
  int slots = (int)max((size_t)4, min((size_t)32, ((size_t)256 << 20) / slot));
  const size_t need = slot * slots + 256;
  if (g_sssc_pool_bytes[dev] < need) {
    if (g_sssc_pool[dev]) TVO_CUDA(cudaFree(g_sssc_pool[dev]));
    g_sssc_pool[dev] = nullptr;
    g_sssc_pool_bytes[dev] = 0;
    TVO_CUDA(cudaMalloc((void **)&g_sssc_pool[dev], need));
    TVO_CUDA(cudaMemset(g_sssc_pool[dev], 0, 256));
    g_sssc_pool_bytes[dev] = need;
    g_sssc_pool_slots[dev] = slots;
  } else {
    slots = min(slots, g_sssc_pool_slots[dev]);
  }
  p.pool_locks = (int *)g_sssc_pool[dev];
  p.pool = g_sssc_pool[dev] + 256;
  p.pool_slots = slots;
  return TVO_OK;
}

template <typename T>
int sssc_prepare(const T *W, const T *sigma2, const T *mus, const T *Psi, const T *pies, int H, int D, void *ws,
                 void *stream) {
  TVO_REQUIRE(W && sigma2 && mus && Psi && pies && ws, "sssc_prepare: null pointer");
  TVO_REQUIRE(H > 0 && D > 0, "sssc_prepare: bad sizes");
  const size_t total = (size_t)H * max(sssc_dpad(D), H);
  sssc_prepare_kernel<T><<<(int)max((size_t)1, min((size_t)4 * sm_count(), (total + 255) / 256)), 256, 0,
                           (cudaStream_t)stream>>>(W, sigma2, mus, Psi, pies, H, D, ws);
  TVO_AFTER_LAUNCH("sssc_prepare");
  return TVO_OK;
}

// scratch: A (cap x H, T) for lpj / E-step; X, E (cap x H dbl each) [+ Yd cap x D dbl for f32] for the M-step
template <typename T>
int sssc_lpj(const T *Y, int64_t ldY, const uint64_t *K, int64_t k_stride_n, const int64_t *idx, const void *ws,
             T *lpj_out, int64_t lpj_stride_n, void *scratch, int64_t scratch_bytes, int *overflow, int64_t B, int S, int H,
             int D, void *stream) {
  TVO_REQUIRE(Y && K && ws && lpj_out && scratch && overflow, "sssc_lpj: null pointer");
  TVO_REQUIRE(B >= 0 && S > 0 && H > 0 && D > 0 && ldY < (1ll << 31), "sssc_lpj: bad sizes");
  const int64_t cap = scratch_bytes / ((int64_t)H * sizeof(T));
  TVO_REQUIRE(cap >= 1, "sssc_lpj: scratch too small");
  if (B == 0) return TVO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  cublasHandle_t hb;
  int rc = sssc_cublas(st, hb);
  if (rc) return rc;
  const SsscTheta<T> th = sssc_theta_view<T>(ws, H, D);
  typename SsscModel<T>::Params p{Y, ldY, (const T *)scratch, ws, H, D, nullptr, overflow, nullptr, nullptr, 0};
  {
    int rcp = sssc_pool_attach(p, H);
    if (rcp) return rcp;
  }
  const size_t msm = align16(SsscModel<T>::smem_bytes(p));
  int wpb = 8;
  while (wpb > 1 && msm * wpb > (size_t)227 * 1024) wpb >>= 1;
  TVO_REQUIRE(msm * wpb <= (size_t)227 * 1024, "sssc_lpj: H=%d too large for shared memory", H);
  auto kern = sssc_lpj_kernel<T>;
  TVO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(msm * wpb)));
  const int CH = S <= 32 ? S : 32;
  const int64_t max_grid = (int64_t)sm_count() * 4;
  rc = sssc_big_alloc(st, max_grid * wpb, &p.big);
  if (rc) return rc;
  for (int64_t c0 = 0; c0 < B; c0 += cap) {
    const int Bc = (int)min(cap, B - c0);
    rc = sssc_gemm_A(hb, th.Wt, th.Dpad, Y + c0 * ldY, ldY, (T *)scratch, Bc, H, D, th.Wg, st);
    if (rc) return rc;
    p.Y = Y + c0 * ldY;
    const int64_t ntask = (int64_t)Bc * ((S + CH - 1) / CH);
    const int grid = (int)max((int64_t)1, min((ntask + wpb - 1) / wpb, max_grid));
    // rows of K / lpj are addressed through idx (or consecutively from c0)
    kern<<<grid, wpb * 32, msm * wpb, st>>>(p, idx ? K : K + c0 * k_stride_n, k_stride_n, idx ? idx + c0 : nullptr,
                                            idx ? lpj_out : lpj_out + c0 * lpj_stride_n, lpj_stride_n, Bc, S, CH, msm);
    TVO_AFTER_LAUNCH("sssc_lpj");
  }
  TVO_CUDA(cudaFreeAsync(p.big, st));
  return TVO_OK;
}

template <typename T>
int sssc_estep(const T *Y, int64_t ldY, const int64_t *idx, uint64_t *K, T *lpj, const void *ws, const tvo_evo_config *cfg,
               const double *sparsity, void *scratch, int64_t scratch_bytes, int *overflow, int64_t B, int S, int H, int D,
               int64_t *nsubs, void *stream) {
  TVO_REQUIRE(Y && K && lpj && ws && cfg && nsubs && scratch && overflow, "estep_evo_sssc: null pointer");
  TVO_REQUIRE(B >= 0 && H > 0 && D > 0 && ldY < (1ll << 31), "estep_evo_sssc: bad sizes");
  const int64_t cap = scratch_bytes / ((int64_t)H * sizeof(T));
  TVO_REQUIRE(cap >= 1, "estep_evo_sssc: scratch too small");
  if (B == 0) return TVO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  cublasHandle_t hb;
  int rc = sssc_cublas(st, hb);
  if (rc) return rc;
  EvoParams ep;
  EstepDims dims;
  rc = evo_params_from_cfg(cfg, S, H, ep, dims, "estep_evo_sssc");
  if (rc) return rc;
  const SsscTheta<T> th = sssc_theta_view<T>(ws, H, D);
  typename SsscModel<T>::Params p{Y, ldY, (const T *)scratch, ws, H, D, nullptr, overflow, nullptr, nullptr, 0};
  {
    int rcp = sssc_pool_attach(p, H);
    if (rcp) return rcp;
  }
  rc = sssc_big_alloc(st, (int64_t)sm_count() * 3 * 8, &p.big);  // estep_evo_fused_kernel: <= 3 CTAs x 8 warps per SM
  if (rc) return rc;
  for (int64_t c0 = 0; c0 < B; c0 += cap) {
    const int Bc = (int)min(cap, B - c0);
    rc = sssc_gemm_A(hb, th.Wt, th.Dpad, Y + c0 * ldY, ldY, (T *)scratch, Bc, H, D, th.Wg, st);
    if (rc) return rc;
    p.Y = Y + c0 * ldY;
    rc = launch_estep_evo_fused<T, SsscModel<T>>(p, idx ? idx + c0 : nullptr, c0, K, lpj, cfg, sparsity, Bc, S, H, nsubs, st,
                                                 "estep_evo_sssc");
    if (rc) break;
  }
  TVO_CUDA(cudaFreeAsync(p.big, st));
  return rc;
}

template <typename T>
int sssc_mstep(const T *Y, int64_t ldY, const int64_t *idx, const uint64_t *K, int64_t k_stride_n, const T *lpj,
               int64_t lpj_stride_n, const void *ws, double *stats, void *scratch, int64_t scratch_bytes, int *overflow,
               int64_t B, int S, int H, int D, int with_ssz, void *stream) {
  TVO_REQUIRE(Y && K && lpj && ws && stats && scratch && overflow, "sssc_mstep: null pointer");
  TVO_REQUIRE(B >= 0 && S > 0 && H > 0 && D > 0 && ldY < (1ll << 31), "sssc_mstep: bad sizes");
  const bool f64 = sizeof(T) == 8;
  const int64_t per = (int64_t)H * sizeof(T) + 2 * (int64_t)H * 8 + (f64 ? 0 : (int64_t)D * 8);
  const int64_t cap = scratch_bytes / per;
  TVO_REQUIRE(cap >= 1, "sssc_mstep: scratch too small (need %lld B per datapoint)", (long long)per);
  if (B == 0) return TVO_OK;
  cudaStream_t st = (cudaStream_t)stream;
  cublasHandle_t hb;
  int rc = sssc_cublas(st, hb);
  if (rc) return rc;
  const SsscTheta<T> th = sssc_theta_view<T>(ws, H, D);
  const int W64 = n_words(H);
  const size_t extra = (size_t)2 * H * 8 + (size_t)S * 8 + (size_t)S * W64 * 8;
  typename SsscModel<T>::Params p{Y, ldY, (const T *)scratch, ws, H, D, nullptr, overflow, nullptr, nullptr, 0};
  {
    int rcp = sssc_pool_attach(p, H);
    if (rcp) return rcp;
  }
  const size_t wbytes = align16(SsscModel<T>::smem_bytes(p) + extra);
  int wpb = 8;
  while (wpb > 1 && wbytes * wpb > (size_t)227 * 1024) wpb >>= 1;
  TVO_REQUIRE(wbytes * wpb <= (size_t)227 * 1024, "sssc_mstep: working set too large for shared memory");
  auto kern = sssc_mstep_kernel<T>;
  TVO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(wbytes * wpb)));
  int occ = 1;
  TVO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpb * 32, wbytes * wpb));
  if (wpb == 8 && occ <= 1 && wbytes * 9 <= (size_t)227 * 1024) {  // one CTA per SM: as many warps as fit (<= 16)
    wpb = (int)min((size_t)16, (size_t)227 * 1024 / wbytes);
    TVO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(wbytes * wpb)));
    occ = 1;
  }
  rc = sssc_big_alloc(st, (int64_t)sm_count() * max(occ, 1) * wpb, &p.big);
  if (rc) return rc;
  const size_t HD = (size_t)H * D, HH = (size_t)H * H;
  const double one = 1.0;
  for (int64_t c0 = 0; c0 < B; c0 += cap) {
    const int Bc = (int)min(cap, B - c0);
    T *A = (T *)scratch;
    double *X = (double *)((unsigned char *)scratch + align16((size_t)cap * H * sizeof(T)));
    double *E = X + (size_t)cap * H;
    double *Yd = f64 ? nullptr : E + (size_t)cap * H;
    rc = sssc_gemm_A(hb, th.Wt, th.Dpad, Y + c0 * ldY, ldY, A, Bc, H, D, th.Wg, st);
    if (rc) return rc;
    p.Y = Y + c0 * ldY;
    const int grid = (int)max((int64_t)1, min((int64_t)(Bc + wpb - 1) / wpb, (int64_t)sm_count() * max(occ, 1)));
    kern<<<grid, wpb * 32, wbytes * wpb, st>>>(p, idx ? idx + c0 : nullptr, c0, K, k_stride_n, lpj, lpj_stride_n, stats, X, E,
                                              Yd, Bc, S, wbytes, (with_ssz & 2) ? 1 : 0);
    TVO_AFTER_LAUNCH("sssc_mstep");
    const double *Yp = f64 ? (const double *)(Y + c0 * ldY) : Yd;
    const int ldy = f64 ? (int)ldY : D;
    // y_szT^T (H x D row-major) += X^T Y                     (sssc.py:446)
    rc = sssc_gemm_EtZ(hb, st, X, Yp, ldy, stats, D, Bc, H, D);
    if (rc) return rc;
    // sz_szT (H x H) += X^T X                                 (sssc.py:443)
    rc = sssc_gemm_EtZ(hb, st, X, X, H, stats + HD, H, Bc, H, H);
    if (rc) return rc;
    // ssz (H x H row-major, [h_s][h_kappa]) += E^T X         (sssc.py:448-450); col-major view: ssz^T = X^T E
    if (with_ssz & 1) {
      rc = sssc_gemm_EtZ(hb, st, E, X, H, stats + HD + 3 * HH, H, Bc, H, H);
      if (rc) return rc;
    }
  }
  TVO_CUDA(cudaFreeAsync(p.big, st));
  return TVO_OK;
}

}  // namespace tvo

using namespace tvo;

extern "C" {
int64_t tvo_sssc_theta_ws_bytes(int32_t H, int32_t D, int32_t is_f64) {
  return (int64_t)(is_f64 ? sssc_ws_bytes<double>(H, D) : sssc_ws_bytes<float>(H, D));
}
int64_t tvo_sssc_stats_len(int32_t H, int32_t D) { return (int64_t)H * D + 4 * (int64_t)H * H + 2 * H + D + 2; }
int64_t tvo_sssc_scratch_bytes(int64_t B, int32_t H, int32_t D, int32_t is_f64) {
  return 16 + B * ((int64_t)H * (is_f64 ? 8 : 4) + 2 * (int64_t)H * 8 + (is_f64 ? 0 : (int64_t)D * 8));
}
#define TVO_SSSC_WRAP(sfx, T)                                                                                              \
  int tvo_sssc_prepare_##sfx(const T *W, const T *sigma2, const T *mus, const T *Psi, const T *pies, int32_t H, int32_t D, \
                             void *ws, void *stream) {                                                                     \
    return sssc_prepare<T>(W, sigma2, mus, Psi, pies, H, D, ws, stream);                                                   \
  }                                                                                                                        \
  int tvo_sssc_lpj_##sfx(const T *Y, int64_t ldY, const uint64_t *K, int64_t k_stride_n, const int64_t *idx,               \
                         const void *ws, T *lpj_out, int64_t lpj_stride_n, void *scratch, int64_t scratch_bytes,           \
                         int32_t *overflow, int64_t B, int32_t S, int32_t H, int32_t D, void *stream) {                    \
    return sssc_lpj<T>(Y, ldY, K, k_stride_n, idx, ws, lpj_out, lpj_stride_n, scratch, scratch_bytes, overflow, B, S, H,   \
                       D, stream);                                                                                         \
  }                                                                                                                        \
  int tvo_estep_evo_sssc_##sfx(const T *Y, int64_t ldY, const int64_t *idx, uint64_t *K, T *lpj, const void *ws,           \
                               const tvo_evo_config *cfg, const double *sparsity, void *scratch, int64_t scratch_bytes,    \
                               int32_t *overflow, int64_t B, int32_t S, int32_t H, int32_t D, int64_t *nsubs,              \
                               void *stream) {                                                                             \
    return sssc_estep<T>(Y, ldY, idx, K, lpj, ws, cfg, sparsity, scratch, scratch_bytes, overflow, B, S, H, D, nsubs,      \
                         stream);                                                                                          \
  }                                                                                                                        \
  int tvo_sssc_mstep_##sfx(const T *Y, int64_t ldY, const int64_t *idx, const uint64_t *K, int64_t k_stride_n,             \
                           const T *lpj, int64_t lpj_stride_n, const void *ws, double *stats, void *scratch,               \
                           int64_t scratch_bytes, int32_t *overflow, int64_t B, int32_t S, int32_t H, int32_t D,           \
                           int32_t with_ssz, void *stream) {                                                               \
    return sssc_mstep<T>(Y, ldY, idx, K, k_stride_n, lpj, lpj_stride_n, ws, stats, scratch, scratch_bytes, overflow, B, S, \
                         H, D, with_ssz, stream);                                                                          \
  }
TVO_SSSC_WRAP(f64, double)
TVO_SSSC_WRAP(f32, float)
}
