// tvo_b200 — the W update of BSC.update_param_epoch (bsc.py:177-181: W^T = my_Wq^-1 my_Wp^T) as two hand-written kernels.
//
// my_Wq = sum_n <s s^T>_n is symmetric positive (semi-)definite, so the system is solved by a Cholesky factorisation
// instead of the QR behind torch.linalg.lstsq (the reference's call; it stays the fallback when the factorisation
// fails or its pivots signal an ill-conditioned matrix).  H <= 256 — the sizes where the epoch update is a FIXED
// cost of every EM iteration that does not shrink when the datapoints are sharded over more GPUs (strong scaling):
// the library route (potrf + potrs through torch: ~45 launches, 0.40 ms at H = 256) was 18 % of the EM iteration
// at 125k datapoints per GPU.
//
//   bsc_chol_factor_kernel   ONE CTA.  Blocked left-looking Cholesky, panels of 32 columns: the already factored
//                            columns of the rows below are staged in shared memory by TMA bulk copies (one per row,
//                            one mbarrier), the panel update runs on the FP64 tensor pipe (mma.sync.m8n8k4.f64, 8 x 32
//                            tiles dealt to the 8 warps), warp 0 factors the 32 x 32 diagonal block in registers
//                            (shuffles, fully unrolled), then inverts it for the solve kernel while the other warps
//                            solve their panel rows by substitution.  Writes L, L^T and the inverted diagonal blocks.
//   bsc_chol_solve_kernel    one CTA per right-hand side (a column of my_Wp^T): forward and backward substitution
//                            in blocks of 32 — the diagonal step is a 32 x 32 mat-vec with the inverted block
//                            (no dependent chain), the update a contiguous 256-byte row segment per thread.
// Both are deterministic (fixed summation order), so every rank of a data-parallel run computes bit-identical W from
// the all-reduced statistics and no theta broadcast is needed (the library solve was not guaranteed to be).
#pragma once
#include "bsc_gemm.cuh"  // mbarrier / TMA bulk copy / DMMA helpers

namespace tvo {

__host__ __device__ inline int chol_hp(int H) { return (H + 31) & ~31; }
// work (doubles): L (Hp x Hp row-major, lower) | Lt (Hp x Hp, Lt[k][i] = L[i][k]) | invD (Hp/32 blocks of 32 x 32)
__host__ __device__ inline size_t chol_work_doubles(int H) {
  const size_t Hp = (size_t)chol_hp(H);
  return 2 * Hp * Hp + (Hp / 32) * 1024;
}
__host__ inline size_t chol_factor_smem(int H) {
  const int Hp = chol_hp(H);
  size_t stage = 0;
  for (int c0 = 32; c0 < Hp; c0 += 32) stage = max(stage, (size_t)(Hp - c0) * (c0 + 4));
  return (stage + (size_t)Hp * 33 + 32 + 2) * 8;
}

// U: the accumulated UPPER triangle of my_Wq (H x H row-major, U[r][c] valid for r <= c).
// flags[0] = 0, or 1 + index of the first non-positive pivot (the solve kernel sets -1 when the solution is not
// finite);  flags[1] = min_h L_hh / max_h L_hh.
__global__ void __launch_bounds__(256, 1)
bsc_chol_factor_kernel(const double *__restrict__ U, int H, double *__restrict__ work, double *__restrict__ flags) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Hp = chol_hp(H), np = Hp / 32;
  double *Lg = work, *Lt = work + (size_t)Hp * Hp, *invD = Lt + (size_t)Hp * Hp;
  size_t stage_max = 0;
  for (int c0 = 32; c0 < Hp; c0 += 32) stage_max = max(stage_max, (size_t)(Hp - c0) * (c0 + 4));
  double *Ls = (double *)smem_raw;   // staged L[c0:Hp, 0:c0], row stride c0 + 4 doubles (8 words mod 32: conflict-free fragments)
  double *Ps = Ls + stage_max;       // the panel, rows c0..Hp-1, row stride 33
  double *dv = Ps + (size_t)Hp * 33;  // reciprocals of the diagonal block's pivots
  uint64_t *bar = (uint64_t *)(dv + 32);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  int info = 0;
  double dmin = INFINITY, dmax = 0.0;
#ifdef TVO_CHOL_TIMING
  long long tph[6] = {0, 0, 0, 0, 0, 0}, tq = clock64();
#define TVO_PH(i) { __syncthreads(); long long now = clock64(); tph[i] += now - tq; tq = now; }
#else
#define TVO_PH(i)
#endif
  for (int p = 0; p < np; ++p) {
    const int c0 = 32 * p, R = Hp - c0, ls = c0 + 4;
    // 1. stage the factored columns of the rows at and below the panel: one TMA bulk copy per row (thread r copies
    //    row r), all in flight at once behind one mbarrier — the L2 round trip is paid once per panel
    if (p > 0) {
      if (t == 0) mbar_expect_tx(bar, (uint32_t)(R * c0 * 8));
      __syncthreads();
      if (t < R) tma_bulk_g2s(Ls + (size_t)t * ls, Lg + (size_t)(c0 + t) * Hp, (uint32_t)(c0 * 8), bar);
      mbar_wait(bar, (uint32_t)((p - 1) & 1));
    }
    TVO_PH(0)
    // 2. panel = A[c0:, c0:c0+32] - L[c0:, :c0] L[c0:c0+32, :c0]^T on the FP64 tensor pipe (mma.sync.m8n8k4.f64): tiles of
    //    8 rows x 32 columns dealt round-robin to the 8 warps (all four FP64 pipes busy down to the last panel), two
    //    tiles per pass for independent MMA chains
    {
      const int g = lane >> 2, q = lane & 3, ntile = R >> 3;
      const double *Br = Ls + (size_t)g * ls + q;
      for (int tile0 = warp; tile0 < ntile; tile0 += 16) {
        const bool two = tile0 + 8 < ntile;
        const int rowA = 8 * tile0 + g, rowB = two ? rowA + 64 : rowA;
        double acc[2][4][2], u[2][4][2];  // u: the entries of A, requested now, needed after the MMA loop
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int gi = c0 + (i == 0 ? rowA : rowB);
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int gj = c0 + 8 * j + 2 * q + e;
              acc[i][j][e] = 0.0;
              u[i][j][e] = (gi >= gj && gi < H) ? U[(size_t)gj * H + gi] : (gi == gj ? 1.0 : 0.0);  // gj <= gi < H, or padding
            }
        }
        const double *ArA = Ls + (size_t)rowA * ls + q, *ArB = Ls + (size_t)rowB * ls + q;
#ifdef TVO_CHOL_TIMING
        const long long tk = clock64();
#endif
#pragma unroll 4
        for (int k0 = 0; k0 < c0; k0 += 4) {
          double b[4];
          const double aA = ArA[k0], aB = ArB[k0];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = Br[(size_t)(8 * j) * ls + k0];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            dmma884(acc[0][j][0], acc[0][j][1], aA, b[j]);
            dmma884(acc[1][j][0], acc[1][j][1], aB, b[j]);
          }
        }
#ifdef TVO_CHOL_TIMING
        tph[5] += clock64() - tk;
#endif
        // C fragment: row g, columns 2q and 2q + 1 of every 8 x 8 block
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (i == 1 && !two) break;
          const int r = i == 0 ? rowA : rowB, gi = c0 + r;
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int jj = 8 * j + 2 * q + e, gj = c0 + jj;
              Ps[r * 33 + jj] = gi >= gj ? u[i][j][e] - acc[i][j][e] : 0.0;
            }
        }
      }
    }
    __syncthreads();
    TVO_PH(1)
    // 3. warp 0: Cholesky of the 32 x 32 diagonal block, row `lane` in registers (shuffles, fully unrolled)
    if (warp == 0) {
      double a[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) a[c] = Ps[lane * 33 + c];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const double djj = __shfl_sync(FULL, a[j], j);
        if (!(djj > 0.0) && info == 0) info = c0 + j + 1;
        const double inv = rsqrt(djj), ljj = djj * inv;
        if (c0 + j < H) {
          dmin = fmin(dmin, ljj);
          dmax = fmax(dmax, ljj);
        }
        if (lane == j) dv[j] = inv;
        a[j] = lane == j ? ljj : (lane > j ? a[j] * inv : 0.0);
#pragma unroll
        for (int c = j + 1; c < 32; ++c) {
          const double lcj = __shfl_sync(FULL, a[j], c);
          a[c] = fma(-a[j], lcj, a[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) Ps[lane * 33 + c] = c <= lane ? a[c] : 0.0;
    }
    __syncthreads();
    TVO_PH(2)
    // 4. warps 1-7: the panel rows below the diagonal block, L[r][:] Lpp^T = P[r][:] by forward substitution (one row
    //    per thread, Lpp broadcast from shared memory);  warp 0 meanwhile: X = Lpp^-1 for the solve kernel (lane c owns
    //    column c, x[i] = (delta_ic - sum_{k<i} L[i][k] x[k]) / L[i][i])
    if (warp == 0) {
      double x[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        double s = (i == lane) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) s = fma(-Ps[i * 33 + k], x[k], s);
        x[i] = s * dv[i];
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) invD[(size_t)p * 1024 + i * 32 + lane] = x[i];
    } else {
      const int r = t;  // rows 32 .. 255 of the panel
      if (r < R) {
        double pr[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) pr[c] = Ps[r * 33 + c];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const double x = pr[c] * dv[c];
          pr[c] = x;
#pragma unroll
          for (int j = c + 1; j < 32; ++j) pr[j] = fma(-x, Ps[j * 33 + c], pr[j]);
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) Ps[r * 33 + c] = pr[c];
      }
    }
    __syncthreads();
    TVO_PH(3)
    // 5. write the panel: L (rows contiguous) and L^T (columns of the panel as rows)
    for (int e = t; e < R * 32; e += 256) {
      const int r = e >> 5, j = e & 31;
      Lg[(size_t)(c0 + r) * Hp + c0 + j] = Ps[r * 33 + j];
    }
    for (int j = warp; j < 32; j += 8)
      for (int r = lane; r < R; r += 32) Lt[(size_t)(c0 + j) * Hp + c0 + r] = Ps[r * 33 + j];
    asm volatile("fence.proxy.async;" ::: "memory");  // these global writes are read back by the next panel's bulk copies
    __syncthreads();
    TVO_PH(4)
  }
#ifdef TVO_CHOL_TIMING
  if (t == 0) printf("chol phases (cycles): stage %lld update %lld (k-loops of warp 0: %lld) diag %lld panel+inv %lld write %lld\n", tph[0], tph[1], tph[5], tph[2], tph[3], tph[4]);
#endif
  if (t == 0) {  // warp 0 holds the pivots (every lane saw every one)
    flags[0] = (double)info;
    flags[1] = dmin / dmax;
  }
}

// sol (H x D row-major) = my_Wq^-1 B, B = WpT (H x D row-major); one CTA per column d of B.
__host__ inline size_t chol_solve_smem(int H) { return ((size_t)(chol_hp(H) / 32) * 32 * 33 + 256) * 8; }

__device__ __forceinline__ void chol_prefetch_l1(const double *p) {  // a 256-byte row segment = two lines
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p + 16));
}

__global__ void __launch_bounds__(256, 1)
bsc_chol_solve_kernel(const double *__restrict__ work, const double *__restrict__ B, double *__restrict__ sol,
                      double *__restrict__ flags, int H, int D) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Hp = chol_hp(H), nb = Hp / 32;
  const double *Lg = work, *Lt = work + (size_t)Hp * Hp, *invD = Lt + (size_t)Hp * Hp;
  double *Xs = (double *)smem_raw;            // the inverted diagonal blocks, row stride 33
  double *xs = Xs + (size_t)nb * 32 * 33;     // the solved entries
  const int d = blockIdx.x, i = threadIdx.x, lane = i & 31, warp = i >> 5;
  for (int e = i; e < nb * 1024; e += 256) Xs[(e >> 5) * 33 + (e & 31)] = invD[e];
  double b = (i < H) ? B[(size_t)i * D + d] : 0.0;
  if (i >= 32 && i < Hp) chol_prefetch_l1(Lg + (size_t)i * Hp);
  __syncthreads();
  // ---- forward: L y = b ----
  for (int blk = 0; blk < nb; ++blk) {
    double seg[32];  // this thread's row segment of L under the block (rows below it only)
    const bool below = i >= (blk + 1) * 32 && i < Hp;
    if (below) {
      const double2 *src = reinterpret_cast<const double2 *>(Lg + (size_t)i * Hp + blk * 32);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const double2 v = src[k];
        seg[2 * k] = v.x;
        seg[2 * k + 1] = v.y;
      }
      if (i >= (blk + 2) * 32) chol_prefetch_l1(Lg + (size_t)i * Hp + (blk + 1) * 32);
    }
    if (blk == nb - 1 && i < Hp - 32) chol_prefetch_l1(Lt + (size_t)i * Hp + (nb - 1) * 32);  // first backward segment
    if (warp == blk) {  // y_blk = X b_blk
      const double *ir = Xs + (size_t)(blk * 32 + lane) * 33;
      double y = 0.0;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const double bk = __shfl_sync(FULL, b, k);
        y = fma(ir[k], bk, y);  // X[lane][k] = 0 for k > lane
      }
      b = y;
      xs[i] = y;
    }
    __syncthreads();
    if (below) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 32; ++k) s = fma(seg[k], xs[blk * 32 + k], s);
      b -= s;
    }
  }
  __syncthreads();
  // ---- backward: L^T x = y ----
  for (int blk = nb - 1; blk >= 0; --blk) {
    double seg[32];
    const bool above = i < blk * 32;
    if (above) {
      const double2 *src = reinterpret_cast<const double2 *>(Lt + (size_t)i * Hp + blk * 32);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const double2 v = src[k];
        seg[2 * k] = v.x;
        seg[2 * k + 1] = v.y;
      }
      if (i < (blk - 1) * 32) chol_prefetch_l1(Lt + (size_t)i * Hp + (blk - 1) * 32);
    }
    if (warp == blk) {  // x_blk = X^T y_blk
      const double *ic = Xs + (size_t)blk * 32 * 33 + lane;
      double x = 0.0;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const double yk = __shfl_sync(FULL, b, k);
        x = fma(ic[k * 33], yk, x);  // X[k][lane] = 0 for k < lane
      }
      b = x;
      xs[i] = x;
    }
    __syncthreads();
    if (above) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 32; ++k) s = fma(seg[k], xs[blk * 32 + k], s);
      b -= s;
    }
  }
  if (i < H) {
    sol[(size_t)i * D + d] = b;
    if (!isfinite(b)) flags[0] = -1.0;  // non-finite statistics: the caller takes the reference's fallback route
  }
}

__host__ inline bool chol_supported(int H) { return H >= 1 && H <= 256; }

}  // namespace tvo
