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

// ------------------------------------------------------------------------------------ the same factorisation on a cluster
// bsc_chol_factor_cluster_kernel: ONE thread-block cluster of 8 CTAs.  The rows below the diagonal block of a panel
// are dealt to the CTAs in tiles of 8 (tile t -> CTA t mod 8); every CTA stages the 32 panel rows and its own tiles
// (plain L2 loads, all in flight at once), updates them on the FP64 tensor pipe, factors the diagonal block
// REDUNDANTLY (warp 0, rows in registers, the pivot column through shared memory: the one sequential part, ~230
// cycles per pivot), solves its own rows against it (one lane per row) and writes them; one cluster barrier per panel.  CTA 0 also writes the diagonal block
// (the solve kernel inverts the diagonal blocks itself).  What is left of the single-CTA kernel's time is the pivot chain.
constexpr int kChCl = 8;  // CTAs of the cluster
__host__ inline size_t chol_cluster_smem(int H) {
  const int Hp = chol_hp(H);
  return ((size_t)64 * (Hp - 32 + 4) + 64 * 33 + 64) * 8;
}

__global__ void __cluster_dims__(kChCl, 1, 1) __launch_bounds__(256, 1)
bsc_chol_factor_cluster_kernel(const double *__restrict__ U, int H, double *__restrict__ work, double *__restrict__ flags) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Hp = chol_hp(H), np = Hp / 32;
  double *Lg = work, *Lt = work + (size_t)Hp * Hp, *invD = Lt + (size_t)Hp * Hp;
  double *Ls = (double *)smem_raw;                    // 64 staged rows (32 panel rows, then this CTA's tiles), stride c0 + 4
  double *Ps = Ls + (size_t)64 * (Hp - 32 + 4);       // the panel entries of those 64 rows, stride 33
  double *dv = Ps + 64 * 33;                          // reciprocals of the diagonal block's pivots (32), pivot column (32)
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, rank = blockIdx.x % kChCl;
  const int g = lane >> 2, q = lane & 3;
  int info = 0;
  double dmin = INFINITY, dmax = 0.0;
#ifdef TVO_CHOL_TIMING
  long long tpc[6] = {0, 0, 0, 0, 0, 0}, tqc = clock64();
#define TVO_PC(i) { __syncthreads(); long long now = clock64(); tpc[i] += now - tqc; tqc = now; }
#else
#define TVO_PC(i)
#endif
  for (int p = 0; p < np; ++p) {
    const int c0 = 32 * p, R = Hp - c0, ls = c0 + 4;
    const int ntile = (R - 32) >> 3;                                     // tiles of 8 rows below the diagonal block
    const int nmy = rank < ntile ? (ntile - rank + kChCl - 1) / kChCl : 0;  // this CTA's tiles (<= 4)
    // global row of local row lr (0..31: the panel's own rows; 32 + 8 lt + g: tile rank + 8 lt)
    auto grow = [&](int lr) { return lr < 32 ? c0 + lr : c0 + 32 + 8 * (rank + kChCl * ((lr - 32) >> 3)) + (lr & 7); };
    // the entries of A this warp's tile needs (warps 0-3: a tile of the diagonal block, warps 4-7: one of this CTA's
    // tiles): requested first, needed after the MMA loop
    const int lr0 = warp < 4 ? 8 * warp : 32 + 8 * (warp - 4);
    const bool has_tile = warp < 4 || warp - 4 < nmy;
    const int gi = grow(lr0 + g);
    double u[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int gj = c0 + 8 * j + 2 * q + e;
        u[j][e] = (has_tile && gi >= gj && gi < H) ? U[(size_t)gj * H + gi] : (gi == gj ? 1.0 : 0.0);
      }
    // 1. stage: warp w loads local rows 8 w .. 8 w + 7 (up to 56 independent 256-byte reads in flight)
    if (p > 0) {
      double v[8][7];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int lr = 8 * warp + u;
        const bool ok = lr < 32 + 8 * nmy;
        const double *src = Lg + (size_t)(ok ? grow(lr) : 0) * Hp + lane;
#pragma unroll
        for (int m = 0; m < 7; ++m) v[u][m] = (ok && m < p) ? __ldcg(src + 32 * m) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int m = 0; m < 7; ++m)
          if (m < p) Ls[(size_t)(8 * warp + u) * ls + 32 * m + lane] = v[u][m];
    }
    __syncthreads();
    TVO_PC(0)
    // 2. update: warps 0-3 the diagonal block (8 rows each), warps 4-7 this CTA's tiles
    {
      if (has_tile) {
        double acc[4][2], acc2[4][2];  // two independent accumulator sets: eight MMA chains in flight per warp
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc2[j][0] = acc2[j][1] = 0.0;
        const double *Ar = Ls + (size_t)(lr0 + g) * ls + q, *Br = Ls + (size_t)g * ls + q;
#pragma unroll 2
        for (int k0 = 0; k0 < c0; k0 += 8) {  // c0 is a multiple of 32
          const double a = Ar[k0], a2 = Ar[k0 + 4];
          double b[4], b2[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            b[j] = Br[(size_t)(8 * j) * ls + k0];
            b2[j] = Br[(size_t)(8 * j) * ls + k0 + 4];
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            dmma884(acc[j][0], acc[j][1], a, b[j]);
            dmma884(acc2[j][0], acc2[j][1], a2, b2[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[j][0] += acc2[j][0];
          acc[j][1] += acc2[j][1];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int jj = 8 * j + 2 * q + e;
            Ps[(lr0 + g) * 33 + jj] = gi >= c0 + jj ? u[j][e] - acc[j][e] : 0.0;
          }
      }
    }
    __syncthreads();
    TVO_PC(1)
    // 3. every CTA: Cholesky of the diagonal block (warp 0, row `lane` in registers)
    if (warp == 0) {
      // row `lane` in registers; the pivot column goes through shared memory (one STS.64 per lane, then broadcast
      // LDS.128 of two entries at a time) instead of two shuffles per entry: a single warp's issue rate bounds this loop
      double a[32];
      double *col = dv + 32;  // 32 doubles
#pragma unroll
      for (int c = 0; c < 32; ++c) a[c] = Ps[lane * 33 + c];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        col[lane] = a[j];
        __syncwarp();
        const double djj = col[j];
        if (!(djj > 0.0) && info == 0) info = c0 + j + 1;
        const double inv = rsqrt(djj), ljj = djj * inv;
        if (c0 + j < H) {
          dmin = fmin(dmin, ljj);
          dmax = fmax(dmax, ljj);
        }
        if (lane == j) dv[j] = inv;
        const double mine = lane > j ? a[j] * inv : 0.0;  // L[lane][j]
        a[j] = lane == j ? ljj : mine;
        const double mi = -mine * inv;                    // a[c] -= L[lane][j] L[c][j],  L[c][j] = col[c] inv
        if ((j & 1) == 0) a[(j + 1) & 31] = fma(mi, col[(j + 1) & 31], a[(j + 1) & 31]);  // j even: the odd entry j + 1 alone
#pragma unroll
        for (int c = (j + 2) & ~1; c < 32; c += 2) {
          const double2 cc = *reinterpret_cast<const double2 *>(col + c);
          a[c] = fma(mi, cc.x, a[c]);
          a[c + 1] = fma(mi, cc.y, a[c + 1]);
        }
        __syncwarp();
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) Ps[lane * 33 + c] = c <= lane ? a[c] : 0.0;
    }
    __syncthreads();
    TVO_PC(2)
    // 4. this CTA's rows: L[r][:] Lpp^T = P[r][:] by forward substitution
    if (warp == 4 && lane < 8 * nmy) {  // one row per lane (<= 32 rows per CTA), Lpp broadcast from shared memory
      const int lr = 32 + lane;
      double pr[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) pr[c] = Ps[lr * 33 + c];
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const double x = pr[c] * dv[c];
        pr[c] = x;
#pragma unroll
        for (int j = c + 1; j < 32; ++j) pr[j] = fma(-x, Ps[j * 33 + c], pr[j]);
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) Ps[lr * 33 + c] = pr[c];
    }
    __syncthreads();
    TVO_PC(3)
    // 5. write this CTA's rows (CTA 0: the diagonal block as well) to L and L^T
    {
      const int lo = rank == 0 ? 0 : 32, hi = 32 + 8 * nmy;
      for (int e = 32 * lo + t; e < 32 * hi; e += 256) {
        const int lr = e >> 5, j = e & 31;
        Lg[(size_t)grow(lr) * Hp + c0 + j] = Ps[lr * 33 + j];
      }
      for (int j = warp; j < 32; j += 8)
        for (int lr = lo + lane; lr < hi; lr += 32) Lt[(size_t)(c0 + j) * Hp + grow(lr)] = Ps[lr * 33 + j];
    }
    // every thread releases its own global writes at cluster scope; the peers' staging loads follow their acquire
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    TVO_PC(4)
  }
#ifdef TVO_CHOL_TIMING
  if (rank == 0 && t == 0) printf("cluster phases (cycles): stage %lld update %lld diag %lld panel+inv %lld write+barrier %lld\n", tpc[0], tpc[1], tpc[2], tpc[3], tpc[4]);
#endif
  if (rank == 0 && t == 0) {
    flags[0] = (double)info;
    flags[1] = dmin / dmax;
  }
}

// sol (H x D row-major) = my_Wq^-1 B, B = WpT (H x D row-major); one CTA per column d of B.
__host__ inline size_t chol_solve_smem(int H) { return ((size_t)(chol_hp(H) / 32) * 2 * 32 * 33 + 256) * 8; }

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
  double *Lb = xs + 256;                      // the diagonal blocks of L, row stride 33
  const int d = blockIdx.x, i = threadIdx.x, lane = i & 31, warp = i >> 5;
  double b = (i < H) ? B[(size_t)i * D + d] : 0.0;
  if (i >= 32 && i < Hp) chol_prefetch_l1(Lg + (size_t)i * Hp);
  // prologue: warp w inverts diagonal block w (X = Lpp^-1, lane c owns column c; rows of the block staged in shared
  // memory for broadcast reads).  Every CTA does this for itself: ~3 us here against 8 x that on the factorisation's
  // critical path.
  if (warp < nb) {
    double *Lw = Lb + (size_t)warp * 32 * 33, *Xw = Xs + (size_t)warp * 32 * 33;
    const double *src = Lg + (size_t)(32 * warp) * Hp + 32 * warp;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) Lw[r * 33 + lane] = src[(size_t)r * Hp + lane];
    __syncwarp();
    Xw[lane] = 1.0 / Lw[lane * 33 + lane];  // reciprocal pivots, parked in the output block until it is written
    __syncwarp();
    double x[32], rp[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) rp[r] = Xw[r];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      double s0 = (r == lane) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
      for (int k = 0; k < r; ++k) {
        if (k & 1)
          s1 = fma(-Lw[r * 33 + k], x[k], s1);
        else
          s0 = fma(-Lw[r * 33 + k], x[k], s0);
      }
      x[r] = (s0 + s1) * rp[r];
    }
#pragma unroll
    for (int r = 0; r < 32; ++r) Xw[r * 33 + lane] = x[r];
  }
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
