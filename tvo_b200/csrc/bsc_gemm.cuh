// tvo_b200 — hand-written FP64 tensor-core GEMMs of the BSC hot path (sm_100a), fed by TMA bulk copies.
//
//   gemm_yw :  A (B x H) = Y (B x D) W (D x H)  and  yy_b = ||y_b||^2       (bsc.py:105-107, Gram form of the E-step)
//   gemm_ety:  WpT (H x D) += E^T (H x B) Y (B x D)                          (bsc.py:145,153, M-step)
//
// Both run DMMA (mma.sync.m8n8k4.f64: tcgen05 has no f64 kind) from shared-memory tiles that a producer warp fills
// with cp.async.bulk (TMA, SASS: UBLKCP) behind an mbarrier ring; the MMA warps never touch global memory for their
// operands.  Tile layouts are chosen so that every fragment read is a conflict-free LDS.64:
//   * W is kept in the theta workspace as Wg[k/4][h][k%4] (bsc_device.cuh), so a K-chunk of all H columns is ONE
//     contiguous bulk copy and lane T's B fragment B[k = T%4][n = T/4] sits at 8 T bytes;
//   * rows of Y (and E) land in shared memory with a row stride that is 8 words mod 32, so the 8 rows x 4 k of an A
//     fragment spread over all banks per half warp.
#pragma once
#include "bsc_device.cuh"

namespace tvo {

// ------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on `bar` (size: multiple of 16, both 16-byte aligned)
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------ A = Y W, yy = ||y||^2
// CTA tile: 64 datapoints x one PANEL of 32 NB <= 256 columns (16 MMA warps as 4 (M) x 4 (N) since the end of round 2:
// twice the warps per scheduler hide the fragment loads — 82 -> ~90 % of the DGEMM peak; 8 warps as 2 x 4 before) (all H columns for H <= 256; H = 512 / 768 / 1024 —
// BASELINE configs[4] — as 2 / 3 / 4 panels of 256), K in chunks of 16.  4 WM MMA warps as WM (M) x 4 (N): warp tile
// 64 / WM x 8 NB; the last warp is the TMA producer.  Persistent over the (row tile, panel) work items, the panels of a row tile
// adjacent in the order so that the tile of Y is re-read from L2.
constexpr int kGyBM = 64, kGyKC = 16, kGyStages = 4;
constexpr int kGyARow = kGyKC + 4;  // doubles per shared-memory row of the Y chunk: 160 B = 8 words mod 32

__host__ __device__ inline size_t gemm_yw_stage_bytes(int H) {
  return (size_t)kGyBM * kGyARow * 8 + (size_t)kGyKC * H * 8;
}
__host__ __device__ inline size_t gemm_yw_smem_bytes(int H) { return kGyStages * gemm_yw_stage_bytes(H) + 2 * kGyStages * 8 + 16; }

template <int NB, int WM = 4>  // NB n-blocks per warp: panel = 32 NB columns;  WM x 4 MMA warps, 64 / WM rows each
__global__ void __launch_bounds__(32 * (4 * WM + 1), 1)
bsc_gemm_yw_kernel(const double *__restrict__ Y, int64_t ldY, const double *__restrict__ Wg, double *__restrict__ A,
                   double *__restrict__ yy_out, int B, int D, int Hfull) {
  constexpr int H = 32 * NB;  // columns of a panel
  constexpr int MI = kGyBM / WM / 8, NWARP = 4 * WM;  // 8-row blocks per warp; MMA warps
  const int npanel = Hfull / H;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const size_t stage_bytes = gemm_yw_stage_bytes(H);
  uint64_t *full = (uint64_t *)(smem_raw + kGyStages * stage_bytes);
  uint64_t *empty = full + kGyStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = bsc_kpad16(D) / kGyKC;
  const int ntile = ((B + kGyBM - 1) / kGyBM) * npanel;  // work items
  if (threadIdx.x == 0) {
    for (int s = 0; s < kGyStages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, NWARP);
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (warp == NWARP) {
    // ---- producer: per stage 64 row segments of Y (128 B each) + the K-chunk of the panel's columns of Wg (one
    //      contiguous copy when the panel is all of H, else one per k/4 slice) ----
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < ntile; t += gridDim.x) {
      const int r0 = (t / npanel) * kGyBM, rows = min(kGyBM, B - r0), n0 = (t % npanel) * H;
      for (int c = 0; c < nchunk; ++c) {
        mbar_wait(empty + stage, phase ^ 1u);
        double *As = (double *)(smem_raw + stage * stage_bytes);
        double *Bs = As + kGyBM * kGyARow;
        if (lane == 0) mbar_expect_tx(full + stage, (uint32_t)(rows * kGyKC * 8 + kGyKC * H * 8));
        __syncwarp();
        for (int r = lane; r < rows; r += 32)
          tma_bulk_g2s(As + r * kGyARow, Y + (int64_t)(r0 + r) * ldY + c * kGyKC, kGyKC * 8, full + stage);
        if (npanel == 1) {
          if (lane == 0) tma_bulk_g2s(Bs, Wg + (size_t)c * kGyKC * H, kGyKC * H * 8, full + stage);
        } else if (lane < kGyKC / 4) {
          tma_bulk_g2s(Bs + (size_t)lane * H * 4, Wg + ((size_t)(c * (kGyKC / 4) + lane) * Hfull + n0) * 4, H * 4 * 8,
                       full + stage);
        }
        if (++stage == kGyStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    return;
  }
  // ---- MMA warps ----
  const int wm = warp >> 2, wn = warp & 3;  // WM x 4
  const int g = lane >> 2, q = lane & 3;    // fragment row / column group, k index
  int stage = 0;
  uint32_t phase = 0;
  for (int t = blockIdx.x; t < ntile; t += gridDim.x) {
    const int r0 = (t / npanel) * kGyBM, n0 = (t % npanel) * H;
    double acc[MI][NB][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NB; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double ysq[MI];
#pragma unroll
    for (int i = 0; i < MI; ++i) ysq[i] = 0.0;
    for (int c = 0; c < nchunk; ++c) {
      mbar_wait(full + stage, phase);
      const double *As = (const double *)(smem_raw + stage * stage_bytes);
      const double *Bs = As + kGyBM * kGyARow;
#pragma unroll
      for (int k4 = 0; k4 < kGyKC / 4; ++k4) {
        double a[MI], b[NB];
#pragma unroll
        for (int i = 0; i < MI; ++i) a[i] = As[(wm * (8 * MI) + i * 8 + g) * kGyARow + k4 * 4 + q];
#pragma unroll
        for (int j = 0; j < NB; ++j) b[j] = Bs[((size_t)k4 * H + wn * (8 * NB) + j * 8 + g) * 4 + q];
        if (wn == 0) {
#pragma unroll
          for (int i = 0; i < MI; ++i) ysq[i] = fma(a[i], a[i], ysq[i]);
        }
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NB; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + stage);
      if (++stage == kGyStages) {
        stage = 0;
        phase ^= 1u;
      }
    }
    // epilogue: C fragment rows g, columns 2q, 2q+1 of every 8 x 8 block
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      const int row = r0 + wm * (8 * MI) + i * 8 + g;
      if (row < B) {
#pragma unroll
        for (int j = 0; j < NB; ++j)
          *reinterpret_cast<double2 *>(A + (size_t)row * Hfull + n0 + wn * (8 * NB) + j * 8 + 2 * q) =
              make_double2(acc[i][j][0], acc[i][j][1]);
      }
    }
    if (wn == 0 && yy_out && n0 == 0) {
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        double v = ysq[i];
        v += __shfl_xor_sync(FULL, v, 1);
        v += __shfl_xor_sync(FULL, v, 2);
        const int row = r0 + wm * (8 * MI) + i * 8 + g;
        if (q == 0 && row < B) yy_out[row] = v;
      }
    }
  }
}

// host: true if the shape is covered (else the caller uses the library GEMM)
__host__ inline bool gemm_yw_supported(int H, int D, int64_t ldY, const void *Y) {
  return (H == 64 || H == 128 || (H % 256 == 0 && H <= 4096)) && (D % kGyKC) == 0 && (ldY % 2) == 0 && ((uintptr_t)Y % 16) == 0 &&
         gemm_yw_smem_bytes(H > 256 ? 256 : H) <= (size_t)227 * 1024;
}

__host__ inline cudaError_t launch_gemm_yw(const double *Y, int64_t ldY, const double *Wg, double *A, double *yy, int B, int H,
                                           int D, int sms, cudaStream_t st) {
  const int Hp = H > 256 ? 256 : H;  // panel width
  const size_t smem = gemm_yw_smem_bytes(Hp);
  const int grid = min(sms, ((B + kGyBM - 1) / kGyBM) * (H / Hp));
  cudaError_t e;
#define TVO_GYW(NBV)                                                                                          \
  e = cudaFuncSetAttribute(bsc_gemm_yw_kernel<NBV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
  if (e != cudaSuccess) return e;                                                                            \
  bsc_gemm_yw_kernel<NBV><<<grid, 32 * 17, smem, st>>>(Y, ldY, Wg, A, yy, B, D, H);
  if (Hp == 256) {
    TVO_GYW(8)
  } else if (H == 128) {
    TVO_GYW(4)
  } else {
    TVO_GYW(2)
  }
#undef TVO_GYW
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------ WpT += E^T Y
// C (H x D) += sum_b E[b][:]^T Y[b][:]: M = h, N = d, K = b (the datapoints).  The grid is 2 x nslab CTAs: CTA (half,
// slab) accumulates H x D/2 over its contiguous range of datapoints in registers (NW MMA warps x 8 MB rows x all D/2
// columns) and writes the partial to `part[slab][half]`; bsc_gemm_ety_reduce_kernel adds the partials to the statistics
// in slab order (deterministic, unlike a split-K with atomics).  K-chunk: 16 datapoints = 16 rows of E (2 kB each)
// + 16 half rows of Y per stage, one TMA bulk copy per row; row strides are 8 words mod 32 (conflict-free fragments).
constexpr int kGeKC = 16, kGeStages = 4;
__host__ __device__ inline int gemm_ety_erow(int H) { return H + 4; }                 // doubles per E row in shared memory
__host__ __device__ inline int gemm_ety_yrow(int Dh) { return Dh + ((4 - Dh % 16) + 16) % 16; }  // = 4 mod 16 doubles
__host__ __device__ inline size_t gemm_ety_stage_bytes(int H, int Dh) {
  return (size_t)kGeKC * (gemm_ety_erow(H) + gemm_ety_yrow(Dh)) * 8;
}
__host__ __device__ inline size_t gemm_ety_smem_bytes(int H, int Dh) {
  return kGeStages * gemm_ety_stage_bytes(H, Dh) + 2 * kGeStages * 8 + 16;
}

template <int MB, int NBH, int NW = 8>  // NW MMA warps of 8 MB rows each: H = 8 MB NW;  D = 16 NBH
__global__ void __launch_bounds__(32 * NW + 32, 1)
bsc_gemm_ety_kernel(const double *__restrict__ E, const double *__restrict__ Y, int64_t ldY, double *__restrict__ part, int B) {
  constexpr int H = 8 * MB * NW, Dh = 8 * NBH;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int erow = gemm_ety_erow(H), yrow = gemm_ety_yrow(Dh);
  const size_t stage_bytes = gemm_ety_stage_bytes(H, Dh);
  uint64_t *full = (uint64_t *)(smem_raw + kGeStages * stage_bytes);
  uint64_t *empty = full + kGeStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = blockIdx.x & 1, slab = blockIdx.x >> 1, nslab = gridDim.x >> 1;
  // contiguous datapoint range of this slab, in whole K-chunks
  const int nchunk_all = (B + kGeKC - 1) / kGeKC;
  const int c_lo = (int)((int64_t)nchunk_all * slab / nslab), c_hi = (int)((int64_t)nchunk_all * (slab + 1) / nslab);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kGeStages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, NW);
    }
    mbar_fence_init();
  }
  __syncthreads();
  if (warp == NW) {
    int stage = 0;
    uint32_t phase = 0;
    for (int c = c_lo; c < c_hi; ++c) {
      mbar_wait(empty + stage, phase ^ 1u);
      double *Es = (double *)(smem_raw + stage * stage_bytes);
      double *Ys = Es + kGeKC * erow;
      const int b0 = c * kGeKC, rows = min(kGeKC, B - b0);
      if (lane == 0) mbar_expect_tx(full + stage, (uint32_t)(rows * (H + Dh) * 8));
      __syncwarp();
      if (lane < rows) tma_bulk_g2s(Es + lane * erow, E + (size_t)(b0 + lane) * H, H * 8, full + stage);
      if (lane >= 16 && lane - 16 < rows)
        tma_bulk_g2s(Ys + (lane - 16) * yrow, Y + (int64_t)(b0 + lane - 16) * ldY + half * Dh, Dh * 8, full + stage);
      if (++stage == kGeStages) {
        stage = 0;
        phase ^= 1u;
      }
    }
    return;
  }
  const int g = lane >> 2, q = lane & 3;
  double acc[MB][NBH][2];
#pragma unroll
  for (int i = 0; i < MB; ++i)
#pragma unroll
    for (int j = 0; j < NBH; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  int stage = 0;
  uint32_t phase = 0;
  for (int c = c_lo; c < c_hi; ++c) {
    mbar_wait(full + stage, phase);
    const double *Es = (const double *)(smem_raw + stage * stage_bytes);
    const double *Ys = Es + kGeKC * erow;
    const int rows = min(kGeKC, B - c * kGeKC);
#pragma unroll
    for (int k4 = 0; k4 < kGeKC / 4; ++k4) {
      const int kr = k4 * 4 + q;  // this lane's datapoint of the chunk
      const bool ok = kr < rows;  // the tail chunk's missing rows contribute zeros
      double a[MB], b[NBH];
#pragma unroll
      for (int i = 0; i < MB; ++i) a[i] = ok ? Es[kr * erow + warp * (8 * MB) + i * 8 + g] : 0.0;
#pragma unroll
      for (int j = 0; j < NBH; ++j) b[j] = ok ? Ys[kr * yrow + j * 8 + g] : 0.0;
#pragma unroll
      for (int i = 0; i < MB; ++i)
#pragma unroll
        for (int j = 0; j < NBH; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + stage);
    if (++stage == kGeStages) {
      stage = 0;
      phase ^= 1u;
    }
  }
  double *out = part + ((size_t)slab * 2 + half) * H * Dh;
#pragma unroll
  for (int i = 0; i < MB; ++i)
#pragma unroll
    for (int j = 0; j < NBH; ++j)
      *reinterpret_cast<double2 *>(out + (size_t)(warp * (8 * MB) + i * 8 + g) * Dh + j * 8 + 2 * q) =
          make_double2(acc[i][j][0], acc[i][j][1]);
}

// C (H x D, row stride ldC) += sum over slabs, in slab order
static __global__ void bsc_gemm_ety_reduce_kernel(const double *__restrict__ part, double *__restrict__ C, int ldC, int nslab,
                                                  int H, int D) {
  const int Dh = D / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * D) return;
  const int h = i / D, d = i % D, half = d >= Dh ? 1 : 0, dd = d - half * Dh;
  double s = 0.0;
  for (int sl = 0; sl < nslab; ++sl) s += part[((size_t)sl * 2 + half) * H * Dh + (size_t)h * Dh + dd];
  C[(size_t)h * ldC + d] += s;
}

__host__ inline bool gemm_ety_supported(int H, int D, int64_t ldY, const void *Y, const void *E) {
  return H == 256 && (D == 144 || D == 128 || D == 64) && (ldY % 2) == 0 && ((uintptr_t)Y % 16) == 0 &&
         ((uintptr_t)E % 16) == 0 && gemm_ety_smem_bytes(H, D / 2) <= (size_t)227 * 1024;
}
__host__ inline size_t gemm_ety_part_bytes(int H, int D, int sms) { return (size_t)(sms / 2) * 2 * H * (D / 2) * 8; }

// C (H x D, row stride ldC) += E^T (H x B) Y (B x D, row stride ldY);  E is B x H contiguous
__host__ inline cudaError_t launch_gemm_ety(const double *E, const double *Y, int64_t ldY, double *part, double *C, int ldC,
                                            int B, int H, int D, int sms, cudaStream_t st) {
  const size_t smem = gemm_ety_smem_bytes(H, D / 2);
  const int nslab = max(1, min(sms / 2, (B + kGeKC - 1) / kGeKC));
  cudaError_t e;
#define TVO_GETY(MBV, NBV, NWV)                                                                                          \
  e = cudaFuncSetAttribute(bsc_gemm_ety_kernel<MBV, NBV, NWV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
  if (e != cudaSuccess) return e;                                                                                       \
  bsc_gemm_ety_kernel<MBV, NBV, NWV><<<2 * nslab, 32 * NWV + 32, smem, st>>>(E, Y, ldY, part, B);
  if (D == 144) {
    TVO_GETY(2, 9, 16)
  } else if (D == 128) {
    TVO_GETY(2, 8, 16)
  } else {
    TVO_GETY(2, 4, 16)
  }
#undef TVO_GETY
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  bsc_gemm_ety_reduce_kernel<<<(H * D + 255) / 256, 256, 0, st>>>(part, C, ldC, nslab, H, D);
  return cudaGetLastError();
}

}  // namespace tvo
