// tvo_b200 — fp32 mode of the Gram-form E-step:  A (B x H) = Y (B x D) W (D x H)  and  yy_b = ||y_b||^2  on the
// 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM), fed by TMA bulk copies.
//
// fp32 accuracy from TF32 products (3xTF32): every fp32 operand x is split as x = hi + lo with hi = x truncated to
// TF32 (the tensor core reads the upper 19 bits of a 32-bit word) and lo = x - hi (exact in fp32); then
//     A B  ~=  A_hi B_hi + A_hi B_lo + A_lo B_hi          (the dropped lo x lo term is ~2^-22 relative)
// accumulated in fp32 in TMEM — three MMAs per K step instead of one fp32 FMA chain per output.
//
// CTA = 6 warps, persistent over 128-row tiles of Y, K in blocks of 32 (one 128-byte swizzle span):
//   warp 4     TMA producer: per stage 128 row segments of Y (raw fp32) and the two pre-split, pre-swizzled images of the
//              W block (bsc_prepare writes them into the theta workspace), all as cp.async.bulk on one mbarrier;
//   warps 0-3  split the raw Y block into the hi tile and (in place) the lo tile in the canonical K-major SWIZZLE_128B
//              layout, accumulate ||y||^2, fence.proxy.async, arrive; after the tile's last block they are the epilogue:
//              tcgen05.ld of their 32 TMEM lanes, 128-byte row stores to global memory;
//   warp 5     one lane issues the tcgen05.mma (M = 128, N = H, K = 8 per instruction) from shared-memory
//              descriptors and commits to the stage's "free" mbarrier / the tile's "accumulator full" mbarrier.
#pragma once
#include "bsc_gemm.cuh"

namespace tvo {

constexpr int kTfBM = 128, kTfBK = 32, kTfStages = 2;

__host__ __device__ inline int tf32_nkb(int D) { return (D + kTfBK - 1) / kTfBK; }
// bytes of the W images in the theta workspace: per K block a hi and a lo tile of H rows x 128 B
__host__ __device__ inline size_t tf32_w_bytes(int H, int D) { return (size_t)tf32_nkb(D) * 2 * H * 128; }
// byte offset of (row, byte b of its 128-byte K span) in a K-major SWIZZLE_128B tile (8-row groups of 1024 B)
__host__ __device__ inline uint32_t sw128(uint32_t row, uint32_t b) {
  return (row >> 3) * 1024u + (row & 7u) * 128u + ((((b >> 4) ^ (row & 7u)) << 4) | (b & 15u));
}
__host__ __device__ inline size_t tf32_stage_bytes(int H) { return (size_t)2 * kTfBM * 128 + (size_t)2 * H * 128; }
__host__ __device__ inline size_t tf32_smem_bytes(int H) { return kTfStages * tf32_stage_bytes(H) + 1024 + 256; }

// W (D x H row-major, fp32) -> split + swizzled images, run by bsc_prepare for the fp32 model
static __global__ void bsc_w_tf32_images_kernel(const float *__restrict__ W, unsigned char *__restrict__ img, int H, int D) {
  const int nkb = tf32_nkb(D);
  const size_t total = (size_t)nkb * H * kTfBK;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % kTfBK), n = (int)((i / kTfBK) % H), kb = (int)(i / ((size_t)kTfBK * H));
    const int d = kb * kTfBK + k;
    const float v = d < D ? W[(size_t)d * H + n] : 0.f;
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    const float lo = v - hi;
    unsigned char *base = img + (size_t)kb * 2 * H * 128;
    *(float *)(base + sw128((uint32_t)n, (uint32_t)k * 4)) = hi;
    *(float *)(base + (size_t)H * 128 + sw128((uint32_t)n, (uint32_t)k * 4)) = lo;
  }
}

// shared-memory matrix descriptor of a K-major SWIZZLE_128B tile (cute::UMMA::SmemDescriptor): start address >> 4,
// leading byte offset 1 (unused for swizzled K-major), stride byte offset 1024 >> 4 between 8-row groups, version 1,
// layout type 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A and B TF32, both K-major, N >> 3, M >> 4
__host__ __device__ inline uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int N>  // N = H (64, 128 or 256): the MMA's N and the number of TMEM columns
__global__ void __launch_bounds__(192, 1)
bsc_gemm_yw_tf32_kernel(const float *__restrict__ Y, int64_t ldY, const unsigned char *__restrict__ Wimg, float *__restrict__ A,
                        double *__restrict__ yy_out, int B, int D) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char *smem = (unsigned char *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B tiles: 1024-byte aligned
  const size_t stage_bytes = tf32_stage_bytes(N);
  uint64_t *bars = (uint64_t *)(smem + kTfStages * stage_bytes);
  uint64_t *raw_full = bars, *a_ready = bars + kTfStages, *stage_free = bars + 2 * kTfStages;
  uint64_t *acc_full = bars + 3 * kTfStages, *acc_empty = acc_full + 1;
  uint32_t *tmem_slot = (uint32_t *)(acc_empty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = tf32_nkb(D), ntile = (B + kTfBM - 1) / kTfBM;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kTfStages; ++s) {
      mbar_init(raw_full + s, 1);
      mbar_init(a_ready + s, 128);
      mbar_init(stage_free + s, 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 128);
    mbar_fence_init();
  }
  if (warp == 5) {  // TMEM: N fp32 accumulator columns, allocated (and later freed) by this warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ---------------------------------------------------------------- TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < ntile; t += gridDim.x) {
      const int r0 = t * kTfBM, rows = min(kTfBM, B - r0);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(stage_free + stage, phase ^ 1u);
        unsigned char *st = smem + stage * stage_bytes;
        unsigned char *raw = st + kTfBM * 128;      // becomes the lo tile
        unsigned char *bimg = st + 2 * kTfBM * 128;  // hi image, then lo image
        const int kcols = min(kTfBK, D - kb * kTfBK);
        if (lane == 0) mbar_expect_tx(raw_full + stage, (uint32_t)(rows * kcols * 4 + 2 * N * 128));
        __syncwarp();
        for (int r = lane; r < rows; r += 32)
          tma_bulk_g2s(raw + r * 128, Y + (int64_t)(r0 + r) * ldY + kb * kTfBK, (uint32_t)kcols * 4, raw_full + stage);
        if (lane == 0) tma_bulk_g2s(bimg, Wimg + (size_t)kb * 2 * N * 128, 2 * N * 128, raw_full + stage);
        if (++stage == kTfStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 5) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idesc = umma_idesc_tf32(kTfBM, N);
    int stage = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int t = blockIdx.x; t < ntile; t += gridDim.x) {
      mbar_wait(acc_empty, acc_phase ^ 1u);  // the epilogue has drained the previous tile's accumulator
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(raw_full + stage, phase);
        mbar_wait(a_ready + stage, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t st = smem_u32(smem + stage * stage_bytes);
          const uint32_t a_hi = st, a_lo = st + kTfBM * 128, b_hi = st + 2 * kTfBM * 128, b_lo = b_hi + N * 128;
          const int ksteps = min(kTfBK, D - kb * kTfBK) / 8;
          for (int k = 0; k < ksteps; ++k) {
            const uint32_t o = (uint32_t)k * 32u;  // 8 TF32 along K inside the 128-byte swizzle span
            umma_tf32(tmem, umma_desc_sw128(a_hi + o), umma_desc_sw128(b_hi + o), idesc, (kb | k) ? 1u : 0u);
            umma_tf32(tmem, umma_desc_sw128(a_hi + o), umma_desc_sw128(b_lo + o), idesc, 1u);
            umma_tf32(tmem, umma_desc_sw128(a_lo + o), umma_desc_sw128(b_hi + o), idesc, 1u);
          }
          umma_commit(stage_free + stage);  // the stage's buffers are free once these MMAs have read them
          if (kb == nkb - 1) umma_commit(acc_full);
        }
        __syncwarp();
        if (++stage == kTfStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      acc_phase ^= 1u;
    }
  } else {
    // ---------------------------------------------------------------- split (hi / lo tiles) + epilogue, warps 0-3
    const int tid = threadIdx.x;         // 0..127
    const int c = tid & 7, rsub = tid >> 3;  // 16-byte chunk of the 128-byte row; row within a 16-row pass
    int stage = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int t = blockIdx.x; t < ntile; t += gridDim.x) {
      const int r0 = t * kTfBM;
      double ysq[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(raw_full + stage, phase);
        unsigned char *st = smem + stage * stage_bytes;
        unsigned char *hi_t = st, *raw = st + kTfBM * 128;
        const int kcols = min(kTfBK, D - kb * kTfBK);
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 16 + rsub;
          v[i] = *reinterpret_cast<const float4 *>(raw + r * 128 + c * 16);
          if (4 * c >= kcols || r0 + r >= B) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);  // beyond D / beyond the batch
        }
        __syncwarp();  // a row's 8 chunks are held by 8 lanes of this warp: all read before the in-place writes
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 16 + rsub;
          const float4 x = v[i];
          ysq[i] += (double)x.x * x.x + (double)x.y * x.y + (double)x.z * x.z + (double)x.w * x.w;
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
          h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
          h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
          h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
          l = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
          const uint32_t o = sw128((uint32_t)r, (uint32_t)c * 16);
          *reinterpret_cast<float4 *>(hi_t + o) = h;
          *reinterpret_cast<float4 *>(raw + o) = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
        mbar_arrive(a_ready + stage);
        if (++stage == kTfStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      // ||y||^2: the 8 chunk lanes of a row
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        double s = ysq[i];
        s += __shfl_xor_sync(FULL, s, 1);
        s += __shfl_xor_sync(FULL, s, 2);
        s += __shfl_xor_sync(FULL, s, 4);
        const int row = r0 + i * 16 + rsub;
        if (c == 0 && row < B) yy_out[row] = s;
      }
      // epilogue: this warp's 32 TMEM lanes = rows 32 warp .. 32 warp + 31 of the tile
      mbar_wait(acc_full, acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = r0 + warp * 32 + lane;
#pragma unroll 1
      for (int cb = 0; cb < N / 32; ++cb) {
        uint32_t acc[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cb * 32), acc);
        if (row < B) {
          float4 *dst = reinterpret_cast<float4 *>(A + (size_t)row * N + cb * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            dst[q] = make_float4(__uint_as_float(acc[4 * q]), __uint_as_float(acc[4 * q + 1]), __uint_as_float(acc[4 * q + 2]),
                                 __uint_as_float(acc[4 * q + 3]));
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(acc_empty);
      acc_phase ^= 1u;
    }
  }
  __syncthreads();
  if (warp == 5) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)N) : "memory");
}

__host__ inline bool gemm_yw_tf32_supported(int H, int D, int64_t ldY, const void *Y) {
  return (H == 64 || H == 128 || H == 256) && (D % 8) == 0 && (ldY % 4) == 0 && ((uintptr_t)Y % 16) == 0 &&
         tf32_smem_bytes(H) <= (size_t)227 * 1024;
}

__host__ inline cudaError_t launch_gemm_yw_tf32(const float *Y, int64_t ldY, const unsigned char *Wimg, float *A, double *yy, int B,
                                                int H, int D, int sms, cudaStream_t st) {
  const size_t smem = tf32_smem_bytes(H);
  const int grid = min(sms, (B + kTfBM - 1) / kTfBM);
  cudaError_t e;
#define TVO_GTF(NV)                                                                                                  \
  e = cudaFuncSetAttribute(bsc_gemm_yw_tf32_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
  if (e != cudaSuccess) return e;                                                                                    \
  bsc_gemm_yw_tf32_kernel<NV><<<grid, 192, smem, st>>>(Y, ldY, Wimg, A, yy, B, D);
  if (H == 256) {
    TVO_GTF(256)
  } else if (H == 128) {
    TVO_GTF(128)
  } else {
    TVO_GTF(64)
  }
#undef TVO_GTF
  return cudaGetLastError();
}

}  // namespace tvo
