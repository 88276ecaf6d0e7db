// tvo_b200 — shared host/device utilities (error reporting, Philox stream, bit helpers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tvo_b200.h"

namespace tvo {

int set_error(int code, const char *fmt, ...);  // capi.cu; returns `code`
void count_launch(int n = 1);
int sm_count();  // cached cudaDevAttrMultiProcessorCount of the current device
// opt-in per-kernel timing (tvo_kernel_timing_enable): event pair on the launching stream
bool kernel_timing_on();
void kernel_timing_begin(int which, cudaStream_t st, void **token);
void kernel_timing_end(void *token, cudaStream_t st);

#define TVO_REQUIRE(cond, ...)                                          \
  do {                                                                  \
    if (!(cond)) return ::tvo::set_error(TVO_EINVAL, __VA_ARGS__);      \
  } while (0)
#define TVO_CUDA(expr)                                                                            \
  do {                                                                                            \
    cudaError_t e_ = (expr);                                                                      \
    if (e_ != cudaSuccess) return ::tvo::set_error(TVO_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)
#define TVO_AFTER_LAUNCH(name)                                                                        \
  do {                                                                                                \
    cudaError_t e_ = cudaGetLastError();                                                              \
    if (e_ != cudaSuccess) return ::tvo::set_error(TVO_ECUDA, "launch %s: %s", name, cudaGetErrorString(e_)); \
    ::tvo::count_launch();                                                                            \
  } while (0)

constexpr unsigned FULL = 0xffffffffu;

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }
__host__ __device__ inline int n_words(int H) { return (H + 63) >> 6; }
__host__ __device__ inline int pad32(int D) { return (D + 31) & ~31; }

// ------------------------------------------------------------------------------------ Philox
// Philox4x32 with kPhiloxRounds = 7 rounds (Salmon et al. SC'11: 7 is the smallest round count that passes BigCrush
// with margin, "Philox4x32-7"; Random123's default of 10 adds safety margin only).  Same stream definition as
// oracle/philox.py — the round count is part of the stream specification.
constexpr int kPhiloxRounds = 7;
// __noinline__: ~75 SASS instructions and a dozen call sites — one shared copy keeps the fused kernels
// inside the instruction cache (ncu: 35% of warp stalls were instruction-fetch misses when inlined)
static __device__ __noinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < kPhiloxRounds; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

enum : uint32_t { PUR_SELECT = 0, PUR_RANDFLIP = 1, PUR_FLIP0 = 2, PUR_FLIP1 = 3, PUR_CROSS = 4, PUR_SAMPLE = 5 };

__host__ __device__ inline uint32_t make_tag(uint32_t step, uint32_t gen, uint32_t purpose) {
  return ((step & 0x7FFFFFu) << 9) | ((gen & 63u) << 3) | (purpose & 7u);
}

struct Stream {  // one (n, tag) family of sub-streams
  uint2 key;
  uint32_t n, tag;
  __device__ __forceinline__ uint4 block(uint32_t slot, uint32_t blk) const {
    return philox4x32_10(make_uint4(blk, slot, n, tag), key);
  }
  __device__ __forceinline__ uint32_t u32(uint32_t slot, uint32_t i) const {
    const uint4 r = block(slot, i >> 2);
    const uint32_t l = i & 3u;
    return l == 0 ? r.x : l == 1 ? r.y : l == 2 ? r.z : r.w;
  }
  __device__ __forceinline__ uint64_t u64(uint32_t slot, uint32_t i) const {
    const uint4 r = block(slot, i >> 1);
    return (i & 1u) ? ((uint64_t)r.w << 32 | r.z) : ((uint64_t)r.y << 32 | r.x);
  }
};

__device__ __forceinline__ uint32_t randint_u32(uint32_t r, uint32_t span) { return __umulhi(r, span); }
__device__ __forceinline__ double unit_double(uint64_t w) { return (double)(w >> 11) * 0x1.0p-53; }

// Threshold T with  U < p  <=>  floor(U*2^64) < T.   flag: 0 = use T, 1 = always, T==0 = never.
__device__ __forceinline__ void prob_to_fixed64(double p, uint64_t &T, uint32_t &always) {
  always = 0;
  T = 0;
  if (!(p > 0.0)) return;
  if (p >= 1.0) {
    always = 1;
    return;
  }
  T = __double2ull_rd(p * 18446744073709551616.0);
}

// 64 independent Bernoulli bits restricted to `und`; lane b's uniform is revealed MSB first from
// u64 stream elements 0,1,2,... of sub-stream `slot` (see oracle/philox.py:bernoulli_mask64).
__device__ __forceinline__ uint64_t bernoulli_mask64(const Stream &st, uint32_t slot, uint64_t T, uint32_t always,
                                                     uint64_t und) {
  if (always) return und;
  uint64_t res = 0;
  for (uint32_t j = 0; j < 64 && und; j += 2) {
    if ((T << j) == 0) break;
    const uint4 r = st.block(slot, j >> 1);
    uint64_t w = (uint64_t)r.y << 32 | r.x;
    if ((T >> (63 - j)) & 1) {
      res |= und & ~w;
      und &= w;
    } else {
      und &= ~w;
    }
    if (!und || (T << (j + 1)) == 0) break;
    w = (uint64_t)r.w << 32 | r.z;
    if ((T >> (62 - j)) & 1) {
      res |= und & ~w;
      und &= w;
    } else {
      und &= ~w;
    }
  }
  return res;
}

// ------------------------------------------------------------------------------------ misc
// order-preserving map of an lpj value to u64 (larger lpj -> larger key; NaN above +inf, as torch.topk)
__device__ __forceinline__ uint64_t sort_key(double x) {
  if (x != x) return ~0ull;
  if (x == 0.0) x = 0.0;  // -0.0 and +0.0 compare equal in torch.topk: one key
  const uint64_t b = (uint64_t)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ uint32_t hash_words(const uint64_t *s, int W64) {
  uint64_t h = 0x9E3779B97F4A7C15ull;
  int w = 0;
  for (; w + 4 <= W64; w += 4) {  // four loads in flight (the states may sit in global memory), mixed in order
    const uint64_t a0 = s[w], a1 = s[w + 1], a2 = s[w + 2], a3 = s[w + 3];
    h = (h ^ a0) * 0xff51afd7ed558ccdull;
    h ^= h >> 32;
    h = (h ^ a1) * 0xff51afd7ed558ccdull;
    h ^= h >> 32;
    h = (h ^ a2) * 0xff51afd7ed558ccdull;
    h ^= h >> 32;
    h = (h ^ a3) * 0xff51afd7ed558ccdull;
    h ^= h >> 32;
  }
  for (; w < W64; ++w) {
    h = (h ^ s[w]) * 0xff51afd7ed558ccdull;
    h ^= h >> 32;
  }
  return (uint32_t)h;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T u = __shfl_xor_sync(FULL, v, o);
    v = u > v ? u : v;
  }
  return v;
}

template <typename T>
struct LowLpj;
template <>
struct LowLpj<double> {
  __device__ static double v() { return (double)TVO_LOW_LPJ_F32; }
};
template <>
struct LowLpj<float> {
  __device__ static float v() { return TVO_LOW_LPJ_F32; }
};

}  // namespace tvo
