// tvo_b200 — list-based fused EVO E-step for BSC (EVOVariationalStates.update, evo.py:80-115): the fast path for
// sparse states, 64 <= H <= 256, randparents + sparseflip.
//
// TVO states are sparse bit vectors (|s| ~ 1-6 of H).  Inside this kernel a state is not a bit row but a SORTED
// POSITION LIST: 15 ascending one-byte positions + the count in byte 15 = one uint4 (`PosList`).  With that
// representation
//   * the old states are turned into lists once (one lane per state, one find-first-set walk) and scored in the same
//     walk:  lpj = coef ||y||^2 + sum_{h in s} (gp_h - 2 coef a_h) + sum_{h<h' in s} G2_hh'   (bsc.py:100-119 in Gram
//     form; a = W^T y from the chunk GEMM, gp / G2 from the theta workspace);
//   * a sparseflip child is (parent list, removed entries, added positions): the mutation draws positions, so no bit
//     row is ever searched, and the child's lpj is the parent's plus the flipped terms
//         lpj(s') = lpj(s) - sum_{h in R} d_h - sum_{h in R, j in K} G2_hj - sum_{{h,h'} in R} G2_hh'
//                          + sum_{z in A} d_z + sum_{z in A, j in K} G2_zj + sum_{{z,z'} in A} G2_zz'
//     (R removed, A added, K kept units, d_h = gp_h - 2 coef a_h) — |R u A| (|K| + 1) gathers instead of |s'|^2 / 2;
//   * duplicates are list-equal (canonical: ascending, zero padded), the dedup hash is three multiplies;
//   * bit rows are rebuilt only for the S states that are written back.
// Same random stream, same draws and the same results as the generic kernel (estep_device.cuh) and the oracle:
// candidate selection, flip counts, rejection order and the n-th-remaining rule are restated on lists.
//
// Datapoints the lists cannot hold — a state or child with more than 15 active units,
// a NaN in y (nansum form), the (practically unreachable) rejection cap — are appended to an overflow list and
// processed by the generic kernel in a second launch; nothing has been written for them.
#pragma once
#include "bsc_device.cuh"

namespace tvo {

constexpr int kV2MaxList = 15;  // positions per list
constexpr int kV2MaxAdd = 15;   // added units per child (sorted positions in a 16-entry shared-memory slot of the lane)

// Position type PT: one byte for H <= 256 (a list = one uint4), two bytes for 256 < H <= 1024 (BASELINE configs[4]:
// a list = two uint4).  Entry 15 holds the count either way.
template <typename PT>
struct PLT;
template <>
struct PLT<unsigned char> {
  static constexpr int kBytes = 16;
  static constexpr uint32_t kNone = 256u;  // larger than any position
};
template <>
struct PLT<unsigned short> {
  static constexpr int kBytes = 32;
  static constexpr uint32_t kNone = 65536u;
};
__host__ __device__ inline int v2_list_bytes(int H) { return H <= 256 ? 16 : 32; }
// d_h = gp_h - 2 coef a_h staged in shared memory for all H units (H <= 256), or gathered from the a row on demand
// (H > 256: a datapoint touches a few hundred of the 1024 units, staging would cost 8 kB per warp and halve the
// resident warps)
__host__ __device__ inline bool v2_stage_d(int H) { return H <= 256; }

struct V2Dims {
  int S, Snew, Sgen, P, C, G, W64, H;
  // filled by v2_finish(): per-warp shared-memory layout (bytes) and the 16-bit magic of j / C (exact for j < 256)
  int off_lpj, off_hash, off_iscr, off_d, warp_bytes, cinv;
};

// hash region: the dedup hashes, afterwards the survivors of the merge (8-byte composite keys or 4-byte indices)
__host__ __device__ inline size_t v2_hash_bytes(const V2Dims &d) {
  const size_t h = (size_t)hash_len(d.S, d.Snew) * 4, k = (size_t)d.Snew * 8;
  const size_t ab = (size_t)32 * v2_list_bytes(d.H);  // during mutation: one slot of added positions per lane
  const size_t m = h > k ? h : k;
  return align16(m > ab ? m : ab);
}
__host__ __device__ inline size_t v2_warp_bytes(const V2Dims &d) {
  const int nI = d.P > d.S ? d.P : d.S;
  return (size_t)(d.S + d.Snew) * v2_list_bytes(d.H) + align16((size_t)(d.S + d.Snew) * 8) + v2_hash_bytes(d) +
         align16((size_t)nI * 4) + (v2_stage_d(d.H) ? align16((size_t)d.H * 8) : 0);
}

__host__ inline void v2_finish(V2Dims &d) {
  d.off_lpj = (d.S + d.Snew) * v2_list_bytes(d.H);
  d.off_hash = d.off_lpj + (int)align16((size_t)(d.S + d.Snew) * 8);
  d.off_iscr = d.off_hash + (int)v2_hash_bytes(d);
  d.off_d = d.off_iscr + (int)align16((size_t)(d.P > d.S ? d.P : d.S) * 4);
  d.warp_bytes = (int)v2_warp_bytes(d);
  d.cinv = (65536 + d.C - 1) / d.C;
}

__host__ inline bool v2_supported(const EvoParams &ep, const V2Dims &d) {
  return ep.sel == TVO_SEL_RANDPARENTS && ep.mut == TVO_MUT_SPARSEFLIP && ep.flip_tab != nullptr && d.H >= 64 &&
         d.H <= 1024 && d.S <= 128 && d.Sgen <= 128 && d.Snew <= 128;
}

// ---- position lists ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool pl_equal(const uint4 &a, const uint4 &b) {
  return ((a.x ^ b.x) | (a.y ^ b.y) | (a.z ^ b.z) | (a.w ^ b.w)) == 0u;
}
__device__ __forceinline__ uint32_t pl_mix(const uint4 &l) {
  return (l.x * 0x9E3779B1u) ^ (l.y * 0x85EBCA77u) ^ (l.z * 0xC2B2AE3Du) ^ (l.w * 0x27D4EB2Fu);
}
// lists of a warp: `base` + i * PLT<PT>::kBytes
template <typename PT>
__device__ __forceinline__ PT *pl_at(unsigned char *base, int i) {
  return (PT *)(base + (size_t)i * PLT<PT>::kBytes);
}
template <typename PT>
__device__ __forceinline__ const PT *pl_at(const unsigned char *base, int i) {
  return (const PT *)(base + (size_t)i * PLT<PT>::kBytes);
}
template <typename PT>
__device__ __forceinline__ void pl_clear(unsigned char *base, int i) {
  uint4 *q = (uint4 *)(base + (size_t)i * PLT<PT>::kBytes);
  q[0] = make_uint4(0u, 0u, 0u, 0u);
  if (PLT<PT>::kBytes == 32) q[1] = make_uint4(0u, 0u, 0u, 0u);
}
template <typename PT>
__device__ __forceinline__ bool pl_same(const unsigned char *base, int i, int j) {
  const uint4 *a = (const uint4 *)(base + (size_t)i * PLT<PT>::kBytes), *b = (const uint4 *)(base + (size_t)j * PLT<PT>::kBytes);
  bool eq = pl_equal(a[0], b[0]);
  if (PLT<PT>::kBytes == 32) eq = eq && pl_equal(a[1], b[1]);
  return eq;
}
template <typename PT>
__device__ __forceinline__ uint32_t pl_hash(const unsigned char *base, int i) {
  const uint4 *a = (const uint4 *)(base + (size_t)i * PLT<PT>::kBytes);
  uint32_t h = pl_mix(a[0]);
  if (PLT<PT>::kBytes == 32) h = (h * 0x165667B1u) ^ pl_mix(a[1]) ^ (h >> 13);
  h ^= h >> 15;
  return h | 1u;  // odd: the zero padding of the hash runs never matches
}

// ---- dedup (same scan as mark_redundant_cpl, states compared as lists) ------------------------------------------
template <typename T, int CPL, typename PT>
__device__ __forceinline__ void v2_mark_redundant_cpl(const unsigned char *lists, double *lpjs, const uint32_t *hash, int S,
                                                      int Snew, int lane) {
  const int Sp = hash_old_pad(S), Snp = (Snew + 3) & ~3;
  const uint32_t *ho = hash, *hn = hash + Sp;
#pragma unroll 1
  for (int c0 = 0; c0 < Snew; c0 += 32 * CPL) {
    uint32_t hc[CPL];
    int hit[CPL];
#pragma unroll
    for (int t = 0; t < CPL; ++t) {
      const int c = c0 + 32 * t + lane;
      hc[t] = c < Snew ? hn[c] : 0u;
      hit[t] = -1;
    }
#pragma unroll 2
    for (int j = 0; j < Sp; j += 4) {
      const uint4 h4 = *reinterpret_cast<const uint4 *>(ho + j);
#pragma unroll
      for (int t = 0; t < CPL; ++t) {
        if (h4.x == hc[t]) hit[t] = j;
        if (h4.y == hc[t]) hit[t] = j + 1;
        if (h4.z == hc[t]) hit[t] = j + 2;
        if (h4.w == hc[t]) hit[t] = j + 3;
      }
    }
#pragma unroll
    for (int g = 1; g < CPL; ++g) {  // children of later 32-groups of this sweep
      const int jb = c0 + 32 * g, je = min(jb + 32, Snp);
      for (int j = jb; j < je; j += 4) {
        const uint4 h4 = *reinterpret_cast<const uint4 *>(hn + j);
#pragma unroll
        for (int t = 0; t < CPL; ++t) {
          if (t < g) {
            if (h4.x == hc[t]) hit[t] = S + j;
            if (h4.y == hc[t]) hit[t] = S + j + 1;
            if (h4.z == hc[t]) hit[t] = S + j + 2;
            if (h4.w == hc[t]) hit[t] = S + j + 3;
          }
        }
      }
    }
    for (int j = c0 + 32 * CPL; j < Snp; j += 4) {  // children of later sweeps
      const uint4 h4 = *reinterpret_cast<const uint4 *>(hn + j);
#pragma unroll
      for (int t = 0; t < CPL; ++t) {
        if (h4.x == hc[t]) hit[t] = S + j;
        if (h4.y == hc[t]) hit[t] = S + j + 1;
        if (h4.z == hc[t]) hit[t] = S + j + 2;
        if (h4.w == hc[t]) hit[t] = S + j + 3;
      }
    }
#pragma unroll
    for (int t = 0; t < CPL; ++t) {  // later children of the same 32-group
      const unsigned later = __match_any_sync(FULL, hc[t]) & ~((2u << lane) - 1u);
      if (later && hc[t]) hit[t] = S + c0 + 32 * t + __ffs(later) - 1;
    }
#pragma unroll
    for (int t = 0; t < CPL; ++t) {
      const int c = c0 + 32 * t + lane;
      if (c < Snew && hit[t] >= 0) {
        bool dup = pl_same<PT>(lists, hit[t], S + c);
        if (!dup) {  // hash collision: exhaustive rescan (practically never)
          for (int j2 = 0; j2 < S && !dup; ++j2)
            if (ho[j2] == hc[t]) dup = pl_same<PT>(lists, j2, S + c);
          for (int j2 = c + 1; j2 < Snew && !dup; ++j2)
            if (hn[j2] == hc[t]) dup = pl_same<PT>(lists, S + j2, S + c);
        }
        if (dup) lpjs[S + c] = (double)LowLpj<T>::v();
      }
    }
  }
}

// ---- top-S merge (select_merge of estep_device.cuh with the keys taken from the fp64 lpj row) -------------------
template <typename T>
__device__ __forceinline__ uint64_t v2_key(const double *lpjs, int i) {
  return sort_key((double)(T)lpjs[i]);  // ordered by the value that is stored (T), like the generic kernel
}

template <typename T, int EPL>
__device__ __forceinline__ void v2_sort_survivors(const double *lpjs, int32_t *slist, int ns, int lane) {
  uint64_t key[EPL];
  int src[EPL];
#pragma unroll
  for (int r = 0; r < EPL; ++r) {
    const int e = lane * EPL + r;
    src[r] = e < ns ? slist[e] : kPadSrc;
    key[r] = e < ns ? v2_key<T>(lpjs, src[r]) : 0ull;
  }
  __syncwarp();
  bitonic_network<EPL>(key, src, lane, 2);
#pragma unroll
  for (int r = 0; r < EPL; ++r) {
    const int e = lane * EPL + r;
    if (e < ns) slist[e] = src[r];
  }
  __syncwarp();
}

// order[pos] = concatenated index of the state stored at position pos (ascending lpj, best last); lpj_row written
template <typename T, int EPL>
__device__ __forceinline__ void v2_select_merge(const double *lpjs, int32_t *slist, int32_t *order, int S, int Snew,
                                                T *lpj_row, int lane, int &nsub) {
  constexpr int MP = 32 * EPL;
  uint64_t key[EPL];
  int src[EPL];
#pragma unroll
  for (int r = 0; r < EPL; ++r) {
    const int e = lane * EPL + r;
    key[r] = e < S ? v2_key<T>(lpjs, e) : 0ull;
    src[r] = e < S ? e : kPadSrc;
  }
  bitonic_network<EPL>(key, src, lane, 2);
  uint64_t wkey = 0;
#pragma unroll
  for (int r = 0; r < EPL; ++r)
    if (((S - 1) % EPL) == r) wkey = key[r];
  wkey = __shfl_sync(FULL, wkey, (S - 1) / EPL);
  int ns = 0;
  for (int c0 = 0; c0 < Snew; c0 += 32) {
    const int c = c0 + lane;
    const bool keep = c < Snew && v2_key<T>(lpjs, S + c) > wkey;  // ties go to the old state (lower index)
    const unsigned m = __ballot_sync(FULL, keep);
    if (keep) slist[ns + __popc(m & ((1u << lane) - 1u))] = S + c;
    ns += __popc(m);
  }
  __syncwarp();
  if (ns > 0) {  // ns <= Snew <= 128
    if (ns <= 32)
      v2_sort_survivors<T, 1>(lpjs, slist, ns, lane);
    else if (ns <= 64)
      v2_sort_survivors<T, 2>(lpjs, slist, ns, lane);
    else
      v2_sort_survivors<T, 4>(lpjs, slist, ns, lane);
#pragma unroll
    for (int r = 0; r < EPL; ++r) {
      const int bi = MP - 1 - (lane * EPL + r);
      if (bi < ns) {
        const int bs = slist[bi];
        const uint64_t bk = v2_key<T>(lpjs, bs);
        if (ranks_before(bk, bs, key[r], src[r])) {
          key[r] = bk;
          src[r] = bs;
        }
      }
    }
    bitonic_network<EPL>(key, src, lane, MP);
  }
  nsub = 0;
#pragma unroll
  for (int r = 0; r < EPL; ++r) {
    const int e = lane * EPL + r;
    if (e < S) {
      const int pos = S - 1 - e;
      lpj_row[pos] = (T)lpjs[src[r]];
      order[pos] = src[r];
      nsub += (src[r] >= S) ? 1 : 0;
    }
  }
}

// ---- top-S merge on COMPOSITE keys ---------------------------------------------------------------------------
// ckey = the 64-bit sort key with its low 8 bits replaced by 255 - (concatenated index): all keys of a datapoint are
// distinct, "ranks before" is one unsigned 64-bit compare and the index rides in the key (no payload register, no
// 96-bit compare).  Dropping the 8 low mantissa bits can only matter when two candidates agree in the upper 56 bits;
// the merge checks exactly that for every adjacent pair of the result and for the best excluded candidate, and
// reports `false` (nothing written) so that the caller runs the exact merge instead.  True ties order by index
// either way, so equal lpj values pass the check as well when they are exactly equal — they are sent to the exact
// path too, which costs time, not correctness.
template <typename T>
__device__ __forceinline__ uint64_t v2_ckey(const double *lpjs, int i) {
  return (sort_key((double)(T)lpjs[i]) & ~0xFFull) | (uint64_t)(255 - i);
}

// Fully unrolled (the rolled network spends half of its instructions on loop control and direction flags; the three
// instances a shape class needs are ~600 instructions).  KFIRST = 2 sorts, KFIRST = 32*EPL merges a bitonic sequence.
template <int EPL, int KFIRST>
__device__ __forceinline__ void bitonic_desc_u64(uint64_t (&key)[EPL], int lane) {
  constexpr int MP = 32 * EPL;
#pragma unroll
  for (int k = KFIRST; k <= MP; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j >= EPL; j >>= 1) {  // partner in another lane
      const int lm = j / EPL;
      // j and k are multiples of EPL here, so the direction is a property of the lane, not of the element
      const bool want_max = (((lane / lm) ^ (lane / (k / EPL))) & 1) == 0;
#pragma unroll
      for (int r = 0; r < EPL; ++r) {
        const uint64_t o = __shfl_xor_sync(FULL, key[r], lm);
        key[r] = ((o > key[r]) == want_max) ? o : key[r];
      }
    }
#pragma unroll
    for (int jj = EPL >> 1; jj > 0; jj >>= 1) {  // partner in the same lane
      if (jj < k) {
#pragma unroll
        for (int r = 0; r < EPL; ++r) {
          if ((r & jj) == 0) {
            const bool up = k >= MP ? true : ((lane * EPL + r) & k) == 0;  // best-first block
            const uint64_t hi = key[r] > key[r ^ jj] ? key[r] : key[r ^ jj];
            const uint64_t lo = key[r] > key[r ^ jj] ? key[r ^ jj] : key[r];
            key[r] = up ? hi : lo;
            key[r ^ jj] = up ? lo : hi;
          }
        }
      }
    }
  }
}

template <int EPL>
__device__ __forceinline__ void v2_sort_survivors_ck(uint64_t *sk, int ns, int lane) {
  uint64_t key[EPL];
#pragma unroll
  for (int r = 0; r < EPL; ++r) {
    const int e = lane * EPL + r;
    key[r] = e < ns ? sk[e] : 0ull;
  }
  __syncwarp();
  bitonic_desc_u64<EPL, 2>(key, lane);
#pragma unroll
  for (int r = 0; r < EPL; ++r) {
    const int e = lane * EPL + r;
    if (e < ns) sk[e] = key[r];
  }
  __syncwarp();
}

__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
  const uint32_t hi = __reduce_max_sync(FULL, (uint32_t)(v >> 32));
  const uint32_t lo = __reduce_max_sync(FULL, (uint32_t)(v >> 32) == hi ? (uint32_t)v : 0u);
  return (uint64_t)hi << 32 | lo;
}

// EPL: 32*EPL >= S;  EPLN: 32*EPLN >= Snew.  sk: Snew composite keys of scratch.
template <typename T, int EPL, int EPLN>
__device__ __forceinline__ bool v2_select_merge_ck(const double *lpjs, uint64_t *sk, int32_t *order, int S, int Snew,
                                                   T *lpj_row, int lane, int &nsub) {
  constexpr int MP = 32 * EPL;
  uint64_t key[EPL];
#pragma unroll
  for (int r = 0; r < EPL; ++r) {
    const int e = lane * EPL + r;
    key[r] = e < S ? v2_ckey<T>(lpjs, e) : 0ull;
  }
  bitonic_desc_u64<EPL, 2>(key, lane);
  uint64_t wkey = 0;
#pragma unroll
  for (int r = 0; r < EPL; ++r)
    if (((S - 1) % EPL) == r) wkey = key[r];
  wkey = __shfl_sync(FULL, wkey, (S - 1) / EPL) & ~0xFFull;
  // children whose upper 56 bits reach the worst old state's (conservative: an exact tie loses to the old state,
  // a 56-bit tie is resolved by the check below)
  int ns = 0;
  for (int c0 = 0; c0 < Snew; c0 += 32) {
    const int c = c0 + lane;
    const uint64_t ck = c < Snew ? v2_ckey<T>(lpjs, S + c) : 0ull;
    const bool keep = c < Snew && ck >= wkey;
    const unsigned m = __ballot_sync(FULL, keep);
    if (keep) sk[ns + __popc(m & ((1u << lane) - 1u))] = ck;
    ns += __popc(m);
  }
  __syncwarp();
  uint64_t dropped = 0;  // best candidate that does not make the top MP
  if (ns > 0) {
    if (EPLN == 1 || ns <= 32)
      v2_sort_survivors_ck<1>(sk, ns, lane);
    else if (EPLN == 2 || ns <= 64)
      v2_sort_survivors_ck<(EPLN >= 2 ? 2 : 1)>(sk, ns, lane);
    else
      v2_sort_survivors_ck<(EPLN >= 4 ? 4 : 1)>(sk, ns, lane);
#pragma unroll
    for (int r = 0; r < EPL; ++r) {
      const int bi = MP - 1 - (lane * EPL + r);
      if (bi < ns) {
        const uint64_t bk = sk[bi];
        const uint64_t lo = bk > key[r] ? key[r] : bk;
        key[r] = bk > key[r] ? bk : key[r];
        dropped = lo > dropped ? lo : dropped;
      }
    }
    for (int bi = MP + lane; bi < ns; bi += 32) dropped = sk[bi] > dropped ? sk[bi] : dropped;  // survivors beyond MP
    bitonic_desc_u64<EPL, MP>(key, lane);
  }
  // 56-bit ties: adjacent elements of the result (ranks 0..S, the first excluded one included when it is held here) ...
  bool amb = false;
#pragma unroll
  for (int r = 0; r < EPL; ++r) {
    const int e = lane * EPL + r;
    uint64_t nxt;
    if (r + 1 < EPL) {
      nxt = key[r + 1];
    } else {
      nxt = __shfl_down_sync(FULL, key[0], 1);
      if (lane == 31) nxt = 0ull;
    }
    amb |= e < S && nxt != 0ull && (key[r] >> 8) == (nxt >> 8);
  }
  // ... and the last selected element against the best candidate dropped by the merge
  uint64_t last = 0;
#pragma unroll
  for (int r = 0; r < EPL; ++r)
    if (((S - 1) % EPL) == r) last = key[r];
  last = __shfl_sync(FULL, last, (S - 1) / EPL);
  dropped = warp_max_u64(dropped);
  amb |= dropped != 0ull && (dropped >> 8) == (last >> 8);
  if (__any_sync(FULL, amb)) return false;
  nsub = 0;
#pragma unroll
  for (int r = 0; r < EPL; ++r) {
    const int e = lane * EPL + r;
    if (e < S) {
      const int pos = S - 1 - e, src = 255 - (int)(key[r] & 0xFFull);
      lpj_row[pos] = (T)lpjs[src];
      order[pos] = src;
      nsub += (src >= S) ? 1 : 0;
    }
  }
  return true;
}

// ---- the kernel -----------------------------------------------------------------------------------------------
// ES / EG / EN: elements per lane that hold S, S_gen and S_new (1, 2 or 4): only the networks of one shape class are
// compiled into a kernel (the instruction cache is the scarce resource here)
template <typename T, int ES, int EG, int EN, typename PT = unsigned char>
__global__ void __launch_bounds__(256, 4)
estep_evo_bsc_v2_kernel(const T *__restrict__ Y, int64_t ldY, const T *__restrict__ A, const int64_t *__restrict__ idx,
                        int64_t row0, uint64_t *__restrict__ K, T *__restrict__ lpj, const void *ws, EvoParams ep,
                        V2Dims dims, int64_t B, int D, unsigned long long *nsubs, int32_t *ovf_count,
                        int32_t *__restrict__ ovf_list, uint32_t sync_mask, const double *__restrict__ yy_in,
                        int64_t ovf_base) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int S = dims.S, Snew = dims.Snew, Sgen = dims.Sgen, W64 = dims.W64, H = dims.H;
  const BscTheta<T> th = bsc_theta_view<T>(ws, H, D);
  unsigned char *p = smem_raw + wib * dims.warp_bytes;
  constexpr bool kStageD = sizeof(PT) == 1;         // == v2_stage_d(H)
  unsigned char *lists = p;                         // S old + Snew children, PLT<PT>::kBytes each
  double *lpjs = (double *)(p + dims.off_lpj);      // their lpj, fp64
  uint32_t *hash = (uint32_t *)(p + dims.off_hash);  // dedup hashes; survivor list of the merge afterwards
  int32_t *iscr = (int32_t *)(p + dims.off_iscr);   // chosen parents (P); order of the merged set (S)
  double *dsm = (double *)(p + dims.off_d);         // d_h = gp_h - 2 coef a_h of this datapoint, all H units (kStageD)
  const double coef = th.hdr->coef, m2c = -2.0 * coef;
  const double *__restrict__ G2 = th.G2;
  const double *__restrict__ gp = th.gp;
  unsigned long long my_subs = 0;

  // the warps of a CTA walk the phases in lockstep (instruction-cache locality, see bsc_fast.cu)
  int sync_id = 0;
#define V2_PHASE_SYNC()                          \
  do {                                           \
    if ((sync_mask >> sync_id) & 1u) __syncthreads(); \
    ++sync_id;                                   \
  } while (0)
  for (int64_t b0 = (int64_t)blockIdx.x * wpb; b0 < B; b0 += (int64_t)gridDim.x * wpb) {
    sync_id = 0;
    const bool active = b0 + wib < B;
    const int64_t b = active ? b0 + wib : 0;
    const int64_t row = idx ? idx[b] : row0 + b;
    const uint32_t n_glob = ep.n_offset + (uint32_t)row;
    uint64_t *Krow = K + row * (int64_t)S * W64;
    const T *__restrict__ y = Y + b * ldY;
    const T *__restrict__ a_row = A + b * (int64_t)H;
    double yy = 0.0;
    bool bail = !active;  // idle warps only keep the barriers company
    if (yy_in) {  // ||y||^2 from the GEMM's epilogue (NaN if y has a NaN): y itself is not read here
      yy = yy_in[b];
      bail |= (yy != yy);
    } else {
      for (int d = lane; d < D; d += 32) {
        const double v = (double)y[d];
        bail |= (v != v);
        yy += v * v;
      }
      yy = warp_sum(yy);
    }
    const double yyc = yy * coef;
    __syncwarp();  // the previous datapoint's readers of this warp's shared memory are done
#define V2_D(h) (kStageD ? dsm[h] : fma(m2c, (double)a_row[h], gp[h]))
    if (!kStageD) {
      // d_h gathered on demand
    } else if ((H & 1) == 0 && sizeof(T) == 8) {  // coalesced read of the a row, two units per lane
      for (int h = 2 * lane; h < H; h += 64) {
        const double2 av = *reinterpret_cast<const double2 *>(a_row + h), gv = *reinterpret_cast<const double2 *>(gp + h);
        *reinterpret_cast<double2 *>(dsm + h) = make_double2(fma(m2c, av.x, gv.x), fma(m2c, av.y, gv.y));
      }
    } else {
      for (int h = lane; h < H; h += 32) dsm[h] = fma(m2c, (double)a_row[h], gp[h]);
    }
    __syncwarp();
    V2_PHASE_SYNC();

    // ---- old states: bit rows -> lists, scored in the same walk (evo.py:95) ----
    for (int s0 = 0; s0 < S && active; s0 += 32) {
      const int s = s0 + lane;
      if (s < S) {
        pl_clear<PT>(lists, s);
        PT *lb = pl_at<PT>(lists, s);
        const uint32_t *st = (const uint32_t *)(Krow + (size_t)s * W64);  // 2 W64 half words
        // which half words are non-empty and how many units are active; then ONE loop over the active units that
        // re-reads a half word (an L1 hit) only when the current one is used up: the warp iterates max-over-lanes |s|
        // times instead of once per non-empty half word of any lane
        uint32_t nz = 0;
        int ntot = 0;
        if ((W64 & 1) == 0) {
          for (int q = 0; q < W64 / 2; ++q) {
            const uint4 v = *reinterpret_cast<const uint4 *>(st + 4 * q);
            nz |= ((v.x ? 1u : 0u) | (v.y ? 2u : 0u) | (v.z ? 4u : 0u) | (v.w ? 8u : 0u)) << (4 * q);
            ntot += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
          }
        } else {
          for (int hw = 0; hw < 2 * W64; ++hw) {
            const uint32_t v = st[hw];
            nz |= (v ? 1u : 0u) << hw;
            ntot += __popc(v);
          }
        }
        double acc = yyc;
        if (ntot > kV2MaxList) {
          bail = true;
          ntot = 0;
        }
        uint32_t m = 0;
        int hw = 0;
#pragma unroll 1
        for (int n = 0; n < ntot; ++n) {
          if (m == 0) {
            hw = __ffs((int)nz) - 1;
            nz &= nz - 1;
            m = st[hw];
          }
          const int h = (hw << 5) + __ffs((int)m) - 1;
          m &= m - 1;
          acc += V2_D(h);
          const double *g2 = G2 + (size_t)h * H;
#pragma unroll 1
          for (int i = 0; i < n; ++i) acc += g2[lb[i]];
          lb[n] = (PT)h;
        }
        lb[15] = (PT)ntot;
        lpjs[s] = acc;
      }
    }
    bail = __any_sync(FULL, bail);

    // ---- G generations: select, mutate, score incrementally ----
    for (int g = 0; g < ep.G; ++g) {
      V2_PHASE_SYNC();
      if (bail) {
        V2_PHASE_SYNC();
        continue;
      }
      const int cand_base = g == 0 ? 0 : S + (g - 1) * Sgen;
      const int S_c = g == 0 ? S : Sgen;
      Stream st;
      st.key = ep.key;
      st.n = n_glob;
      st.tag = make_tag(ep.step, g, PUR_SELECT);
      const int bits = index_bits(S_c);
      __syncwarp();
      if (g == 0)
        randparents_sort<ES, true>(S_c, ep.P, bits, st, iscr, lane);
      else
        randparents_sort<EG, true>(S_c, ep.P, bits, st, iscr, lane);
      __syncwarp();
      V2_PHASE_SYNC();
      st.tag = make_tag(ep.step, g, PUR_FLIP0);
      for (int j0 = 0; j0 < Sgen; j0 += 32) {
        const int j = j0 + lane;
        if (j < Sgen) {
          const int src = cand_base + iscr[(j * dims.cinv) >> 16];  // j / C
          const PT *pb = pl_at<PT>(lists, src);  // the parent's list
          const int a = (int)pb[15];
          // flip counts (batch_sparseflip evo.py:310-329 as Binomial counts, mutate_sparseflip of estep_device.cuh)
          const FlipEntry *te = ep.flip_tab + a;
          FlipClass f0, f1;
          f0.pmf0 = __ldg(&te->c[0].pmf0);
          f0.r = __ldg(&te->c[0].r);
          f0.kconst = __ldg(&te->c[0].kconst);
          f0.flip = __ldg(&te->c[0].flip);
          f1.pmf0 = __ldg(&te->c[1].pmf0);
          f1.r = __ldg(&te->c[1].r);
          f1.kconst = __ldg(&te->c[1].kconst);
          f1.flip = __ldg(&te->c[1].flip);
          int k0 = f0.kconst, k1 = f1.kconst;
          if (k0 < 0 || k1 < 0) {
            const uint4 b0 = st.block((uint32_t)j, 0);
            if (k0 < 0) k0 = flip_class_sample(f0, unit_double((uint64_t)b0.y << 32 | b0.x), H - a);
            if (k1 < 0) k1 = flip_class_sample(f1, unit_double((uint64_t)b0.w << 32 | b0.z), a);
          }
          uint32_t t = 0;
          uint4 r = make_uint4(0u, 0u, 0u, 0u);
          // zero bits: k0 of the H - a candidates by rejection over all H positions (H - a >= 49: the dense class,
          // never its complement form)
          PT *ab = pl_at<PT>((unsigned char *)hash, lane);  // added positions, ascending (the hash region is idle here)
          int ka = 0;
          if (k0 > kV2MaxAdd) {
            bail = true;
            k0 = 0;
            k1 = 0;
          }
          if (k0 > 0) {
            const uint32_t cap = 64u + 64u * (uint32_t)k0;
            while (ka < k0 && t < cap) {
              if ((t & 3u) == 0u) r = st.block((uint32_t)j, 1u + (t >> 2));
              const uint32_t ri = (t & 3u) == 0u ? r.x : (t & 3u) == 1u ? r.y : (t & 3u) == 2u ? r.z : r.w;
              ++t;
              const uint32_t pos = randint_u32(ri, (uint32_t)H);
              bool taken = false;
#pragma unroll 1
              for (int i = 0; i < a; ++i) taken |= ((uint32_t)pb[i] == pos);
              int below = 0;
#pragma unroll 1
              for (int q = 0; q < ka; ++q) {
                const uint32_t z = ab[q];
                taken |= (z == pos);
                below += (z < pos) ? 1 : 0;
              }
              if (!taken) {  // sorted insert at byte `below`
#pragma unroll 1
                for (int q = ka; q > below; --q) ab[q] = ab[q - 1];
                ab[below] = (PT)pos;
                ++ka;
              }
            }
            if (ka < k0) bail = true;  // rejection cap (the generic kernel continues with the n-th-remaining rule)
          }
          // one bits: k1 of the a candidates, each draw the idx-th REMAINING entry (2a < H: the sparse class)
          uint32_t fl = 0;  // entries of L that are flipped (removed)
          if (k1 > 0) {
            const bool comp = 2 * k1 > a;
            const int draws = k1 >= a ? 0 : (comp ? a - k1 : k1);
            uint32_t dr = 0;
            for (int got = 0; got < draws; ++got) {
              if ((t & 3u) == 0u) r = st.block((uint32_t)j, 1u + (t >> 2));
              const uint32_t ri = (t & 3u) == 0u ? r.x : (t & 3u) == 1u ? r.y : (t & 3u) == 2u ? r.z : r.w;
              ++t;
              int ix = (int)randint_u32(ri, (uint32_t)(a - got));
              int e = 0;
              while (true) {
                if (!((dr >> e) & 1u)) {
                  if (ix == 0) break;
                  --ix;
                }
                ++e;
              }
              dr |= 1u << e;
            }
            const uint32_t valid = (1u << a) - 1u;
            fl = (k1 >= a || comp) ? (valid & ~dr) : dr;
          }
          const int nc = a - __popc(fl) + ka;
          if (nc > kV2MaxList) bail = true;
          // lpj(child) - lpj(parent): the removed units, then the added ones, each against the units present
          const uint32_t kept = ((1u << a) - 1u) & ~fl;
          double dl = 0.0;
          for (uint32_t f = fl; f;) {
            const int i = __ffs((int)f) - 1;
            f &= f - 1;
            const uint32_t h = pb[i];
            double t1 = V2_D(h);
            const double *g2 = G2 + (size_t)h * H;
            for (uint32_t mm = kept | f; mm;) {  // kept units and the removed ones after i
              const int i2 = __ffs((int)mm) - 1;
              mm &= mm - 1;
              t1 += g2[pb[i2]];
            }
            dl -= t1;
          }
          {
#pragma unroll 1
            for (int q = 0; q < ka; ++q) {
              const uint32_t z = ab[q];
              double t1 = V2_D(z);
              const double *g2 = G2 + (size_t)z * H;
              for (uint32_t mm = kept; mm;) {
                const int i2 = __ffs((int)mm) - 1;
                mm &= mm - 1;
                t1 += g2[pb[i2]];
              }
#pragma unroll 1
              for (int q2 = 0; q2 < q; ++q2) t1 += g2[ab[q2]];
              dl += t1;
            }
          }
          // the child's list: kept entries merged with the added positions, ascending
          const int child = S + g * Sgen + j;
          pl_clear<PT>(lists, child);
          if (nc <= kV2MaxList) {
            PT *cb = pl_at<PT>(lists, child);
            int m = 0, q = 0;
            uint32_t nextz = ka ? (uint32_t)ab[0] : PLT<PT>::kNone;  // the next added position not yet written
            for (uint32_t mm = kept; mm;) {
              const int i = __ffs((int)mm) - 1;
              mm &= mm - 1;
              const uint32_t e = pb[i];
              while (nextz < e) {
                cb[m++] = (PT)nextz;
                ++q;
                nextz = q < ka ? (uint32_t)ab[q] : PLT<PT>::kNone;
              }
              cb[m++] = (PT)e;
            }
#pragma unroll 1
            for (; q < ka; ++q) cb[m++] = ab[q];
            cb[15] = (PT)m;
          }
          lpjs[child] = lpjs[src] + dl;
        }
      }
      bail = __any_sync(FULL, bail);
    }
    // -> the generic kernel; row0 + b is the datapoint's index in the CALL (row0 = first datapoint of this chunk)
    if (bail && active && lane == 0) ovf_list[atomicAdd(ovf_count, 1)] = (int32_t)(ovf_base + b);
    __syncwarp();
    V2_PHASE_SYNC();
    if (!bail)

    // ---- dedup: a child equal to a later child or to an old state is redundant (_set_redundant_lpj_to_low) ----
    {
      const int Sp = hash_old_pad(S), Snp = (Snew + 3) & ~3;
      for (int i = lane; i < Sp; i += 32) hash[i] = i < S ? pl_hash<PT>(lists, i) : 0u;
      for (int i = lane; i < Snp; i += 32) hash[Sp + i] = i < Snew ? pl_hash<PT>(lists, S + i) : 0u;
      __syncwarp();
      v2_mark_redundant_cpl<T, EN, PT>(lists, lpjs, hash, S, Snew, lane);
      __syncwarp();
    }

    V2_PHASE_SYNC();
    if (bail) continue;
    // ---- top-S merge (update_states_for_batch) and write-back ----
    int nsub = 0;
    T *lrow = lpj + row * (int64_t)S;
    if (!v2_select_merge_ck<T, ES, EN>(lpjs, (uint64_t *)hash, iscr, S, Snew, lrow, lane, nsub)) {
      __syncwarp();
      v2_select_merge<T, ES>(lpjs, (int32_t *)hash, iscr, S, Snew, lrow, lane, nsub);  // exact keys (56-bit tie)
    }
    my_subs += (unsigned long long)warp_sum(nsub);
    __syncwarp();
    for (int p0 = 0; p0 < S; p0 += 32) {
      const int pos = p0 + lane;
      if (pos < S) {
        const PT *ob = pl_at<PT>(lists, iscr[pos]);
        const int n = (int)ob[15];
        uint64_t *dst = Krow + (size_t)pos * W64;
        if (sizeof(PT) == 1) {  // W64 <= 4
          uint64_t wd[4] = {0ull, 0ull, 0ull, 0ull};
          for (int i = 0; i < n; ++i) {
            const uint32_t h = ob[i];
            const uint64_t bit = 1ull << (h & 63u);
            const uint32_t w = h >> 6;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (w == (uint32_t)k) wd[k] |= bit;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < W64) dst[k] = wd[k];
        } else {  // the list is ascending: one pass over the words, two per 16-byte store when the row allows it
          int i = 0;
          uint32_t h = n ? (uint32_t)ob[0] : PLT<PT>::kNone;
          if ((W64 & 1) == 0) {
            for (int k = 0; k < W64; k += 2) {
              uint64_t w0 = 0ull, w1 = 0ull;
              while ((h >> 6) == (uint32_t)k) {
                w0 |= 1ull << (h & 63u);
                h = ++i < n ? (uint32_t)ob[i] : PLT<PT>::kNone;
              }
              while ((h >> 6) == (uint32_t)(k + 1)) {
                w1 |= 1ull << (h & 63u);
                h = ++i < n ? (uint32_t)ob[i] : PLT<PT>::kNone;
              }
              *reinterpret_cast<ulonglong2 *>(dst + k) = make_ulonglong2(w0, w1);
            }
          } else {
            for (int k = 0; k < W64; ++k) {
              uint64_t w0 = 0ull;
              while ((h >> 6) == (uint32_t)k) {
                w0 |= 1ull << (h & 63u);
                h = ++i < n ? (uint32_t)ob[i] : PLT<PT>::kNone;
              }
              dst[k] = w0;
            }
          }
        }
      }
    }
  }
#undef V2_PHASE_SYNC
#undef V2_D
  if (lane == 0 && my_subs) atomicAdd(nsubs, my_subs);
}

}  // namespace tvo
