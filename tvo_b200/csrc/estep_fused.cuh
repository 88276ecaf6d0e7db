// tvo_b200 — model-generic fused EVO E-step kernel (EVOVariationalStates.update, evo.py:80-115).
//
// `Model` supplies the log-pseudo-joint:
//     struct Model {
//       struct Params {...};                                   // plain data passed by value
//       static size_t smem_bytes(const Params&);               // extra per-warp shared memory
//       struct Ctx { __device__ void load(const Params&, int64_t b, unsigned char* smem, int lane);
//                    __device__ void score(const Params&, const uint64_t* states, int count, int W64,
//                                          T* out, int lane); };   // warp-cooperative
//     };
// Everything else — Philox candidate generation, dedup, top-S merge, phase-lockstep scheduling — is
// shared with the BSC kernels (estep_device.cuh).
#pragma once
#include "estep_device.cuh"

namespace tvo {

// Two register / residency tiers, picked by how many 8-warp CTAs the working set allows per SM:
//   <8, 3>  three or more CTAs of 8 warps (80 registers; NoisyOR);
//   <16, 1> ONE CTA per SM with as many warps as its shared memory holds (up to 16; 128 registers): SSSC at H = 256.
// Measured at C3 (E-step per 200k datapoints): 8 warps 19.1 ms, 10 warps 14.9, 12 warps 12.6, 14 warps 11.1; two CTAs
// of 8 warps at 128 registers were slower than one wide CTA with fewer warps (13.3 ms with 16 warps against 11.1 ms
// with 14), so there is no two-CTA tier.
template <typename T, typename Model, int MAXW, int MINB>
__global__ void __launch_bounds__(32 * MAXW, MINB)
estep_evo_fused_kernel(typename Model::Params mp, const int64_t *__restrict__ idx, int64_t row0,
                       uint64_t *__restrict__ K, T *__restrict__ lpj, EvoParams ep,
                       const double *__restrict__ sparsity_p, EstepDims dims, size_t model_smem, int64_t B, int H,
                       unsigned long long *nsubs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const size_t wbytes = warp_buf_bytes<T>(dims) + align16(model_smem);
  unsigned char *base = smem_raw + (size_t)wib * wbytes;
  const WarpBuf<T> wb = carve_warp_buf<T>(base, dims);
  unsigned char *msm = base + warp_buf_bytes<T>(dims);
  const int S = dims.S, Snew = dims.Snew, Sgen = dims.Sgen, W64 = dims.W64;
  const double sparsity = sparsity_p ? *sparsity_p : 0.0;
  unsigned long long my_subs = 0;
  typename Model::Ctx ctx;
  // warps of a CTA move through the phases in lockstep (instruction-cache locality, see bsc_fast.cu)
  for (int64_t b0 = (int64_t)blockIdx.x * wpb; b0 < B; b0 += (int64_t)gridDim.x * wpb) {
    const int64_t b = b0 + wib;
    const bool active = b < B;
    const int64_t row = active ? (idx ? idx[b] : row0 + b) : 0;
    const uint32_t n_glob = ep.n_offset + (uint32_t)row;
    uint64_t *Krow = K + row * (int64_t)S * W64;
    T *lrow = lpj + row * (int64_t)S;
    if (active) {
      ctx.load(mp, b, msm, lane);
      for (int t = lane; t < S * W64; t += 32) wb.oldK[t] = Krow[t];
      __syncwarp();
    }
    __syncthreads();
    if (active) {  // lpj[idx] = lpj_fn(batch, K[idx])  (evo.py:95), or already there (batch_fitparents: the normaliser
                   // of generation 0 is a reduction over exactly these values)
      if (ep.lpj_current) {
        for (int i = lane; i < S; i += 32) wb.lpj[i] = lrow[i];
        __syncwarp();
      } else {
        ctx.score(mp, wb.oldK, S, W64, wb.lpj, lane);
      }
    }
    for (int g = 0; g < ep.G; ++g) {
      const uint64_t *cand = g == 0 ? wb.oldK : wb.newK + (size_t)(g - 1) * Sgen * W64;
      const T *cand_lpj = g == 0 ? wb.lpj : wb.lpj + S + (g - 1) * Sgen;
      uint64_t *children = wb.newK + (size_t)g * Sgen * W64;
      __syncthreads();
      if (active)
        evo_generation<T>(wb, cand, cand_lpj, g == 0 ? S : Sgen, children, ep, g, n_glob, W64, H, sparsity,
                          ep.fit_norm ? ep.fit_norm[0] : 0.0, ep.fit_norm ? ep.fit_norm[1] : 1.0, lane);
      __syncthreads();
      if (active) ctx.score(mp, children, Sgen, W64, wb.lpj + S + g * Sgen, lane);
    }
    __syncthreads();
    if (active) mark_redundant<T>(wb, S, Snew, W64, lane);
    __syncthreads();
    if (active) my_subs += (unsigned long long)select_topS<T>(wb, S, Snew, W64, Krow, lrow, lane);
  }
  if (lane == 0 && my_subs) atomicAdd(nsubs, my_subs);
}

int evo_params_from_cfg(const tvo_evo_config *cfg, int S, int H, EvoParams &ep, EstepDims &dims, const char *who);

// Host launcher shared by the models.
template <typename T, typename Model>
int launch_estep_evo_fused(const typename Model::Params &mp, const int64_t *idx, int64_t row0, uint64_t *K, T *lpj,
                           const tvo_evo_config *cfg, const double *sparsity, int64_t B, int S, int H, int64_t *nsubs,
                           cudaStream_t stream, const char *who) {
  EvoParams ep;
  EstepDims dims;
  int rc = evo_params_from_cfg(cfg, S, H, ep, dims, who);
  if (rc) return rc;
  TVO_REQUIRE_FITPARENTS(ep, who);
  TVO_REQUIRE(!(ep.mut == TVO_MUT_SPARSEFLIP || ep.mut == TVO_MUT_CROSS_SPARSEFLIP) || sparsity,
              "%s: sparseflip needs the device scalar `sparsity`", who);
  if (B == 0) return TVO_OK;
  const size_t msm = Model::smem_bytes(mp);
  const size_t wbytes = warp_buf_bytes<T>(dims) + align16(msm);
  int wpb = 8;
  while (wpb > 1 && wbytes * wpb > (size_t)227 * 1024) wpb >>= 1;
  TVO_REQUIRE(wbytes * wpb <= (size_t)227 * 1024, "%s: per-datapoint working set %zu B exceeds shared memory", who, wbytes);
  size_t smem = wbytes * wpb;
  auto kern = estep_evo_fused_kernel<T, Model, 8, 3>;
  const size_t fit8 = wpb == 8 ? ((size_t)227 * 1024) / (smem + 1024) : 0;  // 8-warp CTAs per SM by shared memory
  if (fit8 < 3) {  // fewer than three 8-warp CTAs per SM: ONE wide CTA with as many warps as fit (up to 16)
    wpb = (int)max((size_t)1, min((size_t)16, (size_t)227 * 1024 / wbytes));
    smem = wbytes * wpb;
    kern = estep_evo_fused_kernel<T, Model, 16, 1>;
  }
  TVO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  TVO_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, wpb * 32, smem));
  occ = max(occ, 1);
  const int grid = (int)max((int64_t)1, min((B + wpb - 1) / wpb, (int64_t)sm_count() * occ));
  kern<<<grid, wpb * 32, smem, stream>>>(mp, idx, row0, K, lpj, ep, sparsity, dims, msm, B, H,
                                         (unsigned long long *)nsubs);
  TVO_AFTER_LAUNCH(who);
  return TVO_OK;
}

}  // namespace tvo
