/* tvo_b200 — C ABI of the B200-native TVO hot path (truncated E-step + M-step accumulation).
 *
 * This is the drop-in boundary.  The reference (tvlearn/tvo) has no FFI for this path: its
 * boundary is a duck-typed Python protocol (tvo/utils/model_protocols.py:15-229,
 * tvo/variational/TVOVariationalStates.py:47-56) plus ONE native routine
 * (tvo/variational/_set_redundant_lpj_to_low_CPU.pyx:13).  Each entry point below names the
 * reference code it replaces.  INTEGRATION.md shows the ctypes binding a maintainer of the
 * reference would add.
 *
 * Conventions
 *  - every function returns 0 on success, a TVO_E* code otherwise; tvo_last_error() returns a
 *    thread-local message.  Nothing throws across the boundary.
 *  - all pointers are DEVICE pointers unless the name ends in _host; buffers are borrowed for the
 *    duration of the call (the caller owns them; typically torch tensors).
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls only
 *    enqueue work; they never synchronise unless documented.
 *  - states are bit-packed: state row = W64 = ceil(H/64) uint64 words, bit (h%64) of word (h/64)
 *    is s_h.  K is (N, S, W64) row-major; lpj is (N, S) row-major.
 *  - `idx` (int64, may be NULL) selects rows of the dataset-sized arrays K/lpj for a batch;
 *    NULL means rows 0..B-1.  The global datapoint index that addresses the random stream is
 *    n_offset + (idx ? idx[b] : b).
 *  - suffix _f64 / _f32 is the storage AND arithmetic type of Y, theta and lpj ("precision" in
 *    the reference).  M-step statistics are always accumulated in fp64.
 */
#ifndef TVO_B200_H
#define TVO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TVO_OK 0
#define TVO_EINVAL 1     /* bad argument (message says which) */
#define TVO_ECUDA 2      /* CUDA runtime error (message carries cudaGetErrorString) */
#define TVO_EUNSUPPORTED 3
#define TVO_ENONFINITE 4 /* reserved for host-side checks of the nonfinite flag */

#define TVO_ABI_VERSION 2

/* parent selection (tvo/variational/evo.py:223) */
#define TVO_SEL_RANDPARENTS 0      /* batch_randparents  evo.py:436-451 */
#define TVO_SEL_FITPARENTS 1       /* batch_fitparents   evo.py:397-433, batch-global normaliser (quirk Q3) */
#define TVO_SEL_FITPARENTS_FIXED 2 /* per-datapoint normaliser + inverse-CDF index (documented correction) */
/* mutation (evo.py:224-230) */
#define TVO_MUT_RANDFLIP 0
#define TVO_MUT_SPARSEFLIP 1
#define TVO_MUT_CROSS 2
#define TVO_MUT_CROSS_RANDFLIP 3
#define TVO_MUT_CROSS_SPARSEFLIP 4
/* models */
#define TVO_MODEL_BSC 0
#define TVO_MODEL_NOISYOR 1
#define TVO_MODEL_SSSC 2

typedef struct {
  int32_t parent_selection; /* TVO_SEL_* */
  int32_t mutation;         /* TVO_MUT_* */
  int32_t n_parents;
  int32_t n_children; /* per parent (ignored for cross*, where it is n_parents-1) */
  int32_t n_generations;
  int32_t flags;     /* TVO_EVO_LPJ_CURRENT: lpj[idx] already holds lpj_fn(batch, K[idx]) under the current theta (the fused
                        E-step then skips its re-score, evo.py:95) */
  double p_bf;       /* bitflip_frequency (sparseflip) */
  uint64_t seed;     /* Philox key */
  uint32_t step;     /* E-step counter (epoch * batches + batch), part of the Philox counter */
  uint32_t n_offset; /* global index of local datapoint 0 (rank offset) */
  const double *fit_norm; /* device {m, T} of tvo_fit_norm over the batch's lpj: required by the fused E-steps for
                             TVO_SEL_FITPARENTS (the reference's batch-global normaliser, evo.py:404-406), which they
                             support for ONE generation (later generations need the normaliser of the previous
                             generation's children over the whole batch: stand-alone kernels).  NULL otherwise. */
} tvo_evo_config;
#define TVO_EVO_LPJ_CURRENT 1

/* dedup sentinel written for redundant children: (float)-1e20, as the reference's CPU path
 * (_set_redundant_lpj_to_low_CPU.pyx:17) */
#define TVO_LOW_LPJ_F32 (-1e20f)

int tvo_abi_version(void);
const char *tvo_last_error(void);
/* number of kernels launched by this library in the calling process since load / last reset */
int64_t tvo_launch_count(void);
void tvo_reset_launch_count(void);
/* Opt-in timing of the dominant kernels (measurement aid, no reference counterpart): while enabled, every
 * launch of a timed kernel is bracketed by a cudaEvent pair on the stream it is launched on.
 * tvo_kernel_timing_enable(on) also discards earlier samples; tvo_kernel_timing_read synchronises the
 * recorded events and returns the summed duration and the number of launches of kernel `which`. */
#define TVO_TIMED_ESTEP_EVO_BSC 1 /* estep_evo_bsc_gram_kernel / estep_evo_bsc_kernel */
#define TVO_TIMED_BSC_MSTEP 2     /* bsc_mstep_kernel */
#define TVO_TIMED_BSC_GEMM_YW 3   /* bsc_gemm_yw_kernel: a = Y W of the Gram-form E-step */
#define TVO_TIMED_BSC_GEMM_ETY 4  /* bsc_gemm_ety_kernel: WpT += E^T Y of the M-step */
void tvo_kernel_timing_enable(int32_t on);
int tvo_kernel_timing_read(int32_t which, double *total_ms, int64_t *launches);
int tvo_device_sm_count(int *out);

/* ---- state layout: replaces `states.to(dtype=...)` casts (bsc.py:102, noisyor.py:78) ---- */
int tvo_pack_states(const uint8_t *K_u8, uint64_t *K_packed, int64_t n_states, int32_t H, void *stream);
int tvo_unpack_states(const uint64_t *K_packed, uint8_t *K_u8, int64_t n_states, int32_t H, void *stream);
/* all 2^H states, row i = binary of i MSB first (fullem.py:14-28); out is (2^H, W64) */
int tvo_state_matrix(uint64_t *K_packed, int32_t H, void *stream);

/* ---- workspace sizes (bytes) ---- */
int64_t tvo_bsc_theta_ws_bytes(int32_t H, int32_t D, int32_t is_f64);

/* ---- BSC (tvo/models/bsc.py) ---- */
/* Digest theta into the kernel-friendly form (padded W^T, log-odds prior, constants).  Runs on
 * the device, no host sync.  W is (D,H) row-major as in the reference; pies has n_pies = H or 1
 * entries (individual_priors, bsc.py:107-112); sigma2 has 1.  Must be re-run whenever theta
 * changes (theta is user-rebindable; nothing is cached across calls). */
int tvo_bsc_prepare_f64(const double *W, const double *pies, int32_t n_pies, const double *sigma2, int32_t H,
                        int32_t D, void *theta_ws, void *stream);
int tvo_bsc_prepare_f32(const float *W, const float *pies, int32_t n_pies, const float *sigma2, int32_t H,
                        int32_t D, void *theta_ws, void *stream);

/* log_pseudo_joint (bsc.py:100-119).  Y is (B, D) with row stride ldY; K row for datapoint b is
 * K + (idx?idx[b]:b) * k_stride_n words (k_stride_n = 0 broadcasts one state set, as FullEM's
 * stride-0 view fullem.py:48); out lpj_out[(idx?idx[b]:b) * lpj_stride_n + s].  NaN entries of Y
 * are skipped (to.nansum, bsc.py:115). */
int tvo_bsc_lpj_f64(const double *Y, int64_t ldY, const uint64_t *K, int64_t k_stride_n, const int64_t *idx,
                    const void *theta_ws, double *lpj_out, int64_t lpj_stride_n, int64_t B, int32_t S,
                    int32_t H, int32_t D, void *stream);
int tvo_bsc_lpj_f32(const float *Y, int64_t ldY, const uint64_t *K, int64_t k_stride_n, const int64_t *idx,
                    const void *theta_ws, float *lpj_out, int64_t lpj_stride_n, int64_t B, int32_t S, int32_t H,
                    int32_t D, void *stream);

/* ---- candidate generation (tvo/variational/evo.py:118-451) ---- */
/* One EA generation for a batch: select n_parents from `cand` (B rows of S_c states, lpj cand_lpj)
 * and mutate them into `children` (B, S_gen, W64), S_gen = P*C or P*(P-1).  cand rows are addressed
 * through idx/cand_stride_n like K above; cand_lpj may be NULL for randparents.  `sparsity` is a
 * device scalar (double) = mean(pies) (evo.py:109), read by sparseflip.  fit_norm (device, 2
 * doubles {m, T}) is the batch-global fitness normaliser of evo.py:404-406, produced by
 * tvo_fit_norm_*; required for TVO_SEL_FITPARENTS only. */
int tvo_evo_generation_f64(const uint64_t *cand, int64_t cand_stride_n, const double *cand_lpj,
                           int64_t cand_lpj_stride_n, const int64_t *idx, int32_t cand_indexed, int32_t S_c,
                           uint64_t *children, const tvo_evo_config *cfg, int32_t gen, const double *sparsity,
                           const double *fit_norm, int64_t B, int32_t H, void *stream);
int tvo_evo_generation_f32(const uint64_t *cand, int64_t cand_stride_n, const float *cand_lpj,
                           int64_t cand_lpj_stride_n, const int64_t *idx, int32_t cand_indexed, int32_t S_c,
                           uint64_t *children, const tvo_evo_config *cfg, int32_t gen, const double *sparsity,
                           const double *fit_norm, int64_t B, int32_t H, void *stream);
/* {m, T} of evo.py:404-406 over a (B, S_c) lpj block (rows through idx if indexed). */
int tvo_fit_norm_f64(const double *lpj, int64_t stride_n, const int64_t *idx, int64_t B, int32_t S_c,
                     double *fit_norm_out, void *stream);
int tvo_fit_norm_f32(const float *lpj, int64_t stride_n, const int64_t *idx, int64_t B, int32_t S_c,
                     double *fit_norm_out, void *stream);
/* RandomSampledVarStates.py:57-59: every bit ~ Bernoulli(sparsity); out (B, S_new, W64) */
int tvo_random_states(uint64_t *out, const int64_t *idx, int64_t B, int32_t S_new, int32_t H, double sparsity,
                      uint64_t seed, uint32_t step, uint32_t n_offset, void *stream);

/* ---- dedup: replaces set_redundant_lpj_to_low_CPU (the .pyx, :13-30) and
 *      _set_redundant_lpj_to_low_GPU (_utils.py:59-83).  new_lpj (B,S_new) is modified in place;
 *      old states are K rows through idx. ---- */
int tvo_mark_redundant_f64(const uint64_t *new_states, double *new_lpj, const uint64_t *K, int64_t k_stride_n,
                           const int64_t *idx, int64_t B, int32_t S_new, int32_t S, int32_t H, void *stream);
int tvo_mark_redundant_f32(const uint64_t *new_states, float *new_lpj, const uint64_t *K, int64_t k_stride_n,
                           const int64_t *idx, int64_t B, int32_t S_new, int32_t S, int32_t H, void *stream);

/* ---- top-S merge: replaces update_states_for_batch (_utils.py:124-179).  K/lpj rows (through
 *      idx) are overwritten with the S best of old+new in ascending lpj order (best last); ties go
 *      to the lower concatenated index.  *nsubs (device int64) is INCREMENTED by the number of
 *      selected children. ---- */
int tvo_select_topS_f64(const uint64_t *new_states, const double *new_lpj, uint64_t *K, double *lpj,
                        const int64_t *idx, int64_t B, int32_t S_new, int32_t S, int32_t H, int64_t *nsubs,
                        void *stream);
int tvo_select_topS_f32(const uint64_t *new_states, const float *new_lpj, uint64_t *K, float *lpj,
                        const int64_t *idx, int64_t B, int32_t S_new, int32_t S, int32_t H, int64_t *nsubs,
                        void *stream);

/* ---- BSC.update_param_epoch (bsc.py:160-204) on the device.  tvo_bsc_symmetrize expands the accumulated upper
 *      triangle of my_Wq to the full matrix the solve wants; tvo_bsc_epoch_finish turns the solution sol (H x D,
 *      row-major, = W^T; NULL when the solve failed) and the reduced statistics into the new theta IN PLACE, with
 *      fix_theta (sanity.py:76-88) applied: non-finite W entries <- W_noisy (bsc.py:174), pi clamped to
 *      [1e-5, 1-1e-5], sigma2 >= 1e-5, non-finite pi / sigma2 keep their old values. ---- */
int tvo_bsc_symmetrize(const double *upper, double *full, int32_t H, void *stream);
/*      tvo_bsc_epoch_solve: the solve itself (bsc.py:177-181, W^T = my_Wq^-1 my_Wp^T) for H <= 256 as a hand-written
 *      blocked Cholesky (one CTA) + substitution (one CTA per column) straight from the packed statistics `stats`
 *      ([WpT (H*D) | WqU (H*H upper) | ...], tvo_bsc_stats_len).  `work`: tvo_bsc_epoch_solve_work_doubles(H) doubles
 *      (0: H not covered); `sol`: H x D;  `flags` (2 device doubles): [0] = 0 or 1 + index of the first non-positive
 *      pivot, [1] = min/max of the factor's diagonal — the caller falls back to lstsq, as the reference does on a
 *      failed solve, when [0] != 0 or [1] is tiny.  Deterministic: identical on every rank given identical stats. */
int64_t tvo_bsc_epoch_solve_work_doubles(int32_t H);
/*      tvo_bsc_epoch_solve_variant(1) selects the one-CTA factorisation instead of the 8-CTA cluster (A/B tests). */
int tvo_bsc_epoch_solve_variant(int32_t single_cta);
int tvo_bsc_epoch_solve(const double *stats, double *work, double *sol, double *flags, int32_t H, int32_t D, void *stream);
/*      `guard` (device, may be NULL) = the flags of tvo_bsc_epoch_solve: when they report an unusable factorisation the
 *      call writes nothing, so it can be queued before the host has read the flags. */
int tvo_bsc_epoch_finish_f64(const double *stats, const double *sol, const double *W_noisy, double *W, double *pies,
                             int32_t n_pies, double *sigma2, int32_t H, int32_t D, const double *guard, void *stream);
int tvo_bsc_epoch_finish_f32(const double *stats, const double *sol, const float *W_noisy, float *W, float *pies,
                             int32_t n_pies, float *sigma2, int32_t H, int32_t D, const double *guard, void *stream);

/* ---- fused E-step for BSC: replaces EVOVariationalStates.update (evo.py:80-115) end to end:
 *      re-score K under theta, G generations of select+mutate+score with the children kept in
 *      shared memory, dedup against the original K, top-S merge, nsubs.
 *      `scratch` (device, scratch_bytes; see tvo_bsc_estep_scratch_bytes) enables the Gram form
 *      ||y||^2 - 2 (W^T y).s + s^T (W^T W) s: the batch is processed in chunks of as many datapoints
 *      as fit (one cuBLAS GEMM a = Y W plus one kernel per chunk).  scratch == NULL selects the direct
 *      column-walk form (D <= 256).  Datapoints with NaN in y always take the direct form.
 *      TVO_SEL_FITPARENTS (batch-global normaliser): one generation, with cfg->fit_norm and
 *      TVO_EVO_LPJ_CURRENT (re-score with tvo_bsc_lpj, reduce with tvo_fit_norm, then this call). ---- */
int64_t tvo_bsc_estep_scratch_bytes(int64_t B, int32_t H, int32_t is_f64);
/* Sparse states at 64 <= H <= 256 with randparents + sparseflip run the list-based kernel (states held as
 * sorted position lists, children scored incrementally); datapoints it cannot hold and all other shapes run
 * the generic kernel.  Both produce the same K.  generic_only != 0 disables the list-based kernel (tests). */
int tvo_bsc_estep_variant(int32_t generic_only);
int tvo_estep_evo_bsc_f64(const double *Y, int64_t ldY, const int64_t *idx, uint64_t *K, double *lpj,
                          const void *theta_ws, const tvo_evo_config *cfg, const double *sparsity, void *scratch,
                          int64_t scratch_bytes, int64_t B, int32_t S, int32_t H, int32_t D, int64_t *nsubs,
                          void *stream);
int tvo_estep_evo_bsc_f32(const float *Y, int64_t ldY, const int64_t *idx, uint64_t *K, float *lpj,
                          const void *theta_ws, const tvo_evo_config *cfg, const double *sparsity, void *scratch,
                          int64_t scratch_bytes, int64_t B, int32_t S, int32_t H, int32_t D, int64_t *nsubs,
                          void *stream);

/* ---- TVS (tvo/variational/tvs.py:47-81) and BSC.data_estimator (tvo/models/bsc.py:241-252) ----
 * tvo_state_marginals: out[b, h] = sum_s softmax(lpj[row])_s * K[row, s, h]  (mean_posterior of the states
 *      themselves, tvs.py:66-70), row = idx ? idx[b] : b; K packed with row stride k_stride_n words.
 * tvo_sample_states_probs: new states (b, S_off + j), j < S_new, of the packed buffer `out` (row stride
 *      out_stride_n words) with bit h ~ Bernoulli(probs[b * probs_stride_n + h]) (stride 0 = one row for all
 *      datapoints: the prior pies, tvs.py:60-63); 24-bit uniforms from the counter stream (seed, step, group,
 *      global datapoint index n_offset + row, j, h).  group < 64 separates the prior and marginal draws. */
int tvo_state_marginals_f64(const uint64_t *K, int64_t k_stride_n, const double *lpj, int64_t lpj_stride_n,
                            const int64_t *idx, double *out, int64_t B, int32_t S, int32_t H, void *stream);
int tvo_state_marginals_f32(const uint64_t *K, int64_t k_stride_n, const float *lpj, int64_t lpj_stride_n,
                            const int64_t *idx, float *out, int64_t B, int32_t S, int32_t H, void *stream);
int tvo_sample_states_probs_f64(uint64_t *out, int64_t out_stride_n, int32_t S_off, const double *probs,
                                int64_t probs_stride_n, const int64_t *idx, int64_t B, int32_t S_new, int32_t H,
                                uint64_t seed, uint32_t step, uint32_t group, uint32_t n_offset, void *stream);
int tvo_sample_states_probs_f32(uint64_t *out, int64_t out_stride_n, int32_t S_off, const float *probs,
                                int64_t probs_stride_n, const int64_t *idx, int64_t B, int32_t S_new, int32_t H,
                                uint64_t seed, uint32_t step, uint32_t group, uint32_t n_offset, void *stream);

/* ---- single-cause mixture models under full EM ----
 * GMM (tvo/models/gmm.py:91-180) and PMM (tvo/models/pmm.py:76-153) with FullEMSingleCauseModels
 * (tvo/variational/fullem.py:61-82: K_n = identity, every cause evaluated for every datapoint).
 * kind: 0 = GMM, 1 = PMM.  W is (D,H) row-major, pies (H), sigma2 (1, GMM only), all in the model precision.
 *   tvo_mixture_prepare: digest theta into `ws` (tvo_mixture_ws_bytes).
 *   tvo_mixture_lpj:     lpj_out[row, c] for all c < H, row = idx ? idx[b] : b  (log_pseudo_joint with K = identity).
 *   tvo_mixture_mstep:   update_param_batch: stats (tvo_mixture_stats_len doubles)
 *                        [Wp^T (H*D) | sum_n q (H) = my_Wq = my_pies | my_sigma2 | N | F]  +=  this batch;
 *                        `scratch` (tvo_mixture_scratch_bytes) holds the posterior weights of one chunk for the
 *                        Wp += Y^T q GEMM.
 *   tvo_mixture_free_energy: *F_out += sum_b logsumexp_c log_joint  (model_protocols.py:179-193). */
int64_t tvo_mixture_ws_bytes(int32_t H, int32_t D, int32_t is_f64);
int64_t tvo_mixture_stats_len(int32_t H, int32_t D);
int64_t tvo_mixture_scratch_bytes(int64_t B, int32_t H, int32_t D, int32_t is_f64);
int tvo_mixture_prepare_f64(int32_t kind, const double *W, const double *pies, const double *sigma2, int32_t H, int32_t D,
                            void *ws, void *stream);
int tvo_mixture_prepare_f32(int32_t kind, const float *W, const float *pies, const float *sigma2, int32_t H, int32_t D,
                            void *ws, void *stream);
int tvo_mixture_lpj_f64(int32_t kind, const double *Y, int64_t ldY, const int64_t *idx, const void *ws, double *lpj_out,
                        int64_t lpj_stride_n, int64_t B, int32_t H, int32_t D, void *stream);
int tvo_mixture_lpj_f32(int32_t kind, const float *Y, int64_t ldY, const int64_t *idx, const void *ws, float *lpj_out,
                        int64_t lpj_stride_n, int64_t B, int32_t H, int32_t D, void *stream);
int tvo_mixture_mstep_f64(int32_t kind, const double *Y, int64_t ldY, const int64_t *idx, const double *lpj,
                          int64_t lpj_stride_n, const void *ws, double *stats, void *scratch, int64_t scratch_bytes,
                          int64_t B, int32_t H, int32_t D, void *stream);
int tvo_mixture_mstep_f32(int32_t kind, const float *Y, int64_t ldY, const int64_t *idx, const float *lpj,
                          int64_t lpj_stride_n, const void *ws, double *stats, void *scratch, int64_t scratch_bytes,
                          int64_t B, int32_t H, int32_t D, void *stream);
int tvo_mixture_free_energy_f64(int32_t kind, const double *Y, int64_t ldY, const int64_t *idx, const double *lpj,
                                int64_t lpj_stride_n, const void *ws, int64_t B, int32_t H, int32_t D, double *F_out,
                                void *stream);
int tvo_mixture_free_energy_f32(int32_t kind, const float *Y, int64_t ldY, const int64_t *idx, const float *lpj,
                                int64_t lpj_stride_n, const void *ws, int64_t B, int32_t H, int32_t D, double *F_out,
                                void *stream);

/* ---- posterior weights: lpj2pjc / mean_posterior (_utils.py:182-230) ---- */
int tvo_lpj2pjc_f64(const double *lpj, double *pjc, int64_t N, int32_t S, void *stream);
int tvo_lpj2pjc_f32(const float *lpj, float *pjc, int64_t N, int32_t S, void *stream);
/* out[n, :] = sum_s softmax(lpj[n,:])_s g[n, s, :]   (g is (N,S,G)) */
int tvo_mean_posterior_f64(const double *g, const double *lpj, double *out, int64_t N, int32_t S, int64_t G,
                           void *stream);
int tvo_mean_posterior_f32(const float *g, const float *lpj, float *out, int64_t N, int32_t S, int64_t G,
                           void *stream);

/* ---- free energy: replaces Optimized.free_energy (model_protocols.py:179-193) with the model's
 *      log_joint constant (bsc.py:121-132).  *F_out (device double) is INCREMENTED by
 *      sum_b logsumexp_s(lpj[b,s]) + const_b, const_b = sum_h log(1-pi_h) - D_b/2 log(2 pi sigma2),
 *      D_b = #non-NaN of y_b (Y may be NULL => D_b = D). ---- */
int tvo_bsc_free_energy_f64(const double *Y, int64_t ldY, const double *lpj, int64_t lpj_stride_n,
                            const int64_t *idx, const void *theta_ws, int64_t B, int32_t S, int32_t D,
                            double *F_out, void *stream);
int tvo_bsc_free_energy_f32(const float *Y, int64_t ldY, const float *lpj, int64_t lpj_stride_n,
                            const int64_t *idx, const void *theta_ws, int64_t B, int32_t S, int32_t D,
                            double *F_out, void *stream);

/* ---- BSC M-step accumulation: replaces BSC.update_param_batch (bsc.py:134-158).  `stats` is one
 *      packed fp64 buffer [WpT (H*D, my_Wp transposed) | WqU (H*H, UPPER triangle incl. diagonal of
 *      the symmetric my_Wq; the strict lower triangle stays 0) | pies (H) | sigma2 (1) | N (1) | F (1)]
 *      that is INCREMENTED; q(s) = softmax(lpj) lives on chip only.  F gets the batch free energy
 *      (same value as tvo_bsc_free_energy).  `scratch` (see tvo_bsc_mstep_scratch_bytes) holds
 *      E = <s>_q for one chunk of datapoints so that WpT += E^T Y is one cuBLAS GEMM per chunk.
 *      exact_residual = 1 recomputes ||y - W s||^2 from Y and theta as the reference does
 *      (bsc.py:140-150); 0 derives it from lpj, valid when lpj was produced from the same Y and theta
 *      (the Trainer's E-step/M-step order, Trainer.py:239-249). ---- */
int64_t tvo_bsc_stats_len(int32_t H, int32_t D);
int64_t tvo_bsc_mstep_scratch_bytes(int64_t B, int32_t H, int32_t D, int32_t is_f64);
int tvo_bsc_mstep_f64(const double *Y, int64_t ldY, const int64_t *idx, const uint64_t *K, int64_t k_stride_n,
                      const double *lpj, int64_t lpj_stride_n, const void *theta_ws, double *stats, void *scratch,
                      int64_t scratch_bytes, int64_t B, int32_t S, int32_t H, int32_t D, int32_t exact_residual,
                      void *stream);
int tvo_bsc_mstep_f32(const float *Y, int64_t ldY, const int64_t *idx, const uint64_t *K, int64_t k_stride_n,
                      const float *lpj, int64_t lpj_stride_n, const void *theta_ws, double *stats, void *scratch,
                      int64_t scratch_bytes, int64_t B, int32_t S, int32_t H, int32_t D, int32_t exact_residual,
                      void *stream);

/* ---- NoisyOR (tvo/models/noisyor.py).  Y is uint8 (B, D), as the reference requires (noisyor.py:71).
 *      theta digestion: W (D,H), pies (H) -> (1 - W)^T padded, log-odds prior, sum log(1-pi).
 *      lpj: noisyor.py:67-106 incl. the all-zero-state rule (0 if y == 0 else -1e30).
 *      fused EVO E-step: same contract as tvo_estep_evo_bsc (direct form; no scratch).
 *      M-step: noisyor.py:108-138 into stats = [BtildeT (H*D) | CtildeT (H*D) | new_pi (H) | N | F]
 *      (fp64, INCREMENTED; F = sum logsumexp(lpj) + sum_h log(1-pi_h)). ---- */
int64_t tvo_noisyor_theta_ws_bytes(int32_t H, int32_t D, int32_t is_f64);
int64_t tvo_noisyor_stats_len(int32_t H, int32_t D);
int tvo_noisyor_prepare_f64(const double *W, const double *pies, int32_t H, int32_t D, void *theta_ws, void *stream);
int tvo_noisyor_prepare_f32(const float *W, const float *pies, int32_t H, int32_t D, void *theta_ws, void *stream);
int tvo_noisyor_lpj_f64(const uint8_t *Y, int64_t ldY, const uint64_t *K, int64_t k_stride_n, const int64_t *idx,
                        const void *theta_ws, double *lpj_out, int64_t lpj_stride_n, int64_t B, int32_t S, int32_t H,
                        int32_t D, void *stream);
int tvo_noisyor_lpj_f32(const uint8_t *Y, int64_t ldY, const uint64_t *K, int64_t k_stride_n, const int64_t *idx,
                        const void *theta_ws, float *lpj_out, int64_t lpj_stride_n, int64_t B, int32_t S, int32_t H,
                        int32_t D, void *stream);
int tvo_estep_evo_noisyor_f64(const uint8_t *Y, int64_t ldY, const int64_t *idx, uint64_t *K, double *lpj,
                              const void *theta_ws, const tvo_evo_config *cfg, const double *sparsity, int64_t B,
                              int32_t S, int32_t H, int32_t D, int64_t *nsubs, void *stream);
int tvo_estep_evo_noisyor_f32(const uint8_t *Y, int64_t ldY, const int64_t *idx, uint64_t *K, float *lpj,
                              const void *theta_ws, const tvo_evo_config *cfg, const double *sparsity, int64_t B,
                              int32_t S, int32_t H, int32_t D, int64_t *nsubs, void *stream);
int tvo_noisyor_mstep_f64(const uint8_t *Y, int64_t ldY, const int64_t *idx, const uint64_t *K, int64_t k_stride_n,
                          const double *lpj, int64_t lpj_stride_n, const void *theta_ws, double *stats, int64_t B,
                          int32_t S, int32_t H, int32_t D, void *stream);
int tvo_noisyor_mstep_f32(const uint8_t *Y, int64_t ldY, const int64_t *idx, const uint64_t *K, int64_t k_stride_n,
                          const float *lpj, int64_t lpj_stride_n, const void *theta_ws, double *stats, int64_t B,
                          int32_t S, int32_t H, int32_t D, void *stream);
/* *out (device double) += sum_b logsumexp_s lpj[b, s]  (model_protocols.py:191-193) */
int tvo_logsumexp_sum_f64(const double *lpj, int64_t stride_n, const int64_t *idx, int64_t B, int32_t S, double *out,
                          void *stream);
int tvo_logsumexp_sum_f32(const float *lpj, int64_t stride_n, const int64_t *idx, int64_t B, int32_t S, double *out,
                          void *stream);

/* ---- SSSC (tvo/models/sssc.py).  theta: W (D,H), sigma2 (1), mus (H), Psi (H,H), pies (H).
 *      lpj: the reformulated form (sssc.py:266-337) evaluated per state by Cholesky factorisations of
 *      the a x a active sub-blocks (a = popcount; at most 32, fewer if shared memory is short) — see
 *      csrc/sssc.cu.  States with more active units get lpj = finfo.min and set *overflow (device int)
 *      so that the host can fail loudly.  `scratch` (tvo_sssc_scratch_bytes) holds a = W^T y (and, for
 *      the M-step, <kappa>_q and <s>_q) for one chunk of datapoints; GEMMs on it are cuBLAS.
 *      M-step (sssc.py:366-452) stats layout (fp64, INCREMENTED):
 *      [y_szT^T (H*D) | sz_szT (H*H) | szszT (H*H) | ssT (H*H, upper triangle) | ssz (H*H) | s (H) | sz (H)
 *       | diag_yyT (D) | N | F].  `with_ssz` is a bit set: 1 = also accumulate ssz (reformulated_psi_update),
 *      2 = rows of Y with NaN take the restricted moments (data_estimator on incomplete data, sssc.py:563-597;
 *      the statistics are then not meaningful and the caller discards them). ---- */
int64_t tvo_sssc_theta_ws_bytes(int32_t H, int32_t D, int32_t is_f64);
int64_t tvo_sssc_stats_len(int32_t H, int32_t D);
int64_t tvo_sssc_scratch_bytes(int64_t B, int32_t H, int32_t D, int32_t is_f64);
int tvo_sssc_prepare_f64(const double *W, const double *sigma2, const double *mus, const double *Psi, const double *pies,
                         int32_t H, int32_t D, void *theta_ws, void *stream);
int tvo_sssc_prepare_f32(const float *W, const float *sigma2, const float *mus, const float *Psi, const float *pies,
                         int32_t H, int32_t D, void *theta_ws, void *stream);
int tvo_sssc_lpj_f64(const double *Y, int64_t ldY, const uint64_t *K, int64_t k_stride_n, const int64_t *idx,
                     const void *theta_ws, double *lpj_out, int64_t lpj_stride_n, void *scratch, int64_t scratch_bytes,
                     int32_t *overflow, int64_t B, int32_t S, int32_t H, int32_t D, void *stream);
int tvo_sssc_lpj_f32(const float *Y, int64_t ldY, const uint64_t *K, int64_t k_stride_n, const int64_t *idx,
                     const void *theta_ws, float *lpj_out, int64_t lpj_stride_n, void *scratch, int64_t scratch_bytes,
                     int32_t *overflow, int64_t B, int32_t S, int32_t H, int32_t D, void *stream);
int tvo_estep_evo_sssc_f64(const double *Y, int64_t ldY, const int64_t *idx, uint64_t *K, double *lpj,
                           const void *theta_ws, const tvo_evo_config *cfg, const double *sparsity, void *scratch,
                           int64_t scratch_bytes, int32_t *overflow, int64_t B, int32_t S, int32_t H, int32_t D,
                           int64_t *nsubs, void *stream);
int tvo_estep_evo_sssc_f32(const float *Y, int64_t ldY, const int64_t *idx, uint64_t *K, float *lpj,
                           const void *theta_ws, const tvo_evo_config *cfg, const double *sparsity, void *scratch,
                           int64_t scratch_bytes, int32_t *overflow, int64_t B, int32_t S, int32_t H, int32_t D,
                           int64_t *nsubs, void *stream);
int tvo_sssc_mstep_f64(const double *Y, int64_t ldY, const int64_t *idx, const uint64_t *K, int64_t k_stride_n,
                       const double *lpj, int64_t lpj_stride_n, const void *theta_ws, double *stats, void *scratch,
                       int64_t scratch_bytes, int32_t *overflow, int64_t B, int32_t S, int32_t H, int32_t D,
                       int32_t with_ssz, void *stream);
int tvo_sssc_mstep_f32(const float *Y, int64_t ldY, const int64_t *idx, const uint64_t *K, int64_t k_stride_n,
                       const float *lpj, int64_t lpj_stride_n, const void *theta_ws, double *stats, void *scratch,
                       int64_t scratch_bytes, int32_t *overflow, int64_t B, int32_t S, int32_t H, int32_t D,
                       int32_t with_ssz, void *stream);

/* ---- TVAE (tvo/models/tvae.py): the decoder pieces that touch the packed states.  The reference casts the uint8
 *      states to float and runs `states @ W_0 + b_0` (tvae.py:434-459, 672-701) and
 *      `states @ log(pi/(1-pi))` (tvae.py:337, 600) as dense products.
 *      tvo_tvae_layer0:       out (n_states x H1) = b0 + sum_{h in state} W0[h, :]   (W0: H0 x H1 row-major, b0 may
 *                             be NULL); states: n_states x ceil(H0/64) words, contiguous.
 *      tvo_tvae_layer0_grad:  gW0 (H0 x H1, INCREMENTED) += sum over states with unit h active of g (n_states x H1):
 *                             the weight gradient of the first layer (what autograd's `states.T @ g` would compute).
 *      tvo_tvae_lpj:          lpj[b, s] = sum_{h in state} prior[h] - s1 with prior = log(pi/(1-pi)),
 *                             kind 0 (GaussianTVAE, tvae.py:335-338): s1 = nansum_d (y - out)^2 / (2 sigma2[0]);
 *                             kind 1 (BernoulliTVAE, tvae.py:594-601): s1 = nansum_d BCE(out, y) (logs clamped at
 *                             -100 as torch does).  Y: B x D (row stride ldY); out: (B*S) x D contiguous decoder
 *                             output; states row b starts at states + b * st_stride_b; lpj likewise. ---- */
int tvo_tvae_layer0_f64(const uint64_t *states, int64_t n_states, int32_t H0, const double *W0, const double *b0,
                        int32_t H1, double *out, void *stream);
int tvo_tvae_layer0_f32(const uint64_t *states, int64_t n_states, int32_t H0, const float *W0, const float *b0,
                        int32_t H1, float *out, void *stream);
int tvo_tvae_layer0_grad_f64(const uint64_t *states, int64_t n_states, int32_t H0, const double *g, int32_t H1,
                             double *gW0, void *stream);
int tvo_tvae_layer0_grad_f32(const uint64_t *states, int64_t n_states, int32_t H0, const float *g, int32_t H1,
                             float *gW0, void *stream);
int tvo_tvae_lpj_f64(int32_t kind, const double *Y, int64_t ldY, const double *out, const uint64_t *states,
                     int64_t st_stride_b, const double *prior, const double *sigma2, double *lpj, int64_t lpj_stride_b,
                     int64_t B, int32_t S, int32_t H0, int32_t D, void *stream);
int tvo_tvae_lpj_f32(int32_t kind, const float *Y, int64_t ldY, const float *out, const uint64_t *states,
                     int64_t st_stride_b, const float *prior, const float *sigma2, float *lpj, int64_t lpj_stride_b,
                     int64_t B, int32_t S, int32_t H0, int32_t D, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TVO_B200_H */
