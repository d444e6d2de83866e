/*
 * stif_b200.h -- C ABI of libstif_b200.so: the B200 (sm_100a) implementation of STIF's
 * two data-parallel kernels.  Plain C, plain pointers and sizes, no torch types.
 *
 * The reference (johanneskopton/stif) is pure Python + numba and has no FFI of its
 * own; the seam this library plugs into is the set of Python call sites listed
 * beside each entry point (file:line into the reference tree).  INTEGRATION.md shows
 * the ctypes stub a maintainer would add at each of them.
 *
 * Conventions
 *   - every function returns 0 on success or a negative STIF_E* code and never throws;
 *     stif_last_error() returns a thread-local message for the last failure.
 *   - "_dev" entry points take DEVICE pointers (e.g. torch tensors' data_ptr()) and a
 *     cudaStream_t passed as void* (NULL = the context's own stream; pass cudaStreamLegacy,
 *     (void*)1, for the legacy default stream); they enqueue work and return without
 *     synchronising.  The matching host-pointer entry points copy
 *     in, run, copy out and synchronise.
 *   - observations are structure-of-arrays float64 (x, y, t, value).  int64 times and
 *     bool values are converted by the caller; the conversion is exact below 2^53.
 *   - a context is bound to one CUDA device and is not thread-safe; different contexts may be
 *     used from different host threads concurrently (one process per GPU under torchrun, or one
 *     thread per GPU in a single process: stif_b200/multi.py).
 */
#ifndef STIF_B200_H
#define STIF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STIF_OK 0
#define STIF_EINVAL (-1)     /* bad argument                      -> ValueError        */
#define STIF_ECUDA (-2)      /* CUDA runtime error                -> RuntimeError      */
#define STIF_ENOMEM (-3)     /* device allocation failed          -> MemoryError       */
#define STIF_ESTATE (-4)     /* call-order misuse (no obs/model)  -> ValueError        */
#define STIF_ENONFINITE (-5) /* non-finite kriging system         -> LinAlgError       */
/* STIF_ENONFINITE is returned by stif_kriging_predict_host when a kriging system holds NaN or
 * Inf: the reference's lstsq raises np.linalg.LinAlgError there (stif/predictor.py:587, numba's
 * _check_finite_matrix).  The outputs are still written (NaN for the affected targets). */

typedef struct stif_ctx stif_ctx;

/* ---- context ---------------------------------------------------------------- */

/* Number of CUDA devices visible to the process (0 if there is none / no driver). */
int stif_device_count(void);
/* Create a context on CUDA device `device` (-1 = current device).  SURVEY 8(b) proposed one
 * context over a device list, stif_ctx_create(devices, ndev); the C ABI keeps ONE device per
 * context (its calls enqueue on one stream and return), and a process that wants every GPU of
 * the box creates one context per device and drives them from one host thread each --
 * stif_b200/multi.py (MultiContext) does exactly that behind Predictor, no torchrun needed. */
int stif_ctx_create(int device, stif_ctx **out);
int stif_ctx_destroy(stif_ctx *ctx);
const char *stif_last_error(void);
/* Library/ABI version: major*10000 + minor*100 + patch. */
int stif_version(void);
/* Number of SMs of the context's device (grid sizing is a multiple of this). */
int stif_sm_count(stif_ctx *ctx);

/* ---- empirical space-time variogram / covariogram ----------------------------
 * Replaces stif.utils.get_variogram (stif/utils.py:104-190) and get_covariogram
 * (stif/utils.py:193-279), all-pairs branch of _pair_index_generator
 * (stif/utils.py:87-92), as called from Predictor.calc_empirical_variogram
 * (stif/predictor.py:398-408) and calc_empirical_covariogram (:471-481).
 *
 * Outputs are the RAW accumulators: counts[ns*nt] (uint64, row-major [space][time])
 * and sums[ns*nt] (double: sum of (vi-vj)^2, or of (vi-mean)(vj-mean) in covariogram
 * mode).  The caller forms variogram = sums*0.5/counts (or sums/counts/var), NaN where
 * counts == 0, exactly as stif/utils.py:178-188.
 *
 * mode:   0 = variogram, 1 = covariogram (`mean` = np.mean(val), utils.py:234).
 * flags:  STIF_VARIO_PEEL  reproduce the reference's executed arithmetic for its first
 *                          pair (rows 0 and 1), which numba/LLVM peels out of the loop
 *                          and bins with a true division instead of the reciprocal
 *                          multiply used for every other pair;
 *         STIF_VARIO_NO_CULL  visit every tile pair (disables the time-sorted tile
 *                          culling; results are identical, only slower);
 *         STIF_VARIO_SPACE_SORT / STIF_VARIO_TIME_SORT  force / forbid the space-time sort
 *                          mode (observations ordered by time slab and space cell, tile
 *                          pairs culled by their bounding boxes in space and time; default:
 *                          used from 163 840 observations).  Counts are identical in every
 *                          mode, the sums differ by the summation order only.
 * part / nparts: this call accumulates only the work items of partition `part` of
 *         `nparts` (multi-GPU: one partition per rank; the caller sums the partial
 *         histograms, e.g. with one NCCL all-reduce).  The PEEL fix-up is applied by
 *         partition 0 only.
 */
#define STIF_VARIO_PEEL 1
#define STIF_VARIO_NO_CULL 2
#define STIF_VARIO_SPACE_SORT 4
#define STIF_VARIO_TIME_SORT 8
#define STIF_VARIO_DETERMINISTIC 32 /* bit-reproducible sums: every term is added as a 62-bit fixed-point
                                    integer (quantum = 2^-60 of the largest possible term) into integer
                                    histograms instead of FP64 atomics, whose order -- hence the last
                                    bits of the sums -- depends on the scheduling.  ~25 % slower. */
#define STIF_VARIO_COUNTS_F64 16 /* _dev only: counts_dev receives float64 values (exact below 2^53,
                                    the dtype of the reference's own `norm`), so that counts and sums
                                    of all ranks meet in ONE all-reduce of one float64 buffer */

int stif_variogram_dev(stif_ctx *ctx, const double *x, const double *y, const double *t,
                       const double *v, int64_t n, double space_dist_max,
                       double time_dist_max, int n_space_bins, int n_time_bins, int mode,
                       double mean, int flags, int part, int nparts,
                       uint64_t *counts_dev, double *sums_dev, void *stream);

int stif_variogram_host(stif_ctx *ctx, const double *x, const double *y, const double *t,
                        const double *v, int64_t n, double space_dist_max,
                        double time_dist_max, int n_space_bins, int n_time_bins, int mode,
                        double mean, int flags, int part, int nparts,
                        uint64_t *counts_host, double *sums_host);

/* The same over the BOUND observation set (stif_kriging_set_obs / stif_kriging_set_residuals:
 * coordinates and residuals already on the device), optionally restricted to idx_host[n_idx]
 * (host int64 indices into the bound set, numpy-style negatives allowed; NULL = all of it, n_idx
 * ignored).  What Predictor.calc_cross_validation needs per fold (stif/predictor.py:293-296,
 * calc_empirical_variogram(train)): the coordinates are uploaded once per Predictor, a fold that
 * refits the covariate model replaces the residuals only, and a subset travels as 8 B per
 * selected observation.  STIF_VARIO_PEEL refers to rows idx[0], idx[1]. */
int stif_variogram_bound(stif_ctx *ctx, const int64_t *idx_host, int64_t n_idx, double space_dist_max,
                         double time_dist_max, int n_space_bins, int n_time_bins, int mode, double mean,
                         int flags, int part, int nparts, uint64_t *counts_host, double *sums_host);

/* Sampled-pair branch of _pair_index_generator (stif/utils.py:94-101), reached when the
 * caller passes el_max: n_samples draws of an ordered pair (i, j) with replacement, each
 * index uniform in [0, n); draws with i == j are skipped but still count.  The reference's
 * RNG (numba MT19937) is seeded from the OS, so its results are not reproducible and parity
 * is statistical only; this implementation is reproducible for a given `seed`
 * (counter-based Philox4x32-10, one counter per draw). */
int stif_variogram_sampled_dev(stif_ctx *ctx, const double *x, const double *y, const double *t,
                               const double *v, int64_t n, double space_dist_max,
                               double time_dist_max, int n_space_bins, int n_time_bins, int mode,
                               double mean, uint64_t n_samples, uint64_t seed,
                               uint64_t *counts_dev, double *sums_dev, void *stream);

int stif_variogram_sampled_host(stif_ctx *ctx, const double *x, const double *y, const double *t,
                                const double *v, int64_t n, double space_dist_max,
                                double time_dist_max, int n_space_bins, int n_time_bins, int mode,
                                double mean, uint64_t n_samples, uint64_t seed,
                                uint64_t *counts_host, double *sums_host);

/* Statistics of the last variogram call on this context (after a synchronise):
 * out[0] = algorithmic pairs of this partition's share (n(n-1)/2 when nparts == 1),
 * out[1] = pairs actually evaluated (tile pairs visited x tile area),
 * out[2] = pairs that passed both cut-offs (binned or dropped by the bounds check),
 * out[3] = work items (tile pairs) executed. */
int stif_variogram_stats(stif_ctx *ctx, uint64_t out[4]);

/* ---- space-time ordinary kriging ----------------------------------------------
 * Replaces Data.get_kriging_idxs (stif/data.py:170-212), the nd_kriging closure
 * (stif/predictor.py:602-660), calc_kriging_weights (:566-593), the variogram_model
 * closure (:547-556 over stif/variogram_models.py:7-218) and the epilogue of
 * _get_kriging_prediction_batch (:882-906), as called from
 * Predictor._get_kriging_weights (:777-804) and get_kriging_prediction (:806-880).
 */

/* Space-time model ids (stif/variogram_models.py:401-409). */
#define STIF_ST_SUM 0
#define STIF_ST_PRODUCT 1
#define STIF_ST_PRODUCT_SUM 2
#define STIF_ST_METRIC 3
#define STIF_ST_SUM_METRIC 4
/* Marginal model ids. */
#define STIF_M_SPHERICAL 0
#define STIF_M_GAUSSIAN 1

/* Upload (host pointers) / bind (device pointers are copied) the observation set:
 * coordinates, times and residuals of the training data (stif/predictor.py:799-804
 * passes self._data.space_coords / time_coords; :900 reads self._residuals). */
int stif_kriging_set_obs(stif_ctx *ctx, const double *x, const double *y, const double *t,
                         const double *resid, int64_t n_obs, int on_device);

/* Data.get_kriging_idxs (stif/data.py:170-212) as a stand-alone call: every (observation,
 * target) pair with  ||s_o - s_t||^2 < space_dist_max^2,  |t_o - t_t| < time_dist_max,
 * t_o <= t_t  and  o != leave_out[target], found through the space-time grid (never the dense
 * n_obs x n_targets matrices), sorted by observation index then target index like
 * np.nonzero.  Targets are host pointers.  _count runs the search and returns the number of
 * pairs; _fetch copies them out as int64 [n_pairs][2] = {obs_idx, target_idx}. */
int stif_kriging_pairs_count(stif_ctx *ctx, const double *tx, const double *ty, const double *tt,
                             const int64_t *leave_out, int64_t n_targets, double space_dist_max,
                             double time_dist_max, int64_t *n_pairs);
int stif_kriging_pairs_fetch(stif_ctx *ctx, int64_t *pairs_out);

/* Replace only the residuals of the bound observation set (same n_obs); coordinates and the
 * cached space-time grid stay.  Cross-validation refits the covariate model per fold
 * (stif/predictor.py:275, :211-229), which changes the residuals but not the coordinates. */
int stif_kriging_set_residuals(stif_ctx *ctx, const double *resid, int64_t n_obs, int on_device);

/* Fitted variogram model: ids + parameter vector x of Predictor._variogram_fit.x
 * (stif/predictor.py:755-758).  nparams: sum/product 6, product_sum 7, metric 4,
 * sum_metric 10. */
int stif_kriging_set_model(stif_ctx *ctx, int st_model, int space_model, int time_model,
                           int metric_model, const double *params, int nparams);

/* Evaluate gamma(h, t) of the current model elementwise (host pointers); replaces the
 * numba ufunc of stif/predictor.py:547-556 for _get_variogram_model_grid (:766-775). */
int stif_variogram_model_eval(stif_ctx *ctx, const double *h, const double *t, int64_t n,
                              double *gamma_out);

/* calc_kriging_weights (stif/predictor.py:566-593) as a stand-alone call -- the callable the
 * reference installs as Predictor._kriging_weights_function: ordinary-kriging weights (float64) of ONE
 * neighbour set from its coordinates xs, ys, ts [k] and gamma-vector [k] under the current model.
 * Solved like np.linalg.lstsq (SVD of the bordered matrix [Gamma 1; 1' 0], minimum-norm solution if it
 * is rank-deficient).  Host pointers, k <= 128.  Not a hot path: get_kriging_prediction assembles and
 * solves its systems inside stif_kriging_predict_*. */
int stif_kriging_weights_host(stif_ctx *ctx, const double *xs, const double *ys, const double *ts,
                              const double *gamma_vec, int k, double *w_out);

/* Objective of the variogram-model fit (stif/variogram_models.py:221-308: prediction_grid +
 * weighted_mean_square_error, called by scipy's Nelder-Mead from Predictor.fit_variogram_model,
 * stif/predictor.py:738-753).  The optimiser stays on the host; only the objective moves.
 * stif_fit_set_grid uploads the lag-bin centres, the empirical variogram (NaN allowed) and the
 * weights (host pointers, row-major [n_space][n_time]) once per fit; stif_fit_wmse evaluates
 *   sum(w * (empirical - gamma(h, t; params))^2) / sum(w)
 * for one parameter vector (model ids as in stif_kriging_set_model) and returns it in *out
 * (one small kernel + an 8-byte read-back; NaN propagates like np.average). */
int stif_fit_set_grid(stif_ctx *ctx, const double *bins_space, int n_space, const double *bins_time,
                      int n_time, const double *empirical, const double *weights);
int stif_fit_wmse(stif_ctx *ctx, int st_model, int space_model, int time_model, int metric_model,
                  const double *params, int nparams, double *out);

/* flags for stif_kriging_predict_* */
#define STIF_KRIG_F32_STORAGE 1 /* round weights and gamma-vector to float32 before the
                                   mean/std reductions and reproduce numpy's pairwise
                                   f32/f64 summation (stock reference, predictor.py:610-618,
                                   :900-905).  Without it everything stays float64. */
#define STIF_KRIG_FORCE_LU 2    /* solve every system with the pivoted LU (skip the covariance-form
                                   Cholesky fast path that valid variogram models take) */
#define STIF_KRIG_TRY_CHOL 4    /* try the Cholesky fast path for any model (targets whose
                                   covariance matrix is not positive definite still go to the LU) */
#define STIF_KRIG_TRY_LDL 8     /* solve every model's systems with the symmetric indefinite solver
                                   (Bunch-Kaufman LDL' of the bordered system, one warp per target) that
                                   models which are not valid variograms -- product, product_sum,
                                   a negative total nugget -- take by default; targets with a tiny pivot
                                   or a non-finite entry still go to the LU.  STIF_KRIG_FORCE_LU wins. */

/* Predict n_targets points.  leave_out may be NULL; entry < 0 = none
 * (stif/data.py:204-205).  mean_out/std_out: n_targets doubles.  Optional outputs
 * (NULL to skip): w_out, gamma_out float32 [n_targets*max_pts], idx_out uint32
 * [n_targets*max_pts] (zero padded, same row order as the reference: ascending
 * gamma when truncated to max_pts, else ascending observation index), n_used_out
 * int32 [n_targets] (-1 = fewer than min_pts candidates -> zero weights).
 * Pointers are host pointers for _host, device pointers for _dev. */
int stif_kriging_predict_host(stif_ctx *ctx, const double *tx, const double *ty,
                              const double *tt, const int64_t *leave_out, int64_t n_targets,
                              int min_pts, int max_pts, double space_dist_max,
                              double time_dist_max, int flags, double *mean_out,
                              double *std_out, float *w_out, float *gamma_out,
                              uint32_t *idx_out, int32_t *n_used_out);

int stif_kriging_predict_dev(stif_ctx *ctx, const double *tx, const double *ty,
                             const double *tt, const int64_t *leave_out, int64_t n_targets,
                             int min_pts, int max_pts, double space_dist_max,
                             double time_dist_max, int flags, double *mean_out,
                             double *std_out, float *w_out, float *gamma_out,
                             uint32_t *idx_out, int32_t *n_used_out, void *stream);

/* Statistics of the last kriging call (after a synchronise):
 * out[0] = candidates examined by the neighbour search, out[1] = candidates inside the
 * search window, out[2] = targets with fewer than min_pts candidates,
 * out[3] = targets whose system was rank-deficient (duplicate / co-located observations with a
 *          zero total nugget): they are solved by the minimum-norm path. */
int stif_kriging_stats(stif_ctx *ctx, uint64_t out[4]);

/* out[0] = targets of the last call that the first-pass solver (the covariance-form Cholesky
 * kernel, or the symmetric indefinite LDL' kernel) handed to the pivoted LU (not positive
 * definite / tiny pivot / non-finite); out[1] = bit 0: the Cholesky path was enabled for the
 * call's model, bit 1: the LDL' path was. */
int stif_kriging_solver_stats(stif_ctx *ctx, uint64_t out[2]);

/* Device time in ms of the phases of the last call, measured with CUDA events on the
 * launching stream: variogram: [0]=sort/prepare [1]=pair kernel [2]=reduce;
 * kriging: [0]=grid build [1]=neighbour search [2]=assemble+solve [3]=epilogue.
 * Synchronises the stream. */
int stif_last_phase_ms(stif_ctx *ctx, float out[4]);

/* Test hook: the device implementation of numpy's pairwise row summation that the float32-storage
 * epilogue uses for np.sum(kriging_weights * residuals[idx], axis=1) and
 * np.sum(kriging_weights * kriging_vectors, axis=1) (stif/predictor.py:900-905), applied to n_rows
 * rows of k elements (host pointers; float32 if is_f32 else float64; k <= 128).  Lets the tests
 * compare the DEVICE code with np.sum bit for bit. */
int stif_debug_numpy_row_sum(stif_ctx *ctx, const void *rows, int n_rows, int k, int is_f32, void *out);

/* ---- measurement helpers ------------------------------------------------------ */

/* Register-resident FP64 FMA microbenchmark: the measured FP64 issue peak that the
 * roofline fractions are quoted against (MEASURED_PEAKS.json has no FP64 figure).
 * Returns achieved FP64 lane-instructions per second in *lane_instr_per_s. */
int stif_bench_dfma(stif_ctx *ctx, int iters, double *lane_instr_per_s, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* STIF_B200_H */
