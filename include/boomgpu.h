/*
 * boomgpu.h -- C ABI of the B200-native auxiliary-mixture Gibbs hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, int
 * status (0 = ok) plus boomgpu_last_error().  The C++ sampler classes that keep
 * BOOM's PosteriorSampler::draw() surface (boom_b200/host, boom_b200/boom_adapter)
 * are the only intended callers; INTEGRATION.md shows the binding.
 *
 * What each entry point replaces in the reference (steve-the-bayesian/BOOM,
 * paths relative to its root):
 *
 *   boomgpu_upload_binomial / _poisson
 *       the AoS of heap objects behind IID_DataPolicy::dat()
 *       (Models/Policies/IID_DataPolicy.hpp:43-58, Models/Glm/BinomialRegressionData.hpp:25-55,
 *        Models/Glm/PoissonRegressionData.hpp:25-62) becomes one row-major matrix in HBM.
 *   boomgpu_set_logit_mixture
 *       BinomialLogitDataImputer::mixture_approximation
 *       (Models/Glm/PosteriorSamplers/BinomialLogitDataImputer.hpp:51, NormalMixtureApproximation.cpp:416-424)
 *   boomgpu_set_poisson_table
 *       PoissonDataImputer::mixture_table_ (PoissonDataImputer.hpp:99, serialize format
 *       NormalMixtureApproximation.cpp:393-399,534-542)
 *   boomgpu_logit_step
 *       LatentDataSampler::impute_latent_data for the logit samplers, i.e. the worker-pool loop
 *       Models/PosteriorSamplers/Imputer.hpp:175-180 + Imputer.cpp:27-66 over
 *       ImputeWorker::impute_latent_data_point (BinomialLogitAuxmixSampler.cpp:77-97):
 *       GlmCoefs::predict (GlmCoefs.cpp:134-156), BinomialLogitCltDataImputer::impute
 *       (BinomialLogitDataImputer.cpp:122-210) and SufficientStatistics::update/combine
 *       (BinomialLogitAuxmixSampler.cpp:44-67).
 *   boomgpu_poisson_step
 *       the same framework over PoissonRegressionDataImputer::impute_latent_data_point
 *       (PoissonRegressionAuxMixSampler.cpp:58-81): PoissonDataImputer::impute
 *       (PoissonDataImputer.cpp:36-98) and WeightedRegSuf::add_data/combine
 *       (Models/Glm/WeightedRegressionModel.cpp:69-87,161-169).
 *   boomgpu_binomial_loglike / boomgpu_poisson_loglike (value only), boomgpu_*_loglike_derivs (value, gradient, Hessian)
 *       BinomialLogitModel::log_likelihood (Models/Glm/BinomialLogitModel.cpp:140-180),
 *       PoissonRegressionModel::log_likelihood (Models/Glm/PoissonRegressionModel.cpp:56-95).
 *
 * Threading: one host thread per context; a context is bound to one CUDA device.
 * Multi-GPU = one context (one process) per GPU, rows sharded, the packed
 * statistics of the *_step_device variants all-reduced by the caller (NCCL).
 * Ownership: the caller owns every host buffer; the context owns device memory
 * it allocated (uploaded data, workspaces) but NOT adopted device pointers.
 * There is no CPU fallback: every compute entry point fails if CUDA fails.
 */
#ifndef BOOMGPU_H
#define BOOMGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#pragma GCC visibility push(default)

typedef struct boomgpu_ctx boomgpu_ctx;

#define BOOMGPU_OK 0
#define BOOMGPU_ERR_CUDA 1
#define BOOMGPU_ERR_ARG 2
#define BOOMGPU_ERR_STATE 3
#define BOOMGPU_ERR_DATA 4 /* device-side input validation failed (y > n, NaN eta, nu missing from table ...) */

/* ---- lifetime ------------------------------------------------------------------------ */
int boomgpu_create(boomgpu_ctx **ctx, int device);
void boomgpu_destroy(boomgpu_ctx *ctx);
/* message of the last failing call on ctx (ctx == NULL: last failing boomgpu_create) */
const char *boomgpu_last_error(const boomgpu_ctx *ctx);
const char *boomgpu_version(void);

/* cudaStream_t to launch on (default: a stream owned by the context). */
int boomgpu_set_stream(boomgpu_ctx *ctx, void *cuda_stream);
/* global index of this shard's first row: keys the Philox counters so draws do not depend on the sharding */
int boomgpu_set_row_offset(boomgpu_ctx *ctx, uint64_t first_global_row);
/* options: "path" = 0 auto | 1 fused single pass (p <= 64) | 2 two-pass imputer + DMMA SYRK;
 *          "small_variant" = 0 auto (the TMA-fed warp-autonomous kernel when X has an even leading dimension and a 16-byte
 *                            aligned base; for 32 < p <= 48, and the Poisson model up to p = 56, its 12-warp form that parks
 *                            the accumulators in tensor memory between DMMA phases) | 1 force the cp.async kernel |
 *                            2 the TMA kernel, never parked | 3 the warp-specialised kernel for 32 < p <= 64 (accumulate
 *                            warps + draw warps: measured no faster, kept as an experiment) | 4 parked wherever 32 < p <= 64;
 *          "single_launch" = 1 (default) the small-p step is one kernel whose last CTA sums the per-CTA partials | 0 a
 *                            separate reduction kernel;
 *          "gather" = 0 auto (two-pass path: a beta with fewer than p / 4 non-zeros reads only those columns of X in the
 *                     imputer pass) | 1 never | 2 whenever beta has a zero;
 *          "syrk_order" = 1 (default) the split-K SYRK's off-diagonal regions are scheduled first and the cheaper diagonal
 *                         regions last | 0 k-slice major | 2 k-slice major over uniform work items (diagonal regions in pairs on
 *                         one CTA: less DRAM traffic with short CTAs, 1 % slower); "syrk_waves" = CTAs per SM the split-K aims
 *                         for (default 30); "syrk_rdiag" = 1 (default) the ragged last column block (p not a multiple of 128)
 *                         runs in its own strip-form kernel | 0 in the main kernel's 128 x 128 unit form;
 *          "tma_promotion" = 3 (default; 0 none, 1 64 B, 2 128 B, 3 256 B): L2 promotion of the SYRK's tensor map;
 *          "syrk_diag" = 0 (default) strip form for whole diagonal regions | 1 unit form; "syrk_filter": profiling aid;
 *          "syrk_cluster" = c > 1: launch the SYRK with thread-block clusters of c CTAs (experiment; measured slower);
 *          "timing" = 1 records CUDA events around every kernel (boomgpu_get_timings) */
int boomgpu_set_option(boomgpu_ctx *ctx, const char *name, int64_t value);

/* ---- data ---------------------------------------------------------------------------- */
/* host -> device; X row major, n x p, leading dimension ldx >= p */
int boomgpu_upload_binomial(boomgpu_ctx *ctx, int64_t n, int p, const double *X, int64_t ldx,
                            const double *y, const double *ntrials);
int boomgpu_upload_poisson(boomgpu_ctx *ctx, int64_t n, int p, const double *X, int64_t ldx,
                           const int64_t *y, const double *exposure);
/* The same upload a chunk of rows at a time, for callers whose rows are not one contiguous host matrix (BOOM: one heap
 * object per observation, IID_DataPolicy.hpp:57-58): begin allocates, rows copies rows [row0, row0 + nrows) and returns when
 * the caller's chunk buffers may be re-used, end makes the data usable.  y: nrows doubles (binomial) or int64 (poisson);
 * aux: trials / exposure.  poisson = 2: plain regression rows for the Student-t sibling (y doubles, aux ignored / NULL). */
int boomgpu_upload_begin(boomgpu_ctx *ctx, int poisson, int64_t n, int p);
int boomgpu_upload_rows(boomgpu_ctx *ctx, int64_t row0, int64_t nrows, const double *X, int64_t ldx, const void *y, const double *aux);
int boomgpu_upload_end(boomgpu_ctx *ctx);
/* data already resident in HBM (device pointers; caller keeps them alive).  Used in place when TMA can describe them
 * (16-byte aligned base, even ldx); for p > 64 rows that do not qualify (e.g. a contiguous n x 501 tensor) are copied ONCE,
 * at the first step, into a padded device buffer owned by the context -- a second copy of X in HBM. */
int boomgpu_adopt_binomial(boomgpu_ctx *ctx, int64_t n, int p, const double *dX, int64_t ldx,
                           const double *dy, const double *dntrials);
int boomgpu_adopt_poisson(boomgpu_ctx *ctx, int64_t n, int p, const double *dX, int64_t ldx,
                          const int64_t *dy, const double *dexposure);

int boomgpu_set_logit_mixture(boomgpu_ctx *ctx, int K, const double *mu, const double *sigma,
                              const double *weights);
/* ntab entries sorted by nu; entry e owns components offset[e] .. offset[e+1]-1 of the flat arrays */
int boomgpu_set_poisson_table(boomgpu_ctx *ctx, int ntab, const int64_t *nu, const int32_t *offset,
                              const double *weights, const double *mu, const double *sigma,
                              int64_t gaussian_cutoff);

/* present[v] = 1 iff some uploaded / adopted Poisson row has count y == v, for 0 <= v < len (host array of len bytes).
 * The reference extends its table lazily, per observation, inside the draw (NormalMixtureApproximationTable::approximate,
 * NormalMixtureApproximation.cpp:472-532); here the host asks once per data set which counts occur, adds the missing
 * entries by the same rule, and re-states the table. */
int boomgpu_poisson_counts_present(boomgpu_ctx *ctx, unsigned char *present, int64_t len);

/* ---- the hot path -------------------------------------------------------------------- */
/* One imputation pass.  xtx: p x p (symmetric, both triangles filled), xty: p.  Synchronous. */
int boomgpu_logit_step(boomgpu_ctx *ctx, const double *beta, int clt_threshold, uint64_t seed,
                       uint64_t iteration, double *xtx, double *xty, int64_t *sample_size);
/* scalars = {n, y'Wy, sum w, sum log w} as WeightedRegSuf keeps them */
int boomgpu_poisson_step(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration,
                         double *xtwx, double *xtwy, double scalars[4]);

/* The probit sibling -- BinomialProbitSpikeSlabSampler::impute_latent_data / refresh_xtx
 * (Models/Glm/PosteriorSamplers/BinomialProbitSpikeSlabSampler.cpp:58-83) over BinomialProbitDataImputer::impute
 * (BinomialProbitDataImputer.cpp:31-68), on binomial data uploaded with boomgpu_upload_binomial / _adopt_binomial:
 * xtx = sum n_i x x' (it does not depend on beta: pass xtx = NULL after the first call and only X'z -- one HBM-bound
 * pass over X -- is computed), xtz = sum x_i (sum of the n_i latent normals of observation i). */
int boomgpu_probit_step(boomgpu_ctx *ctx, const double *beta, int clt_threshold, uint64_t seed, uint64_t iteration,
                        double *xtx, double *xtz, int64_t *sample_size);
int boomgpu_probit_step_device(boomgpu_ctx *ctx, const double *beta, int clt_threshold, uint64_t seed, uint64_t iteration,
                               double *suf_dev, int xtz_only);

/* The Student-t sibling -- TRegressionSampler::impute_latent_data (Models/Glm/PosteriorSamplers/TRegressionSampler.cpp:128-143)
 * over TDataImputer::impute (TDataImputer.cpp:25-29), on plain regression rows (boomgpu_upload_regression / _adopt_regression,
 * or boomgpu_upload_begin with kind 2): w_i ~ Gamma((nu + 1)/2, rate (nu + ((y_i - x_i'beta)/sigma)^2)/2), then
 * WeightedRegSuf::add_data(x_i, y_i, w_i) (WeightedRegressionModel.cpp:161-169): xtwx = X'WX, xtwy = X'Wy,
 * scalars = {n, y'Wy, sum w, sum log w} (the last two are also the GammaSuf of the sampler's weight model). */
int boomgpu_upload_regression(boomgpu_ctx *ctx, int64_t n, int p, const double *X, int64_t ldx, const double *y);
int boomgpu_adopt_regression(boomgpu_ctx *ctx, int64_t n, int p, const double *dX, int64_t ldx, const double *dy);
int boomgpu_student_step(boomgpu_ctx *ctx, const double *beta, double sigma, double nu, uint64_t seed, uint64_t iteration,
                         double *xtwx, double *xtwy, double scalars[4]);
int boomgpu_student_step_device(boomgpu_ctx *ctx, const double *beta, double sigma, double nu, uint64_t seed, uint64_t iteration,
                                double *suf_dev);
/* active-set form, as boomgpu_logit_step_active below (p > 64): G = (X'WX)[:, active], diag, X'Wy; scalars as above */
int boomgpu_student_step_active(boomgpu_ctx *ctx, const double *beta, double sigma, double nu, uint64_t seed, uint64_t iteration,
                                const int32_t *active, int k, double *G, double *diag, double *xty, double scalars[4]);
/* TRegressionModel::log_likelihood(beta, sigsq, nu) (Models/Glm/TRegression.cpp:74-86) = sum_i log dstudent(y_i; x_i'beta,
 * sigma, nu), all-reduced over the shards.  beta != NULL: one pass over X, the residuals stay on the device; beta == NULL:
 * the residuals of the previous call are reused (8 n bytes per evaluation): what the slice sampler on nu
 * (TRegressionSampler.cpp:173-176) needs, several evaluations per draw with beta and sigma fixed. */
int boomgpu_student_loglike(boomgpu_ctx *ctx, const double *beta, double sigma, double nu, double *loglike);

/* ACTIVE-SET form of the step (p > 64).  A sweep over the inclusion indicators reads of X'WX only the columns of the
 * variables in the model and the diagonal (log_model_prob selects sub-blocks, BinomialLogitSpikeSlabSampler.cpp:88-117), so the
 * device computes, for the column set `active` (k <= 128 columns, normally the included variables):
 *   G[j * k + a] = (X'WX)[j, active[a]]  (p x k, row major),  diag[j] = (X'WX)[j, j],  xty = X'Wz
 * -- n p (2 k + 4) flops and one read of X instead of the n p (p + 1) flops of the full matrix.  The latents (w_i, s_i) stay
 * in HBM until the next step: boomgpu_weighted_column returns column j of X'WX for them (what the sweep needs when it ADDS a
 * variable outside `active`: one more pass over X), boomgpu_full_statistics the whole matrix (for suf() consumers). */
int boomgpu_logit_step_active(boomgpu_ctx *ctx, const double *beta, int clt_threshold, uint64_t seed, uint64_t iteration,
                              const int32_t *active, int k, double *G, double *diag, double *xty, int64_t *sample_size);
int boomgpu_poisson_step_active(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, const int32_t *active,
                                int k, double *G, double *diag, double *xty, double scalars[4]);
int boomgpu_weighted_column(boomgpu_ctx *ctx, int j, double *column);
int boomgpu_full_statistics(boomgpu_ctx *ctx, double *xtx, double *xty);

/* Page-locks a caller-owned host range (cudaHostRegister).  When the xtx / xtwx argument of the synchronous steps points
 * into such a range, the p x p matrix is copied device->host straight into it (at p = 4000 it is 128 MB: the staging
 * copy it saves costs ~15 ms per iteration and per rank).  ctx-free: errors are reported through boomgpu_last_error(NULL). */
int boomgpu_pin_host(void *ptr, uint64_t bytes);
int boomgpu_unpin_host(void *ptr);

/* Asynchronous variants: the packed statistics stay in HBM at suf_dev (device pointer,
 * boomgpu_suf_len(p) doubles: [p*p matrix | p vector | 4 scalars]; logit scalars = {sample_size,0,0,0}),
 * ready for an in-place all-reduce on the same stream.  beta is a HOST pointer. */
int64_t boomgpu_suf_len(int p);
int boomgpu_logit_step_device(boomgpu_ctx *ctx, const double *beta, int clt_threshold, uint64_t seed,
                              uint64_t iteration, double *suf_dev);
int boomgpu_poisson_step_device(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration,
                                double *suf_dev);
/* waits for the context's stream and reports device-side validation errors of the steps since the last call */
int boomgpu_synchronize(boomgpu_ctx *ctx);
/* a context-owned device buffer of boomgpu_suf_len(p) doubles for the *_step_device variants */
int boomgpu_suf_buffer(boomgpu_ctx *ctx, double **suf_dev);
/* device -> host copy of count doubles on the context's stream (through pinned staging), then
 * boomgpu_synchronize: the read side of a step_device + all-reduce sequence */
int boomgpu_download(boomgpu_ctx *ctx, const double *src_dev, double *dst_host, int64_t count);

/* ---- multi-GPU: one context per GPU, rows sharded, ONE all-reduce of the packed statistics per iteration ---------
 * The reference combines its workers' statistics under a mutex (Models/PosteriorSamplers/Imputer.cpp:40-64 over
 * BinomialLogitAuxmixSampler.cpp:44-49 / WeightedRegressionModel.cpp:79-87); here a worker is a GPU and combine() is
 * ncclAllReduce(ncclDouble, ncclSum) over NVLink.  NCCL is loaded at run time (libnccl.so.2, whichever copy the process
 * already holds), so the library has no link-time dependency on it.
 *   boomgpu_comm_unique_id   rank 0 makes the 128-byte id and hands it to the other ranks by its own means
 *   boomgpu_comm_init        every rank, after boomgpu_create: joins the communicator (ncclCommInitRank)
 *   boomgpu_allreduce        sums count doubles at a device pointer over all ranks, in place, on the context's stream
 * With a communicator attached, the synchronous boomgpu_logit_step / boomgpu_poisson_step all-reduce before they copy
 * the statistics to the host; the *_step_device variants leave that to the caller (boomgpu_allreduce). */
#define BOOMGPU_COMM_ID_BYTES 128
int boomgpu_comm_unique_id(char id[BOOMGPU_COMM_ID_BYTES]);
int boomgpu_comm_init(boomgpu_ctx *ctx, const char id[BOOMGPU_COMM_ID_BYTES], int nranks, int rank);
int boomgpu_comm_destroy(boomgpu_ctx *ctx);
int boomgpu_allreduce(boomgpu_ctx *ctx, double *dev, int64_t count);

/* ---- parity / test hooks --------------------------------------------------------------- */
/* deterministic accumulation from caller supplied latents (host arrays of length n) */
int boomgpu_accumulate(boomgpu_ctx *ctx, const double *weight, const double *weighted_value,
                       double *xtx, double *xty);
/* per-row latent draws without accumulation: sum_out/info_out host arrays of length n */
int boomgpu_logit_draw(boomgpu_ctx *ctx, const double *beta, int clt_threshold, uint64_t seed,
                       uint64_t iteration, double *sum_out, double *info_out);
/* out6 (n x 6) = {z_int, mu_int, w_int, z_ext, mu_ext, w_ext}; kout2 (n x 2, may be NULL) component indices */
int boomgpu_poisson_draw(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration,
                         double *out6, int32_t *kout2);
/* per-row sums of the latent probit normals (host array of length n) */
int boomgpu_probit_draw(boomgpu_ctx *ctx, const double *beta, int clt_threshold, uint64_t seed, uint64_t iteration,
                        double *sum_z_out);
/* the latent weights of the Student-t sibling, one per row (host array of length n) */
int boomgpu_student_draw(boomgpu_ctx *ctx, const double *beta, double sigma, double nu, uint64_t seed, uint64_t iteration,
                         double *weight_out);
int boomgpu_binomial_loglike(boomgpu_ctx *ctx, const double *beta, double *loglike);
int boomgpu_poisson_loglike(boomgpu_ctx *ctx, const double *beta, double *loglike);
/* log likelihood with gradient (p, may be NULL) and Hessian (p x p, symmetric, may be NULL) in one pass over the rows:
 * BinomialLogitModel::log_likelihood(beta, g, h) (Models/Glm/BinomialLogitModel.cpp:140-180; eta = x'beta - log_alpha,
 * g = sum (y - n p) x, h = -sum n p q x x') and PoissonRegressionModel::log_likelihood(beta, g, h)
 * (Models/Glm/PoissonRegressionModel.cpp:56-95; g = sum (y - E lambda) x, h = -sum lambda x x', exposure-free as there).
 * beta is the FULL coefficient vector (zeros at excluded positions); callers select the included sub-blocks. */
int boomgpu_binomial_loglike_derivs(boomgpu_ctx *ctx, const double *beta, double log_alpha, double *loglike, double *gradient,
                                    double *hessian);
int boomgpu_poisson_loglike_derivs(boomgpu_ctx *ctx, const double *beta, double *loglike, double *gradient, double *hessian);
/* With a communicator attached (boomgpu_comm_init) the four log-likelihood entry points above all-reduce over the ranks
 * before they answer: every rank gets the likelihood of ALL rows.  The *_device variants leave this rank's packed result
 * [ -(Hessian) p*p | gradient p | {., log likelihood, ., .} ] at suf_dev for a caller-driven all-reduce. */
int boomgpu_binomial_loglike_derivs_device(boomgpu_ctx *ctx, const double *beta, double log_alpha, double *suf_dev);
int boomgpu_poisson_loglike_derivs_device(boomgpu_ctx *ctx, const double *beta, double *suf_dev);

/* The same restricted to a set of columns -- the included variables of a spike-and-slab model.  The reference's chunk log
 * posterior (BinomialLogitLogPostChunk, Models/Glm/PosteriorSamplers/BinomialLogitCompositeSpikeSlabSampler.cpp:34-74) selects
 * the included columns from every observation on every evaluation; here boomgpu_select_columns gathers X_gamma (n x k) once
 * per model (cached until another selection or new data), and every evaluation is one pass over those k columns:
 * beta_selected, gradient: k; hessian: k x k.  With a communicator attached the results are all-reduced over the ranks;
 * the _device variant leaves [ -H k*k | g k | {., log likelihood, ., .} ] (boomgpu_suf_len(k) doubles) at suf_dev. */
int boomgpu_select_columns(boomgpu_ctx *ctx, const int32_t *columns, int k);
int boomgpu_binomial_loglike_derivs_selected(boomgpu_ctx *ctx, const double *beta_selected, double log_alpha, double *loglike,
                                             double *gradient, double *hessian);
int boomgpu_poisson_loglike_derivs_selected(boomgpu_ctx *ctx, const double *beta_selected, double *loglike, double *gradient,
                                            double *hessian);
int boomgpu_binomial_loglike_derivs_selected_device(boomgpu_ctx *ctx, const double *beta_selected, double log_alpha, double *suf_dev);

/* ---- instrumentation ------------------------------------------------------------------- */
/* number of kernels this context has launched so far */
int64_t boomgpu_kernel_launches(const boomgpu_ctx *ctx);
/* with option "timing" = 1: accumulated CUDA-event milliseconds and launch counts per kernel class
 * since the last reset; classes: 0 fused small-p, 1 imputer pass, 2 DMMA SYRK, 3 reductions, 4 other */
#define BOOMGPU_NUM_KERNEL_CLASSES 5
int boomgpu_get_timings(boomgpu_ctx *ctx, double ms[BOOMGPU_NUM_KERNEL_CLASSES],
                        int64_t launches[BOOMGPU_NUM_KERNEL_CLASSES], int reset);

#pragma GCC visibility pop
#ifdef __cplusplus
}
#endif
#endif
