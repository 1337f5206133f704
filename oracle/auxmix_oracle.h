/*
 * auxmix_oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C CPU restatement of the reference's (steve-the-bayesian/BOOM)
 * auxiliary-mixture data-augmentation hot path.  Every function cites the
 * reference file:line it follows (paths relative to the reference root).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * call this.  The product path (boom_b200/csrc) never does.
 *
 * Parity status: PINNED.  The deterministic parts (unmix, truncated-logistic
 * inverse CDF, truncated normal moments, sufficient-statistic updates, the
 * log likelihood) are checked against golden vectors produced by the compiled
 * reference itself (oracle/ref_driver.cpp -> tests/golden/), see
 * tests/test_oracle_golden.py.
 *
 * Random numbers: the reference draws from a per-worker std::mt19937_64, so
 * bit identity of draws is impossible by construction.
 * The oracle and the CUDA path share ONE counter-based generator instead
 * (Philox4x32-10 keyed by seed / iteration / global row / slot), so the
 * device draws can be compared with the oracle's value by value, and the
 * oracle's draws are compared with the reference's *distributions*.
 */
#ifndef BOOM_B200_AUXMIX_ORACLE_H
#define BOOM_B200_AUXMIX_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- counter based RNG shared with the device path ------------------------------- */
void bo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* two uniforms in the open interval (0,1) for (seed, iteration, global row, slot) */
void bo_uniform_pair(uint64_t seed, uint64_t iteration, uint64_t row, uint32_t slot, double u[2]);

/* ---- finite normal mixtures ------------------------------------------------------- */
typedef struct {
  int K;
  const double *mu;          /* K */
  const double *sigma;       /* K */
  const double *weights;     /* K */
  const double *log_weights; /* K  (= log(weights), computed by bo_log_array: glibc log like the reference) */
} bo_mixture;

/* elementwise natural log with the C library (so fixtures need not trust numpy's SIMD log) */
void bo_log_array(int n, const double *x, double *out);

/* NormalMixtureApproximation::unmix given the uniform (returns component index;
 * post, if not NULL, receives the K normalised posterior probabilities). */
int bo_unmix(const bo_mixture *mix, double residual, double unif, double *post);

/* rtrun_logit_mt given its uniform. */
double bo_rtrun_logit(double eta, int success, double unif);

/* trun_norm_moments. */
void bo_trun_norm_moments(double mu, double sigma, double cutpoint, int positive_support,
                          double *mean, double *variance);

/* exact binomial draw from ONE uniform by chop-down search from the mode
 * (replaces Bmath rbinom inside rmultinom; same distribution). */
int64_t bo_binomial_from_uniform(int64_t n, double p, double unif);

/* ---- per-observation imputers ----------------------------------------------------- */
/* BinomialLogitCltDataImputer::impute: returns 0 ok, nonzero = invalid input.
 * kcount (optional, K ints) accumulates the mixture-indicator histogram. */
int bo_logit_impute(const bo_mixture *mix, int clt_threshold, double ntrials, double y, double eta,
                    uint64_t seed, uint64_t iteration, uint64_t row,
                    double *sum, double *info, int64_t *kcount);

/* Poisson mixture table: ntab entries sorted by nu, CSR offsets into flat arrays. */
typedef struct {
  int ntab;
  const int64_t *nu;      /* ntab   */
  const int32_t *offset;  /* ntab+1 */
  const double *mu, *sigma, *log_weights;
  int64_t gaussian_cutoff; /* table->largest_index() */
} bo_poisson_table;

/* PoissonDataImputer::impute.  out6 = {z_int, mu_int, w_int, z_ext, mu_ext, w_ext}
 * (internal triple untouched when y == 0); kout2 = component indices (-1 = none).
 * returns 0 ok, 1 = nu missing from table, 2 = invalid input. */
int bo_poisson_impute(const bo_poisson_table *tab, int64_t y, double exposure, double eta,
                      uint64_t seed, uint64_t iteration, uint64_t row, double out6[6], int kout2[2]);

/* ---- sufficient statistics -------------------------------------------------------- */
/* BinomialLogit::SufficientStatistics::update -- xtx is p x p column major, upper triangle only */
void bo_logit_suf_update(int p, double *xtx, double *xty, const double *x, double weighted_value, double weight);
/* WeightedRegSuf::add_data -- scalars = {n, yWy, sumw, sumlogw} */
void bo_weighted_reg_suf_add(int p, double *xtwx, double *xtwy, double scalars[4], const double *x, double y, double w);
/* SpdMatrix::reflect (upper -> lower) */
void bo_reflect(int p, double *a);

/* ---- whole steps (the loops of Imputer.hpp:175-180) -------------------------------- */
/* accumulate from caller supplied latents (deterministic parity) */
void bo_accumulate(int64_t n, int p, const double *X, int64_t ldx, const double *weight,
                   const double *weighted_value, double *xtx, double *xty);

int bo_logit_step(int64_t n, int p, const double *X, int64_t ldx, const double *y, const double *ntrials,
                  const double *beta, int clt_threshold, const bo_mixture *mix, uint64_t seed,
                  uint64_t iteration, uint64_t row_offset, double *xtx, double *xty, int64_t *sample_size,
                  int64_t *kcount);

int bo_poisson_step(int64_t n, int p, const double *X, int64_t ldx, const int64_t *y, const double *exposure,
                    const double *beta, const bo_poisson_table *tab, uint64_t seed, uint64_t iteration,
                    uint64_t row_offset, double *xtwx, double *xtwy, double scalars[4]);

/* per-row latent draws without accumulation (KS / chi-square tests) */
int bo_logit_draw(int64_t n, int p, const double *X, int64_t ldx, const double *y, const double *ntrials,
                  const double *beta, int clt_threshold, const bo_mixture *mix, uint64_t seed,
                  uint64_t iteration, uint64_t row_offset, double *sum_out, double *info_out);
int bo_poisson_draw(int64_t n, int p, const double *X, int64_t ldx, const int64_t *y, const double *exposure,
                    const double *beta, const bo_poisson_table *tab, uint64_t seed, uint64_t iteration,
                    uint64_t row_offset, double *out6, int32_t *kout2);

/* ---- probit sibling (BinomialProbitSpikeSlabSampler) ---------------------------------- */
double bo_rtrun_norm_unit(double eta, int positive, double unif);
int bo_probit_impute(int clt_threshold, double ntrials, double nsuccess, double eta, uint64_t seed, uint64_t iteration,
                     uint64_t row, double *sum_z);
int bo_probit_step(int64_t n, int p, const double *X, int64_t ldx, const double *y, const double *ntrials, const double *beta,
                   int clt_threshold, uint64_t seed, uint64_t iteration, uint64_t row_offset, double *xtx, double *xtz,
                   double *draws);

/* ---- Student-t sibling (TRegressionSampler) ------------------------------------------- */
int bo_rgamma(double shape, double rate, uint64_t seed, uint64_t iteration, uint64_t row, double *out);
int bo_student_impute(double residual, double sigma, double nu, uint64_t seed, uint64_t iteration, uint64_t row, double *weight);
int bo_student_step(int64_t n, int p, const double *X, int64_t ldx, const double *y, const double *beta, double sigma, double nu,
                    uint64_t seed, uint64_t iteration, uint64_t row_offset, double *xtwx, double *xtwy, double scalars[4],
                    double *weights);
double bo_student_loglike(int64_t n, int p, const double *X, int64_t ldx, const double *y, const double *beta, double sigma,
                          double nu);
void bo_synth_student_y(int64_t n, int p, const double *X, int64_t ldx, const double *beta, double sigma, double nu,
                        uint64_t seed, uint64_t row_offset, double *y);

/* BinomialLogitModel::log_likelihood (value only, log_alpha = 0) */
double bo_dbinom_log(double x, double n, double p);
double bo_binomial_logit_loglike(int64_t n, int p, const double *X, int64_t ldx, const double *y,
                                 const double *ntrials, const double *beta);
/* PoissonRegressionModel::log_likelihood (value only) */
double bo_poisson_loglike(int64_t n, int p, const double *X, int64_t ldx, const int64_t *y,
                          const double *exposure, const double *beta);

/* the same with gradient g (p) and Hessian h (p x p): BinomialLogitModel.cpp:140-180, PoissonRegressionModel.cpp:56-95 */
double bo_binomial_logit_loglike_derivs(int64_t n, int p, const double *X, int64_t ldx, const double *y,
                                        const double *ntrials, const double *beta, double log_alpha, double *g, double *h);
double bo_poisson_loglike_derivs(int64_t n, int p, const double *X, int64_t ldx, const int64_t *y, const double *exposure,
                                 const double *beta, double *g, double *h);

/* ---- synthetic data shared by every arm -------------------------- */
void bo_synth_beta(int p, int nonzero, double intercept, double *beta);
void bo_synth_x(int64_t n, int p, uint64_t seed, double xscale, uint64_t row_offset, double *X, int64_t ldx);
void bo_synth_binomial_y(int64_t n, int p, const double *X, int64_t ldx, const double *beta, uint64_t seed,
                         int max_trials, uint64_t row_offset, double *y, double *ntrials);
void bo_synth_poisson_y(int64_t n, int p, const double *X, int64_t ldx, const double *beta, uint64_t seed,
                        uint64_t row_offset, int64_t *y, double *exposure);

#ifdef __cplusplus
}
#endif
#endif
