// boom_b200_adapter.hpp -- the drop-in for an existing BOOM program.
//
// Samplers derived from BOOM::PosteriorSampler (Models/PosteriorSamplers/PosteriorSampler.hpp:44-107)
// that are constructed on the reference's OWN model objects and attach with model->set_method(sampler):
//
//   BOOM::BinomialLogitAuxmixSampler        -> BOOM::B200::BinomialLogitAuxmixSampler
//   BOOM::BinomialLogitSpikeSlabSampler     -> BOOM::B200::BinomialLogitSpikeSlabSampler
//   BOOM::BinomialLogitCompositeSpikeSlabSampler -> BOOM::B200::BinomialLogitCompositeSpikeSlabSampler  (what R's logit.spike builds)
//   BOOM::BinomialProbitSpikeSlabSampler    -> BOOM::B200::BinomialProbitSpikeSlabSampler   (sibling)
//   BOOM::TRegressionSampler                -> BOOM::B200::TRegressionSampler               (sibling)
//   BOOM::TRegressionSpikeSlabSampler       -> BOOM::B200::TRegressionSpikeSlabSampler      (what lm.spike builds for Student errors)
//   BOOM::PoissonRegressionAuxMixSampler    -> BOOM::B200::PoissonRegressionAuxMixSampler
//   BOOM::PoissonRegressionSpikeSlabSampler -> BOOM::B200::PoissonRegressionSpikeSlabSampler
//
// Same constructor arguments and public methods as the originals
// (Models/Glm/PosteriorSamplers/BinomialLogitAuxmixSampler.hpp:109-152, BinomialLogitSpikeSlabSampler.hpp:27-96,
// PoissonRegressionAuxMixSampler.hpp:60-125, PoissonRegressionSpikeSlabSampler.hpp), the public accessors included:
// suf() returns the reference's own BinomialLogit::SufficientStatistics, complete_data_sufficient_statistics() its
// WeightedRegSuf.  They cannot subclass the originals (the worker-pool base LatentDataSampler<...> would come along),
// so they derive from PosteriorSampler directly.
//
// Where the statistics live.  The device step lands X'WX / X'Wz in ONE host object, a BOOM_B200::WeightedRegSuf whose
// p x p storage is page-locked once for large p (the device->host copy goes straight into it), and the host steps --
// beta | statistics, the inclusion sweep -- read that object in place: no p x p matrix is copied per draw (at p = 4000
// each copy is 128 MB).  The reference-typed objects behind suf() / complete_data_sufficient_statistics() are filled
// from it on demand, when somebody asks.
//
// impute_latent_data() is one device step through the C ABI (include/boomgpu.h) on rows packed from model->dat() a chunk
// at a time (no second n x p host copy), re-packed when the model's data change (IID_DataPolicy::add_observer,
// IID_DataPolicy.hpp:43-45); in-place edits of a row (set_x / set_y / set_n) are seen after observe_rows(true), or
// with reassign_data_each_time(true), or after refresh_data().  The small-state steps run on the host: beta by BOOM's own
// rmvn_suf_mt, the inclusion sweep by the shared bordered-Cholesky evaluator of boom_b200/host.  All randomness derives
// from the sampler's BOOM::RNG, so set_seed() repeats a chain.  Mixture tables are read from the live reference objects
// (BinomialLogitDataImputer::mixture_approximation, create_poisson_mixture_approximation_table()).
//
// Needs the BOOM headers and library: built only where the reference exists.
#pragma once

#include <functional>
#include <memory>

#include "Models/Glm/BinomialLogitModel.hpp"
#include "Models/Glm/BinomialProbitModel.hpp"
#include "Models/Glm/PoissonRegressionModel.hpp"
#include "Models/Glm/PosteriorSamplers/BinomialLogitAuxmixSampler.hpp"
#include "Models/Glm/TRegression.hpp"
#include "Models/Glm/VariableSelectionPrior.hpp"
#include "Models/GammaModel.hpp"
#include "Models/PosteriorSamplers/GenericGaussianVarianceSampler.hpp"
#include "Models/ScaledChisqModel.hpp"
#include "Samplers/ScalarSliceSampler.hpp"
#include "Models/Glm/WeightedRegressionModel.hpp"
#include "Models/MvnBase.hpp"
#include "Models/PosteriorSamplers/PosteriorSampler.hpp"
#include "Samplers/MoveAccounting.hpp"
#include "cpputil/math_utils.hpp"

#include "../host/boom_b200.hpp"

namespace BOOM {
namespace B200 {

// Shared machinery: packed rows on the device, the step, the statistics, the host steps.
class DeviceImputerBase : public PosteriorSampler {
 public:
  ~DeviceImputerBase() override;
  // LatentDataSampler surface (Models/PosteriorSamplers/Imputer.hpp:260-314)
  void impute_latent_data();
  void fix_latent_data(bool fixed = true) { latent_data_fixed_ = fixed; }
  void set_number_of_workers(int) {}
  void reassign_data_each_time(bool tf) { repack_each_time_ = tf; }
  void clear_complete_data_sufficient_statistics() { hsuf_.clear(); statistics_changed(); }
  // the rows are re-packed before the next draw (use after editing observations in place)
  void refresh_data() { stale_ = true; }
  // true: every observation (and its x / y parts) gets an observer, so in-place edits (set_x, set_y, set_n) re-pack the
  // rows before the next draw -- three map nodes per observation, meant for data sets that fit that comfortably
  void observe_rows(bool tf);
  // multi-GPU / placement
  void set_device(int device);
  void set_row_offset(uint64_t first_global_row) { row_offset_ = first_global_row; stale_ = true; }
  void set_allreduce(const BOOM_B200::AllReduceFn &fn) { allreduce_ = fn; }
  // or natively: join an NCCL communicator (id from BOOM_B200::GlmModelBase::comm_unique_id() on rank 0)
  void set_communicator(const std::string &id, int nranks, int rank) { comm_id_ = id; comm_ranks_ = nranks; comm_rank_ = rank; comm_dirty_ = true; }
  // Active-set statistics (off by default; spike-and-slab samplers, p > 64, no all-reduce hook): the device computes
  // X'WX for the columns of the included variables only (+ diagonal, X'Wz), the sweep fetches a column when it adds a variable;
  // the same chain as with the full matrix.  The reference-typed statistics accessors still answer with the full matrix,
  // computed on demand from the latents that stay in HBM.
  void set_active_set_statistics(bool tf) { active_.enabled = tf; }
  int64_t active_set_columns_fetched() const { return active_.columns_fetched; }
  // wall-clock seconds this sampler has spent in: the device step (incl. copies), the host small-state steps
  double seconds_in_device_step() const { return secs_device_; }
  double seconds_in_host_steps() const { return secs_host_; }

 protected:
  DeviceImputerBase(int xdim, RNG &seeding_rng);
  // walks model->dat(): n rows, handed to the device a chunk at a time through emit(row0, nrows, X, y, aux)
  typedef std::function<void(int64_t row0, int64_t nrows, const double *X, const void *y, const double *aux)> ChunkSink;
  virtual int64_t row_count() const = 0;
  virtual bool rows_are_poisson() const = 0;
  virtual int row_kind() const { return rows_are_poisson() ? 1 : 0; }   // boomgpu_upload_begin: 0 binomial, 1 Poisson, 2 plain regression
  virtual void install_tables(boomgpu_ctx *ctx) = 0;                     // mixture tables, before the rows
  virtual void pack_rows(int64_t row0, int64_t nrows, double *X, void *y, double *aux) const = 0;
  virtual void observe_row_objects(bool tf) = 0;
  virtual int device_step(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, double *suf_dev) = 0;
  // the synchronous form: all-reduces natively when a communicator is attached, lands the statistics at xtx / xty / scalars
  virtual int device_step_sync(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, double *xtx, double *xty,
                               double scalars[4]) = 0;
  virtual const Vector &current_beta() const = 0;
  virtual int device_step_active(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, const int32_t *cols, int k,
                                 double *G, double *diag, double *xty, double scalars[4]);   // default: not provided
  bool impute_latent_data_active(const Selector &inc);   // false: does not apply here, run the full step
  void materialize_full_statistics() const;              // after an active-set step: the full matrix from the latents in HBM
  // log likelihood with gradient / Hessian at a full coefficient vector, one device pass (boomgpu_*_loglike_derivs)
  virtual int device_loglike_derivs(boomgpu_ctx *ctx, const double *beta, double *loglike, double *g, double *h) = 0;
  virtual int device_loglike_derivs_device(boomgpu_ctx *ctx, const double *beta, double *suf_dev) = 0;
  double loglike_derivs(const BOOM_B200::Vector &beta, BOOM_B200::Vector *g, BOOM_B200::SpdMatrix *h);   // all ranks' rows
  // the same over the INCLUDED columns only (k = inc.nvars(); beta, g: k, h: k x k): X_gamma is gathered on the device once
  // per inclusion pattern, an evaluation is one pass over it (boomgpu_select_columns / boomgpu_*_loglike_derivs_selected)
  double included_loglike_derivs(const Selector &inc, const Vector &beta_included, Vector *g, Matrix *h);
  virtual int device_loglike_derivs_selected(boomgpu_ctx *ctx, const double *beta, double *loglike, double *g, double *h) = 0;
  virtual int device_loglike_derivs_selected_device(boomgpu_ctx *ctx, const double *beta, double *suf_dev) = 0;
  void ensure_device_rows();
  // find_posterior_mode of the spike-and-slab samplers (BinomialLogitSpikeSlabSampler.cpp:147-177,
  // PoissonRegressionSpikeSlabSampler.cpp:69-106): Newton-Raphson on the included coefficients, derivatives from the device
  bool find_mode(GlmCoefs &coef, const Ptr<MvnBase> &slab, const Ptr<VariableSelectionPrior> &spike, double epsilon, double *value);
  void mark_stale() { stale_ = true; }
  boomgpu_ctx *device_ctx() { return ctx_; }   // valid after ensure_device_rows()
  uint64_t device_iteration() const { return iteration_; }
  void check(int rc) const;
  virtual void statistics_changed() {}   // the reference-typed views are out of date
  // host steps on hsuf_
  void draw_beta_full_model(GlmCoefs &coef, const MvnBase &prior);
  // the shared spike-and-slab core over the CURRENT slab / spike (host copies cached; see prior_cache())
  const BOOM_B200::SpikeSlabCore &core(const Ptr<MvnBase> &slab, const Ptr<VariableSelectionPrior> &spike, bool fisher_yates) const;
  void sweep_indicators(GlmCoefs &coef, const BOOM_B200::SpikeSlabCore &core);
  void draw_included_beta(GlmCoefs &coef, const BOOM_B200::SpikeSlabCore &core);
  double model_log_prob(const Selector &g, const BOOM_B200::SpikeSlabCore &core) const;
  double spike_slab_logpri(const GlmCoefs &coef, const MvnBase &slab, const VariableSelectionPrior &spike) const;
  void priors_changed() { ++prior_version_; }
  void *observer_key() { return static_cast<void *>(this); }

  mutable BOOM_B200::WeightedRegSuf hsuf_;   // where the device step lands and the host steps read
  mutable BOOM_B200::ActiveSetState active_;
  std::unique_ptr<BOOM_B200::StatView> view_;   // the active-set view of the current iteration (sweep, then beta)
  int xdim_;

 private:
  struct PriorCache;
  boomgpu_ctx *ctx_ = nullptr;
  int device_ = 0;
  bool stale_ = true, latent_data_fixed_ = false, repack_each_time_ = false, observing_rows_ = false;
  uint64_t row_offset_ = 0, iteration_ = 0;
  BOOM_B200::AllReduceFn allreduce_;
  std::string comm_id_;
  int comm_ranks_ = 1, comm_rank_ = 0;
  bool comm_dirty_ = false;
  std::vector<double> packed_;
  mutable std::unique_ptr<PriorCache> prior_cache_;
  uint64_t prior_version_ = 1;
  double secs_device_ = 0, secs_host_ = 0;
};

class BinomialLogitAuxmixSampler : public DeviceImputerBase {
 public:
  BinomialLogitAuxmixSampler(BinomialLogitModel *model, const Ptr<MvnBase> &prior, int clt_threshold = 10,
                             RNG &seeding_rng = GlobalRng::rng);
  void draw() override;
  double logpri() const override;
  void draw_params();
  // the reference's own statistics type (BinomialLogitAuxmixSampler.hpp:39-67), filled from the device result on demand
  const BinomialLogit::SufficientStatistics &suf() const;
  int clt_threshold() const { return clt_threshold_; }
  void update_complete_data_sufficient_statistics(double precision_weighted_sum, double total_precision, const Vector &x);

 protected:
  int64_t row_count() const override { return (int64_t)model_->dat().size(); }
  bool rows_are_poisson() const override { return false; }
  void install_tables(boomgpu_ctx *ctx) override;
  void pack_rows(int64_t row0, int64_t nrows, double *X, void *y, double *aux) const override;
  void observe_row_objects(bool tf) override;
  int device_step(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, double *suf_dev) override;
  int device_step_sync(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, double *xtx, double *xty,
                       double scalars[4]) override;
  int device_loglike_derivs(boomgpu_ctx *ctx, const double *beta, double *loglike, double *g, double *h) override;
  int device_loglike_derivs_device(boomgpu_ctx *ctx, const double *beta, double *suf_dev) override;
  int device_loglike_derivs_selected(boomgpu_ctx *ctx, const double *beta, double *loglike, double *g, double *h) override;
  int device_loglike_derivs_selected_device(boomgpu_ctx *ctx, const double *beta, double *suf_dev) override;
  int device_step_active(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, const int32_t *cols, int k,
                         double *G, double *diag, double *xty, double scalars[4]) override;
  const Vector &current_beta() const override { return model_->Beta(); }
  void statistics_changed() override { suf_synced_ = false; }
  BinomialLogitModel *model_;
  Ptr<MvnBase> prior_;

 private:
  int clt_threshold_;
  mutable BinomialLogit::SufficientStatistics suf_;
  mutable bool suf_synced_ = false;
};

class BinomialLogitSpikeSlabSampler : public BinomialLogitAuxmixSampler {
 public:
  BinomialLogitSpikeSlabSampler(BinomialLogitModel *model, const Ptr<MvnBase> &slab, const Ptr<VariableSelectionPrior> &spike,
                                int clt_threshold, RNG &seeding_rng = GlobalRng::rng);
  BinomialLogitSpikeSlabSampler *clone_to_new_host(Model *model) const override;
  void draw() override;
  double logpri() const override;
  void draw_model_indicators();                         // BinomialLogitSpikeSlabSampler.cpp:180-211
  virtual void draw_beta();                             // .cpp:56-75
  double log_model_prob(const Selector &gamma) const;   // .cpp:88-117
  void allow_model_selection(bool tf) { allow_model_selection_ = tf; }
  void limit_model_selection(int max_flips) { max_flips_ = max_flips; }
  void set_spike(const Ptr<VariableSelectionPrior> &spike);
  void set_slab(const Ptr<MvnBase> &slab);
  int xdim() const { return model_->xdim(); }
  void find_posterior_mode(double epsilon = 1e-5) override;
  bool can_find_posterior_mode() const override { return true; }
  bool posterior_mode_found() const { return posterior_mode_found_; }
  double log_posterior_at_mode() const { return log_posterior_at_mode_; }

 private:
  Ptr<MvnBase> slab_;
  Ptr<VariableSelectionPrior> spike_;
  bool allow_model_selection_ = true;
  int max_flips_ = -1;
  bool posterior_mode_found_ = false;
  double log_posterior_at_mode_ = negative_infinity();
};

// BinomialLogitLogPostChunk (BinomialLogitCompositeSpikeSlabSampler.hpp:27-47, .cpp:26-74): log posterior of a chunk of the
// included coefficients given the rest, with gradient and Hessian with respect to the chunk.  The reference walks all n
// observations on the host per evaluation; here an evaluation is one device pass over the included columns.
class BinomialLogitCompositeSpikeSlabSampler;
class BinomialLogitLogPostChunk {
 public:
  BinomialLogitLogPostChunk(BinomialLogitCompositeSpikeSlabSampler *sampler, int chunk_size, int chunk_number);
  double operator()(const Vector &beta_chunk) const;
  double operator()(const Vector &beta_chunk, Vector &grad, Matrix &hess, int nd) const;

 private:
  BinomialLogitCompositeSpikeSlabSampler *sampler_;
  int start_, chunk_size_;
};

// BinomialLogitCompositeSpikeSlabSampler (.hpp:49-103, .cpp:76-265): every draw() is, with the given weights, a
// data-augmentation move (the auxiliary-mixture Gibbs step above), a random-walk Metropolis sweep over chunks of the included
// coefficients, or a tailored-independence-Metropolis sweep (BOOM's own TIM, Samplers/TIM.cpp, driven by the device-evaluated
// chunk log posterior).
class BinomialLogitCompositeSpikeSlabSampler : public BinomialLogitSpikeSlabSampler {
 public:
  BinomialLogitCompositeSpikeSlabSampler(BinomialLogitModel *model, const Ptr<MvnBase> &prior, const Ptr<VariableSelectionPrior> &vpri,
                                         int clt_threshold, double tdf, int max_tim_chunk_size, int max_rwm_chunk_size = 1,
                                         double rwm_variance_scale_factor = 1.0, RNG &seeding_rng = GlobalRng::rng);
  void draw() override;
  void rwm_draw();
  void tim_draw();
  void rwm_draw_chunk(int chunk);
  BinomialLogitLogPostChunk log_posterior(int chunk_number, int max_chunk_size) const;
  void set_sampler_weights(double da_weight, double rwm_weight, double tim_weight);
  std::ostream &time_report(std::ostream &out) const;

 private:
  friend class BinomialLogitLogPostChunk;
  double chunk_loglike(const Vector &beta_included, Vector *g, Matrix *h) { return included_loglike_derivs(m_->coef().inc(), beta_included, g, h); }
  BinomialLogitModel *m_;
  Ptr<MvnBase> pri_;
  double tdf_;
  int max_tim_chunk_size_, max_rwm_chunk_size_;
  double rwm_variance_scale_factor_;
  MoveAccounting move_accounting_;
  Vector sampler_weights_;
  int compute_chunk_size(int max_chunk_size) const;
  int compute_number_of_chunks(int max_chunk_size) const;
};

class PoissonRegressionAuxMixSampler : public DeviceImputerBase {
 public:
  PoissonRegressionAuxMixSampler(PoissonRegressionModel *model, const Ptr<MvnBase> &prior, int number_of_threads = 1,
                                 RNG &seeding_rng = GlobalRng::rng);
  void draw() override;
  double logpri() const override;
  void draw_beta_given_complete_data();
  const WeightedRegSuf &complete_data_sufficient_statistics() const;   // BOOM's own type, filled on demand
  void update_complete_data_sufficient_statistics(double precision_weighted_sum, double total_precision, const Vector &x);

 protected:
  int64_t row_count() const override { return (int64_t)model_->dat().size(); }
  bool rows_are_poisson() const override { return true; }
  void install_tables(boomgpu_ctx *ctx) override;
  void pack_rows(int64_t row0, int64_t nrows, double *X, void *y, double *aux) const override;
  void observe_row_objects(bool tf) override;
  int device_step_active(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, const int32_t *cols, int k,
                         double *G, double *diag, double *xty, double scalars[4]) override;
  int device_step(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, double *suf_dev) override;
  int device_step_sync(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, double *xtx, double *xty,
                       double scalars[4]) override;
  int device_loglike_derivs(boomgpu_ctx *ctx, const double *beta, double *loglike, double *g, double *h) override;
  int device_loglike_derivs_device(boomgpu_ctx *ctx, const double *beta, double *suf_dev) override;
  int device_loglike_derivs_selected(boomgpu_ctx *ctx, const double *beta, double *loglike, double *g, double *h) override;
  int device_loglike_derivs_selected_device(boomgpu_ctx *ctx, const double *beta, double *suf_dev) override;
  const Vector &current_beta() const override { return model_->Beta(); }
  void statistics_changed() override { suf_synced_ = false; }
  PoissonRegressionModel *model_;
  Ptr<MvnBase> prior_;

 private:
  mutable WeightedRegSuf suf_;
  mutable bool suf_synced_ = false;
};

// BinomialProbitSpikeSlabSampler (Models/Glm/PosteriorSamplers/BinomialProbitSpikeSlabSampler.hpp:34-68).  X'NX is computed
// when the rows are (re)packed, every later iteration computes X'z alone.
class BinomialProbitSpikeSlabSampler : public DeviceImputerBase {
 public:
  BinomialProbitSpikeSlabSampler(BinomialProbitModel *model, const Ptr<MvnBase> &slab_prior, const Ptr<VariableSelectionPrior> &spike_prior,
                                 int clt_threshold = 10, RNG &seeding_rng = GlobalRng::rng);
  void draw() override;
  double logpri() const override;
  void allow_model_selection(bool tf) { allow_model_selection_ = tf; }
  void limit_model_selection(int max_flips) { max_flips_ = max_flips; }
  void impute_latent_data();
  void refresh_xtx() { want_xtx_ = true; }
  WeightedRegSuf complete_data_sufficient_statistics() const;   // by value, as in the reference

 protected:
  int64_t row_count() const override { return (int64_t)model_->dat().size(); }
  bool rows_are_poisson() const override { return false; }
  void install_tables(boomgpu_ctx *) override { want_xtx_ = true; }   // called exactly when the rows are (re)packed
  void pack_rows(int64_t row0, int64_t nrows, double *X, void *y, double *aux) const override;
  void observe_row_objects(bool tf) override;
  int device_step(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, double *suf_dev) override;
  int device_step_sync(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, double *xtx, double *xty,
                       double scalars[4]) override;
  int device_loglike_derivs(boomgpu_ctx *, const double *, double *, double *, double *) override;
  int device_loglike_derivs_device(boomgpu_ctx *, const double *, double *) override;
  int device_loglike_derivs_selected(boomgpu_ctx *, const double *, double *, double *, double *) override;
  int device_loglike_derivs_selected_device(boomgpu_ctx *, const double *, double *) override;
  const Vector &current_beta() const override { return model_->Beta(); }

 private:
  BinomialProbitModel *model_;
  Ptr<MvnBase> slab_;
  Ptr<VariableSelectionPrior> spike_;
  int clt_threshold_;
  bool allow_model_selection_ = true, want_xtx_ = true;
  int max_flips_ = -1;
};

// TRegressionSampler (Models/Glm/PosteriorSamplers/TRegressionSampler.hpp:34-137): Student-t regression by the scale-mixture
// augmentation.  The weights w_i | residual and WeightedRegSuf::add_data(x_i, y_i, w_i) are one device step; beta, sigma^2
// and nu are drawn on the host with BOOM's own rmvn_suf_mt, GenericGaussianVarianceSampler and ScalarSliceSampler -- the
// slice sampler's target (prior + observed-data Student log likelihood, TRegressionSampler.cpp:31-49) is evaluated on the
// device from residuals it keeps in HBM: one pass over X per draw of nu, then 8 n bytes per candidate.
class TRegressionSampler : public DeviceImputerBase {
 public:
  TRegressionSampler(TRegressionModel *model, const Ptr<MvnBase> &coefficient_prior, const Ptr<GammaModelBase> &siginv_prior,
                     const Ptr<DoubleModel> &nu_prior, RNG &seeding_rng = GlobalRng::rng);
  void draw() override;
  double logpri() const override;
  void draw_beta_full_conditional();
  void draw_sigsq_full_conditional();
  void draw_nu_given_complete_data();
  void draw_nu_given_observed_data();
  void set_sigma_upper_limit(double max_sigma) { sigsq_sampler_.set_sigma_max(max_sigma); }
  const WeightedRegSuf &complete_data_sufficient_statistics() const;   // the reference's own type, filled on demand
  void clear_complete_data_sufficient_statistics();
  void update_complete_data_sufficient_statistics(double y, const Vector &x, double weight);
  // TRegressionModel::log_likelihood(beta, sigsq, nu) (TRegression.cpp:74-86) evaluated on the device, all shards' rows
  double log_likelihood(const Vector &beta, double sigsq, double nu);
  int64_t likelihood_evaluations() const { return ll_evals_; }

 protected:
  int64_t row_count() const override { return (int64_t)model_->dat().size(); }
  bool rows_are_poisson() const override { return false; }
  int row_kind() const override { return 2; }
  void install_tables(boomgpu_ctx *) override {}
  void pack_rows(int64_t row0, int64_t nrows, double *X, void *y, double *aux) const override;
  void observe_row_objects(bool tf) override;
  int device_step(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, double *suf_dev) override;
  int device_step_sync(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, double *xtx, double *xty,
                       double scalars[4]) override;
  int device_step_active(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, const int32_t *cols, int k,
                         double *G, double *diag, double *xty, double scalars[4]) override;
  int device_loglike_derivs(boomgpu_ctx *, const double *, double *, double *, double *) override;
  int device_loglike_derivs_device(boomgpu_ctx *, const double *, double *) override;
  int device_loglike_derivs_selected(boomgpu_ctx *, const double *, double *, double *, double *) override;
  int device_loglike_derivs_selected_device(boomgpu_ctx *, const double *, double *) override;
  const Vector &current_beta() const override { return model_->Beta(); }
  void statistics_changed() override { suf_synced_ = false; }
  void coefficients_changed() { residuals_current_ = false; }
  virtual double weighted_sum_of_squared_errors();   // from hsuf_; the spike-and-slab sampler answers from its active-set view
  TRegressionModel *model_;
  Ptr<MvnBase> coefficient_prior_;
  Ptr<GammaModelBase> siginv_prior_;
  Ptr<DoubleModel> nu_prior_;

 private:
  double nu_log_posterior(double nu);
  Ptr<ScaledChisqModel> weight_model_;
  GenericGaussianVarianceSampler sigsq_sampler_;
  ScalarSliceSampler nu_observed_data_sampler_, nu_complete_data_sampler_;
  mutable WeightedRegSuf suf_;
  mutable bool suf_synced_ = false;
  bool residuals_current_ = false;
  int64_t ll_evals_ = 0;
};

// TRegressionSpikeSlabSampler (Models/Glm/PosteriorSamplers/TRegressionSpikeSlabSampler.hpp:30-72): the Student-t sampler with the
// coefficient draw replaced by the generic spike-and-slab steps at the model's residual variance.
class TRegressionSpikeSlabSampler : public TRegressionSampler {
 public:
  TRegressionSpikeSlabSampler(TRegressionModel *model, const Ptr<MvnBase> &coefficient_slab_prior,
                              const Ptr<VariableSelectionPrior> &coefficient_spike_prior, const Ptr<GammaModelBase> &siginv_prior,
                              const Ptr<DoubleModel> &nu_prior, RNG &seeding_rng = GlobalRng::rng);
  void draw() override;
  double logpri() const override;
  void draw_model_indicators();
  void draw_included_coefficients();
  void allow_model_selection(bool allow) { allow_model_selection_ = allow; }
  void limit_model_selection(int max_flips) { max_flips_ = max_flips; }
  // set_active_set_statistics(true) (DeviceImputerBase) applies here too: the view then carries the statistics divided by sigsq

 protected:
  double weighted_sum_of_squared_errors() override;

 private:
  const BOOM_B200::WeightedRegSuf &scaled_statistics();   // X'WX / sigsq, X'Wy / sigsq (SpikeSlabSampler.cpp:131-134,188-193)
  BOOM_B200::StatView &scaled_view();
  Ptr<VariableSelectionPrior> spike_;
  BOOM_B200::WeightedRegSuf scaled_;
  std::unique_ptr<BOOM_B200::StatView> tview_;
  BOOM_B200::Vector tview_xty_;
  double tview_sigsq_ = 0.0;
  uint64_t tview_iteration_ = ~0ull;
  bool allow_model_selection_ = true;
  int max_flips_ = -1;
};

class PoissonRegressionSpikeSlabSampler : public PoissonRegressionAuxMixSampler {
 public:
  PoissonRegressionSpikeSlabSampler(PoissonRegressionModel *model, const Ptr<MvnBase> &slab,
                                    const Ptr<VariableSelectionPrior> &spike, int number_of_threads = 1,
                                    RNG &seeding_rng = GlobalRng::rng);
  void draw() override;
  double logpri() const override;
  void allow_model_selection(bool tf) { allow_model_selection_ = tf; }
  void limit_model_selection(int max_flips) { max_flips_ = max_flips; }
  void draw_model_indicators();                         // SpikeSlabSampler.cpp:40-100 with sigsq = 1
  void draw_beta();
  double log_model_prob(const Selector &gamma) const;
  PoissonRegressionSpikeSlabSampler *clone_to_new_host(Model *new_host) const override;
  void find_posterior_mode(double epsilon = 1e-5) override;
  bool can_find_posterior_mode() const override { return true; }
  double log_posterior_at_mode() const { return log_posterior_at_mode_; }

 private:
  Ptr<MvnBase> slab_;
  Ptr<VariableSelectionPrior> spike_;
  bool allow_model_selection_ = true;
  int max_flips_ = -1;
  double log_posterior_at_mode_ = negative_infinity();
};

}  // namespace B200
}  // namespace BOOM
