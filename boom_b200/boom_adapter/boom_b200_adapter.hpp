// boom_b200_adapter.hpp -- the drop-in for an existing BOOM program.
//
// Samplers derived from BOOM::PosteriorSampler (Models/PosteriorSamplers/PosteriorSampler.hpp:44-107)
// that are constructed on the reference's OWN model objects and attach with model->set_method(sampler):
//
//   BOOM::BinomialLogitAuxmixSampler        -> BOOM::B200::BinomialLogitAuxmixSampler
//   BOOM::BinomialLogitSpikeSlabSampler     -> BOOM::B200::BinomialLogitSpikeSlabSampler
//   BOOM::PoissonRegressionAuxMixSampler    -> BOOM::B200::PoissonRegressionAuxMixSampler
//   BOOM::PoissonRegressionSpikeSlabSampler -> BOOM::B200::PoissonRegressionSpikeSlabSampler
//
// Same constructor arguments and public methods as the originals
// (Models/Glm/PosteriorSamplers/BinomialLogitAuxmixSampler.hpp:109-152, BinomialLogitSpikeSlabSampler.hpp:27-96,
// PoissonRegressionAuxMixSampler.hpp:60-125, PoissonRegressionSpikeSlabSampler.hpp).  They cannot subclass the
// originals: the statistics those fill are private with no bulk setter, so they derive from
// PosteriorSampler directly and keep their statistics in a BOOM::WeightedRegSuf (bulk-loaded with reset()).
//
// impute_latent_data() is one device step through the C ABI (include/boomgpu.h) on rows packed once from
// model->dat() (re-packed when the model's data change: IID_DataPolicy::add_observer, IID_DataPolicy.hpp:43-45);
// the small-state steps run on the host: beta by BOOM's own rmvn_suf_mt, the inclusion sweep by the shared
// bordered-Cholesky evaluator of boom_b200/host.  All randomness derives from the sampler's BOOM::RNG, so
// set_seed() repeats a chain.  Mixture tables are read from the live reference objects
// (BinomialLogitDataImputer::mixture_approximation, create_poisson_mixture_approximation_table()).
//
// Needs the BOOM headers and library: built only where the reference exists.
#pragma once

#include <functional>
#include <memory>

#include "Models/Glm/BinomialLogitModel.hpp"
#include "Models/Glm/PoissonRegressionModel.hpp"
#include "Models/Glm/VariableSelectionPrior.hpp"
#include "Models/Glm/WeightedRegressionModel.hpp"
#include "Models/MvnBase.hpp"
#include "Models/PosteriorSamplers/PosteriorSampler.hpp"
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
  void clear_complete_data_sufficient_statistics() { suf_.clear(); }
  // multi-GPU / placement
  void set_device(int device);
  void set_row_offset(uint64_t first_global_row) { row_offset_ = first_global_row; stale_ = true; }
  void set_allreduce(const BOOM_B200::AllReduceFn &fn) { allreduce_ = fn; }
  // or natively: join an NCCL communicator (id from BOOM_B200::GlmModelBase::comm_unique_id() on rank 0)
  void set_communicator(const std::string &id, int nranks, int rank) { comm_id_ = id; comm_ranks_ = nranks; comm_rank_ = rank; comm_dirty_ = true; }

 protected:
  DeviceImputerBase(int xdim, RNG &seeding_rng);
  virtual void pack_and_upload(boomgpu_ctx *ctx) = 0;   // walks model->dat()
  virtual int device_step(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, double *suf_dev) = 0;
  virtual const Vector &current_beta() const = 0;
  // log likelihood with gradient / Hessian at a full coefficient vector, one device pass (boomgpu_*_loglike_derivs)
  virtual int device_loglike_derivs(boomgpu_ctx *ctx, const double *beta, double *loglike, double *g, double *h) = 0;
  void ensure_device_rows();
  // find_posterior_mode of the spike-and-slab samplers (BinomialLogitSpikeSlabSampler.cpp:147-177,
  // PoissonRegressionSpikeSlabSampler.cpp:69-106): Newton-Raphson on the included coefficients, derivatives from the device
  bool find_mode(GlmCoefs &coef, const MvnBase &slab, const VariableSelectionPrior &spike, double epsilon, double *value);
  void mark_stale() { stale_ = true; }
  void check(int rc) const;
  // host steps on suf_
  void draw_beta_full_model(GlmCoefs &coef, const MvnBase &prior);
  void spike_slab_draw(GlmCoefs &coef, const MvnBase &slab, const VariableSelectionPrior &spike, bool select, int max_flips,
                       bool fisher_yates);
  double spike_slab_logpri(const GlmCoefs &coef, const MvnBase &slab, const VariableSelectionPrior &spike) const;

  WeightedRegSuf suf_;
  int xdim_;

 private:
  boomgpu_ctx *ctx_ = nullptr;
  int device_ = 0;
  bool stale_ = true, latent_data_fixed_ = false, repack_each_time_ = false;
  uint64_t row_offset_ = 0, iteration_ = 0;
  BOOM_B200::AllReduceFn allreduce_;
  std::string comm_id_;
  int comm_ranks_ = 1, comm_rank_ = 0;
  bool comm_dirty_ = false;
  std::vector<double> packed_;
};

class BinomialLogitAuxmixSampler : public DeviceImputerBase {
 public:
  BinomialLogitAuxmixSampler(BinomialLogitModel *model, const Ptr<MvnBase> &prior, int clt_threshold = 10,
                             RNG &seeding_rng = GlobalRng::rng);
  void draw() override;
  double logpri() const override;
  void draw_params();
  const WeightedRegSuf &suf() const { return suf_; }
  int clt_threshold() const { return clt_threshold_; }
  void update_complete_data_sufficient_statistics(double precision_weighted_sum, double total_precision, const Vector &x);

 protected:
  void pack_and_upload(boomgpu_ctx *ctx) override;
  int device_step(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, double *suf_dev) override;
  int device_loglike_derivs(boomgpu_ctx *ctx, const double *beta, double *loglike, double *g, double *h) override;
  const Vector &current_beta() const override { return model_->Beta(); }
  BinomialLogitModel *model_;
  Ptr<MvnBase> prior_;

 private:
  int clt_threshold_;
};

class BinomialLogitSpikeSlabSampler : public BinomialLogitAuxmixSampler {
 public:
  BinomialLogitSpikeSlabSampler(BinomialLogitModel *model, const Ptr<MvnBase> &slab, const Ptr<VariableSelectionPrior> &spike,
                                int clt_threshold, RNG &seeding_rng = GlobalRng::rng);
  BinomialLogitSpikeSlabSampler *clone_to_new_host(Model *model) const override;
  void draw() override;
  double logpri() const override;
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

class PoissonRegressionAuxMixSampler : public DeviceImputerBase {
 public:
  PoissonRegressionAuxMixSampler(PoissonRegressionModel *model, const Ptr<MvnBase> &prior, int number_of_threads = 1,
                                 RNG &seeding_rng = GlobalRng::rng);
  void draw() override;
  double logpri() const override;
  void draw_beta_given_complete_data();
  const WeightedRegSuf &complete_data_sufficient_statistics() const { return suf_; }
  void update_complete_data_sufficient_statistics(double precision_weighted_sum, double total_precision, const Vector &x);

 protected:
  void pack_and_upload(boomgpu_ctx *ctx) override;
  int device_step(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, double *suf_dev) override;
  int device_loglike_derivs(boomgpu_ctx *ctx, const double *beta, double *loglike, double *g, double *h) override;
  const Vector &current_beta() const override { return model_->Beta(); }
  PoissonRegressionModel *model_;
  Ptr<MvnBase> prior_;
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
