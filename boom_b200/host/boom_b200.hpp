// boom_b200.hpp -- C++ host side of the B200-native auxiliary-mixture Gibbs samplers.
//
// The classes keep the surface of the reference (steve-the-bayesian/BOOM; file:line relative to
// its root) for this path so that a BOOM user finds the calls they know:
//
//   Model::set_method / sample_posterior        Models/ModelTypes.hpp:81-100, Models/Policies/PriorPolicy.cpp:25-39
//   PosteriorSampler::{draw,logpri,rng,set_seed} Models/PosteriorSamplers/PosteriorSampler.hpp:44-107
//   BinomialLogitModel                          Models/Glm/BinomialLogitModel.hpp:33-87
//   PoissonRegressionModel                      Models/Glm/PoissonRegressionModel.hpp:35-94
//   MvnModel (as the MvnBase prior)             Models/MvnBase.hpp:92-139
//   VariableSelectionPrior                      Models/Glm/VariableSelectionPrior.hpp:99-101, .cpp:271-300
//   BinomialLogitAuxmixSampler                  Models/Glm/PosteriorSamplers/BinomialLogitAuxmixSampler.hpp:109-152
//   BinomialLogitSpikeSlabSampler               Models/Glm/PosteriorSamplers/BinomialLogitSpikeSlabSampler.hpp:27-96
//   PoissonRegressionAuxMixSampler              Models/Glm/PosteriorSamplers/PoissonRegressionAuxMixSampler.hpp:37-125
//   PoissonRegressionSpikeSlabSampler           Models/Glm/PosteriorSamplers/PoissonRegressionSpikeSlabSampler.cpp:55-59
//
// What changes underneath: the worker pool of Models/PosteriorSamplers/Imputer.hpp:241-343 (one
// virtual call and one rank-1 update per observation) is ONE device step through the C ABI of
// include/boomgpu.h.  The small-state steps (Cholesky draw of beta, the sweep over the inclusion
// indicators) stay on the host as in the reference.  There is no CPU fallback for the device step.
//
// This layer is standalone (no BOOM headers) so it builds and runs on the GPU box; the adapter in
// boom_b200/boom_adapter derives the same samplers from BOOM::PosteriorSampler for a true drop-in.
#pragma once

#include <cmath>
#include <cstdint>
#include <functional>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

struct boomgpu_ctx;

namespace BOOM_B200 {

typedef std::vector<double> Vector;

// report_error (cpputil/report_error.cpp:29-31): everything surfaces as std::runtime_error.
[[noreturn]] void report_error(const std::string &msg);

// distributions/rng.hpp:28-54
class RNG {
 public:
  typedef std::uint_fast64_t RngIntType;
  RNG() : generator_(8675309) {}
  explicit RNG(RngIntType seed) : generator_(seed) {}
  void seed(RngIntType s) { generator_.seed(s); }
  double operator()() { return dist_(generator_); }
  std::mt19937_64 &generator() { return generator_; }

 private:
  std::mt19937_64 generator_;
  std::uniform_real_distribution<double> dist_;
};
struct GlobalRng { static RNG rng; };
RNG::RngIntType seed_rng(RNG &rng = GlobalRng::rng);  // distributions/rng.cpp:39-47
double runif_mt(RNG &rng, double lo = 0, double hi = 1);
double rnorm_mt(RNG &rng, double mu = 0, double sd = 1);
int random_int_mt(RNG &rng, int lo, int hi);            // distributions/random_int.cpp:26-29

// Dense symmetric matrix, row major (== column major), p x p.
struct SpdMatrix {
  int dim = 0;
  Vector a;
  SpdMatrix() {}
  explicit SpdMatrix(int p, double diag = 0.0) : dim(p), a((size_t)p * p, 0.0) {
    for (int i = 0; i < p; ++i) a[(size_t)i * p + i] = diag;
  }
  double &operator()(int i, int j) { return a[(size_t)i * dim + j]; }
  double operator()(int i, int j) const { return a[(size_t)i * dim + j]; }
  int nrow() const { return dim; }
};

// LinAlg/Selector.hpp: which coefficients are in the model.
class Selector {
 public:
  Selector() {}
  explicit Selector(int n, bool all = true) : inc_(n, all) {}
  int nvars_possible() const { return (int)inc_.size(); }
  int nvars() const;
  bool operator[](int i) const { return inc_[i]; }
  void flip(int i) { inc_[i] = !inc_[i]; }
  void add(int i) { inc_[i] = true; }
  void drop(int i) { inc_[i] = false; }
  void drop_all() { inc_.assign(inc_.size(), false); }
  void add_all() { inc_.assign(inc_.size(), true); }
  std::vector<int> included_positions() const;
  Vector select(const Vector &v) const;
  SpdMatrix select(const SpdMatrix &m) const;  // LinAlg/Selector.cpp:411-426
  Vector expand(const Vector &sub) const;
  const std::vector<bool> &bits() const { return inc_; }

 private:
  std::vector<bool> inc_;
};

// Models/MvnBase.hpp:92-139 -- the prior is read through mu() and siginv().
class MvnBase {
 public:
  virtual ~MvnBase() {}
  virtual int dim() const = 0;
  virtual const Vector &mu() const = 0;
  virtual const SpdMatrix &siginv() const = 0;
  double logp(const Vector &x) const;
};

class MvnModel : public MvnBase {
 public:
  // ivar = false: V is the variance matrix (inverted here); true: V is the precision.
  MvnModel(const Vector &mean, const SpdMatrix &V, bool ivar = false);
  int dim() const override { return (int)mu_.size(); }
  const Vector &mu() const override { return mu_; }
  const SpdMatrix &siginv() const override { return siginv_; }

 private:
  Vector mu_;
  SpdMatrix siginv_;
};

// Models/Glm/VariableSelectionPrior.cpp:271-300
class VariableSelectionPrior {
 public:
  VariableSelectionPrior(int n, double inclusion_probability = 1.0);
  explicit VariableSelectionPrior(const Vector &prior_inclusion_probabilities);
  double logp(const Selector &inc) const;
  double flip_delta(int j, bool included_now) const;
  int max_model_size() const { return max_model_size_; }
  void make_valid(Selector &inc) const;
  const Vector &prior_inclusion_probabilities() const { return probs_; }
  void set_max_model_size(int m) { max_model_size_ = m; }
  int potential_nvars() const { return (int)probs_.size(); }

 private:
  Vector probs_, log_p_, log_q_;
  int max_model_size_ = -1;
};

// Models/Glm/GlmCoefs.hpp:26-140 -- beta with exact zeros at excluded positions.
class GlmCoefs {
 public:
  explicit GlmCoefs(int p, bool all = true) : beta_(p, 0.0), inc_(p, all) {}
  const Vector &Beta() const { return beta_; }
  void set_Beta(const Vector &b);
  const Selector &inc() const { return inc_; }
  void set_inc(const Selector &g);
  Vector included_coefficients() const { return inc_.select(beta_); }
  void set_included_coefficients(const Vector &b);
  void drop_all() { inc_.drop_all(); beta_.assign(beta_.size(), 0.0); }
  void add_all() { inc_.add_all(); }
  void add(int i) { inc_.add(i); }
  void drop(int i) { inc_.drop(i); beta_[i] = 0.0; }
  int nvars() const { return inc_.nvars(); }
  int nvars_possible() const { return inc_.nvars_possible(); }

 private:
  Vector beta_;
  Selector inc_;
};

class PosteriorSampler;

// The hook a multi-GPU launcher installs: sum the packed statistics (count doubles at a DEVICE
// pointer, on the CUDA stream returned by DeviceData::stream()) over all ranks, in place.
typedef std::function<void(double *suf_dev, int64_t count)> AllReduceFn;

// Device residency of one model's data: the replacement of the AoS behind IID_DataPolicy::dat()
// (Models/Policies/IID_DataPolicy.hpp:43-58).  Rows are packed once and re-packed when data change.
class DeviceData {
 public:
  explicit DeviceData(int device = 0);
  ~DeviceData();
  DeviceData(const DeviceData &) = delete;
  boomgpu_ctx *ctx() { return ctx_; }
  void check(int rc) const;  // boomgpu status -> report_error
  int device() const { return device_; }

 private:
  boomgpu_ctx *ctx_ = nullptr;
  int device_ = 0;
};

// Common base of the two regression models: SoA data + coefficients + sampler list.
class GlmModelBase {
 public:
  explicit GlmModelBase(int xdim, bool all = true) : coef_(xdim, all) {}
  virtual ~GlmModelBase() {}
  int xdim() const { return coef_.nvars_possible(); }
  int64_t sample_size() const { return (int64_t)(x_.size() / (size_t)std::max(1, xdim())); }
  GlmCoefs &coef() { return coef_; }
  const GlmCoefs &coef() const { return coef_; }
  const Vector &Beta() const { return coef_.Beta(); }
  void set_Beta(const Vector &b) { coef_.set_Beta(b); }
  Vector included_coefficients() const { return coef_.included_coefficients(); }
  void set_included_coefficients(const Vector &b) { coef_.set_included_coefficients(b); }
  void drop_all() { coef_.drop_all(); }

  // Model::set_method / sample_posterior (Models/Policies/PriorPolicy.cpp:25-39)
  void set_method(const std::shared_ptr<PosteriorSampler> &sampler);
  void clear_methods() { samplers_.clear(); }
  void sample_posterior();
  double logpri() const;

  // log likelihood at a FULL coefficient vector with gradient (p) and Hessian (p x p) from one device pass
  // (BinomialLogitModel.cpp:140-180, PoissonRegressionModel.cpp:56-95); g / h may be null
  virtual double log_likelihood_derivs(const Vector &beta, Vector *g, SpdMatrix *h) = 0;

  // ---- device residency
  void set_device(int device);            // which GPU holds this model's rows (default 0)
  void set_row_offset(uint64_t first_global_row);  // this shard's first global row (multi-GPU)
  void set_stream(void *cuda_stream);     // cudaStream_t the device step (and the all-reduce hook) runs on
  void set_allreduce(const AllReduceFn &fn) { allreduce_ = fn; }
  // native alternative to the hook: join an NCCL communicator through the C ABI (boomgpu_comm_init); id = the 128 bytes
  // rank 0 obtained from comm_unique_id() and handed to the other ranks
  static std::string comm_unique_id();
  void set_communicator(const std::string &id, int nranks, int rank);
  const AllReduceFn &allreduce() const { return allreduce_; }
  // rows already resident in HBM (device pointers): nothing is copied; the caller keeps them alive
  DeviceData &device_data();              // packs / uploads when stale
  const Vector &x_rows() const { return x_; }
  uint64_t data_version() const { return data_version_; }
  // instrumentation of the device context (include/boomgpu.h: boomgpu_set_option / _kernel_launches / _get_timings)
  void set_device_option(const std::string &name, int64_t value);
  int64_t kernel_launches();
  // per kernel class {fused small-p, imputer pass, DMMA SYRK, reductions, other}: CUDA-event ms and launch counts
  void kernel_timings(double ms[5], int64_t launches[5], bool reset);

 protected:
  virtual void upload(DeviceData &dev) = 0;
  void touch() { ++data_version_; }
  Vector x_;  // row major n x p
  bool adopted_ = false, borrowed_ = false;
  int64_t adopted_n_ = 0;
  std::shared_ptr<void> keepalive_;

 private:
  GlmCoefs coef_;
  std::vector<std::shared_ptr<PosteriorSampler>> samplers_;
  std::unique_ptr<DeviceData> dev_;
  int device_ = 0;
  uint64_t row_offset_ = 0;
  uint64_t data_version_ = 1, uploaded_version_ = 0;
  AllReduceFn allreduce_;
  void *stream_ = nullptr;
  bool have_stream_ = false;
  std::vector<std::pair<std::string, int64_t>> options_;
  std::string comm_id_;
  int comm_ranks_ = 1, comm_rank_ = 0;
};

class BinomialLogitModel : public GlmModelBase {
 public:
  explicit BinomialLogitModel(int xdim, bool all = true) : GlmModelBase(xdim, all) {}
  // bulk constructor (Models/Glm/BinomialLogitModel.cpp:40-49): X row major n x p
  BinomialLogitModel(int64_t n, int p, const double *X, const double *y, const double *ntrials);
  // model->add_data(new BinomialRegressionData(y, n, x))  (BinomialRegressionData.hpp:25-55)
  void add_data(double y, double n, const Vector &x);
  void adopt_device_data(int64_t n, const double *dX, int64_t ldx, const double *dy, const double *dntrials);
  // rows that stay in the CALLER's host memory (no copy into the model: at n = 1e7, p = 500 a second 40 GB copy matters);
  // read when the rows are uploaded; keepalive owns whatever keeps the three arrays alive
  void borrow_host_data(int64_t n, const double *X, int64_t ldx, const double *y, const double *ntrials,
                        std::shared_ptr<void> keepalive);
  int64_t nobs() const { return (adopted_ || borrowed_) ? adopted_n_ : (int64_t)y_.size(); }
  // Models/Glm/BinomialLogitModel.cpp:140-180, value only, evaluated on the device
  double log_likelihood(const Vector &beta);
  double log_likelihood() { return log_likelihood(Beta()); }
  double log_likelihood_derivs(const Vector &beta, Vector *g, SpdMatrix *h) override;
  // nonevent down-sampling offset (BinomialLogitModel.hpp:82-86): enters the likelihood, not the auxmix draws
  void set_nonevent_sampling_prob(double alpha);
  double log_alpha() const { return log_alpha_; }

 protected:
  void upload(DeviceData &dev) override;

 private:
  double log_alpha_ = 0.0;
  Vector y_, n_;
  const double *dX_ = nullptr, *dy_ = nullptr, *dn_ = nullptr;
  int64_t dldx_ = 0;
};

class PoissonRegressionModel : public GlmModelBase {
 public:
  explicit PoissonRegressionModel(int xdim, bool all = true) : GlmModelBase(xdim, all) {}
  PoissonRegressionModel(int64_t n, int p, const double *X, const int64_t *y, const double *exposure);
  // model->add_data(new PoissonRegressionData(y, x, exposure))  (PoissonRegressionData.hpp:25-62)
  void add_data(int64_t y, const Vector &x, double exposure = 1.0);
  void adopt_device_data(int64_t n, const double *dX, int64_t ldx, const int64_t *dy, const double *dexposure);
  void borrow_host_data(int64_t n, const double *X, int64_t ldx, const int64_t *y, const double *exposure,
                        std::shared_ptr<void> keepalive);
  int64_t nobs() const { return (adopted_ || borrowed_) ? adopted_n_ : (int64_t)y_.size(); }
  double log_likelihood(const Vector &beta);
  double log_likelihood() { return log_likelihood(Beta()); }
  double log_likelihood_derivs(const Vector &beta, Vector *g, SpdMatrix *h) override;
  const std::vector<int64_t> &y() const { return y_; }

 protected:
  void upload(DeviceData &dev) override;

 private:
  std::vector<int64_t> y_;
  Vector exposure_;
  const double *dX_ = nullptr, *dexp_ = nullptr;
  const int64_t *dy_ = nullptr;
  int64_t dldx_ = 0;
};

// Models/PosteriorSamplers/PosteriorSampler.hpp:44-107
class PosteriorSampler {
 public:
  explicit PosteriorSampler(RNG &seeding_rng) : rng_(seed_rng(seeding_rng)) {}
  virtual ~PosteriorSampler() {}
  virtual void draw() = 0;
  virtual double logpri() const = 0;
  RNG &rng() const { return rng_; }
  void set_seed(unsigned long s) { rng_.seed(s); on_seed(); }

 protected:
  virtual void on_seed() {}

 private:
  mutable RNG rng_;
};

// BinomialLogit::SufficientStatistics (BinomialLogitAuxmixSampler.hpp:39-67) and, with the four
// scalars, WeightedRegSuf (Models/Glm/WeightedRegressionModel.hpp:40-110).
class WeightedRegSuf {
 public:
  explicit WeightedRegSuf(int p = 0) : xtx_(p), xty_(p, 0.0) {}
  WeightedRegSuf(const WeightedRegSuf &rhs) : xtx_(rhs.xtx_), xty_(rhs.xty_), n_(rhs.n_), yty_(rhs.yty_), sumw_(rhs.sumw_),
                                              sumlogw_(rhs.sumlogw_) {}   // a copy is never page-locked
  WeightedRegSuf &operator=(const WeightedRegSuf &rhs);
  ~WeightedRegSuf();
  void clear();
  // rank-1 update on the host: the path state-space callers drive (fix_latent_data(true))
  void add_data(const Vector &x, double y, double w);            // WeightedRegressionModel.cpp:161-169
  void update(const Vector &x, double weighted_value, double weight);  // BinomialLogitAuxmixSampler.cpp:61-67
  // bulk load of a device result (WeightedRegSuf::reset, WeightedRegressionModel.cpp:148-157)
  void reset(const double *packed, int p);
  // in-place bulk load: the device step writes X'WX and X'Wz straight into this object's storage
  // (at p = 4000 the matrix is 128 MB: every intermediate copy costs ~10 ms)
  double *xtx_storage(int p);
  double *xty_storage() { return xty_.data(); }
  void set_scalars(double n, double yty, double sumw, double sumlogw) { n_ = n; yty_ = yty; sumw_ = sumw; sumlogw_ = sumlogw; }
  const SpdMatrix &xtx() const { return xtx_; }
  const Vector &xty() const { return xty_; }
  double n() const { return n_; }
  double yty() const { return yty_; }
  double sumw() const { return sumw_; }
  double sumlogw() const { return sumlogw_; }
  int64_t sample_size() const { return (int64_t)(n_ + 0.5); }

 private:
  void unpin();
  SpdMatrix xtx_;
  Vector xty_;
  double n_ = 0, yty_ = 0, sumw_ = 0, sumlogw_ = 0;
  double *pinned_ = nullptr;   // xtx_'s storage while it is page-locked for direct device->host copies (large p)
};
typedef WeightedRegSuf SufficientStatistics;

// What the spike-and-slab host steps read of the complete-data statistics: entries of X'WX and X'Wz.  Either the full
// matrix (a WeightedRegSuf), or the ACTIVE-SET form: the columns of the variables in the model, the diagonal
// and X'Wz, as an active-set device step returns them, with any further column fetched from the device on demand
// (boomgpu_weighted_column: the latents stay in HBM until the next step).  A sweep reads X'WX[j, gamma + {j}] only, so the
// active form answers every proposal from what it holds; a column is fetched when the sweep ADDS a variable outside it.
class StatView {
 public:
  explicit StatView(const WeightedRegSuf &full);
  // active form: cols = the columns held (G is p x cols.size(), row major), fetch(j, out) writes column j of X'WX (p doubles)
  StatView(int p, const std::vector<int> &cols, const Vector &G, const Vector &diag, const Vector &xty,
           std::function<void(int, double *)> fetch);
  double at(int i, int j);               // (X'WX)[i, j]
  double xty(int j) const { return xty_[j]; }
  void ensure_column(int j);             // active form: make column j resident (no-op in the full form / when it is held)
  int columns_fetched() const { return fetched_; }
  bool is_full() const { return full_ != nullptr; }

 private:
  const double *full_ = nullptr;
  int p_ = 0, ld_ = 0, k_ = 0, fetched_ = 0;
  std::vector<int> where_;               // column -> slot in G_, or -1
  Vector G_, diag_;
  const double *xty_ = nullptr;
  std::function<void(int, double *)> fetch_;
  Vector column_;
};

// The host-side spike-and-slab steps shared by the logit and Poisson samplers
// (BinomialLogitSpikeSlabSampler.cpp:56-117,180-222; SpikeSlabSampler.cpp:40-216 with sigsq = 1).
class SpikeSlabCore {
 public:
  SpikeSlabCore(const std::shared_ptr<MvnBase> &slab, const std::shared_ptr<VariableSelectionPrior> &spike,
                bool fisher_yates);
  double log_model_prob(const Selector &g, const WeightedRegSuf &suf) const;
  void draw_model_indicators(RNG &rng, GlmCoefs &coef, const WeightedRegSuf &suf) const;
  void draw_beta(RNG &rng, GlmCoefs &coef, const WeightedRegSuf &suf) const;
  // the same on a view of the statistics (full or active-set form)
  double log_model_prob(const Selector &g, StatView &stats) const;
  void draw_model_indicators(RNG &rng, GlmCoefs &coef, StatView &stats) const;
  void draw_beta(RNG &rng, GlmCoefs &coef, StatView &stats) const;
  double logpri(const GlmCoefs &coef) const;
  void allow_model_selection(bool tf) { allow_model_selection_ = tf; }
  void limit_model_selection(int max_flips) { max_flips_ = max_flips; }
  bool model_selection_allowed() const { return allow_model_selection_; }
  int max_flips() const { return max_flips_; }
  // BinomialLogitSpikeSlabSampler::set_spike / set_slab (.hpp:68-74): the dimension is checked against the other prior
  void set_spike(const std::shared_ptr<VariableSelectionPrior> &spike);
  void set_slab(const std::shared_ptr<MvnBase> &slab);
  const std::shared_ptr<MvnBase> &slab() const { return slab_; }
  const std::shared_ptr<VariableSelectionPrior> &spike() const { return spike_; }
  // Newton-Raphson (with step halving) on the included coefficients of log slab(beta_gamma) + log likelihood, the
  // objective of BinomialLogitSpikeSlabSampler::find_posterior_mode (.cpp:123-177) / PoissonRegressionSpikeSlabSampler
  // (.cpp:69-106); sets the model's included coefficients on success.  Returns false when gamma is empty or on failure.
  bool find_posterior_mode(GlmModelBase &model, double epsilon, double *log_posterior_at_mode) const;
  // test hook: log_model_prob of every proposal along a path of flips (accept[i]: commit flip i), evaluated the
  // way the sweep evaluates it (bordered Cholesky factors); the last entry is the final model's value
  Vector flip_path_log_probs(const Selector &start, const WeightedRegSuf &suf, const std::vector<int> &flips,
                             const std::vector<bool> &accept) const;
  Vector flip_path_log_probs(const Selector &start, StatView &stats, const std::vector<int> &flips,
                             const std::vector<bool> &accept) const;

 private:
  class FlipEvaluator;
  std::shared_ptr<MvnBase> slab_;
  std::shared_ptr<VariableSelectionPrior> spike_;
  bool fisher_yates_;
  bool allow_model_selection_ = true;
  int max_flips_ = -1;
};

// State of the active-set form of a spike-and-slab sampler's statistics (see StatView): what the last device step returned
// for the columns of the model's included variables.
struct ActiveSetState {
  bool enabled = false;        // the sampler option
  bool valid = false;          // the last imputation ran in active-set form (suf_ then lags until somebody asks for it)
  std::vector<int> cols;
  Vector G, diag, xty;
  double scalars[4] = {0, 0, 0, 0};
  int64_t columns_fetched = 0; // over the life of the sampler (a sweep fetches one per accepted add outside the set)
};

// ---------------------------------------------------------------------------------------------
class BinomialLogitAuxmixSampler : public PosteriorSampler {
 public:
  BinomialLogitAuxmixSampler(BinomialLogitModel *model, const std::shared_ptr<MvnBase> &prior,
                             int clt_threshold = 10, RNG &seeding_rng = GlobalRng::rng);
  void draw() override;                // impute_latent_data(); draw_params();   (.cpp:115-118)
  double logpri() const override;
  void impute_latent_data();           // ONE device step instead of the worker pool
  void draw_params();                  // .cpp:125-130
  // the full statistics; after an active-set step they are computed on demand from the latents still in HBM
  const SufficientStatistics &suf() const { if (active_.valid) materialize_full_statistics(); return suf_; }
  int clt_threshold() const { return clt_threshold_; }
  void clear_complete_data_sufficient_statistics() { active_.valid = false; suf_.clear(); }
  void update_complete_data_sufficient_statistics(double precision_weighted_sum, double total_precision,
                                                  const Vector &x) { active_.valid = false; suf_.update(x, precision_weighted_sum, total_precision); }
  // LatentDataSampler surface (Models/PosteriorSamplers/Imputer.hpp:260-314)
  void fix_latent_data(bool fixed = true) { latent_data_fixed_ = fixed; }
  void set_number_of_workers(int) {}       // the device replaces the worker pool
  void reassign_data_each_time(bool) {}    // rows are re-packed whenever the model's data change
  uint64_t iteration() const { return iteration_; }

 protected:
  void on_seed() override;
  // active-set form of the imputation (p > 64, no caller-supplied all-reduce hook): statistics for the columns `cols`;
  // returns false (and does nothing) where it does not apply, and the caller runs the full step
  bool impute_latent_data_active(const std::vector<int> &cols);
  // the statistics the host steps read: the active-set form when the last imputation produced it, the full matrix otherwise
  std::unique_ptr<StatView> statistics_view();
  void materialize_full_statistics() const;
  BinomialLogitModel *model_;
  std::shared_ptr<MvnBase> prior_;
  mutable ActiveSetState active_;

 private:
  mutable SufficientStatistics suf_;
  int clt_threshold_;
  bool latent_data_fixed_ = false;
  uint64_t device_seed_, iteration_ = 0;
  Vector packed_;
};

class BinomialLogitSpikeSlabSampler : public BinomialLogitAuxmixSampler {
 public:
  BinomialLogitSpikeSlabSampler(BinomialLogitModel *model, const std::shared_ptr<MvnBase> &slab,
                                const std::shared_ptr<VariableSelectionPrior> &spike, int clt_threshold = 5,
                                RNG &seeding_rng = GlobalRng::rng);
  void draw() override;                // .cpp:50-54
  double logpri() const override;
  void draw_model_indicators();        // .cpp:180-211
  void draw_beta();                    // .cpp:56-75
  double log_model_prob(const Selector &g) const;  // .cpp:88-117
  void allow_model_selection(bool tf) { core_.allow_model_selection(tf); }
  void limit_model_selection(int max_flips) { core_.limit_model_selection(max_flips); }
  void set_spike(const std::shared_ptr<VariableSelectionPrior> &spike) { core_.set_spike(spike); }
  void set_slab(const std::shared_ptr<MvnBase> &slab) { core_.set_slab(slab); prior_ = slab; }
  // Active-set statistics (off by default): each iteration the device computes X'WX only for the columns of the
  // variables in the model (plus the diagonal and X'Wz) and the sweep fetches a further column when it adds a variable --
  // the same chain as with the full matrix (the sweep reads nothing else), for n p (2 |gamma| + 4) instead of n p (p + 1)
  // flops.  suf() still answers with the full matrix (computed on demand).  Applies to p > 64 without an all-reduce hook.
  void set_active_set_statistics(bool tf) { active_.enabled = tf; }
  bool active_set_statistics() const { return active_.enabled; }
  int64_t active_set_columns_fetched() const { return active_.columns_fetched; }
  // clone_to_new_host (.cpp:42-48): the same priors and settings on another model, seeded from this sampler's stream
  std::shared_ptr<BinomialLogitSpikeSlabSampler> clone_to_new_host(BinomialLogitModel *new_host) const;
  int xdim() const { return model_->xdim(); }
  void find_posterior_mode(double epsilon = 1e-5);   // .cpp:147-177
  bool can_find_posterior_mode() const { return true; }
  bool posterior_mode_found() const { return posterior_mode_found_; }
  double log_posterior_at_mode() const { return log_posterior_at_mode_; }

 private:
  SpikeSlabCore core_;
  std::unique_ptr<StatView> kept_view_;
  bool posterior_mode_found_ = false;
  double log_posterior_at_mode_ = -1.0 / 0.0;
};

// NormalMixtureApproximation / NormalMixtureApproximationTable (Models/Glm/PosteriorSamplers/NormalMixtureApproximation.hpp:
// 60-200, 262-330), as far as the Poisson samplers use them: a table of finite normal mixtures approximating the
// -log Gamma(nu, 1) densities, extended on demand for counts off its grid (mixture_table.cpp).
struct NormalMixtureApproximation {
  Vector mu, sigma, weights;
  double kullback_leibler = -1.0 / 0.0;
  int number_of_function_evaluations = -1;
  int dim() const { return (int)mu.size(); }
  void order_by_mu();
};
double kullback_leibler_neg_log_gamma(double nu, const NormalMixtureApproximation &approx);   // .cpp:296-319 with NegLogGamma(nu)
NormalMixtureApproximation fit_neg_log_gamma(double nu, const NormalMixtureApproximation &start, double precision, int max_evals,
                                             double stepsize);
class NormalMixtureApproximationTable {
 public:
  void deserialize(const Vector &serialized);   // [nu, K, w[K], sigma[K], mu[K]] ... (.cpp:393-399, 544-558)
  Vector serialize() const;
  void add(int64_t nu, const NormalMixtureApproximation &approximation);
  bool contains(int64_t nu) const;
  bool empty() const { return index_.empty(); }
  size_t size() const { return index_.size(); }
  int64_t smallest_index() const { return index_.front(); }
  int64_t largest_index() const { return index_.back(); }
  // the entry for nu; when nu is off the grid it is interpolated or fitted, ADDED to the table, and returned (.cpp:472-532)
  const NormalMixtureApproximation &approximate(int64_t nu);
  const std::vector<int64_t> &index() const { return index_; }
  const NormalMixtureApproximation &entry(size_t e) const { return approximations_[e]; }

 private:
  void insert(int64_t nu, const NormalMixtureApproximation &approximation, bool derived);
  std::vector<int64_t> index_;
  std::vector<NormalMixtureApproximation> approximations_;
  std::vector<char> derived_;   // entries approximate() added (never used as interpolation neighbours: see mixture_table.cpp)
};

class PoissonRegressionAuxMixSampler : public PosteriorSampler {
 public:
  PoissonRegressionAuxMixSampler(PoissonRegressionModel *model, const std::shared_ptr<MvnBase> &prior,
                                 int number_of_threads = 1, RNG &seeding_rng = GlobalRng::rng);
  void draw() override;                // .cpp:108-111
  double logpri() const override;
  void impute_latent_data();
  void draw_beta_given_complete_data();  // .cpp:123-128
  // the full statistics; after an active-set step they are computed on demand from the latents still in HBM
  const WeightedRegSuf &complete_data_sufficient_statistics() const { if (active_.valid) materialize_full_statistics(); return suf_; }
  void clear_complete_data_sufficient_statistics() { active_.valid = false; suf_.clear(); }
  void update_complete_data_sufficient_statistics(double precision_weighted_sum, double total_precision,
                                                  const Vector &x) {  // .cpp:153-158
    active_.valid = false;
    suf_.add_data(x, precision_weighted_sum / total_precision, total_precision);
  }
  void fix_latent_data(bool fixed = true) { latent_data_fixed_ = fixed; }
  void set_number_of_workers(int) {}
  // The mixture table (PoissonDataImputer::mixture_table_): serialized as
  // NormalMixtureApproximationTable::serialize() writes it (NormalMixtureApproximation.cpp:534-542).
  // Must hold every distinct count of the data below largest_index.
  static void set_mixture_table(const Vector &serialized, int64_t largest_index);
  static bool mixture_table_is_set();
  // the table as it stands now (entries added for off-grid counts included), in the reference's serialize() layout
  static Vector mixture_table();
  // PoissonDataImputer's use of NormalMixtureApproximationTable::approximate (poisson_mixture_approximation_table.cpp:44-61)
  // for one count: returns the entry, adding it to the table when nu is off the grid
  static NormalMixtureApproximation approximate(int64_t nu);

 protected:
  void on_seed() override;
  // active-set form of the imputation, as on the logit samplers (BinomialLogitAuxmixSampler::impute_latent_data_active)
  bool impute_latent_data_active(const std::vector<int> &cols);
  std::unique_ptr<StatView> statistics_view();
  void materialize_full_statistics() const;
  PoissonRegressionModel *model_;
  std::shared_ptr<MvnBase> prior_;
  mutable WeightedRegSuf suf_;
  mutable ActiveSetState active_;

 private:
  int ensure_table(boomgpu_ctx *ctx);   // the table on the device holds every count of the data (extends it where the grid lacks one)
  bool latent_data_fixed_ = false;
  uint64_t device_seed_, iteration_ = 0;
  uint64_t counts_checked_data_version_ = 0, counts_checked_table_version_ = 0;
  const void *counts_checked_ctx_ = nullptr;
  Vector packed_;
};

class PoissonRegressionSpikeSlabSampler : public PoissonRegressionAuxMixSampler {
 public:
  PoissonRegressionSpikeSlabSampler(PoissonRegressionModel *model, const std::shared_ptr<MvnBase> &slab,
                                    const std::shared_ptr<VariableSelectionPrior> &spike, int number_of_threads = 1,
                                    RNG &seeding_rng = GlobalRng::rng);
  void draw() override;                // PoissonRegressionSpikeSlabSampler.cpp:55-59
  double logpri() const override;
  void allow_model_selection(bool tf) { core_.allow_model_selection(tf); }
  void limit_model_selection(int max_flips) { core_.limit_model_selection(max_flips); }
  void draw_model_indicators();        // SpikeSlabSampler.cpp:40-100 with sigsq = 1
  void draw_beta();                    // SpikeSlabSampler.cpp:102-140
  double log_model_prob(const Selector &g) const;
  // Active-set statistics, as on BinomialLogitSpikeSlabSampler (off by default; the same chain as with the full matrix)
  void set_active_set_statistics(bool tf) { active_.enabled = tf; }
  bool active_set_statistics() const { return active_.enabled; }
  int64_t active_set_columns_fetched() const { return active_.columns_fetched; }
  std::shared_ptr<PoissonRegressionSpikeSlabSampler> clone_to_new_host(PoissonRegressionModel *new_host) const;   // .cpp:44-50
  void find_posterior_mode(double epsilon = 1e-5);   // PoissonRegressionSpikeSlabSampler.cpp:69-106
  bool can_find_posterior_mode() const { return true; }
  double log_posterior_at_mode() const { return log_posterior_at_mode_; }

 private:
  SpikeSlabCore core_;
  std::unique_ptr<StatView> kept_view_;
  double log_posterior_at_mode_ = -1.0 / 0.0;
};

// ---- the probit sibling --------------------------------------------------------
// BinomialProbitModel (Models/Glm/BinomialProbitModel.hpp): the same data as the logit model, probit link.  Only what the
// sampler needs is provided here: the data, the coefficients, device residency.
class BinomialProbitModel : public BinomialLogitModel {
 public:
  explicit BinomialProbitModel(int xdim, bool all = true) : BinomialLogitModel(xdim, all) {}
  BinomialProbitModel(int64_t n, int p, const double *X, const double *y, const double *ntrials) : BinomialLogitModel(n, p, X, y, ntrials) {}
};

// BinomialProbitSpikeSlabSampler (Models/Glm/PosteriorSamplers/BinomialProbitSpikeSlabSampler.hpp:34-68, .cpp:30-92):
// z_ij ~ N(x_i' beta, 1) truncated by the sign of trial j; the complete-data statistics are X'NX (N = diag(n_i): it does not
// depend on beta and is computed ONCE, refresh_xtx) and X'z (one HBM-bound pass over X per iteration), followed by the
// generic spike-and-slab steps with residual variance 1 (SpikeSlabSampler.cpp:40-216).
class BinomialProbitSpikeSlabSampler : public PosteriorSampler {
 public:
  BinomialProbitSpikeSlabSampler(BinomialProbitModel *model, const std::shared_ptr<MvnBase> &slab,
                                 const std::shared_ptr<VariableSelectionPrior> &spike, int clt_threshold = 10,
                                 RNG &seeding_rng = GlobalRng::rng);
  void draw() override;                 // .cpp:42-47
  double logpri() const override;
  void allow_model_selection(bool tf) { core_.allow_model_selection(tf); }
  void limit_model_selection(int max_flips) { core_.limit_model_selection(max_flips); }
  void impute_latent_data();            // .cpp:58-70: one device step
  void refresh_xtx();                   // .cpp:72-78: the next impute_latent_data recomputes X'NX as well
  WeightedRegSuf complete_data_sufficient_statistics() const { return suf_; }   // by value, as in the reference (.cpp:85-90)
  int clt_threshold() const { return clt_threshold_; }

 protected:
  void on_seed() override;

 private:
  BinomialProbitModel *model_;
  SpikeSlabCore core_;
  WeightedRegSuf suf_;
  int clt_threshold_;
  uint64_t device_seed_, iteration_ = 0, xtx_data_version_ = 0;
  Vector packed_;
};

// ---- the Student-t sibling -----------------------------------------------------
// Models/DoubleModel.hpp: a prior on a scalar is read through logp().
class DoubleModel {
 public:
  virtual ~DoubleModel() {}
  virtual double logp(double x) const = 0;
};
// Models/UniformModel.hpp:71: flat on [lo, hi]
class UniformModel : public DoubleModel {
 public:
  explicit UniformModel(double lo = 0, double hi = 1);
  double logp(double x) const override;
  double lo() const { return lo_; }
  double hi() const { return hi_; }

 private:
  double lo_, hi_;
};
// Models/GammaModel.hpp: GammaModelBase is read through alpha() (shape) and beta() (rate); logp = dgamma(x, a, b, log)
class GammaModelBase : public DoubleModel {
 public:
  virtual double alpha() const = 0;
  virtual double beta() const = 0;
  double logp(double x) const override;
};
class GammaModel : public GammaModelBase {
 public:
  GammaModel(double a, double b);
  double alpha() const override { return a_; }
  double beta() const override { return b_; }

 private:
  double a_, b_;
};
// Models/ChisqModel.hpp: the prior "df observations with standard deviation sigma_estimate" on 1 / sigma^2
class ChisqModel : public GammaModel {
 public:
  ChisqModel(double df, double sigma_estimate) : GammaModel(df / 2.0, df * sigma_estimate * sigma_estimate / 2.0) {}
};
double rgamma_mt(RNG &rng, double a, double b);                   // shape a, RATE b (distributions/Rmath_dist.cpp:72-74)
double rtrun_gamma_mt(RNG &rng, double a, double b, double cut);  // Gamma(a, b) given x > cut (distributions/trun_gamma.cpp:73-106)
double rexp_mt(RNG &rng, double lambda);

// Models/PosteriorSamplers/GenericGaussianVarianceSampler.{hpp,cpp}: sigma^2 | data_df, data_ss under a Gamma prior on
// 1 / sigma^2, optionally truncated to sigma <= sigma_max.
class GenericGaussianVarianceSampler {
 public:
  explicit GenericGaussianVarianceSampler(const std::shared_ptr<GammaModelBase> &prior, double sigma_max = 1.0 / 0.0);
  void set_sigma_max(double sigma_max);
  double sigma_max() const { return sigma_max_; }
  double draw(RNG &rng, double data_df, double data_ss, double prior_sigma_guess_scale_factor = 1.0) const;   // .cpp:44-63
  double posterior_mode(double data_df, double data_ss) const;
  double log_prior(double sigsq) const;                                                                    // .cpp:82-92

 private:
  std::shared_ptr<GammaModelBase> prior_;
  double sigma_max_;
};

// Samplers/ScalarSliceSampler.{hpp,cpp}: Neal's (2003) slice sampler for a scalar log density, with optional finite
// limits; stepping out by doubling (and, for a possibly multimodal target, the randomised doubling of .cpp:199-236).
class ScalarSliceSampler {
 public:
  typedef std::function<double(double)> Fun;
  ScalarSliceSampler(const Fun &logf, bool unimodal = false, double suggested_dx = 1.0, RNG *rng = nullptr);
  void set_rng(RNG *rng) { rng_ = rng; }
  void set_suggested_dx(double dx) { suggested_dx_ = dx; }
  void set_min_dx(double dx) { min_dx_ = dx; }
  void estimate_dx(bool yn) { estimate_dx_ = yn; }
  void set_limits(double lo, double hi) { set_lower_limit(lo); set_upper_limit(hi); }
  void set_lower_limit(double lo);
  void set_upper_limit(double hi);
  void unset_limits() { lo_set_ = hi_set_ = false; }
  double draw(double x);
  double logp(double x) const { return logf_(x); }
  int64_t function_evaluations() const { return evals_; }

 private:
  double f(double x) { ++evals_; return logf_(x); }
  void find_limits(double x);
  bool find_limits_unbounded(double x);
  bool find_upper_limit(double x);
  bool find_lower_limit(double x);
  void double_hi(double x);
  void double_lo(double x);
  void contract(double x, double x_cand, double logp);
  [[noreturn]] void handle_error(const std::string &msg, double x) const;
  Fun logf_;
  RNG *rng_;
  double suggested_dx_, min_dx_ = -1;
  double lo_ = 0, hi_ = 0, lower_bound_ = 0, upper_bound_ = 0, logplo_ = 0, logphi_ = 0, logp_slice_ = 0;
  bool lo_set_ = false, hi_set_ = false, unimodal_, estimate_dx_ = true;
  int64_t evals_ = 0;
};

// TRegressionModel (Models/Glm/TRegression.{hpp,cpp}): y_i = x_i'beta + sigma t_nu.  Parameters beta, sigsq, nu as in the
// reference's ParamPolicy_3 (defaults sigsq = 1, nu = 30: TRegression.cpp:33-35); the rows live in HBM.
class TRegressionModel : public GlmModelBase {
 public:
  explicit TRegressionModel(int xdim) : GlmModelBase(xdim) {}
  TRegressionModel(int64_t n, int p, const double *X, const double *y);   // TRegression.cpp:41-51: X row major n x p
  void add_data(double y, const Vector &x);                              // model->add_data(new RegressionData(y, x))
  void adopt_device_data(int64_t n, const double *dX, int64_t ldx, const double *dy);
  void borrow_host_data(int64_t n, const double *X, int64_t ldx, const double *y, std::shared_ptr<void> keepalive);
  int64_t nobs() const { return (adopted_ || borrowed_) ? adopted_n_ : (int64_t)y_.size(); }
  double sigsq() const { return sigsq_; }
  double sigma() const { return std::sqrt(sigsq_); }
  void set_sigsq(double s2);
  double nu() const { return nu_; }
  void set_nu(double nu);
  // TRegression.cpp:74-86 on the device (one pass over X; the residuals stay in HBM) ...
  double log_likelihood(const Vector &beta, double sigsq, double nu);
  double log_likelihood() { return log_likelihood(Beta(), sigsq_, nu_); }
  // ... and at another (sigsq, nu) for the SAME beta as the previous call: 8 n bytes instead of a pass over X
  double log_likelihood_same_beta(double sigsq, double nu);
  double log_likelihood_derivs(const Vector &beta, Vector *g, SpdMatrix *h) override;   // not provided (TRegression.cpp:118-121)

 protected:
  void upload(DeviceData &dev) override;

 private:
  double sigsq_ = 1.0, nu_ = 30.0;
  Vector y_;
  const double *dX_ = nullptr, *dy_ = nullptr;
  int64_t dldx_ = 0;
};

// TRegressionSampler (Models/Glm/PosteriorSamplers/TRegressionSampler.{hpp,cpp}): the scale-mixture Gibbs sampler.
//   impute_latent_data     w_i ~ Gamma((nu+1)/2, (nu + delta_i^2)/2) and WeightedRegSuf::add_data(x_i, y_i, w_i): ONE device step
//   draw_beta_full_conditional    N((Ominv + X'WX/sigsq)^-1 (Ominv b + X'Wy/sigsq), .)  on the host (.cpp:153-160)
//   draw_sigsq_full_conditional   GenericGaussianVarianceSampler on n and the weighted SSE (.cpp:165-171)
//   draw_nu_given_observed_data   slice sampler on nu over the Student log likelihood: the device keeps y - X beta and
//                                 evaluates each candidate nu from 8 n bytes (.cpp:178-181)
class TRegressionSampler : public PosteriorSampler {
 public:
  TRegressionSampler(TRegressionModel *model, const std::shared_ptr<MvnBase> &coefficient_prior,
                     const std::shared_ptr<GammaModelBase> &siginv_prior, const std::shared_ptr<DoubleModel> &nu_prior,
                     RNG &seeding_rng = GlobalRng::rng);
  void draw() override;                 // .cpp:114-119
  double logpri() const override;       // .cpp:121-126
  void impute_latent_data();            // .cpp:128-143
  void draw_beta_full_conditional();
  void draw_sigsq_full_conditional();
  void draw_nu_given_complete_data();   // .cpp:173-176: from the weights' GammaSuf alone (ScaledChisqModel.cpp:52-82)
  void draw_nu_given_observed_data();
  void set_sigma_upper_limit(double max_sigma) { sigsq_sampler_.set_sigma_max(max_sigma); }
  const WeightedRegSuf &complete_data_sufficient_statistics() const { materialize_full_statistics(); return suf_; }
  void fix_latent_data(bool fixed = true) { latent_data_fixed_ = fixed; }
  void clear_complete_data_sufficient_statistics() { suf_.clear(); }
  void update_complete_data_sufficient_statistics(double y, const Vector &x, double weight) { suf_.add_data(x, y, weight); }
  uint64_t iteration() const { return iteration_; }
  int64_t likelihood_evaluations() const { return nu_observed_.function_evaluations(); }

 protected:
  void on_seed() override;
  void coefficients_changed() { residuals_current_ = false; }
  // sum_i w_i (y_i - x_i'beta)^2 from the statistics (WeightedRegressionModel.cpp:89-95); the spike-and-slab sampler answers
  // from its active-set view when the last imputation ran in that form
  virtual double weighted_sum_of_squared_errors();
  virtual void materialize_full_statistics() const {}
  bool latent_data_is_fixed() const { return latent_data_fixed_; }
  void next_device_key(uint64_t *seed, uint64_t *iteration) { *seed = device_seed_; *iteration = iteration_++; }
  TRegressionModel *model_;
  std::shared_ptr<MvnBase> coefficient_prior_;
  std::shared_ptr<GammaModelBase> siginv_prior_;
  std::shared_ptr<DoubleModel> nu_prior_;
  WeightedRegSuf suf_;                  // its sumw / sumlogw / n are the weight model's GammaSuf as well

 private:
  GenericGaussianVarianceSampler sigsq_sampler_;
  ScalarSliceSampler nu_observed_, nu_complete_;
  bool latent_data_fixed_ = false, residuals_current_ = false;
  uint64_t device_seed_, iteration_ = 0;
  Vector packed_;
};

// TRegressionSpikeSlabSampler (Models/Glm/PosteriorSamplers/TRegressionSpikeSlabSampler.{hpp,cpp}; what BoomSpikeSlab's lm.spike
// builds for Student errors): TRegressionSampler with the coefficient draw replaced by the generic spike-and-slab steps
// (SpikeSlabSampler.cpp:40-216) at the model's residual variance -- sigsq enters them only as X'WX / sigsq, X'Wy / sigsq.
class TRegressionSpikeSlabSampler : public TRegressionSampler {
 public:
  TRegressionSpikeSlabSampler(TRegressionModel *model, const std::shared_ptr<MvnBase> &coefficient_slab_prior,
                              const std::shared_ptr<VariableSelectionPrior> &coefficient_spike_prior,
                              const std::shared_ptr<GammaModelBase> &siginv_prior, const std::shared_ptr<DoubleModel> &nu_prior,
                              RNG &seeding_rng = GlobalRng::rng);
  void draw() override;                 // .cpp:41-47
  double logpri() const override;       // .cpp:49-52
  void draw_model_indicators();
  void draw_included_coefficients();
  void allow_model_selection(bool allow) { core_.allow_model_selection(allow); }
  void limit_model_selection(int max_flips) { core_.limit_model_selection(max_flips); }
  double log_model_prob(const Selector &g);
  // Active-set statistics, as on BinomialLogitSpikeSlabSampler (off by default; p > 64, no all-reduce hook): per iteration the
  // device computes X'WX for the included columns + diagonal + X'Wy, the sweep fetches a column when it adds a variable; the
  // same chain as with the full matrix.  complete_data_sufficient_statistics() still answers with the full matrix (on demand).
  void set_active_set_statistics(bool tf) { active_.enabled = tf; }
  bool active_set_statistics() const { return active_.enabled; }
  int64_t active_set_columns_fetched() const { return active_.columns_fetched; }

 protected:
  double weighted_sum_of_squared_errors() override;
  void materialize_full_statistics() const override;

 private:
  const WeightedRegSuf &scaled_statistics();   // X'WX / sigsq, X'Wy / sigsq at the model's current sigsq
  bool impute_latent_data_active(const std::vector<int> &cols);
  StatView &scaled_view();                      // the active-set arrays divided by the model's current sigsq
  SpikeSlabCore core_;
  WeightedRegSuf scaled_;
  mutable ActiveSetState active_;
  std::unique_ptr<StatView> view_;
  Vector view_xty_;
  double view_sigsq_ = 0.0;
};

// The logit mixture every BinomialLogit sampler uploads; defaults to the 9-component table of
// Fruhwirth-Schnatter & Fruhwirth that the reference hard-codes (NormalMixtureApproximation.cpp:416-424);
// the BOOM adapter overwrites it from the live BinomialLogitDataImputer::mixture_approximation.
void set_logit_mixture(const Vector &mu, const Vector &sigma, const Vector &weights);

// ---- host linear algebra used by the small-state steps (exposed for tests) -----------------
// in-place lower Cholesky of a row-major n x n matrix; returns false when not positive definite
bool cholesky_lower(double *a, int n);
void lsolve_inplace(const double *L, int n, double *b);   // L x = b
void ltsolve_inplace(const double *L, int n, double *b);  // L' x = b
Vector rmvn_suf_mt(RNG &rng, const SpdMatrix &ivar, const Vector &ivar_mu);  // distributions/mvn.cpp:128-136

}  // namespace BOOM_B200
