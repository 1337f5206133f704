// boom_b200_adapter.cpp -- see boom_b200_adapter.hpp.
#include "boom_b200_adapter.hpp"

#include <map>
#include <set>

#include "../../include/boomgpu.h"
#include "Models/Glm/PosteriorSamplers/BinomialLogitDataImputer.hpp"
#include "Models/Glm/PosteriorSamplers/NormalMixtureApproximation.hpp"
#include "Models/Glm/PosteriorSamplers/poisson_mixture_approximation_table.hpp"
#include "cpputil/report_error.hpp"
#include "distributions.hpp"

namespace BOOM {
namespace B200 {

namespace {
// BOOM::SpdMatrix is column major and symmetric: the same bytes as the row-major BOOM_B200::SpdMatrix.
BOOM_B200::SpdMatrix to_host(const SpdMatrix &m) {
  BOOM_B200::SpdMatrix out((int)m.nrow());
  std::copy(m.data(), m.data() + (size_t)m.nrow() * m.nrow(), out.a.begin());
  return out;
}
BOOM_B200::Vector to_host(const Vector &v) { return BOOM_B200::Vector(v.begin(), v.end()); }

std::shared_ptr<BOOM_B200::VariableSelectionPrior> to_host(const VariableSelectionPrior &spike) {
  auto out = std::make_shared<BOOM_B200::VariableSelectionPrior>(to_host(spike.prior_inclusion_probabilities()));
  out->set_max_model_size((int)spike.max_model_size());
  return out;
}
}  // namespace

DeviceImputerBase::DeviceImputerBase(int xdim, RNG &seeding_rng) : PosteriorSampler(seeding_rng), suf_(xdim), xdim_(xdim) {}

DeviceImputerBase::~DeviceImputerBase() { boomgpu_destroy(ctx_); }

void DeviceImputerBase::check(int rc) const {
  if (rc) report_error(std::string("boomgpu: ") + boomgpu_last_error(ctx_));
}

void DeviceImputerBase::set_device(int device) {
  if (ctx_ && device != device_) { boomgpu_destroy(ctx_); ctx_ = nullptr; }
  device_ = device;
  stale_ = true;
}

void DeviceImputerBase::ensure_device_rows() {
  if (!ctx_) {
    if (boomgpu_create(&ctx_, device_)) report_error(std::string("boomgpu_create: ") + boomgpu_last_error(nullptr));
    stale_ = true;
    comm_dirty_ = !comm_id_.empty();
  }
  if (comm_dirty_) {
    check(boomgpu_comm_init(ctx_, comm_id_.data(), comm_ranks_, comm_rank_));
    comm_dirty_ = false;
  }
  if (stale_ || repack_each_time_) {
    pack_and_upload(ctx_);
    check(boomgpu_set_row_offset(ctx_, row_offset_));
    stale_ = false;
  }
}

namespace {
// SpikeSlabCore::find_posterior_mode works on a BOOM_B200::GlmModelBase: a proxy whose derivatives come from the adapter
class DerivsProxy : public BOOM_B200::GlmModelBase {
 public:
  typedef std::function<double(const BOOM_B200::Vector &, BOOM_B200::Vector *, BOOM_B200::SpdMatrix *)> Fn;
  DerivsProxy(int p, Fn f) : BOOM_B200::GlmModelBase(p, false), f_(std::move(f)) {}
  double log_likelihood_derivs(const BOOM_B200::Vector &b, BOOM_B200::Vector *g, BOOM_B200::SpdMatrix *h) override { return f_(b, g, h); }

 protected:
  void upload(BOOM_B200::DeviceData &) override {}

 private:
  Fn f_;
};
}  // namespace

bool DeviceImputerBase::find_mode(GlmCoefs &coef, const MvnBase &slab, const VariableSelectionPrior &spike, double epsilon,
                                  double *value) {
  ensure_device_rows();
  const int p = xdim_;
  DerivsProxy proxy(p, [&](const BOOM_B200::Vector &b, BOOM_B200::Vector *g, BOOM_B200::SpdMatrix *h) {
    double ll = 0;
    if (g) g->assign(p, 0.0);
    if (h && h->dim != p) *h = BOOM_B200::SpdMatrix(p);
    check(device_loglike_derivs(ctx_, b.data(), &ll, g ? g->data() : nullptr, h ? h->a.data() : nullptr));
    return ll;
  });
  BOOM_B200::Selector g(p, false);
  for (int i = 0; i < p; ++i) if (coef.inc()[i]) g.add(i);
  proxy.coef().set_inc(g);
  proxy.coef().set_Beta(to_host(coef.Beta()));
  auto hslab = std::make_shared<BOOM_B200::MvnModel>(to_host(slab.mu()), to_host(slab.siginv()), true);
  BOOM_B200::SpikeSlabCore core(hslab, to_host(spike), false);
  const bool ok = core.find_posterior_mode(proxy, epsilon, value);
  if (ok) coef.set_Beta(Vector(proxy.Beta().begin(), proxy.Beta().end()));
  return ok;
}

void DeviceImputerBase::impute_latent_data() {
  if (latent_data_fixed_) return;  // statistics are under external control (Imputer.hpp:282-299)
  ensure_device_rows();
  const int64_t len = boomgpu_suf_len(xdim_);
  packed_.resize((size_t)len);
  double *suf_dev = nullptr;
  check(boomgpu_suf_buffer(ctx_, &suf_dev));
  // fresh Philox key from the sampler's own stream (set_seed() repeats the chain).  Raw generator bits, not
  // BOOM::seed_rng: its llround(U * 2^64) overflows for half of all U and returns one fixed value for them.
  const uint64_t seed = rng().generator()();
  check(device_step(ctx_, current_beta().data(), seed, iteration_++, suf_dev));
  if (allreduce_) allreduce_(suf_dev, len);
  else check(boomgpu_allreduce(ctx_, suf_dev, len));   // no-op without a communicator
  check(boomgpu_download(ctx_, suf_dev, packed_.data(), len));
  const int p = xdim_;
  SpdMatrix xtx(p);
  std::copy(packed_.begin(), packed_.begin() + (size_t)p * p, xtx.data());
  Vector xty(packed_.begin() + (size_t)p * p, packed_.begin() + (size_t)p * p + p);
  const double *sc = packed_.data() + (size_t)p * p + p;
  suf_.reset(xtx, xty, sc[1], sc[0], sc[2], sc[3]);   // WeightedRegressionModel.cpp:148-157
}

void DeviceImputerBase::draw_beta_full_model(GlmCoefs &coef, const MvnBase &prior) {
  SpdMatrix ivar = prior.siginv() + suf_.xtx();
  Vector ivar_mu = suf_.xty() + prior.siginv() * prior.mu();
  coef.set_Beta(rmvn_suf_mt(rng(), ivar, ivar_mu));   // distributions/mvn.cpp:128-136, BOOM's own
}

void DeviceImputerBase::spike_slab_draw(GlmCoefs &coef, const MvnBase &slab, const VariableSelectionPrior &spike, bool select,
                                        int max_flips, bool fisher_yates) {
  // host small-state step on BOOM's objects through the shared evaluator of boom_b200/host
  auto hslab = std::make_shared<BOOM_B200::MvnModel>(to_host(slab.mu()), to_host(slab.siginv()), true);
  BOOM_B200::SpikeSlabCore core(hslab, to_host(spike), fisher_yates);
  core.allow_model_selection(select);
  core.limit_model_selection(max_flips);
  // from suf_ (not the last device result): externally driven statistics (fix_latent_data) must be honoured
  const int p = xdim_;
  packed_.resize((size_t)p * p + p + 4);
  const SpdMatrix xtx = suf_.xtx();   // by value in the reference (WeightedRegressionModel.cpp:192-195): once per draw here
  std::copy(xtx.data(), xtx.data() + (size_t)p * p, packed_.begin());
  const Vector xty = suf_.xty();
  std::copy(xty.begin(), xty.end(), packed_.begin() + (size_t)p * p);
  double *sc = packed_.data() + (size_t)p * p + p;
  sc[0] = suf_.n(); sc[1] = suf_.yty(); sc[2] = suf_.sumw(); sc[3] = suf_.sumlogw();
  BOOM_B200::WeightedRegSuf hsuf(xdim_);
  hsuf.reset(packed_.data(), xdim_);
  BOOM_B200::GlmCoefs hcoef(xdim_, false);
  BOOM_B200::Selector g(xdim_, false);
  for (int i = 0; i < xdim_; ++i) if (coef.inc()[i]) g.add(i);
  hcoef.set_inc(g);
  BOOM_B200::RNG local(rng().generator()());   // see impute_latent_data() on why not seed_rng()
  core.draw_model_indicators(local, hcoef, hsuf);
  core.draw_beta(local, hcoef, hsuf);
  std::vector<bool> bits(xdim_);
  for (int i = 0; i < xdim_; ++i) bits[i] = hcoef.inc()[i];
  coef.set_inc(Selector(bits));
  coef.set_Beta(Vector(hcoef.Beta().begin(), hcoef.Beta().end()));
}

double DeviceImputerBase::spike_slab_logpri(const GlmCoefs &coef, const MvnBase &slab, const VariableSelectionPrior &spike) const {
  const Selector &g(coef.inc());
  double ans = spike.logp(g);
  if (ans == negative_infinity()) return ans;
  if (g.nvars() > 0) {
    ans += dmvn(g.select(coef.Beta()), g.select(slab.mu()), g.select(slab.siginv()), true);
  }
  return ans;
}

// ---------------------------------------------------------------------------------------------
BinomialLogitAuxmixSampler::BinomialLogitAuxmixSampler(BinomialLogitModel *model, const Ptr<MvnBase> &prior, int clt_threshold,
                                                       RNG &seeding_rng)
    : DeviceImputerBase(model->xdim(), seeding_rng), model_(model), prior_(prior), clt_threshold_(clt_threshold) {
  if (prior_->dim() != model_->xdim()) report_error("Prior does not match model dimension.");
  model_->add_observer([this]() { this->mark_stale(); });
}

void BinomialLogitAuxmixSampler::pack_and_upload(boomgpu_ctx *ctx) {
  const std::vector<Ptr<BinomialRegressionData>> &data(model_->dat());
  const int64_t n = (int64_t)data.size();
  const int p = xdim_;
  std::vector<double> X((size_t)n * p), y(n), nt(n);
  for (int64_t i = 0; i < n; ++i) {
    const Vector &x(data[i]->x());
    std::copy(x.begin(), x.end(), X.begin() + (size_t)i * p);
    y[i] = data[i]->y();
    nt[i] = data[i]->n();
  }
  const NormalMixtureApproximation &mix(BinomialLogitDataImputer::mixture_approximation);   // BinomialLogitDataImputer.hpp:51
  check(boomgpu_set_logit_mixture(ctx, mix.dim(), mix.mu().data(), mix.sigma().data(), mix.weights().data()));
  check(boomgpu_upload_binomial(ctx, n, p, X.data(), p, y.data(), nt.data()));
}

int BinomialLogitAuxmixSampler::device_step(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration,
                                            double *suf_dev) {
  return boomgpu_logit_step_device(ctx, beta, clt_threshold_, seed, iteration, suf_dev);
}

int BinomialLogitAuxmixSampler::device_loglike_derivs(boomgpu_ctx *ctx, const double *beta, double *loglike, double *g, double *h) {
  return boomgpu_binomial_loglike_derivs(ctx, beta, model_->log_alpha(), loglike, g, h);   // BinomialLogitModel.cpp:168
}

void BinomialLogitAuxmixSampler::draw() {
  impute_latent_data();
  draw_params();
}
void BinomialLogitAuxmixSampler::draw_params() { draw_beta_full_model(model_->coef(), *prior_); }
double BinomialLogitAuxmixSampler::logpri() const { return prior_->logp(model_->Beta()); }
void BinomialLogitAuxmixSampler::update_complete_data_sufficient_statistics(double precision_weighted_sum, double total_precision,
                                                                            const Vector &x) {
  // the logit statistics are (sum, information): BinomialLogitAuxmixSampler.cpp:61-67
  suf_.add_data(x, total_precision > 0 ? precision_weighted_sum / total_precision : 0.0, total_precision);
}

BinomialLogitSpikeSlabSampler::BinomialLogitSpikeSlabSampler(BinomialLogitModel *model, const Ptr<MvnBase> &slab,
                                                             const Ptr<VariableSelectionPrior> &spike, int clt_threshold,
                                                             RNG &seeding_rng)
    : BinomialLogitAuxmixSampler(model, slab, clt_threshold, seeding_rng), slab_(slab), spike_(spike) {
  if ((int)spike_->potential_nvars() != model_->xdim()) report_error("Spike does not match model dimension.");
}
BinomialLogitSpikeSlabSampler *BinomialLogitSpikeSlabSampler::clone_to_new_host(Model *new_host) const {
  return new BinomialLogitSpikeSlabSampler(dynamic_cast<BinomialLogitModel *>(new_host), slab_->clone(), spike_->clone(),
                                           clt_threshold(), rng());
}
void BinomialLogitSpikeSlabSampler::draw() {   // BinomialLogitSpikeSlabSampler.cpp:50-54
  impute_latent_data();
  spike_slab_draw(model_->coef(), *slab_, *spike_, allow_model_selection_, max_flips_, false);
}
double BinomialLogitSpikeSlabSampler::logpri() const { return spike_slab_logpri(model_->coef(), *slab_, *spike_); }
void BinomialLogitSpikeSlabSampler::find_posterior_mode(double epsilon) {
  posterior_mode_found_ = find_mode(model_->coef(), *slab_, *spike_, epsilon, &log_posterior_at_mode_);
}
void BinomialLogitSpikeSlabSampler::set_spike(const Ptr<VariableSelectionPrior> &spike) {
  if ((int)spike->potential_nvars() != model_->xdim()) report_error("Spike does not match model dimension.");
  spike_ = spike;
}
void BinomialLogitSpikeSlabSampler::set_slab(const Ptr<MvnBase> &slab) {
  if (slab->dim() != model_->xdim()) report_error("Slab does not match model dimension.");
  slab_ = slab;
}

// ---------------------------------------------------------------------------------------------
PoissonRegressionAuxMixSampler::PoissonRegressionAuxMixSampler(PoissonRegressionModel *model, const Ptr<MvnBase> &prior, int,
                                                               RNG &seeding_rng)
    : DeviceImputerBase(model->xdim(), seeding_rng), model_(model), prior_(prior) {
  if (prior_->dim() != model_->xdim()) report_error("Prior does not match model dimension.");
  model_->add_observer([this]() { this->mark_stale(); });
}

void PoissonRegressionAuxMixSampler::pack_and_upload(boomgpu_ctx *ctx) {
  const std::vector<Ptr<PoissonRegressionData>> &data(model_->dat());
  const int64_t n = (int64_t)data.size();
  const int p = xdim_;
  std::vector<double> X((size_t)n * p), ex(n);
  std::vector<int64_t> y(n);
  std::set<int64_t> distinct;
  for (int64_t i = 0; i < n; ++i) {
    const Vector &x(data[i]->x());
    std::copy(x.begin(), x.end(), X.begin() + (size_t)i * p);
    y[i] = data[i]->y();
    ex[i] = data[i]->exposure();
    if (y[i] > 0) distinct.insert(y[i]);
  }
  // The table is materialised by the reference's own code for every count in the data (first touch of an
  // off-grid value may run its Powell re-fit on the host), then uploaded in its serialized layout.
  static NormalMixtureApproximationTable table = create_poisson_mixture_approximation_table();
  table.approximate(1);
  for (int64_t v : distinct) if (v < table.largest_index()) table.approximate((int)v);
  const Vector ser = table.serialize();   // [nu, K, w[K], sigma[K], mu[K]] ...  NormalMixtureApproximation.cpp:393-399,534-542
  std::map<int64_t, size_t> entries;      // last entry wins for duplicate keys, like the reference's lookup
  for (size_t i = 0; i < ser.size();) {
    const int K = (int)lround(ser[i + 1]);
    entries[(int64_t)llround(ser[i])] = i;
    i += 2 + 3 * (size_t)K;
  }
  std::vector<int64_t> nu;
  std::vector<int32_t> offset(1, 0);
  std::vector<double> w, mu, sigma;
  for (auto &e : entries) {
    const size_t i = e.second;
    const int K = (int)lround(ser[i + 1]);
    nu.push_back(e.first);
    for (int k = 0; k < K; ++k) { w.push_back(ser[i + 2 + k]); sigma.push_back(ser[i + 2 + K + k]); mu.push_back(ser[i + 2 + 2 * K + k]); }
    offset.push_back(offset.back() + K);
  }
  check(boomgpu_set_poisson_table(ctx, (int)nu.size(), nu.data(), offset.data(), w.data(), mu.data(), sigma.data(),
                                  table.largest_index()));
  check(boomgpu_upload_poisson(ctx, n, p, X.data(), p, y.data(), ex.data()));
}

int PoissonRegressionAuxMixSampler::device_step(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration,
                                                double *suf_dev) {
  return boomgpu_poisson_step_device(ctx, beta, seed, iteration, suf_dev);
}
int PoissonRegressionAuxMixSampler::device_loglike_derivs(boomgpu_ctx *ctx, const double *beta, double *loglike, double *g, double *h) {
  return boomgpu_poisson_loglike_derivs(ctx, beta, loglike, g, h);
}
void PoissonRegressionAuxMixSampler::draw() {
  impute_latent_data();
  draw_beta_given_complete_data();
}
void PoissonRegressionAuxMixSampler::draw_beta_given_complete_data() { draw_beta_full_model(model_->coef(), *prior_); }
double PoissonRegressionAuxMixSampler::logpri() const { return prior_->logp(model_->Beta()); }
void PoissonRegressionAuxMixSampler::update_complete_data_sufficient_statistics(double precision_weighted_sum,
                                                                                double total_precision, const Vector &x) {
  suf_.add_data(x, precision_weighted_sum / total_precision, total_precision);   // PoissonRegressionAuxMixSampler.cpp:153-158
}

PoissonRegressionSpikeSlabSampler::PoissonRegressionSpikeSlabSampler(PoissonRegressionModel *model, const Ptr<MvnBase> &slab,
                                                                     const Ptr<VariableSelectionPrior> &spike, int nthreads,
                                                                     RNG &seeding_rng)
    : PoissonRegressionAuxMixSampler(model, slab, nthreads, seeding_rng), slab_(slab), spike_(spike) {
  if ((int)spike_->potential_nvars() != model_->xdim()) report_error("Spike does not match model dimension.");
}
void PoissonRegressionSpikeSlabSampler::draw() {   // PoissonRegressionSpikeSlabSampler.cpp:55-59
  impute_latent_data();
  spike_slab_draw(model_->coef(), *slab_, *spike_, allow_model_selection_, max_flips_, true);
}
double PoissonRegressionSpikeSlabSampler::logpri() const { return spike_slab_logpri(model_->coef(), *slab_, *spike_); }
PoissonRegressionSpikeSlabSampler *PoissonRegressionSpikeSlabSampler::clone_to_new_host(Model *new_host) const {
  auto *s = new PoissonRegressionSpikeSlabSampler(dynamic_cast<PoissonRegressionModel *>(new_host), slab_->clone(), spike_->clone(),
                                                  1, rng());
  s->allow_model_selection(allow_model_selection_);
  s->limit_model_selection(max_flips_);
  return s;
}
void PoissonRegressionSpikeSlabSampler::find_posterior_mode(double epsilon) {
  find_mode(model_->coef(), *slab_, *spike_, epsilon, &log_posterior_at_mode_);
}

}  // namespace B200
}  // namespace BOOM
