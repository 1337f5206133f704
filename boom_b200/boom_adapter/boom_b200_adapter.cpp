// boom_b200_adapter.cpp -- see boom_b200_adapter.hpp.
#include "boom_b200_adapter.hpp"

#include <chrono>
#include <map>
#include <set>

#include "../../include/boomgpu.h"
#include "Models/Glm/PosteriorSamplers/BinomialLogitDataImputer.hpp"
#include "Models/Glm/PosteriorSamplers/NormalMixtureApproximation.hpp"
#include "Models/Glm/PosteriorSamplers/poisson_mixture_approximation_table.hpp"
#include "Models/MvnModel.hpp"
#include "Samplers/TIM.hpp"
#include "cpputil/math_utils.hpp"
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

// What the host steps need from the BOOM priors, converted once and kept until the priors change.  The reference reads
// slab_->siginv() / spike_ live in every draw; here the p x p precision is copied into the host layout only when (a) the
// sampler is handed another prior object (set_slab / set_spike), or (b) the slab is a BOOM::MvnModel whose parameters
// signal a change (observers on Mu_prm / Sigma_prm), or (c) the slab is some other MvnBase, whose changes cannot be
// observed -- then it is re-read every draw, as the reference does.
struct DeviceImputerBase::PriorCache {
  const MvnBase *slab = nullptr;
  const VariableSelectionPrior *spike = nullptr;
  uint64_t version = 0;
  bool fisher_yates = false, observable = false;
  Ptr<VectorParams> mu_prm;
  Ptr<SpdParams> sigma_prm;
  std::unique_ptr<BOOM_B200::SpikeSlabCore> core;
};

DeviceImputerBase::DeviceImputerBase(int xdim, RNG &seeding_rng) : PosteriorSampler(seeding_rng), hsuf_(xdim), xdim_(xdim) {}

DeviceImputerBase::~DeviceImputerBase() {
  if (prior_cache_ && prior_cache_->observable) {
    prior_cache_->mu_prm->remove_observer(observer_key());
    prior_cache_->sigma_prm->remove_observer(observer_key());
  }
  boomgpu_destroy(ctx_);
}

void DeviceImputerBase::check(int rc) const {
  if (rc) report_error(std::string("boomgpu: ") + boomgpu_last_error(ctx_));
}

void DeviceImputerBase::set_device(int device) {
  if (ctx_ && device != device_) { boomgpu_destroy(ctx_); ctx_ = nullptr; }
  device_ = device;
  stale_ = true;
}

void DeviceImputerBase::observe_rows(bool tf) {
  if (tf == observing_rows_) return;
  observe_row_objects(tf);
  observing_rows_ = tf;
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
    install_tables(ctx_);
    // rows go to the device a chunk at a time (about 32 MB of X): no second n x p copy on the host
    const int64_t n = row_count();
    const int p = xdim_;
    const int64_t chunk = std::max<int64_t>(1024, (int64_t)(32u << 20) / (8 * (int64_t)p));
    std::vector<double> X((size_t)std::min(n, chunk) * p), aux((size_t)std::min(n, chunk));
    std::vector<int64_t> y((size_t)std::min(n, chunk));   // 8 bytes per row: doubles (binomial) or int64 (poisson)
    check(boomgpu_upload_begin(ctx_, row_kind(), n, p));
    for (int64_t row0 = 0; row0 < n; row0 += chunk) {
      const int64_t rows = std::min(chunk, n - row0);
      pack_rows(row0, rows, X.data(), y.data(), aux.data());
      check(boomgpu_upload_rows(ctx_, row0, rows, X.data(), p, y.data(), aux.data()));
    }
    check(boomgpu_upload_end(ctx_));
    check(boomgpu_set_row_offset(ctx_, row_offset_));
    stale_ = false;
    if (observing_rows_) observe_row_objects(true);   // rows added since the last pack are observed too
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

struct Stopwatch {
  double &acc;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  explicit Stopwatch(double &a) : acc(a) {}
  ~Stopwatch() { acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};
}  // namespace

// log likelihood, gradient and Hessian over ALL ranks' rows: natively all-reduced inside the C ABI when a communicator is
// attached; through the caller's hook otherwise (packed [-H | g | {., ll, ., .}] on the device, like a Gibbs step)
double DeviceImputerBase::loglike_derivs(const BOOM_B200::Vector &b, BOOM_B200::Vector *g, BOOM_B200::SpdMatrix *h) {
  const int p = xdim_;
  double ll = 0;
  if (g) g->assign(p, 0.0);
  if (h && h->dim != p) *h = BOOM_B200::SpdMatrix(p);
  if (!allreduce_) {
    check(device_loglike_derivs(ctx_, b.data(), &ll, g ? g->data() : nullptr, h ? h->a.data() : nullptr));
    return ll;
  }
  const int64_t len = boomgpu_suf_len(p);
  packed_.resize((size_t)len);
  double *suf_dev = nullptr;
  check(boomgpu_suf_buffer(ctx_, &suf_dev));
  check(device_loglike_derivs_device(ctx_, b.data(), suf_dev));
  allreduce_(suf_dev, len);
  check(boomgpu_download(ctx_, suf_dev, packed_.data(), len));
  const size_t mat = (size_t)p * p;
  if (g) g->assign(packed_.begin() + mat, packed_.begin() + mat + p);
  if (h) for (size_t e = 0; e < mat; ++e) h->a[e] = -packed_[e];
  return packed_[mat + p + 1];
}

double DeviceImputerBase::included_loglike_derivs(const Selector &inc, const Vector &beta_included, Vector *g, Matrix *h) {
  Stopwatch sw(secs_device_);
  ensure_device_rows();
  const int k = (int)inc.nvars();
  if ((int)beta_included.size() != k) report_error("included_loglike_derivs: beta does not match the inclusion pattern");
  std::vector<int32_t> cols;
  for (int j = 0; j < (int)inc.nvars_possible(); ++j) if (inc[j]) cols.push_back(j);
  check(boomgpu_select_columns(ctx_, cols.data(), k));   // gathers X_gamma only when the pattern (or the data) changed
  double ll = 0;
  std::vector<double> hbuf(h ? (size_t)k * k : 0);
  if (g) g->resize(k);
  if (!allreduce_) {
    check(device_loglike_derivs_selected(ctx_, beta_included.data(), &ll, g ? g->data() : nullptr, h ? hbuf.data() : nullptr));
  } else {
    const int64_t len = boomgpu_suf_len(k);
    packed_.resize((size_t)len);
    double *suf_dev = nullptr;
    check(boomgpu_suf_buffer(ctx_, &suf_dev));   // sized for the full p: large enough for k <= p
    check(device_loglike_derivs_selected_device(ctx_, beta_included.data(), suf_dev));
    allreduce_(suf_dev, len);
    check(boomgpu_download(ctx_, suf_dev, packed_.data(), len));
    const size_t mat = (size_t)k * k;
    ll = packed_[mat + k + 1];
    if (g) std::copy(packed_.begin() + mat, packed_.begin() + mat + k, g->begin());
    if (h) for (size_t e = 0; e < mat; ++e) hbuf[e] = -packed_[e];
  }
  if (h) {
    *h = Matrix(k, k);
    std::copy(hbuf.begin(), hbuf.end(), h->data());   // symmetric: row major == column major
  }
  return ll;
}

bool DeviceImputerBase::find_mode(GlmCoefs &coef, const Ptr<MvnBase> &slab, const Ptr<VariableSelectionPrior> &spike, double epsilon,
                                  double *value) {
  ensure_device_rows();
  const int p = xdim_;
  DerivsProxy proxy(p, [&](const BOOM_B200::Vector &b, BOOM_B200::Vector *g, BOOM_B200::SpdMatrix *h) { return loglike_derivs(b, g, h); });
  BOOM_B200::Selector g(p, false);
  for (int i = 0; i < p; ++i) if (coef.inc()[i]) g.add(i);
  proxy.coef().set_inc(g);
  proxy.coef().set_Beta(to_host(coef.Beta()));
  const bool ok = core(slab, spike, false).find_posterior_mode(proxy, epsilon, value);
  if (ok) coef.set_Beta(Vector(proxy.Beta().begin(), proxy.Beta().end()));
  return ok;
}

void DeviceImputerBase::impute_latent_data() {
  if (latent_data_fixed_) return;  // statistics are under external control (Imputer.hpp:282-299)
  active_.valid = false;
  view_.reset();
  Stopwatch sw(secs_device_);
  ensure_device_rows();
  const int p = xdim_;
  // fresh Philox key from the sampler's own stream (set_seed() repeats the chain).  Raw generator bits, not
  // BOOM::seed_rng: its llround(U * 2^64) overflows for half of all U and returns one fixed value for them.
  const uint64_t seed = rng().generator()();
  if (allreduce_) {   // caller-supplied all-reduce: the packed statistics stay on the device for the hook
    const int64_t len = boomgpu_suf_len(p);
    packed_.resize((size_t)len);
    double *suf_dev = nullptr;
    check(boomgpu_suf_buffer(ctx_, &suf_dev));
    check(device_step(ctx_, current_beta().data(), seed, iteration_++, suf_dev));
    allreduce_(suf_dev, len);
    check(boomgpu_download(ctx_, suf_dev, packed_.data(), len));
    hsuf_.reset(packed_.data(), p);
  } else {            // straight into the statistics object (native all-reduce inside when a communicator is attached)
    double sc[4] = {0, 0, 0, 0};
    check(device_step_sync(ctx_, current_beta().data(), seed, iteration_++, hsuf_.xtx_storage(p), hsuf_.xty_storage(), sc));
    hsuf_.set_scalars(sc[0], sc[1], sc[2], sc[3]);
  }
  statistics_changed();
}

int DeviceImputerBase::device_step_active(boomgpu_ctx *, const double *, uint64_t, uint64_t, const int32_t *, int, double *, double *,
                                          double *, double *) {
  return -1;   // this sampler has no active-set form
}

bool DeviceImputerBase::impute_latent_data_active(const Selector &inc) {
  const int p = xdim_;
  if (!active_.enabled || latent_data_fixed_ || p <= 64 || allreduce_ || inc.nvars() > 128) return false;
  Stopwatch sw(secs_device_);
  ensure_device_rows();
  std::vector<int32_t> cols;
  for (int j = 0; j < p; ++j) if (inc[j]) cols.push_back(j);
  if (cols.empty()) cols.push_back(0);
  const int k = (int)cols.size();
  active_.cols.assign(cols.begin(), cols.end());
  active_.G.resize((size_t)p * k); active_.diag.resize(p); active_.xty.resize(p);
  const uint64_t seed = rng().generator()();
  const int rc = device_step_active(ctx_, current_beta().data(), seed, iteration_, cols.data(), k, active_.G.data(), active_.diag.data(),
                                    active_.xty.data(), active_.scalars);
  if (rc == -1) return false;
  ++iteration_;
  check(rc);
  active_.valid = true;
  boomgpu_ctx *ctx = ctx_;
  BOOM_B200::ActiveSetState *st = &active_;
  view_.reset(new BOOM_B200::StatView(p, active_.cols, active_.G, active_.diag, active_.xty, [this, ctx, st](int j, double *out) {
    check(boomgpu_weighted_column(ctx, j, out));
    ++st->columns_fetched;
  }));
  statistics_changed();
  return true;
}

void DeviceImputerBase::materialize_full_statistics() const {
  if (!active_.valid) return;
  if (int rc = boomgpu_full_statistics(ctx_, hsuf_.xtx_storage(xdim_), hsuf_.xty_storage()))
    report_error(std::string("boomgpu: ") + boomgpu_last_error(ctx_));
  hsuf_.set_scalars(active_.scalars[0], active_.scalars[1], active_.scalars[2], active_.scalars[3]);
  active_.valid = false;
}

void DeviceImputerBase::draw_beta_full_model(GlmCoefs &coef, const MvnBase &prior) {
  materialize_full_statistics();
  Stopwatch sw(secs_host_);
  // ivar = Omega^-1 + X'WX, ivar_mu = X'Wz + Omega^-1 mu_0, then BOOM's own rmvn_suf_mt (distributions/mvn.cpp:128-136).
  // BOOM::SpdMatrix is column major and symmetric: the same bytes as the row-major host matrix.
  const int p = xdim_;
  SpdMatrix ivar = prior.siginv();
  Vector ivar_mu = prior.siginv() * prior.mu();
  const double *xtx = hsuf_.xtx().a.data();
  double *iv = ivar.data();
  for (size_t e = 0; e < (size_t)p * p; ++e) iv[e] += xtx[e];
  for (int i = 0; i < p; ++i) ivar_mu[i] += hsuf_.xty()[i];
  coef.set_Beta(rmvn_suf_mt(rng(), ivar, ivar_mu));
}

const BOOM_B200::SpikeSlabCore &DeviceImputerBase::core(const Ptr<MvnBase> &slab, const Ptr<VariableSelectionPrior> &spike,
                                                      bool fisher_yates) const {
  if (!prior_cache_) prior_cache_.reset(new PriorCache);
  PriorCache &c(*prior_cache_);
  const bool same = c.core && c.slab == slab.get() && c.spike == spike.get() && c.fisher_yates == fisher_yates && c.observable &&
                    c.version == prior_version_;
  if (same) return *c.core;
  DeviceImputerBase *self = const_cast<DeviceImputerBase *>(this);
  if (c.observable && c.slab != slab.get()) {
    c.mu_prm->remove_observer(self->observer_key());
    c.sigma_prm->remove_observer(self->observer_key());
    c.observable = false;
  }
  if (!c.observable) {
    if (MvnModel *mvn = dynamic_cast<MvnModel *>(slab.get())) {   // its parameters signal when they are set
      c.mu_prm = mvn->Mu_prm();
      c.sigma_prm = mvn->Sigma_prm();
      c.mu_prm->add_observer(self->observer_key(), [self]() { self->priors_changed(); });
      c.sigma_prm->add_observer(self->observer_key(), [self]() { self->priors_changed(); });
      c.observable = true;
    }
  }
  auto hslab = std::make_shared<BOOM_B200::MvnModel>(to_host(slab->mu()), to_host(slab->siginv()), true);
  c.core.reset(new BOOM_B200::SpikeSlabCore(hslab, to_host(*spike), fisher_yates));
  c.slab = slab.get(); c.spike = spike.get(); c.fisher_yates = fisher_yates; c.version = prior_version_;
  return *c.core;
}

namespace {
BOOM_B200::GlmCoefs host_coefs(const GlmCoefs &coef, int p, bool with_beta) {
  BOOM_B200::GlmCoefs h(p, false);
  BOOM_B200::Selector g(p, false);
  for (int i = 0; i < p; ++i) if (coef.inc()[i]) g.add(i);
  h.set_inc(g);
  if (with_beta) h.set_Beta(to_host(coef.Beta()));
  return h;
}
void write_back(GlmCoefs &coef, const BOOM_B200::GlmCoefs &h, int p, bool beta) {
  std::vector<bool> bits(p);
  for (int i = 0; i < p; ++i) bits[i] = h.inc()[i];
  coef.set_inc(Selector(bits));
  if (beta) coef.set_Beta(Vector(h.Beta().begin(), h.Beta().end()));
}
}  // namespace

void DeviceImputerBase::sweep_indicators(GlmCoefs &coef, const BOOM_B200::SpikeSlabCore &c) {
  Stopwatch sw(secs_host_);
  BOOM_B200::GlmCoefs h = host_coefs(coef, xdim_, true);
  BOOM_B200::RNG local(rng().generator()());   // see impute_latent_data() on why not seed_rng()
  if (active_.valid && view_) c.draw_model_indicators(local, h, *view_);   // active-set form: fetches a column per accepted add
  else c.draw_model_indicators(local, h, hsuf_);   // reads the statistics in place: externally driven ones (fix_latent_data) too
  write_back(coef, h, xdim_, false);
}

void DeviceImputerBase::draw_included_beta(GlmCoefs &coef, const BOOM_B200::SpikeSlabCore &c) {
  Stopwatch sw(secs_host_);
  BOOM_B200::GlmCoefs h = host_coefs(coef, xdim_, false);
  BOOM_B200::RNG local(rng().generator()());
  if (active_.valid && view_) c.draw_beta(local, h, *view_);
  else c.draw_beta(local, h, hsuf_);
  write_back(coef, h, xdim_, true);
}

double DeviceImputerBase::model_log_prob(const Selector &g, const BOOM_B200::SpikeSlabCore &c) const {
  materialize_full_statistics();
  BOOM_B200::Selector h((int)g.nvars_possible(), false);
  for (int i = 0; i < (int)g.nvars_possible(); ++i) if (g[i]) h.add(i);
  return c.log_model_prob(h, hsuf_);
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

// ---- the reference's statistics types, filled in bulk ------------------------------------------------------------
// BinomialLogit::SufficientStatistics (BinomialLogitAuxmixSampler.hpp:39-67) keeps xtx_, xty_, sym_ and sample_size_ private
// and offers only per-observation update(): there is no way to hand it a finished p x p matrix through its interface.  The
// drop-in must nevertheless return THAT type from suf().  The members are reached through explicit template instantiation,
// whose arguments are exempt from access checking ([temp.explicit]): standard C++, no change to the BOOM sources, and a
// compile error -- not silent misbehaviour -- should BOOM ever rename a member.
namespace {
template <class Tag, typename Tag::type Member>
struct MemberAccess { friend typename Tag::type member_pointer(Tag) { return Member; } };
struct SufXtx { typedef SpdMatrix BinomialLogit::SufficientStatistics::*type; friend type member_pointer(SufXtx); };
struct SufXty { typedef Vector BinomialLogit::SufficientStatistics::*type; friend type member_pointer(SufXty); };
struct SufSym { typedef bool BinomialLogit::SufficientStatistics::*type; friend type member_pointer(SufSym); };
struct SufSize { typedef int BinomialLogit::SufficientStatistics::*type; friend type member_pointer(SufSize); };
template struct MemberAccess<SufXtx, &BinomialLogit::SufficientStatistics::xtx_>;
template struct MemberAccess<SufXty, &BinomialLogit::SufficientStatistics::xty_>;
template struct MemberAccess<SufSym, &BinomialLogit::SufficientStatistics::sym_>;
template struct MemberAccess<SufSize, &BinomialLogit::SufficientStatistics::sample_size_>;
}  // namespace

// ---------------------------------------------------------------------------------------------
BinomialLogitAuxmixSampler::BinomialLogitAuxmixSampler(BinomialLogitModel *model, const Ptr<MvnBase> &prior, int clt_threshold,
                                                       RNG &seeding_rng)
    : DeviceImputerBase(model->xdim(), seeding_rng), model_(model), prior_(prior), clt_threshold_(clt_threshold), suf_(model->xdim()) {
  if (prior_->dim() != model_->xdim()) report_error("Prior does not match model dimension.");
  model_->add_observer([this]() { this->mark_stale(); });
}

const BinomialLogit::SufficientStatistics &BinomialLogitAuxmixSampler::suf() const {
  materialize_full_statistics();
  if (!suf_synced_) {
    const int p = xdim_;
    SpdMatrix &xtx(suf_.*member_pointer(SufXtx()));
    Vector &xty(suf_.*member_pointer(SufXty()));
    if ((int)xtx.nrow() != p) { xtx = SpdMatrix(p); xty = Vector(p); }
    std::copy(hsuf_.xtx().a.begin(), hsuf_.xtx().a.end(), xtx.data());   // symmetric: row major == column major
    std::copy(hsuf_.xty().begin(), hsuf_.xty().end(), xty.begin());
    suf_.*member_pointer(SufSym()) = true;                                // both triangles are filled
    suf_.*member_pointer(SufSize()) = (int)hsuf_.sample_size();
    suf_synced_ = true;
  }
  return suf_;
}

void BinomialLogitAuxmixSampler::install_tables(boomgpu_ctx *ctx) {
  const NormalMixtureApproximation &mix(BinomialLogitDataImputer::mixture_approximation);   // BinomialLogitDataImputer.hpp:51
  check(boomgpu_set_logit_mixture(ctx, mix.dim(), mix.mu().data(), mix.sigma().data(), mix.weights().data()));
}

void BinomialLogitAuxmixSampler::pack_rows(int64_t row0, int64_t nrows, double *X, void *y, double *aux) const {
  const std::vector<Ptr<BinomialRegressionData>> &data(model_->dat());
  const int p = xdim_;
  double *yd = static_cast<double *>(y);
  for (int64_t i = 0; i < nrows; ++i) {
    const BinomialRegressionData &d(*data[row0 + i]);
    const Vector &x(d.x());
    std::copy(x.begin(), x.end(), X + (size_t)i * p);
    yd[i] = d.y();
    aux[i] = d.n();
  }
}

void BinomialLogitAuxmixSampler::observe_row_objects(bool tf) {
  for (const Ptr<BinomialRegressionData> &d : model_->dat()) {
    d->remove_observer(observer_key());
    d->Xptr()->remove_observer(observer_key());
    if (tf) {
      d->add_observer(observer_key(), [this]() { this->mark_stale(); });            // set_y / set_n signal on the observation
      d->Xptr()->add_observer(observer_key(), [this]() { this->mark_stale(); });    // set_x signals on its VectorData
    }
  }
}

int BinomialLogitAuxmixSampler::device_step(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration,
                                            double *suf_dev) {
  return boomgpu_logit_step_device(ctx, beta, clt_threshold_, seed, iteration, suf_dev);
}

int BinomialLogitAuxmixSampler::device_step_sync(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, double *xtx,
                                                 double *xty, double scalars[4]) {
  int64_t ss = 0;
  const int rc = boomgpu_logit_step(ctx, beta, clt_threshold_, seed, iteration, xtx, xty, &ss);
  scalars[0] = (double)ss; scalars[1] = scalars[2] = scalars[3] = 0.0;
  return rc;
}

int BinomialLogitAuxmixSampler::device_step_active(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration,
                                                   const int32_t *cols, int k, double *G, double *diag, double *xty, double scalars[4]) {
  int64_t ss = 0;
  const int rc = boomgpu_logit_step_active(ctx, beta, clt_threshold_, seed, iteration, cols, k, G, diag, xty, &ss);
  scalars[0] = (double)ss; scalars[1] = scalars[2] = scalars[3] = 0.0;
  return rc;
}

int BinomialLogitAuxmixSampler::device_loglike_derivs(boomgpu_ctx *ctx, const double *beta, double *loglike, double *g, double *h) {
  return boomgpu_binomial_loglike_derivs(ctx, beta, model_->log_alpha(), loglike, g, h);   // BinomialLogitModel.cpp:168
}
int BinomialLogitAuxmixSampler::device_loglike_derivs_device(boomgpu_ctx *ctx, const double *beta, double *suf_dev) {
  return boomgpu_binomial_loglike_derivs_device(ctx, beta, model_->log_alpha(), suf_dev);
}
int BinomialLogitAuxmixSampler::device_loglike_derivs_selected(boomgpu_ctx *ctx, const double *beta, double *loglike, double *g, double *h) {
  return boomgpu_binomial_loglike_derivs_selected(ctx, beta, model_->log_alpha(), loglike, g, h);
}
int BinomialLogitAuxmixSampler::device_loglike_derivs_selected_device(boomgpu_ctx *ctx, const double *beta, double *suf_dev) {
  return boomgpu_binomial_loglike_derivs_selected_device(ctx, beta, model_->log_alpha(), suf_dev);
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
  hsuf_.update(to_host(x), precision_weighted_sum, total_precision);
  statistics_changed();
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
  if (!impute_latent_data_active(model_->coef().inc())) impute_latent_data();
  if (allow_model_selection_) draw_model_indicators();
  draw_beta();
}
void BinomialLogitSpikeSlabSampler::draw_model_indicators() {
  BOOM_B200::SpikeSlabCore c(core(slab_, spike_, false));   // a copy: the sweep limits are per call
  c.allow_model_selection(true);
  c.limit_model_selection(max_flips_);
  sweep_indicators(model_->coef(), c);
}
void BinomialLogitSpikeSlabSampler::draw_beta() { draw_included_beta(model_->coef(), core(slab_, spike_, false)); }
double BinomialLogitSpikeSlabSampler::log_model_prob(const Selector &gamma) const {
  return model_log_prob(gamma, core(slab_, spike_, false));
}
double BinomialLogitSpikeSlabSampler::logpri() const { return spike_slab_logpri(model_->coef(), *slab_, *spike_); }
void BinomialLogitSpikeSlabSampler::find_posterior_mode(double epsilon) {
  posterior_mode_found_ = find_mode(model_->coef(), slab_, spike_, epsilon, &log_posterior_at_mode_);
}
void BinomialLogitSpikeSlabSampler::set_spike(const Ptr<VariableSelectionPrior> &spike) {
  if ((int)spike->potential_nvars() != model_->xdim()) report_error("Spike does not match model dimension.");
  spike_ = spike;
  priors_changed();
}
void BinomialLogitSpikeSlabSampler::set_slab(const Ptr<MvnBase> &slab) {
  if (slab->dim() != model_->xdim()) report_error("Slab does not match model dimension.");
  slab_ = slab;
  priors_changed();
}

// ---------------------------------------------------------------------------------------------
BinomialLogitLogPostChunk::BinomialLogitLogPostChunk(BinomialLogitCompositeSpikeSlabSampler *sampler, int chunk_size, int chunk_number)
    : sampler_(sampler), start_(chunk_size * chunk_number) {
  const int nvars = (int)sampler_->m_->coef().nvars();
  chunk_size_ = std::min(chunk_size, nvars - start_);
}

double BinomialLogitLogPostChunk::operator()(const Vector &beta_chunk) const {
  Vector g;
  Matrix h;
  return (*this)(beta_chunk, g, h, 0);
}

// .cpp:34-74: log prior of ALL included coefficients with the chunk replaced + log likelihood; derivatives w.r.t. the chunk
double BinomialLogitLogPostChunk::operator()(const Vector &beta_chunk, Vector &grad, Matrix &hess, int nd) const {
  BinomialLogitModel *m = sampler_->m_;
  Vector nonzero_beta = m->included_coefficients();
  VectorView(nonzero_beta, start_, chunk_size_) = beta_chunk;
  const Selector &inc(m->coef().inc());
  const SpdMatrix siginv(inc.select(sampler_->pri_->siginv()));
  const Vector mu(inc.select(sampler_->pri_->mu()));
  double ans = dmvn(nonzero_beta, mu, siginv, 0.0, true);
  Selector chunk_selector(nonzero_beta.size(), false);
  for (int i = start_; i < start_ + chunk_size_; ++i) chunk_selector.add(i);
  Vector g_full;
  Matrix h_full;
  ans += sampler_->chunk_loglike(nonzero_beta, nd > 0 ? &g_full : nullptr, nd > 1 ? &h_full : nullptr);
  if (nd > 0) {
    grad = -1 * chunk_selector.select(siginv * (nonzero_beta - mu));
    grad += chunk_selector.select(g_full);
    if (nd > 1) {
      hess = chunk_selector.select(siginv);
      hess *= -1;
      for (int i = 0; i < chunk_size_; ++i)
        for (int j = 0; j < chunk_size_; ++j) hess(i, j) += h_full(start_ + i, start_ + j);
    }
  }
  return ans;
}

typedef BinomialLogitCompositeSpikeSlabSampler BLCSSS;
BLCSSS::BinomialLogitCompositeSpikeSlabSampler(BinomialLogitModel *model, const Ptr<MvnBase> &prior,
                                               const Ptr<VariableSelectionPrior> &vpri, int clt_threshold, double tdf,
                                               int max_tim_chunk_size, int max_rwm_chunk_size, double rwm_variance_scale_factor,
                                               RNG &seeding_rng)
    : BinomialLogitSpikeSlabSampler(model, prior, vpri, clt_threshold, seeding_rng), m_(model), pri_(prior), tdf_(tdf),
      max_tim_chunk_size_(max_tim_chunk_size), max_rwm_chunk_size_(max_rwm_chunk_size),
      rwm_variance_scale_factor_(rwm_variance_scale_factor) {
  set_sampler_weights(1.0, 1.0, 1.0);
}

void BLCSSS::draw() {   // .cpp:93-116
  enum SamplingMethod { DATA_AUGMENTATION = 0, RWM = 1, TIM_METHOD = 2 };
  const SamplingMethod method = SamplingMethod(rmulti_mt(rng(), sampler_weights_));
  switch (method) {
    case DATA_AUGMENTATION: {
      MoveTimer timer = move_accounting_.start_time("auxmix");
      BinomialLogitSpikeSlabSampler::draw();
      move_accounting_.record_acceptance("auxmix");
      break;
    }
    case RWM: {
      MoveTimer timer = move_accounting_.start_time("rwm (total time)");
      rwm_draw();
      break;
    }
    case TIM_METHOD: {
      MoveTimer timer = move_accounting_.start_time("TIM (total time)");
      tim_draw();
      break;
    }
    default:
      report_error("Unknown method in BinomialLogitSpikeSlabSampler::draw.");
  }
}

void BLCSSS::rwm_draw() {
  if (m_->coef().nvars() == 0) return;
  const int total_number_of_chunks = compute_number_of_chunks(max_rwm_chunk_size_);
  for (int chunk = 0; chunk < total_number_of_chunks; ++chunk) rwm_draw_chunk(chunk);
}

// .cpp:127-184.  The proposal precision is the chunk of the negative log-posterior Hessian at the current beta (prior
// precision + sum_i n_i p_i q_i x x'; the reference's loop weights by p_i q_i, which is the same thing for Bernoulli rows),
// from the same device pass that gives the current log posterior.
void BLCSSS::rwm_draw_chunk(int chunk) {
  const Selector &inc(m_->coef().inc());
  const int nvars = (int)inc.nvars();
  Vector full_nonzero_beta = m_->included_coefficients();
  const Vector mu(inc.select(pri_->mu()));
  const SpdMatrix siginv(inc.select(pri_->siginv()));
  double original_logpost = dmvn(full_nonzero_beta, mu, siginv, 0, true);
  const int full_chunk_size = compute_chunk_size(max_rwm_chunk_size_);
  const int chunk_start = chunk * full_chunk_size;
  const int this_chunk_size = std::min(nvars - chunk_start, full_chunk_size);
  Selector chunk_selector(nvars, false);
  for (int i = chunk_start; i < chunk_start + this_chunk_size; ++i) chunk_selector.add(i);
  SpdMatrix proposal_ivar = chunk_selector.select(siginv);
  Matrix h_full;
  original_logpost += chunk_loglike(full_nonzero_beta, nullptr, &h_full);
  for (int i = 0; i < this_chunk_size; ++i)
    for (int j = 0; j < this_chunk_size; ++j) proposal_ivar(i, j) -= h_full(chunk_start + i, chunk_start + j);
  VectorView beta_chunk(full_nonzero_beta, chunk_start, this_chunk_size);
  if (tdf_ > 0) beta_chunk = rmvt_ivar_mt(rng(), beta_chunk, proposal_ivar / rwm_variance_scale_factor_, tdf_);
  else beta_chunk = rmvn_ivar_mt(rng(), beta_chunk, proposal_ivar / rwm_variance_scale_factor_);
  double logpost = dmvn(full_nonzero_beta, mu, siginv, 0, true);
  logpost += chunk_loglike(full_nonzero_beta, nullptr, nullptr);
  const double log_alpha = logpost - original_logpost;
  const double logu = log(runif_mt(rng()));
  if (logu < log_alpha) {
    m_->set_included_coefficients(full_nonzero_beta);
    move_accounting_.record_acceptance("rwm_chunk");
  } else {
    move_accounting_.record_rejection("rwm_chunk");
  }
}

void BLCSSS::tim_draw() {   // .cpp:187-221
  const int nvars = (int)m_->coef().nvars();
  if (nvars == 0) return;
  const int chunk_size = compute_chunk_size(max_tim_chunk_size_);
  const int number_of_chunks = compute_number_of_chunks(max_tim_chunk_size_);
  for (int chunk = 0; chunk < number_of_chunks; ++chunk) {
    clock_t mode_start = clock();
    TIM tim_sampler(log_posterior(chunk, max_tim_chunk_size_), tdf_, &rng());
    Vector beta = m_->included_coefficients();
    const int start = chunk_size * chunk;
    VectorView beta_chunk(beta, start, std::min(nvars - start, chunk_size));
    const bool ok = tim_sampler.locate_mode(beta_chunk);
    move_accounting_.stop_time("tim mode finding", mode_start);
    if (ok) {
      move_accounting_.record_acceptance("tim mode finding");
      tim_sampler.fix_mode(true);
      MoveTimer timer = move_accounting_.start_time("TIM chunk");
      beta_chunk = tim_sampler.draw(beta_chunk);
      m_->set_included_coefficients(beta);
      if (tim_sampler.last_draw_was_accepted()) move_accounting_.record_acceptance("TIM chunk");
      else move_accounting_.record_rejection("TIM chunk");
    } else {
      move_accounting_.record_rejection("tim mode finding");
      rwm_draw_chunk(chunk);
    }
  }
}

BinomialLogitLogPostChunk BLCSSS::log_posterior(int chunk, int max_chunk_size) const {
  return BinomialLogitLogPostChunk(const_cast<BLCSSS *>(this), compute_chunk_size(max_chunk_size), chunk);
}

void BLCSSS::set_sampler_weights(double da_weight, double rwm_weight, double tim_weight) {
  if (da_weight < 0 || rwm_weight < 0 || tim_weight < 0) report_error("All three weights must be non-negative.");
  if (da_weight <= 0 && rwm_weight <= 0 && tim_weight <= 0) report_error("At least one weight must be positive.");
  sampler_weights_.resize(3);
  sampler_weights_[0] = da_weight;
  sampler_weights_[1] = rwm_weight;
  sampler_weights_[2] = tim_weight;
  sampler_weights_ /= sum(sampler_weights_);
}

int BLCSSS::compute_chunk_size(int max_chunk_size) const {
  const int nvars = (int)m_->coef().nvars();
  if (max_chunk_size <= 0) return nvars;
  const int number_of_full_chunks = nvars / max_chunk_size;
  const bool has_partial_chunk = number_of_full_chunks * max_chunk_size < nvars;
  const int total_chunks = number_of_full_chunks + has_partial_chunk;
  return divide_rounding_up(nvars, total_chunks);
}

int BLCSSS::compute_number_of_chunks(int max_chunk_size) const {
  if (max_chunk_size <= 0) return 1;
  const int nvars = (int)m_->coef().nvars();
  const int number_of_full_chunks = nvars / max_chunk_size;
  const bool has_partial_chunk = number_of_full_chunks * max_chunk_size < nvars;
  return number_of_full_chunks + has_partial_chunk;
}

std::ostream &BLCSSS::time_report(std::ostream &out) const {
  out << move_accounting_.to_matrix();
  return out;
}

// ---------------------------------------------------------------------------------------------
BinomialProbitSpikeSlabSampler::BinomialProbitSpikeSlabSampler(BinomialProbitModel *model, const Ptr<MvnBase> &slab,
                                                               const Ptr<VariableSelectionPrior> &spike, int clt_threshold,
                                                               RNG &seeding_rng)
    : DeviceImputerBase(model->xdim(), seeding_rng), model_(model), slab_(slab), spike_(spike), clt_threshold_(clt_threshold) {
  if (slab_->dim() != model_->xdim() || (int)spike_->potential_nvars() != model_->xdim())
    report_error("Prior does not match model dimension.");
  model_->add_observer([this]() { this->mark_stale(); });
}
void BinomialProbitSpikeSlabSampler::pack_rows(int64_t row0, int64_t nrows, double *X, void *y, double *aux) const {
  const std::vector<Ptr<BinomialRegressionData>> &data(model_->dat());
  const int p = xdim_;
  double *yd = static_cast<double *>(y);
  for (int64_t i = 0; i < nrows; ++i) {
    const BinomialRegressionData &d(*data[row0 + i]);
    const Vector &x(d.x());
    std::copy(x.begin(), x.end(), X + (size_t)i * p);
    yd[i] = d.y();
    aux[i] = d.n();
  }
}
void BinomialProbitSpikeSlabSampler::observe_row_objects(bool tf) {
  for (const Ptr<BinomialRegressionData> &d : model_->dat()) {
    d->remove_observer(observer_key());
    d->Xptr()->remove_observer(observer_key());
    if (tf) {
      d->add_observer(observer_key(), [this]() { this->mark_stale(); });
      d->Xptr()->add_observer(observer_key(), [this]() { this->mark_stale(); });
    }
  }
}
int BinomialProbitSpikeSlabSampler::device_step(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, double *suf_dev) {
  return boomgpu_probit_step_device(ctx, beta, clt_threshold_, seed, iteration, suf_dev, 0);   // hook path: the full statistics
}
int BinomialProbitSpikeSlabSampler::device_step_sync(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, double *xtx,
                                                     double *xty, double scalars[4]) {
  int64_t ss = 0;
  const int rc = boomgpu_probit_step(ctx, beta, clt_threshold_, seed, iteration, want_xtx_ ? xtx : nullptr, xty, &ss);
  scalars[0] = (double)ss; scalars[1] = scalars[2] = scalars[3] = 0.0;
  return rc;
}
int BinomialProbitSpikeSlabSampler::device_loglike_derivs(boomgpu_ctx *, const double *, double *, double *, double *) {
  report_error("the probit sibling provides the Gibbs step only");
  return 1;
}
int BinomialProbitSpikeSlabSampler::device_loglike_derivs_device(boomgpu_ctx *, const double *, double *) {
  report_error("the probit sibling provides the Gibbs step only");
  return 1;
}
int BinomialProbitSpikeSlabSampler::device_loglike_derivs_selected(boomgpu_ctx *, const double *, double *, double *, double *) {
  report_error("the probit sibling provides the Gibbs step only");
  return 1;
}
int BinomialProbitSpikeSlabSampler::device_loglike_derivs_selected_device(boomgpu_ctx *, const double *, double *) {
  report_error("the probit sibling provides the Gibbs step only");
  return 1;
}
void BinomialProbitSpikeSlabSampler::impute_latent_data() {
  DeviceImputerBase::impute_latent_data();   // (re)packs when needed -> install_tables() -> want_xtx_
  want_xtx_ = false;
}
void BinomialProbitSpikeSlabSampler::draw() {   // BinomialProbitSpikeSlabSampler.cpp:42-47
  impute_latent_data();
  if (allow_model_selection_) {
    BOOM_B200::SpikeSlabCore c(core(slab_, spike_, true));
    c.allow_model_selection(true);
    c.limit_model_selection(max_flips_);
    sweep_indicators(model_->coef(), c);
  }
  draw_included_beta(model_->coef(), core(slab_, spike_, true));
}
double BinomialProbitSpikeSlabSampler::logpri() const { return spike_slab_logpri(model_->coef(), *slab_, *spike_); }
WeightedRegSuf BinomialProbitSpikeSlabSampler::complete_data_sufficient_statistics() const {
  const int p = xdim_;
  WeightedRegSuf suf(p);
  SpdMatrix xtx(p);
  std::copy(hsuf_.xtx().a.begin(), hsuf_.xtx().a.end(), xtx.data());
  suf.set_xtwx(xtx);
  suf.set_xtwy(Vector(hsuf_.xty().begin(), hsuf_.xty().end()));
  return suf;
}

// ---------------------------------------------------------------------------------------------
// The Student-t sibling.
namespace {
// the functors BOOM's ScalarSliceSampler holds (it copies them: they carry only a pointer)
struct NuObservedTarget {
  std::function<double(double)> f;
  double operator()(double nu) const { return f(nu); }
};
}  // namespace

TRegressionSampler::TRegressionSampler(TRegressionModel *model, const Ptr<MvnBase> &coefficient_prior,
                                       const Ptr<GammaModelBase> &siginv_prior, const Ptr<DoubleModel> &nu_prior, RNG &seeding_rng)
    : DeviceImputerBase(model->xdim(), seeding_rng), model_(model), coefficient_prior_(coefficient_prior),
      siginv_prior_(siginv_prior), nu_prior_(nu_prior), weight_model_(new ScaledChisqModel(model->nu())),
      sigsq_sampler_(siginv_prior),
      nu_observed_data_sampler_(NuObservedTarget{[this](double nu) { return this->nu_log_posterior(nu); }}, false, 1.0, &rng()),
      nu_complete_data_sampler_(NuObservedTarget{[this](double nu) {
                                  // TRegressionCompleteDataLogPosterior (TRegressionSampler.cpp:51-72)
                                  if (nu <= 0.0) return negative_infinity();
                                  const double ans = nu_prior_->logp(nu);
                                  if (ans <= negative_infinity()) return ans;
                                  return ans + weight_model_->log_likelihood(nu);
                                }}, false, 1.0, &rng()),
      suf_(model->xdim()) {
  if (coefficient_prior_->dim() != model_->xdim()) report_error("Prior does not match model dimension.");
  nu_observed_data_sampler_.set_lower_limit(0.0);
  nu_complete_data_sampler_.set_lower_limit(0.0);
  model_->add_observer([this]() { this->mark_stale(); });
}

void TRegressionSampler::pack_rows(int64_t row0, int64_t nrows, double *X, void *y, double *) const {
  const std::vector<Ptr<RegressionData>> &data(model_->dat());
  const int p = xdim_;
  double *yd = static_cast<double *>(y);
  for (int64_t i = 0; i < nrows; ++i) {
    const RegressionData &d(*data[row0 + i]);
    const Vector &x(d.x());
    std::copy(x.begin(), x.end(), X + (size_t)i * p);
    yd[i] = d.y();
  }
}
void TRegressionSampler::observe_row_objects(bool tf) {
  for (const Ptr<RegressionData> &d : model_->dat()) {
    d->remove_observer(observer_key());
    d->Xptr()->remove_observer(observer_key());
    if (tf) {
      d->add_observer(observer_key(), [this]() { this->mark_stale(); });
      d->Xptr()->add_observer(observer_key(), [this]() { this->mark_stale(); });
    }
  }
}
int TRegressionSampler::device_step(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, double *suf_dev) {
  return boomgpu_student_step_device(ctx, beta, model_->sigma(), model_->nu(), seed, iteration, suf_dev);
}
int TRegressionSampler::device_step_sync(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, double *xtx,
                                         double *xty, double scalars[4]) {
  return boomgpu_student_step(ctx, beta, model_->sigma(), model_->nu(), seed, iteration, xtx, xty, scalars);
}
int TRegressionSampler::device_step_active(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration, const int32_t *cols,
                                           int k, double *G, double *diag, double *xty, double scalars[4]) {
  return boomgpu_student_step_active(ctx, beta, model_->sigma(), model_->nu(), seed, iteration, cols, k, G, diag, xty, scalars);
}
int TRegressionSampler::device_loglike_derivs(boomgpu_ctx *, const double *, double *, double *, double *) {
  report_error("derivatives of the Student-t log likelihood are not provided (TRegression.cpp:118-121)");
  return 1;
}
int TRegressionSampler::device_loglike_derivs_device(boomgpu_ctx *, const double *, double *) {
  report_error("derivatives of the Student-t log likelihood are not provided (TRegression.cpp:118-121)");
  return 1;
}
int TRegressionSampler::device_loglike_derivs_selected(boomgpu_ctx *, const double *, double *, double *, double *) {
  report_error("derivatives of the Student-t log likelihood are not provided (TRegression.cpp:118-121)");
  return 1;
}
int TRegressionSampler::device_loglike_derivs_selected_device(boomgpu_ctx *, const double *, double *) {
  report_error("derivatives of the Student-t log likelihood are not provided (TRegression.cpp:118-121)");
  return 1;
}

void TRegressionSampler::draw() {   // TRegressionSampler.cpp:114-119
  impute_latent_data();
  draw_beta_full_conditional();
  draw_sigsq_full_conditional();
  draw_nu_given_observed_data();
}
double TRegressionSampler::logpri() const {   // .cpp:121-126
  return nu_prior_->logp(model_->nu()) + sigsq_sampler_.log_prior(model_->sigsq()) + coefficient_prior_->logp(model_->Beta());
}

const WeightedRegSuf &TRegressionSampler::complete_data_sufficient_statistics() const {
  materialize_full_statistics();
  if (!suf_synced_) {
    const int p = xdim_;
    SpdMatrix xtx(p);
    std::copy(hsuf_.xtx().a.begin(), hsuf_.xtx().a.end(), xtx.data());
    suf_.reset(xtx, Vector(hsuf_.xty().begin(), hsuf_.xty().end()), hsuf_.yty(), hsuf_.n(), hsuf_.sumw(), hsuf_.sumlogw());
    suf_synced_ = true;
  }
  return suf_;
}
void TRegressionSampler::clear_complete_data_sufficient_statistics() {
  DeviceImputerBase::clear_complete_data_sufficient_statistics();
  weight_model_->clear_data();
}
void TRegressionSampler::update_complete_data_sufficient_statistics(double y, const Vector &x, double weight) {
  hsuf_.add_data(BOOM_B200::Vector(x.begin(), x.end()), y, weight);
  statistics_changed();
}

// draw_beta_full_conditional_impl (.cpp:74-85): precision = Ominv + X'WX / sigsq, scaled mean = Ominv b + X'Wy / sigsq
void TRegressionSampler::draw_beta_full_conditional() {
  const int p = xdim_;
  const double sigsq = model_->sigsq();
  SpdMatrix precision(coefficient_prior_->siginv());
  Vector scaled_mean = coefficient_prior_->siginv() * coefficient_prior_->mu();
  const std::vector<double> &a(hsuf_.xtx().a);
  for (int i = 0; i < p; ++i) {
    for (int j = 0; j < p; ++j) precision(i, j) += a[(size_t)i * p + j] / sigsq;
    scaled_mean[i] += hsuf_.xty()[i] / sigsq;
  }
  model_->set_Beta(rmvn_suf_mt(rng(), precision, scaled_mean));
  residuals_current_ = false;
}
// .cpp:165-171 with WeightedRegSuf::weighted_sum_of_squared_errors (WeightedRegressionModel.cpp:89-95) on the landed statistics
void TRegressionSampler::draw_sigsq_full_conditional() {
  model_->set_sigsq(sigsq_sampler_.draw(rng(), hsuf_.n(), weighted_sum_of_squared_errors()));
}
double TRegressionSampler::weighted_sum_of_squared_errors() {
  const int p = xdim_;
  const Vector &b(model_->Beta());
  const std::vector<double> &a(hsuf_.xtx().a);
  double bxy = 0, bxxb = 0;
  for (int i = 0; i < p; ++i) {
    if (b[i] == 0.0) continue;
    bxy += b[i] * hsuf_.xty()[i];
    double s = 0;
    for (int j = 0; j < p; ++j) s += a[(size_t)i * p + j] * b[j];
    bxxb += b[i] * s;
  }
  return hsuf_.yty() - 2 * bxy + bxxb;
}
void TRegressionSampler::draw_nu_given_complete_data() {   // .cpp:173-176: the weights' GammaSuf = (n, sum w, sum log w)
  weight_model_->suf()->set(hsuf_.sumw(), hsuf_.sumlogw(), hsuf_.n());
  model_->set_nu(nu_complete_data_sampler_.draw(model_->nu()));
}
void TRegressionSampler::draw_nu_given_observed_data() {   // .cpp:178-181
  residuals_current_ = false;
  model_->set_nu(nu_observed_data_sampler_.draw(model_->nu()));
}
double TRegressionSampler::nu_log_posterior(double nu) {   // TRegressionLogPosterior (.cpp:31-49)
  double ans = nu_prior_->logp(nu);
  if (ans <= negative_infinity()) return ans;
  if (!(nu > 0)) return negative_infinity();
  ensure_device_rows();
  double ll = 0;
  ++ll_evals_;
  check(boomgpu_student_loglike(device_ctx(), residuals_current_ ? nullptr : model_->Beta().data(), model_->sigma(), nu, &ll));
  residuals_current_ = true;
  return ans + ll;
}
double TRegressionSampler::log_likelihood(const Vector &beta, double sigsq, double nu) {
  if (!(nu > 0) || !(sigsq > 0)) return negative_infinity();
  ensure_device_rows();
  double ll = 0;
  residuals_current_ = false;
  check(boomgpu_student_loglike(device_ctx(), beta.data(), std::sqrt(sigsq), nu, &ll));
  return ll;
}

// ---- TRegressionSpikeSlabSampler
TRegressionSpikeSlabSampler::TRegressionSpikeSlabSampler(TRegressionModel *model, const Ptr<MvnBase> &slab,
                                                         const Ptr<VariableSelectionPrior> &spike, const Ptr<GammaModelBase> &siginv_prior,
                                                         const Ptr<DoubleModel> &nu_prior, RNG &seeding_rng)
    : TRegressionSampler(model, slab, siginv_prior, nu_prior, seeding_rng), spike_(spike), scaled_(model->xdim()) {
  if ((int)spike_->potential_nvars() != model->xdim()) report_error("Prior does not match model dimension.");
}
const BOOM_B200::WeightedRegSuf &TRegressionSpikeSlabSampler::scaled_statistics() {
  const int p = xdim_;
  const double inv = 1.0 / model_->sigsq();
  double *a = scaled_.xtx_storage(p), *b = scaled_.xty_storage();
  const std::vector<double> &src(hsuf_.xtx().a);
  for (size_t e = 0; e < (size_t)p * p; ++e) a[e] = src[e] * inv;
  for (int j = 0; j < p; ++j) b[j] = hsuf_.xty()[j] * inv;
  scaled_.set_scalars(hsuf_.n(), hsuf_.yty() * inv, hsuf_.sumw(), hsuf_.sumlogw());
  return scaled_;
}
// the active-set arrays of this iteration divided by the model's current sigsq, every fetched column scaled alike
BOOM_B200::StatView &TRegressionSpikeSlabSampler::scaled_view() {
  const double sigsq = model_->sigsq();
  if (tview_ && tview_sigsq_ == sigsq && tview_iteration_ == device_iteration()) return *tview_;
  const int p = xdim_;
  const double inv = 1.0 / sigsq;
  BOOM_B200::Vector G(active_.G), diag(active_.diag);
  for (double &v : G) v *= inv;
  for (double &v : diag) v *= inv;
  tview_xty_ = active_.xty;
  for (double &v : tview_xty_) v *= inv;
  boomgpu_ctx *ctx = device_ctx();
  BOOM_B200::ActiveSetState *st = &active_;
  tview_.reset(new BOOM_B200::StatView(p, active_.cols, G, diag, tview_xty_, [this, ctx, st, inv, p](int j, double *out) {
    check(boomgpu_weighted_column(ctx, j, out));
    for (int i = 0; i < p; ++i) out[i] *= inv;
    ++st->columns_fetched;
  }));
  tview_sigsq_ = sigsq;
  tview_iteration_ = device_iteration();
  return *tview_;
}
double TRegressionSpikeSlabSampler::weighted_sum_of_squared_errors() {
  if (!active_.valid) return TRegressionSampler::weighted_sum_of_squared_errors();
  BOOM_B200::StatView &v(scaled_view());
  const double sigsq = tview_sigsq_;
  const Vector &b(model_->Beta());
  const Selector &inc(model_->coef().inc());
  double bxy = 0, bxxb = 0;
  for (int a = 0; a < (int)inc.nvars(); ++a) {
    const int i = inc.indx(a);
    bxy += b[i] * active_.xty[i];
    double s = 0;
    for (int c = 0; c < (int)inc.nvars(); ++c) { const int j = inc.indx(c); s += v.at(i, j) * b[j]; }
    bxxb += b[i] * s;
  }
  return active_.scalars[1] - 2 * bxy + bxxb * sigsq;
}
void TRegressionSpikeSlabSampler::draw() {   // TRegressionSpikeSlabSampler.cpp:41-47
  if (impute_latent_data_active(model_->coef().inc())) {
    // n, y'Wy, sum w, sum log w are complete in this form too: the sigsq draw and nu | weights read them from hsuf_
    hsuf_.set_scalars(active_.scalars[0], active_.scalars[1], active_.scalars[2], active_.scalars[3]);
  } else {
    impute_latent_data();
  }
  draw_model_indicators();
  draw_included_coefficients();
  draw_sigsq_full_conditional();
  draw_nu_given_observed_data();
}
double TRegressionSpikeSlabSampler::logpri() const {   // .cpp:49-52
  return spike_slab_logpri(model_->coef(), *coefficient_prior_, *spike_) + nu_prior_->logp(model_->nu()) +
         siginv_prior_->logp(1.0 / model_->sigsq());
}
void TRegressionSpikeSlabSampler::draw_model_indicators() {
  if (!allow_model_selection_) return;
  BOOM_B200::SpikeSlabCore c(core(coefficient_prior_, spike_, true));
  c.allow_model_selection(true);
  c.limit_model_selection(max_flips_);
  BOOM_B200::GlmCoefs h = host_coefs(model_->coef(), xdim_, true);
  BOOM_B200::RNG local(rng().generator()());
  if (active_.valid) c.draw_model_indicators(local, h, scaled_view());
  else c.draw_model_indicators(local, h, scaled_statistics());
  write_back(model_->coef(), h, xdim_, false);
  coefficients_changed();
}
void TRegressionSpikeSlabSampler::draw_included_coefficients() {
  BOOM_B200::GlmCoefs h = host_coefs(model_->coef(), xdim_, false);
  BOOM_B200::RNG local(rng().generator()());
  if (active_.valid) core(coefficient_prior_, spike_, true).draw_beta(local, h, scaled_view());
  else core(coefficient_prior_, spike_, true).draw_beta(local, h, scaled_statistics());
  write_back(model_->coef(), h, xdim_, true);
  coefficients_changed();
}

// ---------------------------------------------------------------------------------------------
PoissonRegressionAuxMixSampler::PoissonRegressionAuxMixSampler(PoissonRegressionModel *model, const Ptr<MvnBase> &prior, int,
                                                               RNG &seeding_rng)
    : DeviceImputerBase(model->xdim(), seeding_rng), model_(model), prior_(prior), suf_(model->xdim()) {
  if (prior_->dim() != model_->xdim()) report_error("Prior does not match model dimension.");
  model_->add_observer([this]() { this->mark_stale(); });
}

const WeightedRegSuf &PoissonRegressionAuxMixSampler::complete_data_sufficient_statistics() const {
  materialize_full_statistics();
  if (!suf_synced_) {
    const int p = xdim_;
    SpdMatrix xtx(p);
    std::copy(hsuf_.xtx().a.begin(), hsuf_.xtx().a.end(), xtx.data());
    Vector xty(hsuf_.xty().begin(), hsuf_.xty().end());
    suf_.reset(xtx, xty, hsuf_.yty(), hsuf_.n(), hsuf_.sumw(), hsuf_.sumlogw());   // WeightedRegressionModel.cpp:148-157
    suf_synced_ = true;
  }
  return suf_;
}

void PoissonRegressionAuxMixSampler::install_tables(boomgpu_ctx *ctx) {
  // The table is materialised by the reference's own code for every count in the data (first touch of an
  // off-grid value may run its Powell re-fit on the host), then stated in its serialized layout.
  static NormalMixtureApproximationTable table = create_poisson_mixture_approximation_table();
  table.approximate(1);
  std::set<int64_t> distinct;
  for (const Ptr<PoissonRegressionData> &d : model_->dat()) if (d->y() > 0) distinct.insert(d->y());
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
}

void PoissonRegressionAuxMixSampler::pack_rows(int64_t row0, int64_t nrows, double *X, void *y, double *aux) const {
  const std::vector<Ptr<PoissonRegressionData>> &data(model_->dat());
  const int p = xdim_;
  int64_t *yi = static_cast<int64_t *>(y);
  for (int64_t i = 0; i < nrows; ++i) {
    const PoissonRegressionData &d(*data[row0 + i]);
    const Vector &x(d.x());
    std::copy(x.begin(), x.end(), X + (size_t)i * p);
    yi[i] = d.y();
    aux[i] = d.exposure();
  }
}

void PoissonRegressionAuxMixSampler::observe_row_objects(bool tf) {
  for (const Ptr<PoissonRegressionData> &d : model_->dat()) {
    d->remove_observer(observer_key());
    d->Xptr()->remove_observer(observer_key());
    if (tf) {
      d->add_observer(observer_key(), [this]() { this->mark_stale(); });
      d->Xptr()->add_observer(observer_key(), [this]() { this->mark_stale(); });
    }
  }
}

int PoissonRegressionAuxMixSampler::device_step(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration,
                                                double *suf_dev) {
  return boomgpu_poisson_step_device(ctx, beta, seed, iteration, suf_dev);
}
int PoissonRegressionAuxMixSampler::device_step_sync(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration,
                                                     double *xtx, double *xty, double scalars[4]) {
  return boomgpu_poisson_step(ctx, beta, seed, iteration, xtx, xty, scalars);
}
int PoissonRegressionAuxMixSampler::device_step_active(boomgpu_ctx *ctx, const double *beta, uint64_t seed, uint64_t iteration,
                                                       const int32_t *cols, int k, double *G, double *diag, double *xty, double scalars[4]) {
  return boomgpu_poisson_step_active(ctx, beta, seed, iteration, cols, k, G, diag, xty, scalars);
}
int PoissonRegressionAuxMixSampler::device_loglike_derivs(boomgpu_ctx *ctx, const double *beta, double *loglike, double *g, double *h) {
  return boomgpu_poisson_loglike_derivs(ctx, beta, loglike, g, h);
}
int PoissonRegressionAuxMixSampler::device_loglike_derivs_device(boomgpu_ctx *ctx, const double *beta, double *suf_dev) {
  return boomgpu_poisson_loglike_derivs_device(ctx, beta, suf_dev);
}
int PoissonRegressionAuxMixSampler::device_loglike_derivs_selected(boomgpu_ctx *ctx, const double *beta, double *loglike, double *g, double *h) {
  return boomgpu_poisson_loglike_derivs_selected(ctx, beta, loglike, g, h);
}
int PoissonRegressionAuxMixSampler::device_loglike_derivs_selected_device(boomgpu_ctx *, const double *, double *) {
  report_error("the selected-column Poisson derivatives have no hook-driven form; attach a native communicator instead");
  return 1;
}
void PoissonRegressionAuxMixSampler::draw() {
  impute_latent_data();
  draw_beta_given_complete_data();
}
void PoissonRegressionAuxMixSampler::draw_beta_given_complete_data() { draw_beta_full_model(model_->coef(), *prior_); }
double PoissonRegressionAuxMixSampler::logpri() const { return prior_->logp(model_->Beta()); }
void PoissonRegressionAuxMixSampler::update_complete_data_sufficient_statistics(double precision_weighted_sum,
                                                                                double total_precision, const Vector &x) {
  hsuf_.add_data(to_host(x), precision_weighted_sum / total_precision, total_precision);   // PoissonRegressionAuxMixSampler.cpp:153-158
  statistics_changed();
}

PoissonRegressionSpikeSlabSampler::PoissonRegressionSpikeSlabSampler(PoissonRegressionModel *model, const Ptr<MvnBase> &slab,
                                                                     const Ptr<VariableSelectionPrior> &spike, int nthreads,
                                                                     RNG &seeding_rng)
    : PoissonRegressionAuxMixSampler(model, slab, nthreads, seeding_rng), slab_(slab), spike_(spike) {
  if ((int)spike_->potential_nvars() != model_->xdim()) report_error("Spike does not match model dimension.");
}
void PoissonRegressionSpikeSlabSampler::draw() {   // PoissonRegressionSpikeSlabSampler.cpp:55-59
  if (!impute_latent_data_active(model_->coef().inc())) impute_latent_data();
  if (allow_model_selection_) draw_model_indicators();
  draw_beta();
}
void PoissonRegressionSpikeSlabSampler::draw_model_indicators() {
  BOOM_B200::SpikeSlabCore c(core(slab_, spike_, true));
  c.allow_model_selection(true);
  c.limit_model_selection(max_flips_);
  sweep_indicators(model_->coef(), c);
}
void PoissonRegressionSpikeSlabSampler::draw_beta() { draw_included_beta(model_->coef(), core(slab_, spike_, true)); }
double PoissonRegressionSpikeSlabSampler::log_model_prob(const Selector &gamma) const {
  return model_log_prob(gamma, core(slab_, spike_, true));
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
  find_mode(model_->coef(), slab_, spike_, epsilon, &log_posterior_at_mode_);
}

}  // namespace B200
}  // namespace BOOM
