// boom_b200.cpp -- host side of the B200-native auxiliary-mixture samplers (see boom_b200.hpp).
#include "boom_b200.hpp"

#include <algorithm>
#include <cmath>
#include <limits>
#include <mutex>
#include <thread>
#include <sstream>

#include "../../include/boomgpu.h"

namespace BOOM_B200 {

void report_error(const std::string &msg) { throw std::runtime_error(msg); }

RNG GlobalRng::rng(8675309);

// distributions/rng.cpp:39-47 draws the seed as llround(U * 2^64), which overflows long long for half of all U
// (every such seed collapses to one value).  Same contract here -- a seed > 2 from the parent stream -- but taken
// from the generator's 64 raw bits.
RNG::RngIntType seed_rng(RNG &rng) {
  RNG::RngIntType ans = 0;
  while (ans <= 2) ans = rng.generator()();
  return ans;
}

double runif_mt(RNG &rng, double lo, double hi) { return lo + (hi - lo) * rng(); }

double rnorm_mt(RNG &rng, double mu, double sd) {
  // Box-Muller on the sampler's own stream (the reference uses Kinderman-Ramage, Bmath/snorm.cpp:77;
  // only the distribution has to agree).
  double u1 = rng(), u2 = rng();
  while (u1 <= 0) u1 = rng();
  return mu + sd * std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925286766559 * u2);
}

int random_int_mt(RNG &rng, int lo, int hi) { return (int)std::floor(runif_mt(rng, lo, hi + 1)); }

// ---------------------------------------------------------------------------------------------
// ---- Cholesky ---------------------------------------------------------------------------------
// Left-looking, blocked, row major (every inner product runs over contiguous memory).  For block column J:
//   (1) A[i][J] -= L[i][0:jb] . L[J][0:jb]'   for all rows i >= jb  -- p^3 / 3 of the flops, as 4 x 4 register tiles of
//       dot products (vectorised over k; AVX-512 / AVX2 clones chosen at load time), split over threads when large;
//   (2) unblocked factorisation of the diagonal block;  (3) triangular solve of the rows below it.
// The full-model beta draw needs it at p = 500 (42 MFLOP) to p = 4000 (21 GFLOP); the reference leans on Eigen's LLT.
namespace {
constexpr int kCholBlock = 64;

// C[i][j] -= sum_k A[i][k] B[j][k]  for i in [0, mi), j in [0, nj): rows of A, B and C are lda / ldb / ldc apart
__attribute__((target_clones("avx512f", "avx2,fma", "default")))
void gemm_nt_minus(int mi, int nj, int kk, const double *A, size_t lda, const double *B, size_t ldb, double *C, size_t ldc,
                   bool lower_only, int i_off, int j_off) {
  for (int i0 = 0; i0 < mi; i0 += 4) {
    const int ib = std::min(4, mi - i0);
    for (int j0 = 0; j0 < nj; j0 += 4) {
      if (lower_only && j_off + j0 > i_off + i0 + ib - 1) break;   // the tile lies strictly above the diagonal
      const int jb = std::min(4, nj - j0);
      if (ib == 4 && jb == 4) {
        const double *a0 = A + (size_t)i0 * lda, *a1 = a0 + lda, *a2 = a1 + lda, *a3 = a2 + lda;
        const double *b0 = B + (size_t)j0 * ldb, *b1 = b0 + ldb, *b2 = b1 + ldb, *b3 = b2 + ldb;
        double c00 = 0, c01 = 0, c02 = 0, c03 = 0, c10 = 0, c11 = 0, c12 = 0, c13 = 0;
        double c20 = 0, c21 = 0, c22 = 0, c23 = 0, c30 = 0, c31 = 0, c32 = 0, c33 = 0;
#pragma omp simd reduction(+ : c00, c01, c02, c03, c10, c11, c12, c13, c20, c21, c22, c23, c30, c31, c32, c33)
        for (int k = 0; k < kk; ++k) {
          const double x0 = a0[k], x1 = a1[k], x2 = a2[k], x3 = a3[k];
          const double y0 = b0[k], y1 = b1[k], y2 = b2[k], y3 = b3[k];
          c00 += x0 * y0; c01 += x0 * y1; c02 += x0 * y2; c03 += x0 * y3;
          c10 += x1 * y0; c11 += x1 * y1; c12 += x1 * y2; c13 += x1 * y3;
          c20 += x2 * y0; c21 += x2 * y1; c22 += x2 * y2; c23 += x2 * y3;
          c30 += x3 * y0; c31 += x3 * y1; c32 += x3 * y2; c33 += x3 * y3;
        }
        double *c = C + (size_t)i0 * ldc + j0;
        c[0] -= c00; c[1] -= c01; c[2] -= c02; c[3] -= c03; c += ldc;
        c[0] -= c10; c[1] -= c11; c[2] -= c12; c[3] -= c13; c += ldc;
        c[0] -= c20; c[1] -= c21; c[2] -= c22; c[3] -= c23; c += ldc;
        c[0] -= c30; c[1] -= c31; c[2] -= c32; c[3] -= c33;
      } else {
        for (int i = 0; i < ib; ++i)
          for (int j = 0; j < jb; ++j) {
            const double *ar = A + (size_t)(i0 + i) * lda, *br = B + (size_t)(j0 + j) * ldb;
            double acc = 0;
            for (int k = 0; k < kk; ++k) acc += ar[k] * br[k];
            C[(size_t)(i0 + i) * ldc + j0 + j] -= acc;
          }
      }
    }
  }
}

template <class F>
void parallel_rows(int begin, int end, int chunk, double work, F f) {
  const int nchunks = (end - begin + chunk - 1) / chunk;
  unsigned hw = std::thread::hardware_concurrency();
  int nthreads = (int)std::min<unsigned>(hw ? hw : 1, 16);
  if (work < 2e7 || nchunks < 2 || nthreads < 2) { f(begin, end); return; }   // small: thread start-up would dominate
  nthreads = std::min(nthreads, nchunks);
  std::vector<std::thread> pool;
  const int per = (nchunks + nthreads - 1) / nthreads * chunk;
  for (int t = 0; t < nthreads; ++t) {
    const int b = begin + t * per, e = std::min(end, b + per);
    if (b >= e) break;
    pool.emplace_back([=]() { f(b, e); });
  }
  for (auto &th : pool) th.join();
}
}  // namespace

bool cholesky_lower(double *a, int n) {
  const size_t ld = (size_t)n;
  for (int jb = 0; jb < n; jb += kCholBlock) {
    const int bw = std::min(kCholBlock, n - jb);
    // (1) rows jb .. n of block column J minus the contribution of the columns already factorised
    if (jb > 0) {
      // rows of the diagonal block first (lower triangle only), then the rows below it in parallel
      gemm_nt_minus(bw, bw, jb, a + (size_t)jb * ld, ld, a + (size_t)jb * ld, ld, a + (size_t)jb * ld + jb, ld, true, 0, 0);
      const int below = jb + bw;
      parallel_rows(below, n, 64, 2.0 * (double)(n - below) * bw * jb, [&](int r0, int r1) {
        gemm_nt_minus(r1 - r0, bw, jb, a + (size_t)r0 * ld, ld, a + (size_t)jb * ld, ld, a + (size_t)r0 * ld + jb, ld, false, 0, 0);
      });
    }
    // (2) the diagonal block, unblocked
    for (int j = jb; j < jb + bw; ++j) {
      double *rj = a + (size_t)j * ld;
      double d = rj[j];
      for (int k = jb; k < j; ++k) d -= rj[k] * rj[k];
      if (!(d > 0) || !std::isfinite(d)) return false;
      d = std::sqrt(d);
      rj[j] = d;
      const double inv = 1.0 / d;
      for (int i = j + 1; i < jb + bw; ++i) {
        double *ri = a + (size_t)i * ld;
        double s = ri[j];
        for (int k = jb; k < j; ++k) s -= ri[k] * rj[k];
        ri[j] = s * inv;
      }
    }
    // (3) rows below: L[i][J] = A[i][J] L[J][J]^-T
    for (int i = jb + bw; i < n; ++i) {
      double *ri = a + (size_t)i * ld;
      for (int j = jb; j < jb + bw; ++j) {
        const double *rj = a + (size_t)j * ld;
        double s = ri[j];
        for (int k = jb; k < j; ++k) s -= ri[k] * rj[k];
        ri[j] = s / rj[j];
      }
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) a[(size_t)i * n + j] = 0.0;
  return true;
}

void lsolve_inplace(const double *L, int n, double *b) {
  for (int i = 0; i < n; ++i) {
    const double *ri = L + (size_t)i * n;
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= ri[k] * b[k];
    b[i] = s / ri[i];
  }
}

void ltsolve_inplace(const double *L, int n, double *b) {
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i] / L[(size_t)i * n + i];
    b[i] = s;
    const double *ri = L + (size_t)i * n;
    for (int k = 0; k < i; ++k) b[k] -= ri[k] * s;
  }
}

// rmvn_suf_mt (distributions/mvn.cpp:128-136): beta = L^-T z + (L L^T)^-1 ivar_mu
Vector rmvn_suf_mt(RNG &rng, const SpdMatrix &ivar, const Vector &ivar_mu) {
  const int n = ivar.dim;
  Vector L(ivar.a);
  if (!cholesky_lower(L.data(), n)) report_error("Cholesky decomposition failed in rmvn_suf_mt.");
  Vector z(n);
  for (int i = 0; i < n; ++i) z[i] = rnorm_mt(rng);
  ltsolve_inplace(L.data(), n, z.data());
  Vector m(ivar_mu);
  lsolve_inplace(L.data(), n, m.data());
  ltsolve_inplace(L.data(), n, m.data());
  for (int i = 0; i < n; ++i) z[i] += m[i];
  return z;
}

static double logdet_spd(const SpdMatrix &m, bool *ok) {
  Vector L(m.a);
  if (!cholesky_lower(L.data(), m.dim)) { *ok = false; return -std::numeric_limits<double>::infinity(); }
  *ok = true;
  double s = 0;
  for (int i = 0; i < m.dim; ++i) s += std::log(L[(size_t)i * m.dim + i]);
  return 2 * s;
}

// ---------------------------------------------------------------------------------------------
int Selector::nvars() const { return (int)std::count(inc_.begin(), inc_.end(), true); }

std::vector<int> Selector::included_positions() const {
  std::vector<int> pos;
  for (int i = 0; i < (int)inc_.size(); ++i) if (inc_[i]) pos.push_back(i);
  return pos;
}

Vector Selector::select(const Vector &v) const {
  if ((int)v.size() != nvars_possible()) report_error("Selector::select: wrong size vector");
  Vector out;
  for (int i = 0; i < (int)inc_.size(); ++i) if (inc_[i]) out.push_back(v[i]);
  return out;
}

SpdMatrix Selector::select(const SpdMatrix &m) const {
  const std::vector<int> pos = included_positions();
  const int k = (int)pos.size();
  SpdMatrix out(k);
  for (int i = 0; i < k; ++i) {
    const double *row = m.a.data() + (size_t)pos[i] * m.dim;
    for (int j = 0; j < k; ++j) out.a[(size_t)i * k + j] = row[pos[j]];
  }
  return out;
}

Vector Selector::expand(const Vector &sub) const {
  Vector out(inc_.size(), 0.0);
  int k = 0;
  for (int i = 0; i < (int)inc_.size(); ++i) if (inc_[i]) out[i] = sub[k++];
  return out;
}

// ---------------------------------------------------------------------------------------------
double MvnBase::logp(const Vector &x) const {
  const int p = dim();
  const SpdMatrix &P(siginv());
  bool ok = true;
  const double ld = logdet_spd(P, &ok);
  if (!ok) return -std::numeric_limits<double>::infinity();
  double q = 0;
  for (int i = 0; i < p; ++i) {
    double s = 0;
    for (int j = 0; j < p; ++j) s += P(i, j) * (x[j] - mu()[j]);
    q += s * (x[i] - mu()[i]);
  }
  return -0.5 * p * std::log(6.283185307179586476925286766559) + 0.5 * ld - 0.5 * q;
}

MvnModel::MvnModel(const Vector &mean, const SpdMatrix &V, bool ivar) : mu_(mean), siginv_(V) {
  if ((int)mean.size() != V.dim) report_error("MvnModel: mean and variance dimensions differ");
  if (!ivar) {
    const int p = V.dim;
    bool diagonal = true;
    for (int i = 0; i < p && diagonal; ++i)
      for (int j = 0; j < p; ++j)
        if (i != j && V(i, j) != 0.0) { diagonal = false; break; }
    if (diagonal) {   // the usual N(0, s^2 I) slab: no factorisation needed (at p = 4000 the general path is 10^11 flops)
      for (int i = 0; i < p; ++i) {
        if (!(V(i, i) > 0)) report_error("MvnModel: variance matrix is not positive definite");
        siginv_(i, i) = 1.0 / V(i, i);
      }
      return;
    }
    Vector L(V.a);
    if (!cholesky_lower(L.data(), p)) report_error("MvnModel: variance matrix is not positive definite");
    // V^-1 = L^-T L^-1, column by column; column c of L^-1 is zero above row c
    for (int c = 0; c < p; ++c) {
      Vector e(p, 0.0);
      e[c] = 1.0;
      for (int i = c; i < p; ++i) {   // forward substitution from row c
        const double *ri = L.data() + (size_t)i * p;
        double t = e[i];
        for (int k = c; k < i; ++k) t -= ri[k] * e[k];
        e[i] = t / ri[i];
      }
      ltsolve_inplace(L.data(), p, e.data());
      for (int r = 0; r < p; ++r) siginv_(r, c) = e[r];
    }
  }
}

VariableSelectionPrior::VariableSelectionPrior(int n, double pr) : VariableSelectionPrior(Vector(n, pr)) {}

VariableSelectionPrior::VariableSelectionPrior(const Vector &probs) : probs_(probs), log_p_(probs.size()), log_q_(probs.size()) {
  for (size_t i = 0; i < probs.size(); ++i) {
    if (!(probs[i] >= 0 && probs[i] <= 1)) report_error("prior inclusion probabilities must lie in [0, 1]");
    log_p_[i] = probs[i] > 0 ? std::log(probs[i]) : -std::numeric_limits<double>::infinity();
    log_q_[i] = probs[i] < 1 ? std::log(1 - probs[i]) : -std::numeric_limits<double>::infinity();
  }
}

double VariableSelectionPrior::logp(const Selector &inc) const {
  if (max_model_size_ >= 0 && inc.nvars() > max_model_size_) return -std::numeric_limits<double>::infinity();
  double ans = 0;
  for (int i = 0; i < inc.nvars_possible(); ++i) {
    ans += inc[i] ? log_p_[i] : log_q_[i];
    if (!std::isfinite(ans)) return -std::numeric_limits<double>::infinity();
  }
  return ans;
}

// change of logp when variable j flips (included_now: its current state); -inf / nan never escape:
// a flip into a zero-probability state returns -inf
double VariableSelectionPrior::flip_delta(int j, bool included_now) const {
  const double to = included_now ? log_q_[j] : log_p_[j], from = included_now ? log_p_[j] : log_q_[j];
  if (!std::isfinite(to)) return -std::numeric_limits<double>::infinity();
  return to - from;
}

void VariableSelectionPrior::make_valid(Selector &inc) const {
  if (inc.nvars_possible() != (int)probs_.size()) report_error("Wrong size Selector passed to make_valid.");
  for (size_t i = 0; i < probs_.size(); ++i) {
    if (probs_[i] <= 0.0 && inc[i]) inc.flip(i);
    if (probs_[i] >= 1.0 && !inc[i]) inc.flip(i);
  }
}

// ---------------------------------------------------------------------------------------------
void GlmCoefs::set_Beta(const Vector &b) {
  if (b.size() != beta_.size()) report_error("GlmCoefs::set_Beta: wrong size");
  beta_ = b;
}

void GlmCoefs::set_inc(const Selector &g) {
  if (g.nvars_possible() != inc_.nvars_possible()) report_error("GlmCoefs::set_inc: wrong size");
  inc_ = g;
  for (size_t i = 0; i < beta_.size(); ++i) if (!inc_[i]) beta_[i] = 0.0;
}

void GlmCoefs::set_included_coefficients(const Vector &b) {
  if ((int)b.size() != inc_.nvars()) report_error("GlmCoefs::set_included_coefficients: wrong size");
  beta_ = inc_.expand(b);
}

// ---------------------------------------------------------------------------------------------
DeviceData::DeviceData(int device) : device_(device) {
  if (boomgpu_create(&ctx_, device)) report_error(std::string("boomgpu_create: ") + boomgpu_last_error(nullptr));
}
DeviceData::~DeviceData() { boomgpu_destroy(ctx_); }
void DeviceData::check(int rc) const {
  if (rc) report_error(std::string("boomgpu: ") + boomgpu_last_error(ctx_));
}

void GlmModelBase::set_method(const std::shared_ptr<PosteriorSampler> &sampler) { samplers_.push_back(sampler); }
void GlmModelBase::sample_posterior() {
  if (samplers_.empty()) report_error("sample_posterior() called with no sampler: call set_method first");
  for (auto &s : samplers_) s->draw();
}
double GlmModelBase::logpri() const {
  double ans = 0;
  for (auto &s : samplers_) ans += s->logpri();
  return ans;
}
void GlmModelBase::set_device(int device) {
  if (device != device_) { dev_.reset(); uploaded_version_ = 0; }
  device_ = device;
}
void GlmModelBase::set_stream(void *cuda_stream) {
  stream_ = cuda_stream; have_stream_ = true;
  if (dev_) dev_->check(boomgpu_set_stream(dev_->ctx(), cuda_stream));
}
void GlmModelBase::set_row_offset(uint64_t r) {
  row_offset_ = r;
  if (dev_) dev_->check(boomgpu_set_row_offset(dev_->ctx(), r));
}
DeviceData &GlmModelBase::device_data() {
  if (!dev_) {
    dev_.reset(new DeviceData(device_));
    uploaded_version_ = 0;
    if (have_stream_) dev_->check(boomgpu_set_stream(dev_->ctx(), stream_));
    for (auto &o : options_) dev_->check(boomgpu_set_option(dev_->ctx(), o.first.c_str(), o.second));
    if (!comm_id_.empty()) dev_->check(boomgpu_comm_init(dev_->ctx(), comm_id_.data(), comm_ranks_, comm_rank_));
  }
  if (uploaded_version_ != data_version_) {
    upload(*dev_);
    dev_->check(boomgpu_set_row_offset(dev_->ctx(), row_offset_));
    uploaded_version_ = data_version_;
  }
  return *dev_;
}

std::string GlmModelBase::comm_unique_id() {
  std::string id(BOOMGPU_COMM_ID_BYTES, '\0');
  if (boomgpu_comm_unique_id(&id[0])) report_error(std::string("boomgpu_comm_unique_id: ") + boomgpu_last_error(nullptr));
  return id;
}
void GlmModelBase::set_communicator(const std::string &id, int nranks, int rank) {
  if ((int)id.size() != BOOMGPU_COMM_ID_BYTES) report_error("set_communicator: the id must be the 128 bytes of comm_unique_id()");
  comm_id_ = id; comm_ranks_ = nranks; comm_rank_ = rank;
  if (dev_) dev_->check(boomgpu_comm_init(dev_->ctx(), comm_id_.data(), comm_ranks_, comm_rank_));
}
void GlmModelBase::set_device_option(const std::string &name, int64_t value) {
  bool known = false;
  for (auto &o : options_) if (o.first == name) { o.second = value; known = true; }
  if (!known) options_.emplace_back(name, value);
  if (dev_) dev_->check(boomgpu_set_option(dev_->ctx(), name.c_str(), value));
}
int64_t GlmModelBase::kernel_launches() { return dev_ ? boomgpu_kernel_launches(dev_->ctx()) : 0; }
void GlmModelBase::kernel_timings(double ms[5], int64_t launches[5], bool reset) {
  for (int c = 0; c < 5; ++c) { ms[c] = 0; launches[c] = 0; }
  if (dev_) dev_->check(boomgpu_get_timings(dev_->ctx(), ms, launches, reset ? 1 : 0));
}

BinomialLogitModel::BinomialLogitModel(int64_t n, int p, const double *X, const double *y, const double *nt)
    : GlmModelBase(p) {
  x_.assign(X, X + (size_t)n * p);
  y_.assign(y, y + n);
  n_.assign(nt, nt + n);
  for (int64_t i = 0; i < n; ++i)
    if (y_[i] > n_[i] || y_[i] < 0) report_error("BinomialRegressionData: y must lie in [0, n]");
}
void BinomialLogitModel::add_data(double y, double n, const Vector &x) {
  if ((int)x.size() != xdim()) report_error("BinomialLogitModel::add_data: wrong size x");
  if (y > n || y < 0 || n < 0) report_error("BinomialRegressionData: y must lie in [0, n]");
  if (adopted_ || borrowed_) report_error("add_data on a model whose data live in adopted / borrowed memory");
  x_.insert(x_.end(), x.begin(), x.end());
  y_.push_back(y);
  n_.push_back(n);
  touch();
}
void BinomialLogitModel::adopt_device_data(int64_t n, const double *dX, int64_t ldx, const double *dy, const double *dn) {
  adopted_ = true; adopted_n_ = n; dX_ = dX; dldx_ = ldx; dy_ = dy; dn_ = dn;
  touch();
}
void BinomialLogitModel::borrow_host_data(int64_t n, const double *X, int64_t ldx, const double *y, const double *nt,
                                          std::shared_ptr<void> keepalive) {
  borrowed_ = true; adopted_ = false; adopted_n_ = n; dX_ = X; dldx_ = ldx; dy_ = y; dn_ = nt;
  keepalive_ = std::move(keepalive);
  touch();
}
void BinomialLogitModel::upload(DeviceData &dev) {
  if (borrowed_) { dev.check(boomgpu_upload_binomial(dev.ctx(), adopted_n_, xdim(), dX_, dldx_, dy_, dn_)); return; }
  if (adopted_) dev.check(boomgpu_adopt_binomial(dev.ctx(), adopted_n_, xdim(), dX_, dldx_, dy_, dn_));
  else dev.check(boomgpu_upload_binomial(dev.ctx(), (int64_t)y_.size(), xdim(), x_.data(), xdim(), y_.data(), n_.data()));
}
// Rows sharded over ranks with a caller-supplied all-reduce hook: this rank's packed [ -H | g | {., ll, ., .} ] stays on the
// device for the hook, exactly like a Gibbs step's statistics (with a native communicator the C ABI all-reduces itself).
// Every rank therefore sees the likelihood of ALL rows: the mode finder and the MH moves run replicated and stay in step.
template <class StepDeviceFn>
static double loglike_through_hook(GlmModelBase &model, StepDeviceFn step_device, Vector *g, SpdMatrix *h) {
  DeviceData &dev(model.device_data());
  const int p = model.xdim();
  const int64_t len = boomgpu_suf_len(p);
  Vector packed((size_t)len);
  double *suf_dev = nullptr;
  dev.check(boomgpu_suf_buffer(dev.ctx(), &suf_dev));
  dev.check(step_device(dev.ctx(), suf_dev));
  model.allreduce()(suf_dev, len);
  dev.check(boomgpu_download(dev.ctx(), suf_dev, packed.data(), len));
  const size_t mat = (size_t)p * p;
  if (g) g->assign(packed.begin() + mat, packed.begin() + mat + p);
  if (h) {
    if (h->dim != p) *h = SpdMatrix(p);
    for (size_t e = 0; e < mat; ++e) h->a[e] = -packed[e];
  }
  return packed[mat + p + 1];
}

double BinomialLogitModel::log_likelihood(const Vector &beta) {
  if ((int)beta.size() != xdim()) report_error("log_likelihood: wrong size beta");
  if (log_alpha_ != 0.0 || allreduce()) return log_likelihood_derivs(beta, nullptr, nullptr);   // the offset enters eta (BinomialLogitModel.cpp:168)
  DeviceData &dev(device_data());
  double ans = 0;
  dev.check(boomgpu_binomial_loglike(dev.ctx(), beta.data(), &ans));
  return ans;
}

double BinomialLogitModel::log_likelihood_derivs(const Vector &beta, Vector *g, SpdMatrix *h) {
  if ((int)beta.size() != xdim()) report_error("log_likelihood: wrong size beta");
  if (allreduce()) {
    const double la = log_alpha_;
    return loglike_through_hook(*this, [&](boomgpu_ctx *ctx, double *suf_dev) {
      return boomgpu_binomial_loglike_derivs_device(ctx, beta.data(), la, suf_dev); }, g, h);
  }
  DeviceData &dev(device_data());
  double ans = 0;
  if (g) g->assign(xdim(), 0.0);
  if (h && h->dim != xdim()) *h = SpdMatrix(xdim());
  dev.check(boomgpu_binomial_loglike_derivs(dev.ctx(), beta.data(), log_alpha_, &ans, g ? g->data() : nullptr,
                                            h ? h->a.data() : nullptr));
  return ans;
}
void BinomialLogitModel::set_nonevent_sampling_prob(double alpha) {
  if (!(alpha > 0 && alpha <= 1)) report_error("alpha (proportion of non-events retained in the data) must be in (0, 1]");
  log_alpha_ = std::log(alpha);
}

PoissonRegressionModel::PoissonRegressionModel(int64_t n, int p, const double *X, const int64_t *y, const double *ex)
    : GlmModelBase(p) {
  x_.assign(X, X + (size_t)n * p);
  y_.assign(y, y + n);
  exposure_.assign(ex, ex + n);
}
void PoissonRegressionModel::add_data(int64_t y, const Vector &x, double exposure) {
  if ((int)x.size() != xdim()) report_error("PoissonRegressionModel::add_data: wrong size x");
  if (y < 0 || exposure < 0) report_error("PoissonRegressionData: y and exposure must be non-negative");
  if (adopted_ || borrowed_) report_error("add_data on a model whose data live in adopted / borrowed memory");
  x_.insert(x_.end(), x.begin(), x.end());
  y_.push_back(y);
  exposure_.push_back(exposure);
  touch();
}
void PoissonRegressionModel::adopt_device_data(int64_t n, const double *dX, int64_t ldx, const int64_t *dy, const double *dex) {
  adopted_ = true; adopted_n_ = n; dX_ = dX; dldx_ = ldx; dy_ = dy; dexp_ = dex;
  touch();
}
void PoissonRegressionModel::borrow_host_data(int64_t n, const double *X, int64_t ldx, const int64_t *y, const double *ex,
                                              std::shared_ptr<void> keepalive) {
  borrowed_ = true; adopted_ = false; adopted_n_ = n; dX_ = X; dldx_ = ldx; dy_ = y; dexp_ = ex;
  keepalive_ = std::move(keepalive);
  touch();
}
void PoissonRegressionModel::upload(DeviceData &dev) {
  if (borrowed_) { dev.check(boomgpu_upload_poisson(dev.ctx(), adopted_n_, xdim(), dX_, dldx_, dy_, dexp_)); return; }
  if (adopted_) dev.check(boomgpu_adopt_poisson(dev.ctx(), adopted_n_, xdim(), dX_, dldx_, dy_, dexp_));
  else dev.check(boomgpu_upload_poisson(dev.ctx(), (int64_t)y_.size(), xdim(), x_.data(), xdim(), y_.data(), exposure_.data()));
}
double PoissonRegressionModel::log_likelihood(const Vector &beta) {
  if ((int)beta.size() != xdim()) report_error("log_likelihood: wrong size beta");
  if (allreduce()) return log_likelihood_derivs(beta, nullptr, nullptr);
  DeviceData &dev(device_data());
  double ans = 0;
  dev.check(boomgpu_poisson_loglike(dev.ctx(), beta.data(), &ans));
  return ans;
}

double PoissonRegressionModel::log_likelihood_derivs(const Vector &beta, Vector *g, SpdMatrix *h) {
  if ((int)beta.size() != xdim()) report_error("log_likelihood: wrong size beta");
  if (allreduce()) {
    return loglike_through_hook(*this, [&](boomgpu_ctx *ctx, double *suf_dev) {
      return boomgpu_poisson_loglike_derivs_device(ctx, beta.data(), suf_dev); }, g, h);
  }
  DeviceData &dev(device_data());
  double ans = 0;
  if (g) g->assign(xdim(), 0.0);
  if (h && h->dim != xdim()) *h = SpdMatrix(xdim());
  dev.check(boomgpu_poisson_loglike_derivs(dev.ctx(), beta.data(), &ans, g ? g->data() : nullptr, h ? h->a.data() : nullptr));
  return ans;
}

// ---------------------------------------------------------------------------------------------
void WeightedRegSuf::clear() {
  std::fill(xtx_.a.begin(), xtx_.a.end(), 0.0);
  std::fill(xty_.begin(), xty_.end(), 0.0);
  n_ = yty_ = sumw_ = sumlogw_ = 0;
}
void WeightedRegSuf::update(const Vector &x, double weighted_value, double weight) {
  const int p = xtx_.dim;
  if ((int)x.size() != p) report_error("sufficient statistics: wrong size x");
  for (int i = 0; i < p; ++i) {
    const double a = weight * x[i];
    for (int j = 0; j < p; ++j) xtx_.a[(size_t)i * p + j] += a * x[j];
    xty_[i] += x[i] * weighted_value;
  }
  n_ += 1;
}
void WeightedRegSuf::add_data(const Vector &x, double y, double w) {
  update(x, w * y, w);
  yty_ += w * y * y;
  sumw_ += w;
  sumlogw_ += std::log(w);
}
WeightedRegSuf::~WeightedRegSuf() { unpin(); }
void WeightedRegSuf::unpin() {
  if (pinned_) { boomgpu_unpin_host(pinned_); pinned_ = nullptr; }
}
WeightedRegSuf &WeightedRegSuf::operator=(const WeightedRegSuf &rhs) {
  if (this != &rhs) {
    if (rhs.xtx_.a.size() != xtx_.a.size()) unpin();   // the assignment below may reallocate
    xtx_ = rhs.xtx_; xty_ = rhs.xty_;
    n_ = rhs.n_; yty_ = rhs.yty_; sumw_ = rhs.sumw_; sumlogw_ = rhs.sumlogw_;
    if (pinned_ && pinned_ != xtx_.a.data()) unpin();
  }
  return *this;
}
double *WeightedRegSuf::xtx_storage(int p) {
  if (xtx_.dim != p) { unpin(); xtx_ = SpdMatrix(p); xty_.assign(p, 0.0); }
  // from 1 MB up the device->host copy of the matrix lands here directly: page-lock the storage once
  const size_t bytes = xtx_.a.size() * sizeof(double);
  if (!pinned_ && bytes >= ((size_t)1 << 20) && boomgpu_pin_host(xtx_.a.data(), bytes) == 0) pinned_ = xtx_.a.data();
  return xtx_.a.data();
}
void WeightedRegSuf::reset(const double *packed, int p) {
  if (xtx_.dim != p) { unpin(); xtx_ = SpdMatrix(p); xty_.assign(p, 0.0); }
  std::copy(packed, packed + (size_t)p * p, xtx_.a.begin());
  std::copy(packed + (size_t)p * p, packed + (size_t)p * p + p, xty_.begin());
  const double *sc = packed + (size_t)p * p + p;
  n_ = sc[0]; yty_ = sc[1]; sumw_ = sc[2]; sumlogw_ = sc[3];
}

// ---------------------------------------------------------------------------------------------
StatView::StatView(const WeightedRegSuf &full) : full_(full.xtx().a.data()), p_(full.xtx().dim), xty_(full.xty().data()) {}

StatView::StatView(int p, const std::vector<int> &cols, const Vector &G, const Vector &diag, const Vector &xty,
                   std::function<void(int, double *)> fetch)
    : p_(p), k_((int)cols.size()), where_(p, -1), diag_(diag), xty_(xty.data()), fetch_(std::move(fetch)), column_(p) {
  ld_ = k_ + 16;                       // room for the columns a sweep may add
  G_.assign((size_t)p * ld_, 0.0);
  for (int j = 0; j < p; ++j) std::copy(G.begin() + (size_t)j * k_, G.begin() + (size_t)(j + 1) * k_, G_.begin() + (size_t)j * ld_);
  for (int a = 0; a < k_; ++a) where_[cols[a]] = a;
}

void StatView::ensure_column(int j) {
  if (full_ || where_[j] >= 0) return;
  if (!fetch_) report_error("StatView: column not held and no way to fetch it");
  fetch_(j, column_.data());
  ++fetched_;
  if (k_ == ld_) {                     // grow the row pitch
    const int nld = ld_ + 32;
    Vector G((size_t)p_ * nld, 0.0);
    for (int r = 0; r < p_; ++r) std::copy(G_.begin() + (size_t)r * ld_, G_.begin() + (size_t)r * ld_ + k_, G.begin() + (size_t)r * nld);
    G_.swap(G);
    ld_ = nld;
  }
  for (int r = 0; r < p_; ++r) G_[(size_t)r * ld_ + k_] = column_[r];
  where_[j] = k_++;
}

double StatView::at(int i, int j) {
  if (full_) return full_[(size_t)i * p_ + j];
  if (i == j) return diag_[i];
  int a = where_[j];
  if (a >= 0) return G_[(size_t)i * ld_ + a];
  a = where_[i];
  if (a >= 0) return G_[(size_t)j * ld_ + a];   // symmetric
  ensure_column(j);
  return G_[(size_t)i * ld_ + where_[j]];
}

SpikeSlabCore::SpikeSlabCore(const std::shared_ptr<MvnBase> &slab, const std::shared_ptr<VariableSelectionPrior> &spike,
                             bool fisher_yates)
    : slab_(slab), spike_(spike), fisher_yates_(fisher_yates) {
  if (!slab || !spike) report_error("spike and slab priors must not be null");
  if (slab->dim() != spike->potential_nvars()) report_error("Slab and spike dimensions differ.");
}

// BinomialLogitSpikeSlabSampler::log_model_prob (.cpp:88-117) == SpikeSlabSampler::log_model_prob with sigsq = 1
double SpikeSlabCore::log_model_prob(const Selector &g, const WeightedRegSuf &suf) const {
  StatView v(suf);
  return log_model_prob(g, v);
}
double SpikeSlabCore::log_model_prob(const Selector &g, StatView &stats) const {
  const double neg_inf = -std::numeric_limits<double>::infinity();
  double num = spike_->logp(g);
  if (num == neg_inf || g.nvars() == 0) return num;
  SpdMatrix ivar = g.select(slab_->siginv());
  bool ok = true;
  num += .5 * logdet_spd(ivar, &ok);
  if (!ok || num == neg_inf) return neg_inf;
  const int k = ivar.dim;
  Vector mu = g.select(slab_->mu());
  Vector ivar_mu(k, 0.0);
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) ivar_mu[i] += ivar(i, j) * mu[j];
  double q = 0;
  for (int i = 0; i < k; ++i) q += mu[i] * ivar_mu[i];
  num -= .5 * q;
  const std::vector<int> pos = g.included_positions();
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) ivar.a[(size_t)i * k + j] += stats.at(pos[i], pos[j]);
  if (!cholesky_lower(ivar.a.data(), k)) return neg_inf;
  double denom = 0;
  for (int i = 0; i < k; ++i) denom += std::log(ivar(i, i));  // = .5 log |ivar|
  Vector S(k);
  for (int i = 0; i < k; ++i) S[i] = stats.xty(pos[i]) + ivar_mu[i];
  lsolve_inplace(ivar.a.data(), k, S.data());
  double nsq = 0;
  for (int i = 0; i < k; ++i) nsq += S[i] * S[i];
  denom -= .5 * nsq;
  return num - denom;
}

// log_model_prob along a path of single flips.  The value is the one log_model_prob() computes from
// scratch (the posterior only depends on the SET of included variables), but an "add j" proposal -- all
// but a handful of the p proposals of a sweep at a sparse model -- is evaluated by bordering the two
// Cholesky factors of the current model with one row: O(k^2) instead of two O(k^3) factorisations and
// a handful of heap allocations.  Included variables are kept in insertion order.
class SpikeSlabCore::FlipEvaluator {
 public:
  FlipEvaluator(const SpikeSlabCore &core, StatView &stats, const Selector &g)
      : slab_mu_(core.slab_->mu()), siginv_(core.slab_->siginv()), spike_(*core.spike_), stats_(stats),
        g_(g), p_(g.nvars_possible()) {
    pos_ = g.included_positions();
    reserve(std::min(p_, std::max(32, 2 * (int)pos_.size() + 8)));
    logp_ = rebuild();
  }
  double logp() const { return logp_; }
  const Selector &selector() const { return g_; }

  // log_model_prob of the current model with variable j flipped; commit() makes it current
  double propose(int j) {
    prop_ = j;
    const double neg_inf = -std::numeric_limits<double>::infinity();
    const int k_new = (int)pos_.size() + (g_[j] ? -1 : 1);
    prop_spike_ = (spike_.max_model_size() >= 0 && k_new > spike_.max_model_size()) ? neg_inf
                                                                                   : spike_logp_ + spike_.flip_delta(j, g_[j]);
    if (g_[j]) {  // drop: evaluated from scratch on the smaller model
      prop_is_add_ = false;
      return prop_logp_ = neg_inf == prop_spike_ ? neg_inf : scratch_without(j);
    }
    prop_is_add_ = true;
    if (prop_spike_ == neg_inf) return prop_logp_ = neg_inf;
    const int k = (int)pos_.size();
    const double *sj = siginv_.a.data() + (size_t)j * p_;
    // border rows: l1 = Lp^-1 Siginv[gamma, j], l2 = Lq^-1 (Siginv + XtX)[gamma, j]
    double n1 = 0, n2 = 0, cross = 0;
    for (int i = 0; i < k; ++i) {
      const int pi = pos_[i];
      double s1 = sj[pi], s2 = sj[pi] + stats_.at(j, pi);
      const double *r1 = Lp_.data() + (size_t)i * ld_, *r2 = Lq_.data() + (size_t)i * ld_;
      for (int c = 0; c < i; ++c) { s1 -= r1[c] * l1_[c]; s2 -= r2[c] * l2_[c]; }
      l1_[i] = s1 / r1[i]; l2_[i] = s2 / r2[i];
      n1 += l1_[i] * l1_[i]; n2 += l2_[i] * l2_[i];
      cross += sj[pi] * slab_mu_[pi];   // Siginv[j, gamma] mu_gamma
    }
    const double d1 = sj[j] - n1, d2 = sj[j] + stats_.at(j, j) - n2;
    if (!(d1 > 0) || !(d2 > 0)) return prop_logp_ = neg_inf;
    d1_ = std::sqrt(d1); d2_ = std::sqrt(d2);
    const double mj = slab_mu_[j];
    const double q_new = q_ + 2 * mj * cross + sj[j] * mj * mj;
    double nsq;
    if (mj == 0.0) {  // b_gamma unchanged: only the last forward-substitution row is new
      bj_ = stats_.xty(j) + cross;
      double t = bj_;
      for (int c = 0; c < k; ++c) t -= l2_[c] * u_[c];
      uj_ = t / d2_;
      nsq = usq_ + uj_ * uj_;
    } else {  // every entry of b changes: b_i += Siginv[i, j] mu_j, then a full forward solve
      bj_ = stats_.xty(j) + cross + sj[j] * mj;
      nsq = 0;
      for (int i = 0; i < k; ++i) {
        double t = b_[i] + sj[pos_[i]] * mj;
        const double *r2 = Lq_.data() + (size_t)i * ld_;
        for (int c = 0; c < i; ++c) t -= r2[c] * tmp_[c];
        tmp_[i] = t / r2[i];
        nsq += tmp_[i] * tmp_[i];
      }
      double t = bj_;
      for (int c = 0; c < k; ++c) t -= l2_[c] * tmp_[c];
      uj_ = t / d2_;
      nsq += uj_ * uj_;
    }
    prop_q_ = q_new; prop_usq_ = nsq;
    prop_hldp_ = hldp_ + std::log(d1_); prop_hldq_ = hldq_ + std::log(d2_);
    return prop_logp_ = prop_spike_ + prop_hldp_ - .5 * q_new - (prop_hldq_ - .5 * nsq);
  }

  void commit() {
    const int j = prop_;
    if (!prop_is_add_) {
      g_.drop(j);
      pos_.erase(std::find(pos_.begin(), pos_.end(), j));
      logp_ = rebuild();   // also re-sums the spike prior from scratch: no drift along the path
      return;
    }
    const int k = (int)pos_.size();
    const double mj = slab_mu_[j];
    stats_.ensure_column(j);   // active-set form: later proposals read X'WX[., j]
    if (k + 1 > ld_) grow(std::min(p_, 2 * ld_));
    double *r1 = Lp_.data() + (size_t)k * ld_, *r2 = Lq_.data() + (size_t)k * ld_;
    for (int c = 0; c < k; ++c) { r1[c] = l1_[c]; r2[c] = l2_[c]; }
    r1[k] = d1_; r2[k] = d2_;
    if (mj != 0.0) {
      const double *sj = siginv_.a.data() + (size_t)j * p_;
      for (int i = 0; i < k; ++i) { b_[i] += sj[pos_[i]] * mj; u_[i] = tmp_[i]; }
    }
    b_[k] = bj_; u_[k] = uj_;
    g_.add(j);
    pos_.push_back(j);
    q_ = prop_q_; usq_ = prop_usq_; hldp_ = prop_hldp_; hldq_ = prop_hldq_;
    spike_logp_ = prop_spike_;
    logp_ = prop_logp_;
  }

 private:
  // workspace for models of up to cap variables (the factors are cap x cap, lower triangles)
  void reserve(int cap) {
    ld_ = cap;
    Lp_.resize((size_t)cap * cap); Lq_.resize((size_t)cap * cap);
    b_.resize(cap); u_.resize(cap); l1_.resize(cap); l2_.resize(cap); im_.resize(cap); tmp_.resize(cap);
  }
  void grow(int cap) {  // re-lays the factors out with the new leading dimension; pending border rows live in l1_/l2_
    const int k = (int)pos_.size(), old = ld_;
    Vector Lp(std::move(Lp_)), Lq(std::move(Lq_)), l1(l1_), l2(l2_), b(b_), u(u_), tmp(tmp_);
    Lp_.clear(); Lq_.clear();
    reserve(cap);
    for (int i = 0; i < k; ++i)
      for (int c = 0; c <= i; ++c) { Lp_[(size_t)i * ld_ + c] = Lp[(size_t)i * old + c]; Lq_[(size_t)i * ld_ + c] = Lq[(size_t)i * old + c]; }
    std::copy(l1.begin(), l1.end(), l1_.begin()); std::copy(l2.begin(), l2.end(), l2_.begin());
    std::copy(b.begin(), b.end(), b_.begin()); std::copy(u.begin(), u.end(), u_.begin());
    std::copy(tmp.begin(), tmp.end(), tmp_.begin());
  }
  // factorises the current model from scratch into the workspace; returns log_model_prob
  double rebuild() {
    const double neg_inf = -std::numeric_limits<double>::infinity();
    spike_logp_ = spike_.logp(g_);
    const int k = (int)pos_.size();
    q_ = usq_ = hldp_ = hldq_ = 0;
    if (spike_logp_ == neg_inf || k == 0) return spike_logp_;
    return factor(pos_, Lp_.data(), Lq_.data(), b_.data(), u_.data(), &q_, &usq_, &hldp_, &hldq_, spike_logp_);
  }
  double scratch_without(int j) {
    scratch_pos_.clear();
    for (int v : pos_) if (v != j) scratch_pos_.push_back(v);
    if (scratch_pos_.empty()) return prop_spike_;
    const size_t need = (size_t)ld_ * ld_;
    if (S1_.size() < need) { S1_.resize(need); S2_.resize(need); sb_.resize(ld_); su_.resize(ld_); }
    double q, usq, h1, h2;
    return factor(scratch_pos_, S1_.data(), S2_.data(), sb_.data(), su_.data(), &q, &usq, &h1, &h2, prop_spike_);
  }
  double factor(const std::vector<int> &pos, double *Lp, double *Lq, double *b, double *u, double *q, double *usq, double *hldp,
                double *hldq, double spike_logp) {
    const double neg_inf = -std::numeric_limits<double>::infinity();
    const int k = (int)pos.size();
    *q = 0;
    for (int i = 0; i < k; ++i) {
      const double *si = siginv_.a.data() + (size_t)pos[i] * p_;
      double m = 0;
      for (int c = 0; c < k; ++c) m += si[pos[c]] * slab_mu_[pos[c]];
      im_[i] = m;
      *q += m * slab_mu_[pos[i]];
      b[i] = stats_.xty(pos[i]) + m;
      for (int c = 0; c <= i; ++c) { Lp[(size_t)i * ld_ + c] = si[pos[c]]; Lq[(size_t)i * ld_ + c] = si[pos[c]] + stats_.at(pos[i], pos[c]); }
    }
    if (!chol_ld(Lp, k) || !chol_ld(Lq, k)) return neg_inf;
    *hldp = *hldq = *usq = 0;
    for (int i = 0; i < k; ++i) {
      *hldp += std::log(Lp[(size_t)i * ld_ + i]);
      *hldq += std::log(Lq[(size_t)i * ld_ + i]);
      double t = b[i];
      const double *r = Lq + (size_t)i * ld_;
      for (int c = 0; c < i; ++c) t -= r[c] * u[c];
      u[i] = t / r[i];
      *usq += u[i] * u[i];
    }
    return spike_logp + *hldp - .5 * *q - (*hldq - .5 * *usq);
  }
  // in-place lower Cholesky of the leading k x k block (lower triangle stored, leading dimension ld_)
  bool chol_ld(double *a, int k) const {
    for (int j = 0; j < k; ++j) {
      double *rj = a + (size_t)j * ld_;
      double d = rj[j];
      for (int c = 0; c < j; ++c) d -= rj[c] * rj[c];
      if (!(d > 0) || !std::isfinite(d)) return false;
      d = std::sqrt(d);
      rj[j] = d;
      for (int i = j + 1; i < k; ++i) {
        double *ri = a + (size_t)i * ld_;
        double s = ri[j];
        for (int c = 0; c < j; ++c) s -= ri[c] * rj[c];
        ri[j] = s / d;
      }
    }
    return true;
  }

  const Vector &slab_mu_;
  const SpdMatrix &siginv_;
  const VariableSelectionPrior &spike_;
  StatView &stats_;
  Selector g_;
  int p_, ld_;
  std::vector<int> pos_, scratch_pos_;
  Vector Lp_, Lq_, b_, u_, l1_, l2_, im_, tmp_, S1_, S2_, sb_, su_;
  double q_ = 0, usq_ = 0, hldp_ = 0, hldq_ = 0, spike_logp_ = 0, logp_ = 0;
  // pending proposal
  int prop_ = -1;
  bool prop_is_add_ = false;
  double prop_logp_ = 0, prop_spike_ = 0, prop_q_ = 0, prop_usq_ = 0, prop_hldp_ = 0, prop_hldq_ = 0, d1_ = 0, d2_ = 0, bj_ = 0, uj_ = 0;
};

Vector SpikeSlabCore::flip_path_log_probs(const Selector &start, const WeightedRegSuf &suf, const std::vector<int> &flips,
                                          const std::vector<bool> &accept) const {
  StatView v(suf);
  return flip_path_log_probs(start, v, flips, accept);
}
Vector SpikeSlabCore::flip_path_log_probs(const Selector &start, StatView &stats, const std::vector<int> &flips,
                                          const std::vector<bool> &accept) const {
  FlipEvaluator ev(*this, stats, start);
  Vector out;
  for (size_t i = 0; i < flips.size(); ++i) {
    out.push_back(ev.propose(flips[i]));
    if (accept[i] && std::isfinite(out.back())) ev.commit();
  }
  out.push_back(ev.logp());
  return out;
}

void SpikeSlabCore::draw_model_indicators(RNG &rng, GlmCoefs &coef, const WeightedRegSuf &suf) const {
  StatView v(suf);
  draw_model_indicators(rng, coef, v);
}
void SpikeSlabCore::draw_model_indicators(RNG &rng, GlmCoefs &coef, StatView &suf) const {
  if (!allow_model_selection_) return;
  Selector g = coef.inc();
  const int nv = g.nvars_possible();
  std::vector<int> indx(nv);
  for (int i = 0; i < nv; ++i) indx[i] = i;
  if (fisher_yates_) {  // SpikeSlabSampler.cpp:52-57
    for (int i = nv - 1; i > 0; --i) {
      int j = random_int_mt(rng, 0, i);
      if (j != i) std::swap(indx[i], indx[j]);
    }
  } else {  // BinomialLogitSpikeSlabSampler.cpp:185-188
    for (int i = 0; i < nv; ++i) {
      int j = random_int_mt(rng, 0, nv - 1);
      std::swap(indx[i], indx[j]);
    }
  }
  double logp = log_model_prob(g, suf);
  if (!std::isfinite(logp)) {
    spike_->make_valid(g);
    logp = log_model_prob(g, suf);
  }
  if (!std::isfinite(logp)) report_error("The spike and slab sampler did not start with a legal configuration.");
  int n = nv;
  if (max_flips_ > 0) n = std::min(n, max_flips_);
  FlipEvaluator ev(*this, suf, g);
  for (int i = 0; i < n; ++i) {  // mcmc_one_flip (.cpp:213-222)
    const double logp_new = ev.propose(indx[i]);
    const double u = runif_mt(rng, 0, 1);
    if (std::log(u) <= logp_new - ev.logp()) ev.commit();
  }
  coef.set_inc(ev.selector());
}

// BinomialLogitSpikeSlabSampler::draw_beta (.cpp:56-75)
void SpikeSlabCore::draw_beta(RNG &rng, GlmCoefs &coef, const WeightedRegSuf &suf) const {
  StatView v(suf);
  draw_beta(rng, coef, v);
}
void SpikeSlabCore::draw_beta(RNG &rng, GlmCoefs &coef, StatView &suf) const {
  const Selector &g(coef.inc());
  if (g.nvars() == 0) { coef.drop_all(); return; }
  SpdMatrix precision = g.select(slab_->siginv());
  const int k = precision.dim;
  Vector mu = g.select(slab_->mu());
  Vector scaled_mean(k, 0.0);
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) scaled_mean[i] += precision(i, j) * mu[j];
  const std::vector<int> pos = g.included_positions();
  for (int i = 0; i < k; ++i) {
    for (int j = 0; j < k; ++j) precision.a[(size_t)i * k + j] += suf.at(pos[i], pos[j]);
    scaled_mean[i] += suf.xty(pos[i]);
  }
  if (!cholesky_lower(precision.a.data(), k)) report_error("Cholesky decomposition failed in draw_beta.");
  lsolve_inplace(precision.a.data(), k, scaled_mean.data());
  ltsolve_inplace(precision.a.data(), k, scaled_mean.data());  // posterior mean
  Vector z(k);
  for (int i = 0; i < k; ++i) z[i] = rnorm_mt(rng, 0, 1);
  ltsolve_inplace(precision.a.data(), k, z.data());  // rmvn_precision_upper_cholesky_mt (mvn.cpp:114-122)
  for (int i = 0; i < k; ++i) z[i] += scaled_mean[i];
  coef.set_included_coefficients(z);
}

bool SpikeSlabCore::find_posterior_mode(GlmModelBase &model, double epsilon, double *log_posterior_at_mode) const {
  const double neg_inf = -std::numeric_limits<double>::infinity();
  *log_posterior_at_mode = neg_inf;
  GlmCoefs &coef(model.coef());
  const Selector g = coef.inc();
  const int k = g.nvars(), p = g.nvars_possible();
  if (k == 0) return false;   // the reference declines the empty model as well (.cpp:154-159)
  const std::vector<int> pos = g.included_positions();
  const MvnModel slab_g(g.select(slab_->mu()), g.select(slab_->siginv()), true);
  Vector beta = coef.included_coefficients();
  Vector grad_full;
  SpdMatrix hess_full;
  // objective, gradient and NEGATIVE Hessian on the included coordinates
  auto evaluate = [&](const Vector &b, Vector *grad, Vector *neg_hess) {
    const Vector full = g.expand(b);
    double f = model.log_likelihood_derivs(full, grad ? &grad_full : nullptr, neg_hess ? &hess_full : nullptr);
    f += slab_g.logp(b);
    if (grad) {
      grad->assign(k, 0.0);
      for (int i = 0; i < k; ++i) {
        double s = 0;
        for (int j = 0; j < k; ++j) s += slab_g.siginv()(i, j) * (b[j] - slab_g.mu()[j]);
        (*grad)[i] = grad_full[pos[i]] - s;
      }
    }
    if (neg_hess) {
      neg_hess->assign((size_t)k * k, 0.0);
      for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) (*neg_hess)[(size_t)i * k + j] = slab_g.siginv()(i, j) - hess_full.a[(size_t)pos[i] * p + pos[j]];
    }
    return f;
  };
  Vector grad, nh;
  double f = evaluate(beta, &grad, &nh);
  if (!std::isfinite(f)) return false;
  for (int iter = 0; iter < 100; ++iter) {
    if (!cholesky_lower(nh.data(), k)) return false;   // the log posterior is concave: -H is positive definite
    Vector step(grad);
    lsolve_inplace(nh.data(), k, step.data());
    ltsolve_inplace(nh.data(), k, step.data());        // step = (-H)^-1 grad
    double decrement = 0;
    for (int i = 0; i < k; ++i) decrement += step[i] * grad[i];
    double scale = 1.0, f_new = neg_inf;
    Vector cand(k), grad_new, nh_new;
    for (int half = 0; half < 40; ++half, scale *= 0.5) {
      for (int i = 0; i < k; ++i) cand[i] = beta[i] + scale * step[i];
      f_new = evaluate(cand, &grad_new, &nh_new);
      if (std::isfinite(f_new) && f_new >= f - 1e-12 * std::fabs(f)) break;
    }
    if (!std::isfinite(f_new) || f_new < f - 1e-12 * std::fabs(f)) return false;
    const double gain = f_new - f;
    beta = cand; f = f_new; grad = grad_new; nh = nh_new;
    if (gain < epsilon && decrement < 2 * epsilon) {
      *log_posterior_at_mode = f;
      coef.set_included_coefficients(beta);
      return true;
    }
  }
  return false;
}

void SpikeSlabCore::set_spike(const std::shared_ptr<VariableSelectionPrior> &spike) {
  if (!spike || spike->potential_nvars() != slab_->dim()) report_error("Spike does not match model dimension.");
  spike_ = spike;
}
void SpikeSlabCore::set_slab(const std::shared_ptr<MvnBase> &slab) {
  if (!slab || slab->dim() != spike_->potential_nvars()) report_error("Slab does not match model dimension.");
  slab_ = slab;
}

double SpikeSlabCore::logpri(const GlmCoefs &coef) const {
  const Selector &g(coef.inc());
  double ans = spike_->logp(g);
  if (!std::isfinite(ans)) return ans;
  if (g.nvars() > 0) {
    MvnModel sub(g.select(slab_->mu()), g.select(slab_->siginv()), true);
    ans += sub.logp(coef.included_coefficients());
  }
  return ans;
}

// ---------------------------------------------------------------------------------------------
namespace {
struct LogitMixtureStore {
  Vector mu = Vector(9, 0.0);
  // Fruhwirth-Schnatter & Fruhwirth scale mixture for the logistic (values as in
  // Models/Glm/PosteriorSamplers/NormalMixtureApproximation.cpp:416-424)
  Vector sigma = {0.88437229872213, 1.16097607474416, 1.28021991084306, 1.3592552924727, 1.67589879794907,
                  2.20287232043947, 2.20507148325819, 2.91944313615144, 3.90807611741308};
  Vector weights = {0.038483985581272, 0.13389889791451, 0.0657842076622429, 0.105680086433879, 0.345939491553619,
                    0.0442261124345564, 0.193289780660134, 0.068173066865908, 0.00452437089387876};
};
LogitMixtureStore &logit_mixture_store() { static LogitMixtureStore s; return s; }

// PoissonDataImputer::mixture_table_ (PoissonDataImputer.hpp:99): one table per process, grown on demand, plus its
// flattened (CSR) form for boomgpu_set_poisson_table
struct PoissonTableStore {
  NormalMixtureApproximationTable table;
  std::vector<int64_t> nu;
  std::vector<int32_t> offset;
  Vector weights, mu, sigma;
  int64_t largest = 0;
  uint64_t version = 0;
  std::mutex lock;
  void flatten() {
    nu = table.index();
    offset.assign(1, 0);
    weights.clear(); mu.clear(); sigma.clear();
    for (size_t e = 0; e < table.size(); ++e) {
      const NormalMixtureApproximation &a(table.entry(e));
      weights.insert(weights.end(), a.weights.begin(), a.weights.end());
      mu.insert(mu.end(), a.mu.begin(), a.mu.end());
      sigma.insert(sigma.end(), a.sigma.begin(), a.sigma.end());
      offset.push_back(offset.back() + a.dim());
    }
  }
};
PoissonTableStore &poisson_table_store() { static PoissonTableStore s; return s; }

// device step -> all-reduce -> packed statistics [p*p | p | 4] on the host.
// With a caller-supplied all-reduce hook the statistics stay on the device for it (step_device + hook + download);
// otherwise the synchronous C-ABI step does everything (native NCCL when a communicator is attached, and for p <= 64 on
// one GPU the reduction writes straight into pinned host memory).
template <class StepDeviceFn, class StepSyncFn>
void run_device_step(GlmModelBase &model, WeightedRegSuf &suf, Vector &packed, StepDeviceFn step_device, StepSyncFn step_sync) {
  DeviceData &dev(model.device_data());
  const int p = model.xdim();
  const int64_t len = boomgpu_suf_len(p);
  if (model.allreduce()) {
    packed.resize((size_t)len);
    double *suf_dev = nullptr;
    dev.check(boomgpu_suf_buffer(dev.ctx(), &suf_dev));
    dev.check(step_device(dev.ctx(), suf_dev));
    model.allreduce()(suf_dev, len);
    dev.check(boomgpu_download(dev.ctx(), suf_dev, packed.data(), len));
    suf.reset(packed.data(), p);
  } else {
    double scalars[4] = {0, 0, 0, 0};
    dev.check(step_sync(dev.ctx(), suf.xtx_storage(p), suf.xty_storage(), scalars));   // straight into the statistics object
    suf.set_scalars(scalars[0], scalars[1], scalars[2], scalars[3]);
  }
}
}  // namespace

void set_logit_mixture(const Vector &mu, const Vector &sigma, const Vector &weights) {
  if (sigma.empty() || sigma.size() != weights.size() || mu.size() != sigma.size())
    report_error("set_logit_mixture: mu, sigma and weights must have the same positive length");
  LogitMixtureStore &s(logit_mixture_store());
  s.mu = mu; s.sigma = sigma; s.weights = weights;
}

// ---------------------------------------------------------------------------------------------
BinomialLogitAuxmixSampler::BinomialLogitAuxmixSampler(BinomialLogitModel *model, const std::shared_ptr<MvnBase> &prior,
                                                       int clt_threshold, RNG &seeding_rng)
    : PosteriorSampler(seeding_rng), model_(model), prior_(prior), suf_(model ? model->xdim() : 0),
      clt_threshold_(clt_threshold) {
  if (!model) report_error("BinomialLogitAuxmixSampler: null model");
  if (!prior || prior->dim() != model->xdim()) report_error("Prior does not match model dimension.");
  device_seed_ = seed_rng(rng());
}

void BinomialLogitAuxmixSampler::on_seed() { device_seed_ = seed_rng(rng()); iteration_ = 0; }

double BinomialLogitAuxmixSampler::logpri() const { return prior_->logp(model_->Beta()); }

void BinomialLogitAuxmixSampler::draw() {
  impute_latent_data();
  draw_params();
}

void BinomialLogitAuxmixSampler::impute_latent_data() {
  if (latent_data_fixed_) return;  // statistics are under external control (Imputer.hpp:282-299)
  active_.valid = false;
  const LogitMixtureStore &mix(logit_mixture_store());
  const int clt = clt_threshold_;
  const uint64_t seed = device_seed_, it = iteration_++;
  const Vector &beta(model_->Beta());
  run_device_step(
      *model_, suf_, packed_,
      [&](boomgpu_ctx *ctx, double *suf_dev) {
        int rc = boomgpu_set_logit_mixture(ctx, (int)mix.sigma.size(), mix.mu.data(), mix.sigma.data(), mix.weights.data());
        if (rc) return rc;
        return boomgpu_logit_step_device(ctx, beta.data(), clt, seed, it, suf_dev);
      },
      [&](boomgpu_ctx *ctx, double *xtx, double *xty, double *scalars) {
        int rc = boomgpu_set_logit_mixture(ctx, (int)mix.sigma.size(), mix.mu.data(), mix.sigma.data(), mix.weights.data());
        if (rc) return rc;
        int64_t ss = 0;
        rc = boomgpu_logit_step(ctx, beta.data(), clt, seed, it, xtx, xty, &ss);
        scalars[0] = (double)ss; scalars[1] = scalars[2] = scalars[3] = 0.0;
        return rc;
      });
}

bool BinomialLogitAuxmixSampler::impute_latent_data_active(const std::vector<int> &cols_in) {
  const int p = model_->xdim();
  if (latent_data_fixed_ || p <= 64 || model_->allreduce() || cols_in.size() > 128) return false;
  std::vector<int> cols(cols_in);
  if (cols.empty()) cols.push_back(0);
  const LogitMixtureStore &mix(logit_mixture_store());
  DeviceData &dev(model_->device_data());
  dev.check(boomgpu_set_logit_mixture(dev.ctx(), (int)mix.sigma.size(), mix.mu.data(), mix.sigma.data(), mix.weights.data()));
  const int k = (int)cols.size();
  active_.cols = cols;
  active_.G.resize((size_t)p * k); active_.diag.resize(p); active_.xty.resize(p);
  int64_t ss = 0;
  std::vector<int32_t> c32(cols.begin(), cols.end());
  dev.check(boomgpu_logit_step_active(dev.ctx(), model_->Beta().data(), clt_threshold_, device_seed_, iteration_++, c32.data(), k,
                                      active_.G.data(), active_.diag.data(), active_.xty.data(), &ss));
  active_.scalars[0] = (double)ss; active_.scalars[1] = active_.scalars[2] = active_.scalars[3] = 0;
  active_.valid = true;
  return true;
}

std::unique_ptr<StatView> BinomialLogitAuxmixSampler::statistics_view() {
  if (!active_.valid) return std::unique_ptr<StatView>(new StatView(suf_));
  DeviceData *dev = &model_->device_data();
  ActiveSetState *st = &active_;
  return std::unique_ptr<StatView>(new StatView(model_->xdim(), active_.cols, active_.G, active_.diag, active_.xty,
                                               [dev, st](int j, double *out) {
                                                 dev->check(boomgpu_weighted_column(dev->ctx(), j, out));
                                                 ++st->columns_fetched;
                                               }));
}

void BinomialLogitAuxmixSampler::materialize_full_statistics() const {
  DeviceData &dev(model_->device_data());
  const int p = model_->xdim();
  dev.check(boomgpu_full_statistics(dev.ctx(), suf_.xtx_storage(p), suf_.xty_storage()));
  suf_.set_scalars(active_.scalars[0], active_.scalars[1], active_.scalars[2], active_.scalars[3]);
  active_.valid = false;
}

void BinomialLogitAuxmixSampler::draw_params() {
  if (active_.valid) materialize_full_statistics();
  const int p = model_->xdim();
  SpdMatrix ivar(prior_->siginv());
  Vector ivar_mu(suf_.xty());
  for (int i = 0; i < p; ++i) {
    double s = 0;
    for (int j = 0; j < p; ++j) {
      ivar.a[(size_t)i * p + j] += suf_.xtx()(i, j);
      s += prior_->siginv()(i, j) * prior_->mu()[j];
    }
    ivar_mu[i] += s;
  }
  model_->set_Beta(rmvn_suf_mt(rng(), ivar, ivar_mu));
}

BinomialLogitSpikeSlabSampler::BinomialLogitSpikeSlabSampler(BinomialLogitModel *model, const std::shared_ptr<MvnBase> &slab,
                                                             const std::shared_ptr<VariableSelectionPrior> &spike,
                                                             int clt_threshold, RNG &seeding_rng)
    : BinomialLogitAuxmixSampler(model, slab, clt_threshold, seeding_rng), core_(slab, spike, false) {
  if (spike->potential_nvars() != model->xdim()) report_error("Spike does not match model dimension.");
}

void BinomialLogitSpikeSlabSampler::draw() {
  if (!(active_.enabled && impute_latent_data_active(model_->coef().inc().included_positions()))) impute_latent_data();
  if (core_.model_selection_allowed()) draw_model_indicators();
  draw_beta();
}
double BinomialLogitSpikeSlabSampler::logpri() const { return core_.logpri(model_->coef()); }
void BinomialLogitSpikeSlabSampler::draw_model_indicators() {
  std::unique_ptr<StatView> v(statistics_view());
  core_.draw_model_indicators(rng(), model_->coef(), *v);
  if (active_.valid) {   // the columns the sweep fetched stay with the state: draw_beta reads them next
    // (the view copied the active arrays; re-run of statistics_view() for draw_beta would refetch, so keep what it holds)
    kept_view_ = std::move(v);
  }
}
void BinomialLogitSpikeSlabSampler::draw_beta() {
  if (active_.valid && kept_view_) { core_.draw_beta(rng(), model_->coef(), *kept_view_); kept_view_.reset(); return; }
  std::unique_ptr<StatView> v(statistics_view());
  core_.draw_beta(rng(), model_->coef(), *v);
}
double BinomialLogitSpikeSlabSampler::log_model_prob(const Selector &g) const { return core_.log_model_prob(g, suf()); }
std::shared_ptr<BinomialLogitSpikeSlabSampler> BinomialLogitSpikeSlabSampler::clone_to_new_host(BinomialLogitModel *new_host) const {
  auto s = std::make_shared<BinomialLogitSpikeSlabSampler>(new_host, core_.slab(), core_.spike(), clt_threshold(), rng());
  s->allow_model_selection(core_.model_selection_allowed());
  s->limit_model_selection(core_.max_flips());
  return s;
}
void BinomialLogitSpikeSlabSampler::find_posterior_mode(double epsilon) {
  posterior_mode_found_ = core_.find_posterior_mode(*model_, epsilon, &log_posterior_at_mode_);
}

// ---------------------------------------------------------------------------------------------
void PoissonRegressionAuxMixSampler::set_mixture_table(const Vector &ser, int64_t largest_index) {
  NormalMixtureApproximationTable t;
  t.deserialize(ser);
  if (t.empty()) report_error("set_mixture_table: empty table");
  PoissonTableStore &st(poisson_table_store());
  std::lock_guard<std::mutex> guard(st.lock);
  st.table = t;
  st.largest = largest_index;
  st.flatten();
  ++st.version;
}
Vector PoissonRegressionAuxMixSampler::mixture_table() {
  PoissonTableStore &st(poisson_table_store());
  std::lock_guard<std::mutex> guard(st.lock);
  return st.table.serialize();
}
NormalMixtureApproximation PoissonRegressionAuxMixSampler::approximate(int64_t nu) {
  PoissonTableStore &st(poisson_table_store());
  std::lock_guard<std::mutex> guard(st.lock);
  if (st.table.empty()) report_error("PoissonRegressionAuxMixSampler: no mixture table");
  const size_t before = st.table.size();
  NormalMixtureApproximation a = st.table.approximate(nu);
  if (st.table.size() != before) { st.flatten(); ++st.version; }
  return a;
}
bool PoissonRegressionAuxMixSampler::mixture_table_is_set() { return !poisson_table_store().table.empty(); }

PoissonRegressionAuxMixSampler::PoissonRegressionAuxMixSampler(PoissonRegressionModel *model, const std::shared_ptr<MvnBase> &prior,
                                                               int, RNG &seeding_rng)
    : PosteriorSampler(seeding_rng), model_(model), prior_(prior), suf_(model ? model->xdim() : 0) {
  if (!model) report_error("PoissonRegressionAuxMixSampler: null model");
  if (!prior || prior->dim() != model->xdim()) report_error("Prior does not match model dimension.");
  device_seed_ = seed_rng(rng());
}
void PoissonRegressionAuxMixSampler::on_seed() { device_seed_ = seed_rng(rng()); iteration_ = 0; }
double PoissonRegressionAuxMixSampler::logpri() const { return prior_->logp(model_->Beta()); }
void PoissonRegressionAuxMixSampler::draw() {
  impute_latent_data();
  draw_beta_given_complete_data();
}
// The reference extends its table per observation inside the draw (NormalMixtureApproximationTable::approximate,
// NormalMixtureApproximation.cpp:472-532).  Here: once per (data, table, context), ask the device which counts occur,
// add the entries the grid lacks by the same rule, then state the table (the C side skips an unchanged upload).
int PoissonRegressionAuxMixSampler::ensure_table(boomgpu_ctx *ctx) {
  PoissonTableStore &t(poisson_table_store());
  if (t.nu.empty())
    report_error("PoissonRegressionAuxMixSampler: no mixture table; call set_mixture_table with the serialized "
                 "NormalMixtureApproximationTable (see boom_b200/data/poisson_mixture_table.json)");
  std::lock_guard<std::mutex> guard(t.lock);
  if (counts_checked_ctx_ != ctx || counts_checked_data_version_ != model_->data_version() || counts_checked_table_version_ != t.version) {
    std::vector<unsigned char> present((size_t)std::max<int64_t>(t.largest, 1), 0);
    if (int rc = boomgpu_poisson_counts_present(ctx, present.data(), (int64_t)present.size())) return rc;
    const size_t before = t.table.size();
    for (int64_t v = std::max<int64_t>(1, t.table.smallest_index()); v < (int64_t)present.size() && v < t.table.largest_index(); ++v)
      if (present[(size_t)v] && !t.table.contains(v)) t.table.approximate(v);
    if (t.table.size() != before) { t.flatten(); ++t.version; }
    counts_checked_ctx_ = ctx; counts_checked_data_version_ = model_->data_version(); counts_checked_table_version_ = t.version;
  }
  return boomgpu_set_poisson_table(ctx, (int)t.nu.size(), t.nu.data(), t.offset.data(), t.weights.data(), t.mu.data(),
                                   t.sigma.data(), t.largest);
}

void PoissonRegressionAuxMixSampler::impute_latent_data() {
  if (latent_data_fixed_) return;
  active_.valid = false;
  const uint64_t seed = device_seed_, it = iteration_++;
  const Vector &beta(model_->Beta());
  run_device_step(
      *model_, suf_, packed_,
      [&](boomgpu_ctx *ctx, double *suf_dev) {
        if (int rc = ensure_table(ctx)) return rc;
        return boomgpu_poisson_step_device(ctx, beta.data(), seed, it, suf_dev);
      },
      [&](boomgpu_ctx *ctx, double *xtwx, double *xtwy, double *scalars) {
        if (int rc = ensure_table(ctx)) return rc;
        return boomgpu_poisson_step(ctx, beta.data(), seed, it, xtwx, xtwy, scalars);
      });
}

bool PoissonRegressionAuxMixSampler::impute_latent_data_active(const std::vector<int> &cols_in) {
  const int p = model_->xdim();
  if (latent_data_fixed_ || p <= 64 || model_->allreduce() || cols_in.size() > 128) return false;
  std::vector<int> cols(cols_in);
  if (cols.empty()) cols.push_back(0);
  DeviceData &dev(model_->device_data());
  dev.check(ensure_table(dev.ctx()));
  const int k = (int)cols.size();
  active_.cols = cols;
  active_.G.resize((size_t)p * k); active_.diag.resize(p); active_.xty.resize(p);
  std::vector<int32_t> c32(cols.begin(), cols.end());
  dev.check(boomgpu_poisson_step_active(dev.ctx(), model_->Beta().data(), device_seed_, iteration_++, c32.data(), k,
                                        active_.G.data(), active_.diag.data(), active_.xty.data(), active_.scalars));
  active_.valid = true;
  return true;
}

std::unique_ptr<StatView> PoissonRegressionAuxMixSampler::statistics_view() {
  if (!active_.valid) return std::unique_ptr<StatView>(new StatView(suf_));
  DeviceData *dev = &model_->device_data();
  ActiveSetState *st = &active_;
  return std::unique_ptr<StatView>(new StatView(model_->xdim(), active_.cols, active_.G, active_.diag, active_.xty,
                                               [dev, st](int j, double *out) {
                                                 dev->check(boomgpu_weighted_column(dev->ctx(), j, out));
                                                 ++st->columns_fetched;
                                               }));
}

void PoissonRegressionAuxMixSampler::materialize_full_statistics() const {
  DeviceData &dev(model_->device_data());
  const int p = model_->xdim();
  dev.check(boomgpu_full_statistics(dev.ctx(), suf_.xtx_storage(p), suf_.xty_storage()));
  suf_.set_scalars(active_.scalars[0], active_.scalars[1], active_.scalars[2], active_.scalars[3]);
  active_.valid = false;
}

void PoissonRegressionAuxMixSampler::draw_beta_given_complete_data() {
  if (active_.valid) materialize_full_statistics();
  const int p = model_->xdim();
  SpdMatrix ivar(prior_->siginv());
  Vector ivar_mu(suf_.xty());
  for (int i = 0; i < p; ++i) {
    double s = 0;
    for (int j = 0; j < p; ++j) {
      ivar.a[(size_t)i * p + j] += suf_.xtx()(i, j);
      s += prior_->siginv()(i, j) * prior_->mu()[j];
    }
    ivar_mu[i] += s;
  }
  model_->set_Beta(rmvn_suf_mt(rng(), ivar, ivar_mu));
}

PoissonRegressionSpikeSlabSampler::PoissonRegressionSpikeSlabSampler(PoissonRegressionModel *model, const std::shared_ptr<MvnBase> &slab,
                                                                     const std::shared_ptr<VariableSelectionPrior> &spike,
                                                                     int number_of_threads, RNG &seeding_rng)
    : PoissonRegressionAuxMixSampler(model, slab, number_of_threads, seeding_rng), core_(slab, spike, true) {
  if (spike->potential_nvars() != model->xdim()) report_error("Spike does not match model dimension.");
}
void PoissonRegressionSpikeSlabSampler::draw() {
  if (!(active_.enabled && impute_latent_data_active(model_->coef().inc().included_positions()))) impute_latent_data();
  draw_model_indicators();
  draw_beta();
}
void PoissonRegressionSpikeSlabSampler::draw_model_indicators() {
  std::unique_ptr<StatView> v(statistics_view());
  core_.draw_model_indicators(rng(), model_->coef(), *v);
  if (active_.valid) kept_view_ = std::move(v);   // the columns the sweep fetched: draw_beta reads them next
}
void PoissonRegressionSpikeSlabSampler::draw_beta() {
  if (active_.valid && kept_view_) { core_.draw_beta(rng(), model_->coef(), *kept_view_); kept_view_.reset(); return; }
  std::unique_ptr<StatView> v(statistics_view());
  core_.draw_beta(rng(), model_->coef(), *v);
}
double PoissonRegressionSpikeSlabSampler::log_model_prob(const Selector &g) const {
  return core_.log_model_prob(g, complete_data_sufficient_statistics());
}
double PoissonRegressionSpikeSlabSampler::logpri() const { return core_.logpri(model_->coef()); }
std::shared_ptr<PoissonRegressionSpikeSlabSampler> PoissonRegressionSpikeSlabSampler::clone_to_new_host(
    PoissonRegressionModel *new_host) const {
  auto s = std::make_shared<PoissonRegressionSpikeSlabSampler>(new_host, core_.slab(), core_.spike(), 1, rng());
  s->allow_model_selection(core_.model_selection_allowed());
  s->limit_model_selection(core_.max_flips());
  return s;
}
void PoissonRegressionSpikeSlabSampler::find_posterior_mode(double epsilon) {
  core_.find_posterior_mode(*model_, epsilon, &log_posterior_at_mode_);
}

// ---------------------------------------------------------------------------------------------
BinomialProbitSpikeSlabSampler::BinomialProbitSpikeSlabSampler(BinomialProbitModel *model, const std::shared_ptr<MvnBase> &slab,
                                                               const std::shared_ptr<VariableSelectionPrior> &spike, int clt_threshold,
                                                               RNG &seeding_rng)
    : PosteriorSampler(seeding_rng), model_(model), core_(slab, spike, true), suf_(model ? model->xdim() : 0),
      clt_threshold_(clt_threshold) {
  if (!model) report_error("BinomialProbitSpikeSlabSampler: null model");
  if (slab->dim() != model->xdim() || spike->potential_nvars() != model->xdim()) report_error("Prior does not match model dimension.");
  device_seed_ = seed_rng(rng());
}
void BinomialProbitSpikeSlabSampler::on_seed() { device_seed_ = seed_rng(rng()); iteration_ = 0; }
double BinomialProbitSpikeSlabSampler::logpri() const { return core_.logpri(model_->coef()); }
void BinomialProbitSpikeSlabSampler::refresh_xtx() { xtx_data_version_ = 0; }

void BinomialProbitSpikeSlabSampler::draw() {
  impute_latent_data();
  core_.draw_model_indicators(rng(), model_->coef(), suf_);
  core_.draw_beta(rng(), model_->coef(), suf_);
}

void BinomialProbitSpikeSlabSampler::impute_latent_data() {
  DeviceData &dev(model_->device_data());
  const int p = model_->xdim();
  const bool want_xtx = xtx_data_version_ != model_->data_version();   // X'NX: once per data set
  const uint64_t seed = device_seed_, it = iteration_++;
  const Vector &beta(model_->Beta());
  if (model_->allreduce()) {
    const int64_t len = boomgpu_suf_len(p);
    packed_.resize((size_t)len);
    double *suf_dev = nullptr;
    dev.check(boomgpu_suf_buffer(dev.ctx(), &suf_dev));
    dev.check(boomgpu_probit_step_device(dev.ctx(), beta.data(), clt_threshold_, seed, it, suf_dev, want_xtx ? 0 : 1));
    const size_t mat = (size_t)p * p;
    if (want_xtx) {
      model_->allreduce()(suf_dev, len);
      dev.check(boomgpu_download(dev.ctx(), suf_dev, packed_.data(), len));
      suf_.reset(packed_.data(), p);
    } else {
      model_->allreduce()(suf_dev + mat, p + 4);
      dev.check(boomgpu_download(dev.ctx(), suf_dev + mat, packed_.data() + mat, p + 4));
      std::copy(packed_.begin() + mat, packed_.begin() + mat + p, suf_.xty_storage());
    }
    suf_.set_scalars(packed_[mat + p], 0, 0, 0);
  } else {
    int64_t ss = 0;
    dev.check(boomgpu_probit_step(dev.ctx(), beta.data(), clt_threshold_, seed, it, want_xtx ? suf_.xtx_storage(p) : nullptr,
                                  suf_.xty_storage(), &ss));
    suf_.set_scalars((double)ss, 0, 0, 0);
  }
  xtx_data_version_ = model_->data_version();
}

// =============================================================================================
// The Student-t sibling: TRegressionModel + TRegressionSampler and the small scalar samplers it owns.
// =============================================================================================
UniformModel::UniformModel(double lo, double hi) : lo_(lo), hi_(hi) {
  if (!(hi > lo)) report_error("UniformModel: hi must exceed lo");
}
double UniformModel::logp(double x) const {
  return (x >= lo_ && x <= hi_) ? -std::log(hi_ - lo_) : -std::numeric_limits<double>::infinity();
}

GammaModel::GammaModel(double a, double b) : a_(a), b_(b) {
  if (!(a > 0) || !(b > 0)) report_error("GammaModel: shape and rate must be positive");
}
double GammaModelBase::logp(double x) const {   // dgamma(x, a, b, log) with b a rate
  const double a = alpha(), b = beta();
  if (x < 0) return -std::numeric_limits<double>::infinity();
  if (x == 0) return a < 1 ? std::numeric_limits<double>::infinity() : a == 1 ? std::log(b) : -std::numeric_limits<double>::infinity();
  return a * std::log(b) - std::lgamma(a) + (a - 1.0) * std::log(x) - b * x;
}

double rgamma_mt(RNG &rng, double a, double b) {
  if (!(a > 0) || !(b > 0)) report_error("rgamma_mt: shape and rate must be positive");
  std::gamma_distribution<double> d(a, 1.0 / b);
  return d(rng.generator());
}
double rexp_mt(RNG &rng, double lambda) {
  double u = rng();
  while (u <= 0) u = rng();
  return -std::log(u) / lambda;
}
// Gamma(a, b) given x > cut.  Below the mode: plain rejection, as the reference (trun_gamma.cpp:76-81).  Beyond it the
// reference runs an adaptive rejection sampler (a > 1) or slice steps; here an exact exponential-envelope rejection:
// x = cut + Exp(lambda), lambda = b - (a-1)/cut (a > 1) or b (a <= 1), accepted with probability
// (x/cut)^(a-1) exp(-(a-1)(x-cut)/cut) (resp. (x/cut)^(a-1)), both <= 1.
double rtrun_gamma_mt(RNG &rng, double a, double b, double cut) {
  if (!(cut > 0)) return rgamma_mt(rng, a, b);
  const double mode = (a - 1.0) / b;
  if (cut < mode) {
    double x;
    do { x = rgamma_mt(rng, a, b); } while (x < cut);
    return x;
  }
  const double slope = a > 1.0 ? (a - 1.0) / cut : 0.0;
  const double lambda = b - slope;
  for (int tries = 0; tries < 100000; ++tries) {
    const double x = cut + rexp_mt(rng, lambda);
    const double log_accept = (a - 1.0) * std::log(x / cut) - slope * (x - cut);
    if (std::log(std::max(rng(), 1e-300)) <= log_accept) return x;
  }
  report_error("rtrun_gamma_mt: rejection sampler failed");
}

GenericGaussianVarianceSampler::GenericGaussianVarianceSampler(const std::shared_ptr<GammaModelBase> &prior, double sigma_max)
    : prior_(prior), sigma_max_(std::numeric_limits<double>::infinity()) { set_sigma_max(sigma_max); }
void GenericGaussianVarianceSampler::set_sigma_max(double sigma_max) {
  if (sigma_max < 0) report_error("sigma_max must be non-negative.");
  sigma_max_ = sigma_max;
}
double GenericGaussianVarianceSampler::draw(RNG &rng, double data_df, double data_ss, double scale) const {
  if (!prior_) report_error("GenericGaussianVarianceSampler is disabled because it was built with a null prior.");
  const double DF = data_df + 2 * prior_->alpha();
  const double SS = data_ss + 2 * prior_->beta() * scale * scale;
  if (sigma_max_ == 0.0) return 0.0;
  if (std::isinf(sigma_max_)) return 1.0 / rgamma_mt(rng, DF / 2, SS / 2);
  return 1.0 / rtrun_gamma_mt(rng, DF / 2, SS / 2, 1.0 / (sigma_max_ * sigma_max_));
}
double GenericGaussianVarianceSampler::posterior_mode(double data_df, double data_ss) const {
  if (!prior_) report_error("GenericGaussianVarianceSampler is disabled because it was built with a null prior.");
  const double alpha = (data_df + 2 * prior_->alpha()) / 2, beta = (data_ss + 2 * prior_->beta()) / 2;
  const double mode = beta / (alpha + 1), cap = sigma_max_ * sigma_max_;
  return mode > cap ? cap : mode;
}
double GenericGaussianVarianceSampler::log_prior(double sigsq) const {
  if (!prior_) report_error("GenericGaussianVarianceSampler is disabled because it was built with a null prior.");
  return prior_->logp(1.0 / sigsq) - 2 * std::log(sigsq);
}

// ---- ScalarSliceSampler (Samplers/ScalarSliceSampler.cpp) -------------------------------------
ScalarSliceSampler::ScalarSliceSampler(const Fun &logf, bool unimodal, double dx, RNG *rng)
    : logf_(logf), rng_(rng), suggested_dx_(dx), unimodal_(unimodal) {}
void ScalarSliceSampler::set_lower_limit(double lo) {
  if (std::isfinite(lo)) { lo_ = lower_bound_ = lo; lo_set_ = true; } else lo_set_ = false;
}
void ScalarSliceSampler::set_upper_limit(double hi) {
  if (std::isfinite(hi)) { hi_ = upper_bound_ = hi; hi_set_ = true; } else hi_set_ = false;
}
void ScalarSliceSampler::handle_error(const std::string &msg, double x) const {
  std::ostringstream err;
  err << msg << " in ScalarSliceSampler\nlo = " << lo_ << "  logp(lo) = " << logplo_ << "\nhi = " << hi_ << "  logp(hi) = " << logphi_
      << "\nx  = " << x << "  logp(x)  = " << logp_slice_ << "\n";
  report_error(err.str());
}
void ScalarSliceSampler::double_hi(double x) {
  hi_ = x + 2 * (hi_ - x);
  if (!std::isfinite(hi_)) handle_error("infinite upper limit", x);
  logphi_ = f(hi_);
}
void ScalarSliceSampler::double_lo(double x) {
  lo_ = x - 2 * (x - lo_);
  if (!std::isfinite(lo_)) handle_error("infinite lower limit", x);
  logplo_ = f(lo_);
}
bool ScalarSliceSampler::find_upper_limit(double x) {   // .cpp:185-201
  hi_ = x + suggested_dx_;
  logphi_ = f(hi_);
  int doubling_count = 0;
  while (logphi_ >= logp_slice_ || (!unimodal_ && runif_mt(*rng_) > .5)) {
    double_hi(x);
    if (++doubling_count > 100) return false;
  }
  if (x > hi_ || std::isnan(logphi_)) handle_error("problem with the upper limit", x);
  return true;
}
bool ScalarSliceSampler::find_lower_limit(double x) {   // .cpp:203-219
  lo_ = x - suggested_dx_;
  logplo_ = f(lo_);
  int doubling_count = 0;
  while (logplo_ >= logp_slice_ || (!unimodal_ && runif_mt(*rng_) > .5)) {
    double_lo(x);
    if (++doubling_count > 100) return false;
  }
  if (x < lo_ || std::isnan(logplo_)) handle_error("problem with the lower limit", x);
  return true;
}
bool ScalarSliceSampler::find_limits_unbounded(double x) {   // .cpp:138-170
  hi_ = x + suggested_dx_;
  lo_ = x - suggested_dx_;
  logphi_ = f(hi_);
  logplo_ = f(lo_);
  if (unimodal_) {
    while (logphi_ >= logp_slice_) double_hi(x);
    while (logplo_ >= logp_slice_) double_lo(x);
    return true;
  }
  int doubling_count = 0;
  while (!((logphi_ < logp_slice_) && (logplo_ < logp_slice_))) {
    if (runif_mt(*rng_, -1, 1) > 0) double_hi(x); else double_lo(x);
    if (++doubling_count > 100) return false;
  }
  return true;
}
void ScalarSliceSampler::find_limits(double x) {   // .cpp:108-133
  logp_slice_ = f(x) - rexp_mt(*rng_, 1.0);
  if (!std::isfinite(logp_slice_)) handle_error("initial value leads to infinite probability", x);
  bool found = true;
  if (lo_set_ && hi_set_) {
    lo_ = lower_bound_; logplo_ = f(lo_);
    hi_ = upper_bound_; logphi_ = f(hi_);
  } else if (lo_set_) {
    lo_ = lower_bound_; logplo_ = f(lo_);
    found = find_upper_limit(x);
  } else if (hi_set_) {
    found = find_lower_limit(x);
    hi_ = upper_bound_; logphi_ = f(hi_);
  } else {
    found = find_limits_unbounded(x);
  }
  if (x < lo_ || x > hi_) handle_error("problem building slice:  x out of bounds", x);
  if (found) {
    const bool logood = lo_set_ || (logplo_ <= logp_slice_), higood = hi_set_ || (logphi_ <= logp_slice_);
    if (!(logood && higood)) handle_error("problem with probabilities", x);
  }
}
void ScalarSliceSampler::contract(double x, double x_cand, double logp) {   // .cpp:92-104
  if (x_cand > x) { hi_ = x_cand; logphi_ = logp; } else { lo_ = x_cand; logplo_ = logp; }
  if (estimate_dx_) {
    suggested_dx_ = hi_ - lo_;
    if (suggested_dx_ < min_dx_) suggested_dx_ = min_dx_;
  }
}
double ScalarSliceSampler::draw(double x) {   // .cpp:67-88
  if (!rng_) report_error("ScalarSliceSampler: no random number generator");
  find_limits(x);
  for (int tries = 0; tries <= 100; ++tries) {
    const double x_cand = runif_mt(*rng_, lo_, hi_);
    const double logp_cand = f(x_cand);
    if (logp_cand >= logp_slice_) return x_cand;
    contract(x, x_cand, logp_cand);
  }
  handle_error("number of tries exceeded", x);
}

// ---- TRegressionModel -------------------------------------------------------------------------
TRegressionModel::TRegressionModel(int64_t n, int p, const double *X, const double *y) : GlmModelBase(p) {
  x_.assign(X, X + (size_t)n * p);
  y_.assign(y, y + n);
}
void TRegressionModel::add_data(double y, const Vector &x) {
  if ((int)x.size() != xdim()) report_error("TRegressionModel::add_data: wrong size x");
  if (adopted_ || borrowed_) report_error("add_data on a model whose data live in adopted / borrowed memory");
  x_.insert(x_.end(), x.begin(), x.end());
  y_.push_back(y);
  touch();
}
void TRegressionModel::adopt_device_data(int64_t n, const double *dX, int64_t ldx, const double *dy) {
  adopted_ = true; adopted_n_ = n; dX_ = dX; dldx_ = ldx; dy_ = dy;
  touch();
}
void TRegressionModel::borrow_host_data(int64_t n, const double *X, int64_t ldx, const double *y, std::shared_ptr<void> keepalive) {
  borrowed_ = true; adopted_ = false; adopted_n_ = n; dX_ = X; dldx_ = ldx; dy_ = y;
  keepalive_ = std::move(keepalive);
  touch();
}
void TRegressionModel::upload(DeviceData &dev) {
  if (borrowed_) { dev.check(boomgpu_upload_regression(dev.ctx(), adopted_n_, xdim(), dX_, dldx_, dy_)); return; }
  if (adopted_) dev.check(boomgpu_adopt_regression(dev.ctx(), adopted_n_, xdim(), dX_, dldx_, dy_));
  else dev.check(boomgpu_upload_regression(dev.ctx(), (int64_t)y_.size(), xdim(), x_.data(), xdim(), y_.data()));
}
void TRegressionModel::set_sigsq(double s2) {
  if (!(s2 > 0)) report_error("TRegressionModel::set_sigsq: sigsq must be positive");
  sigsq_ = s2;
}
void TRegressionModel::set_nu(double nu) {
  if (!(nu > 0)) report_error("TRegressionModel::set_nu: nu must be positive");
  nu_ = nu;
}
double TRegressionModel::log_likelihood(const Vector &beta, double sigsq, double nu) {
  if ((int)beta.size() != xdim()) report_error("log_likelihood: wrong size beta");
  if (!(nu > 0) || !(sigsq > 0)) return -std::numeric_limits<double>::infinity();
  if (allreduce()) report_error("TRegressionModel::log_likelihood: shard through set_communicator (the all-reduce hook is not supported here)");
  DeviceData &dev(device_data());
  double ans = 0;
  dev.check(boomgpu_student_loglike(dev.ctx(), beta.data(), std::sqrt(sigsq), nu, &ans));
  return ans;
}
double TRegressionModel::log_likelihood_same_beta(double sigsq, double nu) {
  if (!(nu > 0) || !(sigsq > 0)) return -std::numeric_limits<double>::infinity();
  DeviceData &dev(device_data());
  double ans = 0;
  dev.check(boomgpu_student_loglike(dev.ctx(), nullptr, std::sqrt(sigsq), nu, &ans));
  return ans;
}
double TRegressionModel::log_likelihood_derivs(const Vector &, Vector *, SpdMatrix *) {
  report_error("derivatives of the TRegressionModel log likelihood are not provided (TRegression.cpp:118-121)");
}

// ---- TRegressionSampler -----------------------------------------------------------------------
TRegressionSampler::TRegressionSampler(TRegressionModel *model, const std::shared_ptr<MvnBase> &coefficient_prior,
                                       const std::shared_ptr<GammaModelBase> &siginv_prior,
                                       const std::shared_ptr<DoubleModel> &nu_prior, RNG &seeding_rng)
    : PosteriorSampler(seeding_rng), model_(model), coefficient_prior_(coefficient_prior), siginv_prior_(siginv_prior),
      nu_prior_(nu_prior), suf_(model ? model->xdim() : 0), sigsq_sampler_(siginv_prior),
      nu_observed_([this](double nu) {
                     // TRegressionLogPosterior (.cpp:31-49); the residuals of the current beta are computed once per draw
                     double ans = nu_prior_->logp(nu);
                     if (!(ans > -std::numeric_limits<double>::infinity())) return ans;
                     if (!residuals_current_) {
                       residuals_current_ = true;
                       return ans + model_->log_likelihood(model_->Beta(), model_->sigsq(), nu);
                     }
                     return ans + model_->log_likelihood_same_beta(model_->sigsq(), nu);
                   }, false, 1.0, &rng()),
      nu_complete_([this](double nu) {
                     // TRegressionCompleteDataLogPosterior (.cpp:51-72) over ScaledChisqModel::Loglike (ScaledChisqModel.cpp:52-82)
                     if (nu <= 0.0) return -std::numeric_limits<double>::infinity();
                     double ans = nu_prior_->logp(nu);
                     if (!(ans > -std::numeric_limits<double>::infinity())) return ans;
                     const double nu2 = nu / 2.0;
                     return ans + suf_.n() * (nu2 * std::log(nu2) - std::lgamma(nu2)) + (nu2 - 1) * suf_.sumlogw() - nu2 * suf_.sumw();
                   }, false, 1.0, &rng()) {
  if (!model) report_error("TRegressionSampler: null model");
  if (!coefficient_prior || coefficient_prior->dim() != model->xdim()) report_error("Prior does not match model dimension.");
  if (!siginv_prior || !nu_prior) report_error("TRegressionSampler: null prior");
  nu_observed_.set_lower_limit(0.0);
  nu_complete_.set_lower_limit(0.0);
  device_seed_ = seed_rng(rng());
}
void TRegressionSampler::on_seed() { device_seed_ = seed_rng(rng()); iteration_ = 0; }

void TRegressionSampler::draw() {
  impute_latent_data();
  draw_beta_full_conditional();
  draw_sigsq_full_conditional();
  draw_nu_given_observed_data();
}
double TRegressionSampler::logpri() const {
  return nu_prior_->logp(model_->nu()) + sigsq_sampler_.log_prior(model_->sigsq()) + coefficient_prior_->logp(model_->Beta());
}
void TRegressionSampler::impute_latent_data() {
  if (latent_data_fixed_) return;
  const uint64_t seed = device_seed_, it = iteration_++;
  const Vector &beta(model_->Beta());
  const double sigma = model_->sigma(), nu = model_->nu();
  run_device_step(
      *model_, suf_, packed_,
      [&](boomgpu_ctx *ctx, double *suf_dev) { return boomgpu_student_step_device(ctx, beta.data(), sigma, nu, seed, it, suf_dev); },
      [&](boomgpu_ctx *ctx, double *xtx, double *xty, double *scalars) {
        return boomgpu_student_step(ctx, beta.data(), sigma, nu, seed, it, xtx, xty, scalars);
      });
}
void TRegressionSampler::draw_beta_full_conditional() {   // draw_beta_full_conditional_impl, .cpp:74-85
  const int p = model_->xdim();
  const double sigsq = model_->sigsq();
  SpdMatrix precision(coefficient_prior_->siginv());
  Vector scaled_mean(p, 0.0);
  for (int i = 0; i < p; ++i) {
    double s = 0;
    for (int j = 0; j < p; ++j) {
      precision.a[(size_t)i * p + j] += suf_.xtx()(i, j) / sigsq;
      s += coefficient_prior_->siginv()(i, j) * coefficient_prior_->mu()[j];
    }
    scaled_mean[i] = s + suf_.xty()[i] / sigsq;
  }
  model_->set_Beta(rmvn_suf_mt(rng(), precision, scaled_mean));
  residuals_current_ = false;
}

// ---- TRegressionSpikeSlabSampler ---------------------------------------------------------------
TRegressionSpikeSlabSampler::TRegressionSpikeSlabSampler(TRegressionModel *model, const std::shared_ptr<MvnBase> &slab,
                                                         const std::shared_ptr<VariableSelectionPrior> &spike,
                                                         const std::shared_ptr<GammaModelBase> &siginv_prior,
                                                         const std::shared_ptr<DoubleModel> &nu_prior, RNG &seeding_rng)
    : TRegressionSampler(model, slab, siginv_prior, nu_prior, seeding_rng), core_(slab, spike, true), scaled_(model ? model->xdim() : 0) {
  if (!spike || spike->potential_nvars() != model->xdim()) report_error("Prior does not match model dimension.");
}
const WeightedRegSuf &TRegressionSpikeSlabSampler::scaled_statistics() {
  const int p = model_->xdim();
  const double inv = 1.0 / model_->sigsq();
  double *a = scaled_.xtx_storage(p), *b = scaled_.xty_storage();
  const Vector &src(suf_.xtx().a);
  for (size_t e = 0; e < (size_t)p * p; ++e) a[e] = src[e] * inv;
  for (int j = 0; j < p; ++j) b[j] = suf_.xty()[j] * inv;
  scaled_.set_scalars(suf_.n(), suf_.yty() * inv, suf_.sumw(), suf_.sumlogw());
  return scaled_;
}
bool TRegressionSpikeSlabSampler::impute_latent_data_active(const std::vector<int> &cols_in) {
  const int p = model_->xdim();
  if (latent_data_is_fixed() || p <= 64 || model_->allreduce() || cols_in.size() > 128) return false;
  std::vector<int> cols(cols_in);
  if (cols.empty()) cols.push_back(0);
  DeviceData &dev(model_->device_data());
  const int k = (int)cols.size();
  active_.cols = cols;
  active_.G.resize((size_t)p * k); active_.diag.resize(p); active_.xty.resize(p);
  std::vector<int32_t> c32(cols.begin(), cols.end());
  uint64_t seed, it;
  next_device_key(&seed, &it);
  dev.check(boomgpu_student_step_active(dev.ctx(), model_->Beta().data(), model_->sigma(), model_->nu(), seed, it, c32.data(), k,
                                        active_.G.data(), active_.diag.data(), active_.xty.data(), active_.scalars));
  // the scalar statistics (n, y'Wy, sum w, sum log w) are complete in this form too: sigsq and nu | weights read them from suf_
  suf_.set_scalars(active_.scalars[0], active_.scalars[1], active_.scalars[2], active_.scalars[3]);
  active_.valid = true;
  view_.reset();
  return true;
}
// The spike-and-slab steps read the statistics divided by sigsq: the view holds scaled copies of the active arrays and scales
// every column it fetches.  Rebuilt when sigsq has changed since it was made (it has not between the sweep and the beta draw).
StatView &TRegressionSpikeSlabSampler::scaled_view() {
  const double sigsq = model_->sigsq();
  if (view_ && view_sigsq_ == sigsq) return *view_;
  const int p = model_->xdim();
  const double inv = 1.0 / sigsq;
  Vector G(active_.G), diag(active_.diag);
  for (double &v : G) v *= inv;
  for (double &v : diag) v *= inv;
  view_xty_ = active_.xty;
  for (double &v : view_xty_) v *= inv;
  DeviceData *dev = &model_->device_data();
  ActiveSetState *st = &active_;
  view_.reset(new StatView(p, active_.cols, G, diag, view_xty_, [dev, st, inv, p](int j, double *out) {
    dev->check(boomgpu_weighted_column(dev->ctx(), j, out));
    for (int i = 0; i < p; ++i) out[i] *= inv;
    ++st->columns_fetched;
  }));
  view_sigsq_ = sigsq;
  return *view_;
}
void TRegressionSpikeSlabSampler::materialize_full_statistics() const {
  if (!active_.valid) return;
  DeviceData &dev(model_->device_data());
  const int p = model_->xdim();
  WeightedRegSuf &suf(const_cast<WeightedRegSuf &>(suf_));
  dev.check(boomgpu_full_statistics(dev.ctx(), suf.xtx_storage(p), suf.xty_storage()));
  suf.set_scalars(active_.scalars[0], active_.scalars[1], active_.scalars[2], active_.scalars[3]);
  active_.valid = false;
}
double TRegressionSpikeSlabSampler::weighted_sum_of_squared_errors() {
  if (!active_.valid) return TRegressionSampler::weighted_sum_of_squared_errors();
  // every included variable's column is in the view (the model's columns at the start of the iteration + the fetched adds)
  StatView &v(scaled_view());
  const double sigsq = view_sigsq_;
  const Vector &b(model_->Beta());
  const std::vector<int> inc(model_->coef().inc().included_positions());
  double bxy = 0, bxxb = 0;
  for (int i : inc) {
    bxy += b[i] * active_.xty[i];
    double s = 0;
    for (int j : inc) s += v.at(i, j) * b[j];
    bxxb += b[i] * s;
  }
  return active_.scalars[1] - 2 * bxy + bxxb * sigsq;
}
void TRegressionSpikeSlabSampler::draw() {
  if (!(active_.enabled && impute_latent_data_active(model_->coef().inc().included_positions()))) {
    active_.valid = false;
    view_.reset();
    impute_latent_data();
  }
  draw_model_indicators();
  draw_included_coefficients();
  draw_sigsq_full_conditional();
  draw_nu_given_observed_data();
}
double TRegressionSpikeSlabSampler::logpri() const {
  return core_.logpri(model_->coef()) + nu_prior_->logp(model_->nu()) + siginv_prior_->logp(1.0 / model_->sigsq());
}
void TRegressionSpikeSlabSampler::draw_model_indicators() {
  if (active_.valid) core_.draw_model_indicators(rng(), model_->coef(), scaled_view());
  else core_.draw_model_indicators(rng(), model_->coef(), scaled_statistics());
  coefficients_changed();
}
void TRegressionSpikeSlabSampler::draw_included_coefficients() {
  if (active_.valid) core_.draw_beta(rng(), model_->coef(), scaled_view());
  else core_.draw_beta(rng(), model_->coef(), scaled_statistics());
  coefficients_changed();
}
double TRegressionSpikeSlabSampler::log_model_prob(const Selector &g) {
  materialize_full_statistics();
  return core_.log_model_prob(g, scaled_statistics());
}

void TRegressionSampler::draw_sigsq_full_conditional() {   // .cpp:165-171
  model_->set_sigsq(sigsq_sampler_.draw(rng(), suf_.n(), weighted_sum_of_squared_errors()));
}
double TRegressionSampler::weighted_sum_of_squared_errors() {   // SSE = y'Wy - 2 b'X'Wy + b'X'WXb (WeightedRegressionModel.cpp:89-95)
  const int p = model_->xdim();
  const Vector &b(model_->Beta());
  double bxy = 0, bxxb = 0;
  for (int i = 0; i < p; ++i) {
    if (b[i] == 0.0) continue;
    bxy += b[i] * suf_.xty()[i];
    double s = 0;
    for (int j = 0; j < p; ++j) s += suf_.xtx()(i, j) * b[j];
    bxxb += b[i] * s;
  }
  return suf_.yty() - 2 * bxy + bxxb;
}
void TRegressionSampler::draw_nu_given_complete_data() { model_->set_nu(nu_complete_.draw(model_->nu())); }
void TRegressionSampler::draw_nu_given_observed_data() {
  residuals_current_ = false;   // beta may have been set from outside since the last draw
  model_->set_nu(nu_observed_.draw(model_->nu()));
}

}  // namespace BOOM_B200
