// mixture_table.cpp -- the Poisson auxiliary-mixture table on the host, including the entries the data need but the
// shipped grid does not hold.
//
// Reference: NormalMixtureApproximationTable (Models/Glm/PosteriorSamplers/NormalMixtureApproximation.hpp:262-330,
// .cpp:426-560).  Its approximate(nu) (.cpp:472-532) is called per observation inside the draw and MUTATES the table:
//   * nu on the grid                        -> that entry;
//   * neighbours nu0 < nu < nu1 of equal size -> the linear interpolation of (mu, sigma, weights) with weight
//     (nu - nu0) / (nu1 - nu0), kept when its Kullback-Leibler divergence from the -log Gamma(nu, 1) density is below 1e-5;
//   * otherwise                              -> a direct fit with the lower neighbour's number of components that minimises the
//     same divergence (the reference runs Powell's method from a symmetric start).
// The device cannot grow a table in the middle of a kernel, so here the host asks the device once per data set which counts
// occur (boomgpu_poisson_counts_present), adds the missing entries by the rule above and re-states the table
// (boomgpu_set_poisson_table skips the upload when nothing changed).
// Interpolated entries are the reference's numbers to rounding.  A directly fitted entry is the minimiser found by a
// Nelder-Mead search started from the lower neighbour (rescaled to nu): an equally valid mixture -- the sampler only needs
// the divergence to be negligible -- but not the same iterate as the reference's Powell run; tests/test_host_logic.py checks
// its divergence against the reference's own fits (tests/golden/poisson_offgrid.json).
#include <algorithm>
#include <cmath>
#include <functional>
#include <limits>
#include <numeric>

#include "boom_b200.hpp"

namespace BOOM_B200 {

namespace {

constexpr double kLnSqrt2Pi = 0.918938533204672741780329736406;

// log of the mixture density (NormalMixtureApproximation::logp, .cpp:270-277)
double mixture_logp(const NormalMixtureApproximation &a, double y) {
  double mx = -std::numeric_limits<double>::infinity();
  double lp[32];
  const int K = a.dim();
  for (int s = 0; s < K; ++s) {
    const double z = (y - a.mu[s]) / a.sigma[s];
    lp[s] = std::log(a.weights[s]) - kLnSqrt2Pi - std::log(a.sigma[s]) - 0.5 * z * z;
    mx = std::max(mx, lp[s]);
  }
  if (!std::isfinite(mx)) return mx;
  double tot = 0;
  for (int s = 0; s < K; ++s) tot += std::exp(lp[s] - mx);
  return mx + std::log(tot);
}

// NegLogGamma (NormalMixtureApproximation.hpp:210-219): log density of -log Gamma(nu, 1)
double neg_log_gamma(double nu, double y) { return -nu * y - std::exp(-y) - std::lgamma(nu); }

// adaptive Gauss-Kronrod (7, 15) on [a, b]
double gk15(const std::function<double(double)> &f, double a, double b, double tol, int depth) {
  static const double xgk[8] = {0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
                                0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
                                0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                                0.207784955007898467600689403773245, 0.000000000000000000000000000000000};
  static const double wgk[8] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                                0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                                0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                                0.204432940075298892414161999234649, 0.209482141084727828012999174891714};
  static const double wg[4] = {0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
                               0.381830050505118944950369775488975, 0.417959183673469387755102040816327};
  const double c = 0.5 * (a + b), h = 0.5 * (b - a);
  const double fc = f(c);
  double rk = fc * wgk[7], rg = fc * wg[3];
  for (int j = 0; j < 7; ++j) {
    const double dx = h * xgk[j];
    const double s = f(c - dx) + f(c + dx);
    rk += wgk[j] * s;
    if (j % 2 == 1) rg += wg[j / 2] * s;
  }
  rk *= h; rg *= h;
  // (the tolerance is NOT halved per level: rounding noise in f (log f - log approx) puts a floor of ~1e-16 under the
  // error estimate, and a halved tolerance would chase it through all 2^depth leaves)
  if (depth <= 0 || std::fabs(rk - rg) <= tol) return rk;
  return gk15(f, a, c, tol, depth - 1) + gk15(f, c, b, tol, depth - 1);
}

}  // namespace

// NormalMixtureApproximation::kullback_leibler(target) for target = NegLogGamma(nu) (.cpp:296-319): the integration limits
// are where the target has dropped 30 log units below its mode (at -log nu), found in unit steps as there.
double kullback_leibler_neg_log_gamma(double nu, const NormalMixtureApproximation &approx) {
  const double mode = -std::log(nu);
  const double top = neg_log_gamma(nu, mode);
  double lo = mode - 1, hi = mode + 1;
  while (top - neg_log_gamma(nu, lo) < 30) lo -= 1;
  while (top - neg_log_gamma(nu, hi) < 30) hi += 1;
  auto integrand = [&](double x) {
    const double lf = neg_log_gamma(nu, x);
    return std::exp(lf) * (lf - mixture_logp(approx, x));
  };
  return gk15(integrand, lo, mode, 1e-12, 10) + gk15(integrand, mode, hi, 1e-12, 10);
}

void NormalMixtureApproximation::order_by_mu() {
  std::vector<int> idx(mu.size());
  std::iota(idx.begin(), idx.end(), 0);
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return mu[a] < mu[b]; });
  Vector m(mu.size()), s(mu.size()), w(mu.size());
  for (size_t i = 0; i < idx.size(); ++i) { m[i] = mu[idx[i]]; s[i] = sigma[idx[i]]; w[i] = weights[idx[i]]; }
  mu = m; sigma = s; weights = w;
}

namespace {

// theta = (mu[K], log sigma[K], log(w[1..K-1] / w[0])) as in the reference's parameterisation (.cpp:38-52,160-175)
NormalMixtureApproximation from_theta(const Vector &theta, int K) {
  NormalMixtureApproximation a;
  a.mu.assign(theta.begin(), theta.begin() + K);
  a.sigma.resize(K);
  for (int k = 0; k < K; ++k) a.sigma[k] = std::exp(theta[K + k]);
  a.weights.assign(K, 1.0);
  double tot = 1.0;
  for (int k = 1; k < K; ++k) { a.weights[k] = std::exp(theta[2 * K + k - 1]); tot += a.weights[k]; }
  for (int k = 0; k < K; ++k) a.weights[k] /= tot;
  return a;
}

// Nelder-Mead with restarts; returns the minimum found
double nelder_mead(const std::function<double(const Vector &)> &f, Vector &x, double step, double precision, int max_evals, int *evals) {
  const int d = (int)x.size();
  double best = f(x);
  *evals = 1;
  for (int restart = 0; restart < 12 && *evals < max_evals; ++restart) {
    std::vector<Vector> sx(d + 1, x);
    Vector fx(d + 1, best);
    for (int i = 0; i < d; ++i) { sx[i + 1][i] += step; fx[i + 1] = f(sx[i + 1]); ++*evals; }
    for (;;) {
      std::vector<int> ord(d + 1);
      std::iota(ord.begin(), ord.end(), 0);
      std::sort(ord.begin(), ord.end(), [&](int a, int b) { return fx[a] < fx[b]; });
      const int lo = ord[0], hi = ord[d], nhi = ord[d - 1];
      if (std::fabs(fx[hi] - fx[lo]) <= 0.01 * precision * (std::fabs(fx[lo]) + 1e-12) || *evals >= max_evals) break;
      Vector cen(d, 0.0);
      for (int i = 0; i <= d; ++i) if (i != hi) for (int j = 0; j < d; ++j) cen[j] += sx[i][j] / d;
      auto along = [&](double t) { Vector y(d); for (int j = 0; j < d; ++j) y[j] = cen[j] + t * (sx[hi][j] - cen[j]); return y; };
      Vector xr = along(-1.0);
      const double fr = f(xr); ++*evals;
      if (fr < fx[lo]) {
        Vector xe = along(-2.0);
        const double fe = f(xe); ++*evals;
        if (fe < fr) { sx[hi] = xe; fx[hi] = fe; } else { sx[hi] = xr; fx[hi] = fr; }
      } else if (fr < fx[nhi]) {
        sx[hi] = xr; fx[hi] = fr;
      } else {
        Vector xc = along(fr < fx[hi] ? -0.5 : 0.5);
        const double fc = f(xc); ++*evals;
        if (fc < std::min(fr, fx[hi])) { sx[hi] = xc; fx[hi] = fc; }
        else {
          for (int i = 0; i <= d; ++i) if (i != lo) {
            for (int j = 0; j < d; ++j) sx[i][j] = sx[lo][j] + 0.5 * (sx[i][j] - sx[lo][j]);
            fx[i] = f(sx[i]); ++*evals;
          }
        }
      }
    }
    const int lo = (int)(std::min_element(fx.begin(), fx.end()) - fx.begin());
    const double gain = best - fx[lo];
    if (fx[lo] < best) { best = fx[lo]; x = sx[lo]; }
    if (restart > 0 && gain <= precision * (std::fabs(best) + 1e-12)) break;
    step *= 0.5;
  }
  return best;
}

}  // namespace

NormalMixtureApproximation fit_neg_log_gamma(double nu, const NormalMixtureApproximation &start, double precision, int max_evals,
                                             double stepsize) {
  const int K = start.dim();
  Vector theta(3 * K - 1);
  for (int k = 0; k < K; ++k) { theta[k] = start.mu[k]; theta[K + k] = std::log(start.sigma[k]); }
  for (int k = 1; k < K; ++k) theta[2 * K + k - 1] = std::log(start.weights[k] / start.weights[0]);
  auto objective = [&](const Vector &t) {
    for (int k = 0; k < K; ++k) if (!(std::fabs(t[K + k]) < 50)) return 1e300;
    const double v = kullback_leibler_neg_log_gamma(nu, from_theta(t, K));
    return std::isfinite(v) ? v : 1e300;
  };
  int evals = 0;
  const double kl = nelder_mead(objective, theta, stepsize, precision, max_evals, &evals);
  NormalMixtureApproximation a = from_theta(theta, K);
  a.order_by_mu();
  a.kullback_leibler = kl;
  a.number_of_function_evaluations = evals;
  return a;
}

// ---------------------------------------------------------------------------------------------
void NormalMixtureApproximationTable::deserialize(const Vector &ser) {
  index_.clear(); approximations_.clear(); derived_.clear();
  size_t i = 0;
  while (i < ser.size()) {
    if (i + 1 >= ser.size()) report_error("NormalMixtureApproximationTable::deserialize: truncated table");
    const int64_t nu = std::llround(ser[i]);
    const int K = (int)std::llround(ser[i + 1]);
    if (K < 1 || i + 2 + 3 * (size_t)K > ser.size()) report_error("NormalMixtureApproximationTable::deserialize: malformed table");
    NormalMixtureApproximation a;
    a.weights.assign(ser.begin() + i + 2, ser.begin() + i + 2 + K);
    a.sigma.assign(ser.begin() + i + 2 + K, ser.begin() + i + 2 + 2 * K);
    a.mu.assign(ser.begin() + i + 2 + 2 * K, ser.begin() + i + 2 + 3 * K);
    index_.push_back(nu);
    approximations_.push_back(a);
    derived_.push_back(0);
    i += 2 + 3 * (size_t)K;
  }
}

Vector NormalMixtureApproximationTable::serialize() const {
  Vector ans;
  for (size_t e = 0; e < index_.size(); ++e) {
    const NormalMixtureApproximation &a(approximations_[e]);
    ans.push_back((double)index_[e]);
    ans.push_back((double)a.dim());
    ans.insert(ans.end(), a.weights.begin(), a.weights.end());
    ans.insert(ans.end(), a.sigma.begin(), a.sigma.end());
    ans.insert(ans.end(), a.mu.begin(), a.mu.end());
  }
  return ans;
}

void NormalMixtureApproximationTable::add(int64_t nu, const NormalMixtureApproximation &a) { insert(nu, a, false); }

void NormalMixtureApproximationTable::insert(int64_t nu, const NormalMixtureApproximation &a, bool derived) {
  auto it = std::lower_bound(index_.begin(), index_.end(), nu);
  const size_t pos = (size_t)(it - index_.begin());
  index_.insert(it, nu);
  approximations_.insert(approximations_.begin() + pos, a);
  derived_.insert(derived_.begin() + pos, derived ? 1 : 0);
}

bool NormalMixtureApproximationTable::contains(int64_t nu) const {
  auto it = std::lower_bound(index_.begin(), index_.end(), nu);
  return it != index_.end() && *it == nu;
}

const NormalMixtureApproximation &NormalMixtureApproximationTable::approximate(int64_t nu) {
  if (index_.empty()) report_error("NormalMixtureApproximationTable::approximate: empty table");
  auto it = std::lower_bound(index_.begin(), index_.end(), nu);
  const size_t pos = (size_t)(it - index_.begin());
  if (it != index_.end() && *it == nu) return approximations_[pos];
  // Neighbours: the nearest entries that were NOT themselves derived by this function.  (The reference takes whatever sits
  // next to nu at the time of the call, so its result depends on the order in which counts were first seen; with rows
  // sharded over GPUs that order differs per rank, and every rank must derive the same entry for the same count.)
  size_t lo = pos, hi = pos;
  while (lo > 0 && derived_[lo - 1]) --lo;
  while (hi < index_.size() && derived_[hi]) ++hi;
  if (lo == 0 || hi >= index_.size())
    report_error("NormalMixtureApproximationTable::approximate: nu outside the table's range");   // the caller handles nu >= largest_index
  const int64_t nu0 = index_[lo - 1], nu1 = index_[hi];
  const NormalMixtureApproximation a0 = approximations_[lo - 1], a1 = approximations_[hi];
  const double weight = (double)(nu - nu0) / (1.0 * (double)(nu1 - nu0));
  const double precision = 1e-6;
  const int max_evals = 20000;
  const double stepsize = .5 / std::sqrt((double)nu);
  if (a0.dim() == a1.dim()) {
    NormalMixtureApproximation a;
    const int K = a0.dim();
    a.mu.resize(K); a.sigma.resize(K); a.weights.resize(K);
    for (int k = 0; k < K; ++k) {
      a.mu[k] = (1 - weight) * a0.mu[k] + weight * a1.mu[k];
      a.sigma[k] = (1 - weight) * a0.sigma[k] + weight * a1.sigma[k];
      a.weights[k] = (1 - weight) * a0.weights[k] + weight * a1.weights[k];
    }
    a.order_by_mu();
    a.kullback_leibler = kullback_leibler_neg_log_gamma((double)nu, a);
    if (a.kullback_leibler < 1e-5) {
      insert(nu, a, true);
      return approximate(nu);
    }
  }
  // direct fit with the lower neighbour's size, started from the lower neighbour moved to nu's location and scale
  NormalMixtureApproximation start = a0;
  const double shift = -std::log((double)nu) + std::log((double)nu0), scale = std::sqrt((double)nu0 / (double)nu);
  for (int k = 0; k < start.dim(); ++k) { start.mu[k] += shift; start.sigma[k] *= scale; }
  insert(nu, fit_neg_log_gamma((double)nu, start, precision, max_evals, stepsize), true);
  return approximate(nu);
}

}  // namespace BOOM_B200
