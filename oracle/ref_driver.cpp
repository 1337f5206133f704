// ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Links the UNMODIFIED reference (steve-the-bayesian/BOOM, compiled from
// /root/reference by oracle/build_ref.sh into build/boomref/libboom_ref.a) and
//   golden <dir>   writes the golden vectors the oracle and the CUDA path are pinned to
//   bench ...      times the reference's own samplers (bench.py --impl reference, cpu_baseline)
// The binary lands in oracle/_ref/ (git-ignored, travels to the GPU box).
// Synthetic data come from the oracle's bo_synth_* helpers so every arm sees
// the same numbers; nothing else of the oracle is used here.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "Models/Glm/BinomialLogitModel.hpp"
#include "Models/Glm/BinomialRegressionData.hpp"
#include "Models/Glm/PoissonRegressionData.hpp"
#include "Models/Glm/PoissonRegressionModel.hpp"
#include "Models/Glm/PosteriorSamplers/BinomialLogitAuxmixSampler.hpp"
#include "Models/Glm/PosteriorSamplers/BinomialLogitDataImputer.hpp"
#include "Models/Glm/PosteriorSamplers/BinomialLogitSpikeSlabSampler.hpp"
#include "Models/Glm/PosteriorSamplers/NormalMixtureApproximation.hpp"
#include "Models/Glm/PosteriorSamplers/PoissonDataImputer.hpp"
#include "Models/Glm/PosteriorSamplers/PoissonRegressionAuxMixSampler.hpp"
#include "Models/Glm/PosteriorSamplers/poisson_mixture_approximation_table.hpp"
#include "Models/ChisqModel.hpp"
#include "Models/Glm/PosteriorSamplers/TDataImputer.hpp"
#include "Models/Glm/PosteriorSamplers/TRegressionSampler.hpp"
#include "Models/Glm/PosteriorSamplers/TRegressionSpikeSlabSampler.hpp"
#include "Models/Glm/TRegression.hpp"
#include "Models/Glm/VariableSelectionPrior.hpp"
#include "Models/UniformModel.hpp"
#include "Models/Glm/WeightedRegressionModel.hpp"
#include "Models/MvnModel.hpp"
#include "distributions.hpp"
#include "distributions/trun_logit.hpp"

#include "auxmix_oracle.h"

using namespace BOOM;

namespace {

struct Json {
  std::ostringstream s;
  bool first = true;
  Json() { s << std::setprecision(17); s << "{"; }
  void key(const std::string &k) { if (!first) s << ",\n"; first = false; s << "\"" << k << "\": "; }
  void num(const std::string &k, double v) { key(k); put(v); }
  void put(double v) {
    if (std::isfinite(v)) s << v; else if (std::isnan(v)) s << "NaN"; else s << (v > 0 ? "Infinity" : "-Infinity");
  }
  template <class V> void arr(const std::string &k, const V &v) {
    key(k); s << "[";
    for (size_t i = 0; i < (size_t)v.size(); ++i) { if (i) s << ", "; put((double)v[i]); }
    s << "]";
  }
  void raw(const std::string &k, const std::string &v) { key(k); s << v; }
  std::string str() { return s.str() + "}\n"; }
};

std::string row_json(const std::vector<std::pair<std::string, double>> &kv) {
  std::ostringstream o; o << std::setprecision(17) << "{";
  for (size_t i = 0; i < kv.size(); ++i) {
    if (i) o << ", ";
    o << "\"" << kv[i].first << "\": ";
    double v = kv[i].second;
    if (std::isfinite(v)) o << v; else if (std::isnan(v)) o << "NaN"; else o << (v > 0 ? "Infinity" : "-Infinity");
  }
  o << "}";
  return o.str();
}

std::string list_json(const std::vector<std::string> &rows) {
  std::ostringstream o; o << "[\n";
  for (size_t i = 0; i < rows.size(); ++i) { o << "  " << rows[i] << (i + 1 < rows.size() ? ",\n" : "\n"); }
  o << "]";
  return o.str();
}

void write_file(const std::string &path, const std::string &body) {
  std::ofstream f(path);
  f << body;
  fprintf(stderr, "wrote %s (%zu bytes)\n", path.c_str(), body.size());
}

int sigma_index(const Vector &sigma, double sigsq) {
  int best = 0; double bd = 1e300;
  for (int k = 0; k < (int)sigma.size(); ++k) {
    double d = std::fabs(sigma[k] * sigma[k] - sigsq);
    if (d < bd) { bd = d; best = k; }
  }
  return best;
}

// ------------------------------------------------------------------------------------
void golden_mixture(const std::string &dir) {
  const LogitMixtureApproximation &m(BinomialLogitDataImputer::mixture_approximation);
  Json j;
  j.arr("mu", m.mu()); j.arr("sigma", m.sigma()); j.arr("weights", m.weights()); j.arr("log_weights", m.log_weights());
  write_file(dir + "/logit_mixture.json", j.str());

  // unmix(rng, u): seed, call, re-seed and read the uniform the call consumed.
  std::vector<std::string> rows;
  RNG rng(1);
  for (int c = 0; c < 600; ++c) {
    double u = -9.0 + 18.0 * (c % 200) / 199.0 + 0.013 * (c / 200);
    unsigned long seed = 1000 + c;
    rng.seed(seed);
    double mu, sigsq;
    m.unmix(rng, u, &mu, &sigsq);
    rng.seed(seed);
    double U = rng();
    rows.push_back(row_json({{"u", u}, {"U", U}, {"k", (double)sigma_index(m.sigma(), sigsq)}, {"sigsq", sigsq}}));
  }
  write_file(dir + "/unmix_logit.json", list_json(rows));

  rows.clear();
  for (int c = 0; c < 400; ++c) {
    double eta = -12.0 + 24.0 * (c % 100) / 99.0;
    bool above = (c / 100) % 2 == 0;
    unsigned long seed = 5000 + c;
    rng.seed(seed);
    double z = rtrun_logit_mt(rng, eta, 0, above);
    rng.seed(seed);
    double U = rng();
    rows.push_back(row_json({{"eta", eta}, {"above", above ? 1.0 : 0.0}, {"U", U}, {"z", z}}));
  }
  write_file(dir + "/rtrun_logit.json", list_json(rows));

  // BinomialLogitCltDataImputer::impute, small-sample branch, uniforms recovered the same way
  rows.clear();
  BinomialLogitCltDataImputer imputer(10);
  for (int c = 0; c < 300; ++c) {
    double eta = -6.0 + 12.0 * (c % 60) / 59.0;
    int nt = 1 + (c / 60) % 4;          // 1..4 trials
    int y = (c * 7) % (nt + 1);
    unsigned long seed = 9000 + c;
    rng.seed(seed);
    std::pair<double, double> ans = imputer.impute(rng, nt, y, eta);
    rng.seed(seed);
    std::vector<double> U(2 * nt);
    for (int i = 0; i < 2 * nt; ++i) U[i] = rng();
    std::ostringstream o; o << std::setprecision(17);
    o << "{\"eta\": " << eta << ", \"ntrials\": " << nt << ", \"y\": " << y << ", \"sum\": " << ans.first
      << ", \"info\": " << ans.second << ", \"U\": [";
    for (int i = 0; i < 2 * nt; ++i) o << (i ? ", " : "") << U[i];
    o << "]}";
    rows.push_back(o.str());
  }
  write_file(dir + "/logit_impute_small.json", list_json(rows));

  rows.clear();
  for (int c = 0; c < 240; ++c) {
    double mu = -9.0 + 18.0 * (c % 60) / 59.0;
    double sigma = m.sigma()[(c / 60) * 2 + 1];
    for (int pos = 0; pos < 2; ++pos) {
      double mean, var;
      trun_norm_moments(mu, sigma, 0.0, pos == 1, &mean, &var);
      rows.push_back(row_json({{"mu", mu}, {"sigma", sigma}, {"positive", (double)pos}, {"mean", mean}, {"variance", var}}));
    }
  }
  write_file(dir + "/trun_norm_moments.json", list_json(rows));
}

// ------------------------------------------------------------------------------------
void golden_suf(const std::string &dir) {
  const int n = 64, p = 7;
  std::vector<double> X(n * p), w(n), s(n), yy(n);
  bo_synth_x(n, p, 77, 1.0, 0, X.data(), p);
  for (int i = 0; i < n; ++i) {
    double u[2]; bo_uniform_pair(78, 0, i, 0, u);
    w[i] = 0.05 + 1.3 * u[0];
    s[i] = 6.0 * (u[1] - 0.5);
    yy[i] = 4.0 * (u[1] - 0.3);
  }
  BinomialLogit::SufficientStatistics suf(p);
  WeightedRegSuf wsuf(p);
  for (int i = 0; i < n; ++i) {
    Vector x(p); for (int j = 0; j < p; ++j) x[j] = X[i * p + j];
    suf.update(x, s[i], w[i]);
    wsuf.add_data(x, yy[i], w[i]);
  }
  Json j;
  j.num("n", n); j.num("p", p);
  j.arr("X", X); j.arr("weight", w); j.arr("weighted_value", s); j.arr("y", yy);
  SpdMatrix xtx = suf.xtx();
  std::vector<double> flat(xtx.data(), xtx.data() + p * p);
  j.arr("xtx_colmajor", flat); j.arr("xty", suf.xty()); j.num("sample_size", suf.sample_size());
  SpdMatrix wx = wsuf.xtx();
  std::vector<double> wflat(wx.data(), wx.data() + p * p);
  j.arr("w_xtwx_colmajor", wflat); j.arr("w_xtwy", wsuf.xty());
  std::vector<double> sc = {wsuf.n(), wsuf.yty(), wsuf.sumw(), wsuf.sumlogw()};
  j.arr("w_scalars", sc);
  write_file(dir + "/suf.json", j.str());
}

// ------------------------------------------------------------------------------------
void golden_poisson_table(const std::string &dir) {
  NormalMixtureApproximationTable table = create_poisson_mixture_approximation_table();
  Vector grid = table.serialize();
  // Materialise every nu up to 300 the way the reference does on first touch
  // (interpolation or Powell re-fit, NormalMixtureApproximation.cpp:472-532).
  for (int nu = 1; nu <= 300; ++nu) table.approximate(nu);
  Vector ser = table.serialize();
  Json j;
  j.num("smallest_index", table.smallest_index());
  j.num("largest_index", table.largest_index());
  j.num("grid_serialized_length", grid.size());
  j.arr("serialized", ser);
  write_file(dir + "/poisson_mixture_table.json", j.str());

  std::vector<std::string> rows;
  RNG rng(3);
  int nus[] = {1, 2, 3, 5, 9, 19, 20, 33, 49, 50, 77, 100, 103, 150, 199, 250, 300, 500, 1000, 5000, 20000, 30000, 45000};
  int c = 0;
  for (int nu : nus) {
    for (int r = 0; r < 12; ++r, ++c) {
      double center = -std::log((double)nu);
      double sd = 1.0 / std::sqrt((double)nu);
      double resid = center + sd * (-3.5 + 7.0 * r / 11.0);
      unsigned long seed = 20000 + c;
      rng.seed(seed);
      double mu, sigsq;
      unmix_poisson_augmented_data(rng, resid, nu, &mu, &sigsq, &table);
      rng.seed(seed);
      double U = rng();
      rows.push_back(row_json({{"nu", (double)nu}, {"resid", resid}, {"U", U}, {"mu", mu}, {"sigsq", sigsq}}));
    }
  }
  write_file(dir + "/unmix_poisson.json", list_json(rows));
}

// ------------------------------------------------------------------------------------
// Off-grid counts: what NormalMixtureApproximationTable::approximate(nu) (NormalMixtureApproximation.cpp:472-532) returns from a
// FRESH table (neighbours = grid entries) -- interpolated entries, Powell re-fits -- with its Kullback-Leibler divergence.
void golden_poisson_offgrid(const std::string &dir) {
  std::vector<std::string> rows;
  int nus[] = {305, 333, 477, 495, 777, 1234, 2950, 4321, 5500, 12345, 29500};
  for (int nu : nus) {
    NormalMixtureApproximationTable table = create_poisson_mixture_approximation_table();
    std::vector<int> idx;
    {
      Vector ser = table.serialize();
      size_t i = 0;
      while (i < ser.size()) { idx.push_back((int)lround(ser[i])); i += 2 + 3 * (size_t)lround(ser[i + 1]); }
    }
    auto ub = std::upper_bound(idx.begin(), idx.end(), nu);
    int nu1 = *ub, nu0 = *(ub - 1);
    int k0 = table.approximate(nu0).dim(), k1 = table.approximate(nu1).dim();
    NormalMixtureApproximation a = table.approximate(nu);
    NegLogGamma target(nu);
    double kl = a.kullback_leibler(target);
    std::ostringstream o; o << std::setprecision(17) << "{\"nu\": " << nu << ", \"nu0\": " << nu0 << ", \"nu1\": " << nu1
                            << ", \"k0\": " << k0 << ", \"k1\": " << k1 << ", \"kl\": " << kl << ", \"mu\": [";
    for (int k = 0; k < a.dim(); ++k) o << (k ? ", " : "") << a.mu()[k];
    o << "], \"sigma\": [";
    for (int k = 0; k < a.dim(); ++k) o << (k ? ", " : "") << a.sigma()[k];
    o << "], \"weights\": [";
    for (int k = 0; k < a.dim(); ++k) o << (k ? ", " : "") << a.weights()[k];
    o << "]}";
    rows.push_back(o.str());
  }
  write_file(dir + "/poisson_offgrid.json", list_json(rows));
}

// ------------------------------------------------------------------------------------
void golden_loglike(const std::string &dir) {
  const int n = 48, p = 4;
  std::vector<double> X(n * p), y(n), nt(n), beta(p);
  bo_synth_x(n, p, 91, 1.0, 0, X.data(), p);
  bo_synth_beta(p, 2, -0.7, beta.data());
  bo_synth_binomial_y(n, p, X.data(), p, beta.data(), 92, 30, 0, y.data(), nt.data());
  NEW(BinomialLogitModel, model)(p);
  for (int i = 0; i < n; ++i) {
    Vector x(p); for (int j = 0; j < p; ++j) x[j] = X[i * p + j];
    NEW(BinomialRegressionData, dp)(y[i], nt[i], x);
    model->add_data(dp);
  }
  Vector b(p); for (int j = 0; j < p; ++j) b[j] = beta[j] * 0.8 + 0.05;
  Json j;
  j.num("n", n); j.num("p", p); j.arr("X", X); j.arr("y", y); j.arr("ntrials", nt); j.arr("beta", b);
  j.num("binomial_loglike", model->log_likelihood(b, nullptr, nullptr));
  {
    Vector g; Matrix h;
    double ll = model->log_likelihood(b, &g, &h);
    j.num("binomial_loglike_d", ll); j.arr("binomial_gradient", g);
    std::vector<double> hh; for (int a = 0; a < p; ++a) for (int c = 0; c < p; ++c) hh.push_back(h(a, c));
    j.arr("binomial_hessian", hh);
    model->set_nonevent_sampling_prob(0.25);   // log_alpha = log(0.25): BinomialLogitModel.cpp:168
    double ll2 = model->log_likelihood(b, &g, &h);
    j.num("binomial_log_alpha", std::log(0.25)); j.num("binomial_loglike_alpha", ll2); j.arr("binomial_gradient_alpha", g);
    model->set_nonevent_sampling_prob(1.0);
  }

  std::vector<int64_t> yp(n); std::vector<double> ex(n), bp(p);
  bo_synth_beta(p, 2, 0.5, bp.data());
  std::vector<double> Xp(n * p);
  bo_synth_x(n, p, 93, 0.3, 0, Xp.data(), p);
  bo_synth_poisson_y(n, p, Xp.data(), p, bp.data(), 94, 0, yp.data(), ex.data());
  NEW(PoissonRegressionModel, pm)(p);
  for (int i = 0; i < n; ++i) {
    Vector x(p); for (int jj = 0; jj < p; ++jj) x[jj] = Xp[i * p + jj];
    ex[i] = 0.5 + 0.1 * (i % 7);
    NEW(PoissonRegressionData, dp)(yp[i], x, ex[i]);
    pm->add_data(dp);
  }
  Vector b2(p); for (int jj = 0; jj < p; ++jj) b2[jj] = bp[jj] * 0.9 - 0.02;
  j.arr("poisson_X", Xp); j.arr("poisson_y", yp); j.arr("poisson_exposure", ex); j.arr("poisson_beta", b2);
  j.num("poisson_loglike", pm->log_likelihood(b2, nullptr, nullptr));
  {
    Vector g; Matrix h;
    double ll = pm->log_likelihood(b2, &g, &h);
    j.num("poisson_loglike_d", ll); j.arr("poisson_gradient", g);
    std::vector<double> hh; for (int a = 0; a < p; ++a) for (int c = 0; c < p; ++c) hh.push_back(h(a, c));
    j.arr("poisson_hessian", hh);
  }

  std::vector<std::string> rows;
  double ns[] = {1, 1, 2, 5, 12, 16, 30, 40, 100, 700};
  for (double nn : ns)
    for (double pr : {1e-4, 0.03, 0.2, 0.5, 0.77, 0.95, 0.9999})
      for (double frac : {0.0, 0.25, 0.5, 1.0}) {
        double x = std::floor(frac * nn);
        rows.push_back(row_json({{"x", x}, {"n", nn}, {"p", pr}, {"logd", dbinom(x, nn, pr, true)}}));
      }
  j.raw("dbinom", list_json(rows));
  write_file(dir + "/loglike.json", j.str());
}

// ------------------------------------------------------------------------------------
// Monte Carlo summaries of the reference's own imputers (their RNG, their algorithms).
// 19 quantiles (5 %, ..., 95 %) of a sample, for two-sample chi-square tests on quantile bins
std::string quantiles_json(std::vector<double> v) {
  std::sort(v.begin(), v.end());
  std::ostringstream o; o << std::setprecision(17) << "[";
  for (int j = 1; j < 20; ++j) o << (j > 1 ? ", " : "") << v[(size_t)((double)j / 20.0 * (v.size() - 1))];
  o << "]";
  return o.str();
}

void golden_draw_stats(const std::string &dir) {
  const LogitMixtureApproximation &m(BinomialLogitDataImputer::mixture_approximation);
  RNG rng(424242);
  BinomialLogitCltDataImputer imputer(10);
  std::vector<std::string> rows;
  const int N = 400000;
  for (double eta : {-3.0, -1.0, 0.0, 0.5, 2.0}) {
    for (int y = 0; y < 2; ++y) {
      std::vector<double> kc(9, 0.0);
      std::vector<double> zs; zs.reserve(N);
      double s1 = 0, s2 = 0, w1 = 0, w2 = 0;
      for (int i = 0; i < N; ++i) {
        std::pair<double, double> a = imputer.impute(rng, 1, y, eta);
        double info = a.second, z = a.first / info;
        zs.push_back(z);
        kc[sigma_index(m.sigma(), 1.0 / info)] += 1;
        s1 += z; s2 += z * z; w1 += info; w2 += info * info;
      }
      std::ostringstream o; o << std::setprecision(17);
      o << "{\"eta\": " << eta << ", \"y\": " << y << ", \"N\": " << N << ", \"z_mean\": " << s1 / N
        << ", \"z_var\": " << s2 / N - (s1 / N) * (s1 / N) << ", \"info_mean\": " << w1 / N << ", \"info_var\": "
        << w2 / N - (w1 / N) * (w1 / N) << ", \"kcount\": [";
      for (int k = 0; k < 9; ++k) o << (k ? ", " : "") << kc[k];
      o << "], \"z_quantiles\": " << quantiles_json(zs) << "}";
      rows.push_back(o.str());
    }
  }
  write_file(dir + "/ref_logit_small_stats.json", list_json(rows));

  rows.clear();
  struct C { double n, y, eta; };
  const int N2 = 200000;
  for (C c : {C{25, 7, -0.8}, C{200, 150, 1.2}, C{12, 0, -2.0}, C{11, 11, 3.0}, C{1000, 480, -0.1}}) {
    double s1 = 0, s2 = 0, w1 = 0, w2 = 0;
    std::vector<double> sums, infos;
    for (int i = 0; i < N2; ++i) {
      std::pair<double, double> a = imputer.impute(rng, c.n, c.y, c.eta);
      s1 += a.first; s2 += a.first * a.first; w1 += a.second; w2 += a.second * a.second;
      sums.push_back(a.first); infos.push_back(a.second);
    }
    std::string r = row_json({{"ntrials", c.n}, {"y", c.y}, {"eta", c.eta}, {"N", (double)N2}, {"sum_mean", s1 / N2},
                              {"sum_var", s2 / N2 - (s1 / N2) * (s1 / N2)}, {"info_mean", w1 / N2},
                              {"info_var", w2 / N2 - (w1 / N2) * (w1 / N2)}});
    r.pop_back();
    r += ", \"sum_quantiles\": " + quantiles_json(sums) + ", \"info_quantiles\": " + quantiles_json(infos) + "}";
    rows.push_back(r);
  }
  write_file(dir + "/ref_logit_clt_stats.json", list_json(rows));

  rows.clear();
  PoissonDataImputer pimp;
  struct P { int y; double E, eta; };
  for (P c : {P{0, 1.0, -1.0}, P{0, 2.5, 0.7}, P{1, 1.0, 0.0}, P{3, 1.0, 0.7}, P{3, 2.5, -1.0}, P{12, 1.0, 2.0},
              P{60, 2.5, 3.0}, P{150, 1.0, 5.0}}) {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    std::vector<double> zes, zis;
    for (int i = 0; i < N2; ++i) {
      double zi = 0, mi = 0, wi = 0, ze = 0, me = 0, we = 0;
      pimp.impute(rng, c.y, c.E, c.eta, &zi, &mi, &wi, &ze, &me, &we);
      zes.push_back(ze); if (c.y > 0) zis.push_back(zi);
      acc[0] += ze; acc[1] += ze * ze; acc[2] += we; acc[3] += (ze - me) * we;
      if (c.y > 0) { acc[4] += zi; acc[5] += zi * zi; acc[6] += wi; acc[7] += (zi - mi) * wi; }
    }
    for (double &a : acc) a /= N2;
    std::string r = row_json({{"y", (double)c.y}, {"exposure", c.E}, {"eta", c.eta}, {"N", (double)N2},
                              {"zext_mean", acc[0]}, {"zext_var", acc[1] - acc[0] * acc[0]}, {"wext_mean", acc[2]},
                              {"rwext_mean", acc[3]}, {"zint_mean", acc[4]}, {"zint_var", acc[5] - acc[4] * acc[4]},
                              {"wint_mean", acc[6]}, {"rwint_mean", acc[7]}});
    r.pop_back();
    r += ", \"zext_quantiles\": " + quantiles_json(zes);
    if (c.y > 0) r += ", \"zint_quantiles\": " + quantiles_json(zis);
    r += "}";
    rows.push_back(r);
  }
  write_file(dir + "/ref_poisson_stats.json", list_json(rows));
}

// ------------------------------------------------------------------------------------
struct Moments {
  std::vector<double> s1, s2; int n = 0;
  explicit Moments(int p) : s1(p, 0.0), s2(p, 0.0) {}
  void add(const Vector &b) { for (size_t j = 0; j < s1.size(); ++j) { s1[j] += b[j]; s2[j] += b[j] * b[j]; } ++n; }
  std::vector<double> mean() const { std::vector<double> m(s1); for (double &v : m) v /= n; return m; }
  std::vector<double> sd() const {
    std::vector<double> m = mean(), o(s1.size());
    for (size_t j = 0; j < o.size(); ++j) o[j] = std::sqrt(std::max(0.0, s2[j] / n - m[j] * m[j]));
    return o;
  }
};

Ptr<BinomialLogitModel> make_logit_model(int64_t n, int p, int nonzero, uint64_t seed, int max_trials,
                                         std::vector<double> *beta_true) {
  std::vector<double> X((size_t)n * p), y(n), nt(n), beta(p);
  bo_synth_x(n, p, seed, 1.0, 0, X.data(), p);
  bo_synth_beta(p, nonzero, -1.0, beta.data());
  bo_synth_binomial_y(n, p, X.data(), p, beta.data(), seed, max_trials, 0, y.data(), nt.data());
  NEW(BinomialLogitModel, model)(p);
  Vector x(p);
  for (int64_t i = 0; i < n; ++i) {
    for (int j = 0; j < p; ++j) x[j] = X[i * p + j];
    NEW(BinomialRegressionData, dp)(y[i], nt[i], x);
    model->add_data(dp);
  }
  if (beta_true) *beta_true = beta;
  return model;
}

Ptr<PoissonRegressionModel> make_poisson_model(int64_t n, int p, int nonzero, uint64_t seed,
                                               std::vector<double> *beta_true) {
  std::vector<double> X((size_t)n * p), ex(n), beta(p);
  std::vector<int64_t> y(n);
  bo_synth_x(n, p, seed, 0.3, 0, X.data(), p);
  bo_synth_beta(p, nonzero, 0.5, beta.data());
  bo_synth_poisson_y(n, p, X.data(), p, beta.data(), seed, 0, y.data(), ex.data());
  NEW(PoissonRegressionModel, model)(p);
  Vector x(p);
  for (int64_t i = 0; i < n; ++i) {
    for (int j = 0; j < p; ++j) x[j] = X[i * p + j];
    NEW(PoissonRegressionData, dp)(y[i], x, ex[i]);
    model->add_data(dp);
  }
  if (beta_true) *beta_true = beta;
  return model;
}

void golden_chains(const std::string &dir) {
  Json j;
  {
    const int n = 4000, p = 5, iters = 6000, burn = 1000;
    std::vector<double> bt;
    Ptr<BinomialLogitModel> model = make_logit_model(n, p, 3, 1234, 1, &bt);
    NEW(MvnModel, prior)(Vector(p, 0.0), SpdMatrix(p, 1.0));
    GlobalRng::rng.seed(8675309);
    NEW(BinomialLogitAuxmixSampler, sampler)(model.get(), prior, 10);
    model->set_method(sampler);
    Moments mo(p);
    for (int it = 0; it < iters; ++it) { model->sample_posterior(); if (it >= burn) mo.add(model->Beta()); }
    j.raw("logit_auxmix", "{\"n\": 4000, \"p\": 5, \"nonzero\": 3, \"seed\": 1234, \"max_trials\": 1, \"iters\": 6000, \"burn\": 1000}");
    j.arr("logit_auxmix_beta_true", bt); j.arr("logit_auxmix_mean", mo.mean()); j.arr("logit_auxmix_sd", mo.sd());
  }
  {
    // binomial responses with up to 40 trials: exercises the CLT branch (clt_threshold 10)
    const int n = 1500, p = 4, iters = 6000, burn = 1000;
    std::vector<double> bt;
    Ptr<BinomialLogitModel> model = make_logit_model(n, p, 2, 4321, 40, &bt);
    NEW(MvnModel, prior)(Vector(p, 0.0), SpdMatrix(p, 1.0));
    NEW(BinomialLogitAuxmixSampler, sampler)(model.get(), prior, 10);
    model->set_method(sampler);
    Moments mo(p);
    for (int it = 0; it < iters; ++it) { model->sample_posterior(); if (it >= burn) mo.add(model->Beta()); }
    j.raw("logit_binomial", "{\"n\": 1500, \"p\": 4, \"nonzero\": 2, \"seed\": 4321, \"max_trials\": 40, \"iters\": 6000, \"burn\": 1000}");
    j.arr("logit_binomial_beta_true", bt); j.arr("logit_binomial_mean", mo.mean()); j.arr("logit_binomial_sd", mo.sd());
  }
  {
    const int n = 3000, p = 12, iters = 6000, burn = 1000;
    std::vector<double> bt;
    Ptr<BinomialLogitModel> model = make_logit_model(n, p, 3, 2468, 1, &bt);
    NEW(MvnModel, slab)(Vector(p, 0.0), SpdMatrix(p, 1.0));
    NEW(VariableSelectionPrior, spike)(p, 0.25);
    NEW(BinomialLogitSpikeSlabSampler, sampler)(model.get(), slab, spike, 10);
    model->set_method(sampler);
    Moments mo(p), inc(p);
    for (int it = 0; it < iters; ++it) {
      model->sample_posterior();
      if (it >= burn) {
        mo.add(model->Beta());
        Vector g(p); for (int k = 0; k < p; ++k) g[k] = model->coef().inc()[k] ? 1.0 : 0.0;
        inc.add(g);
      }
    }
    j.raw("logit_spike_slab", "{\"n\": 3000, \"p\": 12, \"nonzero\": 3, \"seed\": 2468, \"max_trials\": 1, \"iters\": 6000, \"burn\": 1000, \"prior_inclusion\": 0.25}");
    j.arr("logit_spike_slab_beta_true", bt); j.arr("logit_spike_slab_mean", mo.mean());
    j.arr("logit_spike_slab_sd", mo.sd()); j.arr("logit_spike_slab_inclusion", inc.mean());
  }
  {
    const int n = 3000, p = 5, iters = 6000, burn = 1000;
    std::vector<double> bt;
    Ptr<PoissonRegressionModel> model = make_poisson_model(n, p, 3, 1357, &bt);
    NEW(MvnModel, prior)(Vector(p, 0.0), SpdMatrix(p, 1.0));
    NEW(PoissonRegressionAuxMixSampler, sampler)(model.get(), prior, 1);
    model->set_method(sampler);
    Moments mo(p);
    for (int it = 0; it < iters; ++it) { model->sample_posterior(); if (it >= burn) mo.add(model->Beta()); }
    j.raw("poisson_auxmix", "{\"n\": 3000, \"p\": 5, \"nonzero\": 3, \"seed\": 1357, \"iters\": 6000, \"burn\": 1000}");
    j.arr("poisson_auxmix_beta_true", bt); j.arr("poisson_auxmix_mean", mo.mean()); j.arr("poisson_auxmix_sd", mo.sd());
  }
  write_file(dir + "/ref_chains.json", j.str());
}

// ------------------------------------------------------------------------------------
// The Student-t sibling: the reference's TDataImputer draws, TRegressionModel::log_likelihood and a TRegressionSampler chain.
void golden_student(const std::string &dir) {
  Json j;
  {
    RNG rng(20261017);
    TDataImputer imputer;
    struct C { double residual, sigma, nu; };
    std::vector<std::string> rows;
    const int N = 200000;
    for (C c : {C{0.0, 1.0, 4.0}, C{2.5, 1.5, 4.0}, C{-7.0, 0.8, 2.0}, C{0.3, 2.0, 30.0}, C{1.0, 1.0, 0.6}, C{40.0, 1.0, 1.0}}) {
      std::vector<double> w; w.reserve(N);
      double s1 = 0, s2 = 0, sl = 0;
      for (int i = 0; i < N; ++i) {
        double v = imputer.impute(rng, c.residual, c.sigma, c.nu);
        w.push_back(v); s1 += v; s2 += v * v; sl += std::log(v);
      }
      std::string r = row_json({{"residual", c.residual}, {"sigma", c.sigma}, {"nu", c.nu}, {"N", (double)N}, {"w_mean", s1 / N},
                                {"w_var", s2 / N - (s1 / N) * (s1 / N)}, {"logw_mean", sl / N}});
      r.pop_back();
      r += ", \"w_quantiles\": " + quantiles_json(w) + "}";
      rows.push_back(r);
    }
    j.raw("draw_stats", list_json(rows));
  }
  const double sigma_true = 1.5, nu_true = 4.0;
  auto make_model = [&](int64_t n, int p, int nonzero, uint64_t seed, std::vector<double> *beta_true) {
    std::vector<double> X((size_t)n * p), y(n), beta(p);
    bo_synth_x(n, p, seed, 1.0, 0, X.data(), p);
    bo_synth_beta(p, nonzero, 0.5, beta.data());
    bo_synth_student_y(n, p, X.data(), p, beta.data(), sigma_true, nu_true, seed, 0, y.data());
    NEW(TRegressionModel, model)(p);
    Vector x(p);
    for (int64_t i = 0; i < n; ++i) {
      for (int k = 0; k < p; ++k) x[k] = X[i * p + k];
      NEW(RegressionData, dp)(y[i], x);
      model->add_data(dp);
    }
    if (beta_true) *beta_true = beta;
    return model;
  };
  {
    // TRegressionModel::log_likelihood(beta, sigsq, nu) on synth_student(n = 500, p = 7, nonzero = 3, seed = 99)
    std::vector<double> bt;
    Ptr<TRegressionModel> model = make_model(500, 7, 3, 99, &bt);
    Vector beta(bt.size());
    for (size_t k = 0; k < bt.size(); ++k) beta[k] = bt[k] * 0.9 + 0.05;
    std::vector<std::string> rows;
    for (double sigma : {0.7, 1.5, 3.0}) for (double nu : {0.8, 4.0, 30.0, 200.0})
      rows.push_back(row_json({{"sigma", sigma}, {"nu", nu}, {"loglike", model->log_likelihood(beta, sigma * sigma, nu)}}));
    j.raw("loglike", "{\"n\": 500, \"p\": 7, \"nonzero\": 3, \"seed\": 99, \"beta_scale\": 0.9, \"beta_shift\": 0.05, \"rows\": " + list_json(rows) + "}");
  }
  {
    const int n = 3000, p = 5, iters = 8000, burn = 1000;
    std::vector<double> bt;
    Ptr<TRegressionModel> model = make_model(n, p, 3, 777, &bt);
    NEW(MvnModel, prior)(Vector(p, 0.0), SpdMatrix(p, 100.0));
    NEW(ChisqModel, siginv_prior)(1.0, 1.0);
    NEW(UniformModel, nu_prior)(0.5, 60.0);
    GlobalRng::rng.seed(8675309);
    NEW(TRegressionSampler, sampler)(model.get(), prior, siginv_prior, nu_prior);
    model->set_method(sampler);
    Moments mo(p), sn(2);
    for (int it = 0; it < iters; ++it) {
      model->sample_posterior();
      if (it >= burn) { mo.add(model->Beta()); Vector v(2); v[0] = model->sigma(); v[1] = model->nu(); sn.add(v); }
    }
    j.raw("chain", "{\"n\": 3000, \"p\": 5, \"nonzero\": 3, \"seed\": 777, \"sigma_true\": 1.5, \"nu_true\": 4.0, \"iters\": 8000, \"burn\": 1000, "
                   "\"beta_prior_variance\": 100.0, \"siginv_prior\": [1.0, 1.0], \"nu_prior\": [0.5, 60.0]}");
    j.arr("chain_beta_true", bt); j.arr("chain_beta_mean", mo.mean()); j.arr("chain_beta_sd", mo.sd());
    j.arr("chain_sigma_nu_mean", sn.mean()); j.arr("chain_sigma_nu_sd", sn.sd());
  }
  {
    // TRegressionSpikeSlabSampler (what lm.spike builds for Student errors) on synth_student(n = 3000, p = 12, 3 non-zero slopes)
    const int n = 3000, p = 12, iters = 8000, burn = 1000;
    std::vector<double> bt;
    Ptr<TRegressionModel> model = make_model(n, p, 3, 778, &bt);
    NEW(MvnModel, slab)(Vector(p, 0.0), SpdMatrix(p, 4.0));
    NEW(VariableSelectionPrior, spike)(p, 0.3);
    NEW(ChisqModel, siginv_prior)(1.0, 1.0);
    NEW(UniformModel, nu_prior)(0.5, 60.0);
    model->coef().drop_all(); model->coef().add(0);
    NEW(TRegressionSpikeSlabSampler, sampler)(model.get(), slab, spike, siginv_prior, nu_prior);
    model->set_method(sampler);
    Moments mo(p), inc(p), sn(2);
    for (int it = 0; it < iters; ++it) {
      model->sample_posterior();
      if (it >= burn) {
        mo.add(model->Beta());
        Vector g(p); for (int k = 0; k < p; ++k) g[k] = model->coef().inc()[k] ? 1.0 : 0.0;
        inc.add(g);
        Vector v(2); v[0] = model->sigma(); v[1] = model->nu(); sn.add(v);
      }
    }
    j.raw("spike_chain", "{\"n\": 3000, \"p\": 12, \"nonzero\": 3, \"seed\": 778, \"sigma_true\": 1.5, \"nu_true\": 4.0, \"iters\": 8000, \"burn\": 1000, "
                         "\"slab_variance\": 4.0, \"prior_inclusion\": 0.3, \"siginv_prior\": [1.0, 1.0], \"nu_prior\": [0.5, 60.0]}");
    j.arr("spike_chain_beta_true", bt); j.arr("spike_chain_beta_mean", mo.mean()); j.arr("spike_chain_beta_sd", mo.sd());
    j.arr("spike_chain_inclusion", inc.mean());
    j.arr("spike_chain_sigma_nu_mean", sn.mean()); j.arr("spike_chain_sigma_nu_sd", sn.sd());
  }
  write_file(dir + "/ref_student.json", j.str());
}

// ------------------------------------------------------------------------------------
// bench <logit|spike|poisson> n p nonzero threads iters warmup
int run_bench(int argc, char **argv) {
  if (argc < 9) { fprintf(stderr, "usage: bench model n p nonzero threads iters warmup [zellner]\n"); return 2; }
  std::string kind = argv[2];
  int64_t n = atoll(argv[3]); int p = atoi(argv[4]); int nonzero = atoi(argv[5]);
  int threads = atoi(argv[6]); int iters = atoi(argv[7]); int warm = atoi(argv[8]);
  const bool zellner = argc >= 10 && std::string(argv[9]) == "zellner";
  auto t0 = std::chrono::steady_clock::now();
  double secs = 0;
  GlobalRng::rng.seed(8675309);
  auto timed = [&](auto &model) {
    for (int i = 0; i < warm; ++i) model->sample_posterior();
    auto a = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; ++i) model->sample_posterior();
    secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
  };
  if (kind == "poisson") {
    Ptr<PoissonRegressionModel> model = make_poisson_model(n, p, nonzero, 20261017, nullptr);
    NEW(MvnModel, prior)(Vector(p, 0.0), SpdMatrix(p, 1.0));
    NEW(PoissonRegressionAuxMixSampler, sampler)(model.get(), prior, threads);
    model->set_method(sampler);
    timed(model);
  } else {
    Ptr<BinomialLogitModel> model = make_logit_model(n, p, nonzero, 20261017, 1, nullptr);
    Ptr<MvnModel> prior(new MvnModel(Vector(p, 0.0), SpdMatrix(p, 1.0)));
    if (zellner) {
      // LogitZellnerPrior (Interfaces/python/spikeslab/BayesBoom/spikeslab/priors.py:385-462) with its defaults:
      // precision = X'X / n with the off-diagonal halved (diagonal_shrinkage = .5), mean = (trimmed logit of mean(y/n), 0, ...)
      const std::vector<Ptr<BinomialRegressionData>> &dat(model->dat());
      int nt = std::max(1, threads);
      std::vector<std::vector<double>> part(nt, std::vector<double>((size_t)p * p, 0.0));
      std::vector<double> ysum(nt, 0.0);
      std::vector<std::thread> pool;
      for (int t = 0; t < nt; ++t) pool.emplace_back([&, t]() {
        std::vector<double> &a(part[t]);
        for (int64_t r = n * t / nt; r < n * (t + 1) / nt; ++r) {
          const Vector &x(dat[r]->x());
          ysum[t] += dat[r]->y() / dat[r]->n();
          for (int i = 0; i < p; ++i) { const double xi = x[i]; double *row = a.data() + (size_t)i * p; for (int j = 0; j <= i; ++j) row[j] += xi * x[j]; }
        }
      });
      for (auto &th : pool) th.join();
      SpdMatrix prec(p, 0.0);
      double ys = 0;
      for (int t = 0; t < nt; ++t) ys += ysum[t];
      for (int i = 0; i < p; ++i) for (int j = 0; j <= i; ++j) {
        double v = 0;
        for (int t = 0; t < nt; ++t) v += part[t][(size_t)i * p + j];
        v /= (double)n;
        if (i != j) v *= 0.5;
        prec(i, j) = v; prec(j, i) = v;
      }
      Vector mean(p, 0.0);
      double ph = std::min(.999, std::max(.001, ys / (double)n));
      mean[0] = std::log(ph / (1 - ph));
      prior = new MvnModel(mean, prec, true);
    }
    if (kind == "spike") {
      model->coef().drop_all(); model->coef().add(0);
      NEW(VariableSelectionPrior, spike)(p, std::min(1.0, (double)std::max(nonzero, 1) / p));
      NEW(BinomialLogitSpikeSlabSampler, sampler)(model.get(), prior, spike, 10);
      sampler->set_number_of_workers(threads);
      model->set_method(sampler);
    } else {
      NEW(BinomialLogitAuxmixSampler, sampler)(model.get(), prior, 10);
      sampler->set_number_of_workers(threads);
      model->set_method(sampler);
    }
    timed(model);
  }
  double setup = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() - secs;
  printf("{\"kind\": \"%s\", \"n\": %lld, \"p\": %d, \"threads\": %d, \"iters\": %d, \"warmup\": %d, \"seconds\": %.6f, "
         "\"iters_per_sec\": %.6f, \"obs_per_sec\": %.3f, \"setup_seconds\": %.3f, \"hw_threads\": %u}\n",
         kind.c_str(), (long long)n, p, threads, iters, warm, secs, iters / secs, iters / secs * n, setup,
         std::thread::hardware_concurrency());
  return 0;
}

}  // namespace

int main(int argc, char **argv) {
  try {
    if (argc >= 3 && std::string(argv[1]) == "golden") {
      std::string dir = argv[2];
      std::string what = argc >= 4 ? argv[3] : "all";
      if (what == "all" || what == "mixture") golden_mixture(dir);
      if (what == "all" || what == "suf") golden_suf(dir);
      if (what == "all" || what == "table") golden_poisson_table(dir);
      if (what == "all" || what == "offgrid") golden_poisson_offgrid(dir);
      if (what == "all" || what == "loglike") golden_loglike(dir);
      if (what == "all" || what == "stats") golden_draw_stats(dir);
      if (what == "all" || what == "chains") golden_chains(dir);
      if (what == "all" || what == "student") golden_student(dir);
      return 0;
    }
    if (argc >= 2 && std::string(argv[1]) == "bench") return run_bench(argc, argv);
  } catch (std::exception &e) {
    fprintf(stderr, "reference threw: %s\n", e.what());
    return 1;
  }
  fprintf(stderr, "usage: %s golden <dir> [what] | bench <logit|spike|poisson> n p nonzero threads iters warmup\n", argv[0]);
  return 2;
}
