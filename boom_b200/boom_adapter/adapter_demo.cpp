// adapter_demo.cpp -- TEST DRIVER (links the compiled reference; the binary lands in oracle/_ref/).
//
// The drop-in claim, executed: a BOOM program builds BOOM's own BinomialLogitModel / PoissonRegressionModel,
// then attaches EITHER the reference sampler OR the B200 sampler with model->set_method(sampler) and calls
// model->sample_posterior().  Prints one JSON line with the posterior summaries of both chains on the same
// data; tests/test_gpu_adapter.py compares them within Monte Carlo error.
//   usage: boom_adapter_demo <logit|spike|poisson|pspike|mode|pmode|fixed|pfixed|bench|api|composite|chunk|probit|treg|tspike|tactive|active|pactive> n p nonzero iters burn
//          (mode / pmode: find_posterior_mode; fixed / pfixed: externally driven statistics, host steps only, no GPU;
//           bench: ms per iteration of the adapter beside the standalone classes; api: the public surface beyond draw())
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <sstream>
#include <string>
#include <vector>

#include "Models/Glm/BinomialLogitModel.hpp"
#include "Models/Glm/BinomialRegressionData.hpp"
#include "Models/Glm/PoissonRegressionData.hpp"
#include "Models/Glm/PoissonRegressionModel.hpp"
#include "Models/Glm/PosteriorSamplers/BinomialLogitAuxmixSampler.hpp"
#include "Models/Glm/PosteriorSamplers/BinomialLogitCompositeSpikeSlabSampler.hpp"
#include "Models/Glm/PosteriorSamplers/BinomialLogitSpikeSlabSampler.hpp"
#include "Models/Glm/BinomialProbitModel.hpp"
#include "Models/Glm/PosteriorSamplers/BinomialProbitSpikeSlabSampler.hpp"
#include "Models/Glm/PosteriorSamplers/PoissonRegressionAuxMixSampler.hpp"
#include "Models/Glm/PosteriorSamplers/PoissonRegressionSpikeSlabSampler.hpp"
#include "Models/ChisqModel.hpp"
#include "Models/Glm/PosteriorSamplers/TRegressionSampler.hpp"
#include "Models/Glm/PosteriorSamplers/TRegressionSpikeSlabSampler.hpp"
#include "Models/Glm/TRegression.hpp"
#include "Models/Glm/VariableSelectionPrior.hpp"
#include "Models/MvnModel.hpp"
#include "Models/UniformModel.hpp"
#include "distributions.hpp"

#include "boom_b200_adapter.hpp"

using namespace BOOM;

namespace {
struct Summary { Vector mean, sd, inc; double secs; };

template <class MODEL>
Summary run(const Ptr<MODEL> &model, int iters, int burn) {
  const int p = model->xdim();
  Vector s1(p, 0.0), s2(p, 0.0), inc(p, 0.0);
  auto t0 = std::chrono::steady_clock::now();
  for (int it = 0; it < iters; ++it) {
    model->sample_posterior();
    if (it >= burn) {
      const Vector &b(model->Beta());
      for (int j = 0; j < p; ++j) { s1[j] += b[j]; s2[j] += b[j] * b[j]; inc[j] += model->coef().inc()[j]; }
    }
  }
  Summary out;
  out.secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  const double m = iters - burn;
  out.mean = s1 / m; out.sd = Vector(p); out.inc = inc / m;
  for (int j = 0; j < p; ++j) out.sd[j] = std::sqrt(std::max(0.0, s2[j] / m - out.mean[j] * out.mean[j]));
  return out;
}

void print_vec(const char *name, const Vector &v, bool comma = true) {
  printf("\"%s\": [", name);
  for (size_t i = 0; i < v.size(); ++i) printf("%s%.10g", i ? ", " : "", v[i]);
  printf("]%s", comma ? ", " : "");
}
void print_summary(const char *tag, const Summary &s) {
  printf("\"%s\": {", tag);
  print_vec("mean", s.mean); print_vec("sd", s.sd); print_vec("inclusion", s.inc);
  printf("\"seconds\": %.4f}", s.secs);
}
}  // namespace

int main(int argc, char **argv) {
  if (argc < 7) { fprintf(stderr, "usage: %s <logit|spike|poisson|pspike> n p nonzero iters burn\n", argv[0]); return 2; }
  const std::string kind = argv[1];
  const int n = atoi(argv[2]), p = atoi(argv[3]), nonzero = atoi(argv[4]), iters = atoi(argv[5]), burn = atoi(argv[6]);
  try {
    GlobalRng::rng.seed(20261017);
    const bool poisson = kind == "poisson" || kind == "pspike" || kind == "pmode" || kind == "pfixed" || kind == "pactive";
    Vector beta(p, 0.0);
    beta[0] = poisson ? 0.5 : -1.0;
    for (int j = 1; j <= nonzero && j < p; ++j) beta[j] = (j % 2) ? 0.5 : -0.5;
    std::vector<Vector> xs;
    std::vector<double> ys;
    for (int i = 0; i < n; ++i) {
      Vector x(p);
      x[0] = 1.0;
      for (int j = 1; j < p; ++j) x[j] = rnorm(0, poisson ? 0.3 : 1.0);
      const double eta = x.dot(beta);
      xs.push_back(x);
      ys.push_back(poisson ? rpois(exp(eta)) : (runif() < plogis(eta) ? 1.0 : 0.0));
    }
    NEW(MvnModel, slab)(Vector(p, 0.0), SpdMatrix(p, 1.0));
    NEW(VariableSelectionPrior, spike)(p, std::min(1.0, (nonzero + 1.0) / p));
    if (kind == "bench") {
      // ms per Gibbs iteration of the adapter on BOOM's own model (n heap objects) beside the standalone host classes on
      // the same rows: what the BOOM-typed surface costs on top of the device step (bench.py: e2e_adapter).
      NEW(BinomialLogitModel, model)(p);
      for (int i = 0; i < n; ++i) model->add_data(new BinomialRegressionData(ys[i], 1.0, xs[i]));
      model->coef().drop_all(); model->coef().add(0);
      RNG seeder(5);
      Ptr<B200::BinomialLogitSpikeSlabSampler> s(new B200::BinomialLogitSpikeSlabSampler(model.get(), slab, spike, 10, seeder));
      model->set_method(s);
      auto tp = std::chrono::steady_clock::now();
      model->sample_posterior();   // packs + uploads the rows, first iteration
      const double first = std::chrono::duration<double>(std::chrono::steady_clock::now() - tp).count();
      for (int i = 0; i < burn; ++i) model->sample_posterior();
      const double d0 = s->seconds_in_device_step(), h0 = s->seconds_in_host_steps();
      auto t0 = std::chrono::steady_clock::now();
      for (int i = 0; i < iters; ++i) model->sample_posterior();
      const double adapter_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      const double dev_s = s->seconds_in_device_step() - d0, host_s = s->seconds_in_host_steps() - h0;
      const int nvars = (int)model->coef().inc().nvars();
      // the standalone classes on the same rows
      std::vector<double> X((size_t)n * p), y(n), nt(n, 1.0);
      for (int i = 0; i < n; ++i) { std::copy(xs[i].begin(), xs[i].end(), X.begin() + (size_t)i * p); y[i] = ys[i]; }
      BOOM_B200::BinomialLogitModel hm(n, p, X.data(), y.data(), nt.data());
      hm.coef().drop_all(); hm.coef().add(0);
      BOOM_B200::RNG hseed(5);
      auto hslab = std::make_shared<BOOM_B200::MvnModel>(BOOM_B200::Vector(p, 0.0), BOOM_B200::SpdMatrix(p, 1.0));
      auto hspike = std::make_shared<BOOM_B200::VariableSelectionPrior>(p, std::min(1.0, (nonzero + 1.0) / p));
      auto hs = std::make_shared<BOOM_B200::BinomialLogitSpikeSlabSampler>(&hm, hslab, hspike, 10, hseed);
      hm.set_method(hs);
      for (int i = 0; i < burn + 1; ++i) hm.sample_posterior();
      auto t1 = std::chrono::steady_clock::now();
      for (int i = 0; i < iters; ++i) hm.sample_posterior();
      const double standalone_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
      printf("{\"kind\": \"bench\", \"n\": %d, \"p\": %d, \"iters\": %d, \"adapter_ms_per_iter\": %.4f, \"adapter_device_ms\": %.4f, "
             "\"adapter_host_ms\": %.4f, \"standalone_ms_per_iter\": %.4f, \"adapter_first_iteration_with_upload_s\": %.3f, "
             "\"model_size_at_end\": %d}\n",
             n, p, iters, 1e3 * adapter_s / iters, 1e3 * dev_s / iters, 1e3 * host_s / iters, 1e3 * standalone_s / iters, first, nvars);
      return 0;
    }
    if (kind == "composite" || kind == "chunk") {
      // BinomialLogitCompositeSpikeSlabSampler (what R's logit.spike builds): reference vs B200 on the same data.
      // "chunk": the chunk log posterior with gradient and Hessian, value by value; "composite": the chains.
      const double tdf = 3.0;
      const int max_tim = 4, max_rwm = 2;
      if (kind == "chunk") {
        NEW(BinomialLogitModel, model)(p);
        for (int i = 0; i < n; ++i) model->add_data(new BinomialRegressionData(ys[i], 1.0 + (i % 3), xs[i]));
        model->coef().drop_all();
        for (int j = 0; j < p; j += 2) model->coef().add(j);
        Vector b = model->included_coefficients();
        for (size_t j = 0; j < b.size(); ++j) b[j] = 0.1 * (j % 3) - 0.1;
        model->set_included_coefficients(b);
        NEW(BinomialLogitCompositeSpikeSlabSampler, r)(model.get(), slab, spike, 10, tdf, max_tim, max_rwm);
        Ptr<B200::BinomialLogitCompositeSpikeSlabSampler> g(
            new B200::BinomialLogitCompositeSpikeSlabSampler(model.get(), slab, spike, 10, tdf, max_tim, max_rwm));
        printf("{\"kind\": \"chunk\", \"cases\": [");
        const int nvars = (int)model->coef().nvars();
        bool first = true;
        for (int max_chunk : {2, 3, 0}) {
          const int nchunks = max_chunk <= 0 ? 1 : (nvars + max_chunk - 1) / max_chunk;
          for (int chunk = 0; chunk < nchunks; ++chunk) {
            BinomialLogitLogPostChunk fr = r->log_posterior(chunk, max_chunk);
            B200::BinomialLogitLogPostChunk fg = g->log_posterior(chunk, max_chunk);
            // chunk size as the samplers compute it
            Vector gr, gg; Matrix hr, hg;
            int cs = 0;
            { Vector probe = model->included_coefficients(); int per = max_chunk <= 0 ? nvars : (nvars + nchunks - 1) / nchunks;
              cs = std::min(per, nvars - per * chunk); }
            Vector bc(cs);
            for (int j = 0; j < cs; ++j) bc[j] = 0.05 * j - 0.02 * chunk;
            const double vr = fr(bc, gr, hr, 2), vg = fg(bc, gg, hg, 2);
            printf("%s{\"value_ref\": %.15g, \"value_b200\": %.15g, \"grad_diff\": %.3g, \"hess_diff\": %.3g, \"hess_scale\": %.3g}",
                   first ? "" : ", ", vr, vg, (gr - gg).max_abs(), (hr - hg).max_abs(), hr.max_abs());
            first = false;
          }
        }
        printf("]}\n");
        return 0;
      }
      Summary out[2];
      std::string report;
      for (int arm = 0; arm < 2; ++arm) {
        NEW(BinomialLogitModel, model)(p);
        for (int i = 0; i < n; ++i) model->add_data(new BinomialRegressionData(ys[i], 1.0, xs[i]));
        model->coef().drop_all(); model->coef().add(0);
        RNG seeder(arm == 0 ? 41 : 42);
        if (arm == 0) {
          NEW(BinomialLogitCompositeSpikeSlabSampler, s)(model.get(), slab, spike, 10, tdf, max_tim, max_rwm, 1.0, seeder);
          model->set_method(s);
          out[arm] = run(model, iters, burn);
        } else {
          Ptr<B200::BinomialLogitCompositeSpikeSlabSampler> s(
              new B200::BinomialLogitCompositeSpikeSlabSampler(model.get(), slab, spike, 10, tdf, max_tim, max_rwm, 1.0, seeder));
          model->set_method(s);
          out[arm] = run(model, iters, burn);
          std::ostringstream os;
          s->time_report(os);
          report = os.str();
        }
      }
      printf("{\"kind\": \"composite\", \"n\": %d, \"p\": %d, \"iters\": %d, \"burn\": %d, ", n, p, iters, burn);
      print_vec("beta_true", beta);
      print_summary("reference", out[0]); printf(", ");
      print_summary("b200", out[1]);
      printf(", \"time_report_lines\": %d}\n", (int)std::count(report.begin(), report.end(), '\n'));
      return 0;
    }
    if (kind == "active" || kind == "pactive") {
      // the active-set option on the adapter (logit / Poisson spike-and-slab): same seed, full statistics vs active-set
      // statistics -> the same chain
      std::vector<Vector> chain[2];
      long long fetched = 0;
      double xtx_diff = 0;
      SpdMatrix full_last;
      for (int arm = 0; arm < 2; ++arm) {
        RNG seeder(61);
        if (poisson) {
          NEW(PoissonRegressionModel, model)(p);
          for (int i = 0; i < n; ++i) model->add_data(new PoissonRegressionData((int64_t)ys[i], xs[i], 1.0));
          model->coef().drop_all(); model->coef().add(0);
          Ptr<B200::PoissonRegressionSpikeSlabSampler> s(new B200::PoissonRegressionSpikeSlabSampler(model.get(), slab, spike, 1, seeder));
          s->set_active_set_statistics(arm == 1);
          model->set_method(s);
          for (int it = 0; it < iters; ++it) { model->sample_posterior(); chain[arm].push_back(model->Beta()); }
          if (arm == 0) full_last = s->complete_data_sufficient_statistics().xtx();
          else {
            fetched = s->active_set_columns_fetched();
            xtx_diff = (s->complete_data_sufficient_statistics().xtx() - full_last).max_abs() / full_last.max_abs();
          }
        } else {
          NEW(BinomialLogitModel, model)(p);
          for (int i = 0; i < n; ++i) model->add_data(new BinomialRegressionData(ys[i], 1.0, xs[i]));
          model->coef().drop_all(); model->coef().add(0);
          Ptr<B200::BinomialLogitSpikeSlabSampler> s(new B200::BinomialLogitSpikeSlabSampler(model.get(), slab, spike, 10, seeder));
          s->set_active_set_statistics(arm == 1);
          model->set_method(s);
          for (int it = 0; it < iters; ++it) { model->sample_posterior(); chain[arm].push_back(model->Beta()); }
          if (arm == 0) full_last = s->suf().xtx();
          else { fetched = s->active_set_columns_fetched(); xtx_diff = (s->suf().xtx() - full_last).max_abs() / full_last.max_abs(); }
        }
      }
      double dmax = 0; bool same_model = true;
      for (int it = 0; it < iters; ++it)
        for (int j = 0; j < p; ++j) {
          dmax = std::max(dmax, std::fabs(chain[0][it][j] - chain[1][it][j]));
          same_model = same_model && ((chain[0][it][j] != 0) == (chain[1][it][j] != 0));
        }
      printf("{\"kind\": \"%s\", \"n\": %d, \"p\": %d, \"iters\": %d, \"chain_max_abs_diff\": %.3g, \"same_model\": %s, "
             "\"columns_fetched\": %lld, \"suf_xtx_rel_diff\": %.3g}\n", kind.c_str(), n, p, iters, dmax, same_model ? "true" : "false", fetched,
             xtx_diff);
      return 0;
    }
    if (kind == "probit") {
      // the sibling sampler: BinomialProbitSpikeSlabSampler on BOOM's BinomialProbitModel, reference vs B200 (binomial rows,
      // n_i in {1, 2, 3, 15}: both the per-trial draws and the CLT branch)
      Summary out[2];
      std::vector<double> yp(n), np_(n);
      for (int i = 0; i < n; ++i) {
        np_[i] = (i % 7 == 0) ? 15.0 : 1.0 + (i % 3);
        yp[i] = rbinom((int)np_[i], pnorm(xs[i].dot(beta)));
      }
      for (int arm = 0; arm < 2; ++arm) {
        NEW(BinomialProbitModel, model)(p);
        for (int i = 0; i < n; ++i) model->add_data(new BinomialRegressionData(yp[i], np_[i], xs[i]));
        model->coef().drop_all(); model->coef().add(0);
        RNG seeder(arm == 0 ? 51 : 52);
        Ptr<PosteriorSampler> sampler;
        if (arm == 0) sampler = new BinomialProbitSpikeSlabSampler(model.get(), slab, spike, 10, seeder);
        else sampler = new B200::BinomialProbitSpikeSlabSampler(model.get(), slab, spike, 10, seeder);
        model->set_method(sampler);
        out[arm] = run(model, iters, burn);
      }
      printf("{\"kind\": \"probit\", \"n\": %d, \"p\": %d, \"iters\": %d, \"burn\": %d, ", n, p, iters, burn);
      print_vec("beta_true", beta);
      print_summary("reference", out[0]); printf(", ");
      print_summary("b200", out[1]);
      printf("}\n");
      return 0;
    }
    if (kind == "tactive") {
      // the active-set option on B200::TRegressionSpikeSlabSampler: same seed, full statistics vs active-set statistics
      std::vector<Vector> chain[2];
      long long fetched = 0;
      double xtx_diff = 0;
      SpdMatrix full_last;
      std::vector<double> yt(n);
      for (int i = 0; i < n; ++i) yt[i] = xs[i].dot(beta) + 1.5 * rnorm() / std::sqrt(rgamma(2.0, 2.0));
      NEW(MvnModel, tslab)(Vector(p, 0.0), SpdMatrix(p, 4.0));
      NEW(ChisqModel, siginv_prior)(1.0, 1.0);
      NEW(UniformModel, nu_prior)(0.5, 60.0);
      for (int arm = 0; arm < 2; ++arm) {
        RNG seeder(81);
        NEW(TRegressionModel, model)(p);
        for (int i = 0; i < n; ++i) model->add_data(new RegressionData(yt[i], xs[i]));
        model->coef().drop_all(); model->coef().add(0);
        Ptr<B200::TRegressionSpikeSlabSampler> s(new B200::TRegressionSpikeSlabSampler(model.get(), tslab, spike, siginv_prior, nu_prior, seeder));
        s->set_active_set_statistics(arm == 1);
        model->set_method(s);
        for (int it = 0; it < iters; ++it) {
          model->sample_posterior();
          chain[arm].push_back(concat(model->Beta(), Vector{model->sigsq(), model->nu()}));
        }
        if (arm == 0) full_last = s->complete_data_sufficient_statistics().xtx();
        else {
          fetched = s->active_set_columns_fetched();
          xtx_diff = (s->complete_data_sufficient_statistics().xtx() - full_last).max_abs() / full_last.max_abs();
        }
      }
      double dmax = 0; bool same_model = true;
      for (int it = 0; it < iters; ++it)
        for (int j = 0; j < p + 2; ++j) {
          dmax = std::max(dmax, std::fabs(chain[0][it][j] - chain[1][it][j]));
          if (j < p) same_model = same_model && ((chain[0][it][j] != 0) == (chain[1][it][j] != 0));
        }
      printf("{\"kind\": \"tactive\", \"n\": %d, \"p\": %d, \"iters\": %d, \"chain_max_abs_diff\": %.3g, \"same_model\": %s, "
             "\"columns_fetched\": %lld, \"suf_xtx_rel_diff\": %.3g}\n", n, p, iters, dmax, same_model ? "true" : "false", fetched, xtx_diff);
      return 0;
    }
    if (kind == "tspike") {
      // TRegressionSpikeSlabSampler (lm.spike with Student errors) on BOOM's TRegressionModel: reference vs B200
      Summary out[2];
      std::vector<double> yt(n);
      for (int i = 0; i < n; ++i) yt[i] = xs[i].dot(beta) + 1.5 * rnorm() / std::sqrt(rgamma(2.0, 2.0));
      NEW(MvnModel, tslab)(Vector(p, 0.0), SpdMatrix(p, 4.0));
      NEW(ChisqModel, siginv_prior)(1.0, 1.0);
      NEW(UniformModel, nu_prior)(0.5, 60.0);
      Vector sn_mean[2], sn_sd[2];
      for (int arm = 0; arm < 2; ++arm) {
        NEW(TRegressionModel, model)(p);
        for (int i = 0; i < n; ++i) model->add_data(new RegressionData(yt[i], xs[i]));
        model->coef().drop_all(); model->coef().add(0);
        RNG seeder(arm == 0 ? 71 : 72);
        Ptr<PosteriorSampler> sampler;
        if (arm == 0) sampler = new TRegressionSpikeSlabSampler(model.get(), tslab, spike, siginv_prior, nu_prior, seeder);
        else sampler = new B200::TRegressionSpikeSlabSampler(model.get(), tslab, spike, siginv_prior, nu_prior, seeder);
        model->set_method(sampler);
        Vector s1(p, 0.0), s2(p, 0.0), inc(p, 0.0), t1(2, 0.0), t2(2, 0.0);
        auto t0 = std::chrono::steady_clock::now();
        for (int it = 0; it < iters; ++it) {
          model->sample_posterior();
          if (it >= burn) {
            const Vector &b(model->Beta());
            for (int j = 0; j < p; ++j) { s1[j] += b[j]; s2[j] += b[j] * b[j]; inc[j] += model->coef().inc()[j]; }
            const double v[2] = {model->sigma(), model->nu()};
            for (int j = 0; j < 2; ++j) { t1[j] += v[j]; t2[j] += v[j] * v[j]; }
          }
        }
        out[arm].secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const double m = iters - burn;
        out[arm].mean = s1 / m; out[arm].sd = Vector(p); out[arm].inc = inc / m;
        for (int j = 0; j < p; ++j) out[arm].sd[j] = std::sqrt(std::max(0.0, s2[j] / m - out[arm].mean[j] * out[arm].mean[j]));
        sn_mean[arm] = t1 / m; sn_sd[arm] = Vector(2);
        for (int j = 0; j < 2; ++j) sn_sd[arm][j] = std::sqrt(std::max(0.0, t2[j] / m - sn_mean[arm][j] * sn_mean[arm][j]));
      }
      printf("{\"kind\": \"tspike\", \"n\": %d, \"p\": %d, \"iters\": %d, \"burn\": %d, ", n, p, iters, burn);
      print_vec("beta_true", beta);
      print_vec("reference_sigma_nu_mean", sn_mean[0]); print_vec("reference_sigma_nu_sd", sn_sd[0]);
      print_vec("b200_sigma_nu_mean", sn_mean[1]); print_vec("b200_sigma_nu_sd", sn_sd[1]);
      print_summary("reference", out[0]); printf(", ");
      print_summary("b200", out[1]);
      printf("}\n");
      return 0;
    }
    if (kind == "treg") {
      // the Student-t sibling: TRegressionSampler on BOOM's TRegressionModel, reference vs B200, y = x'beta + 1.5 t_4.
      // The summaries carry (beta, sigma, nu); the B200 arm also checks its public pieces against the reference's own.
      Summary out[2];
      std::vector<double> yt(n);
      for (int i = 0; i < n; ++i) yt[i] = xs[i].dot(beta) + 1.5 * rnorm() / std::sqrt(rgamma(2.0, 2.0));
      NEW(MvnModel, prior)(Vector(p, 0.0), SpdMatrix(p, 100.0));
      NEW(ChisqModel, siginv_prior)(1.0, 1.0);
      NEW(UniformModel, nu_prior)(0.5, 60.0);
      double ll_ref = 0, ll_b200 = 0, suf_diff = -1;
      for (int arm = 0; arm < 2; ++arm) {
        NEW(TRegressionModel, model)(p);
        for (int i = 0; i < n; ++i) model->add_data(new RegressionData(yt[i], xs[i]));
        RNG seeder(arm == 0 ? 61 : 62);
        Ptr<PosteriorSampler> sampler;
        Ptr<B200::TRegressionSampler> b200;
        if (arm == 0) sampler = new TRegressionSampler(model.get(), prior, siginv_prior, nu_prior, seeder);
        else { b200 = new B200::TRegressionSampler(model.get(), prior, siginv_prior, nu_prior, seeder); sampler = b200; }
        model->set_method(sampler);
        Vector s1(p + 2, 0.0), s2(p + 2, 0.0);
        auto t0 = std::chrono::steady_clock::now();
        for (int it = 0; it < iters; ++it) {
          model->sample_posterior();
          if (it >= burn) {
            Vector v = concat(model->Beta(), Vector{model->sigma(), model->nu()});
            for (int j = 0; j < p + 2; ++j) { s1[j] += v[j]; s2[j] += v[j] * v[j]; }
          }
        }
        out[arm].secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const double m = iters - burn;
        out[arm].mean = s1 / m; out[arm].sd = Vector(p + 2); out[arm].inc = Vector(p + 2, 1.0);
        for (int j = 0; j < p + 2; ++j) out[arm].sd[j] = std::sqrt(std::max(0.0, s2[j] / m - out[arm].mean[j] * out[arm].mean[j]));
        if (arm == 1) {
          // device log likelihood against the reference model's own; the reference-typed statistics against a host recomputation
          // is not possible (the weights stay on the device), so check their internal consistency instead
          ll_ref = model->log_likelihood(model->Beta(), model->sigsq(), model->nu());
          ll_b200 = b200->log_likelihood(model->Beta(), model->sigsq(), model->nu());
          const WeightedRegSuf &suf(b200->complete_data_sufficient_statistics());
          suf_diff = std::fabs(suf.n() - n) + (suf.sumw() > 0 ? 0.0 : 1.0) + (suf.xtx()(0, 0) == suf.sumw() ? 0.0 : std::fabs(suf.xtx()(0, 0) - suf.sumw()) / suf.sumw());
          b200->draw_nu_given_complete_data();
          if (!(model->nu() > 0.5 && model->nu() < 60.0)) suf_diff += 10;
        }
      }
      printf("{\"kind\": \"treg\", \"n\": %d, \"p\": %d, \"iters\": %d, \"burn\": %d, \"loglike_reference\": %.15g, \"loglike_b200\": %.15g, "
             "\"suf_consistency\": %.3g, ", n, p, iters, burn, ll_ref, ll_b200, suf_diff);
      print_vec("beta_true", beta);
      print_summary("reference", out[0]); printf(", ");
      print_summary("b200", out[1]);
      printf("}\n");
      return 0;
    }
    if (kind == "api") {
      // the public surface beyond draw(): draw_model_indicators / draw_beta / log_model_prob, suf() in the reference's own
      // type, in-place row edits.  Compared with the reference sampler fed the SAME externally driven statistics.
      NEW(BinomialLogitModel, model)(p);
      for (int i = 0; i < n; ++i) model->add_data(new BinomialRegressionData(ys[i], 1.0, xs[i]));
      model->coef().drop_all(); model->coef().add(0);
      RNG seeder(9);
      Ptr<B200::BinomialLogitSpikeSlabSampler> s(new B200::BinomialLogitSpikeSlabSampler(model.get(), slab, spike, 10, seeder));
      s->impute_latent_data();
      const BinomialLogit::SufficientStatistics &suf(s->suf());          // the reference's type
      // feed the reference sampler the same statistics through its own update hook: x = e_j pieces are not enough for a
      // full matrix, so compare log_model_prob on statistics BOTH samplers build from the same external updates instead
      NEW(BinomialLogitModel, m2)(p);
      m2->coef().drop_all(); m2->coef().add(0);
      NEW(BinomialLogitSpikeSlabSampler, r)(m2.get(), slab, spike, 10, seeder);
      Ptr<B200::BinomialLogitSpikeSlabSampler> s2(new B200::BinomialLogitSpikeSlabSampler(m2.get(), slab, spike, 10, seeder));
      r->fix_latent_data(true); s2->fix_latent_data(true);
      r->clear_complete_data_sufficient_statistics(); s2->clear_complete_data_sufficient_statistics();
      RNG data_rng(3);
      for (int i = 0; i < n; ++i) {
        const double w = 0.1 + runif_mt(data_rng), z = xs[i].dot(beta) + rnorm_mt(data_rng) / sqrt(w);
        r->update_complete_data_sufficient_statistics(w * z, w, xs[i]);
        s2->update_complete_data_sufficient_statistics(w * z, w, xs[i]);
      }
      Vector lref, lgpu;
      for (int trial = 0; trial < 6; ++trial) {
        Selector g(p, false);
        g.add(0);
        for (int j = 1; j < p; ++j) if ((j + trial) % 3 == 0 || j <= trial) g.add(j);
        lref.push_back(r->log_model_prob(g)); lgpu.push_back(s2->log_model_prob(g));
      }
      const double xtx_diff = (r->suf().xtx() - s2->suf().xtx()).max_abs(), xty_diff = (r->suf().xty() - s2->suf().xty()).max_abs();
      s2->draw_model_indicators();
      s2->draw_beta();
      const int nv = (int)m2->coef().inc().nvars();
      // in-place edit of one row is seen after observe_rows(true)
      s->observe_rows(true);
      const double before = s->suf().xtx()(1, 1);
      Vector x0 = model->dat()[0]->x();
      x0[1] += 100.0;
      model->dat()[0]->set_x(x0);
      s->impute_latent_data();
      const double after = s->suf().xtx()(1, 1);
      printf("{\"kind\": \"api\", \"suf_sample_size\": %d, \"suf_xtx00\": %.10g, ", suf.sample_size(), before);
      print_vec("log_model_prob_reference", lref); print_vec("log_model_prob_b200", lgpu);
      printf("\"external_xtx_max_abs_diff\": %.3g, \"external_xty_max_abs_diff\": %.3g, \"nvars_after_sweep\": %d, "
             "\"xtx11_before_edit\": %.10g, \"xtx11_after_edit\": %.10g, \"reference_sample_size\": %d, \"b200_sample_size\": %d}\n",
             xtx_diff, xty_diff, nv, before, after, r->suf().sample_size(), s2->suf().sample_size());
      return 0;
    }
    if (kind == "mode" || kind == "pmode") {
      // find_posterior_mode of both spike-and-slab samplers from the same start, on the model with the first
      // nonzero + 1 coefficients included
      Vector bref, bgpu;
      double vref = 0, vgpu = 0;
      for (int arm = 0; arm < 2; ++arm) {
        if (kind == "mode") {
          NEW(BinomialLogitModel, model)(p);
          for (int i = 0; i < n; ++i) model->add_data(new BinomialRegressionData(ys[i], 1.0, xs[i]));
          model->coef().drop_all();
          for (int j = 0; j <= nonzero && j < p; ++j) model->coef().add(j);
          if (arm == 0) {
            NEW(BinomialLogitSpikeSlabSampler, s)(model.get(), slab, spike, 10);
            s->find_posterior_mode(1e-9); bref = model->Beta(); vref = s->log_posterior_at_mode();
          } else {
            Ptr<B200::BinomialLogitSpikeSlabSampler> s(new B200::BinomialLogitSpikeSlabSampler(model.get(), slab, spike, 10));
            s->find_posterior_mode(1e-9); bgpu = model->Beta(); vgpu = s->log_posterior_at_mode();
          }
        } else {
          NEW(PoissonRegressionModel, model)(p);
          for (int i = 0; i < n; ++i) model->add_data(new PoissonRegressionData((int64_t)ys[i], xs[i], 1.0));
          model->coef().drop_all();
          for (int j = 0; j <= nonzero && j < p; ++j) model->coef().add(j);
          if (arm == 0) {
            NEW(PoissonRegressionSpikeSlabSampler, s)(model.get(), slab, spike, 1);
            s->find_posterior_mode(1e-9); bref = model->Beta(); vref = s->log_posterior_at_mode();
          } else {
            Ptr<B200::PoissonRegressionSpikeSlabSampler> s(new B200::PoissonRegressionSpikeSlabSampler(model.get(), slab, spike, 1));
            s->find_posterior_mode(1e-9); bgpu = model->Beta(); vgpu = s->log_posterior_at_mode();
          }
        }
      }
      printf("{\"kind\": \"%s\", ", kind.c_str());
      print_vec("reference_mode", bref); print_vec("b200_mode", bgpu);
      printf("\"reference_log_posterior\": %.12g, \"b200_log_posterior\": %.12g}\n", vref, vgpu);
      return 0;
    }
    if (kind == "fixed" || kind == "pfixed") {
      // The state-space callers' mode (fix_latent_data(true), statistics pushed from outside): both samplers run their
      // host steps only -- no device needed -- on the SAME externally driven complete-data statistics.
      Summary out[2];
      for (int arm = 0; arm < 2; ++arm) {
        RNG seeder(arm == 0 ? 31 : 32), data_rng(77);
        auto push = [&](auto &sampler) {
          sampler->fix_latent_data(true);
          sampler->clear_complete_data_sufficient_statistics();
          for (int i = 0; i < n; ++i) {
            const double w = 0.1 + runif_mt(data_rng), z = xs[i].dot(beta) + rnorm_mt(data_rng) / sqrt(w);
            sampler->update_complete_data_sufficient_statistics(w * z, w, xs[i]);
          }
        };
        if (kind == "fixed") {
          NEW(BinomialLogitModel, model)(p);
          model->coef().drop_all(); model->coef().add(0);
          if (arm == 0) {
            NEW(BinomialLogitSpikeSlabSampler, s)(model.get(), slab, spike, 10, seeder);
            push(s); model->set_method(s); out[arm] = run(model, iters, burn);
          } else {
            Ptr<B200::BinomialLogitSpikeSlabSampler> s(new B200::BinomialLogitSpikeSlabSampler(model.get(), slab, spike, 10, seeder));
            push(s); model->set_method(s); out[arm] = run(model, iters, burn);
          }
        } else {
          NEW(PoissonRegressionModel, model)(p);
          model->coef().drop_all(); model->coef().add(0);
          if (arm == 0) {
            NEW(PoissonRegressionSpikeSlabSampler, s)(model.get(), slab, spike, 1, seeder);
            push(s); model->set_method(s); out[arm] = run(model, iters, burn);
          } else {
            Ptr<B200::PoissonRegressionSpikeSlabSampler> s(new B200::PoissonRegressionSpikeSlabSampler(model.get(), slab, spike, 1, seeder));
            push(s); model->set_method(s); out[arm] = run(model, iters, burn);
          }
        }
      }
      printf("{\"kind\": \"%s\", \"n\": %d, \"p\": %d, \"iters\": %d, \"burn\": %d, ", kind.c_str(), n, p, iters, burn);
      print_vec("beta_true", beta);
      print_summary("reference", out[0]); printf(", ");
      print_summary("b200", out[1]);
      printf("}\n");
      return 0;
    }
    Summary ref, gpu;
    for (int arm = 0; arm < 2; ++arm) {
      if (!poisson) {
        NEW(BinomialLogitModel, model)(p);
        for (int i = 0; i < n; ++i) model->add_data(new BinomialRegressionData(ys[i], 1.0, xs[i]));
        if (kind == "spike") { model->coef().drop_all(); model->coef().add(0); }
        Ptr<PosteriorSampler> sampler;
        RNG seeder(arm == 0 ? 11 : 12);
        if (kind == "spike") {
          if (arm == 0) sampler = new BinomialLogitSpikeSlabSampler(model.get(), slab, spike, 10, seeder);
          else sampler = new B200::BinomialLogitSpikeSlabSampler(model.get(), slab, spike, 10, seeder);
        } else {
          if (arm == 0) sampler = new BinomialLogitAuxmixSampler(model.get(), slab, 10, seeder);
          else sampler = new B200::BinomialLogitAuxmixSampler(model.get(), slab, 10, seeder);
        }
        model->set_method(sampler);
        (arm == 0 ? ref : gpu) = run(model, iters, burn);
      } else {
        NEW(PoissonRegressionModel, model)(p);
        for (int i = 0; i < n; ++i) model->add_data(new PoissonRegressionData((int64_t)ys[i], xs[i], 1.0));
        if (kind == "pspike") { model->coef().drop_all(); model->coef().add(0); }
        Ptr<PosteriorSampler> sampler;
        RNG seeder(arm == 0 ? 21 : 22);
        if (kind == "pspike") {
          if (arm == 0) sampler = new PoissonRegressionSpikeSlabSampler(model.get(), slab, spike, 1, seeder);
          else sampler = new B200::PoissonRegressionSpikeSlabSampler(model.get(), slab, spike, 1, seeder);
        } else {
          if (arm == 0) sampler = new PoissonRegressionAuxMixSampler(model.get(), slab, 1, seeder);
          else sampler = new B200::PoissonRegressionAuxMixSampler(model.get(), slab, 1, seeder);
        }
        model->set_method(sampler);
        (arm == 0 ? ref : gpu) = run(model, iters, burn);
      }
    }
    printf("{\"kind\": \"%s\", \"n\": %d, \"p\": %d, \"iters\": %d, \"burn\": %d, ", kind.c_str(), n, p, iters, burn);
    print_vec("beta_true", beta);
    print_summary("reference", ref); printf(", ");
    print_summary("b200", gpu);
    printf("}\n");
  } catch (std::exception &e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
