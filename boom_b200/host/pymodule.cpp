// pymodule.cpp -- pybind11 view of boom_b200.hpp, shaped like the reference's own bindings
// (Interfaces/python/BayesBoom/Models/Glm/GlmModel_def.cpp:935-1055): models, priors, samplers,
// model.set_method(sampler), model.sample_posterior().
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <tuple>

#include "boom_b200.hpp"

namespace py = pybind11;
using namespace BOOM_B200;

namespace {
typedef py::array_t<double, py::array::c_style | py::array::forcecast> NpD;
typedef py::array_t<int64_t, py::array::c_style | py::array::forcecast> NpI;

SpdMatrix to_spd(const NpD &a) {
  if (a.ndim() != 2 || a.shape(0) != a.shape(1)) report_error("expected a square matrix");
  SpdMatrix m((int)a.shape(0));
  std::copy(a.data(), a.data() + a.size(), m.a.begin());
  return m;
}
NpD from_spd(const SpdMatrix &m) {
  NpD out({m.dim, m.dim});
  std::copy(m.a.begin(), m.a.end(), out.mutable_data());
  return out;
}
Vector to_vec(const NpD &a) { return Vector(a.data(), a.data() + a.size()); }
NpD from_vec(const Vector &v) {
  NpD out((py::ssize_t)v.size());
  std::copy(v.begin(), v.end(), out.mutable_data());
  return out;
}

template <class M>
void bind_model_core(py::class_<M, std::shared_ptr<M>> &c) {
  c.def_property_readonly("xdim", &M::xdim)
      .def_property_readonly("sample_size", [](const M &m) { return m.nobs(); })
      .def_property("Beta", [](const M &m) { return from_vec(m.Beta()); }, [](M &m, const NpD &b) { m.set_Beta(to_vec(b)); })
      .def("set_Beta", [](M &m, const NpD &b) { m.set_Beta(to_vec(b)); })
      .def_property_readonly("inc", [](const M &m) {
        py::array_t<bool> out((py::ssize_t)m.xdim());
        for (int i = 0; i < m.xdim(); ++i) out.mutable_data()[i] = m.coef().inc()[i];
        return out;
      })
      .def("set_inc", [](M &m, const std::vector<bool> &bits) {
        Selector g((int)bits.size(), false);
        for (size_t i = 0; i < bits.size(); ++i) if (bits[i]) g.add((int)i);
        m.coef().set_inc(g);
      })
      .def("drop_all", &M::drop_all)
      .def("add", [](M &m, int i) { m.coef().add(i); })
      .def("set_method", [](M &m, const std::shared_ptr<PosteriorSampler> &s) { m.set_method(s); })
      .def("clear_methods", &M::clear_methods)
      .def("sample_posterior", [](M &m) { py::gil_scoped_release rel; m.sample_posterior(); })
      .def("logpri", &M::logpri)
      .def("set_device", &M::set_device)
      .def("set_row_offset", &M::set_row_offset)
      .def("set_stream", [](M &m, uintptr_t s) { m.set_stream(reinterpret_cast<void *>(s)); })
      .def("set_allreduce", [](M &m, py::object fn) {
        if (fn.is_none()) { m.set_allreduce(AllReduceFn()); return; }
        m.set_allreduce([fn](double *dev, int64_t count) {
          py::gil_scoped_acquire acq;
          fn(reinterpret_cast<uintptr_t>(dev), count);
        });
      })
      .def_static("comm_unique_id", []() { return py::bytes(M::comm_unique_id()); })
      .def("set_communicator", [](M &m, const py::bytes &id, int nranks, int rank) { m.set_communicator(std::string(id), nranks, rank); })
      .def("set_device_option", &M::set_device_option)
      .def("kernel_launches", &M::kernel_launches)
      .def("kernel_timings", [](M &m, bool reset) {
        double ms[5]; int64_t cnt[5];
        m.kernel_timings(ms, cnt, reset);
        static const char *names[5] = {"fused_small", "impute_rows", "syrk_dmma", "reduce", "other"};
        py::dict out;
        for (int c = 0; c < 5; ++c) out[names[c]] = py::make_tuple(ms[c], cnt[c]);
        return out;
      }, py::arg("reset") = false);
}

template <class M>
void bind_model_common(py::class_<M, std::shared_ptr<M>> &c) {
  bind_model_core(c);
  c.def("log_likelihood_derivs", [](M &m, const NpD &b) {
        Vector g; SpdMatrix h;
        double ll = m.log_likelihood_derivs(to_vec(b), &g, &h);
        return py::make_tuple(ll, from_vec(g), from_spd(h));
      })
      .def("log_likelihood", [](M &m) { return m.log_likelihood(); })
      .def("log_likelihood", [](M &m, const NpD &b) { return m.log_likelihood(to_vec(b)); });
}
}  // namespace

PYBIND11_MODULE(_host, m) {
  m.doc() = "B200-native auxiliary-mixture Gibbs samplers behind BOOM's sampler surface";

  py::class_<RNG>(m, "RNG").def(py::init<>()).def(py::init<RNG::RngIntType>()).def("seed", [](RNG &r, RNG::RngIntType s) { r.seed(s); })
      .def("__call__", [](RNG &r) { return r(); });
  m.def("global_rng", []() -> RNG & { return GlobalRng::rng; }, py::return_value_policy::reference);
  m.def("set_logit_mixture", [](const NpD &mu, const NpD &sigma, const NpD &w) { set_logit_mixture(to_vec(mu), to_vec(sigma), to_vec(w)); });
  m.def("poisson_mixture_table_is_set", []() { return PoissonRegressionAuxMixSampler::mixture_table_is_set(); });
  m.def("set_poisson_mixture_table", [](const NpD &ser, int64_t largest) {
    PoissonRegressionAuxMixSampler::set_mixture_table(to_vec(ser), largest);
  });

  m.def("poisson_mixture_table", []() { return from_vec(PoissonRegressionAuxMixSampler::mixture_table()); });
  // (mu, sigma, weights, kl) of the table entry for nu, interpolated / fitted and added when nu is off the grid
  m.def("poisson_mixture_approximate", [](int64_t nu) {
    NormalMixtureApproximation a = PoissonRegressionAuxMixSampler::approximate(nu);
    return py::make_tuple(from_vec(a.mu), from_vec(a.sigma), from_vec(a.weights), kullback_leibler_neg_log_gamma((double)nu, a));
  });

  py::class_<MvnBase, std::shared_ptr<MvnBase>>(m, "MvnBase");
  py::class_<MvnModel, MvnBase, std::shared_ptr<MvnModel>>(m, "MvnModel")
      .def(py::init([](const NpD &mu, const NpD &V, bool ivar) { return std::make_shared<MvnModel>(to_vec(mu), to_spd(V), ivar); }),
           py::arg("mu"), py::arg("Sigma"), py::arg("ivar") = false)
      .def_property_readonly("mu", [](const MvnModel &p) { return from_vec(p.mu()); })
      .def_property_readonly("siginv", [](const MvnModel &p) { return from_spd(p.siginv()); })
      .def("logp", [](const MvnModel &p, const NpD &x) { return p.logp(to_vec(x)); });
  py::class_<VariableSelectionPrior, std::shared_ptr<VariableSelectionPrior>>(m, "VariableSelectionPrior")
      .def(py::init<int, double>(), py::arg("n"), py::arg("inclusion_probability") = 1.0)
      .def(py::init([](const NpD &probs) { return std::make_shared<VariableSelectionPrior>(to_vec(probs)); }))
      .def("set_max_model_size", &VariableSelectionPrior::set_max_model_size);

  py::class_<BinomialLogitModel, std::shared_ptr<BinomialLogitModel>> blm(m, "BinomialLogitModel");
  blm.def(py::init<int, bool>(), py::arg("xdim"), py::arg("all") = true)
      .def(py::init([](const NpD &X, const NpD &y, const NpD &n) {
        if (X.ndim() != 2 || y.size() != X.shape(0) || n.size() != X.shape(0)) report_error("BinomialLogitModel(X, y, n): shape mismatch");
        return std::make_shared<BinomialLogitModel>((int64_t)X.shape(0), (int)X.shape(1), X.data(), y.data(), n.data());
      }))
      .def("add_data", [](BinomialLogitModel &mo, double y, double n, const NpD &x) { mo.add_data(y, n, to_vec(x)); })
      .def("borrow_host_data", [](BinomialLogitModel &mo, NpD X, NpD y, NpD n) {
        // zero copy: the model keeps references to the three arrays (they must already be C-contiguous float64:
        // forcecast would otherwise have made private copies, which is fine too) and reads them when it uploads
        if (X.ndim() != 2 || (int)X.shape(1) != mo.xdim() || y.size() != X.shape(0) || n.size() != X.shape(0))
          report_error("borrow_host_data(X, y, n): shape mismatch");
        auto keep = std::make_shared<std::tuple<NpD, NpD, NpD>>(X, y, n);
        mo.borrow_host_data((int64_t)X.shape(0), X.data(), (int64_t)X.shape(1), y.data(), n.data(),
                            std::shared_ptr<void>(keep, keep.get()));
      })
      .def("adopt_device_data", [](BinomialLogitModel &mo, int64_t n, uintptr_t dX, int64_t ldx, uintptr_t dy, uintptr_t dn) {
        mo.adopt_device_data(n, reinterpret_cast<const double *>(dX), ldx, reinterpret_cast<const double *>(dy),
                             reinterpret_cast<const double *>(dn));
      });
  blm.def("set_nonevent_sampling_prob", &BinomialLogitModel::set_nonevent_sampling_prob);
  bind_model_common(blm);

  py::class_<BinomialProbitModel, BinomialLogitModel, std::shared_ptr<BinomialProbitModel>>(m, "BinomialProbitModel")
      .def(py::init<int, bool>(), py::arg("xdim"), py::arg("all") = true)
      .def(py::init([](const NpD &X, const NpD &y, const NpD &nt) {
        if (X.ndim() != 2 || y.size() != X.shape(0) || nt.size() != X.shape(0)) report_error("BinomialProbitModel(X, y, n): shape mismatch");
        return std::make_shared<BinomialProbitModel>((int64_t)X.shape(0), (int)X.shape(1), X.data(), y.data(), nt.data());
      }));

  py::class_<PoissonRegressionModel, std::shared_ptr<PoissonRegressionModel>> prm(m, "PoissonRegressionModel");
  prm.def(py::init<int, bool>(), py::arg("xdim"), py::arg("all") = true)
      .def(py::init([](const NpD &X, const NpI &y, const NpD &e) {
        if (X.ndim() != 2 || y.size() != X.shape(0) || e.size() != X.shape(0)) report_error("PoissonRegressionModel(X, y, exposure): shape mismatch");
        return std::make_shared<PoissonRegressionModel>((int64_t)X.shape(0), (int)X.shape(1), X.data(), y.data(), e.data());
      }))
      .def("add_data", [](PoissonRegressionModel &mo, int64_t y, const NpD &x, double e) { mo.add_data(y, to_vec(x), e); },
           py::arg("y"), py::arg("x"), py::arg("exposure") = 1.0)
      .def("borrow_host_data", [](PoissonRegressionModel &mo, NpD X, NpI y, NpD e) {
        if (X.ndim() != 2 || (int)X.shape(1) != mo.xdim() || y.size() != X.shape(0) || e.size() != X.shape(0))
          report_error("borrow_host_data(X, y, exposure): shape mismatch");
        auto keep = std::make_shared<std::tuple<NpD, NpI, NpD>>(X, y, e);
        mo.borrow_host_data((int64_t)X.shape(0), X.data(), (int64_t)X.shape(1), y.data(), e.data(),
                            std::shared_ptr<void>(keep, keep.get()));
      })
      .def("adopt_device_data", [](PoissonRegressionModel &mo, int64_t n, uintptr_t dX, int64_t ldx, uintptr_t dy, uintptr_t de) {
        mo.adopt_device_data(n, reinterpret_cast<const double *>(dX), ldx, reinterpret_cast<const int64_t *>(dy),
                             reinterpret_cast<const double *>(de));
      });
  bind_model_common(prm);

  py::class_<WeightedRegSuf>(m, "WeightedRegSuf")
      .def_property_readonly("xtx", [](const WeightedRegSuf &s) { return from_spd(s.xtx()); })
      .def_property_readonly("xty", [](const WeightedRegSuf &s) { return from_vec(s.xty()); })
      .def_property_readonly("n", &WeightedRegSuf::n)
      .def_property_readonly("yty", &WeightedRegSuf::yty)
      .def_property_readonly("sumw", &WeightedRegSuf::sumw)
      .def_property_readonly("sumlogw", &WeightedRegSuf::sumlogw)
      .def_property_readonly("sample_size", &WeightedRegSuf::sample_size);

  py::class_<PosteriorSampler, std::shared_ptr<PosteriorSampler>>(m, "PosteriorSampler")
      .def("draw", [](PosteriorSampler &s) { py::gil_scoped_release rel; s.draw(); })
      .def("logpri", &PosteriorSampler::logpri)
      .def("set_seed", &PosteriorSampler::set_seed);

  py::class_<BinomialLogitAuxmixSampler, PosteriorSampler, std::shared_ptr<BinomialLogitAuxmixSampler>>(m, "BinomialLogitAuxmixSampler")
      .def(py::init([](BinomialLogitModel *model, const std::shared_ptr<MvnBase> &prior, int clt, RNG &rng) {
             return std::make_shared<BinomialLogitAuxmixSampler>(model, prior, clt, rng);
           }),
           py::arg("model"), py::arg("prior"), py::arg("clt_threshold") = 10, py::arg("seeding_rng") = std::ref(GlobalRng::rng),
           py::keep_alive<1, 2>())
      .def("impute_latent_data", [](BinomialLogitAuxmixSampler &s) { py::gil_scoped_release rel; s.impute_latent_data(); })
      .def("draw_params", &BinomialLogitAuxmixSampler::draw_params)
      .def_property_readonly("suf", &BinomialLogitAuxmixSampler::suf, py::return_value_policy::reference_internal)
      .def_property_readonly("clt_threshold", &BinomialLogitAuxmixSampler::clt_threshold)
      .def("clear_complete_data_sufficient_statistics", &BinomialLogitAuxmixSampler::clear_complete_data_sufficient_statistics)
      .def("update_complete_data_sufficient_statistics",
           [](BinomialLogitAuxmixSampler &s, double sum, double prec, const NpD &x) {
             s.update_complete_data_sufficient_statistics(sum, prec, to_vec(x));
           })
      .def("fix_latent_data", &BinomialLogitAuxmixSampler::fix_latent_data, py::arg("fixed") = true)
      .def("set_number_of_workers", &BinomialLogitAuxmixSampler::set_number_of_workers);

  py::class_<BinomialLogitSpikeSlabSampler, BinomialLogitAuxmixSampler, std::shared_ptr<BinomialLogitSpikeSlabSampler>>(
      m, "BinomialLogitSpikeSlabSampler")
      .def(py::init([](BinomialLogitModel *model, const std::shared_ptr<MvnBase> &slab,
                       const std::shared_ptr<VariableSelectionPrior> &spike, int clt, RNG &rng) {
             return std::make_shared<BinomialLogitSpikeSlabSampler>(model, slab, spike, clt, rng);
           }),
           py::arg("model"), py::arg("slab"), py::arg("spike"), py::arg("clt_threshold") = 5,
           py::arg("seeding_rng") = std::ref(GlobalRng::rng), py::keep_alive<1, 2>())
      .def("draw_model_indicators", &BinomialLogitSpikeSlabSampler::draw_model_indicators)
      .def("draw_beta", &BinomialLogitSpikeSlabSampler::draw_beta)
      .def("log_model_prob", [](const BinomialLogitSpikeSlabSampler &s, const std::vector<bool> &bits) {
        Selector g((int)bits.size(), false);
        for (size_t i = 0; i < bits.size(); ++i) if (bits[i]) g.add((int)i);
        return s.log_model_prob(g);
      })
      .def("set_spike", &BinomialLogitSpikeSlabSampler::set_spike)
      .def("set_slab", &BinomialLogitSpikeSlabSampler::set_slab)
      .def("clone_to_new_host", &BinomialLogitSpikeSlabSampler::clone_to_new_host, py::keep_alive<0, 2>())
      .def("find_posterior_mode", &BinomialLogitSpikeSlabSampler::find_posterior_mode, py::arg("epsilon") = 1e-5)
      .def_property_readonly("posterior_mode_found", &BinomialLogitSpikeSlabSampler::posterior_mode_found)
      .def_property_readonly("log_posterior_at_mode", &BinomialLogitSpikeSlabSampler::log_posterior_at_mode)
      .def("allow_model_selection", &BinomialLogitSpikeSlabSampler::allow_model_selection)
      .def("limit_model_selection", &BinomialLogitSpikeSlabSampler::limit_model_selection)
      .def("set_active_set_statistics", &BinomialLogitSpikeSlabSampler::set_active_set_statistics)
      .def_property_readonly("active_set_statistics", &BinomialLogitSpikeSlabSampler::active_set_statistics)
      .def_property_readonly("active_set_columns_fetched", &BinomialLogitSpikeSlabSampler::active_set_columns_fetched);

  py::class_<BinomialProbitSpikeSlabSampler, PosteriorSampler, std::shared_ptr<BinomialProbitSpikeSlabSampler>>(
      m, "BinomialProbitSpikeSlabSampler")
      .def(py::init([](BinomialProbitModel *model, const std::shared_ptr<MvnBase> &slab,
                       const std::shared_ptr<VariableSelectionPrior> &spike, int clt, RNG &rng) {
             return std::make_shared<BinomialProbitSpikeSlabSampler>(model, slab, spike, clt, rng);
           }),
           py::arg("model"), py::arg("slab"), py::arg("spike"), py::arg("clt_threshold") = 10,
           py::arg("seeding_rng") = std::ref(GlobalRng::rng), py::keep_alive<1, 2>())
      .def("impute_latent_data", [](BinomialProbitSpikeSlabSampler &s) { py::gil_scoped_release rel; s.impute_latent_data(); })
      .def("refresh_xtx", &BinomialProbitSpikeSlabSampler::refresh_xtx)
      .def("complete_data_sufficient_statistics", &BinomialProbitSpikeSlabSampler::complete_data_sufficient_statistics)
      .def("allow_model_selection", &BinomialProbitSpikeSlabSampler::allow_model_selection)
      .def("limit_model_selection", &BinomialProbitSpikeSlabSampler::limit_model_selection);

  // ---- the Student-t sibling
  py::class_<DoubleModel, std::shared_ptr<DoubleModel>>(m, "DoubleModel").def("logp", &DoubleModel::logp);
  py::class_<UniformModel, DoubleModel, std::shared_ptr<UniformModel>>(m, "UniformModel")
      .def(py::init<double, double>(), py::arg("lo") = 0.0, py::arg("hi") = 1.0);
  py::class_<GammaModelBase, DoubleModel, std::shared_ptr<GammaModelBase>>(m, "GammaModelBase")
      .def_property_readonly("alpha", &GammaModelBase::alpha).def_property_readonly("beta", &GammaModelBase::beta);
  py::class_<GammaModel, GammaModelBase, std::shared_ptr<GammaModel>>(m, "GammaModel").def(py::init<double, double>(), py::arg("a"), py::arg("b"));
  py::class_<ChisqModel, GammaModel, std::shared_ptr<ChisqModel>>(m, "ChisqModel")
      .def(py::init<double, double>(), py::arg("df"), py::arg("sigma_estimate"));
  m.def("rgamma_mt", [](RNG &rng, double a, double b) { return rgamma_mt(rng, a, b); });
  m.def("rtrun_gamma_mt", [](RNG &rng, double a, double b, double cut) { return rtrun_gamma_mt(rng, a, b, cut); });
  // test hook: n successive draws of ScalarSliceSampler on a Python log density
  m.def("slice_sample", [](py::function logf, double x0, int n, double lo, double hi, bool unimodal, RNG &rng) {
    ScalarSliceSampler s([&logf](double x) { return logf(x).cast<double>(); }, unimodal, 1.0, &rng);
    s.set_lower_limit(lo); s.set_upper_limit(hi);
    Vector out((size_t)n);
    double x = x0;
    for (int i = 0; i < n; ++i) { x = s.draw(x); out[i] = x; }
    return from_vec(out);
  }, py::arg("logf"), py::arg("x0"), py::arg("n"), py::arg("lo") = -1.0 / 0.0, py::arg("hi") = 1.0 / 0.0, py::arg("unimodal") = false,
     py::arg("rng") = std::ref(GlobalRng::rng));
  m.def("draw_sigsq", [](const std::shared_ptr<GammaModelBase> &prior, double sigma_max, double df, double ss, int n, RNG &rng) {
    GenericGaussianVarianceSampler s(prior, sigma_max);
    Vector out((size_t)n);
    for (int i = 0; i < n; ++i) out[i] = s.draw(rng, df, ss);
    return from_vec(out);
  });

  py::class_<TRegressionModel, std::shared_ptr<TRegressionModel>> trm(m, "TRegressionModel");
  trm.def(py::init<int>(), py::arg("xdim"))
      .def(py::init([](const NpD &X, const NpD &y) {
        if (X.ndim() != 2 || y.size() != X.shape(0)) report_error("TRegressionModel(X, y): shape mismatch");
        return std::make_shared<TRegressionModel>((int64_t)X.shape(0), (int)X.shape(1), X.data(), y.data());
      }))
      .def("add_data", [](TRegressionModel &mo, double y, const NpD &x) { mo.add_data(y, to_vec(x)); })
      .def("borrow_host_data", [](TRegressionModel &mo, NpD X, NpD y) {
        if (X.ndim() != 2 || (int)X.shape(1) != mo.xdim() || y.size() != X.shape(0)) report_error("borrow_host_data(X, y): shape mismatch");
        auto keep = std::make_shared<std::tuple<NpD, NpD>>(X, y);
        mo.borrow_host_data((int64_t)X.shape(0), X.data(), (int64_t)X.shape(1), y.data(), std::shared_ptr<void>(keep, keep.get()));
      })
      .def("adopt_device_data", [](TRegressionModel &mo, int64_t n, uintptr_t dX, int64_t ldx, uintptr_t dy) {
        mo.adopt_device_data(n, reinterpret_cast<const double *>(dX), ldx, reinterpret_cast<const double *>(dy));
      })
      .def_property("sigsq", &TRegressionModel::sigsq, &TRegressionModel::set_sigsq)
      .def_property_readonly("sigma", &TRegressionModel::sigma)
      .def_property("nu", &TRegressionModel::nu, &TRegressionModel::set_nu)
      .def("set_sigsq", &TRegressionModel::set_sigsq)
      .def("set_nu", &TRegressionModel::set_nu)
      .def("log_likelihood", [](TRegressionModel &mo) { return mo.log_likelihood(); })
      .def("log_likelihood", [](TRegressionModel &mo, const NpD &b, double sigsq, double nu) { return mo.log_likelihood(to_vec(b), sigsq, nu); })
      .def("log_likelihood_same_beta", &TRegressionModel::log_likelihood_same_beta);
  bind_model_core(trm);

  py::class_<TRegressionSampler, PosteriorSampler, std::shared_ptr<TRegressionSampler>>(m, "TRegressionSampler")
      .def(py::init([](TRegressionModel *model, const std::shared_ptr<MvnBase> &coefficient_prior,
                       const std::shared_ptr<GammaModelBase> &siginv_prior, const std::shared_ptr<DoubleModel> &nu_prior, RNG &rng) {
             return std::make_shared<TRegressionSampler>(model, coefficient_prior, siginv_prior, nu_prior, rng);
           }),
           py::arg("model"), py::arg("coefficient_prior"), py::arg("siginv_prior"), py::arg("nu_prior"),
           py::arg("seeding_rng") = std::ref(GlobalRng::rng), py::keep_alive<1, 2>())
      .def("impute_latent_data", [](TRegressionSampler &s) { py::gil_scoped_release rel; s.impute_latent_data(); })
      .def("draw_beta_full_conditional", &TRegressionSampler::draw_beta_full_conditional)
      .def("draw_sigsq_full_conditional", &TRegressionSampler::draw_sigsq_full_conditional)
      .def("draw_nu_given_complete_data", &TRegressionSampler::draw_nu_given_complete_data)
      .def("draw_nu_given_observed_data", [](TRegressionSampler &s) { py::gil_scoped_release rel; s.draw_nu_given_observed_data(); })
      .def("set_sigma_upper_limit", &TRegressionSampler::set_sigma_upper_limit)
      .def("fix_latent_data", &TRegressionSampler::fix_latent_data, py::arg("fixed") = true)
      .def("clear_complete_data_sufficient_statistics", &TRegressionSampler::clear_complete_data_sufficient_statistics)
      .def("update_complete_data_sufficient_statistics", [](TRegressionSampler &s, double y, const NpD &x, double w) {
        s.update_complete_data_sufficient_statistics(y, to_vec(x), w); })
      .def_property_readonly("complete_data_sufficient_statistics", &TRegressionSampler::complete_data_sufficient_statistics,
                             py::return_value_policy::reference_internal)
      .def_property_readonly("likelihood_evaluations", &TRegressionSampler::likelihood_evaluations);

  py::class_<TRegressionSpikeSlabSampler, TRegressionSampler, std::shared_ptr<TRegressionSpikeSlabSampler>>(m, "TRegressionSpikeSlabSampler")
      .def(py::init([](TRegressionModel *model, const std::shared_ptr<MvnBase> &slab, const std::shared_ptr<VariableSelectionPrior> &spike,
                       const std::shared_ptr<GammaModelBase> &siginv_prior, const std::shared_ptr<DoubleModel> &nu_prior, RNG &rng) {
             return std::make_shared<TRegressionSpikeSlabSampler>(model, slab, spike, siginv_prior, nu_prior, rng);
           }),
           py::arg("model"), py::arg("coefficient_slab_prior"), py::arg("coefficient_spike_prior"), py::arg("siginv_prior"),
           py::arg("nu_prior"), py::arg("seeding_rng") = std::ref(GlobalRng::rng), py::keep_alive<1, 2>())
      .def("draw_model_indicators", &TRegressionSpikeSlabSampler::draw_model_indicators)
      .def("draw_included_coefficients", &TRegressionSpikeSlabSampler::draw_included_coefficients)
      .def("set_active_set_statistics", &TRegressionSpikeSlabSampler::set_active_set_statistics)
      .def_property_readonly("active_set_statistics", &TRegressionSpikeSlabSampler::active_set_statistics)
      .def_property_readonly("active_set_columns_fetched", &TRegressionSpikeSlabSampler::active_set_columns_fetched)
      .def("allow_model_selection", &TRegressionSpikeSlabSampler::allow_model_selection)
      .def("limit_model_selection", &TRegressionSpikeSlabSampler::limit_model_selection)
      .def("log_model_prob", [](TRegressionSpikeSlabSampler &s, const std::vector<bool> &bits) {
        Selector g((int)bits.size(), false);
        for (size_t i = 0; i < bits.size(); ++i) if (bits[i]) g.add((int)i);
        return s.log_model_prob(g);
      });

  py::class_<PoissonRegressionAuxMixSampler, PosteriorSampler, std::shared_ptr<PoissonRegressionAuxMixSampler>>(
      m, "PoissonRegressionAuxMixSampler")
      .def(py::init([](PoissonRegressionModel *model, const std::shared_ptr<MvnBase> &prior, int nthreads, RNG &rng) {
             return std::make_shared<PoissonRegressionAuxMixSampler>(model, prior, nthreads, rng);
           }),
           py::arg("model"), py::arg("prior"), py::arg("number_of_threads") = 1, py::arg("seeding_rng") = std::ref(GlobalRng::rng),
           py::keep_alive<1, 2>())
      .def("impute_latent_data", [](PoissonRegressionAuxMixSampler &s) { py::gil_scoped_release rel; s.impute_latent_data(); })
      .def("draw_beta_given_complete_data", &PoissonRegressionAuxMixSampler::draw_beta_given_complete_data)
      .def_property_readonly("complete_data_sufficient_statistics", &PoissonRegressionAuxMixSampler::complete_data_sufficient_statistics,
                             py::return_value_policy::reference_internal)
      .def("clear_complete_data_sufficient_statistics", &PoissonRegressionAuxMixSampler::clear_complete_data_sufficient_statistics)
      .def("update_complete_data_sufficient_statistics",
           [](PoissonRegressionAuxMixSampler &s, double sum, double prec, const NpD &x) {
             s.update_complete_data_sufficient_statistics(sum, prec, to_vec(x));
           })
      .def("fix_latent_data", &PoissonRegressionAuxMixSampler::fix_latent_data, py::arg("fixed") = true)
      .def("set_number_of_workers", &PoissonRegressionAuxMixSampler::set_number_of_workers);

  py::class_<PoissonRegressionSpikeSlabSampler, PoissonRegressionAuxMixSampler, std::shared_ptr<PoissonRegressionSpikeSlabSampler>>(
      m, "PoissonRegressionSpikeSlabSampler")
      .def(py::init([](PoissonRegressionModel *model, const std::shared_ptr<MvnBase> &slab,
                       const std::shared_ptr<VariableSelectionPrior> &spike, int nthreads, RNG &rng) {
             return std::make_shared<PoissonRegressionSpikeSlabSampler>(model, slab, spike, nthreads, rng);
           }),
           py::arg("model"), py::arg("slab"), py::arg("spike"), py::arg("number_of_threads") = 1,
           py::arg("seeding_rng") = std::ref(GlobalRng::rng), py::keep_alive<1, 2>())
      .def("clone_to_new_host", &PoissonRegressionSpikeSlabSampler::clone_to_new_host, py::keep_alive<0, 2>())
      .def("find_posterior_mode", &PoissonRegressionSpikeSlabSampler::find_posterior_mode, py::arg("epsilon") = 1e-5)
      .def_property_readonly("log_posterior_at_mode", &PoissonRegressionSpikeSlabSampler::log_posterior_at_mode)
      .def("allow_model_selection", &PoissonRegressionSpikeSlabSampler::allow_model_selection)
      .def("limit_model_selection", &PoissonRegressionSpikeSlabSampler::limit_model_selection)
      .def("draw_model_indicators", &PoissonRegressionSpikeSlabSampler::draw_model_indicators)
      .def("draw_beta", &PoissonRegressionSpikeSlabSampler::draw_beta)
      .def("log_model_prob", [](const PoissonRegressionSpikeSlabSampler &s, std::vector<bool> bits) {
        Selector g((int)bits.size(), false);
        for (size_t i = 0; i < bits.size(); ++i) if (bits[i]) g.add((int)i);
        return s.log_model_prob(g);
      })
      .def("set_active_set_statistics", &PoissonRegressionSpikeSlabSampler::set_active_set_statistics)
      .def_property_readonly("active_set_statistics", &PoissonRegressionSpikeSlabSampler::active_set_statistics)
      .def_property_readonly("active_set_columns_fetched", &PoissonRegressionSpikeSlabSampler::active_set_columns_fetched);

  // CPU-side test hook for the mode finder: the derivatives come from a Python callable (tests pass the oracle's), so
  // SpikeSlabCore::find_posterior_mode runs without a device
  m.def("find_posterior_mode_with", [](py::function derivs, const std::shared_ptr<MvnBase> &slab,
                                       const std::shared_ptr<VariableSelectionPrior> &spike, std::vector<bool> bits, const NpD &start,
                                       double epsilon) {
    struct CallbackModel : GlmModelBase {
      py::function f;
      CallbackModel(int p, py::function fn) : GlmModelBase(p, false), f(std::move(fn)) {}
      double log_likelihood_derivs(const Vector &b, Vector *g, SpdMatrix *h) override {
        py::tuple r = f(from_vec(b));
        if (g) *g = to_vec(r[1].cast<NpD>());
        if (h) *h = to_spd(r[2].cast<NpD>());
        return r[0].cast<double>();
      }
      void upload(DeviceData &) override {}
    };
    const int p = (int)bits.size();
    CallbackModel model(p, derivs);
    Selector g(p, false);
    for (int i = 0; i < p; ++i) if (bits[i]) g.add(i);
    model.coef().set_inc(g);
    model.coef().set_Beta(to_vec(start));
    double value = 0;
    const bool ok = SpikeSlabCore(slab, spike, false).find_posterior_mode(model, epsilon, &value);
    return py::make_tuple(ok, from_vec(model.Beta()), value);
  });

  // host linear algebra, exposed for the CPU-side tests
  m.def("cholesky_lower", [](const NpD &a) {
    SpdMatrix s = to_spd(a);
    bool ok = cholesky_lower(s.a.data(), s.dim);
    return py::make_tuple(ok, from_spd(s));
  });
  m.def("rmvn_suf", [](RNG &rng, const NpD &ivar, const NpD &ivar_mu) { return from_vec(rmvn_suf_mt(rng, to_spd(ivar), to_vec(ivar_mu))); });
  m.def("spike_slab_sweep", [](RNG &rng, const NpD &xtx, const NpD &xty, const std::shared_ptr<MvnBase> &slab,
                               const std::shared_ptr<VariableSelectionPrior> &spike, std::vector<bool> bits, int sweeps, bool fisher_yates) {
    // host-only driver of the inclusion sweep + beta draw on fixed statistics (CPU tests of the small-state steps)
    const int p = (int)xty.size();
    Vector packed((size_t)p * p + p + 4, 0.0);
    std::copy(xtx.data(), xtx.data() + (size_t)p * p, packed.begin());
    std::copy(xty.data(), xty.data() + p, packed.begin() + (size_t)p * p);
    WeightedRegSuf suf(p);
    suf.reset(packed.data(), p);
    GlmCoefs coef(p, false);
    Selector g(p, false);
    for (int i = 0; i < p; ++i) if (bits[i]) g.add(i);
    coef.set_inc(g);
    SpikeSlabCore core(slab, spike, fisher_yates);
    Vector inc_sum(p, 0.0), beta_sum(p, 0.0);
    for (int s = 0; s < sweeps; ++s) {
      core.draw_model_indicators(rng, coef, suf);
      core.draw_beta(rng, coef, suf);
      for (int i = 0; i < p; ++i) { inc_sum[i] += coef.inc()[i]; beta_sum[i] += coef.Beta()[i]; }
    }
    for (int i = 0; i < p; ++i) { inc_sum[i] /= sweeps; beta_sum[i] /= sweeps; }
    return py::make_tuple(from_vec(inc_sum), from_vec(beta_sum));
  });
  m.def("spike_slab_sweep_active", [](RNG &rng, const NpD &xtx, const NpD &xty, const std::shared_ptr<MvnBase> &slab,
                                      const std::shared_ptr<VariableSelectionPrior> &spike, std::vector<bool> bits, int sweeps,
                                      bool fisher_yates) {
    // the same driver on the ACTIVE-SET view of the statistics: before every sweep the view holds only the
    // columns of the included variables, the diagonal and X'Wz, and fetches a further column (here: from the same matrix)
    // when the sweep adds a variable.  Returns the chain of (inclusion bits, beta) per sweep and the number of fetches.
    const int p = (int)xty.size();
    const double *A = xtx.data();
    Vector xty_v(xty.data(), xty.data() + p), diag(p);
    for (int i = 0; i < p; ++i) diag[i] = A[(size_t)i * p + i];
    GlmCoefs coef(p, false);
    Selector g(p, false);
    for (int i = 0; i < p; ++i) if (bits[i]) g.add(i);
    coef.set_inc(g);
    SpikeSlabCore core(slab, spike, fisher_yates);
    py::list chain;
    int fetched = 0;
    for (int s = 0; s < sweeps; ++s) {
      std::vector<int> cols = coef.inc().included_positions();
      if (cols.empty()) cols.push_back(0);
      const int k = (int)cols.size();
      Vector G((size_t)p * k);
      for (int i = 0; i < p; ++i) for (int a = 0; a < k; ++a) G[(size_t)i * k + a] = A[(size_t)i * p + cols[a]];
      StatView view(p, cols, G, diag, xty_v, [&](int j, double *out) {
        for (int i = 0; i < p; ++i) out[i] = A[(size_t)i * p + j];
        ++fetched;
      });
      core.draw_model_indicators(rng, coef, view);
      core.draw_beta(rng, coef, view);
      std::vector<bool> inc(p);
      for (int i = 0; i < p; ++i) inc[i] = coef.inc()[i];
      chain.append(py::make_tuple(inc, from_vec(coef.Beta())));
    }
    return py::make_tuple(chain, fetched);
  });
  m.def("spike_slab_chain", [](RNG &rng, const NpD &xtx, const NpD &xty, const std::shared_ptr<MvnBase> &slab,
                               const std::shared_ptr<VariableSelectionPrior> &spike, std::vector<bool> bits, int sweeps, bool fisher_yates) {
    // spike_slab_sweep returning the chain itself (full statistics): what spike_slab_sweep_active is compared with
    const int p = (int)xty.size();
    Vector packed((size_t)p * p + p + 4, 0.0);
    std::copy(xtx.data(), xtx.data() + (size_t)p * p, packed.begin());
    std::copy(xty.data(), xty.data() + p, packed.begin() + (size_t)p * p);
    WeightedRegSuf suf(p);
    suf.reset(packed.data(), p);
    GlmCoefs coef(p, false);
    Selector g(p, false);
    for (int i = 0; i < p; ++i) if (bits[i]) g.add(i);
    coef.set_inc(g);
    SpikeSlabCore core(slab, spike, fisher_yates);
    py::list chain;
    for (int s = 0; s < sweeps; ++s) {
      core.draw_model_indicators(rng, coef, suf);
      core.draw_beta(rng, coef, suf);
      std::vector<bool> inc(p);
      for (int i = 0; i < p; ++i) inc[i] = coef.inc()[i];
      chain.append(py::make_tuple(inc, from_vec(coef.Beta())));
    }
    return chain;
  });
  m.def("flip_path_log_probs", [](const NpD &xtx, const NpD &xty, const std::shared_ptr<MvnBase> &slab,
                                  const std::shared_ptr<VariableSelectionPrior> &spike, std::vector<bool> bits,
                                  std::vector<int> flips, std::vector<bool> accept) {
    const int p = (int)xty.size();
    Vector packed((size_t)p * p + p + 4, 0.0);
    std::copy(xtx.data(), xtx.data() + (size_t)p * p, packed.begin());
    std::copy(xty.data(), xty.data() + p, packed.begin() + (size_t)p * p);
    WeightedRegSuf suf(p);
    suf.reset(packed.data(), p);
    Selector g(p, false);
    for (int i = 0; i < p; ++i) if (bits[i]) g.add(i);
    return from_vec(SpikeSlabCore(slab, spike, false).flip_path_log_probs(g, suf, flips, accept));
  });
  m.def("log_model_prob", [](const NpD &xtx, const NpD &xty, const std::shared_ptr<MvnBase> &slab,
                             const std::shared_ptr<VariableSelectionPrior> &spike, std::vector<bool> bits) {
    const int p = (int)xty.size();
    Vector packed((size_t)p * p + p + 4, 0.0);
    std::copy(xtx.data(), xtx.data() + (size_t)p * p, packed.begin());
    std::copy(xty.data(), xty.data() + p, packed.begin() + (size_t)p * p);
    WeightedRegSuf suf(p);
    suf.reset(packed.data(), p);
    Selector g(p, false);
    for (int i = 0; i < p; ++i) if (bits[i]) g.add(i);
    return SpikeSlabCore(slab, spike, false).log_model_prob(g, suf);
  });
}
