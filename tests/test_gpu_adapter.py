"""The drop-in claim, executed on the GPU box: oracle/_ref/boom_adapter_demo (built here from
boom_b200/boom_adapter + the compiled, unmodified reference) constructs BOOM's OWN model classes, attaches first the
reference sampler and then the B200 sampler with model->set_method(), and runs model->sample_posterior() on the same
data.  Posterior means / sds / inclusion probabilities of the two chains must agree within Monte Carlo error
(BASELINE.json north_star, 'Posterior')."""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "boom_adapter_demo")
TAU = 10.0   # integrated autocorrelation time assumed for the standard errors


def _demo(kind, n, p, nonzero, iters, burn):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/boom_adapter_demo not built (needs the reference sources: __graft_entry__.build())")
    out = subprocess.run([EXE, kind, str(n), str(p), str(nonzero), str(iters), str(burn)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr
    return json.loads(out.stdout.strip().splitlines()[-1])


def _agree(r, iters, burn, strong_only=False):
    ref, gpu = r["reference"], r["b200"]
    m0, m1 = np.array(ref["mean"]), np.array(gpu["mean"])
    s0, s1 = np.array(ref["sd"]), np.array(gpu["sd"])
    i0, i1 = np.array(ref["inclusion"]), np.array(gpu["inclusion"])
    # inclusion indicators of borderline variables are sticky: allow 5 standard errors at an autocorrelation time of 30 sweeps
    pi = 0.5 * (i0 + i1)
    tol = 5 * np.sqrt(pi * (1 - pi) * 30.0 / (iters - burn)) + 0.01
    assert np.all(np.abs(i0 - i1) < tol), (i0, i1, tol)
    sel = (i0 > 0.95) if strong_only else np.ones_like(i0, dtype=bool)
    se = np.sqrt((s0 ** 2 + s1 ** 2) * TAU / (iters - burn))
    assert np.all(np.abs(m0 - m1)[sel] < 4 * se[sel] + 1e-3), (m0, m1, se)
    big = sel & (s0 > 0.01)
    np.testing.assert_allclose(s1[big], s0[big], rtol=0.2)


@pytest.mark.parametrize("kind,n,p,nonzero", [("logit", 3000, 5, 3), ("poisson", 2000, 4, 2)])
def test_full_model_samplers_on_boom_models(kind, n, p, nonzero):
    iters, burn = 3000, 500
    _agree(_demo(kind, n, p, nonzero, iters, burn), iters, burn)


@pytest.mark.parametrize("kind,n,p,nonzero,iters", [("spike", 3000, 8, 3, 20000), ("pspike", 1500, 8, 3, 12000)])
def test_spike_slab_samplers_on_boom_models(kind, n, p, nonzero, iters):
    burn = 2000
    r = _demo(kind, n, p, nonzero, iters, burn)
    _agree(r, iters, burn, strong_only=True)
    inc = np.array(r["b200"]["inclusion"])
    assert np.all(inc[:nonzero + 1] > 0.9)    # the true variables are found


@pytest.mark.parametrize("kind", ["mode", "pmode"])
def test_find_posterior_mode_on_boom_models(kind):
    """find_posterior_mode of the reference samplers (max_nd2_careful on BOOM's own derivatives) and of the B200 samplers
    (Newton-Raphson on the device derivatives) reach the same mode and the same un-normalised log posterior."""
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/boom_adapter_demo not built")
    out = subprocess.run([EXE, kind, "3000", "7", "3", "0", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    np.testing.assert_allclose(r["b200_mode"], r["reference_mode"], rtol=1e-5, atol=1e-6)
    assert r["b200_log_posterior"] == pytest.approx(r["reference_log_posterior"], rel=1e-9)
    assert np.count_nonzero(r["b200_mode"]) == 4


def test_public_surface_beyond_draw():
    """draw_model_indicators / draw_beta / log_model_prob on the adapter (BinomialLogitSpikeSlabSampler.hpp:41-43), suf() in the
    reference's own BinomialLogit::SufficientStatistics type, externally driven statistics equal to the reference sampler's,
    and in-place row edits seen after observe_rows(true)."""
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/boom_adapter_demo not built")
    n = 2000
    out = subprocess.run([EXE, "api", str(n), "7", "3", "0", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["suf_sample_size"] == n
    np.testing.assert_allclose(r["log_model_prob_b200"], r["log_model_prob_reference"], rtol=1e-10)
    assert r["external_xtx_max_abs_diff"] < 1e-9 and r["external_xty_max_abs_diff"] < 1e-9
    assert r["reference_sample_size"] == r["b200_sample_size"] == n
    assert 1 <= r["nvars_after_sweep"] <= 7
    # x_01 += 100 on one row adds ~ info * (x^2 difference) >> the previous value of X'WX[1, 1]
    assert r["xtx11_after_edit"] > r["xtx11_before_edit"] + 100.0


def test_adapter_costs_what_the_standalone_classes_cost():
    """The BOOM-typed adapter (n heap objects behind model->dat()) and the standalone classes on the same rows: one Gibbs
    iteration takes the same time to within a few percent (p = 500: the statistics land in place, nothing p x p is copied)."""
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/boom_adapter_demo not built")
    out = subprocess.run([EXE, "bench", "200000", "500", "20", "8", "3"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["adapter_ms_per_iter"] < 1.15 * r["standalone_ms_per_iter"] + 0.3, r


def test_chunk_log_posterior_matches_the_reference():
    """BinomialLogitLogPostChunk (BinomialLogitCompositeSpikeSlabSampler.cpp:34-74): value, gradient and Hessian of the chunk
    log posterior from one device pass over the included columns against the reference's host loop, chunk by chunk
    (binomial rows with n_i in {1, 2, 3}; chunk sizes 2, 3 and the whole model)."""
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/boom_adapter_demo not built")
    out = subprocess.run([EXE, "chunk", "4000", "9", "4", "0", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert len(r["cases"]) >= 6
    for c in r["cases"]:
        assert c["value_b200"] == pytest.approx(c["value_ref"], rel=1e-11)
        assert c["grad_diff"] < 1e-8 and c["hess_diff"] < 1e-9 * max(1.0, c["hess_scale"])


def test_composite_sampler_on_boom_models():
    """BinomialLogitCompositeSpikeSlabSampler -- the sampler R's logit.spike constructs (logit_spike_slab_wrapper.cc:72-81):
    data augmentation, random-walk Metropolis and tailored-independence-Metropolis moves with equal weights; reference vs
    B200 chains on the same data agree within Monte Carlo error."""
    iters, burn = 9000, 1000
    r = _demo("composite", 2500, 8, 3, iters, burn)
    _agree(r, iters, burn, strong_only=True)
    inc = np.array(r["b200"]["inclusion"])
    assert np.all(inc[:4] > 0.9)
    assert r["time_report_lines"] >= 3


def test_probit_sibling_on_boom_models():
    """BinomialProbitSpikeSlabSampler on BOOM's BinomialProbitModel, binomial rows with n_i in {1, 2, 3, 15}:
    reference vs B200 chains agree within Monte Carlo error."""
    iters, burn = 8000, 1000
    r = _demo("probit", 2500, 8, 3, iters, burn)
    _agree(r, iters, burn, strong_only=True)
    assert np.all(np.array(r["b200"]["inclusion"])[:4] > 0.9)


@pytest.mark.parametrize("mode", ["active", "pactive"])
def test_active_set_option_on_the_adapter(mode):
    """set_active_set_statistics(true) on B200::BinomialLogitSpikeSlabSampler / B200::PoissonRegressionSpikeSlabSampler: the same
    chain as with the full statistics (same seed), columns fetched only when the sweep adds a variable, suf() still the full
    matrix."""
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/boom_adapter_demo not built")
    out = subprocess.run([EXE, mode, "20000", "120", "5", "30", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["same_model"] is True and r["chain_max_abs_diff"] < 1e-7
    assert r["columns_fetched"] >= 5 and r["suf_xtx_rel_diff"] < 1e-9


def test_student_t_sibling_on_boom_models():
    """TRegressionSampler on BOOM's TRegressionModel, y = x'beta + 1.5 t_4: the reference sampler and the B200
    sampler (BOOM's own rmvn / GenericGaussianVarianceSampler / ScalarSliceSampler around the device step) give the same
    posterior of (beta, sigma, nu) within Monte Carlo error; the device log likelihood equals TRegressionModel's own."""
    iters, burn = 6000, 1000
    r = _demo("treg", 2500, 5, 3, iters, burn)
    ref, gpu = r["reference"], r["b200"]
    m0, m1, s0, s1 = (np.array(ref["mean"]), np.array(gpu["mean"]), np.array(ref["sd"]), np.array(gpu["sd"]))
    tau = np.r_[np.full(len(m0) - 2, TAU), 40.0, 40.0]      # sigma and nu mix slower than beta
    se = np.sqrt((s0 ** 2 + s1 ** 2) * tau / (iters - burn))
    assert np.all(np.abs(m0 - m1) < 4 * se + 1e-3), (m0, m1, se)
    np.testing.assert_allclose(s1, s0, rtol=0.2)
    assert abs(m1[-2] - 1.5) < 0.2 and 2.5 < m1[-1] < 7.0      # sigma and nu near the truth
    assert r["loglike_b200"] == pytest.approx(r["loglike_reference"], rel=1e-12)
    assert r["suf_consistency"] < 1e-10


def test_student_t_spike_slab_on_boom_models():
    """TRegressionSpikeSlabSampler (what lm.spike builds for Student errors) on BOOM's TRegressionModel: reference vs B200 --
    inclusion probabilities, the coefficients of the strongly included variables, sigma and nu."""
    iters, burn = 8000, 1000
    r = _demo("tspike", 2500, 10, 3, iters, burn)
    _agree(r, iters, burn, strong_only=True)
    assert np.all(np.array(r["b200"]["inclusion"])[:4] > 0.9)
    m0, m1 = np.array(r["reference_sigma_nu_mean"]), np.array(r["b200_sigma_nu_mean"])
    s0, s1 = np.array(r["reference_sigma_nu_sd"]), np.array(r["b200_sigma_nu_sd"])
    se = np.sqrt((s0 ** 2 + s1 ** 2) * 40.0 / (iters - burn))
    assert np.all(np.abs(m0 - m1) < 4 * se + 1e-3), (m0, m1, se)


def test_active_set_option_on_the_student_t_adapter():
    """set_active_set_statistics(true) on B200::TRegressionSpikeSlabSampler: the same chain of (beta, sigsq, nu) as with the full
    statistics (same seed), columns fetched when the sweep adds a variable, complete_data_sufficient_statistics() still the full
    matrix."""
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/boom_adapter_demo not built")
    out = subprocess.run([EXE, "tactive", "12000", "120", "5", "25", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["same_model"] is True and r["chain_max_abs_diff"] < 1e-6
    assert r["columns_fetched"] >= 5 and r["suf_xtx_rel_diff"] < 1e-9
