"""The Student-t sibling: TRegressionSampler's data-augmentation step and what surrounds it.

Reference side: tests/golden/ref_student.json, written by oracle/ref_driver.cpp (golden_student) from the unmodified
reference: 2 x 10^5 draws of TDataImputer::impute per (residual, sigma, nu) case, TRegressionModel::log_likelihood values,
and a TRegressionSampler chain on the shared synthetic data (oracle.synth_student).

CPU (-m "not gpu"): the oracle's gamma draw against the exact Gamma law and the reference's quantile bins; the oracle's log
likelihood against the reference's; the host-side scalar samplers (slice sampler, variance draw, truncated gamma) against
their exact laws; the host full conditionals of beta and sigma^2 on fixed statistics.
GPU (-m gpu, through the C ABI): the device weights value by value against the oracle on the same Philox stream, the
statistics <= 1e-12 normwise, the log likelihood <= 1e-12, the draws against the reference's distribution, and the
posterior of a chain run through model.set_method(sampler); model.sample_posterior() against the reference chain."""
import numpy as np
import pytest
from scipy import stats

from oracle import oracle as O
from tests.helpers import normwise_err, vec_err

ALPHA = 1e-6
M = 200_000


def _quantile_bin_chi2(x, ref_quantiles, n_ref):
    edges = np.asarray(ref_quantiles)
    counts = np.bincount(np.searchsorted(edges, x, side="right"), minlength=len(edges) + 1)
    expect = len(x) / (len(edges) + 1.0)
    chi2 = ((counts - expect) ** 2 / expect).sum() / (1.0 + len(x) / n_ref)
    assert chi2 < stats.chi2.ppf(1 - ALPHA, len(edges)), ("quantile-bin chi-square", chi2, counts)


def check_weight_draws(draw, golden):
    """draw(residual, sigma, nu) -> M weights; against Gamma((nu+1)/2, rate (nu + delta^2)/2) and the reference's draws."""
    for c in golden("ref_student.json")["draw_stats"]:
        w = draw(c["residual"], c["sigma"], c["nu"])
        shape = 0.5 * (c["nu"] + 1.0)
        rate = 0.5 * (c["nu"] + (c["residual"] / c["sigma"]) ** 2)
        assert w.min() > 0
        assert stats.kstest(w, "gamma", args=(shape, 0, 1.0 / rate)).pvalue > ALPHA, ("KS vs Gamma", c)
        _quantile_bin_chi2(w, c["w_quantiles"], int(c["N"]))
        se = np.sqrt(w.var() / len(w) + c["w_var"] / c["N"])
        assert abs(w.mean() - c["w_mean"]) < stats.norm.ppf(1 - ALPHA / 2) * se
        assert abs(np.log(w).mean() - c["logw_mean"]) < 6 * np.log(w).std() * np.sqrt(1.0 / len(w) + 1.0 / c["N"])


def _single_column_problem(residual, m=M):
    """m rows of the one-column design x = 1 with y = residual and beta = 0: every row has the same conditional law."""
    return np.ones((m, 1)), np.full(m, residual), np.zeros(1)


# ------------------------------------------------------------------------------------------ CPU
def test_oracle_weights_follow_the_reference_law(golden):
    def draw(residual, sigma, nu):
        X, y, beta = _single_column_problem(residual)
        return O.student_step(X, y, beta, sigma, nu, 20261017, 5)[3]
    check_weight_draws(draw, golden)


def test_oracle_statistics_are_weighted_reg_suf():
    X, y, beta = O.synth_student(4000, 6, 3, 31)
    xtwx, xtwy, sc, w = O.student_step(X, y, beta * 0.8, 1.3, 5.0, 9, 2)
    np.testing.assert_allclose(xtwx, (X * w[:, None]).T @ X, rtol=1e-12)
    np.testing.assert_allclose(xtwy, X.T @ (w * y), rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(sc, [len(y), (w * y * y).sum(), w.sum(), np.log(w).sum()], rtol=1e-11)
    # the stream is keyed by the global row: a shard draws what the whole problem draws
    w2 = O.student_step(X[1000:], y[1000:], beta * 0.8, 1.3, 5.0, 9, 2, row_offset=1000)[3]
    assert np.array_equal(w2, w[1000:])


def test_oracle_loglike_matches_reference(golden):
    g = golden("ref_student.json")["loglike"]
    X, y, bt = O.synth_student(g["n"], g["p"], g["nonzero"], g["seed"])
    beta = bt * g["beta_scale"] + g["beta_shift"]
    for r in g["rows"]:
        assert O.student_loglike(X, y, beta, r["sigma"], r["nu"]) == pytest.approx(r["loglike"], rel=1e-12)


def test_slice_sampler_targets_its_density():
    import boom_b200
    H = boom_b200.host()
    # Gamma(3, rate 2) on (0, inf): lower limit only (the configuration of the nu samplers), and doubly bounded
    logf = lambda x: 2.0 * np.log(x) - 2.0 * x if x > 0 else -np.inf      # noqa: E731
    d = H.slice_sample(logf, 1.0, 40000, lo=0.0, rng=boom_b200.RNG(5))[::8]
    assert stats.kstest(d, "gamma", args=(3.0, 0, 0.5)).pvalue > 1e-4
    d = H.slice_sample(logf, 1.0, 40000, lo=0.5, hi=2.0, rng=boom_b200.RNG(6))[::8]
    g = stats.gamma(3.0, 0, 0.5)
    cdf = lambda t: (g.cdf(t) - g.cdf(0.5)) / (g.cdf(2.0) - g.cdf(0.5))   # noqa: E731
    assert d.min() >= 0.5 and d.max() <= 2.0 and stats.kstest(d, cdf).pvalue > 1e-4
    # unbounded, bimodal: the doubling procedure has to cross the valley
    logf2 = lambda x: np.logaddexp(-0.5 * (x + 2.5) ** 2, -0.5 * (x - 2.5) ** 2)   # noqa: E731
    d = H.slice_sample(logf2, 0.3, 60000, rng=boom_b200.RNG(7))[::10]
    cdf2 = lambda t: 0.5 * (stats.norm.cdf(t + 2.5) + stats.norm.cdf(t - 2.5))     # noqa: E731
    assert stats.kstest(d, cdf2).pvalue > 1e-4


def test_variance_draw_and_truncated_gamma():
    import boom_b200
    H = boom_b200.host()
    prior = boom_b200.ChisqModel(3.0, 1.2)          # Gamma(1.5, 1.5 * 1.44) on 1 / sigsq
    assert prior.alpha == 1.5 and prior.beta == pytest.approx(2.16)
    df, ss = 40.0, 55.0
    a, b = (df + 2 * prior.alpha) / 2, (ss + 2 * prior.beta) / 2
    d = H.draw_sigsq(prior, np.inf, df, ss, 50000, boom_b200.RNG(11))
    assert stats.kstest(1.0 / d, "gamma", args=(a, 0, 1.0 / b)).pvalue > 1e-4
    # sigma <= sigma_max: 1 / sigsq ~ Gamma(a, b) given > 1 / sigma_max^2, on either side of the mode
    g = stats.gamma(a, 0, 1.0 / b)
    for smax in (1.4, 1.0):
        cut = 1.0 / smax ** 2
        d = H.draw_sigsq(prior, smax, df, ss, 50000, boom_b200.RNG(12))
        assert d.max() <= smax ** 2 * (1 + 1e-12)
        cdf = lambda t: (g.cdf(t) - g.cdf(cut)) / g.sf(cut)                # noqa: E731
        assert stats.kstest(1.0 / d, cdf).pvalue > 1e-4, smax
    rng = boom_b200.RNG(13)
    d = np.array([H.rtrun_gamma_mt(rng, 0.7, 1.5, 0.4) for _ in range(30000)])   # shape below one
    g = stats.gamma(0.7, 0, 1 / 1.5)
    assert stats.kstest(d, lambda t: (g.cdf(t) - g.cdf(0.4)) / g.sf(0.4)).pvalue > 1e-4


def test_host_full_conditionals_on_fixed_statistics():
    """beta | suf, sigsq ~ N(V (Ominv b + X'Wy / sigsq), V), V^-1 = Ominv + X'WX / sigsq (TRegressionSampler.cpp:74-85), and
    1 / sigsq | suf, beta ~ Gamma(a + n / 2, b + SSE / 2) (.cpp:165-171), with the statistics supplied from outside
    (fix_latent_data: no device involved)."""
    import boom_b200
    p, n = 4, 300
    X, y, _ = O.synth_student(n, p, 2, 17)
    w = O.student_step(X, y, np.zeros(p), 1.0, 4.0, 3, 0)[3]
    model = boom_b200.TRegressionModel(p)
    prior = boom_b200.MvnModel(np.full(p, 0.2), 4.0 * np.eye(p))
    siginv_prior = boom_b200.GammaModel(2.0, 3.0)
    sampler = boom_b200.TRegressionSampler(model, prior, siginv_prior, boom_b200.UniformModel(0.5, 60.0), boom_b200.RNG(8))
    sampler.fix_latent_data(True)
    for i in range(n):
        sampler.update_complete_data_sufficient_statistics(y[i], X[i], w[i])
    suf = sampler.complete_data_sufficient_statistics
    np.testing.assert_allclose(suf.xtx, (X * w[:, None]).T @ X, rtol=1e-12)
    assert suf.sumw == pytest.approx(w.sum()) and suf.sumlogw == pytest.approx(np.log(w).sum()) and suf.n == n
    model.sigsq = 1.7
    draws = []
    for _ in range(20000):
        sampler.draw_beta_full_conditional()
        draws.append(model.Beta.copy())
    draws = np.array(draws)
    prec = np.eye(p) / 4.0 + suf.xtx / 1.7
    V = np.linalg.inv(prec)
    mean = V @ (np.full(p, 0.2) / 4.0 + suf.xty / 1.7)
    assert np.all(np.abs(draws.mean(0) - mean) < 5 * np.sqrt(np.diag(V) / len(draws)))
    np.testing.assert_allclose(np.cov(draws.T), V, atol=0.06 * np.max(np.diag(V)))
    beta = mean.copy()
    model.Beta = beta
    sse = (w * (y - X @ beta) ** 2).sum()
    s2 = []
    for _ in range(20000):
        sampler.draw_sigsq_full_conditional()
        s2.append(model.sigsq)
    assert stats.kstest(1.0 / np.array(s2), "gamma", args=(2.0 + n / 2, 0, 1.0 / (3.0 + sse / 2))).pvalue > 1e-4
    # nu | weights: the complete-data draw stays inside the prior's support and moves
    nus = []
    for _ in range(300):
        sampler.draw_nu_given_complete_data()
        nus.append(model.nu)
    assert 0.5 <= min(nus) and max(nus) <= 60.0 and np.std(nus) > 0
    lp = sampler.logpri()
    assert np.isfinite(lp)


def test_spike_slab_steps_see_the_statistics_scaled_by_sigsq():
    """TRegressionSpikeSlabSampler: SpikeSlabSampler::log_model_prob(gamma, suf, sigsq) (SpikeSlabSampler.cpp:176-199) =
    log pi(gamma) + 1/2 log|Om_g| - 1/2 mu_g' Om_g mu_g - sum log diag chol(Om_g + X'WX_gg / sigsq)
    + 1/2 |L^-1 (X'Wy_g / sigsq + Om_g mu_g)|^2, on externally supplied statistics (no device)."""
    import boom_b200
    p, n = 6, 400
    X, y, _ = O.synth_student(n, p, 2, 23)
    w = O.student_step(X, y, np.zeros(p), 1.0, 4.0, 3, 0)[3]
    model = boom_b200.TRegressionModel(p)
    mu0 = np.linspace(-0.2, 0.3, p)
    slab = boom_b200.MvnModel(mu0, 2.0 * np.eye(p))
    pi = np.array([0.9, 0.5, 0.3, 0.2, 0.4, 0.1])
    spike = boom_b200.VariableSelectionPrior(pi)
    s = boom_b200.TRegressionSpikeSlabSampler(model, slab, spike, boom_b200.GammaModel(2.0, 3.0), boom_b200.UniformModel(0.5, 60.0),
                                              boom_b200.RNG(4))
    s.fix_latent_data(True)
    for i in range(n):
        s.update_complete_data_sufficient_statistics(y[i], X[i], w[i])
    xtx, xty = (X * w[:, None]).T @ X, X.T @ (w * y)
    for sigsq in (0.6, 2.3):
        model.sigsq = sigsq
        for g in ([1, 0, 0, 0, 0, 0], [1, 1, 0, 1, 0, 0], [1, 1, 1, 1, 1, 1], [0, 0, 1, 0, 0, 1]):
            g = np.array(g, dtype=bool)
            om = np.eye(int(g.sum())) / 2.0
            prec = om + xtx[np.ix_(g, g)] / sigsq
            L = np.linalg.cholesky(prec)
            S = np.linalg.solve(L, xty[g] / sigsq + om @ mu0[g])
            want = (np.log(pi[g]).sum() + np.log1p(-pi[~g]).sum() + 0.5 * np.linalg.slogdet(om)[1] - 0.5 * mu0[g] @ om @ mu0[g]
                    - np.log(np.diag(L)).sum() + 0.5 * S @ S)
            assert s.log_model_prob(list(g)) == pytest.approx(want, rel=1e-12)
    # the sweep and the coefficient draw run on those statistics and leave excluded coefficients at exactly zero
    model.drop_all(); model.add(0)
    incs = []
    for _ in range(300):
        s.draw_model_indicators(); s.draw_included_coefficients(); s.draw_sigsq_full_conditional()
        incs.append(model.inc.copy())
        assert np.all(model.Beta[~model.inc] == 0.0)
    assert np.mean(incs, axis=0)[0] > 0.9 and np.isfinite(s.logpri())
    s.allow_model_selection(False)
    before = model.inc.copy()
    s.draw_model_indicators()
    assert np.array_equal(before, model.inc)


def test_error_behaviour_of_the_student_t_host_classes():
    """report_error -> RuntimeError, as in the reference: dimension mismatches at construction, invalid parameters, a slice sampler
    started where the target is -inf (ScalarSliceSampler.cpp:check_finite)."""
    import boom_b200
    H = boom_b200.host()
    p = 3
    model = boom_b200.TRegressionModel(p)
    good = (boom_b200.MvnModel(np.zeros(p), np.eye(p)), boom_b200.GammaModel(1.0, 1.0), boom_b200.UniformModel(0.5, 60.0))
    with pytest.raises(RuntimeError):
        boom_b200.TRegressionSampler(model, boom_b200.MvnModel(np.zeros(p + 1), np.eye(p + 1)), good[1], good[2], boom_b200.RNG(1))
    with pytest.raises(RuntimeError):
        boom_b200.TRegressionSpikeSlabSampler(model, good[0], boom_b200.VariableSelectionPrior(p + 2, 0.5), good[1], good[2], boom_b200.RNG(1))
    for bad in (lambda: model.set_sigsq(0.0), lambda: model.set_nu(-1.0), lambda: model.add_data(1.0, np.zeros(p + 1)),
                lambda: boom_b200.GammaModel(0.0, 1.0), lambda: boom_b200.UniformModel(2.0, 1.0)):
        with pytest.raises(RuntimeError):
            bad()
    with pytest.raises(RuntimeError):                                  # the target is -inf at the starting point
        H.slice_sample(lambda x: -np.inf if x < 1.0 else -x, 0.5, 3, lo=0.0, rng=boom_b200.RNG(2))
    s = boom_b200.TRegressionSampler(model, *good, boom_b200.RNG(1))
    with pytest.raises(RuntimeError):
        s.set_sigma_upper_limit(-1.0)
    u = boom_b200.UniformModel(0.5, 60.0)
    assert u.logp(0.4) == -np.inf and u.logp(1.0) == pytest.approx(-np.log(59.5))
    g = boom_b200.GammaModel(2.0, 3.0)
    assert g.logp(-1.0) == -np.inf and g.logp(0.7) == pytest.approx(stats.gamma.logpdf(0.7, 2.0, scale=1 / 3.0))
    # defaults of the model's parameters (TRegression.cpp:33-35) and their setters
    assert model.sigsq == 1.0 and model.nu == 30.0
    model.sigsq = 2.25; model.nu = 7.0
    assert model.sigma == 1.5 and model.nu == 7.0


# ------------------------------------------------------------------------------------------ GPU
def _ctx(X, y, path=0):
    import boom_b200
    ctx = boom_b200.Context(0)
    ctx.set_option("path", path)
    ctx.upload_regression(X, y)
    return ctx


@pytest.mark.gpu
@pytest.mark.parametrize("n,p,path", [(1, 1, 0), (777, 3, 0), (5000, 16, 0), (5000, 16, 2), (4099, 37, 0), (3000, 50, 0), (2500, 64, 0),
                                      (2048, 65, 0), (3001, 130, 0), (1500, 500, 0)])
def test_student_step_matches_oracle(n, p, path):
    X, y, bt = O.synth_student(n, p, min(p, 5), 100 + p)
    beta = bt * 0.9
    sigma, nu, seed, it = 1.4, 3.5, 99, 4
    ctx = _ctx(X, y, path)
    w = ctx.student_draw(beta, sigma, nu, seed, it)
    xtwx, xtwy, sc = ctx.student_step(beta, sigma, nu, seed, it)
    ctx.close()
    rxtwx, rxtwy, rsc, rw = O.student_step(X, y, beta, sigma, nu, seed, it)
    # the weights value by value: the device inverts the normal cdf with CUDA's normcdfinv, the oracle with AS241
    np.testing.assert_allclose(w, rw, rtol=1e-11)
    assert normwise_err(xtwx, rxtwx) < 1e-12
    assert vec_err(xtwy, rxtwy) < 1e-11
    assert sc[0] == n
    np.testing.assert_allclose(sc[1:], rsc[1:], rtol=1e-10, atol=1e-9)


@pytest.mark.gpu
def test_student_step_edge_cases():
    import boom_b200
    X, y, bt = O.synth_student(600, 5, 2, 3)
    ctx = _ctx(X, y)
    for sigma, nu in [(1.0, 0.4), (0.2, 1.0), (5.0, 250.0)]:      # shape below one (the boost draw), heavy and light tails
        w = ctx.student_draw(bt, sigma, nu, 5, 0)
        np.testing.assert_allclose(w, O.student_step(X, y, bt, sigma, nu, 5, 0)[3], rtol=1e-11)
    w0 = ctx.student_draw(bt, 1.0, 4.0, 5, 0)
    assert np.array_equal(w0, ctx.student_draw(bt, 1.0, 4.0, 5, 0))          # same (seed, iteration): same bits
    assert not np.array_equal(w0, ctx.student_draw(bt, 1.0, 4.0, 5, 1))
    ctx.set_row_offset(1000)                                                 # sharding: the stream follows the global row
    np.testing.assert_allclose(ctx.student_draw(bt, 1.0, 4.0, 5, 0), O.student_step(X, y, bt, 1.0, 4.0, 5, 0, row_offset=1000)[3], rtol=1e-11)
    for bad in [(0.0, 4.0), (1.0, 0.0), (1.0, float("nan")), (-1.0, 3.0)]:
        with pytest.raises(boom_b200.BoomGpuError):
            ctx.student_step(bt, bad[0], bad[1], 1, 0)
    ybad = y.copy(); ybad[17] = np.inf
    ctx.upload_regression(X, ybad)
    with pytest.raises(boom_b200.BoomGpuError):                              # a non-finite residual is reported, not accumulated
        ctx.student_step(bt, 1.0, 4.0, 1, 0)
    ctx.close()
    # binomial data do not answer to the Student-t entry points, and the other way round
    Xb, yb, nb, bb = O.synth_binomial(100, 3, 2, 1)
    cb = boom_b200.Context(0); cb.upload_binomial(Xb, yb, nb)
    with pytest.raises(boom_b200.BoomGpuError):
        cb.student_step(bb, 1.0, 4.0, 1, 0)
    cb.close()
    # chunked upload of regression rows (what the BOOM adapter does) and empty data
    c2 = boom_b200.Context(0)
    c2.upload_chunked(X, y, np.zeros(len(y)), 128, poisson=2)
    np.testing.assert_allclose(c2.student_draw(bt, 1.0, 4.0, 5, 0), O.student_step(X, y, bt, 1.0, 4.0, 5, 0)[3], rtol=1e-11)
    c2.upload_regression(X[:0], y[:0])
    xtwx, xtwy, sc = c2.student_step(bt, 1.0, 4.0, 5, 0)
    assert not xtwx.any() and not xtwy.any() and not sc.any()
    c2.close()


@pytest.mark.gpu
def test_student_loglike_matches_reference_and_oracle(golden):
    g = golden("ref_student.json")["loglike"]
    X, y, bt = O.synth_student(g["n"], g["p"], g["nonzero"], g["seed"])
    beta = bt * g["beta_scale"] + g["beta_shift"]
    ctx = _ctx(X, y)
    for i, r in enumerate(g["rows"]):
        ll = ctx.student_loglike(beta if i == 0 else None, r["sigma"], r["nu"])   # one pass over X, then stored residuals
        assert ll == pytest.approx(r["loglike"], rel=1e-12)
    ctx.close()
    X, y, bt = O.synth_student(300_000, 33, 5, 8)
    ctx = _ctx(X, y)
    assert ctx.student_loglike(bt, 1.5, 4.0) == pytest.approx(O.student_loglike(X, y, bt, 1.5, 4.0), rel=1e-12)
    assert ctx.student_loglike(None, 1.1, 7.0) == pytest.approx(O.student_loglike(X, y, bt, 1.1, 7.0), rel=1e-12)
    ctx.close()
    import boom_b200
    c = boom_b200.Context(0); c.upload_regression(X[:10], y[:10])
    with pytest.raises(boom_b200.BoomGpuError):
        c.student_loglike(None, 1.0, 4.0)            # no residuals yet
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", [0, 2])
def test_device_weights_follow_the_reference_law(golden, path):
    def draw(residual, sigma, nu):
        X, y, beta = _single_column_problem(residual)
        ctx = _ctx(X, y, path)
        w = ctx.student_draw(beta, sigma, nu, 424242, 1)
        ctx.close()
        return w
    check_weight_draws(draw, golden)


@pytest.mark.gpu
def test_student_chain_matches_reference(golden):
    """Posterior of (beta, sigma, nu) through the drop-in surface against the reference's TRegressionSampler chain on the same
    data and priors, within Monte Carlo standard error (tau = 10 for beta; sigma and nu mix slower: tau = 40)."""
    import boom_b200
    g = golden("ref_student.json"); c = g["chain"]
    p = c["p"]
    X, y, _ = O.synth_student(c["n"], p, c["nonzero"], c["seed"], c["sigma_true"], c["nu_true"])
    model = boom_b200.TRegressionModel(X, y)
    assert model.sigsq == 1.0 and model.nu == 30.0                    # TRegression.cpp:33-35
    sampler = boom_b200.TRegressionSampler(model, boom_b200.MvnModel(np.zeros(p), c["beta_prior_variance"] * np.eye(p)),
                                           boom_b200.ChisqModel(*c["siginv_prior"]), boom_b200.UniformModel(*c["nu_prior"]),
                                           boom_b200.RNG(77))
    model.set_method(sampler)
    iters, burn = 5000, 800
    betas, sn = [], []
    for it in range(iters):
        model.sample_posterior()
        if it >= burn:
            betas.append(model.Beta.copy()); sn.append((model.sigma, model.nu))
    betas, sn = np.array(betas), np.array(sn)
    n_ref = c["iters"] - c["burn"]

    def check(draws, ref_mean, ref_sd, tau):
        se = np.sqrt(ref_sd ** 2 * tau / n_ref + draws.std(0) ** 2 * tau / len(draws))
        assert np.all(np.abs(draws.mean(0) - ref_mean) < 4 * se + 1e-4), (draws.mean(0), ref_mean, se)
        np.testing.assert_allclose(draws.std(0), ref_sd, rtol=0.2)
    check(betas, np.array(g["chain_beta_mean"]), np.array(g["chain_beta_sd"]), 10.0)
    check(sn, np.array(g["chain_sigma_nu_mean"]), np.array(g["chain_sigma_nu_sd"]), 40.0)
    suf = sampler.complete_data_sufficient_statistics
    assert suf.n == c["n"] and suf.sumw > 0 and np.isfinite(suf.sumlogw)
    assert sampler.likelihood_evaluations >= 3 * iters
    # externally driven statistics: with the latent data fixed the device is not asked again
    sampler.fix_latent_data(True)
    before = model.kernel_launches()
    sampler.impute_latent_data()
    assert model.kernel_launches() == before
    # same seed, same chain
    def chain(seed):
        m = boom_b200.TRegressionModel(X[:500], y[:500])
        s = boom_b200.TRegressionSampler(m, boom_b200.MvnModel(np.zeros(p), 100.0 * np.eye(p)), boom_b200.ChisqModel(1.0, 1.0),
                                         boom_b200.UniformModel(0.5, 60.0), boom_b200.RNG(seed))
        m.set_method(s)
        for _ in range(20):
            m.sample_posterior()
        return np.r_[m.Beta, m.sigsq, m.nu]
    assert np.array_equal(chain(5), chain(5)) and not np.array_equal(chain(5), chain(6))


@pytest.mark.gpu
def test_student_spike_slab_chain_matches_reference(golden):
    """TRegressionSpikeSlabSampler through the drop-in surface against the reference's chain on the same data and priors:
    inclusion probabilities, coefficients of the strongly included variables, sigma and nu."""
    import boom_b200
    g = golden("ref_student.json"); c = g["spike_chain"]
    p = c["p"]
    X, y, _ = O.synth_student(c["n"], p, c["nonzero"], c["seed"], c["sigma_true"], c["nu_true"])
    model = boom_b200.TRegressionModel(X, y)
    model.drop_all(); model.add(0)
    sampler = boom_b200.TRegressionSpikeSlabSampler(model, boom_b200.MvnModel(np.zeros(p), c["slab_variance"] * np.eye(p)),
                                                    boom_b200.VariableSelectionPrior(p, c["prior_inclusion"]),
                                                    boom_b200.ChisqModel(*c["siginv_prior"]), boom_b200.UniformModel(*c["nu_prior"]),
                                                    boom_b200.RNG(78))
    model.set_method(sampler)
    iters, burn = 5000, 800
    betas, incs, sn = [], [], []
    for it in range(iters):
        model.sample_posterior()
        if it >= burn:
            betas.append(model.Beta.copy()); incs.append(model.inc.copy()); sn.append((model.sigma, model.nu))
    betas, incs, sn = np.array(betas), np.array(incs, dtype=float), np.array(sn)
    ref_inc = np.array(g["spike_chain_inclusion"])
    assert np.max(np.abs(incs.mean(0) - ref_inc)) < 0.05
    strong = ref_inc > 0.95
    n_ref = c["iters"] - c["burn"]
    rm, rs = np.array(g["spike_chain_beta_mean"]), np.array(g["spike_chain_beta_sd"])
    se = np.sqrt(rs ** 2 * 10.0 / n_ref + betas.std(0) ** 2 * 10.0 / len(betas))
    assert np.all(np.abs(betas.mean(0) - rm)[strong] < 4 * se[strong] + 1e-4)
    np.testing.assert_allclose(betas.std(0)[strong], rs[strong], rtol=0.2)
    rm2, rs2 = np.array(g["spike_chain_sigma_nu_mean"]), np.array(g["spike_chain_sigma_nu_sd"])
    se2 = np.sqrt(rs2 ** 2 * 40.0 / n_ref + sn.std(0) ** 2 * 40.0 / len(sn))
    assert np.all(np.abs(sn.mean(0) - rm2) < 4 * se2 + 1e-4)


@pytest.mark.gpu
def test_student_active_set_matches_the_full_statistics():
    """boomgpu_student_step_active: the columns of the active set, the diagonal and X'Wy of the SAME weights as the full step
    (same seed / iteration), a further column on demand, the full matrix from the weights kept in HBM."""
    n, p = 6000, 150
    X, y, bt = O.synth_student(n, p, 5, 61)
    ctx = _ctx(X, y)
    xtwx, xtwy, sc = ctx.student_step(bt, 1.3, 4.5, 11, 3)
    act = [0, 1, 2, 3, 4, 5, 77, 149]
    G, diag, xty, sc2 = ctx.student_step_active(bt, 1.3, 4.5, 11, 3, act)
    d = np.sqrt(np.diag(xtwx))
    assert np.max(np.abs(G - xtwx[:, act]) / np.outer(d, d[act])) < 1e-12
    np.testing.assert_allclose(diag, np.diag(xtwx), rtol=1e-12)
    assert vec_err(xty, xtwy) < 1e-11
    np.testing.assert_allclose(sc2, sc, rtol=1e-11)
    col = ctx.weighted_column(33)
    assert np.max(np.abs(col - xtwx[:, 33]) / (d * d[33])) < 1e-12
    fxx, fxy = ctx.full_statistics()
    assert normwise_err(fxx, xtwx) < 1e-12 and vec_err(fxy, xtwy) < 1e-11
    ctx.close()


@pytest.mark.gpu
def test_student_active_set_chain_is_the_full_statistics_chain():
    """TRegressionSpikeSlabSampler.set_active_set_statistics(True): the sweep, the coefficient draw and the sigma^2 draw read the
    same numbers (up to the summation order of the two device kernels), so the chain is the chain of the full-matrix sampler."""
    import boom_b200
    n, p = 8000, 140
    X, y, _ = O.synth_student(n, p, 6, 62)

    def chain(active, iters=25):
        model = boom_b200.TRegressionModel(X, y)
        model.drop_all(); model.add(0)
        s = boom_b200.TRegressionSpikeSlabSampler(model, boom_b200.MvnModel(np.zeros(p), 4.0 * np.eye(p)),
                                                  boom_b200.VariableSelectionPrior(p, 0.05), boom_b200.ChisqModel(1.0, 1.0),
                                                  boom_b200.UniformModel(0.5, 60.0), boom_b200.RNG(9))
        s.set_active_set_statistics(active)
        model.set_method(s)
        out = []
        for _ in range(iters):
            model.sample_posterior()
            out.append(np.r_[model.Beta, model.sigsq, model.nu])
        return np.array(out), s, model
    full, _, _ = chain(False)
    act, s, model = chain(True)
    assert np.array_equal(full[:, :p] != 0, act[:, :p] != 0)            # the same models, iteration by iteration
    np.testing.assert_allclose(act, full, rtol=1e-7, atol=1e-9)
    assert s.active_set_columns_fetched >= 5                             # the true variables entered through fetched columns
    suf = s.complete_data_sufficient_statistics                          # the full matrix on demand, from the weights in HBM
    assert suf.n == n and np.all(np.isfinite(suf.xtx)) and suf.xtx[0, 0] == pytest.approx(suf.sumw, rel=1e-12)
