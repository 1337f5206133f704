"""Posterior agreement with the reference's own CPU chains (tests/golden/ref_chains.json, written by
oracle/ref_driver.cpp from the unmodified reference) on the same synthetic data, through the
drop-in sampler surface: model.set_method(sampler); model.sample_posterior().

Tolerance (BASELINE.json north_star: 'within Monte Carlo standard error'): chains are autocorrelated,
so the standard error of a posterior mean is taken as sd * sqrt(tau / N) with tau = 10."""
import numpy as np
import pytest

import boom_b200
from oracle import oracle as O

pytestmark = pytest.mark.gpu

TAU = 10.0


def _check_moments(draws, ref_mean, ref_sd, n_ref):
    mean, sd = draws.mean(0), draws.std(0)
    se = np.sqrt(ref_sd ** 2 * TAU / n_ref + sd ** 2 * TAU / len(draws))
    assert np.all(np.abs(mean - ref_mean) < 4 * se + 1e-4), (mean, ref_mean, se)
    big = ref_sd > 0.01
    np.testing.assert_allclose(sd[big], ref_sd[big], rtol=0.15)


def _run(model, iters, burn):
    out = []
    for it in range(iters):
        model.sample_posterior()
        if it >= burn:
            out.append(model.Beta.copy())
    return np.array(out)


def test_logit_auxmix_posterior(golden):
    g = golden("ref_chains.json"); c = g["logit_auxmix"]
    X, y, nt, _ = O.synth_binomial(c["n"], c["p"], c["nonzero"], c["seed"], c["max_trials"])
    model = boom_b200.BinomialLogitModel(X, y, nt)
    prior = boom_b200.MvnModel(np.zeros(c["p"]), np.eye(c["p"]))
    sampler = boom_b200.BinomialLogitAuxmixSampler(model, prior, 10, boom_b200.RNG(42))
    model.set_method(sampler)
    draws = _run(model, 4000, 500)
    _check_moments(draws, np.array(g["logit_auxmix_mean"]), np.array(g["logit_auxmix_sd"]), c["iters"] - c["burn"])
    assert sampler.suf.sample_size == c["n"]


def test_logit_binomial_clt_posterior(golden):
    g = golden("ref_chains.json"); c = g["logit_binomial"]
    X, y, nt, _ = O.synth_binomial(c["n"], c["p"], c["nonzero"], c["seed"], c["max_trials"])
    model = boom_b200.BinomialLogitModel(X, y, nt)
    sampler = boom_b200.BinomialLogitAuxmixSampler(model, boom_b200.MvnModel(np.zeros(c["p"]), np.eye(c["p"])), 10,
                                                   boom_b200.RNG(43))
    model.set_method(sampler)
    draws = _run(model, 4000, 500)
    _check_moments(draws, np.array(g["logit_binomial_mean"]), np.array(g["logit_binomial_sd"]), c["iters"] - c["burn"])


def test_logit_spike_slab_posterior(golden):
    g = golden("ref_chains.json"); c = g["logit_spike_slab"]
    p = c["p"]
    X, y, nt, _ = O.synth_binomial(c["n"], p, c["nonzero"], c["seed"], c["max_trials"])
    model = boom_b200.BinomialLogitModel(X, y, nt)
    slab = boom_b200.MvnModel(np.zeros(p), np.eye(p))
    spike = boom_b200.VariableSelectionPrior(p, c["prior_inclusion"])
    sampler = boom_b200.BinomialLogitSpikeSlabSampler(model, slab, spike, 10, boom_b200.RNG(44))
    model.set_method(sampler)
    betas, incs = [], []
    for it in range(4000):
        model.sample_posterior()
        if it >= 500:
            betas.append(model.Beta.copy()); incs.append(model.inc.copy())
    betas, incs = np.array(betas), np.array(incs, dtype=float)
    ref_inc = np.array(g["logit_spike_slab_inclusion"])
    assert np.max(np.abs(incs.mean(0) - ref_inc)) < 0.05
    strong = ref_inc > 0.95
    _check_moments(betas[:, strong], np.array(g["logit_spike_slab_mean"])[strong], np.array(g["logit_spike_slab_sd"])[strong],
                   c["iters"] - c["burn"])
    # excluded coefficients are exact zeros (GlmCoefs.cpp:304-309)
    assert np.all(betas[incs == 0] == 0)


def test_poisson_auxmix_posterior(golden):
    g = golden("ref_chains.json"); c = g["poisson_auxmix"]
    X, y, ex, _ = O.synth_poisson(c["n"], c["p"], c["nonzero"], c["seed"])
    boom_b200.load_poisson_mixture_table()
    model = boom_b200.PoissonRegressionModel(X, y, ex)
    sampler = boom_b200.PoissonRegressionAuxMixSampler(model, boom_b200.MvnModel(np.zeros(c["p"]), np.eye(c["p"])), 1,
                                                       boom_b200.RNG(45))
    model.set_method(sampler)
    draws = _run(model, 4000, 500)
    _check_moments(draws, np.array(g["poisson_auxmix_mean"]), np.array(g["poisson_auxmix_sd"]), c["iters"] - c["burn"])
    suf = sampler.complete_data_sufficient_statistics
    assert suf.n == c["n"] + np.count_nonzero(y)


def test_poisson_spike_slab_recovers_signal():
    """The R test of the reference (Interfaces/R/BoomSpikeSlab/tests/testthat/test-poisson.R:70-124):
    the true variables are found, the noise variables are not."""
    n, p = 4000, 10
    X, y, ex, beta = O.synth_poisson(n, p, 3, seed=99)
    boom_b200.load_poisson_mixture_table()
    model = boom_b200.PoissonRegressionModel(X, y, ex)
    sampler = boom_b200.PoissonRegressionSpikeSlabSampler(model, boom_b200.MvnModel(np.zeros(p), np.eye(p)),
                                                          boom_b200.VariableSelectionPrior(p, 0.3), 1, boom_b200.RNG(46))
    model.set_method(sampler)
    incs = []
    for it in range(1500):
        model.sample_posterior()
        if it >= 300:
            incs.append(model.inc.copy())
    inc = np.array(incs, dtype=float).mean(0)
    assert np.all(inc[:4] > 0.9) and np.all(inc[4:] < 0.2)


def test_same_seed_repeatability():
    """test-logit.R:23-45 of the reference: the same seed gives the same chain."""
    X, y, nt, _ = O.synth_binomial(2000, 6, 3, seed=8)
    out = []
    for _ in range(2):
        model = boom_b200.BinomialLogitModel(X, y, nt)
        sampler = boom_b200.BinomialLogitSpikeSlabSampler(model, boom_b200.MvnModel(np.zeros(6), np.eye(6)),
                                                          boom_b200.VariableSelectionPrior(6, 0.5), 10, boom_b200.RNG(5))
        model.set_method(sampler)
        out.append(_run(model, 30, 0))
    np.testing.assert_array_equal(out[0], out[1])


def test_externally_driven_statistics():
    """The state-space callers' path (StateSpaceLogitPosteriorSampler.cpp:58,111-123): latent data fixed,
    statistics pushed row by row, then draw_params()."""
    X, y, nt, _ = O.synth_binomial(300, 3, 2, seed=4)
    model = boom_b200.BinomialLogitModel(X, y, nt)
    sampler = boom_b200.BinomialLogitAuxmixSampler(model, boom_b200.MvnModel(np.zeros(3), np.eye(3)), 10, boom_b200.RNG(3))
    sampler.fix_latent_data(True)
    sampler.clear_complete_data_sufficient_statistics()
    rng = np.random.default_rng(0)
    w = 0.2 + rng.random(300); s = rng.normal(size=300)
    for i in range(300):
        sampler.update_complete_data_sufficient_statistics(s[i], w[i], X[i])
    sampler.impute_latent_data()   # must be a no-op now
    xtx, xty = O.accumulate(X, w, s)
    np.testing.assert_allclose(sampler.suf.xtx, xtx, rtol=1e-12)
    np.testing.assert_allclose(sampler.suf.xty, xty, rtol=1e-12)
    sampler.draw_params()
    assert np.all(np.isfinite(model.Beta))


def test_data_added_after_construction_is_picked_up():
    X, y, nt, _ = O.synth_binomial(500, 4, 2, seed=6)
    model = boom_b200.BinomialLogitModel(4)
    for i in range(250):
        model.add_data(y[i], nt[i], X[i])
    sampler = boom_b200.BinomialLogitAuxmixSampler(model, boom_b200.MvnModel(np.zeros(4), np.eye(4)), 10, boom_b200.RNG(2))
    model.set_method(sampler)
    model.sample_posterior()
    assert sampler.suf.sample_size == 250
    for i in range(250, 500):
        model.add_data(y[i], nt[i], X[i])
    model.sample_posterior()
    assert sampler.suf.sample_size == 500
    assert model.log_likelihood() == pytest.approx(O.binomial_logit_loglike(X, y, nt, model.Beta), rel=1e-12)


def test_find_posterior_mode_matches_newton_on_oracle_derivatives():
    """BinomialLogitSpikeSlabSampler::find_posterior_mode (.cpp:123-177): the mode of log slab + log likelihood over the
    included coefficients; checked against Newton-Raphson in numpy on the oracle's gradient / Hessian."""
    n, p = 4000, 9
    X, y, nt, _ = O.synth_binomial(n, p, 4, seed=15, max_trials=3)
    inc = np.array([1, 1, 0, 1, 1, 0, 0, 1, 0], dtype=bool)
    mu = np.linspace(-0.1, 0.1, p)
    A = np.random.default_rng(2).normal(size=(p, p)); siginv = A @ A.T / p + np.eye(p)
    model = boom_b200.BinomialLogitModel(X, y, nt)
    model.set_inc(list(inc))
    sampler = boom_b200.BinomialLogitSpikeSlabSampler(model, boom_b200.MvnModel(mu, siginv, True),
                                                      boom_b200.VariableSelectionPrior(p, 0.5), 10, boom_b200.RNG(1))
    model.set_method(sampler)
    sampler.find_posterior_mode(1e-8)
    assert sampler.posterior_mode_found
    idx = np.flatnonzero(inc)
    b = np.zeros(len(idx))
    P = siginv[np.ix_(idx, idx)]
    for _ in range(50):
        full = np.zeros(p); full[idx] = b
        _, g, h = O.binomial_logit_loglike_derivs(X, y, nt, full)
        grad = g[idx] - P @ (b - mu[idx])
        step = np.linalg.solve(P - h[np.ix_(idx, idx)], grad)
        b = b + step
        if np.max(np.abs(step)) < 1e-13:
            break
    beta = np.array(model.Beta)
    np.testing.assert_allclose(beta[idx], b, rtol=1e-7, atol=1e-9)
    assert np.all(beta[~inc] == 0)
    full = np.zeros(p); full[idx] = b
    ll = O.binomial_logit_loglike(X, y, nt, full)
    k = len(idx)
    logprior = -0.5 * k * np.log(2 * np.pi) + 0.5 * np.linalg.slogdet(P)[1] - 0.5 * (b - mu[idx]) @ P @ (b - mu[idx])
    assert sampler.log_posterior_at_mode == pytest.approx(ll + logprior, rel=1e-10)


def test_poisson_find_posterior_mode_zero_gradient():
    n, p = 3000, 6
    X, y, ex, _ = O.synth_poisson(n, p, 3, seed=16)
    boom_b200.load_poisson_mixture_table()
    model = boom_b200.PoissonRegressionModel(X, y, ex)
    sampler = boom_b200.PoissonRegressionSpikeSlabSampler(model, boom_b200.MvnModel(np.zeros(p), np.eye(p)),
                                                          boom_b200.VariableSelectionPrior(p, 0.5), 1, boom_b200.RNG(1))
    model.set_method(sampler)
    sampler.find_posterior_mode(1e-9)
    beta = np.array(model.Beta)
    _, g, _ = O.poisson_loglike_derivs(X, y, ex, beta)
    assert np.max(np.abs(g - beta)) < 1e-5 * n     # gradient of the log posterior: g - Siginv (beta - 0) = 0
    assert np.isfinite(sampler.log_posterior_at_mode)


def test_borrowed_host_rows_give_the_same_chain():
    """borrow_host_data: rows stay in the caller's arrays (no second host copy); same chain as the copying constructor."""
    X, y, nt, _ = O.synth_binomial(3000, 5, 2, seed=12)
    out = []
    for borrow in (False, True):
        if borrow:
            model = boom_b200.BinomialLogitModel(5)
            model.borrow_host_data(X, y, nt)
        else:
            model = boom_b200.BinomialLogitModel(X, y, nt)
        sampler = boom_b200.BinomialLogitAuxmixSampler(model, boom_b200.MvnModel(np.zeros(5), np.eye(5)), 10, boom_b200.RNG(9))
        model.set_method(sampler)
        out.append(_run(model, 20, 0))
        assert model.sample_size == 3000
    np.testing.assert_array_equal(out[0], out[1])


def test_standalone_probit_sampler_recovers_the_coefficients():
    """BinomialProbitModel + BinomialProbitSpikeSlabSampler through the Python surface: X'NX is computed once, every further
    iteration is the X'z pass; the chain finds the true variables and their values."""
    import boom_b200
    from scipy import stats
    rng = np.random.default_rng(3)
    n, p = 20000, 10
    X = rng.normal(size=(n, p))
    X[:, 0] = 1.0
    beta = np.zeros(p)
    beta[:4] = [-0.5, 0.8, -0.6, 0.4]
    nt = rng.integers(1, 4, size=n).astype(float)
    y = rng.binomial(nt.astype(int), stats.norm.cdf(X @ beta)).astype(float)
    model = boom_b200.BinomialProbitModel(X, y, nt)
    s = boom_b200.BinomialProbitSpikeSlabSampler(model, boom_b200.MvnModel(np.zeros(p), np.eye(p)),
                                                 boom_b200.VariableSelectionPrior(p, 0.3), 10, boom_b200.RNG(7))
    model.set_method(s)
    model.drop_all()
    model.add(0)
    draws = []
    for it in range(600):
        model.sample_posterior()
        if it >= 100:
            draws.append(np.array(model.Beta))
    draws = np.array(draws)
    suf = s.complete_data_sufficient_statistics()
    np.testing.assert_allclose(np.array(suf.xtx), (X.T * nt) @ X, rtol=1e-11)
    inc = np.mean(draws != 0, axis=0)
    assert np.all(inc[:4] > 0.95) and np.all(inc[4:] < 0.3)
    np.testing.assert_allclose(draws.mean(axis=0)[:4], beta[:4], atol=0.05)


def test_active_set_statistics_give_the_same_chain():
    """set_active_set_statistics(True): the device computes X'WX only for the included columns (+ diagonal, X'Wz) and the sweep
    fetches a column when it adds a variable.  Same seed -> the same chain as with the full matrix (the sweep reads nothing
    else of it), and suf still answers with the full statistics, computed on demand from the latents in HBM."""
    import boom_b200
    from oracle import oracle as O
    n, p = 30_000, 150
    X, y, nt, beta_true = O.synth_binomial(n, p, 6, seed=77, max_trials=1)
    chains = []
    fetched = 0
    for active in (False, True):
        model = boom_b200.BinomialLogitModel(X, y, nt)
        s = boom_b200.BinomialLogitSpikeSlabSampler(model, boom_b200.MvnModel(np.zeros(p), np.eye(p)),
                                                    boom_b200.VariableSelectionPrior(p, 6.0 / p), 10, boom_b200.RNG(21))
        s.set_active_set_statistics(active)
        model.set_method(s)
        model.drop_all()
        model.add(0)
        out = []
        for it in range(40):
            model.sample_posterior()
            out.append(np.array(model.Beta))
        chains.append(np.array(out))
        if active:
            fetched = s.active_set_columns_fetched
            xtx_on_demand = np.array(s.suf.xtx)          # the full matrix of the last iteration's latents
            assert xtx_on_demand.shape == (p, p) and np.all(np.isfinite(xtx_on_demand))
            np.testing.assert_array_equal(xtx_on_demand, xtx_on_demand.T)
            assert s.suf.sample_size == n
        else:
            full_last = np.array(s.suf.xtx)
    # identical decisions, coefficients equal up to the summation order of the two device kernels
    assert np.array_equal(chains[0] != 0, chains[1] != 0)
    np.testing.assert_allclose(chains[1], chains[0], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(xtx_on_demand, full_last, rtol=1e-9, atol=1e-9)
    assert fetched >= 6        # the six true variables entered one by one, each through a fetched column
    assert np.count_nonzero(chains[1][-1]) >= 6


def test_poisson_active_set_statistics_give_the_same_chain():
    """The same option on PoissonRegressionSpikeSlabSampler (boomgpu_poisson_step_active): identical decisions, coefficients
    equal up to summation order, the full WeightedRegSuf (with its four scalars) on demand."""
    import boom_b200
    from oracle import oracle as O
    n, p = 20_000, 120
    X, y, ex, beta_true = O.synth_poisson(n, p, 5, seed=78)
    boom_b200.load_poisson_mixture_table()
    chains, sufs = [], []
    fetched = 0
    for active in (False, True):
        model = boom_b200.PoissonRegressionModel(X, y, ex)
        s = boom_b200.PoissonRegressionSpikeSlabSampler(model, boom_b200.MvnModel(np.zeros(p), np.eye(p)),
                                                        boom_b200.VariableSelectionPrior(p, 5.0 / p), 1, boom_b200.RNG(22))
        s.set_active_set_statistics(active)
        assert s.active_set_statistics == active
        model.set_method(s)
        model.drop_all()
        model.add(0)
        out = []
        for it in range(40):
            model.sample_posterior()
            out.append(np.array(model.Beta))
        chains.append(np.array(out))
        suf = s.complete_data_sufficient_statistics
        sufs.append((np.array(suf.xtx), np.array(suf.xty), suf.n, suf.yty, suf.sumw, suf.sumlogw))
        if active:
            fetched = s.active_set_columns_fetched
    assert np.array_equal(chains[0] != 0, chains[1] != 0)
    np.testing.assert_allclose(chains[1], chains[0], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(sufs[1][0], sufs[0][0], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(sufs[1][1], sufs[0][1], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(sufs[1][2:], sufs[0][2:], rtol=1e-10)
    assert fetched >= 3
    assert np.count_nonzero(chains[1][-1]) >= 4
