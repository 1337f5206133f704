"""CPU tests of the host-side small-state steps (the parts that stay on the host as in the
reference): Cholesky, the beta draw, log_model_prob and the inclusion sweep, on FIXED
sufficient statistics.  No GPU, no oracle needed: the checks are against numpy / exact enumeration.
Reference: BinomialLogitSpikeSlabSampler.cpp:56-117,180-222; distributions/mvn.cpp:128-136."""
import itertools

import numpy as np
import pytest

import boom_b200


def _suf(p, n=400, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(n, p)); X[:, 0] = 1
    w = 0.1 + rng.random(n)
    beta = np.zeros(p); beta[:3] = [0.8, -0.7, 0.6]
    z = X @ beta + rng.normal(size=n) / np.sqrt(w)
    return (X.T * w) @ X, X.T @ (w * z)


def test_cholesky_matches_numpy():
    h = boom_b200.host()
    xtx, _ = _suf(40)
    ok, L = h.cholesky_lower(xtx)
    assert ok
    np.testing.assert_allclose(L, np.linalg.cholesky(xtx), rtol=1e-12, atol=1e-12)
    # the blocked / vectorised / threaded code paths: sizes around the 64-wide block, ragged 4 x 4 tiles, and large
    # enough (p = 900) for the row-parallel update
    for p in (1, 3, 63, 64, 65, 130, 259, 900):
        A = np.random.default_rng(p).normal(size=(p + 30, p))
        xtx = A.T @ A + np.eye(p)
        ok, L = h.cholesky_lower(xtx)
        ref = np.linalg.cholesky(xtx)
        assert ok and np.max(np.abs(L - ref)) < 1e-11 * np.max(np.abs(ref)), p
        assert np.all(np.triu(L, 1) == 0)
    ok, _ = h.cholesky_lower(-np.eye(3))
    assert not ok


def test_rmvn_suf_moments():
    h = boom_b200.host()
    p = 4
    xtx, xty = _suf(p)
    rng = boom_b200.RNG(11)
    draws = np.array([h.rmvn_suf(rng, xtx, xty) for _ in range(20000)])
    cov = np.linalg.inv(xtx)
    mean = cov @ xty
    se = np.sqrt(np.diag(cov) / len(draws))
    assert np.all(np.abs(draws.mean(0) - mean) < 5 * se)
    mc = 5 * np.sqrt(np.outer(np.diag(cov), np.diag(cov)) * 2 / len(draws))   # 5 sigma of a covariance estimate
    assert np.all(np.abs(np.cov(draws.T) - cov) < mc)


def _log_model_prob_numpy(xtx, xty, mu, siginv, probs, g):
    idx = np.flatnonzero(g)
    num = np.sum(np.where(g, np.log(probs), np.log1p(-probs)))
    if len(idx) == 0:
        return num
    iv = siginv[np.ix_(idx, idx)]
    num += 0.5 * np.linalg.slogdet(iv)[1]
    m = mu[idx]
    num -= 0.5 * m @ iv @ m
    post = iv + xtx[np.ix_(idx, idx)]
    L = np.linalg.cholesky(post)
    S = np.linalg.solve(L, xty[idx] + iv @ m)
    denom = np.sum(np.log(np.diag(L))) - 0.5 * S @ S
    return num - denom


def test_log_model_prob_matches_formula():
    h = boom_b200.host()
    p = 7
    xtx, xty = _suf(p, seed=3)
    mu = np.linspace(-0.2, 0.2, p)
    A = np.random.default_rng(1).normal(size=(p, p)); siginv = A @ A.T + p * np.eye(p)
    probs = np.full(p, 0.3)
    slab = boom_b200.MvnModel(mu, siginv, True)
    spike = boom_b200.VariableSelectionPrior(probs)
    for g in ([1, 0, 0, 0, 0, 0, 0], [1, 1, 1, 0, 0, 1, 0], [0] * 7, [1] * 7):
        g = np.array(g, dtype=bool)
        assert h.log_model_prob(xtx, xty, slab, spike, list(g)) == pytest.approx(
            _log_model_prob_numpy(xtx, xty, mu, siginv, probs, g), rel=1e-11, abs=1e-9)


@pytest.mark.parametrize("fisher_yates", [False, True])
def test_inclusion_sweep_matches_exact_enumeration(fisher_yates):
    """Marginal inclusion probabilities of the Gibbs sweep on fixed statistics vs the exact
    posterior over all 2^p models (p = 6)."""
    h = boom_b200.host()
    p = 6
    xtx, xty = _suf(p, n=60, seed=5)
    mu = np.zeros(p); siginv = np.eye(p)
    probs = np.full(p, 0.4)
    slab = boom_b200.MvnModel(mu, siginv, True)
    spike = boom_b200.VariableSelectionPrior(probs)
    lp, gs = [], []
    for bits in itertools.product([0, 1], repeat=p):
        g = np.array(bits, dtype=bool)
        gs.append(g); lp.append(_log_model_prob_numpy(xtx, xty, mu, siginv, probs, g))
    lp = np.array(lp); w = np.exp(lp - lp.max()); w /= w.sum()
    exact = (np.array(gs) * w[:, None]).sum(0)
    inc, _ = h.spike_slab_sweep(boom_b200.RNG(7), xtx, xty, slab, spike, [True] + [False] * (p - 1), 30000, fisher_yates)
    assert np.max(np.abs(inc - exact)) < 0.02


def test_spike_prior_limits():
    spike = boom_b200.VariableSelectionPrior(np.array([1.0, 0.5, 0.0]))
    slab = boom_b200.MvnModel(np.zeros(3), np.eye(3))
    xtx, xty = _suf(3)
    h = boom_b200.host()
    # variable 2 has prior probability 0: including it is a zero-support point
    assert h.log_model_prob(xtx, xty, slab, spike, [True, False, True]) == -np.inf
    inc, _ = h.spike_slab_sweep(boom_b200.RNG(1), xtx, xty, slab, spike, [True, False, False], 200, False)
    assert inc[0] == 1.0 and inc[2] == 0.0


def test_errors_surface_as_exceptions():
    with pytest.raises(RuntimeError, match="wrong size"):
        boom_b200.BinomialLogitModel(3).add_data(1, 1, np.zeros(2))
    with pytest.raises(RuntimeError, match=r"\[0, n\]"):
        boom_b200.BinomialLogitModel(2).add_data(3, 1, np.zeros(2))
    with pytest.raises(RuntimeError, match="dimension"):
        boom_b200.BinomialLogitAuxmixSampler(boom_b200.BinomialLogitModel(3), boom_b200.MvnModel(np.zeros(2), np.eye(2)))
    m = boom_b200.PoissonRegressionModel(2)
    with pytest.raises(RuntimeError, match="no sampler"):
        m.sample_posterior()


def test_bordered_flip_evaluator_matches_from_scratch():
    """The sweep evaluates 'add j' proposals by bordering the Cholesky factors of the current model; along a
    random path of proposals/acceptances (dense prior precision, non-zero prior mean, unequal spike
    probabilities, a model-size cap) every value equals log_model_prob computed from scratch."""
    h = boom_b200.host()
    rng = np.random.default_rng(12)
    p = 12
    X = rng.normal(size=(300, p)); w = 0.1 + rng.random(300); z = rng.normal(size=300)
    xtx = (X.T * w) @ X; xty = X.T @ (w * z)
    A = rng.normal(size=(p, p)); siginv = A @ A.T + p * np.eye(p)
    for mu, cap in ((np.zeros(p), -1), (rng.normal(size=p) * 0.3, -1), (rng.normal(size=p) * 0.3, 7)):
        slab = boom_b200.MvnModel(mu, siginv, True)
        spike = boom_b200.VariableSelectionPrior(rng.uniform(0.2, 0.8, size=p))
        if cap > 0:
            spike.set_max_model_size(cap)
        g = [bool(b) for b in rng.integers(0, 2, size=p)]
        if cap > 0:
            g = [True] * 3 + [False] * (p - 3)
        flips = [int(j) for j in rng.integers(0, p, size=600)]
        acc = [bool(a) for a in rng.integers(0, 2, size=600)]
        vals = h.flip_path_log_probs(xtx, xty, slab, spike, g, flips, acc)
        for f, a, v in zip(flips, acc, vals):
            g2 = list(g); g2[f] = not g2[f]
            ref = h.log_model_prob(xtx, xty, slab, spike, g2)
            if np.isinf(ref):
                assert v == ref
            else:
                assert v == pytest.approx(ref, rel=1e-12, abs=1e-10)
                if a:
                    g = g2
        assert vals[-1] == pytest.approx(h.log_model_prob(xtx, xty, slab, spike, g), rel=1e-12, abs=1e-10)


def test_spike_slab_surface_setters_and_clone():
    """set_spike / set_slab check dimensions (BinomialLogitSpikeSlabSampler.hpp:68-74, .cpp:224-240); clone_to_new_host carries
    the priors and settings to another model (.cpp:42-48)."""
    p = 4
    m1, m2 = boom_b200.BinomialLogitModel(p), boom_b200.BinomialLogitModel(p)
    slab = boom_b200.MvnModel(np.zeros(p), np.eye(p))
    spike = boom_b200.VariableSelectionPrior(p, 0.3)
    s = boom_b200.BinomialLogitSpikeSlabSampler(m1, slab, spike, 7, boom_b200.RNG(2))
    s.limit_model_selection(2)
    s.set_spike(boom_b200.VariableSelectionPrior(p, 0.5))
    s.set_slab(boom_b200.MvnModel(np.ones(p), 2 * np.eye(p)))
    with pytest.raises(RuntimeError, match="dimension"):
        s.set_spike(boom_b200.VariableSelectionPrior(p + 1, 0.5))
    with pytest.raises(RuntimeError, match="dimension"):
        s.set_slab(boom_b200.MvnModel(np.zeros(p - 1), np.eye(p - 1)))
    c = s.clone_to_new_host(m2)
    assert c.clt_threshold == 7 and isinstance(c, boom_b200.BinomialLogitSpikeSlabSampler)
    pm = boom_b200.PoissonRegressionModel(p)
    ps = boom_b200.PoissonRegressionSpikeSlabSampler(pm, slab, spike, 1, boom_b200.RNG(3))
    assert isinstance(ps.clone_to_new_host(boom_b200.PoissonRegressionModel(p)), boom_b200.PoissonRegressionSpikeSlabSampler)


def test_find_posterior_mode_on_oracle_derivatives():
    """SpikeSlabCore::find_posterior_mode (the objective of BinomialLogitSpikeSlabSampler.cpp:123-177) driven by the ORACLE's
    log likelihood / gradient / Hessian instead of the device's: the Newton iteration, the step halving and the selection
    of the included sub-blocks are checked on CPU against numpy."""
    from oracle import oracle as O
    h = boom_b200.host()
    n, p = 1500, 7
    X, y, nt, _ = O.synth_binomial(n, p, 3, seed=33, max_trials=4)
    inc = np.array([1, 0, 1, 1, 0, 1, 0], dtype=bool)
    mu = np.linspace(-0.2, 0.2, p)
    A = np.random.default_rng(3).normal(size=(p, p)); siginv = A @ A.T / p + np.eye(p)
    slab = boom_b200.MvnModel(mu, siginv, True)
    spike = boom_b200.VariableSelectionPrior(p, 0.5)
    ok, beta, value = h.find_posterior_mode_with(lambda b: O.binomial_logit_loglike_derivs(X, y, nt, b), slab, spike, list(inc),
                                                 np.zeros(p), 1e-10)
    assert ok
    idx = np.flatnonzero(inc)
    P = siginv[np.ix_(idx, idx)]
    b = np.zeros(len(idx))
    for _ in range(60):
        full = np.zeros(p); full[idx] = b
        _, g, hh = O.binomial_logit_loglike_derivs(X, y, nt, full)
        step = np.linalg.solve(P - hh[np.ix_(idx, idx)], g[idx] - P @ (b - mu[idx]))
        b = b + step
        if np.max(np.abs(step)) < 1e-13:
            break
    np.testing.assert_allclose(beta[idx], b, rtol=1e-8, atol=1e-10)
    assert np.all(beta[~inc] == 0)
    full = np.zeros(p); full[idx] = b
    k = len(idx)
    logprior = -0.5 * k * np.log(2 * np.pi) + 0.5 * np.linalg.slogdet(P)[1] - 0.5 * (b - mu[idx]) @ P @ (b - mu[idx])
    assert value == pytest.approx(O.binomial_logit_loglike(X, y, nt, full) + logprior, rel=1e-11)
    # a far-away start still converges (the full Newton step overshoots into -inf likelihood values: step halving), an
    # empty model is declined as in the reference
    ok2, beta2, _ = h.find_posterior_mode_with(lambda b: O.binomial_logit_loglike_derivs(X, y, nt, b), slab, spike, list(inc),
                                               np.where(inc, 3.0, 0.0), 1e-10)
    assert ok2
    np.testing.assert_allclose(beta2[idx], b, rtol=1e-7, atol=1e-9)
    ok3, _, _ = h.find_posterior_mode_with(lambda b: O.binomial_logit_loglike_derivs(X, y, nt, b), slab, spike, [False] * p,
                                           np.zeros(p), 1e-10)
    assert not ok3


# ---------------------------------------------------------------- Poisson mixture table: counts off the shipped grid
def test_poisson_table_offgrid_entries_follow_the_reference_rule(golden):
    """NormalMixtureApproximationTable::approximate (NormalMixtureApproximation.cpp:472-532) restated on the host
    (boom_b200/host/mixture_table.cpp) against the reference's own answers from a fresh table
    (tests/golden/poisson_offgrid.json, oracle/ref_driver.cpp: golden_poisson_offgrid): the same branch is taken for every
    count; interpolated entries are the reference's numbers; directly fitted entries (the reference runs Powell, here
    Nelder-Mead from the rescaled lower neighbour) reach a Kullback-Leibler divergence at least as small."""
    import boom_b200
    boom_b200.load_poisson_mixture_table()
    h = boom_b200.host()
    nu_grid, off, w, mu, sig, cut = boom_b200.poisson_mixture_table_arrays()
    for r in golden("poisson_offgrid.json"):
        nu = int(r["nu"])
        m, s, wt, kl = h.poisson_mixture_approximate(nu)
        K = len(r["mu"])
        assert len(m) == K == int(r["k0"])            # the lower neighbour's number of components
        assert abs(wt.sum() - 1.0) < 1e-6 and np.all(s > 0) and np.all(np.diff(m) >= 0)
        i0, i1 = np.searchsorted(nu_grid, int(r["nu0"])), np.searchsorted(nu_grid, int(r["nu1"]))
        t = (nu - r["nu0"]) / (r["nu1"] - r["nu0"])
        interp_mu = ((1 - t) * mu[off[i0]:off[i0 + 1]] + t * mu[off[i1]:off[i1 + 1]]) if r["k0"] == r["k1"] else None
        interpolated_in_ref = interp_mu is not None and np.allclose(interp_mu, r["mu"], atol=1e-12)
        if interpolated_in_ref:
            np.testing.assert_allclose(m, r["mu"], rtol=0, atol=1e-13)
            np.testing.assert_allclose(s, r["sigma"], rtol=0, atol=1e-13)
            np.testing.assert_allclose(wt, r["weights"], rtol=0, atol=1e-13)
            assert kl == pytest.approx(r["kl"], abs=2e-8) and kl < 1e-5
        else:
            assert interp_mu is None or not np.allclose(m, interp_mu, atol=1e-9)     # fitted here as well
            assert kl <= max(r["kl"], 1e-7) * 1.05
    # the table now holds the added entries, and asking again returns them unchanged
    ser = h.poisson_mixture_table()
    again = h.poisson_mixture_approximate(1234)
    assert np.array_equal(h.poisson_mixture_table(), ser)
    assert again[3] < 1e-6
    boom_b200.load_poisson_mixture_table()    # back to the shipped table for the other tests


@pytest.mark.parametrize("fisher_yates", [False, True])
def test_sweep_on_the_active_set_view_is_the_sweep_on_the_full_matrix(fisher_yates):
    """StatView: a sweep over the inclusion indicators reads of X'WX only the columns of the included variables
    and the diagonal; with those columns, the diagonal and X'Wz held, and a further column fetched whenever the sweep ADDS a
    variable, the chain is bit for bit the one the full matrix gives (same random stream, same arithmetic) -- on the host
    alone, no device: the fetch callback reads the same matrix."""
    h = boom_b200.host()
    p = 40
    rng = np.random.default_rng(11)
    X = rng.normal(size=(400, p)); X[:, 0] = 1.0
    w = 0.2 + rng.random(400)
    beta = np.zeros(p); beta[[0, 3, 7, 19, 33]] = [0.5, 1.0, -1.0, 0.8, -0.7]
    z = X @ beta + rng.normal(size=400) / np.sqrt(w)
    xtx = (X.T * w) @ X
    xty = (X.T * w) @ z
    slab = boom_b200.MvnModel(np.zeros(p), np.eye(p), True)
    spike = boom_b200.VariableSelectionPrior(np.full(p, 0.1))
    start = [True] + [False] * (p - 1)
    full = h.spike_slab_chain(boom_b200.RNG(5), xtx, xty, slab, spike, start, 60, fisher_yates)
    active, fetched = h.spike_slab_sweep_active(boom_b200.RNG(5), xtx, xty, slab, spike, start, 60, fisher_yates)
    assert len(full) == len(active) == 60
    for (g0, b0), (g1, b1) in zip(full, active):
        assert list(g0) == list(g1)
        np.testing.assert_array_equal(np.asarray(b0), np.asarray(b1))
    assert fetched >= 4                                   # the four true variables outside the start model came in through fetches
    assert sum(full[-1][0]) >= 5
