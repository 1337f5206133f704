"""Pins the C oracle (oracle/auxmix_oracle.c) to golden vectors written by the
compiled, unmodified reference (oracle/ref_driver.cpp -> tests/golden/*.json).

Deterministic functions must agree to rounding; the reference's own draws are
recovered by re-seeding its RNG and reading the uniform it consumed, so the
mixture indicator must agree EXACTLY.
"""
import numpy as np
import pytest

from oracle import oracle as O


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert O.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_uniforms_open_interval_and_distinct():
    us = np.array([O.uniform_pair(7, it, row, slot) for it in range(3) for row in range(50) for slot in range(3)])
    assert us.min() > 0 and us.max() < 1
    assert len(np.unique(us)) == us.size
    assert abs(us.mean() - 0.5) < 0.05


def test_logit_mixture_matches_reference_constants(golden):
    g = golden("logit_mixture.json")
    mix = O.logit_mixture()
    assert mix.K == 9
    np.testing.assert_array_equal(mix.log_weights, np.array(g["log_weights"]))  # glibc log both sides
    assert abs(sum(g["weights"]) - 1) < 1e-12
    assert abs(np.dot(g["weights"], np.square(g["sigma"])) - np.pi ** 2 / 3) < 1e-4


def test_unmix_indicator_exact(golden):
    mix = O.logit_mixture()
    cases = golden("unmix_logit.json")
    ks = [O.unmix(mix, c["u"], c["U"]) for c in cases]
    assert ks == [int(c["k"]) for c in cases]
    assert len(set(ks)) >= 7  # the grid exercises most components


def test_unmix_posterior_is_normalised():
    mix = O.logit_mixture()
    k, post = O.unmix(mix, 0.3, 0.5, want_post=True)
    assert abs(post.sum() - 1) < 1e-14 and 0 <= k < 9


def test_rtrun_logit_matches_reference(golden):
    for c in golden("rtrun_logit.json"):
        z = O.rtrun_logit(c["eta"], c["above"] > 0, c["U"])
        assert z == pytest.approx(c["z"], rel=1e-13, abs=1e-13)
        assert (z > 0) == (c["above"] > 0)


def test_logit_impute_small_sample_matches_reference(golden):
    """sum/info of BinomialLogitCltDataImputer::impute rebuilt from the reference's own uniforms."""
    mix = O.logit_mixture()
    for c in golden("logit_impute_small.json"):
        s = w = 0.0
        for i in range(int(c["ntrials"])):
            z = O.rtrun_logit(c["eta"], i < c["y"], c["U"][2 * i])
            k = O.unmix(mix, z - c["eta"], c["U"][2 * i + 1])
            cw = 1.0 / (mix.sigma[k] * mix.sigma[k])
            w += cw
            s += z * cw
        assert w == pytest.approx(c["info"], rel=1e-14)
        assert s == pytest.approx(c["sum"], rel=1e-12, abs=1e-12)


def test_trun_norm_moments_match_reference(golden):
    for c in golden("trun_norm_moments.json"):
        m, v = O.trun_norm_moments(c["mu"], c["sigma"], 0.0, c["positive"] > 0)
        assert m == pytest.approx(c["mean"], rel=1e-9, abs=1e-12)
        assert v == pytest.approx(c["variance"], rel=1e-6, abs=1e-12)


def test_sufficient_statistics_bit_exact(golden):
    g = golden("suf.json")
    n, p = int(g["n"]), int(g["p"])
    X = np.array(g["X"]).reshape(n, p)
    xtx, xty = O.accumulate(X, g["weight"], g["weighted_value"])
    ref = np.array(g["xtx_colmajor"]).reshape(p, p).T
    np.testing.assert_array_equal(xtx, ref)          # same operations in the same order
    np.testing.assert_array_equal(xty, np.array(g["xty"]))
    assert int(g["sample_size"]) == n


def test_weighted_reg_suf_bit_exact(golden):
    import ctypes as C
    g = golden("suf.json")
    n, p = int(g["n"]), int(g["p"])
    X = np.array(g["X"]).reshape(n, p)
    xtwx = np.zeros((p, p)); xtwy = np.zeros(p); sc = np.zeros(4)
    L = O.lib()
    for i in range(n):
        x = np.ascontiguousarray(X[i])
        L.bo_weighted_reg_suf_add(C.c_int(p), O._dp(xtwx), O._dp(xtwy), O._dp(sc), O._dp(x),
                                  C.c_double(g["y"][i]), C.c_double(g["weight"][i]))
    L.bo_reflect(C.c_int(p), O._dp(xtwx))
    np.testing.assert_array_equal(xtwx.T, np.array(g["w_xtwx_colmajor"]).reshape(p, p).T)
    np.testing.assert_array_equal(xtwy, np.array(g["w_xtwy"]))
    np.testing.assert_allclose(sc, np.array(g["w_scalars"]), rtol=1e-15)


def test_poisson_table_structure(golden):
    g = golden("poisson_mixture_table.json")
    tab = O.poisson_table()
    assert tab.gaussian_cutoff == 30000 and int(g["smallest_index"]) == 1
    assert int(g["grid_serialized_length"]) == 2894          # SURVEY.md App. A.3 [probe]
    assert set(range(1, 301)) <= set(tab.nu.tolist())
    for nu, K in [(1, 10), (19, 10), (20, 4), (49, 4), (50, 3), (490, 3), (1000, 2)]:
        assert tab.entry(nu).K == K
    e = tab.entry(7)
    assert abs(e.weights.sum() - 1) < 1e-6 and np.all(np.diff(e.mu) >= 0)  # stored sorted by mu


def test_unmix_poisson_matches_reference(golden):
    tab = O.poisson_table()
    for c in golden("unmix_poisson.json"):
        nu = int(c["nu"])
        if nu >= tab.gaussian_cutoff:
            assert c["mu"] == pytest.approx(-np.log(nu), rel=1e-15) and c["sigsq"] == pytest.approx(1.0 / nu)
            continue
        mix = tab.entry(nu)
        k = O.unmix(mix, c["resid"], c["U"])
        assert mix.mu[k] == c["mu"]
        assert mix.sigma[k] ** 2 == pytest.approx(c["sigsq"], rel=1e-15)


def test_loglike_matches_reference(golden):
    g = golden("loglike.json")
    n, p = int(g["n"]), int(g["p"])
    X = np.array(g["X"]).reshape(n, p)
    ll = O.binomial_logit_loglike(X, g["y"], g["ntrials"], g["beta"])
    assert ll == pytest.approx(g["binomial_loglike"], rel=1e-13)
    Xp = np.array(g["poisson_X"]).reshape(n, p)
    llp = O.poisson_loglike(Xp, g["poisson_y"], g["poisson_exposure"], g["poisson_beta"])
    assert llp == pytest.approx(g["poisson_loglike"], rel=1e-13)
    for c in g["dbinom"]:
        assert O.lib().bo_dbinom_log(c["x"], c["n"], c["p"]) == pytest.approx(c["logd"], rel=1e-11, abs=1e-13)


def test_binomial_from_uniform_is_the_inverse_cdf():
    from scipy import stats
    rng = np.random.default_rng(5)
    for n, p in [(1, 0.3), (7, 0.5), (40, 0.07), (1000, 0.48), (5000, 0.999), (25, 1e-9)]:
        u = rng.random(4000)
        k = np.array([O.lib().bo_binomial_from_uniform(n, p, float(v)) for v in u])
        assert k.min() >= 0 and k.max() <= n
        # chop-down from the mode is a measure preserving rearrangement of inversion: compare frequencies
        vals, cnt = np.unique(k, return_counts=True)
        expect = stats.binom.pmf(vals, n, p) * len(u)
        keep = expect > 5
        if keep.sum() > 1:
            chi2 = ((cnt[keep] - expect[keep]) ** 2 / expect[keep]).sum()
            assert chi2 < stats.chi2.ppf(1 - 1e-6, keep.sum())


def test_loglike_derivatives_match_reference(golden):
    """gradient / Hessian of the log likelihood as the reference computes them (incl. the log_alpha offset of
    BinomialLogitModel.cpp:168 and the exposure-free Hessian weight of PoissonRegressionModel.cpp:84)."""
    g = golden("loglike.json")
    n, p = int(g["n"]), int(g["p"])
    X = np.array(g["X"]).reshape(n, p)
    ll, gr, h = O.binomial_logit_loglike_derivs(X, g["y"], g["ntrials"], g["beta"])
    assert ll == pytest.approx(g["binomial_loglike_d"], rel=1e-13)
    np.testing.assert_allclose(gr, g["binomial_gradient"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(h, np.array(g["binomial_hessian"]).reshape(p, p), rtol=1e-12, atol=1e-12)
    ll, gr, _ = O.binomial_logit_loglike_derivs(X, g["y"], g["ntrials"], g["beta"], g["binomial_log_alpha"])
    assert ll == pytest.approx(g["binomial_loglike_alpha"], rel=1e-13)
    np.testing.assert_allclose(gr, g["binomial_gradient_alpha"], rtol=1e-12, atol=1e-12)
    Xp = np.array(g["poisson_X"]).reshape(n, p)
    ll, gr, h = O.poisson_loglike_derivs(Xp, g["poisson_y"], g["poisson_exposure"], g["poisson_beta"])
    assert ll == pytest.approx(g["poisson_loglike_d"], rel=1e-13)
    np.testing.assert_allclose(gr, g["poisson_gradient"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(h, np.array(g["poisson_hessian"]).reshape(p, p), rtol=1e-12, atol=1e-12)
