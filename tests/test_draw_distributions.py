"""BASELINE.json north_star, 'Random draws': draws cannot be bit-identical to Bmath's RNG stream; they must pass
KS / chi-square tests against the reference's conditional distributions.

Reference side: tests/golden/ref_*_stats.json, written by the compiled, unmodified reference
(oracle/ref_driver.cpp golden_draw_stats): for a grid of (eta, y) the moments, the histogram of mixture
indicators and 19 quantiles of 2-4 x 10^5 draws of BinomialLogitCltDataImputer::impute / PoissonDataImputer::impute.
Tested side: the oracle (CPU, -m "not gpu") and the CUDA path through the C ABI (-m gpu), 2 x 10^5 draws each:
  * logit small-sample branch: KS against the exact truncated-logistic law, two-sample
    chi-square of the indicator histogram against the reference's, chi-square on the reference's quantile bins;
  * logit CLT branch and Poisson: chi-square on the reference's quantile bins (two-sample, variance inflated by
    1 + M / N_ref for the sampling error of the reference quantiles), z-tests of the means.
All tests use a 1e-6 significance level: a correct sampler fails one of the ~100 comparisons once in 10^4 runs."""
import numpy as np
import pytest
from scipy import stats

from oracle import oracle as O

M = 200_000
ALPHA = 1e-6


def _quantile_bin_chi2(x, ref_quantiles, n_ref):
    edges = np.asarray(ref_quantiles)
    counts = np.bincount(np.searchsorted(edges, x, side="right"), minlength=len(edges) + 1)
    expect = len(x) / (len(edges) + 1.0)
    chi2 = ((counts - expect) ** 2 / expect).sum() / (1.0 + len(x) / n_ref)
    assert chi2 < stats.chi2.ppf(1 - ALPHA, len(edges)), ("quantile-bin chi-square", chi2, counts)


def _mean_ztest(x, ref_mean, ref_var, n_ref, what):
    se = np.sqrt(x.var() / len(x) + ref_var / n_ref)
    assert abs(x.mean() - ref_mean) < stats.norm.ppf(1 - ALPHA / 2) * se + 1e-12, (what, x.mean(), ref_mean, se)


def _logistic_cdf(t):
    return 1.0 / (1.0 + np.exp(-t))


def check_logit_small(draw, golden):
    """draw(eta, y) -> (z, info) arrays of M draws with n_i = 1."""
    mix = O.logit_mixture()
    inv = 1.0 / mix.sigma ** 2
    for c in golden("ref_logit_small_stats.json"):
        eta, y, n_ref = c["eta"], int(c["y"]), int(c["N"])
        z, info = draw(eta, y)
        f0 = _logistic_cdf(-eta)
        if y == 1:
            assert z.min() > 0
            cdf = lambda t: (_logistic_cdf(t - eta) - f0) / (1.0 - f0)       # noqa: E731
        else:
            assert z.max() < 0
            cdf = lambda t: _logistic_cdf(t - eta) / f0                      # noqa: E731
        assert stats.kstest(z, cdf).pvalue > ALPHA, ("KS vs truncated logistic", eta, y)
        _quantile_bin_chi2(z, c["z_quantiles"], n_ref)
        _mean_ztest(z, c["z_mean"], c["z_var"], n_ref, "z mean")
        _mean_ztest(info, c["info_mean"], c["info_var"], n_ref, "information mean")
        k = np.argmin(np.abs(info[:, None] - inv[None, :]), axis=1)
        table = np.array([np.bincount(k, minlength=9), np.array(c["kcount"], dtype=np.int64)])
        assert stats.chi2_contingency(table).pvalue > ALPHA, ("indicator histogram", eta, y, table)


def check_logit_clt(draw, golden):
    """draw(ntrials, y, eta) -> (sum, info)."""
    for c in golden("ref_logit_clt_stats.json"):
        s, info = draw(c["ntrials"], c["y"], c["eta"])
        n_ref = int(c["N"])
        _quantile_bin_chi2(s, c["sum_quantiles"], n_ref)
        _mean_ztest(s, c["sum_mean"], c["sum_var"], n_ref, "sum mean")
        _mean_ztest(info, c["info_mean"], c["info_var"], n_ref, "information mean")
        # the information takes few distinct values at small n: compare its variance instead of binning it
        assert info.var() == pytest.approx(c["info_var"], rel=0.05)


def check_poisson(draw, golden):
    """draw(y, exposure, eta) -> out6 (M x 6) = z_int, mu_int, w_int, z_ext, mu_ext, w_ext."""
    for c in golden("ref_poisson_stats.json"):
        y, n_ref = int(c["y"]), int(c["N"])
        o = draw(y, c["exposure"], c["eta"])
        _quantile_bin_chi2(o[:, 3], c["zext_quantiles"], n_ref)
        _mean_ztest(o[:, 3], c["zext_mean"], c["zext_var"], n_ref, "z_ext mean")
        assert o[:, 5].mean() == pytest.approx(c["wext_mean"], rel=0.02)
        assert ((o[:, 3] - o[:, 4]) * o[:, 5]).mean() == pytest.approx(c["rwext_mean"], rel=0.03, abs=0.02)
        if y > 0:
            _quantile_bin_chi2(o[:, 0], c["zint_quantiles"], n_ref)
            _mean_ztest(o[:, 0], c["zint_mean"], c["zint_var"], n_ref, "z_int mean")
            assert o[:, 2].mean() == pytest.approx(c["wint_mean"], rel=0.02)
            assert ((o[:, 0] - o[:, 1]) * o[:, 2]).mean() == pytest.approx(c["rwint_mean"], rel=0.03, abs=0.02 * max(1, y))


# ---------------------------------------------------------------------------------------- oracle (CPU)
ONES = np.ones((M, 1))


def test_oracle_logit_small_sample_draws(golden):
    mix = O.logit_mixture()

    def draw(eta, y):
        s, w = O.logit_draw(ONES, np.full(M, float(y)), np.ones(M), [eta], 10, mix, 1234, int(10 * eta) + 50 + y)
        return s / w, w
    check_logit_small(draw, golden)


def test_oracle_logit_clt_draws(golden):
    mix = O.logit_mixture()

    def draw(nt, y, eta):
        return O.logit_draw(ONES, np.full(M, float(y)), np.full(M, float(nt)), [eta], 10, mix, 4321, int(nt))
    check_logit_clt(draw, golden)


def test_oracle_poisson_draws(golden):
    tab = O.poisson_table()

    def draw(y, ex, eta):
        return O.poisson_draw(ONES, np.full(M, y, dtype=np.int64), np.full(M, ex), [eta], tab, 99, y)[0]
    check_poisson(draw, golden)


# ---------------------------------------------------------------------------------------- CUDA path (C ABI)
@pytest.mark.gpu
def test_device_logit_small_sample_draws(golden):
    from tests.helpers import logit_ctx

    def draw(eta, y):
        ctx, _ = logit_ctx(ONES, np.full(M, float(y)), np.ones(M))
        s, w = ctx.logit_draw([eta], 10, seed=777, iteration=int(10 * eta) + 50 + y)
        ctx.close()
        return s / w, w
    check_logit_small(draw, golden)


@pytest.mark.gpu
def test_device_logit_clt_draws(golden):
    from tests.helpers import logit_ctx

    def draw(nt, y, eta):
        ctx, _ = logit_ctx(ONES, np.full(M, float(y)), np.full(M, float(nt)))
        out = ctx.logit_draw([eta], 10, seed=778, iteration=int(nt))
        ctx.close()
        return out
    check_logit_clt(draw, golden)


@pytest.mark.gpu
def test_device_poisson_draws(golden):
    from tests.helpers import poisson_ctx

    def draw(y, ex, eta):
        ctx, _ = poisson_ctx(ONES, np.full(M, y, dtype=np.int64), np.full(M, ex))
        out, _ = ctx.poisson_draw([eta], seed=779, iteration=y)
        ctx.close()
        return out
    check_poisson(draw, golden)


@pytest.mark.gpu
def test_device_probit_draws_follow_the_truncated_normal_law():
    """The probit sibling's latent draws (one trial per row) against the exact N(eta, 1) law truncated by the response:
    KS at 1e-6 per response class and linear-predictor bin (BinomialProbitDataImputer.cpp:55-66 draws the same law by rejection)."""
    from scipy import stats
    from oracle import oracle as O
    from tests.helpers import logit_ctx
    n, p = 200_000, 3
    X = O.synth_x(n, p, seed=91)
    X[:, 1] = np.repeat([-2.5, -0.7, 0.4, 3.0], n // 4)   # four linear predictors
    X[:, 2] = 0.0
    beta = np.array([0.2, 1.0, 0.0])
    rng = np.random.default_rng(5)
    y = (rng.random(n) < 0.5).astype(float)
    nt = np.ones(n)
    ctx, _ = logit_ctx(X, y, nt)
    z = ctx.probit_draw(beta, 10, seed=17, iteration=1)
    ctx.close()
    eta = X @ beta
    for e in np.unique(eta):
        for resp in (0.0, 1.0):
            sel = (eta == e) & (y == resp)
            lo, hi = (0.0, np.inf) if resp else (-np.inf, 0.0)
            law = stats.truncnorm(lo - e, hi - e, loc=e, scale=1.0)
            assert np.all((z[sel] > lo) & (z[sel] < hi))
            assert stats.kstest(z[sel], law.cdf).pvalue > 1e-6, (e, resp)
