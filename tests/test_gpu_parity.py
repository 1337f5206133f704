"""GPU parity tests proper: the CUDA path, driven through the C ABI (include/boomgpu.h),
against the C oracle on the same seeded inputs.  Tolerances (BASELINE.json north_star):
deterministic statistics 1e-12 (normwise for X'WX), mixture-indicator counting exact.
"""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import logit_ctx, normwise_err, poisson_ctx, vec_err

pytestmark = pytest.mark.gpu

TOL = 1e-12


def _latents(n, seed):
    rng = np.random.default_rng(seed)
    return 0.05 + 1.5 * rng.random(n), 6.0 * (rng.random(n) - 0.5)


# ---------------------------------------------------------------- deterministic accumulation
@pytest.mark.parametrize("n,p,path", [
    (1, 3, 0), (37, 5, 0), (1000, 16, 0), (5003, 20, 0), (777, 33, 0), (2049, 50, 0), (1500, 64, 0), (300, 7, 0),
    (1000, 16, 2), (37, 5, 2), (4097, 50, 2), (3000, 100, 0), (2500, 128, 0), (2100, 130, 0), (5000, 260, 0),
    (6000, 500, 0), (1, 200, 0), (15, 129, 0),
])
def test_accumulate_matches_oracle(n, p, path):
    X = O.synth_x(n, p, seed=100 + p)
    w, s = _latents(n, p)
    ctx, _ = logit_ctx(X, np.zeros(n), np.ones(n), path=path)
    xtx, xty = ctx.accumulate(w, s)
    ref_xtx, ref_xty = O.accumulate(X, w, s)
    assert normwise_err(xtx, ref_xtx) < TOL
    assert vec_err(xty, ref_xty) < TOL
    np.testing.assert_array_equal(xtx, xtx.T)   # both triangles filled, exactly symmetric
    ctx.close()


@pytest.mark.parametrize("n,p", [(3000, 65), (3000, 72), (2500, 97), (2500, 120), (2000, 127), (2000, 192), (2000, 248), (1000, 385),
                                 (1500, 456)])
def test_ragged_diagonal_region_kernel(n, p):
    """The ragged last diagonal region (p not a multiple of 128; the whole matrix for 64 < p < 128) runs in syrk_rdiag_kernel,
    strip form over A = ceil((p mod 128) / 8) atom columns (A = 9, 9, 13, 15, 16 -> whole, 8, 15, 1, 9 here); the 128 x 128
    unit form of the main kernel stays available (option syrk_rdiag = 0).  Both against the oracle."""
    X = O.synth_x(n, p, seed=500 + p)
    w, s = _latents(n, p)
    ctx, _ = logit_ctx(X, np.zeros(n), np.ones(n))
    ref_xtx, ref_xty = O.accumulate(X, w, s)
    for rdiag in (1, 0):
        ctx.set_option("syrk_rdiag", rdiag)
        xtx, xty = ctx.accumulate(w, s)
        assert normwise_err(xtx, ref_xtx) < TOL, rdiag
        assert vec_err(xty, ref_xty) < TOL, rdiag
        np.testing.assert_array_equal(xtx, xtx.T)
    ctx.close()


def test_accumulate_golden_fixture(golden):
    g = golden("suf.json")
    n, p = int(g["n"]), int(g["p"])
    X = np.array(g["X"]).reshape(n, p)
    ctx, _ = logit_ctx(X, np.zeros(n), np.ones(n))
    xtx, xty = ctx.accumulate(g["weight"], g["weighted_value"])
    ref = np.array(g["xtx_colmajor"]).reshape(p, p).T
    assert normwise_err(xtx, ref) < TOL and vec_err(xty, np.array(g["xty"])) < TOL
    ctx.close()


# ---------------------------------------------------------------- latent draws, value by value
@pytest.mark.parametrize("n,p,max_trials,path", [(4000, 6, 1, 0), (3000, 20, 4, 0), (2000, 8, 40, 0), (1500, 70, 1, 0),
                                                 (1200, 90, 30, 0)])
def test_logit_draws_match_oracle(n, p, max_trials, path):
    X, y, nt, beta_true = O.synth_binomial(n, p, 3, seed=11 + p, max_trials=max_trials)
    beta = beta_true * 0.9 + 0.01
    ctx, mix = logit_ctx(X, y, nt, path=path)
    s, w = ctx.logit_draw(beta, 10, seed=99, iteration=3)
    rs, rw = O.logit_draw(X, y, nt, beta, 10, mix, 99, 3)
    # information = sum of 1/sigma_k^2 over the drawn indicators: any indicator mismatch shows here
    np.testing.assert_allclose(w, rw, rtol=1e-13)
    np.testing.assert_allclose(s, rs, rtol=1e-9, atol=1e-9)
    ctx.close()


def test_logit_indicator_counts_exact():
    """Histogram of the mixture indicators (decoded from info = 1/sigma_k^2) is bit exact vs the oracle."""
    n, p = 20000, 4
    X, y, nt, beta = O.synth_binomial(n, p, 2, seed=5)
    ctx, mix = logit_ctx(X, y, nt)
    _, w = ctx.logit_draw(beta, 10, seed=1, iteration=0)
    _, _, _, kc = O.logit_step(X, y, nt, beta, 10, mix, 1, 0)
    inv = 1.0 / (mix.sigma * mix.sigma)
    k_dev = np.argmin(np.abs(w[:, None] - inv[None, :]), axis=1)
    assert np.array_equal(np.bincount(k_dev, minlength=9), kc)
    ctx.close()


def test_poisson_draws_match_oracle():
    n, p = 5000, 6
    X, y, ex, beta_true = O.synth_poisson(n, p, 3, seed=21)
    y[:7] = [0, 1, 2, 40, 120, 299, 35000]      # table edges and the gaussian cutoff
    ex[:4] = [0.5, 2.0, 1.5, 3.0]
    ctx, tab = poisson_ctx(X, y, ex)
    out, k2 = ctx.poisson_draw(beta_true, seed=8, iteration=2)
    ref, rk2 = O.poisson_draw(X, y, ex, beta_true, tab, 8, 2)
    assert np.array_equal(k2, rk2)
    np.testing.assert_allclose(out, ref, rtol=1e-10, atol=1e-12)
    ctx.close()


# ---------------------------------------------------------------- whole steps
@pytest.mark.parametrize("n,p,max_trials", [(3001, 20, 1), (2000, 16, 25), (2500, 50, 1), (2200, 100, 1), (2000, 200, 12)])
def test_logit_step_matches_oracle(n, p, max_trials):
    X, y, nt, beta = O.synth_binomial(n, p, 5, seed=31 + p, max_trials=max_trials)
    ctx, mix = logit_ctx(X, y, nt)
    xtx, xty, ss = ctx.logit_step(beta, 10, seed=77, iteration=5)
    rxtx, rxty, rss, _ = O.logit_step(X, y, nt, beta, 10, mix, 77, 5)
    assert ss == rss == n
    assert normwise_err(xtx, rxtx) < 1e-11     # draws differ by libm-vs-CUDA ulps before the sum
    assert vec_err(xty, rxty) < 1e-10
    ctx.close()


@pytest.mark.parametrize("n,p", [(3000, 5), (2500, 50), (2000, 80)])
def test_poisson_step_matches_oracle(n, p):
    X, y, ex, beta = O.synth_poisson(n, p, 3, seed=41 + p)
    ctx, tab = poisson_ctx(X, y, ex)
    xtx, xty, sc = ctx.poisson_step(beta, seed=13, iteration=1)
    rxtx, rxty, rsc = O.poisson_step(X, y, ex, beta, tab, 13, 1)
    assert sc[0] == rsc[0] == n + np.count_nonzero(y)
    assert normwise_err(xtx, rxtx) < 1e-11
    assert vec_err(xty, rxty) < 1e-10
    np.testing.assert_allclose(sc, rsc, rtol=1e-10)
    ctx.close()


def test_steps_are_reproducible_and_iteration_keyed():
    X, y, nt, beta = O.synth_binomial(5000, 12, 3, seed=3)
    ctx, _ = logit_ctx(X, y, nt)
    a = ctx.logit_step(beta, 10, 5, 9)
    b = ctx.logit_step(beta, 10, 5, 9)
    c = ctx.logit_step(beta, 10, 5, 10)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])
    assert np.max(np.abs(a[1] - c[1])) > 0
    ctx.close()


@pytest.mark.parametrize("p", [10, 150])
def test_sharding_invariance(p):
    """Rows split over two contexts with their global row offsets give the same statistics as one context."""
    n = 4000
    X, y, nt, beta = O.synth_binomial(n, p, 3, seed=17)
    full, _ = logit_ctx(X, y, nt)
    fx, fy, _ = full.logit_step(beta, 10, 123, 4)
    cut = 1777
    a, _ = logit_ctx(X[:cut], y[:cut], nt[:cut])
    b, _ = logit_ctx(X[cut:], y[cut:], nt[cut:])
    b.set_row_offset(cut)
    ax, ay, an = a.logit_step(beta, 10, 123, 4)
    bx, by, bn = b.logit_step(beta, 10, 123, 4)
    assert an + bn == n
    assert normwise_err(ax + bx, fx) < 1e-13 and vec_err(ay + by, fy) < 1e-12
    for c in (full, a, b):
        c.close()


# ---------------------------------------------------------------- log likelihood
def test_loglike_matches_oracle_and_golden(golden):
    g = golden("loglike.json")
    n, p = int(g["n"]), int(g["p"])
    X = np.array(g["X"]).reshape(n, p)
    ctx, _ = logit_ctx(X, g["y"], g["ntrials"])
    assert ctx.binomial_loglike(g["beta"]) == pytest.approx(g["binomial_loglike"], rel=1e-12)
    ctx.close()
    Xp = np.array(g["poisson_X"]).reshape(n, p)
    ctx, _ = poisson_ctx(Xp, g["poisson_y"], g["poisson_exposure"])
    assert ctx.poisson_loglike(g["poisson_beta"]) == pytest.approx(g["poisson_loglike"], rel=1e-12)
    ctx.close()
    X, y, nt, beta = O.synth_binomial(20000, 30, 4, seed=2, max_trials=15)
    ctx, _ = logit_ctx(X, y, nt)
    assert ctx.binomial_loglike(beta) == pytest.approx(O.binomial_logit_loglike(X, y, nt, beta), rel=1e-12)
    ctx.close()


# ---------------------------------------------------------------- errors
def test_invalid_data_is_reported():
    import boom_b200
    X, y, nt, beta = O.synth_binomial(500, 4, 2, seed=1)
    y[100] = 3.0  # successes > trials
    ctx, _ = logit_ctx(X, y, nt)
    with pytest.raises(boom_b200.BoomGpuError, match="invalid observation"):
        ctx.logit_step(beta, 10, 1, 0)
    ctx.close()
    Xp, yp, ex, bp = O.synth_poisson(500, 4, 2, seed=1)
    yp[3] = 12345  # off-grid count that the fixture table does not hold
    ctx, _ = poisson_ctx(Xp, yp, ex)
    with pytest.raises(boom_b200.BoomGpuError, match="missing from the Poisson mixture table"):
        ctx.poisson_step(bp, 1, 0)
    ctx.close()


def test_state_errors():
    import boom_b200
    ctx = boom_b200.Context(0)
    with pytest.raises(boom_b200.BoomGpuError, match="no binomial data"):
        ctx.p = 3
        ctx.logit_step(np.zeros(3), 10, 1, 0)
    ctx.close()


# ---------------------------------------------------------------- certified FP32 selection of the mixture indicator
def test_logit_indicators_exact_at_scale():
    """draws.cuh selects the indicator in FP32 when the decision is provably the FP64 one and falls back to the
    FP64 rule otherwise: over 400k draws (several hundred of them on the fallback) EVERY indicator equals the oracle's."""
    n, p = 400_000, 3
    X, y, nt, beta = O.synth_binomial(n, p, 2, seed=77)
    ctx, mix = logit_ctx(X, y, nt)
    _, w = ctx.logit_draw(beta, 10, seed=5, iteration=11)
    _, rw = O.logit_draw(X, y, nt, beta, 10, mix, 5, 11)
    np.testing.assert_array_equal(w, rw)     # info = 1/sigma_k^2 of the drawn k: bitwise equal or a different k
    ctx.close()


def test_poisson_indicators_exact_at_scale():
    n, p = 200_000, 3
    X, y, ex, beta = O.synth_poisson(n, p, 2, seed=78)
    y[:20000] = np.random.default_rng(1).integers(0, 300, size=20000)    # every table size K = 10, 4, 3 and off-grid counts
    ctx, tab = poisson_ctx(X, y, ex)
    _, k2 = ctx.poisson_draw(beta, seed=6, iteration=4)
    _, rk2 = O.poisson_draw(X, y, ex, beta, tab, 6, 4)
    assert np.array_equal(k2, rk2)
    ctx.close()


@pytest.mark.parametrize("n,p,kind", [(5003, 20, "logit"), (40_000, 16, "logit"), (3000, 64, "logit"), (7001, 50, "poisson"),
                                      (33, 1, "logit"), (100_000, 7, "logit"), (60_001, 40, "poisson"), (50_003, 48, "logit")])
def test_small_p_variants_agree_with_oracle(n, p, kind):
    """p <= 64: the TMA-fed warp-autonomous kernel (0: automatic choice between its 8-warp form and, for wide tiles, the
    12-warp form with the accumulators parked in tensor memory; 2: never parked; 4: parked wherever 32 < p <= 64), the cp.async
    kernel (1: X that TMA cannot describe) and the warp-specialised kernel (3)."""
    for variant in (0, 1, 2, 3, 4):
        if kind == "logit":
            X, y, nt, beta = O.synth_binomial(n, p, min(3, p - 1), seed=50 + p, max_trials=2)
            ctx, mix = logit_ctx(X, y, nt)
            ctx.set_option("small_variant", variant)
            xtx, xty, ss = ctx.logit_step(beta, 10, seed=3, iteration=2)
            rxtx, rxty, rss, _ = O.logit_step(X, y, nt, beta, 10, mix, 3, 2)
            assert ss == rss == n
        else:
            X, y, ex, beta = O.synth_poisson(n, p, 3, seed=60 + p)
            ctx, tab = poisson_ctx(X, y, ex)
            ctx.set_option("small_variant", variant)
            xtx, xty, sc = ctx.poisson_step(beta, seed=3, iteration=2)
            rxtx, rxty, rsc = O.poisson_step(X, y, ex, beta, tab, 3, 2)
            np.testing.assert_allclose(sc, rsc, rtol=1e-10)
        assert normwise_err(xtx, rxtx) < 1e-11, ("variant", variant)
        assert vec_err(xty, rxty) < 1e-10, ("variant", variant)
        np.testing.assert_array_equal(xtx, xtx.T)
        ctx.close()


# ---------------------------------------------------------------- log likelihood with gradient and Hessian
@pytest.mark.parametrize("n,p", [(3000, 6), (2500, 50), (2000, 130)])
def test_loglike_derivatives_match_oracle(n, p):
    X, y, nt, beta = O.synth_binomial(n, p, 3, seed=70 + p, max_trials=6)
    ctx, _ = logit_ctx(X, y, nt)
    for log_alpha in (0.0, np.log(0.3)):
        ll, g, h = ctx.binomial_loglike_derivs(beta * 0.7, log_alpha)
        rll, rg, rh = O.binomial_logit_loglike_derivs(X, y, nt, beta * 0.7, log_alpha)
        assert ll == pytest.approx(rll, rel=1e-12)
        assert vec_err(g, rg) < 1e-12
        assert normwise_err(-h, -rh) < 1e-12
    ctx.close()
    Xp, yp, ex, bp = O.synth_poisson(n, p, 3, seed=80 + p)
    ex = 0.5 + 0.1 * (np.arange(n) % 7)
    ctx, _ = poisson_ctx(Xp, yp, ex)
    ll, g, h = ctx.poisson_loglike_derivs(bp * 0.9)
    rll, rg, rh = O.poisson_loglike_derivs(Xp, yp, ex, bp * 0.9)
    assert ll == pytest.approx(rll, rel=1e-12)
    assert vec_err(g, rg) < 1e-12
    assert normwise_err(-h, -rh) < 1e-12
    ctx.close()


def test_loglike_derivatives_golden(golden):
    g = golden("loglike.json")
    n, p = int(g["n"]), int(g["p"])
    ctx, _ = logit_ctx(np.array(g["X"]).reshape(n, p), g["y"], g["ntrials"])
    ll, gr, h = ctx.binomial_loglike_derivs(g["beta"])
    assert ll == pytest.approx(g["binomial_loglike_d"], rel=1e-12)
    np.testing.assert_allclose(gr, g["binomial_gradient"], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(h, np.array(g["binomial_hessian"]).reshape(p, p), rtol=1e-11, atol=1e-11)
    ll, gr, _ = ctx.binomial_loglike_derivs(g["beta"], g["binomial_log_alpha"])
    assert ll == pytest.approx(g["binomial_loglike_alpha"], rel=1e-12)
    np.testing.assert_allclose(gr, g["binomial_gradient_alpha"], rtol=1e-11, atol=1e-11)
    ctx.close()
    ctx, _ = poisson_ctx(np.array(g["poisson_X"]).reshape(n, p), g["poisson_y"], g["poisson_exposure"])
    ll, gr, h = ctx.poisson_loglike_derivs(g["poisson_beta"])
    assert ll == pytest.approx(g["poisson_loglike_d"], rel=1e-12)
    np.testing.assert_allclose(gr, g["poisson_gradient"], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(h, np.array(g["poisson_hessian"]).reshape(p, p), rtol=1e-11, atol=1e-11)
    ctx.close()


def test_matrix_lands_directly_in_page_locked_host_memory():
    """Large p: when the caller's X'WX buffer is page-locked the matrix is copied device->host straight into it
    (no staging copy); the result is bit-identical to the staged path."""
    import boom_b200
    n, p = 3000, 400          # p * p * 8 = 1.28 MB >= the 1 MB threshold
    X, y, nt, beta = O.synth_binomial(n, p, 5, seed=90)
    ctx, _ = logit_ctx(X, y, nt)
    ref_xtx, ref_xty, ref_ss = ctx.logit_step(beta, 10, seed=4, iteration=1)
    out = np.full((p, p), np.nan)
    boom_b200.Context.pin_host(out)
    try:
        xtx, xty, ss = ctx.logit_step(beta, 10, seed=4, iteration=1, xtx=out)
    finally:
        boom_b200.Context.unpin_host(out)
    assert xtx is out and ss == ref_ss
    np.testing.assert_array_equal(out, ref_xtx)
    np.testing.assert_array_equal(xty, ref_xty)
    ctx.close()


# ---------------------------------------------------------------- probit sibling
@pytest.mark.parametrize("n,p,max_trials,path", [(4000, 6, 1, 0), (3000, 20, 25, 0), (2500, 70, 1, 0), (2000, 130, 40, 0), (1500, 8, 3, 2)])
def test_probit_step_matches_oracle(n, p, max_trials, path):
    """BinomialProbitSpikeSlabSampler::impute_latent_data / refresh_xtx (BinomialProbitSpikeSlabSampler.cpp:58-83): per-row sums
    of the latent normals value by value on the shared Philox stream, X'NX and X'z to 1e-12 / 1e-10; the X'z-only pass
    (xtx not requested) gives the same X'z."""
    X, y, nt, beta_true = O.synth_binomial(n, p, 3, seed=400 + p, max_trials=max_trials)
    beta = beta_true * 0.6
    ctx, _ = logit_ctx(X, y, nt, path=path)
    draws = ctx.probit_draw(beta, 10, seed=5, iteration=2)
    rxtx, rxtz, rdraws = O.probit_step(X, y, nt, beta, 10, 5, 2)
    np.testing.assert_allclose(draws, rdraws, rtol=1e-9, atol=1e-9)
    xtx, xtz, ss = ctx.probit_step(beta, 10, seed=5, iteration=2)
    assert ss == n
    assert normwise_err(xtx, rxtx) < TOL
    assert vec_err(xtz, rxtz) < 1e-10
    _, xtz2, ss2 = ctx.probit_step(beta, 10, seed=5, iteration=2, want_xtx=False)
    assert ss2 == n and vec_err(xtz2, rxtz) < 1e-10
    ctx.close()
