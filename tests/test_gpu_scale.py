"""GPU parity at the shapes bench.py runs at (VERDICT r01 "what's weak" 1, 3), through the C ABI.

The small parity cases (tests/test_gpu_parity.py) never reach the launch geometry of the benchmark configs:
C3 runs the split-K SYRK with ksplit = 444 (4440 CTAs), C4 with 32 column blocks / 528 regions and TMA tiles at
column offsets >= 512, C5 streams 25-200 M rows through the single-pass kernel.  Here the same oracle is run
on disjoint row blocks on all host cores (tests/helpers.py: *_blocked, block results summed in long double) so
that the comparison stays in seconds:

  * accumulate + logit_step at n = 1.2 M, p = 500  (ksplit = 444, the C3 geometry per CTA);
  * p in {520, 1000, 4000} (nblk 5 / 8 / 32, ragged last block), incl. the page-locked direct-copy landing;
  * a C5-shaped single pass, n = 50 M, p = 16: sample_size, X'WX, X'Wz and the indicator histogram;
  * the branches the small cases never take: Poisson |eta| >= 600 (PoissonDataImputer.cpp:55-79), exposure 0,
    binomial rows with n_i = 0 (counted in sample_size, no trials) and fractional n_i / y_i (the reference's loop
    bounds are doubles, BinomialLogitDataImputer.cpp:139-141).

Tolerances: deterministic statistics 1e-12 normwise (north_star); steps that draw 1e-11 (libm vs CUDA ulps enter
before the sum); indicator-derived integers exact.
"""
import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as H
from tests.helpers import logit_ctx, normwise_err, poisson_ctx, vec_err

pytestmark = pytest.mark.gpu


def _latents(n, seed):
    rng = np.random.default_rng(seed)
    return 0.05 + 1.5 * rng.random(n), 6.0 * (rng.random(n) - 0.5)


def test_c3_geometry_accumulate_and_step():
    """n = 1.2 M, p = 500: launch_syrk picks ksplit = min(444, n / 2048) = 444, as at C3 (n = 10 M)."""
    n, p = 1_200_000, 500
    X, y, nt, beta = H.synth_binomial_parallel(n, p, 20, seed=301)
    w, s = _latents(n, 5)
    ctx, mix = logit_ctx(X, y, nt)
    xtx, xty = ctx.accumulate(w, s)
    rxtx, rxty = H.accumulate_blocked(X, w, s)
    assert normwise_err(xtx, rxtx) < 1e-12
    assert vec_err(xty, rxty) < 1e-12
    np.testing.assert_array_equal(xtx, xtx.T)
    ctx.set_option("syrk_order", 2)        # C3's geometry under the paired-diagonal grid: (0,0)+(1,1) and (2,2)+(3,3 ragged)
    xtx2, xty2 = ctx.accumulate(w, s)
    assert normwise_err(xtx2, rxtx) < 1e-12 and vec_err(xty2, rxty) < 1e-12
    ctx.set_option("syrk_order", 1)
    xtx, xty, ss = ctx.logit_step(beta, 10, seed=41, iteration=7)
    rxtx, rxty, rss, _ = H.logit_step_blocked(X, y, nt, beta, 10, mix, 41, 7)
    assert ss == rss == n
    assert normwise_err(xtx, rxtx) < 1e-11
    assert vec_err(xty, rxty) < 1e-10
    ctx.close()


@pytest.mark.parametrize("n,p", [(20_011, 520), (20_011, 1000), (6_007, 4000), (9_001, 2049), (12_345, 130), (15_000, 1600)])
def test_wide_p_accumulate_and_step(n, p):
    """nblk = 5 / 8 / 32 / 17 column blocks with a ragged last block (p = 2049: ONE column in it); TMA tiles at column
    offsets >= 512; at p >= 1000 the p x p matrix also lands through the page-locked direct-copy path."""
    import boom_b200
    X, y, nt, beta = H.synth_binomial_parallel(n, p, 40, seed=310 + p)
    w, s = _latents(n, p)
    ctx, mix = logit_ctx(X, y, nt)
    xtx, xty = ctx.accumulate(w, s)
    rxtx, rxty = H.accumulate_blocked(X, w, s)
    assert normwise_err(xtx, rxtx) < 1e-12
    assert vec_err(xty, rxty) < 1e-12
    np.testing.assert_array_equal(xtx, xtx.T)
    # the order in which the CTAs are scheduled (off-diagonal regions first / k-slice major) and the split-K depth change who
    # computes what when, never the sums: same partial per (k-slice, region), same reduction order -> the same bits
    ctx.set_option("syrk_order", 0)
    xtx0, xty0 = ctx.accumulate(w, s)
    np.testing.assert_array_equal(xtx0, xtx)
    np.testing.assert_array_equal(xty0, xty)
    ctx.set_option("syrk_order", 1)
    ctx.set_option("syrk_waves", 7)
    xtx0, xty0 = ctx.accumulate(w, s)
    assert normwise_err(xtx0, rxtx) < 1e-12 and vec_err(xty0, rxty) < 1e-12
    # order 2: diagonal regions in pairs on one CTA (strip form on each of its two panels, a ragged last region through the
    # zero columns TMA delivers; an odd last diagonal region alone), super-tiled region order
    for waves in (30, 3):
        ctx.set_option("syrk_order", 2)
        ctx.set_option("syrk_waves", waves)
        xtx0, xty0 = ctx.accumulate(w, s)
        assert normwise_err(xtx0, rxtx) < 1e-12 and vec_err(xty0, rxty) < 1e-12
        np.testing.assert_array_equal(xtx0, xtx0.T)
    ctx.set_option("syrk_order", 1)
    ctx.set_option("syrk_waves", 30)
    del xtx, xtx0
    rxtx, rxty, rss, _ = H.logit_step_blocked(X, y, nt, beta, 10, mix, 43, 2)
    out = np.full((p, p), np.nan)
    boom_b200.Context.pin_host(out)
    try:
        xtx, xty, ss = ctx.logit_step(beta, 10, seed=43, iteration=2, xtx=out)
        assert xtx is out and ss == rss == n
        assert normwise_err(xtx, rxtx) < 1e-11
        assert vec_err(xty, rxty) < 1e-10
        np.testing.assert_array_equal(xtx, xtx.T)
    finally:
        boom_b200.Context.unpin_host(out)
    ctx.close()


def test_c5_shaped_single_pass():
    """n = 50 M, p = 16 (C5 holds 25 M rows per GPU on 8 GPUs, 200 M on one): the persistent single-pass kernel over
    1.56 M slices.  sample_size exact, the indicator histogram exact (decoded from sum of info is not possible here, so
    through the statistics: X'WX's (0, 0) entry is sum_i 1/sigma_k(i)^2 because x_i0 = 1), statistics to 1e-11."""
    n, p = 50_000_000, 16
    X, y, nt, beta = H.synth_binomial_parallel(n, p, 5, seed=320)
    ctx, mix = logit_ctx(X, y, nt)
    xtx, xty, ss = ctx.logit_step(beta, 10, seed=47, iteration=3)
    rxtx, rxty, rss, kc = H.logit_step_blocked(X, y, nt, beta, 10, mix, 47, 3)
    assert ss == rss == n
    assert normwise_err(xtx, rxtx) < 1e-11
    assert vec_err(xty, rxty) < 1e-10
    # sum of the drawn information = sum_k count_k / sigma_k^2: ties the device's statistics to the oracle's INTEGER histogram
    info_from_counts = float(np.sum(kc.astype(np.longdouble) / (mix.sigma.astype(np.longdouble) ** 2)))
    assert int(kc.sum()) == n
    assert abs(xtx[0, 0] - info_from_counts) <= 1e-12 * info_from_counts
    ctx.close()


def test_logit_zero_and_fractional_trials():
    """n_i = 0 rows draw nothing but count in sample_size; fractional n_i / y_i follow the reference's double loop bounds
    (i < n_i, success iff i < y_i): n_i = 2.5, y_i = 1.5 is three trials, two of them successes."""
    n, p = 6000, 9
    X, y, nt, beta = O.synth_binomial(n, p, 3, seed=330, max_trials=6)
    nt[1::11] = 2.5
    y[1::11] = 1.5
    nt[2::13] = 9.75     # still <= clt_threshold = 10: ten trials
    y[2::13] = 0.25      # one success
    nt[::7] = 0.0
    y[::7] = 0.0
    for path in (0, 2):
        ctx, mix = logit_ctx(X, y, nt, path=path)
        s, w = ctx.logit_draw(beta, 10, seed=3, iteration=1)
        rs, rw = O.logit_draw(X, y, nt, beta, 10, mix, 3, 1)
        np.testing.assert_allclose(w, rw, rtol=1e-13)
        np.testing.assert_allclose(s, rs, rtol=1e-9, atol=1e-9)
        assert np.all(w[::7] == 0.0) and np.all(s[::7] == 0.0)
        xtx, xty, ss = ctx.logit_step(beta, 10, seed=3, iteration=1)
        rxtx, rxty, rss, _ = O.logit_step(X, y, nt, beta, 10, mix, 3, 1)
        assert ss == rss == n
        assert normwise_err(xtx, rxtx) < 1e-11 and vec_err(xty, rxty) < 1e-10
        ctx.close()


def test_logit_mixture_too_wide_for_clt_branch_is_an_error():
    """A mixture with more than 9 components cannot take the CLT branch (slot layout of the conditional binomials):
    reported as an error, in the oracle and on the device alike."""
    import boom_b200
    n, p = 200, 3
    X, y, nt, beta = O.synth_binomial(n, p, 2, seed=331, max_trials=40)
    mix = O.logit_mixture()
    mu = np.zeros(10)
    sigma = np.concatenate([mix.sigma, [5.0]])
    weights = np.concatenate([mix.weights * 0.99, [0.01]])
    ctx = boom_b200.Context(0)
    ctx.set_logit_mixture(mu, sigma, weights)
    ctx.upload_binomial(X, y, nt)
    with pytest.raises(boom_b200.BoomGpuError, match="invalid observation"):
        ctx.logit_step(beta, 10, 1, 0)
    xtx, _, ss = ctx.logit_step(beta, 1000, 1, 0)    # every row on the per-trial branch: fine with K = 10
    assert ss == n and np.all(np.isfinite(xtx))
    ctx.close()


def test_poisson_extreme_linear_predictors():
    """|eta| >= 600: the lse2 statement (delta > 0) and the eta + extreme-value statement (delta <= 0: exposure 0);
    value by value and indicator by indicator against the oracle, both paths."""
    n, p = 4000, 5
    X, y, ex, beta = O.synth_poisson(n, p, 2, seed=340)
    beta = beta.copy()
    beta[1] = 0.5
    X[:40, 1] = 1400.0        # eta ~ +700
    X[40:80, 1] = -1400.0     # eta ~ -700
    X[80:100, 1] = 1199.0     # eta just below 600
    X[100:120, 1] = 1201.0    # just above
    y[:80:2] = 0
    y[1:80:2] = np.arange(1, 80, 2) % 17 + 1
    ex[:80:5] = 0.0           # delta = 0 when y = 0 as well
    y[:80:5] = 0
    for path in (0, 2):
        ctx, tab = poisson_ctx(X, y, ex, path=path)
        out, k2 = ctx.poisson_draw(beta, seed=9, iteration=4)
        ref, rk2 = O.poisson_draw(X, y, ex, beta, tab, 9, 4)
        assert np.array_equal(k2, rk2)
        assert np.all(np.isfinite(out))
        np.testing.assert_allclose(out, ref, rtol=1e-10, atol=1e-12)
        xtx, xty, sc = ctx.poisson_step(beta, seed=9, iteration=4)
        rxtx, rxty, rsc = O.poisson_step(X, y, ex, beta, tab, 9, 4)
        assert sc[0] == rsc[0] == n + np.count_nonzero(y)
        assert normwise_err(xtx, rxtx) < 1e-11 and vec_err(xty, rxty) < 1e-10
        np.testing.assert_allclose(sc, rsc, rtol=1e-10)
        ctx.close()


def test_c2_shaped_poisson_step():
    """C2's shape (n = 1 M, p = 50, NB = 7 tiles) through the single-pass kernel against the blocked oracle."""
    n, p = 1_000_000, 50
    X = H.synth_x_parallel(n, p, seed=350, xscale=0.3)
    beta = O.synth_beta(p, 5, 0.5)
    y = np.empty(n, dtype=np.int64)
    ex = np.empty(n)

    def fill(a, b):
        _, y[a:b], ex[a:b], _ = O.synth_poisson(b - a, p, 5, 350, row_offset=a)
    H.parallel_rows(n, fill)
    ctx, tab = poisson_ctx(X, y, ex)
    xtx, xty, sc = ctx.poisson_step(beta, seed=13, iteration=6)
    parts = H.parallel_rows(n, lambda a, b: O.poisson_step(X[a:b], y[a:b], ex[a:b], beta, tab, 13, 6, row_offset=a))
    rxtx, rxty = H.sum_long_double([q[0] for q in parts]), H.sum_long_double([q[1] for q in parts])
    rsc = H.sum_long_double([q[2] for q in parts])
    assert sc[0] == rsc[0] == n + np.count_nonzero(y)
    assert normwise_err(xtx, rxtx) < 1e-11 and vec_err(xty, rxty) < 1e-10
    np.testing.assert_allclose(sc, rsc, rtol=1e-10)
    ctx.close()


# ---------------------------------------------------------------- boundary fixes
def test_standalone_poisson_sampler_extends_the_table_for_offgrid_counts():
    """Counts off the shipped grid (305, 1234, 12345 ...) used to fail the standalone samplers with BOOMGPU_ERR_DATA; now the
    host adds the entries by the reference's rule (NormalMixtureApproximation.cpp:472-532) before the first draw, and the
    device statistics equal the oracle's with the same (extended) table."""
    import boom_b200
    boom_b200.load_poisson_mixture_table()
    h = boom_b200.host()
    n, p = 3000, 4
    X, y, ex, beta = O.synth_poisson(n, p, 2, seed=360)
    y[:8] = [305, 1234, 12345, 29999, 777, 495, 45000, 301]
    model = boom_b200.PoissonRegressionModel(X, y, ex)
    model.set_Beta(beta)
    s = boom_b200.PoissonRegressionAuxMixSampler(model, boom_b200.MvnModel(np.zeros(p), np.eye(p)), 1, boom_b200.RNG(5))
    model.set_method(s)
    s.impute_latent_data()
    suf = s.complete_data_sufficient_statistics
    assert suf.n == n + np.count_nonzero(y) and np.all(np.isfinite(np.array(suf.xtx)))
    ser = np.array(h.poisson_mixture_table())
    tab = O.PoissonTableSpec(ser, 30000)
    for v in (305, 1234, 12345, 29999, 777, 495, 301):
        tab.entry(v)                                  # present now
    # the same table through the C ABI against the oracle
    ctx = boom_b200.Context(0)
    ctx.set_poisson_table(tab.nu, tab.offset, tab.weights, tab.mu, tab.sigma, tab.gaussian_cutoff)
    ctx.upload_poisson(X, y, ex)
    out, k2 = ctx.poisson_draw(beta, seed=2, iteration=1)
    ref, rk2 = O.poisson_draw(X, y, ex, beta, tab, 2, 1)
    assert np.array_equal(k2, rk2)
    np.testing.assert_allclose(out, ref, rtol=1e-10, atol=1e-12)
    present = ctx.poisson_counts_present(30000)
    assert set(np.nonzero(present)[0]) == set(int(v) for v in np.unique(y) if v < 30000)
    ctx.close()
    for _ in range(3):
        model.sample_posterior()
    assert np.all(np.isfinite(model.Beta))
    boom_b200.load_poisson_mixture_table()


def test_adopted_rows_that_tma_cannot_describe():
    """p > 64 with an odd leading dimension / a base pointer that is not 16-byte aligned: the rows are copied once into a
    padded device buffer and the two-pass path runs on that (it used to refuse them)."""
    import torch
    import boom_b200
    n, p = 3000, 101
    X, y, nt, beta = O.synth_binomial(n, p, 5, seed=370)
    mix = O.logit_mixture()
    rxtx, rxty, rss, _ = O.logit_step(X, y, nt, beta, 10, mix, 11, 2)
    dev = torch.device("cuda", 0)
    buf = torch.zeros(n * p + 1, dtype=torch.float64, device=dev)
    for shift in (0, 1):        # contiguous n x 101 (odd ldx), and the same shifted by 8 bytes (unaligned base)
        Xd = buf[shift:shift + n * p].view(n, p)
        Xd.copy_(torch.from_numpy(X))
        yd, ntd = torch.from_numpy(y).to(dev), torch.from_numpy(nt).to(dev)
        torch.cuda.synchronize()
        ctx = boom_b200.Context(0)
        ctx.set_logit_mixture(mix.mu, mix.sigma, mix.weights)
        ctx.adopt_binomial(n, p, Xd.data_ptr(), p, yd.data_ptr(), ntd.data_ptr(), keepalive=(Xd, yd, ntd))
        xtx, xty, ss = ctx.logit_step(beta, 10, seed=11, iteration=2)
        assert ss == rss == n
        assert normwise_err(xtx, rxtx) < 1e-11 and vec_err(xty, rxty) < 1e-10
        ctx.close()


def test_clt_threshold_beyond_the_slot_field_is_rejected():
    import boom_b200
    X, y, nt, beta = O.synth_binomial(100, 3, 2, seed=1)
    ctx, _ = logit_ctx(X, y, nt)
    with pytest.raises(boom_b200.BoomGpuError, match="clt_threshold"):
        ctx.logit_step(beta, 70000, 1, 0)
    ctx.logit_step(beta, 65535, 1, 0)
    ctx.close()


@pytest.mark.parametrize("n,p,nnz", [(5003, 200, 0), (5003, 200, 1), (4001, 500, 21), (3000, 130, 32), (3000, 130, 129)])
def test_gather_imputer_pass_matches_the_dense_pass(n, p, nnz):
    """Two-pass path with a sparse beta (spike and slab): the imputer pass that reads only the included columns (option
    gather: 0 auto, 1 never, 2 whenever beta has a zero) against the dense pass and against the oracle."""
    X, y, nt, _ = O.synth_binomial(n, p, 5, seed=380 + nnz, max_trials=3)
    rng = np.random.default_rng(nnz)
    beta = np.zeros(p)
    cols = rng.choice(p, size=min(nnz, p), replace=False)
    beta[cols] = rng.normal(size=len(cols)) * 0.4
    mix = O.logit_mixture()
    rs, rw = O.logit_draw(X, y, nt, beta, 10, mix, 7, 3)
    rxtx, rxty, rss, _ = O.logit_step(X, y, nt, beta, 10, mix, 7, 3)
    for gather in (1, 2, 0):
        ctx, _ = logit_ctx(X, y, nt, path=2)
        ctx.set_option("gather", gather)
        s, w = ctx.logit_draw(beta, 10, seed=7, iteration=3)
        np.testing.assert_allclose(w, rw, rtol=1e-13)
        np.testing.assert_allclose(s, rs, rtol=1e-9, atol=1e-9)
        xtx, xty, ss = ctx.logit_step(beta, 10, seed=7, iteration=3)
        assert ss == rss == n
        assert normwise_err(xtx, rxtx) < 1e-11 and vec_err(xty, rxty) < 1e-10
        ctx.close()


# ---------------------------------------------------------------- active-set statistics
@pytest.mark.parametrize("n,p,k", [(5003, 200, 1), (4001, 500, 21), (3000, 130, 40), (2500, 1000, 128), (6000, 300, 9)])
def test_active_set_step_matches_the_full_statistics(n, p, k):
    """boomgpu_logit_step_active: G = (X'WX)[:, active], the diagonal and X'Wz of the SAME latents as the full step (same seed,
    same iteration) -- against the oracle's full matrix; boomgpu_weighted_column returns any other column for those latents,
    boomgpu_full_statistics the whole matrix."""
    X, y, nt, _ = O.synth_binomial(n, p, 5, seed=500 + k, max_trials=3)
    rng = np.random.default_rng(k)
    active = np.sort(rng.choice(p, size=k, replace=False))
    beta = np.zeros(p)
    beta[active] = rng.normal(size=k) * 0.3
    mix = O.logit_mixture()
    rxtx, rxty, rss, _ = O.logit_step(X, y, nt, beta, 10, mix, 19, 4)
    d = np.sqrt(np.diag(rxtx))
    ctx, _ = logit_ctx(X, y, nt)
    G, diag, xty, ss = ctx.logit_step_active(beta, 10, 19, 4, active)
    assert ss == rss == n
    assert np.max(np.abs(G - rxtx[:, active]) / np.outer(d, d[active])) < 1e-11
    assert np.max(np.abs(diag - np.diag(rxtx)) / (d * d)) < 1e-11
    assert vec_err(xty, rxty) < 1e-10
    for j in (0, p // 2, p - 1):
        col = ctx.weighted_column(j)
        assert np.max(np.abs(col - rxtx[:, j]) / (d * d[j])) < 1e-11
    xtx, xty2 = ctx.full_statistics()
    assert normwise_err(xtx, rxtx) < 1e-11 and vec_err(xty2, rxty) < 1e-10
    ctx.close()


def test_active_set_poisson_step():
    n, p, k = 4000, 150, 12
    X, y, ex, _ = O.synth_poisson(n, p, 4, seed=520)
    rng = np.random.default_rng(2)
    active = np.sort(rng.choice(p, size=k, replace=False))
    beta = np.zeros(p)
    beta[active] = rng.normal(size=k) * 0.2
    ctx, tab = poisson_ctx(X, y, ex)
    G, diag, xty, sc = ctx.poisson_step_active(beta, 23, 2, active)
    rxtx, rxty, rsc = O.poisson_step(X, y, ex, beta, tab, 23, 2)
    d = np.sqrt(np.diag(rxtx))
    assert np.max(np.abs(G - rxtx[:, active]) / np.outer(d, d[active])) < 1e-11
    assert np.max(np.abs(diag - np.diag(rxtx)) / (d * d)) < 1e-11
    assert vec_err(xty, rxty) < 1e-10
    np.testing.assert_allclose(sc, rsc, rtol=1e-10)
    ctx.close()


def test_large_pageable_upload_with_a_leading_dimension():
    """Uploads of 256 MB and more from pageable host memory go through page-locked staging buffers filled by several host
    threads (upload_x_staged); here with a host leading dimension larger than p (a row-strided view), which takes the
    row-by-row packing branch.  The statistics must be those of the compact matrix."""
    n, p, ld = 300_000, 120, 128
    rng = np.random.default_rng(9)
    wide = rng.normal(size=(n, ld))
    wide[:, 0] = 1.0
    X = wide[:, :p]                                        # strides (8 * 128, 8): passed with ldx = 128
    assert not X.flags.c_contiguous and X.nbytes >= 256 << 20
    y = (rng.random(n) < 0.4).astype(float)
    nt = np.ones(n)
    w, s = _latents(n, 3)
    ctx, _ = logit_ctx(X, y, nt)
    xtx, xty = ctx.accumulate(w, s)
    Xc = np.ascontiguousarray(X)
    rxtx, rxty = H.accumulate_blocked(Xc, w, s)
    assert normwise_err(xtx, rxtx) < 1e-12 and vec_err(xty, rxty) < 1e-12
    ctx.close()
