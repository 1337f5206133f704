import numpy as np

from oracle import oracle as O


def normwise_err(a, ref):
    """max |a_ij - ref_ij| / sqrt(ref_ii ref_jj)  ('hard parts': the parity metric for X'WX)."""
    d = np.sqrt(np.abs(np.diag(ref)))
    d[d == 0] = 1.0
    return float(np.max(np.abs(a - ref) / np.outer(d, d)))


def vec_err(a, ref):
    return float(np.max(np.abs(a - ref)) / max(1e-300, np.max(np.abs(ref))))


def logit_ctx(X, y, nt, path=0, device=0):
    import boom_b200
    ctx = boom_b200.Context(device)
    mix = O.logit_mixture()
    ctx.set_logit_mixture(mix.mu, mix.sigma, mix.weights)
    ctx.set_option("path", path)
    ctx.upload_binomial(X, y, nt)
    return ctx, mix


def poisson_ctx(X, y, ex, path=0, device=0):
    import boom_b200
    ctx = boom_b200.Context(device)
    tab = O.poisson_table()
    ctx.set_poisson_table(tab.nu, tab.offset, tab.weights, tab.mu, tab.sigma, tab.gaussian_cutoff)
    ctx.set_option("path", path)
    ctx.upload_poisson(X, y, ex)
    return ctx, tab
