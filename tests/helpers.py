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


# ---- the oracle over disjoint row blocks on all host cores (ctypes releases the GIL): the additive statistics of the
# blocks are summed in long double.  What the at-scale parity tests (tests/test_gpu_scale.py) compare the device with.
def _blocks(n, nblocks):
    cuts = [n * b // nblocks for b in range(nblocks + 1)]
    return [(cuts[b], cuts[b + 1]) for b in range(nblocks) if cuts[b + 1] > cuts[b]]


def host_threads():
    import os
    return max(1, min(32, os.cpu_count() or 1))


def parallel_rows(n, fn, nblocks=None):
    """fn(row0, row1) for disjoint row blocks on a thread pool; results in block order."""
    from concurrent.futures import ThreadPoolExecutor
    nblocks = nblocks or 4 * host_threads()
    with ThreadPoolExecutor(host_threads()) as ex:
        return list(ex.map(lambda ab: fn(*ab), _blocks(n, nblocks)))


def synth_x_parallel(n, p, seed, xscale=1.0):
    X = np.empty((n, p))

    def fill(a, b):
        X[a:b] = O.synth_x(b - a, p, seed, xscale, row_offset=a)
    parallel_rows(n, fill)
    return X


def synth_binomial_parallel(n, p, nonzero, seed, max_trials=1):
    X = np.empty((n, p))
    y = np.empty(n)
    nt = np.empty(n)
    beta = O.synth_beta(p, nonzero, -1.0)

    def fill(a, b):
        X[a:b], y[a:b], nt[a:b], _ = O.synth_binomial(b - a, p, nonzero, seed, max_trials=max_trials, row_offset=a)
    parallel_rows(n, fill)
    return X, y, nt, beta


def sum_long_double(parts):
    acc = np.zeros(parts[0].shape, dtype=np.longdouble)
    for q in parts:
        acc += q
    return np.asarray(acc, dtype=np.float64)


def accumulate_blocked(X, w, s):
    """O.accumulate on disjoint row blocks, block results summed in long double."""
    parts = parallel_rows(X.shape[0], lambda a, b: O.accumulate(X[a:b], w[a:b], s[a:b]))
    return sum_long_double([q[0] for q in parts]), sum_long_double([q[1] for q in parts])


def logit_step_blocked(X, y, nt, beta, clt, mix, seed, iteration):
    parts = parallel_rows(X.shape[0], lambda a, b: O.logit_step(X[a:b], y[a:b], nt[a:b], beta, clt, mix, seed, iteration, row_offset=a))
    return (sum_long_double([q[0] for q in parts]), sum_long_double([q[1] for q in parts]), sum(q[2] for q in parts),
            sum(q[3] for q in parts))
