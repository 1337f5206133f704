"""ctypes front-end of the C oracle (oracle/auxmix_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never by boom_b200.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libauxmix_oracle.so")

c_double_p = C.POINTER(C.c_double)
c_i64_p = C.POINTER(C.c_int64)
c_i32_p = C.POINTER(C.c_int32)


def build(force=False):
    src = os.path.join(_HERE, "auxmix_oracle.c")
    hdr = os.path.join(_HERE, "auxmix_oracle.h")
    if (force or not os.path.exists(_LIB_PATH)
            or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _LIB_PATH, src, "-lm"])
    return _LIB_PATH


class Mixture(C.Structure):
    _fields_ = [("K", C.c_int), ("mu", c_double_p), ("sigma", c_double_p), ("weights", c_double_p),
                ("log_weights", c_double_p)]


class PoissonTable(C.Structure):
    _fields_ = [("ntab", C.c_int), ("nu", c_i64_p), ("offset", c_i32_p), ("mu", c_double_p),
                ("sigma", c_double_p), ("log_weights", c_double_p), ("gaussian_cutoff", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.bo_rtrun_logit.restype = C.c_double
        _lib.bo_rtrun_logit.argtypes = [C.c_double, C.c_int, C.c_double]
        _lib.bo_unmix.restype = C.c_int
        _lib.bo_unmix.argtypes = [C.POINTER(Mixture), C.c_double, C.c_double, c_double_p]
        _lib.bo_binomial_from_uniform.restype = C.c_int64
        _lib.bo_binomial_from_uniform.argtypes = [C.c_int64, C.c_double, C.c_double]
        _lib.bo_dbinom_log.restype = C.c_double
        _lib.bo_dbinom_log.argtypes = [C.c_double] * 3
        _lib.bo_binomial_logit_loglike.restype = C.c_double
        _lib.bo_poisson_loglike.restype = C.c_double
        _lib.bo_binomial_logit_loglike_derivs.restype = C.c_double
        _lib.bo_poisson_loglike_derivs.restype = C.c_double
    return _lib


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class MixtureSpec:
    """A finite normal mixture (NormalMixtureApproximation) in host arrays."""

    def __init__(self, mu, sigma, weights):
        self.mu = _f64(mu)
        self.sigma = _f64(sigma)
        self.weights = _f64(weights)
        self.log_weights = np.empty_like(self.weights)
        lib().bo_log_array(C.c_int(len(self.weights)), _dp(self.weights), _dp(self.log_weights))
        self.K = len(self.sigma)
        self.c = Mixture(self.K, _dp(self.mu), _dp(self.sigma), _dp(self.weights), _dp(self.log_weights))


def golden_dir():
    return os.path.join(os.path.dirname(_HERE), "tests", "golden")


def logit_mixture():
    """The 9-component logit table as dumped from the live reference object
    (BinomialLogitDataImputer::mixture_approximation, BinomialLogitDataImputer.hpp:51)."""
    with open(os.path.join(golden_dir(), "logit_mixture.json")) as f:
        g = json.load(f)
    return MixtureSpec(g["mu"], g["sigma"], g["weights"])


class PoissonTableSpec:
    """NormalMixtureApproximationTable::serialize() -> CSR arrays."""

    def __init__(self, serialized, gaussian_cutoff):
        s = np.asarray(serialized, dtype=np.float64)
        nu, off, w, sig, mu = [], [0], [], [], []
        i = 0
        while i < len(s):
            nu.append(int(round(s[i])))
            K = int(round(s[i + 1]))
            w.extend(s[i + 2:i + 2 + K])
            sig.extend(s[i + 2 + K:i + 2 + 2 * K])
            mu.extend(s[i + 2 + 2 * K:i + 2 + 3 * K])
            off.append(off[-1] + K)
            i += 2 + 3 * K
        self.nu = np.asarray(nu, dtype=np.int64)
        self.offset = np.asarray(off, dtype=np.int32)
        self.weights = _f64(w)
        self.sigma = _f64(sig)
        self.mu = _f64(mu)
        self.log_weights = np.empty_like(self.weights)
        lib().bo_log_array(C.c_int(len(self.weights)), _dp(self.weights), _dp(self.log_weights))
        self.gaussian_cutoff = int(gaussian_cutoff)
        self.c = PoissonTable(len(self.nu), self.nu.ctypes.data_as(c_i64_p), self.offset.ctypes.data_as(c_i32_p),
                              _dp(self.mu), _dp(self.sigma), _dp(self.log_weights), self.gaussian_cutoff)

    def entry(self, nu):
        idx = int(np.searchsorted(self.nu, nu))
        if idx >= len(self.nu) or self.nu[idx] != nu:
            raise KeyError(nu)
        a, b = self.offset[idx], self.offset[idx + 1]
        return MixtureSpec(self.mu[a:b], self.sigma[a:b], self.weights[a:b])


def poisson_table():
    with open(os.path.join(golden_dir(), "poisson_mixture_table.json")) as f:
        g = json.load(f)
    return PoissonTableSpec(g["serialized"], g["largest_index"])


# ---------------------------------------------------------------------------- primitives
def uniform_pair(seed, iteration, row, slot):
    u = (C.c_double * 2)()
    lib().bo_uniform_pair(C.c_uint64(seed), C.c_uint64(iteration), C.c_uint64(row), C.c_uint32(slot), u)
    return u[0], u[1]


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().bo_philox4x32_10(c, k, o)
    return list(o)


def unmix(mix, residual, unif, want_post=False):
    post = np.zeros(mix.K)
    k = lib().bo_unmix(C.byref(mix.c), residual, unif, _dp(post))
    return (k, post) if want_post else k


def rtrun_logit(eta, success, unif):
    return lib().bo_rtrun_logit(eta, int(success), unif)


def trun_norm_moments(mu, sigma, cutpoint, positive):
    m, v = C.c_double(), C.c_double()
    lib().bo_trun_norm_moments(C.c_double(mu), C.c_double(sigma), C.c_double(cutpoint), C.c_int(int(positive)),
                               C.byref(m), C.byref(v))
    return m.value, v.value


def logit_impute(mix, clt_threshold, ntrials, y, eta, seed, iteration, row):
    s, w = C.c_double(), C.c_double()
    kc = np.zeros(mix.K, dtype=np.int64)
    rc = lib().bo_logit_impute(C.byref(mix.c), C.c_int(clt_threshold), C.c_double(ntrials), C.c_double(y),
                               C.c_double(eta), C.c_uint64(seed), C.c_uint64(iteration), C.c_uint64(row),
                               C.byref(s), C.byref(w), kc.ctypes.data_as(c_i64_p))
    if rc:
        raise ValueError("bo_logit_impute rc=%d" % rc)
    return s.value, w.value, kc


# ---------------------------------------------------------------------------- steps
def accumulate(X, weight, weighted_value):
    X = _f64(X)
    n, p = X.shape
    xtx = np.zeros((p, p))
    xty = np.zeros(p)
    w, s = _f64(weight), _f64(weighted_value)
    lib().bo_accumulate(C.c_int64(n), C.c_int(p), _dp(X), C.c_int64(p), _dp(w), _dp(s), _dp(xtx), _dp(xty))
    return xtx.T.copy(), xty  # column major -> numpy (symmetric anyway)


def logit_step(X, y, ntrials, beta, clt_threshold, mix, seed, iteration, row_offset=0):
    X, y, ntrials, beta = _f64(X), _f64(y), _f64(ntrials), _f64(beta)
    n, p = X.shape
    xtx = np.zeros((p, p))
    xty = np.zeros(p)
    ss = C.c_int64()
    kc = np.zeros(mix.K, dtype=np.int64)
    rc = lib().bo_logit_step(C.c_int64(n), C.c_int(p), _dp(X), C.c_int64(p), _dp(y), _dp(ntrials), _dp(beta),
                             C.c_int(clt_threshold), C.byref(mix.c), C.c_uint64(seed), C.c_uint64(iteration),
                             C.c_uint64(row_offset), _dp(xtx), _dp(xty), C.byref(ss), kc.ctypes.data_as(c_i64_p))
    if rc:
        raise ValueError("bo_logit_step rc=%d" % rc)
    return xtx.T.copy(), xty, ss.value, kc


def logit_draw(X, y, ntrials, beta, clt_threshold, mix, seed, iteration, row_offset=0):
    X, y, ntrials, beta = _f64(X), _f64(y), _f64(ntrials), _f64(beta)
    n, p = X.shape
    s = np.zeros(n)
    w = np.zeros(n)
    rc = lib().bo_logit_draw(C.c_int64(n), C.c_int(p), _dp(X), C.c_int64(p), _dp(y), _dp(ntrials), _dp(beta),
                             C.c_int(clt_threshold), C.byref(mix.c), C.c_uint64(seed), C.c_uint64(iteration),
                             C.c_uint64(row_offset), _dp(s), _dp(w))
    if rc:
        raise ValueError("bo_logit_draw rc=%d" % rc)
    return s, w


def poisson_step(X, y, exposure, beta, tab, seed, iteration, row_offset=0):
    X, exposure, beta = _f64(X), _f64(exposure), _f64(beta)
    y = np.ascontiguousarray(y, dtype=np.int64)
    n, p = X.shape
    xtx = np.zeros((p, p))
    xty = np.zeros(p)
    sc = np.zeros(4)
    rc = lib().bo_poisson_step(C.c_int64(n), C.c_int(p), _dp(X), C.c_int64(p), y.ctypes.data_as(c_i64_p),
                               _dp(exposure), _dp(beta), C.byref(tab.c), C.c_uint64(seed), C.c_uint64(iteration),
                               C.c_uint64(row_offset), _dp(xtx), _dp(xty), _dp(sc))
    if rc:
        raise ValueError("bo_poisson_step rc=%d" % rc)
    return xtx.T.copy(), xty, sc


def poisson_draw(X, y, exposure, beta, tab, seed, iteration, row_offset=0):
    X, exposure, beta = _f64(X), _f64(exposure), _f64(beta)
    y = np.ascontiguousarray(y, dtype=np.int64)
    n, p = X.shape
    out = np.zeros((n, 6))
    k2 = np.zeros((n, 2), dtype=np.int32)
    rc = lib().bo_poisson_draw(C.c_int64(n), C.c_int(p), _dp(X), C.c_int64(p), y.ctypes.data_as(c_i64_p),
                               _dp(exposure), _dp(beta), C.byref(tab.c), C.c_uint64(seed), C.c_uint64(iteration),
                               C.c_uint64(row_offset), _dp(out), k2.ctypes.data_as(c_i32_p))
    if rc:
        raise ValueError("bo_poisson_draw rc=%d" % rc)
    return out, k2


def rtrun_norm_unit(eta, positive, unif):
    lib().bo_rtrun_norm_unit.restype = C.c_double
    return lib().bo_rtrun_norm_unit(C.c_double(eta), C.c_int(int(positive)), C.c_double(unif))


def probit_step(X, y, ntrials, beta, clt_threshold, seed, iteration, row_offset=0):
    """(xtx, xtz, per-row sum of latent z) of BinomialProbitSpikeSlabSampler::impute_latent_data on the shared Philox stream."""
    X, y, ntrials, beta = _f64(X), _f64(y), _f64(ntrials), _f64(beta)
    n, p = X.shape
    xtx = np.zeros((p, p))
    xtz = np.zeros(p)
    draws = np.zeros(n)
    rc = lib().bo_probit_step(C.c_int64(n), C.c_int(p), _dp(X), C.c_int64(p), _dp(y), _dp(ntrials), _dp(beta), C.c_int(clt_threshold),
                              C.c_uint64(seed), C.c_uint64(iteration), C.c_uint64(row_offset), _dp(xtx), _dp(xtz), _dp(draws))
    if rc:
        raise ValueError("bo_probit_step rc=%d" % rc)
    return xtx.T.copy(), xtz, draws


def rgamma(shape, rate, seed, iteration, row):
    """Gamma(shape, rate) from the row's Philox blocks (Marsaglia-Tsang): the draw behind TDataImputer::impute."""
    out = C.c_double(0.0)
    rc = lib().bo_rgamma(C.c_double(shape), C.c_double(rate), C.c_uint64(seed), C.c_uint64(iteration), C.c_uint64(row), C.byref(out))
    if rc:
        raise ValueError("bo_rgamma rc=%d" % rc)
    return out.value


def student_step(X, y, beta, sigma, nu, seed, iteration, row_offset=0):
    """(xtwx, xtwy, scalars[n, y'Wy, sum w, sum log w], weights) of TRegressionSampler::impute_latent_data on the shared stream."""
    X, y, beta = _f64(X), _f64(y), _f64(beta)
    n, p = X.shape
    xtwx, xtwy, sc, w = np.zeros((p, p)), np.zeros(p), np.zeros(4), np.zeros(n)
    rc = lib().bo_student_step(C.c_int64(n), C.c_int(p), _dp(X), C.c_int64(p), _dp(y), _dp(beta), C.c_double(sigma), C.c_double(nu),
                               C.c_uint64(seed), C.c_uint64(iteration), C.c_uint64(row_offset), _dp(xtwx), _dp(xtwy), _dp(sc), _dp(w))
    if rc:
        raise ValueError("bo_student_step rc=%d" % rc)
    return xtwx.T.copy(), xtwy, sc, w


def student_loglike(X, y, beta, sigma, nu):
    X, y, beta = _f64(X), _f64(y), _f64(beta)
    n, p = X.shape
    lib().bo_student_loglike.restype = C.c_double
    return lib().bo_student_loglike(C.c_int64(n), C.c_int(p), _dp(X), C.c_int64(p), _dp(y), _dp(beta), C.c_double(sigma), C.c_double(nu))


def synth_student(n, p, nonzero, seed, sigma=1.5, nu=4.0, intercept=0.5, row_offset=0):
    """X, y = X beta + sigma t_nu, beta: the synthetic Student-t regression every arm shares."""
    X = synth_x(n, p, seed, 1.0, row_offset)
    beta = synth_beta(p, nonzero, intercept)
    y = np.zeros(n)
    lib().bo_synth_student_y(C.c_int64(n), C.c_int(p), _dp(X), C.c_int64(p), _dp(beta), C.c_double(sigma), C.c_double(nu),
                             C.c_uint64(seed), C.c_uint64(row_offset), _dp(y))
    return X, y, beta


def binomial_logit_loglike(X, y, ntrials, beta):
    X, y, ntrials, beta = _f64(X), _f64(y), _f64(ntrials), _f64(beta)
    n, p = X.shape
    return lib().bo_binomial_logit_loglike(C.c_int64(n), C.c_int(p), _dp(X), C.c_int64(p), _dp(y), _dp(ntrials),
                                           _dp(beta))


def poisson_loglike(X, y, exposure, beta):
    X, exposure, beta = _f64(X), _f64(exposure), _f64(beta)
    y = np.ascontiguousarray(y, dtype=np.int64)
    n, p = X.shape
    return lib().bo_poisson_loglike(C.c_int64(n), C.c_int(p), _dp(X), C.c_int64(p), y.ctypes.data_as(c_i64_p),
                                    _dp(exposure), _dp(beta))


def binomial_logit_loglike_derivs(X, y, ntrials, beta, log_alpha=0.0):
    """(loglike, gradient, hessian) as BinomialLogitModel::log_likelihood(beta, &g, &h) (BinomialLogitModel.cpp:140-180)."""
    X, y, ntrials, beta = _f64(X), _f64(y), _f64(ntrials), _f64(beta)
    n, p = X.shape
    g, h = np.empty(p), np.empty((p, p))
    ll = lib().bo_binomial_logit_loglike_derivs(C.c_int64(n), C.c_int(p), _dp(X), C.c_int64(p), _dp(y), _dp(ntrials), _dp(beta),
                                                C.c_double(log_alpha), _dp(g), _dp(h))
    return ll, g, h


def poisson_loglike_derivs(X, y, exposure, beta):
    X, exposure, beta = _f64(X), _f64(exposure), _f64(beta)
    y = np.ascontiguousarray(y, dtype=np.int64)
    n, p = X.shape
    g, h = np.empty(p), np.empty((p, p))
    ll = lib().bo_poisson_loglike_derivs(C.c_int64(n), C.c_int(p), _dp(X), C.c_int64(p), y.ctypes.data_as(c_i64_p), _dp(exposure),
                                         _dp(beta), _dp(g), _dp(h))
    return ll, g, h


# ---------------------------------------------------------------------------- synthetic data
def synth_beta(p, nonzero, intercept):
    b = np.zeros(p)
    lib().bo_synth_beta(C.c_int(p), C.c_int(nonzero), C.c_double(intercept), _dp(b))
    return b


def synth_x(n, p, seed, xscale=1.0, row_offset=0):
    X = np.zeros((n, p))
    lib().bo_synth_x(C.c_int64(n), C.c_int(p), C.c_uint64(seed), C.c_double(xscale), C.c_uint64(row_offset), _dp(X),
                     C.c_int64(p))
    return X


def synth_binomial(n, p, nonzero, seed, max_trials=1, intercept=-1.0, row_offset=0):
    X = synth_x(n, p, seed, 1.0, row_offset)
    beta = synth_beta(p, nonzero, intercept)
    y = np.zeros(n)
    nt = np.zeros(n)
    lib().bo_synth_binomial_y(C.c_int64(n), C.c_int(p), _dp(X), C.c_int64(p), _dp(beta), C.c_uint64(seed),
                              C.c_int(max_trials), C.c_uint64(row_offset), _dp(y), _dp(nt))
    return X, y, nt, beta


def synth_poisson(n, p, nonzero, seed, intercept=0.5, row_offset=0):
    X = synth_x(n, p, seed, 0.3, row_offset)
    beta = synth_beta(p, nonzero, intercept)
    y = np.zeros(n, dtype=np.int64)
    ex = np.zeros(n)
    lib().bo_synth_poisson_y(C.c_int64(n), C.c_int(p), _dp(X), C.c_int64(p), _dp(beta), C.c_uint64(seed),
                             C.c_uint64(row_offset), y.ctypes.data_as(c_i64_p), _dp(ex))
    return X, y, ex, beta
