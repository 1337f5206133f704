"""ctypes binding of include/boomgpu.h.  Thin by design: every method is one C call."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_double_p = C.POINTER(C.c_double)
c_i64_p = C.POINTER(C.c_int64)
c_i32_p = C.POINTER(C.c_int32)

NUM_KERNEL_CLASSES = 5
KERNEL_CLASSES = ("fused_small", "impute_rows", "syrk_dmma", "reduce", "other")

# every symbol include/boomgpu.h declares (tests/test_capi_symbols.py checks the library exports them all)
SYMBOLS = (
    "boomgpu_create", "boomgpu_destroy", "boomgpu_last_error", "boomgpu_version", "boomgpu_set_stream",
    "boomgpu_set_row_offset", "boomgpu_set_option", "boomgpu_upload_binomial", "boomgpu_upload_poisson",
    "boomgpu_upload_begin", "boomgpu_upload_rows", "boomgpu_upload_end", "boomgpu_adopt_binomial", "boomgpu_adopt_poisson", "boomgpu_set_logit_mixture", "boomgpu_set_poisson_table",
    "boomgpu_logit_step", "boomgpu_poisson_step", "boomgpu_logit_step_active", "boomgpu_poisson_step_active",
    "boomgpu_weighted_column", "boomgpu_full_statistics", "boomgpu_probit_step", "boomgpu_probit_step_device", "boomgpu_probit_draw", "boomgpu_suf_len", "boomgpu_logit_step_device",
    "boomgpu_poisson_step_device", "boomgpu_synchronize", "boomgpu_suf_buffer", "boomgpu_download", "boomgpu_accumulate", "boomgpu_logit_draw",
    "boomgpu_poisson_draw", "boomgpu_binomial_loglike", "boomgpu_poisson_loglike", "boomgpu_binomial_loglike_derivs",
    "boomgpu_poisson_loglike_derivs", "boomgpu_binomial_loglike_derivs_device", "boomgpu_poisson_loglike_derivs_device",
    "boomgpu_poisson_counts_present", "boomgpu_select_columns", "boomgpu_binomial_loglike_derivs_selected",
    "boomgpu_poisson_loglike_derivs_selected", "boomgpu_binomial_loglike_derivs_selected_device", "boomgpu_pin_host", "boomgpu_unpin_host", "boomgpu_comm_unique_id", "boomgpu_comm_init", "boomgpu_comm_destroy", "boomgpu_allreduce", "boomgpu_kernel_launches",
    "boomgpu_get_timings",
    "boomgpu_upload_regression", "boomgpu_adopt_regression", "boomgpu_student_step", "boomgpu_student_step_device",
    "boomgpu_student_draw", "boomgpu_student_loglike", "boomgpu_student_step_active",
)


class BoomGpuError(RuntimeError):
    pass


def library_path():
    # BOOMGPU_LIBRARY: a differently tuned build of the same library (profiles/tune_small.sh); never a fallback
    return os.environ.get("BOOMGPU_LIBRARY") or os.path.join(_HERE, "libboomgpu.so")


def load_library():
    """Loads libboomgpu.so; fails loudly when it has not been built (no fallback exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise BoomGpuError(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or make -C boom_b200/csrc). boom_b200 has no CPU fallback." % path)
    lib = C.CDLL(path)
    lib.boomgpu_last_error.restype = C.c_char_p
    lib.boomgpu_last_error.argtypes = [C.c_void_p]
    lib.boomgpu_version.restype = C.c_char_p
    lib.boomgpu_suf_len.restype = C.c_int64
    lib.boomgpu_suf_len.argtypes = [C.c_int]
    lib.boomgpu_kernel_launches.restype = C.c_int64
    lib.boomgpu_kernel_launches.argtypes = [C.c_void_p]
    lib.boomgpu_destroy.restype = None
    lib.boomgpu_destroy.argtypes = [C.c_void_p]
    _LIB = lib
    return lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _dp(a):
    return a.ctypes.data_as(c_double_p)


class Context:
    """One boomgpu context = one CUDA device."""

    def __init__(self, device=0):
        self._lib = load_library()
        self._h = C.c_void_p()
        rc = self._lib.boomgpu_create(C.byref(self._h), C.c_int(device))
        if rc:
            raise BoomGpuError(self._lib.boomgpu_last_error(None).decode())
        self.n = 0
        self.p = 0
        self._keep = []

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.boomgpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc:
            raise BoomGpuError("boomgpu error %d: %s" % (rc, self._lib.boomgpu_last_error(self._h).decode()))

    # ---- configuration
    def set_option(self, name, value):
        self._check(self._lib.boomgpu_set_option(self._h, name.encode(), C.c_int64(int(value))))

    def set_stream(self, cuda_stream):
        self._check(self._lib.boomgpu_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def set_row_offset(self, first_global_row):
        self._check(self._lib.boomgpu_set_row_offset(self._h, C.c_uint64(int(first_global_row))))

    def set_logit_mixture(self, mu, sigma, weights):
        mu, sigma, weights = _f64(mu), _f64(sigma), _f64(weights)
        self._check(self._lib.boomgpu_set_logit_mixture(self._h, C.c_int(len(sigma)), _dp(mu), _dp(sigma), _dp(weights)))

    def set_poisson_table(self, nu, offset, weights, mu, sigma, gaussian_cutoff):
        nu = np.ascontiguousarray(nu, dtype=np.int64)
        offset = np.ascontiguousarray(offset, dtype=np.int32)
        weights, mu, sigma = _f64(weights), _f64(mu), _f64(sigma)
        self._check(self._lib.boomgpu_set_poisson_table(
            self._h, C.c_int(len(nu)), nu.ctypes.data_as(c_i64_p), offset.ctypes.data_as(c_i32_p), _dp(weights), _dp(mu),
            _dp(sigma), C.c_int64(int(gaussian_cutoff))))

    # ---- data
    def upload_binomial(self, X, y, ntrials):
        """X: (n, p) float64; a row-strided view (X_wide[:, :p]) is passed as it is, with its leading dimension."""
        X = np.asarray(X)
        if X.dtype == np.float64 and X.ndim == 2 and X.shape[1] and X.strides[1] == 8 and X.strides[0] % 8 == 0 and X.strides[0] >= 8 * X.shape[1]:
            ldx = X.strides[0] // 8
        else:
            X = _f64(X)
            ldx = X.shape[1]
        y, ntrials = _f64(y), _f64(ntrials)
        n, p = X.shape
        self._check(self._lib.boomgpu_upload_binomial(self._h, C.c_int64(n), C.c_int(p), _dp(X),
                                                      C.c_int64(ldx), _dp(y), _dp(ntrials)))
        self.n, self.p = n, p

    def upload_poisson(self, X, y, exposure):
        X, exposure = _f64(X), _f64(exposure)
        y = np.ascontiguousarray(y, dtype=np.int64)
        n, p = X.shape
        self._check(self._lib.boomgpu_upload_poisson(self._h, C.c_int64(n), C.c_int(p), _dp(X), C.c_int64(p),
                                                     y.ctypes.data_as(c_i64_p), _dp(exposure)))
        self.n, self.p = n, p

    def upload_chunked(self, X, y, aux, chunk_rows, poisson=False):
        """boomgpu_upload_begin / _rows / _end over row chunks (what the BOOM adapter does while walking model->dat()).
        poisson: False / 0 binomial rows, True / 1 Poisson rows, 2 plain regression rows (aux ignored)."""
        X, aux = _f64(X), _f64(aux)
        y = np.ascontiguousarray(y, dtype=np.int64 if int(poisson) == 1 else np.float64)
        n, p = X.shape
        self._check(self._lib.boomgpu_upload_begin(self._h, C.c_int(int(poisson)), C.c_int64(n), C.c_int(p)))
        for a in range(0, n, chunk_rows):
            b = min(n, a + chunk_rows)
            xc, yc, ac = np.ascontiguousarray(X[a:b]), np.ascontiguousarray(y[a:b]), np.ascontiguousarray(aux[a:b])
            self._check(self._lib.boomgpu_upload_rows(self._h, C.c_int64(a), C.c_int64(b - a), _dp(xc), C.c_int64(p),
                                                      yc.ctypes.data_as(C.c_void_p), _dp(ac)))
        self._check(self._lib.boomgpu_upload_end(self._h))
        self.n, self.p = n, p

    def adopt_binomial(self, n, p, dX, ldx, dy, dntrials, keepalive=()):
        """dX/dy/dntrials: device pointers (ints), e.g. torch tensor .data_ptr()."""
        self._check(self._lib.boomgpu_adopt_binomial(self._h, C.c_int64(n), C.c_int(p), C.c_void_p(dX), C.c_int64(ldx),
                                                     C.c_void_p(dy), C.c_void_p(dntrials)))
        self.n, self.p = n, p
        self._keep = list(keepalive)

    def adopt_poisson(self, n, p, dX, ldx, dy, dexposure, keepalive=()):
        self._check(self._lib.boomgpu_adopt_poisson(self._h, C.c_int64(n), C.c_int(p), C.c_void_p(dX), C.c_int64(ldx),
                                                    C.c_void_p(dy), C.c_void_p(dexposure)))
        self.n, self.p = n, p
        self._keep = list(keepalive)

    @staticmethod
    def pin_host(array):
        """Page-locks a numpy array's buffer (boomgpu_pin_host); pair with unpin_host."""
        lib = load_library()
        rc = lib.boomgpu_pin_host(C.c_void_p(array.ctypes.data), C.c_uint64(array.nbytes))
        if rc:
            raise BoomGpuError(lib.boomgpu_last_error(None).decode())

    @staticmethod
    def unpin_host(array):
        lib = load_library()
        rc = lib.boomgpu_unpin_host(C.c_void_p(array.ctypes.data))
        if rc:
            raise BoomGpuError(lib.boomgpu_last_error(None).decode())

    # ---- multi-GPU (NCCL bound at run time inside the library)
    @staticmethod
    def comm_unique_id():
        lib = load_library()
        buf = C.create_string_buffer(128)
        rc = lib.boomgpu_comm_unique_id(buf)
        if rc:
            raise BoomGpuError(lib.boomgpu_last_error(None).decode())
        return buf.raw

    def comm_init(self, unique_id, nranks, rank):
        self._check(self._lib.boomgpu_comm_init(self._h, C.c_char_p(unique_id), C.c_int(nranks), C.c_int(rank)))

    def comm_destroy(self):
        self._check(self._lib.boomgpu_comm_destroy(self._h))

    def allreduce(self, dev_ptr, count):
        self._check(self._lib.boomgpu_allreduce(self._h, C.c_void_p(dev_ptr), C.c_int64(int(count))))

    # ---- hot path
    def logit_step(self, beta, clt_threshold, seed, iteration, xtx=None):
        """xtx: optional preallocated (p, p) float64 array; when it is page-locked (pin_host) and at least 1 MB the matrix
        is copied device->host straight into it."""
        p = self.p
        beta = _f64(beta)
        if xtx is None:
            xtx = np.empty((p, p))
        assert xtx.shape == (p, p) and xtx.dtype == np.float64 and xtx.flags.c_contiguous
        xty = np.empty(p)
        ss = C.c_int64()
        self._check(self._lib.boomgpu_logit_step(self._h, _dp(beta), C.c_int(clt_threshold), C.c_uint64(seed),
                                                 C.c_uint64(iteration), _dp(xtx), _dp(xty), C.byref(ss)))
        return xtx, xty, ss.value

    def poisson_step(self, beta, seed, iteration):
        p = self.p
        beta = _f64(beta)
        xtx = np.empty((p, p))
        xty = np.empty(p)
        sc = np.empty(4)
        self._check(self._lib.boomgpu_poisson_step(self._h, _dp(beta), C.c_uint64(seed), C.c_uint64(iteration), _dp(xtx),
                                                   _dp(xty), _dp(sc)))
        return xtx, xty, sc

    def logit_step_active(self, beta, clt_threshold, seed, iteration, active):
        """(G p x k, diag p, xty p, sample_size): the columns `active` of X'WX, its diagonal and X'Wz."""
        p = self.p
        beta = _f64(beta)
        act = np.ascontiguousarray(active, dtype=np.int32)
        k = len(act)
        G, diag, xty = np.empty((p, k)), np.empty(p), np.empty(p)
        ss = C.c_int64()
        self._check(self._lib.boomgpu_logit_step_active(self._h, _dp(beta), C.c_int(clt_threshold), C.c_uint64(seed), C.c_uint64(iteration),
                                                        act.ctypes.data_as(c_i32_p), C.c_int(k), _dp(G), _dp(diag), _dp(xty), C.byref(ss)))
        return G, diag, xty, ss.value

    def poisson_step_active(self, beta, seed, iteration, active):
        p = self.p
        beta = _f64(beta)
        act = np.ascontiguousarray(active, dtype=np.int32)
        k = len(act)
        G, diag, xty, sc = np.empty((p, k)), np.empty(p), np.empty(p), np.empty(4)
        self._check(self._lib.boomgpu_poisson_step_active(self._h, _dp(beta), C.c_uint64(seed), C.c_uint64(iteration),
                                                          act.ctypes.data_as(c_i32_p), C.c_int(k), _dp(G), _dp(diag), _dp(xty), _dp(sc)))
        return G, diag, xty, sc

    def weighted_column(self, j):
        out = np.empty(self.p)
        self._check(self._lib.boomgpu_weighted_column(self._h, C.c_int(int(j)), _dp(out)))
        return out

    def full_statistics(self):
        p = self.p
        xtx, xty = np.empty((p, p)), np.empty(p)
        self._check(self._lib.boomgpu_full_statistics(self._h, _dp(xtx), _dp(xty)))
        return xtx, xty

    def probit_step(self, beta, clt_threshold, seed, iteration, want_xtx=True):
        """(xtx or None, xtz, sample_size) of the probit sibling; want_xtx=False computes X'z alone."""
        p = self.p
        beta = _f64(beta)
        xtx = np.empty((p, p)) if want_xtx else None
        xtz = np.empty(p)
        ss = C.c_int64()
        self._check(self._lib.boomgpu_probit_step(self._h, _dp(beta), C.c_int(clt_threshold), C.c_uint64(seed), C.c_uint64(iteration),
                                                  _dp(xtx) if want_xtx else None, _dp(xtz), C.byref(ss)))
        return xtx, xtz, ss.value

    def probit_draw(self, beta, clt_threshold, seed, iteration):
        beta = _f64(beta)
        out = np.empty(self.n)
        self._check(self._lib.boomgpu_probit_draw(self._h, _dp(beta), C.c_int(clt_threshold), C.c_uint64(seed), C.c_uint64(iteration),
                                                  _dp(out)))
        return out

    # ---- Student-t sibling (TRegressionSampler)
    def upload_regression(self, X, y):
        X, y = _f64(X), _f64(y)
        n, p = X.shape
        self._check(self._lib.boomgpu_upload_regression(self._h, C.c_int64(n), C.c_int(p), _dp(X), C.c_int64(X.strides[0] // 8), _dp(y)))
        self.n, self.p = n, p

    def adopt_regression(self, n, p, dX, ldx, dy, keepalive=()):
        self._check(self._lib.boomgpu_adopt_regression(self._h, C.c_int64(n), C.c_int(p), C.c_void_p(dX), C.c_int64(ldx), C.c_void_p(dy)))
        self.n, self.p = n, p
        self._keep = list(keepalive)

    def student_step(self, beta, sigma, nu, seed, iteration):
        """(xtwx, xtwy, scalars[n, y'Wy, sum w, sum log w]) of TRegressionSampler::impute_latent_data."""
        p = self.p
        beta = _f64(beta)
        xtwx, xtwy, sc = np.empty((p, p)), np.empty(p), np.empty(4)
        self._check(self._lib.boomgpu_student_step(self._h, _dp(beta), C.c_double(sigma), C.c_double(nu), C.c_uint64(seed),
                                                   C.c_uint64(iteration), _dp(xtwx), _dp(xtwy), _dp(sc)))
        return xtwx, xtwy, sc

    def student_step_active(self, beta, sigma, nu, seed, iteration, active):
        """(G p x k, diag, xtwy, scalars) for the column set `active` (p > 64), as logit_step_active."""
        p = self.p
        beta = _f64(beta)
        act = np.ascontiguousarray(active, dtype=np.int32)
        k = len(act)
        G, diag, xty, sc = np.empty((p, k)), np.empty(p), np.empty(p), np.empty(4)
        self._check(self._lib.boomgpu_student_step_active(self._h, _dp(beta), C.c_double(sigma), C.c_double(nu), C.c_uint64(seed),
                                                          C.c_uint64(iteration), act.ctypes.data_as(c_i32_p), C.c_int(k), _dp(G), _dp(diag),
                                                          _dp(xty), _dp(sc)))
        return G, diag, xty, sc

    def student_step_device(self, beta, sigma, nu, seed, iteration, suf_dev_ptr):
        beta = _f64(beta)
        self._check(self._lib.boomgpu_student_step_device(self._h, _dp(beta), C.c_double(sigma), C.c_double(nu), C.c_uint64(seed),
                                                          C.c_uint64(iteration), C.c_void_p(suf_dev_ptr)))

    def student_draw(self, beta, sigma, nu, seed, iteration):
        beta = _f64(beta)
        out = np.empty(self.n)
        self._check(self._lib.boomgpu_student_draw(self._h, _dp(beta), C.c_double(sigma), C.c_double(nu), C.c_uint64(seed),
                                                   C.c_uint64(iteration), _dp(out)))
        return out

    def student_loglike(self, beta, sigma, nu):
        """beta=None reuses the residuals of the previous call (the slice sampler on nu)."""
        out = C.c_double()
        b = None if beta is None else _f64(beta)
        self._check(self._lib.boomgpu_student_loglike(self._h, None if b is None else _dp(b), C.c_double(sigma), C.c_double(nu),
                                                      C.byref(out)))
        return out.value

    def suf_len(self):
        return int(self._lib.boomgpu_suf_len(C.c_int(self.p)))

    def logit_step_device(self, beta, clt_threshold, seed, iteration, suf_dev_ptr):
        beta = _f64(beta)
        self._check(self._lib.boomgpu_logit_step_device(self._h, _dp(beta), C.c_int(clt_threshold), C.c_uint64(seed),
                                                        C.c_uint64(iteration), C.c_void_p(suf_dev_ptr)))

    def poisson_step_device(self, beta, seed, iteration, suf_dev_ptr):
        beta = _f64(beta)
        self._check(self._lib.boomgpu_poisson_step_device(self._h, _dp(beta), C.c_uint64(seed), C.c_uint64(iteration),
                                                          C.c_void_p(suf_dev_ptr)))

    def synchronize(self):
        self._check(self._lib.boomgpu_synchronize(self._h))

    def suf_buffer(self):
        ptr = C.c_void_p()
        self._check(self._lib.boomgpu_suf_buffer(self._h, C.byref(ptr)))
        return ptr.value

    def download(self, src_dev_ptr, count):
        out = np.empty(int(count))
        self._check(self._lib.boomgpu_download(self._h, C.c_void_p(src_dev_ptr), _dp(out), C.c_int64(int(count))))
        return out

    # ---- parity hooks
    def accumulate(self, weight, weighted_value):
        p = self.p
        w, s = _f64(weight), _f64(weighted_value)
        assert len(w) == self.n and len(s) == self.n
        xtx = np.empty((p, p))
        xty = np.empty(p)
        self._check(self._lib.boomgpu_accumulate(self._h, _dp(w), _dp(s), _dp(xtx), _dp(xty)))
        return xtx, xty

    def logit_draw(self, beta, clt_threshold, seed, iteration):
        beta = _f64(beta)
        s = np.empty(self.n)
        w = np.empty(self.n)
        self._check(self._lib.boomgpu_logit_draw(self._h, _dp(beta), C.c_int(clt_threshold), C.c_uint64(seed),
                                                 C.c_uint64(iteration), _dp(s), _dp(w)))
        return s, w

    def poisson_draw(self, beta, seed, iteration):
        beta = _f64(beta)
        out = np.empty((self.n, 6))
        k2 = np.empty((self.n, 2), dtype=np.int32)
        self._check(self._lib.boomgpu_poisson_draw(self._h, _dp(beta), C.c_uint64(seed), C.c_uint64(iteration), _dp(out),
                                                   k2.ctypes.data_as(c_i32_p)))
        return out, k2

    def binomial_loglike(self, beta):
        beta = _f64(beta)
        out = C.c_double()
        self._check(self._lib.boomgpu_binomial_loglike(self._h, _dp(beta), C.byref(out)))
        return out.value

    def poisson_loglike(self, beta):
        beta = _f64(beta)
        out = C.c_double()
        self._check(self._lib.boomgpu_poisson_loglike(self._h, _dp(beta), C.byref(out)))
        return out.value

    def binomial_loglike_derivs(self, beta, log_alpha=0.0):
        beta = _f64(beta)
        ll = C.c_double()
        g, h = np.empty(self.p), np.empty((self.p, self.p))
        self._check(self._lib.boomgpu_binomial_loglike_derivs(self._h, _dp(beta), C.c_double(log_alpha), C.byref(ll), _dp(g), _dp(h)))
        return ll.value, g, h

    def poisson_loglike_derivs(self, beta):
        beta = _f64(beta)
        ll = C.c_double()
        g, h = np.empty(self.p), np.empty((self.p, self.p))
        self._check(self._lib.boomgpu_poisson_loglike_derivs(self._h, _dp(beta), C.byref(ll), _dp(g), _dp(h)))
        return ll.value, g, h

    def select_columns(self, cols):
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        self._check(self._lib.boomgpu_select_columns(self._h, cols.ctypes.data_as(c_i32_p), C.c_int(len(cols))))
        self._k = len(cols)

    def binomial_loglike_derivs_selected(self, beta_selected, log_alpha=0.0):
        b = _f64(beta_selected)
        k = len(b)
        ll = C.c_double()
        g, h = np.empty(k), np.empty((k, k))
        self._check(self._lib.boomgpu_binomial_loglike_derivs_selected(self._h, _dp(b), C.c_double(log_alpha), C.byref(ll), _dp(g), _dp(h)))
        return ll.value, g, h

    def poisson_loglike_derivs_selected(self, beta_selected):
        b = _f64(beta_selected)
        k = len(b)
        ll = C.c_double()
        g, h = np.empty(k), np.empty((k, k))
        self._check(self._lib.boomgpu_poisson_loglike_derivs_selected(self._h, _dp(b), C.byref(ll), _dp(g), _dp(h)))
        return ll.value, g, h

    def poisson_counts_present(self, length):
        out = np.zeros(int(length), dtype=np.uint8)
        self._check(self._lib.boomgpu_poisson_counts_present(self._h, out.ctypes.data_as(C.c_void_p), C.c_int64(int(length))))
        return out

    # ---- instrumentation
    def kernel_launches(self):
        return int(self._lib.boomgpu_kernel_launches(self._h))

    def timings(self, reset=False):
        ms = (C.c_double * NUM_KERNEL_CLASSES)()
        cnt = (C.c_int64 * NUM_KERNEL_CLASSES)()
        self._check(self._lib.boomgpu_get_timings(self._h, ms, cnt, C.c_int(int(reset))))
        return {KERNEL_CLASSES[i]: (ms[i], int(cnt[i])) for i in range(NUM_KERNEL_CLASSES)}
