"""boom_b200 -- B200-native auxiliary-mixture Gibbs hot path behind BOOM's sampler surface.

Layers:
  csrc/         CUDA kernels for sm_100a + the C ABI of include/boomgpu.h  -> libboomgpu.so
  capi.py       ctypes binding of that C ABI (parity tests and the kernel-level bench drive it)
  host/         C++ samplers with BOOM's PosteriorSampler::draw() / Model::set_method() surface
                -> _host (pybind11), re-exported here: BinomialLogitModel, MvnModel, ...
  data/         the Poisson mixture table as serialized by the reference (a data fixture)
There is no CPU fallback: importing works anywhere, computing needs the CUDA library and a GPU.
"""
import json
import os

import numpy as np

from .capi import BoomGpuError, Context, library_path, load_library  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_HOST_NAMES = (
    "RNG", "global_rng", "MvnModel", "MvnBase", "VariableSelectionPrior", "BinomialLogitModel", "PoissonRegressionModel",
    "PosteriorSampler", "BinomialLogitAuxmixSampler", "BinomialLogitSpikeSlabSampler", "PoissonRegressionAuxMixSampler",
    "PoissonRegressionSpikeSlabSampler", "BinomialProbitModel", "BinomialProbitSpikeSlabSampler", "TRegressionModel", "TRegressionSampler", "TRegressionSpikeSlabSampler",
    "DoubleModel", "UniformModel", "GammaModelBase", "GammaModel", "ChisqModel", "WeightedRegSuf", "set_logit_mixture", "set_poisson_mixture_table",
)

__all__ = ["BoomGpuError", "Context", "library_path", "load_library", "host", "load_poisson_mixture_table", "default_logit_mixture",
           "poisson_mixture_table_arrays"] + list(_HOST_NAMES)

_host_module = None


def host():
    """The pybind11 module of the C++ host samplers (fails loudly if it has not been built)."""
    global _host_module
    if _host_module is None:
        load_library()  # libboomgpu.so first: _host links against it
        import importlib
        try:
            _host_module = importlib.import_module("boom_b200._host")
        except ImportError as e:
            raise BoomGpuError("boom_b200._host is not built (make -C boom_b200/host, or __graft_entry__.build()): %s" % e)
    return _host_module


def __getattr__(name):
    if name in _HOST_NAMES:
        if name.startswith("PoissonRegression") and name.endswith("Sampler") and not host().poisson_mixture_table_is_set():
            load_poisson_mixture_table()   # the reference's table ships with the package: install it on first use
        return getattr(host(), name)
    raise AttributeError(name)


def load_poisson_mixture_table(path=None):
    """Installs the Poisson mixture table (NormalMixtureApproximationTable::serialize() format) for the
    PoissonRegression samplers; default: the table dumped from the reference, every count 1..300 materialised."""
    path = path or os.path.join(_HERE, "data", "poisson_mixture_table.json")
    with open(path) as f:
        g = json.load(f)
    host().set_poisson_mixture_table(np.asarray(g["serialized"], dtype=np.float64), int(g["largest_index"]))
    return path


def default_logit_mixture():
    """(mu, sigma, weights) of the 9-component logistic scale mixture the reference hard-codes
    (Models/Glm/PosteriorSamplers/NormalMixtureApproximation.cpp:416-424): what Context.set_logit_mixture takes."""
    sigma = np.array([0.88437229872213, 1.16097607474416, 1.28021991084306, 1.3592552924727, 1.67589879794907,
                      2.20287232043947, 2.20507148325819, 2.91944313615144, 3.90807611741308])
    weights = np.array([0.038483985581272, 0.13389889791451, 0.0657842076622429, 0.105680086433879, 0.345939491553619,
                        0.0442261124345564, 0.193289780660134, 0.068173066865908, 0.00452437089387876])
    return np.zeros(9), sigma, weights


def poisson_mixture_table_arrays(path=None):
    """The Poisson mixture table (data/poisson_mixture_table.json, NormalMixtureApproximationTable::serialize() layout
    [nu, K, w[K], sigma[K], mu[K]] ...) as the arrays Context.set_poisson_table takes:
    (nu, offset, weights, mu, sigma, gaussian_cutoff)."""
    path = path or os.path.join(_HERE, "data", "poisson_mixture_table.json")
    with open(path) as f:
        g = json.load(f)
    s = np.asarray(g["serialized"], dtype=np.float64)
    nu, off, w, sig, mu = [], [0], [], [], []
    i = 0
    while i < len(s):
        K = int(round(s[i + 1]))
        nu.append(int(round(s[i])))
        w.extend(s[i + 2:i + 2 + K]); sig.extend(s[i + 2 + K:i + 2 + 2 * K]); mu.extend(s[i + 2 + 2 * K:i + 2 + 3 * K])
        off.append(off[-1] + K)
        i += 2 + 3 * K
    return (np.asarray(nu, dtype=np.int64), np.asarray(off, dtype=np.int32), np.asarray(w), np.asarray(mu), np.asarray(sig),
            int(g["largest_index"]))
