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
    "PoissonRegressionSpikeSlabSampler", "WeightedRegSuf", "set_logit_mixture", "set_poisson_mixture_table",
)

__all__ = ["BoomGpuError", "Context", "library_path", "load_library", "host", "load_poisson_mixture_table"] + list(_HOST_NAMES)

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
