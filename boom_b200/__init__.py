"""boom_b200 -- B200-native auxiliary-mixture Gibbs hot path behind BOOM's sampler surface.

Layers:
  csrc/         CUDA kernels for sm_100a + the C ABI of include/boomgpu.h  -> libboomgpu.so
  capi.py       ctypes binding of that C ABI (what tests and bench.py drive)
  host/         C++ host samplers mirroring BOOM's PosteriorSampler::draw() surface (+ pybind module)
There is no CPU fallback: importing works anywhere, computing needs the CUDA library and a GPU.
"""
from .capi import BoomGpuError, Context, library_path, load_library  # noqa: F401

__all__ = ["BoomGpuError", "Context", "library_path", "load_library"]
