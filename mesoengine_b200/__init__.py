"""mesoengine_b200 -- B200-native voxel hot path behind MesoEngine's API surface.

csrc/    hand-written CUDA kernels for sm_100a + the C ABI (include/meso_cuda.h) -> libmeso_b200.so
host/    C++ mirror of the reference's engine-side types and frame loop, calling the C ABI
capi.py  ctypes plumbing used by tests/ and bench.py

No CPU fallback: importing the package needs the built shared library, and creating a Context needs an sm_100 GPU.
"""
from . import capi  # noqa: F401  (fails loudly if libmeso_b200.so is missing)
from .capi import Context, MesoError  # noqa: F401
