"""opencalibration_b200 -- B200-native (sm_100a) matching + RANSAC-scoring hot path of jkflying/opencalibration.

Only what the path needs lives here:
  csrc/    hand-written CUDA kernels (K1 Hamming top-2, K2/K3 MSAC scoring) and the C-ABI layer -> libocb.so
  host/    C++ mirror of the reference's entry points (match_features_subset, ransac<Model>, ...) -> libocb_host.so
  capi.py  ctypes binding of include/ocb.h (what a Python caller uses; fails loudly without the CUDA library)
  host.py  ctypes binding of the C++ mirror's flat test shim
  build.py in-tree nvcc / g++ build of both libraries
There is no CPU fallback in this package.
"""
from .capi import OcbError, lib_path  # noqa: F401

__version__ = "0.1"
