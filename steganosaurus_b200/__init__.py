"""steganosaurus_b200 -- B200-native (sm_100a) implementation of TurtleFFT's spectral hot path.

Layout (only what the path needs):
    csrc/            hand-written CUDA kernels + the C ABI (include/tfft.h), host C++ (csrc/host)
    _lib.py          ctypes loader for libtfft_b200.so (no fallback)
    api.py           Python face of the C ABI
    host.py          Python face of the host-side C++ (KDF, AEAD, turtlewalk, framing)
    build.py         in-tree build (nvcc -gencode arch=compute_100a,code=sm_100a)
"""
from .api import (CapacityError, Context, DEFAULTS, TfftError, next_pow2, pack_bins)  # noqa: F401

__all__ = ["Context", "TfftError", "CapacityError", "DEFAULTS", "next_pow2", "pack_bins"]
