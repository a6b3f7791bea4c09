"""ctypes loader for the in-tree CUDA library (libtfft_b200.so).

There is deliberately NO fallback: if the extension is missing or cannot be loaded the
import of the device API fails loudly (the product path is the CUDA path or nothing).
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TFFT_LIB") or os.path.join(PKG, "libtfft_b200.so")  # TFFT_LIB: experiment builds of the same CUDA library

# every symbol include/tfft.h declares
SYMBOLS = [
    "tfft_create", "tfft_destroy", "tfft_abi_version", "tfft_strerror", "tfft_last_cuda_error",
    "tfft_set_workspace_limit", "tfft_set_adaptive_alpha", "tfft_host_alloc", "tfft_host_free", "tfft_launch_count",
    "tfft_embed_batch", "tfft_embed_batch_packed", "tfft_embed_batch_dev", "tfft_extract_bits", "tfft_extract_bits_dev",
    "tfft_forward_batch", "tfft_read_bits", "tfft_forward_spectrum", "tfft_fft2d", "tfft_fft2d_dev",
    "tfft_fft_pass_dev", "tfft_median_capacity_dev", "tfft_extract_frame", "tfft_extract_frame_dev",
    "tfft_profile_enable", "tfft_profile_reset", "tfft_profile_read", "tfft_kind_name", "tfft_bin_window",
    "tfft_slab_sizes", "tfft_slab_rows_forward_dev", "tfft_slab_cols_dev", "tfft_slab_embed_dev", "tfft_slab_read_dev",
    "tfft_slab_rows_inverse_dev", "tfft_ipc_export", "tfft_ipc_open", "tfft_ipc_close",
]

_lib = None


class TfftLibraryMissing(ImportError):
    pass


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TfftLibraryMissing(
            f"{LIB_PATH} not built -- run `python -m steganosaurus_b200.build` (needs nvcc); "
            "there is no CPU fallback for the hot path")
    L = C.CDLL(LIB_PATH)
    vp, i, d, sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
    L.tfft_create.argtypes = [i, C.POINTER(vp)]
    L.tfft_create.restype = i
    L.tfft_destroy.argtypes = [vp]
    L.tfft_destroy.restype = None
    L.tfft_abi_version.restype = i
    L.tfft_strerror.argtypes = [i]
    L.tfft_strerror.restype = C.c_char_p
    L.tfft_last_cuda_error.argtypes = [vp]
    L.tfft_last_cuda_error.restype = C.c_char_p
    L.tfft_set_workspace_limit.argtypes = [vp, sz]
    L.tfft_set_workspace_limit.restype = i
    L.tfft_set_adaptive_alpha.argtypes = [vp, i]
    L.tfft_set_adaptive_alpha.restype = i
    L.tfft_host_alloc.argtypes = [sz]
    L.tfft_host_alloc.restype = vp
    L.tfft_host_free.argtypes = [vp]
    L.tfft_host_free.restype = None
    L.tfft_launch_count.argtypes = [vp]
    L.tfft_launch_count.restype = C.c_uint64
    emb = [vp, vp, i, i, i, vp, vp, sz, vp, d, i, d, d, d, vp, vp, vp]
    L.tfft_embed_batch.argtypes = emb
    L.tfft_embed_batch.restype = i
    L.tfft_embed_batch_packed.argtypes = emb
    L.tfft_embed_batch_packed.restype = i
    L.tfft_embed_batch_dev.argtypes = emb + [vp]
    L.tfft_embed_batch_dev.restype = i
    ext = [vp, vp, i, i, i, vp, sz, i, vp, d, i, vp, vp]
    L.tfft_extract_bits.argtypes = ext
    L.tfft_extract_bits.restype = i
    L.tfft_extract_bits_dev.argtypes = ext + [vp]
    L.tfft_extract_bits_dev.restype = i
    frm = [vp, vp, i, i, i, vp, sz, sz, vp, d, i, vp, vp, vp]
    L.tfft_extract_frame.argtypes = frm
    L.tfft_extract_frame.restype = i
    L.tfft_extract_frame_dev.argtypes = frm + [vp]
    L.tfft_extract_frame_dev.restype = i
    L.tfft_profile_enable.argtypes = [vp, i]
    L.tfft_profile_enable.restype = i
    L.tfft_profile_reset.argtypes = [vp]
    L.tfft_profile_reset.restype = i
    L.tfft_profile_read.argtypes = [vp, i, C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.tfft_profile_read.restype = i
    L.tfft_kind_name.argtypes = [i]
    L.tfft_kind_name.restype = C.c_char_p
    L.tfft_forward_batch.argtypes = [vp, vp, i, i, i, i]
    L.tfft_forward_batch.restype = i
    L.tfft_read_bits.argtypes = [vp, vp, sz, i, vp, d, vp, vp]
    L.tfft_read_bits.restype = i
    L.tfft_forward_spectrum.argtypes = [vp, vp, i, i, i, vp]
    L.tfft_forward_spectrum.restype = i
    L.tfft_fft2d.argtypes = [vp, vp, i, i, i, i]
    L.tfft_fft2d.restype = i
    L.tfft_fft2d_dev.argtypes = [vp, vp, i, i, i, i, vp]
    L.tfft_fft2d_dev.restype = i
    L.tfft_fft_pass_dev.argtypes = [vp, vp, i, i, i, i, i, vp]
    L.tfft_fft_pass_dev.restype = i
    L.tfft_median_capacity_dev.argtypes = [vp, vp, i, i, i, d, d, d, vp, vp, vp]
    L.tfft_median_capacity_dev.restype = i
    L.tfft_bin_window.argtypes = [vp, sz, i, i, i, C.POINTER(i), C.POINTER(i), C.POINTER(i)]
    L.tfft_bin_window.restype = i
    ip = C.POINTER(i)
    L.tfft_slab_sizes.argtypes = [i, i, i, ip, ip, ip, ip, ip]
    L.tfft_slab_sizes.restype = i
    L.tfft_slab_rows_forward_dev.argtypes = [vp, vp, i, i, i, i, i, i, C.POINTER(vp), sz, i, vp]
    L.tfft_slab_rows_forward_dev.restype = i
    L.tfft_slab_cols_dev.argtypes = [vp, vp, i, i, i, i, vp]
    L.tfft_slab_cols_dev.restype = i
    L.tfft_slab_embed_dev.argtypes = [vp, vp, i, i, i, i, vp, vp, sz, d, vp]
    L.tfft_slab_embed_dev.restype = i
    L.tfft_slab_read_dev.argtypes = [vp, vp, i, i, i, i, vp, sz, d, vp, vp]
    L.tfft_slab_read_dev.restype = i
    L.tfft_slab_rows_inverse_dev.argtypes = [vp, vp, i, i, i, i, i, i, vp, vp]
    L.tfft_slab_rows_inverse_dev.restype = i
    L.tfft_ipc_export.argtypes = [vp, C.c_char_p]
    L.tfft_ipc_export.restype = i
    L.tfft_ipc_open.argtypes = [i, C.c_char_p, C.POINTER(vp)]
    L.tfft_ipc_open.restype = i
    L.tfft_ipc_close.argtypes = [vp]
    L.tfft_ipc_close.restype = i
    _lib = L
    return L
