"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/tfft.h declares (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols(name):
    src = open(os.path.join(ROOT, "include", name)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tfft_[a-z0-9_]+)\s*\(", src)))


def test_cuda_library_exports_header_symbols():
    from steganosaurus_b200 import _lib
    L = _lib.load()
    syms = header_symbols("tfft.h")
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(L, s), f"libtfft_b200.so lacks {s}"
    assert sorted(_lib.SYMBOLS) == syms
    assert L.tfft_abi_version() == 1


def test_strerror_and_invalid_args_without_gpu():
    from steganosaurus_b200 import _lib
    L = _lib.load()
    assert L.tfft_strerror(0) == b"ok"
    assert b"too large" in L.tfft_strerror(3)
    # null ctx is rejected before any CUDA call
    assert L.tfft_embed_batch(None, None, 0, 0, 0, None, None, 0, None, 0.5, 0, 0.01, 0.05, 0.45, None, None, None) == 1
    assert L.tfft_extract_bits(None, None, 0, 0, 0, None, 0, 3, None, 0.5, 0, None, None) == 1
    assert L.tfft_read_bits(None, None, 0, 3, None, 0.5, None, None) == 1


def test_no_cpu_fallback_in_product():
    """The product package must not import the oracle."""
    pkg = os.path.join(ROOT, "steganosaurus_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "tfft_oracle" not in txt, f
                assert "libtfft_ref" not in txt, f


def test_context_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import steganosaurus_b200 as sb
    with pytest.raises(sb.TfftError):
        sb.Context(0)
