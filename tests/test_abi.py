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


def test_the_two_libraries_share_no_symbol():
    """libtfft_b200.so (tfft_host_alloc / tfft_host_free = pinned CUDA memory) and libtfft_host.so (malloc'd PNG pixels,
    tfft_hostlib_free) are linked into one process by the CLI: a shared name would make the free() that runs depend on
    link order."""
    import subprocess
    from steganosaurus_b200 import _lib, host
    def exported(path):
        out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
        return {ln.split()[-1] for ln in out.splitlines() if " T " in ln and ln.split()[-1].startswith("tfft_")}
    a, b = exported(_lib.LIB_PATH), exported(host.LIB_PATH)
    assert a and b and not (a & b), a & b


def test_strerror_and_invalid_args_without_gpu():
    from steganosaurus_b200 import _lib
    L = _lib.load()
    assert L.tfft_strerror(0) == b"ok"
    assert b"too large" in L.tfft_strerror(3)
    # null ctx is rejected before any CUDA call
    assert L.tfft_embed_batch(None, None, 0, 0, 0, None, None, 0, None, 0.5, 0, 0.01, 0.05, 0.45, None, None, None) == 1
    assert L.tfft_embed_batch_packed(None, None, 0, 0, 0, None, None, 0, None, 0.5, 0, 0.01, 0.05, 0.45, None, None, None) == 1
    assert L.tfft_extract_bits(None, None, 0, 0, 0, None, 0, 3, None, 0.5, 0, None, None) == 1
    assert L.tfft_read_bits(None, None, 0, 3, None, 0.5, None, None) == 1


def test_bin_window_host_logic():
    """The window of the stored spectrum an extract sizes its forward column pass with (host-only entry point):
    largest row / column as listed, or through the Hermitian mirror for bins right of the Nyquist column."""
    import numpy as np
    from steganosaurus_b200 import _lib, synth
    L = _lib.load()

    def window(bins, W, H, half):
        r, c, m = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        b = np.ascontiguousarray(bins, np.uint32)
        rc = L.tfft_bin_window(b.ctypes.data, b.size, W, H, half, ctypes.byref(r), ctypes.byref(c), ctypes.byref(m))
        return rc, r.value, c.value, m.value

    def want(bins, PW, PH, half):
        lin = bins & np.uint32(0x3FFFFFFF)
        y, x = (lin // PW).astype(np.int64), (lin % PW).astype(np.int64)
        mir = half and bool((x > PW // 2).any())
        if mir:
            far = x > PW // 2
            y = np.where(far, (PH - y) % PH, y)
            x = np.where(far, PW - x, x)
        return int(y.max()) + 1, int(x.max()) + 1, int(mir)

    # the reference's walk stays in the annulus corner: 0.45 * 4096 = 1843 -> rows / cols 1844, first 8 of 16 row blocks
    bins = synth.random_bins(4096, 4096, 50000, 3)
    assert window(bins, 3840, 2160, 1) == (0, *want(bins, 4096, 4096, True))
    rc, r, c, m = window(bins, 3840, 2160, 1)
    assert r <= 1844 and c <= 1844 and m == 0
    rng = np.random.default_rng(5)
    for (W, H) in ((600, 4096), (1000, 300), (64, 64)):
        PW, PH = synth.next_pow2(W), synth.next_pow2(H)
        for _ in range(5):
            n = int(rng.integers(1, 2000))
            y = rng.integers(0, PH, n).astype(np.uint32)
            x = rng.integers(0, rng.integers(1, PW + 1), n).astype(np.uint32)
            b = (rng.integers(0, 3, n).astype(np.uint32) << np.uint32(30)) | (y * np.uint32(PW) + x)
            for half in (0, 1):
                assert window(b, W, H, half) == (0, *want(b, PW, PH, bool(half))), (W, H, half)
    assert window(np.zeros(0, np.uint32), 64, 64, 1) == (0, 0, 0, 0)
    assert window(np.array([3 << 30], np.uint32), 64, 64, 1)[0] == 1            # plane 3
    assert window(np.array([64 * 64], np.uint32), 64, 64, 0)[0] == 1            # outside the plane
    assert window(np.array([5], np.uint32), 20000, 64, 0)[0] == 5               # padded width above TFFT_MAX_DIM


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
