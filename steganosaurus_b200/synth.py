"""Seeded synthetic inputs for tests and bench.py (no network: no datasets).

* covers follow the reference's only fixture, tools/gen_png.cpp:8-17 (R = 180 + 40x/W + n,
  G = 180 + 40y/H + n, B = 200 + n, n in [-10, 9]) with a seeded PRNG instead of rand();
* bin lists are either the real keyed turtlewalk (steganosaurus_b200.host) or, for pure
  hot-path tests, a seeded random subset of the valid quarter-annulus bins -- the hot path
  accepts any unique, axis-free, conjugate-free bin list.
"""
from __future__ import annotations

import numpy as np


def next_pow2(v: int) -> int:
    p = 1
    while p < v:
        p <<= 1
    return p


def gen_cover(W: int, H: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = np.arange(W, dtype=np.int64)[None, :]
    y = np.arange(H, dtype=np.int64)[:, None]
    n = rng.integers(-10, 10, size=(H, W), dtype=np.int64)
    img = np.empty((H, W, 3), np.int64)
    img[:, :, 0] = 180 + (x * 40) // W + n
    img[:, :, 1] = 180 + (y * 40) // H + n
    img[:, :, 2] = 200 + n
    return np.clip(img, 0, 255).astype(np.uint8)


def gen_texture(W: int, H: int, seed: int = 0, sigma: float = 40.0) -> np.ndarray:
    """Busier cover (gradient + wide noise, clipped): exercises the clamp path."""
    rng = np.random.default_rng(seed)
    x = np.linspace(0, 255, W)[None, :, None]
    y = np.linspace(0, 255, H)[:, None, None]
    img = 0.5 * x + 0.5 * y + rng.normal(0, sigma, size=(H, W, 3))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def valid_bins(PH: int, PW: int, rmin: float = 0.05, rmax: float = 0.45) -> np.ndarray:
    """All (plane,y,x) the walk may select (Turtle::advance_to_valid, S:793-801), packed."""
    m = min(PH, PW)
    lo, hi = rmin * m, rmax * m
    ymax, xmax = min(PH - 1, int(np.floor(hi))), min(PW - 1, int(np.floor(hi)))
    y, x = np.mgrid[0:ymax + 1, 0:xmax + 1]
    r = np.hypot(y.astype(np.float64), x.astype(np.float64))
    ok = (y != 0) & (x != 0) & (y != PH // 2) & (x != PW // 2) & (r >= lo) & (r <= hi)
    # a bin and its conjugate must not both be listed
    cy, cx = (PH - y) % PH, (PW - x) % PW
    ok &= (y * PW + x) < (cy * PW + cx)
    lin = (y[ok].astype(np.uint64) * PW + x[ok].astype(np.uint64))
    out = np.concatenate([(np.uint64(p) << np.uint64(30)) | lin for p in range(3)])
    return out.astype(np.uint32)


def random_bins(PH: int, PW: int, nbits: int, seed: int = 0, rmin: float = 0.05, rmax: float = 0.45) -> np.ndarray:
    allb = valid_bins(PH, PW, rmin, rmax)
    if nbits > allb.size:
        raise ValueError(f"nbits {nbits} > {allb.size} valid bins")
    rng = np.random.default_rng(seed)
    return allb[rng.permutation(allb.size)[:nbits]].copy()


def random_bits(n: int, nbits: int, seed: int = 0) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 2, size=(n, nbits), dtype=np.uint8)


def frame_len(payload_bytes: int) -> int:
    """nbits = Rep3(38-byte header) + Rep7(ct||tag) = 912 + 56*(len+16)  (S:986-995)."""
    return 912 + 56 * (payload_bytes + 16)
