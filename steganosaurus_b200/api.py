"""Python face of the C ABI in include/tfft.h (ctypes; numpy for host buffers, torch tensors
for the *_dev entry points).  Names follow the reference's hot-path vocabulary
(steganosaurus/src/steganosaur.cpp: fft2d S:359, median_abs S:404, write_bit_on_bin S:712,
read_bit_from_bin S:734, rep3/rep7_decode_bits S:468/S:501): bins, bits, planes, spectra.

Nothing here computes: every method is one call into libtfft_b200.so.  If the library is
missing, constructing a Context raises -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

OK, E_INVALID, E_CUDA, E_CAPACITY, E_NOMEM, E_UNSUPPORTED, E_STATE = range(7)

# Params defaults, S:375-381
DEFAULTS = dict(alpha=0.50, rmin=0.05, rmax=0.45, magmin=0.01, density=0.7, jitter=0.0,
                center=False, pbkdf2_iter=600000)


class TfftError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


class CapacityError(TfftError):
    """nbits > usable for at least one image (reference: 'Message too large.' S:1010)."""

    def __init__(self, nbits, usable, stego=None, median=None):
        self.nbits, self.usable, self.stego, self.median = nbits, usable, stego, median
        worst = int(np.min(usable)) if usable is not None and len(usable) else 0
        super().__init__(E_CAPACITY, f"Message too large. Need {nbits} bits (after ECC), capacity ~{worst} bits.")


def next_pow2(v: int) -> int:
    p = 1
    while p < v:
        p <<= 1
    return p


def pack_bins(plane, y, x, PW):
    """plane<<30 | y*PW + x (include/tfft.h)."""
    return ((np.asarray(plane, np.uint64) << 30) | (np.asarray(y, np.uint64) * PW + np.asarray(x, np.uint64))).astype(np.uint32)


def _ptr(a):
    return None if a is None else a.ctypes.data


def _dptr(t):
    return None if t is None else t.data_ptr()


class Context:
    """One per GPU per host thread (tfft_ctx is not thread-safe)."""

    def __init__(self, device: int = 0):
        self.L = _lib.load()
        h = C.c_void_p()
        rc = self.L.tfft_create(device, C.byref(h))
        if rc != OK:
            raise TfftError(rc, f"tfft_create(device={device}): {self.L.tfft_strerror(rc).decode()}")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.tfft_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != OK:
            msg = self.L.tfft_strerror(rc).decode()
            if rc == E_CUDA:
                msg += " -- " + self.L.tfft_last_cuda_error(self.h).decode()
            raise TfftError(rc, msg)

    @property
    def launches(self) -> int:
        return int(self.L.tfft_launch_count(self.h))

    def set_adaptive_alpha(self, on: bool):
        """Params.adaptive_alpha (S:379, S:704-710): alpha scaled by |F| / median per bin in later embeds / extracts."""
        self._check(self.L.tfft_set_adaptive_alpha(self.h, int(on)))

    def set_workspace_limit(self, nbytes: int):
        self._check(self.L.tfft_set_workspace_limit(self.h, nbytes))

    # ------------------------------------------------------------------ host-buffer entry points
    def embed_batch(self, cover, bins, bits, alpha=0.5, center=False, magmin=0.01, rmin=0.05, rmax=0.45,
                    jitter=None, out=None, packed=False):
        """cover u8 [n,H,W,3]; bins u32 [nbits]; bits u8 [n,nbits] (packed: [n, ceil(nbits/8)], MSB first as
        bytes_from_bits S:447) -> (stego, usable[n], median[n,3])."""
        cover = np.ascontiguousarray(cover, np.uint8)
        n, H, W, ch = cover.shape
        assert ch == 3
        bins = np.ascontiguousarray(bins, np.uint32)
        width = (bins.size + 7) // 8 if packed else bins.size
        bits = np.ascontiguousarray(bits, np.uint8).reshape(n, -1) if n else np.zeros((0, width), np.uint8)
        assert bits.shape[1] == width, "bits must be [n, nbits] (packed: [n, ceil(nbits / 8)])"
        jit = None if jitter is None else np.ascontiguousarray(jitter, np.float64)
        stego = np.empty_like(cover) if out is None else out
        usable = np.zeros(max(n, 1), np.uint64)
        median = np.zeros((max(n, 1), 3), np.float64)
        fn = self.L.tfft_embed_batch_packed if packed else self.L.tfft_embed_batch
        rc = fn(self.h, _ptr(cover), n, W, H, _ptr(bins), _ptr(bits), bins.size, _ptr(jit),
                alpha, int(center), magmin, rmin, rmax, _ptr(stego), _ptr(usable), _ptr(median))
        if rc == E_CAPACITY:
            raise CapacityError(bins.size, usable[:n], stego, median[:n])
        self._check(rc)
        return stego, usable[:n], median[:n]

    def extract_bits(self, stego, bins, rep, alpha=0.5, center=False, jitter=None, want_raw=True):
        """stego u8 [n,H,W,3]; bins u32 [nbins] -> (bytes u8 [n, ceil(nbins//rep/8)], raw u8 [n,nbins] | None)."""
        stego = np.ascontiguousarray(stego, np.uint8)
        n, H, W, ch = stego.shape
        assert ch == 3
        bins = np.ascontiguousarray(bins, np.uint32)
        jit = None if jitter is None else np.ascontiguousarray(jitter, np.float64)
        nb = (bins.size // rep + 7) // 8
        out = np.zeros((n, nb), np.uint8)
        raw = np.zeros((n, bins.size), np.uint8) if want_raw else None
        self._check(self.L.tfft_extract_bits(self.h, _ptr(stego), n, W, H, _ptr(bins), bins.size, rep, _ptr(jit),
                                             alpha, int(center), _ptr(out) if nb else None,
                                             _ptr(raw) if (want_raw and bins.size) else None))
        return out, raw

    def extract_frame(self, stego, bins, nhdr_bins=912, alpha=0.5, center=False, jitter=None, want_raw=False):
        """One forward FFT per image; header bins voted rep-3, the rest rep-7 -> (hdr bytes, payload bytes, raw|None)."""
        stego = np.ascontiguousarray(stego, np.uint8)
        n, H, W, _ = stego.shape
        bins = np.ascontiguousarray(bins, np.uint32)
        jit = None if jitter is None else np.ascontiguousarray(jitter, np.float64)
        nb, nbp = (nhdr_bins // 3 + 7) // 8, ((bins.size - nhdr_bins) // 7 + 7) // 8
        hdr = np.zeros((n, nb), np.uint8)
        pay = np.zeros((n, nbp), np.uint8)
        raw = np.zeros((n, bins.size), np.uint8) if want_raw else None
        self._check(self.L.tfft_extract_frame(self.h, _ptr(stego), n, W, H, _ptr(bins), bins.size, nhdr_bins, _ptr(jit),
                                              alpha, int(center), _ptr(hdr), _ptr(pay) if nbp else None, _ptr(raw)))
        return hdr, pay, raw

    def forward_batch(self, img, center=False):
        img = np.ascontiguousarray(img, np.uint8)
        n, H, W, _ = img.shape
        self._check(self.L.tfft_forward_batch(self.h, _ptr(img), n, W, H, int(center)))
        self._res_n = n

    def read_bits(self, bins, rep, alpha=0.5, jitter=None, want_raw=True):
        bins = np.ascontiguousarray(bins, np.uint32)
        jit = None if jitter is None else np.ascontiguousarray(jitter, np.float64)
        n = getattr(self, "_res_n", 0)
        nb = (bins.size // rep + 7) // 8
        out = np.zeros((max(n, 1), nb), np.uint8)
        raw = np.zeros((max(n, 1), bins.size), np.uint8) if want_raw else None
        self._check(self.L.tfft_read_bits(self.h, _ptr(bins), bins.size, rep, _ptr(jit), alpha,
                                          _ptr(out) if nb else None, _ptr(raw) if (want_raw and bins.size) else None))
        return out[:n], (raw[:n] if want_raw else None)

    def forward_spectrum(self, img, center=False):
        """img u8 [H,W,3] -> complex128 [3,PH,PW] (reference sign convention)."""
        img = np.ascontiguousarray(img, np.uint8)
        H, W, _ = img.shape
        out = np.empty((3, next_pow2(H), next_pow2(W)), np.complex128)
        self._check(self.L.tfft_forward_spectrum(self.h, _ptr(img), W, H, int(center), _ptr(out)))
        return out

    def fft2d(self, data, inverse=False):
        """complex128 [n,PH,PW] (or [PH,PW]) -> transformed copy (fft2d S:359 conventions)."""
        a = np.array(data, np.complex128, order="C", copy=True)
        shp = a.shape
        b = a.reshape((-1,) + shp[-2:])
        self._check(self.L.tfft_fft2d(self.h, _ptr(b), b.shape[0], b.shape[1], b.shape[2], int(inverse)))
        return b.reshape(shp)

    # ------------------------------------------------------------------ device-pointer entry points
    @staticmethod
    def _stream():
        import torch
        return torch.cuda.current_stream().cuda_stream

    def embed_batch_dev(self, cover, bins, bits, stego, alpha=0.5, center=False, magmin=0.01, rmin=0.05, rmax=0.45,
                        jitter=None, usable=None, median=None):
        """torch CUDA tensors; enqueues on torch's current stream, does not synchronise."""
        n, H, W, _ = cover.shape
        self._check(self.L.tfft_embed_batch_dev(self.h, _dptr(cover), n, W, H, _dptr(bins), _dptr(bits), bins.numel(),
                                                _dptr(jitter), alpha, int(center), magmin, rmin, rmax, _dptr(stego),
                                                _dptr(usable), _dptr(median), self._stream()))

    def extract_bits_dev(self, stego, bins, rep, out_bytes, raw_bits=None, alpha=0.5, center=False, jitter=None):
        n, H, W, _ = stego.shape
        self._check(self.L.tfft_extract_bits_dev(self.h, _dptr(stego), n, W, H, _dptr(bins), bins.numel(), rep,
                                                 _dptr(jitter), alpha, int(center), _dptr(out_bytes), _dptr(raw_bits),
                                                 self._stream()))

    def extract_frame_dev(self, stego, bins, nhdr_bins, out_hdr, out_payload, raw_bits=None, alpha=0.5, center=False,
                          jitter=None):
        n, H, W, _ = stego.shape
        self._check(self.L.tfft_extract_frame_dev(self.h, _dptr(stego), n, W, H, _dptr(bins), bins.numel(), nhdr_bins,
                                                  _dptr(jitter), alpha, int(center), _dptr(out_hdr), _dptr(out_payload),
                                                  _dptr(raw_bits), self._stream()))

    # ------------------------------------------------------------------ per-kernel timing
    KINDS = ["row_fwd_u8", "col_fwd", "median_capacity", "embed_scatter", "col_inv", "row_inv_u8", "extract_vote", "c2c_pass", "col_fwd_window", "col_embed_fused", "slab_glue"]

    def profile_enable(self, on=True):
        self._check(self.L.tfft_profile_enable(self.h, int(on)))

    def profile_reset(self):
        self._check(self.L.tfft_profile_reset(self.h))

    def profile_read(self):
        """{kind: (groups, total_ms, algorithmic_bytes)} -- call after synchronising the streams used."""
        out = {}
        for k, name in enumerate(self.KINDS):
            g, ms, by = C.c_uint64(), C.c_double(), C.c_double()
            self._check(self.L.tfft_profile_read(self.h, k, C.byref(g), C.byref(ms), C.byref(by)))
            out[name] = (int(g.value), float(ms.value), float(by.value))
        return out

    def fft2d_dev(self, data, inverse=False):
        """complex128 CUDA tensor [n,PH,PW], in place."""
        n, PH, PW = data.shape
        self._check(self.L.tfft_fft2d_dev(self.h, data.data_ptr(), n, PH, PW, int(inverse), self._stream()))

    def fft_pass_dev(self, data, axis, inverse=False):
        n, PH, PW = data.shape
        self._check(self.L.tfft_fft_pass_dev(self.h, data.data_ptr(), n, PH, PW, axis, int(inverse), self._stream()))

    def median_capacity_dev(self, spec, median, usable=None, magmin=0.01, rmin=0.05, rmax=0.45):
        """spec complex128 CUDA [n,3,PH,PW]; median f64 [n,3]; usable u64-as-int64 [n]."""
        n, _, PH, PW = spec.shape
        self._check(self.L.tfft_median_capacity_dev(self.h, spec.data_ptr(), n, PH, PW, magmin, rmin, rmax,
                                                    median.data_ptr(), _dptr(usable), self._stream()))
