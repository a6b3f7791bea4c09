"""Python face of the host-side C++ (libtfft_host.so, include/tfft_host.h): KDF, AEAD, turtlewalk,
framing, PNG.  These stay on the CPU by design; they produce the bins/bits the CUDA path consumes.
`embed_image` / `extract_image` mirror the reference's do_embed / do_extract (S:907, S:1112) on top
of a steganosaurus_b200.Context, with walk caching per (pass, dims, params)."""
from __future__ import annotations

import ctypes as C
import os
from functools import lru_cache

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libtfft_host.so")
SYMBOLS = ["tfft_host_sha256", "tfft_host_hmac_sha256", "tfft_host_hkdf_expand", "tfft_host_pbkdf2", "tfft_host_seal",
           "tfft_host_open", "tfft_host_seal_rfc8439", "tfft_host_derive_keys", "tfft_host_turtle_keys", "tfft_host_walk", "tfft_host_jitter",
           "tfft_host_frame_bits", "tfft_host_parse_header", "tfft_host_open_payload", "tfft_host_png_load",
           "tfft_host_png_save", "tfft_hostlib_free", "tfft_host_derive_keys_raw", "tfft_host_frame_bits_key",
           "tfft_host_open_payload_key", "tfft_host_key_decode"]
_lib = None
_u8 = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u32 = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f64 = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not built -- run `python -m steganosaurus_b200.build`")
    L = C.CDLL(LIB_PATH)
    cp, sz, u32, i, d = C.c_char_p, C.c_size_t, C.c_uint32, C.c_int, C.c_double
    L.tfft_host_sha256.argtypes = [cp, sz, _u8]
    L.tfft_host_hmac_sha256.argtypes = [cp, sz, cp, sz, _u8]
    L.tfft_host_hkdf_expand.argtypes = [cp, cp, sz, _u8, sz]
    L.tfft_host_pbkdf2.argtypes = [cp, sz, cp, sz, u32, _u8, sz]
    L.tfft_host_seal.argtypes = [cp, cp, cp, sz, _u8, sz, _u8]
    L.tfft_host_seal_rfc8439.argtypes = [cp, cp, cp, sz, _u8, sz, _u8]
    L.tfft_host_open.argtypes = [cp, cp, cp, sz, _u8, sz, cp]
    L.tfft_host_open.restype = i
    L.tfft_host_derive_keys.argtypes = [cp, sz, cp, u32, _u8, _u8]
    L.tfft_host_turtle_keys.argtypes = [cp, sz, _u8, _u8]
    L.tfft_host_walk.argtypes = [cp, i, i, d, d, d, sz, _u32, C.POINTER(i), C.POINTER(u32), C.c_uint64]
    L.tfft_host_walk.restype = i
    L.tfft_host_jitter.argtypes = [cp, _u32, sz, d, _f64]
    L.tfft_host_frame_bits.argtypes = [cp, sz, cp, u32, cp, sz, _u8, _u8]
    L.tfft_host_frame_bits.restype = sz
    L.tfft_host_parse_header.argtypes = [cp, C.POINTER(u32), _u8, _u8]
    L.tfft_host_parse_header.restype = i
    L.tfft_host_open_payload.argtypes = [cp, sz, u32, cp, _u8, u32]
    L.tfft_host_open_payload.restype = i
    L.tfft_host_png_load.argtypes = [cp, C.POINTER(i), C.POINTER(i)]
    L.tfft_host_png_load.restype = C.POINTER(C.c_uint8)
    L.tfft_host_png_save.argtypes = [cp, _u8, i, i]
    L.tfft_host_png_save.restype = i
    L.tfft_hostlib_free.argtypes = [C.c_void_p]
    L.tfft_host_derive_keys_raw.argtypes = [cp, cp, _u8, _u8]
    L.tfft_host_frame_bits_key.argtypes = [cp, cp, cp, sz, _u8, _u8]
    L.tfft_host_frame_bits_key.restype = sz
    L.tfft_host_open_payload_key.argtypes = [cp, cp, _u8, u32]
    L.tfft_host_open_payload_key.restype = i
    L.tfft_host_key_decode.argtypes = [cp, cp, u32, _u8]
    L.tfft_host_key_decode.restype = i
    _lib = L
    return L


def _buf(n):
    return np.zeros(n, np.uint8)


def sha256(d: bytes) -> bytes:
    o = _buf(32); load().tfft_host_sha256(d, len(d), o); return o.tobytes()


def hmac_sha256(k: bytes, m: bytes) -> bytes:
    o = _buf(32); load().tfft_host_hmac_sha256(k, len(k), m, len(m), o); return o.tobytes()


def hkdf_expand(prk: bytes, info: bytes, L: int) -> bytes:
    o = _buf(L); load().tfft_host_hkdf_expand(prk, info, len(info), o, L); return o.tobytes()


def pbkdf2(pw: bytes, salt: bytes, iters: int, dk: int) -> bytes:
    o = _buf(dk); load().tfft_host_pbkdf2(pw, len(pw), salt, len(salt), iters, o, dk); return o.tobytes()


def seal(key: bytes, nonce: bytes, aad: bytes, pt: bytes, rfc: bool = False):
    """rfc=False: the reference's (non-standard) tag; rfc=True: RFC 8439 tag."""
    d = np.frombuffer(pt, np.uint8).copy() if pt else _buf(1)
    t = _buf(16)
    (load().tfft_host_seal_rfc8439 if rfc else load().tfft_host_seal)(key, nonce, aad, len(aad), d, len(pt), t)
    return d[:len(pt)].tobytes(), t.tobytes()


def open_(key: bytes, nonce: bytes, aad: bytes, ct: bytes, tag: bytes):
    d = np.frombuffer(ct, np.uint8).copy() if ct else _buf(1)
    ok = load().tfft_host_open(key, nonce, aad, len(aad), d, len(ct), tag)
    return bool(ok), d[:len(ct)].tobytes()


def derive_keys(pw: bytes, salt: bytes, iters: int):
    k, n = _buf(32), _buf(12)
    load().tfft_host_derive_keys(pw, len(pw), salt, iters, k, n)
    return k.tobytes(), n.tobytes()


def turtle_keys(pw: bytes):
    pk, sub = _buf(32), _buf(128)
    load().tfft_host_turtle_keys(pw, len(pw), pk, sub)
    return pk.tobytes(), sub.tobytes()


class WalkExhausted(RuntimeError):
    pass


def walk(pw: bytes, PH: int, PW: int, nbits: int, rmin=0.05, rmax=0.45, density=0.7, max_steps=0):
    """(bins u32[nbits], start (plane,y,x), ks_walk.ctr) -- same walk as the reference (S:1071-1097)."""
    _, sub = turtle_keys(pw)
    bins = np.zeros(max(nbits, 1), np.uint32)
    start = (C.c_int * 3)()
    ctr = C.c_uint32()
    rc = load().tfft_host_walk(sub[:32], PH, PW, rmin, rmax, density, nbits, bins, start, C.byref(ctr), max_steps)
    if rc != 0:
        raise WalkExhausted(f"turtlewalk cannot deliver {nbits} bins on {PW}x{PH} (rc={rc})")
    return bins[:nbits].copy(), tuple(start), int(ctr.value)


@lru_cache(maxsize=16)
def cached_walk(pw: bytes, PH: int, PW: int, nbits: int, rmin: float, rmax: float, density: float):
    """The bin list is cover-independent (S:797-799): one walk serves every image of a shape."""
    return walk(pw, PH, PW, nbits, rmin, rmax, density)[0]


def jitter_values(pw: bytes, bins, maxj: float):
    _, sub = turtle_keys(pw)
    bins = np.ascontiguousarray(bins, np.uint32)
    out = np.zeros(max(bins.size, 1), np.float64)
    load().tfft_host_jitter(sub, bins, bins.size, maxj, out)
    return out[:bins.size]


def frame_bits(pw: bytes, salt: bytes, iters: int, secret: bytes):
    n = 912 + 56 * (len(secret) + 16)
    bits, hdr = _buf(n), _buf(38)
    got = load().tfft_host_frame_bits(pw, len(pw), salt, iters, secret, len(secret), bits, hdr)
    assert got == n
    return bits, hdr.tobytes()


# ---- --key path (S:576-591, S:603-662): a 32-byte master key instead of a passphrase; the walk uses
# turtle_keys(master_key) (path_key = SHA256(master_key), S:1036)
def key_decode(key_b64: str, wrap_pass: str = "", iters: int = 600000) -> bytes:
    out = _buf(32)
    rc = load().tfft_host_key_decode(key_b64.encode(), wrap_pass.encode(), iters, out)
    if rc != 1:
        raise ValueError("Failed to decode/unwrap key from --key argument")
    return out.tobytes()


def derive_keys_raw(master: bytes, salt: bytes):
    k, n = _buf(32), _buf(12)
    load().tfft_host_derive_keys_raw(master, salt, k, n)
    return k.tobytes(), n.tobytes()


def frame_bits_key(master: bytes, salt: bytes, secret: bytes):
    n = 912 + 56 * (len(secret) + 16)
    bits, hdr = _buf(n), _buf(38)
    got = load().tfft_host_frame_bits_key(master, salt, secret, len(secret), bits, hdr)
    assert got == n
    return bits, hdr.tobytes()


def open_payload_key(master: bytes, hdr: bytes, payload: bytes, clen: int):
    d = np.frombuffer(payload, np.uint8).copy()
    ok = load().tfft_host_open_payload_key(master, hdr, d, clen)
    return bool(ok), d[:clen].tobytes()


def parse_header(hdr: bytes):
    """-> (rc, clen, salt, nonce); rc 1 = 'Magic not found.', 2 = unsupported version (S:1237-1238)."""
    clen = C.c_uint32()
    salt, nonce = _buf(16), _buf(12)
    rc = load().tfft_host_parse_header(hdr, C.byref(clen), salt, nonce)
    return rc, int(clen.value), salt.tobytes(), nonce.tobytes()


def open_payload(pw: bytes, iters: int, hdr: bytes, payload: bytes, clen: int):
    d = np.frombuffer(payload, np.uint8).copy()
    ok = load().tfft_host_open_payload(pw, len(pw), iters, hdr, d, clen)
    return bool(ok), d[:clen].tobytes()


def png_load(path: str) -> np.ndarray:
    W, H = C.c_int(), C.c_int()
    p = load().tfft_host_png_load(path.encode(), C.byref(W), C.byref(H))
    if not p:
        raise IOError(f"Failed to load {path}")
    try:
        return np.ctypeslib.as_array(p, shape=(H.value, W.value, 3)).copy()
    finally:
        load().tfft_hostlib_free(p)


def png_save(path: str, rgb) -> None:
    rgb = np.ascontiguousarray(rgb, np.uint8)
    H, W, _ = rgb.shape
    if not load().tfft_host_png_save(path.encode(), rgb, W, H):
        raise IOError(f"PNG write failed: {path}")


# ------------------------------------------------------------------------------------------------
class ExtractError(RuntimeError):
    """Carries the reference's message: 'Magic not found.', 'Auth failed (wrong pass or data corrupted).' ..."""


def embed_image(ctx, cover, secret: bytes, pw: bytes, alpha=0.5, density=0.7, rmin=0.05, rmax=0.45, magmin=0.01,
                center=False, pbkdf2_iter=600000, jitter=0.0, salt: bytes | None = None, key: bytes | None = None):
    """do_embed (S:907-1109) for one decoded cover [H,W,3]; returns (stego, nbits).  key (32 bytes): the --key path,
    pw is then ignored."""
    from .api import next_pow2
    H, W, _ = cover.shape
    PH, PW = next_pow2(H), next_pow2(W)
    salt = os.urandom(16) if salt is None else salt  # std::random_device upstream (S:927-929)
    if key is not None:
        bits, _ = frame_bits_key(key, salt, secret)
        pw = key  # path_key = SHA256(master_key) (S:1036)
    else:
        bits, _ = frame_bits(pw, salt, pbkdf2_iter, secret)
    bins = cached_walk(pw, PH, PW, bits.size, rmin, rmax, density)
    jit = jitter_values(pw, bins, jitter) if jitter else None
    stego, usable, _ = ctx.embed_batch(cover[None], bins, bits[None], alpha, center, magmin, rmin, rmax, jitter=jit)
    return stego[0], bits.size


def extract_image(ctx, stego, pw: bytes, alpha=0.5, density=0.7, rmin=0.05, rmax=0.45, center=False, pbkdf2_iter=600000,
                  jitter=0.0, key: bytes | None = None) -> bytes:
    """do_extract (S:1112-1312) for one decoded stego image [H,W,3]; returns the plaintext."""
    from .api import next_pow2
    H, W, _ = stego.shape
    PH, PW = next_pow2(H), next_pow2(W)
    if key is not None:
        pw = key
    ctx.forward_batch(stego[None], center)
    hb = cached_walk(pw, PH, PW, 912, rmin, rmax, density)
    hj = jitter_values(pw, hb, jitter) if jitter else None
    hdr, _ = ctx.read_bits(hb, 3, alpha, jitter=hj, want_raw=False)
    hdr = hdr[0].tobytes()
    rc, clen, _, _ = parse_header(hdr)
    if rc == 1:
        raise ExtractError("Magic not found.")
    if rc == 2:
        raise ExtractError(f"Unsupported version ({hdr[4]}).")
    nb = 912 + 56 * (clen + 16)
    if nb > 3 * PH * PW // 2:  # a length the annulus cannot hold is not a frame (same guard as the CLI and the pipeline)
        raise ExtractError("Payload truncated after ECC decode.")
    try:  # the reference walks forever on a garbage clen (App. D-8); here the walk is bounded
        allb = cached_walk(pw, PH, PW, nb, rmin, rmax, density)
    except WalkExhausted:
        raise ExtractError("Payload truncated after ECC decode.")
    aj = jitter_values(pw, allb, jitter) if jitter else None
    pay, _ = ctx.read_bits(allb[912:], 7, alpha, jitter=None if aj is None else aj[912:], want_raw=False)
    if key is not None:
        ok, pt = open_payload_key(key, hdr, pay[0].tobytes(), clen)
    else:
        ok, pt = open_payload(pw, pbkdf2_iter, hdr, pay[0].tobytes(), clen)
    if not ok:
        raise ExtractError("Auth failed (wrong pass or data corrupted).")
    return pt
