"""CPU tests of the host-side C++ (libtfft_host.so) against the reference's own functions
(oracle/_ref, when built) and the known answers of SURVEY App. B / tests/golden/walk_kat.npz."""
import os

import numpy as np
import pytest

from steganosaurus_b200 import host
from util import GOLDEN

PW_CHBS = b"correct horse battery staple"


def test_exports():
    import re
    L = host.load()
    src = open(os.path.join(os.path.dirname(GOLDEN), "..", "include", "tfft_host.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    syms = sorted(set(re.findall(r"\b(tfft_host[a-z0-9_]+)\s*\(", src)))
    assert syms == sorted(host.SYMBOLS)
    for s in syms:
        assert hasattr(L, s)


def test_sha256_kat():
    assert host.sha256(b"abc").hex() == "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"
    assert host.sha256(b"").hex() == "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"
    assert host.sha256(b"pw").hex() == "30c952fab122c3f9759f02a6d95c3758b246b4fee239957b2d4fee46e26170c4"  # App. B path_key


def test_rfc_vectors():
    import hashlib, hmac
    rng = np.random.default_rng(0)
    for n in (0, 1, 55, 56, 63, 64, 65, 200, 1000):
        d = rng.bytes(n)
        assert host.sha256(d) == hashlib.sha256(d).digest()
        k = rng.bytes(int(rng.integers(0, 100)))
        assert host.hmac_sha256(k, d) == hmac.new(k, d, hashlib.sha256).digest()
    assert host.pbkdf2(b"password", b"salt", 4096, 40) == hashlib.pbkdf2_hmac("sha256", b"password", b"salt", 4096, 40)
    # RFC 8439 section 2.8.2 AEAD vector
    key = bytes(range(0x80, 0xa0)); nonce = bytes.fromhex("070000004041424344454647"); aad = bytes.fromhex("50515253c0c1c2c3c4c5c6c7")
    pt = b"Ladies and Gentlemen of the class of '99: If I could offer you only one tip for the future, sunscreen would be it."
    ct, tag = host.seal(key, nonce, aad, pt, rfc=True)
    assert tag.hex() == "1ae10b594f09e26a7e902ecbd0600691"
    assert ct[:16].hex() == "d31a8d34648e60db7b86afbc53ef7ec2"
    # default mode = the reference's tag (S:261-264 recombines limbs without truncation): self-consistent
    ct2, tag2 = host.seal(key, nonce, aad, pt)
    assert ct2 == ct
    ok, back = host.open_(key, nonce, aad, ct2, tag2)
    assert ok and back == pt
    bad = bytes([tag2[0] ^ 1]) + tag2[1:]
    assert not host.open_(key, nonce, aad, ct2, bad)[0]


def test_app_b_key_schedule():
    pk, sub = host.turtle_keys(b"pw")
    assert pk.hex() == "30c952fab122c3f9759f02a6d95c3758b246b4fee239957b2d4fee46e26170c4"
    assert sub[:32].hex() == "2cd8bdcea80a322c1cdae01cde5577fbe97899f56efd8a3fa0dc947d7af48c31"
    assert sub[32:64].hex() == "afab1f4b18cad06f92def0ad37f1c536cd50532605a56f0f83464cc81ca082e2"
    pk, sub = host.turtle_keys(PW_CHBS)
    assert pk.hex() == "c4bbcb1fbec99d65bf59d85c8cb62ee2db963f0fe106f483d9afa73bd4e39a8a"
    assert sub[:32].hex() == "7da48dc044262597287d7431a4995ee17546650ce7c17432130d416abc418c79"
    key, nonce = host.derive_keys(b"pw", bytes(range(16)), 1000)
    assert key.hex() == "56ca31fd1f6086f1ceb95c5f81251fd388f20cb327bba299d95265030aa5e9d0"
    assert nonce.hex() == "182f46212bb3ac59080e0b08"
    bits, hdr = host.frame_bits(b"pw", bytes(range(16)), 1000, b"the eagle has landed")
    assert hdr.hex() == "465454470200" + bytes(range(16)).hex() + "182f46212bb3ac59080e0b08" + "00000014"
    assert bits.size == 2928
    payload = np.packbits(bits[912:].reshape(-1, 7)[:, 0]).tobytes()
    assert payload[:20].hex() == "f230a8b1c52f6d1357f5cd52a21c3fddab54997c" and payload[20:].hex() == "27588da9565b9bbc279189a1d549e7fe"
    assert np.array_equal(np.packbits(bits[:912].reshape(-1, 3)[:, 0]).tobytes(), hdr)
    assert host.parse_header(hdr)[:2] == (0, 20)
    ok, pt = host.open_payload(b"pw", 1000, hdr, payload, 20)
    assert ok and pt == b"the eagle has landed"
    assert not host.open_payload(b"pw", 999, hdr, payload, 20)[0]       # iterations differ -> auth fails (S:1308)
    assert host.parse_header(b"XTTG" + hdr[4:])[0] == 1 and host.parse_header(hdr[:4] + b"\x03" + hdr[5:])[0] == 2


def test_walk_known_answers():
    z = np.load(os.path.join(GOLDEN, "walk_kat.npz"))
    for pw, tag in ((b"pw", "pw"), (PW_CHBS, "chbs")):
        for n in (512, 4096):
            bins, start, ctr = host.walk(pw, n, n, 2928)
            assert tuple(z[f"{tag}_{n}_start"]) == start
            assert int(z[f"{tag}_{n}_ctr"]) == ctr
            assert np.array_equal(bins[:64], z[f"{tag}_{n}_bins64"])
            hv = 0xcbf29ce484222325
            p, y, x = bins >> 30, (bins & 0x3FFFFFFF) // n, (bins & 0x3FFFFFFF) % n
            for b in np.stack([p, y, x], 1).astype("<u4").tobytes():
                hv = ((hv ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
            assert hv == int(z[f"{tag}_{n}_fnv"])
    # SURVEY App. B literal values
    bins, start, ctr = host.walk(b"pw", 512, 512, 2928)
    assert start == (0, 109, 332) and ctr == 782
    first = [(int(b >> 30), int((b & 0x3FFFFFFF) // 512), int((b & 0x3FFFFFFF) % 512)) for b in bins[:4]]
    assert first == [(2, 1, 230), (2, 2, 229), (2, 4, 230), (2, 4, 229)]


@pytest.mark.parametrize("PH,PW,nbits,rmin,rmax,density", [(256, 256, 2480, 0.05, 0.45, 0.7), (512, 1024, 20000, 0.1, 0.3, 0.5),
                                                            (2048, 2048, 100000, 0.05, 0.45, 0.7), (128, 64, 300, 0.05, 0.45, 0.9)])
def test_walk_matches_reference(ref, PH, PW, nbits, rmin, rmax, density):
    for pw in (b"pw", b"another pass phrase"):
        want, wstart, wctr = ref.walk(pw, PH, PW, nbits, rmin, rmax, density)
        got, start, ctr = host.walk(pw, PH, PW, nbits, rmin, rmax, density)
        assert start == wstart and ctr == wctr
        assert np.array_equal(got, want)
        assert np.unique(got).size == nbits  # no bin twice


def test_walk_exhaustion_is_bounded():
    with pytest.raises(host.WalkExhausted):
        host.walk(b"pw", 64, 64, 5000, max_steps=2_000_000)  # more bins than the annulus holds: the reference spins forever


def test_primitives_match_reference(ref):
    rng = np.random.default_rng(1)
    for _ in range(5):
        pw, salt = rng.bytes(int(rng.integers(1, 40))), rng.bytes(16)
        assert host.derive_keys(pw, salt, 37) == ref.derive_keys(pw, salt, 37)
        assert host.turtle_keys(pw) == ref.turtle_keys(pw)
        key, nonce, aad, pt = rng.bytes(32), rng.bytes(12), rng.bytes(int(rng.integers(0, 60))), rng.bytes(int(rng.integers(0, 300)))
        assert host.seal(key, nonce, aad, pt) == ref.seal(key, nonce, aad, pt)
        secret = rng.bytes(int(rng.integers(1, 100)))
        hb, hh = host.frame_bits(pw, salt, 11, secret)
        rb, rh = ref.frame_bits(pw, salt, 11, secret)
        assert hh == rh and np.array_equal(hb, rb)
        assert host.hkdf_expand(rng.bytes(32), b"turtle_keys", 128) is not None
    prk = rng.bytes(32)
    assert host.hkdf_expand(prk, b"info", 100) == ref.hkdf_expand(prk, b"info", 100)
    assert host.pbkdf2(b"p", b"s" * 16, 1000, 44) == ref.pbkdf2(b"p", b"s" * 16, 1000, 44)


def test_png_roundtrip_and_pil_interop(tmp_path):
    from PIL import Image
    rng = np.random.default_rng(2)
    for (H, W) in ((1, 1), (7, 5), (64, 48), (300, 211)):
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        p = str(tmp_path / f"a_{H}x{W}.png")
        host.png_save(p, img)
        assert np.array_equal(host.png_load(p), img)
        assert np.array_equal(np.asarray(Image.open(p).convert("RGB")), img)  # a standard decoder reads our files
    # files written by a standard encoder in other colour types decode like stbi_load(...,3)
    base = rng.integers(0, 256, (33, 21, 4), dtype=np.uint8)
    Image.fromarray(base, "RGBA").save(str(tmp_path / "rgba.png"))
    assert np.array_equal(host.png_load(str(tmp_path / "rgba.png")), base[:, :, :3])           # alpha dropped
    Image.fromarray(base[:, :, 0], "L").save(str(tmp_path / "gray.png"))
    assert np.array_equal(host.png_load(str(tmp_path / "gray.png")), np.repeat(base[:, :, :1], 3, 2))
    Image.fromarray(base[:, :, :3], "RGB").quantize(16).save(str(tmp_path / "pal.png"))
    assert np.array_equal(host.png_load(str(tmp_path / "pal.png")), np.asarray(Image.open(str(tmp_path / "pal.png")).convert("RGB")))
    Image.fromarray(base[:, :, :3], "RGB").save(str(tmp_path / "inter.png"), interlace=1) if False else None
    g16 = (rng.integers(0, 65536, (9, 13))).astype(np.uint16)
    Image.fromarray(g16, "I;16").save(str(tmp_path / "g16.png"))
    assert np.array_equal(host.png_load(str(tmp_path / "g16.png"))[:, :, 0], (g16 >> 8).astype(np.uint8))
    with pytest.raises(IOError):
        host.png_load(str(tmp_path / "missing.png"))


def test_key_path_against_reference_cli(tmp_path):
    """--key (S:576-591, S:603-662, S:1020-1040): the reference CLI embeds with a raw / a passphrase-wrapped master key;
    our host functions (key decode, path keys, AEAD keys) plus the reference's own hot path read the message back."""
    import subprocess
    from oracle import pyoracle as O
    from steganosaurus_b200 import synth
    if not (O.have_ref() and os.path.exists(O.REF_CLI)):
        pytest.skip("oracle/_ref not built")
    r = O.ref()
    cover = str(tmp_path / "c.png")
    host.png_save(cover, synth.gen_cover(256, 256, 3))
    raw_b64 = "AAECAwQFBgcICQoLDA0ODxAREhMUFRYXGBkaGxwdHh8="  # bytes(range(32))
    assert host.key_decode(raw_b64) == bytes(range(32))
    with pytest.raises(ValueError):
        host.key_decode("not base64 !!")
    with pytest.raises(ValueError):
        host.key_decode("QUJD")  # 3 bytes
    # a wrapped key made by the reference's gen-key
    kf = str(tmp_path / "k.b64")
    p = subprocess.run([O.REF_CLI, "gen-key", "--key-out", kf, "--wrap-pass", "wrapme", "--pbkdf2_iter", "1000"],
                       capture_output=True, text=True, timeout=60)
    assert p.returncode == 0, p.stderr
    shown = [ln.split("Base64:")[1].strip() for ln in p.stdout.splitlines() if "Base64:" in ln][0]
    wrapped = open(kf).read().strip()
    master = host.key_decode(wrapped, "wrapme", 1000)
    assert master == host.key_decode(shown)
    with pytest.raises(ValueError):
        host.key_decode(wrapped, "wrong", 1000)
    with pytest.raises(ValueError):
        host.key_decode(wrapped, "", 1000)
    for key_b64, extra, key in ((raw_b64, [], bytes(range(32))), (wrapped, ["--wrap-pass", "wrapme", "--pbkdf2_iter", "1000"], master)):
        s = str(tmp_path / "s.png")
        msg = b"keyed message"
        p = subprocess.run([O.REF_CLI, "embed", "--in", cover, "--out", s, "--secret", msg.decode(), "--key", key_b64, *extra],
                           capture_output=True, text=True, timeout=60)
        assert p.returncode == 0, p.stderr
        stego = host.png_load(s)
        nb = 912 + 56 * (len(msg) + 16)
        bins = host.walk(key, 256, 256, nb)[0]  # path_key = SHA256(master_key)
        _, raw = r.extract(stego, bins, 1)
        hdr = r.rep_decode(raw[:912], 3).tobytes()
        rc, clen, salt, nonce = host.parse_header(hdr)
        assert (rc, clen) == (0, len(msg))
        assert host.derive_keys_raw(key, salt)[1] == nonce
        ok, pt = host.open_payload_key(key, hdr, r.rep_decode(raw[912:], 7).tobytes(), clen)
        assert ok and pt == msg
        # and our framing with the header's salt reproduces the embedded bit stream's header
        bits, hdr2 = host.frame_bits_key(key, salt, msg)
        assert hdr2 == hdr and bits.size == nb


def test_png_load_rejects_hostile_headers(tmp_path):
    import struct, zlib
    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
    sig = b"\x89PNG\r\n\x1a\n"
    huge = sig + chunk(b"IHDR", struct.pack(">IIBBBBB", 2**31 - 1, 2**31 - 1, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(b"\0" * 16)) + chunk(b"IEND", b"")
    early = sig + chunk(b"IDAT", zlib.compress(b"\0" * 16)) + chunk(b"IHDR", struct.pack(">IIBBBBB", 2, 2, 8, 2, 0, 0, 0)) + chunk(b"IEND", b"")
    for i, blob in enumerate((huge, early)):
        p = str(tmp_path / f"bad{i}.png")
        open(p, "wb").write(blob)
        with pytest.raises(IOError):
            host.png_load(p)
