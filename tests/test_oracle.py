"""CPU tests: pin the oracle port (oracle/tfft_oracle.c) against the reference's own outputs --
the committed golden fixtures (generated from oracle/_ref by tests/golden/make_golden.py) and,
when oracle/_ref is present, the reference TU called live.  SURVEY App. B known answers."""
import numpy as np
import pytest

from oracle import pyoracle as O
from steganosaurus_b200 import synth
from util import golden_cases, load_golden, spec_err


def test_fft1d_known_answers(port):
    # forward is e^{+i}
    got = port.fft1d(np.array([0, 1, 0, 0], np.complex128))
    assert np.allclose(got, [1, 1j, -1, -1j], atol=1e-15)
    got = port.fft1d(np.arange(8, dtype=np.complex128))
    s = 9.6568542494923797
    t = 1.6568542494923797
    want = [28, -4 - s * 1j, -4 - 4j, -4 - t * 1j, -4, -4 + t * 1j, -4 + 4j, -4 + s * 1j]
    assert np.allclose(got, want, atol=1e-13)
    inv = port.fft1d(got, inverse=True)
    assert np.allclose(inv, np.arange(8), atol=1e-14)


def test_fft2d_known_answer(port):
    A = (4 * np.arange(4)[:, None] + np.arange(4)[None, :]).astype(np.complex128)
    F = port.fft2d(A)
    assert np.allclose(F[0], [120, -8 - 8j, -8, -8 + 8j], atol=1e-12)
    assert np.allclose(F[:, 0], [120, -32 - 32j, -32, -32 + 32j], atol=1e-12)
    assert np.allclose(F[1:, 1:], 0, atol=1e-12)


def test_numpy_sign_convention(port):
    img = synth.gen_cover(48, 40, 7)
    a = port.forward_spectrum(img)
    b = O.numpy_forward_spectrum(img)
    assert spec_err(a, b)[0] < 1e-12


def test_read_bit_ties(port):
    # ties read as 1
    for re, im in [(1.0, 0.0), (-1.0, 0.0), (-1.0, -0.0), (0.0, 0.0)]:
        assert port.read_bit(re, im) == 1
    assert port.read_bit(1.0, 1e-3) == 1 and port.read_bit(1.0, -1e-3) == 0
    assert port.read_bit(-1.0, 1e-3) == 1 and port.read_bit(-1.0, -1e-3) == 0


def test_rep_decode(port):
    bits = np.array([1, 1, 0, 0, 0, 1, 1, 0, 1, 0, 0, 0], np.uint8)
    assert port.rep_decode(bits, 3).tolist() == [0b10100000]
    b7 = np.array([1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 1, 1, 1, 0], np.uint8)
    assert port.rep_decode(b7, 7).tolist() == [0b10000000]
    assert port.rep_decode(np.array([1, 0, 1, 1, 0, 0, 0, 1, 1], np.uint8), 1).tolist() == [0b10110001, 0b10000000]


@pytest.mark.parametrize("name", golden_cases())
def test_port_matches_golden(port, name):
    g = load_golden(name)
    F = port.forward_spectrum(g["cover"], g["center"])
    sy, sx = g["spec_sample_yx"]
    rms = np.sqrt(np.mean(np.abs(F) ** 2))
    assert np.abs(F[:, sy, sx] - g["spec_sample"]).max() / rms < 1e-12
    assert np.allclose(np.abs(F).sum(axis=(1, 2)), g["spec_abs_sum"], rtol=1e-12)
    e = port.embed(g["cover"], g["bins"], g["bits"], g["alpha"], g["center"], 0.01, g["rmin"], g["rmax"], want_spectrum=True)
    assert e["rc"] == 0
    assert e["usable"] == g["usable"]
    assert np.allclose(e["medians"], g["medians"], rtol=1e-12)
    assert np.abs(e["spectrum"][:, sy, sx] - g["spec_after_sample"]).max() / rms < 1e-12
    assert np.array_equal(e["stego"], g["stego"])  # bit-exact pixels on these inputs
    _, raw = port.extract(g["stego"], g["bins"], 1, g["alpha"], g["center"])
    assert np.array_equal(raw, g["raw_all"])
    dec3, _ = port.extract(g["stego"], g["bins"][: g["hdr_n"]], 3, g["alpha"], g["center"])
    assert np.array_equal(dec3, g["dec3"])
    rest = g["bins"][g["hdr_n"]:]
    dec7, _ = port.extract(g["stego"], rest[: rest.size // 7 * 7], 7, g["alpha"], g["center"])
    assert np.array_equal(dec7, g["dec7"])


def test_golden_roundtrip_property():
    # pow2 fixtures: the reference's own raw BER is small and the vote repairs it
    g = load_golden("g512_walk")
    assert (g["raw_all"] != g["bits"]).mean() < 0.01


@pytest.mark.parametrize("W,H,nbits,center", [(64, 64, 300, False), (100, 60, 500, True), (256, 128, 900, False)])
def test_port_vs_reference_live(port, ref, W, H, nbits, center):
    PH, PW = synth.next_pow2(H), synth.next_pow2(W)
    cover = synth.gen_texture(W, H, W + H)
    bins = synth.random_bins(PH, PW, nbits, 5)
    bits = synth.random_bits(1, nbits, 6)[0]
    a = port.embed(cover, bins, bits, 0.5, center, want_spectrum=True)
    b = ref.embed(cover, bins, bits, 0.5, center, want_spectrum=True)
    assert a["usable"] == b["usable"]
    assert np.allclose(a["medians"], b["medians"], rtol=1e-12)
    assert spec_err(a["spectrum"], b["spectrum"])[0] < 1e-12
    assert np.array_equal(a["stego"], b["stego"])
    ra = port.extract(b["stego"], bins, 1, 0.5, center)[1]
    rb = ref.extract(b["stego"], bins, 1, 0.5, center)[1]
    assert np.array_equal(ra, rb)


def test_port_fft_vs_reference_live(port, ref):
    rng = np.random.default_rng(3)
    for n in (2, 16, 512, 4096):
        a = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        for inv in (False, True):
            assert np.abs(port.fft1d(a, inv) - ref.fft1d(a, inv)).max() < 1e-11 * np.sqrt(n)


def test_capacity_rc(port):
    cover = synth.gen_cover(64, 64, 1)
    allb = synth.valid_bins(64, 64)
    bits = np.zeros(allb.size, np.uint8)
    e = port.embed(cover, allb, bits)
    assert e["rc"] == 1 and e["usable"] < allb.size  # usable under-counts by 2x (S:1006)


def test_write_read_single_bin(ref, port):
    for re, im in [(3.0, 4.0), (-2.0, 0.5), (0.0, 0.0), (1e-15, 0)]:
        for bit in (0, 1):
            z = ref.write_bit(re, im, bit, 0.5)
            mag = max(1e-12, np.hypot(re, im))
            want = complex(mag * np.cos(0.5), mag * np.sin(0.5) * (1 if bit else -1))
            assert abs(z - want) <= 1e-15 * mag
            assert ref.read_bit(z.real, z.imag, 0.5) == bit == port.read_bit(z.real, z.imag, 0.5)
