"""BASELINE.json configs as GPU parity tests (C3 is tests/test_gpu_parity.py::test_uhd_* / test_full_size_*; C1 and C2 are
tests/test_cli_crosstool.py):

C4  extract-only sweep: stego images PRODUCED BY THE REFERENCE (oracle embed of a keyed-walk frame at 50 % of the
    capacity) for N in {512, 1024, 2048, 4096}; our voted header / payload bytes and raw bits must equal the oracle's on
    the same image -- whether or not the channel happened to be clean.
C5  slab-decomposed 2-D FFT of one large image over 2+ GPUs: see tests/test_slab_gpu.py.
"""
import numpy as np
import pytest

from oracle import pyoracle as O
from steganosaurus_b200 import host, synth

pytestmark = pytest.mark.gpu
PASS = b"correct horse battery staple"


@pytest.mark.parametrize("N", [512, 1024, 2048, 4096])
def test_c4_extract_only_matches_oracle_on_reference_stego(ctx, N):
    o = O.best()
    cover = synth.gen_cover(N, N, 7)
    cap = o.embed(cover, np.zeros(0, np.uint32), np.zeros(0, np.uint8))["usable"]
    plen = max(16, (cap // 2 - 912) // 56 - 16)  # payload = 50 % of the capacity
    nbits = synth.frame_len(plen)
    bins = host.walk(PASS, N, N, nbits)[0]
    rng = np.random.default_rng(N)
    raw1 = rng.integers(0, 2, size=304 + 8 * (plen + 16), dtype=np.uint8)
    bits = np.concatenate([np.repeat(raw1[:304], 3), np.repeat(raw1[304:], 7)])
    stego = o.embed(cover, bins, bits)["stego"]           # the reference's own embed
    batch = np.stack([stego, cover])                       # a clean cover rides along: garbage, but the same garbage
    hdr, pay, raw = ctx.extract_frame(batch, bins, 912, want_raw=True)
    for i in range(2):
        whdr, wraw_h = o.extract(batch[i], bins[:912], 3)
        wpay, wraw_p = o.extract(batch[i], bins[912:], 7)
        assert np.array_equal(hdr[i], whdr), (N, i)
        assert np.array_equal(pay[i], wpay), (N, i)
        assert np.array_equal(raw[i], np.concatenate([wraw_h, wraw_p])), (N, i)
    # the two-phase flow of do_extract (S:1223-1268: header first, then the payload bins) reads the same bytes
    ctx.forward_batch(batch)
    h2, _ = ctx.read_bits(bins[:912], 3, want_raw=False)
    p2, _ = ctx.read_bits(bins[912:], 7, want_raw=False)
    assert np.array_equal(h2, hdr) and np.array_equal(p2, pay)
    ber = (raw[0] != bits).mean()
    assert ber < 0.02, ber
