"""Every environment switch include/tfft.h documents selects a product path; each one is checked against the oracle here on
two cases (a 4096-row half-spectrum plane, where the fused embed / bin window / sign map live, and an ordinary one)."""
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from steganosaurus_b200 import synth
import steganosaurus_b200 as sb

pytestmark = pytest.mark.gpu

SWITCHES = [("TFFT_FFT_IMPL", "v0"), ("TFFT_SPECTRUM", "full"), ("TFFT_WIDE", "0"), ("TFFT_FUSED_EMBED", "0"),
            ("TFFT_EXTRACT_WINDOW", "0"), ("TFFT_SIGNMAP", "0"), ("TFFT_HOST_CHUNK", "1"), ("TFFT_HOST_SLOTS", "1")]


@pytest.mark.parametrize("W,H,n", [(600, 2160, 2), (1024, 700, 3), (8192, 520, 1)])
@pytest.mark.parametrize("var,val", SWITCHES, ids=[f"{k}={v}" for k, v in SWITCHES])
def test_switch_keeps_oracle_parity(var, val, W, H, n):
    if W == 8192 and var != "TFFT_WIDE":
        pytest.skip("the 8192-pixel-row case only exercises TFFT_WIDE")
    o = O.best()
    PH, PW = synth.next_pow2(H), synth.next_pow2(W)
    nbits = 9000
    covers = np.stack([synth.gen_texture(W, H, 11 * W + H + i) for i in range(n)])
    bins = synth.random_bins(PH, PW, nbits, 5)
    bits = synth.random_bits(n, nbits, 6)
    os.environ[var] = val
    try:
        with sb.Context(0) as c:
            stego, usable, med = c.embed_batch(covers, bins, bits)
            dec, raw = c.extract_bits(stego, bins[: nbits // 7 * 7], 7)
    finally:
        os.environ.pop(var, None)
    for i in range(n):
        want = o.embed(covers[i], bins, bits[i])
        d = np.abs(stego[i].astype(np.int16) - want["stego"].astype(np.int16))
        assert d.max() <= 1 and (d == 0).mean() >= 0.9999, (var, i, int(d.max()))
        assert int(usable[i]) == want["usable"]
        assert np.allclose(med[i], want["medians"], rtol=1e-11)
        wdec, wraw = o.extract(stego[i], bins[: nbits // 7 * 7], 7)
        assert np.array_equal(raw[i], wraw) and np.array_equal(dec[i], wdec), (var, i)
