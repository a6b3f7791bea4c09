"""Extracts of 8192-row planes (GPU): the forward row pass takes the first radix-2 step of the 8192-point column
transform (rows y and y + 4096 leave as A_y and B_y, PassArgs::fold in csrc/tfft_kernels.cuh), so the column pass is two
4096-row sign-map passes per plane instead of a four-step pass.  Oracle = the reference's own read path
(fft2d S:359-366 + read_bit_from_bin S:734-746, votes S:468-508) on the same image and bins: raw bits and decoded
bytes bit-exact."""
import numpy as np
import pytest

from oracle import pyoracle as O
from steganosaurus_b200 import synth

pytestmark = pytest.mark.gpu


# narrow rows (the row pair of one transform is (y, y + 4096)): 1024-, 4096-point kernels, odd H, second rows mostly absent;
# wide rows (8192 pixels, the two units of a CTA take y and y + 4096): C4's 8192^2 geometry
@pytest.mark.parametrize("W,H,center", [(600, 5000, False), (1000, 4097, True), (4000, 4500, True), (520, 8192, False),
                                        (4100, 4200, False), (5000, 4099, True)])
def test_fold_extract_vs_oracle(ctx, W, H, center):
    o = O.best()
    PH, PW = synth.next_pow2(H), synth.next_pow2(W)
    assert PH == 8192
    img = synth.gen_texture(W, H, 5 * W + H)
    nbits = 21000
    bins = synth.random_bins(PH, PW, nbits, 17)          # the walk's quarter annulus: rows and columns < 0.45 * min(PH, PW)
    _, wraw = o.extract(img, bins, 1, 0.5, center)       # ONE reference transform per case (an 8192^2 one takes ~20 s)

    def vote(raw, rep):                                  # rep3/rep7_decode_bits S:468/S:501 + bytes_from_bits S:447
        nd = raw.size // rep
        return np.packbits((raw[:nd * rep].reshape(nd, rep).sum(1) >= rep // 2 + 1).astype(np.uint8))

    for rep in (1, 3, 7):
        dec, raw = ctx.extract_bits(img[None], bins, rep, 0.5, center)
        assert np.array_equal(raw[0], wraw), rep
        assert np.array_equal(dec[0], vote(wraw, rep)), rep
    # header + payload in one call (S:1223-1268)
    nb = 912 + 56 * 300
    hdr, pay, raw = ctx.extract_frame(img[None], bins[:nb], 912, 0.5, center, want_raw=True)
    assert np.array_equal(raw[0], wraw[:nb])
    assert np.array_equal(hdr[0], vote(wraw[:912], 3)) and np.array_equal(pay[0], vote(wraw[912:nb], 7))


def test_fold_falls_back_outside_its_window(ctx):
    """Bins beyond stored row 4095 (or read through the Hermitian mirror) cannot use the folded pass: same bits from the
    general path."""
    o = O.best()
    W, H = 600, 4200
    PH, PW = 8192, 1024
    img = synth.gen_texture(W, H, 99)
    rng = np.random.default_rng(3)
    n = 4998
    for (y0, y1, x0, x1) in ((0, PH, 0, PW // 2), (10, 600, PW // 2 + 1, PW), (0, 4096, 0, 300)):
        y = rng.integers(y0, y1, n).astype(np.uint32)
        x = rng.integers(x0, x1, n).astype(np.uint32)
        pl = rng.integers(0, 3, n).astype(np.uint32)
        bins = (pl << np.uint32(30)) | (y * np.uint32(PW) + x)
        dec, raw = ctx.extract_bits(img[None], bins, 7)
        wdec, wraw = o.extract(img, bins, 7)
        assert np.array_equal(raw[0], wraw) and np.array_equal(dec[0], wdec)
