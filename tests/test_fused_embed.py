"""GPU parity tests of the column-resident embed (pencil_col_embed_w: forward column FFT, phase write and inverse column
FFT in one shared-memory residency; median / capacity on the two word planes of q = |F|^2) against the oracle and
against the three-kernel sequence it replaces (TFFT_FUSED_EMBED=0).  4096-row half-spectrum planes only -- the headline
workload's geometry (3840x2160 pads to 4096x4096)."""
import os

import numpy as np
import pytest

from oracle import pyoracle as O
from steganosaurus_b200 import synth
import steganosaurus_b200 as sb

pytestmark = pytest.mark.gpu


def assert_pixels(a, b):
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    assert d.max() <= 1, f"max pixel diff {d.max()}"
    assert (d == 0).mean() >= 0.9999, f"only {(d == 0).mean():.6f} equal"


def unfused_context():
    os.environ["TFFT_FUSED_EMBED"] = "0"
    try:
        return sb.Context(0)
    finally:
        os.environ.pop("TFFT_FUSED_EMBED", None)


# (W, H): 2048 < H <= 2304 -> 9 input / output row blocks (the UHD case), else all 16
@pytest.mark.parametrize("W,H,nbits,center", [(600, 4096, 60000, False), (700, 2160, 50000, True), (512, 2049, 30000, False),
                                              (1500, 2305, 40000, False), (4096, 2100, 50000, True)])
def test_fused_embed_vs_oracle_and_unfused(ctx, W, H, nbits, center):
    o = O.best()
    PH, PW = synth.next_pow2(H), synth.next_pow2(W)
    assert PH == 4096
    cover = synth.gen_texture(W, H, W + 5 * H) if W != 700 else synth.gen_cover(W, H, 3)
    bins = synth.random_bins(PH, PW, nbits, 21)
    bits = synth.random_bits(1, nbits, 22)
    launches0 = ctx.launches
    ctx.profile_reset(); ctx.profile_enable(True)
    stego, usable, med = ctx.embed_batch(cover[None], bins, bits, 0.5, center)
    prof = ctx.profile_read(); ctx.profile_enable(False)
    assert prof["col_embed_fused"][0] == 1 and prof["col_fwd"][0] == 0 and prof["col_inv"][0] == 0, prof  # the fused pass ran
    want = o.embed(cover, bins, bits[0], 0.5, center)
    assert int(usable[0]) == want["usable"]
    assert np.allclose(med[0], want["medians"], rtol=1e-11)
    assert_pixels(stego[0], want["stego"])
    with unfused_context() as c2:
        s2, u2, m2 = c2.embed_batch(cover[None], bins, bits, 0.5, center)
    # the same transform up to rounding (the fused kernel keeps the textbook radix-16 butterfly, the others fold its constant
    # twiddles into FMAs); |F| at the bins by sqrt(re^2 + im^2) here, hypot() there
    assert_pixels(stego[0], s2[0])
    assert (stego != s2).mean() < 1e-6
    assert np.array_equal(usable, u2)
    assert np.abs(med - m2).max() <= 1e-13 * np.abs(m2).max()
    _, raw = ctx.extract_bits(stego, bins, 1, 0.5, center)
    _, wraw = o.extract(want["stego"], bins, 1, 0.5, center)
    assert np.array_equal(raw[0], wraw)


def test_fused_batch_every_image_vs_oracle(ctx):
    """Several images per call (per-plane offsets of the q planes, the bit masks and the median work lists)."""
    W, H, n, nbits = 600, 2160, 3, 20000
    covers = np.stack([synth.gen_texture(W, H, 500 + i) if i != 1 else synth.gen_cover(W, H, 501) for i in range(n)])
    bins = synth.random_bins(4096, 1024, nbits, 9)
    bits = synth.random_bits(n, nbits, 10)
    stego, usable, med = ctx.embed_batch(covers, bins, bits)
    o = O.best()
    for i in range(n):
        want = o.embed(covers[i], bins, bits[i])
        assert_pixels(stego[i], want["stego"])
        assert int(usable[i]) == want["usable"]
        assert np.allclose(med[i], want["medians"], rtol=1e-11)


def test_fused_over_capacity_passes_the_cover_through(ctx):
    """S:1009-1012: the fused pass embeds speculatively; an image whose capacity is below nbits must leave as its cover,
    with the reference's counts."""
    W, H = 600, 4096
    cover = synth.gen_cover(W, H, 5)
    allb = synth.valid_bins(4096, 1024)
    want_usable = O.best().embed(cover, allb[:10], np.zeros(10, np.uint8))["usable"]
    nbits = want_usable + 1000
    assert nbits <= allb.size
    bins = synth.random_bins(4096, 1024, nbits, 1)
    bits = synth.random_bits(1, nbits, 2)
    with pytest.raises(sb.CapacityError) as ei:
        ctx.embed_batch(cover[None], bins, bits)
    assert int(ei.value.usable[0]) == want_usable
    assert np.array_equal(ei.value.stego[0], cover)
    # one bit fewer than the capacity embeds
    bins2 = bins[:want_usable]
    stego, usable, _ = ctx.embed_batch(cover[None], bins2, bits[:, :want_usable])
    assert int(usable[0]) == want_usable and not np.array_equal(stego[0], cover)


def test_fused_refuses_lists_it_cannot_hold(ctx):
    """Bins on column 0, on or right of the Nyquist column, or with jitter run on the three-kernel sequence: results
    still equal the oracle's."""
    W, H, n = 600, 4096, 1500
    PW = 1024
    cover = synth.gen_texture(W, H, 9)
    rng = np.random.default_rng(3)
    y = rng.permutation(np.arange(1, 2000))[:n].astype(np.uint32)
    x = rng.integers(PW // 2 + 1, PW, n).astype(np.uint32)  # right of the Nyquist column: stored through the mirror
    bins = (rng.integers(0, 3, n).astype(np.uint32) << np.uint32(30)) | (y * np.uint32(PW) + x)
    bits = synth.random_bits(1, n, 4)
    ctx.profile_reset(); ctx.profile_enable(True)
    stego, usable, _ = ctx.embed_batch(cover[None], bins, bits)
    prof = ctx.profile_read(); ctx.profile_enable(False)
    assert prof["col_embed_fused"][0] == 0 and prof["embed_scatter"][0] == 1
    want = O.best().embed(cover, bins, bits[0])
    assert_pixels(stego[0], want["stego"])
    assert int(usable[0]) == want["usable"]


def test_fused_device_pointer_entry(ctx):
    import torch
    W, H, n, nbits = 600, 2160, 2, 30000
    covers = np.stack([synth.gen_texture(W, H, 40 + i) for i in range(n)])
    bins = synth.random_bins(4096, 1024, nbits, 3)
    bits = synth.random_bits(n, nbits, 4)
    hs, hu, hm = ctx.embed_batch(covers, bins, bits)
    dev = torch.device("cuda:0")
    d_cover = torch.from_numpy(covers).to(dev)
    d_bins = torch.from_numpy(bins.view(np.int32)).to(dev)
    d_bits = torch.from_numpy(bits).to(dev)
    d_stego = torch.empty_like(d_cover)
    d_us = torch.zeros(n, dtype=torch.int64, device=dev)
    d_med = torch.zeros(n, 3, dtype=torch.float64, device=dev)
    ctx.embed_batch_dev(d_cover, d_bins, d_bits, d_stego, usable=d_us, median=d_med)
    torch.cuda.synchronize()
    assert np.array_equal(d_stego.cpu().numpy(), hs)
    assert np.array_equal(d_us.cpu().numpy().astype(np.uint64), hu)
    assert np.array_equal(d_med.cpu().numpy(), hm)
