"""Regenerate tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libtfft_ref.so, built by
oracle/Makefile from the unmodified sources under /root/reference).  Run in the dev container:

    python tests/golden/make_golden.py

Each fixture stores inputs (cover, bins, bits, params) and the reference's outputs (stego pixels,
raw phase bits, decoded bytes, medians, usable, a checksum and a few samples of the spectrum), so
the oracle port and the CUDA path can both be pinned without /root/reference being present.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import pyoracle as O  # noqa: E402
from steganosaurus_b200 import synth  # noqa: E402

PASS = b"correct horse battery staple"


def case(name, W, H, nbits, seed, center=False, alpha=0.5, walk=False, texture=False, rmin=0.05, rmax=0.45):
    r = O.ref()
    PH, PW = synth.next_pow2(H), synth.next_pow2(W)
    cover = (synth.gen_texture if texture else synth.gen_cover)(W, H, seed)
    if walk:
        bins, start, ctr = r.walk(PASS, PH, PW, nbits, rmin, rmax, 0.7)
    else:
        bins = synth.random_bins(PH, PW, nbits, seed, rmin, rmax)
    bits = synth.random_bits(1, nbits, seed + 1)[0]
    e = r.embed(cover, bins, bits, alpha, center, 0.01, rmin, rmax, want_spectrum=True)
    F0 = r.forward_spectrum(cover, center)
    hdr_n = min(nbits, 912) // 3 * 3
    dec3, raw = r.extract(e["stego"], bins[:hdr_n], 3, alpha, center)
    _, raw_all = r.extract(e["stego"], bins, 1, alpha, center)
    dec7 = r.rep_decode(raw_all[hdr_n:hdr_n + (nbits - hdr_n) // 7 * 7], 7)
    rng = np.random.default_rng(seed + 2)
    sy, sx = rng.integers(0, PH, 64), rng.integers(0, PW, 64)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        W=W, H=H, seed=seed, center=int(center), alpha=alpha, rmin=rmin, rmax=rmax, texture=int(texture),
        bins=bins, bits=bits, stego=e["stego"], medians=e["medians"], usable=e["usable"],
        raw_all=np.packbits(raw_all), dec3=dec3, dec7=dec7, hdr_n=hdr_n,
        spec_sample_yx=np.stack([sy, sx]), spec_sample=F0[:, sy, sx],
        spec_after_sample=e["spectrum"][:, sy, sx],
        spec_abs_sum=np.abs(F0).sum(axis=(1, 2)), spec_sum=F0.sum(axis=(1, 2)),
    )
    ber = float((raw_all != bits).mean())
    print(f"{name}: {W}x{H}->{PW}x{PH} nbits={nbits} usable={e['usable']} raw BER={ber:.4f} "
          f"changed px={(e['stego'] != cover).mean():.3f}")


if __name__ == "__main__":
    case("g256_walk", 256, 256, 2480, 1, walk=True)             # doc/HARDENING.md:467 (12-byte message)
    case("g512_walk", 512, 512, 2928, 2, walk=True)             # C1: "the eagle has landed"
    case("g96x80_center", 96, 80, 700, 3, center=True, alpha=0.3)  # non-pow2 (pad 128x128), centre on
    case("g128x64_texture", 128, 64, 600, 4, texture=True)      # non-square pow2, clamp path
    print("walk KATs")
    r = O.ref()
    out = {}
    for pw, tag in ((b"pw", "pw"), (PASS, "chbs")):
        for n in (512, 4096):
            bins, start, ctr = r.walk(pw, n, n, 2928)
            out[f"{tag}_{n}_bins64"] = bins[:64]
            out[f"{tag}_{n}_start"] = np.array(start)
            out[f"{tag}_{n}_ctr"] = ctr
            h = np.uint64(0xcbf29ce484222325)
            p, y, x = O.unpack_bins(bins, n)
            data = np.stack([p, y, x], 1).astype("<u4").tobytes()
            hv = 0xcbf29ce484222325
            for b in data:
                hv = ((hv ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
            out[f"{tag}_{n}_fnv"] = np.uint64(hv)
            print(tag, n, start, ctr, hex(hv))
    np.savez_compressed(os.path.join(HERE, "walk_kat.npz"), **out)
