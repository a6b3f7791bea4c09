#!/usr/bin/env python
"""A few small embed+extract round trips that touch every kernel family once -- meant to be run under
`compute-sanitizer --tool memcheck python tools/sanity_small.py` (minutes, not hours)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import steganosaurus_b200 as sb  # noqa: E402
from steganosaurus_b200 import synth  # noqa: E402

CASES = [(64, 48, False), (700, 300, True), (3001, 601, True), (3840, 2160, False), (4097, 513, False), (600, 4200, False)]
with sb.Context(0) as ctx:
    for W, H, center in CASES:
        PH, PW = synth.next_pow2(H), synth.next_pow2(W)
        nbits = 3000
        cover = synth.gen_texture(W, H, W + H)
        bins = synth.random_bins(PH, PW, nbits, 3)
        bits = synth.random_bits(2, nbits, 4)
        stego, usable, med = ctx.embed_batch(np.stack([cover, cover]), bins, bits, center=center)
        dec, raw = ctx.extract_bits(stego, bins, 1, center=center)
        print(W, H, center, "usable", int(usable[0]), "raw BER", float((raw != bits).mean()), flush=True)
print("sanity ok")
