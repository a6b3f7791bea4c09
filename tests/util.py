"""Shared helpers for the test-suite."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "g*.npz")))


def load_golden(name):
    from steganosaurus_b200 import synth
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    W, H, seed = int(g["W"]), int(g["H"]), int(g["seed"])
    g["cover"] = (synth.gen_texture if int(g["texture"]) else synth.gen_cover)(W, H, seed)
    g["raw_all"] = np.unpackbits(g["raw_all"])[: g["bins"].size]
    for k in ("W", "H", "seed", "center", "hdr_n", "usable"):
        g[k] = int(g[k])
    for k in ("alpha", "rmin", "rmax"):
        g[k] = float(g[k])
    return g


def spec_err(a, b):
    """(max|d|/rms|b|, L2 relative) -- SURVEY section 6.2 recommended assertion."""
    d = np.abs(a - b)
    rms = np.sqrt(np.mean(np.abs(b) ** 2))
    l2 = np.sqrt((d ** 2).sum() / max((np.abs(b) ** 2).sum(), 1e-300))
    return float(d.max() / max(rms, 1e-300)), float(l2)
