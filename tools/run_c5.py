#!/usr/bin/env python
"""BASELINE config 5: ONE large RGB image, slab-decomposed 2-D FFT across the ranks with an all-to-all
transpose over NVLink (steganosaurus_b200/slab.py).  Launch with torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/run_c5.py --size 16384

Checks the slab path against the single-GPU path of the same library (spectrum, stego pixels, raw
bits) and prints one JSON line with device-side timings (CUDA events, max over ranks).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import steganosaurus_b200 as sb  # noqa: E402
from steganosaurus_b200 import host, slab, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--payload", type=int, default=30720)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-check", action="store_true")
    a = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    N = a.size
    W = H = PW = PH = N
    ctx = sb.Context(local)
    sf = slab.SlabFFT2D(dist, PH, PW, slab.library_pass_fn(ctx), dev)
    pw = b"correct horse battery staple"
    nbits = synth.frame_len(a.payload)
    t0 = time.time()
    bins_np = host.walk(pw, PH, PW, nbits)[0]
    t_walk = time.time() - t0
    bits_np = host.frame_bits(pw, bytes(range(16)), 1000, bytes(np.random.default_rng(5).integers(32, 127, a.payload, dtype=np.uint8)))[0]
    bins = torch.from_numpy(bins_np.astype(np.int64)).to(dev)
    bits = torch.from_numpy(bits_np.astype(np.int64)).to(dev)
    # every rank builds only its own rows of the (seeded) cover
    cover = synth.gen_cover(W, H, 5)
    my_rows = torch.from_numpy(cover[rank * sf.rows:(rank + 1) * sf.rows]).to(dev)

    def run_embed():
        x = sf.planes_from_u8_rows(my_rows, W, H)
        y = sf.forward(x)
        sf.embed_on_cols(y, bins, bits, 0.5)
        back = sf.inverse(y)
        return sf.u8_rows_from_planes(back, W, H)

    def run_extract(stego_rows):
        x = sf.planes_from_u8_rows(stego_rows, W, H)
        y = sf.forward(x)
        return sf.read_on_cols(y, bins)

    stego_rows = run_embed()  # warm-up + result
    raw = run_extract(stego_rows)
    torch.cuda.synchronize()
    dist.barrier()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(a.reps):
        stego_rows = run_embed()
    e[1].record()
    for _ in range(a.reps):
        raw = run_extract(stego_rows)
    e[2].record()
    torch.cuda.synchronize()
    t = torch.tensor([e[0].elapsed_time(e[1]) / a.reps, e[1].elapsed_time(e[2]) / a.reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ber = float((raw.cpu().numpy() != bits_np).mean())

    res = {"config": f"C5 single {N}x{N} RGB image, slab FFT on {world} GPU(s), {a.payload}-byte frame ({nbits} bits)",
           "n_gpus": world, "embed_ms": float(t[0]), "extract_ms": float(t[1]),
           "MP_per_s_embed_extract": N * N / 1e6 / ((float(t[0]) + float(t[1])) / 1e3), "raw_ber": ber, "walk_s": round(t_walk, 2)}
    # parity against the single-GPU path (rank 0 runs it on the whole image)
    gathered = [torch.empty_like(stego_rows) for _ in range(world)] if rank == 0 else None
    dist.gather(stego_rows, gathered, dst=0)
    if rank == 0 and not a.no_check:
        slab_stego = torch.cat(gathered, 0).cpu().numpy()
        one, usable, _ = ctx.embed_batch(cover[None], bins_np, bits_np[None])
        d = np.abs(slab_stego.astype(np.int16) - one[0].astype(np.int16))
        res["stego_max_diff_vs_single_gpu"] = int(d.max())
        res["stego_equal_frac"] = float((d == 0).mean())
        _, raw1 = ctx.extract_bits(one, bins_np, 1)
        res["raw_bits_equal_single_gpu"] = bool(np.array_equal(raw1[0], raw.cpu().numpy()))
        hdr, pay, _ = ctx.extract_frame(slab_stego[None], bins_np, 912)
        ok, pt = host.open_payload(pw, 1000, hdr[0].tobytes(), pay[0].tobytes(), a.payload)
        res["plaintext_recovered_from_slab_stego"] = bool(ok)
        assert d.max() <= 1 and res["stego_equal_frac"] > 0.9999 and res["raw_bits_equal_single_gpu"] and ok, res
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
