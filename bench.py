#!/usr/bin/env python
"""bench.py -- embed+extract megapixels/s for a batch of 4K UHD RGB covers (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over a batch of B synthetic 3840x2160 RGB covers
(pad 4096^2), each with its own 30 720-byte frame (1 722 128 embedded bits, one shared bin list):
embed (fwd 2-D FFT, median+capacity, phase scatter, inv 2-D FFT, u8 quantise) followed by extract
(fwd 2-D FFT, phase gather, Rep-3/Rep-7 vote).  MP = W*H image pixels per image, counted once per
embed+extract round.  Every rank processes its own batch (weak scaling, no collective on the data
path); the only torch.distributed traffic is the timing barrier and the max-over-ranks reduce.

Printed line (rank 0): see the task contract -- value (inputs resident in HBM), e2e (host buffers
through the C-ABI, H2D/D2H inside the timed region), roofline (dominant kernel, CUDA events inside
the timed region, against MEASURED_PEAKS.json), cpu_baseline (the reference's own hot-path code on
the host cores), clocks, gpu_launches.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W_UHD, H_UHD = 3840, 2160
PAYLOAD = 30720
METRIC = "embed+extract megapixels/sec (4K RGB batch)"
UNIT = "MP/s"
PARAMS = dict(alpha=0.5, center=False, magmin=0.01, rmin=0.05, rmax=0.45)


# ------------------------------------------------------------------------------------------------
def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def device_for_rank(local_rank: int, world: int) -> int:
    """Rank -> CUDA device.  On this pool's 8-GPU nodes GPUs 0-3 and GPUs 4-7 each share one host path (measured:
    profiles/r2_pcie_ceiling_8gpu_node.jsonl -- four concurrent pinned H2D streams on GPUs 0-3 get 115 GB/s together, on
    GPUs 0,4,1,5 212 GB/s), so a 2- or 4-rank job on such a node takes its GPUs from both halves.  The device-resident
    numbers do not depend on this; TFFT_BENCH_DEVICE_ORDER=identity switches it off."""
    import torch
    n = torch.cuda.device_count()
    order = os.environ.get("TFFT_BENCH_DEVICE_ORDER", "interleave")
    if order == "interleave" and n == 8 and world in (2, 4):
        return [0, 4, 1, 5, 2, 6, 3, 7][local_rank]
    return local_rank % max(n, 1)


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(pw)), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def make_frame_bits(n: int, payload: int, seed: int) -> np.ndarray:
    """[n, 912 + 56*(payload+16)] bits with the reference's framing structure (Rep-3 header, Rep-7 payload)."""
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 2, size=(n, 304 + 8 * (payload + 16)), dtype=np.uint8)
    return np.concatenate([np.repeat(raw[:, :304], 3, axis=1), np.repeat(raw[:, 304:], 7, axis=1)], axis=1)


def make_covers(batch: int, W: int, H: int, distinct: int = 8) -> np.ndarray:
    from steganosaurus_b200 import synth
    base = [synth.gen_cover(W, H, 1000 + i) for i in range(min(distinct, batch))]
    return np.stack([base[i % len(base)] for i in range(batch)])


# ------------------------------------------------------------------------------------------------
# CPU reference arm: the reference's own hot-path functions (oracle/_ref) or the C port, one image
# per worker process, all host cores.
_CPU_STATE = None


def _cpu_init(W, H, payload, use_ref, mode="embed+extract"):
    """Pool initializer: every worker process builds its own image, bin list and frame once (mode "extract": also the
    stego image the timed extract reads, made by the same checker outside the timed region)."""
    global _CPU_STATE
    sys.path.insert(0, ROOT)
    from oracle import pyoracle as O
    from steganosaurus_b200 import synth
    idx = os.getpid() % 64
    PH, PW = synth.next_pow2(H), synth.next_pow2(W)
    nbits = synth.frame_len(payload)
    cover = synth.gen_cover(W, H, 1000 + idx)
    from steganosaurus_b200 import host
    bins = host.walk(b"correct horse battery staple", PH, PW, nbits, PARAMS["rmin"], PARAMS["rmax"], 0.7)[0]
    bits = make_frame_bits(1, payload, 2000 + idx)[0]
    o = O.ref() if use_ref else O.port()
    stego = o.embed(cover, bins, bits, PARAMS["alpha"], PARAMS["center"], PARAMS["magmin"], PARAMS["rmin"], PARAMS["rmax"])["stego"] if mode == "extract" else None
    _CPU_STATE = (o, cover, bins, bits, mode, stego)


def _cpu_ready(_):
    time.sleep(0.2)
    return _CPU_STATE is not None


def _cpu_step(_):
    o, cover, bins, bits, mode, stego = _CPU_STATE
    t0 = time.time()
    if mode == "extract":  # forward FFT + phase read + Rep-3 / Rep-7 vote of one stego image (S:1116-1268)
        if o.kind == "reference":
            o.time_extract(stego, bins, PARAMS["alpha"], PARAMS["center"])
        else:
            o.extract(stego, bins[:912], 3, PARAMS["alpha"], PARAMS["center"])
        return t0, time.time(), 0.0, 0.0
    if o.kind == "reference":
        te, stego = o.time_embed(cover, bins, bits, PARAMS["alpha"], PARAMS["center"], PARAMS["magmin"], PARAMS["rmin"], PARAMS["rmax"])
        tx, _ = o.time_extract(stego, bins, PARAMS["alpha"], PARAMS["center"])
    else:
        e = o.embed(cover, bins, bits, PARAMS["alpha"], PARAMS["center"], PARAMS["magmin"], PARAMS["rmin"], PARAMS["rmax"])
        o.extract(e["stego"], bins[:912], 3, PARAMS["alpha"], PARAMS["center"])
        te = tx = 0.0
    return t0, time.time(), te, tx


class CpuArm:
    def __init__(self, W, H, payload, workers=None, mode="embed+extract"):
        import multiprocessing as mp
        from oracle import pyoracle as O
        self.use_ref = O.have_ref()
        self.kind = "reference" if self.use_ref else "port"
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        try:
            avail = int(open("/proc/meminfo").read().split("MemAvailable:")[1].split()[0]) * 1024
        except Exception:
            avail = 32 << 30
        PH, PW = 1 << (H - 1).bit_length(), 1 << (W - 1).bit_length()
        per = 140 * PH * PW  # reference needs ~130 B per padded bin
        self.workers = max(1, min(ncpu, int(avail * 0.7 // per))) if workers is None else workers
        self.W, self.H, self.payload, self.mode = W, H, payload, mode
        self.pool = mp.get_context("spawn").Pool(self.workers, initializer=_cpu_init, initargs=(W, H, payload, self.use_ref, mode))
        self.pool.map(_cpu_ready, range(self.workers), chunksize=1)  # wait until every worker is initialised

    def step(self):
        """All workers run one embed+extract of their own image concurrently; returns wall seconds."""
        r = self.pool.map(_cpu_step, range(self.workers), chunksize=1)
        return max(x[1] for x in r) - min(x[0] for x in r)

    def close(self):
        self.pool.close()
        self.pool.join()

    def describe(self, value):
        return {"value": value, "unit": UNIT, "cores": self.workers, "kind": self.kind,
                "sample": f"{self.workers} concurrent single-threaded processes x 1 image {self.W}x{self.H} "
                          f"({self.payload}-byte frame) {self.mode} per step; hot-path stages only "
                          f"(to_planes..from_planes incl. median, capacity, F3 copy; PNG, KDF, walk excluded)"}


def cpu_measure(arm, warmup, steps):
    """Mean wall seconds of one CPU step after `warmup` untimed ones (both arms of bench.py use this)."""
    for _ in range(warmup):
        arm.step()
    return float(np.mean([arm.step() for _ in range(steps)]))


def run_reference_arm(args):
    rank, local_rank, world = dist_env()
    if rank != 0:
        return 0
    if args.config == "c4":
        return run_c4_reference(args)
    if args.config == "c5":
        print(json.dumps({"impl": "reference", "metric": C5_METRIC, "unavailable": "one 16384^2 embed+extract is ~8 min and 35 GB on one host core (SURVEY 8d)"}), flush=True)
        return 0
    arm = CpuArm(args.width, args.height, args.payload)
    t = cpu_measure(arm, args.warmup, args.steps)
    arm.close()
    mp_per_step = arm.workers * args.width * args.height / 1e6
    value = mp_per_step / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"batch of 4K UHD RGB covers {args.width}x{args.height}, {args.payload}-byte frame "
                               f"({912 + 56 * (args.payload + 16)} bits), embed+extract; CPU step = {arm.workers} images",
                   "images_per_step": arm.workers},
        "cpu_baseline": arm.describe(value),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0



# ------------------------------------------------------------------------------------------------
# BASELINE config 4: extract-only throughput sweep 512^2 -> 8192^2 over synthetic stego batches (forward FFT + phase
# gather + Rep-3 / Rep-7 decode).  One JSON line per N.  SURVEY 8(d): batch = max(8, floor(8 GiB / (48 N^2))) capped at
# 256, payload = 50 % of the embed capacity, real keyed turtlewalk; stego batches are made by our embed (validated against
# the reference in tests/test_configs.py).
C4_METRIC = "extract-only megapixels/sec (synthetic stego batch, forward FFT + phase gather + Rep-3/Rep-7 decode)"
C4_CPU_MAX_N = 4096  # one 8192^2 reference extract is ~60 s of CPU time per image: not sampled


def c4_case(ctx, N):
    from steganosaurus_b200 import host, synth
    batch = min(256, max(8, int((8 << 30) // (48 * N * N))))
    cover1 = synth.gen_cover(N, N, 7)
    _, usable, _ = ctx.embed_batch(cover1[None], np.zeros(0, np.uint32), np.zeros((1, 0), np.uint8))
    cap = int(usable[0])
    plen = max(16, (cap // 2 - 912) // 56 - 16)
    nbits = synth.frame_len(plen)
    bins = host.walk(b"correct horse battery staple", N, N, nbits, PARAMS["rmin"], PARAMS["rmax"], 0.7)[0]
    bits1 = make_frame_bits(1, plen, N)[0]
    distinct = min(batch, 4)
    covers = np.stack([synth.gen_cover(N, N, 7 + i) for i in range(distinct)])
    stego, _, _ = ctx.embed_batch(covers, bins, np.stack([bits1] * distinct))
    return batch, plen, nbits, cap, bins, bits1, np.stack([stego[i % distinct] for i in range(batch)])


def run_c4(args):
    import torch
    import steganosaurus_b200 as sb
    rank, local_rank, world = dist_env()
    dev_index = device_for_rank(local_rank, world)
    torch.cuda.set_device(dev_index)
    dev = torch.device("cuda", dev_index)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    from steganosaurus_b200 import shard
    timing = shard.Timing(dist, dev)
    peak, peak_src = measured_peak_hbm()
    ctx = sb.Context(dev_index)
    for N in [int(x) for x in args.sizes.split(",")]:
        cpu_base = None
        batch, plen, nbits, cap, bins, bits1, stego = c4_case(ctx, N)
        if rank == 0 and world == 1 and not args.no_cpu_baseline and N <= C4_CPU_MAX_N:
            arm = CpuArm(N, N, plen, mode="extract")
            t = cpu_measure(arm, 1, 1)
            arm.close()
            cpu_base = arm.describe(arm.workers * N * N / 1e6 / t)
        d_stego = torch.from_numpy(stego).to(dev)
        d_bins = torch.from_numpy(bins.view(np.int32)).to(dev)
        d_hdr = torch.zeros(batch, 38, dtype=torch.uint8, device=dev)
        d_pay = torch.zeros(batch, plen + 16, dtype=torch.uint8, device=dev)

        def step_dev():
            ctx.extract_frame_dev(d_stego, d_bins, 912, d_hdr, d_pay, alpha=PARAMS["alpha"], center=PARAMS["center"])

        for _ in range(args.warmup):
            step_dev()
        timing.barrier(); torch.cuda.synchronize()
        sampler = ClockSampler(dev_index)
        if rank == 0:
            sampler.start()
        ctx.profile_reset(); ctx.profile_enable(True)
        l0 = ctx.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step_dev()
        e1.record()
        timing.barrier(); torch.cuda.synchronize()
        dev_ms = timing.max_over_ranks(e0.elapsed_time(e1)) / args.steps
        launches = ctx.launches - l0
        prof = ctx.profile_read(); ctx.profile_enable(False)
        clocks = sampler.stop() if rank == 0 else None
        want_hdr = np.packbits(bits1[:912:3]); want_pay = np.packbits(bits1[912::7])
        wrong = int(np.unpackbits(d_hdr[0].cpu().numpy() ^ want_hdr).sum() + np.unpackbits(d_pay[0].cpu().numpy() ^ want_pay).sum())
        # end to end: pinned host stego -> C-ABI (H2D inside) -> decoded bytes on the host
        hs = torch.from_numpy(stego).pin_memory().numpy()
        for _ in range(min(args.warmup, 2)):
            ctx.extract_frame(hs, bins, 912, alpha=PARAMS["alpha"], center=PARAMS["center"])
        timing.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.extract_frame(hs, bins, 912, alpha=PARAMS["alpha"], center=PARAMS["center"])
        torch.cuda.synchronize()
        e2e_ms = timing.max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
        # the literal S:1223-1268 flow through the host API: forward call, header read (912 bins, rep 3), payload read (rep 7)
        def two_phase():
            ctx.forward_batch(hs, PARAMS["center"])
            h_, _ = ctx.read_bits(bins[:912], 3, PARAMS["alpha"], want_raw=False)
            p_, _ = ctx.read_bits(bins[912:], 7, PARAMS["alpha"], want_raw=False)
            return h_, p_
        tp_ms, tp_wrong = None, None
        if batch <= 64:
            try:
                h_, p_ = two_phase()
                tp_wrong = int(np.unpackbits(h_[0] ^ want_hdr).sum() + np.unpackbits(p_[0][:want_pay.size] ^ want_pay).sum())
                timing.barrier(); torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    two_phase()
                torch.cuda.synchronize()
                tp_ms = timing.max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
            except sb.TfftError:
                tp_ms = None  # (the batch does not fit one resident workspace)
        if rank == 0:
            name, (groups, ms, nbytes) = max(prof.items(), key=lambda kv: kv[1][1])
            ach = (nbytes / 1e9) / (ms / 1e3) if ms > 0 else 0.0
            mp = batch * N * N / 1e6
            line = {
                "metric": C4_METRIC, "value": world * mp / (dev_ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"C4 extract-only: batch of {batch} {N}x{N} RGB stego images per GPU, {plen}-byte payload = 50 % of the "
                                       f"capacity ({nbits} of {cap} bits, keyed turtlewalk), header + payload in one call",
                           "N": N, "images_per_gpu_per_step": batch, "wrong_voted_bits_image0": wrong,
                           "l2": "batch larger than L2" if batch * N * N * 3 > (126 << 20) else "batch smaller than L2: rotated over steps"},
                "clocks": clocks,
                "e2e": {"value": world * mp / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int(batch * N * N * 3 + 4 * nbits), "d2h_bytes_per_step": int(batch * (38 + plen + 16)), "steps": args.steps},
                "two_phase": None if tp_ms is None else {"value": world * mp / (tp_ms / 1e3), "unit": UNIT, "ms_per_step": tp_ms,
                                                         "calls": "tfft_forward_batch + tfft_read_bits(912 bins, rep 3) + tfft_read_bits(payload, rep 7), host buffers",
                                                         "wrong_voted_bits_image0": tp_wrong},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if peak else None,
                             "traffic": None, "peak_source": peak_src, "launch_groups": groups,
                             "kernels": {k: {"groups": v[0], "ms": round(v[1], 3), "GBps": round((v[2] / 1e9) / (v[1] / 1e3), 1) if v[1] > 0 else None}
                                         for k, v in prof.items() if v[0]}},
                "cpu_baseline": cpu_base,
            }
            print(json.dumps(line), flush=True)
        del d_stego, hs, stego
        torch.cuda.empty_cache()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_c4_reference(args):
    """--impl reference --config c4: the reference's own extract on the host cores, one line per N (bounded sample)."""
    from steganosaurus_b200 import synth
    for N in [int(x) for x in args.sizes.split(",")]:
        if N > C4_CPU_MAX_N:
            print(json.dumps({"impl": "reference", "metric": C4_METRIC, "config": {"N": N}, "unavailable": "one image is ~60 s of CPU time"}), flush=True)
            continue
        # same payload rule as the GPU arm, from the capacity SURVEY 6.2 quotes for this kind of cover (~0.2355 N^2 bits)
        plen = max(16, (int(0.2355 * N * N) // 2 - 912) // 56 - 16)
        arm = CpuArm(N, N, plen, mode="extract")
        t = cpu_measure(arm, args.warmup, args.steps)
        arm.close()
        value = arm.workers * N * N / 1e6 / t
        print(json.dumps({"impl": "reference", "metric": C4_METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f64", "data": "synthetic", "config": {"workload": f"C4 extract-only {N}x{N}, {plen}-byte payload; CPU step = {arm.workers} images", "N": N},
                          "cpu_baseline": arm.describe(value),
                          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# BASELINE config 5: ONE 16384 x 16384 RGB image, slab-decomposed 2-D FFT over the ranks (steganosaurus_b200/slab.py).
# A step = embed (my rows of the cover -> my rows of the stego image) + extract (stego rows -> raw read bits of the
# whole frame, combined over ranks).  Strong scaling: the image is fixed, the ranks share it.
C5_METRIC = "embed+extract megapixels/sec (single 16K x 16K RGB image, slab-decomposed 2-D FFT)"


def c5_rows(W, y0, nrows, seed):
    """gen_png-style rows (tools/gen_png.cpp:8-17) of a tall image, generated per rank (the full 16K image costs 6 GB of host temporaries)."""
    rng = np.random.default_rng([seed, y0])
    x = np.arange(W, dtype=np.int32)[None, :]
    y = (np.arange(nrows, dtype=np.int32) + y0)[:, None]
    n = rng.integers(-10, 10, size=(nrows, W), dtype=np.int32)
    img = np.empty((nrows, W, 3), np.int32)
    img[:, :, 0] = 180 + (x * 40) // W + n
    img[:, :, 1] = 180 + (y * 40) // 16384 + n
    img[:, :, 2] = 200 + n
    return np.clip(img, 0, 255).astype(np.uint8)


def run_c5(args):
    import torch
    import steganosaurus_b200 as sb
    from steganosaurus_b200 import host, slab, synth, shard
    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)
    timing = shard.Timing(dist, dev)
    N = args.size
    nbits = synth.frame_len(args.payload)
    bins_np = host.walk(b"correct horse battery staple", N, N, nbits, PARAMS["rmin"], PARAMS["rmax"], 0.7)[0]
    bits_np = make_frame_bits(1, args.payload, 2005)[0]
    ctx = sb.Context(local_rank)
    kind = args.transport
    try:
        eng = slab.SlabEngine(ctx, N, N, world, rank, dist=dist, transport_kind=kind)
    except RuntimeError:
        kind = "collective"  # no CUDA IPC / peer access on this box: NCCL all-to-all on the zero-copy send buffer
        eng = slab.SlabEngine(ctx, N, N, world, rank, dist=dist, transport_kind=kind)
    plan = eng.plan
    rows_np = c5_rows(N, plan.y0, plan.nrows, 5)
    d_rows = torch.from_numpy(rows_np).to(dev)
    d_bins = torch.from_numpy(bins_np.view(np.int32)).to(dev)
    d_bits = torch.from_numpy(bits_np).to(dev)

    ev_mid = []

    def step_dev(mark=False):
        stego = eng.embed(d_rows, d_bins, d_bits, PARAMS["alpha"], PARAMS["center"])
        if mark:
            e = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            e[0].record()
        raw_ = eng.extract_raw(stego, d_bins, PARAMS["alpha"], PARAMS["center"], dist=dist)
        if mark:
            e[1].record()
            ev_mid.append(e)
        return stego, raw_

    for _ in range(args.warmup):
        step_dev()
    timing.barrier(); torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ctx.profile_reset(); ctx.profile_enable(True)
    l0 = ctx.launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(args.steps):
        stego, raw = step_dev(mark=True)
    ev[1].record()
    timing.barrier(); torch.cuda.synchronize()
    dev_ms = timing.max_over_ranks(ev[0].elapsed_time(ev[1])) / args.steps
    extract_ms = timing.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev_mid) / args.steps)
    launches = ctx.launches - l0
    prof = ctx.profile_read(); ctx.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    ber = float((raw.cpu().numpy().astype(np.uint8) != bits_np).mean())
    # end to end: pinned host rows in, stego rows and raw bits back on the host, every step
    h_rows = torch.from_numpy(rows_np).pin_memory()
    h_out = torch.empty_like(h_rows).pin_memory()
    h_raw = torch.empty(nbits, dtype=torch.int8).pin_memory()

    def step_e2e():
        d = h_rows.to(dev, non_blocking=True)
        st = eng.embed(d, d_bins, d_bits, PARAMS["alpha"], PARAMS["center"])
        h_out.copy_(st, non_blocking=True)
        h_raw.copy_(eng.extract_raw(st, d_bins, PARAMS["alpha"], PARAMS["center"], dist=dist), non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(min(args.warmup, 2)):
        step_e2e()
    timing.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    e2e_ms = timing.max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        name, (groups, ms, nbytes) = max(prof.items(), key=lambda kv: kv[1][1])
        ach = (nbytes / 1e9) / (ms / 1e3) if ms > 0 else 0.0
        mp = N * N / 1e6
        line = {
            "metric": C5_METRIC, "value": mp / (dev_ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C5 single {N}x{N} RGB image, {args.payload}-byte frame ({nbits} bits, keyed turtlewalk), embed + extract, "
                                   f"row slabs of {plan.R} rows -> column slabs of {plan.cols} half-spectrum columns on {world} GPU(s)",
                       "transport": kind, "raw_ber": ber, "embed_ms": dev_ms - extract_ms, "extract_ms": extract_ms,
                       "nvlink_bytes_per_rank_per_plane_and_exchange": plan.exchange_bytes_per_plane(),
                       "algorithmic_exchange_note": "SURVEY 5.8 quotes 448 MiB per GPU and plane for the dense complex spectrum at G = 8; the Hermitian half moves half of it",
                       "l2": "one 16K image: 6.5 GB of half spectra per direction, far larger than L2"},
            "clocks": clocks,
            "e2e": {"value": mp / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(world * rows_np.nbytes),
                    "d2h_bytes_per_step": int(world * rows_np.nbytes + nbits), "steps": args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if peak else None,
                         "traffic": None, "peak_source": peak_src, "launch_groups": groups,
                         "kernels": {k: {"groups": v[0], "ms": round(v[1], 3), "GBps": round((v[2] / 1e9) / (v[1] / 1e3), 1) if v[1] > 0 else None}
                                     for k, v in prof.items() if v[0]}},
            "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        if hasattr(eng.tr, "close"):
            eng.tr.close()
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    if args.config == "c4":
        return run_c4(args)
    if args.config == "c5":
        return run_c5(args)
    import torch
    import steganosaurus_b200 as sb
    from steganosaurus_b200 import synth

    rank, local_rank, world = dist_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    if args.gpus > 1 and world == 1:
        print("bench.py: --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)", file=sys.stderr)
        return 2
    dev_index = device_for_rank(local_rank, world)
    torch.cuda.set_device(dev_index)
    dev = torch.device("cuda", dev_index)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)

    W, H, B = args.width, args.height, args.batch
    PH, PW = synth.next_pow2(H), synth.next_pow2(W)
    nbits = synth.frame_len(args.payload)
    npay = args.payload + 16

    # CPU baseline first (rank 0, N=1 only), before the GPU is busy
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        arm = CpuArm(W, H, args.payload)
        t = cpu_measure(arm, 1, 1)  # one warm-up step, one timed step (~14 s each): the reference arm's code path
        arm.close()
        cpu_base = arm.describe(arm.workers * W * H / 1e6 / t)

    # host placement (after the CPU baseline, which uses every core): CPU time and pinned staging on the GPU's NUMA node
    from steganosaurus_b200 import shard as _shard
    placement = _shard.bind_host_to_gpu(dev_index) if os.environ.get("TFFT_NO_NUMA_BIND") is None else {"numa_node": None}
    placement["cuda_device"] = dev_index

    # ---- synthetic workload (seeded)
    # the real keyed turtlewalk (host C++, ~1.7 s for 1.72 M bins at 4096^2; cover-independent, shared by the batch)
    from steganosaurus_b200 import host
    bins_np = host.walk(b"correct horse battery staple", PH, PW, nbits, PARAMS["rmin"], PARAMS["rmax"], 0.7)[0]
    bits_np = make_frame_bits(B, args.payload, 2000 + rank)
    covers_np = make_covers(B, W, H)
    ctx = sb.Context(dev_index)

    d_cover = torch.from_numpy(covers_np).to(dev)
    d_bins = torch.from_numpy(bins_np.view(np.int32)).to(dev)
    d_bits = torch.from_numpy(bits_np).to(dev)
    d_stego = torch.empty_like(d_cover)
    d_usable = torch.zeros(B, dtype=torch.int64, device=dev)
    d_median = torch.zeros(B, 3, dtype=torch.float64, device=dev)
    d_hdr = torch.zeros(B, 38, dtype=torch.uint8, device=dev)
    d_pay = torch.zeros(B, npay, dtype=torch.uint8, device=dev)

    def step_dev():
        ctx.embed_batch_dev(d_cover, d_bins, d_bits, d_stego, usable=d_usable, median=d_median, **PARAMS)
        ctx.extract_frame_dev(d_stego, d_bins, 912, d_hdr, d_pay, alpha=PARAMS["alpha"], center=PARAMS["center"])

    from steganosaurus_b200 import shard
    timing = shard.Timing(dist, dev)  # the helpers tests/test_shard_gloo.py exercises with gloo

    def barrier():
        timing.barrier()
        torch.cuda.synchronize()

    max_over_ranks = timing.max_over_ranks

    # ---- device-resident leg ("value")
    for _ in range(args.warmup):
        step_dev()
    barrier()
    sampler = ClockSampler(dev_index)
    if rank == 0:
        sampler.start()
    ctx.profile_reset()
    ctx.profile_enable(True)
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_dev()
    e1.record()
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    launches = ctx.launches - launches0
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    usable_min = int(d_usable.min().item())
    changed = float((d_stego[0] != d_cover[0]).float().mean().item())

    # ---- the plain FFT passes by themselves (the second half of BASELINE.json's metric: "FFT-pass HBM GB/s % peak"): one 1-D
    # pass of fft2d S:359-366 over a batch of dense complex planes, in place, every element read once and written once
    # (32 bytes per element) -- the reference's own formulation of a pass, without the half-spectrum / zero-row / window
    # savings the embed+extract step above takes.  24 planes of PH x PW = 6.4 GB at 4096^2: far larger than L2.
    del d_cover, d_stego, d_bits
    torch.cuda.empty_cache()
    fft_passes = None
    if rank == 0 and max(PH, PW) <= 4096:
        npl = 24
        fft_passes = {"planes": npl, "PH": PH, "PW": PW, "bytes_per_pass": 32 * npl * PH * PW, "unit": "GB/s"}
        try:  # (an extra: whatever happens here must not cost the line its headline numbers)
            planes = torch.view_as_complex(torch.empty(npl, PH, PW, 2, dtype=torch.float64, device=dev).normal_())
            for nm, axis, inv in (("row_fwd", 0, False), ("col_fwd", 1, False), ("col_inv", 1, True), ("row_inv", 0, True)):
                for _ in range(2):
                    ctx.fft_pass_dev(planes, axis, inv)
                p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                p0.record()
                for _ in range(3):
                    ctx.fft_pass_dev(planes, axis, inv)
                p1.record()
                torch.cuda.synchronize()
                gbs = 32 * npl * PH * PW / 1e9 / (p0.elapsed_time(p1) / 3 / 1e3)
                fft_passes[nm] = round(gbs, 1)
            del planes
        except Exception as e:  # noqa: BLE001
            fft_passes["error"] = repr(e)[:200]
        torch.cuda.empty_cache()

    # ---- end-to-end leg: host buffers through the C-ABI, pinned memory, copies inside the timed region
    h_cover = torch.from_numpy(covers_np).pin_memory()
    # frame bits cross the link packed eight to a byte, MSB first (tfft_embed_batch_packed; the order of S:447-459)
    h_bits = torch.from_numpy(np.packbits(bits_np, axis=1)).pin_memory()
    h_stego = torch.empty_like(h_cover).pin_memory()
    del covers_np, bits_np
    hc, hb, hs = h_cover.numpy(), h_bits.numpy(), h_stego.numpy()

    def step_e2e():
        ctx.embed_batch(hc, bins_np, hb, out=hs, packed=True, **PARAMS)
        return ctx.extract_frame(hs, bins_np, 912, alpha=PARAMS["alpha"], center=PARAMS["center"])

    e2e_steps = max(1, args.steps)
    if args.no_e2e:
        e2e_steps = 0
    for _ in range(min(args.warmup, 2) if e2e_steps else 0):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        hdr, pay, _ = step_e2e()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e2e_ms = max_over_ranks((t1 - t0) * 1e3) / max(1, e2e_steps)
    img_bytes = W * H * 3
    h2d = B * (img_bytes + (nbits + 7) // 8) + B * img_bytes + 2 * 4 * nbits
    d2h = B * img_bytes + B * 8 + B * (38 + npay)

    # ---- what the host<->device links of this node allow for exactly these bytes: the same pinned buffers copied up and
    # down concurrently on two streams, no kernels, all ranks at once (the e2e leg cannot beat this)
    ceil_ms = None
    if e2e_steps:
        d_a, d_b = torch.empty_like(h_cover, device=dev), torch.empty_like(h_cover, device=dev)
        d_c = torch.empty_like(h_bits, device=dev)
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

        def copies():
            with torch.cuda.stream(s_up):
                d_a.copy_(h_cover, non_blocking=True); d_c.copy_(h_bits, non_blocking=True); d_a.copy_(h_cover, non_blocking=True)
            with torch.cuda.stream(s_dn):
                h_stego.copy_(d_b, non_blocking=True)

        copies()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            copies()
        torch.cuda.synchronize()
        ceil_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / 3
        del d_a, d_b, d_c

    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0

    # ---- roofline of the dominant kernel (max device time inside the timed region)
    peak, peak_src = measured_peak_hbm()
    dom = max(prof.items(), key=lambda kv: kv[1][1])
    name, (groups, ms, nbytes) = dom
    ach = (nbytes / 1e9) / (ms / 1e3) if ms > 0 else 0.0
    kernels = {k: {"groups": v[0], "ms": round(v[1], 3), "GBps": round((v[2] / 1e9) / (v[1] / 1e3), 1) if v[1] > 0 else None}
               for k, v in prof.items() if v[0]}
    traffic, other_pipes = None, None
    tpath = os.path.join(ROOT, "profiles", "dominant_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath)).get("kernels", {}).get(name)
            if tj and groups:
                # ncu-measured DRAM bytes per image x images in one launch group of this run
                traffic = tj["dram_bytes_per_image"] * (nbytes / groups) / tj["algorithmic_bytes_per_image"]
                other_pipes = tj.get("other_pipes")  # what else the capture says is busy (the kernel is FP64 work)
        except Exception:
            pass
    mp_per_step = B * W * H / 1e6   # per rank; every rank runs the same batch size (weak scaling)
    line = {
        "metric": METRIC, "value": world * mp_per_step / (dev_ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"batch of {B} 4K UHD RGB covers {W}x{H} (pad {PW}x{PH}) per GPU, {args.payload}-byte frame "
                               f"({nbits} bits, one shared bin list), embed+extract", "images_per_gpu_per_step": B,
                   "l2": f"inputs larger than L2 ({B * img_bytes / 1e9:.1f} GB covers + {3 * PW * PH * 16 / 1e9:.2f} GB spectra per image)",
                   "fft_impl": os.environ.get("TFFT_FFT_IMPL", "default"), "usable_min_bits": usable_min,
                   "stego_pixels_changed": round(changed, 4), "bins": "keyed turtlewalk (host), density 0.7",
                   "host_placement": placement},
        "clocks": clocks,
        "e2e": {"value": world * mp_per_step / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                # the copies alone (same buffers, both directions at once, all ranks): the ceiling of this leg on this node
                "copy_ceiling_ms_per_step": ceil_ms, "copy_ceiling_value": world * mp_per_step / (ceil_ms / 1e3) if ceil_ms else None,
                "frac_of_copy_ceiling": (ceil_ms / e2e_ms) if ceil_ms else None} if e2e_steps else None,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s",
                     "frac": ach / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                     "launch_groups": groups, "avg_ms_per_group": ms / groups if groups else None,
                     "algorithmic_bytes_per_group": nbytes / groups if groups else None, "other_pipes_ncu": other_pipes,
                     "kernels": kernels},
        "cpu_baseline": cpu_base,
    }
    if fft_passes:
        fft_passes["frac_of_peak"] = {k: round(v / peak, 3) for k, v in fft_passes.items() if k.endswith(("_fwd", "_inv"))} if peak else None
        line["fft_passes"] = fft_passes
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--width", type=int, default=W_UHD)
    ap.add_argument("--height", type=int, default=H_UHD)
    ap.add_argument("--payload", type=int, default=PAYLOAD)
    ap.add_argument("--config", default="c3", choices=["c3", "c4", "c5"],
                    help="c3 (default): BASELINE's headline, the 4K UHD embed+extract batch; c4: extract-only sweep, one line per N; "
                         "c5: one 16K x 16K image, slab-decomposed 2-D FFT over the ranks (strong scaling)")
    ap.add_argument("--size", type=int, default=16384, help="--config c5: image width = height")
    ap.add_argument("--transport", default="peer", choices=["peer", "collective"],
                    help="--config c5: peer-mapped slabs over NVLink (CUDA IPC) or NCCL all-to-all")
    ap.add_argument("--sizes", default="512,1024,2048,4096,8192", help="--config c4: the N of the N x N stego batches")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg (the line then has no e2e)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print("bench.py: note -- warmup < 3 breaks the timing rules; use only for profiling runs", file=sys.stderr)
    return run_reference_arm(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
