#!/usr/bin/env python
"""Host<->device copy ceiling of the node, one process per GPU (torchrun): what the e2e leg of bench.py can reach at most.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_ceiling.py [--perm interleave]

Every rank moves 1 GiB chunks between pinned host memory and its GPU: H2D alone, D2H alone, both at once, and the
bench's mix (two parts up, one part down, concurrently).  Times are max over ranks between barriers; rank 0 prints one
JSON line with the aggregate GB/s.  --perm interleave maps ranks to devices 0,4,1,5,2,6,3,7 (both halves of an
8-GPU node from N = 2 on)."""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--perm", default="identity", choices=["identity", "interleave"])
    ap.add_argument("--gib", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--wc", action="store_true", help="host buffers from tfft_host_alloc_wc (write-combined pinned memory)")
    a = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    ndev = torch.cuda.device_count()
    perm = [0, 4, 1, 5, 2, 6, 3, 7] if (a.perm == "interleave" and ndev == 8) else list(range(ndev))
    dev_index = perm[local % len(perm)]
    torch.cuda.set_device(dev_index)
    dev = torch.device("cuda", dev_index)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = int(a.gib * (1 << 30))
    if a.wc:  # cudaHostAllocWriteCombined through the runtime torch links (measured: no faster than plain pinned memory here)
        import ctypes
        import numpy as np
        rt = ctypes.CDLL([m.split()[-1] for m in open("/proc/self/maps") if "libcudart" in m][0])

        def wc_tensor():
            p = ctypes.c_void_p()
            assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), 4) == 0
            return torch.from_numpy(np.ctypeslib.as_array((ctypes.c_uint8 * n).from_address(p.value)))
        h_up, h_up2, h_dn = wc_tensor(), wc_tensor(), wc_tensor()
    else:
        h_up, h_up2, h_dn = (torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(3))
    d_up, d_up2, d_dn = (torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(3))
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    flag = torch.zeros(1, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.all_reduce(flag)
            torch.cuda.synchronize()

    def timed(fn):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.reps):
            fn()
        torch.cuda.synchronize()
        t = torch.tensor([(time.perf_counter() - t0) / a.reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def up():
        d_up.copy_(h_up, non_blocking=True)

    def down():
        h_dn.copy_(d_dn, non_blocking=True)

    def both():
        with torch.cuda.stream(s1):
            d_up.copy_(h_up, non_blocking=True)
        with torch.cuda.stream(s2):
            h_dn.copy_(d_dn, non_blocking=True)

    def mix():  # bench.py's e2e step: 13.2 GB up, 6.4 GB down
        with torch.cuda.stream(s1):
            d_up.copy_(h_up, non_blocking=True)
            d_up2.copy_(h_up2, non_blocking=True)
        with torch.cuda.stream(s2):
            h_dn.copy_(d_dn, non_blocking=True)

    t_up, t_dn, t_both, t_mix = timed(up), timed(down), timed(both), timed(mix)
    gb = n / 1e9
    if rank == 0:
        print(json.dumps({"n_gpus": world, "perm": a.perm, "write_combined": bool(a.wc), "devices": perm[:world], "chunk_GiB": a.gib,
                          "h2d_GBps_total": round(world * gb / t_up, 1), "d2h_GBps_total": round(world * gb / t_dn, 1),
                          "both_GBps_total": round(world * 2 * gb / t_both, 1), "mix_2up_1down_GBps_total": round(world * 3 * gb / t_mix, 1),
                          "per_gpu": {"h2d": round(gb / t_up, 1), "d2h": round(gb / t_dn, 1), "both_each_way": round(gb / t_both, 1),
                                      "mix_total": round(3 * gb / t_mix, 1)}}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
