"""world_size-2 gloo tests (CPU) of the N>1 path: image sharding, timing barrier, max-over-ranks
and whole-job throughput aggregation -- the same helpers bench.py uses under torchrun/NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from steganosaurus_b200 import shard


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 256, 257, 1000):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard.shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
                for i in range(lo, hi):
                    assert shard.owner_of(i, n, world) == r
            assert seen == list(range(n))
            sizes = [shard.shard_range(n, r, world) for r in range(world)]
            assert max(h - l for l, h in sizes) - min(h - l for l, h in sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        t = shard.Timing(dist, torch.device("cpu"))
        assert t.world == world and t.rank == rank
        lo, hi = shard.shard_range(257, rank, world)
        t.barrier()
        # rank 1 is "slower": the job time is the max over ranks, the work is the sum
        secs = 1.0 + rank
        assert t.max_over_ranks(secs) == float(world)
        total = t.sum_over_ranks(hi - lo)
        assert total == 257.0
        thr = shard.aggregate_throughput(hi - lo, secs, t)
        assert abs(thr - 257.0 / world) < 1e-12
        # every rank works on its own images only: gather the ranges and check disjointness on rank 0
        ranges = [None] * world
        dist.all_gather_object(ranges, (lo, hi))
        if rank == 0:
            flat = sorted(ranges)
            assert flat[0][0] == 0 and flat[-1][1] == 257
            assert all(flat[i][1] == flat[i + 1][0] for i in range(world - 1))
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_timing_and_sharding():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=90) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
    assert res == {0: "ok", 1: "ok"}, res


# ---------------------------------------------------------------- slab-decomposed 2-D FFT (config 5) on gloo
def _slab_worker(rank, world, port, q):
    """The product's SlabPlan + CollectiveTransport (zero-copy all_to_all per plane) around a numpy restatement of the
    layout kernels (tests/slab_ref.py): forward half spectrum, embed on the owner's slab, inverse, against the dense
    single-process computation."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import math
        import numpy as np
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import slab_ref
        from steganosaurus_b200 import slab
        W, H = 600, 500  # pads to 1024 x 512
        plan = slab.SlabPlan(W, H, world, rank)
        rng = np.random.default_rng(5)
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        tr = slab.CollectiveTransport(dist)
        send = torch.zeros(3, plan.G, plan.R, plan.cols, dtype=torch.complex128)
        colslab = torch.zeros(3, plan.PH, plan.cols, dtype=torch.complex128)
        tiles = torch.zeros(3, plan.G, plan.R, plan.cols, dtype=torch.complex128)
        # the targets the CUDA split kernel is given must describe exactly this send layout
        ptrs, pstride, row_base = tr.forward_targets(plan, 1 << 20, 0)
        assert ptrs == [(1 << 20) + d * plan.R * plan.cols * 16 for d in range(plan.G)]
        assert pstride == plan.G * plan.R * plan.cols and row_base == plan.y0
        Z = slab_ref.row_pass(slab_ref.pack_pairs(img[plan.y0:plan.y0 + plan.nrows], plan, center=True))
        send.copy_(torch.from_numpy(slab_ref.split_scatter(Z, plan)))
        tr.forward_exchange(plan, send, colslab)
        col = colslab.numpy()
        col[...] = np.fft.ifft(col, axis=1) * plan.PH   # column pass
        # dense reference
        pad = np.zeros((3, plan.PH, plan.PW))
        yy, xx = np.mgrid[0:H, 0:W]
        sgn = np.where((xx + yy) & 1, -1.0, 1.0)
        for p in range(3):
            pad[p, :H, :W] = img[:, :, p] * sgn
        F = np.fft.ifft2(pad) * (plan.PH * plan.PW)
        assert np.allclose(col[:, :, :min(plan.cols, plan.PW // 2 + 1 - plan.col0)],
                           F[:, :, plan.col0:plan.col0 + plan.cols][:, :, :min(plan.cols, plan.PW // 2 + 1 - plan.col0)], atol=1e-6)
        # embed one bin per rank's column range on the owner, inverse, compare with the dense embed
        bins = [(0, 3, 5), (1, 7, plan.cols + 3), (2, plan.PH - 2, 9)]
        dense = F.copy()
        for (p, y, x), bit in zip(bins, (1, 0, 1)):
            mag = abs(dense[p, y, x])
            nv = complex(mag * math.cos(0.5), mag * math.sin(0.5) * (1 if bit else -1))
            dense[p, y, x] = nv
            dense[p, (plan.PH - y) % plan.PH, (plan.PW - x) % plan.PW] = nv.conjugate()
            if plan.col0 <= x < plan.col0 + plan.cols:
                col[p, y, x - plan.col0] = nv
        col[...] = np.fft.fft(col, axis=1) / plan.PH     # inverse column pass
        tr.inverse_exchange(plan, colslab, tiles)
        rows = slab_ref.pairs_to_rows(slab_ref.row_pass(slab_ref.merge_tiles(tiles.numpy(), plan), inverse=True), plan)
        want = (np.fft.fft2(dense) / (plan.PH * plan.PW)).real
        got = rows.transpose(2, 0, 1)
        assert np.allclose(got, want[:, plan.y0:plan.y0 + plan.R], atol=1e-6)
        assert plan.exchange_bytes_per_plane() == (world - 1) * plan.R * plan.cols * 16
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_slab_fft_transpose_and_embed():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_slab_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=90) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
    assert res == {0: "ok", 1: "ok"}, res
