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
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from steganosaurus_b200 import slab
        PH, PW, n = 16, 32, 3
        g = torch.Generator().manual_seed(5)
        full = torch.randn(n, PH, PW, dtype=torch.float64, generator=g).to(torch.complex128)  # real planes
        sf = slab.SlabFFT2D(dist, PH, PW, slab.torch_pass_fn())
        mine = full[:, rank * sf.rows:(rank + 1) * sf.rows, :].contiguous().clone()
        ycols = sf.forward(mine)
        want = torch.fft.ifft2(full) * (PH * PW)  # reference forward convention (S:347)
        assert torch.allclose(ycols, want[:, :, rank * sf.cols:(rank + 1) * sf.cols], atol=1e-9)
        # embed two bins (one whose mirror lives on the other rank), inverse, check against the dense computation
        bins = torch.tensor([(0 << 30) | (2 * PW + 3), (1 << 30) | (1 * PW + 5), (2 << 30) | (3 * PW + 2)], dtype=torch.int64)
        bits = torch.tensor([1, 0, 1])
        dense = want.clone()
        import math
        for b, bit in zip(bins.tolist(), bits.tolist()):
            p, y, x = b >> 30, (b & 0x3FFFFFFF) // PW, (b & 0x3FFFFFFF) % PW
            mag = max(1e-12, abs(dense[p, y, x].item()))
            nv = complex(mag * math.cos(0.5), mag * math.sin(0.5) * (1 if bit else -1))
            dense[p, y, x] = nv
            dense[p, (PH - y) % PH, (PW - x) % PW] = nv.conjugate()
        sf.embed_on_cols(ycols, bins, bits, 0.5)
        assert torch.allclose(ycols, dense[:, :, rank * sf.cols:(rank + 1) * sf.cols], atol=1e-9)
        raw = sf.read_on_cols(ycols, bins)
        assert raw.tolist() == bits.tolist()
        back = sf.inverse(ycols)
        want_back = torch.fft.fft2(dense) / (PH * PW)
        assert torch.allclose(back, want_back[:, rank * sf.rows:(rank + 1) * sf.rows, :], atol=1e-9)
        assert back.imag.abs().max() < 1e-9  # Hermitian by construction: the image stays real
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
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
