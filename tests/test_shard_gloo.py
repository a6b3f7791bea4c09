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
