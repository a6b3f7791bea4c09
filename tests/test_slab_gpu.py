"""BASELINE config 5 on the GPU: the slab-decomposed 2-D FFT (steganosaurus_b200/slab.py over csrc/tfft_slab.cu).

One GPU: G virtual ranks in one process (LocalTransport) -- every kernel and the whole index math -- against the oracle.
Two or more GPUs (skipped otherwise): one process per GPU, the peer-memory transport (CUDA IPC, NVLink stores) and the
collective transport (NCCL all-to-all) against the single-GPU path of the library, up to 16384 x 16384."""
import os
import socket

import numpy as np
import pytest

from oracle import pyoracle as O
from steganosaurus_b200 import slab, synth
import steganosaurus_b200 as sb

pytestmark = pytest.mark.gpu


def _pixels_ok(a, b):
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    return d.max() <= 1 and (d == 0).mean() >= 0.9999


@pytest.mark.parametrize("W,H,G,center", [(2048, 2048, 1, False), (2048, 2048, 2, False), (2048, 2048, 8, False), (1500, 900, 4, True),
                                          (600, 4096, 2, False), (5000, 520, 2, True)])
def test_slab_virtual_ranks_vs_oracle(ctx, W, H, G, center):
    import torch
    o = O.best()
    PH, PW = synth.next_pow2(H), synth.next_pow2(W)
    nbits = 20000
    cover = synth.gen_texture(W, H, W + 3 * H)
    bins = synth.random_bins(PH, PW, nbits, 7)
    bits = synth.random_bits(1, nbits, 8)[0]
    dev = torch.device("cuda", 0)
    engines = slab.local_group(ctx, W, H, G)
    stego, raw, spec = slab.run_local(engines, torch.from_numpy(cover).to(dev), torch.from_numpy(bins.view(np.int32)).to(dev),
                                      torch.from_numpy(bits).to(dev), 0.5, center, want_spectrum=True)
    torch.cuda.synchronize()
    want = o.embed(cover, bins, bits, 0.5, center)
    assert _pixels_ok(stego.cpu().numpy(), want["stego"])
    _, wraw = o.extract(want["stego"], bins, 1, 0.5, center)
    assert np.array_equal(raw.cpu().numpy().astype(np.uint8), wraw)
    # the gathered column slabs are the reference's forward spectrum (columns 0 .. PW/2)
    F = o.forward_spectrum(cover, center)
    got = spec.cpu().numpy()[:, :, :PW // 2 + 1]
    rms = np.sqrt(np.mean(np.abs(F) ** 2))
    # (the oracle's own twiddle recurrence carries ~len * eps, S:353: looser for the 8192-point rows; north_star allows 1e-9)
    assert np.abs(got - F[:, :, :PW // 2 + 1]).max() / rms < (1e-11 if max(PH, PW) <= 4096 else 1e-10)
    assert np.abs(spec.cpu().numpy()[:, :, PW // 2 + 1:]).max() == 0.0


def test_slab_16k_rows_and_columns_vs_library_path(ctx):
    """16384-point rows and columns (four-step passes) on one GPU: G = 2 virtual ranks on a 16384 x 700 and a 700 x 16384
    image against the library's ordinary single-image path."""
    import torch
    dev = torch.device("cuda", 0)
    for (W, H) in ((16384, 700), (700, 16384)):
        PH, PW = synth.next_pow2(H), synth.next_pow2(W)
        nbits = 9000
        cover = synth.gen_texture(W, H, 5)
        bins = synth.random_bins(PH, PW, nbits, 3)
        bits = synth.random_bits(1, nbits, 4)
        engines = slab.local_group(ctx, W, H, 2)
        stego, raw = slab.run_local(engines, torch.from_numpy(cover).to(dev), torch.from_numpy(bins.view(np.int32)).to(dev),
                                    torch.from_numpy(bits[0]).to(dev))
        one, _, _ = ctx.embed_batch(cover[None], bins, bits)
        assert _pixels_ok(stego.cpu().numpy(), one[0])
        _, raw1 = ctx.extract_bits(one, bins, 1)
        assert np.array_equal(raw.cpu().numpy().astype(np.uint8), raw1[0])
        del engines
        torch.cuda.empty_cache()


# ------------------------------------------------------------------------------------------------ real ranks
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, N, kind, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        nbits = 30000
        cover = synth.gen_cover(N, N, 5)
        bins = synth.random_bins(N, N, nbits, 3)
        bits = synth.random_bits(1, nbits, 4)
        with sb.Context(rank) as c:
            e = slab.SlabEngine(c, N, N, world, rank, dist=dist, transport_kind=kind)
            d_bins = torch.from_numpy(bins.view(np.int32)).to(dev)
            d_bits = torch.from_numpy(bits[0]).to(dev)
            mine = torch.from_numpy(cover[e.plan.y0:e.plan.y0 + e.plan.nrows]).to(dev)
            for _ in range(2):  # twice: buffers are reused across calls
                rows = e.embed(mine, d_bins, d_bits)
                raw = e.extract_raw(rows, d_bins, dist=dist)
            torch.cuda.synchronize()
            parts = [torch.empty_like(rows) for _ in range(world)] if rank == 0 else None
            dist.gather(rows, parts, dst=0)
            if rank == 0:
                stego = torch.cat(parts, 0).cpu().numpy()
                one, _, _ = c.embed_batch(cover[None], bins, bits)
                _, raw1 = c.extract_bits(one, bins, 1)
                assert _pixels_ok(stego, one[0]), "slab stego differs from the single-GPU path"
                assert np.array_equal(raw.cpu().numpy().astype(np.uint8), raw1[0]), "slab raw bits differ"
            if hasattr(e.tr, "close"):
                dist.barrier()
                e.tr.close()
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N,kind", [(2048, "peer"), (2048, "collective"), (16384, "peer")])
def test_slab_two_gpus_vs_single_gpu_path(N, kind):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world, port = 2, _free_port()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_rank_main, args=(r, world, port, N, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=900) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res
