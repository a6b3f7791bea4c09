"""BASELINE config 5: one image too large to be worth replicating (16384 x 16384 RGB), its 2-D FFT slab-decomposed over
G GPUs with ONE exchange per direction over NVLink (fft2d S:359-366 distributed).

Rank g owns image rows [g R, (g+1) R); after the forward exchange it owns the column slab [3][PH][cols] of the
half spectrum (include/tfft.h, "config 5").  Everything that touches pixels or spectra is a CUDA kernel of the library
(csrc/tfft_slab.cu + the FFT passes); this module only sequences the calls and owns the transport:

* ``PeerTransport``       CUDA-IPC-mapped slabs: the split kernel of the forward row pass stores every element straight
                          into the column slab of the GPU that owns its column (the exchange IS the kernel's store
                          stream, over NVLink); the inverse pushes contiguous tiles with peer copies.  The only
                          collective left is a stream-ordered barrier.
* ``CollectiveTransport`` all_to_all_single per plane on a send buffer laid out [plane][dst][R][cols] so that neither side
                          packs or unpacks (NCCL on GPUs; the same code runs on gloo / CPU tensors in the tests).
* ``LocalTransport``      G virtual ranks in one process on one GPU (tests, and G = 1).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional

import numpy as np

from . import _lib


def next_pow2(v: int) -> int:
    p = 1
    while p < v:
        p <<= 1
    return p


class SlabPlan:
    """Geometry shared by the kernels and the transports (mirrors tfft_slab_sizes)."""

    def __init__(self, W: int, H: int, G: int, g: int = 0):
        if G not in (1, 2, 4, 8) or not (0 <= g < G):
            raise ValueError("G must be 1, 2, 4 or 8 and 0 <= g < G")
        self.W, self.H, self.G, self.g = W, H, G, g
        self.PW, self.PH = next_pow2(W), next_pow2(H)
        self.R = self.PH // G
        self.ld = self.PW // 2 + 16
        if self.R < 2 or self.ld % G:
            raise ValueError("unsupported slab geometry")
        self.cols = self.ld // G
        self.y0 = g * self.R
        self.nrows = max(0, min(self.R, H - self.y0))   # rows of the image inside my slab
        self.col0 = g * self.cols

    # element counts (complex doubles)
    @property
    def colslab_elems(self) -> int:
        return 3 * self.PH * self.cols

    @property
    def tiles_elems(self) -> int:
        return 3 * self.G * self.R * self.cols

    def exchange_bytes_per_plane(self) -> int:
        """Bytes this rank sends to OTHER ranks per plane and direction (the NVLink traffic of the exchange)."""
        return (self.G - 1) * self.R * self.cols * 16


# ------------------------------------------------------------------------------------------------ transports
class CollectiveTransport:
    """One all_to_all_single per plane.  send / recv layouts are chosen so that both ends are zero-copy:
    forward  send[p][d][r][c] = (plane p, my row r, column d*cols + c)      -> colslab[p] = recv viewed [PH][cols]
    inverse  colslab[p] viewed [d][R][cols] is the send buffer              -> tiles[p][s][r][c]"""

    kind = "collective"

    def __init__(self, dist):
        self.dist = dist

    def forward_targets(self, plan: SlabPlan, send_ptr: int, colslab_ptr: int):
        per = plan.R * plan.cols * 16
        return [send_ptr + d * per for d in range(plan.G)], plan.G * plan.R * plan.cols, plan.y0

    def forward_exchange(self, plan, send, colslab):
        for p in range(3):
            self.dist.all_to_all_single(colslab[p].reshape(plan.G, -1), send[p].reshape(plan.G, -1))

    def inverse_exchange(self, plan, colslab, tiles):
        for p in range(3):
            self.dist.all_to_all_single(tiles[p].reshape(plan.G, -1), colslab[p].reshape(plan.G, -1))

    def barrier(self):
        pass

    needs_send_buffer = True


class PeerTransport:
    """Peer-mapped slabs (CUDA IPC): rank g holds device pointers to every rank's column slab and tile buffer."""

    kind = "peer"
    needs_send_buffer = False

    def __init__(self, dist, device_index: int, colslab_ptr: int, tiles_ptr: int):
        import torch
        self.dist, self.torch = dist, torch
        L = _lib.load()
        G, g = dist.get_world_size(), dist.get_rank()
        h1, h2 = C.create_string_buffer(64), C.create_string_buffer(64)
        if L.tfft_ipc_export(colslab_ptr, h1) or L.tfft_ipc_export(tiles_ptr, h2):
            raise RuntimeError("cudaIpcGetMemHandle failed")
        mine = (bytes(h1.raw), bytes(h2.raw))
        allh = [None] * G
        dist.all_gather_object(allh, mine)
        self.col_ptrs, self.tile_ptrs, self._opened = [], [], []
        for r in range(G):
            if r == g:
                self.col_ptrs.append(colslab_ptr); self.tile_ptrs.append(tiles_ptr)
                continue
            ptrs = []
            for h in allh[r]:
                p = C.c_void_p()
                if L.tfft_ipc_open(device_index, h, C.byref(p)):
                    raise RuntimeError("cudaIpcOpenMemHandle failed (no peer access between these GPUs?)")
                ptrs.append(p.value); self._opened.append(p.value)
            self.col_ptrs.append(ptrs[0]); self.tile_ptrs.append(ptrs[1])
        self._flag = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", device_index))

    def forward_targets(self, plan, send_ptr, colslab_ptr):
        return list(self.col_ptrs), plan.PH * plan.cols, 0

    def barrier(self):
        """Stream-ordered: completes on a rank only when every rank has reached it on its stream, i.e. when every kernel
        and copy enqueued before it (the peer stores) has finished."""
        self.dist.all_reduce(self._flag)

    def forward_exchange(self, plan, send, colslab):
        self.barrier()   # the split kernels of all ranks have stored into my slab

    def inverse_exchange(self, plan, colslab, tiles):
        """Push rows [d R, (d+1) R) of my slab into tile g of rank d: 3 contiguous copies of R * cols elements per peer."""
        torch = self.torch
        n = plan.R * plan.cols
        for k in range(plan.G):
            d = (plan.g + k) % plan.G  # start with myself, then round-robin so the ranks do not all hit the same peer
            for p in range(3):
                dst = _as_tensor(self.tile_ptrs[d] + ((p * plan.G + plan.g) * n) * 16, n, colslab.device)
                dst.copy_(colslab[p].reshape(plan.G, n)[d], non_blocking=True)
        self.barrier()

    def close(self):
        L = _lib.load()
        for p in self._opened:
            L.tfft_ipc_close(p)
        self._opened = []


class LocalTransport:
    """All G virtual ranks live in this process on one device: "peer" pointers are plain local pointers."""

    kind = "local"
    needs_send_buffer = False

    def __init__(self):
        self.engines: List["SlabEngine"] = []

    def forward_targets(self, plan, send_ptr, colslab_ptr):
        return [e.colslab.data_ptr() for e in self.engines], plan.PH * plan.cols, 0

    def forward_exchange(self, plan, send, colslab):
        pass

    def inverse_exchange(self, plan, colslab, tiles):
        n = plan.R * plan.cols
        for d, e in enumerate(self.engines):
            for p in range(3):
                e.tiles[p, plan.g].reshape(-1).copy_(colslab[p].reshape(plan.G, n)[d])

    def barrier(self):
        pass


class _Cai:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n, 2), "typestr": "<f8", "data": (ptr, False), "version": 3}


def _as_tensor(ptr: int, n_complex: int, device):
    import torch
    return torch.view_as_complex(torch.as_tensor(_Cai(ptr, n_complex), device=device))


# ------------------------------------------------------------------------------------------------ engine
class SlabEngine:
    """One rank of the slab path.  Buffers are torch CUDA tensors; every compute step is one C-ABI call."""

    def __init__(self, ctx, W: int, H: int, G: int, g: int, transport=None, dist=None, transport_kind: str = "peer"):
        import torch
        self.torch, self.ctx, self.L = torch, ctx, ctx.L
        self.plan = p = SlabPlan(W, H, G, g)
        pw, ph, r, ld, cols = (C.c_int() for _ in range(5))
        rc = self.L.tfft_slab_sizes(W, H, G, *(C.byref(v) for v in (pw, ph, r, ld, cols)))
        if rc or (pw.value, ph.value, r.value, ld.value, cols.value) != (p.PW, p.PH, p.R, p.ld, p.cols):
            raise ValueError(f"slab geometry refused by the library (rc={rc})")
        self.dev = torch.device("cuda", ctx.device)
        self.colslab = torch.zeros(3, p.PH, p.cols, dtype=torch.complex128, device=self.dev)
        self.tiles = torch.zeros(3, G, p.R, p.cols, dtype=torch.complex128, device=self.dev)
        if transport is None:
            if G == 1:
                transport = LocalTransport(); transport.engines.append(self)
            elif transport_kind == "peer":
                transport = PeerTransport(dist, ctx.device, self.colslab.data_ptr(), self.tiles.data_ptr())
            else:
                transport = CollectiveTransport(dist)
        self.tr = transport
        self.send = torch.zeros(3, G, p.R, p.cols, dtype=torch.complex128, device=self.dev) if transport.needs_send_buffer else None

    def _stream(self):
        return self.torch.cuda.current_stream().cuda_stream

    # ---- the six steps (each: one library call or the transport) ------------------------------------------------
    def rows_forward(self, rows_u8, center=False):
        p = self.plan
        ptrs, pstride, row_base = self.tr.forward_targets(p, self.send.data_ptr() if self.send is not None else 0, self.colslab.data_ptr())
        arr = (C.c_void_p * p.G)(*ptrs)
        self.ctx._check(self.L.tfft_slab_rows_forward_dev(self.ctx.h, rows_u8.data_ptr() if p.nrows else None, p.nrows, p.W, p.H, p.G, p.g,
                                                          int(center), arr, pstride, row_base, self._stream()))

    def forward_exchange(self):
        self.tr.forward_exchange(self.plan, self.send, self.colslab)

    def cols(self, inverse: bool):
        p = self.plan
        self.ctx._check(self.L.tfft_slab_cols_dev(self.ctx.h, self.colslab.data_ptr(), p.W, p.H, p.G, int(inverse), self._stream()))

    def embed_bins(self, bins, bits, alpha=0.5):
        p = self.plan
        self.ctx._check(self.L.tfft_slab_embed_dev(self.ctx.h, self.colslab.data_ptr(), p.W, p.H, p.G, p.g, bins.data_ptr(), bits.data_ptr(),
                                                   bins.numel(), alpha, self._stream()))

    def read_bins(self, bins, alpha=0.5):
        p = self.plan
        raw = self.torch.empty(bins.numel(), dtype=self.torch.int8, device=self.dev)
        self.ctx._check(self.L.tfft_slab_read_dev(self.ctx.h, self.colslab.data_ptr(), p.W, p.H, p.G, p.g, bins.data_ptr(), bins.numel(),
                                                  alpha, raw.data_ptr(), self._stream()))
        return raw

    def inverse_exchange(self):
        self.tr.inverse_exchange(self.plan, self.colslab, self.tiles)

    def rows_inverse(self, center=False):
        p = self.plan
        out = self.torch.empty(p.nrows, p.W, 3, dtype=self.torch.uint8, device=self.dev)
        self.ctx._check(self.L.tfft_slab_rows_inverse_dev(self.ctx.h, self.tiles.data_ptr(), p.nrows, p.W, p.H, p.G, p.g, int(center),
                                                          out.data_ptr() if p.nrows else None, self._stream()))
        return out

    # ---- whole operations for one rank of a distributed run (PeerTransport / CollectiveTransport) ----------------
    def embed(self, rows_u8, bins, bits, alpha=0.5, center=False):
        """my rows of the cover -> my rows of the stego image (do_embed S:912-1103 without the capacity gate)."""
        self.tr.barrier()            # nobody still reads the slab / tiles I am about to have overwritten
        self.rows_forward(rows_u8, center)
        self.forward_exchange()
        self.cols(False)
        self.embed_bins(bins, bits, alpha)
        self.cols(True)
        self.inverse_exchange()
        return self.rows_inverse(center)

    def extract_raw(self, rows_u8, bins, alpha=0.5, center=False, dist=None):
        """raw read bits of every bin (S:734-746); with `dist` the ranks' partial lists are combined (MAX over -1 / 0 / 1)."""
        self.tr.barrier()
        self.rows_forward(rows_u8, center)
        self.forward_exchange()
        self.cols(False)
        raw = self.read_bins(bins, alpha)
        if dist is not None and self.plan.G > 1:
            dist.all_reduce(raw, op=dist.ReduceOp.MAX)
        return raw


def run_local(engines: List[SlabEngine], cover_u8, bins, bits, alpha=0.5, center=False, want_spectrum=False):
    """G virtual ranks on one GPU, phase by phase (the tests' stand-in for G processes): returns (stego, raw bits read
    back from it[, the forward half spectrum [3][PH][ld] gathered from the column slabs])."""
    import torch
    G = len(engines)
    rows = [cover_u8[e.plan.y0:e.plan.y0 + e.plan.nrows].contiguous() for e in engines]
    for e, r in zip(engines, rows):
        e.rows_forward(r, center)
    for e in engines:
        e.cols(False)
    spec = torch.cat([e.colslab for e in engines], dim=2).clone() if want_spectrum else None
    for e in engines:
        e.embed_bins(bins, bits, alpha)
        e.cols(True)
    for e in engines:
        e.inverse_exchange()
    stego = torch.cat([e.rows_inverse(center) for e in engines], dim=0)
    srows = [stego[e.plan.y0:e.plan.y0 + e.plan.nrows].contiguous() for e in engines]
    for e, r in zip(engines, srows):
        e.rows_forward(r, center)
    raw = None
    for e in engines:
        e.cols(False)
        part = e.read_bins(bins, alpha)
        raw = part if raw is None else torch.maximum(raw, part)
    return (stego, raw, spec) if want_spectrum else (stego, raw)


def local_group(ctx, W: int, H: int, G: int) -> List[SlabEngine]:
    tr = LocalTransport()
    for g in range(G):
        tr.engines.append(SlabEngine(ctx, W, H, G, g, transport=tr))
    return tr.engines
