"""Slab-decomposed 2-D FFT for single images too large to be worth replicating (BASELINE config 5:
one 16384 x 16384 RGB image on 2/4/8 GPUs).

Rank g of G owns the row slab  rows [g*PH/G, (g+1)*PH/G)  of every plane:
    forward : row FFT (local) -> all-to-all transpose -> column FFT (local)  => column slabs
    inverse : column IFFT (local) -> all-to-all transpose back -> row IFFT (local) => row slabs
The exchange is the only collective on the path (one per 2-D FFT direction); it runs through
torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU test).  The 1-D passes are the library's
own kernels (Context.fft_pass_dev); `pass_fn` is injectable so the exchange logic can be tested on
CPU tensors.

Embedding on column slabs: the owner of column x writes the bins with that x; the Hermitian mirror
(PH-y, PW-x) belongs to another rank, which writes conj(nv) using the magnitude of its own mirror
element (equal to the primary's up to the FFT's ~1e-13 rounding asymmetry) -- no extra exchange.
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import torch


class SlabFFT2D:
    def __init__(self, dist, PH: int, PW: int, pass_fn: Callable, device=None):
        """pass_fn(x[n, rows, cols] complex128 contiguous, axis, inverse) transforms in place."""
        self.dist = dist
        self.G = dist.get_world_size() if dist is not None else 1
        self.g = dist.get_rank() if dist is not None else 0
        if PH % self.G or PW % self.G:
            raise ValueError("PH and PW must be divisible by the number of ranks")
        self.PH, self.PW, self.pass_fn, self.device = PH, PW, pass_fn, device
        self.rows, self.cols = PH // self.G, PW // self.G

    # ---- the transpose exchange ---------------------------------------------------------------
    def rows_to_cols(self, x: torch.Tensor) -> torch.Tensor:
        """[n, PH/G, PW] (my rows, all columns) -> [n, PH, PW/G] (all rows, my columns)."""
        n = x.shape[0]
        if self.G == 1:
            return x
        send = x.view(n, self.rows, self.G, self.cols).permute(2, 0, 1, 3).contiguous()  # [G][n][rows][cols]
        recv = torch.empty_like(send)
        self.dist.all_to_all_single(recv, send)
        return recv.permute(1, 0, 2, 3).reshape(n, self.PH, self.cols).contiguous()

    def cols_to_rows(self, y: torch.Tensor) -> torch.Tensor:
        """[n, PH, PW/G] -> [n, PH/G, PW]."""
        n = y.shape[0]
        if self.G == 1:
            return y
        send = y.view(n, self.G, self.rows, self.cols).permute(1, 0, 2, 3).contiguous()  # [G][n][rows][cols]
        recv = torch.empty_like(send)
        self.dist.all_to_all_single(recv, send)
        return recv.permute(1, 2, 0, 3).reshape(n, self.rows, self.PW).contiguous()

    # ---- distributed transforms -----------------------------------------------------------------
    def forward(self, x_rows: torch.Tensor) -> torch.Tensor:
        self.pass_fn(x_rows, 0, False)
        y = self.rows_to_cols(x_rows)
        self.pass_fn(y, 1, False)
        return y

    def inverse(self, y_cols: torch.Tensor) -> torch.Tensor:
        self.pass_fn(y_cols, 1, True)
        x = self.cols_to_rows(y_cols)
        self.pass_fn(x, 0, True)
        return x

    # ---- image <-> slab helpers ----------------------------------------------------------------
    def planes_from_u8_rows(self, img_rows: torch.Tensor, W: int, H: int, center: bool = False) -> torch.Tensor:
        """my rows of the u8 image [rows_here, W, 3] -> zero-padded complex planes [3, PH/G, PW] (S:383-398)."""
        out = torch.zeros(3, self.rows, self.PW, dtype=torch.complex128, device=img_rows.device)
        r = img_rows.shape[0]
        if r:
            pl = img_rows.permute(2, 0, 1).to(torch.float64)
            if center:
                y0 = self.g * self.rows
                yy = torch.arange(y0, y0 + r, device=img_rows.device)[:, None]
                xx = torch.arange(W, device=img_rows.device)[None, :]
                pl = torch.where(((xx + yy) & 1).bool()[None], -pl, pl)
            out[:, :r, :W] = pl
        return out

    def u8_rows_from_planes(self, x_rows: torch.Tensor, W: int, H: int, center: bool = False) -> torch.Tensor:
        """real part, crop, centre, round half away from zero, clamp (S:399-403, S:387-391)."""
        y0 = self.g * self.rows
        r = max(0, min(self.rows, H - y0))
        v = x_rows.real[:, :r, :W]
        if center and r:
            yy = torch.arange(y0, y0 + r, device=v.device)[:, None]
            xx = torch.arange(W, device=v.device)[None, :]
            v = torch.where(((xx + yy) & 1).bool()[None], -v, v)
        q = torch.sign(v) * torch.floor(torch.abs(v) + 0.5)
        return q.clamp_(0, 255).to(torch.uint8).permute(1, 2, 0).contiguous()

    # ---- phase write / read on column slabs (S:712-746) ---------------------------------------------
    def embed_on_cols(self, y_cols: torch.Tensor, bins: torch.Tensor, bits: torch.Tensor, alpha: float) -> None:
        """bins int64 packed plane<<30 | y*PW + x (shared), bits 0/1; in place on my column slab."""
        PW, PH, x0 = self.PW, self.PH, self.g * self.cols
        p = bins >> 30
        lin = bins & 0x3FFFFFFF
        yy, xx = lin // PW, lin % PW
        ca, sa = math.cos(alpha), math.sin(alpha)
        sgn = bits.to(torch.float64) * 2.0 - 1.0
        # primary bins in my columns
        m = (xx >= x0) & (xx < x0 + self.cols)
        if bool(m.any()):
            z = y_cols[p[m], yy[m], xx[m] - x0]
            mag = torch.clamp(torch.abs(z), min=1e-12)
            y_cols[p[m], yy[m], xx[m] - x0] = torch.complex(mag * ca, mag * sa * sgn[m])
        # mirrors in my columns
        cy, cx = (PH - yy) % PH, (PW - xx) % PW
        m = (cx >= x0) & (cx < x0 + self.cols)
        if bool(m.any()):
            z = y_cols[p[m], cy[m], cx[m] - x0]
            mag = torch.clamp(torch.abs(z), min=1e-12)
            y_cols[p[m], cy[m], cx[m] - x0] = torch.complex(mag * ca, -mag * sa * sgn[m])

    def read_on_cols(self, y_cols: torch.Tensor, bins: torch.Tensor) -> torch.Tensor:
        """raw bits (Im >= 0, ties -> 1) for the bins in my columns, -1 elsewhere; combine with a MAX all-reduce."""
        PW, x0 = self.PW, self.g * self.cols
        p = bins >> 30
        lin = bins & 0x3FFFFFFF
        yy, xx = lin // PW, lin % PW
        out = torch.full(bins.shape, -1, dtype=torch.int32, device=bins.device)
        m = (xx >= x0) & (xx < x0 + self.cols)
        if bool(m.any()):
            z = y_cols[p[m], yy[m], xx[m] - x0]
            out[m] = (~(z.imag < 0)).to(torch.int32)
        if self.dist is not None and self.G > 1:
            self.dist.all_reduce(out, op=self.dist.ReduceOp.MAX)
        return out


def library_pass_fn(ctx):
    """1-D passes through the CUDA library (Context.fft_pass_dev)."""
    def f(x, axis, inverse):
        assert x.is_contiguous() and x.dtype == torch.complex128
        ctx.fft_pass_dev(x, axis, inverse)
    return f


def torch_pass_fn():
    """CPU stand-in for the gloo test of the exchange logic (reference sign: forward = N * ifft)."""
    def f(x, axis, inverse):
        dim = 2 if axis == 0 else 1
        n = x.shape[dim]
        x.copy_(torch.fft.fft(x, dim=dim) / n if inverse else torch.fft.ifft(x, dim=dim) * n)
    return f
