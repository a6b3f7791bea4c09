"""numpy restatement of the data-layout kernels of csrc/tfft_slab.cu (TEST INFRASTRUCTURE): the CPU gloo test runs the
product's CollectiveTransport over these to check the layout conventions the CUDA kernels and the transports share."""
import numpy as np


def pack_pairs(rows_u8, plan, center=False):
    """[nrows][W][3] u8 -> z [3][R/2][PW] complex (slab_pack_pairs)."""
    z = np.zeros((3, plan.R // 2, plan.PW), np.complex128)
    full = np.zeros((plan.R, plan.PW, 3), np.float64)
    full[:plan.nrows, :plan.W] = rows_u8
    if center:
        yy = (np.arange(plan.R) + plan.y0)[:, None]
        xx = np.arange(plan.PW)[None, :]
        full = np.where(((xx + yy) & 1)[..., None].astype(bool), -full, full)
    for p in range(3):
        z[p] = full[0::2, :, p] + 1j * full[1::2, :, p]
    return z


def row_pass(z, inverse=False):
    n = z.shape[-1]
    return np.fft.fft(z, axis=-1) / n if inverse else np.fft.ifft(z, axis=-1) * n  # reference sign: forward = N * ifft (S:347)


def split_scatter(Z, plan):
    """Z [3][R/2][PW] -> send [3][G][R][cols] (slab_split_scatter with the send-buffer targets of CollectiveTransport)."""
    PW, h = plan.PW, plan.PW // 2
    k = np.arange(h + 1)
    zk, zn = Z[..., k], np.conj(Z[..., (PW - k) % PW])
    A, B = (zk + zn) / 2, (zk - zn) / 2j
    half = np.zeros((3, plan.R, plan.ld), np.complex128)
    half[:, 0::2, :h + 1] = A
    half[:, 1::2, :h + 1] = B
    return np.ascontiguousarray(half.reshape(3, plan.R, plan.G, plan.cols).transpose(0, 2, 1, 3))


def merge_tiles(tiles, plan):
    """tiles [3][G][R][cols] -> Z [3][R/2][PW] (slab_merge_tiles)."""
    PW, h = plan.PW, plan.PW // 2
    half = tiles.transpose(0, 2, 1, 3).reshape(3, plan.R, plan.ld)
    A, B = half[:, 0::2], half[:, 1::2]
    Z = np.zeros((3, plan.R // 2, PW), np.complex128)
    Z[..., :h + 1] = A[..., :h + 1] + 1j * B[..., :h + 1]
    k = np.arange(h + 1, PW)
    Z[..., k] = np.conj(A[..., PW - k]) + 1j * np.conj(B[..., PW - k])
    return Z


def pairs_to_rows(z, plan):
    """z [3][R/2][PW] after the inverse row pass -> real rows [R][PW][3]."""
    out = np.zeros((plan.R, plan.PW, 3))
    for p in range(3):
        out[0::2, :, p] = z[p].real
        out[1::2, :, p] = z[p].imag
    return out
