// tfft_slab.cu -- BASELINE config 5: ONE image too large to be worth replicating (16384 x 16384), its 2-D FFT
// slab-decomposed over G GPUs (fft2d S:359-366 distributed).
//
// Rank g owns image rows [g R, (g+1) R), R = PH / G.  Real planes are Hermitian, so the exchanged object is the HALF
// spectrum of the row pass (columns 0 .. PW/2, padded to ld = PW/2 + 16): rank d ends up with the column slab
// [3][PH][cols], cols = ld / G, columns [d cols, (d+1) cols).
//
//   forward : slab_pack_pairs (two image rows -> one complex row z = a + i b, S:383-398 fused)
//             -> row pass over PW points (the library's c2c pass)
//             -> slab_split_scatter: Hermitian split into the two half rows AND the exchange in one kernel -- every
//                16-byte element is stored straight into the column slab of the rank that owns its column (peer-mapped
//                memory over NVLink, or a local send buffer when the transport is NCCL)
//             -> column pass over PH points on [3][PH][cols]
//   embed   : slab_embed_scatter / slab_read_raw on the bins whose column this rank owns (S:712-746)
//   inverse : column pass -> tiles [3][G][R][cols] pushed back (contiguous peer copies) -> slab_merge_tiles (half rows ->
//             complex pair rows) -> row pass -> slab_pairs_to_u8 (S:399-403, S:387-391)
//
// The kernels here are the data-layout glue around the FFT passes of tfft_pencil.cu / tfft_kernels.cu; all of them are
// streaming 16-byte copies with a few flops (HBM-bound; the stores of slab_split_scatter are NVLink-bound).
// Reference citations S:n = steganosaurus/src/steganosaur.cpp line n.
#include "tfft_kernels.cuh"

namespace tfft {

#define TFFT_SLAB_LAUNCH_CHECK(L)                      \
    do {                                               \
        if ((L).launch_counter) ++*(L).launch_counter; \
        cudaError_t e__ = cudaGetLastError();          \
        if (e__ != cudaSuccess) return e__;            \
    } while (0)

// z[plane][j][x] = row(y0 + 2j)[x] + i * row(y0 + 2j + 1)[x]; zero beyond the image (pad_to_fft S:393), centre sign S:392
__global__ void __launch_bounds__(256) slab_pack_pairs(const uint8_t* __restrict__ rows, int nrows, int W, int PW, int npairs, int y0,
                                                       int center, double2* __restrict__ z) {
    const long long total = (long long)3 * npairs * PW;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(t % PW);
        const long long u = t / PW;
        const int j = (int)(u % npairs), p = (int)(u / npairs);
        double a = 0.0, b = 0.0;
        if (x < W) {
            const int ya = 2 * j, yb = 2 * j + 1;
            if (ya < nrows) a = (double)rows[((size_t)ya * W + x) * 3 + p];
            if (yb < nrows) b = (double)rows[((size_t)yb * W + x) * 3 + p];
            if (center) {
                if ((x + y0 + ya) & 1) a = -a;
                if ((x + y0 + yb) & 1) b = -b;
            }
        }
        z[t] = make_double2(a, b);
    }
}

// Z[plane][j][PW] (row pass output of the pair rows) -> half rows A (image row y0 + 2j) and B (y0 + 2j + 1):
//   A[k] = (Z[k] + conj Z[PW-k]) / 2,  B[k] = (Z[k] - conj Z[PW-k]) / 2i,  k = 0 .. PW/2;  pad columns are zero.
// Column k belongs to rank d = k / cols and is stored at dst.p[d] + plane * plane_stride + (row - row_base) * cols + k % cols.
__global__ void __launch_bounds__(256) slab_split_scatter(const double2* __restrict__ Z, int PW, int ld, int cols, int npairs, int y0,
                                                          SlabDst dst, size_t plane_stride, int row_base) {
    const long long total = (long long)3 * npairs * ld;
    const int h = PW >> 1;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t % ld);
        const long long u = t / ld;
        const int j = (int)(u % npairs), p = (int)(u / npairs);
        double2 A = make_double2(0.0, 0.0), B = A;
        if (k <= h) {
            const double2* zr = Z + ((size_t)p * npairs + j) * PW;
            const double2 zk = zr[k], zn = zr[(PW - k) & (PW - 1)];
            A = make_double2(0.5 * (zk.x + zn.x), 0.5 * (zk.y - zn.y));
            B = make_double2(0.5 * (zk.y + zn.y), 0.5 * (zn.x - zk.x));
        }
        const int d = k / cols, lc = k - d * cols;
        double2* o = dst.p[d] + (size_t)p * plane_stride + (size_t)(y0 + 2 * j - row_base) * cols + lc;
        o[0] = A;
        o[cols] = B;
    }
}

// tiles[plane][src][R][cols] (half rows of MY image rows, one tile per column owner) -> Z[plane][j][PW]:
//   Z[k] = A[k] + i B[k] (k <= PW/2),  conj(A[PW-k]) + i conj(B[PW-k]) (k > PW/2)
__global__ void __launch_bounds__(256) slab_merge_tiles(const double2* __restrict__ tiles, int PW, int cols, int G, int R, int npairs,
                                                        double2* __restrict__ Z) {
    const long long total = (long long)3 * npairs * PW;
    const int h = PW >> 1;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(t % PW);
        const long long u = t / PW;
        const int j = (int)(u % npairs), p = (int)(u / npairs);
        const int ks = k <= h ? k : PW - k;
        const int s = ks / cols, lc = ks - s * cols;
        const double2* a = tiles + (((size_t)p * G + s) * R + 2 * j) * cols + lc;
        const double2 A = a[0], B = a[cols];
        Z[t] = k <= h ? make_double2(A.x - B.y, A.y + B.x) : make_double2(A.x + B.y, B.x - A.y);
    }
}

__device__ __forceinline__ uint8_t slab_clamp8(double v) {  // from_planes_u8 S:389: round half away from zero, clamp
    double r = round(v);
    r = fmax(0.0, fmin(255.0, r));
    return (uint8_t)r;
}
// z[plane][j][PW] after the inverse row pass: Re -> image row 2j, Im -> row 2j + 1 (ifft_crop S:399, centre S:1102, S:387)
__global__ void __launch_bounds__(256) slab_pairs_to_u8(const double2* __restrict__ z, int nrows, int W, int PW, int npairs, int y0,
                                                        int center, uint8_t* __restrict__ rows) {
    const long long total = (long long)nrows * W * 3;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(t % 3);
        const long long px = t / 3;
        const int x = (int)(px % W), y = (int)(px / W);
        const double2 v = z[((size_t)p * npairs + (y >> 1)) * PW + x];
        double r = (y & 1) ? v.y : v.x;
        if (center && ((x + y0 + y) & 1)) r = -r;
        rows[t] = slab_clamp8(r);
    }
}

// write_bit_on_bin (S:712-732) on a column slab [3][PH][cols] holding columns [col0, col0 + cols) of the half spectrum:
// a bin (y, x) with x <= PW/2 is stored itself, one right of the Nyquist column through its mirror ((PH-y)%PH, PW-x) as
// the conjugate; columns 0 and PW/2 hold both.  Every stored element has exactly one owner rank.
__global__ void __launch_bounds__(256) slab_embed_scatter(double2* slab, int PH, int PW, int cols, int col0, const uint32_t* __restrict__ bins,
                                                          const uint8_t* __restrict__ bits, size_t nbits, double cos_a, double sin_a) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbits) return;
    const uint32_t b = bins[i];
    const int p = (int)(b >> 30);
    const uint32_t lin = b & 0x3FFFFFFFu;
    const int y = (int)(lin / (uint32_t)PW), x = (int)(lin % (uint32_t)PW);
    const int cy = (PH - y) & (PH - 1), cx = (PW - x) & (PW - 1);
    const int h = PW >> 1;
    const bool have_bin = x <= h, have_mir = cx <= h;
    const int xs = have_bin ? x : cx;  // (bin and mirror, when both are stored, share the column: x in {0, PW/2})
    if (xs < col0 || xs >= col0 + cols) return;
    double2* pl = slab + (size_t)p * PH * cols;
    const double2 z = have_bin ? pl[(size_t)y * cols + (x - col0)] : pl[(size_t)cy * cols + (cx - col0)];
    const double mag = fmax(1e-12, hypot(z.x, z.y));
    const double s = bits[i] ? sin_a : -sin_a;
    const double2 nv = make_double2(mag * cos_a, mag * s);
    if (cy == y && cx == x) {
        pl[(size_t)y * cols + (x - col0)] = make_double2(mag, 0.0);  // S:727
    } else {
        if (have_bin) pl[(size_t)y * cols + (x - col0)] = nv;
        if (have_mir) pl[(size_t)cy * cols + (cx - col0)] = make_double2(nv.x, -nv.y);
    }
}

// read_bit_from_bin (S:734-746, ties -> 1) for the bins this rank owns; -1 for the others (the ranks' lists are
// combined with a MAX reduction)
__global__ void __launch_bounds__(256) slab_read_raw(const double2* __restrict__ slab, int PH, int PW, int cols, int col0,
                                                     const uint32_t* __restrict__ bins, size_t nbins, double alpha, int8_t* __restrict__ raw) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbins) return;
    const uint32_t b = bins[i];
    const int p = (int)(b >> 30);
    const uint32_t lin = b & 0x3FFFFFFFu;
    int y = (int)(lin / (uint32_t)PW), x = (int)(lin % (uint32_t)PW);
    bool conj = false;
    if (x > (PW >> 1)) { y = (PH - y) & (PH - 1); x = PW - x; conj = true; }
    if (x < col0 || x >= col0 + cols) { raw[i] = -1; return; }
    double2 z = slab[((size_t)p * PH + y) * cols + (x - col0)];
    if (conj) z.y = -z.y;
    const double PI = 3.14159265358979323846;
    const double th = atan2(z.y, z.x);
    double dp = fmod(th - alpha + PI, 2 * PI);
    if (dp < 0) dp += 2 * PI;
    double dn = fmod(th + alpha + PI, 2 * PI);
    if (dn < 0) dn += 2 * PI;
    raw[i] = fabs(dp - PI) <= fabs(dn - PI) ? 1 : 0;
}

static inline unsigned slab_grid(const Launcher& L, long long total) {
    long long b = (total + 255) / 256, cap = (long long)L.sm_count * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

cudaError_t launch_slab_pack(const Launcher& L, const uint8_t* rows, int nrows, int W, int PW, int npairs, int y0, int center, double2* z) {
    slab_pack_pairs<<<slab_grid(L, (long long)3 * npairs * PW), 256, 0, L.stream>>>(rows, nrows, W, PW, npairs, y0, center, z);
    TFFT_SLAB_LAUNCH_CHECK(L);
    return cudaSuccess;
}
cudaError_t launch_slab_split(const Launcher& L, const double2* Z, int PW, int ld, int cols, int npairs, int y0, const SlabDst& dst,
                              size_t plane_stride, int row_base) {
    slab_split_scatter<<<slab_grid(L, (long long)3 * npairs * ld), 256, 0, L.stream>>>(Z, PW, ld, cols, npairs, y0, dst, plane_stride, row_base);
    TFFT_SLAB_LAUNCH_CHECK(L);
    return cudaSuccess;
}
cudaError_t launch_slab_merge(const Launcher& L, const double2* tiles, int PW, int cols, int G, int R, int npairs, double2* Z) {
    slab_merge_tiles<<<slab_grid(L, (long long)3 * npairs * PW), 256, 0, L.stream>>>(tiles, PW, cols, G, R, npairs, Z);
    TFFT_SLAB_LAUNCH_CHECK(L);
    return cudaSuccess;
}
cudaError_t launch_slab_to_u8(const Launcher& L, const double2* z, int nrows, int W, int PW, int npairs, int y0, int center, uint8_t* rows) {
    if (nrows <= 0) return cudaSuccess;
    slab_pairs_to_u8<<<slab_grid(L, (long long)nrows * W * 3), 256, 0, L.stream>>>(z, nrows, W, PW, npairs, y0, center, rows);
    TFFT_SLAB_LAUNCH_CHECK(L);
    return cudaSuccess;
}
cudaError_t launch_slab_embed(const Launcher& L, double2* slab, int PH, int PW, int cols, int col0, const uint32_t* bins, const uint8_t* bits,
                              size_t nbits, double cos_a, double sin_a) {
    if (!nbits) return cudaSuccess;
    slab_embed_scatter<<<(unsigned)((nbits + 255) / 256), 256, 0, L.stream>>>(slab, PH, PW, cols, col0, bins, bits, nbits, cos_a, sin_a);
    TFFT_SLAB_LAUNCH_CHECK(L);
    return cudaSuccess;
}
cudaError_t launch_slab_read(const Launcher& L, const double2* slab, int PH, int PW, int cols, int col0, const uint32_t* bins, size_t nbins,
                             double alpha, int8_t* raw) {
    if (!nbins) return cudaSuccess;
    slab_read_raw<<<(unsigned)((nbins + 255) / 256), 256, 0, L.stream>>>(slab, PH, PW, cols, col0, bins, nbins, alpha, raw);
    TFFT_SLAB_LAUNCH_CHECK(L);
    return cudaSuccess;
}

}  // namespace tfft
