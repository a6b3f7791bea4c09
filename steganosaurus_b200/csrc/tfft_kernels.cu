// tfft_kernels.cu -- sm_100a kernels of the TurtleFFT hot path: baseline FFT pass (v0),
// median/capacity, embed scatter, extract gather+vote.
// Reference citations S:n = steganosaurus/src/steganosaur.cpp line n.
#include "tfft_kernels.cuh"

#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <mutex>
#include <vector>

namespace tfft {

#define TFFT_LAUNCH_CHECK(L)              \
    do {                                  \
        if ((L).launch_counter) ++*(L).launch_counter; \
        cudaError_t e__ = cudaGetLastError(); \
        if (e__ != cudaSuccess) return e__;   \
    } while (0)

// --------------------------------------------------------------------------------------------
// Twiddles
// --------------------------------------------------------------------------------------------
cudaError_t build_twiddles(double2* d_tw, cudaStream_t s) {
    std::vector<double2> h(TW_N / 2);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int k = 0; k < TW_N / 2; k++) {
        // exact octant symmetries keep cos/sin of k/TW_N consistent to the last bit
        long double a = two_pi * (long double)k / (long double)TW_N;
        h[k].x = (double)cosl(a);
        h[k].y = (double)sinl(a);
    }
    h[0] = make_double2(1.0, 0.0);
    h[TW_N / 4] = make_double2(0.0, 1.0);
    cudaError_t e = cudaMemcpyAsync(d_tw, h.data(), sizeof(double2) * h.size(), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(s);
}

// --------------------------------------------------------------------------------------------
// Shared device helpers
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// from_planes_u8 clamp8 (S:389): round() half away from zero, clamp to [0,255], cast.
__device__ __forceinline__ uint8_t clamp8(double v) {
    double r = round(v);
    r = fmax(0.0, fmin(255.0, r));
    return (uint8_t)r;
}

// --------------------------------------------------------------------------------------------
// v0 FFT pass: one CTA transforms `cpc` pencils held entirely in shared memory with an
// in-place radix-2 DIT (same butterfly order as fft1d S:341-358, but table twiddles).
// Kept as the simple always-correct path and as the cross-check for the pencil kernels.
// --------------------------------------------------------------------------------------------
template <int IN, int OUT>
__global__ void __launch_bounds__(512) fft_pass_v0(PassArgs a, int cpc) {
    extern __shared__ double2 sm[];
    const int n = 1 << a.log2n;
    const int tid = threadIdx.x, T = blockDim.x;
    const long long q0 = (long long)blockIdx.x * cpc;  // first pencil of this CTA
    const int len_other = a.axis == 0 ? a.PH : a.PW;   // pencils per plane
    const int ip = (int)(q0 / len_other);              // plane index (image*3 + plane)
    const int o0 = (int)(q0 % len_other);              // first row (axis 0) / column (axis 1)
    double2* plane = a.spec + (size_t)ip * a.PH * a.PW;
    const int img = ip / 3, ch = ip % 3;

    if (a.axis == 0) {
        // rows beyond out_rows are never consumed; rows beyond in_rows are all-zero on input
        if (o0 >= a.out_rows) return;
        if (o0 >= a.in_rows) {
            if (OUT == OUT_C64)
                for (int idx = tid; idx < cpc * n; idx += T) plane[(size_t)(o0 + idx / n) * a.PW + (idx % n)] = make_double2(0.0, 0.0);
            return;
        }
    }

    // ---- load, bit-reversed placement (S:343-345)
    for (int idx = tid; idx < cpc * n; idx += T) {
        int c, k;
        if (a.axis == 0) { c = idx / n; k = idx % n; } else { c = idx % cpc; k = idx / cpc; }
        double2 v;
        if (IN == IN_U8) {
            const int y = o0 + c;
            double r = 0.0;
            if (y < a.H && k < a.W) {
                r = (double)a.img_in[((size_t)((size_t)img * a.H + y) * a.W + k) * 3 + ch];
                if (a.center && ((k + y) & 1)) r = -r;  // apply_center S:392
            }
            v = make_double2(r, 0.0);
        } else if (a.axis == 0) {
            const int y = o0 + c;
            v = (y < a.in_rows) ? plane[(size_t)y * a.PW + k] : make_double2(0.0, 0.0);
        } else {
            v = (k < a.in_rows) ? plane[(size_t)k * a.PW + o0 + c] : make_double2(0.0, 0.0);
        }
        sm[c * n + (int)(__brev((unsigned)k) >> (32 - a.log2n))] = v;
    }
    __syncthreads();

    // ---- butterflies
    const int half_n = n >> 1;
    for (int s = 1; s <= a.log2n; s++) {
        const int half = 1 << (s - 1);
        for (int b = tid; b < cpc * half_n; b += T) {
            const int c = b >> (a.log2n - 1);
            const int j = b & (half_n - 1);
            const int jj = j & (half - 1);
            const int pos = c * n + (((j >> (s - 1)) << s) | jj);
            double2 w = a.tw[(size_t)jj << (TW_LOG2 - s)];
            if (a.inverse) w.y = -w.y;
            const double2 u = sm[pos];
            const double2 v = cmul(sm[pos + half], w);
            sm[pos] = make_double2(u.x + v.x, u.y + v.y);
            sm[pos + half] = make_double2(u.x - v.x, u.y - v.y);
        }
        __syncthreads();
    }

    // ---- store
    const double scale = a.inverse ? 1.0 / (double)n : 1.0;  // S:357 (exact: n is a power of two)
    for (int idx = tid; idx < cpc * n; idx += T) {
        int c, k;
        if (a.axis == 0) { c = idx / n; k = idx % n; } else { c = idx % cpc; k = idx / cpc; }
        double2 v = sm[c * n + k];
        v.x *= scale; v.y *= scale;
        if (OUT == OUT_U8) {
            const int y = o0 + c;
            if (y < a.H && k < a.W) {
                double r = v.x;                               // ifft_crop S:401 keeps the real part
                if (a.center && ((k + y) & 1)) r = -r;        // S:1102
                a.img_out[((size_t)((size_t)img * a.H + y) * a.W + k) * 3 + ch] = clamp8(r);
            }
        } else if (a.axis == 0) {
            plane[(size_t)(o0 + c) * a.PW + k] = v;
        } else if (k < a.out_rows) {
            plane[(size_t)k * a.PW + o0 + c] = v;
        }
    }
}

cudaError_t launch_fft_pass_pencil(const Launcher& L, const PassArgs& a, bool* handled);

static cudaError_t launch_v0(const Launcher& L, const PassArgs& a) {
    const int n = 1 << a.log2n;
    const size_t per = (size_t)n * sizeof(double2);
    const int other = a.axis == 0 ? a.PH : a.PW;
    int cpc = 1;
    // as many pencils per CTA as fit (columns: wider contiguous chunks; rows: more work per CTA)
    const size_t budget = L.smem_optin < 200 * 1024 ? L.smem_optin : 200 * 1024;
    while (cpc < 8 && (size_t)(cpc * 2) * per <= budget && other % (cpc * 2) == 0) cpc *= 2;
    if (per * cpc > L.smem_optin) return cudaErrorInvalidConfiguration;
    const size_t smem = per * cpc;
    const long long npencils = (long long)a.nplanes * other;
    const unsigned grid = (unsigned)(npencils / cpc);
    void (*k)(PassArgs, int) = nullptr;
    const int in = a.img_in ? IN_U8 : IN_C64, out = a.img_out ? OUT_U8 : OUT_C64;
    if (in == IN_C64 && out == OUT_C64) k = fft_pass_v0<IN_C64, OUT_C64>;
    else if (in == IN_U8 && out == OUT_C64) k = fft_pass_v0<IN_U8, OUT_C64>;
    else if (in == IN_C64 && out == OUT_U8) k = fft_pass_v0<IN_C64, OUT_U8>;
    else return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k<<<grid, 512, smem, L.stream>>>(a, cpc);
    TFFT_LAUNCH_CHECK(L);
    return cudaSuccess;
}

// --------------------------------------------------------------------------------------------
// Four-step scheme for pencils longer than 4096 points (8192 = 2*4096, 16384 = 4*4096).
//   x[R*m + r]  --M-point FFT over m for every r-->  Y_r[k]          (in place: a length-N pencil
//                 viewed as an [M][R] matrix makes the R strided sub-sequences its columns, so the
//                 sub-transforms are exactly a column pass of the 4096-point pencil kernel)
//   X[k + M*q] = sum_r w_R^{rq} * w_N^{rk} * Y_r[k]                   (radix-R combine, spec -> tmp)
// and the result is copied back so the pass stays in place for its callers.
// --------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256) fourstep_combine(const double2* __restrict__ src, double2* __restrict__ dst,
                                                        const double2* __restrict__ tw, int axis, int PH, int PW,
                                                        int log2n, int inverse, long long total) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int N = 1 << log2n, M = N / R;
    size_t in0, in_stride, out0, out_stride;
    int k;
    if (axis == 0) {  // pencil = one row of PW = N points; consecutive threads -> consecutive k
        k = (int)(t % M);
        const size_t row = (size_t)(t / M);  // plane*PH + y
        in0 = row * N + (size_t)R * k; in_stride = 1;
        out0 = row * N + k; out_stride = M;
    } else {          // pencil = one column of PH = N points; consecutive threads -> consecutive x
        const int x = (int)(t % PW);
        const long long u = t / PW;
        k = (int)(u % M);
        const size_t plane = (size_t)(u / M);
        in0 = plane * (size_t)N * PW + (size_t)R * k * PW + x; in_stride = PW;
        out0 = plane * (size_t)N * PW + (size_t)k * PW + x; out_stride = (size_t)M * PW;
    }
    double2 y[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        double2 v = src[in0 + (size_t)r * in_stride];
        if (r) {  // w_N^{r k}: table holds exp(+2 pi i j / 16384), j < 8192; the other half by w^{j+8192} = -w^j
            unsigned j = (unsigned)(r * k) << (TW_LOG2 - log2n);
            double2 w = tw[j & (TW_N / 2 - 1)];
            if (j & (TW_N / 2)) { w.x = -w.x; w.y = -w.y; }
            if (inverse) w.y = -w.y;
            v = make_double2(v.x * w.x - v.y * w.y, v.x * w.y + v.y * w.x);
        }
        y[r] = v;
    }
    const double sc = inverse ? 1.0 / R : 1.0;  // the sub-transforms already scaled by 1/M (S:357)
    if (R == 2) {
        dst[out0] = make_double2((y[0].x + y[1].x) * sc, (y[0].y + y[1].y) * sc);
        dst[out0 + out_stride] = make_double2((y[0].x - y[1].x) * sc, (y[0].y - y[1].y) * sc);
    } else {
        const double s = inverse ? -1.0 : 1.0;  // w_4 = s*i
        const double2 t0 = make_double2(y[0].x + y[2].x, y[0].y + y[2].y), t1 = make_double2(y[0].x - y[2].x, y[0].y - y[2].y);
        const double2 t2 = make_double2(y[1].x + y[3].x, y[1].y + y[3].y);
        const double2 d = make_double2(y[1].x - y[3].x, y[1].y - y[3].y);
        const double2 t3 = make_double2(-s * d.y, s * d.x);  // s*i*(y1 - y3)
        dst[out0] = make_double2((t0.x + t2.x) * sc, (t0.y + t2.y) * sc);
        dst[out0 + out_stride] = make_double2((t1.x + t3.x) * sc, (t1.y + t3.y) * sc);
        dst[out0 + 2 * out_stride] = make_double2((t0.x - t2.x) * sc, (t0.y - t2.y) * sc);
        dst[out0 + 3 * out_stride] = make_double2((t1.x - t3.x) * sc, (t1.y - t3.y) * sc);
    }
}

static cudaError_t launch_fourstep(const Launcher& L, const PassArgs& a) {
    const int R = 1 << (a.log2n - 12), M = 4096;
    // sub-transforms: column pass of the reshaped batch
    PassArgs s = a;
    s.axis = 1; s.log2n = 12; s.half = 0; s.img_in = nullptr; s.img_out = nullptr;
    if (a.axis == 0) { s.nplanes = a.nplanes * a.PH; s.PH = M; s.PW = R; }          // every row is an [M][R] matrix
    else             { s.PH = M; s.PW = R * a.PW; }                               // plane [N][PW] seen as [M][R*PW]
    s.ld = s.PW; s.in_rows = M; s.out_rows = M; s.W = s.PW; s.H = s.PH;
    s.col_limit = 0; s.fourstep_sub_only = 0;
    bool handled = false;
    cudaError_t e = cudaSuccess;
    if (a.axis == 1 && a.fourstep_sub_only && a.col_limit > 0 && a.col_limit < a.PW) {
        // only the first col_limit columns are read downstream: in the [M][R * PW] view they are R separate column ranges
        for (int r = 0; r < R && e == cudaSuccess; r++) {
            PassArgs sr = s;
            sr.spec = a.spec + (size_t)r * a.PW;
            sr.col_limit = a.col_limit;
            e = launch_fft_pass_pencil(L, sr, &handled);
            if (e == cudaSuccess && !handled) e = cudaErrorNotSupported;
        }
        return e;
    }
    e = launch_fft_pass_pencil(L, s, &handled);
    if (e != cudaSuccess) return e;
    if (!handled) return cudaErrorNotSupported;
    if (a.axis == 1 && a.fourstep_sub_only) return cudaSuccess;
    const long long total = (long long)a.nplanes * a.PH * a.PW / R;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (R == 2) fourstep_combine<2><<<grid, 256, 0, L.stream>>>(a.spec, a.tmp, a.tw, a.axis, a.PH, a.PW, a.log2n, a.inverse, total);
    else        fourstep_combine<4><<<grid, 256, 0, L.stream>>>(a.spec, a.tmp, a.tw, a.axis, a.PH, a.PW, a.log2n, a.inverse, total);
    TFFT_LAUNCH_CHECK(L);
    if (a.leave_in_tmp) return cudaSuccess;
    return cudaMemcpyAsync(a.spec, a.tmp, (size_t)a.nplanes * a.PH * a.PW * sizeof(double2), cudaMemcpyDeviceToDevice, L.stream);
}

cudaError_t launch_fft_pass(const Launcher& L, const PassArgs& a) {
    if (L.fft_impl != 0) {
        if (a.log2n > 12 && a.log2n <= 14 && (a.tmp || (a.fourstep_sub_only && a.axis == 1)) && !a.img_in && !a.img_out && !a.half)
            return launch_fourstep(L, a);
        bool handled = false;
        cudaError_t e = launch_fft_pass_pencil(L, a, &handled);
        if (e == cudaErrorNotSupported && !a.img_in && !a.img_out && !a.signmap && !a.fused_embed && !a.half) {
            cudaGetLastError();  // a plain c2c pass whose column count no pencil kernel tiles: the generic kernel takes it
        } else if (e != cudaSuccess || handled) {
            return e;
        }
    }
    return launch_v0(L, a);
}

// --------------------------------------------------------------------------------------------
// Median of |F| (median_abs S:404-409): exact selection of the element of rank P/2 by an MSB
// radix select on the IEEE bit pattern of hypot(re,im) (non-negative doubles order like
// unsigned integers).  Generic path: six 11-bit histogram passes over the plane (always exact,
// used as the on-device fallback).  Fast path: see "sampled bracket" below.
// --------------------------------------------------------------------------------------------
constexpr int RADIX_BITS = 11;
constexpr int RADIX = 1 << RADIX_BITS;
// digit d (0 = most significant) covers key bits [63-11d-1 .. 63-11d-11] of the 63 value bits
__device__ __forceinline__ int key_shift(int d) { int s = 63 - RADIX_BITS * (d + 1); return s < 0 ? 0 : s; }
__device__ __forceinline__ int key_width(int d) { int s = 63 - RADIX_BITS * (d + 1); return s < 0 ? RADIX_BITS + s : RADIX_BITS; }
constexpr int NUM_DIGITS = 6;  // 5*11 + 8 = 63

__device__ __forceinline__ uint64_t mag_key(double2 z) {
    return (uint64_t)__double_as_longlong(hypot(z.x, z.y));  // std::abs(complex) S:406
}

// multiplicity of stored column x in the full PH x PW multiset (half layout: Hermitian mirror)
__device__ __forceinline__ int col_weight(const SpecLayout& s, int x) {
    if (!s.half) return 1;
    const int h = s.PW >> 1;
    return (x == 0 || x == h) ? 1 : (x < h ? 2 : 0);
}
// element (y,x) of the FULL spectrum
// stored element (y, x) of a plane whose column pass stopped after the four-step sub-transforms (SpecLayout::fs_r)
__device__ __forceinline__ double2 spec_load_fs(const double2* __restrict__ pl, const SpecLayout& s, int y, int x) {
    const int M = s.PH / s.fs_r, k = y & (M - 1);
    int lg = 0;
    while ((1 << lg) < s.PH) lg++;
    const double2* p0 = pl + (size_t)s.fs_r * k * s.ld + x;
    double2 acc = p0[0];
    for (int r = 1; r < s.fs_r; r++) {
        const unsigned j = (unsigned)(r * y) << (TW_LOG2 - lg);  // w_PH^{r y} from the table of exp(+2 pi i j / 16384), j < 8192
        double2 w = s.fs_tw[j & (TW_N / 2 - 1)];
        if (j & (TW_N / 2)) { w.x = -w.x; w.y = -w.y; }
        const double2 v = p0[(size_t)r * s.ld];
        acc.x += v.x * w.x - v.y * w.y;
        acc.y += v.x * w.y + v.y * w.x;
    }
    return acc;
}
__device__ __forceinline__ double2 spec_load(const double2* __restrict__ pl, const SpecLayout& s, int y, int x) {
    if (!s.half || x <= (s.PW >> 1)) return s.fs_r ? spec_load_fs(pl, s, y, x) : pl[(size_t)y * s.ld + x];
    const int cy = (s.PH - y) & (s.PH - 1), cx = s.PW - x;
    double2 z = s.fs_r ? spec_load_fs(pl, s, cy, cx) : pl[(size_t)cy * s.ld + cx];
    z.y = -z.y;
    return z;
}

size_t median_work_bytes(int nplanes, uint32_t cand_cap) {
    size_t b = 0;
    b += (size_t)nplanes * RADIX * sizeof(uint32_t);
    b += (size_t)nplanes * sizeof(uint64_t) * 6;  // prefix, rank, counts, bracket (2), cap_below
    b += (size_t)nplanes * cand_cap * sizeof(uint64_t);
    b += (size_t)nplanes * CAP_UNC_MAX * sizeof(uint64_t);
    b += (size_t)nplanes * CAND_B_MAX * sizeof(uint64_t);
    b += sizeof(uint64_t);                          // ann_total
    b += (size_t)(3 * nplanes + 4) * sizeof(uint32_t);  // cand_n, cap_unc_n, cand_b_n, flags
    return (b + 255) & ~(size_t)255;
}
void median_work_carve(MedianWork& w, void* base, int nplanes, uint32_t cand_cap) {
    char* p = (char*)base;
    w.prefix = (uint64_t*)p; p += (size_t)nplanes * sizeof(uint64_t);
    w.rank = (uint64_t*)p; p += (size_t)nplanes * sizeof(uint64_t);
    w.counts = (uint64_t*)p; p += (size_t)nplanes * sizeof(uint64_t);
    w.prefix2 = (uint64_t*)p; p += (size_t)nplanes * 2 * sizeof(uint64_t);
    w.cap_below = (uint64_t*)p; p += (size_t)nplanes * sizeof(uint64_t);
    w.ann_total = (uint64_t*)p; p += sizeof(uint64_t);
    w.cand = (uint64_t*)p; p += (size_t)nplanes * cand_cap * sizeof(uint64_t);
    w.cap_unc = (uint64_t*)p; p += (size_t)nplanes * CAP_UNC_MAX * sizeof(uint64_t);
    w.cand_b = (uint64_t*)p; p += (size_t)nplanes * CAND_B_MAX * sizeof(uint64_t);
    w.hist = (uint32_t*)p; p += (size_t)nplanes * RADIX * sizeof(uint32_t);
    w.cand_n = (uint32_t*)p; p += (size_t)nplanes * sizeof(uint32_t);
    w.cap_unc_n = (uint32_t*)p; p += (size_t)nplanes * sizeof(uint32_t);
    w.cand_b_n = (uint32_t*)p; p += (size_t)nplanes * sizeof(uint32_t);
    w.flags = (int*)p;
    w.cand_cap = cand_cap;
}

__global__ void median_init(MedianWork w, int nplanes, uint64_t P) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nplanes * RADIX) w.hist[i] = 0;
    if (i < nplanes) { w.prefix[i] = 0; w.rank[i] = P / 2; w.cand_n[i] = 0; w.counts[i] = 0; w.cap_below[i] = 0; w.cap_unc_n[i] = 0; w.cand_b_n[i] = 0; }
    if (i < 2) w.flags[i] = 0;
}

// histogram of digit d over keys whose higher digits equal prefix; grid = (chunks, nplanes)
__global__ void __launch_bounds__(512) median_hist(const double2* __restrict__ spec, SpecLayout lay, int d, MedianWork w, const int* gate) {
    __shared__ uint32_t sh[RADIX];
    if (gate && !*gate) return;  // degenerate-spectrum fallback not needed
    const int ip = blockIdx.y;
    for (int i = threadIdx.x; i < RADIX; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint64_t E = lay.plane_elems();
    const double2* pl = spec + (size_t)ip * E;
    const int sft = key_shift(d), wid = key_width(d);
    const uint64_t prefix = w.prefix[ip];
    const int hi_sft = sft + wid;  // bits above this digit
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (uint64_t)gridDim.x * blockDim.x) {
        const int wt = col_weight(lay, (int)(i % (uint64_t)lay.ld));
        if (!wt) continue;
        const uint64_t k = mag_key(pl[i]);
        if (d == 0 || (k >> hi_sft) == (prefix >> hi_sft))
            atomicAdd(&sh[(unsigned)(k >> sft) & ((1u << wid) - 1)], (unsigned)wt);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RADIX; i += blockDim.x)
        if (sh[i]) atomicAdd(&w.hist[ip * RADIX + i], sh[i]);
}

// one CTA per plane: find the bucket holding the wanted rank, extend prefix, clear histogram
__global__ void median_pick(int d, MedianWork w, const int* gate) {
    if (gate && !*gate) return;
    const int ip = blockIdx.x;
    if (threadIdx.x == 0) {
        uint64_t r = w.rank[ip];
        uint32_t* h = w.hist + ip * RADIX;
        const int wid = key_width(d);
        int b = 0;
        for (; b < (1 << wid) - 1; b++) {
            if (r < h[b]) break;
            r -= h[b];
        }
        w.rank[ip] = r;
        w.prefix[ip] |= (uint64_t)b << key_shift(d);
        if (!gate) w.cand_n[ip] = h[b];  // survivors after this digit
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RADIX; i += blockDim.x) w.hist[ip * RADIX + i] = 0;
}

// generic path completion: the prefix is complete after all digits
__global__ void median_from_prefix(MedianWork w, double* median, int nplanes, const int* gate) {
    if (!*gate) return;
    int ip = blockIdx.x * blockDim.x + threadIdx.x;
    if (ip < nplanes) median[ip] = __longlong_as_double((long long)w.prefix[ip]);  // exact for every plane
}
// ---- fast path: sampled bracket + one exact scan ---------------------------------------------
// (1) a strided sample of |F| gives two order statistics lo <= hi that bracket the median with
// overwhelming probability; (2) ONE pass over the plane counts the elements strictly below the
// bracket and compacts the few (~1 %) inside it -- membership is decided on q = re^2+im^2 (two
// FP64 instructions) against thresholds widened by 1e-9, hypot() is only evaluated for members;
// (3) one CTA per plane selects the exact element of rank P/2 - below among the members.  The
// result is the same double median_abs (S:404-409) returns.  Whenever the bracket misses or the
// members do not fit (flat images, adversarial spectra) a device-side flag routes the plane
// batch through the generic radix passes above -- the answer is exact either way.
constexpr uint32_t SAMPLE_MAX = 1u << 18;
constexpr uint32_t RANK_GUARD = 8;

struct Bracket { double qlo, qhi; };

// block-wide exact selection of the key of rank `rank` (0-based) among n keys (blockDim.x == 1024).
// The bits shared by every key (clz(min ^ max)) are skipped, the rest is consumed 11 bits at a time
// with a shared-memory histogram and a parallel bucket search.
struct SelectScratch {
    uint32_t hist[RADIX];
    uint32_t warp_tot[32];
    uint64_t red[64];
    uint64_t prefix, rank;
    int top;
};
// The multiset is { every key of c[0..n) with multiplicity wa } minus { every key of cb[0..nb) once } (cb is a
// sub-multiset: keys whose true multiplicity is below wa).  rank is 0-based in that multiset.
__device__ uint64_t select_rank(const uint64_t* __restrict__ c, uint32_t n, uint32_t wa, const uint64_t* __restrict__ cb, uint32_t nb,
                                uint64_t rank, SelectScratch* sc) {
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    // min / max
    uint64_t mn = ~0ull, mx = 0;
    for (uint32_t i0 = tid; i0 < n; i0 += blockDim.x * 8) {
        uint64_t k[8];
#pragma unroll
        for (int u = 0; u < 8; u++) k[u] = (i0 + u * blockDim.x < n) ? c[i0 + u * blockDim.x] : c[tid < n ? tid : 0];
#pragma unroll
        for (int u = 0; u < 8; u++) { mn = k[u] < mn ? k[u] : mn; mx = k[u] > mx ? k[u] : mx; }
    }
    for (int o = 16; o; o >>= 1) {
        const uint64_t a = __shfl_xor_sync(0xffffffffu, mn, o), b2 = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn; mx = b2 > mx ? b2 : mx;
    }
    if (lane == 0) { sc->red[wrp] = mn; sc->red[32 + wrp] = mx; }
    __syncthreads();
    if (tid == 0) {
        for (int w2 = 1; w2 < (int)(blockDim.x >> 5); w2++) { mn = sc->red[w2] < mn ? sc->red[w2] : mn; mx = sc->red[32 + w2] > mx ? sc->red[32 + w2] : mx; }
        const uint64_t diff = mn ^ mx;
        sc->top = diff ? 64 - __clzll((long long)diff) : 0;  // number of low bits that still vary
        sc->prefix = sc->top == 64 ? 0 : (mn >> sc->top) << sc->top;
        sc->rank = rank;
    }
    __syncthreads();
    while (true) {
        const int top = sc->top;
        if (top == 0) break;
        const int wid = top < RADIX_BITS ? top : RADIX_BITS, sft = top - wid;
        const uint64_t prefix = sc->prefix;
        for (int i = tid; i < RADIX; i += blockDim.x) sc->hist[i] = 0;
        __syncthreads();
        for (uint32_t i0 = tid; i0 < n; i0 += blockDim.x * 8) {
            uint64_t k[8];
#pragma unroll
            for (int u = 0; u < 8; u++) k[u] = (i0 + u * blockDim.x < n) ? c[i0 + u * blockDim.x] : 0;
#pragma unroll
            for (int u = 0; u < 8; u++)
                if (i0 + u * blockDim.x < n && (top == 64 || (k[u] >> top) == (prefix >> top)))
                    atomicAdd(&sc->hist[(unsigned)(k[u] >> sft) & ((1u << wid) - 1)], wa);
        }
        __syncthreads();  // every +wa lands before the -1 of the same key: bins never go negative
        for (uint32_t i = tid; i < nb; i += blockDim.x) {
            const uint64_t k = cb[i];
            if (top == 64 || (k >> top) == (prefix >> top)) atomicSub(&sc->hist[(unsigned)(k >> sft) & ((1u << wid) - 1)], 1u);
        }
        __syncthreads();
        // parallel bucket search: thread t owns bins 2t, 2t+1
        const uint32_t h0 = sc->hist[2 * tid], h1 = sc->hist[2 * tid + 1];
        uint32_t incl = h0 + h1;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) sc->warp_tot[wrp] = incl;
        __syncthreads();
        if (wrp == 0) {
            uint32_t t = sc->warp_tot[lane];
            for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += v; }
            sc->warp_tot[lane] = t;  // inclusive over warps
        }
        __syncthreads();
        const uint64_t r = sc->rank;
        const uint64_t excl = (uint64_t)(wrp ? sc->warp_tot[wrp - 1] : 0) + (incl - h0 - h1);
        __syncthreads();
        if (r >= excl && r < excl + h0 + h1) {  // exactly one thread
            const int bkt = (r < excl + h0) ? 2 * tid : 2 * tid + 1;
            sc->rank = r - (bkt == 2 * tid ? excl : excl + h0);
            sc->prefix = prefix | ((uint64_t)bkt << sft);
            sc->top = sft;
        }
        __syncthreads();
    }
    const uint64_t res = sc->prefix;
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(256) median_sample(const double2* __restrict__ spec, SpecLayout lay, uint32_t S, uint64_t stride, MedianWork w) {
    const int ip = blockIdx.y;
    const double2* pl = spec + (size_t)ip * lay.plane_elems();
    const int ncl = lay.half ? lay.PW >> 1 : lay.PW;  // logical columns sampled (weight-2 region of a half plane)
    // stratified pseudo-random positions: one element per stride block at a hashed offset (a fixed
    // offset would alias with the periodic leakage pattern that zero padding imprints on the spectrum)
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < S; j += gridDim.x * blockDim.x) {
        uint32_t h = (j + 0x9E3779B9u * (uint32_t)(ip + 1)) * 2654435761u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        const uint64_t li = (uint64_t)j * stride + (uint64_t)(h % (uint32_t)stride);
        const uint64_t y = li / (uint64_t)ncl, x = li % (uint64_t)ncl;
        w.cand[(size_t)ip * w.cand_cap + j] = mag_key(pl[y * (uint64_t)lay.ld + x]);
    }
}

// small planes (P <= cand_cap): no sampling, every element is a "member"
__global__ void median_bracket_all(MedianWork w, int nplanes, Bracket* br) {
    const int ip = blockIdx.x * blockDim.x + threadIdx.x;
    if (ip < nplanes) { br[ip].qlo = 0.0; br[ip].qhi = __longlong_as_double(0x7ff0000000000000LL); w.cand_n[ip] = 0; }
}

// one CTA per plane: sample quantiles 0.5 +- 6 sigma (sigma = 0.5/sqrt(S)) -> q-space bracket.
// The bracket only has to CONTAIN the median, so the two order statistics are resolved to the
// top 22 key bits (exponent + 11 mantissa bits, 0.05 % in magnitude) with two histogram passes
// over the sample and then rounded outwards.
__global__ void __launch_bounds__(1024) median_bracket(MedianWork w, uint32_t S, Bracket* br, int qkeys) {
    __shared__ uint32_t h1[RADIX], h2lo[RADIX], h2hi[RADIX];
    __shared__ uint32_t s_blo, s_bhi, s_rlo, s_rhi, s_slo, s_shi;
    const int ip = blockIdx.x;
    const uint64_t* c = w.cand + (size_t)ip * w.cand_cap;
    const double delta = 3.0 / sqrt((double)S);
    long long rlo = (long long)floor((0.5 - delta) * S), rhi = (long long)ceil((0.5 + delta) * S);
    if (rlo < 0) rlo = 0;
    if (rhi > (long long)S - 1) rhi = S - 1;
    for (int i = threadIdx.x; i < RADIX; i += blockDim.x) { h1[i] = 0; h2lo[i] = 0; h2hi[i] = 0; }
    __syncthreads();
    // level 1 = the 11 exponent bits: a spectrum lives in a handful of binades, so the increments of a warp are
    // aggregated per distinct bin (match_any) instead of serialising on the same few counters; loads are
    // batched 8 deep (the sample comes from L2)
    for (uint32_t i0 = threadIdx.x; i0 < S; i0 += blockDim.x * 8) {
        uint64_t k[8];
#pragma unroll
        for (int u = 0; u < 8; u++) k[u] = (i0 + u * blockDim.x < S) ? c[i0 + u * blockDim.x] : ~0ull;
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const bool live = i0 + u * blockDim.x < S;
            const unsigned bin = live ? (unsigned)(k[u] >> 52) & (RADIX - 1) : 0xffffffffu;
            const unsigned peers = __match_any_sync(0xffffffffu, bin);
            if (live && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&h1[bin], (unsigned)__popc(peers));
        }
    }
    __syncthreads();
    if (threadIdx.x < 2) {  // thread 0: lower rank, thread 1: upper rank
        uint32_t r = threadIdx.x == 0 ? (uint32_t)rlo : (uint32_t)rhi;
        int b = 0;
        for (; b < RADIX - 1; b++) { if (r < h1[b]) break; r -= h1[b]; }
        if (threadIdx.x == 0) { s_blo = b; s_rlo = r; } else { s_bhi = b; s_rhi = r; }
    }
    __syncthreads();
    const uint32_t blo = s_blo, bhi = s_bhi;
    for (uint32_t i0 = threadIdx.x; i0 < S; i0 += blockDim.x * 8) {
        uint64_t k[8];
#pragma unroll
        for (int u = 0; u < 8; u++) k[u] = (i0 + u * blockDim.x < S) ? c[i0 + u * blockDim.x] : ~0ull;
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (i0 + u * blockDim.x >= S) continue;
            const uint32_t top = (unsigned)(k[u] >> 52) & (RADIX - 1), sub = (unsigned)(k[u] >> 41) & (RADIX - 1);
            if (top == blo) atomicAdd(&h2lo[sub], 1u);
            if (top == bhi) atomicAdd(&h2hi[sub], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        const uint32_t* h = threadIdx.x == 0 ? h2lo : h2hi;
        uint32_t r = threadIdx.x == 0 ? s_rlo : s_rhi;
        int b = 0;
        for (; b < RADIX - 1; b++) { if (r < h[b]) break; r -= h[b]; }
        if (threadIdx.x == 0) s_slo = b; else s_shi = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint64_t klo = ((uint64_t)blo << 52) | ((uint64_t)s_slo << 41);                              // lower edge
        const uint64_t khi = ((((uint64_t)bhi << 11) | (uint64_t)s_shi) + 1 << 41) - 1;                      // upper edge
        const double lo = __longlong_as_double((long long)klo), hi = __longlong_as_double((long long)khi);
        // keys are magnitudes (gathered sample) or already q = |F|^2 (sample dropped by the column pass)
        br[ip].qlo = (qkeys ? lo : lo * lo) * (1.0 - 1e-9);
        br[ip].qhi = (khi >= 0x7ff0000000000000ull) ? __longlong_as_double(0x7ff0000000000000LL) : (qkeys ? hi : hi * hi) * (1.0 + 1e-9);
        w.cand_n[ip] = 0;  // becomes the member fill counter
    }
}

constexpr int SCAN_UNROLL = 8;
constexpr int SCAN_THREADS = 512;
constexpr uint32_t SCAN_TILE = SCAN_THREADS * SCAN_UNROLL;
constexpr uint32_t SCAN_SBUF = 3072;  // members staged per CTA before one global reservation (24 KB)

// Fused capacity (count_plane S:999-1007): an annulus bin fails |F| >= magmin*median only when its q is
// ~magmin^2 times the median's, i.e. almost never.  The scan counts the bins that certainly fail
// (q < qcap_lo) and stages the keys of the undecided ones (qcap_lo <= q <= qcap_hi, the image of the median
// bracket); capacity_resolve settles those against the exact median.  Off (on = 0) when the annulus
// reaches columns the workspace does not store or every element is a member.
struct ScanCap {
    int on;
    double magmin2;   // magmin^2
    double rlo, rhi;  // annulus radii in bins
};

// A warp with at least one bracket member: ballot-compact the keys into the CTA's staging buffer.  Every stored
// element is staged ONCE here; multiplicities are settled by the weighted select (see median_scan).
__device__ __forceinline__ void scan_stage(bool member, double2 z, MedianWork& w, int ip, uint64_t* s_buf, unsigned* s_cnt) {
    const int lane = threadIdx.x & 31;
    const unsigned m = __ballot_sync(0xffffffffu, member);
    unsigned b0 = 0;
    if (lane == 0) b0 = atomicAdd(s_cnt, (unsigned)__popc(m));
    b0 = __shfl_sync(0xffffffffu, b0, 0);
    if (member) {
        const unsigned slot = b0 + __popc(m & ((1u << lane) - 1));
        const uint64_t key = mag_key(z);
        if (slot < SCAN_SBUF) s_buf[slot] = key;
        else {  // staging full (pathological density): reserve directly
            const unsigned g = atomicAdd(&w.cand_n[ip], 1u);
            if (g < w.cand_cap) w.cand[(size_t)ip * w.cand_cap + g] = key;
        }
    }
}

// the same with the key given (q-plane scan: keys are the bit patterns of q = |F|^2)
__device__ __forceinline__ void scan_stage_key(bool member, uint64_t key, MedianWork& w, int ip, uint64_t* s_buf, unsigned* s_cnt) {
    const int lane = threadIdx.x & 31;
    const unsigned m = __ballot_sync(0xffffffffu, member);
    unsigned b0 = 0;
    if (lane == 0) b0 = atomicAdd(s_cnt, (unsigned)__popc(m));
    b0 = __shfl_sync(0xffffffffu, b0, 0);
    if (member) {
        const unsigned slot = b0 + __popc(m & ((1u << lane) - 1));
        if (slot < SCAN_SBUF) s_buf[slot] = key;
        else {
            const unsigned g = atomicAdd(&w.cand_n[ip], 1u);
            if (g < w.cand_cap) w.cand[(size_t)ip * w.cand_cap + g] = key;
        }
    }
}

// Truly rare (~1e-4 of the bins at magmin = 0.01): a capacity-relevant magnitude.
__device__ __noinline__ void scan_cap_rare(double2 z, double q, uint64_t i, const SpecLayout& lay, const ScanCap& cap, double qcap_lo,
                                           MedianWork& w, int ip, unsigned& capb) {
    const int y = (int)(i / (uint64_t)lay.ld);
    const int x = (int)(i - (uint64_t)y * lay.ld);
    if (!col_weight(lay, x)) return;  // pad column
    const bool axis = y == 0 || x == 0 || y == (lay.PH >> 1) || x == (lay.PW >> 1);  // on_axis S:698 (even sizes)
    const double r = sqrt((double)((long long)y * y + (long long)x * x));        // == hypot for exact integer sums
    if (axis || r < cap.rlo || r > cap.rhi) return;
    if (q < qcap_lo) capb++;
    else {
        const unsigned g = atomicAdd(&w.cap_unc_n[ip], 1u);
        if (g < CAP_UNC_MAX) w.cap_unc[(size_t)ip * CAP_UNC_MAX + g] = mag_key(z);
    }
}

// One streaming read of the plane: weighted count of the elements below the bracket, members staged, and
// (cap.on) the capacity verdicts.  Every stored element is counted / staged with the layout's interior
// weight (2 for a half plane: a bin and its Hermitian mirror) in the flat main loop; the edge columns
// (weight 1) and the zero pad columns (weight 0) are corrected per row afterwards -- counts directly, members
// through the excess list cand_b that the weighted select subtracts -- so the hot loop carries no column
// arithmetic.
__global__ void __launch_bounds__(SCAN_THREADS, 2) median_scan(const double2* __restrict__ spec, SpecLayout lay, MedianWork w,
                                                            const Bracket* __restrict__ br, ScanCap cap) {
    __shared__ uint64_t s_buf[SCAN_SBUF];
    __shared__ unsigned s_cnt, s_base;
    __shared__ long long ws[SCAN_THREADS / 32];
    __shared__ unsigned wc[SCAN_THREADS / 32];
    const int ip = blockIdx.y;
    const uint64_t E = lay.plane_elems();
    const double2* pl = spec + (size_t)ip * E;
    const double qlo = br[ip].qlo, qhi = br[ip].qhi;
    const double qcap_lo = cap.on ? cap.magmin2 * qlo * (1.0 - 1e-9) : 0.0;
    const double qcap_hi = cap.on ? cap.magmin2 * qhi * (1.0 + 1e-9) : -1.0;  // -1: no q qualifies
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    unsigned below = 0, capb = 0;
    const int lane = threadIdx.x & 31;
    const uint64_t ntiles = (E + SCAN_TILE - 1) / SCAN_TILE;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    // (a double-buffered tile loop with 256-thread CTAs measured slower: 9.3 vs 7.3 ms per 64 UHD images for the group)
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint64_t base = t * SCAN_TILE + threadIdx.x;
        double2 z[SCAN_UNROLL];
        if (base - threadIdx.x + SCAN_TILE <= E) {
#pragma unroll
            for (int u = 0; u < SCAN_UNROLL; u++) z[u] = __ldcs(pl + base + (uint64_t)u * SCAN_THREADS);
        } else {
#pragma unroll
            for (int u = 0; u < SCAN_UNROLL; u++) {
                const uint64_t i = base + (uint64_t)u * SCAN_THREADS;
                z[u] = i < E ? __ldcs(pl + i) : make_double2(qnan, 0.0);  // NaN: neither below, member nor tiny
            }
        }
#pragma unroll
        for (int u = 0; u < SCAN_UNROLL; u++) {
            const double q = fma(z[u].x, z[u].x, z[u].y * z[u].y);
            const bool lowq = q < qlo;
            const bool member = !lowq && q <= qhi;
            const bool tiny = q <= qcap_hi;
            below += lowq ? 1u : 0u;
            if (__any_sync(0xffffffffu, member)) scan_stage(member, z[u], w, ip, s_buf, &s_cnt);
            if (tiny) scan_cap_rare(z[u], q, base + (uint64_t)u * SCAN_THREADS, lay, cap, qcap_lo, w, ip, capb);
        }
    }
    long long acc = (long long)below * (lay.half ? 2 : 1);
    if (lay.half) {  // edge columns 0 and PW/2 carry weight 1, pad columns weight 0
        const int h = lay.PW >> 1, nfix = lay.ld - h + 1;  // x = 0, h, h+1 .. ld-1
        // rows blockIdx.x, blockIdx.x + gridDim.x, ...; the (row, fix column) pairs are spread over the whole CTA
        const int nrows_mine = ((int)blockIdx.x < lay.PH) ? (lay.PH - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
        for (int e = threadIdx.x; e < nrows_mine * nfix; e += blockDim.x) {
            const int r = e / nfix, c = e - r * nfix;
            const int y = (int)blockIdx.x + r * (int)gridDim.x;
            const int x = c == 0 ? 0 : h + c - 1;
            const int over = (x == 0 || x == h) ? 1 : 2;
            const double2 v = pl[(size_t)y * lay.ld + x];
            const double q = fma(v.x, v.x, v.y * v.y);
            if (q < qlo) acc -= over;
            else if (q <= qhi) {  // a member staged with the interior weight: list the excess
                const unsigned g = atomicAdd(&w.cand_b_n[ip], (unsigned)over);
                const uint64_t key = mag_key(v);
                for (int k = 0; k < over; k++)
                    if (g + k < CAND_B_MAX) w.cand_b[(size_t)ip * CAND_B_MAX + g + k] = key;
            }
        }
    }
    for (int o = 16; o; o >>= 1) { acc += __shfl_down_sync(0xffffffffu, acc, o); capb += __shfl_down_sync(0xffffffffu, capb, o); }
    if (lane == 0) { ws[threadIdx.x >> 5] = acc; wc[threadIdx.x >> 5] = capb; }
    __syncthreads();
    const unsigned nloc = s_cnt < SCAN_SBUF ? s_cnt : SCAN_SBUF;
    if (threadIdx.x == 0) {
        long long t = 0;
        unsigned c = 0;
        for (int k = 0; k < SCAN_THREADS / 32; k++) { t += ws[k]; c += wc[k]; }
        if (t) atomicAdd((unsigned long long*)&w.counts[ip], (unsigned long long)t);  // wraps correctly for negative partial sums
        if (c) atomicAdd((unsigned long long*)&w.cap_below[ip], (unsigned long long)c);
        s_base = nloc ? atomicAdd(&w.cand_n[ip], nloc) : 0;
    }
    __syncthreads();
    uint64_t* cand = w.cand + (size_t)ip * w.cand_cap;
    for (unsigned i = threadIdx.x; i < nloc; i += blockDim.x)
        if (s_base + i < w.cand_cap) cand[s_base + i] = s_buf[i];
}

// The same scan over the float copy of q the 4096-row column pass left behind (PassArgs::q32, float4 index
// ((g * 16 + k1) * 4 + j) * 32 + lane, lane = 2 m + c: rows k1 + 16 m + 256 (4 j + 0..3) of column 2 g + c): 4 bytes per
// element instead of 16.  A float carries q to 6e-8, so everything outside [qlo (1 - 1e-6), qhi (1 + 1e-6)] is decided
// on the float alone.  The ~1 % inside (and the rare capacity-relevant magnitudes) are queued per warp and tile and then
// settled on the exact spectrum values in one dense pass (independent gathers, no block barrier in the loop).  Column weights are applied directly (a
// float4 holds four rows of ONE column), pad columns are skipped.
constexpr uint32_t Q32_WQ = 256;  // queued elements per warp and tile (1024 elements, ~13 expected); overflow -> device fallback flags
__global__ void __launch_bounds__(SCAN_THREADS, 2) median_scan_q32(const float4* __restrict__ q32, const double2* __restrict__ spec, SpecLayout lay,
                                                                MedianWork w, const Bracket* __restrict__ br, ScanCap cap) {
    __shared__ uint64_t s_buf[SCAN_SBUF];
    // per-warp queues (no block barrier inside the tile loop):
    // element index (24 bits) | float already counted it as below (bit 28) | weight (bits 30..31)
    __shared__ uint32_t s_q[SCAN_THREADS / 32][Q32_WQ];
    __shared__ unsigned s_cnt, s_base;
    __shared__ long long ws[SCAN_THREADS / 32];
    __shared__ unsigned wc[SCAN_THREADS / 32];
    const int ip = blockIdx.y;
    const uint64_t E = lay.plane_elems(), E4 = E / 4;
    const double2* pl = spec + (size_t)ip * E;
    const float4* qp = q32 + (size_t)ip * E4;
    const double qlo = br[ip].qlo, qhi = br[ip].qhi;
    const double qcap_lo = cap.on ? cap.magmin2 * qlo * (1.0 - 1e-9) : 0.0;
    const double qcap_hi = cap.on ? cap.magmin2 * qhi * (1.0 + 1e-9) : -1.0;
    // float thresholds, rounded outwards: below flo -> certainly q < qlo; above fhi -> certainly q > qhi; at most fcap -> look
    const float flo = __double2float_rd(qlo * (1.0 - 1e-6)), fhi = __double2float_ru(qhi * (1.0 + 1e-6));
    const float fcap = cap.on ? __double2float_ru(qcap_hi * (1.0 + 1e-6)) : -1.0f;
    const int hcols = lay.PW >> 1;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    long long acc = 0;
    unsigned capb = 0;
    const int lane = threadIdx.x & 31;
    uint32_t* myq = s_q[threadIdx.x >> 5];
    constexpr uint32_t TILE4 = SCAN_THREADS * SCAN_UNROLL;
    const uint64_t ntiles = (E4 + TILE4 - 1) / TILE4;
    const float fnan = __int_as_float(0x7fc00000);
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint64_t base = t * TILE4 + threadIdx.x;
        float4 f[SCAN_UNROLL];
#pragma unroll
        for (int u = 0; u < SCAN_UNROLL; u++) {
            const uint64_t i4 = base + (uint64_t)u * SCAN_THREADS;
            f[u] = i4 < E4 ? __ldcs(qp + i4) : make_float4(fnan, fnan, fnan, fnan);
        }
        // ---- pass 1: decide on the floats, remember which elements need the exact value (bit 4 u + e)
        unsigned look = 0, counted = 0;
#pragma unroll
        for (int u = 0; u < SCAN_UNROLL; u++) {
            const uint64_t i4 = base + (uint64_t)u * SCAN_THREADS;
            const int x = 2 * (int)(i4 >> 11) + (int)(i4 & 1);
            const int wgt = (x == 0 || x == hcols) ? 1 : (x < hcols ? 2 : 0);  // Hermitian multiplicity of the column
            const float qf[4] = {f[u].x, f[u].y, f[u].z, f[u].w};
            unsigned nb = 0;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const bool b = qf[e] < flo;
                nb += b ? 1u : 0u;
                counted |= b ? (1u << (4 * u + e)) : 0u;
                look |= (wgt != 0 && ((qf[e] >= flo && qf[e] <= fhi) || qf[e] <= fcap)) ? (1u << (4 * u + e)) : 0u;
            }
            acc += (long long)(nb * (unsigned)wgt);
        }
        // the warp's queue: exclusive positions by a prefix sum over the lanes
        const unsigned cnt = __popc(look);
        unsigned incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const unsigned nq_all = __shfl_sync(0xffffffffu, incl, 31);
        if (nq_all == 0) continue;  // (warp-uniform)
        unsigned pos = incl - cnt;
        while (look) {
            const int bit = __ffs(look) - 1;
            look &= look - 1;
            const uint64_t i4 = base + (uint64_t)(bit >> 2) * SCAN_THREADS;
            const int ln = (int)(i4 & 31), j = (int)((i4 >> 5) & 3), k1 = (int)((i4 >> 7) & 15), g = (int)(i4 >> 11);
            const int x = 2 * g + (ln & 1), y = k1 + 16 * (ln >> 1) + 1024 * j + 256 * (bit & 3);
            const unsigned wgt = (x == 0 || x == hcols) ? 1u : 2u;  // (pad columns never get here)
            if (pos < Q32_WQ) myq[pos] = (unsigned)(y * lay.ld + x) | (((counted >> bit) & 1u) << 28) | (wgt << 30);
            pos++;
        }
        __syncwarp();
        // ---- pass 2: the queued elements on their exact values
        const unsigned nq = nq_all < Q32_WQ ? nq_all : Q32_WQ;
        if (nq_all > Q32_WQ && lane == 0) { w.flags[0] = 1; w.flags[1] = 1; }  // hopeless bracket: the exact generic passes take over
        for (unsigned i = lane; i - lane < nq; i += 32) {
            const bool valid = i < nq;
            const unsigned ent = valid ? myq[i] : 0u;
            const unsigned li = ent & 0xFFFFFFu, wgt = ent >> 30;
            bool member = false;
            double2 z = make_double2(0.0, 0.0);
            if (valid) {
                z = pl[li];
                const double q = fma(z.x, z.x, z.y * z.y);
                const bool lowq = q < qlo;
                member = !lowq && q <= qhi;
                if (lowq && !((ent >> 28) & 1u)) acc += wgt;  // (the float alone did not count it)
                if (q <= qcap_hi) scan_cap_rare(z, q, li, lay, cap, qcap_lo, w, ip, capb);
                if (member && wgt == 1) {  // staged with the interior weight below: list the excess
                    const unsigned gb = atomicAdd(&w.cand_b_n[ip], 1u);
                    if (gb < CAND_B_MAX) w.cand_b[(size_t)ip * CAND_B_MAX + gb] = mag_key(z);
                }
            }
            if (__any_sync(0xffffffffu, member)) scan_stage(member, z, w, ip, s_buf, &s_cnt);
        }
        __syncwarp();  // the queue is free for the next tile
    }
    for (int o = 16; o; o >>= 1) { acc += __shfl_down_sync(0xffffffffu, acc, o); capb += __shfl_down_sync(0xffffffffu, capb, o); }
    if (lane == 0) { ws[threadIdx.x >> 5] = acc; wc[threadIdx.x >> 5] = capb; }
    __syncthreads();
    const unsigned nloc = s_cnt < SCAN_SBUF ? s_cnt : SCAN_SBUF;
    if (threadIdx.x == 0) {
        long long t = 0;
        unsigned c = 0;
        for (int k = 0; k < SCAN_THREADS / 32; k++) { t += ws[k]; c += wc[k]; }
        if (t) atomicAdd((unsigned long long*)&w.counts[ip], (unsigned long long)t);
        if (c) atomicAdd((unsigned long long*)&w.cap_below[ip], (unsigned long long)c);
        s_base = nloc ? atomicAdd(&w.cand_n[ip], nloc) : 0;
    }
    __syncthreads();
    uint64_t* cand = w.cand + (size_t)ip * w.cand_cap;
    for (unsigned i = threadIdx.x; i < nloc; i += blockDim.x)
        if (s_base + i < w.cand_cap) cand[s_base + i] = s_buf[i];
}


// The scan over the two 32-bit planes of q the column-resident embed pass leaves behind (PassArgs::qhi / qlo, the layout
// of q32: word index ((g * 16 + k1) * 4 + j) * 128 + 4 lane + e holds row k1 + 16 m + 256 (4 j + e) of column 2 g + c,
// lane = 2 m + c).  Non-negative doubles order like their bit patterns, so the top word alone decides every element
// whose top word differs from those of the bracket edges; the rest (and the capacity-relevant tiny ones) are queued per
// warp and settled on the exact double (top and low word gathered) -- the same counts and members as a scan of q itself.
// Keys are the bit patterns of q; the median is sqrt() of the selected one (median_members / capacity_resolve, qkeys = 1).
__device__ __forceinline__ void q64_coords(unsigned li, int& y, int& x) {
    const unsigned e = li & 3u, i4 = li >> 2;
    const int ln = (int)(i4 & 31u), j = (int)((i4 >> 5) & 3u), k1 = (int)((i4 >> 7) & 15u), g = (int)(i4 >> 11);
    x = 2 * g + (ln & 1);
    y = k1 + 16 * (ln >> 1) + 1024 * j + 256 * (int)e;
}
__device__ __noinline__ void scan_cap_rare_q(uint64_t key, double q, unsigned li, const SpecLayout& lay, const ScanCap& cap, double qcap_lo,
                                             MedianWork& w, int ip, unsigned& capb) {
    int y, x;
    q64_coords(li, y, x);
    if (!col_weight(lay, x)) return;  // pad column
    const bool axis = y == 0 || x == 0 || y == (lay.PH >> 1) || x == (lay.PW >> 1);  // on_axis S:698 (even sizes)
    const double r = sqrt((double)((long long)y * y + (long long)x * x));
    if (axis || r < cap.rlo || r > cap.rhi) return;
    if (q < qcap_lo) capb++;
    else {
        const unsigned g = atomicAdd(&w.cap_unc_n[ip], 1u);
        if (g < CAP_UNC_MAX) w.cap_unc[(size_t)ip * CAP_UNC_MAX + g] = key;
    }
}
__global__ void __launch_bounds__(SCAN_THREADS, 2) median_scan_q64(const uint4* __restrict__ qhi, const uint32_t* __restrict__ qlo, SpecLayout lay,
                                                                MedianWork w, const Bracket* __restrict__ br, ScanCap cap) {
    __shared__ uint64_t s_buf[SCAN_SBUF];
    __shared__ uint32_t s_q[SCAN_THREADS / 32][Q32_WQ];  // element index (24 bits) | counted as below (bit 28) | weight (bits 30..31)
    __shared__ unsigned s_cnt, s_base;
    __shared__ long long ws[SCAN_THREADS / 32];
    __shared__ unsigned wc[SCAN_THREADS / 32];
    const int ip = blockIdx.y;
    const uint64_t E = lay.plane_elems(), E4 = E / 4;
    const uint4* hp = qhi + (size_t)ip * E4;
    const uint32_t* hw = (const uint32_t*)hp;
    const uint32_t* lw = qlo + (size_t)ip * E;
    const double qlo_d = br[ip].qlo, qhi_d = br[ip].qhi;
    const double qcap_lo = cap.on ? cap.magmin2 * qlo_d * (1.0 - 1e-9) : 0.0;
    const double qcap_hi = cap.on ? cap.magmin2 * qhi_d * (1.0 + 1e-9) : -1.0;
    // top words of the edges: below hlo -> certainly q < qlo; above hhi -> certainly q > qhi; at most hcap -> look
    const unsigned hlo = (unsigned)__double2hiint(qlo_d), hhi = (unsigned)__double2hiint(qhi_d);
    const bool capon = cap.on && qcap_hi >= 0.0;
    const unsigned hcap = capon ? (unsigned)__double2hiint(qcap_hi) : 0u;
    const int hcols = lay.PW >> 1;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    long long acc = 0;
    unsigned capb = 0;
    const int lane = threadIdx.x & 31;
    uint32_t* myq = s_q[threadIdx.x >> 5];
    constexpr uint32_t TILE4 = SCAN_THREADS * SCAN_UNROLL;
    const uint64_t ntiles = (E4 + TILE4 - 1) / TILE4;
    for (uint64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const uint64_t base = t * TILE4 + threadIdx.x;
        uint4 f[SCAN_UNROLL];
#pragma unroll
        for (int u = 0; u < SCAN_UNROLL; u++) {
            const uint64_t i4 = base + (uint64_t)u * SCAN_THREADS;
            f[u] = i4 < E4 ? __ldcs(hp + i4) : make_uint4(~0u, ~0u, ~0u, ~0u);  // all ones: above everything
        }
        unsigned look = 0, counted = 0;
#pragma unroll
        for (int u = 0; u < SCAN_UNROLL; u++) {
            const uint64_t i4 = base + (uint64_t)u * SCAN_THREADS;
            const int x = 2 * (int)(i4 >> 11) + (int)(i4 & 1);
            const int wgt = (x == 0 || x == hcols) ? 1 : (x < hcols ? 2 : 0);
            const unsigned hv[4] = {f[u].x, f[u].y, f[u].z, f[u].w};
            unsigned nb = 0;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const bool b = hv[e] < hlo;
                nb += b ? 1u : 0u;
                counted |= b ? (1u << (4 * u + e)) : 0u;
                look |= (wgt != 0 && ((hv[e] >= hlo && hv[e] <= hhi) || (capon && hv[e] <= hcap))) ? (1u << (4 * u + e)) : 0u;
            }
            acc += (long long)(nb * (unsigned)wgt);
        }
        const unsigned cnt = __popc(look);
        unsigned incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const unsigned nq_all = __shfl_sync(0xffffffffu, incl, 31);
        if (nq_all == 0) continue;  // (warp-uniform)
        unsigned pos = incl - cnt;
        while (look) {
            const int bit = __ffs(look) - 1;
            look &= look - 1;
            const uint64_t i4 = base + (uint64_t)(bit >> 2) * SCAN_THREADS;
            const int x = 2 * (int)(i4 >> 11) + (int)(i4 & 1);
            const unsigned wgt = (x == 0 || x == hcols) ? 1u : 2u;  // (pad columns never get here)
            if (pos < Q32_WQ) myq[pos] = (unsigned)(i4 * 4 + (unsigned)(bit & 3)) | (((counted >> bit) & 1u) << 28) | (wgt << 30);
            pos++;
        }
        __syncwarp();
        const unsigned nq = nq_all < Q32_WQ ? nq_all : Q32_WQ;
        if (nq_all > Q32_WQ && lane == 0) { w.flags[0] = 1; w.flags[1] = 1; }  // hopeless bracket: the exact generic passes take over
        for (unsigned i = lane; i - lane < nq; i += 32) {
            const bool valid = i < nq;
            const unsigned ent = valid ? myq[i] : 0u;
            const unsigned li = ent & 0xFFFFFFu, wgt = ent >> 30;
            bool member = false;
            uint64_t key = 0;
            if (valid) {
                const unsigned hi = hw[li], lo = lw[li];
                key = ((uint64_t)hi << 32) | lo;
                const double q = __hiloint2double((int)hi, (int)lo);
                const bool lowq = q < qlo_d;
                member = !lowq && q <= qhi_d;
                if (lowq && !((ent >> 28) & 1u)) acc += wgt;  // (the top word alone did not count it)
                if (q <= qcap_hi) scan_cap_rare_q(key, q, li, lay, cap, qcap_lo, w, ip, capb);
                if (member && wgt == 1) {
                    const unsigned gb = atomicAdd(&w.cand_b_n[ip], 1u);
                    if (gb < CAND_B_MAX) w.cand_b[(size_t)ip * CAND_B_MAX + gb] = key;
                }
            }
            if (__any_sync(0xffffffffu, member)) scan_stage_key(member, key, w, ip, s_buf, &s_cnt);
        }
        __syncwarp();
    }
    for (int o = 16; o; o >>= 1) { acc += __shfl_down_sync(0xffffffffu, acc, o); capb += __shfl_down_sync(0xffffffffu, capb, o); }
    if (lane == 0) { ws[threadIdx.x >> 5] = acc; wc[threadIdx.x >> 5] = capb; }
    __syncthreads();
    const unsigned nloc = s_cnt < SCAN_SBUF ? s_cnt : SCAN_SBUF;
    if (threadIdx.x == 0) {
        long long t = 0;
        unsigned c = 0;
        for (int k = 0; k < SCAN_THREADS / 32; k++) { t += ws[k]; c += wc[k]; }
        if (t) atomicAdd((unsigned long long*)&w.counts[ip], (unsigned long long)t);
        if (c) atomicAdd((unsigned long long*)&w.cap_below[ip], (unsigned long long)c);
        s_base = nloc ? atomicAdd(&w.cand_n[ip], nloc) : 0;
    }
    __syncthreads();
    uint64_t* cand = w.cand + (size_t)ip * w.cand_cap;
    for (unsigned i = threadIdx.x; i < nloc; i += blockDim.x)
        if (s_base + i < w.cand_cap) cand[s_base + i] = s_buf[i];
}

// generic (fallback) radix pass over the q planes: histogram of digit d of the keys (bit patterns of q)
__global__ void __launch_bounds__(512) median_hist_q64(const uint32_t* __restrict__ qhi, const uint32_t* __restrict__ qlo, SpecLayout lay, int d,
                                                       MedianWork w, const int* gate) {
    __shared__ uint32_t sh[RADIX];
    if (gate && !*gate) return;
    const int ip = blockIdx.y;
    for (int i = threadIdx.x; i < RADIX; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint64_t E = lay.plane_elems();
    const uint32_t* hw = qhi + (size_t)ip * E;
    const uint32_t* lw = qlo + (size_t)ip * E;
    const int sft = key_shift(d), wid = key_width(d);
    const uint64_t prefix = w.prefix[ip];
    const int hi_sft = sft + wid;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < E; i += (uint64_t)gridDim.x * blockDim.x) {
        int y, x;
        q64_coords((unsigned)i, y, x);
        const int wt = col_weight(lay, x);
        if (!wt) continue;
        const uint64_t k = ((uint64_t)hw[i] << 32) | lw[i];
        if (d == 0 || (k >> hi_sft) == (prefix >> hi_sft))
            atomicAdd(&sh[(unsigned)(k >> sft) & ((1u << wid) - 1)], (unsigned)wt);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RADIX; i += blockDim.x)
        if (sh[i]) atomicAdd(&w.hist[ip * RADIX + i], sh[i]);
}
// q keys -> magnitudes, for every plane (the selection ran on the bit patterns of q)
__global__ void median_sqrt_keys(double* median, int nplanes) {
    const int ip = blockIdx.x * blockDim.x + threadIdx.x;
    if (ip < nplanes) median[ip] = sqrt(median[ip]);
}
// element index of (y, x) in the q planes of a 4096-row half plane
__device__ __forceinline__ unsigned q64_index(int y, int x) {
    const unsigned g = (unsigned)x >> 1, c = (unsigned)x & 1u, k1 = (unsigned)y & 15u, m = ((unsigned)y >> 4) & 15u, k3 = (unsigned)y >> 8;
    const unsigned i4 = ((g * 16 + k1) * 4 + (k3 >> 2)) * 32 + 2 * m + c;
    return i4 * 4 + (k3 & 3u);
}
__global__ void __launch_bounds__(256) capacity_count_q64(const uint32_t* __restrict__ qhi, const uint32_t* __restrict__ qlo, SpecLayout lay, int ymax,
                                                          int xmax, double rlo, double rhi, double magmin, const double* __restrict__ median,
                                                          uint64_t* counts, const int* gate) {
    if (gate && !*gate) return;
    const int ip = blockIdx.y;
    const int PH = lay.PH, PW = lay.PW;
    const double thr = magmin * median[ip];
    const uint64_t E = lay.plane_elems();
    const uint32_t* hw = qhi + (size_t)ip * E;
    const uint32_t* lw = qlo + (size_t)ip * E;
    const long long box = (long long)(ymax + 1) * (xmax + 1);
    unsigned local = 0;
    const unsigned bw = (unsigned)(xmax + 1);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < box; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)((unsigned)i / bw), x = (int)((unsigned)i % bw);
        if (y == 0 || x == 0 || y == PH / 2 || x == PW / 2) continue;
        const double r = sqrt((double)((long long)y * y + (long long)x * x));
        if (r < rlo || r > rhi) continue;
        const unsigned li = q64_index(y, x);
        if (sqrt(__hiloint2double((int)hw[li], (int)lw[li])) < thr) continue;  // abs(F) < t (S:1004)
        local++;
    }
    for (int o = 16; o; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    __shared__ unsigned ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += ws[i];
        if (t) atomicAdd((unsigned long long*)&counts[ip], (unsigned long long)t);
    }
}

// one CTA per plane: exact rank among the members, or raise the fallback flag
__global__ void __launch_bounds__(1024) median_members(MedianWork w, uint64_t P, uint32_t wa, double* median, int* flag, uint32_t guard) {
    __shared__ SelectScratch sc;
    const int ip = blockIdx.x;
    const uint32_t n = w.cand_n[ip], nb = w.cand_b_n[ip];
    const uint64_t below = w.counts[ip], rank = P / 2;
    const uint64_t total = (uint64_t)n * wa - nb;  // weighted member count
    const bool ok = n <= w.cand_cap && nb <= CAND_B_MAX && nb <= (uint64_t)n * wa && rank >= below + guard && rank + guard < below + total;
    if (!ok) {
        if (threadIdx.x == 0) *flag = 1;
        return;
    }
    const uint64_t k = select_rank(w.cand + (size_t)ip * w.cand_cap, n, wa, w.cand_b + (size_t)ip * CAND_B_MAX, nb, rank - below, &sc);
    if (threadIdx.x == 0) median[ip] = __longlong_as_double((long long)k);
}

// Capacity count (S:999-1007): annulus bins in index space (radius from bin (0,0), scaled by
// min(PH,PW)), off the axes (on_axis S:698), |F| >= magmin*median, conjugate distinct.
// Only the quarter-disc y,x <= rhi can satisfy the radius test, so only that box is scanned.
__global__ void __launch_bounds__(256) capacity_count(const double2* __restrict__ spec, SpecLayout lay, int ymax, int xmax,
                                                      double rlo, double rhi, double magmin,
                                                      const double* __restrict__ median, uint64_t* counts, const int* gate) {
    if (gate && !*gate) return;  // the fused count of the median scan stands
    const int ip = blockIdx.y;
    const int PH = lay.PH, PW = lay.PW;
    const double thr = magmin * median[ip];
    const double thr2_lo = thr * thr * (1.0 - 1e-12), thr2_hi = thr * thr * (1.0 + 1e-12);
    const double2* pl = spec + (size_t)ip * lay.plane_elems();
    const long long box = (long long)(ymax + 1) * (xmax + 1);
    unsigned local = 0;
    const unsigned bw = (unsigned)(xmax + 1);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < box; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)((unsigned)i / bw), x = (int)((unsigned)i % bw);  // box < 2^32
        if (y == 0 || x == 0 || y == PH / 2 || x == PW / 2) continue;  // PH, PW are even powers of two
        const double r = sqrt((double)((long long)y * y + (long long)x * x));  // == hypot for exact integer sums
        if (r < rlo || r > rhi) continue;
        const double2 z = spec_load(pl, lay, y, x);
        // |F| >= thr decided on q = re^2+im^2 (2 FP64 ops); hypot() only inside a 1e-12 guard band around thr^2
        const double q = fma(z.x, z.x, z.y * z.y);
        if (q < thr2_lo) continue;
        if (q <= thr2_hi && hypot(z.x, z.y) < thr) continue;
        local++;  // conjugate (PH-y, PW-x) != (y,x) is implied by the axis test
    }
    // warp + block reduce
    for (int o = 16; o; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    __shared__ unsigned ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += ws[i];
        if (t) atomicAdd((unsigned long long*)&counts[ip], (unsigned long long)t);
    }
}
__global__ void capacity_finish(const uint64_t* counts, uint64_t* usable, int nimg) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nimg) usable[i] = counts[3 * i] / 2 + counts[3 * i + 1] / 2 + counts[3 * i + 2] / 2;  // S:1006, S:1008
}

// Fused capacity, step 2 (one warp per plane): settle the undecided annulus bins against the exact
// threshold magmin*median (abs(F) < t, S:1004).  A plane whose list overflowed raises flags[1] and the
// whole batch is recounted by capacity_count.
__global__ void __launch_bounds__(128) capacity_resolve(MedianWork w, int nplanes, double magmin, const double* __restrict__ median,
                                                        uint64_t ann_total, int qkeys) {
    const int ip = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (ip >= nplanes) return;
    const uint32_t n = w.cap_unc_n[ip];
    if (n > CAP_UNC_MAX) {
        if (lane == 0) { w.flags[1] = 1; w.counts[ip] = 0; }
        return;
    }
    const double thr = magmin * median[ip];
    unsigned c = 0;
    for (uint32_t i = lane; i < n; i += 32) {
        double v = __longlong_as_double((long long)w.cap_unc[(size_t)ip * CAP_UNC_MAX + i]);
        if (qkeys) v = sqrt(v);  // keys from the q planes are |F|^2
        c += v < thr ? 1u : 0u;
    }
    for (int o = 16; o; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if (lane == 0) w.counts[ip] = ann_total - w.cap_below[ip] - c;
}
__global__ void capacity_zero_gated(uint64_t* counts, int nplanes, const int* gate) {
    if (!*gate) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nplanes) counts[i] = 0;
}

// number of off-axis bins inside the annulus (geometry only; the loop of count_plane S:999-1003 without the
// magnitude test), cached per geometry
static uint64_t annulus_total_host(int PH, int PW, int ymax, int xmax, double rlo, double rhi) {
    struct Key { int PH, PW; double rlo, rhi; uint64_t n; };
    static std::vector<Key> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    for (const Key& k : cache)
        if (k.PH == PH && k.PW == PW && k.rlo == rlo && k.rhi == rhi) return k.n;
    uint64_t n = 0;
    for (int y = 1; y <= ymax; y++) {
        if (y == PH / 2) continue;
        for (int x = 1; x <= xmax; x++) {
            if (x == PW / 2) continue;
            const double r = sqrt((double)((long long)y * y + (long long)x * x));
            if (r >= rlo && r <= rhi) n++;
        }
    }
    if (cache.size() > 64) cache.clear();
    cache.push_back(Key{PH, PW, rlo, rhi, n});
    return n;
}

cudaError_t launch_median_capacity(const Launcher& L, const double2* spec, int nplanes, SpecLayout lay,
                                   double magmin, double rlo, double rhi, MedianWork w,
                                   double* d_median, uint64_t* d_usable, unsigned presampled, const float* q32,
                                   const uint32_t* qhi, const uint32_t* qlo) {
    const bool q64 = qhi != nullptr && qlo != nullptr;  // no spectrum: everything runs on the two word planes of q
    if (q64 && !(lay.half && lay.PH == 4096 && presampled)) return cudaErrorInvalidValue;
    const int PH = lay.PH, PW = lay.PW;
    const uint64_t P = (uint64_t)PH * PW;        // size of the full multiset (ranks refer to it)
    const uint64_t E = lay.plane_elems();        // stored elements per plane
    median_init<<<(nplanes * RADIX + 255) / 256, 256, 0, L.stream>>>(w, nplanes, P);
    TFFT_LAUNCH_CHECK(L);
    int* d_flag = w.flags;
    Bracket* br = (Bracket*)w.prefix2;
    const bool all = P <= (uint64_t)w.cand_cap && P <= (uint64_t)SAMPLE_MAX;  // small plane: everything is a member
    if (all) {
        median_bracket_all<<<(nplanes + 255) / 256, 256, 0, L.stream>>>(w, nplanes, br);
        TFFT_LAUNCH_CHECK(L);
    } else if (presampled && presampled <= w.cand_cap) {
        median_bracket<<<nplanes, 1024, 0, L.stream>>>(w, presampled, br, 1);
        TFFT_LAUNCH_CHECK(L);
    } else {
        const uint64_t logical = lay.half ? P / 2 : P;
        const uint32_t S = SAMPLE_MAX < w.cand_cap ? SAMPLE_MAX : w.cand_cap;
        const uint64_t stride = logical / S;
        median_sample<<<dim3((S + 255) / 256 > 256 ? 256 : (S + 255) / 256, (unsigned)nplanes), 256, 0, L.stream>>>(spec, lay, S, stride, w);
        TFFT_LAUNCH_CHECK(L);
        median_bracket<<<nplanes, 1024, 0, L.stream>>>(w, S, br, 0);
        TFFT_LAUNCH_CHECK(L);
    }
    int ymax = (int)floor(rhi), xmax = (int)floor(rhi);
    if (ymax > PH - 1) ymax = PH - 1;
    if (xmax > PW - 1) xmax = PW - 1;
    if (ymax < 0) ymax = 0;
    if (xmax < 0) xmax = 0;
    // the capacity count rides on the scan when every annulus bin is a stored element (x <= rhi < PW/2 for a half plane)
    ScanCap cap;
    cap.on = (d_usable && !all && magmin >= 0.0 && (!lay.half || xmax < PW / 2)) ? 1 : 0;
    cap.magmin2 = magmin * magmin; cap.rlo = rlo; cap.rhi = rhi;
    {
        const uint64_t ntiles = (E + SCAN_TILE - 1) / SCAN_TILE;
        // ~7 tiles (28 K elements, ~340 members) per CTA: the members fit the CTA's staging buffer (an overflowing CTA
        // falls back to one global atomic per member, which is what made fewer, longer CTAs slower)
        unsigned cc = (unsigned)((ntiles + 6) / 7);
        if (cc < 1) cc = 1;
        const unsigned per_plane = (unsigned)(ntiles < cc ? ntiles : cc);
        if (q64) {
            const uint64_t nt4 = (E / 4 + SCAN_TILE - 1) / SCAN_TILE;
            unsigned c4 = (unsigned)((nt4 + 1) / 2);
            if (c4 < 1) c4 = 1;
            median_scan_q64<<<dim3((unsigned)(nt4 < c4 ? nt4 : c4), (unsigned)nplanes), SCAN_THREADS, 0, L.stream>>>((const uint4*)qhi, qlo, lay, w, br, cap);
        } else if (q32 && lay.half && PH == 4096 && !all) {
            const uint64_t nt4 = (E / 4 + SCAN_TILE - 1) / SCAN_TILE;  // tiles of float4 (four elements each)
            unsigned c4 = (unsigned)((nt4 + 1) / 2);  // ~2 tiles (32 K elements) per CTA, as above
            if (c4 < 1) c4 = 1;
            median_scan_q32<<<dim3((unsigned)(nt4 < c4 ? nt4 : c4), (unsigned)nplanes), SCAN_THREADS, 0, L.stream>>>((const float4*)q32, spec, lay, w, br, cap);
        } else {
            median_scan<<<dim3(per_plane, (unsigned)nplanes), SCAN_THREADS, 0, L.stream>>>(spec, lay, w, br, cap);
        }
        TFFT_LAUNCH_CHECK(L);
    }
    median_members<<<nplanes, 1024, 0, L.stream>>>(w, P, lay.half ? 2u : 1u, d_median, d_flag, all ? 0u : RANK_GUARD);
    TFFT_LAUNCH_CHECK(L);
    // Fallback, decided on the device (no host sync): the generic radix passes are gated by the
    // flag and exit immediately in the normal case.
    {
        const dim3 fgrid(8, (unsigned)nplanes);  // rarely does real work: keep the gated launches cheap
        for (int d = 0; d < NUM_DIGITS; d++) {
            if (q64) median_hist_q64<<<fgrid, 512, 0, L.stream>>>(qhi, qlo, lay, d, w, d_flag);
            else median_hist<<<fgrid, 512, 0, L.stream>>>(spec, lay, d, w, d_flag);
            TFFT_LAUNCH_CHECK(L);
            median_pick<<<nplanes, 256, 0, L.stream>>>(d, w, d_flag);
            TFFT_LAUNCH_CHECK(L);
        }
        median_from_prefix<<<(nplanes + 255) / 256, 256, 0, L.stream>>>(w, d_median, nplanes, d_flag);
        TFFT_LAUNCH_CHECK(L);
        if (q64) {  // the selection ran on q = |F|^2
            median_sqrt_keys<<<(nplanes + 255) / 256, 256, 0, L.stream>>>(d_median, nplanes);
            TFFT_LAUNCH_CHECK(L);
        }
    }
    if (d_usable) {
        const long long box = (long long)(ymax + 1) * (xmax + 1);
        int cb = (int)((box + 256 * 8 - 1) / (256 * 8));
        if (cb > 1184) cb = 1184;
        if (cb < 1) cb = 1;
        const dim3 cgrid((unsigned)cb, (unsigned)nplanes);
        if (cap.on) {
            const uint64_t ann = annulus_total_host(PH, PW, ymax, xmax, rlo, rhi);
            capacity_resolve<<<(nplanes + 3) / 4, 128, 0, L.stream>>>(w, nplanes, magmin, d_median, ann, q64 ? 1 : 0);
            TFFT_LAUNCH_CHECK(L);
            capacity_zero_gated<<<(nplanes + 255) / 256, 256, 0, L.stream>>>(w.counts, nplanes, d_flag + 1);
            TFFT_LAUNCH_CHECK(L);
            // gated recount: almost never does real work, so keep the launch small (grid-stride loop inside)
            if (q64) capacity_count_q64<<<dim3(cgrid.x < 16 ? cgrid.x : 16, cgrid.y), 256, 0, L.stream>>>(qhi, qlo, lay, ymax, xmax, rlo, rhi, magmin, d_median, w.counts, d_flag + 1);
            else capacity_count<<<dim3(cgrid.x < 16 ? cgrid.x : 16, cgrid.y), 256, 0, L.stream>>>(spec, lay, ymax, xmax, rlo, rhi, magmin, d_median, w.counts, d_flag + 1);
            TFFT_LAUNCH_CHECK(L);
        } else {
            cudaError_t e = cudaMemsetAsync(w.counts, 0, sizeof(uint64_t) * nplanes, L.stream);
            if (e != cudaSuccess) return e;
            if (q64) capacity_count_q64<<<cgrid, 256, 0, L.stream>>>(qhi, qlo, lay, ymax, xmax, rlo, rhi, magmin, d_median, w.counts, nullptr);
            else capacity_count<<<cgrid, 256, 0, L.stream>>>(spec, lay, ymax, xmax, rlo, rhi, magmin, d_median, w.counts, nullptr);
            TFFT_LAUNCH_CHECK(L);
        }
        capacity_finish<<<(nplanes / 3 + 255) / 256, 256, 0, L.stream>>>(w.counts, d_usable, nplanes / 3);
        TFFT_LAUNCH_CHECK(L);
    }
    return cudaSuccess;
}

// --------------------------------------------------------------------------------------------
// Embed scatter (write_bit_on_bin S:712-732): |F| kept (floored at 1e-12), phase set to
// +-alpha (+ jitter), conjugate bin mirrored so the plane stays real.  Bins are unique and
// never alias a conjugate (Turtle::mark_here S:805-809), so the scatter is conflict-free.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_scatter(double2* spec, SpecLayout lay, const uint32_t* __restrict__ bins,
                                                     const uint8_t* __restrict__ bits, size_t nbits,
                                                     const double* __restrict__ jitter, double alpha,
                                                     double cos_a, double sin_a, const uint64_t* __restrict__ usable,
                                                     const double* __restrict__ median /*adaptive alpha: [nimg*3], else null*/) {
    const int img = blockIdx.y;
    if (usable && usable[img] < (uint64_t)nbits) return;  // S:1009: over capacity -> image untouched
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbits) return;
    const int PH = lay.PH, PW = lay.PW;
    const uint32_t b = bins[i];
    const int p = (int)(b >> 30);
    const uint32_t lin = b & 0x3FFFFFFFu;
    const int y = (int)(lin / (uint32_t)PW), x = (int)(lin % (uint32_t)PW);
    const int cy = (PH - y) & (PH - 1), cx = (PW - x) & (PW - 1);  // conj_idx S:370-372 (powers of two)
    double2* pl = spec + (size_t)(img * 3 + p) * lay.plane_elems();
    // which of (bin, mirror) live in the workspace: full layout both; half layout the one(s) with x <= PW/2
    const bool have_bin = !lay.half || x <= (PW >> 1);
    const bool have_mir = !lay.half || cx <= (PW >> 1);
    const double2 z = have_bin ? pl[(size_t)y * lay.ld + x] : pl[(size_t)cy * lay.ld + cx];  // |conj| == |z|
    const double mag = fmax(1e-12, hypot(z.x, z.y));
    const int bit = bits[(size_t)img * nbits + i];
    double c, s;
    if (jitter || median) {
        double al = alpha;
        if (median) al *= fmin(2.0, fmax(0.5, mag / fmax(1e-12, median[img * 3 + p])));  // compute_adaptive_alpha S:704-710
        const double theta = (bit ? al : -al) + (jitter ? jitter[i] : 0.0);
        sincos(theta, &s, &c);
    } else {
        c = cos_a;                 // cos(-a) == cos(a), sin(-a) == -sin(a) exactly
        s = bit ? sin_a : -sin_a;
    }
    const double2 nv = make_double2(mag * c, mag * s);  // std::polar(mag, theta)
    if (cy == y && cx == x) {
        pl[(size_t)y * lay.ld + x] = make_double2(mag, 0.0);  // S:727
    } else {
        if (have_bin) pl[(size_t)y * lay.ld + x] = nv;
        if (have_mir) pl[(size_t)cy * lay.ld + cx] = make_double2(nv.x, -nv.y);
    }
}

cudaError_t launch_embed(const Launcher& L, double2* spec, int nimg, SpecLayout lay,
                         const uint32_t* bins, const uint8_t* bits, size_t nbits, const double* jitter,
                         double alpha, double cos_a, double sin_a, const uint64_t* usable, const double* adaptive_median) {
    if (nbits == 0 || nimg == 0) return cudaSuccess;
    dim3 grid((unsigned)((nbits + 255) / 256), (unsigned)nimg);
    embed_scatter<<<grid, 256, 0, L.stream>>>(spec, lay, bins, bits, nbits, jitter, alpha, cos_a, sin_a, usable, adaptive_median);
    TFFT_LAUNCH_CHECK(L);
    return cudaSuccess;
}

// --------------------------------------------------------------------------------------------
// Inputs of the column-resident embed pass (pencil_col_embed_w, tfft_pencil.cu): the bin list as one 16-bit mask per
// thread and column pair.  Thread tid = 2 (16 k1 + m) + c of pair g owns rows k1 + 16 m + 256 k3 of column 2 g + c, so
// bin (plane, y, x) is bit k3 = y >> 8 of slot ((plane * groups + x >> 1) * 512 + 2 (16 (y & 15) + (y >> 4 & 15)) + (x & 1)).
// pres: which bins exist (built once per call, the list is shared by the batch); val: the bit to write (per image).
// Bins are unique (Turtle::mark_here S:805-809), so every mask bit has one writer.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ void embed_slot(uint32_t b, int lw, int groups, size_t& word, unsigned& shift) {
    const uint32_t lin = b & 0x3FFFFFFFu, p = b >> 30;
    const unsigned y = lin >> lw, x = lin & ((1u << lw) - 1u);
    const unsigned tid = ((((y & 15u) << 4) | ((y >> 4) & 15u)) << 1) | (x & 1u);
    const size_t s = ((size_t)p * groups + (x >> 1)) * 512 + tid;
    word = s >> 1;
    shift = (y >> 8) + 16u * (unsigned)(s & 1);
}
__global__ void __launch_bounds__(256) embed_pres_build(const uint32_t* __restrict__ bins, size_t nbits, int lw, int groups, uint32_t* pres32) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbits) return;
    size_t word; unsigned shift;
    embed_slot(bins[i], lw, groups, word, shift);
    atomicOr(pres32 + word, 1u << shift);
}
__global__ void __launch_bounds__(256) embed_val_build(const uint32_t* __restrict__ bins, const uint8_t* __restrict__ bits, size_t nbits, int lw,
                                                       int groups, uint32_t* val32) {
    const int img = blockIdx.y;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbits || !bits[(size_t)img * nbits + i]) return;
    size_t word; unsigned shift;
    embed_slot(bins[i], lw, groups, word, shift);
    atomicOr(val32 + (size_t)img * 3 * groups * 256 + word, 1u << shift);
}
// what a device bin list needs from the fused pass: out[0] = some bin outside 0 < x < PW/2 (or a bad plane / index),
// out[1] = largest row
__global__ void __launch_bounds__(256) embed_bins_check(const uint32_t* __restrict__ bins, size_t nbits, SpecLayout lay, unsigned* out) {
    unsigned bad = 0, ry = 0;
    const uint32_t P = (uint32_t)min((size_t)lay.PH * lay.PW, (size_t)0x40000000u);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbits; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t lin = bins[i] & 0x3FFFFFFFu;
        const unsigned y = lin / (uint32_t)lay.PW, x = lin % (uint32_t)lay.PW;
        bad |= ((bins[i] >> 30) > 2u || lin >= P || x == 0u || x >= (unsigned)(lay.PW >> 1)) ? 1u : 0u;
        ry = max(ry, y);
    }
    bad = __reduce_max_sync(0xffffffffu, bad);
    ry = __reduce_max_sync(0xffffffffu, ry);
    if ((threadIdx.x & 31) == 0) { if (bad) atomicMax(out, 1u); atomicMax(out + 1, ry); }
}
cudaError_t launch_embed_bins_check(const Launcher& L, const uint32_t* bins, size_t nbits, SpecLayout lay, unsigned* d_out2) {
    cudaError_t e = cudaMemsetAsync(d_out2, 0, 2 * sizeof(unsigned), L.stream);
    if (e != cudaSuccess || nbits == 0) return e;
    const unsigned grid = (unsigned)std::min<size_t>((nbits + 255) / 256, 4 * (size_t)L.sm_count);
    embed_bins_check<<<grid, 256, 0, L.stream>>>(bins, nbits, lay, d_out2);
    TFFT_LAUNCH_CHECK(L);
    return cudaSuccess;
}
// per (plane % 3, column pair): does any thread of the pair hold a bin?  Pairs without bins skip the phase write, the inverse
// transform and the store of the column-resident pass (their columns come back unchanged).  [3 * groups] bytes behind pres.
__global__ void __launch_bounds__(256) embed_pair_flags(const uint32_t* __restrict__ pres32, uint8_t* __restrict__ flags) {
    const uint32_t w = pres32[(size_t)blockIdx.x * 256 + threadIdx.x];
    const int any = __syncthreads_or(w != 0u);
    if (threadIdx.x == 0) flags[blockIdx.x] = any ? 1 : 0;
}
size_t embed_mask_bytes(int ld, int nplanes) { return (size_t)nplanes * (ld / 2) * 512 * sizeof(uint16_t); }
size_t embed_pres_bytes(int ld) { return embed_mask_bytes(ld, 3) + (size_t)3 * (ld / 2); }  // masks + pair flags
cudaError_t launch_embed_pres(const Launcher& L, const uint32_t* bins, size_t nbits, SpecLayout lay, uint16_t* pres) {
    cudaError_t e = cudaMemsetAsync(pres, 0, embed_pres_bytes(lay.ld), L.stream);
    if (e != cudaSuccess || nbits == 0) return e;
    int lw = 0;
    while ((1 << lw) < lay.PW) lw++;
    embed_pres_build<<<(unsigned)((nbits + 255) / 256), 256, 0, L.stream>>>(bins, nbits, lw, lay.ld / 2, (uint32_t*)pres);
    TFFT_LAUNCH_CHECK(L);
    embed_pair_flags<<<(unsigned)(3 * (lay.ld / 2)), 256, 0, L.stream>>>((const uint32_t*)pres, (uint8_t*)pres + embed_mask_bytes(lay.ld, 3));
    TFFT_LAUNCH_CHECK(L);
    return cudaSuccess;
}
cudaError_t launch_embed_val(const Launcher& L, const uint32_t* bins, const uint8_t* bits, size_t nbits, int nimg, SpecLayout lay, uint16_t* val) {
    cudaError_t e = cudaMemsetAsync(val, 0, embed_mask_bytes(lay.ld, nimg * 3), L.stream);
    if (e != cudaSuccess || nbits == 0 || nimg == 0) return e;
    int lw = 0;
    while ((1 << lw) < lay.PW) lw++;
    embed_val_build<<<dim3((unsigned)((nbits + 255) / 256), (unsigned)nimg), 256, 0, L.stream>>>(bins, bits, nbits, lw, lay.ld / 2, (uint32_t*)val);
    TFFT_LAUNCH_CHECK(L);
    return cudaSuccess;
}

// S:1009-1012 for the speculative (column-resident) embed: an image over capacity leaves as its cover
__global__ void __launch_bounds__(256) passthrough_over_capacity(const uint8_t* __restrict__ cover, uint8_t* __restrict__ stego, size_t img_bytes,
                                                                 const uint64_t* __restrict__ usable, uint64_t nbits) {
    const int img = blockIdx.y;
    if (usable[img] >= nbits) return;
    const uint8_t* src = cover + (size_t)img * img_bytes;
    uint8_t* dst = stego + (size_t)img * img_bytes;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < img_bytes; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
cudaError_t launch_passthrough(const Launcher& L, const uint8_t* cover, uint8_t* stego, size_t img_bytes, int nimg, const uint64_t* usable, size_t nbits) {
    if (nimg == 0 || nbits == 0) return cudaSuccess;
    passthrough_over_capacity<<<dim3(64, (unsigned)nimg), 256, 0, L.stream>>>(cover, stego, img_bytes, usable, (uint64_t)nbits);
    TFFT_LAUNCH_CHECK(L);
    return cudaSuccess;
}

// packed frame bits (MSB first, the order of bytes_from_bits / bits_from_bytes S:447-459) -> one bit per byte, the form
// the embed kernels read; packed rows are `pstride` bytes apart
__global__ void __launch_bounds__(256) unpack_bits(const uint8_t* __restrict__ packed, size_t pstride, uint8_t* __restrict__ bits, size_t nbits) {
    const int img = blockIdx.y;
    const uint8_t* src = packed + (size_t)img * pstride;
    uint8_t* dst = bits + (size_t)img * nbits;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbits; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = (uint8_t)((src[i >> 3] >> (7 - (i & 7))) & 1u);
}
cudaError_t launch_unpack_bits(const Launcher& L, const uint8_t* packed, size_t pstride, uint8_t* bits, size_t nbits, int nimg) {
    if (nimg == 0 || nbits == 0) return cudaSuccess;
    const unsigned gx = (unsigned)std::min<size_t>((nbits + 255) / 256, 1024);
    unpack_bits<<<dim3(gx, (unsigned)nimg), 256, 0, L.stream>>>(packed, pstride, bits, nbits);
    TFFT_LAUNCH_CHECK(L);
    return cudaSuccess;
}

// --------------------------------------------------------------------------------------------
// Extract (read_bit_from_bin S:734-746 restated in full so ties behave like the reference,
// rep3/rep7 majority S:468-474 / S:501-508, MSB-first packing S:447-454).
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ double ang_diff(double a, double b) {
    const double PI = 3.14159265358979323846;
    double d = fmod(a - b + PI, 2 * PI);
    if (d < 0) d += 2 * PI;
    return fabs(d - PI);
}
__device__ __forceinline__ int read_bit(double2 z, double alpha, double jit) {
    const double th = atan2(z.y, z.x);
    return ang_diff(th, jit + alpha) <= ang_diff(th, jit - alpha) ? 1 : 0;
}
// compute_adaptive_alpha (S:704-710) on the read side (S:737-738): alpha scaled by |F| / median, clamped to [0.5, 2]
__device__ __forceinline__ double adaptive_alpha(double2 z, double alpha, const double* __restrict__ median, int img, uint32_t b) {
    if (!median) return alpha;
    const double mag = fmax(1e-12, hypot(z.x, z.y));
    return alpha * fmin(2.0, fmax(0.5, mag / fmax(1e-12, median[img * 3 + (int)(b >> 30)])));
}
__device__ __forceinline__ double2 load_bin(const double2* __restrict__ spec, int img, const SpecLayout& lay, uint32_t b) {
    const uint32_t lin = b & 0x3FFFFFFFu;
    return spec_load(spec + (size_t)(img * 3 + (int)(b >> 30)) * lay.plane_elems(), lay, (int)(lin / (uint32_t)lay.PW), (int)(lin % (uint32_t)lay.PW));
}

__global__ void __launch_bounds__(256) extract_raw(const double2* __restrict__ spec, SpecLayout P, const uint32_t* __restrict__ bins,
                                                   size_t nbins, const double* __restrict__ jitter, double alpha,
                                                   uint8_t* raw_bits, size_t raw_stride, const double* __restrict__ median) {
    const int img = blockIdx.y;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbins) return;
    const double2 z = load_bin(spec, img, P, bins[i]);
    raw_bits[(size_t)img * raw_stride + i] = (uint8_t)read_bit(z, adaptive_alpha(z, alpha, median, img, bins[i]), jitter ? jitter[i] : 0.0);
}

// one thread per decoded bit; a warp packs 32 decoded bits into 4 bytes with a ballot
__global__ void __launch_bounds__(256) extract_vote(const double2* __restrict__ spec, SpecLayout P, const uint32_t* __restrict__ bins,
                                                    size_t ndec, int rep, const double* __restrict__ jitter, double alpha,
                                                    uint8_t* out_bytes, size_t nbytes, const double* __restrict__ median) {
    const int img = blockIdx.y;
    const size_t d = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int bit = 0;
    if (d < ndec) {
        int s = 0;
        for (int j = 0; j < rep; j++) {
            const size_t i = d * rep + j;
            const double2 z = load_bin(spec, img, P, bins[i]);
            s += read_bit(z, adaptive_alpha(z, alpha, median, img, bins[i]), jitter ? jitter[i] : 0.0);
        }
        bit = (s >= rep / 2 + 1) ? 1 : 0;  // >=2 of 3, >=4 of 7 (rep 1: the bit itself)
    }
    const unsigned m = __brev(__ballot_sync(0xffffffffu, bit));  // lane 0 -> MSB (bytes_from_bits S:450)
    const int lane = threadIdx.x & 31;
    const size_t byte0 = (d - lane) / 8;
    if (lane < 4 && byte0 + lane < nbytes) out_bytes[(size_t)img * nbytes + byte0 + lane] = (uint8_t)(m >> (24 - 8 * lane));
}

// ---- the same two kernels reading the sign map the forward column pass of an extract left behind (PassArgs::signmap)
// fold = 1 (PassArgs::fold): an 8192-row plane was transformed as two 4096-row planes, row y of the spectrum is row y >> 1
// of map plane 2 * plane + (y & 1)
__device__ __forceinline__ int map_bit(const uint32_t* __restrict__ bm, int groups, int img, int PW, uint32_t b, int fold) {
    const uint32_t lin = b & 0x3FFFFFFFu;
    int y = (int)(lin / (uint32_t)PW);
    const int x = (int)(lin % (uint32_t)PW);
    size_t plane = (size_t)(img * 3 + (int)(b >> 30));
    if (fold) { plane = 2 * plane + (size_t)(y & 1); y >>= 1; }
    const uint32_t w = bm[((plane * groups + (x >> 1)) * 16 + (y & 15)) * 8 + (y >> 8)];
    return (int)((w >> ((((y >> 4) & 15) << 1) | (x & 1))) & 1u);
}
__global__ void __launch_bounds__(256) extract_raw_map(const uint32_t* __restrict__ bm, int groups, int PW, const uint32_t* __restrict__ bins,
                                                       size_t nbins, uint8_t* raw_bits, size_t raw_stride, int fold) {
    const int img = blockIdx.y;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nbins) return;
    raw_bits[(size_t)img * raw_stride + i] = (uint8_t)map_bit(bm, groups, img, PW, bins[i], fold);
}
__global__ void __launch_bounds__(256) extract_vote_map(const uint32_t* __restrict__ bm, int groups, int PW, const uint32_t* __restrict__ bins,
                                                        size_t ndec, int rep, uint8_t* out_bytes, size_t nbytes, int fold) {
    const int img = blockIdx.y;
    const size_t d = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int bit = 0;
    if (d < ndec) {
        int s = 0;
        for (int j = 0; j < rep; j++) s += map_bit(bm, groups, img, PW, bins[d * rep + j], fold);
        bit = (s >= rep / 2 + 1) ? 1 : 0;
    }
    const unsigned m = __brev(__ballot_sync(0xffffffffu, bit));
    const int lane = threadIdx.x & 31;
    const size_t byte0 = (d - lane) / 8;
    if (lane < 4 && byte0 + lane < nbytes) out_bytes[(size_t)img * nbytes + byte0 + lane] = (uint8_t)(m >> (24 - 8 * lane));
}
cudaError_t launch_extract_signmap(const Launcher& L, const uint32_t* signmap, int cols, int nimg, SpecLayout lay,
                                   const uint32_t* bins, size_t nbins, int rep, uint8_t* out_bytes, uint8_t* raw_bits, size_t raw_stride,
                                   int fold) {
    if (nimg == 0) return cudaSuccess;
    const int groups = cols / 2;
    if (raw_bits && nbins) {
        extract_raw_map<<<dim3((unsigned)((nbins + 255) / 256), (unsigned)nimg), 256, 0, L.stream>>>(signmap, groups, lay.PW, bins, nbins, raw_bits, raw_stride ? raw_stride : nbins, fold);
        TFFT_LAUNCH_CHECK(L);
    }
    const size_t ndec = nbins / (size_t)rep;
    const size_t nbytes = (ndec + 7) / 8;
    if (out_bytes && nbytes) {
        extract_vote_map<<<dim3((unsigned)((ndec + 255) / 256), (unsigned)nimg), 256, 0, L.stream>>>(signmap, groups, lay.PW, bins, ndec, rep, out_bytes, nbytes, fold);
        TFFT_LAUNCH_CHECK(L);
    }
    return cudaSuccess;
}

// Window of the workspace a bin list touches: out[0] = 1 + largest stored row, out[1] = 1 + largest stored column,
// out[2] = 1 when a bin is read through its mirror (half layout: bins right of the Nyquist column, spec_load).
__global__ void __launch_bounds__(256) bins_window(const uint32_t* __restrict__ bins, size_t nbins, SpecLayout lay, unsigned* out) {
    unsigned ry = 0, rx = 0, mir = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbins; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t lin = bins[i] & 0x3FFFFFFFu;
        int y = (int)(lin / (uint32_t)lay.PW), x = (int)(lin % (uint32_t)lay.PW);
        if (lay.half && x > (lay.PW >> 1)) { y = (lay.PH - y) & (lay.PH - 1); x = lay.PW - x; mir = 1; }
        ry = max(ry, (unsigned)y + 1u);
        rx = max(rx, (unsigned)x + 1u);
    }
    ry = __reduce_max_sync(0xffffffffu, ry);
    rx = __reduce_max_sync(0xffffffffu, rx);
    mir = __reduce_max_sync(0xffffffffu, mir);
    if ((threadIdx.x & 31) == 0) { atomicMax(out, ry); atomicMax(out + 1, rx); if (mir) atomicMax(out + 2, 1u); }
}
cudaError_t launch_bins_window(const Launcher& L, const uint32_t* bins, size_t nbins, SpecLayout lay, unsigned* d_out2) {
    cudaError_t e = cudaMemsetAsync(d_out2, 0, 3 * sizeof(unsigned), L.stream);
    if (e != cudaSuccess || nbins == 0) return e;
    const unsigned grid = (unsigned)std::min<size_t>((nbins + 255) / 256, 4 * (size_t)L.sm_count);
    bins_window<<<grid, 256, 0, L.stream>>>(bins, nbins, lay, d_out2);
    TFFT_LAUNCH_CHECK(L);
    return cudaSuccess;
}

cudaError_t launch_extract(const Launcher& L, const double2* spec, int nimg, SpecLayout P,
                           const uint32_t* bins, size_t nbins, int rep, const double* jitter, double alpha,
                           uint8_t* out_bytes, uint8_t* raw_bits, size_t raw_stride, const double* adaptive_median) {
    if (nimg == 0) return cudaSuccess;
    if (raw_bits && nbins) {
        extract_raw<<<dim3((unsigned)((nbins + 255) / 256), (unsigned)nimg), 256, 0, L.stream>>>(spec, P, bins, nbins, jitter, alpha, raw_bits, raw_stride ? raw_stride : nbins, adaptive_median);
        TFFT_LAUNCH_CHECK(L);
    }
    const size_t ndec = nbins / (size_t)rep;
    const size_t nbytes = (ndec + 7) / 8;
    if (out_bytes && nbytes) {
        extract_vote<<<dim3((unsigned)((ndec + 255) / 256), (unsigned)nimg), 256, 0, L.stream>>>(spec, P, bins, ndec, rep, jitter, alpha, out_bytes, nbytes, adaptive_median);
        TFFT_LAUNCH_CHECK(L);
    }
    return cudaSuccess;
}


// ---- unfused conversions for the large-size path (to_planes_u8 S:383 + apply_center S:392 + pad_to_fft S:393;
//      ifft_crop S:399 + apply_center S:1102 + from_planes_u8 S:387)
__global__ void __launch_bounds__(256) u8_to_planes(const uint8_t* __restrict__ img, double2* __restrict__ spec, int W, int H, int PW, int PH,
                                                    int center, long long total) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int x = (int)(t % PW);
    const long long u = t / PW;
    const int y = (int)(u % PH);
    const long long ip = u / PH;  // image*3 + plane
    double v = 0.0;
    if (x < W && y < H) {
        v = (double)img[(((size_t)(ip / 3) * H + y) * W + x) * 3 + (ip % 3)];
        if (center && ((x + y) & 1)) v = -v;
    }
    spec[t] = make_double2(v, 0.0);
}
__global__ void __launch_bounds__(256) planes_to_u8(const double2* __restrict__ spec, uint8_t* __restrict__ img, int W, int H, int PW, int PH,
                                                    int center, long long total) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per output byte
    if (t >= total) return;
    const int ch = (int)(t % 3);
    const long long px = t / 3;
    const int x = (int)(px % W);
    const long long u = px / W;
    const int y = (int)(u % H);
    const long long im = u / H;
    double v = spec[(((size_t)im * 3 + ch) * PH + y) * PW + x].x;
    if (center && ((x + y) & 1)) v = -v;
    img[t] = clamp8(v);
}
cudaError_t launch_u8_to_planes(const Launcher& L, const uint8_t* img, double2* spec, int nimg, int W, int H, int PW, int PH, int center) {
    const long long total = (long long)nimg * 3 * PH * PW;
    u8_to_planes<<<(unsigned)((total + 255) / 256), 256, 0, L.stream>>>(img, spec, W, H, PW, PH, center, total);
    TFFT_LAUNCH_CHECK(L);
    return cudaSuccess;
}
cudaError_t launch_planes_to_u8(const Launcher& L, const double2* spec, uint8_t* img, int nimg, int W, int H, int PW, int PH, int center) {
    const long long total = (long long)nimg * H * W * 3;
    planes_to_u8<<<(unsigned)((total + 255) / 256), 256, 0, L.stream>>>(spec, img, W, H, PW, PH, center, total);
    TFFT_LAUNCH_CHECK(L);
    return cudaSuccess;
}

// full[y][x] from a half-spectrum workspace (parity hook only)
__global__ void expand_half(const double2* __restrict__ hs, double2* __restrict__ fs, SpecLayout lay) {
    const int ip = blockIdx.y;
    const size_t P = (size_t)lay.PH * lay.PW;
    const double2* pl = hs + (size_t)ip * lay.plane_elems();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (size_t)gridDim.x * blockDim.x)
        fs[(size_t)ip * P + i] = spec_load(pl, lay, (int)(i / lay.PW), (int)(i % lay.PW));
}
cudaError_t launch_expand_half(const Launcher& L, const double2* half_spec, double2* full_spec, int nplanes, SpecLayout lay) {
    expand_half<<<dim3(1024, (unsigned)nplanes), 256, 0, L.stream>>>(half_spec, full_spec, lay);
    TFFT_LAUNCH_CHECK(L);
    return cudaSuccess;
}

}  // namespace tfft
