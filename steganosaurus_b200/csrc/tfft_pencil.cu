// tfft_pencil.cu -- the fast FFT passes ("pencil" kernels) for N = 512 .. 4096 (and 8192-pixel rows).
//
// A pencil is one row or one column of a padded plane.  N = R1 * 16 * 16 with R1 = N/256 in
// {2,4,8,16}: three Cooley-Tukey stages (decimation in frequency) whose butterflies live in
// registers (16 complex doubles per thread); data moves between stages through shared memory:
//
//   landing buffer L (rows: cp.async 16 B; columns: TMA box loads), dense 16 B entries
//   stage 1 (radix R1, stride 256)  in place in L                                   [1 barrier]
//   stage 2 (radix 16, stride 16)   reads L, then L is free: the NEXT pencil's loads are issued here
//   exchange 2                      rows: padded / XOR-swizzled buffer; 4096-point columns: a warp-private
//                                   4 KB transpose (after stage 1 the sixteen 256-point problems are independent)
//   stage 3 (radix 16, stride 1)    results go to the epilogue: Hermitian split + STG (forward rows), u8 quantiser
//                                   (inverse rows), staging + TMA box stores (columns)
//
// Kernels:
//   pencil_u8_fwd_r2c / pencil_u8_inv_c2r   fused u8 RGB <-> half-spectrum rows; two image rows share one complex
//                                           transform, or (WIDE) one 8192-pixel row is packed into one; FOLD (forward,
//                                           8192-row planes of an extract): rows y and y + 4096 leave as the two inputs
//                                           of two 4096-point column transforms (first radix-2 step of the column pass)
//   pencil_col_embed_w                      column-resident embed: forward columns, |F|^2 for the median, phase write on
//                                           the registers that own the bins, inverse columns -- one residency per pair
//   pencil_col_tma_w                        4096-point column pairs on the TMA engine, mbarrier hand-offs, zero-block skipping;
//                                           unfused embed forward: also the median sample and a float copy of |F|^2;
//                                           extract forward (SIGN): no spectrum, one read bit per element
//   pencil_col_tma, pencil_c2c, pencil_u8_fwd/inv   other sizes / full-spectrum variants / LSU fallback
//
// A "unit" is the set of threads that owns one buffer set and works through its own stream of pencils (persistent,
// static round-robin); row CTAs host two units that only synchronise among themselves (named barriers).
// What bounds them (profiles/r2_experiments.txt): the FP64 pipe (~730 instructions per thread and
// transform) and the shared-memory pipe (~100 16-byte accesses) carry about the same load and take turns, because the
// warps of a CTA sit in the same phase; HBM is 35 % busy.  What pays is taking work off either pipe.
// Reference citations S:n = steganosaurus/src/steganosaur.cpp line n.
#include "tfft_kernels.cuh"

#include <cstdlib>
#include <cstring>

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

namespace tfft {
namespace pk {

// ------------------------------------------------------------------------------------------
// complex helpers (double2 = re, im)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
// multiply by S*i
template <int S>
__device__ __forceinline__ double2 mul_si(double2 a) {
    return S > 0 ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}
// multiply by the constant (wr, S*wi)
template <int S>
__device__ __forceinline__ double2 cmulc(double2 a, double wr, double wi) {
    const double si = S > 0 ? wi : -wi;
    return make_double2(fma(a.x, wr, -a.y * si), fma(a.x, si, a.y * wr));
}

constexpr long long STAGGER_CYCLES = 700;
constexpr double C8 = 0.70710678118654752440;   // cos(pi/4)
constexpr double C16 = 0.92387953251128673848;  // cos(pi/8)
constexpr double S16 = 0.38268343236508978178;  // sin(pi/8)

// X[k] = sum_n x[n] (S*i)^{nk}, in place, natural order
template <int S>
__device__ __forceinline__ void dft4(double2& a0, double2& a1, double2& a2, double2& a3) {
    const double2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_si<S>(csub(a1, a3));
    a0 = cadd(t0, t2); a2 = csub(t0, t2); a1 = cadd(t1, t3); a3 = csub(t1, t3);
}

// a (1 + i tau)
__device__ __forceinline__ double2 rot_t(double2 a, double tau) { return make_double2(fma(-tau, a.y, a.x), fma(tau, a.x, a.y)); }
// dft4 of (a0, c1 u1, c2 u2, c3 u3) with r = c3 / c1: the scale factors of twiddled inputs w = c (1 + i tau), u = a (1 + i tau),
// ride on the butterfly's own additions as FMAs -- 22 instructions for three twiddles and the butterfly instead of 28
template <int S>
__device__ __forceinline__ void dft4_tw(double2& a0, double2& a1, double2& a2, double2& a3, double c1, double2 u1, double c2, double2 u2,
                                        double r, double2 u3) {
    const double2 t0 = make_double2(fma(c2, u2.x, a0.x), fma(c2, u2.y, a0.y));
    const double2 t1 = make_double2(fma(-c2, u2.x, a0.x), fma(-c2, u2.y, a0.y));
    const double2 p = make_double2(fma(r, u3.x, u1.x), fma(r, u3.y, u1.y));
    const double2 q = mul_si<S>(make_double2(fma(-r, u3.x, u1.x), fma(-r, u3.y, u1.y)));
    a0 = make_double2(fma(c1, p.x, t0.x), fma(c1, p.y, t0.y));
    a2 = make_double2(fma(-c1, p.x, t0.x), fma(-c1, p.y, t0.y));
    a1 = make_double2(fma(c1, q.x, t1.x), fma(c1, q.y, t1.y));
    a3 = make_double2(fma(-c1, q.x, t1.x), fma(-c1, q.y, t1.y));
}

// In-register DFT of R points with w = exp(S*2*pi*i/R).  Output X[k] is left at x[oidx<R>(k)].
template <int R>
__device__ __host__ constexpr int oidx(int k) {
    return R == 16 ? 4 * (k & 3) + (k >> 2) : R == 8 ? 2 * (k & 3) + (k >> 2) : k;
}

// PLAIN (R = 16): the textbook second layer instead of the folded one -- same results within rounding; the column-resident
// embed kernel keeps it (the folded form costs it a spilled register and 6 % of its time, profiles/r2_experiments.txt)
template <int S, int R, bool PLAIN = false>
__device__ __forceinline__ void dft(double2* x) {
#ifdef TFFT_EXP_NO_FP64  // timing experiment only (results are garbage): data movement and barriers without the butterflies
    return;
#endif
    if constexpr (R == 2) {
        const double2 a = x[0], b = x[1];
        x[0] = cadd(a, b); x[1] = csub(a, b);
    } else if constexpr (R == 4) {
        dft4<S>(x[0], x[1], x[2], x[3]);
    } else if constexpr (R == 8) {
        // n = 2 n1 + n2, k = k1 + 4 k2
        dft4<S>(x[0], x[2], x[4], x[6]);
        dft4<S>(x[1], x[3], x[5], x[7]);  // y[n2][k1] at x[2 k1 + n2]
        x[3] = cmulc<S>(x[3], C8, C8);    // w8^1
        x[5] = mul_si<S>(x[5]);           // w8^2
        x[7] = cmulc<S>(x[7], -C8, C8);   // w8^3
#pragma unroll
        for (int k1 = 0; k1 < 4; k1++) {
            const double2 a = x[2 * k1], b = x[2 * k1 + 1];
            x[2 * k1] = cadd(a, b); x[2 * k1 + 1] = csub(a, b);
        }
    } else {
        static_assert(R == 16, "radix");
        // n = 4 n1 + n2, k = k1 + 4 k2
#pragma unroll
        for (int n2 = 0; n2 < 4; n2++) dft4<S>(x[n2], x[4 + n2], x[8 + n2], x[12 + n2]);  // y[n2][k1] at x[4 k1 + n2]
#ifdef TFFT_DFT16_PLAIN
        constexpr bool plain = true;
#else
        constexpr bool plain = PLAIN;
#endif
        if constexpr (plain) {  // the textbook form: nine constant twiddles w16^{n2 k1}, then four plain 4-point butterflies
        x[5] = cmulc<S>(x[5], C16, S16);    // 1
        x[6] = cmulc<S>(x[6], C8, C8);      // 2
        x[7] = cmulc<S>(x[7], S16, C16);    // 3
        x[9] = cmulc<S>(x[9], C8, C8);      // 2
        x[10] = mul_si<S>(x[10]);           // 4
        x[11] = cmulc<S>(x[11], -C8, C8);   // 6
        x[13] = cmulc<S>(x[13], S16, C16);  // 3
        x[14] = cmulc<S>(x[14], -C8, C8);   // 6
        x[15] = cmulc<S>(x[15], -C16, -S16);  // 9
#pragma unroll
        for (int k1 = 0; k1 < 4; k1++) dft4<S>(x[4 * k1], x[4 * k1 + 1], x[4 * k1 + 2], x[4 * k1 + 3]);
        } else {
        // the constant twiddles w16^{n2 k1} = c (1 + i tau) folded into the second layer (dft4_tw): 16 instructions fewer
        constexpr double T1 = 0.41421356237309504880;  // tan(pi/8) = S16 / C16
        constexpr double T3 = 2.41421356237309504880;  // cot(pi/8) = C16 / S16
        dft4<S>(x[0], x[1], x[2], x[3]);
        dft4_tw<S>(x[4], x[5], x[6], x[7], C16, rot_t(x[5], S * T1), C8, rot_t(x[6], S * 1.0), T1, rot_t(x[7], S * T3));            // w^1, w^2, w^3
        dft4_tw<S>(x[8], x[9], x[10], x[11], C8, rot_t(x[9], S * 1.0), 1.0, mul_si<S>(x[10]), -1.0, rot_t(x[11], -S * 1.0));         // w^2, w^4, w^6
        dft4_tw<S>(x[12], x[13], x[14], x[15], S16, rot_t(x[13], S * T3), -C8, rot_t(x[14], -S * 1.0), -T3, rot_t(x[15], S * T1));   // w^3, w^6, w^9
        }
    }
}

// X[k] *= w1^k for k = 1..R-1 (outputs addressed through oidx)
template <int R>
__device__ __forceinline__ void twiddle(double2* x, double2 w1) {
#ifdef TFFT_EXP_NO_FP64
    x[0].x += w1.x;  // keep the table load alive
    return;
#endif
    double2 w = w1;
#pragma unroll
    for (int k = 1; k < R; k++) {
        x[oidx<R>(k)] = cmul(x[oidx<R>(k)], w);
        if (k + 1 < R) w = cmul(w, w1);
    }
}

// ------------------------------------------------------------------------------------------
// async copy + named barrier primitives
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void unit_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- TMA (cp.async.bulk.tensor) + mbarrier primitives ------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TFFT_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TFFT_MBAR_DONE;\n"
        "bra TFFT_MBAR_WAIT;\n"
        "TFFT_MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                 ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];\n"
                 ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ uint8_t clamp8(double v) {  // from_planes_u8 S:389
    double r = round(v);
    r = fmax(0.0, fmin(255.0, r));
    return (uint8_t)r;
}

// ------------------------------------------------------------------------------------------
// compile-time geometry of one unit
// ------------------------------------------------------------------------------------------
template <int LOG2N, int VEC>
struct Geo {
    static constexpr int N = 1 << LOG2N;
    static constexpr int R1 = N / 256;          // first-stage radix
    static constexpr int TP = N / 16;           // threads per pencil
    static constexpr int UT = TP * VEC;         // threads per unit
    static constexpr int J1 = 16 / R1;          // stage-1 butterflies per thread
    static constexpr size_t L_BYTES = (size_t)N * VEC * 16;
    static constexpr size_t X_BYTES = (size_t)N * VEC * 8;
};

enum Mode { M_C2C_ROW = 0, M_C2C_COL = 1, M_U8_FWD = 2, M_U8_INV = 3 };

// XOR swizzle of exchange-2 positions p = k1*256 + k2*16 + low4: low4 ^= (k1 + R1*k2) & 15
template <int R1>
__device__ __forceinline__ int swz(int k1, int k2, int low) { return (k1 << 8) | (k2 << 4) | (low ^ ((k1 + R1 * k2) & 15)); }

// Exchange-2 position used by the fused u8 row kernels.  N = 4096 (R1 = 16): every 256 entries are padded by
// one (row stride 257), which keeps the stride-16 writers and the stride-256 readers conflict-free AND gives
// every access a compile-time offset from one per-thread base (the XOR swizzle costs two integer
// instructions per access; these kernels are issue-bound).  Smaller N keep the XOR swizzle.
template <int R1>
__device__ __forceinline__ int xpos(int k1, int k2, int low) {
    if constexpr (R1 == 16) return k1 * 257 + (k2 << 4) + low;
    else return swz<R1>(k1, k2, low);
}
template <int LOG2N>
constexpr int xpad() { return LOG2N == 12 ? 16 : 0; }  // extra entries of a padded exchange buffer

// exact u8 -> double without a conversion instruction: 2^52 + v has v in its low mantissa word
__device__ __forceinline__ double u8_to_double(unsigned v) {
#ifdef TFFT_U8_I2F
    return (double)v;
#else
    return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
#endif
}

// exact u8 -> double of HALF the value: 2^51 + v/2 has v in its low mantissa word.  The forward half-spectrum row kernel
// feeds its transform with halved pixels, so the Hermitian split's factors 1/2 are already in (a power of two commutes
// with every rounding on the way: bit-identical results, 32 multiplications fewer per thread and plane)
__device__ __forceinline__ double u8_to_half_double(unsigned v) {
    return __hiloint2double(0x43200000, (int)v) - 2251799813685248.0;
}

// ---- stage 1: radix R1, stride 256, in place in L (or from the u8 row for M_U8_FWD) --------
// NZ: input blocks n >= NZ (rows n*256 ..) are known to be zero: not read, and their butterflies fold away
// TW_SHIFT: tw[m << TW_SHIFT] = w_N^m (the global table holds w_16384^k; a shared-memory copy of w_N^m, m < 256, has shift 0)
// TW_READY: the table already holds the twiddles of this direction (conjugated for the inverse)
template <int S, int LOG2N, int VEC, bool FROM_U8, int NZ = (1 << LOG2N) / 256, int TW_SHIFT = TW_LOG2 - LOG2N, bool TW_READY = false,
          bool PLAIN = false>
__device__ __forceinline__ void stage1(double2* L, int tt, int c, const double2* __restrict__ tw,
                                       const uint8_t* urow, int W, int ch, int negmask /*center: parity of y, or -1*/) {
    using G = Geo<LOG2N, VEC>;
#pragma unroll
    for (int j = 0; j < G::J1; j++) {
        const int m = tt + j * G::TP;  // 0..255
        double2 x[G::R1];
#pragma unroll
        for (int n = 0; n < G::R1; n++) {
            if constexpr (FROM_U8) {
                const int xc = n * 256 + m;
                double v = 0.0;
                if (xc < W) {
                    v = (double)urow[xc * 3 + ch];
                    if (negmask >= 0 && ((xc + negmask) & 1)) v = -v;  // apply_center S:392
                }
                x[n] = make_double2(v, 0.0);
            } else {
                x[n] = n < NZ ? L[(n * 256 + m) * VEC + c] : make_double2(0.0, 0.0);
            }
        }
        dft<S, G::R1, PLAIN>(x);
        double2 w1 = tw[(size_t)m << TW_SHIFT];
        if (S < 0 && !TW_READY) w1.y = -w1.y;
        twiddle<G::R1>(x, w1);
#pragma unroll
        for (int k = 0; k < G::R1; k++) L[(k * 256 + m) * VEC + c] = x[oidx<G::R1>(k)];
    }
}

// Per-thread twiddle bases w^m (stage 1) and w256^(m & 15) (stage 2).  They are re-fetched from the
// L2-resident table for every pencil: the kernels sit at the 128-register cap, and keeping these
// constants live across the item loop spilled ~200 B/thread (measured: column pass 3x slower).
template <int S, int LOG2N, int VEC>
struct ThreadTw {
    const double2* tw;
    int tt;
    __device__ __forceinline__ void load(const double2* __restrict__ t, int tt_) { tw = t; tt = tt_; }
    __device__ __forceinline__ double2 s2v() const {
        double2 w = tw[(size_t)(tt & 15) << (TW_LOG2 - 8)];
        if (S < 0) w.y = -w.y;
        return w;
    }
};

// ---- stage 2 load: radix 16, stride 16 -------------------------------------------------------
template <int LOG2N, int VEC>
__device__ __forceinline__ void stage2_load(const double2* L, int tt, int c, double2* x) {
    const int k1 = tt >> 4, m = tt & 15;
#pragma unroll
    for (int n = 0; n < 16; n++) x[n] = L[(k1 * 256 + n * 16 + m) * VEC + c];
}

// ---- stage 2 compute + exchange 2 through X + stage 3 compute -------------------------------
// on return x[oidx<16>(k3)] holds output element tt + TP*k3
template <int S, int LOG2N, int VEC>
__device__ __forceinline__ void stage23(double* X, int tt, int c, double2 w1, double2* x, int bar_id) {
    using G = Geo<LOG2N, VEC>;
    const int k1 = tt >> 4, m = tt & 15;
    dft<S, 16>(x);
    twiddle<16>(x, w1);
    // thread of stage 3: tt = k1' + R1*k2'
    const int k1r = tt & (G::R1 - 1), k2r = tt / G::R1;
    double2 z[16];
#pragma unroll
    for (int k2 = 0; k2 < 16; k2++) X[swz<G::R1>(k1, k2, m) * VEC + c] = x[oidx<16>(k2)].x;
    unit_bar(bar_id, G::UT);
#pragma unroll
    for (int n = 0; n < 16; n++) z[n].x = X[swz<G::R1>(k1r, k2r, n) * VEC + c];
    unit_bar(bar_id, G::UT);
#pragma unroll
    for (int k2 = 0; k2 < 16; k2++) X[swz<G::R1>(k1, k2, m) * VEC + c] = x[oidx<16>(k2)].y;
    unit_bar(bar_id, G::UT);
#pragma unroll
    for (int n = 0; n < 16; n++) z[n].y = X[swz<G::R1>(k1r, k2r, n) * VEC + c];
    dft<S, 16>(z);
#pragma unroll
    for (int n = 0; n < 16; n++) x[n] = z[n];
}

// Variant for passes whose L buffer is not a landing buffer (u8 forward rows): exchange 2 runs in
// place in L with 16 B entries and the same XOR swizzle -> one barrier instead of three.
template <int S, int LOG2N, bool PADDED = false>
__device__ __forceinline__ void stage23_inplace(double2* L, int tt, double2 w1, double2* x, int bar_id) {
    using G = Geo<LOG2N, 1>;
    const int k1 = tt >> 4, m = tt & 15;
    dft<S, 16>(x);
    twiddle<16>(x, w1);
    const int k1r = tt & (G::R1 - 1), k2r = tt / G::R1;
    [[maybe_unused]] double2* Lw = L + xpos<G::R1>(k1, 0, m);   // R1 == 16: the 16 accesses are Lw[16 k2] / Lr[n] (immediate offsets)
    [[maybe_unused]] const double2* Lr = L + xpos<G::R1>(k1r, k2r, 0);
#pragma unroll
    for (int k2 = 0; k2 < 16; k2++) {
        if constexpr (PADDED && G::R1 == 16) Lw[k2 << 4] = x[oidx<16>(k2)];
        else L[(PADDED ? xpos<G::R1>(k1, k2, m) : swz<G::R1>(k1, k2, m))] = x[oidx<16>(k2)];
    }
    unit_bar(bar_id, G::UT);
#pragma unroll
    for (int n = 0; n < 16; n++) {
        if constexpr (PADDED && G::R1 == 16) x[n] = Lr[n];
        else x[n] = L[(PADDED ? xpos<G::R1>(k1r, k2r, n) : swz<G::R1>(k1r, k2r, n))];
    }
    dft<S, 16>(x);
}

// ------------------------------------------------------------------------------------------
// C2C pass (rows: VEC = 1, columns: VEC columns per unit)
// ------------------------------------------------------------------------------------------
struct C2CArgs {
    double2* spec;
    const double2* tw;
    long long nitems;   // pencil groups: rows: nplanes*PH ; columns: nplanes*PW/VEC
    int PW, PH;
    int in_rows, out_rows;
};

template <int S, int LOG2N, int VEC, int MODE, int UNITS>
__global__ void __launch_bounds__(Geo<LOG2N, VEC>::UT* UNITS, 1) pencil_c2c(C2CArgs a) {
    using G = Geo<LOG2N, VEC>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int unit = threadIdx.x / G::UT, ut = threadIdx.x % G::UT;
    const int c = ut % VEC, tt = ut / VEC;
    double2* L = (double2*)(smem_raw + (size_t)unit * (G::L_BYTES + G::X_BYTES));
    double* X = (double*)((unsigned char*)L + G::L_BYTES);
    const int bar_id = 1 + unit;
    const long long stride = (long long)gridDim.x * UNITS;
    long long item = (long long)blockIdx.x * UNITS + unit;
    const double scale = S < 0 ? 1.0 / (double)G::N : 1.0;  // S:357
    ThreadTw<S, LOG2N, VEC> ttw;
    ttw.load(a.tw, tt);

    // element e = k*VEC + c of item -> global pointer; returns validity (zero rows are not read)
    auto gptr = [&](long long it, int k, int cc, bool& valid) -> double2* {
        if constexpr (MODE == M_C2C_ROW) {
            const long long ip = it / a.PH;
            const int y = (int)(it % a.PH);
            valid = y < a.in_rows;
            return a.spec + ((size_t)ip * a.PH + y) * a.PW + k;
        } else {
            const int gpp = a.PW / VEC;
            const long long ip = it / gpp;
            const int x0 = (int)(it % gpp) * VEC;
            valid = k < a.in_rows;
            return a.spec + ((size_t)ip * a.PH + k) * a.PW + x0 + cc;
        }
    };
    auto issue_loads = [&](long long it) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int e = ut + i * G::UT;
            bool valid;
            double2* g = gptr(it, e / VEC, e % VEC, valid);
            cp_async16(&L[e], valid ? (const void*)g : (const void*)a.spec, valid ? 16 : 0);
        }
        cp_async_commit();
    };

    if (unit & 1) {  // de-phase odd units once: FP64-heavy and LSU-heavy phases of neighbours interleave
        const long long t0 = clock64();
        while (clock64() - t0 < STAGGER_CYCLES) {}
    }
    if (item < a.nitems) issue_loads(item);
    for (; item < a.nitems; item += stride) {
        cp_async_wait_all();
        unit_bar(bar_id, G::UT);
        stage1<S, LOG2N, VEC, false>(L, tt, c, a.tw, nullptr, 0, 0, -1);
        unit_bar(bar_id, G::UT);
        double2 x[16];
        stage2_load<LOG2N, VEC>(L, tt, c, x);
        unit_bar(bar_id, G::UT);  // L is free
        if (item + stride < a.nitems) issue_loads(item + stride);
        stage23<S, LOG2N, VEC>(X, tt, c, ttw.s2v(), x, bar_id);
#pragma unroll
        for (int k3 = 0; k3 < 16; k3++) {
            const int k = tt + G::TP * k3;
            bool valid;
            double2* g = gptr(item, k, c, valid);
            double2 v = x[oidx<16>(k3)];
            v.x *= scale; v.y *= scale;
            if (MODE == M_C2C_ROW || k < a.out_rows) *g = v;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Column pass on the TMA engine.  One unit per CTA works on VEC adjacent columns
// (N*VEC = 8192 elements: 128 KB landing buffer + 64 KB exchange/staging buffer, 512 threads).
// Global traffic never touches the LSU: box loads {VEC complex x 256 rows} land the columns densely
// as [row][column]; results are staged in X half a pencil at a time and leave through box stores.
// Zero-structure: the load map only spans in_rows rows (TMA zero-fills the rest without reading),
// the store map only spans out_rows rows (TMA clips the rest).
// ------------------------------------------------------------------------------------------
struct ColTmaArgs {
    const double2* tw;
    long long nitems;      // nplanes * PW / VEC
    int groups_per_plane;  // PW / VEC
    // pencil_col_tma_w, forward only: stratified sample of q = |F|^2 for the median bracket (null: off).  Column pairs
    // g < sample_groups each contribute 256 values: thread (k1, m) of column g & 1 picks one of its 16 row blocks by hash.
    unsigned long long* sample_q;
    unsigned sample_stride;   // entries per plane
    int sample_groups;        // column pairs that hold interior columns (PW_full / 4)
    // with sample_q (optional): q of every element as a float; as float4 [plane][g][k1][j][lane = 2 m + c] holding rows
    // k1 + 16 m + 256 (4 j + 0..3) of column 2 g + c: four fully coalesced 512-byte warp stores per pair
    float* q32;
    // pencil_col_tma_w<.., SIGN = true>: no spectrum is stored, only the bit read_bit_from_bin (S:734-746) would read at
    // every element, packed as signmap[((plane * map_groups + g) * 16 + k1) * 8 + k3] bit (2 m + c) for row
    // k1 + 16 m + 256 k3 of column 2 g + c
    uint32_t* signmap;
    int map_groups;           // column pairs per plane in the map (PW / 2)
    double alpha;
};

// read_bit_from_bin (S:734-746) in full: atan2, circular distances to +alpha and -alpha, ties read 1
__device__ __noinline__ int read_bit_full(double re, double im, double alpha) {
    const double PI = 3.14159265358979323846;
    const double th = atan2(im, re);
    double dp = fmod(th - alpha + PI, 2 * PI);
    if (dp < 0) dp += 2 * PI;
    double dn = fmod(th + alpha + PI, 2 * PI);
    if (dn < 0) dn += 2 * PI;
    return fabs(dp - PI) <= fabs(dn - PI) ? 1 : 0;
}
// For alpha in [1e-6, pi - 1e-6] the two distances differ by 2 min(|th|, alpha, pi - |th|, pi - alpha), so away from
// the real axis the verdict is simply the sign of the imaginary part; within 1e-9 of it the full formula decides.
__device__ __forceinline__ int read_bit_sign(double2 z, double alpha) {
    if (fabs(z.y) > 1e-9 * fabs(z.x)) return z.y > 0.0 ? 1 : 0;
    return read_bit_full(z.x, z.y, alpha);
}

template <int S, int LOG2N, int VEC>
__global__ void __launch_bounds__(512, 1) pencil_col_tma(const __grid_constant__ CUtensorMap in_map,
                                                         const __grid_constant__ CUtensorMap out_map, ColTmaArgs a) {
    using G = Geo<LOG2N, VEC>;
    static_assert(G::UT == 512, "one 512-thread unit per CTA");
    constexpr int BOX_ROWS = 256;
    constexpr int NBOX = G::N / BOX_ROWS;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t full_bar;
    double2* L = (double2*)smem_raw;
    double* X = (double*)(smem_raw + G::L_BYTES);
    double2* STG = (double2*)X;  // staging view of X: (N/2) rows x VEC x 16 B
    const int tid = threadIdx.x, c = tid % VEC, tt = tid / VEC;
    const double scale = S < 0 ? 1.0 / (double)G::N : 1.0;  // S:357
    ThreadTw<S, LOG2N, VEC> ttw;
    ttw.load(a.tw, tt);

    if (tid == 0) {
        mbar_init(&full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    auto issue_load = [&](long long it) {
        const int plane = (int)(it / a.groups_per_plane), g = (int)(it % a.groups_per_plane);
        mbar_expect_tx(&full_bar, (unsigned)G::L_BYTES);
#pragma unroll 1
        for (int j = 0; j < NBOX; j++) tma_load_3d(L + (size_t)j * BOX_ROWS * VEC, &in_map, &full_bar, g * VEC * 2, j * BOX_ROWS, plane);
    };

    const long long stride = gridDim.x;
    long long item = blockIdx.x;
    if (tid == 0 && item < a.nitems) issue_load(item);
    unsigned parity = 0;
    for (; item < a.nitems; item += stride) {
        mbar_wait(&full_bar, parity);
        parity ^= 1;
        stage1<S, LOG2N, VEC, false>(L, tt, c, a.tw, nullptr, 0, 0, -1);
        __syncthreads();
        double2 x[16];
        stage2_load<LOG2N, VEC>(L, tt, c, x);
        if (tid == 0) tma_wait_read_all();  // the previous item's last box store has finished reading X
        __syncthreads();                    // L is free, X is free
        if (tid == 0 && item + stride < a.nitems) {
            fence_async_proxy();
            issue_load(item + stride);
        }
        stage23<S, LOG2N, VEC>(X, tt, c, ttw.s2v(), x, 0);
        const int plane = (int)(item / a.groups_per_plane), g = (int)(item % a.groups_per_plane);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            __syncthreads();  // h=0: every thread is past its last read of X; h=1: first half has been read out
#pragma unroll
            for (int k3 = 8 * h; k3 < 8 * h + 8; k3++) {
                double2 v = x[oidx<16>(k3)];
                v.x *= scale; v.y *= scale;
                STG[(size_t)(tt + G::TP * (k3 - 8 * h)) * VEC + c] = v;
            }
            fence_async_proxy();
            __syncthreads();
            if (tid == 0) {
#pragma unroll 1
                for (int j = 0; j < NBOX / 2; j++)
                    tma_store_3d(&out_map, STG + (size_t)j * BOX_ROWS * VEC, g * VEC * 2, h * (G::N / 2) + j * BOX_ROWS, plane);
                tma_commit();
                if (h == 0) tma_wait_read_all();
            }
        }
    }
    if (tid == 0) tma_wait_all();
}

// ------------------------------------------------------------------------------------------
// Column pass, N = 4096, two columns per CTA, warp-local second exchange.
// After stage 1 the transform splits into sixteen independent 256-point problems (one per k1), and
// with two lane-interleaved columns the 16 threads x 2 columns of one k1 are exactly one warp.  So
// exchange 2 needs no block barrier: every warp transposes inside its private 4 KB slice of X
// (re then im, XOR-swizzled) with __syncwarp, keeps (k1, k2) for stage 3 and stages its results in the
// same slice as [k3][k2][column].  The row order that leaves behind (row = k1 + 16 k2 + 256 k3) is undone
// by the store itself: a rank-5 tensor map (column pair, k2, k3, k1, plane) lets ONE TMA box store per
// half pencil scatter the [k1][k3][k2] staging image to its rows.  Block barriers per pair: 5 (was 9).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];\n"
                 ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

// NZ : only the first NZ 256-row blocks of the input are non-zero (forward pass of a padded image): the
//      other boxes are neither loaded nor read, and stage 1 folds the zero butterflies away.
// K3N: only the first K3N 256-row blocks of the output are kept (inverse pass before the crop): the
//      other results are neither computed (dead code) nor staged.
// SIGN: forward pass of an extract without jitter -- nothing but the read bit of every element is kept (K3N <= 8).
template <int S, int NZ, int K3N, bool SIGN = false>
__global__ void __launch_bounds__(512, 1) pencil_col_tma_w(const __grid_constant__ CUtensorMap in_map,
                                                           const __grid_constant__ CUtensorMap out_map, ColTmaArgs a) {
    static_assert(!SIGN || (S > 0 && K3N <= 8), "sign map: forward pass, at most 2048 rows");
    constexpr int LOG2N = 12, VEC = 2;
    using G = Geo<LOG2N, VEC>;
    constexpr int BOX_ROWS = 256, NBOX = G::N / BOX_ROWS;
    static_assert(NZ >= 1 && NZ <= NBOX && K3N >= 1 && K3N <= 16, "block counts");
    // Only the exchange between stage 1 and stage 2 is a block-wide barrier.  Everything else is a hand-off to ONE
    // thread that talks to the TMA engine, so the other 511 only arrive (non-blocking) on an mbarrier and move on:
    //   lfree   (512 arrivals)  every thread has its stage-2 inputs      -> TL issues the next pair's box loads
    //   xfree_a (1)             the previous pair's last store left X    -> everybody may write exchange 2
    //   staged0 / staged1 (512) half h of the results is staged          -> TS0 / TS1 issue the box store
    //   xfree_b (1)             the first half has been read out         -> everybody may stage the second half
    // The three duties sit in three different warps, so no single warp carries all the waiting.
    constexpr int TL = 0, TS0 = 160, TS1 = 320;
    constexpr int TLAST = K3N > 8 ? TS1 : TS0;  // the thread that issues a pair's last store
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t full_bar, lfree, xfree_a, xfree_b, staged0, staged1;
    double2* L = (double2*)smem_raw;
    double* X = (double*)(smem_raw + G::L_BYTES);
    const int tid = threadIdx.x, c = tid & 1, tt = tid >> 1;
    const int k1 = tt >> 4, m = tt & 15;            // warp index == k1; m doubles as k2 in stage 3
    double* Xw = X + (size_t)k1 * 256 * VEC;        // this warp's private 4 KB slice (8 B entries)
    double2* Sw = (double2*)Xw;                     // the same slice as staging: [k3 (8)][k2 (16)][c] 16 B entries
    const double scale = S < 0 ? 1.0 / (double)G::N : 1.0;  // S:357
    __shared__ __align__(16) double2 tw_s[256];     // w_4096^j, j < 256 (conjugated for the inverse): every twiddle base of the pass

    if (tid == 0) {
        mbar_init(&full_bar, 1);
        mbar_init(&lfree, 512);
        mbar_init(&xfree_a, 1);
        mbar_init(&xfree_b, 1);
        mbar_init(&staged0, 512);
        mbar_init(&staged1, 512);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (tid < 256) {
        double2 w = a.tw[(size_t)tid << (TW_LOG2 - LOG2N)];
        if (S < 0) w.y = -w.y;
        tw_s[tid] = w;
    }
    __syncthreads();

    const int gpp = a.groups_per_plane;
    auto issue_load = [&](int plane, int g) {
        mbar_expect_tx(&full_bar, (unsigned)((size_t)NZ * BOX_ROWS * VEC * 16));
#pragma unroll 1
        for (int j = 0; j < NZ; j++) tma_load_3d(L + (size_t)j * BOX_ROWS * VEC, &in_map, &full_bar, g * VEC * 2, j * BOX_ROWS, plane);
    };

    // position of item = plane * gpp + g, advanced without divisions
    int plane = (int)(blockIdx.x / (unsigned)gpp), g = (int)(blockIdx.x % (unsigned)gpp);
    const int dplane = (int)(gridDim.x / (unsigned)gpp), dg = (int)(gridDim.x % (unsigned)gpp);
    const long long stride = gridDim.x;
    long long item = blockIdx.x;
    if (tid == TL && item < a.nitems) issue_load(plane, g);
    unsigned parity = 0;
    for (; item < a.nitems; item += stride, parity ^= 1) {
        int nplane = plane + dplane, ng = g + dg;
        if (ng >= gpp) { ng -= gpp; nplane++; }
        mbar_wait(&full_bar, parity);
        stage1<S, LOG2N, VEC, false, NZ, 0, true>(L, tt, c, tw_s, nullptr, 0, 0, -1);
        __syncthreads();
        double2 x[16];
        stage2_load<LOG2N, VEC>(L, tt, c, x);
        mbar_arrive(&lfree);
        if (!SIGN && tid == TLAST) {  // its own store of the previous pair's last half has finished reading X
            tma_wait_read_all();
            mbar_arrive(&xfree_a);
        }
        if (tid == TL && item + stride < a.nitems) {
            mbar_wait(&lfree, parity);  // L is free
            fence_async_proxy();
            issue_load(nplane, ng);
        }
        // ---- stage 2, warp-local transpose (write [k2][m ^ k2], read [k2 = m][n ^ m]), stage 3
        dft<S, 16>(x);
        twiddle<16>(x, tw_s[16 * m]);  // w_256^m
        double2 z[16];
        if constexpr (!SIGN) mbar_wait(&xfree_a, parity);  // (SIGN: X is only ever touched by its own warp)
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) Xw[((k2 << 4) | (m ^ k2)) * VEC + c] = x[oidx<16>(k2)].x;
        __syncwarp();
#pragma unroll
        for (int n = 0; n < 16; n++) z[n].x = Xw[((m << 4) | (n ^ m)) * VEC + c];
        __syncwarp();
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) Xw[((k2 << 4) | (m ^ k2)) * VEC + c] = x[oidx<16>(k2)].y;
        __syncwarp();
#pragma unroll
        for (int n = 0; n < 16; n++) z[n].y = Xw[((m << 4) | (n ^ m)) * VEC + c];
        dft<S, 16>(z);  // z[oidx(k3)] = output row k1 + 16*m + 256*k3 of column c
        if constexpr (SIGN) {
            // lane = 2 m + c: one ballot per row block gives the 16 rows x 2 columns of this warp; lane k3 stores word k3.
            // Off the real axis the read bit is the sign of the imaginary part (branch-free); the rare element within 1e-9 of
            // the axis goes through the full formula afterwards.
            unsigned mine = 0, near = 0;
#pragma unroll
            for (int k3 = 0; k3 < K3N; k3++) {
                const double2 v = z[oidx<16>(k3)];
                near |= (fabs(v.y) > 1e-9 * fabs(v.x)) ? 0u : (1u << k3);
                const unsigned w = __ballot_sync(0xffffffffu, v.y > 0.0);
                if ((tid & 31) == k3) mine = w;
            }
            if (__any_sync(0xffffffffu, near != 0u)) {  // (cold)
#pragma unroll
                for (int k3 = 0; k3 < K3N; k3++) {
                    const bool nr = (near >> k3) & 1u;
                    const int bit = nr ? read_bit_full(z[oidx<16>(k3)].x, z[oidx<16>(k3)].y, a.alpha) : 0;
                    const unsigned fix = __ballot_sync(0xffffffffu, nr), val = __ballot_sync(0xffffffffu, nr && bit);
                    if ((tid & 31) == k3) mine = (mine & ~fix) | val;
                }
            }
            if ((tid & 31) < 8) a.signmap[(((size_t)plane * a.map_groups + g) * 16 + k1) * 8 + (tid & 31)] = mine;
            __syncwarp();  // the next pair's exchange writes come after every lane's reads of the slice
            plane = nplane; g = ng;
            continue;
        }
        if constexpr (S > 0 && K3N == 16) {
            if (a.q32 != nullptr) {  // (uniform) float copy of q = |F|^2 for the median scan
                float4* dst = (float4*)a.q32 + (((size_t)plane * a.map_groups + g) * 16 + k1) * 128 + (tid & 31);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    float f[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const double2 v = z[oidx<16>(4 * j + u)];
                        f[u] = __double2float_rn(fma(v.x, v.x, v.y * v.y));
                    }
                    dst[j * 32] = make_float4(f[0], f[1], f[2], f[3]);
                }
            }
        }
        // ---- first half of the results: rows with k3 < 8
        __syncwarp();  // every lane is past its reads of the slice
#pragma unroll
        for (int k3 = 0; k3 < 8; k3++) {
            if (k3 >= K3N) continue;  // rows the store map clips anyway
            double2 v = z[oidx<16>(k3)];
            v.x *= scale; v.y *= scale;
            Sw[(k3 * 16 + m) * VEC + c] = v;
        }
        // median sample: this thread's row block for this pair, read back from its own staging entries
        const bool samp = S > 0 && a.sample_q != nullptr && g < a.sample_groups && c == (g & 1);
        const unsigned k3s = (((unsigned)tt * 2654435761u) >> 15 ^ ((unsigned)g * 0x9E3779B9u) >> 11 ^ (unsigned)plane * 7u) & 15u;
        if (samp && k3s < 8 && (int)k3s < K3N) {
            const double2 v = Sw[(k3s * 16 + m) * VEC + c];
            a.sample_q[(size_t)plane * a.sample_stride + (size_t)g * 256 + tt] = (unsigned long long)__double_as_longlong(fma(v.x, v.x, v.y * v.y));
        }
        fence_async_proxy();
        mbar_arrive(&staged0);
        if (tid == TS0) {
            mbar_wait(&staged0, parity);
            // smem image [k1][k3][k2][2 columns] -> rows k1 + 16 k2 + 256 k3 (map dims: col, k2, k3, k1, plane)
            tma_store_5d(&out_map, X, g * VEC * 2, 0, 0, 0, plane);
            tma_commit();
            if constexpr (K3N > 8) {  // (K3N <= 8: this was the last store of the pair, TS0 waits for it at the top of the next one)
                tma_wait_read_all();
                mbar_arrive(&xfree_b);
            }
        }
        // ---- second half: rows with 8 <= k3 < K3N
        if constexpr (K3N > 8) {
            mbar_wait(&xfree_b, parity);  // the first half has been read out by the TMA engine
#pragma unroll
            for (int k3 = 8; k3 < 16; k3++) {
                if (k3 >= K3N) continue;
                double2 v = z[oidx<16>(k3)];
                v.x *= scale; v.y *= scale;
                Sw[((k3 - 8) * 16 + m) * VEC + c] = v;
            }
            if (samp && k3s >= 8 && (int)k3s < K3N) {
                const double2 v = Sw[((k3s - 8) * 16 + m) * VEC + c];
                a.sample_q[(size_t)plane * a.sample_stride + (size_t)g * 256 + tt] = (unsigned long long)__double_as_longlong(fma(v.x, v.x, v.y * v.y));
            }
            fence_async_proxy();
            mbar_arrive(&staged1);
            if (tid == TS1) {
                mbar_wait(&staged1, parity);
                tma_store_5d(&out_map, X, g * VEC * 2, 0, 8, 0, plane);
                tma_commit();
            }
        }
        plane = nplane; g = ng;
    }
    if (tid == TS0 || tid == TS1) tma_wait_all();
}

// ------------------------------------------------------------------------------------------
// Column-resident embed (N = 4096, two columns per CTA): forward column FFT -> median inputs -> phase write at this
// pair's bins -> inverse column FFT, in ONE shared-memory residency.  Replaces the three-kernel sequence
// pencil_col_tma_w<+1> -> embed_scatter -> pencil_col_tma_w<-1> (write_bit_on_bin S:712-732 between the column halves of
// fft2d S:359-366 forward and inverse): the spectrum is never written to HBM, only
//   * q = |F|^2 of every element as its exact double, split into two 32-bit planes (qhi: sign/exponent/20 mantissa
//     bits, qlo: the other 32) in the warp order of the float copy above -- the median / capacity scan streams qhi (4 bytes
//     per element) and fetches qlo only for the ~1 % of elements its bracket cannot decide on the top word,
//   * the stratified median sample, and
//   * rows < H of the inverse column transform, in place over the row pass's output.
// Ownership makes the middle free: after the forward pass thread (k1, m, c) holds rows k1 + 16 m + 256 k3 (k3 = 0..15) of
// column c -- exactly the sixteen stride-256 inputs of the inverse pass's first radix-16 butterfly at position
// k1 + 16 m.  So the phase write happens on registers (the half-spectrum layout stores exactly one element per bin for
// 0 < x < PW/2; the host refuses other lists for this path) and the inverse starts without an exchange.
// Bins arrive as two 16-bit masks per thread and pair: pres (bit k3: a bin sits at my row block k3; shared by the batch,
// the walk is cover-independent S:797-799) and val (the bit to write; per image).
// Shared memory: L = 128 KB (TMA landing + forward exchange 1, free for the NEXT pair's loads as soon as the stage-2
// inputs are in registers) and X = 80 KB (warp-private slices for the two warp-local exchanges and the store staging, and
// -- between them, together with the tail of L -- the inverse's block-wide exchange 1, see XE_OFFSET).
// The embed is speculative with respect to capacity (S:1009-1012 needs the median, which needs the whole spectrum): the
// caller passes the cover through afterwards for images whose usable < nbits.
// ------------------------------------------------------------------------------------------
struct ColEmbedArgs {
    const double2* tw;
    long long nitems;      // nplanes * groups_per_plane
    int groups_per_plane;  // ld / 2 column pairs
    unsigned long long* sample_q;  // median sample (null: off), as pencil_col_tma_w
    unsigned sample_stride;
    int sample_groups;
    uint32_t* qhi;         // [plane][g][k1][j][lane] uint4: top words of q for rows k1 + 16 m + 256 (4 j + 0..3), lane = 2 m + c
    uint32_t* qlo;         // the low words, same layout
    const uint8_t* pair_has_bins;  // [plane % 3][g]: some thread of the pair holds a bin (else: forward + q only)
    const uint16_t* pres;  // [plane % 3][g][tid]   bit k3: a bin at (row k1 + 16 m + 256 k3, column 2 g + c)
    const uint16_t* val;   // [plane][g][tid]       the bits to write there
    int k3max;             // largest row block that holds a bin
    double cos_a, sin_a;   // of alpha (host libm: bit-exact with the reference's polar())
};

// Inverse exchange 1 (block-wide, 16-byte entries): entry p = k * 256 + position sits at row xe_pos(p) of a padded
// [4352][2 columns] image (one pad row per 16: the stride-16 writers and the stride-1 readers are both conflict-free).
// The image (136 KB) spans the last 56 KB of L -- row blocks 9..15, which no TMA load of a <= 2304-row input ever
// lands in -- and the 80 KB of X behind it, so the next pair's loads can still be issued as early as in the unfused pass.
constexpr int XE_ROWS = 4096 + 256;
constexpr size_t XE_OFFSET = (size_t)9 * 256 * 2 * 16;                       // 72 KB into L
constexpr size_t X_EMBED_BYTES = (size_t)XE_ROWS * 32 - ((size_t)131072 - XE_OFFSET);  // 80 KB
__device__ __forceinline__ int xe_pos(int p) { return p + (p >> 4); }

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}

template <int NZ, int K3N>
__global__ void __launch_bounds__(512, 1) pencil_col_embed_w(const __grid_constant__ CUtensorMap in_map,
                                                             const __grid_constant__ CUtensorMap out_map,
                                                             const __grid_constant__ CUtensorMap out_map2, ColEmbedArgs a) {
    constexpr int LOG2N = 12, VEC = 2;
    using G = Geo<LOG2N, VEC>;
    constexpr int BOX_ROWS = 256, NBOX = G::N / BOX_ROWS;
    static_assert(NZ >= 1 && NZ <= NBOX && K3N >= 1 && K3N <= 16, "block counts");
    // its six radix-16 butterflies keep the textbook second layer: folding the forward three costs a spilled register
    // (+6 % kernel time), folding the inverse three changes nothing (profiles/r2_experiments.txt)
    constexpr bool FWD_PLAIN = true, INV_PLAIN = true;
    constexpr bool EARLY_PREFETCH = NZ <= 9;  // the landing area stays clear of the inverse exchange image
    // Results leave through the warp slices of X, eight row blocks (64 KB) at a time.  One or two extra blocks (UHD: the
    // ninth) are staged behind them in the 16 KB X has left, so both box stores of a pair are issued together; more than
    // that waits for the first store to have read the slices (second round, as pencil_col_tma_w).
    constexpr int K3B = K3N > 8 ? K3N - 8 : 0;        // row blocks beyond the first eight
    constexpr bool ONE_ROUND = K3B <= 2;
    constexpr int TL = 0, TS0 = 160, TS1 = 320;
    constexpr int TLAST = (K3N > 8 && !ONE_ROUND) ? TS1 : TS0;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t full_bar, lfree, xfree_a, xfree_b, staged0, staged1;
    __shared__ __align__(16) double2 tw_s[256];       // w_4096^j, j < 256: every twiddle base of the three stages
    __shared__ __align__(16) uint32_t s_pres[256], s_val[256];  // this pair's bin masks (two threads per word)
    double2* L = (double2*)smem_raw;
    double* X = (double*)(smem_raw + G::L_BYTES);
    double2* X2 = (double2*)(smem_raw + G::L_BYTES + 65536);  // staging of the row blocks beyond the eighth (ONE_ROUND)
    double2* XE = (double2*)(smem_raw + XE_OFFSET);
    const int tid = threadIdx.x, c = tid & 1, tt = tid >> 1;
    const int k1 = tt >> 4, m = tt & 15;
    const int mp = k1 + 16 * m;                      // my position in the inverse pass's first stage (row mod 256)
    double* Xw = X + (size_t)k1 * 256 * VEC;
    double2* Sw = (double2*)Xw;
    const double scale = 1.0 / (double)G::N;        // S:357

    if (tid == 0) {
        mbar_init(&full_bar, 1);
        mbar_init(&lfree, 512);
        mbar_init(&xfree_a, 1);
        mbar_init(&xfree_b, 1);
        mbar_init(&staged0, 512);
        mbar_init(&staged1, 512);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (tid < 256) tw_s[tid] = a.tw[(size_t)tid << (TW_LOG2 - LOG2N)];
    __syncthreads();

    const int gpp = a.groups_per_plane;
    auto issue_load = [&](int plane, int g) {
        mbar_expect_tx(&full_bar, (unsigned)((size_t)NZ * BOX_ROWS * VEC * 16));
#pragma unroll 1
        for (int j = 0; j < NZ; j++) tma_load_3d(L + (size_t)j * BOX_ROWS * VEC, &in_map, &full_bar, g * VEC * 2, j * BOX_ROWS, plane);
    };
    // position of item = plane * gpp + g, advanced without divisions
    int plane = (int)(blockIdx.x / (unsigned)gpp), g = (int)(blockIdx.x % (unsigned)gpp);
    const int dplane = (int)(gridDim.x / (unsigned)gpp), dg = (int)(gridDim.x % (unsigned)gpp);
    const long long stride = gridDim.x;
    long long item = blockIdx.x;
    if (tid == TL && item < a.nitems) issue_load(plane, g);
    unsigned parity = 0, par_st = 0;  // par_st: phase of the store hand-offs (they only advance for pairs that are stored)
    for (; item < a.nitems; item += stride, parity ^= 1) {
        int nplane = plane + dplane, ng = g + dg;
        if (ng >= gpp) { ng -= gpp; nplane++; }
        // this pair's bin masks: 4-byte async copies straight to shared memory (no registers held across the forward pass);
        // completed before the barrier below, read at the phase write
        const bool has_bins = a.pres != nullptr && a.pair_has_bins[(plane % 3) * gpp + g] != 0;  // (uniform; used after the forward pass)
        if (a.pres) {
            const uint32_t* src = tid < 256 ? (const uint32_t*)a.pres + ((size_t)(plane % 3) * gpp + g) * 256 + tid
                                            : (const uint32_t*)a.val + ((size_t)plane * gpp + g) * 256 + (tid - 256);
            cp_async4(tid < 256 ? &s_pres[tid] : &s_val[tid - 256], src);
            cp_async_commit();
        }
        // ================= forward column transform (as pencil_col_tma_w<+1, NZ, 16>) =================
        mbar_wait(&full_bar, parity);
        stage1<+1, LOG2N, VEC, false, NZ, 0, false, FWD_PLAIN>(L, tt, c, tw_s, nullptr, 0, 0, -1);
        cp_async_wait_all();
        __syncthreads();
        double2 x[16];
        stage2_load<LOG2N, VEC>(L, tt, c, x);
        mbar_arrive(&lfree);
        if (tid == TLAST) {  // its own store of the previous pair's last half has finished reading X
            tma_wait_read_all();
            mbar_arrive(&xfree_a);
        }
        if (EARLY_PREFETCH && tid == TL && item + stride < a.nitems) {
            mbar_wait(&lfree, parity);  // L is free
            fence_async_proxy();
            issue_load(nplane, ng);
        }
        dft<+1, 16, FWD_PLAIN>(x);
        twiddle<16>(x, tw_s[16 * m]);  // w_256^m
        double2 z[16];
        mbar_wait(&xfree_a, parity);
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) Xw[((k2 << 4) | (m ^ k2)) * VEC + c] = x[oidx<16>(k2)].x;
        __syncwarp();
#pragma unroll
        for (int n = 0; n < 16; n++) z[n].x = Xw[((m << 4) | (n ^ m)) * VEC + c];
        __syncwarp();
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) Xw[((k2 << 4) | (m ^ k2)) * VEC + c] = x[oidx<16>(k2)].y;
        __syncwarp();
#pragma unroll
        for (int n = 0; n < 16; n++) z[n].y = Xw[((m << 4) | (n ^ m)) * VEC + c];
        dft<+1, 16, FWD_PLAIN>(z);  // z[oidx(k3)] = F[row k1 + 16 m + 256 k3][column 2 g + c]
        // ================= median inputs: q of every element (exact double, two word planes) + the sample =================
        {
            const size_t qbase = ((((size_t)plane * gpp + g) * 16 + k1) * 128 + (tid & 31)) * 4;  // in words
            const bool samp = a.sample_q != nullptr && g < a.sample_groups && c == (g & 1);
            const unsigned k3s = (((unsigned)tt * 2654435761u) >> 15 ^ ((unsigned)g * 0x9E3779B9u) >> 11 ^ (unsigned)plane * 7u) & 15u;
            unsigned long long qs = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                unsigned hi[4], lo[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const double2 v = z[oidx<16>(4 * j + u)];
                    const double q = fma(v.x, v.x, v.y * v.y);
                    hi[u] = (unsigned)__double2hiint(q);
                    lo[u] = (unsigned)__double2loint(q);
                    if (k3s == (unsigned)(4 * j + u)) qs = (unsigned long long)__double_as_longlong(q);
                }
                *(uint4*)(a.qhi + qbase + (size_t)j * 128) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *(uint4*)(a.qlo + qbase + (size_t)j * 128) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            if (samp) a.sample_q[(size_t)plane * a.sample_stride + (size_t)g * 256 + tt] = qs;
        }
        // A pair without bins comes back from the inverse pass as it went in (up to rounding): leave its columns alone.
        // (uniform per CTA; with no bins at all -- a capacity-only call -- every pair takes this exit)
        if (!has_bins) {
            if (!EARLY_PREFETCH && tid == TL && item + stride < a.nitems) {
                mbar_wait(&lfree, parity);  // every thread has its stage-2 inputs: L is free
                fence_async_proxy();
                issue_load(nplane, ng);
            }
            plane = nplane; g = ng;
            continue;
        }
        // ================= phase write (write_bit_on_bin S:712-732) on the registers that own the bins =================
        {
            const unsigned pres = (s_pres[tt] >> (16 * c)) & 0xFFFFu, val = s_val[tt] >> (16 * c);
            if (__any_sync(0xffffffffu, pres != 0u)) {
#pragma unroll
                for (int k3 = 0; k3 < 16; k3++) {
                    if (k3 > a.k3max) break;  // (uniform)
                    if ((pres >> k3) & 1u) {
                        double2& v = z[oidx<16>(k3)];
                        // |F| (S:716).  sqrt of the correctly rounded sum of squares: within one ulp of abs(complex) and an order
                        // of magnitude cheaper than hypot() under divergence; spectra here are far from over/underflow, and a
                        // magnitude below 1e-154 ends at the 1e-12 floor either way
                        const double mag = fmax(1e-12, sqrt(fma(v.x, v.x, v.y * v.y)));
                        const double s = ((val >> k3) & 1u) ? a.sin_a : -a.sin_a;  // theta = +-alpha (S:718-720)
                        v = make_double2(mag * a.cos_a, mag * s);          // std::polar(mag, theta); the mirror is not stored
                    }
                }
            }
        }
        // ================= inverse column transform =================
        // stage 1 in registers: inputs rows mp + 256 n = z[oidx(n)]
        {
#pragma unroll
            for (int n = 0; n < 16; n++) x[n] = z[oidx<16>(n)];
            dft<-1, 16, INV_PLAIN>(x);
            double2 w1 = tw_s[mp];
            w1.y = -w1.y;
            twiddle<16>(x, w1);  // x[oidx(k)] = exchange-1 entry k * 256 + mp
        }
        __syncthreads();  // every warp is past the reads of its private slice of X: the exchange image may cover it
#pragma unroll
        for (int k = 0; k < 16; k++) XE[xe_pos(k * 256 + mp) * VEC + c] = x[oidx<16>(k)];
        __syncthreads();
#pragma unroll
        for (int n = 0; n < 16; n++) z[n] = XE[xe_pos(k1 * 256 + n * 16 + m) * VEC + c];
        __syncthreads();  // block-wide reads done: the slices of X are private again, L's tail is free
        if (!EARLY_PREFETCH && tid == TL && item + stride < a.nitems) {  // (all 16 landing blocks overlap the image)
            fence_async_proxy();
            issue_load(nplane, ng);
        }
        // stage 2, warp-local transpose, stage 3 (as pencil_col_tma_w<-1>)
        dft<-1, 16, INV_PLAIN>(z);
        {
            double2 w2 = tw_s[16 * m];
            w2.y = -w2.y;
            twiddle<16>(z, w2);
        }
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) Xw[((k2 << 4) | (m ^ k2)) * VEC + c] = z[oidx<16>(k2)].x;
        __syncwarp();
#pragma unroll
        for (int n = 0; n < 16; n++) x[n].x = Xw[((m << 4) | (n ^ m)) * VEC + c];
        __syncwarp();
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) Xw[((k2 << 4) | (m ^ k2)) * VEC + c] = z[oidx<16>(k2)].y;
        __syncwarp();
#pragma unroll
        for (int n = 0; n < 16; n++) x[n].y = Xw[((m << 4) | (n ^ m)) * VEC + c];
        dft<-1, 16, INV_PLAIN>(x);  // x[oidx(k3)] = output row k1 + 16 m + 256 k3 of column c
        // ---- first half of the results: rows with k3 < 8 (and, ONE_ROUND, the one or two blocks behind them)
        __syncwarp();
#pragma unroll
        for (int k3 = 0; k3 < 8; k3++) {
            if (k3 >= K3N) continue;
            double2 v = x[oidx<16>(k3)];
            v.x *= scale; v.y *= scale;
            Sw[(k3 * 16 + m) * VEC + c] = v;
        }
        if constexpr (K3N > 8 && ONE_ROUND) {
#pragma unroll
            for (int k3 = 8; k3 < K3N; k3++) {
                double2 v = x[oidx<16>(k3)];
                v.x *= scale; v.y *= scale;
                X2[((k1 * K3B + (k3 - 8)) * 16 + m) * VEC + c] = v;  // image [k1][k3 - 8][k2][column] of the second box
            }
        }
        fence_async_proxy();
        mbar_arrive(&staged0);
        if (tid == TS0) {
            mbar_wait(&staged0, par_st);
            tma_store_5d(&out_map, X, g * VEC * 2, 0, 0, 0, plane);
            if constexpr (K3N > 8 && ONE_ROUND) tma_store_5d(&out_map2, X2, g * VEC * 2, 0, 8, 0, plane);
            tma_commit();
            if constexpr (K3N > 8 && !ONE_ROUND) {
                tma_wait_read_all();
                mbar_arrive(&xfree_b);
            }
        }
        if constexpr (K3N > 8 && !ONE_ROUND) {
            mbar_wait(&xfree_b, par_st);
#pragma unroll
            for (int k3 = 8; k3 < 16; k3++) {
                if (k3 >= K3N) continue;
                double2 v = x[oidx<16>(k3)];
                v.x *= scale; v.y *= scale;
                Sw[((k3 - 8) * 16 + m) * VEC + c] = v;
            }
            fence_async_proxy();
            mbar_arrive(&staged1);
            if (tid == TS1) {
                mbar_wait(&staged1, par_st);
                tma_store_5d(&out_map, X, g * VEC * 2, 0, 8, 0, plane);
                tma_commit();
            }
        }
        par_st ^= 1;
        plane = nplane; g = ng;
    }
    if (tid == TS0 || tid == TS1) tma_wait_all();
}

// ------------------------------------------------------------------------------------------
// u8 RGB row -> three forward row pencils.  item = (image, y < H)
// ------------------------------------------------------------------------------------------
struct U8Args {
    double2* spec;
    const double2* tw;
    const uint8_t* img_in;
    uint8_t* img_out;
    long long nitems;  // nimg * H
    int W, H, PW, PH, center;
};

template <int LOG2N>
struct U8Geo {
    static constexpr int N = 1 << LOG2N;
    static constexpr size_t U_BYTES = (size_t)N * 3 + 32;  // aligned span of one RGB row (W <= N)
};

template <int LOG2N, int UNITS>
__global__ void __launch_bounds__(Geo<LOG2N, 1>::UT* UNITS, 1) pencil_u8_fwd(U8Args a) {
    using G = Geo<LOG2N, 1>;
    constexpr size_t UB = U8Geo<LOG2N>::U_BYTES;
    constexpr size_t UNIT_BYTES = G::L_BYTES + 2 * UB;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int unit = threadIdx.x / G::UT, tt = threadIdx.x % G::UT;
    unsigned char* base = smem_raw + (size_t)unit * UNIT_BYTES;
    double2* L = (double2*)base;  // both exchanges run in place in L (nothing lands here)
    unsigned char* U[2] = {base + G::L_BYTES, base + G::L_BYTES + UB};
    const int bar_id = 1 + unit;
    const long long stride = (long long)gridDim.x * UNITS;
    long long item = (long long)blockIdx.x * UNITS + unit;
    const size_t row_bytes = (size_t)a.W * 3;
    const uintptr_t img_base = (uintptr_t)a.img_in;
    const uintptr_t img_end = img_base + (size_t)(a.nitems)*row_bytes;
    ThreadTw<+1, LOG2N, 1> ttw;
    ttw.load(a.tw, tt);

    auto issue_row = [&](long long it, unsigned char* dst) {
        const uintptr_t start = img_base + (size_t)it * row_bytes;
        const uintptr_t a0 = start & ~(uintptr_t)15;
        const int nchunks = (int)((start + row_bytes - a0 + 15) >> 4);
        for (int i = tt; i < nchunks; i += G::UT) {
            const uintptr_t src = a0 + (size_t)i * 16;
            long long avail = (long long)(img_end - src);
            int nb = avail >= 16 ? 16 : (avail > 0 ? (int)avail : 0);
            cp_async16(dst + (size_t)i * 16, (const void*)(nb ? src : img_base), nb);
        }
        cp_async_commit();
    };

    if (unit & 1) {  // de-phase odd units once: FP64-heavy and LSU-heavy phases of neighbours interleave
        const long long t0 = clock64();
        while (clock64() - t0 < STAGGER_CYCLES) {}
    }
    int buf = 0;
    if (item < a.nitems) issue_row(item, U[0]);
    for (; item < a.nitems; item += stride, buf ^= 1) {
        cp_async_wait_all();
        unit_bar(bar_id, G::UT);
        if (item + stride < a.nitems) issue_row(item + stride, U[buf ^ 1]);
        const long long img = item / a.H;
        const int y = (int)(item % a.H);
        const uint8_t* urow = U[buf] + ((img_base + (size_t)item * row_bytes) & 15);
        for (int ch = 0; ch < 3; ch++) {
            stage1<+1, LOG2N, 1, true>(L, tt, 0, a.tw, urow, a.W, ch, a.center ? (y & 1) : -1);
            unit_bar(bar_id, G::UT);
            double2 x[16];
            stage2_load<LOG2N, 1>(L, tt, 0, x);
            unit_bar(bar_id, G::UT);
            stage23_inplace<+1, LOG2N>(L, tt, ttw.s2v(), x, bar_id);
            double2* out = a.spec + (((size_t)img * 3 + ch) * a.PH + y) * a.PW;
#pragma unroll
            for (int k3 = 0; k3 < 16; k3++) out[tt + G::TP * k3] = x[oidx<16>(k3)];
            unit_bar(bar_id, G::UT);  // L is rewritten by the next plane's stage 1
        }
    }
}

// ------------------------------------------------------------------------------------------
// three inverse row pencils -> u8 RGB row.  item = (image, y < H)
// ------------------------------------------------------------------------------------------
template <int LOG2N, int UNITS>
__global__ void __launch_bounds__(Geo<LOG2N, 1>::UT* UNITS, 1) pencil_u8_inv(U8Args a) {
    using G = Geo<LOG2N, 1>;
    constexpr size_t UB = U8Geo<LOG2N>::U_BYTES;
    constexpr size_t UNIT_BYTES = G::L_BYTES + G::X_BYTES + UB;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int unit = threadIdx.x / G::UT, tt = threadIdx.x % G::UT;
    unsigned char* base = smem_raw + (size_t)unit * UNIT_BYTES;
    double2* L = (double2*)base;
    double* X = (double*)(base + G::L_BYTES);
    unsigned char* O = base + G::L_BYTES + G::X_BYTES;
    const int bar_id = 1 + unit;
    const long long stride = (long long)gridDim.x * UNITS;
    long long item = (long long)blockIdx.x * UNITS + unit;
    const size_t row_bytes = (size_t)a.W * 3;
    const double scale = 1.0 / (double)G::N;
    ThreadTw<-1, LOG2N, 1> ttw;
    ttw.load(a.tw, tt);

    auto src_row = [&](long long it, int ch) -> const double2* {
        const long long img = it / a.H;
        const int y = (int)(it % a.H);
        return a.spec + (((size_t)img * 3 + ch) * a.PH + y) * a.PW;
    };
    auto issue_loads = [&](const double2* src) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int e = tt + i * G::UT;
            cp_async16(&L[e], src + e, 16);
        }
        cp_async_commit();
    };

    if (unit & 1) {  // de-phase odd units once: FP64-heavy and LSU-heavy phases of neighbours interleave
        const long long t0 = clock64();
        while (clock64() - t0 < STAGGER_CYCLES) {}
    }
    if (item < a.nitems) issue_loads(src_row(item, 0));
    for (; item < a.nitems; item += stride) {
        const int y = (int)(item % a.H);
        const uintptr_t gstart = (uintptr_t)a.img_out + (size_t)item * row_bytes;
        const int off = (int)(gstart & 15);
        for (int ch = 0; ch < 3; ch++) {
            cp_async_wait_all();
            unit_bar(bar_id, G::UT);
            stage1<-1, LOG2N, 1, false>(L, tt, 0, a.tw, nullptr, 0, 0, -1);
            unit_bar(bar_id, G::UT);
            double2 x[16];
            stage2_load<LOG2N, 1>(L, tt, 0, x);
            unit_bar(bar_id, G::UT);  // L is free: prefetch the next plane / next item's first plane
            if (ch < 2) issue_loads(src_row(item, ch + 1));
            else if (item + stride < a.nitems) issue_loads(src_row(item + stride, 0));
            stage23<-1, LOG2N, 1>(X, tt, 0, ttw.s2v(), x, bar_id);
#pragma unroll
            for (int k3 = 0; k3 < 16; k3++) {
                const int k = tt + G::TP * k3;
                if (k < a.W) {
                    double v = x[oidx<16>(k3)].x * scale;           // ifft_crop S:401: real part
                    if (a.center && ((k + y) & 1)) v = -v;           // S:1102
                    O[off + k * 3 + ch] = clamp8(v);
                }
            }
        }
        unit_bar(bar_id, G::UT);
        // O[off .. off+row_bytes) -> global; aligned 16 B chunks in the middle, bytes at the edges
        {
            unsigned char* gdst = (unsigned char*)(gstart - off);
            const int total = off + (int)row_bytes;
            const int nchunks = (total + 15) >> 4;
            for (int i = tt; i < nchunks; i += G::UT) {
                const int b0 = i * 16, b1 = b0 + 16;
                if (b0 >= off && b1 <= total) {
                    *(uint4*)(gdst + b0) = *(const uint4*)(O + b0);
                } else {
                    for (int b = (b0 > off ? b0 : off); b < (b1 < total ? b1 : total); b++) gdst[b] = O[b];
                }
            }
        }
        unit_bar(bar_id, G::UT);  // O is rewritten by the next item
    }
}

// ==========================================================================================
// Real-input symmetry: half-spectrum passes.
// A real plane's spectrum is Hermitian, F[y][x] = conj(F[(PH-y)%PH][(PW-x)%PW]), so only columns
// x = 0 .. PW/2 are kept (row stride ld = PW/2 + 16; the 15 pad columns are zero).  Two image rows
// share one complex FFT: z = row_a + i*row_b.  Halves the column passes, the workspace, the median
// scan and the row-pass FP64 work.  Results are the same within rounding (the reference's spectrum
// is Hermitian to ~1e-12, and S:401 keeps only the real part of the inverse).
// ==========================================================================================
struct R2CArgs {
    double2* spec;      // [nplanes][PH][ld]
    const double2* tw;
    const uint8_t* img_in;
    uint8_t* img_out;
    long long nitems;   // nimg * ceil(H/2) row pairs
    int W, H, PW, PH, ld, center;
    long long stagger;  // cycles the odd unit of a CTA waits once at start (de-phases the two units)
};

// FOLD (forward rows of a plane with PH = 8192 whose column pass only has to serve an extract): the first radix-2 step
// of the 8192-point COLUMN transform (decimation in frequency) is taken here, where rows y and y + 4096 meet anyway:
//   A_y[k] = F_y[k] + F_{y+4096}[k]              -> stored row y          column FFT_4096 over y gives F2d[2 y'][k]
//   B_y[k] = (F_y[k] - F_{y+4096}[k]) w_8192^y   -> stored row y + 4096   column FFT_4096 over y gives F2d[2 y' + 1][k]
// so the column pass of such a plane is two independent 4096-point passes (the 4096-row kernels, sign map included)
// instead of a four-step pass.  Narrow rows: the row pair of one transform is (y, y + 4096) instead of (y, y + 1) and
// the fold happens on the registers that hold the two spectra.  WIDE rows (one 8192-pixel row per transform): the two
// units of a CTA take rows y and y + 4096 and swap halves of their spectra through shared memory.

// ---- forward: two u8 rows x 3 planes -> 2 x 3 half-spectrum rows ----------------------------
// Staging of a row pair: each row has its own zero-padded region of RS = 3N + 32 bytes, filled by a FIXED
// number of 16 B cp.async chunks (source size 0 beyond the row end -> zero fill), so stage 1 reads pixel
// x < N without any bounds test and an absent second row (odd H) is simply all zeros.
template <int LOG2N>
struct R2CGeo {
    static constexpr int N = 1 << LOG2N;
    static constexpr size_t RS = (size_t)N * 3 + 32;                        // staging bytes per row
    static constexpr size_t UB = 2 * RS;                                    // per buffer (row pair)
    static constexpr size_t LF_BYTES = (size_t)(N + xpad<LOG2N>()) * 16;    // forward: L with padded exchange 2
    static constexpr size_t FWD_UNIT = LF_BYTES + 2 * UB;
};

// WIDE (LOG2N = 12 only): ONE real row of 2N = 8192 pixels per item through the N-point complex transform of
// z[n] = x[2n] + i*x[2n+1]:  with E/O the spectra of the even/odd samples (the same Hermitian split as the row-pair
// case), X[k] = E[k] + W^k O[k] and X[N-k] = conj(E[k] - W^k O[k]), W = exp(+2 pi i / 2N), k = 0..N/2, X[N/2] = Z[N/2].
// The half-spectrum row then has N+1 = 4097 columns (ld = N + 16).
template <int LOG2N, int UNITS, bool CENTER, bool WIDE = false, bool FOLD = false>
__global__ void __launch_bounds__(Geo<LOG2N, 1>::UT* UNITS, 1) pencil_u8_fwd_r2c(R2CArgs a) {
    using G = Geo<LOG2N, 1>;
    using RG = R2CGeo<LOG2N>;
    static_assert(!WIDE || LOG2N == 12, "the wide variant packs an 8192-pixel row into a 4096-point transform");
    static_assert(!(WIDE && FOLD) || UNITS == 2, "folding wide rows pairs the two units of a CTA");
    constexpr int FR = 4096;                    // FOLD: distance of the two rows that meet (PH / 2)
    constexpr bool XF = WIDE && FOLD;           // cross-unit fold
    constexpr int N = G::N, NH = N / 2;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int unit = threadIdx.x / G::UT, tt = threadIdx.x % G::UT;
    unsigned char* base = smem_raw + (size_t)unit * RG::FWD_UNIT;
    double2* L = (double2*)base;
    unsigned char* Ub = base + RG::LF_BYTES;            // two row-pair buffers of UB bytes
    const int bar_id = 1 + unit;
    // (cross-unit fold: both units of a CTA walk the same items, unit u takes the row FR * u below)
    const long long stride = XF ? (long long)gridDim.x : (long long)gridDim.x * UNITS;
    long long item = XF ? (long long)blockIdx.x : (long long)blockIdx.x * UNITS + unit;
    const int HP = FOLD ? FR : (WIDE ? a.H : (a.H + 1) / 2);  // items (row pairs, or single wide rows) per image
    const int YS = (WIDE || FOLD) ? 1 : 2;      // first rows of consecutive items
    const int Y2 = FOLD ? FR : 1;               // narrow rows: distance of the two rows of a pair
    const int YU = XF ? unit * FR : 0;
    const size_t row_bytes = (size_t)a.W * 3;
    const int nch = (int)((row_bytes + 15) >> 4) + 1;   // chunks per row: covers every 16 B phase of the row start
    const uintptr_t img_base = (uintptr_t)a.img_in;
    ThreadTw<+1, LOG2N, 1> ttw;
    ttw.load(a.tw, tt);

    auto pair_start = [&](long long it) -> uintptr_t {
        const long long img = it / HP;
        const int y0 = YS * (int)(it % HP) + YU;
        return img_base + ((size_t)img * a.H + y0) * row_bytes;
    };
    // rows of the item that exist: narrow 1 or 2 (the second row of the last pair of an odd H / of a folded pair may be
    // absent), wide 1 (cross-unit fold: 0 when this unit's row lies below the image)
    auto pair_rows = [&](long long it) -> int {
        const int y0 = YS * (int)(it % HP) + YU;
        if constexpr (WIDE) return y0 < a.H ? 1 : 0;
        else return y0 + Y2 < a.H ? 2 : 1;
    };
    auto issue_rows = [&](long long it, int buf) {
        const uintptr_t s0 = pair_start(it);
        const int nrows = pair_rows(it);
#pragma unroll
        for (int r = 0; r < (WIDE ? 1 : 2); r++) {  // wide: one row of up to 2N pixels fills the whole buffer
            const uintptr_t start = s0 + (size_t)r * Y2 * row_bytes;
            const unsigned char* a0 = (const unsigned char*)(start & ~(uintptr_t)15);
            const int span = r < nrows ? (int)(start & 15) + (int)row_bytes : 0;  // bytes from a0 to the row end; absent row: all zero-fill
            unsigned char* dst = Ub + (size_t)buf * RG::UB + (size_t)r * RG::RS;
            for (int i = tt; i < nch; i += G::UT) {
                const int avail = span - i * 16;
                const int nb = avail >= 16 ? 16 : (avail > 0 ? avail : 0);
                cp_async16(dst + i * 16, nb ? a0 + i * 16 : (const unsigned char*)a.img_in, nb);  // nb == 0: nothing is read (pure zero fill; the address stays inside the batch)
            }
        }
        cp_async_commit();
    };

    // zero the staging tails once (bytes beyond the chunks are never written again)
    for (size_t i = (size_t)tt * 16; i < 2 * RG::UB; i += (size_t)G::UT * 16) *(uint4*)(Ub + i) = make_uint4(0, 0, 0, 0);
    if (unit & 1) {
        const long long t0 = clock64();
        while (clock64() - t0 < a.stagger) {}
    }
    unit_bar(bar_id, G::UT);
    int buf = 0;
    if (item < a.nitems) issue_rows(item, 0);
    for (; item < a.nitems; item += stride, buf ^= 1) {
        cp_async_wait_all();
        unit_bar(bar_id, G::UT);
        if (item + stride < a.nitems) issue_rows(item + stride, buf ^ 1);
        const long long img = item / HP;
        const int y0 = YS * (int)(item % HP) + YU;
        const int nrows = pair_rows(item);
        const uintptr_t s0 = pair_start(item);
        // shared-space byte addresses of the two real inputs of z[tt] (channel 0): pixel tt of row 0 / row 1, or
        // (wide) pixels 2 tt and 2 tt + 1 of the one row
        const uint8_t* r0 = smem_raw + (size_t)unit * RG::FWD_UNIT + RG::LF_BYTES + (size_t)buf * RG::UB + (s0 & 15) + (size_t)tt * (WIDE ? 6 : 3);
        const uint8_t* r1 = WIDE ? r0 + 3
                                 : smem_raw + (size_t)unit * RG::FWD_UNIT + RG::LF_BYTES + (size_t)buf * RG::UB + RG::RS + ((s0 + (size_t)Y2 * row_bytes) & 15) + (size_t)tt * 3;
        [[maybe_unused]] double2 wy = make_double2(1.0, 0.0);   // FOLD: w_8192^y of this item's first row
        if constexpr (FOLD) wy = a.tw[(size_t)(y0 & (FR - 1)) << (TW_LOG2 - 13)];
        for (int ch = 0; ch < 3; ch++) {
            // ---- stage 1 on z = row0 + i*row1 (plane split, centre sign, zero pad fused; S:383-398)
            double2 x[16];
#pragma unroll
            for (int j = 0; j < G::J1; j++) {
                const int m = tt + j * G::TP;
                double2* xj = x + j * G::R1;
#pragma unroll
                for (int n = 0; n < G::R1; n++) {
                    const unsigned o = (unsigned)((n * 256 + j * G::TP) * (WIDE ? 6 : 3)) + (unsigned)ch;
                    double v0 = u8_to_half_double(r0[o]), v1 = u8_to_half_double(r1[o]);  // Z / 2 from here on
                    if constexpr (CENTER) {  // apply_center S:392: (-1)^(x+y); pair: x parity == m parity; wide: x = 2n, 2n+1
                        if constexpr (FOLD && !WIDE) {  // rows y0 and y0 + 4096 have the same parity
                            if ((m + y0) & 1) { v0 = -v0; v1 = -v1; }
                        } else {
                            if (((WIDE ? 0 : m) + y0) & 1) v0 = -v0; else v1 = -v1;
                        }
                    }
                    xj[n] = make_double2(v0, v1);
                }
                dft<+1, G::R1>(xj);
                double2 w1 = a.tw[(size_t)m << (TW_LOG2 - LOG2N)];
                twiddle<G::R1>(xj, w1);
            }
#pragma unroll
            for (int j = 0; j < G::J1; j++)
#pragma unroll
                for (int k = 0; k < G::R1; k++) L[k * 256 + tt + j * G::TP] = x[j * G::R1 + oidx<G::R1>(k)];
            unit_bar(bar_id, G::UT);
            stage2_load<LOG2N, 1>(L, tt, 0, x);
            unit_bar(bar_id, G::UT);
            stage23_inplace<+1, LOG2N, true>(L, tt, ttw.s2v(), x, bar_id);  // x[oidx(k3)] = Z[tt + TP*k3]
            unit_bar(bar_id, G::UT);          // all reads of L done
            // ---- split Z into the spectra of the two real rows: partners Z[N-k] through L
#pragma unroll
            for (int k3 = 8; k3 < 16; k3++) L[tt + G::TP * (k3 - 8)] = x[oidx<16>(k3)];  // Z[N/2 + j] at L[j]
            if (tt == 0) L[NH] = x[oidx<16>(0)];                                          // Z[0] is its own partner
            unit_bar(bar_id, G::UT);
            double2* out0 = a.spec + (((size_t)img * 3 + ch) * a.PH + y0) * a.ld;
            double2* out1 = out0 + (size_t)Y2 * a.ld;
            const double2* Lp = L + NH - tt;
            if constexpr (XF) {
                // this unit's half-spectrum row X[0..N] as in the wide case below, but nothing is stored yet: unit 0 keeps X[k]
                // (k = tt + 256 k3) and parks X[N-k] in the free tail of its L, unit 1 the other way round; after the CTA
                // barrier each unit holds both rows' values of its columns and stores A (row y) and B (row y + 4096)
                const double2 wb = a.tw[(size_t)tt << (TW_LOG2 - LOG2N - 1)];
                double2 keep[8];
                double2* Sm = L + NH + 1;                                   // 2049 entries, clear of the split's L[0..NH]
                const double2* So = (const double2*)(smem_raw + (size_t)(unit ^ 1) * RG::FWD_UNIT) + NH + 1;
#pragma unroll
                for (int k3 = 0; k3 < 8; k3++) {
                    constexpr double C32[8] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                                               0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785};
                    constexpr double S32[8] = {0.0, 0.19509032201612826785, 0.38268343236508977173, 0.55557023301960222474,
                                               0.70710678118654752440, 0.83146961230254523708, 0.92387953251128675613, 0.98078528040323044913};
                    const double2 z = x[oidx<16>(k3)];
                    const double2 zn = Lp[-G::TP * k3];
                    const double2 E = make_double2(z.x + zn.x, z.y - zn.y);   // (z, zn hold Z / 2)
                    const double2 O = make_double2(z.y + zn.y, zn.x - z.x);
                    const double2 w = k3 == 0 ? wb : cmulc<+1>(wb, C32[k3], S32[k3]);
                    const double2 t = cmul(w, O);
                    const double2 XK = make_double2(E.x + t.x, E.y + t.y);   // X[k]
                    const double2 XM = make_double2(E.x - t.x, t.y - E.y);   // X[N-k] = conj(E - t)
                    keep[k3] = unit == 0 ? XK : XM;
                    Sm[k3 * G::TP + tt] = unit == 0 ? XM : XK;
                }
                const double2 xnh = make_double2(2.0 * x[oidx<16>(8)].x, 2.0 * x[oidx<16>(8)].y);  // X[N/2] = Z[N/2]
                if (tt == 0) Sm[8 * G::TP] = xnh;
                __syncthreads();
                const size_t yA = (size_t)(y0 & (FR - 1));
                double2* rowA = a.spec + (((size_t)img * 3 + ch) * a.PH + yA) * a.ld;
                double2* rowB = rowA + (size_t)FR * a.ld;
                const int col0 = unit == 0 ? tt : N - tt, cstep = unit == 0 ? G::TP : -G::TP;
#pragma unroll
                for (int k3 = 0; k3 < 8; k3++) {
                    const double2 o = So[k3 * G::TP + tt];
                    const double2 X0 = unit == 0 ? keep[k3] : o, X1 = unit == 0 ? o : keep[k3];
                    rowA[col0 + cstep * k3] = cadd(X0, X1);
                    rowB[col0 + cstep * k3] = cmul(csub(X0, X1), wy);
                }
                if (unit == 0 && tt == 0) {
                    const double2 X0 = xnh, X1 = So[8 * G::TP];
                    rowA[NH] = cadd(X0, X1);
                    rowB[NH] = cmul(csub(X0, X1), wy);
                }
                if (unit == 0 && tt >= 1 && tt < 16) {                        // pad columns N+1 .. N+15
                    rowA[N + tt] = make_double2(0.0, 0.0);
                    rowB[N + tt] = make_double2(0.0, 0.0);
                }
                __syncthreads();  // the other unit has read my parked half: the next plane's stage 1 may overwrite L
                continue;
            }
            if constexpr (WIDE) {
                // W^k = W^tt * (W^TP)^k3 with W = exp(2 pi i / 2N): table entry 2 tt, then constant rotations by 2 pi k3 / 32
                const double2 wb = a.tw[(size_t)tt << (TW_LOG2 - LOG2N - 1)];
                double2* outm = out0 + N - tt;  // X[N - k]
#pragma unroll
                for (int k3 = 0; k3 < 8; k3++) {
                    constexpr double C32[8] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                                               0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785};
                    constexpr double S32[8] = {0.0, 0.19509032201612826785, 0.38268343236508977173, 0.55557023301960222474,
                                               0.70710678118654752440, 0.83146961230254523708, 0.92387953251128675613, 0.98078528040323044913};
                    const double2 z = x[oidx<16>(k3)];
                    const double2 zn = Lp[-G::TP * k3];
                    const double2 E = make_double2(z.x + zn.x, z.y - zn.y);   // (z, zn hold Z / 2)
                    const double2 O = make_double2(z.y + zn.y, zn.x - z.x);
                    const double2 w = k3 == 0 ? wb : cmulc<+1>(wb, C32[k3], S32[k3]);
                    const double2 t = cmul(w, O);
                    out0[tt + G::TP * k3] = make_double2(E.x + t.x, E.y + t.y);       // X[k]
                    outm[-G::TP * k3] = make_double2(E.x - t.x, t.y - E.y);           // X[N-k] = conj(E - t)
                }
                if (tt == 0) out0[NH] = make_double2(2.0 * x[oidx<16>(8)].x, 2.0 * x[oidx<16>(8)].y);  // X[N/2] = Z[N/2]
                if (tt >= 1 && tt < 16) out0[N + tt] = make_double2(0.0, 0.0);        // pad columns N+1 .. N+15
            } else {
#pragma unroll
                for (int k3 = 0; k3 < 8; k3++) {
                    const int k = tt + G::TP * k3;
                    const double2 z = x[oidx<16>(k3)];
                    const double2 zn = Lp[-G::TP * k3];  // Z[N-k] = L[(N-k) - N/2]; k = 0 reads L[NH] = Z[0]
                    const double2 F0 = make_double2(z.x + zn.x, z.y - zn.y);   // (Z[k] + conj Z[N-k]) / 2  (z, zn hold Z / 2)
                    const double2 F1 = make_double2(z.y + zn.y, zn.x - z.x);   // (Z[k] - conj Z[N-k]) / 2i
                    if constexpr (FOLD) {  // rows y0 and y0 + 4096 of the plane: A and B (an absent second row is zero)
                        out0[k] = nrows == 2 ? cadd(F0, F1) : F0;
                        out1[k] = cmul(nrows == 2 ? csub(F0, F1) : F0, wy);
                    } else {
                        out0[k] = F0;
                        if (nrows == 2) out1[k] = F1;
                    }
                }
                if (tt < 16) {  // Nyquist column (real) and the zero pad columns N/2+1 .. N/2+15
                    const double2 z8 = make_double2(2.0 * x[oidx<16>(8)].x, 2.0 * x[oidx<16>(8)].y);  // Z[N/2]
                    if constexpr (FOLD) {
                        const double f1 = nrows == 2 ? z8.y : 0.0;
                        out0[NH + tt] = tt == 0 ? make_double2(z8.x + f1, 0.0) : make_double2(0.0, 0.0);
                        out1[NH + tt] = tt == 0 ? make_double2((z8.x - f1) * wy.x, (z8.x - f1) * wy.y) : make_double2(0.0, 0.0);
                    } else {
                        out0[NH + tt] = tt == 0 ? make_double2(z8.x, 0.0) : make_double2(0.0, 0.0);
                        if (nrows == 2) out1[NH + tt] = tt == 0 ? make_double2(z8.y, 0.0) : make_double2(0.0, 0.0);
                    }
                }
            }
            unit_bar(bar_id, G::UT);  // L is rewritten by the next plane's stage 1
        }
    }
}

// ---- inverse: 2 x 3 half-spectrum rows -> two u8 rows (C2R + scale/crop/centre/round/clamp) ----
// from_planes_u8's clamp8 (S:389): (uint8_t)max(0, min(255, round(v))) with round() = half away from zero.
// For v >= 0, round(v) == trunc(v + pred(0.5)) exactly (pred(0.5) = 0.5 - 2^-54: a tie still rounds up to the
// next integer, anything below the tie cannot reach it); negative v clamp to 0 and the unsigned conversion
// saturates them (and NaN) to 0.  Three instructions instead of round/min/max/cast.
__device__ __forceinline__ uint8_t clamp8_fast(double v) {
    const unsigned u = __double2uint_rz(v + 0.49999999999999994);
    return (uint8_t)min(u, 255u);
}

// stage 2 + exchange 2 through the split re/im buffer X (8 B entries, xpos layout) + stage 3, VEC = 1
template <int S, int LOG2N>
__device__ __forceinline__ void stage23_x(double* X, int tt, double2 w1, double2* x, int bar_id) {
    using G = Geo<LOG2N, 1>;
    constexpr bool PD = G::R1 == 16;
    const int k1 = tt >> 4, m = tt & 15;
    dft<S, 16>(x);
    twiddle<16>(x, w1);
    const int k1r = tt & (G::R1 - 1), k2r = tt / G::R1;
    [[maybe_unused]] double* Xw = X + xpos<G::R1>(k1, 0, m);
    [[maybe_unused]] const double* Xr = X + xpos<G::R1>(k1r, k2r, 0);
    double2 z[16];
#pragma unroll
    for (int k2 = 0; k2 < 16; k2++) {
        if constexpr (PD) Xw[k2 << 4] = x[oidx<16>(k2)].x; else X[xpos<G::R1>(k1, k2, m)] = x[oidx<16>(k2)].x;
    }
    unit_bar(bar_id, G::UT);
#pragma unroll
    for (int n = 0; n < 16; n++) {
        if constexpr (PD) z[n].x = Xr[n]; else z[n].x = X[xpos<G::R1>(k1r, k2r, n)];
    }
    unit_bar(bar_id, G::UT);
#pragma unroll
    for (int k2 = 0; k2 < 16; k2++) {
        if constexpr (PD) Xw[k2 << 4] = x[oidx<16>(k2)].y; else X[xpos<G::R1>(k1, k2, m)] = x[oidx<16>(k2)].y;
    }
    unit_bar(bar_id, G::UT);
#pragma unroll
    for (int n = 0; n < 16; n++) {
        if constexpr (PD) z[n].y = Xr[n]; else z[n].y = X[xpos<G::R1>(k1r, k2r, n)];
    }
    dft<S, 16>(z);
#pragma unroll
    for (int n = 0; n < 16; n++) x[n] = z[n];
}

template <int LOG2N>
struct C2RGeo {
    static constexpr int N = 1 << LOG2N;
    static constexpr size_t LB = (size_t)(N + 2) * 16 + 32;              // landing: A half then B half; exchange 1 reuses [0,N)
    static constexpr size_t XB = (size_t)(N + xpad<LOG2N>()) * 8;        // split exchange-2 buffer
    static constexpr size_t UNIT = LB + ((XB + 15) & ~(size_t)15);
};

// WIDE (LOG2N = 12 only): ONE half-spectrum row X[0..N] of a 2N = 8192-pixel image row per item.  With
// E = (X[k] + conj X[N-k])/2 and O = W^-k (X[k] - conj X[N-k])/2 (W = exp(+2 pi i / 2N)) the N-point inverse transform of
// Z = E + iO returns the even samples in the real part and the odd samples in the imaginary part.
template <int LOG2N, int UNITS, bool CENTER, bool WIDE = false>
__global__ void __launch_bounds__(Geo<LOG2N, 1>::UT* UNITS, 1) pencil_u8_inv_c2r(R2CArgs a) {
    using G = Geo<LOG2N, 1>;
    using CG = C2RGeo<LOG2N>;
    static_assert(!WIDE || LOG2N == 12, "the wide variant unpacks a 4096-point transform into an 8192-pixel row");
    constexpr int N = G::N, NH = N / 2, HL = NH + 1;  // HL entries per half row
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int unit = threadIdx.x / G::UT, tt = threadIdx.x % G::UT;
    unsigned char* base = smem_raw + (size_t)unit * CG::UNIT;
    double2* L = (double2*)base;
    double* X = (double*)(base + CG::LB);
    const int bar_id = 1 + unit;
    const long long stride = (long long)gridDim.x * UNITS;
    long long item = (long long)blockIdx.x * UNITS + unit;
    const int HP = WIDE ? a.H : (a.H + 1) / 2;
    const int YS = WIDE ? 1 : 2;
    const size_t row_bytes = (size_t)a.W * 3;
    const double scale = (WIDE ? 0.5 : 1.0) / (double)N;  // S:357 (wide: the halves of E and O folded in)
    ThreadTw<-1, LOG2N, 1> ttw;
    ttw.load(a.tw, tt);

    auto src_rows = [&](long long it, int ch) -> const double2* {
        const long long img = it / HP;
        const int y0 = YS * (int)(it % HP);
        return a.spec + (((size_t)img * 3 + ch) * a.PH + y0) * a.ld;
    };
    auto issue_loads = [&](long long it, int ch) {
        const double2* A = src_rows(it, ch);
        if constexpr (WIDE) {
            for (int e = tt; e < N + 1; e += G::UT) cp_async16(&L[e], A + e, 16);
        } else {
            const bool two = 2 * (int)(it % HP) + 1 < a.H;  // odd-H tail: the second row is absent (zero)
            for (int e = tt; e < 2 * HL; e += G::UT) {
                const bool isB = e >= HL;
                const double2* src = isB ? A + a.ld + (e - HL) : A + e;
                const bool valid = !isB || two;
                cp_async16(&L[e], valid ? (const void*)src : (const void*)a.spec, valid ? 16 : 0);
            }
        }
        cp_async_commit();
    };

    if (unit & 1) {
        const long long t0 = clock64();
        while (clock64() - t0 < a.stagger) {}
    }
    if (item < a.nitems) issue_loads(item, 0);
    for (; item < a.nitems; item += stride) {
        const long long img = item / HP;
        const int y0 = YS * (int)(item % HP);
        const bool two = !WIDE && y0 + 1 < a.H;
        uint8_t* o0 = a.img_out + ((size_t)img * a.H + y0) * row_bytes + (size_t)tt * (WIDE ? 6 : 3);
        uint8_t* o1 = o0 + row_bytes;
        for (int ch = 0; ch < 3; ch++) {
            cp_async_wait_all();
            unit_bar(bar_id, G::UT);
            // ---- gather Z[k] = A[k] + i B[k] (k <= N/2) or conj(A[N-k]) + i conj(B[N-k]) (k > N/2); k = 256 n + m:
            // blocks n < R1/2 take the first form, n > R1/2 the second, block R1/2 the first only at k = N/2 (m = 0)
            double2 x[16];
#pragma unroll
            for (int j = 0; j < G::J1; j++) {
                const int m = tt + j * G::TP;
                const double2* Lf = L + m;        // form 1: A = Lf[256 n], B = Lf[HL + 256 n]
                const double2* Lb = L + N - m;    // form 2: A = Lb[-256 n], B = Lb[HL - 256 n]
                [[maybe_unused]] const double smid = (m == 0) ? 1.0 : -1.0;
                if constexpr (WIDE) {
                    // cos / sin of 2 pi n / 32: W^-k = conj(W^m) * conj(W^(256 n))
                    constexpr double C32[16] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                                                0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785,
                                                0.0, -0.19509032201612826785, -0.38268343236508977173, -0.55557023301960222474,
                                                -0.70710678118654752440, -0.83146961230254523708, -0.92387953251128675613, -0.98078528040323044913};
                    constexpr double S32[16] = {0.0, 0.19509032201612826785, 0.38268343236508977173, 0.55557023301960222474,
                                                0.70710678118654752440, 0.83146961230254523708, 0.92387953251128675613, 0.98078528040323044913,
                                                1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                                                0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785};
                    double2 wb = a.tw[(size_t)m << (TW_LOG2 - LOG2N - 1)];
                    wb.y = -wb.y;  // conj(W^m)
#pragma unroll
                    for (int n = 0; n < G::R1; n++) {
                        const double2 A = Lf[256 * n], B = Lb[-256 * n];                         // X[k], X[N-k]
                        const double2 E = make_double2(A.x + B.x, A.y - B.y);                     // X[k] + conj X[N-k]
                        const double2 D = make_double2(A.x - B.x, A.y + B.y);                     // X[k] - conj X[N-k]
                        const double2 w = n == 0 ? wb : cmulc<-1>(wb, C32[n], S32[n]);
                        const double2 O = cmul(w, D);
                        x[j * G::R1 + n] = make_double2(E.x - O.y, E.y + O.x);                    // E + iO
                    }
                } else {
#pragma unroll
                    for (int n = 0; n < G::R1; n++) {
                        if (n < G::R1 / 2) {
                            const double2 A = Lf[256 * n], B = Lf[HL + 256 * n];
                            x[j * G::R1 + n] = make_double2(A.x - B.y, A.y + B.x);
                        } else if (n > G::R1 / 2) {
                            const double2 A = Lb[-256 * n], B = Lb[HL - 256 * n];
                            x[j * G::R1 + n] = make_double2(A.x + B.y, B.x - A.y);
                        } else {  // k = N/2 + m: element N/2 - m of both halves; m = 0 is the Nyquist bin itself (first form)
                            const double2 A = Lb[-256 * n], B = Lb[HL - 256 * n];
                            x[j * G::R1 + n] = make_double2(fma(-smid, B.y, A.x), fma(smid, A.y, B.x));
                        }
                    }
                }
            }
            unit_bar(bar_id, G::UT);  // every thread has its inputs: exchange 1 may overwrite the landing area
#pragma unroll
            for (int j = 0; j < G::J1; j++) {
                const int m = tt + j * G::TP;
                double2* xj = x + j * G::R1;
                dft<-1, G::R1>(xj);
                double2 w1 = a.tw[(size_t)m << (TW_LOG2 - LOG2N)];
                w1.y = -w1.y;
                twiddle<G::R1>(xj, w1);
#pragma unroll
                for (int k = 0; k < G::R1; k++) L[k * 256 + m] = xj[oidx<G::R1>(k)];
            }
            unit_bar(bar_id, G::UT);
            stage2_load<LOG2N, 1>(L, tt, 0, x);
            unit_bar(bar_id, G::UT);  // L is free: prefetch the next plane / the next item's first plane
            if (ch < 2) issue_loads(item, ch + 1);
            else if (item + stride < a.nitems) issue_loads(item + stride, 0);
            stage23_x<-1, LOG2N>(X, tt, ttw.s2v(), x, bar_id);
            // ---- epilogue (ifft_crop S:399, apply_center S:1102, from_planes_u8 S:387): Re -> row y0, Im -> row y0+1
#pragma unroll
            for (int k3 = 0; k3 < 16; k3++) {
                const int k = tt + G::TP * k3;
                double v0 = x[oidx<16>(k3)].x * scale, v1 = x[oidx<16>(k3)].y * scale;
                if constexpr (WIDE) {  // Re -> pixel 2k, Im -> pixel 2k+1 of the one row
                    if constexpr (CENTER) {
                        if (y0 & 1) v0 = -v0; else v1 = -v1;
                    }
                    if (2 * k < a.W) o0[G::TP * k3 * 6 + ch] = clamp8_fast(v0);
                    if (2 * k + 1 < a.W) o0[G::TP * k3 * 6 + 3 + ch] = clamp8_fast(v1);
                } else if (k < a.W) {
                    if constexpr (CENTER) {
                        if ((k + y0) & 1) v0 = -v0; else v1 = -v1;
                    }
                    o0[G::TP * k3 * 3 + ch] = clamp8_fast(v0);
                    if (two) o1[G::TP * k3 * 3 + ch] = clamp8_fast(v1);
                }
            }
        }
    }
}

}  // namespace pk

// ------------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------------
namespace {

template <typename K>
cudaError_t set_smem(K kern, size_t bytes) {
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

inline unsigned grid_for(const Launcher& L, long long nitems, int units) {
    long long ctas = (nitems + units - 1) / units;
    if (ctas > L.sm_count) ctas = L.sm_count;
    if (ctas < 1) ctas = 1;
    return (unsigned)ctas;
}

template <int S, int LOG2N, int VEC, int MODE, int UNITS>
cudaError_t run_c2c(const Launcher& L, const PassArgs& p) {
    using G = pk::Geo<LOG2N, VEC>;
    pk::C2CArgs a;
    a.spec = p.spec; a.tw = p.tw; a.PW = p.PW; a.PH = p.PH; a.in_rows = p.in_rows; a.out_rows = p.out_rows;
    a.nitems = MODE == pk::M_C2C_ROW ? (long long)p.nplanes * p.PH : (long long)p.nplanes * (p.PW / VEC);
    const size_t smem = (G::L_BYTES + G::X_BYTES) * UNITS;
    auto kern = pk::pencil_c2c<S, LOG2N, VEC, MODE, UNITS>;
    cudaError_t e = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    kern<<<grid_for(L, a.nitems, UNITS), G::UT * UNITS, smem, L.stream>>>(a);
    if (L.launch_counter) ++*L.launch_counter;
    return cudaGetLastError();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encoder() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

// planes viewed as doubles: [nplanes][rows][2*PW]; box = {2*VEC doubles, 256 rows, 1 plane}
bool make_col_map(CUtensorMap* m, const double2* spec, int nplanes, int PH, int PW, int rows, int vec) {
    EncodeTiledFn enc = get_encoder();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)2 * PW, (cuuint64_t)rows, (cuuint64_t)nplanes};
    cuuint64_t strides[2] = {(cuuint64_t)PW * 16, (cuuint64_t)PH * PW * 16};
    cuuint32_t box[3] = {(cuuint32_t)(2 * vec), 256, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)spec, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// store map of pencil_col_tma_w (N = 4096, VEC = 2): dims (fastest first) column doubles, k2, k3, k1, plane with
// row = k1 + 16 k2 + 256 k3; the k3 extent clips the rows that are not needed (multiples of 256 rows)
bool make_col_store_map5(CUtensorMap* m, const double2* spec, int nplanes, int PH, int PW, int out_rows, int k3box = 8) {
    EncodeTiledFn enc = get_encoder();
    if (!enc) return false;
    const cuuint64_t rowb = (cuuint64_t)PW * 16;
    int k3n = (out_rows + 255) / 256;
    if (k3n > 16) k3n = 16;
    if (k3n < 1) k3n = 1;
    cuuint64_t dims[5] = {(cuuint64_t)2 * PW, 16, (cuuint64_t)k3n, 16, (cuuint64_t)nplanes};
    cuuint64_t strides[4] = {16 * rowb, 256 * rowb, rowb, (cuuint64_t)PH * rowb};
    cuuint32_t box[5] = {4, 16, (cuuint32_t)k3box, 16, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, (void*)spec, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// column groups of `vec` columns a pass has to transform (PassArgs::col_limit trims the right-hand side of the plane)
inline int col_groups(const PassArgs& p, int vec) {
    const int all = p.PW / vec;
    if (p.col_limit <= 0) return all;
    const int need = (p.col_limit + vec - 1) / vec;
    return need < all ? need : all;
}

template <int S, int NZ, int K3N, bool SIGN = false>
cudaError_t run_col_tma_w(const Launcher& L, const PassArgs& p, bool* ok) {
    using G = pk::Geo<12, 2>;
    CUtensorMap in_map, out_map;
    *ok = make_col_map(&in_map, p.spec, p.nplanes, p.PH, p.PW, p.in_rows, 2) &&
          make_col_store_map5(&out_map, p.spec, p.nplanes, p.PH, p.PW, p.out_rows);
    if (!*ok) return cudaSuccess;
    pk::ColTmaArgs a;
    a.tw = p.tw; a.groups_per_plane = col_groups(p, 2); a.nitems = (long long)p.nplanes * a.groups_per_plane;
    a.sample_q = S > 0 ? p.sample_q : nullptr;
    a.sample_stride = p.sample_stride;
    a.sample_groups = (int)(p.sample_stride ? (p.PW - 16) / 2 : 0);  // p.PW is ld = PW_full/2 + 16 here: pairs below the Nyquist column
    a.signmap = SIGN ? p.signmap : nullptr; a.map_groups = p.PW / 2; a.alpha = p.sign_alpha;
    a.q32 = (S > 0 && K3N == 16 && !SIGN && p.col_limit == 0) ? p.q32 : nullptr;
    const size_t smem = G::L_BYTES + G::X_BYTES;
    auto kern = pk::pencil_col_tma_w<S, NZ, K3N, SIGN>;
    cudaError_t e = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    kern<<<grid_for(L, a.nitems, 1), 512, smem, L.stream>>>(in_map, out_map, a);
    if (L.launch_counter) ++*L.launch_counter;
    return cudaGetLastError();
}

template <int NZ, int K3N>
cudaError_t run_col_embed_w(const Launcher& L, const PassArgs& p, bool* ok) {
    using G = pk::Geo<12, 2>;
    CUtensorMap in_map, out_map, out_map2;
    constexpr int K3B = K3N > 8 ? K3N - 8 : 0;
    *ok = make_col_map(&in_map, p.spec, p.nplanes, p.PH, p.PW, p.in_rows, 2) &&
          make_col_store_map5(&out_map, p.spec, p.nplanes, p.PH, p.PW, p.out_rows) &&
          make_col_store_map5(&out_map2, p.spec, p.nplanes, p.PH, p.PW, p.out_rows, (K3B >= 1 && K3B <= 2) ? K3B : 8);
    if (!*ok) return cudaSuccess;
    pk::ColEmbedArgs a;
    a.tw = p.tw; a.groups_per_plane = p.PW / 2; a.nitems = (long long)p.nplanes * a.groups_per_plane;
    a.sample_q = p.sample_q; a.sample_stride = p.sample_stride;
    a.sample_groups = (int)(p.sample_stride ? (p.PW - 16) / 2 : 0);  // p.PW is ld = PW_full/2 + 16: pairs below the Nyquist column
    a.qhi = p.qhi; a.qlo = p.qlo; a.pres = p.embed_pres; a.val = p.embed_val; a.k3max = p.embed_k3max;
    a.pair_has_bins = p.embed_pres ? (const uint8_t*)p.embed_pres + embed_mask_bytes(p.PW, 3) : nullptr;  // (p.PW is ld here)
    a.cos_a = p.embed_cos; a.sin_a = p.embed_sin;
    const size_t smem = G::L_BYTES + pk::X_EMBED_BYTES;
    auto kern = pk::pencil_col_embed_w<NZ, K3N>;
    cudaError_t e = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    kern<<<grid_for(L, a.nitems, 1), 512, smem, L.stream>>>(in_map, out_map, out_map2, a);
    if (L.launch_counter) ++*L.launch_counter;
    return cudaGetLastError();
}

template <int S, int LOG2N, int VEC>
cudaError_t run_col_tma(const Launcher& L, const PassArgs& p, bool* ok) {
    using G = pk::Geo<LOG2N, VEC>;
    CUtensorMap in_map, out_map;
    *ok = make_col_map(&in_map, p.spec, p.nplanes, p.PH, p.PW, p.in_rows, VEC) &&
          make_col_map(&out_map, p.spec, p.nplanes, p.PH, p.PW, p.out_rows, VEC);
    if (!*ok) return cudaSuccess;
    pk::ColTmaArgs a;
    a.tw = p.tw; a.groups_per_plane = col_groups(p, VEC); a.nitems = (long long)p.nplanes * a.groups_per_plane;
    a.sample_q = nullptr; a.sample_stride = 0; a.sample_groups = 0;
    a.signmap = nullptr; a.map_groups = 0; a.alpha = 0.0; a.q32 = nullptr;
    const size_t smem = G::L_BYTES + G::X_BYTES;
    auto kern = pk::pencil_col_tma<S, LOG2N, VEC>;
    cudaError_t e = set_smem(kern, smem);
    if (e != cudaSuccess) return e;
    kern<<<grid_for(L, a.nitems, 1), 512, smem, L.stream>>>(in_map, out_map, a);
    if (L.launch_counter) ++*L.launch_counter;
    return cudaGetLastError();
}

template <int LOG2N, int UNITS, bool INV>
cudaError_t run_u8(const Launcher& L, const PassArgs& p) {
    using G = pk::Geo<LOG2N, 1>;
    pk::U8Args a;
    a.spec = p.spec; a.tw = p.tw; a.img_in = p.img_in; a.img_out = p.img_out;
    a.W = p.W; a.H = p.H; a.PW = p.PW; a.PH = p.PH; a.center = p.center;
    a.nitems = (long long)(p.nplanes / 3) * p.H;
    constexpr size_t UB = pk::U8Geo<LOG2N>::U_BYTES;
    const size_t smem = (INV ? (G::L_BYTES + G::X_BYTES + UB) : (G::L_BYTES + 2 * UB)) * UNITS;
    cudaError_t e;
    if constexpr (INV) {
        auto kern = pk::pencil_u8_inv<LOG2N, UNITS>;
        if ((e = set_smem(kern, smem)) != cudaSuccess) return e;
        kern<<<grid_for(L, a.nitems, UNITS), G::UT * UNITS, smem, L.stream>>>(a);
    } else {
        auto kern = pk::pencil_u8_fwd<LOG2N, UNITS>;
        if ((e = set_smem(kern, smem)) != cudaSuccess) return e;
        kern<<<grid_for(L, a.nitems, UNITS), G::UT * UNITS, smem, L.stream>>>(a);
    }
    if (L.launch_counter) ++*L.launch_counter;
    return cudaGetLastError();
}

template <int LOG2N, int UNITS, bool INV, bool CENTER, bool WIDE, bool FOLD = false>
cudaError_t run_r2c_c(const Launcher& L, const pk::R2CArgs& a) {
    using G = pk::Geo<LOG2N, 1>;
    const size_t smem = (INV ? pk::C2RGeo<LOG2N>::UNIT : pk::R2CGeo<LOG2N>::FWD_UNIT) * UNITS;
    cudaError_t e;
    if constexpr (INV) {
        auto kern = pk::pencil_u8_inv_c2r<LOG2N, UNITS, CENTER, WIDE>;
        if ((e = set_smem(kern, smem)) != cudaSuccess) return e;
        kern<<<grid_for(L, a.nitems, UNITS), G::UT * UNITS, smem, L.stream>>>(a);
    } else {
        auto kern = pk::pencil_u8_fwd_r2c<LOG2N, UNITS, CENTER, WIDE, FOLD>;
        if ((e = set_smem(kern, smem)) != cudaSuccess) return e;
        // (folded wide rows: one item per CTA at a time, its two units take the two rows of the item)
        kern<<<grid_for(L, a.nitems, (WIDE && FOLD) ? 1 : UNITS), G::UT * UNITS, smem, L.stream>>>(a);
    }
    if (L.launch_counter) ++*L.launch_counter;
    return cudaGetLastError();
}

// WIDE: p.PW = 8192 runs on the 4096-point kernels (LOG2N = 12), one image row per item
// FOLD: PassArgs::fold (forward rows of an 8192-row plane: rows y and y + 4096 leave as A_y and B_y, see R2CArgs)
template <int LOG2N, int UNITS, bool INV, bool WIDE = false, bool FOLD = false>
cudaError_t run_r2c(const Launcher& L, const PassArgs& p) {
    pk::R2CArgs a;
    a.spec = p.spec; a.tw = p.tw; a.img_in = p.img_in; a.img_out = p.img_out;
    a.W = p.W; a.H = p.H; a.PW = p.PW; a.PH = p.PH; a.ld = p.ld; a.center = p.center;
    a.nitems = (long long)(p.nplanes / 3) * (FOLD ? 4096 : (WIDE ? p.H : (p.H + 1) / 2));
    a.stagger = pk::STAGGER_CYCLES;  // (measured: 0 .. 12000 cycles make no difference to either row kernel, profiles/r2_experiments.txt)
    return p.center ? run_r2c_c<LOG2N, UNITS, INV, true, WIDE, FOLD>(L, a) : run_r2c_c<LOG2N, UNITS, INV, false, WIDE, FOLD>(L, a);
}

// per-size unit counts: rows: (N*24 B) per unit, columns: VEC so that one unit fills ~192 KB
template <int LOG2N>
struct Cfg;
template <> struct Cfg<12> { static constexpr int ROW_UNITS = 2, COL_VEC = 2, COL_UNITS = 1, U8F_UNITS = 2, U8I_UNITS = 2, TMA_VEC = 2; };
template <> struct Cfg<11> { static constexpr int ROW_UNITS = 4, COL_VEC = 4, COL_UNITS = 1, U8F_UNITS = 4, U8I_UNITS = 4, TMA_VEC = 4; };
template <> struct Cfg<10> { static constexpr int ROW_UNITS = 8, COL_VEC = 4, COL_UNITS = 2, U8F_UNITS = 8, U8I_UNITS = 8, TMA_VEC = 8; };
template <> struct Cfg<9>  { static constexpr int ROW_UNITS = 8, COL_VEC = 4, COL_UNITS = 4, U8F_UNITS = 8, U8I_UNITS = 8, TMA_VEC = 16; };

template <int LOG2N>
cudaError_t dispatch(const Launcher& L, const PassArgs& p) {
    using C = Cfg<LOG2N>;
    if (p.fold) {
        if (!(p.half && p.img_in && p.PH == 8192 && p.H > 4096)) return cudaErrorNotSupported;
        return run_r2c<LOG2N, C::U8F_UNITS, false, false, true>(L, p);
    }
    if (p.half && p.img_in) return run_r2c<LOG2N, C::U8F_UNITS, false>(L, p);
    if (p.half && p.img_out) return run_r2c<LOG2N, C::U8I_UNITS, true>(L, p);
    if (p.img_in) return run_u8<LOG2N, C::U8F_UNITS, false>(L, p);
    if (p.img_out) return run_u8<LOG2N, C::U8I_UNITS, true>(L, p);
    if (p.axis == 0)
        return p.inverse ? run_c2c<-1, LOG2N, 1, pk::M_C2C_ROW, C::ROW_UNITS>(L, p)
                         : run_c2c<+1, LOG2N, 1, pk::M_C2C_ROW, C::ROW_UNITS>(L, p);
    if constexpr (LOG2N == 12) {  // warp-local exchange + permuting store
        if (p.fused_embed) {  // forward + phase write + inverse in one residency (the caller checked fused_embed_supported)
            if (L.fft_impl != 1 || !p.qhi || !p.qlo || p.PW < 2 || p.in_rows != p.out_rows) return cudaErrorNotSupported;
            bool ok = false;
            // (a 4096-row plane has more than 2048 image rows: 9 row blocks for UHD, else all 16)
            cudaError_t e = p.in_rows <= 9 * 256 ? run_col_embed_w<9, 9>(L, p, &ok) : run_col_embed_w<16, 16>(L, p, &ok);
            return (e == cudaSuccess && !ok) ? cudaErrorNotSupported : e;
        }
        if (p.signmap) {  // extract without jitter: the pass only leaves the read bits behind (the caller checked signmap_supported)
            if (L.fft_impl != 1 || p.inverse || p.out_rows > 8 * 256 || p.PW < 2) return cudaErrorNotSupported;
            bool ok = false;
            cudaError_t e = p.in_rows <= 9 * 256 ? run_col_tma_w<+1, 9, 8, true>(L, p, &ok) : run_col_tma_w<+1, 16, 8, true>(L, p, &ok);
            return (e == cudaSuccess && !ok) ? cudaErrorNotSupported : e;
        }
        if (p.PW >= 2 && p.PW % 2 == 0) {
            bool ok = false;
            // zero structure of a padded image (UHD: 2160 of 4096 rows): 9 of 16 row blocks carry data
            // forward pass of an extract: only the rows that hold bins are kept (the default annulus ends at row 0.45 * 4096)
            const bool few_in = p.in_rows <= 9 * 256, few_out = p.out_rows <= 9 * 256, half_out = p.out_rows <= 8 * 256;
            cudaError_t e = p.inverse ? (few_out ? run_col_tma_w<-1, 16, 9>(L, p, &ok) : run_col_tma_w<-1, 16, 16>(L, p, &ok))
                            : half_out ? (few_in ? run_col_tma_w<+1, 9, 8>(L, p, &ok) : run_col_tma_w<+1, 16, 8>(L, p, &ok))
                                       : (few_in ? run_col_tma_w<+1, 9, 16>(L, p, &ok) : run_col_tma_w<+1, 16, 16>(L, p, &ok));
            if (e != cudaSuccess || ok) return e;
        }
    }
    if (p.PW >= C::TMA_VEC && p.PW % C::TMA_VEC == 0) {  // (the cp.async / STG column kernel below only runs where no tensor map can be had)
        bool ok = false;
        cudaError_t e = p.inverse ? run_col_tma<-1, LOG2N, C::TMA_VEC>(L, p, &ok) : run_col_tma<+1, LOG2N, C::TMA_VEC>(L, p, &ok);
        if (e != cudaSuccess || ok) return e;
    }
    if (p.PW % C::COL_VEC) return cudaErrorNotSupported;  // (the caller falls back to the generic kernel)
    return p.inverse ? run_c2c<-1, LOG2N, C::COL_VEC, pk::M_C2C_COL, C::COL_UNITS>(L, p)
                     : run_c2c<+1, LOG2N, C::COL_VEC, pk::M_C2C_COL, C::COL_UNITS>(L, p);
}

}  // namespace

bool signmap_supported(const Launcher& L) { return L.fft_impl == 1 && get_encoder() != nullptr; }

bool fused_embed_supported(const Launcher& L) { return signmap_supported(L); }

cudaError_t launch_fft_pass_pencil(const Launcher& L, const PassArgs& p, bool* handled) {
    *handled = true;
    if ((p.signmap || p.fused_embed) && !(p.log2n == 12 && p.axis == 1)) return cudaErrorNotSupported;
    if (p.fold && (p.axis != 0 || p.log2n < 9 || p.log2n > 13)) return cudaErrorNotSupported;
    // the fused u8 passes need W <= PW == N (always true) and run along x only
    switch (p.log2n) {
        case 12: return dispatch<12>(L, p);
        case 11: return dispatch<11>(L, p);
        case 10: return dispatch<10>(L, p);
        case 9: return dispatch<9>(L, p);
        case 13:  // 8192-pixel rows of a half-spectrum workspace: packed into the 4096-point fused u8 kernels
            if (p.fold) {
                if (!(p.half && p.axis == 0 && p.img_in && p.PH == 8192 && p.H > 4096)) return cudaErrorNotSupported;
                return run_r2c<12, Cfg<12>::U8F_UNITS, false, true, true>(L, p);
            }
            if (p.half && p.axis == 0 && p.img_in) return run_r2c<12, Cfg<12>::U8F_UNITS, false, true>(L, p);
            if (p.half && p.axis == 0 && p.img_out) return run_r2c<12, Cfg<12>::U8I_UNITS, true, true>(L, p);
            *handled = false;
            return cudaSuccess;
        default: *handled = false; return cudaSuccess;
    }
}

}  // namespace tfft
