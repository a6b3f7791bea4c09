// tfft_kernels.cuh -- device-side interface of the TurtleFFT hot path (sm_100a).
// Reference citations S:n = steganosaurus/src/steganosaur.cpp line n.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tfft {

// Twiddle table: tw[k] = exp(+2*pi*i*k / TW_N), k in [0, TW_N/2).  The reference's forward
// transform uses the +i sign (S:347); inverse passes conjugate on the fly.
constexpr int TW_LOG2 = 14;  // 16384 = TFFT_MAX_DIM
constexpr int TW_N = 1 << TW_LOG2;

enum PassIn { IN_C64 = 0, IN_U8 = 1 };
enum PassOut { OUT_C64 = 0, OUT_U8 = 1 };

// One 1-D FFT pass over a batch of planes.  A "pencil" is one row (axis 0) or one column
// (axis 1) of one padded plane; planes are [PH][PW] complex<double>, nplanes = 3 * nimages.
struct PassArgs {
    double2* spec;           // [nplanes][PH][PW]
    const uint8_t* img_in;   // IN_U8 : [n][H][W][3]  (to_planes_u8 S:383 + pad_to_fft S:393 fused)
    uint8_t* img_out;        // OUT_U8: [n][H][W][3]  (ifft_crop S:399 + from_planes_u8 S:387 fused)
    const double2* tw;       // twiddle table (device)
    const uint64_t* usable;  // OUT_U8 only, may be null (unused by the pass itself)
    int nplanes;
    int W, H, PW, PH;
    int log2n;    // pencil length = 1 << log2n
    int axis;     // 0: along x (rows), 1: along y (columns)
    int inverse;  // 0: e^{+i} (reference forward), 1: e^{-i} and scale by 1/n (S:357)
    int center;   // apply_center S:392 on the image side of IN_U8 / OUT_U8
    // Zero-structure hints in plane-row coordinates (never change results, only skip work):
    int in_rows;  // plane rows y >= in_rows are all-zero on input and are NOT read
                  //   (axis 0: such pencils are skipped; axis 1: those elements are zero-filled)
    int out_rows; // plane rows y >= out_rows are not needed downstream and are NOT written
    int col_limit; // axis 1 only, 0 = all: columns x >= col_limit (rounded up to the kernel's column group) are not needed
                   //   downstream and are neither read nor transformed (extract: the bins sit in a corner of the plane)
    // Half-spectrum mode (real planes): the workspace keeps columns 0..PW/2 with row stride ld =
    // PW/2 + 16.  u8 passes get half = 1 (PW = full width, ld = stride); column passes simply see a
    // plane of PW := ld columns.
    int half;
    int ld;
    // Scratch plane batch of the same size as spec: needed by passes longer than 4096 points
    // (four-step scheme: strided 4096-point sub-transforms in place, radix-2/4 combine through tmp).
    double2* tmp;
    int leave_in_tmp;  // four-step passes only: skip the copy back, the result stays in tmp (the caller ping-pongs)
    int fourstep_sub_only;  // four-step column passes only: stop after the in-place sub-transforms (SpecLayout::fs_r readers)
    // Forward 4096-point column pass only (optional): the kernel drops a stratified sample of q = re^2 + im^2 (one
    // element per thread and column pair, as IEEE bit patterns) into sample_q[plane * sample_stride + ...] -- the
    // median bracket is then built from it without a separate gather pass over the spectrum.
    unsigned long long* sample_q;
    unsigned sample_stride;
    // ... and (optional, with sample_q) q of EVERY element as a float, q32[plane * PH * PW + ...] in the kernel's own
    // warp order (float4 [g][k1][j][lane], tfft_pencil.cu): the median scan then reads 4 bytes per element instead of 16
    // and goes back to the spectrum only for the ~1 % of elements whose float cannot decide (median_scan_q32).
    float* q32;
    // Forward 4096-point column pass of an extract without jitter (optional, out_rows <= 2048): instead of the spectrum
    // the pass leaves one bit per element -- what read_bit_from_bin (S:734-746) reads there with this alpha -- in
    // signmap (sign_map_words() uint32 per plane, layout in tfft_pencil.cu); spec rows are NOT written.
    uint32_t* signmap;
    double sign_alpha;
    // Forward u8 row pass of a half-spectrum workspace with PH = 8192 (axis 0, img_in, H > 4096; extract only): rows y and
    // y + 4096 leave as A_y = F_y + F_{y+4096} (stored row y) and B_y = (F_y - F_{y+4096}) w_8192^y (stored row y + 4096),
    // the first radix-2 step of the column transform, so the column pass is two 4096-point passes per plane: the workspace
    // is then read as 2 * nplanes planes of 4096 rows whose column transforms hold rows 2 y' and 2 y' + 1 of the spectrum.
    int fold;
    // Column-resident embed (fused_embed = 1; axis 1, 4096-row half planes, in_rows = out_rows = H): forward column pass,
    // phase write (write_bit_on_bin S:712-732) and inverse column pass in one shared-memory residency (pencil_col_embed_w in
    // tfft_pencil.cu).  The spectrum is not stored; q = |F|^2 of every element leaves as two 32-bit planes qhi / qlo (the
    // exact double, layout of q32) next to the sample.  embed_pres [3][ld/2][512] / embed_val [nplanes][ld/2][512]: per
    // thread and column pair, bit k3 = a bin at row k1 + 16 m + 256 k3 / the bit to write there (embed_masks_* below).
    int fused_embed;
    const uint16_t* embed_pres;   // (followed by the pair flags, embed_pres_bytes)
    const uint16_t* embed_val;
    int embed_k3max;
    double embed_cos, embed_sin;
    uint32_t* qhi;
    uint32_t* qlo;
};
// uint32 words per plane of the sign map of a 4096-row plane with `cols` stored columns
inline size_t sign_map_words(int cols) { return (size_t)(cols / 2) * 16 * 8; }
// number of samples per plane the 4096-point column pass delivers for a half-spectrum plane of PW columns (0: none)
inline unsigned col_pass_samples(int PH, int PW_full, int half) { return (PH == 4096 && half) ? (unsigned)(PW_full / 4) * 256u : 0u; }

// How a plane's spectrum is stored.  full: [PH][PW], ld = PW.  half (real planes, Hermitian):
// columns 0..PW/2 only, row stride ld = PW/2 + 16 (pad columns are zero); element (y,x) with
// x > PW/2 is conj of the stored element ((PH-y)%PH, PW-x).
// fs_r > 0 (extract of a plane taller than 4096 rows): the column pass stopped after the sub-transforms of its four-step
// scheme (PH = fs_r * M), so stored row fs_r * k + r holds Y_r[k], the M-point transform of the rows r, r + fs_r, ...;
// readers evaluate X[y] = sum_r w_PH^{r y} Y_r[y mod M] at the bins they need (spec_load) instead of a combine pass
// over the whole plane.
struct SpecLayout {
    int PH, PW, ld, half;
    int fs_r;
    const double2* fs_tw;
    __host__ __device__ size_t plane_elems() const { return (size_t)PH * ld; }
};

struct Launcher {
    cudaStream_t stream;
    uint64_t* launch_counter;  // host-side counter of kernels launched
    int sm_count;
    size_t smem_optin;
    int fft_impl;  // 0 = v0 (simple shared-memory radix-2), 1 = pencil kernels
};

bool signmap_supported(const Launcher& L);  // PassArgs::signmap is available (TMA column kernel in use)
cudaError_t build_twiddles(double2* d_tw, cudaStream_t s);  // fills TW_N/2 entries
cudaError_t launch_fft_pass(const Launcher& L, const PassArgs& a);

// ---- median + capacity (median_abs S:404-409, count_plane S:999-1007) ----------------------
struct MedianWork {
    uint32_t* hist;      // [nplanes][2048]
    uint64_t* prefix;    // [nplanes] key prefix selected so far
    uint64_t* rank;      // [nplanes] remaining rank inside the prefix bucket
    uint64_t* cand;      // [nplanes][cand_cap] candidate keys (sample, then bracket members: each stored element once)
    uint32_t* cand_n;    // [nplanes]
    uint32_t cand_cap;
    uint64_t* cand_b;    // [nplanes][CAND_B_MAX] members whose multiplicity is below the interior weight (edge / pad columns
    uint32_t* cand_b_n;  // [nplanes]              of a half plane), one entry per unit of excess
    uint64_t* counts;    // [nplanes] below-bracket counts, then capacity counts before halving
    uint64_t* prefix2;   // [nplanes][2] median bracket (qlo, qhi as doubles)
    // capacity fused into the median scan: annulus bins certainly below magmin*median are counted,
    // the few whose verdict needs the exact median are staged (keys) and resolved afterwards
    uint64_t* cap_below; // [nplanes]
    uint64_t* cap_unc;   // [nplanes][CAP_UNC_MAX] keys of undecided annulus bins
    uint32_t* cap_unc_n; // [nplanes]
    uint64_t* ann_total; // [1] number of off-axis annulus bins of the geometry
    int* flags;          // [0] median fallback, [1] capacity fallback
};
constexpr uint32_t CAP_UNC_MAX = 1024;
constexpr uint32_t CAND_B_MAX = 1u << 15;
size_t median_work_bytes(int nplanes, uint32_t cand_cap);
void median_work_carve(MedianWork& w, void* base, int nplanes, uint32_t cand_cap);
// d_median: [nplanes]; d_usable: [nplanes/3] (sum over the 3 planes of count/2)
// presampled > 0: w.cand already holds that many q-keys per plane (written by the column pass), no gather pass
// q32 != null (4096-row half planes): float copy of q = |F|^2 of every element left by the column pass (PassArgs::q32)
cudaError_t launch_median_capacity(const Launcher& L, const double2* spec, int nplanes, SpecLayout lay,
                                   double magmin, double rlo, double rhi, MedianWork w,
                                   double* d_median, uint64_t* d_usable, unsigned presampled = 0, const float* q32 = nullptr,
                                   const uint32_t* qhi = nullptr, const uint32_t* qlo = nullptr);
// qhi / qlo != null (column-resident embed, 4096-row half planes, presampled > 0): the two 32-bit planes of q = |F|^2 the
// fused pass left (PassArgs::qhi / qlo); spec is not read (it no longer holds the spectrum).  The selection then runs on
// the bit patterns of q and the median is sqrt() of the selected element (within 1 ulp of hypot() of the same element).

// ---- column-resident embed (pencil_col_embed_w): bin masks, bin-list check, capacity pass-through ------------------
size_t embed_mask_bytes(int ld, int nplanes);  // bytes of a [nplanes][ld/2][512] uint16 mask array
size_t embed_pres_bytes(int ld);               // pres masks [3][ld/2][512] + per-pair "holds a bin" flags [3][ld/2] behind them
cudaError_t launch_embed_pres(const Launcher& L, const uint32_t* bins, size_t nbits, SpecLayout lay, uint16_t* pres /*[3][ld/2][512]*/);
cudaError_t launch_embed_val(const Launcher& L, const uint32_t* bins, const uint8_t* bits, size_t nbits, int nimg, SpecLayout lay,
                             uint16_t* val /*[nimg*3][ld/2][512]*/);
// d_out2[0] = 1 when some bin is outside 0 < x < PW/2 (or invalid), d_out2[1] = largest row of the list
cudaError_t launch_embed_bins_check(const Launcher& L, const uint32_t* bins, size_t nbits, SpecLayout lay, unsigned* d_out2);
// stego[img] = cover[img] for every image with usable[img] < nbits (S:1009-1012)
// frame bits packed MSB first (S:447-459), rows pstride bytes apart -> [nimg][nbits] one bit per byte
cudaError_t launch_unpack_bits(const Launcher& L, const uint8_t* packed, size_t pstride, uint8_t* bits, size_t nbits, int nimg);
cudaError_t launch_passthrough(const Launcher& L, const uint8_t* cover, uint8_t* stego, size_t img_bytes, int nimg, const uint64_t* usable, size_t nbits);
bool fused_embed_supported(const Launcher& L);  // PassArgs::fused_embed is available (TMA column kernels in use)

// ---- embed scatter (write_bit_on_bin S:712-732) -----------------------------------------
cudaError_t launch_embed(const Launcher& L, double2* spec, int nimg, SpecLayout lay,
                         const uint32_t* bins, const uint8_t* bits, size_t nbits, const double* jitter,
                         double alpha, double cos_a, double sin_a, const uint64_t* usable,
                         const double* adaptive_median = nullptr /* [nimg*3]: alpha scaled by |F| / median (S:704-710) */);

// ---- extract gather + vote (read_bit_from_bin S:734-746, rep3/7 S:468/S:501, pack S:447) --
cudaError_t launch_extract(const Launcher& L, const double2* spec, int nimg, SpecLayout lay,
                           const uint32_t* bins, size_t nbins, int rep, const double* jitter, double alpha,
                           uint8_t* out_bytes, uint8_t* raw_bits, size_t raw_stride = 0,
                           const double* adaptive_median = nullptr /* [nimg*3], S:737-738 */);

// the same votes from a sign map (PassArgs::signmap) of planes with `cols` stored columns; bins must not need the mirror
cudaError_t launch_extract_signmap(const Launcher& L, const uint32_t* signmap, int cols, int nimg, SpecLayout lay,
                                   const uint32_t* bins, size_t nbins, int rep, uint8_t* out_bytes, uint8_t* raw_bits, size_t raw_stride = 0,
                                   int fold = 0 /* the map is that of a folded 8192-row plane (PassArgs::fold) */);

// window of the workspace a bin list touches: d_out2[0] = 1 + largest stored row, [1] = 1 + largest stored column,
// [2] = some bin is read through its Hermitian mirror (three unsigned)
cudaError_t launch_bins_window(const Launcher& L, const uint32_t* bins, size_t nbins, SpecLayout lay, unsigned* d_out2);

// unfused image <-> plane conversion for sizes the fused row passes do not cover (PW or PH > 4096)
cudaError_t launch_u8_to_planes(const Launcher& L, const uint8_t* img, double2* spec, int nimg, int W, int H, int PW, int PH, int center);
cudaError_t launch_planes_to_u8(const Launcher& L, const double2* spec, uint8_t* img, int nimg, int W, int H, int PW, int PH, int center);

// ---- config 5: slab-decomposed 2-D FFT over G GPUs (tfft_slab.cu) ----------------------------------------------------
constexpr int SLAB_MAX_RANKS = 8;
struct SlabDst { double2* p[SLAB_MAX_RANKS]; };  // per destination rank: where its columns of my half rows go (local or peer-mapped)
cudaError_t launch_slab_pack(const Launcher& L, const uint8_t* rows, int nrows, int W, int PW, int npairs, int y0, int center, double2* z);
cudaError_t launch_slab_split(const Launcher& L, const double2* Z, int PW, int ld, int cols, int npairs, int y0, const SlabDst& dst,
                              size_t plane_stride, int row_base);
cudaError_t launch_slab_merge(const Launcher& L, const double2* tiles, int PW, int cols, int G, int R, int npairs, double2* Z);
cudaError_t launch_slab_to_u8(const Launcher& L, const double2* z, int nrows, int W, int PW, int npairs, int y0, int center, uint8_t* rows);
cudaError_t launch_slab_embed(const Launcher& L, double2* slab, int PH, int PW, int cols, int col0, const uint32_t* bins, const uint8_t* bits,
                              size_t nbits, double cos_a, double sin_a);
cudaError_t launch_slab_read(const Launcher& L, const double2* slab, int PH, int PW, int cols, int col0, const uint32_t* bins, size_t nbins,
                             double alpha, int8_t* raw);

// full[y][x] from the half layout (parity hook tfft_forward_spectrum)
cudaError_t launch_expand_half(const Launcher& L, const double2* half_spec, double2* full_spec, int nplanes, SpecLayout lay);

}  // namespace tfft
