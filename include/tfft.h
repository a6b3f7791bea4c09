/* tfft.h -- C ABI of the B200-native TurtleFFT spectral hot path.
 *
 * The reference (rickenator/steganosaurus) has no plugin/FFI interface: its hot path is a set
 * of `static` functions in steganosaurus/src/steganosaur.cpp (cited below as S:line) called from
 * do_embed (S:907-1109) and do_extract (S:1112-1312).  This header is the seam a maintainer
 * would bind instead of those calls: the host keeps PNG I/O, KDF, AEAD,
 * framing and the turtlewalk, emits a bin-index array and a bit array, and hands them over.
 *
 * Conventions
 *   images   uint8 [n][H][W][3] interleaved RGB, exactly what stbi_load(...,3) returns (S:909)
 *   bins     uint32, one per embedded/read bit: plane<<30 | (y*PW + x) with PW = next_pow2(W),
 *            PH = next_pow2(H) (S:393-394); ONE bin list is shared by the whole batch (the
 *            walk is cover-independent, S:797-799)
 *   bits     uint8 0/1, one per byte (S:455-459), [n][nbits] -- every image has its own bits
 *   spectra  complex<double> as (re,im) pairs, [3][PH][PW] row-major, reference sign
 *            convention: forward = sum x[n] e^{+2*pi*i*nk/N} (S:347)
 *   errors   integer codes, never exit(); tfft_strerror() names them
 *   threads  a tfft_ctx is NOT thread-safe: one ctx per GPU per host thread
 *   memory   the caller owns every buffer passed in; the library owns its device workspace
 *
 * Entry points ending in _dev take DEVICE pointers and a cudaStream_t (as void*), enqueue
 * their work on that stream and return without waiting for it (they may synchronise that stream ONCE at entry to
 * look at the device bin list).  They return TFFT_OK once everything is enqueued: the capacity verdict of an embed
 * is NOT in the return code -- pass d_usable and compare it with nbits after synchronising (images over capacity
 * leave as their covers).  Device bin lists are trusted: a plane above 2 or an index outside the padded plane is
 * the caller's bug, as with any device pointer.  One context works on ONE stream at a time (its workspaces are
 * shared): synchronise, or order the streams with an event, before passing a different one.  The others take HOST
 * pointers (pinned memory from tfft_host_alloc recommended), copy in/out on internal
 * streams and are synchronous at return.
 *
 * Environment switches (read once at tfft_create; the defaults are the fast paths, none changes a result beyond
 * rounding, and every one has a parity test in tests/test_env_variants.py):
 *   TFFT_FFT_IMPL=v0        every FFT pass on the simple shared-memory radix-2 kernel (cross-check of the pencil kernels)
 *   TFFT_SPECTRUM=full      no Hermitian halving: full PH x PW spectra
 *   TFFT_WIDE=0             8192-pixel rows and tall images on the unfused four-step path
 *   TFFT_FUSED_EMBED=0      embed: forward columns -> embed_scatter -> inverse columns instead of the column-resident pass
 *   TFFT_EXTRACT_WINDOW=0   extract transforms every column and keeps every row
 *   TFFT_SIGNMAP=0          extract keeps spectra instead of read bits
 *   TFFT_HOST_CHUNK, TFFT_HOST_SLOTS   host pipeline of the host-buffer entry points: images per chunk, chunks in flight
 */
#ifndef TFFT_H
#define TFFT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFFT_ABI_VERSION 1

enum {
    TFFT_OK = 0,
    TFFT_E_INVALID = 1,     /* bad argument (null pointer, non-positive size, rep not in 1,3,7 ...) */
    TFFT_E_CUDA = 2,        /* a CUDA call failed; tfft_last_cuda_error() has the text */
    TFFT_E_CAPACITY = 3,    /* nbits > usable for at least one image (S:1009-1012) */
    TFFT_E_NOMEM = 4,       /* workspace does not fit the device */
    TFFT_E_UNSUPPORTED = 5, /* padded dimension outside [TFFT_MIN_DIM, TFFT_MAX_DIM] */
    TFFT_E_STATE = 6        /* tfft_read_bits without resident spectra */
};

#define TFFT_MIN_DIM 16
#define TFFT_MAX_DIM 16384

typedef struct tfft_ctx tfft_ctx;

/* ---- lifecycle ------------------------------------------------------------------------- */
/* Create a context on CUDA device `device`; owns streams, twiddle tables, workspaces. */
int tfft_create(int device, tfft_ctx** out);
void tfft_destroy(tfft_ctx* ctx);
int tfft_abi_version(void);
const char* tfft_strerror(int code);
const char* tfft_last_cuda_error(const tfft_ctx* ctx);
/* Upper bound for the spectrum workspace in bytes (default: 40% of device memory). */
int tfft_set_workspace_limit(tfft_ctx* ctx, size_t bytes);
/* Params.adaptive_alpha (S:379, S:704-710; experimental and off by default upstream): when on, every later embed writes
 * and every later extract reads bin (y,x) of plane p with alpha * clamp(|F[y][x]| / median_p, 0.5, 2) instead of alpha,
 * exactly as write_bit_on_bin / read_bit_from_bin do with adaptive_alpha = true (the embed uses the cover's medians,
 * the extract the stego image's own, S:922 / S:1124).  These calls run on the general path: no fused embed, no bin
 * window, no sign map. */
int tfft_set_adaptive_alpha(tfft_ctx* ctx, int on);
/* Pinned host memory helpers (cudaHostAlloc / cudaFreeHost). */
void* tfft_host_alloc(size_t bytes);
void tfft_host_free(void* p);
/* Number of kernels this library has launched through ctx since creation. */
uint64_t tfft_launch_count(const tfft_ctx* ctx);

/* ---- embed: replaces S:912-923, S:997-1012, S:1015, S:1086, S:1100-1103 ------------------
 * For each image: plane split (to_planes_u8 S:383), optional (-1)^(x+y) (apply_center S:392),
 * implicit zero pad to PHxPW (pad_to_fft S:393), forward 2-D FFT (fft2d S:359), median |F| per
 * plane (median_abs S:404), capacity count (S:999-1007), phase write at the bins
 * (write_bit_on_bin S:712: |F| kept, phase = +-alpha (+ jitter[i]), conjugate bin mirrored),
 * inverse 2-D FFT, crop (ifft_crop S:399), centre, round/clamp/interleave (from_planes_u8 S:387).
 *   jitter  NULL or [nbits] radians added to the target phase (KS::jitter, S:690)
 *   usable  NULL or [n]     capacity in bits as the reference counts it
 *   median  NULL or [n][3]  median |F| per plane
 * Images with nbits > usable are passed through unmodified (their spectrum is not touched)
 * and the call returns TFFT_E_CAPACITY after finishing the rest of the batch. */
int tfft_embed_batch(tfft_ctx* ctx, const uint8_t* cover, int n, int W, int H,
                     const uint32_t* bins, const uint8_t* bits, size_t nbits, const double* jitter,
                     double alpha, int center, double magmin, double rmin, double rmax,
                     uint8_t* stego, uint64_t* usable, double* median);
/* The same call with the frame bits packed eight to a byte, MSB first -- the order bits_from_bytes / bytes_from_bits use
 * (S:447-459): bits_packed = [n][ceil(nbits / 8)], bit i of image j = (bits_packed[j][i / 8] >> (7 - i % 8)) & 1.  One eighth
 * of the host->device bytes of the bit array (a 4K batch's frame bits are 3 % of an embed+extract step's uploads). */
int tfft_embed_batch_packed(tfft_ctx* ctx, const uint8_t* cover, int n, int W, int H,
                            const uint32_t* bins, const uint8_t* bits_packed, size_t nbits, const double* jitter,
                            double alpha, int center, double magmin, double rmin, double rmax,
                            uint8_t* stego, uint64_t* usable, double* median);
int tfft_embed_batch_dev(tfft_ctx* ctx, const uint8_t* d_cover, int n, int W, int H,
                         const uint32_t* d_bins, const uint8_t* d_bits, size_t nbits,
                         const double* d_jitter, double alpha, int center, double magmin,
                         double rmin, double rmax, uint8_t* d_stego, uint64_t* d_usable,
                         double* d_median, void* stream);

/* ---- extract: replaces S:1116-1123, S:1209, S:1228, S:1266-1268 --------------------------
 * Forward 2-D FFT of every image, phase read at the bins (read_bit_from_bin S:734, ties -> 1),
 * majority vote over `rep` consecutive bins (rep3/rep7_decode_bits S:468/S:501; rep=1: none)
 * and MSB-first packing (bytes_from_bits S:447).
 *   out_bytes  [n][ceil(floor(nbins/rep)/8)] decoded bytes (may be NULL)
 *   raw_bits   [n][nbins] pre-vote bits (may be NULL)
 * The forward column pass of an extract only covers the part of the spectrum the bin list reads (the walk stays
 * inside r <= rmax*min(PH,PW), S:771-774).  The _dev variants reduce the bin list on the device for that and
 * synchronise `stream` ONCE at entry to read the two numbers back; everything after it is enqueued only. */
int tfft_extract_bits(tfft_ctx* ctx, const uint8_t* stego, int n, int W, int H,
                      const uint32_t* bins, size_t nbins, int rep, const double* jitter,
                      double alpha, int center, uint8_t* out_bytes, uint8_t* raw_bits);
int tfft_extract_bits_dev(tfft_ctx* ctx, const uint8_t* d_stego, int n, int W, int H,
                          const uint32_t* d_bins, size_t nbins, int rep, const double* d_jitter,
                          double alpha, int center, uint8_t* d_out_bytes, uint8_t* d_raw_bits,
                          void* stream);

/* Whole-frame extract when the payload length is known (or bounded): ONE forward FFT per image,
 * then the first nhdr_bins bins are voted rep-3 into out_hdr [n][ceil(nhdr_bins/3/8)] (the 38-byte
 * header is nhdr_bins = 912, S:1223-1230) and the remaining bins rep-7 into out_payload
 * [n][ceil((nbins-nhdr_bins)/7/8)] (S:1260-1268).  raw_bits [n][nbins] optional. */
int tfft_extract_frame(tfft_ctx* ctx, const uint8_t* stego, int n, int W, int H,
                       const uint32_t* bins, size_t nbins, size_t nhdr_bins, const double* jitter,
                       double alpha, int center, uint8_t* out_hdr, uint8_t* out_payload,
                       uint8_t* raw_bits);
int tfft_extract_frame_dev(tfft_ctx* ctx, const uint8_t* d_stego, int n, int W, int H,
                           const uint32_t* d_bins, size_t nbins, size_t nhdr_bins,
                           const double* d_jitter, double alpha, int center, uint8_t* d_out_hdr,
                           uint8_t* d_out_payload, uint8_t* d_raw_bits, void* stream);

/* Two-phase extract for the data dependency at S:1253 (payload length is only known after the
 * header has been decoded): tfft_forward_batch keeps the transform of the batch resident in the
 * context; tfft_read_bits may then be called any number of times (header: 912 bins rep 3,
 * payload: 56*(clen+16) bins rep 7).  The batch must fit the workspace (TFFT_E_NOMEM if not).
 * What stays resident is the library's business and never changes a result: the spectra, or -- on 4096-row planes --
 * the output of the row pass, from which the first tfft_read_bits makes either the read bits of the whole quarter plane
 * for its alpha (no jitter, bins within stored rows 0..2047 and left of the Nyquist column; later reads with the same
 * alpha only gather and vote) or, for any other list, the spectra.  Any other entry point called in between drops the
 * resident state (tfft_read_bits then returns TFFT_E_STATE). */
int tfft_forward_batch(tfft_ctx* ctx, const uint8_t* img, int n, int W, int H, int center);
int tfft_read_bits(tfft_ctx* ctx, const uint32_t* bins, size_t nbins, int rep,
                   const double* jitter, double alpha, uint8_t* out_bytes, uint8_t* raw_bits);

/* ---- parity / measurement hooks ----------------------------------------------------------
 * Forward spectra (S:359-366 after S:383-398) of one image: out_c64 = [3][PH][PW] (re,im). */
int tfft_forward_spectrum(tfft_ctx* ctx, const uint8_t* img, int W, int H, int center,
                          double* out_c64);
/* Plain batched complex 2-D FFT, in place, reference sign/scale conventions (fft2d S:359). */
int tfft_fft2d(tfft_ctx* ctx, double* data_c64, int n, int PH, int PW, int inverse);
int tfft_fft2d_dev(tfft_ctx* ctx, double* d_data_c64, int n, int PH, int PW, int inverse, void* stream);
/* One 1-D pass only, on a device array of n planes [PH][PW]: axis 0 = along rows (x),
 * axis 1 = along columns (y).  Used by bench.py to time a single pass for the roofline. */
int tfft_fft_pass_dev(tfft_ctx* ctx, double* d_data_c64, int n, int PH, int PW, int axis,
                      int inverse, void* stream);
/* median |F| and capacity count of device spectra [n][3][PH][PW] (S:404-409, S:999-1007). */
int tfft_median_capacity_dev(tfft_ctx* ctx, const double* d_spec_c64, int n, int PH, int PW,
                             double magmin, double rmin, double rmax, double* d_median,
                             uint64_t* d_usable, void* stream);


/* Host-only (no GPU, no context): the part of the stored spectrum a bin list reads, as the extract entry points
 * compute it before sizing their forward column pass.  half = 1: half-spectrum workspace (columns 0..PW/2; a bin
 * right of the Nyquist column is read through its Hermitian mirror ((PH-y)%PH, PW-x), S:370-372), half = 0: full.
 * rows / cols = 1 + largest stored row / column, mirrored = some bin needs the mirror.  TFFT_E_INVALID for a
 * plane index above 2 or a bin outside the PH x PW plane. */
int tfft_bin_window(const uint32_t* bins, size_t nbins, int W, int H, int half, int* rows, int* cols, int* mirrored);

/* ---- config 5: ONE image too large to be worth replicating, its 2-D FFT slab-decomposed over G GPUs ------------------
 * (fft2d S:359-366 distributed).  One process / context per GPU.  Rank g owns image rows
 * [g R, (g+1) R), R = PH / G; after the exchange it owns the COLUMN slab [3][PH][cols] of the half spectrum (real
 * planes are Hermitian: columns 0..PW/2, padded to ld = PW/2 + 16; cols = ld / G, columns [g cols, (g+1) cols)).
 * The library does the passes and the data movement; the caller provides the transport: either peer-mapped
 * destination pointers (tfft_ipc_*; the split kernel then stores straight into the other GPUs' slabs over NVLink) or
 * a local send buffer that it hands to a collective.  All pointers are device pointers, all calls enqueue on `stream`.
 * G in {1,2,4,8}; PW, PH in [512, 16384]; PH / G >= 2; no capacity gate on this path (a frame that fits an image this
 * size is orders of magnitude below its capacity; S:1009 is checked by the caller if at all). */
int tfft_slab_sizes(int W, int H, int G, int* PW, int* PH, int* R, int* ld, int* cols);
/* forward, step 1: my image rows [nrows][W][3] (image rows y0 = g R ...; nrows = rows of the image inside my slab,
 * possibly 0) -> pair rows -> row FFT -> Hermitian split; element (plane, image row y, column k) is stored at
 * dst[k / cols] + plane * plane_stride + (y - row_base) * cols + k % cols  (units: complex doubles).
 *   peer slabs:   dst[d] = rank d's column slab, plane_stride = PH * cols, row_base = 0
 *   send buffer:  dst[d] = send + d * R * cols,  plane_stride = G * R * cols, row_base = g * R   (per plane [G][R][cols]) */
int tfft_slab_rows_forward_dev(tfft_ctx* ctx, const uint8_t* d_rows, int nrows, int W, int H, int G, int g, int center,
                               double* const* dst, size_t plane_stride, int row_base, void* stream);
/* column pass over PH points on my column slab [3][PH][cols], in place (inverse: e^{-i}, scaled by 1/PH) */
int tfft_slab_cols_dev(tfft_ctx* ctx, double* d_colslab, int W, int H, int G, int inverse, void* stream);
/* phase write (S:712-732) / phase read (S:734-746; d_raw[i] = 0/1, or -1 when bin i lives on another rank) on my slab */
int tfft_slab_embed_dev(tfft_ctx* ctx, double* d_colslab, int W, int H, int G, int g, const uint32_t* d_bins,
                        const uint8_t* d_bits, size_t nbits, double alpha, void* stream);
int tfft_slab_read_dev(tfft_ctx* ctx, const double* d_colslab, int W, int H, int G, int g, const uint32_t* d_bins,
                       size_t nbins, double alpha, int8_t* d_raw, void* stream);
/* inverse, last step: tiles [3][G][R][cols] (tile s = my rows of rank s's column slab after its inverse column pass)
 * -> pair rows -> inverse row FFT -> crop / centre / round / clamp -> my rows of the stego image [nrows][W][3] */
int tfft_slab_rows_inverse_dev(tfft_ctx* ctx, const double* d_tiles, int nrows, int W, int H, int G, int g, int center,
                               uint8_t* d_rows_out, void* stream);
/* CUDA IPC for the peer transport: export a cudaMalloc'd buffer / map another process's buffer (peer access is enabled
 * on demand) / unmap it.  handle = 64 bytes (cudaIpcMemHandle_t). */
int tfft_ipc_export(const void* d_ptr, unsigned char handle[64]);
int tfft_ipc_open(int device, const unsigned char handle[64], void** d_ptr);
int tfft_ipc_close(void* d_ptr);

/* ---- per-kernel timing (CUDA events on the launching stream) -------------------------------
 * When enabled, every kernel group the library launches is bracketed by a pair of events; after
 * the caller has synchronised, tfft_profile_read() returns launches and summed device time per
 * kind.  bench.py uses this for the roofline of the dominant kernel inside the timed region. */
enum {
    TFFT_K_ROW_FWD = 0,  /* row FFT, u8 -> c64 (fused plane split / centre / zero pad) */
    TFFT_K_COL_FWD = 1,  /* column FFT c64 -> c64 */
    TFFT_K_MEDIAN = 2,   /* median |F| + capacity count */
    TFFT_K_EMBED = 3,    /* phase scatter (column-resident embed: the per-image bit masks of the fused pass) */
    TFFT_K_COL_INV = 4,  /* column IFFT */
    TFFT_K_ROW_INV = 5,  /* row IFFT + scale/round/clamp/interleave/crop epilogue */
    TFFT_K_EXTRACT = 6,  /* phase gather + vote + pack */
    TFFT_K_C2C = 7,      /* plain c64 pass from the fft2d / fft_pass hooks */
    TFFT_K_COL_FWD_WIN = 8, /* column FFT of an extract: only the columns / rows that hold bins */
    TFFT_K_COL_EMBED = 9,   /* column-resident embed: forward columns + phase write + inverse columns in one pass */
    TFFT_K_SLAB = 10,       /* config 5 glue: pair pack, split + scatter (the exchange), tile merge, u8 quantise */
    TFFT_K_COUNT = 11
};
int tfft_profile_enable(tfft_ctx* ctx, int on);
int tfft_profile_reset(tfft_ctx* ctx);
/* total_bytes = algorithmic bytes the implemented groups must move (DESIGN.md lists the formulas) */
int tfft_profile_read(tfft_ctx* ctx, int kind, uint64_t* groups, double* total_ms, double* total_bytes);
const char* tfft_kind_name(int kind);

#ifdef __cplusplus
}
#endif
#endif /* TFFT_H */
