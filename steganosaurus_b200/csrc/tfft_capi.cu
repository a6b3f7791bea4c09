// tfft_capi.cu -- the C ABI declared in include/tfft.h: context, workspaces, chunked
// double-buffered host<->device pipeline, and the kernel sequences for embed / extract.
// Reference citations S:n = steganosaurus/src/steganosaur.cpp line n.
#include "../../include/tfft.h"
#include "tfft_kernels.cuh"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

using namespace tfft;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct Slot {
    cudaStream_t stream = nullptr;
    DevBuf spec, spec2, in, out, bits, med, medians, usable, outbytes, raw;  // spec2: scratch of the four-step passes (dims > 4096)
    DevBuf signmap;  // extract without jitter on 4096-row planes: read bits of every element instead of the spectrum
    DevBuf q32;      // embed on 4096-row half planes: float copy of |F|^2 left by the column pass for the median scan
                     // (column-resident embed: the two 32-bit planes of the exact q instead, qhi then qlo)
    DevBuf val;      // column-resident embed: per-image bit masks of the fused pass (embed_val_build)
    DevBuf bitsp;    // tfft_embed_batch_packed: the chunk's packed frame bits as uploaded
    // pinned staging for the small per-chunk results (capacity verdict, medians, decoded bytes): they
    // are copied to the caller's (possibly pageable) memory only when the chunk is drained, so the
    // asynchronous pipeline never blocks on a pageable cudaMemcpyAsync
    unsigned char* h_stage = nullptr;
    size_t h_stage_cap = 0;
};

}  // namespace

constexpr int NSLOT = 8;   // slots allocated per context
static int HOST_SLOTS = 4;  // host pipeline depth in use (TFFT_HOST_SLOTS): H2D of later chunks, kernels, D2H of earlier chunks overlap

struct tfft_ctx {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    size_t total_mem = 0;
    size_t ws_limit = 0;
    double2* d_tw = nullptr;
    Slot slot[NSLOT];
    DevBuf bins, jitter;
    uint64_t launches = 0;
    int fft_impl = 1;
    bool use_half = true;   // real-input symmetry (half-spectrum workspace); TFFT_SPECTRUM=full disables
    bool col_sample = true; // 4096-row half planes: the median sample is dropped by the forward column pass
    bool use_wide = true;   // 8192-pixel rows on the half-spectrum path; TFFT_WIDE=0 keeps them on the unfused four-step path
    bool use_window = true; // extract: the forward column pass keeps only the rows / columns that hold bins (TFFT_EXTRACT_WINDOW=0: all)
    bool use_q32 = true;     // embed, 4096-row half planes (unfused column stage): the median scan reads a float copy of |F|^2
    bool use_signmap = true; // extract without jitter, 4096-row planes: the column pass leaves read bits, not spectra (TFFT_SIGNMAP=0)
    bool use_fused = true;   // embed, 4096-row half planes, no jitter: forward columns + phase write + inverse columns in ONE pass
                             // (TFFT_FUSED_EMBED=0: the three-kernel sequence col_fwd -> embed_scatter -> col_inv)
    DevBuf pres;             // ... its per-call bin-presence masks [3][ld/2][512] (shared by the batch)
    bool adaptive = false;   // tfft_set_adaptive_alpha: alpha scaled by |F| / median per bin (S:704-710; experimental upstream)
    unsigned* d_win = nullptr;  // device-pointer entry points: bin window reduced on the device ...
    unsigned* h_win = nullptr;  // ... and read back through this pinned pair
    DevBuf full;            // expansion target of the tfft_forward_spectrum hook
    DevBuf slab_z, slab_tmp;  // config 5: pair rows [3][R/2][PW] and the scratch batch of their four-step row passes
    SpecLayout res_lay{0, 0, 0, 0};
    // resident spectra for the two-phase extract
    int res_n = 0, res_PH = 0, res_PW = 0;
    double2* res_spec = nullptr;  // buffer of slot 0 that holds them
    // 4096-row half planes: tfft_forward_batch stops after the row pass (res_stage 1) and the first tfft_read_bits decides
    // what the column pass leaves behind -- the sign map of the quarter plane for this alpha (res_stage 2; the row-pass
    // output stays intact, so a later list outside the map or with jitter can still get the full pass) or the spectrum
    // (res_stage 0, as for every other geometry)
    int res_stage = 0, res_W = 0, res_H = 0;
    double res_alpha = 0.0;
    char cuda_err[256] = {0};
    // per-kernel-kind timing (tfft_profile_*)
    bool prof_on = false;
    struct ProfEvt { cudaEvent_t a, b; int kind; double bytes; };
    std::vector<ProfEvt> prof_pending;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[TFFT_K_COUNT] = {0};
    double prof_bytes[TFFT_K_COUNT] = {0};
    uint64_t prof_groups[TFFT_K_COUNT] = {0};
};

namespace {

// median: sample (<= 2^18 keys) and bracket members per plane.  The bracket holds ~1.2 % of the stored elements, so
// planes beyond 2^25 bins (8192 x 8192 and up) get a larger list instead of falling back to the generic passes.
inline uint32_t cand_cap_for(size_t P) { return P > ((size_t)1 << 25) ? (1u << 21) : (1u << 19); }
constexpr int MAX_CHUNK = 64;
static int HOST_CHUNK = 8;  // host-buffer entry points: small chunks so H2D / kernels / D2H of neighbouring chunks overlap (TFFT_HOST_CHUNK)

int fail_cuda(tfft_ctx* c, cudaError_t e, const char* where) {
    snprintf(c->cuda_err, sizeof(c->cuda_err), "%.160s: %.80s", where, cudaGetErrorString(e));
    return TFFT_E_CUDA;
}
#define CK(call)                                                   \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return fail_cuda(ctx, e_, #call);   \
    } while (0)

int ensure(tfft_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return TFFT_OK;
    if (b.p) { CK(cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return TFFT_E_NOMEM; }
    if (e != cudaSuccess) return fail_cuda(ctx, e, "cudaMalloc");
    b.cap = bytes;
    return TFFT_OK;
}
void release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

inline int next_pow2_i(int v) { int p = 1; while (p < v) p <<= 1; return p; }
inline int ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

struct Geom {
    int W, H, PW, PH, lw, lh;
    size_t P;          // PH*PW
    size_t img_bytes;  // H*W*3
    int half, ld;      // workspace layout (SpecLayout)
    int large;         // full-spectrum path for a padded dimension above 4096: unfused conversion + four-step passes
    int col4;          // half-spectrum workspace whose column passes are four-step (PH = 8192 / 16384): needs the scratch batch
    size_t E;          // stored elements per plane = PH*ld
    SpecLayout lay() const { return SpecLayout{PH, PW, ld, half, 0, nullptr}; }
    // layout of a tall half-spectrum workspace whose column pass stopped after the four-step sub-transforms
    SpecLayout lay_fs(const double2* tw) const { return SpecLayout{PH, PW, ld, half, PH / 4096, tw}; }
};
int make_geom(const tfft_ctx* ctx, int W, int H, Geom& g) {
    if (W <= 0 || H <= 0) return TFFT_E_INVALID;
    g.W = W; g.H = H;
    g.PW = next_pow2_i(W); g.PH = next_pow2_i(H);  // S:394
    // the FFT passes need at least 2 points per axis; tiny images are padded further only by
    // the reference's own rule, so anything below TFFT_MIN_DIM is refused rather than changed.
    if (g.PW > TFFT_MAX_DIM || g.PH > TFFT_MAX_DIM) return TFFT_E_UNSUPPORTED;
    if (g.PW < TFFT_MIN_DIM || g.PH < TFFT_MIN_DIM) return TFFT_E_UNSUPPORTED;
    g.lw = ilog2(g.PW); g.lh = ilog2(g.PH);
    g.P = (size_t)g.PW * g.PH;
    g.img_bytes = (size_t)W * H * 3;
    // half-spectrum workspace whenever both axes run on the pencil kernels (512..4096 points), and for 8192-pixel
    // rows (packed into the 4096-point fused row kernels) with any supported height
    const bool ok = ctx && ctx->use_half && ctx->fft_impl != 0;
    g.half = (ok && g.lw >= 9 && g.lw <= 12 && g.lh >= 9 && g.lh <= 12) ? 1 : 0;
    if (ok && ctx->use_wide && g.lw == 13 && g.lh >= 9 && g.lh <= 14) g.half = 1;
    if (ok && ctx->use_wide && g.lw >= 9 && g.lw <= 12 && g.lh >= 13 && g.lh <= 14) g.half = 1;  // tall: four-step columns
    g.col4 = (g.half && g.lh > 12) ? 1 : 0;
    g.large = (!g.half && ctx && ctx->fft_impl != 0 && (g.lw > 12 || g.lh > 12)) ? 1 : 0;
    g.ld = g.half ? g.PW / 2 + 16 : g.PW;
    g.E = (size_t)g.PH * g.ld;
    return TFFT_OK;
}

Launcher make_launcher(tfft_ctx* ctx, cudaStream_t s) {
    Launcher L;
    L.stream = s;
    L.launch_counter = &ctx->launches;
    L.sm_count = ctx->sm_count;
    L.smem_optin = ctx->smem_optin;
    L.fft_impl = ctx->fft_impl;
    return L;
}

// Brackets one kernel group with events on its stream when profiling is enabled.
struct ProfScope {
    tfft_ctx* c; cudaStream_t s; tfft_ctx::ProfEvt e; bool on;
    static cudaEvent_t get(tfft_ctx* c) {
        if (!c->prof_pool.empty()) { cudaEvent_t ev = c->prof_pool.back(); c->prof_pool.pop_back(); return ev; }
        cudaEvent_t ev = nullptr;
        cudaEventCreate(&ev);
        return ev;
    }
    ProfScope(tfft_ctx* c_, cudaStream_t s_, int kind, double bytes) : c(c_), s(s_), on(c_->prof_on) {
        if (!on) return;
        e.kind = kind; e.bytes = bytes; e.a = get(c); e.b = get(c);
        cudaEventRecord(e.a, s);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(e.b, s);
        c->prof_pending.push_back(e);
    }
};
void prof_drain(tfft_ctx* c) {
    for (auto& e : c->prof_pending) {
        float ms = 0.f;
        cudaEventSynchronize(e.b);
        if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) { c->prof_ms[e.kind] += ms; c->prof_groups[e.kind]++; c->prof_bytes[e.kind] += e.bytes; }
        c->prof_pool.push_back(e.a); c->prof_pool.push_back(e.b);
    }
    c->prof_pending.clear();
}

// images per chunk so that `nslots` spectrum workspaces fit the limit
int chunk_for(const tfft_ctx* ctx, const Geom& g, int n, int nslots) {
    const size_t per_img = 3 * g.E * sizeof(double2) * ((g.large || g.col4) ? 2 : 1);
    size_t c = ctx->ws_limit / nslots / per_img;
    if (c < 1) c = 1;
    if (c > (size_t)MAX_CHUNK) c = MAX_CHUNK;
    if (c > (size_t)n) c = n;
    return (int)c;
}

int ensure_slot(tfft_ctx* ctx, Slot& S, const Geom& g, int chunk, bool need_io, size_t nbits, size_t outbytes, size_t rawbytes) {
    int rc;
    const int nplanes = chunk * 3;
    if ((rc = ensure(ctx, S.spec, (size_t)nplanes * g.E * sizeof(double2)))) return rc;
    if ((g.large || g.col4) && (rc = ensure(ctx, S.spec2, (size_t)nplanes * g.E * sizeof(double2)))) return rc;
    if ((rc = ensure(ctx, S.med, median_work_bytes(nplanes, cand_cap_for(g.P))))) return rc;
    if ((rc = ensure(ctx, S.medians, sizeof(double) * nplanes))) return rc;
    if ((rc = ensure(ctx, S.usable, sizeof(uint64_t) * chunk))) return rc;
    if (need_io) {
        if ((rc = ensure(ctx, S.in, (size_t)chunk * g.img_bytes))) return rc;
        if ((rc = ensure(ctx, S.out, (size_t)chunk * g.img_bytes))) return rc;
        if (nbits && (rc = ensure(ctx, S.bits, (size_t)chunk * nbits))) return rc;
    }
    if (outbytes && (rc = ensure(ctx, S.outbytes, (size_t)chunk * outbytes))) return rc;
    if (rawbytes && (rc = ensure(ctx, S.raw, (size_t)chunk * rawbytes))) return rc;
    return TFFT_OK;
}

int ensure_stage(tfft_ctx* ctx, Slot& S, size_t bytes) {
    if (bytes <= S.h_stage_cap) return TFFT_OK;
    if (S.h_stage) cudaFreeHost(S.h_stage);
    S.h_stage = nullptr; S.h_stage_cap = 0;
    CK(cudaHostAlloc((void**)&S.h_stage, bytes, cudaHostAllocDefault));
    S.h_stage_cap = bytes;
    return TFFT_OK;
}

// ---- kernel sequences -----------------------------------------------------------------------
PassArgs base_args(tfft_ctx* ctx, double2* spec, int nimg, const Geom& g, int center) {
    PassArgs a;
    memset(&a, 0, sizeof(a));
    a.spec = spec;
    a.tw = ctx->d_tw;
    a.nplanes = nimg * 3;
    a.W = g.W; a.H = g.H; a.PW = g.PW; a.PH = g.PH;
    a.center = center;
    a.in_rows = g.PH;
    a.out_rows = g.PH;
    a.half = g.half;
    a.ld = g.ld;
    return a;
}

// The two generic passes of the large-size path (a padded dimension of 8192 / 16384).  A four-step pass leaves its
// result in the scratch batch; instead of copying it back the next pass simply runs on the other buffer, and one copy
// happens at the end only when an odd number of four-step passes ran.
int c2c_two_passes(tfft_ctx* ctx, const Launcher& L, PassArgs a, double2* spec, double2* tmp, int nimg, const Geom& g, bool rows_first) {
    double2 *cur = spec, *other = tmp;
    for (int k = 0; k < 2; k++) {
        const int axis = (k == 0) == rows_first ? 0 : 1;
        a.axis = axis; a.log2n = axis == 0 ? g.lw : g.lh;
        a.spec = cur; a.tmp = other;
        const bool four = a.log2n > 12;
        a.leave_in_tmp = four ? 1 : 0;
        { ProfScope ps(ctx, L.stream, TFFT_K_C2C, (double)nimg * 3.0 * 32.0 * (double)g.P); CK(launch_fft_pass(L, a)); }
        if (four) std::swap(cur, other);
    }
    if (cur != spec) CK(cudaMemcpyAsync(spec, cur, (size_t)nimg * 3 * g.P * sizeof(double2), cudaMemcpyDeviceToDevice, L.stream));
    return TFFT_OK;
}

// Part of the workspace an extract reads: stored rows < rows, stored columns < cols (0: unknown -> everything).  The
// reference's walk only visits the quarter annulus r <= rmax * min(PH, PW) next to index (0,0) (S:771-774), so the
// forward column pass of an extract neither transforms the columns nor stores the rows beyond it.
struct BinWindow { int rows = 0, cols = 0, mirrored = 0; };  // mirrored: some bin sits right of the Nyquist column of a half plane

// optional extras of forward_images
// How an embed call runs its column stage: fused (column-resident, pencil_col_embed_w) when the geometry and the bin list
// allow it, else forward columns -> embed_scatter -> inverse columns.
struct EmbedPlan { bool fused = false; int k3max = 0; };

struct FwdOpts {
    unsigned long long* sample_q = nullptr;  // embed, 4096-row half planes: the column pass drops the median sample here
    unsigned sample_stride = 0;
    float* q32 = nullptr;                    // ... and a float copy of |F|^2 of every element
    const BinWindow* win = nullptr;          // extract: part of the workspace the bin list reads
    bool fs_sub_only = false;                // extract, tall half planes: stop the four-step column pass after its sub-transforms
    uint32_t* signmap = nullptr;             // extract: leave read bits (for this alpha) instead of the column-pass spectrum
    double alpha = 0.0;
    bool rows_only = false, cols_only = false;  // two-phase extract: the row pass now, the column pass when the first bin list arrives
    bool fold = false;                       // extract, 8192-row half planes, with signmap: the row pass takes the first radix-2
                                             // step of the column transform (PassArgs::fold), the columns run as 4096-point passes
};

// forward 2-D FFT of `nimg` u8 images into spec (S:912-921 / S:1116-1123)
// `where` (optional) receives the buffer that holds the spectrum afterwards: spec, or tmp when the four-step column
// pass of a tall half-spectrum workspace left its result in the scratch batch (no copy back)
// `sample_q` (optional, embed only): the 4096-point column pass drops its median sample there (col_pass_samples() > 0).
int forward_images(tfft_ctx* ctx, const Launcher& L, double2* spec, double2* tmp, const uint8_t* d_img, int nimg, const Geom& g, int center,
                   double2** where = nullptr, const FwdOpts& o = FwdOpts{}) {
    const BinWindow* win = o.win;
    if (where) *where = spec;
    PassArgs a = base_args(ctx, spec, nimg, g, center);
    if (g.large) {  // unfused: u8 -> planes (zero pad materialised), then two generic c2c passes
        { ProfScope ps(ctx, L.stream, TFFT_K_ROW_FWD, (double)nimg * 3.0 * ((double)g.W * g.H + 16.0 * (double)g.P)); CK(launch_u8_to_planes(L, d_img, spec, nimg, g.W, g.H, g.PW, g.PH, center)); }
        a.inverse = 0;
        return c2c_two_passes(ctx, L, a, spec, tmp, nimg, g, /*rows_first=*/true);
    }
    const double cols = (double)g.ld;  // columns the workspace keeps (PW, or PW/2+16 in half mode)
    a.img_in = d_img;
    a.axis = 0; a.log2n = g.lw; a.inverse = 0;
    a.in_rows = g.H;  // rows >= H are zero padding (S:395)
    if (o.fold) {
        // 8192-row half planes of an extract that only needs read bits (the caller checked: sign map conditions, window of at
        // most 4096 rows): the row pass folds rows y and y + 4096 (first radix-2 step of the column transform), all 8192
        // stored rows are written, and the column pass is the 4096-row sign-map pass over 2 x 3 planes per image
        a.fold = 1;
        { ProfScope ps(ctx, L.stream, TFFT_K_ROW_FWD, (double)nimg * 3.0 * ((double)g.W * g.H + 16.0 * (double)g.PH * cols)); CK(launch_fft_pass(L, a)); }
        a.fold = 0; a.img_in = nullptr;
        a.axis = 1; a.log2n = 12; a.PW = g.ld; a.half = 0;
        a.nplanes = nimg * 6; a.PH = 4096; a.in_rows = 4096;
        a.out_rows = std::min(2048, (win->rows + 1) / 2);
        a.col_limit = std::min(g.ld, win->cols);
        a.signmap = o.signmap; a.sign_alpha = o.alpha;
        const double ncols = (double)a.col_limit;
        ProfScope ps(ctx, L.stream, TFFT_K_COL_FWD_WIN, (double)nimg * 3.0 * (16.0 * (double)g.PH * ncols + 2.0 * (double)a.out_rows * ncols / 8.0));
        CK(launch_fft_pass(L, a));
        return TFFT_OK;
    }
    if (!o.cols_only) { ProfScope ps(ctx, L.stream, TFFT_K_ROW_FWD, (double)nimg * 3.0 * ((double)g.W * g.H + 16.0 * (double)g.H * cols)); CK(launch_fft_pass(L, a)); }
    if (o.rows_only) return TFFT_OK;
    a.img_in = nullptr;
    a.axis = 1; a.log2n = g.lh;  // in_rows stays H: the row pass left rows >= H unwritten (they are zero)
    if (g.half) { a.PW = g.ld; a.half = 0; }  // the column pass just sees a plane of ld columns
    if (g.col4) {  // four-step columns read every row: materialise the zero rows the row pass skipped
        if (g.H < g.PH)
            CK(cudaMemset2DAsync(spec + (size_t)g.H * g.ld, (size_t)g.PH * g.ld * sizeof(double2), 0,
                                 (size_t)(g.PH - g.H) * g.ld * sizeof(double2), (size_t)nimg * 3, L.stream));
        if (o.fs_sub_only) {  // readers combine at their bins (SpecLayout::fs_r): one in-place pass over the columns they touch
            a.in_rows = g.PH; a.fourstep_sub_only = 1;
            double ncols = cols;
            if (win && ctx->use_window && win->cols > 0 && win->cols < g.ld) { a.col_limit = win->cols; ncols = (double)win->cols; }
            ProfScope ps(ctx, L.stream, TFFT_K_COL_FWD_WIN, (double)nimg * 3.0 * 32.0 * (double)g.PH * ncols);
            CK(launch_fft_pass(L, a));
            return TFFT_OK;
        }
        a.in_rows = g.PH; a.tmp = tmp;
        a.leave_in_tmp = where ? 1 : 0;
        ProfScope ps(ctx, L.stream, TFFT_K_C2C, (double)nimg * 3.0 * 32.0 * (double)g.PH * cols);
        CK(launch_fft_pass(L, a));
        if (where) *where = tmp;
        return TFFT_OK;
    }
    a.sample_q = o.sample_q; a.sample_stride = o.sample_stride; a.q32 = o.q32;
    double out_rows = (double)g.PH, ncols = cols;
    int kind = TFFT_K_COL_FWD;
    if (win && ctx->use_window && win->rows > 0 && win->cols > 0 && (win->rows < g.PH || win->cols < g.ld)) {
        a.out_rows = std::min(g.PH, win->rows);
        a.col_limit = std::min(g.ld, win->cols);
        out_rows = (double)a.out_rows; ncols = (double)a.col_limit;
        kind = TFFT_K_COL_FWD_WIN;
    }
    double out_bytes = 16.0 * out_rows * ncols;
    if (o.signmap && kind == TFFT_K_COL_FWD_WIN) {  // (the caller checked the conditions: 4096 rows, window of at most 2048 rows)
        a.signmap = o.signmap; a.sign_alpha = o.alpha;
        out_bytes = out_rows * ncols / 8.0;
    }
    if (o.q32) out_bytes += 4.0 * out_rows * ncols;
    { ProfScope ps(ctx, L.stream, kind, (double)nimg * 3.0 * (16.0 * (double)g.H * ncols + out_bytes)); CK(launch_fft_pass(L, a)); }
    return TFFT_OK;
}

// inverse 2-D FFT + crop + quantise (S:1100-1103): columns first so that the last pass runs
// along image rows and can emit interleaved u8 directly.
int inverse_images(tfft_ctx* ctx, const Launcher& L, double2* spec, double2* tmp, uint8_t* d_img, int nimg, const Geom& g, int center) {
    PassArgs a = base_args(ctx, spec, nimg, g, center);
    if (g.large) {
        a.inverse = 1;
        { int rc2 = c2c_two_passes(ctx, L, a, spec, tmp, nimg, g, /*rows_first=*/false); if (rc2) return rc2; }
        { ProfScope ps(ctx, L.stream, TFFT_K_ROW_INV, (double)nimg * 3.0 * ((double)g.W * g.H + 16.0 * (double)g.P)); CK(launch_planes_to_u8(L, spec, d_img, nimg, g.W, g.H, g.PW, g.PH, center)); }
        return TFFT_OK;
    }
    const double cols = (double)g.ld;
    a.axis = 1; a.log2n = g.lh; a.inverse = 1;
    a.out_rows = g.H;  // rows >= H are cropped away (S:399-403): the column pass does not store them
    if (g.half) { a.PW = g.ld; a.half = 0; }
    if (g.col4) {  // four-step: the result lands in the other buffer, which the row pass then reads
        a.out_rows = g.PH; a.tmp = tmp; a.leave_in_tmp = 1;
        ProfScope ps(ctx, L.stream, TFFT_K_C2C, (double)nimg * 3.0 * 32.0 * (double)g.PH * cols);
        CK(launch_fft_pass(L, a));
        a.spec = tmp; a.leave_in_tmp = 0;
    } else {
        ProfScope ps(ctx, L.stream, TFFT_K_COL_INV, (double)nimg * 3.0 * 16.0 * ((double)g.PH * cols + (double)g.H * cols));
        CK(launch_fft_pass(L, a));
    }
    a.tmp = nullptr;
    a.PW = g.PW; a.half = g.half;
    a.axis = 0; a.log2n = g.lw;
    a.img_out = d_img;
    { ProfScope ps(ctx, L.stream, TFFT_K_ROW_INV, (double)nimg * 3.0 * ((double)g.W * g.H + 16.0 * (double)g.H * cols)); CK(launch_fft_pass(L, a)); }
    return TFFT_OK;
}

// Column-resident embed of one chunk: row pass -> [bit masks] -> ONE column pass (forward, phase write, inverse; leaves
// q = |F|^2 as two word planes + the median sample) -> median / capacity on the q planes -> inverse row pass ->
// pass-through of the images over capacity (the phase write was speculative, S:1009-1012).
int embed_chunk_fused(tfft_ctx* ctx, const Launcher& L, Slot& S, const uint8_t* d_cover, int nimg, const Geom& g,
                      const uint32_t* d_bins, const uint8_t* d_bits, size_t nbits, double alpha, int center, double magmin,
                      double rmin, double rmax, uint8_t* d_stego, uint64_t* d_usable, double* d_median, const EmbedPlan& plan) {
    int rc;
    MedianWork mw;
    median_work_carve(mw, S.med.p, nimg * 3, cand_cap_for(g.P));
    const double cols = (double)g.ld;
    const size_t E = g.E;
    if ((rc = ensure(ctx, S.q32, (size_t)nimg * 3 * E * 2 * sizeof(uint32_t)))) return rc;
    if (nbits && (rc = ensure(ctx, S.val, embed_mask_bytes(g.ld, nimg * 3)))) return rc;
    uint32_t* qhi = (uint32_t*)S.q32.p;
    uint32_t* qlo = qhi + (size_t)nimg * 3 * E;
    PassArgs a = base_args(ctx, (double2*)S.spec.p, nimg, g, center);
    a.img_in = d_cover; a.axis = 0; a.log2n = g.lw; a.inverse = 0; a.in_rows = g.H;
    { ProfScope ps(ctx, L.stream, TFFT_K_ROW_FWD, (double)nimg * 3.0 * ((double)g.W * g.H + 16.0 * (double)g.H * cols)); CK(launch_fft_pass(L, a)); }
    if (nbits) {
        ProfScope ps(ctx, L.stream, TFFT_K_EMBED, (double)nimg * ((double)nbits * 5.0 + 2.0 * (double)embed_mask_bytes(g.ld, 3)));
        CK(launch_embed_val(L, d_bins, d_bits, nbits, nimg, g.lay(), (uint16_t*)S.val.p));
    }
    a.img_in = nullptr; a.axis = 1; a.log2n = g.lh; a.PW = g.ld; a.half = 0;
    a.in_rows = g.H; a.out_rows = g.H;   // rows >= H: zero on input (S:395), cropped on output (S:399-403)
    a.fused_embed = 1;
    a.sample_q = (unsigned long long*)mw.cand; a.sample_stride = mw.cand_cap;
    a.qhi = qhi; a.qlo = qlo;
    a.embed_pres = nbits ? (const uint16_t*)ctx->pres.p : nullptr;
    a.embed_val = nbits ? (const uint16_t*)S.val.p : nullptr;
    a.embed_k3max = plan.k3max;
    a.embed_cos = cos(alpha); a.embed_sin = sin(alpha);
    { ProfScope ps(ctx, L.stream, TFFT_K_COL_EMBED,
                   (double)nimg * 3.0 * (32.0 * (double)g.H * cols + 8.0 * (double)g.PH * cols + (nbits ? 2.0 * (cols / 2.0) * 512.0 * 2.0 : 0.0)));
      CK(launch_fft_pass(L, a)); }
    const int m = std::min(g.PH, g.PW);
    const unsigned presampled = col_pass_samples(g.PH, g.PW, g.half);
    { ProfScope ps(ctx, L.stream, TFFT_K_MEDIAN, (double)nimg * 3.0 * 4.0 * (double)E);
      CK(launch_median_capacity(L, nullptr, nimg * 3, g.lay(), magmin, rmin * m, rmax * m, mw, d_median, d_usable, presampled, nullptr, qhi, qlo)); }
    a.fused_embed = 0; a.sample_q = nullptr; a.qhi = nullptr; a.qlo = nullptr;
    a.PW = g.PW; a.half = g.half; a.axis = 0; a.log2n = g.lw; a.inverse = 1; a.img_out = d_stego;
    { ProfScope ps(ctx, L.stream, TFFT_K_ROW_INV, (double)nimg * 3.0 * ((double)g.W * g.H + 16.0 * (double)g.H * cols)); CK(launch_fft_pass(L, a)); }
    if (nbits) CK(launch_passthrough(L, d_cover, d_stego, g.img_bytes, nimg, d_usable, nbits));
    return TFFT_OK;
}

// can this call's column stage run fused?  (geometry part; the bin list part is decided by the callers)
bool fused_geometry_ok(const tfft_ctx* ctx, const Launcher& L, const Geom& g, const double* jitter) {
    return ctx->use_fused && !ctx->adaptive && ctx->col_sample && !jitter && !g.large && !g.col4 && g.half && g.lh == 12 && ctx->fft_impl == 1 &&
           fused_embed_supported(L) && col_pass_samples(g.PH, g.PW, g.half) <= cand_cap_for(g.P);
}

int embed_chunk(tfft_ctx* ctx, const Launcher& L, Slot& S, const uint8_t* d_cover, int nimg, const Geom& g,
                const uint32_t* d_bins, const uint8_t* d_bits, size_t nbits, const double* d_jitter,
                double alpha, int center, double magmin, double rmin, double rmax,
                uint8_t* d_stego, uint64_t* d_usable, double* d_median, const EmbedPlan& plan = EmbedPlan{}) {
    if (plan.fused)
        return embed_chunk_fused(ctx, L, S, d_cover, nimg, g, d_bins, d_bits, nbits, alpha, center, magmin, rmin, rmax, d_stego, d_usable,
                                 d_median, plan);
    double2* spec = nullptr;  // whichever of the slot's two buffers holds the spectrum
    MedianWork mw;
    median_work_carve(mw, S.med.p, nimg * 3, cand_cap_for(g.P));
    // the 4096-point forward column kernel can drop the median sample while its results are still on chip
    unsigned presampled = (ctx->col_sample && !g.col4 && g.lh == 12 && g.half && ctx->fft_impl == 1) ? col_pass_samples(g.PH, g.PW, g.half) : 0;
    if (presampled > mw.cand_cap) presampled = 0;
    FwdOpts fo;
    if (presampled) { fo.sample_q = (unsigned long long*)mw.cand; fo.sample_stride = mw.cand_cap; }
    // ... and a float copy of q = |F|^2 of every element: the scan then reads 4 bytes per element instead of 16
    const bool q32 = presampled && ctx->use_q32 && signmap_supported(L);
    if (q32) {
        int rc2 = ensure(ctx, S.q32, (size_t)nimg * 3 * g.E * sizeof(float));
        if (rc2) return rc2;
        fo.q32 = (float*)S.q32.p;
    }
    int rc = forward_images(ctx, L, (double2*)S.spec.p, (double2*)S.spec2.p, d_cover, nimg, g, center, &spec, fo);
    if (rc) return rc;
    double2* other = spec == (double2*)S.spec.p ? (double2*)S.spec2.p : (double2*)S.spec.p;
    const int m = std::min(g.PH, g.PW);
    { ProfScope ps(ctx, L.stream, TFFT_K_MEDIAN, (double)nimg * 3.0 * (q32 ? 4.0 : 16.0) * (double)g.E);
      CK(launch_median_capacity(L, spec, nimg * 3, g.lay(), magmin, rmin * m, rmax * m, mw, d_median, d_usable, presampled, q32 ? (const float*)S.q32.p : nullptr)); }
    { ProfScope ps(ctx, L.stream, TFFT_K_EMBED, (double)nimg * (double)nbits * (16.0 + (g.half ? 16.0 : 32.0) + 5.0));
      CK(launch_embed(L, spec, nimg, g.lay(), d_bins, d_bits, nbits, d_jitter, alpha, cos(alpha), sin(alpha), d_usable,
                      ctx->adaptive ? d_median : nullptr)); }
    return inverse_images(ctx, L, spec, other, d_stego, nimg, g, center);
}

// nhdr == 0: one segment of `rep`; nhdr > 0: rep-3 header segment + rep-7 payload segment (S:1223-1268)
int extract_chunk(tfft_ctx* ctx, const Launcher& L, Slot& S, const uint8_t* d_stego, int nimg, const Geom& g,
                  const uint32_t* d_bins, size_t nbins, int rep, size_t nhdr, const double* d_jitter, double alpha, int center,
                  uint8_t* d_out_bytes, uint8_t* d_out_payload, uint8_t* d_raw, const BinWindow& win) {
    double2* spec = nullptr;
    FwdOpts fo;
    fo.win = &win;
    // Without jitter the read decision is a property of the element alone, so a 4096-row column pass can leave one bit
    // per element instead of 16 bytes (window of at most 2048 rows, no bin behind the Nyquist column, alpha away from
    // 0 and pi so that the decision is the sign of the imaginary part off the real axis).
    const bool sign = ctx->use_signmap && !ctx->adaptive && ctx->use_window && !d_jitter && !g.large && !g.col4 && g.lh == 12 && nbins > 0 &&
                      win.rows > 0 && win.rows <= 2048 && !win.mirrored && alpha >= 1e-6 && alpha <= 3.14159 && signmap_supported(L);
    if (sign) {
        int rc2 = ensure(ctx, S.signmap, (size_t)nimg * 3 * sign_map_words(g.ld) * sizeof(uint32_t));
        if (rc2) return rc2;
        fo.signmap = (uint32_t*)S.signmap.p; fo.alpha = alpha;
    }
    // 8192-row half planes: the same sign map from two 4096-row column passes per plane, the row pass folds (PassArgs::fold)
    const bool fold = ctx->use_signmap && !ctx->adaptive && ctx->use_window && !d_jitter && g.col4 && g.half && g.lh == 13 && nbins > 0 &&
                      win.rows > 0 && win.rows <= 4096 && win.cols > 0 && !win.mirrored && alpha >= 1e-6 && alpha <= 3.14159 && signmap_supported(L);
    if (fold) {
        int rc2 = ensure(ctx, S.signmap, (size_t)nimg * 6 * sign_map_words(g.ld) * sizeof(uint32_t));
        if (rc2) return rc2;
        fo.signmap = (uint32_t*)S.signmap.p; fo.alpha = alpha; fo.fold = true;
    }
    if (ctx->adaptive) fo.win = nullptr;  // the medians (S:1124) need the whole spectrum
    const bool fs = g.col4 && !fold && !ctx->adaptive && ctx->fft_impl == 1;  // tall planes: no combine pass, the readers combine at their bins
    fo.fs_sub_only = fs;
    int rc = forward_images(ctx, L, (double2*)S.spec.p, (double2*)S.spec2.p, d_stego, nimg, g, center, &spec, fo);
    if (rc) return rc;
    const SpecLayout lay = fs ? g.lay_fs(ctx->d_tw) : g.lay();
    const double* amed = nullptr;
    if (ctx->adaptive) {
        MedianWork mw;
        median_work_carve(mw, S.med.p, nimg * 3, cand_cap_for(g.P));
        ProfScope ps(ctx, L.stream, TFFT_K_MEDIAN, (double)nimg * 3.0 * 16.0 * (double)g.E);
        CK(launch_median_capacity(L, spec, nimg * 3, g.lay(), 0.0, 0.0, 0.0, mw, (double*)S.medians.p, nullptr));
        amed = (const double*)S.medians.p;
    }
    if (sign || fold) {
        const uint32_t* bm = (const uint32_t*)S.signmap.p;
        const int fd = fold ? 1 : 0;
        ProfScope ps(ctx, L.stream, TFFT_K_EXTRACT, (double)nimg * (double)nbins * (4.0 + 4.0));
        if (nhdr == 0) {
            CK(launch_extract_signmap(L, bm, g.ld, nimg, g.lay(), d_bins, nbins, rep, d_out_bytes, d_raw, nbins, fd));
        } else {
            CK(launch_extract_signmap(L, bm, g.ld, nimg, g.lay(), d_bins, nhdr, 3, d_out_bytes, d_raw, nbins, fd));
            CK(launch_extract_signmap(L, bm, g.ld, nimg, g.lay(), d_bins + nhdr, nbins - nhdr, 7, d_out_payload, d_raw ? d_raw + nhdr : nullptr, nbins, fd));
        }
        return TFFT_OK;
    }
    ProfScope ps(ctx, L.stream, TFFT_K_EXTRACT, (double)nimg * (double)nbins * (16.0 + 4.0));
    if (nhdr == 0) {
        CK(launch_extract(L, spec, nimg, lay, d_bins, nbins, rep, d_jitter, alpha, d_out_bytes, d_raw, nbins, amed));
    } else {
        CK(launch_extract(L, spec, nimg, lay, d_bins, nhdr, 3, d_jitter, alpha, d_out_bytes, d_raw, nbins, amed));
        CK(launch_extract(L, spec, nimg, lay, d_bins + nhdr, nbins - nhdr, 7, d_jitter ? d_jitter + nhdr : nullptr, alpha,
                          d_out_payload, d_raw ? d_raw + nhdr : nullptr, nbins, amed));
    }
    return TFFT_OK;
}

int upload_bins(tfft_ctx* ctx, const uint32_t* bins, size_t nbins, const double* jitter, cudaStream_t s) {
    int rc;
    if (nbins == 0) return TFFT_OK;
    if ((rc = ensure(ctx, ctx->bins, sizeof(uint32_t) * nbins))) return rc;
    CK(cudaMemcpyAsync(ctx->bins.p, bins, sizeof(uint32_t) * nbins, cudaMemcpyHostToDevice, s));
    if (jitter) {
        if ((rc = ensure(ctx, ctx->jitter, sizeof(double) * nbins))) return rc;
        CK(cudaMemcpyAsync(ctx->jitter.p, jitter, sizeof(double) * nbins, cudaMemcpyHostToDevice, s));
    }
    CK(cudaStreamSynchronize(s));  // both slot streams read the shared bin list
    return TFFT_OK;
}

inline size_t dec_bytes(size_t nbins, int rep) { return (nbins / (size_t)rep + 7) / 8; }

// Chunk sizes of the host pipeline: a short ramp (chunk/4, chunk/2) fills the H2D -> kernels -> D2H pipeline quickly
// and a short last chunk drains it quickly; everything in between is `chunk` images.
inline int next_chunk(int i0, int n, int chunk, int done_chunks) {
    const int rem = n - i0, q = std::max(1, chunk / 4), h = std::max(1, chunk / 2);
    if (n <= 2 * chunk) return std::min(chunk, rem);       // small batches: nothing to ramp
    if (done_chunks == 0) return q;
    if (done_chunks == 1) return h;
    if (rem <= chunk + q) return rem > q ? rem - q : rem;   // tail: (rem - q) then q
    return chunk;
}

// host-side sanity of a bin list: plane in 0..2 and linear index inside the padded plane
bool bins_ok(const uint32_t* bins, size_t n, size_t P) {
    for (size_t i = 0; i < n; i++)
        if ((bins[i] >> 30) > 2 || (size_t)(bins[i] & 0x3FFFFFFFu) >= P) return false;
    return true;
}
// ... plus what the column-resident embed needs: every bin strictly between column 0 and the Nyquist column (the half
// layout then stores exactly the bin itself, never its mirror) and the largest row block
bool bins_ok_fused(const uint32_t* bins, size_t n, const Geom& g, bool& fusable, int& k3max) {
    uint32_t bad = 0, unf = 0, ry = 0;
    const uint32_t P = (uint32_t)std::min<size_t>(g.P, 0x40000000u), xmask = (uint32_t)(g.PW - 1), half = (uint32_t)(g.PW >> 1);
    for (size_t i = 0; i < n; i++) {
        const uint32_t lin = bins[i] & 0x3FFFFFFFu, x = lin & xmask;
        bad |= (uint32_t)((bins[i] >> 30) > 2) | (uint32_t)(lin >= P);
        unf |= (uint32_t)(x == 0) | (uint32_t)(x >= half);
        ry = std::max(ry, lin >> g.lw);
    }
    fusable = !bad && !unf;
    k3max = (int)(ry >> 8);
    return !bad;
}
// the same check plus the window of the workspace the list touches (same rule as the bins_window kernel)
bool bins_ok_window(const uint32_t* bins, size_t n, const Geom& g, BinWindow& w) {
    uint32_t bad = 0, ry = 0, rx = 0;
    const uint32_t P = (uint32_t)std::min<size_t>(g.P, 0x40000000u), xmask = (uint32_t)(g.PW - 1);
    const int lw = g.lw;
    for (size_t i = 0; i < n; i++) {  // branch-free: largest row and column as listed
        const uint32_t lin = bins[i] & 0x3FFFFFFFu;
        bad |= (uint32_t)((bins[i] >> 30) > 2) | (uint32_t)(lin >= P);
        ry = std::max(ry, lin >> lw);
        rx = std::max(rx, lin & xmask);
    }
    w.rows = (int)ry + 1; w.cols = (int)rx + 1; w.mirrored = 0;
    if (n == 0) { w.rows = 0; w.cols = 0; }
    if (g.half && (int)rx > (g.PW >> 1) && !bad) {  // some bin sits behind the Nyquist column: redo with the mirror rule
        int my = 0, mx = 0;
        for (size_t i = 0; i < n; i++) {
            const uint32_t lin = bins[i] & 0x3FFFFFFFu;
            int y = (int)(lin >> lw), x = (int)(lin & xmask);
            if (x > (g.PW >> 1)) { y = (g.PH - y) & (g.PH - 1); x = g.PW - x; }
            my = std::max(my, y + 1);
            mx = std::max(mx, x + 1);
        }
        w.rows = my; w.cols = mx; w.mirrored = 1;
    }
    return !bad;
}
// device bin list: reduce on the device, read the two numbers back (one stream synchronisation per call)
int bins_window_dev(tfft_ctx* ctx, const Launcher& L, const uint32_t* d_bins, size_t nbins, const Geom& g, BinWindow& w) {
    w = BinWindow{};
    if (!ctx->use_window || nbins == 0 || g.large) return TFFT_OK;
    CK(launch_bins_window(L, d_bins, nbins, g.lay(), ctx->d_win));
    CK(cudaMemcpyAsync(ctx->h_win, ctx->d_win, 3 * sizeof(unsigned), cudaMemcpyDeviceToHost, L.stream));
    CK(cudaStreamSynchronize(L.stream));
    w.rows = (int)ctx->h_win[0]; w.cols = (int)ctx->h_win[1]; w.mirrored = (int)ctx->h_win[2];
    return TFFT_OK;
}

}  // namespace

// =============================================================================================
extern "C" {

int tfft_abi_version(void) { return TFFT_ABI_VERSION; }

const char* tfft_strerror(int code) {
    switch (code) {
        case TFFT_OK: return "ok";
        case TFFT_E_INVALID: return "invalid argument";
        case TFFT_E_CUDA: return "CUDA error";
        case TFFT_E_CAPACITY: return "message too large for the cover's capacity";
        case TFFT_E_NOMEM: return "device workspace does not fit";
        case TFFT_E_UNSUPPORTED: return "padded image dimension not supported";
        case TFFT_E_STATE: return "no resident spectra (call tfft_forward_batch first)";
        default: return "unknown error";
    }
}

const char* tfft_last_cuda_error(const tfft_ctx* ctx) { return ctx ? ctx->cuda_err : ""; }
uint64_t tfft_launch_count(const tfft_ctx* ctx) { return ctx ? ctx->launches : 0; }

void* tfft_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void tfft_host_free(void* p) { if (p) cudaFreeHost(p); }

int tfft_create(int device, tfft_ctx** out) {
    if (!out) return TFFT_E_INVALID;
    *out = nullptr;
    tfft_ctx* ctx = new tfft_ctx();
    ctx->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { delete ctx; cudaGetLastError(); return TFFT_E_CUDA; }
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device); ctx->sm_count = v;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device); ctx->smem_optin = (size_t)v;
    size_t fr = 0, tot = 0;
    cudaMemGetInfo(&fr, &tot);
    ctx->total_mem = tot;
    ctx->ws_limit = (size_t)((double)tot * 0.40);
    const char* impl = getenv("TFFT_FFT_IMPL");  // "v0" forces the baseline shared-memory kernel
    ctx->fft_impl = (impl && !strcmp(impl, "v0")) ? 0 : 1;
    if (const char* hc = getenv("TFFT_HOST_CHUNK")) { int v = atoi(hc); if (v >= 1 && v <= MAX_CHUNK) HOST_CHUNK = v; }
    if (const char* hs = getenv("TFFT_HOST_SLOTS")) { int v = atoi(hs); if (v >= 1 && v <= NSLOT) HOST_SLOTS = v; }
    const char* spc = getenv("TFFT_SPECTRUM");  // "full" keeps the complete PH x PW spectrum (no Hermitian halving)
    ctx->use_half = !(spc && !strcmp(spc, "full"));
    if (const char* wd = getenv("TFFT_WIDE")) ctx->use_wide = atoi(wd) != 0;
    if (const char* ew = getenv("TFFT_EXTRACT_WINDOW")) ctx->use_window = atoi(ew) != 0;
    if (const char* sm = getenv("TFFT_SIGNMAP")) ctx->use_signmap = atoi(sm) != 0;
    if (const char* fe = getenv("TFFT_FUSED_EMBED")) ctx->use_fused = atoi(fe) != 0;
    for (int i = 0; i < NSLOT; i++)
        if ((e = cudaStreamCreateWithFlags(&ctx->slot[i].stream, cudaStreamNonBlocking)) != cudaSuccess) break;
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_tw, sizeof(double2) * (TW_N / 2));
    if (e == cudaSuccess) e = build_twiddles(ctx->d_tw, ctx->slot[0].stream);
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_win, 3 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&ctx->h_win, 3 * sizeof(unsigned), cudaHostAllocDefault);
    if (e != cudaSuccess) { tfft_destroy(ctx); cudaGetLastError(); return TFFT_E_CUDA; }
    *out = ctx;
    return TFFT_OK;
}

void tfft_destroy(tfft_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < NSLOT; i++) {
        Slot& S = ctx->slot[i];
        release(S.spec); release(S.spec2); release(S.in); release(S.out); release(S.bits); release(S.bitsp); release(S.med);
        release(S.medians); release(S.usable); release(S.outbytes); release(S.raw); release(S.signmap); release(S.q32); release(S.val);
        if (S.h_stage) cudaFreeHost(S.h_stage);
        if (S.stream) cudaStreamDestroy(S.stream);
    }
    release(ctx->bins); release(ctx->jitter); release(ctx->full); release(ctx->pres); release(ctx->slab_z); release(ctx->slab_tmp);
    prof_drain(ctx);
    for (cudaEvent_t ev : ctx->prof_pool) cudaEventDestroy(ev);
    if (ctx->d_tw) cudaFree(ctx->d_tw);
    if (ctx->d_win) cudaFree(ctx->d_win);
    if (ctx->h_win) cudaFreeHost(ctx->h_win);
    delete ctx;
}

int tfft_profile_enable(tfft_ctx* ctx, int on) {
    if (!ctx) return TFFT_E_INVALID;
    ctx->prof_on = on != 0;
    return TFFT_OK;
}
int tfft_profile_reset(tfft_ctx* ctx) {
    if (!ctx) return TFFT_E_INVALID;
    prof_drain(ctx);
    for (int k = 0; k < TFFT_K_COUNT; k++) { ctx->prof_ms[k] = 0; ctx->prof_groups[k] = 0; ctx->prof_bytes[k] = 0; }
    return TFFT_OK;
}
int tfft_profile_read(tfft_ctx* ctx, int kind, uint64_t* groups, double* total_ms, double* total_bytes) {
    if (!ctx || kind < 0 || kind >= TFFT_K_COUNT) return TFFT_E_INVALID;
    prof_drain(ctx);
    if (groups) *groups = ctx->prof_groups[kind];
    if (total_ms) *total_ms = ctx->prof_ms[kind];
    if (total_bytes) *total_bytes = ctx->prof_bytes[kind];
    return TFFT_OK;
}
const char* tfft_kind_name(int kind) {
    static const char* names[TFFT_K_COUNT] = {"row_fwd_u8", "col_fwd", "median_capacity", "embed_scatter",
                                              "col_inv", "row_inv_u8", "extract_vote", "c2c_pass", "col_fwd_window", "col_embed_fused", "slab_glue"};
    return (kind >= 0 && kind < TFFT_K_COUNT) ? names[kind] : "?";
}

int tfft_set_adaptive_alpha(tfft_ctx* ctx, int on) {
    if (!ctx) return TFFT_E_INVALID;
    ctx->adaptive = on != 0;
    ctx->res_n = 0;  // resident spectra carry no medians
    return TFFT_OK;
}

int tfft_set_workspace_limit(tfft_ctx* ctx, size_t bytes) {
    if (!ctx || bytes == 0) return TFFT_E_INVALID;
    ctx->ws_limit = bytes;
    return TFFT_OK;
}

// ---------------------------------------------------------------------------------------------
int tfft_embed_batch_dev(tfft_ctx* ctx, const uint8_t* d_cover, int n, int W, int H,
                         const uint32_t* d_bins, const uint8_t* d_bits, size_t nbits,
                         const double* d_jitter, double alpha, int center, double magmin,
                         double rmin, double rmax, uint8_t* d_stego, uint64_t* d_usable,
                         double* d_median, void* stream) {
    if (!ctx || !d_cover || !d_stego || n < 0 || (nbits && (!d_bins || !d_bits))) return TFFT_E_INVALID;
    Geom g;
    int rc = make_geom(ctx, W, H, g);
    if (rc) return rc;
    if (n == 0) return TFFT_OK;
    CK(cudaSetDevice(ctx->device));
    ctx->res_n = 0;
    const int chunk = chunk_for(ctx, g, n, 1);
    Slot& S = ctx->slot[0];
    if ((rc = ensure_slot(ctx, S, g, chunk, false, 0, 0, 0))) return rc;
    Launcher L = make_launcher(ctx, (cudaStream_t)stream);
    EmbedPlan plan;
    if (fused_geometry_ok(ctx, L, g, d_jitter)) {
        // the bin list lives on the device: one reduction + ONE stream synchronisation per call decides (as the extracts do)
        plan.fused = true;
        if (nbits) {
            CK(launch_embed_bins_check(L, d_bins, nbits, g.lay(), ctx->d_win));
            CK(cudaMemcpyAsync(ctx->h_win, ctx->d_win, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, L.stream));
            CK(cudaStreamSynchronize(L.stream));
            plan.fused = ctx->h_win[0] == 0;
            plan.k3max = (int)(ctx->h_win[1] >> 8);
            if (plan.fused) {
                if ((rc = ensure(ctx, ctx->pres, embed_pres_bytes(g.ld)))) return rc;
                CK(launch_embed_pres(L, d_bins, nbits, g.lay(), (uint16_t*)ctx->pres.p));
            }
        }
    }
    for (int i0 = 0; i0 < n; i0 += chunk) {
        const int m = std::min(chunk, n - i0);
        uint64_t* us = d_usable ? d_usable + i0 : (uint64_t*)S.usable.p;
        double* med = d_median ? d_median + (size_t)i0 * 3 : (double*)S.medians.p;
        rc = embed_chunk(ctx, L, S, d_cover + (size_t)i0 * g.img_bytes, m, g, d_bins, d_bits + (size_t)i0 * nbits, nbits,
                         d_jitter, alpha, center, magmin, rmin, rmax, d_stego + (size_t)i0 * g.img_bytes, us, med, plan);
        if (rc) return rc;
    }
    return TFFT_OK;
}

// packed: bits = [n][ceil(nbits / 8)] MSB first instead of [n][nbits] one bit per byte
static int embed_host_impl(tfft_ctx* ctx, const uint8_t* cover, int n, int W, int H,
                           const uint32_t* bins, const uint8_t* bits, bool packed, size_t nbits, const double* jitter,
                           double alpha, int center, double magmin, double rmin, double rmax,
                           uint8_t* stego, uint64_t* usable, double* median) {
    if (!ctx || !cover || !stego || n < 0 || (nbits && (!bins || !bits))) return TFFT_E_INVALID;
    const size_t pbytes = (nbits + 7) / 8;
    Geom g;
    int rc = make_geom(ctx, W, H, g);
    if (rc) return rc;
    if (n == 0) return TFFT_OK;
    bool fusable = false;
    EmbedPlan plan;
    if (!bins_ok_fused(bins, nbits, g, fusable, plan.k3max)) return TFFT_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    ctx->res_n = 0;
    plan.fused = fusable && fused_geometry_ok(ctx, make_launcher(ctx, ctx->slot[0].stream), g, jitter);
    const int chunk = std::min(chunk_for(ctx, g, n, HOST_SLOTS), HOST_CHUNK);
    const int nslots = std::min(HOST_SLOTS, (n + chunk - 1) / chunk);
    for (int s = 0; s < nslots; s++) {
        if ((rc = ensure_slot(ctx, ctx->slot[s], g, chunk, true, nbits, 0, 0))) return rc;
        if (packed && nbits && (rc = ensure(ctx, ctx->slot[s].bitsp, (size_t)chunk * pbytes))) return rc;
    }
    if ((rc = upload_bins(ctx, bins, nbits, jitter, ctx->slot[0].stream))) return rc;
    if (plan.fused && nbits) {  // bin-presence masks of the fused pass: once per call, read by every slot stream
        if ((rc = ensure(ctx, ctx->pres, embed_pres_bytes(g.ld)))) return rc;
        CK(launch_embed_pres(make_launcher(ctx, ctx->slot[0].stream), (const uint32_t*)ctx->bins.p, nbits, g.lay(), (uint16_t*)ctx->pres.p));
        CK(cudaStreamSynchronize(ctx->slot[0].stream));
    }
    const size_t stage_bytes = (size_t)chunk * (sizeof(uint64_t) + 3 * sizeof(double));
    for (int s = 0; s < nslots; s++)
        if ((rc = ensure_stage(ctx, ctx->slot[s], stage_bytes))) return rc;
    bool over = false;
    int pend_i0[NSLOT], pend_m[NSLOT];
    for (int s = 0; s < NSLOT; s++) { pend_i0[s] = -1; pend_m[s] = 0; }
    auto drain = [&](int s) -> int {  // wait for the slot's chunk and hand its small results to the caller
        if (pend_i0[s] < 0) return TFFT_OK;
        Slot& S = ctx->slot[s];
        CK(cudaStreamSynchronize(S.stream));
        const uint64_t* hu = (const uint64_t*)S.h_stage;
        const double* hm = (const double*)(S.h_stage + (size_t)chunk * sizeof(uint64_t));
        for (int k = 0; k < pend_m[s]; k++) {
            if (usable) usable[pend_i0[s] + k] = hu[k];
            if (hu[k] < (uint64_t)nbits) over = true;
        }
        if (median) memcpy(median + (size_t)pend_i0[s] * 3, hm, sizeof(double) * 3 * pend_m[s]);
        pend_i0[s] = -1;
        return TFFT_OK;
    };
    int ci = 0;
    for (int i0 = 0, m = 0; i0 < n; i0 += m, ci++) {
        const int sl = ci % nslots;
        Slot& S = ctx->slot[sl];
        m = next_chunk(i0, n, chunk, ci);
        cudaStream_t st = S.stream;
        if ((rc = drain(sl))) return rc;  // the slot's buffers are about to be reused
        CK(cudaMemcpyAsync(S.in.p, cover + (size_t)i0 * g.img_bytes, (size_t)m * g.img_bytes, cudaMemcpyHostToDevice, st));
        Launcher L = make_launcher(ctx, st);
        if (nbits && packed) {
            CK(cudaMemcpyAsync(S.bitsp.p, bits + (size_t)i0 * pbytes, (size_t)m * pbytes, cudaMemcpyHostToDevice, st));
            CK(launch_unpack_bits(L, (const uint8_t*)S.bitsp.p, pbytes, (uint8_t*)S.bits.p, nbits, m));
        } else if (nbits) {
            CK(cudaMemcpyAsync(S.bits.p, bits + (size_t)i0 * nbits, (size_t)m * nbits, cudaMemcpyHostToDevice, st));
        }
        rc = embed_chunk(ctx, L, S, (const uint8_t*)S.in.p, m, g, (const uint32_t*)ctx->bins.p, (const uint8_t*)S.bits.p, nbits,
                         jitter ? (const double*)ctx->jitter.p : nullptr, alpha, center, magmin, rmin, rmax,
                         (uint8_t*)S.out.p, (uint64_t*)S.usable.p, (double*)S.medians.p, plan);
        if (rc) return rc;
        CK(cudaMemcpyAsync(stego + (size_t)i0 * g.img_bytes, S.out.p, (size_t)m * g.img_bytes, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(S.h_stage, S.usable.p, sizeof(uint64_t) * m, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(S.h_stage + (size_t)chunk * sizeof(uint64_t), S.medians.p, sizeof(double) * 3 * m, cudaMemcpyDeviceToHost, st));
        pend_i0[sl] = i0; pend_m[sl] = m;
    }
    for (int k = 0; k < nslots; k++)  // oldest first
        if ((rc = drain((ci + k) % nslots))) return rc;
    return over ? TFFT_E_CAPACITY : TFFT_OK;
}

int tfft_embed_batch(tfft_ctx* ctx, const uint8_t* cover, int n, int W, int H,
                     const uint32_t* bins, const uint8_t* bits, size_t nbits, const double* jitter,
                     double alpha, int center, double magmin, double rmin, double rmax,
                     uint8_t* stego, uint64_t* usable, double* median) {
    return embed_host_impl(ctx, cover, n, W, H, bins, bits, false, nbits, jitter, alpha, center, magmin, rmin, rmax, stego, usable, median);
}
int tfft_embed_batch_packed(tfft_ctx* ctx, const uint8_t* cover, int n, int W, int H,
                            const uint32_t* bins, const uint8_t* bits_packed, size_t nbits, const double* jitter,
                            double alpha, int center, double magmin, double rmin, double rmax,
                            uint8_t* stego, uint64_t* usable, double* median) {
    return embed_host_impl(ctx, cover, n, W, H, bins, bits_packed, true, nbits, jitter, alpha, center, magmin, rmin, rmax, stego, usable, median);
}

// ---------------------------------------------------------------------------------------------
static int extract_dev_impl(tfft_ctx* ctx, const uint8_t* d_stego, int n, int W, int H, const uint32_t* d_bins, size_t nbins,
                            int rep, size_t nhdr, const double* d_jitter, double alpha, int center, uint8_t* d_out,
                            uint8_t* d_out_payload, uint8_t* d_raw_bits, void* stream) {
    if (!ctx || !d_stego || n < 0 || (nbins && !d_bins) || nhdr > nbins) return TFFT_E_INVALID;
    Geom g;
    int rc = make_geom(ctx, W, H, g);
    if (rc) return rc;
    if (n == 0) return TFFT_OK;
    CK(cudaSetDevice(ctx->device));
    ctx->res_n = 0;
    const int chunk = chunk_for(ctx, g, n, 1);
    Slot& S = ctx->slot[0];
    if ((rc = ensure_slot(ctx, S, g, chunk, false, 0, 0, 0))) return rc;
    Launcher L = make_launcher(ctx, (cudaStream_t)stream);
    const size_t nb = nhdr ? dec_bytes(nhdr, 3) : dec_bytes(nbins, rep);
    const size_t nbp = nhdr ? dec_bytes(nbins - nhdr, 7) : 0;
    BinWindow win;
    if ((rc = bins_window_dev(ctx, L, d_bins, nbins, g, win))) return rc;
    for (int i0 = 0; i0 < n; i0 += chunk) {
        const int m = std::min(chunk, n - i0);
        rc = extract_chunk(ctx, L, S, d_stego + (size_t)i0 * g.img_bytes, m, g, d_bins, nbins, rep, nhdr, d_jitter, alpha, center,
                           d_out ? d_out + (size_t)i0 * nb : nullptr,
                           d_out_payload ? d_out_payload + (size_t)i0 * nbp : nullptr,
                           d_raw_bits ? d_raw_bits + (size_t)i0 * nbins : nullptr, win);
        if (rc) return rc;
    }
    return TFFT_OK;
}

static int extract_host_impl(tfft_ctx* ctx, const uint8_t* stego, int n, int W, int H, const uint32_t* bins, size_t nbins,
                             int rep, size_t nhdr, const double* jitter, double alpha, int center, uint8_t* out,
                             uint8_t* out_payload, uint8_t* raw_bits) {
    if (!ctx || !stego || n < 0 || (nbins && !bins) || nhdr > nbins) return TFFT_E_INVALID;
    Geom g;
    int rc = make_geom(ctx, W, H, g);
    if (rc) return rc;
    if (n == 0) return TFFT_OK;
    BinWindow win;
    if (!bins_ok_window(bins, nbins, g, win)) return TFFT_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    ctx->res_n = 0;
    const int chunk = std::min(chunk_for(ctx, g, n, HOST_SLOTS), HOST_CHUNK);
    const int nslots = std::min(HOST_SLOTS, (n + chunk - 1) / chunk);
    const size_t nb = nhdr ? dec_bytes(nhdr, 3) : dec_bytes(nbins, rep);
    const size_t nbp = nhdr ? dec_bytes(nbins - nhdr, 7) : 0;
    // header bytes and payload bytes share one device buffer per slot: [chunk][nb] then [chunk][nbp]
    for (int s = 0; s < nslots; s++)
        if ((rc = ensure_slot(ctx, ctx->slot[s], g, chunk, true, 0, (out ? nb : 0) + (out_payload ? nbp : 0), raw_bits ? nbins : 0))) return rc;
    if ((rc = upload_bins(ctx, bins, nbins, jitter, ctx->slot[0].stream))) return rc;
    const size_t per_img = (out ? nb : 0) + (out_payload ? nbp : 0);
    for (int s = 0; s < nslots; s++)
        if (per_img && (rc = ensure_stage(ctx, ctx->slot[s], (size_t)chunk * per_img))) return rc;
    int pend_i0[NSLOT], pend_m[NSLOT];
    for (int s = 0; s < NSLOT; s++) { pend_i0[s] = -1; pend_m[s] = 0; }
    auto drain = [&](int s) -> int {
        if (pend_i0[s] < 0) return TFFT_OK;
        Slot& S = ctx->slot[s];
        CK(cudaStreamSynchronize(S.stream));
        if (out && nb) memcpy(out + (size_t)pend_i0[s] * nb, S.h_stage, (size_t)pend_m[s] * nb);
        if (out_payload && nbp) memcpy(out_payload + (size_t)pend_i0[s] * nbp, S.h_stage + (out ? (size_t)chunk * nb : 0), (size_t)pend_m[s] * nbp);
        pend_i0[s] = -1;
        return TFFT_OK;
    };
    int ci = 0;
    for (int i0 = 0, m = 0; i0 < n; i0 += m, ci++) {
        const int sl = ci % nslots;
        Slot& S = ctx->slot[sl];
        m = next_chunk(i0, n, chunk, ci);
        cudaStream_t st = S.stream;
        if ((rc = drain(sl))) return rc;
        CK(cudaMemcpyAsync(S.in.p, stego + (size_t)i0 * g.img_bytes, (size_t)m * g.img_bytes, cudaMemcpyHostToDevice, st));
        Launcher L = make_launcher(ctx, st);
        uint8_t* d_out = out ? (uint8_t*)S.outbytes.p : nullptr;
        uint8_t* d_pay = out_payload ? (uint8_t*)S.outbytes.p + (out ? (size_t)chunk * nb : 0) : nullptr;
        rc = extract_chunk(ctx, L, S, (const uint8_t*)S.in.p, m, g, (const uint32_t*)ctx->bins.p, nbins, rep, nhdr,
                           jitter ? (const double*)ctx->jitter.p : nullptr, alpha, center, d_out, d_pay,
                           raw_bits ? (uint8_t*)S.raw.p : nullptr, win);
        if (rc) return rc;
        if (d_out && nb) CK(cudaMemcpyAsync(S.h_stage, d_out, (size_t)m * nb, cudaMemcpyDeviceToHost, st));
        if (d_pay && nbp) CK(cudaMemcpyAsync(S.h_stage + (out ? (size_t)chunk * nb : 0), d_pay, (size_t)m * nbp, cudaMemcpyDeviceToHost, st));
        if (raw_bits && nbins) CK(cudaMemcpyAsync(raw_bits + (size_t)i0 * nbins, S.raw.p, (size_t)m * nbins, cudaMemcpyDeviceToHost, st));
        pend_i0[sl] = i0; pend_m[sl] = m;
    }
    for (int k = 0; k < nslots; k++)
        if ((rc = drain((ci + k) % nslots))) return rc;
    return TFFT_OK;
}

int tfft_extract_bits_dev(tfft_ctx* ctx, const uint8_t* d_stego, int n, int W, int H,
                          const uint32_t* d_bins, size_t nbins, int rep, const double* d_jitter,
                          double alpha, int center, uint8_t* d_out_bytes, uint8_t* d_raw_bits, void* stream) {
    if (!(rep == 1 || rep == 3 || rep == 7)) return TFFT_E_INVALID;
    return extract_dev_impl(ctx, d_stego, n, W, H, d_bins, nbins, rep, 0, d_jitter, alpha, center, d_out_bytes, nullptr, d_raw_bits, stream);
}
int tfft_extract_bits(tfft_ctx* ctx, const uint8_t* stego, int n, int W, int H,
                      const uint32_t* bins, size_t nbins, int rep, const double* jitter,
                      double alpha, int center, uint8_t* out_bytes, uint8_t* raw_bits) {
    if (!(rep == 1 || rep == 3 || rep == 7)) return TFFT_E_INVALID;
    return extract_host_impl(ctx, stego, n, W, H, bins, nbins, rep, 0, jitter, alpha, center, out_bytes, nullptr, raw_bits);
}
int tfft_extract_frame_dev(tfft_ctx* ctx, const uint8_t* d_stego, int n, int W, int H,
                           const uint32_t* d_bins, size_t nbins, size_t nhdr_bins, const double* d_jitter,
                           double alpha, int center, uint8_t* d_out_hdr, uint8_t* d_out_payload,
                           uint8_t* d_raw_bits, void* stream) {
    if (nhdr_bins == 0) return TFFT_E_INVALID;
    return extract_dev_impl(ctx, d_stego, n, W, H, d_bins, nbins, 3, nhdr_bins, d_jitter, alpha, center, d_out_hdr, d_out_payload, d_raw_bits, stream);
}
int tfft_extract_frame(tfft_ctx* ctx, const uint8_t* stego, int n, int W, int H,
                       const uint32_t* bins, size_t nbins, size_t nhdr_bins, const double* jitter,
                       double alpha, int center, uint8_t* out_hdr, uint8_t* out_payload, uint8_t* raw_bits) {
    if (nhdr_bins == 0) return TFFT_E_INVALID;
    return extract_host_impl(ctx, stego, n, W, H, bins, nbins, 3, nhdr_bins, jitter, alpha, center, out_hdr, out_payload, raw_bits);
}

// ---------------------------------------------------------------------------------------------
static int forward_batch_impl(tfft_ctx* ctx, const uint8_t* img, int n, int W, int H, int center, bool eager) {
    if (!ctx || !img || n <= 0) return TFFT_E_INVALID;
    Geom g;
    int rc = make_geom(ctx, W, H, g);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    ctx->res_n = 0;
    if ((size_t)n * 3 * g.E * sizeof(double2) * ((g.large || g.col4) ? 2 : 1) > ctx->ws_limit || n > MAX_CHUNK) return TFFT_E_NOMEM;
    Slot& S = ctx->slot[0];
    if ((rc = ensure_slot(ctx, S, g, n, true, 0, 0, 0))) return rc;
    CK(cudaMemcpyAsync(S.in.p, img, (size_t)n * g.img_bytes, cudaMemcpyHostToDevice, S.stream));
    Launcher L = make_launcher(ctx, S.stream);
    ctx->res_stage = 0; ctx->res_W = W; ctx->res_H = H;
    if (!eager && ctx->use_signmap && ctx->use_window && !ctx->adaptive && g.half && g.lh == 12 && !g.col4 && !g.large && signmap_supported(L)) {
        FwdOpts fo;
        fo.rows_only = true;
        if ((rc = forward_images(ctx, L, (double2*)S.spec.p, (double2*)S.spec2.p, (const uint8_t*)S.in.p, n, g, center, nullptr, fo))) return rc;
        CK(cudaStreamSynchronize(S.stream));
        ctx->res_spec = (double2*)S.spec.p;
        ctx->res_n = n; ctx->res_PH = g.PH; ctx->res_PW = g.PW; ctx->res_lay = g.lay(); ctx->res_stage = 1;
        return TFFT_OK;
    }
    if ((rc = forward_images(ctx, L, (double2*)S.spec.p, (double2*)S.spec2.p, (const uint8_t*)S.in.p, n, g, center, &ctx->res_spec))) return rc;
    if (ctx->adaptive) {  // S:1124: the medians of the stego spectra scale alpha in read_bit_from_bin
        MedianWork mw;
        median_work_carve(mw, S.med.p, n * 3, cand_cap_for(g.P));
        CK(launch_median_capacity(L, ctx->res_spec, n * 3, g.lay(), 0.0, 0.0, 0.0, mw, (double*)S.medians.p, nullptr));
    }
    CK(cudaStreamSynchronize(S.stream));
    ctx->res_n = n; ctx->res_PH = g.PH; ctx->res_PW = g.PW; ctx->res_lay = g.lay();
    return TFFT_OK;
}

int tfft_forward_batch(tfft_ctx* ctx, const uint8_t* img, int n, int W, int H, int center) {
    return forward_batch_impl(ctx, img, n, W, H, center, /*eager=*/false);
}

int tfft_read_bits(tfft_ctx* ctx, const uint32_t* bins, size_t nbins, int rep, const double* jitter,
                   double alpha, uint8_t* out_bytes, uint8_t* raw_bits) {
    if (!ctx || (nbins && !bins) || !(rep == 1 || rep == 3 || rep == 7)) return TFFT_E_INVALID;
    if (ctx->res_n <= 0) return TFFT_E_STATE;
    CK(cudaSetDevice(ctx->device));
    Slot& S = ctx->slot[0];
    const int n = ctx->res_n;
    const size_t P = (size_t)ctx->res_PH * ctx->res_PW;
    if (!bins_ok(bins, nbins, P)) return TFFT_E_INVALID;
    int rc;
    const size_t nb = dec_bytes(nbins, rep);
    if (out_bytes && nb && (rc = ensure(ctx, S.outbytes, (size_t)n * nb))) return rc;
    if (raw_bits && nbins && (rc = ensure(ctx, S.raw, (size_t)n * nbins))) return rc;
    if ((rc = upload_bins(ctx, bins, nbins, jitter, S.stream))) return rc;
    Launcher L = make_launcher(ctx, S.stream);
    if (ctx->res_stage != 0) {  // the column pass is still owed (or left a sign map): decide on this list
        Geom g;
        if ((rc = make_geom(ctx, ctx->res_W, ctx->res_H, g))) return rc;
        BinWindow win;
        if (!bins_ok_window(bins, nbins, g, win)) return TFFT_E_INVALID;
        const bool sign = !jitter && nbins > 0 && win.rows <= 2048 && !win.mirrored && alpha >= 1e-6 && alpha <= 3.14159;
        if (sign) {
            if (ctx->res_stage != 2 || ctx->res_alpha != alpha) {  // read bits of the whole quarter plane (rows < 2048), once per alpha
                if ((rc = ensure(ctx, S.signmap, (size_t)n * 3 * sign_map_words(g.ld) * sizeof(uint32_t)))) return rc;
                BinWindow q;
                q.rows = 2048; q.cols = g.ld;
                FwdOpts fo;
                fo.cols_only = true; fo.win = &q; fo.signmap = (uint32_t*)S.signmap.p; fo.alpha = alpha;
                if ((rc = forward_images(ctx, L, ctx->res_spec, (double2*)S.spec2.p, nullptr, n, g, 0, nullptr, fo))) return rc;
                ctx->res_stage = 2; ctx->res_alpha = alpha;
            }
            ProfScope ps(ctx, S.stream, TFFT_K_EXTRACT, (double)n * (double)nbins * (4.0 + 4.0));
            CK(launch_extract_signmap(L, (const uint32_t*)S.signmap.p, g.ld, n, g.lay(), (const uint32_t*)ctx->bins.p, nbins, rep,
                                      out_bytes ? (uint8_t*)S.outbytes.p : nullptr, raw_bits ? (uint8_t*)S.raw.p : nullptr, 0));
            if (out_bytes && nb) CK(cudaMemcpyAsync(out_bytes, S.outbytes.p, (size_t)n * nb, cudaMemcpyDeviceToHost, S.stream));
            if (raw_bits && nbins) CK(cudaMemcpyAsync(raw_bits, S.raw.p, (size_t)n * nbins, cudaMemcpyDeviceToHost, S.stream));
            CK(cudaStreamSynchronize(S.stream));
            return TFFT_OK;
        }
        // jitter, rows beyond the map, mirrored bins: the full column pass over the row-pass output, spectra from here on
        FwdOpts fo;
        fo.cols_only = true;
        if ((rc = forward_images(ctx, L, ctx->res_spec, (double2*)S.spec2.p, nullptr, n, g, 0, nullptr, fo))) return rc;
        ctx->res_stage = 0;
    }
    ProfScope ps(ctx, S.stream, TFFT_K_EXTRACT, (double)n * (double)nbins * (16.0 + 4.0));
    CK(launch_extract(L, (const double2*)ctx->res_spec, n, ctx->res_lay, (const uint32_t*)ctx->bins.p, nbins, rep,
                      jitter ? (const double*)ctx->jitter.p : nullptr, alpha,
                      out_bytes ? (uint8_t*)S.outbytes.p : nullptr, raw_bits ? (uint8_t*)S.raw.p : nullptr, 0,
                      ctx->adaptive ? (const double*)S.medians.p : nullptr));
    if (out_bytes && nb) CK(cudaMemcpyAsync(out_bytes, S.outbytes.p, (size_t)n * nb, cudaMemcpyDeviceToHost, S.stream));
    if (raw_bits && nbins) CK(cudaMemcpyAsync(raw_bits, S.raw.p, (size_t)n * nbins, cudaMemcpyDeviceToHost, S.stream));
    CK(cudaStreamSynchronize(S.stream));
    return TFFT_OK;
}

int tfft_forward_spectrum(tfft_ctx* ctx, const uint8_t* img, int W, int H, int center, double* out_c64) {
    if (!out_c64) return TFFT_E_INVALID;
    int rc = forward_batch_impl(ctx, img, 1, W, H, center, /*eager=*/true);
    if (rc) return rc;
    Slot& S = ctx->slot[0];
    const size_t full_bytes = 3 * (size_t)ctx->res_PH * ctx->res_PW * sizeof(double2);
    if (ctx->res_lay.half) {  // expand the half-spectrum workspace for the caller
        if ((rc = ensure(ctx, ctx->full, full_bytes))) return rc;
        Launcher L = make_launcher(ctx, S.stream);
        CK(launch_expand_half(L, (const double2*)ctx->res_spec, (double2*)ctx->full.p, 3, ctx->res_lay));
        CK(cudaStreamSynchronize(S.stream));
        CK(cudaMemcpy(out_c64, ctx->full.p, full_bytes, cudaMemcpyDeviceToHost));
    } else {
        CK(cudaMemcpy(out_c64, ctx->res_spec, full_bytes, cudaMemcpyDeviceToHost));
    }
    return TFFT_OK;
}

// ---------------------------------------------------------------------------------------------
static int hook_scratch(tfft_ctx* ctx, int nplanes, int PH, int PW, double2** tmp) {
    *tmp = nullptr;
    if (ctx->fft_impl == 0 || (PH <= 4096 && PW <= 4096)) return TFFT_OK;
    int rc = ensure(ctx, ctx->full, (size_t)nplanes * PH * PW * sizeof(double2));
    if (rc) return rc;
    *tmp = (double2*)ctx->full.p;
    return TFFT_OK;
}

static int fft2d_planes(tfft_ctx* ctx, const Launcher& L, double2* d, int nplanes, int PH, int PW, int inverse) {
    PassArgs a;
    memset(&a, 0, sizeof(a));
    { int rc = hook_scratch(ctx, nplanes, PH, PW, &a.tmp); if (rc) return rc; }
    a.spec = d; a.tw = ctx->d_tw; a.nplanes = nplanes;
    a.W = PW; a.H = PH; a.PW = PW; a.PH = PH;
    a.in_rows = PH; a.out_rows = PH; a.inverse = inverse; a.ld = PW;
    a.axis = 0; a.log2n = ilog2(PW);  // rows, then columns (S:361-365)
    { ProfScope ps(ctx, L.stream, TFFT_K_C2C, (double)a.nplanes * 32.0 * (double)PH * PW); CK(launch_fft_pass(L, a)); }
    a.axis = 1; a.log2n = ilog2(PH);
    { ProfScope ps(ctx, L.stream, TFFT_K_C2C, (double)a.nplanes * 32.0 * (double)PH * PW); CK(launch_fft_pass(L, a)); }
    return TFFT_OK;
}
static int check_dims(int PH, int PW) {
    if (PH < 2 || PW < 2 || (PH & (PH - 1)) || (PW & (PW - 1))) return TFFT_E_INVALID;
    if (PH > TFFT_MAX_DIM || PW > TFFT_MAX_DIM) return TFFT_E_UNSUPPORTED;
    return TFFT_OK;
}

int tfft_fft2d_dev(tfft_ctx* ctx, double* d_data, int n, int PH, int PW, int inverse, void* stream) {
    if (!ctx || !d_data || n < 0) return TFFT_E_INVALID;
    int rc = check_dims(PH, PW);
    if (rc) return rc;
    if (n == 0) return TFFT_OK;
    CK(cudaSetDevice(ctx->device));
    Launcher L = make_launcher(ctx, (cudaStream_t)stream);
    return fft2d_planes(ctx, L, (double2*)d_data, n, PH, PW, inverse);
}

int tfft_fft_pass_dev(tfft_ctx* ctx, double* d_data, int n, int PH, int PW, int axis, int inverse, void* stream) {
    if (!ctx || !d_data || n <= 0 || (axis != 0 && axis != 1)) return TFFT_E_INVALID;
    int rc = check_dims(PH, PW);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    Launcher L = make_launcher(ctx, (cudaStream_t)stream);
    PassArgs a;
    memset(&a, 0, sizeof(a));
    if ((rc = hook_scratch(ctx, n, PH, PW, &a.tmp))) return rc;
    a.spec = (double2*)d_data; a.tw = ctx->d_tw; a.nplanes = n;
    a.W = PW; a.H = PH; a.PW = PW; a.PH = PH;
    a.in_rows = PH; a.out_rows = PH; a.inverse = inverse; a.ld = PW;
    a.axis = axis; a.log2n = ilog2(axis == 0 ? PW : PH);
    { ProfScope ps(ctx, L.stream, TFFT_K_C2C, (double)a.nplanes * 32.0 * (double)PH * PW); CK(launch_fft_pass(L, a)); }
    return TFFT_OK;
}

int tfft_fft2d(tfft_ctx* ctx, double* data, int n, int PH, int PW, int inverse) {
    if (!ctx || !data || n < 0) return TFFT_E_INVALID;
    int rc = check_dims(PH, PW);
    if (rc) return rc;
    if (n == 0) return TFFT_OK;
    CK(cudaSetDevice(ctx->device));
    ctx->res_n = 0;
    const size_t P = (size_t)PH * PW, pb = P * sizeof(double2);
    size_t chunk = ctx->ws_limit / pb;
    if (chunk < 1) chunk = 1;
    if (chunk > (size_t)n) chunk = n;
    Slot& S = ctx->slot[0];
    if ((rc = ensure(ctx, S.spec, chunk * pb))) return rc;
    Launcher L = make_launcher(ctx, S.stream);
    for (size_t i0 = 0; i0 < (size_t)n; i0 += chunk) {
        const size_t m = std::min(chunk, (size_t)n - i0);
        CK(cudaMemcpyAsync(S.spec.p, data + i0 * P * 2, m * pb, cudaMemcpyHostToDevice, S.stream));
        if ((rc = fft2d_planes(ctx, L, (double2*)S.spec.p, (int)m, PH, PW, inverse))) return rc;
        CK(cudaMemcpyAsync(data + i0 * P * 2, S.spec.p, m * pb, cudaMemcpyDeviceToHost, S.stream));
        CK(cudaStreamSynchronize(S.stream));
    }
    return TFFT_OK;
}

// ---------------------------------------------------------------------------------------------
// config 5: slab-decomposed 2-D FFT (kernels in tfft_slab.cu)
namespace {
struct SlabGeom { int W, H, G, PW, PH, R, ld, cols, lw, lh, npairs; };
int slab_geom(int W, int H, int G, SlabGeom& s) {
    if (W <= 0 || H <= 0 || !(G == 1 || G == 2 || G == 4 || G == 8)) return TFFT_E_INVALID;
    s.W = W; s.H = H; s.G = G;
    s.PW = next_pow2_i(W); s.PH = next_pow2_i(H);
    if (s.PW > TFFT_MAX_DIM || s.PH > TFFT_MAX_DIM || s.PW < 512 || s.PH < 512) return TFFT_E_UNSUPPORTED;
    s.lw = ilog2(s.PW); s.lh = ilog2(s.PH);
    s.R = s.PH / G;
    s.ld = s.PW / 2 + 16;
    if (s.R < 2 || s.ld % G) return TFFT_E_UNSUPPORTED;
    s.cols = s.ld / G;
    s.npairs = s.R / 2;
    return TFFT_OK;
}
// one c2c pass over the pair rows [3][npairs][PW] along x; returns the buffer that holds the result (four-step passes
// leave it in the scratch batch instead of copying it back)
int slab_row_pass(tfft_ctx* ctx, const Launcher& L, const SlabGeom& s, int inverse, double2** where) {
    PassArgs a;
    memset(&a, 0, sizeof(a));
    a.spec = (double2*)ctx->slab_z.p; a.tmp = (double2*)ctx->slab_tmp.p; a.tw = ctx->d_tw;
    a.nplanes = 3; a.W = s.PW; a.H = s.npairs; a.PW = s.PW; a.PH = s.npairs;
    a.in_rows = s.npairs; a.out_rows = s.npairs; a.inverse = inverse; a.ld = s.PW;
    a.axis = 0; a.log2n = s.lw;
    const bool four = s.lw > 12;
    a.leave_in_tmp = four ? 1 : 0;
    ProfScope ps(ctx, L.stream, TFFT_K_C2C, 3.0 * 32.0 * (double)s.npairs * s.PW * (four ? 2.0 : 1.0));
    CK(launch_fft_pass(L, a));
    *where = four ? a.tmp : a.spec;
    return TFFT_OK;
}
int slab_ensure_rows(tfft_ctx* ctx, const SlabGeom& s) {
    int rc;
    const size_t b = (size_t)3 * s.npairs * s.PW * sizeof(double2);
    if ((rc = ensure(ctx, ctx->slab_z, b))) return rc;
    if (s.lw > 12 && (rc = ensure(ctx, ctx->slab_tmp, b))) return rc;
    return TFFT_OK;
}
}  // namespace

int tfft_slab_sizes(int W, int H, int G, int* PW, int* PH, int* R, int* ld, int* cols) {
    SlabGeom s;
    int rc = slab_geom(W, H, G, s);
    if (rc) return rc;
    if (PW) *PW = s.PW;
    if (PH) *PH = s.PH;
    if (R) *R = s.R;
    if (ld) *ld = s.ld;
    if (cols) *cols = s.cols;
    return TFFT_OK;
}

int tfft_slab_rows_forward_dev(tfft_ctx* ctx, const uint8_t* d_rows, int nrows, int W, int H, int G, int g, int center,
                               double* const* dst, size_t plane_stride, int row_base, void* stream) {
    SlabGeom s;
    int rc = slab_geom(W, H, G, s);
    if (rc) return rc;
    if (!ctx || !dst || g < 0 || g >= G || nrows < 0 || nrows > s.R || (nrows && !d_rows)) return TFFT_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    if ((rc = slab_ensure_rows(ctx, s))) return rc;
    Launcher L = make_launcher(ctx, (cudaStream_t)stream);
    double2* z = (double2*)ctx->slab_z.p;
    const double pz = 3.0 * 16.0 * (double)s.npairs * s.PW;
    { ProfScope ps(ctx, L.stream, TFFT_K_SLAB, (double)nrows * W * 3 + pz); CK(launch_slab_pack(L, d_rows, nrows, W, s.PW, s.npairs, g * s.R, center, z)); }
    double2* Z = nullptr;
    if ((rc = slab_row_pass(ctx, L, s, 0, &Z))) return rc;
    SlabDst d;
    for (int i = 0; i < SLAB_MAX_RANKS; i++) d.p[i] = i < G ? (double2*)dst[i] : nullptr;
    { ProfScope ps(ctx, L.stream, TFFT_K_SLAB, pz + 3.0 * 16.0 * (double)s.R * s.ld);
      CK(launch_slab_split(L, Z, s.PW, s.ld, s.cols, s.npairs, g * s.R, d, plane_stride, row_base)); }
    return TFFT_OK;
}

int tfft_slab_cols_dev(tfft_ctx* ctx, double* d_colslab, int W, int H, int G, int inverse, void* stream) {
    SlabGeom s;
    int rc = slab_geom(W, H, G, s);
    if (rc) return rc;
    if (!ctx || !d_colslab) return TFFT_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    Launcher L = make_launcher(ctx, (cudaStream_t)stream);
    PassArgs a;
    memset(&a, 0, sizeof(a));
    if (ctx->fft_impl != 0 && s.PH > 4096) {  // four-step columns: scratch batch of the slab's size
        if ((rc = ensure(ctx, ctx->full, (size_t)3 * s.PH * s.cols * sizeof(double2)))) return rc;
        a.tmp = (double2*)ctx->full.p;
    }
    a.spec = (double2*)d_colslab; a.tw = ctx->d_tw; a.nplanes = 3;
    a.W = s.cols; a.H = s.PH; a.PW = s.cols; a.PH = s.PH;  // (the slab is a plane of `cols` columns: any even count)
    a.in_rows = s.PH; a.out_rows = s.PH; a.inverse = inverse; a.ld = s.cols;
    a.axis = 1; a.log2n = s.lh;
    ProfScope ps(ctx, L.stream, TFFT_K_C2C, 3.0 * 32.0 * (double)s.PH * s.cols * (s.PH > 4096 ? 2.0 : 1.0));
    CK(launch_fft_pass(L, a));
    return TFFT_OK;
}

int tfft_slab_embed_dev(tfft_ctx* ctx, double* d_colslab, int W, int H, int G, int g, const uint32_t* d_bins,
                        const uint8_t* d_bits, size_t nbits, double alpha, void* stream) {
    SlabGeom s;
    int rc = slab_geom(W, H, G, s);
    if (rc) return rc;
    if (!ctx || !d_colslab || g < 0 || g >= G || (nbits && (!d_bins || !d_bits))) return TFFT_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    Launcher L = make_launcher(ctx, (cudaStream_t)stream);
    ProfScope ps(ctx, L.stream, TFFT_K_EMBED, (double)nbits * (16.0 + 16.0 + 5.0) / G);
    CK(launch_slab_embed(L, (double2*)d_colslab, s.PH, s.PW, s.cols, g * s.cols, d_bins, d_bits, nbits, cos(alpha), sin(alpha)));
    return TFFT_OK;
}

int tfft_slab_read_dev(tfft_ctx* ctx, const double* d_colslab, int W, int H, int G, int g, const uint32_t* d_bins,
                       size_t nbins, double alpha, int8_t* d_raw, void* stream) {
    SlabGeom s;
    int rc = slab_geom(W, H, G, s);
    if (rc) return rc;
    if (!ctx || !d_colslab || g < 0 || g >= G || (nbins && (!d_bins || !d_raw))) return TFFT_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    Launcher L = make_launcher(ctx, (cudaStream_t)stream);
    ProfScope ps(ctx, L.stream, TFFT_K_EXTRACT, (double)nbins * (16.0 / G + 5.0));
    CK(launch_slab_read(L, (const double2*)d_colslab, s.PH, s.PW, s.cols, g * s.cols, d_bins, nbins, alpha, d_raw));
    return TFFT_OK;
}

int tfft_slab_rows_inverse_dev(tfft_ctx* ctx, const double* d_tiles, int nrows, int W, int H, int G, int g, int center,
                               uint8_t* d_rows_out, void* stream) {
    SlabGeom s;
    int rc = slab_geom(W, H, G, s);
    if (rc) return rc;
    if (!ctx || !d_tiles || g < 0 || g >= G || nrows < 0 || nrows > s.R || (nrows && !d_rows_out)) return TFFT_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    if ((rc = slab_ensure_rows(ctx, s))) return rc;
    Launcher L = make_launcher(ctx, (cudaStream_t)stream);
    const double pz = 3.0 * 16.0 * (double)s.npairs * s.PW;
    { ProfScope ps(ctx, L.stream, TFFT_K_SLAB, pz + 3.0 * 16.0 * (double)s.R * s.ld);
      CK(launch_slab_merge(L, (const double2*)d_tiles, s.PW, s.cols, G, s.R, s.npairs, (double2*)ctx->slab_z.p)); }
    double2* Z = nullptr;
    if ((rc = slab_row_pass(ctx, L, s, 1, &Z))) return rc;
    { ProfScope ps(ctx, L.stream, TFFT_K_SLAB, pz + (double)nrows * W * 3); CK(launch_slab_to_u8(L, Z, nrows, W, s.PW, s.npairs, g * s.R, center, d_rows_out)); }
    return TFFT_OK;
}

int tfft_ipc_export(const void* d_ptr, unsigned char handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!d_ptr || !handle) return TFFT_E_INVALID;
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, (void*)d_ptr) != cudaSuccess) { cudaGetLastError(); return TFFT_E_CUDA; }
    memcpy(handle, &h, 64);
    return TFFT_OK;
}
int tfft_ipc_open(int device, const unsigned char handle[64], void** d_ptr) {
    if (!handle || !d_ptr) return TFFT_E_INVALID;
    *d_ptr = nullptr;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    if (cudaSetDevice(device) != cudaSuccess || cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        return TFFT_E_CUDA;
    }
    return TFFT_OK;
}
int tfft_ipc_close(void* d_ptr) {
    if (!d_ptr) return TFFT_E_INVALID;
    if (cudaIpcCloseMemHandle(d_ptr) != cudaSuccess) { cudaGetLastError(); return TFFT_E_CUDA; }
    return TFFT_OK;
}

int tfft_bin_window(const uint32_t* bins, size_t nbins, int W, int H, int half, int* rows, int* cols, int* mirrored) {
    if ((nbins && !bins) || W <= 0 || H <= 0) return TFFT_E_INVALID;
    Geom g;
    int rc = make_geom(nullptr, W, H, g);  // (no context: full layout) -- only the padded sizes are needed here
    if (rc) return rc;
    g.half = half ? 1 : 0;
    BinWindow w;
    if (!bins_ok_window(bins, nbins, g, w)) return TFFT_E_INVALID;
    if (rows) *rows = w.rows;
    if (cols) *cols = w.cols;
    if (mirrored) *mirrored = w.mirrored;
    return TFFT_OK;
}

int tfft_median_capacity_dev(tfft_ctx* ctx, const double* d_spec, int n, int PH, int PW, double magmin,
                             double rmin, double rmax, double* d_median, uint64_t* d_usable, void* stream) {
    if (!ctx || !d_spec || n <= 0 || !d_median) return TFFT_E_INVALID;
    int rc = check_dims(PH, PW);
    if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    Slot& S = ctx->slot[1];
    if ((rc = ensure(ctx, S.med, median_work_bytes(n * 3, cand_cap_for((size_t)PH * PW))))) return rc;
    MedianWork mw;
    median_work_carve(mw, S.med.p, n * 3, cand_cap_for((size_t)PH * PW));
    Launcher L = make_launcher(ctx, (cudaStream_t)stream);
    const int m = std::min(PH, PW);
    ProfScope ps(ctx, L.stream, TFFT_K_MEDIAN, (double)n * 3.0 * 16.0 * (double)PH * PW);
    CK(launch_median_capacity(L, (const double2*)d_spec, n * 3, SpecLayout{PH, PW, PW, 0}, magmin, rmin * m, rmax * m, mw, d_median, d_usable));
    return TFFT_OK;
}

}  // extern "C"
