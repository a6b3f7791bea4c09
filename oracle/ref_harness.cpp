// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Wraps the UNMODIFIED reference translation unit (steganosaurus/src/steganosaur.cpp,
// included from /root/reference at build time -- never copied into this repo) behind a
// tiny C ABI so that tests and bench.py's CPU-baseline leg can call the reference's own
// `static` functions directly.  Everything the reference does on the hot path is `static`
// in one TU, so `#define main` + `#include` is the only way to reach it.
//
// Built by oracle/Makefile into oracle/_ref/libtfft_ref.so (git-ignored, travels to the
// GPU box with the snapshot).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it.
#define main turtlefft_reference_main
#include "steganosaur.cpp"
#undef main

#include <chrono>

namespace {
using cplx = complex<double>;
using Plane = vector<vector<cplx>>;

Plane plane_from_flat(const double* f, int PH, int PW) {
    Plane F(PH, vector<cplx>(PW));
    for (int y = 0; y < PH; y++)
        for (int x = 0; x < PW; x++) {
            size_t i = ((size_t)y * PW + x) * 2;
            F[y][x] = cplx(f[i], f[i + 1]);
        }
    return F;
}
void plane_to_flat(const Plane& F, double* f) {
    const int PH = (int)F.size(), PW = (int)F[0].size();
    for (int y = 0; y < PH; y++)
        for (int x = 0; x < PW; x++) {
            size_t i = ((size_t)y * PW + x) * 2;
            f[i] = F[y][x].real();
            f[i + 1] = F[y][x].imag();
        }
}
// S:912-921 / S:1116-1123: planes -> centre -> pad -> forward fft2d.
void forward3(const uint8_t* img, int W, int H, bool center, int& PW, int& PH, Plane F[3],
              vector<double>* planes_out = nullptr) {
    vector<double> R, G, B;
    to_planes_u8(img, W, H, 3, R, G, B);
    apply_center(R, W, H, center);
    apply_center(G, W, H, center);
    apply_center(B, W, H, center);
    F[0] = pad_to_fft(R, W, H, PW, PH);
    F[1] = pad_to_fft(G, W, H, PW, PH);
    F[2] = pad_to_fft(B, W, H, PW, PH);
    for (int p = 0; p < 3; p++) fft2d(F[p], false);
    (void)planes_out;
}
inline void unpack_bin(uint32_t b, int PW, int& p, int& y, int& x) {
    p = (int)(b >> 30);
    uint32_t lin = b & 0x3FFFFFFFu;
    y = (int)(lin / (uint32_t)PW);
    x = (int)(lin % (uint32_t)PW);
}
double now_s() {
    return chrono::duration<double>(chrono::steady_clock::now().time_since_epoch()).count();
}
}  // namespace

extern "C" {

int ref_next_pow2(int v) { return (int)next_pow2((size_t)v); }

// fft1d S:341-358 on interleaved (re,im) doubles, in place.
void ref_fft1d(double* a, int n, int inverse) {
    vector<cplx> v(n);
    for (int i = 0; i < n; i++) v[i] = cplx(a[2 * i], a[2 * i + 1]);
    fft1d(v, inverse != 0);
    for (int i = 0; i < n; i++) { a[2 * i] = v[i].real(); a[2 * i + 1] = v[i].imag(); }
}
// fft2d S:359-366 on a flat [PH][PW][2] array, in place.
void ref_fft2d(double* f, int PH, int PW, int inverse) {
    Plane F = plane_from_flat(f, PH, PW);
    fft2d(F, inverse != 0);
    plane_to_flat(F, f);
}
// Forward spectra of the three planes: out is [3][PH][PW][2].
void ref_forward_spectrum(const uint8_t* img, int W, int H, int center, double* out) {
    int PW, PH;
    Plane F[3];
    forward3(img, W, H, center != 0, PW, PH, F);
    for (int p = 0; p < 3; p++) plane_to_flat(F[p], out + (size_t)p * PH * PW * 2);
}
double ref_median_abs(const double* f, int PH, int PW) {
    Plane F = plane_from_flat(f, PH, PW);
    return median_abs(F);
}
// Capacity count, restating the lambda at S:999-1007 around the reference's own helpers
// (on_axis, hypot_idx, conj_idx); the lambda itself is local to do_embed and not callable.
uint64_t ref_count_plane(const double* f, int PH, int PW, double rmin, double rmax, double thr) {
    Plane F = plane_from_flat(f, PH, PW);
    size_t c = 0;
    for (int y = 0; y < PH; y++)
        for (int x = 0; x < PW; x++) {
            if (on_axis(y, x, PH, PW)) continue;
            if (y == 0 && x == 0) continue;
            double r = hypot_idx(y, x);
            if (r < rmin * min(PH, PW) || r > rmax * min(PH, PW)) continue;
            if (abs(F[y][x]) < thr) continue;
            auto [cy, cx] = conj_idx(y, x, PH, PW);
            if (!(cy == y && cx == x)) c++;
        }
    return (uint64_t)(c / 2);
}

// The whole embed hot path (S:912-923, S:997-1008, S:1086, S:1100-1103) for given bins/bits.
// bins: plane<<30 | y*PW+x.  jitter must be 0 (KS::jitter(0) returns +-0).
// Optional outputs (may be NULL): medians[3], usable, spectrum_after [3][PH][PW][2].
// --adaptive_alpha (S:704-710, experimental upstream): when set, ref_embed / ref_extract_raw pass adaptive_alpha = true
// and the per-plane medians to write_bit_on_bin / read_bit_from_bin exactly as do_embed / do_extract do
static bool g_adaptive = false;
void ref_set_adaptive(int on) { g_adaptive = on != 0; }

void ref_embed(const uint8_t* cover, int W, int H, const uint32_t* bins, const uint8_t* bits,
               size_t nbits, double alpha, int center, double magmin, double rmin, double rmax,
               uint8_t* stego, double* medians, uint64_t* usable, double* spectrum_after) {
    int PW, PH;
    Plane F[3];
    forward3(cover, W, H, center != 0, PW, PH, F);
    double med[3];
    for (int p = 0; p < 3; p++) med[p] = median_abs(F[p]);
    if (medians) for (int p = 0; p < 3; p++) medians[p] = med[p];
    if (usable) {
        vector<double> flat((size_t)PH * PW * 2);
        uint64_t u = 0;
        for (int p = 0; p < 3; p++) {
            plane_to_flat(F[p], flat.data());
            u += ref_count_plane(flat.data(), PH, PW, rmin, rmax, magmin * med[p]);
        }
        *usable = u;
    }
    array<uint8_t, 32> zero_key{};
    KS dummy(zero_key);
    for (size_t i = 0; i < nbits; i++) {
        int p, y, x;
        unpack_bin(bins[i], PW, p, y, x);
        write_bit_on_bin(F[p], y, x, bits[i], alpha, 0.0, dummy, med[p], g_adaptive);
    }
    if (spectrum_after)
        for (int p = 0; p < 3; p++) plane_to_flat(F[p], spectrum_after + (size_t)p * PH * PW * 2);
    for (int p = 0; p < 3; p++) fft2d(F[p], true);
    auto R2 = ifft_crop(F[0], W, H), G2 = ifft_crop(F[1], W, H), B2 = ifft_crop(F[2], W, H);
    apply_center(R2, W, H, center != 0);
    apply_center(G2, W, H, center != 0);
    apply_center(B2, W, H, center != 0);
    vector<uint8_t> out;
    from_planes_u8(R2, G2, B2, W, H, out);
    memcpy(stego, out.data(), out.size());
}

// Raw phase bits (S:1123, S:1209) for given bins.
void ref_extract_raw(const uint8_t* stego, int W, int H, const uint32_t* bins, size_t nbins,
                     double alpha, int center, uint8_t* raw_bits) {
    int PW, PH;
    Plane F[3];
    forward3(stego, W, H, center != 0, PW, PH, F);
    double med[3] = {1.0, 1.0, 1.0};
    if (g_adaptive) for (int p = 0; p < 3; p++) med[p] = median_abs(F[p]);  // S:1124
    for (size_t i = 0; i < nbins; i++) {
        int p, y, x;
        unpack_bin(bins[i], PW, p, y, x);
        raw_bits[i] = (uint8_t)read_bit_from_bin(F[p], y, x, alpha, 0.0, med[p], g_adaptive);
    }
}
// Single-bin read (tie behaviour, S:734-746).
int ref_read_bit(double re, double im, double alpha, double jitter) {
    Plane F(1, vector<cplx>(1, cplx(re, im)));
    return read_bit_from_bin(F, 0, 0, alpha, jitter, 1.0, false);
}
// Single-bin write: returns new value of the bin (S:712-722).
void ref_write_bit(double re, double im, int bit, double alpha, double* out_re, double* out_im) {
    Plane F(4, vector<cplx>(4, cplx(0, 0)));
    F[1][1] = cplx(re, im);
    array<uint8_t, 32> zero_key{};
    KS dummy(zero_key);
    write_bit_on_bin(F, 1, 1, bit, alpha, 0.0, dummy, 1.0, false);
    *out_re = F[1][1].real();
    *out_im = F[1][1].imag();
}
// rep-3 / rep-7 majority + MSB-first pack (S:468-474, S:501-508, S:447-454).
// Returns number of bytes written.
size_t ref_rep_decode(const uint8_t* bits, size_t n, int rep, uint8_t* out_bytes) {
    vector<uint8_t> b(bits, bits + n);
    bool ok = true;
    vector<uint8_t> d = (rep == 3) ? rep3_decode_bits(b, ok) : (rep == 7) ? rep7_decode_bits(b, ok) : b;
    vector<uint8_t> by = bytes_from_bits(d);
    memcpy(out_bytes, by.data(), by.size());
    return by.size();
}
void ref_from_planes_u8(const double* R, const double* G, const double* B, int W, int H, uint8_t* out) {
    vector<double> r(R, R + (size_t)W * H), g(G, G + (size_t)W * H), b(B, B + (size_t)W * H);
    vector<uint8_t> o;
    from_planes_u8(r, g, b, W, H, o);
    memcpy(out, o.data(), o.size());
}

// ---------------------------------------------------------------- host-side pieces (L0/L2/L3)
void ref_sha256(const uint8_t* d, size_t n, uint8_t out[32]) {
    auto h = sha256::hash(d, n);
    memcpy(out, h.data(), 32);
}
void ref_hmac_sha256(const uint8_t* k, size_t kl, const uint8_t* m, size_t ml, uint8_t out[32]) {
    sha256::hmac_sha256(k, kl, m, ml, out);
}
void ref_hkdf_expand(const uint8_t prk[32], const uint8_t* info, size_t il, uint8_t* out, size_t L) {
    sha256::hkdf_sha256_expand(prk, info, il, out, L);
}
void ref_pbkdf2(const char* pass, size_t pl, const uint8_t* salt, size_t sl, uint32_t iters,
                uint8_t* out, size_t dk) {
    sha256::pbkdf2_hmac_sha256(string(pass, pl), vector<uint8_t>(salt, salt + sl), iters, out, dk);
}
void ref_derive_keys(const char* pass, size_t pl, const uint8_t salt[16], uint32_t iters,
                     uint8_t aead_key[32], uint8_t nonce[12]) {
    array<uint8_t, 16> s;
    memcpy(s.data(), salt, 16);
    KeyMaterial km = derive_keys(string(pass, pl), s, iters);
    memcpy(aead_key, km.aead_key.data(), 32);
    memcpy(nonce, km.nonce.data(), 12);
}
void ref_seal(const uint8_t key[32], const uint8_t nonce[12], const uint8_t* aad, size_t al,
              uint8_t* data, size_t n, uint8_t tag[16]) {
    chacha_poly::chacha20_poly1305_seal(key, nonce, aad, al, data, n, tag);
}
int ref_open(const uint8_t key[32], const uint8_t nonce[12], const uint8_t* aad, size_t al,
             uint8_t* data, size_t n, const uint8_t tag[16]) {
    return chacha_poly::chacha20_poly1305_open(key, nonce, aad, al, data, n, tag) ? 1 : 0;
}
// Turtle path keys (S:1038, S:1054-1061): out = key_walk | key_r | key_g | key_b (128 B).
void ref_turtle_keys(const char* pass, size_t pl, uint8_t path_key[32], uint8_t sub[128]) {
    auto pk = sha256::hash(string(pass, pl));
    memcpy(path_key, pk.data(), 32);
    const uint8_t info[] = "turtle_keys";
    sha256::hkdf_sha256_expand(pk.data(), info, sizeof(info) - 1, sub, 128);
}
// The embed/extract walk (S:1071-1097 without the write): emits nbits packed bins.
// Returns ks_walk.ctr at the end; start[3] receives the initial (plane,y,x).
uint32_t ref_walk(const char* pass, size_t pl, int PH, int PW, double rmin, double rmax,
                  double density, size_t nbits, uint32_t* bins, int start[3]) {
    uint8_t pk[32], sub[128];
    ref_turtle_keys(pass, pl, pk, sub);
    array<uint8_t, 32> kw, kr, kg, kb;
    memcpy(kw.data(), sub, 32);
    memcpy(kr.data(), sub + 32, 32);
    memcpy(kg.data(), sub + 64, 32);
    memcpy(kb.data(), sub + 96, 32);
    KS ks_walk(kw), ks_r(kr), ks_g(kg), ks_b(kb);
    array<KS*, 3> ks_planes = {&ks_r, &ks_g, &ks_b};
    vector<double> thr = {0, 0, 0};
    Turtle T(PH, PW, &ks_walk, ks_planes, rmin, rmax, nullptr, thr);
    if (start) { start[0] = T.plane; start[1] = T.y; start[2] = T.x; }
    for (size_t i = 0; i < nbits; i++) {
        while (true) {
            T.advance_to_valid();
            if (ks_walk.hit_density(density)) break;
            T.mark_here();
        }
        bins[i] = ((uint32_t)T.plane << 30) | (uint32_t)((size_t)T.y * PW + T.x);
        T.mark_here();
    }
    return ks_walk.ctr;
}
// Frame (S:946-995) with a caller-supplied salt (the CLI draws it from random_device).
// bits_out must hold 912 + 56*(slen+16) entries; header_out 38 B. Returns nbits.
size_t ref_frame_bits(const char* pass, size_t pl, const uint8_t salt[16], uint32_t iters,
                      const uint8_t* secret, size_t slen, uint8_t* bits_out, uint8_t* header_out) {
    array<uint8_t, 16> s;
    memcpy(s.data(), salt, 16);
    KeyMaterial km = derive_keys(string(pass, pl), s, iters);
    Header Hdr;
    Hdr.salt = km.salt;
    Hdr.nonce = km.nonce;
    Hdr.clen = (uint32_t)slen;
    vector<uint8_t> hb = Hdr.to_bytes();
    vector<uint8_t> ct(secret, secret + slen);
    array<uint8_t, 16> tag{};
    chacha_poly::chacha20_poly1305_seal(km.aead_key.data(), km.nonce.data(), hb.data(), hb.size(),
                                        ct.data(), ct.size(), tag.data());
    auto h3 = rep3_encode_bits(bits_from_bytes(hb));
    vector<uint8_t> pay(ct);
    pay.insert(pay.end(), tag.begin(), tag.end());
    auto p7 = rep7_encode_bits(bits_from_bytes(pay));
    memcpy(bits_out, h3.data(), h3.size());
    memcpy(bits_out + h3.size(), p7.data(), p7.size());
    if (header_out) memcpy(header_out, hb.data(), hb.size());
    return h3.size() + p7.size();
}

// ---------------------------------------------------------------- CPU baseline timing legs
// Times the reference's own hot-path stages exactly in do_embed's order (S:912-923, S:997-1008,
// S:1015 deep copy, S:1086 writes, S:1100-1103) -- PNG, KDF, walk excluded.
// Returns seconds; stego written to `stego`.
double ref_time_embed_hotpath(const uint8_t* cover, int W, int H, const uint32_t* bins,
                              const uint8_t* bits, size_t nbits, double alpha, int center,
                              double magmin, double rmin, double rmax, uint8_t* stego) {
    double t0 = now_s();
    vector<double> R, G, B;
    to_planes_u8(cover, W, H, 3, R, G, B);
    apply_center(R, W, H, center); apply_center(G, W, H, center); apply_center(B, W, H, center);
    int PW, PH;
    auto FR = pad_to_fft(R, W, H, PW, PH), FG = pad_to_fft(G, W, H, PW, PH), FB = pad_to_fft(B, W, H, PW, PH);
    fft2d(FR, false); fft2d(FG, false); fft2d(FB, false);
    double med[3] = {median_abs(FR), median_abs(FG), median_abs(FB)};
    size_t usable = 0;
    {
        const Plane* Fs[3] = {&FR, &FG, &FB};
        for (int p = 0; p < 3; p++) {
            const Plane& F = *Fs[p];
            double t = magmin * med[p];
            size_t c = 0;
            for (int y = 0; y < PH; y++) for (int x = 0; x < PW; x++) {
                if (on_axis(y, x, PH, PW)) continue;
                double r = hypot_idx(y, x);
                if (r < rmin * min(PH, PW) || r > rmax * min(PH, PW)) continue;
                if (abs(F[y][x]) < t) continue;
                auto [cy, cx] = conj_idx(y, x, PH, PW);
                if (!(cy == y && cx == x)) c++;
            }
            usable += c / 2;
        }
    }
    vector<Plane> F3 = {FR, FG, FB};  // S:1015 deep copy
    array<uint8_t, 32> zero_key{};
    KS dummy(zero_key);
    if (nbits <= usable)
        for (size_t i = 0; i < nbits; i++) {
            int p, y, x;
            unpack_bin(bins[i], PW, p, y, x);
            write_bit_on_bin(F3[p], y, x, bits[i], alpha, 0.0, dummy, med[p], false);
        }
    fft2d(F3[0], true); fft2d(F3[1], true); fft2d(F3[2], true);
    auto R2 = ifft_crop(F3[0], W, H), G2 = ifft_crop(F3[1], W, H), B2 = ifft_crop(F3[2], W, H);
    apply_center(R2, W, H, center); apply_center(G2, W, H, center); apply_center(B2, W, H, center);
    vector<uint8_t> out;
    from_planes_u8(R2, G2, B2, W, H, out);
    double t1 = now_s();
    memcpy(stego, out.data(), out.size());
    return t1 - t0;
}
// Extract hot path in do_extract's order (S:1116-1132, S:1209, S:1228, S:1266-1268).
// bins = 912 header bins followed by payload bins; decoded bytes to out (38 + npayload_bytes).
double ref_time_extract_hotpath(const uint8_t* stego, int W, int H, const uint32_t* bins,
                                size_t nbins, double alpha, int center, uint8_t* out_bytes) {
    double t0 = now_s();
    vector<double> R, G, B;
    to_planes_u8(stego, W, H, 3, R, G, B);
    apply_center(R, W, H, center); apply_center(G, W, H, center); apply_center(B, W, H, center);
    int PW, PH;
    auto FR = pad_to_fft(R, W, H, PW, PH), FG = pad_to_fft(G, W, H, PW, PH), FB = pad_to_fft(B, W, H, PW, PH);
    fft2d(FR, false); fft2d(FG, false); fft2d(FB, false);
    double med[3] = {median_abs(FR), median_abs(FG), median_abs(FB)};
    vector<Plane> F3 = {FR, FG, FB};  // S:1132
    vector<uint8_t> raw(nbins);
    for (size_t i = 0; i < nbins; i++) {
        int p, y, x;
        unpack_bin(bins[i], PW, p, y, x);
        raw[i] = (uint8_t)read_bit_from_bin(F3[p], y, x, alpha, 0.0, med[p], false);
    }
    size_t nh = min(nbins, (size_t)912);
    bool ok = true;
    auto hb = bytes_from_bits(rep3_decode_bits(vector<uint8_t>(raw.begin(), raw.begin() + nh), ok));
    auto pb = bytes_from_bits(rep7_decode_bits(vector<uint8_t>(raw.begin() + nh, raw.end()), ok));
    double t1 = now_s();
    memcpy(out_bytes, hb.data(), hb.size());
    memcpy(out_bytes + hb.size(), pb.data(), pb.size());
    return t1 - t0;
}

}  // extern "C"
