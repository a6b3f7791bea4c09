// host_capi.cpp -- extern "C" face of the host-side pieces (include/tfft_host.h).
#include "../../../include/tfft_host.h"

#include <cstdlib>

#include "aead.h"
#include "png.h"
#include "sha256.h"
#include "walk.h"

extern "C" {

void tfft_host_sha256(const uint8_t* d, size_t n, uint8_t out[32]) { tfh::sha256(d, n, out); }
void tfft_host_hmac_sha256(const uint8_t* k, size_t kl, const uint8_t* m, size_t ml, uint8_t out[32]) { tfh::hmac_sha256(k, kl, m, ml, out); }
void tfft_host_hkdf_expand(const uint8_t prk[32], const uint8_t* info, size_t il, uint8_t* out, size_t L) { tfh::hkdf_expand(prk, info, il, out, L); }
void tfft_host_pbkdf2(const uint8_t* p, size_t pl, const uint8_t* s, size_t sl, uint32_t it, uint8_t* out, size_t dk) { tfh::pbkdf2(p, pl, s, sl, it, out, dk); }
void tfft_host_seal(const uint8_t key[32], const uint8_t nonce[12], const uint8_t* aad, size_t al, uint8_t* d, size_t n, uint8_t tag[16]) {
    tfh::aead_seal(key, nonce, aad, al, d, n, tag);
}
int tfft_host_open(const uint8_t key[32], const uint8_t nonce[12], const uint8_t* aad, size_t al, uint8_t* d, size_t n, const uint8_t tag[16]) {
    return tfh::aead_open(key, nonce, aad, al, d, n, tag) ? 1 : 0;
}
void tfft_host_seal_rfc8439(const uint8_t key[32], const uint8_t nonce[12], const uint8_t* aad, size_t al, uint8_t* d, size_t n, uint8_t tag[16]) {
    tfh::aead_seal(key, nonce, aad, al, d, n, tag, true);
}
void tfft_host_derive_keys(const uint8_t* p, size_t pl, const uint8_t salt[16], uint32_t it, uint8_t key[32], uint8_t nonce[12]) {
    tfh::derive_keys(p, pl, salt, it, key, nonce);
}
void tfft_host_turtle_keys(const uint8_t* p, size_t pl, uint8_t path_key[32], uint8_t sub[128]) { tfh::turtle_keys(p, pl, path_key, sub); }
int tfft_host_walk(const uint8_t kw[32], int PH, int PW, double rmin, double rmax, double density, size_t nbits, uint32_t* bins,
                   int start[3], uint32_t* ctr_out, uint64_t max_steps) {
    try { return tfh::walk(kw, PH, PW, rmin, rmax, density, nbits, bins, start, ctr_out, max_steps); } catch (...) { return -1; }
}
void tfft_host_jitter(const uint8_t sub[128], const uint32_t* bins, size_t nbits, double maxj, double* out) { tfh::jitter_values(sub, bins, nbits, maxj, out); }
size_t tfft_host_frame_bits(const uint8_t* p, size_t pl, const uint8_t salt[16], uint32_t it, const uint8_t* secret, size_t sl,
                            uint8_t* bits_out, uint8_t header_out[38]) {
    return tfh::frame_bits(p, pl, salt, it, secret, sl, bits_out, header_out);
}
int tfft_host_parse_header(const uint8_t hdr[38], uint32_t* clen, uint8_t salt[16], uint8_t nonce[12]) { return tfh::parse_header(hdr, clen, salt, nonce); }
int tfft_host_open_payload(const uint8_t* p, size_t pl, uint32_t it, const uint8_t hdr[38], uint8_t* payload, uint32_t clen) {
    return tfh::open_payload(p, pl, it, hdr, payload, clen);
}
void tfft_host_derive_keys_raw(const uint8_t master[32], const uint8_t salt[16], uint8_t key[32], uint8_t nonce[12]) {
    tfh::derive_keys_raw(master, salt, key, nonce);
}
size_t tfft_host_frame_bits_key(const uint8_t master[32], const uint8_t salt[16], const uint8_t* secret, size_t sl, uint8_t* bits_out,
                                uint8_t header_out[38]) {
    return tfh::frame_bits_key(master, salt, secret, sl, bits_out, header_out);
}
int tfft_host_open_payload_key(const uint8_t master[32], const uint8_t hdr[38], uint8_t* payload, uint32_t clen) {
    return tfh::open_payload_key(master, hdr, payload, clen);
}
int tfft_host_key_decode(const char* key_b64, const char* wrap_pass, uint32_t iters, uint8_t key_out[32]) {
    try { return tfh::key_decode(key_b64, wrap_pass, iters, key_out); } catch (...) { return 0; }
}
// (the PNG codec allocates: nothing may unwind through the C boundary)
uint8_t* tfft_host_png_load(const char* path, int* W, int* H) {
    try { return tfh::png_load(path, W, H); } catch (...) { return nullptr; }
}
int tfft_host_png_save(const char* path, const uint8_t* rgb, int W, int H) {
    try { return tfh::png_save(path, rgb, W, H); } catch (...) { return 0; }
}
void tfft_hostlib_free(void* p) { free(p); }

}  // extern "C"
