/* tfft_host.h -- C ABI of the host-side (sequential) half of TurtleFFT, re-implemented from
 * the specification in SURVEY.md App. A.  These functions stay on the CPU by design (north_star):
 * PBKDF2/HKDF, ChaCha20-Poly1305, the SHA-256 keyed turtlewalk, framing, PNG I/O.  They produce the
 * bin-index and bit arrays that include/tfft.h consumes.  S:n = steganosaurus/src/steganosaur.cpp:n.
 */
#ifndef TFFT_HOST_H
#define TFFT_HOST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- primitives (S:46-148, S:151-323) ---------------------------------------------------- */
void tfft_host_sha256(const uint8_t* data, size_t n, uint8_t out[32]);
void tfft_host_hmac_sha256(const uint8_t* key, size_t klen, const uint8_t* msg, size_t mlen, uint8_t out[32]);
void tfft_host_hkdf_expand(const uint8_t prk[32], const uint8_t* info, size_t ilen, uint8_t* out, size_t L);
void tfft_host_pbkdf2(const uint8_t* pass, size_t plen, const uint8_t* salt, size_t slen, uint32_t iters,
                      uint8_t* out, size_t dklen);
/* ChaCha20-Poly1305, in place; tag 16 bytes; open returns 1 when the tag verifies.  NOTE: the tag is
 * the one the REFERENCE computes (S:192-270): its Poly1305 finalisation recombines limbs without
 * truncation and is not RFC 8439-conformant; stego images are only interchangeable with that value.
 * tfft_host_seal_rfc8439 produces the standard tag (known-answer tests). */
void tfft_host_seal(const uint8_t key[32], const uint8_t nonce[12], const uint8_t* aad, size_t alen,
                    uint8_t* data, size_t n, uint8_t tag[16]);
int tfft_host_open(const uint8_t key[32], const uint8_t nonce[12], const uint8_t* aad, size_t alen,
                   uint8_t* data, size_t n, const uint8_t tag[16]);
void tfft_host_seal_rfc8439(const uint8_t key[32], const uint8_t nonce[12], const uint8_t* aad, size_t alen,
                            uint8_t* data, size_t n, uint8_t tag[16]);

/* ---- key schedule (S:556-573, S:1038-1061) ----------------------------------------------- */
/* PBKDF2(pass, salt, iters) -> HKDF("fft_turtle:keys") -> aead_key, nonce */
void tfft_host_derive_keys(const uint8_t* pass, size_t plen, const uint8_t salt[16], uint32_t iters,
                           uint8_t aead_key[32], uint8_t nonce[12]);
/* path_key = SHA256(pass); sub = HKDF-Expand(path_key, "turtle_keys", 128) = key_walk|key_r|key_g|key_b */
void tfft_host_turtle_keys(const uint8_t* pass, size_t plen, uint8_t path_key[32], uint8_t sub[128]);

/* ---- turtlewalk (KS S:665-695, Turtle S:749-810, loop S:1074-1097) ------------------------ */
/* Emits nbits bins (plane<<30 | y*PW + x) for padded dims PH x PW.  Returns 0, or -1 when the walk
 * cannot deliver nbits bins (the reference would spin forever): `max_steps`
 * opcode draws without progress abort the walk (0 = default bound).  start[3] (may be NULL)
 * receives the initial (plane,y,x); ctr_out (may be NULL) the walk keystream block counter. */
int tfft_host_walk(const uint8_t key_walk[32], int PH, int PW, double rmin, double rmax, double density,
                   size_t nbits, uint32_t* bins, int start[3], uint32_t* ctr_out, uint64_t max_steps);
/* per-bin jitter values in walk order (KS::jitter S:690, per-plane keystreams): out[nbits] */
void tfft_host_jitter(const uint8_t sub[128], const uint32_t* bins, size_t nbits, double maxj, double* out);

/* ---- framing (Header S:886-904, S:946-995) ------------------------------------------------ */
/* header(38) = "FTTG" 2 0 salt[16] nonce[12] BE32(clen); bits = Rep3(header) | Rep7(ct|tag).
 * bits_out must hold 912 + 56*(slen+16) entries. Returns nbits. */
size_t tfft_host_frame_bits(const uint8_t* pass, size_t plen, const uint8_t salt[16], uint32_t iters,
                            const uint8_t* secret, size_t slen, uint8_t* bits_out, uint8_t header_out[38]);
/* Header checks of S:1236-1253: 0 ok (clen_out set), 1 "Magic not found.", 2 unsupported version */
int tfft_host_parse_header(const uint8_t hdr[38], uint32_t* clen_out, uint8_t salt_out[16], uint8_t nonce_out[12]);
/* S:1270-1311: payload = ct|tag -> plaintext in place in payload[0..clen); returns 1 on success */
int tfft_host_open_payload(const uint8_t* pass, size_t plen, uint32_t iters, const uint8_t hdr[38],
                           uint8_t* payload, uint32_t clen);

/* ---- --key path (S:576-591, S:603-662, S:1020-1040): a raw 32-byte master key instead of a passphrase ------------
 * path keys: tfft_host_turtle_keys(master_key, 32, ...) (path_key = SHA256(master_key), S:1036);
 * AEAD keys: HKDF-Extract(salt, master_key) -> HKDF-Expand("fft_turtle:keys"), no PBKDF2. */
void tfft_host_derive_keys_raw(const uint8_t master_key[32], const uint8_t salt[16], uint8_t aead_key[32], uint8_t nonce[12]);
size_t tfft_host_frame_bits_key(const uint8_t master_key[32], const uint8_t salt[16], const uint8_t* secret, size_t slen,
                                uint8_t* bits_out, uint8_t header_out[38]);
int tfft_host_open_payload_key(const uint8_t master_key[32], const uint8_t hdr[38], uint8_t* payload, uint32_t clen);
/* decode_or_unwrap_key: base64 of the raw key, or of the passphrase-wrapped 80-byte form ("TFKW", needs wrap_pass).
 * 1 ok, 0 undecodable / wrong passphrase, -1 wrapped key without a passphrase. */
int tfft_host_key_decode(const char* key_b64, const char* wrap_pass, uint32_t iters, uint8_t key_out[32]);

/* ---- PNG (replaces stbi_load(...,3) S:909 / stbi_write_png S:1104) ------------------------- */
/* Decodes any non-interlaced or Adam7 8/16-bit PNG to 8-bit RGB (malloc'd; NULL on any failure, including
 * a dimension above 16384 = TFFT_MAX_DIM).  The caller frees it with tfft_hostlib_free -- NOT with
 * tfft_host_free of include/tfft.h, which releases pinned CUDA memory (the two libraries share no symbol). */
uint8_t* tfft_host_png_load(const char* path, int* W, int* H);
int tfft_host_png_save(const char* path, const uint8_t* rgb, int W, int H);
void tfft_hostlib_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
