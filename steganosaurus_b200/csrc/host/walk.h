#pragma once
#include <cstddef>
#include <cstdint>
namespace tfh {
int walk(const uint8_t key_walk[32], int PH, int PW, double rmin, double rmax, double density, size_t nbits, uint32_t* bins,
         int start[3], uint32_t* ctr_out, uint64_t max_steps);
void jitter_values(const uint8_t sub[128], const uint32_t* bins, size_t nbits, double maxj, double* out);
void turtle_keys(const uint8_t* pass, size_t plen, uint8_t path_key[32], uint8_t sub[128]);
void derive_keys(const uint8_t* pass, size_t plen, const uint8_t salt[16], uint32_t iters, uint8_t aead_key[32], uint8_t nonce[12]);
size_t frame_bits(const uint8_t* pass, size_t plen, const uint8_t salt[16], uint32_t iters, const uint8_t* secret, size_t slen,
                  uint8_t* bits_out, uint8_t header_out[38]);
int parse_header(const uint8_t hdr[38], uint32_t* clen, uint8_t salt[16], uint8_t nonce[12]);
int open_payload(const uint8_t* pass, size_t plen, uint32_t iters, const uint8_t hdr[38], uint8_t* payload, uint32_t clen);
void derive_keys_raw(const uint8_t master[32], const uint8_t salt[16], uint8_t aead_key[32], uint8_t nonce[12]);
size_t frame_bits_key(const uint8_t master[32], const uint8_t salt[16], const uint8_t* secret, size_t slen, uint8_t* bits_out,
                      uint8_t header_out[38]);
int open_payload_key(const uint8_t master[32], const uint8_t hdr[38], uint8_t* payload, uint32_t clen);
int key_decode(const char* key_b64, const char* wrap_pass, uint32_t iters, uint8_t key_out[32]);
}  // namespace tfh
