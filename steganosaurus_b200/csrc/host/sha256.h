// sha256.h -- SHA-256 / HMAC / PBKDF2 / HKDF for the host side (FIPS 180-4, RFC 2104, 8018, 5869).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace tfh {

struct Sha256 {
    uint32_t h[8];
    uint8_t buf[64];
    uint64_t total = 0;
    size_t fill = 0;
    Sha256() { reset(); }
    void reset();
    void update(const void* data, size_t n);
    void finish(uint8_t out[32]);
    static void compress(uint32_t h[8], const uint8_t block[64]);
};

inline void sha256(const void* d, size_t n, uint8_t out[32]) { Sha256 s; s.update(d, n); s.finish(out); }

// HMAC with precomputed inner/outer midstates (PBKDF2 reuses them for every iteration)
struct Hmac {
    Sha256 inner0, outer0;
    void init(const uint8_t* key, size_t klen);
    void mac(const uint8_t* msg, size_t mlen, uint8_t out[32]) const;
};

void hmac_sha256(const uint8_t* key, size_t klen, const uint8_t* msg, size_t mlen, uint8_t out[32]);
void pbkdf2(const uint8_t* pass, size_t plen, const uint8_t* salt, size_t slen, uint32_t iters, uint8_t* out, size_t dklen);
void hkdf_expand(const uint8_t prk[32], const uint8_t* info, size_t ilen, uint8_t* out, size_t L);

}  // namespace tfh
