// aead.cpp -- ChaCha20-Poly1305 AEAD (RFC 8439) for the host side.  The reference keeps two copies
// (S:151-323 and crypto/chacha20poly1305.cpp); this one is written from the RFC.
#include "aead.h"

#include <cstring>
#include <vector>

namespace tfh {

static inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static inline void wr32(uint8_t* p, uint32_t v) { p[0] = v; p[1] = v >> 8; p[2] = v >> 16; p[3] = v >> 24; }
static inline uint32_t rol(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }

#define TFH_QR(a, b, c, d) \
    a += b; d ^= a; d = rol(d, 16); c += d; b ^= c; b = rol(b, 12); a += b; d ^= a; d = rol(d, 8); c += d; b ^= c; b = rol(b, 7);

static void chacha_block(const uint8_t key[32], const uint8_t nonce[12], uint32_t counter, uint8_t out[64]) {
    uint32_t s[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574};
    for (int i = 0; i < 8; i++) s[4 + i] = rd32(key + 4 * i);
    s[12] = counter;
    for (int i = 0; i < 3; i++) s[13 + i] = rd32(nonce + 4 * i);
    uint32_t x[16];
    memcpy(x, s, sizeof(x));
    for (int r = 0; r < 10; r++) {
        TFH_QR(x[0], x[4], x[8], x[12]) TFH_QR(x[1], x[5], x[9], x[13]) TFH_QR(x[2], x[6], x[10], x[14]) TFH_QR(x[3], x[7], x[11], x[15])
        TFH_QR(x[0], x[5], x[10], x[15]) TFH_QR(x[1], x[6], x[11], x[12]) TFH_QR(x[2], x[7], x[8], x[13]) TFH_QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; i++) wr32(out + 4 * i, x[i] + s[i]);
}

static void chacha_xor(const uint8_t key[32], const uint8_t nonce[12], uint32_t counter, uint8_t* data, size_t n) {
    uint8_t ks[64];
    for (size_t off = 0; off < n; off += 64) {
        chacha_block(key, nonce, counter++, ks);
        const size_t m = n - off < 64 ? n - off : 64;
        for (size_t i = 0; i < m; i++) data[off + i] ^= ks[i];
    }
}

// Poly1305 over 16-byte blocks, radix 2^26
struct Poly {
    uint32_t r[5], h[5] = {0, 0, 0, 0, 0}, pad[4];
    explicit Poly(const uint8_t key[32]) {
        r[0] = rd32(key) & 0x3ffffff;
        r[1] = (rd32(key + 3) >> 2) & 0x3ffff03;
        r[2] = (rd32(key + 6) >> 4) & 0x3ffc0ff;
        r[3] = (rd32(key + 9) >> 6) & 0x3f03fff;
        r[4] = (rd32(key + 12) >> 8) & 0x00fffff;
        for (int i = 0; i < 4; i++) pad[i] = rd32(key + 16 + 4 * i);
    }
    void block(const uint8_t* m, uint32_t hibit) {
        const uint64_t s1 = r[1] * 5, s2 = r[2] * 5, s3 = r[3] * 5, s4 = r[4] * 5;
        h[0] += rd32(m) & 0x3ffffff;
        h[1] += (rd32(m + 3) >> 2) & 0x3ffffff;
        h[2] += (rd32(m + 6) >> 4) & 0x3ffffff;
        h[3] += (rd32(m + 9) >> 6) & 0x3ffffff;
        h[4] += (rd32(m + 12) >> 8) | hibit;
        const uint64_t d0 = (uint64_t)h[0] * r[0] + h[1] * s4 + h[2] * s3 + h[3] * s2 + h[4] * s1;
        uint64_t d1 = (uint64_t)h[0] * r[1] + (uint64_t)h[1] * r[0] + h[2] * s4 + h[3] * s3 + h[4] * s2;
        uint64_t d2 = (uint64_t)h[0] * r[2] + (uint64_t)h[1] * r[1] + (uint64_t)h[2] * r[0] + h[3] * s4 + h[4] * s3;
        uint64_t d3 = (uint64_t)h[0] * r[3] + (uint64_t)h[1] * r[2] + (uint64_t)h[2] * r[1] + (uint64_t)h[3] * r[0] + h[4] * s4;
        uint64_t d4 = (uint64_t)h[0] * r[4] + (uint64_t)h[1] * r[3] + (uint64_t)h[2] * r[2] + (uint64_t)h[3] * r[1] + (uint64_t)h[4] * r[0];
        uint64_t c = d0 >> 26; h[0] = d0 & 0x3ffffff;
        d1 += c; c = d1 >> 26; h[1] = d1 & 0x3ffffff;
        d2 += c; c = d2 >> 26; h[2] = d2 & 0x3ffffff;
        d3 += c; c = d3 >> 26; h[3] = d3 & 0x3ffffff;
        d4 += c; c = d4 >> 26; h[4] = d4 & 0x3ffffff;
        h[0] += (uint32_t)c * 5; h[1] += h[0] >> 26; h[0] &= 0x3ffffff;
    }
    // ref_quirk: the reference's in-TU Poly1305 (S:261-264) recombines the 26-bit limbs in 64-bit
    // arithmetic WITHOUT truncating (h1<<26) etc. to 32 bits, so the bits that spill past each 32-bit
    // word are added a second time through the carry.  Its tags are therefore not RFC 8439 tags.
    // Cross-tool compatibility needs exactly that value; rfc = true gives the standard tag.
    void finish(uint8_t tag[16], bool rfc) {
        uint32_t c = h[1] >> 26; h[1] &= 0x3ffffff;
        h[2] += c; c = h[2] >> 26; h[2] &= 0x3ffffff;
        h[3] += c; c = h[3] >> 26; h[3] &= 0x3ffffff;
        h[4] += c; c = h[4] >> 26; h[4] &= 0x3ffffff;
        h[0] += c * 5; c = h[0] >> 26; h[0] &= 0x3ffffff;
        h[1] += c;
        uint32_t g[5];
        g[0] = h[0] + 5; c = g[0] >> 26; g[0] &= 0x3ffffff;
        for (int i = 1; i < 4; i++) { g[i] = h[i] + c; c = g[i] >> 26; g[i] &= 0x3ffffff; }
        g[4] = h[4] + c - (1u << 26);
        const uint32_t take_g = (g[4] >> 31) - 1;  // all ones when h >= p (no borrow)
        for (int i = 0; i < 5; i++) h[i] = (h[i] & ~take_g) | (g[i] & take_g);
        h[4] &= 0x3ffffff;
        if (rfc) {
            const uint32_t w0 = h[0] | (h[1] << 26), w1 = (h[1] >> 6) | (h[2] << 20), w2 = (h[2] >> 12) | (h[3] << 14), w3 = (h[3] >> 18) | (h[4] << 8);
            uint64_t f = (uint64_t)w0 + pad[0]; wr32(tag, (uint32_t)f);
            f = (uint64_t)w1 + pad[1] + (f >> 32); wr32(tag + 4, (uint32_t)f);
            f = (uint64_t)w2 + pad[2] + (f >> 32); wr32(tag + 8, (uint32_t)f);
            f = (uint64_t)w3 + pad[3] + (f >> 32); wr32(tag + 12, (uint32_t)f);
        } else {
            const uint64_t H0 = h[0], H1 = h[1], H2 = h[2], H3 = h[3], H4 = (uint64_t)h[4] + (1ull << 26);  // S:253
            uint64_t f0 = (H0 | (H1 << 26)) + pad[0];
            uint64_t f1 = ((H1 >> 6) | (H2 << 20)) + pad[1] + (f0 >> 32);
            uint64_t f2 = ((H2 >> 12) | (H3 << 14)) + pad[2] + (f1 >> 32);
            uint64_t f3 = ((H3 >> 18) | (H4 << 8)) + pad[3] + (f2 >> 32);
            wr32(tag, (uint32_t)f0); wr32(tag + 4, (uint32_t)f1); wr32(tag + 8, (uint32_t)f2); wr32(tag + 12, (uint32_t)f3);
        }
    }
};

static void aead_tag(const uint8_t key[32], const uint8_t nonce[12], const uint8_t* aad, size_t alen, const uint8_t* ct, size_t n,
                     uint8_t tag[16], bool rfc) {
    uint8_t otk[64];
    chacha_block(key, nonce, 0, otk);
    Poly p(otk);
    auto absorb = [&](const uint8_t* d, size_t len) {  // data then zero padding to 16
        size_t off = 0;
        for (; off + 16 <= len; off += 16) p.block(d + off, 1u << 24);
        if (off < len) {
            uint8_t b[16] = {0};
            memcpy(b, d + off, len - off);
            p.block(b, 1u << 24);
        }
    };
    if (aad && alen) absorb(aad, alen);
    if (n) absorb(ct, n);
    uint8_t lens[16];
    for (int i = 0; i < 8; i++) { lens[i] = (uint8_t)((uint64_t)alen >> (8 * i)); lens[8 + i] = (uint8_t)((uint64_t)n >> (8 * i)); }
    p.block(lens, 1u << 24);
    p.finish(tag, rfc);
    memset(otk, 0, sizeof(otk));
}

void aead_seal(const uint8_t key[32], const uint8_t nonce[12], const uint8_t* aad, size_t alen, uint8_t* data, size_t n, uint8_t tag[16], bool rfc) {
    chacha_xor(key, nonce, 1, data, n);
    aead_tag(key, nonce, aad, alen, data, n, tag, rfc);
}

bool aead_open(const uint8_t key[32], const uint8_t nonce[12], const uint8_t* aad, size_t alen, uint8_t* data, size_t n, const uint8_t tag[16], bool rfc) {
    uint8_t t[16];
    aead_tag(key, nonce, aad, alen, data, n, t, rfc);
    uint8_t diff = 0;
    for (int i = 0; i < 16; i++) diff |= t[i] ^ tag[i];
    if (diff) return false;
    chacha_xor(key, nonce, 1, data, n);
    return true;
}

}  // namespace tfh
