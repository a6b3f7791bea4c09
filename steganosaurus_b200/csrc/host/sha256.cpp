// sha256.cpp -- see sha256.h.  Written from the standards; the reference's own copies live at S:46-148.
#include "sha256.h"

#include <vector>

namespace tfh {

static const uint32_t KC[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

static inline uint32_t ror(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

void Sha256::reset() {
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(h, iv, sizeof(h));
    total = 0;
    fill = 0;
}

void Sha256::compress(uint32_t h[8], const uint8_t* p) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        const uint32_t s0 = ror(w[i - 15], 7) ^ ror(w[i - 15], 18) ^ (w[i - 15] >> 3);
        const uint32_t s1 = ror(w[i - 2], 17) ^ ror(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
        const uint32_t S1 = ror(e, 6) ^ ror(e, 11) ^ ror(e, 25);
        const uint32_t t1 = hh + S1 + ((e & f) ^ (~e & g)) + KC[i] + w[i];
        const uint32_t S0 = ror(a, 2) ^ ror(a, 13) ^ ror(a, 22);
        const uint32_t t2 = S0 + ((a & b) ^ (a & c) ^ (b & c));
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

void Sha256::update(const void* data, size_t n) {
    const uint8_t* p = (const uint8_t*)data;
    total += n;
    if (fill) {
        const size_t take = n < 64 - fill ? n : 64 - fill;
        memcpy(buf + fill, p, take);
        fill += take; p += take; n -= take;
        if (fill == 64) { compress(h, buf); fill = 0; }
    }
    while (n >= 64) { compress(h, p); p += 64; n -= 64; }
    if (n) { memcpy(buf, p, n); fill = n; }
}

void Sha256::finish(uint8_t out[32]) {
    const uint64_t bits = total * 8;
    uint8_t pad[72] = {0x80};
    const size_t padlen = (fill < 56) ? 56 - fill : 120 - fill;
    uint8_t len[8];
    for (int i = 0; i < 8; i++) len[i] = (uint8_t)(bits >> (56 - 8 * i));
    update(pad, padlen);
    update(len, 8);
    for (int i = 0; i < 8; i++) { out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i]; }
}

void Hmac::init(const uint8_t* key, size_t klen) {
    uint8_t k0[64] = {0};
    if (klen > 64) sha256(key, klen, k0); else if (klen) memcpy(k0, key, klen);
    uint8_t ip[64], op[64];
    for (int i = 0; i < 64; i++) { ip[i] = k0[i] ^ 0x36; op[i] = k0[i] ^ 0x5c; }
    inner0.reset(); inner0.update(ip, 64);
    outer0.reset(); outer0.update(op, 64);
}
void Hmac::mac(const uint8_t* msg, size_t mlen, uint8_t out[32]) const {
    Sha256 i = inner0;
    i.update(msg, mlen);
    uint8_t ih[32];
    i.finish(ih);
    Sha256 o = outer0;
    o.update(ih, 32);
    o.finish(out);
}
void hmac_sha256(const uint8_t* key, size_t klen, const uint8_t* msg, size_t mlen, uint8_t out[32]) {
    Hmac h;
    h.init(key, klen);
    h.mac(msg, mlen, out);
}

void pbkdf2(const uint8_t* pass, size_t plen, const uint8_t* salt, size_t slen, uint32_t iters, uint8_t* out, size_t dklen) {
    Hmac h;
    h.init(pass, plen);
    const uint32_t blocks = (uint32_t)((dklen + 31) / 32);
    std::vector<uint8_t> msg(slen + 4);
    if (slen) memcpy(msg.data(), salt, slen);
    for (uint32_t i = 1; i <= blocks; i++) {
        msg[slen] = i >> 24; msg[slen + 1] = i >> 16; msg[slen + 2] = i >> 8; msg[slen + 3] = i;
        uint8_t u[32], t[32];
        h.mac(msg.data(), msg.size(), u);
        memcpy(t, u, 32);
        for (uint32_t j = 2; j <= iters; j++) {
            h.mac(u, 32, u);
            for (int k = 0; k < 32; k++) t[k] ^= u[k];
        }
        const size_t off = (size_t)(i - 1) * 32, need = dklen - off < 32 ? dklen - off : 32;
        memcpy(out + off, t, need);
    }
}

void hkdf_expand(const uint8_t prk[32], const uint8_t* info, size_t ilen, uint8_t* out, size_t L) {
    Hmac h;
    h.init(prk, 32);
    uint8_t t[32];
    size_t tlen = 0, pos = 0;
    uint8_t ctr = 1;
    std::vector<uint8_t> msg;
    while (pos < L) {
        msg.assign(t, t + tlen);
        msg.insert(msg.end(), info, info + ilen);
        msg.push_back(ctr++);
        h.mac(msg.data(), msg.size(), t);
        tlen = 32;
        const size_t need = L - pos < 32 ? L - pos : 32;
        memcpy(out + pos, t, need);
        pos += need;
    }
}

}  // namespace tfh
