// walk.cpp -- the keyed turtlewalk and the framing, re-implemented from SURVEY.md App. A
// (reference: KS S:665-695, Turtle S:749-810, embed loop S:1074-1097, Header S:886-904, S:946-995).
// Sequential by nature (a SHA-256 counter-mode keystream drives a data-dependent walk): stays on the host.
#include "walk.h"

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "aead.h"
#include "sha256.h"

namespace tfh {

// ---- KS: block = SHA256(key | 0xAA | LE32(ctr)), 32 bytes per block (S:673-684)
struct Keystream {
    uint8_t key[32];
    uint8_t block[32];
    int pos = 32;
    uint32_t ctr = 0;
    uint32_t pool = 0;
    int bits = 0;
    explicit Keystream(const uint8_t k[32]) { memcpy(key, k, 32); }
    inline uint8_t next_byte() {
        if (pos >= 32) {
            uint8_t m[37];
            memcpy(m, key, 32);
            m[32] = 0xAA;
            m[33] = (uint8_t)ctr; m[34] = (uint8_t)(ctr >> 8); m[35] = (uint8_t)(ctr >> 16); m[36] = (uint8_t)(ctr >> 24);
            sha256(m, 37, block);
            pos = 0;
            ctr++;
        }
        return block[pos++];
    }
    inline int next_opcode3() {  // S:685 (unsigned pool: same low bits as the reference's overflowing int)
        while (bits < 3) { pool = (pool << 8) | next_byte(); bits += 8; }
        const int op = (pool >> (bits - 3)) & 7;
        bits -= 3;
        return op;
    }
    inline double jitter(double maxj) {  // S:690-694: two bytes per call, even when maxj == 0
        const int hi = next_byte(), lo = next_byte();
        const int16_t r = (int16_t)((hi << 8) | lo);
        return (r / 32768.0) * maxj;
    }
};

static inline bool on_axis(int y, int x, int H, int W) {  // S:698-700
    return y == 0 || x == 0 || (H % 2 == 0 && y == H / 2) || (W % 2 == 0 && x == W / 2);
}

int walk(const uint8_t key_walk[32], int PH, int PW, double rmin, double rmax, double density, size_t nbits, uint32_t* bins,
         int start[3], uint32_t* ctr_out, uint64_t max_steps) {
    if (PH <= 0 || PW <= 0 || (size_t)PH * PW >= (1u << 30)) return -2;
    Keystream ks(key_walk);
    // seed (S:764-769): SHA256("seed:" H "x" W "|key:" key_walk)
    std::string seed = "seed:" + std::to_string(PH) + "x" + std::to_string(PW) + "|key:";
    seed.append((const char*)key_walk, 32);
    uint8_t h[32];
    sha256(seed.data(), seed.size(), h);
    uint64_t s = 0;
    for (int i = 0; i < 8; i++) s = (s << 8) | h[i];
    int y = (int)(s % (uint64_t)PH), x = (int)((s >> 16) % (uint64_t)PW), plane = (int)((s >> 32) % 3);
    if (start) { start[0] = plane; start[1] = y; start[2] = x; }
    // visited bitmap: 3 planes x PH x PW bits
    const size_t P = (size_t)PH * PW;
    std::vector<uint64_t> visited((3 * P + 63) / 64, 0);
    auto vis = [&](int p, int yy, int xx) -> bool { const size_t i = (size_t)p * P + (size_t)yy * PW + xx; return (visited[i >> 6] >> (i & 63)) & 1; };
    auto mark = [&](int p, int yy, int xx) { const size_t i = (size_t)p * P + (size_t)yy * PW + xx; visited[i >> 6] |= 1ull << (i & 63); };
    const int m = PH < PW ? PH : PW;
    const double lo = rmin * m, hi = rmax * m;  // annulus_ok S:771-774
    const uint8_t dens = (uint8_t)std::floor(density * 256.0);  // hit_density S:686-689
    if (max_steps == 0) max_steps = 64ull * 3 * P + (1ull << 24);
    uint64_t idle = 0;
    for (size_t i = 0; i < nbits; i++) {
        while (true) {
            // advance_to_valid (S:778-804)
            while (true) {
                if (++idle > max_steps) { if (ctr_out) *ctr_out = ks.ctr; return -1; }
                switch (ks.next_opcode3()) {
                    case 0: plane = (plane + 1) % 3; break;
                    case 1: x = (x + 1) % PW; break;
                    case 2: y = (y + 1) % PH; break;
                    case 3: x = (x - 1 + PW) % PW; break;
                    case 4: y = (y - 1 + PH) % PH; break;
                    case 5: x = (x + 1) % PW; y = (y + 1) % PH; break;
                    case 6: x = (x - 1 + PW) % PW; y = (y + 1) % PH; break;
                    default: break;
                }
                if (on_axis(y, x, PH, PW)) continue;
                if (vis(plane, y, x)) continue;
                const double r = std::hypot((double)y, (double)x);
                if (!(r >= lo && r <= hi)) continue;
                if (vis(plane, (PH - y) % PH, (PW - x) % PW)) continue;
                break;
            }
            const bool hit = ks.next_byte() < dens;
            if (hit) break;
            mark(plane, y, x);  // used-but-empty (S:1080)
            mark(plane, (PH - y) % PH, (PW - x) % PW);
        }
        idle = 0;
        bins[i] = ((uint32_t)plane << 30) | (uint32_t)((size_t)y * PW + x);
        mark(plane, y, x);
        mark(plane, (PH - y) % PH, (PW - x) % PW);
    }
    if (ctr_out) *ctr_out = ks.ctr;
    return 0;
}

void jitter_values(const uint8_t sub[128], const uint32_t* bins, size_t nbits, double maxj, double* out) {
    Keystream kp[3] = {Keystream(sub + 32), Keystream(sub + 64), Keystream(sub + 96)};
    for (size_t i = 0; i < nbits; i++) out[i] = kp[bins[i] >> 30].jitter(maxj);
}

void turtle_keys(const uint8_t* pass, size_t plen, uint8_t path_key[32], uint8_t sub[128]) {
    sha256(pass, plen, path_key);  // S:1038
    static const char info[] = "turtle_keys";
    hkdf_expand(path_key, (const uint8_t*)info, sizeof(info) - 1, sub, 128);  // S:1054-1061
}

void derive_keys(const uint8_t* pass, size_t plen, const uint8_t salt[16], uint32_t iters, uint8_t aead_key[32], uint8_t nonce[12]) {
    uint8_t dk[32], prk[32], out[76];
    pbkdf2(pass, plen, salt, 16, iters, dk, 32);           // S:559
    hmac_sha256(nullptr, 0, dk, 32, prk);                 // HKDF-Extract with an empty salt (S:561)
    static const char info[] = "fft_turtle:keys";
    hkdf_expand(prk, (const uint8_t*)info, sizeof(info) - 1, out, sizeof(out));
    memcpy(aead_key, out + 32, 32);                       // out[0:32] (path_key) is unused upstream (S:565)
    memcpy(nonce, out + 64, 12);
    memset(dk, 0, sizeof(dk)); memset(prk, 0, sizeof(prk)); memset(out, 0, sizeof(out));
}

static void push_bits(std::vector<uint8_t>& bits, const uint8_t* bytes, size_t n, int rep) {
    for (size_t i = 0; i < n; i++)
        for (int b = 7; b >= 0; b--) {  // MSB first (S:455-459)
            const uint8_t v = (bytes[i] >> b) & 1;
            for (int r = 0; r < rep; r++) bits.push_back(v);
        }
}

// --key path (S:576-591): HKDF-Extract(salt, master_key) -> HKDF-Expand("fft_turtle:keys") -> aead_key, nonce
void derive_keys_raw(const uint8_t master[32], const uint8_t salt[16], uint8_t aead_key[32], uint8_t nonce[12]) {
    uint8_t prk[32], out[76];
    hmac_sha256(salt, 16, master, 32, prk);
    static const char info[] = "fft_turtle:keys";
    hkdf_expand(prk, (const uint8_t*)info, sizeof(info) - 1, out, sizeof(out));
    memcpy(aead_key, out + 32, 32);
    memcpy(nonce, out + 64, 12);
    memset(prk, 0, sizeof(prk)); memset(out, 0, sizeof(out));
}

static size_t frame_with(const uint8_t key[32], const uint8_t nonce[12], const uint8_t salt[16], const uint8_t* secret, size_t slen,
                         uint8_t* bits_out, uint8_t header_out[38]) {
    uint8_t hdr[38] = {'F', 'T', 'T', 'G', 2, 0};  // S:886-904
    memcpy(hdr + 6, salt, 16);
    memcpy(hdr + 22, nonce, 12);
    hdr[34] = (uint8_t)(slen >> 24); hdr[35] = (uint8_t)(slen >> 16); hdr[36] = (uint8_t)(slen >> 8); hdr[37] = (uint8_t)slen;
    std::vector<uint8_t> ct(secret, secret + slen);
    ct.resize(slen + 16);
    aead_seal(key, nonce, hdr, 38, ct.data(), slen, ct.data() + slen);  // header is the AAD (S:970)
    std::vector<uint8_t> bits;
    bits.reserve(912 + 56 * (slen + 16));
    push_bits(bits, hdr, 38, 3);             // Rep-3 header (S:987)
    push_bits(bits, ct.data(), ct.size(), 7);  // Rep-7 ct|tag (S:991)
    memcpy(bits_out, bits.data(), bits.size());
    if (header_out) memcpy(header_out, hdr, 38);
    return bits.size();
}

size_t frame_bits(const uint8_t* pass, size_t plen, const uint8_t salt[16], uint32_t iters, const uint8_t* secret, size_t slen,
                  uint8_t* bits_out, uint8_t header_out[38]) {
    uint8_t key[32], nonce[12];
    derive_keys(pass, plen, salt, iters, key, nonce);
    const size_t n = frame_with(key, nonce, salt, secret, slen, bits_out, header_out);
    memset(key, 0, sizeof(key));
    return n;
}

size_t frame_bits_key(const uint8_t master[32], const uint8_t salt[16], const uint8_t* secret, size_t slen, uint8_t* bits_out,
                      uint8_t header_out[38]) {
    uint8_t key[32], nonce[12];
    derive_keys_raw(master, salt, key, nonce);
    const size_t n = frame_with(key, nonce, salt, secret, slen, bits_out, header_out);
    memset(key, 0, sizeof(key));
    return n;
}

int parse_header(const uint8_t hdr[38], uint32_t* clen, uint8_t salt[16], uint8_t nonce[12]) {
    if (!(hdr[0] == 'F' && hdr[1] == 'T' && hdr[2] == 'T' && hdr[3] == 'G')) return 1;  // S:1237
    if (hdr[4] != 2) return 2;                                                          // S:1238
    if (salt) memcpy(salt, hdr + 6, 16);
    if (nonce) memcpy(nonce, hdr + 22, 12);
    if (clen) *clen = (uint32_t)hdr[34] << 24 | (uint32_t)hdr[35] << 16 | (uint32_t)hdr[36] << 8 | hdr[37];
    return 0;
}

int open_payload(const uint8_t* pass, size_t plen, uint32_t iters, const uint8_t hdr[38], uint8_t* payload, uint32_t clen) {
    uint8_t key[32], nonce[12];
    derive_keys(pass, plen, hdr + 6, iters, key, nonce);  // salt from the header (S:1280); nonce re-derived
    const bool ok = aead_open(key, nonce, hdr, 38, payload, clen, payload + clen);  // S:1305
    memset(key, 0, sizeof(key));
    return ok ? 1 : 0;
}

int open_payload_key(const uint8_t master[32], const uint8_t hdr[38], uint8_t* payload, uint32_t clen) {
    uint8_t key[32], nonce[12];
    derive_keys_raw(master, hdr + 6, key, nonce);  // S:1275
    const bool ok = aead_open(key, nonce, hdr, 38, payload, clen, payload + clen);
    memset(key, 0, sizeof(key));
    return ok ? 1 : 0;
}

// decode_or_unwrap_key (S:603-662): base64 of a raw 32-byte key, or of the 80-byte wrapped form
// "TFKW" | salt[16] | nonce[12] | ct[32] | tag[16] (key and nonce = PBKDF2(wrap_pass, salt, iters, 44), no AAD;
// the library AEAD copy the reference wraps with carries the same Poly1305 finalisation quirk as its in-TU copy)
static std::vector<uint8_t> b64_decode(const char* s) {
    std::vector<uint8_t> out;
    uint32_t acc = 0;
    int nb = 0;
    for (; *s; s++) {
        const unsigned char c = (unsigned char)*s;
        int v;
        if (c >= 'A' && c <= 'Z') v = c - 'A';
        else if (c >= 'a' && c <= 'z') v = c - 'a' + 26;
        else if (c >= '0' && c <= '9') v = c - '0' + 52;
        else if (c == '+') v = 62;
        else if (c == '/') v = 63;
        else if (c == '=' || c == '\n' || c == '\r' || c == ' ') continue;
        else return {};
        acc = (acc << 6) | (uint32_t)v;
        nb += 6;
        if (nb >= 8) { nb -= 8; out.push_back((uint8_t)(acc >> nb)); }
    }
    return out;
}
int key_decode(const char* key_b64, const char* wrap_pass, uint32_t iters, uint8_t key_out[32]) {
    const std::vector<uint8_t> d = b64_decode(key_b64);
    if (d.size() == 80 && !memcmp(d.data(), "TFKW", 4)) {
        if (!wrap_pass || !*wrap_pass) return -1;  // "Key is wrapped but no unwrap passphrase provided"
        uint8_t derived[44];
        pbkdf2((const uint8_t*)wrap_pass, strlen(wrap_pass), d.data() + 4, 16, iters, derived, sizeof(derived));
        uint8_t ct[32];
        memcpy(ct, d.data() + 32, 32);
        const bool ok = aead_open(derived, d.data() + 20, nullptr, 0, ct, 32, d.data() + 64);
        if (ok) memcpy(key_out, ct, 32);
        memset(derived, 0, sizeof(derived)); memset(ct, 0, sizeof(ct));
        return ok ? 1 : 0;
    }
    if (d.size() == 32) { memcpy(key_out, d.data(), 32); return 1; }
    return 0;
}

}  // namespace tfh
