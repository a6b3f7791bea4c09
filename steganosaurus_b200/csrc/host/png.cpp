// png.cpp -- minimal PNG codec on top of zlib: decode any 1/2/4/8/16-bit gray/RGB/palette/alpha PNG
// (interlaced or not) to 8-bit RGB exactly like stbi_load(path,&W,&H,&c,3) does (S:909: alpha dropped,
// gray replicated, 16-bit reduced by >> 8, sub-byte gray scaled to 0..255); encode 8-bit RGB (S:1104).
// Pixels, not file bytes, are the contract: any valid lossless PNG is equivalent.
#include "png.h"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace tfh {

static inline uint32_t be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
static inline int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// reverse the scanline filters of one (sub)image in place; rows are [filter byte][rowbytes]
static bool unfilter(uint8_t* d, int rows, size_t rowbytes, int bpp) {
    std::vector<uint8_t> zero(rowbytes, 0);
    const uint8_t* prev = zero.data();
    for (int y = 0; y < rows; y++) {
        uint8_t* r = d + (size_t)y * (rowbytes + 1);
        const int f = r[0];
        uint8_t* cur = r + 1;
        switch (f) {
            case 0: break;
            case 1: for (size_t i = bpp; i < rowbytes; i++) cur[i] += cur[i - bpp]; break;
            case 2: for (size_t i = 0; i < rowbytes; i++) cur[i] += prev[i]; break;
            case 3:
                for (size_t i = 0; i < rowbytes; i++) cur[i] += (uint8_t)(((i >= (size_t)bpp ? cur[i - bpp] : 0) + prev[i]) >> 1);
                break;
            case 4:
                for (size_t i = 0; i < rowbytes; i++)
                    cur[i] += (uint8_t)paeth(i >= (size_t)bpp ? cur[i - bpp] : 0, prev[i], i >= (size_t)bpp ? prev[i - bpp] : 0);
                break;
            default: return false;
        }
        prev = cur;
    }
    return true;
}

struct Info { int W, H, depth, ctype, interlace, channels; };

// one decoded sample row (after unfiltering) -> RGB8 pixels written with stride `step` starting at x0
static void row_to_rgb(const Info& I, const uint8_t* row, int npix, const uint8_t* pal, uint8_t* out_row, int x0, int step) {
    const int d = I.depth;
    auto sample = [&](int idx) -> int {  // idx-th sample of the row, reduced to 8 bits where needed
        if (d == 8) return row[idx];
        if (d == 16) return row[2 * idx];  // high byte (stb: >> 8)
        const int per = 8 / d, b = row[idx / per], sh = (per - 1 - idx % per) * d;
        return (b >> sh) & ((1 << d) - 1);
    };
    static const int scale[5] = {0, 255, 85, 0, 17};
    for (int i = 0; i < npix; i++) {
        uint8_t* o = out_row + (size_t)(x0 + i * step) * 3;
        switch (I.ctype) {
            case 0: { int g = sample(i); if (d < 8) g *= scale[d]; o[0] = o[1] = o[2] = (uint8_t)g; break; }
            case 4: { const int g = sample(2 * i); o[0] = o[1] = o[2] = (uint8_t)g; break; }
            case 2: o[0] = sample(3 * i); o[1] = sample(3 * i + 1); o[2] = sample(3 * i + 2); break;
            case 6: o[0] = sample(4 * i); o[1] = sample(4 * i + 1); o[2] = sample(4 * i + 2); break;
            case 3: { const int p = sample(i); o[0] = pal[3 * p]; o[1] = pal[3 * p + 1]; o[2] = pal[3 * p + 2]; break; }
        }
    }
}

uint8_t* png_load(const char* path, int* Wout, int* Hout) {
    FILE* f = fopen(path, "rb");
    if (!f) return nullptr;
    std::vector<uint8_t> file;
    uint8_t tmp[1 << 16];
    size_t n;
    while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) file.insert(file.end(), tmp, tmp + n);
    fclose(f);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 + 25 || memcmp(file.data(), sig, 8)) return nullptr;
    Info I{};
    uint8_t pal[768] = {0};
    std::vector<uint8_t> idat;
    bool have_ihdr = false;
    for (size_t pos = 8; pos + 12 <= file.size();) {
        const uint32_t len = be32(&file[pos]);
        const uint8_t* type = &file[pos + 4];
        const uint8_t* data = &file[pos + 8];
        if (pos + 12 + (size_t)len > file.size()) return nullptr;
        if (!memcmp(type, "IHDR", 4) && len >= 13) {
            I.W = (int)be32(data); I.H = (int)be32(data + 4); I.depth = data[8]; I.ctype = data[9]; I.interlace = data[12];
            static const int ch[7] = {1, 0, 3, 1, 2, 0, 4};
            if (have_ihdr || I.ctype > 6 || !ch[I.ctype] || I.W <= 0 || I.H <= 0 || data[10] || data[11] || I.interlace > 1) return nullptr;
            if (I.W > 16384 || I.H > 16384) return nullptr;  // TFFT_MAX_DIM: nothing larger can be processed (stb caps at 1 << 24)
            if (!(I.depth == 1 || I.depth == 2 || I.depth == 4 || I.depth == 8 || I.depth == 16)) return nullptr;
            I.channels = ch[I.ctype];
            have_ihdr = true;
        } else if (!have_ihdr) {
            return nullptr;  // IHDR must come first
        } else if (!memcmp(type, "PLTE", 4)) {
            memcpy(pal, data, len < 768 ? len : 768);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || idat.empty()) return nullptr;
    const int bits_pp = I.depth * I.channels, bpp = bits_pp >= 8 ? bits_pp / 8 : 1;
    auto rowbytes_of = [&](int w) -> size_t { return ((size_t)w * bits_pp + 7) / 8; };
    // sub-images: one for non-interlaced, seven Adam7 passes otherwise
    static const int xs[7] = {0, 4, 0, 2, 0, 1, 0}, ys[7] = {0, 0, 4, 0, 2, 0, 1}, dx[7] = {8, 8, 4, 4, 2, 2, 1}, dy[7] = {8, 8, 8, 4, 4, 2, 2};
    struct Pass { int w, h, x0, y0, sx, sy; };
    std::vector<Pass> passes;
    if (!I.interlace) passes.push_back({I.W, I.H, 0, 0, 1, 1});
    else
        for (int p = 0; p < 7; p++) {
            const int w = (I.W - xs[p] + dx[p] - 1) / dx[p], h = (I.H - ys[p] + dy[p] - 1) / dy[p];
            if (w > 0 && h > 0) passes.push_back({w, h, xs[p], ys[p], dx[p], dy[p]});
        }
    size_t raw_size = 0;
    for (auto& p : passes) raw_size += (size_t)p.h * (rowbytes_of(p.w) + 1);
    std::vector<uint8_t> raw(raw_size);
    z_stream zs{};
    if (inflateInit(&zs) != Z_OK) return nullptr;
    size_t produced = 0, fed = 0;
    int zr = Z_OK;
    while (zr == Z_OK && produced < raw_size) {  // sizes may exceed uInt: feed input and output in pieces
        if (zs.avail_in == 0 && fed < idat.size()) {
            const size_t take = idat.size() - fed < (1u << 30) ? idat.size() - fed : (1u << 30);
            zs.next_in = idat.data() + fed; zs.avail_in = (uInt)take; fed += take;
        }
        const size_t want = raw_size - produced < (1u << 30) ? raw_size - produced : (1u << 30);
        zs.next_out = raw.data() + produced; zs.avail_out = (uInt)want;
        zr = inflate(&zs, Z_NO_FLUSH);
        produced += want - zs.avail_out;
        if (zs.avail_in == 0 && fed == idat.size() && zr == Z_OK && zs.avail_out != 0) break;
    }
    inflateEnd(&zs);
    if (produced != raw_size) return nullptr;
    uint8_t* out = (uint8_t*)malloc((size_t)I.W * I.H * 3);
    if (!out) return nullptr;
    size_t off = 0;
    for (auto& p : passes) {
        const size_t rb = rowbytes_of(p.w);
        if (!unfilter(raw.data() + off, p.h, rb, bpp)) { free(out); return nullptr; }
        for (int y = 0; y < p.h; y++)
            row_to_rgb(I, raw.data() + off + (size_t)y * (rb + 1) + 1, p.w, pal, out + (size_t)(p.y0 + y * p.sy) * I.W * 3, p.x0, p.sx);
        off += (size_t)p.h * (rb + 1);
    }
    *Wout = I.W; *Hout = I.H;
    return out;
}

static void put_chunk(FILE* f, const char* type, const uint8_t* data, size_t len) {
    uint8_t hdr[8] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len, (uint8_t)type[0], (uint8_t)type[1], (uint8_t)type[2], (uint8_t)type[3]};
    fwrite(hdr, 1, 8, f);
    if (len) fwrite(data, 1, len, f);
    uLong c = crc32(0L, hdr + 4, 4);
    if (len) c = crc32(c, data, (uInt)len);
    const uint8_t crc[4] = {(uint8_t)(c >> 24), (uint8_t)(c >> 16), (uint8_t)(c >> 8), (uint8_t)c};
    fwrite(crc, 1, 4, f);
}

int png_save(const char* path, const uint8_t* rgb, int W, int H) {
    if (W <= 0 || H <= 0) return 0;
    const size_t rb = (size_t)W * 3;
    std::vector<uint8_t> raw((size_t)H * (rb + 1));
    std::vector<uint8_t> cand(rb);
    const std::vector<uint8_t> zero(rb, 0);
    for (int y = 0; y < H; y++) {
        const uint8_t* cur = rgb + (size_t)y * rb;
        const uint8_t* prev = y ? cur - rb : zero.data();
        int best_f = 0;
        unsigned long best = ~0ul;
        uint8_t* dst = raw.data() + (size_t)y * (rb + 1);
        for (int f = 0; f < 5; f++) {  // minimum-sum-of-absolute-differences heuristic
            unsigned long sum = 0;
            for (size_t i = 0; i < rb; i++) {
                const int a = i >= 3 ? cur[i - 3] : 0, b = prev[i], c = i >= 3 ? prev[i - 3] : 0;
                int v;
                switch (f) {
                    case 0: v = cur[i]; break;
                    case 1: v = cur[i] - a; break;
                    case 2: v = cur[i] - b; break;
                    case 3: v = cur[i] - ((a + b) >> 1); break;
                    default: v = cur[i] - paeth(a, b, c); break;
                }
                cand[i] = (uint8_t)v;
                sum += (unsigned long)abs((int)(int8_t)cand[i]);
            }
            if (sum < best) { best = sum; best_f = f; dst[0] = (uint8_t)f; memcpy(dst + 1, cand.data(), rb); }
        }
        (void)best_f;
    }
    // deflate in one stream (sizes may exceed uInt: loop)
    z_stream zs{};
    // TFFT_PNG_LEVEL (0..9, default 6): any valid lossless PNG is fine for the pipeline, level 1 is ~3x faster
    static const int level = [] { const char* e = getenv("TFFT_PNG_LEVEL"); int v = e ? atoi(e) : 6; return v < 0 ? 0 : (v > 9 ? 9 : v); }();
    if (deflateInit(&zs, level) != Z_OK) return 0;
    std::vector<uint8_t> comp;
    comp.resize(raw.size() / 2 + 4096);
    size_t in_pos = 0, out_pos = 0;
    int zr = Z_OK;
    while (zr != Z_STREAM_END) {
        if (zs.avail_in == 0 && in_pos < raw.size()) {
            const size_t take = raw.size() - in_pos < (1u << 30) ? raw.size() - in_pos : (1u << 30);
            zs.next_in = raw.data() + in_pos; zs.avail_in = (uInt)take; in_pos += take;
        }
        if (out_pos == comp.size()) comp.resize(comp.size() * 2);
        const size_t room = comp.size() - out_pos < (1u << 30) ? comp.size() - out_pos : (1u << 30);
        zs.next_out = comp.data() + out_pos; zs.avail_out = (uInt)room;
        zr = deflate(&zs, in_pos == raw.size() ? Z_FINISH : Z_NO_FLUSH);
        out_pos += room - zs.avail_out;
        if (zr != Z_OK && zr != Z_STREAM_END && zr != Z_BUF_ERROR) { deflateEnd(&zs); return 0; }
    }
    deflateEnd(&zs);
    FILE* f = fopen(path, "wb");
    if (!f) return 0;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    fwrite(sig, 1, 8, f);
    const uint8_t ihdr[13] = {(uint8_t)(W >> 24), (uint8_t)(W >> 16), (uint8_t)(W >> 8), (uint8_t)W, (uint8_t)(H >> 24), (uint8_t)(H >> 16), (uint8_t)(H >> 8), (uint8_t)H, 8, 2, 0, 0, 0};
    put_chunk(f, "IHDR", ihdr, 13);
    for (size_t p = 0; p < out_pos; p += (1u << 20)) put_chunk(f, "IDAT", comp.data() + p, out_pos - p < (1u << 20) ? out_pos - p : (1u << 20));
    put_chunk(f, "IEND", nullptr, 0);
    const bool ok = !ferror(f);
    fclose(f);
    return ok ? 1 : 0;
}

}  // namespace tfh
