// cli_main.cpp -- `turtlefft embed|extract`: drop-in for the reference CLI (S:813-877, S:907-1312)
// on top of the B200 hot path.  Same sub-commands, flags, defaults (Params S:375-381), messages and
// exit codes for the --pass path; the spectral work is two calls into libtfft_b200.so.
// --key / --wrap-pass (a raw or passphrase-wrapped 32-byte master key, S:576-662, S:1020-1040) are supported;
// not carried over (out of scope): gen-key and the experimental --adaptive_alpha /
// --cover_dependent_path (upstream documents both as broken).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <string>
#include <vector>

#include <unistd.h>

#include "../../../include/tfft.h"
#include "../../../include/tfft_host.h"

namespace {

// Every way out of the process after the GPU context thread has been started: flush and _Exit.  exit() would run the
// CUDA runtime's teardown under that thread (a "cannot open the GPU context" line after the real error message), and
// tearing a context down is time a one-shot process does not need to spend.
[[noreturn]] void leave(int code) {
    fflush(stdout);
    fflush(stderr);
    _Exit(code);
}

struct Args {
    std::string mode, in, out, secret, pass, key, wrap_pass;
    double alpha = 0.50, rmin = 0.05, rmax = 0.45, magmin = 0.01, density = 0.7, jitter = 0.0;  // S:375-381
    bool center = false, adaptive = false, cover_dep = false;
    uint32_t iters = 600000;
};

void usage() {
    fprintf(stderr,
            "Usage:\n"
            "  turtlefft embed   --in cover.png --out stego.png --secret TEXT (--pass PW | --key BASE64 [--wrap-pass PW])\n"
            "            [--alpha 0.5 --jitter 0 --density 0.7 --rmin 0.05 --rmax 0.45 --magmin 0.01 --center 0]\n"
            "            [--pbkdf2_iter 600000]\n"
            "  turtlefft extract --in stego.png (--pass PW | --key BASE64 [--wrap-pass PW]) [same options as embed]\n"
            "  (B200 build: spectral path on the GPU; gen-key and the experimental\n"
            "   --adaptive_alpha / --cover_dependent_path switches are not part of this build)\n");
}

bool parse(int argc, char** argv, Args& A) {
    if (argc < 2) return false;
    A.mode = argv[1];
    for (int i = 2; i < argc; i++) {
        const std::string k = argv[i];
        auto need = [&]() -> std::string { return i + 1 < argc ? std::string(argv[++i]) : std::string(); };
        auto truthy = [](const std::string& v) { return v == "1" || v == "true"; };  // S:863-866
        try {
            if (k == "--in") A.in = need();
            else if (k == "--out") A.out = need();
            else if (k == "--secret") A.secret = need();
            else if (k == "--pass") A.pass = need();
            else if (k == "--key") A.key = need();
            else if (k == "--key-out") need();
            else if (k == "--wrap-pass") A.wrap_pass = need();
            else if (k == "--alpha") A.alpha = std::stod(need());
            else if (k == "--jitter") A.jitter = std::stod(need());
            else if (k == "--density") A.density = std::stod(need());
            else if (k == "--rmin") A.rmin = std::stod(need());
            else if (k == "--rmax") A.rmax = std::stod(need());
            else if (k == "--magmin") A.magmin = std::stod(need());
            else if (k == "--center") A.center = truthy(need());
            else if (k == "--pbkdf2_iter") A.iters = (uint32_t)std::stoul(need());
            else if (k == "--adaptive_alpha") A.adaptive = truthy(need());
            else if (k == "--cover_dependent_path") A.cover_dep = truthy(need());
            else { fprintf(stderr, "Unknown arg: %s\n", k.c_str()); return false; }  // S:867
        } catch (...) { return false; }
    }
    if (A.mode == "gen-key") return true;
    if (A.mode != "embed" && A.mode != "extract") return false;
    if (A.in.empty()) return false;
    if (A.pass.empty() && A.key.empty()) return false;
    if (A.mode == "embed" && (A.out.empty() || A.secret.empty())) return false;
    return true;
}

int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// TFFT_CLI_TIMING=1: wall milliseconds of the phases of one invocation on stderr
struct Phases {
    bool on = getenv("TFFT_CLI_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(), last = t0;
    void mark(const char* what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[timing] %-28s %8.1f ms  (at %8.1f ms)\n", what, std::chrono::duration<double, std::milli>(now - last).count(),
                std::chrono::duration<double, std::milli>(now - t0).count());
        last = now;
    }
};

// A one-shot process pays for the CUDA start-up (driver initialisation, context, module) on every call; it is the largest
// part of a small image's wall time.  So the context is opened on a second thread the moment the arguments are parsed,
// under the PNG decode, the KDF and the walk; on a multi-GPU node only the GPU that will be used is made visible (the
// driver otherwise initialises all of them); and the process leaves through _Exit once its output is flushed instead
// of tearing the context down.
std::future<tfft_ctx*> open_ctx_async() {
    const char* d = getenv("TFFT_DEVICE");
    int dev = d ? atoi(d) : 0;
    if (!getenv("CUDA_VISIBLE_DEVICES")) {  // (the environment is settled before the second thread exists)
        setenv("CUDA_VISIBLE_DEVICES", std::to_string(dev).c_str(), 1);
        dev = 0;
    }
    return std::async(std::launch::async, [dev] {
        tfft_ctx* c = nullptr;
        const int rc = tfft_create(dev, &c);
        if (rc) { fprintf(stderr, "turtlefft: cannot open the GPU context: %s\n", tfft_strerror(rc)); leave(1); }
        return c;
    });
}

[[noreturn]] void die_tfft(tfft_ctx* c, int rc) {
    fprintf(stderr, "turtlefft: %s (%s)\n", tfft_strerror(rc), tfft_last_cuda_error(c));
    leave(1);
}

void random_salt(uint8_t salt[16]) {
    // TFFT_TEST_SALT_HEX (32 hex digits): the test-suite pins the salt to compare stego pixels with the oracle.  Never set
    // it otherwise: a repeated salt repeats the AEAD key and nonce for the same passphrase.
    const char* hex = getenv("TFFT_TEST_SALT_HEX");
    if (hex && strlen(hex) == 32) {
        const std::string h(hex);
        for (int i = 0; i < 16; i++) salt[i] = (uint8_t)strtoul(h.substr(2 * i, 2).c_str(), nullptr, 16);
        return;
    }
    FILE* f = fopen("/dev/urandom", "rb");  // std::random_device upstream (S:927-929)
    if (!f || fread(salt, 1, 16, f) != 16) { fprintf(stderr, "turtlefft: no entropy source\n"); leave(1); }
    fclose(f);
}

// --key: decode_or_unwrap_key (S:603-662); false when the passphrase path is in use
bool load_key(const Args& A, uint8_t master[32]) {
    if (A.key.empty()) return false;
    const int rc = tfft_host_key_decode(A.key.c_str(), A.wrap_pass.c_str(), A.iters, master);
    if (rc < 0) fprintf(stderr, "Key is wrapped but no unwrap passphrase provided\n");
    if (rc != 1) { fprintf(stderr, "Failed to decode/unwrap key from --key argument\n"); leave(1); }  // S:936, S:1149
    return true;
}

void do_embed(const Args& A) {
    Phases ph;
    auto ctx_f = open_ctx_async();
    int W, H;
    uint8_t* img = tfft_host_png_load(A.in.c_str(), &W, &H);
    if (!img) { fprintf(stderr, "Failed to load %s\n", A.in.c_str()); leave(1); }  // S:910
    const int PW = next_pow2(W), PH = next_pow2(H);
    uint8_t salt[16];
    random_salt(salt);
    const size_t nbits = 912 + 56 * (A.secret.size() + 16);
    std::vector<uint8_t> bits(nbits);
    uint8_t hdr[38];
    uint8_t path_key[32], sub[128], master[32];
    if (load_key(A, master)) {  // S:933-939, S:1036
        tfft_host_frame_bits_key(master, salt, (const uint8_t*)A.secret.data(), A.secret.size(), bits.data(), hdr);
        tfft_host_turtle_keys(master, 32, path_key, sub);
        memset(master, 0, sizeof(master));
    } else {
        tfft_host_frame_bits((const uint8_t*)A.pass.data(), A.pass.size(), salt, A.iters, (const uint8_t*)A.secret.data(), A.secret.size(),
                             bits.data(), hdr);
        tfft_host_turtle_keys((const uint8_t*)A.pass.data(), A.pass.size(), path_key, sub);
    }
    std::vector<uint32_t> bins(nbits);
    const int wrc = tfft_host_walk(sub, PH, PW, A.rmin, A.rmax, A.density, nbits, bins.data(), nullptr, nullptr, 0);
    std::vector<double> jit;
    if (A.jitter != 0.0 && wrc == 0) { jit.resize(nbits); tfft_host_jitter(sub, bins.data(), nbits, A.jitter, jit.data()); }
    ph.mark("png + kdf + frame + walk");
    tfft_ctx* ctx = ctx_f.get();
    ph.mark("wait for the gpu context");
    std::vector<uint8_t> out((size_t)W * H * 3);
    uint64_t usable = 0;
    double med[3];
    // a walk that cannot place nbits bins means the message does not fit: run the capacity count only
    const size_t n_embed = wrc == 0 ? nbits : 0;
    int rc = tfft_embed_batch(ctx, img, 1, W, H, bins.data(), bits.data(), n_embed, jit.empty() ? nullptr : jit.data(), A.alpha,
                              A.center ? 1 : 0, A.magmin, A.rmin, A.rmax, out.data(), &usable, med);
    if (rc == TFFT_E_CAPACITY || wrc != 0 || (rc == TFFT_OK && nbits > usable)) {  // S:1009-1012
        fprintf(stderr, "Message too large. Need %zu bits (after ECC), capacity ~%zu bits.\n", nbits, (size_t)usable);
        leave(1);
    }
    if (rc) die_tfft(ctx, rc);
    ph.mark("tfft_embed_batch");
    if (!tfft_host_png_save(A.out.c_str(), out.data(), W, H)) { fprintf(stderr, "PNG write failed: %s\n", A.out.c_str()); leave(1); }  // S:1105
    fprintf(stdout, "Embedded %zu bits into %s (payload %u bytes, ver=2, salt/nonce in header)\n", nbits, A.out.c_str(),
            (unsigned)A.secret.size());  // S:1107
    ph.mark("png encode");
    leave(0);
}

void do_extract(const Args& A) {
    Phases ph;
    auto ctx_f = open_ctx_async();
    int W, H;
    uint8_t* img = tfft_host_png_load(A.in.c_str(), &W, &H);
    if (!img) { fprintf(stderr, "Failed to load %s\n", A.in.c_str()); leave(1); }  // S:1115
    const int PW = next_pow2(W), PH = next_pow2(H);
    uint8_t path_key[32], sub[128], master[32];
    const bool raw_key = load_key(A, master);
    if (raw_key) tfft_host_turtle_keys(master, 32, path_key, sub);
    else tfft_host_turtle_keys((const uint8_t*)A.pass.data(), A.pass.size(), path_key, sub);
    std::vector<uint32_t> bins(912);
    if (tfft_host_walk(sub, PH, PW, A.rmin, A.rmax, A.density, 912, bins.data(), nullptr, nullptr, 0)) {
        fprintf(stderr, "Magic not found.\n"); leave(1);  // not even room for a header
    }
    ph.mark("png + keys + header walk");
    tfft_ctx* ctx = ctx_f.get();
    ph.mark("wait for the gpu context");
    int rc = tfft_forward_batch(ctx, img, 1, W, H, A.center ? 1 : 0);
    if (rc) die_tfft(ctx, rc);
    std::vector<double> jit;
    if (A.jitter != 0.0) { jit.resize(912); tfft_host_jitter(sub, bins.data(), 912, A.jitter, jit.data()); }
    uint8_t hdr[38];
    if ((rc = tfft_read_bits(ctx, bins.data(), 912, 3, jit.empty() ? nullptr : jit.data(), A.alpha, hdr, nullptr))) die_tfft(ctx, rc);
    uint32_t clen = 0;
    const int hrc = tfft_host_parse_header(hdr, &clen, nullptr, nullptr);
    if (hrc == 1) { fprintf(stderr, "Magic not found.\n"); leave(1); }                         // S:1237
    if (hrc == 2) { fprintf(stderr, "Unsupported version (%u).\n", hdr[4]); leave(1); }         // S:1238
    const size_t nb = 912 + 56 * ((size_t)clen + 16);
    // a (noise- or attacker-controlled) length beyond what the annulus can hold cannot be a frame: fail like a walk that
    // ran out of bins instead of allocating for it (the reference walks forever here)
    if (nb > (size_t)3 * PH * PW / 2) { fprintf(stderr, "Payload truncated after ECC decode.\n"); leave(1); }
    bins.resize(nb);
    // the same walk continued (S:1260-1264); bounded, unlike upstream
    if (tfft_host_walk(sub, PH, PW, A.rmin, A.rmax, A.density, nb, bins.data(), nullptr, nullptr, 0)) {
        fprintf(stderr, "Payload truncated after ECC decode.\n"); leave(1);  // S:1269
    }
    if (A.jitter != 0.0) { jit.resize(nb); tfft_host_jitter(sub, bins.data(), nb, A.jitter, jit.data()); }
    std::vector<uint8_t> rest((size_t)clen + 16);
    if ((rc = tfft_read_bits(ctx, bins.data() + 912, nb - 912, 7, jit.empty() ? nullptr : jit.data() + 912, A.alpha, rest.data(), nullptr)))
        die_tfft(ctx, rc);
    ph.mark("forward + header + payload");
    const int opened = raw_key ? tfft_host_open_payload_key(master, hdr, rest.data(), clen)
                               : tfft_host_open_payload((const uint8_t*)A.pass.data(), A.pass.size(), A.iters, hdr, rest.data(), clen);
    memset(master, 0, sizeof(master));
    if (!opened) {
        fprintf(stderr, "Auth failed (wrong pass or data corrupted).\n"); leave(1);  // S:1308
    }
    std::string secret((const char*)rest.data(), clen);
    printf("%s\n", secret.c_str());  // S:1311
    ph.mark("open payload (kdf + aead)");
    leave(0);
}

}  // namespace

int main(int argc, char** argv) {
    Args A;
    if (!parse(argc, argv, A)) { usage(); return 1; }
    if (A.mode == "gen-key") {
        fprintf(stderr, "turtlefft (B200 build): gen-key is outside this build; --key takes keys made by the reference tool\n");
        return 1;
    }
    if (A.adaptive || A.cover_dep) {
        fprintf(stderr, "turtlefft (B200 build): --adaptive_alpha / --cover_dependent_path are experimental upstream and not supported\n");
        return 1;
    }
    if (A.mode == "embed") do_embed(A); else do_extract(A);
    return 0;
}
