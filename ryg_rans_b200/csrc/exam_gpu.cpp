// exam_gpu.cpp -- the reference's demo driver, with its hot loops replaced by the bulk C-ABI.
//
// Same shape as main_simd.cpp / main_alias.cpp (read a file or synthesize a buffer, build
// the order-0 model on the host, encode x5, decode x5, memcmp, print sizes and "decode ok!"),
// but the N-way encode loop (main_simd.cpp:287-300) and the SIMD decode loop (:313-332) are
// one rb200_encode / rb200_decode call each.  Host code stays C++, as in the reference.
//
//   exam_gpu [file|-] [word|alias] [chunk_syms] [synthetic_bytes]
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rans_b200.h"

static double now_s()
{
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

static void die(const char* what, int rc, rb200_ctx* ctx)
{
    std::fprintf(stderr, "Error: %s: %s %s\n", what, rb200_strerror(rc), ctx ? rb200_last_cuda_error(ctx) : "");
    std::exit(1);
}

int main(int argc, char** argv)
{
    const std::string path = argc > 1 ? argv[1] : "-";
    const bool alias = argc > 2 && std::string(argv[2]) == "alias";
    const uint32_t chunk = argc > 3 ? static_cast<uint32_t>(std::atoi(argv[3])) : 8192;
    const size_t synth = argc > 4 ? static_cast<size_t>(std::atoll(argv[4])) : (64u << 20);

    std::vector<uint8_t> in_bytes;
    if (path == "-") {                       // seeded skewed bytes (splitmix64), no file needed
        in_bytes.resize(synth);
        uint64_t s = 0x5EED0000ull;
        for (size_t i = 0; i < synth; i++) {
            s += 0x9E3779B97F4A7C15ull;
            uint64_t z = s;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            z ^= z >> 31;
            in_bytes[i] = static_cast<uint8_t>((z & 0xff) & ((z >> 8) & 0xff));   // AND of two uniform bytes: skewed
        }
    } else {
        FILE* f = std::fopen(path.c_str(), "rb");
        if (!f) { std::fprintf(stderr, "Error: file not found: %s\n", path.c_str()); return 1; }
        std::fseek(f, 0, SEEK_END);
        in_bytes.resize(static_cast<size_t>(std::ftell(f)));
        std::fseek(f, 0, SEEK_SET);
        if (std::fread(in_bytes.data(), 1, in_bytes.size(), f) != in_bytes.size()) { std::fprintf(stderr, "Error: read failed\n"); return 1; }
        std::fclose(f);
    }
    const size_t in_size = in_bytes.size();

    // ---- model, on the host like the reference (SymbolStats)
    const uint32_t scale_bits = alias ? 16 : 12;
    uint32_t freqs[256], cum_freqs[257];
    int rc = rb200_count_freqs(in_bytes.data(), in_size, freqs);
    if (rc == RB200_OK) rc = rb200_normalize_freqs(freqs, cum_freqs, 1u << scale_bits);
    if (rc != RB200_OK) die("model", rc, nullptr);

    rb200_ctx* ctx = nullptr;
    rc = rb200_ctx_create(&ctx, 0, nullptr);
    if (rc != RB200_OK) die("rb200_ctx_create (is a GPU visible? there is no CPU fallback)", rc, nullptr);
    rb200_model* model = nullptr;
    rc = rb200_model_create(ctx, alias ? RB200_CODER_ALIAS : RB200_CODER_WORD, scale_bits, freqs, &model);
    if (rc != RB200_OK) die("rb200_model_create", rc, ctx);

    const size_t cap = rb200_encode_bound(in_size, chunk);
    std::vector<uint8_t> out_buf(cap);
    std::vector<uint64_t> dir(rb200_chunk_count(in_size, chunk) + 1);
    std::vector<uint8_t> dec_bytes(in_size, 0xcc);
    size_t out_size = 0;

    std::printf("%s coder, %zu symbols, 32-way chunks of %u symbols\n", alias ? "alias" : "word", in_size, chunk);
    std::printf("GPU rANS encode (host buffers, copies included):\n");
    for (int run = 0; run < 5; run++) {
        const double t0 = now_s();
        rc = rb200_encode(ctx, model, in_bytes.data(), in_size, chunk, out_buf.data(), cap, dir.data(), &out_size, RB200_MEM_HOST);
        if (rc != RB200_OK) die("rb200_encode", rc, ctx);
        const double dt = now_s() - t0;
        std::printf("%.3f ms (%7.1f MiB/s)\n", dt * 1e3, in_size / (dt * 1048576.0));
    }
    std::printf("GPU rANS: %zu bytes\n", out_size);

    for (int run = 0; run < 5; run++) {
        const double t0 = now_s();
        rc = rb200_decode(ctx, model, out_buf.data(), out_size, dir.data(), chunk, dec_bytes.data(), in_size, RB200_MEM_HOST);
        if (rc != RB200_OK) die("rb200_decode", rc, ctx);
        const double dt = now_s() - t0;
        std::printf("%.3f ms (%7.1f MiB/s)\n", dt * 1e3, in_size / (dt * 1048576.0));
    }
    if (std::memcmp(in_bytes.data(), dec_bytes.data(), in_size) == 0)
        std::printf("decode ok!\n");
    else
        std::printf("ERROR: bad decoder!\n");

    rb200_model_destroy(model);
    rb200_ctx_destroy(ctx);
    return 0;
}
