// oracle/_ref harness: rans64.h via main64.cpp.  TEST INFRASTRUCTURE ONLY.
#include "ref_prelude.h"

namespace ref_64 {
#define main ref_driver_main_64
#include "main64.cpp"
#undef main
}
using namespace ref_64;

// main64.cpp:228-248 generalised to nlanes
REF_EXPORT long ref_rans64_encode(const uint8_t* in, size_t n, const uint32_t* freqs, const uint32_t* cum,
                                  uint32_t scale_bits, uint32_t nlanes, uint8_t* out, size_t cap)
{
    Rans64EncSymbol esyms[256];
    for (int i = 0; i < 256; i++) Rans64EncSymbolInit(&esyms[i], cum[i], freqs[i], scale_bits);
    size_t max_words = n + 2 * (size_t)nlanes + 8;
    std::vector<uint32_t> buf(max_words);
    std::vector<Rans64State> rans(nlanes);
    for (uint32_t i = 0; i < nlanes; i++) Rans64EncInit(&rans[i]);
    uint32_t* ptr = buf.data() + max_words;
    for (size_t i = n; i > 0; i--)
        Rans64EncPutSymbol(&rans[(i - 1) % nlanes], &ptr, &esyms[in[i - 1]], scale_bits);
    for (uint32_t i = nlanes; i > 0; i--) Rans64EncFlush(&rans[i - 1], &ptr);
    size_t bytes = (size_t)(buf.data() + max_words - ptr) * 4;
    if (bytes > cap) return -3;
    memcpy(out, ptr, bytes);
    return (long)bytes;
}

// main64.cpp:261-282 generalised
REF_EXPORT long ref_rans64_decode(const uint8_t* stream, size_t size, const uint32_t* freqs, const uint32_t* cum,
                                  uint32_t scale_bits, uint32_t nlanes, uint8_t* out, size_t n)
{
    Rans64DecSymbol dsyms[256];
    for (int i = 0; i < 256; i++) Rans64DecSymbolInit(&dsyms[i], cum[i], freqs[i]);
    std::vector<uint8_t> cum2sym((size_t)1 << scale_bits);
    for (int s = 0; s < 256; s++)
        for (uint32_t i = cum[s]; i < cum[s + 1]; i++) cum2sym[i] = (uint8_t)s;
    std::vector<uint32_t> padded(size / 4 + 8, 0);
    memcpy(padded.data(), stream, size);
    uint32_t* ptr = padded.data();
    std::vector<Rans64State> rans(nlanes);
    for (uint32_t i = 0; i < nlanes; i++) Rans64DecInit(&rans[i], &ptr);
    for (size_t i = 0; i < n; i++) {
        uint32_t s = cum2sym[Rans64DecGet(&rans[i % nlanes], scale_bits)];
        out[i] = (uint8_t)s;
        Rans64DecAdvanceSymbolStep(&rans[i % nlanes], &dsyms[s], scale_bits);
        Rans64DecRenorm(&rans[i % nlanes], &ptr);
    }
    return (long)((uint8_t*)ptr - (uint8_t*)padded.data());
}

// CPU baseline: the reference's 2-way interleaved rans64 loops (main64.cpp:228-282),
// one contiguous slice per thread, best of `runs`.
REF_EXPORT int ref_cpu_baseline_rans64(const uint8_t* in, size_t n, uint32_t prob_bits, int nthreads, int runs,
                                       double* enc_seconds, double* dec_seconds, uint64_t* total_bytes)
{
    if (nthreads < 1) nthreads = 1;
    SymbolStats stats;
    stats.count_freqs(in, n);
    stats.normalize_freqs(1u << prob_bits);
    std::vector<uint8_t> cum2sym((size_t)1 << prob_bits);
    for (int s = 0; s < 256; s++)
        for (uint32_t i = stats.cum_freqs[s]; i < stats.cum_freqs[s + 1]; i++) cum2sym[i] = (uint8_t)s;
    Rans64EncSymbol esyms[256];
    Rans64DecSymbol dsyms[256];
    for (int i = 0; i < 256; i++) {
        Rans64EncSymbolInit(&esyms[i], stats.cum_freqs[i], stats.freqs[i], prob_bits);
        Rans64DecSymbolInit(&dsyms[i], stats.cum_freqs[i], stats.freqs[i]);
    }
    struct Slice { const uint8_t* in; size_t n; std::vector<uint32_t> buf; std::vector<uint8_t> dec; uint32_t* begin; size_t bytes; };
    std::vector<Slice> sl(nthreads);
    size_t per = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        size_t lo = std::min(n, per * t), hi = std::min(n, per * (t + 1));
        sl[t].in = in + lo; sl[t].n = hi - lo;
        sl[t].buf.assign(sl[t].n / 2 + sl[t].n / 4 + 64, 0);
        sl[t].dec.assign(sl[t].n + 16, 0xcc);
    }
    auto enc = [&](Slice& s) {
        size_t in_size = s.n; const uint8_t* in_bytes = s.in;
        Rans64State rans0, rans1;
        Rans64EncInit(&rans0); Rans64EncInit(&rans1);
        uint32_t* ptr = s.buf.data() + s.buf.size();
        if (in_size & 1) Rans64EncPutSymbol(&rans0, &ptr, &esyms[in_bytes[in_size - 1]], prob_bits);
        for (size_t i = (in_size & ~(size_t)1); i > 0; i -= 2) {
            Rans64EncPutSymbol(&rans1, &ptr, &esyms[in_bytes[i - 1]], prob_bits);
            Rans64EncPutSymbol(&rans0, &ptr, &esyms[in_bytes[i - 2]], prob_bits);
        }
        Rans64EncFlush(&rans1, &ptr); Rans64EncFlush(&rans0, &ptr);
        s.begin = ptr; s.bytes = (size_t)(s.buf.data() + s.buf.size() - ptr) * 4;
    };
    auto dec = [&](Slice& s) {
        size_t in_size = s.n; uint8_t* dec_bytes = s.dec.data();
        Rans64State rans0, rans1;
        uint32_t* ptr = s.begin;
        Rans64DecInit(&rans0, &ptr); Rans64DecInit(&rans1, &ptr);
        for (size_t i = 0; i < (in_size & ~(size_t)1); i += 2) {
            uint32_t s0 = cum2sym[Rans64DecGet(&rans0, prob_bits)];
            uint32_t s1 = cum2sym[Rans64DecGet(&rans1, prob_bits)];
            dec_bytes[i + 0] = (uint8_t)s0; dec_bytes[i + 1] = (uint8_t)s1;
            Rans64DecAdvanceSymbolStep(&rans0, &dsyms[s0], prob_bits);
            Rans64DecAdvanceSymbolStep(&rans1, &dsyms[s1], prob_bits);
            Rans64DecRenorm(&rans0, &ptr); Rans64DecRenorm(&rans1, &ptr);
        }
        if (in_size & 1) {
            uint32_t s0 = cum2sym[Rans64DecGet(&rans0, prob_bits)];
            dec_bytes[in_size - 1] = (uint8_t)s0;
            Rans64DecAdvanceSymbol(&rans0, &ptr, &dsyms[s0], prob_bits);
        }
    };
    double best_e = 1e30, best_d = 1e30;
    for (int r = 0; r < runs; r++) {
        double t0 = ref_now();
        { std::vector<std::thread> th; for (int t = 0; t < nthreads; t++) th.emplace_back([&, t] { enc(sl[t]); }); for (auto& x : th) x.join(); }
        double t1 = ref_now();
        { std::vector<std::thread> th; for (int t = 0; t < nthreads; t++) th.emplace_back([&, t] { dec(sl[t]); }); for (auto& x : th) x.join(); }
        double t2 = ref_now();
        best_e = std::min(best_e, t1 - t0); best_d = std::min(best_d, t2 - t1);
    }
    uint64_t tot = 0; int bad = 0;
    for (int t = 0; t < nthreads; t++) { tot += sl[t].bytes; if (memcmp(sl[t].in, sl[t].dec.data(), sl[t].n)) bad = 1; }
    *enc_seconds = best_e; *dec_seconds = best_d; *total_bytes = tot;
    return bad;
}
