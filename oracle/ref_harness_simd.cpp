// oracle/_ref harness: word coder (rans_word_sse41.h) + main_simd.cpp's SymbolStats.
// TEST INFRASTRUCTURE ONLY -- built only where /root/reference exists.
#include "ref_prelude.h"

namespace ref_simd {
#define main ref_driver_main_simd
#include "main_simd.cpp"          // resolved via -I/root/reference
#undef main
}
using namespace ref_simd;

REF_EXPORT void ref_count_freqs(const uint8_t* in, size_t n, uint32_t* freqs)
{
    SymbolStats st;
    st.count_freqs(in, n);
    memcpy(freqs, st.freqs, sizeof st.freqs);
}

// aborts (assert) exactly where the reference would
REF_EXPORT int ref_normalize_freqs(uint32_t* freqs, uint32_t* cum, uint32_t target_total)
{
    SymbolStats st;
    memcpy(st.freqs, freqs, sizeof st.freqs);
    st.normalize_freqs(target_total);
    memcpy(freqs, st.freqs, sizeof st.freqs);
    memcpy(cum, st.cum_freqs, sizeof st.cum_freqs);
    return 0;
}

REF_EXPORT void ref_word_tables(const uint32_t* freqs, const uint32_t* cum, uint32_t* slots, uint8_t* slot2sym)
{
    static RansWordTables tab;
    memset(&tab, 0, sizeof tab);
    for (int s = 0; s < 256; s++)
        RansWordTablesInitSymbol(&tab, (uint8_t)s, cum[s], freqs[s]);
    for (int i = 0; i < 4096; i++) slots[i] = tab.slots[i].u32;
    memcpy(slot2sym, tab.slot2sym, 4096);
}

// main_simd.cpp:287-300 with 8 -> nlanes
REF_EXPORT long ref_word_encode(const uint8_t* in, size_t n, const uint32_t* freqs, const uint32_t* cum,
                                uint32_t nlanes, uint8_t* out, size_t cap)
{
    size_t max_bytes = 2 * n + 4 * (size_t)nlanes + 32;
    std::vector<uint8_t> buf(max_bytes);
    std::vector<RansWordEnc> rans(nlanes);
    for (uint32_t i = 0; i < nlanes; i++) rans[i] = RansWordEncInit();
    uint16_t* ptr = (uint16_t*)(buf.data() + (max_bytes & ~(size_t)1));
    uint16_t* end = ptr;
    for (size_t i = n; i > 0; i--) {
        int s = in[i - 1];
        RansWordEncPut(&rans[(i - 1) % nlanes], &ptr, cum[s], freqs[s]);
    }
    for (uint32_t i = nlanes; i > 0; i--)
        RansWordEncFlush(&rans[i - 1], &ptr);
    size_t bytes = (size_t)(end - ptr) * 2;
    if (bytes > cap) return -3;
    memcpy(out, ptr, bytes);
    return (long)bytes;
}

// scalar N-way decode with the reference primitives (main_simd.cpp:241-263 generalised)
REF_EXPORT long ref_word_decode(const uint8_t* stream, size_t size, const uint32_t* freqs, const uint32_t* cum,
                                uint32_t nlanes, uint8_t* out, size_t n)
{
    static RansWordTables tab;
    for (int s = 0; s < 256; s++)
        RansWordTablesInitSymbol(&tab, (uint8_t)s, cum[s], freqs[s]);
    std::vector<uint8_t> padded(size + 32, 0);
    memcpy(padded.data(), stream, size);
    uint16_t* ptr = (uint16_t*)padded.data();
    std::vector<RansWordDec> rans(nlanes);
    for (uint32_t i = 0; i < nlanes; i++) RansWordDecInit(&rans[i], &ptr);
    for (size_t i = 0; i < n; i++) {
        out[i] = RansWordDecSym(&rans[i % nlanes], &tab);
        RansWordDecRenorm(&rans[i % nlanes], &ptr);
    }
    return (long)((uint8_t*)ptr - padded.data());
}

// the reference's own SSE4.1 8-way decode loop, main_simd.cpp:313-332, verbatim structure
REF_EXPORT long ref_word_decode_simd8(const uint8_t* stream, size_t size, const uint32_t* freqs, const uint32_t* cum,
                                      uint8_t* out, size_t n)
{
    static RansWordTables tab;
    for (int s = 0; s < 256; s++)
        RansWordTablesInitSymbol(&tab, (uint8_t)s, cum[s], freqs[s]);
    std::vector<uint8_t> padded(size + 32, 0);
    memcpy(padded.data(), stream, size);
    std::vector<uint8_t> dec(n + 16);
    RansSimdDec rans0, rans1;
    uint16_t* ptr = (uint16_t*)padded.data();
    RansSimdDecInit(&rans0, &ptr);
    RansSimdDecInit(&rans1, &ptr);
    for (size_t i = 0; i < (n & ~(size_t)7); i += 8) {
        uint32_t s03 = RansSimdDecSym(&rans0, &tab);
        uint32_t s47 = RansSimdDecSym(&rans1, &tab);
        memcpy(dec.data() + i, &s03, 4);
        memcpy(dec.data() + i + 4, &s47, 4);
        RansSimdDecRenorm(&rans0, &ptr);
        RansSimdDecRenorm(&rans1, &ptr);
    }
    for (size_t i = (n & ~(size_t)7); i < n; i++) {
        RansSimdDec* which = (i & 4) != 0 ? &rans1 : &rans0;
        dec[i] = RansWordDecSym(&which->lane[i & 3], &tab);
    }
    memcpy(out, dec.data(), n);
    return (long)((uint8_t*)ptr - padded.data());
}

// ---- CPU baseline: the reference's fastest word-coder paths, timed on host cores.
// Each thread owns one contiguous slice; scalar 8-way encode (main_simd.cpp:287-300)
// then 2x4-lane SSE4.1 decode (:313-332).  Best of `runs`.  Returns 0 if every
// slice round-trips.
struct SimdSlice {
    const uint8_t* in; size_t n; std::vector<uint8_t> buf; std::vector<uint8_t> dec; uint16_t* begin; size_t bytes;
};

static void simd_slice_encode(SimdSlice& sl, const SymbolStats& st)
{
    size_t cap = 2 * sl.n + 64;
    RansWordEnc rans[8];
    for (int i = 0; i < 8; i++) rans[i] = RansWordEncInit();
    uint16_t* ptr = (uint16_t*)(sl.buf.data() + cap);
    for (size_t i = sl.n; i > 0; i--) {
        int s = sl.in[i - 1];
        RansWordEncPut(&rans[(i - 1) & 7], &ptr, st.cum_freqs[s], st.freqs[s]);
    }
    for (int i = 8; i > 0; i--) RansWordEncFlush(&rans[i - 1], &ptr);
    sl.begin = ptr;
    sl.bytes = (size_t)((uint8_t*)(sl.buf.data() + cap) - (uint8_t*)ptr);
}

static void simd_slice_decode(SimdSlice& sl, const RansWordTables& tab)
{
    RansSimdDec rans0, rans1;
    uint16_t* ptr = sl.begin;
    uint8_t* dec_bytes = sl.dec.data();
    size_t in_size = sl.n;
    RansSimdDecInit(&rans0, &ptr);
    RansSimdDecInit(&rans1, &ptr);
    for (size_t i = 0; i < (in_size & ~(size_t)7); i += 8) {
        uint32_t s03 = RansSimdDecSym(&rans0, &tab);
        uint32_t s47 = RansSimdDecSym(&rans1, &tab);
        *(uint32_t*)(dec_bytes + i) = s03;
        *(uint32_t*)(dec_bytes + i + 4) = s47;
        RansSimdDecRenorm(&rans0, &ptr);
        RansSimdDecRenorm(&rans1, &ptr);
    }
    for (size_t i = (in_size & ~(size_t)7); i < in_size; i++) {
        RansSimdDec* which = (i & 4) != 0 ? &rans1 : &rans0;
        dec_bytes[i] = RansWordDecSym(&which->lane[i & 3], &tab);
    }
}

REF_EXPORT int ref_cpu_baseline_simd(const uint8_t* in, size_t n, int nthreads, int runs,
                                     double* enc_seconds, double* dec_seconds, uint64_t* total_bytes)
{
    if (nthreads < 1) nthreads = 1;
    SymbolStats st;
    st.count_freqs(in, n);
    st.normalize_freqs(RANS_WORD_M);
    static RansWordTables tab;
    for (int s = 0; s < 256; s++)
        RansWordTablesInitSymbol(&tab, (uint8_t)s, st.cum_freqs[s], st.freqs[s]);

    std::vector<SimdSlice> sl(nthreads);
    size_t per = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        size_t lo = std::min(n, per * t), hi = std::min(n, per * (t + 1));
        sl[t].in = in + lo; sl[t].n = hi - lo;
        sl[t].buf.assign(2 * sl[t].n + 64 + 32, 0);
        sl[t].dec.assign(sl[t].n + 16, 0xcc);
    }
    double best_e = 1e30, best_d = 1e30;
    for (int r = 0; r < runs; r++) {
        double t0 = ref_now();
        { std::vector<std::thread> th;
          for (int t = 0; t < nthreads; t++) th.emplace_back([&, t] { simd_slice_encode(sl[t], st); });
          for (auto& x : th) x.join(); }
        double t1 = ref_now();
        { std::vector<std::thread> th;
          for (int t = 0; t < nthreads; t++) th.emplace_back([&, t] { simd_slice_decode(sl[t], tab); });
          for (auto& x : th) x.join(); }
        double t2 = ref_now();
        best_e = std::min(best_e, t1 - t0);
        best_d = std::min(best_d, t2 - t1);
    }
    uint64_t tot = 0; int bad = 0;
    for (int t = 0; t < nthreads; t++) {
        tot += sl[t].bytes;
        if (memcmp(sl[t].in, sl[t].dec.data(), sl[t].n) != 0) bad = 1;
    }
    *enc_seconds = best_e; *dec_seconds = best_d; *total_bytes = tot;
    return bad;
}
