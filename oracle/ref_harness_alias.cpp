// oracle/_ref harness: alias coder (main_alias.cpp's SymbolStats::make_alias_table,
// RansEncPutAlias, RansDecGetAlias) over rans_byte.h.  TEST INFRASTRUCTURE ONLY.
#include "ref_prelude.h"

namespace ref_alias {
#define main ref_driver_main_alias
#include "main_alias.cpp"
#undef main
}
using namespace ref_alias;

static void load_stats(SymbolStats& st, const uint32_t* freqs, const uint32_t* cum)
{
    memcpy(st.freqs, freqs, sizeof st.freqs);
    memcpy(st.cum_freqs, cum, sizeof st.cum_freqs);
    st.make_alias_table();
}

// tables out: divider[256], slot_adjust[512], slot_freqs[512], sym_id[512], remap[cum[256]]
REF_EXPORT int ref_alias_build(const uint32_t* freqs, const uint32_t* cum, uint32_t* divider, uint32_t* slot_adjust,
                               uint32_t* slot_freqs, uint8_t* sym_id, uint32_t* remap)
{
    SymbolStats st;
    load_stats(st, freqs, cum);
    memcpy(divider, st.divider, sizeof st.divider);
    memcpy(slot_adjust, st.slot_adjust, sizeof st.slot_adjust);
    memcpy(slot_freqs, st.slot_freqs, sizeof st.slot_freqs);
    memcpy(sym_id, st.sym_id, sizeof st.sym_id);
    memcpy(remap, st.alias_remap, (size_t)cum[256] * sizeof(uint32_t));
    return 0;
}

// main_alias.cpp:353-373 generalised to nlanes
REF_EXPORT long ref_alias_encode(const uint8_t* in, size_t n, const uint32_t* freqs, const uint32_t* cum,
                                 uint32_t scale_bits, uint32_t nlanes, uint8_t* out, size_t cap)
{
    SymbolStats st;
    load_stats(st, freqs, cum);
    size_t max_bytes = 2 * n + 4 * (size_t)nlanes + 32;
    std::vector<uint8_t> buf(max_bytes);
    std::vector<RansState> rans(nlanes);
    for (uint32_t i = 0; i < nlanes; i++) RansEncInit(&rans[i]);
    uint8_t* ptr = buf.data() + max_bytes;
    for (size_t i = n; i > 0; i--)
        RansEncPutAlias(&rans[(i - 1) % nlanes], &ptr, &st, in[i - 1], scale_bits);
    for (uint32_t i = nlanes; i > 0; i--) RansEncFlush(&rans[i - 1], &ptr);
    size_t bytes = (size_t)(buf.data() + max_bytes - ptr);
    if (bytes > cap) return -3;
    memcpy(out, ptr, bytes);
    return (long)bytes;
}

// main_alias.cpp:386-405 generalised
REF_EXPORT long ref_alias_decode(const uint8_t* stream, size_t size, const uint32_t* freqs, const uint32_t* cum,
                                 uint32_t scale_bits, uint32_t nlanes, uint8_t* out, size_t n)
{
    SymbolStats st;
    load_stats(st, freqs, cum);
    std::vector<uint8_t> padded(size + 16, 0);
    memcpy(padded.data(), stream, size);
    uint8_t* ptr = padded.data();
    std::vector<RansState> rans(nlanes);
    for (uint32_t i = 0; i < nlanes; i++) RansDecInit(&rans[i], &ptr);
    for (size_t i = 0; i < n; i++) {
        out[i] = (uint8_t)RansDecGetAlias(&rans[i % nlanes], &st, scale_bits);
        RansDecRenorm(&rans[i % nlanes], &ptr);
    }
    return (long)(ptr - padded.data());
}

// CPU baseline: reference 2-way alias loops (main_alias.cpp:353-405), one slice per thread.
REF_EXPORT int ref_cpu_baseline_alias(const uint8_t* in, size_t n, uint32_t scale_bits, int nthreads, int runs,
                                      double* enc_seconds, double* dec_seconds, uint64_t* total_bytes)
{
    if (nthreads < 1) nthreads = 1;
    SymbolStats st;
    st.count_freqs(in, n);
    st.normalize_freqs(1u << scale_bits);
    st.make_alias_table();
    struct Slice { const uint8_t* in; size_t n; std::vector<uint8_t> buf, dec; uint8_t* begin; size_t bytes; };
    std::vector<Slice> sl(nthreads);
    size_t per = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        size_t lo = std::min(n, per * t), hi = std::min(n, per * (t + 1));
        sl[t].in = in + lo; sl[t].n = hi - lo;
        sl[t].buf.assign(2 * sl[t].n + 64, 0);
        sl[t].dec.assign(sl[t].n + 16, 0xcc);
    }
    auto enc = [&](Slice& s) {
        size_t in_size = s.n; const uint8_t* in_bytes = s.in;
        RansState rans0, rans1;
        RansEncInit(&rans0); RansEncInit(&rans1);
        uint8_t* ptr = s.buf.data() + s.buf.size();
        if (in_size & 1) RansEncPutAlias(&rans0, &ptr, &st, in_bytes[in_size - 1], scale_bits);
        for (size_t i = (in_size & ~(size_t)1); i > 0; i -= 2) {
            RansEncPutAlias(&rans1, &ptr, &st, in_bytes[i - 1], scale_bits);
            RansEncPutAlias(&rans0, &ptr, &st, in_bytes[i - 2], scale_bits);
        }
        RansEncFlush(&rans1, &ptr); RansEncFlush(&rans0, &ptr);
        s.begin = ptr; s.bytes = (size_t)(s.buf.data() + s.buf.size() - ptr);
    };
    auto dec = [&](Slice& s) {
        size_t in_size = s.n; uint8_t* dec_bytes = s.dec.data();
        RansState rans0, rans1;
        uint8_t* ptr = s.begin;
        RansDecInit(&rans0, &ptr); RansDecInit(&rans1, &ptr);
        for (size_t i = 0; i < (in_size & ~(size_t)1); i += 2) {
            uint32_t s0 = RansDecGetAlias(&rans0, &st, scale_bits);
            uint32_t s1 = RansDecGetAlias(&rans1, &st, scale_bits);
            dec_bytes[i + 0] = (uint8_t)s0; dec_bytes[i + 1] = (uint8_t)s1;
            RansDecRenorm(&rans0, &ptr); RansDecRenorm(&rans1, &ptr);
        }
        if (in_size & 1) {
            dec_bytes[in_size - 1] = (uint8_t)RansDecGetAlias(&rans0, &st, scale_bits);
            RansDecRenorm(&rans0, &ptr);
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
