// rans_alias.h -- source-level home of the alias-method rANS coder.
//
// The reference keeps these pieces inside its demo driver: the table fields of `SymbolStats`
// (main_alias.cpp:47-72), `SymbolStats::make_alias_table` (:147-237), `RansEncPutAlias` (:241-250) and
// `RansDecGetAlias` (:252-267).  Here they are a header like rans_byte.h, host + device: same names, same
// argument meaning, own bodies.  The two step functions are templates on the stats type, so they work on
// `RansAliasTables` below AND on a driver's own `SymbolStats` (any type with the reference's field names).
// The bulk GPU path (rb200_encode / rb200_decode with RB200_CODER_ALIAS) does not go through this header; a
// CUDA kernel of a caller can, exactly as with rans_byte.h.
#ifndef RANS_ALIAS_HEADER
#define RANS_ALIAS_HEADER

#include <stdint.h>

#include "rans_byte.h"
#include "rans_hd.h"

// The alias tables of one 256-symbol model (field names as main_alias.cpp:47-62).
struct RansAliasTables {
    static const int LOG2NSYMS = 8;
    static const int NSYMS = 1 << LOG2NSYMS;

    uint32_t freqs[NSYMS];
    uint32_t cum_freqs[NSYMS + 1];

    uint32_t divider[NSYMS];             // decoder: x mod M below this -> the bucket's own symbol
    uint32_t slot_adjust[NSYMS * 2];
    uint32_t slot_freqs[NSYMS * 2];
    uint8_t sym_id[NSYMS * 2];

    uint32_t* alias_remap;               // encoder: cum_freqs[NSYMS] entries, owned by the caller
};

// SymbolStats::make_alias_table (main_alias.cpp:147-237): Vose-style pairing of under-full buckets with over-full
// donors, then the slot hand-out.  freqs / cum_freqs must be filled in, cum_freqs[NSYMS] a multiple of NSYMS
// (the reference asserts it, :151) and alias_remap must have cum_freqs[NSYMS] entries.  Returns 0, or -1 where
// the reference would assert.  Host only (set-up code).
static inline int RansAliasTablesInit(RansAliasTables* t)
{
    const int N = RansAliasTables::NSYMS;
    const uint32_t total = t->cum_freqs[N];
    if (total == 0 || total % N != 0) return -1;
    const uint32_t bucket = total / N;

    uint32_t rest[RansAliasTables::NSYMS];          // what each symbol still has to hand out
    for (int s = 0; s < N; s++) {
        rest[s] = t->freqs[s];
        t->divider[s] = bucket;
        t->sym_id[2 * s] = t->sym_id[2 * s + 1] = (uint8_t)s;
    }
    int donor = 0, taker = 0;
    while (donor < N && rest[donor] < bucket) donor++;
    while (taker < N && rest[taker] >= bucket) taker++;
    int resume = taker + 1;
    while (donor < N && taker < N) {
        t->sym_id[2 * taker] = (uint8_t)donor;
        t->divider[taker] = rest[taker];
        rest[donor] -= bucket - t->divider[taker];
        if (rest[donor] >= bucket || resume <= donor) {
            taker = resume;
            while (taker < N && rest[taker] >= bucket) taker++;
            resume = taker + 1;
        } else {
            taker = donor;                           // the donor fell below a full bucket behind the sweep
        }
        while (donor < N && rest[donor] < bucket) donor++;
    }

    uint32_t given[RansAliasTables::NSYMS];
    for (int s = 0; s < N; s++) given[s] = 0;
    for (int s = 0; s < N; s++) {
        const int other = t->sym_id[2 * s];
        const uint32_t own = t->divider[s], lent = bucket - own;
        const uint32_t own_base = given[s], other_base = given[other];
        const uint32_t lo = (uint32_t)s * bucket;
        t->divider[s] = lo + own;
        t->slot_freqs[2 * s + 1] = t->freqs[s];
        t->slot_freqs[2 * s] = t->freqs[other];
        t->slot_adjust[2 * s + 1] = lo - own_base;
        t->slot_adjust[2 * s] = lo - (other_base - own);
        uint32_t* dst = t->alias_remap + t->cum_freqs[s] + own_base;
        for (uint32_t k = 0; k < own; k++) dst[k] = lo + k;
        dst = t->alias_remap + t->cum_freqs[other] + other_base;
        for (uint32_t k = 0; k < lent; k++) dst[k] = lo + own + k;
        given[s] += own;
        given[other] += lent;
    }
    for (int s = 0; s < N; s++)
        if (given[s] != t->freqs[s]) return -1;
    return 0;
}

// RansEncPutAlias, main_alias.cpp:241-250: renormalise, then x = (x / freq) << scale_bits + alias_remap[x % freq + cum].
template <class Stats>
RANS_HD void RansEncPutAlias(RansState* r, uint8_t** pptr, Stats* const syms, int s, uint32_t scale_bits)
{
    const uint32_t freq = syms->freqs[s];
    const RansState x = RansEncRenorm(*r, pptr, freq, scale_bits);
    const uint32_t q = x / freq;
    *r = (q << scale_bits) + syms->alias_remap[(x - q * freq) + syms->cum_freqs[s]];
}

// RansDecGetAlias, main_alias.cpp:252-267: bucket from the top 8 bits of x mod M, the divider picks one of the
// bucket's two symbols, then the state update; returns the symbol.  The caller renormalises (RansDecRenorm).
template <class Stats>
RANS_HD uint32_t RansDecGetAlias(RansState* r, Stats* const syms, uint32_t scale_bits)
{
    const RansState x = *r;
    const uint32_t xm = x & ((1u << scale_bits) - 1);
    const uint32_t bucket = xm >> (scale_bits - 8);                      // LOG2NSYMS = 8
    const uint32_t b2 = 2 * bucket + (xm < syms->divider[bucket] ? 1u : 0u);
    *r = syms->slot_freqs[b2] * (x >> scale_bits) + xm - syms->slot_adjust[b2];
    return syms->sym_id[b2];
}

#endif  // RANS_ALIAS_HEADER
