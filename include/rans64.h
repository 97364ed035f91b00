// rans64.h -- 64-bit-state rANS step functions (32-bit renormalisation), host + device.
//
// Same role as rans_byte.h in this package: the reference's names, types and signatures
// (rygorous/ryg_rans rans64.h; lines cited) over this repo's own bodies, so main64.cpp
// compiles unchanged and kernels can call the same functions.  State lives in
// [2^31, 2^63); every step moves at most one 32-bit word, so renormalisation never loops.
// The word stream is native-endian u32 (as the reference, README:12).
#ifndef RANS64_HEADER
#define RANS64_HEADER

#include <stdint.h>
#include "rans_hd.h"

#ifdef assert
#define Rans64Assert assert
#else
#define Rans64Assert(x)
#endif

#define RANS64_L (1ull << 31)                                              // ref :59

typedef uint64_t Rans64State;                                              // ref :62

typedef struct {                                                           // ref :142-148
    uint64_t rcp_freq;
    uint32_t freq;
    uint32_t bias;
    uint32_t cmpl_freq;
    uint32_t rcp_shift;
} Rans64EncSymbol;

typedef struct {                                                           // ref :151-154
    uint32_t start;
    uint32_t freq;
} Rans64DecSymbol;

RANS_HD uint64_t Rans64MulHi(uint64_t a, uint64_t b)                       // ref :35-45
{
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#elif defined(_MSC_VER)
    return __umulh(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

namespace rans_detail {

struct Wide64 {
    static RANS_HDM uint64_t limit_for(uint32_t freq, uint32_t scale_bits) { return ((RANS64_L >> scale_bits) << 32) * freq; }
    // emit the low word if x has outgrown the interval for this freq (at most once)
    static RANS_HDM uint64_t shrink(uint64_t x, uint32_t** pptr, uint64_t limit)
    {
        if (x >= limit) {
            *--*pptr = (uint32_t)x;
            x >>= 32;
        }
        return x;
    }
    static RANS_HDM uint64_t grow(uint64_t x, uint32_t** pptr)
    {
        if (x < RANS64_L) {
            x = (x << 32) | *(*pptr)++;
            Rans64Assert(x >= RANS64_L);
        }
        return x;
    }
    static RANS_HDM uint64_t pop(uint64_t x, uint32_t start, uint32_t freq, uint32_t scale_bits)
    {
        return freq * (x >> scale_bits) + (x & ((1ull << scale_bits) - 1)) - start;
    }
};

}  // namespace rans_detail

RANS_HD void Rans64EncInit(Rans64State* r) { *r = RANS64_L; }                                      // ref :65

RANS_HD void Rans64EncPut(Rans64State* r, uint32_t** pptr, uint32_t start, uint32_t freq, uint32_t scale_bits)   // ref :77
{
    Rans64Assert(freq != 0);
    const uint64_t x = rans_detail::Wide64::shrink(*r, pptr, rans_detail::Wide64::limit_for(freq, scale_bits));
    *r = ((x / freq) << scale_bits) + (x % freq) + start;
}

RANS_HD void Rans64EncFlush(Rans64State* r, uint32_t** pptr)                                       // ref :96
{
    uint32_t* p = *pptr - 2;
    p[0] = (uint32_t)*r;
    p[1] = (uint32_t)(*r >> 32);
    *pptr = p;
}

RANS_HD void Rans64DecInit(Rans64State* r, uint32_t** pptr)                                        // ref :107
{
    const uint32_t* p = *pptr;
    *r = (uint64_t)p[0] | ((uint64_t)p[1] << 32);
    *pptr += 2;
}

RANS_HD uint32_t Rans64DecGet(Rans64State* r, uint32_t scale_bits)                                 // ref :118
{
    return (uint32_t)*r & ((1u << scale_bits) - 1);
}

RANS_HD void Rans64DecAdvance(Rans64State* r, uint32_t** pptr, uint32_t start, uint32_t freq, uint32_t scale_bits)   // ref :126
{
    *r = rans_detail::Wide64::grow(rans_detail::Wide64::pop(*r, start, freq, scale_bits), pptr);
}

// ref :167-247.  rcp = ceil(2^(63+k) / freq), k = ceil(log2 freq): exact reciprocal for
// every 63-bit x.  The 128-bit dividend 2^(63+k) + freq - 1 is divided in two 64-bit
// halves (high half 2^(31+k), low half freq - 1), which is valid because the first
// remainder, shifted up by 32, plus freq - 1 still fits 64 bits (freq < 2^31).
RANS_HD void Rans64EncSymbolInit(Rans64EncSymbol* s, uint32_t start, uint32_t freq, uint32_t scale_bits)
{
    Rans64Assert(scale_bits <= 31);
    Rans64Assert(start <= (1u << scale_bits));
    Rans64Assert(freq <= (1u << scale_bits) - start);
    const uint32_t M = 1u << scale_bits;
    s->freq = freq;
    s->cmpl_freq = M - freq;
    if (freq >= 2) {
        uint32_t k = 0;
        while ((1u << k) < freq) k++;
        const uint64_t hi = 1ull << (k + 31);
        const uint64_t q_hi = hi / freq;
        const uint64_t lo = ((hi % freq) << 32) + (freq - 1);
        s->rcp_freq = (q_hi << 32) + lo / freq;
        s->rcp_shift = k - 1;
        s->bias = start;
    } else {                       // freq == 1: q = x - 1, compensate in bias (see rans_byte.h)
        s->rcp_freq = ~0ull;
        s->rcp_shift = 0;
        s->bias = start + M - 1;
    }
}

RANS_HD void Rans64DecSymbolInit(Rans64DecSymbol* s, uint32_t start, uint32_t freq)                // ref :250
{
    Rans64Assert(start <= (1 << 31));
    Rans64Assert(freq <= (1 << 31) - start);
    s->start = start;
    s->freq = freq;
}

RANS_HD void Rans64EncPutSymbol(Rans64State* r, uint32_t** pptr, Rans64EncSymbol const* sym, uint32_t scale_bits)   // ref :262
{
    Rans64Assert(sym->freq != 0);
    const uint64_t x = rans_detail::Wide64::shrink(*r, pptr, rans_detail::Wide64::limit_for(sym->freq, scale_bits));
    const uint64_t q = Rans64MulHi(x, sym->rcp_freq) >> sym->rcp_shift;
    *r = x + sym->bias + q * sym->cmpl_freq;
}

RANS_HD void Rans64DecAdvanceSymbol(Rans64State* r, uint32_t** pptr, Rans64DecSymbol const* sym, uint32_t scale_bits)   // ref :281
{
    Rans64DecAdvance(r, pptr, sym->start, sym->freq, scale_bits);
}

RANS_HD void Rans64DecAdvanceStep(Rans64State* r, uint32_t start, uint32_t freq, uint32_t scale_bits)   // ref :289
{
    *r = rans_detail::Wide64::pop(*r, start, freq, scale_bits);
}

RANS_HD void Rans64DecAdvanceSymbolStep(Rans64State* r, Rans64DecSymbol const* sym, uint32_t scale_bits)   // ref :299
{
    Rans64DecAdvanceStep(r, sym->start, sym->freq, scale_bits);
}

RANS_HD void Rans64DecRenorm(Rans64State* r, uint32_t** pptr)                                      // ref :305
{
    *r = rans_detail::Wide64::grow(*r, pptr);
}

#endif  // RANS64_HEADER
