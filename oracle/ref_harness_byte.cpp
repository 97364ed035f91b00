// oracle/_ref harness: rans_byte.h via main.cpp (RansEncPutSymbol / cum2sym + RansDecAdvanceSymbolStep).
// TEST INFRASTRUCTURE ONLY.
#include "ref_prelude.h"

namespace ref_byte {
#define main ref_driver_main_byte
#include "main.cpp"
#undef main
}
using namespace ref_byte;

// main.cpp:226-246 generalised to nlanes
REF_EXPORT long ref_byte_encode(const uint8_t* in, size_t n, const uint32_t* freqs, const uint32_t* cum,
                                uint32_t scale_bits, uint32_t nlanes, uint8_t* out, size_t cap)
{
    RansEncSymbol esyms[256];
    for (int i = 0; i < 256; i++) RansEncSymbolInit(&esyms[i], cum[i], freqs[i], scale_bits);
    size_t max_bytes = 2 * n + 4 * (size_t)nlanes + 32;
    std::vector<uint8_t> buf(max_bytes);
    std::vector<RansState> rans(nlanes);
    for (uint32_t i = 0; i < nlanes; i++) RansEncInit(&rans[i]);
    uint8_t* ptr = buf.data() + max_bytes;
    for (size_t i = n; i > 0; i--)
        RansEncPutSymbol(&rans[(i - 1) % nlanes], &ptr, &esyms[in[i - 1]]);
    for (uint32_t i = nlanes; i > 0; i--) RansEncFlush(&rans[i - 1], &ptr);
    size_t bytes = (size_t)(buf.data() + max_bytes - ptr);
    if (bytes > cap) return -3;
    memcpy(out, ptr, bytes);
    return (long)bytes;
}

// main.cpp:259-280 generalised
REF_EXPORT long ref_byte_decode(const uint8_t* stream, size_t size, const uint32_t* freqs, const uint32_t* cum,
                                uint32_t scale_bits, uint32_t nlanes, uint8_t* out, size_t n)
{
    RansDecSymbol dsyms[256];
    for (int i = 0; i < 256; i++) RansDecSymbolInit(&dsyms[i], cum[i], freqs[i]);
    std::vector<uint8_t> cum2sym((size_t)1 << scale_bits);
    for (int s = 0; s < 256; s++)
        for (uint32_t i = cum[s]; i < cum[s + 1]; i++) cum2sym[i] = (uint8_t)s;
    std::vector<uint8_t> padded(size + 16, 0);
    memcpy(padded.data(), stream, size);
    uint8_t* ptr = padded.data();
    std::vector<RansState> rans(nlanes);
    for (uint32_t i = 0; i < nlanes; i++) RansDecInit(&rans[i], &ptr);
    for (size_t i = 0; i < n; i++) {
        uint32_t s = cum2sym[RansDecGet(&rans[i % nlanes], scale_bits)];
        out[i] = (uint8_t)s;
        RansDecAdvanceSymbolStep(&rans[i % nlanes], &dsyms[s], scale_bits);
        RansDecRenorm(&rans[i % nlanes], &ptr);
    }
    return (long)(ptr - padded.data());
}
