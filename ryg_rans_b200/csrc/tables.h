// tables.h -- host images of the device-side coding tables (internal).
#pragma once
#include <cstdint>
#include <vector>

namespace rb200 {

// marks a symbol the model gives frequency 0; the encoder flags RB200_E_SYMBOL
constexpr uint32_t kEncBadSymbol = 0x80000000u;

// ---- word coder (rans_word_sse41.h semantics, scale_bits 12)
//
// decode: one u32 per slot, 16 KiB: (freq & 0xfff) << 20 | bias << 8 | symbol.
//   It fuses RansWordSlot{freq,bias} and slot2sym (rans_word_sse41.h:50-61) so the
//   decode step is ONE shared-memory gather.  freq == 4096 (single-symbol model)
//   does not fit 12 bits; `wide` selects a kernel variant that maps 0 -> 4096.
// encode: per symbol {magic, freq | start << 13 | shift << 25}: exact division
//   q = (x + mulhi(x, magic)) >> shift for any 32-bit x (round-up reciprocal),
//   replacing the hardware divide of RansWordEncPut (rans_word_sse41.h:92).
struct WordEncEntry { uint32_t magic, packed; };
struct WordDeviceTables {
    uint32_t dec[4096];
    WordEncEntry enc[256];
    int wide;
};
int build_word_device_tables(const uint32_t freqs[256], WordDeviceTables& t);

// ---- alias coder (main_alias.cpp semantics over the rans_byte.h state machine)
//
// decode: ONE 16-byte entry per bucket (main_alias.cpp:55-59 fused from four gathers to one):
//   w0 = divider[b]
//   w1 = slot_freqs[2b]   | sym_id[2b]   << 17        (taken when xm >= divider)
//   w2 = slot_freqs[2b+1] | sym_id[2b+1] << 17        (taken when xm <  divider)
//   w3 = (slot_adjust[2b] & 0xffff) | (slot_adjust[2b+1] & 0xffff) << 16
//   xm - slot_adjust is the position inside the symbol's range, < freq <= 65536, so 16 bits of
//   the adjust are enough: bias = (xm - adjust16) & 0xffff.
// encode: per symbol {magic, freq, cum, shift} (same exact-division scheme) and
//   alias_remap as u16 (values < 65536) = 128 KiB at scale_bits 16.
struct AliasDecEntry { uint32_t divider, alt0, alt1, adjust; };
struct AliasEncEntry { uint32_t magic, freq, cum, shift; };
struct AliasDeviceTables {
    uint32_t scale_bits;
    AliasDecEntry dec[256];
    AliasEncEntry enc[256];
    std::vector<uint16_t> remap;
};
int build_alias_device_tables(const uint32_t freqs[256], uint32_t scale_bits, AliasDeviceTables& t);

}  // namespace rb200
