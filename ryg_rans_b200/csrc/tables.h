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
//   `enc32` is the same table with a 32-bit reciprocal: the encoder only divides states x < freq << 20
//   (RansWordEncPut renormalises first), and on that range q = mulhi(x, M32) >> s32 with
//   M32 = ceil(2^(32+s32) / freq), s32 = ceil(log2 freq) - 1 is exact for every freq <= 2963 and most above
//   (checked per symbol: (freq * 2^20 - 1) * (M32 * freq - 2^(32+s32)) < 2^(32+s32)); three instructions
//   fewer per symbol.  freq 1 uses M32 = 2^32 - 1, which yields x - 1; the kernel adds the missing
//   1 * (4096 - 1) to `start`.  `enc32_ok` says every symbol of the model passed the check.
struct WordEncEntry { uint32_t magic, packed; };
struct WordDeviceTables {
    uint32_t dec[4096];
    WordEncEntry enc[256];
    WordEncEntry enc32[256];
    int wide;
    int enc32_ok;
};
int build_word_device_tables(const uint32_t freqs[256], WordDeviceTables& t);

// ---- alias coder (main_alias.cpp semantics over the rans_byte.h state machine)
//
// decode: ONE 16-byte entry per bucket (main_alias.cpp:55-59 fused from four gathers to one):
//   w0 = divider[b]
//   w1 = slot_freqs[2b]   << 8 | sym_id[2b]           (taken when xm >= divider; slot_freqs <= 65536: 25 bits)
//   w2 = slot_freqs[2b+1] << 8 | sym_id[2b+1]         (taken when xm <  divider)
//   (the symbol in the low byte is what STG.U8 stores; the frequency is one shift away)
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

// ---- byte coder with cum2sym (main.cpp semantics over rans_byte.h), scale_bits 8..16
//
// decode: cum2sym[1 << scale_bits] (main.cpp:145-148) followed by 256 x u32 {start | freq << 16}
//   (= RansDecSymbol, rans_byte.h:168-171; like the reference's u16 field, freq must be < 65536).
// encode: RansEncSymbol images (rans_byte.h:159-165, built as RansEncSymbolInit :174-243 does)
//   stored in AliasEncEntry's four words: {x_max, rcp_freq, bias, cmpl_freq | rcp_shift << 16}.
struct ByteDeviceTables {
    uint32_t scale_bits;
    std::vector<uint8_t> dec;          // (1 << scale_bits) + 1024 bytes
    AliasEncEntry enc[256];
};
int build_byte_device_tables(const uint32_t freqs[256], uint32_t scale_bits, ByteDeviceTables& t);

// ---- rans64 (main64.cpp semantics over rans64.h), scale_bits 8..16
//
// decode: cum2sym[1 << scale_bits] followed by 256 x {u32 start, u32 freq} (= Rans64DecSymbol, rans64.h:151-154)
// encode: 256 x 32 B: {rcp_lo, rcp_hi, freq, bias}, {cmpl_freq, rcp_shift, 0, 0} (= Rans64EncSymbol, rans64.h:142-148,
//   built as Rans64EncSymbolInit :167-247 does); bit 31 of rcp_shift marks a symbol outside the model.
struct Rans64DeviceTables {
    uint32_t scale_bits;
    std::vector<uint8_t> dec;          // (1 << scale_bits) + 2048 bytes
    uint32_t enc[256][8];
};
int build_rans64_device_tables(const uint32_t freqs[256], uint32_t scale_bits, Rans64DeviceTables& t);

}  // namespace rb200
