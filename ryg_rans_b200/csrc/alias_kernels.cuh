// alias_kernels.cuh -- sm_100a kernels for the alias-table coder (main_alias.cpp semantics):
// rans_byte.h state machine (31-bit state, byte renormalisation, L = 1 << 23) with the
// alias-method symbol lookup of main_alias.cpp:252-267 (decode) and the slot remap of
// main_alias.cpp:241-250 (encode).
//
// Same mapping as the word coder: one warp = one chunk = one 32-way stream.  The byte
// coder moves 0, 1 or 2 bytes per lane per step (scale_bits <= 16: after a decode step
// x >= 2^(23-scale_bits) >= 2^7, and 2^7 << 16 = L; before an encode step x < 2^31 and
// x_max >= 2^15), so the warp-wide renormalisation needs two ballots: "at least one
// byte" and "two bytes".  Within a step lane k's bytes sit before lane k+1's, most
// significant first (RansDecRenorm, rans_byte.h:307-318, called lane by lane at
// main_alias.cpp:396-397).
#pragma once
#include "device_utils.cuh"
#include "tables.h"
#include "word_kernels.cuh"   // StreamWindow

namespace rb200 {

constexpr uint32_t kByteL = 1u << 23;   // RANS_BYTE_L, rans_byte.h:50
constexpr int kAliasDecWarps = 8;
constexpr int kAliasEncWarps = 8;

// RansDecGetAlias (main_alias.cpp:252-267) + RansDecRenorm (rans_byte.h:307-318), warp-wide
__device__ __forceinline__ void alias_dec_step(uint32_t& x, uint32_t& cursor, uint32_t div_tab, uint32_t dec_tab,
                                               uint32_t ring, uint8_t* o, uint32_t lt, uint32_t sb, bool active)
{
    bool n1 = false, n2 = false;
    if (active) {
        const uint32_t xm = x & ((1u << sb) - 1);                          // :258
        const uint32_t bucket = xm >> (sb - 8);                            // :259
        const uint32_t divider = lds_u32_ro(div_tab + 4u * bucket);
        const uint32_t b2 = 2u * bucket + (xm < divider ? 1u : 0u);        // :260-262
        const uint2 e = lds_u64_ro(dec_tab + 8u * b2);                     // {slot_freq | sym << 24, slot_adjust}
        x = (e.x & 0xffffffu) * (x >> sb) + xm - e.y;                      // :265
        *o = static_cast<uint8_t>(e.x >> 24);                              // :266
        n1 = x < kByteL;
        n2 = x < (kByteL >> 8);
    }
    const uint32_t m1 = __ballot_sync(0xffffffffu, n1);
    const uint32_t m2 = __ballot_sync(0xffffffffu, n2);
    const uint32_t a = cursor + __popc(m1 & lt) + __popc(m2 & lt);
    const uint32_t b0 = lds_u8(ring | (a & (kRingBytes - 1)));
    const uint32_t b1 = lds_u8(ring | ((a + 1) & (kRingBytes - 1)));
    if (n1) x = (x << 8) | b0;                                             // rans_byte.h:313
    if (n2) x = (x << 8) | b1;
    cursor += __popc(m1) + __popc(m2);
}

__global__ void __launch_bounds__(kAliasDecWarps * 32, 8)
alias_decode_kernel(const uint8_t* __restrict__ blob, uint64_t blob_size, const uint64_t* __restrict__ offsets, uint32_t sb,
                    const uint32_t* __restrict__ g_divider, const AliasDecEntry* __restrict__ g_dec,
                    uint8_t* __restrict__ out, uint64_t n, uint32_t chunk_syms, uint32_t n_chunks, uint32_t* __restrict__ status)
{
    __shared__ __align__(16) uint32_t s_div[256];
    __shared__ __align__(16) uint2 s_dec[512];
    __shared__ __align__(1024) uint8_t s_ring[kAliasDecWarps][kRingBytes];

    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_div[i] = g_divider[i];
    for (uint32_t i = threadIdx.x; i < 512; i += blockDim.x) s_dec[i] = make_uint2(g_dec[i].freq_sym, g_dec[i].adjust);
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t chunk = blockIdx.x * kAliasDecWarps + warp;
    if (chunk >= n_chunks) return;

    const uint64_t first = static_cast<uint64_t>(chunk) * chunk_syms;
    const uint32_t m = static_cast<uint32_t>(min(static_cast<uint64_t>(chunk_syms), n - first));
    const uint64_t off = offsets[chunk];
    const uint64_t end = offsets[chunk + 1] & ~static_cast<uint64_t>(15);
    if (off + kHeaderBytes > end || end > blob_size) {
        if (lane == 0) atomicOr(status, kStatStream);
        return;
    }

    StreamWindow win;
    win.open(blob, blob_size, off, smem_addr(&s_ring[warp][0]), lane);
    uint32_t cursor = static_cast<uint32_t>(off);
    // RansDecInit x 32 (rans_byte.h:109-122)
    uint32_t x = 0;
#pragma unroll
    for (int b = 3; b >= 0; b--) x = (x << 8) | lds_u8(win.ring | ((cursor + 4 * lane + b) & (kRingBytes - 1)));
    cursor += kHeaderBytes;

    const uint32_t lt = lanemask_lt();
    const uint32_t div_tab = smem_addr(s_div), dec_tab = smem_addr(s_dec);
    uint8_t* o = out + first + lane;
    const uint32_t steps = m >> 5, rem = m & 31;
    uint32_t g = 0;
    for (; g + 4 <= steps; g += 4) {
        win.top_up(cursor, lane);
        alias_dec_step(x, cursor, div_tab, dec_tab, win.ring, o, lt, sb, true);
        alias_dec_step(x, cursor, div_tab, dec_tab, win.ring, o + 32, lt, sb, true);
        alias_dec_step(x, cursor, div_tab, dec_tab, win.ring, o + 64, lt, sb, true);
        alias_dec_step(x, cursor, div_tab, dec_tab, win.ring, o + 96, lt, sb, true);
        o += 128;
    }
    win.top_up(cursor, lane);
    for (; g < steps; g++) {
        alias_dec_step(x, cursor, div_tab, dec_tab, win.ring, o, lt, sb, true);
        o += 32;
    }
    if (rem) alias_dec_step(x, cursor, div_tab, dec_tab, win.ring, o, lt, sb, lane < rem);

    const bool bad = (cursor != static_cast<uint32_t>(end)) || (x != kByteL);
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(status, kStatStream);
}

// RansEncPutAlias (main_alias.cpp:241-250) = RansEncRenorm (rans_byte.h:62-74) + divide + remap gather
__device__ __forceinline__ void alias_enc_step(uint32_t& x, uint32_t& emitted, uint32_t& flags, uint32_t sym, uint32_t tab,
                                               const uint16_t* __restrict__ remap, uint8_t* slot_end, uint32_t gt, uint32_t sb,
                                               bool active)
{
    bool n1 = false, n2 = false;
    uint4 e = make_uint4(0, 1, 0, 0);
    if (active) {
        e = lds_u128_ro(tab + 16u * sym);                    // {magic, freq, cum, shift}
        flags |= e.w;
        const uint32_t x_max = e.y << (31 - sb);             // ((L >> sb) << 8) * freq, rans_byte.h:64
        n1 = x >= x_max;                                     // :65
        n2 = (x >> 8) >= x_max;                              // second trip of the do/while, :67-70
    }
    const uint32_t m1 = __ballot_sync(0xffffffffu, n1);
    const uint32_t m2 = __ballot_sync(0xffffffffu, n2);
    if (n1) {
        // lanes are visited 31..0 (main_alias.cpp:365-370 generalised); each writes downwards
        const uint32_t t = emitted + __popc(m1 & gt) + __popc(m2 & gt);
        slot_end[-static_cast<int64_t>(t + 1)] = static_cast<uint8_t>(x);          // :68
        if (n2) slot_end[-static_cast<int64_t>(t + 2)] = static_cast<uint8_t>(x >> 8);
        x >>= n2 ? 16 : 8;                                                          // :69
    }
    emitted += __popc(m1) + __popc(m2);
    if (active) {
        const uint32_t q = static_cast<uint32_t>((static_cast<uint64_t>(x) + __umulhi(x, e.x)) >> (e.w & 31u));   // x / freq
        const uint32_t r = x - q * e.y;                                                                          // x % freq
        x = (q << sb) + __ldg(remap + r + e.z);                                     // main_alias.cpp:249
    }
}

__global__ void __launch_bounds__(kAliasEncWarps * 32, 4)
alias_encode_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t chunk_syms, uint32_t n_chunks, uint32_t sb,
                    const AliasEncEntry* __restrict__ g_enc, const uint16_t* __restrict__ remap,
                    uint8_t* __restrict__ scratch, uint32_t slot_bytes, uint32_t* __restrict__ sizes,
                    uint32_t* __restrict__ status)
{
    __shared__ __align__(16) uint4 s_enc[256];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        const AliasEncEntry e = g_enc[i];
        // a symbol the model does not contain: keep the step well defined (freq 1), flag it
        s_enc[i] = (e.shift & kEncBadSymbol) ? make_uint4(0, 1, 0, kEncBadSymbol) : make_uint4(e.magic, e.freq, e.cum, e.shift);
    }
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t chunk = blockIdx.x * kAliasEncWarps + warp;
    if (chunk >= n_chunks) return;

    const uint64_t first = static_cast<uint64_t>(chunk) * chunk_syms;
    const uint32_t m = static_cast<uint32_t>(min(static_cast<uint64_t>(chunk_syms), n - first));
    const uint8_t* src = in + first + lane;
    uint8_t* slot_end = scratch + static_cast<uint64_t>(chunk + 1) * slot_bytes;
    const uint32_t tab = smem_addr(s_enc);
    const uint32_t gt = lanemask_gt();

    uint32_t x = kByteL;                       // RansEncInit, rans_byte.h:56-59
    uint32_t emitted = 0, flags = 0;
    const uint32_t steps = m >> 5, rem = m & 31;
    if (rem) {
        const bool active = lane < rem;
        const uint32_t s = active ? src[static_cast<uint64_t>(steps) * 32] : 0;
        alias_enc_step(x, emitted, flags, s, tab, remap, slot_end, gt, sb, active);
    }
    uint32_t g = steps;
    for (; g >= 4; g -= 4) {
        const uint8_t* p = src + static_cast<uint64_t>(g - 4) * 32;
        const uint32_t s3 = p[96], s2 = p[64], s1 = p[32], s0 = p[0];
        alias_enc_step(x, emitted, flags, s3, tab, remap, slot_end, gt, sb, true);
        alias_enc_step(x, emitted, flags, s2, tab, remap, slot_end, gt, sb, true);
        alias_enc_step(x, emitted, flags, s1, tab, remap, slot_end, gt, sb, true);
        alias_enc_step(x, emitted, flags, s0, tab, remap, slot_end, gt, sb, true);
    }
    for (; g >= 1; g--) {
        const uint32_t s = src[static_cast<uint64_t>(g - 1) * 32];
        alias_enc_step(x, emitted, flags, s, tab, remap, slot_end, gt, sb, true);
    }

    // RansEncFlush for lanes 31..0 (rans_byte.h:93-105): 4 little-endian bytes per lane, lane 0 first
    uint8_t* head = slot_end - emitted - kHeaderBytes + 4 * lane;
    head[0] = static_cast<uint8_t>(x);
    head[1] = static_cast<uint8_t>(x >> 8);
    head[2] = static_cast<uint8_t>(x >> 16);
    head[3] = static_cast<uint8_t>(x >> 24);
    if (lane == 0) sizes[chunk] = kHeaderBytes + emitted;
    if (__any_sync(0xffffffffu, (flags & kEncBadSymbol) != 0) && lane == 0) atomicOr(status, kStatSymbol);
}

inline void configure_alias_kernels()
{
    cudaFuncSetAttribute(alias_decode_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

inline int launch_alias_encode(cudaStream_t stream, const uint8_t* d_in, uint64_t n, uint32_t chunk_syms, uint32_t n_chunks,
                               uint32_t sb, const AliasEncEntry* enc, const uint16_t* remap, uint8_t* scratch, uint32_t slot,
                               uint32_t* sizes, uint32_t* status)
{
    const uint32_t grid = (n_chunks + kAliasEncWarps - 1) / kAliasEncWarps;
    alias_encode_kernel<<<grid, kAliasEncWarps * 32, 0, stream>>>(d_in, n, chunk_syms, n_chunks, sb, enc, remap, scratch, slot,
                                                                   sizes, status);
    return 0;
}

inline int launch_alias_decode(cudaStream_t stream, const uint8_t* blob, uint64_t blob_size, const uint64_t* offsets, uint32_t sb,
                               const uint32_t* divider, const AliasDecEntry* dec, uint8_t* out, uint64_t n, uint32_t chunk_syms,
                               uint32_t n_chunks, uint32_t* status)
{
    const uint32_t grid = (n_chunks + kAliasDecWarps - 1) / kAliasDecWarps;
    alias_decode_kernel<<<grid, kAliasDecWarps * 32, 0, stream>>>(blob, blob_size, offsets, sb, divider, dec, out, n, chunk_syms,
                                                                   n_chunks, status);
    return 0;
}

}  // namespace rb200
