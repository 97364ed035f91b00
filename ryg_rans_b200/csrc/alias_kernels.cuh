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
#include "word_decode_tma.cuh"   // persistent decode plumbing: DecPolicy, TmaWindow, tma_issue / tma_advance

namespace rb200 {

constexpr uint32_t kByteL = 1u << 23;   // RANS_BYTE_L, rans_byte.h:50
constexpr int kAliasDecWarps = 16;
constexpr uint32_t kAliasDecReplicas = 8;    // quarter-warp lanes hit 8 different 16-byte bank groups

// One decode step of the byte-renormalising coders for the whole warp.
//   ALIAS:  RansDecGetAlias (main_alias.cpp:252-267); tab_lane = this lane's replica of the 256 x 16 B
//           bucket table.
//   !ALIAS: RansDecGet -> cum2sym -> RansDecAdvanceSymbolStep (rans_byte.h:125-128, main.cpp:200,
//           rans_byte.h:291-304); tab_lane = shared address of cum2sym[1 << sb], followed by the
//           256 x {start | freq << 16} table.
// then RansDecRenorm (rans_byte.h:307-318) for both.
// FULL: all 32 lanes take part (active == true at compile time); only the ragged last step of a chunk does not.
// SB: scale_bits as a compile-time constant (16, 14, 12: masks and shift counts become immediates), 0 = use `sb_rt`.
template <bool ALIAS, bool FULL, uint32_t SB>
__device__ __forceinline__ void alias_dec_step(uint32_t& x, uint32_t& cursor, uint32_t tab_lane, uint32_t ring, uint8_t* o,
                                               uint32_t lt, uint32_t sb_rt, bool active)
{
    const uint32_t sb = SB ? SB : sb_rt;
    bool n1 = false, n2 = false;
    if (active) {
        const uint32_t xm = x & ((1u << sb) - 1);                          // main_alias.cpp:258 / rans_byte.h:127
        if (ALIAS) {
            const uint32_t bucket = xm >> (sb - 8);                        // :259
            const uint4 e = lds_u128_ro(mad_u32(bucket, kAliasDecReplicas * 16, tab_lane));
            const bool own = xm < e.x;                                     // :261 (bucket2 = 2 * bucket + 1)
            const uint32_t fs = own ? e.z : e.y;                           // slot_freqs << 8 | sym_id
            const uint32_t adj = own ? (e.w >> 16) : (e.w & 0xffffu);
            x = (fs >> 8) * (x >> sb) + ((xm - adj) & 0xffffu);            // :265
            *o = static_cast<uint8_t>(fs);                                 // :266
        } else {
            const uint32_t s = lds_u8_ro(tab_lane + xm);                   // cum2sym, main.cpp:200
            const uint32_t ds = lds_u32_ro(tab_lane + (1u << sb) + 4u * s);   // RansDecSymbol {start, freq}
            x = (ds >> 16) * (x >> sb) + xm - (ds & 0xffffu);              // rans_byte.h:297
            *o = static_cast<uint8_t>(s);
        }
        n1 = x < kByteL;
        n2 = x < (kByteL >> 8);
    }
    if (FULL) {
        // RansDecRenorm for the warp as one PTX sequence: two predicates feed the votes, the two ranked byte
        // loads and the state updates (the C++ below costs six more instructions and a branch per step)
        asm volatile(
            "{\n\t"
            ".reg .pred p1, p2;\n\t"
            ".reg .b32 m, r, a, b;\n\t"
            "setp.lt.u32 p1, %0, %2;\n\t"
            "setp.lt.u32 p2, %0, %3;\n\t"
            "vote.sync.ballot.b32 m, p1, 0xffffffff;\n\t"
            "and.b32 r, m, %4;\n\t"
            "popc.b32 r, r;\n\t"
            "add.u32 a, %1, r;\n\t"
            "popc.b32 m, m;\n\t"
            "add.u32 %1, %1, m;\n\t"
            "vote.sync.ballot.b32 m, p2, 0xffffffff;\n\t"
            "and.b32 r, m, %4;\n\t"
            "popc.b32 r, r;\n\t"
            "add.u32 a, a, r;\n\t"
            "popc.b32 m, m;\n\t"
            "add.u32 %1, %1, m;\n\t"
            "add.u32 b, a, 1;\n\t"
            "and.b32 a, a, %5;\n\t"
            "and.b32 b, b, %5;\n\t"
            "or.b32 a, a, %6;\n\t"
            "or.b32 b, b, %6;\n\t"
            "@p1 ld.shared.u8 a, [a];\n\t"
            "@p2 ld.shared.u8 b, [b];\n\t"
            "@p1 mad.lo.u32 %0, %0, 256, a;\n\t"
            "@p2 mad.lo.u32 %0, %0, 256, b;\n\t"
            "}"
            : "+r"(x), "+r"(cursor)
            : "n"(kByteL), "n"(kByteL >> 8), "r"(lt), "n"(kRingBytes - 1), "r"(ring)
            : "memory");
        return;
    }
    const uint32_t m1 = __ballot_sync(0xffffffffu, n1);
    const uint32_t m2 = __ballot_sync(0xffffffffu, n2);
    uint32_t a = cursor + __popc(m1 & lt);
    if (m2) a += __popc(m2 & lt);                                          // warp-uniform: two-byte refills are the rarer case
    const uint32_t b0 = lds_u8(ring | (a & (kRingBytes - 1)));
    if (n1) x = (x << 8) | b0;                                             // rans_byte.h:313
    cursor += __popc(m1);
    if (m2) {
        const uint32_t b1 = lds_u8(ring | ((a + 1) & (kRingBytes - 1)));
        if (n2) x = (x << 8) | b1;
        cursor += __popc(m2);
    }
}

// ALIAS: g_tab = 256 x AliasDecEntry.  !ALIAS: g_tab = cum2sym[1 << sb] followed by 256 x u32 {start | freq << 16}.
template <bool ALIAS, uint32_t SB>
__global__ void __launch_bounds__(kAliasDecWarps * 32, 4)
alias_decode_kernel(const uint8_t* __restrict__ blob, uint64_t blob_size, const uint64_t* __restrict__ offsets, uint32_t sb_rt,
                    const void* __restrict__ g_tab, uint8_t* __restrict__ out, uint64_t n, uint32_t chunk_syms,
                    uint32_t n_chunks, uint32_t* __restrict__ status)
{
    const uint32_t sb = SB ? SB : sb_rt;
    extern __shared__ __align__(1024) uint8_t s_adec[];       // [16 x 1 KiB rings][table]
    uint4* s_tab = reinterpret_cast<uint4*>(s_adec + kAliasDecWarps * kRingBytes);
    if (ALIAS) {
        const AliasDecEntry* g_dec = static_cast<const AliasDecEntry*>(g_tab);
        for (uint32_t i = threadIdx.x; i < 256 * kAliasDecReplicas; i += blockDim.x) {
            const AliasDecEntry e = g_dec[i / kAliasDecReplicas];
            s_tab[i] = make_uint4(e.divider, e.alt0, e.alt1, e.adjust);
        }
    } else {
        const uint32_t vecs = ((1u << sb) + 1024) / 16;
        for (uint32_t i = threadIdx.x; i < vecs; i += blockDim.x) s_tab[i] = static_cast<const uint4*>(g_tab)[i];
    }
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
    win.open(blob, blob_size, off, smem_addr(s_adec) + warp * kRingBytes, lane);
    uint32_t cursor = static_cast<uint32_t>(off);
    // RansDecInit x 32 (rans_byte.h:109-122)
    uint32_t x = 0;
#pragma unroll
    for (int b = 3; b >= 0; b--) x = (x << 8) | lds_u8(win.ring | ((cursor + 4 * lane + b) & (kRingBytes - 1)));
    cursor += kHeaderBytes;

    const uint32_t lt = lanemask_lt();
    const uint32_t tab_lane = smem_addr_pinned(s_tab) + (ALIAS ? (lane & (kAliasDecReplicas - 1)) * 16 : 0);
    uint8_t* o = out + first + lane;
    const uint32_t steps = m >> 5, rem = m & 31;
    uint32_t g = 0;
    for (; g + 8 <= steps; g += 8) {
        win.top_up(cursor, lane);
        alias_dec_step<ALIAS, true, SB>(x, cursor, tab_lane, win.ring, o, lt, sb, true);
        alias_dec_step<ALIAS, true, SB>(x, cursor, tab_lane, win.ring, o + 32, lt, sb, true);
        alias_dec_step<ALIAS, true, SB>(x, cursor, tab_lane, win.ring, o + 64, lt, sb, true);
        alias_dec_step<ALIAS, true, SB>(x, cursor, tab_lane, win.ring, o + 96, lt, sb, true);
        win.top_up(cursor, lane);
        alias_dec_step<ALIAS, true, SB>(x, cursor, tab_lane, win.ring, o + 128, lt, sb, true);
        alias_dec_step<ALIAS, true, SB>(x, cursor, tab_lane, win.ring, o + 160, lt, sb, true);
        alias_dec_step<ALIAS, true, SB>(x, cursor, tab_lane, win.ring, o + 192, lt, sb, true);
        alias_dec_step<ALIAS, true, SB>(x, cursor, tab_lane, win.ring, o + 224, lt, sb, true);
        o += 256;
    }
    if (g + 4 <= steps) {
        win.top_up(cursor, lane);
        alias_dec_step<ALIAS, true, SB>(x, cursor, tab_lane, win.ring, o, lt, sb, true);
        alias_dec_step<ALIAS, true, SB>(x, cursor, tab_lane, win.ring, o + 32, lt, sb, true);
        alias_dec_step<ALIAS, true, SB>(x, cursor, tab_lane, win.ring, o + 64, lt, sb, true);
        alias_dec_step<ALIAS, true, SB>(x, cursor, tab_lane, win.ring, o + 96, lt, sb, true);
        o += 128;
        g += 4;
    }
    win.top_up(cursor, lane);
    for (; g < steps; g++) {
        alias_dec_step<ALIAS, true, SB>(x, cursor, tab_lane, win.ring, o, lt, sb, true);
        o += 32;
    }
    if (rem) alias_dec_step<ALIAS, false, SB>(x, cursor, tab_lane, win.ring, o, lt, sb, lane < rem);

    const bool bad = (cursor != static_cast<uint32_t>(end)) || (x != kByteL);
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(status, kStatStream);
}
constexpr uint32_t kAliasDecSmem = kAliasDecWarps * kRingBytes + 256 * kAliasDecReplicas * 16;   // 48 KiB

// ---------------------------------------------------------------------------
// K3p: the alias decoder on the persistent plumbing of word_decode_tma.cuh -- 2 CTAs of 20 warps per SM stay resident
// and pull chunk ids from the context's atomic counter; the 32 KiB replicated bucket table is built once per CTA; the
// per-warp stream window is the cp.async ring with a mirror behind it, wrapped once per 8 steps, so the two ranked byte
// reads of RansDecRenorm need no address masking (round 1's kernel spends 5 of its ~20 ALU-pipe instructions per step
// on that, and the ALU pipe is what binds it: 85 % busy).  Experiments and ablations: profiles/r2_alias_decode_lab.md.
// ---------------------------------------------------------------------------
using AliasDecShip = DecPolicy<20, 2, 8, kRefillCpAsync, 9, 0, 0, 0, false, 256 * kAliasDecReplicas * 16>;

// Experiment switches of tools/decode_lab.cu --coder alias (the shipped kernel has all of them off).  The word decoder's
// ablation bits kAblNoSymbolStore / kAblNoRingRead / kAblNoRefill apply as they do there; two more are specific to this step:
constexpr int kAblAliasOneByte = 16;   // ABLATION: renormalise by at most one byte (no second vote / rank / load / merge)
// LEAN (not an ablation, bit-exact): the bucket entry is repacked when the CTA stages the table, so that the step needs
// two selects and no byte permute, and compares the bucket-local bits of x at the TOP of a word:
//   w0 = (own_count - 1) << (40 - SB) | own adjust,  w1 = own slot_freqs << 8 | sym,  w2 = other's,  w3 = other's adjust
// "own" <=> x << (40 - SB) <= w0 (the low bits of w0 cannot change the outcome: the left side has zeros there); an empty
// own part (own_count 0 wraps to all ones: always taken) carries the other slot's data.  Needs a compile-time SB > 8.
template <uint32_t SB>
__device__ __forceinline__ uint4 alias_lean_entry(const AliasDecEntry& e, uint32_t bucket)
{
    constexpr uint32_t LB = SB - 8, K = 32 - LB;
    const uint32_t own_count = e.divider - (bucket << LB);                     // 0 .. 1 << LB (model_host.cpp: divider = lo + own)
    const bool none = own_count == 0;
    const uint32_t adj_own = none ? (e.adjust & 0xffffu) : (e.adjust >> 16);
    return make_uint4(((own_count - 1u) << K) | adj_own, none ? e.alt0 : e.alt1, e.alt0, e.adjust & 0xffffu);
}

template <uint32_t SB, int ABL = 0, int LEAN = 0>
__device__ __forceinline__ void alias_dec_step_p(uint32_t& x, uint32_t& cur, uint32_t tab_lane, uint8_t* o, uint32_t lt, uint32_t sb_rt,
                                                 bool active)
{
    const uint32_t sb = SB ? SB : sb_rt;
    if (active) {
        if (LEAN && SB > 8) {
            constexpr uint32_t LB = SB > 8 ? SB - 8 : 1;
            // bucket id: byte 1 of x by one PRMT at scale_bits 16, otherwise masked in place and scaled by the IMAD
            const uint32_t addr = LB == 8 ? mad_u32(__byte_perm(x, 0u, 0x4441u), kAliasDecReplicas * 16, tab_lane)
                                          : mad_u32(x & (0xffu << LB), (kAliasDecReplicas * 16) >> (LB & 7), tab_lane);
            const uint4 e = lds_u128_ro(addr);
            const bool own = mad_u32(x, 1u << (32 - LB), 0u) <= e.x;           // main_alias.cpp:261 on the bucket-local bits
            const uint32_t fs = own ? e.y : e.z;
            const uint32_t adj = own ? e.x : e.w;
            x = mad_u32(fs >> 8, x >> SB, (x - adj) & ((1u << SB) - 1));       // :265
            if (!(ABL & kAblNoSymbolStore)) *o = static_cast<uint8_t>(fs);     // :266
        } else {
            const uint32_t xm = x & ((1u << sb) - 1);                              // main_alias.cpp:258
            const uint32_t bucket = xm >> (sb - 8);                                // :259
            const uint4 e = lds_u128_ro(mad_u32(bucket, kAliasDecReplicas * 16, tab_lane));
            const bool own = xm < e.x;                                             // :261 (bucket2 = 2 * bucket + 1)
            const uint32_t fs = own ? e.z : e.y;                                   // slot_freqs << 8 | sym_id
            const uint32_t adj = __byte_perm(e.w, 0u, own ? 0x4432u : 0x4410u);    // the taken half of the packed adjusts
            x = mad_u32(fs >> 8, x >> sb, (xm - adj) & 0xffffu);                   // :265
            if (!(ABL & kAblNoSymbolStore)) *o = static_cast<uint8_t>(fs);         // :266
        }
    }
    // RansDecRenorm (rans_byte.h:307-318) for the warp: 0, 1 or 2 bytes per lane, lane k's bytes before lane k+1's,
    // most significant first.  Two predicates feed the votes, the ranked byte loads and the merges.
    const uint32_t xr = active ? x : kByteL;
    uint32_t xo = x;
    if (ABL & kAblAliasOneByte) {
        asm volatile(
            "{\n\t"
            ".reg .pred p1;\n\t"
            ".reg .b32 m, r, a, b0;\n\t"
            "setp.lt.u32 p1, %2, %4;\n\t"
            "vote.sync.ballot.b32 m, p1, 0xffffffff;\n\t"
            "and.b32 r, m, %3;\n\t"
            "popc.b32 r, r;\n\t"
            "add.u32 a, %1, r;\n\t"
            "popc.b32 m, m;\n\t"
            "add.u32 %1, %1, m;\n\t"
            "ld.shared.u8 b0, [a];\n\t"
            "@p1 mad.lo.u32 %0, %0, 256, b0;\n\t"
            "}"
            : "+r"(xo), "+r"(cur)
            : "r"(xr), "r"(lt), "n"(kByteL));
    } else if (ABL & kAblNoRingRead) {
        asm volatile(
            "{\n\t"
            ".reg .pred p1, p2;\n\t"
            ".reg .b32 m, r, a, b0;\n\t"
            "setp.lt.u32 p1, %2, %4;\n\t"
            "setp.lt.u32 p2, %2, %5;\n\t"
            "vote.sync.ballot.b32 m, p1, 0xffffffff;\n\t"
            "and.b32 r, m, %3;\n\t"
            "popc.b32 r, r;\n\t"
            "add.u32 a, %1, r;\n\t"
            "popc.b32 m, m;\n\t"
            "add.u32 %1, %1, m;\n\t"
            "vote.sync.ballot.b32 m, p2, 0xffffffff;\n\t"
            "and.b32 r, m, %3;\n\t"
            "popc.b32 r, r;\n\t"
            "add.u32 a, a, r;\n\t"
            "popc.b32 m, m;\n\t"
            "add.u32 %1, %1, m;\n\t"
            "and.b32 b0, a, 255;\n\t"
            "@p1 mad.lo.u32 %0, %0, 256, b0;\n\t"
            "@p2 mad.lo.u32 %0, %0, 256, b0;\n\t"
            "}"
            : "+r"(xo), "+r"(cur)
            : "r"(xr), "r"(lt), "n"(kByteL), "n"(kByteL >> 8));
    } else {
        asm volatile(
            "{\n\t"
            ".reg .pred p1, p2;\n\t"
            ".reg .b32 m, r, a, b0, b1;\n\t"
            "setp.lt.u32 p1, %2, %4;\n\t"
            "setp.lt.u32 p2, %2, %5;\n\t"
            "vote.sync.ballot.b32 m, p1, 0xffffffff;\n\t"
            "and.b32 r, m, %3;\n\t"
            "popc.b32 r, r;\n\t"
            "add.u32 a, %1, r;\n\t"
            "popc.b32 m, m;\n\t"
            "add.u32 %1, %1, m;\n\t"
            "vote.sync.ballot.b32 m, p2, 0xffffffff;\n\t"
            "and.b32 r, m, %3;\n\t"
            "popc.b32 r, r;\n\t"
            "add.u32 a, a, r;\n\t"
            "popc.b32 m, m;\n\t"
            "add.u32 %1, %1, m;\n\t"
            "ld.shared.u8 b0, [a];\n\t"              // unpredicated: ring + mirror make both reads safe for every lane
            "ld.shared.u8 b1, [a+1];\n\t"
            "@p1 mad.lo.u32 %0, %0, 256, b0;\n\t"
            "@p2 mad.lo.u32 %0, %0, 256, b1;\n\t"
            "}"
            : "+r"(xo), "+r"(cur)
            : "r"(xr), "r"(lt), "n"(kByteL), "n"(kByteL >> 8));
    }
    x = xo;
}

// kGroup steps: wrap (the mirror absorbed the previous group's overrun), fill check, steps
template <uint32_t SB, class P, int LEAN>
__device__ __forceinline__ void alias_dec_group(TmaWindow& win, uint32_t& x, uint32_t ring_end, uint32_t tab_lane, uint8_t* og, uint32_t lt,
                                                uint32_t lane, uint32_t sb_rt, uint32_t* __restrict__ status)
{
    if (win.cur >= ring_end) {
        win.cur -= P::kRing;
        win.limit -= P::kRing;
    }
    if (!(P::kAblate & kAblNoRefill) && __any_sync(0xffffffffu, win.cur > win.limit)) tma_advance<P>(win, lane, status);
#pragma unroll
    for (int j = 0; j < P::kGroup; j++) alias_dec_step_p<SB, P::kAblate, LEAN>(x, win.cur, tab_lane, og + 32 * j, lt, sb_rt, true);
}

// P, LEAN: the shipped configuration unless tools/decode_lab.cu instantiates an experiment
template <uint32_t SB, class P = AliasDecShip, int LEAN = 0>
__global__ void __launch_bounds__(P::kWarps * 32, P::kMinBlocks)
alias_decode_persist_kernel(const uint8_t* __restrict__ blob, uint64_t blob_size, const uint64_t* __restrict__ offsets, uint32_t sb_rt,
                            const AliasDecEntry* __restrict__ g_dec, uint8_t* __restrict__ out, uint64_t n, uint32_t chunk_syms,
                            uint32_t n_chunks, DecodeWork* __restrict__ work, uint32_t* __restrict__ status)
{
    extern __shared__ __align__(1024) uint8_t s_adec[];       // [32 KiB table][16 B][kWarps x (ring, mirror)]
    uint4* s_tab = reinterpret_cast<uint4*>(s_adec);
    for (uint32_t i = threadIdx.x; i < 256 * kAliasDecReplicas; i += blockDim.x) {
        const AliasDecEntry e = g_dec[i / kAliasDecReplicas];
        if (LEAN && SB > 8) s_tab[i] = alias_lean_entry<(SB > 8 ? SB : 9)>(e, i / kAliasDecReplicas);
        else s_tab[i] = make_uint4(e.divider, e.alt0, e.alt1, e.adjust);
    }
    __syncthreads();

    const uint32_t base = smem_addr_pinned(s_adec);
    const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t tab_lane = base + (lane & (kAliasDecReplicas - 1)) * 16;
    const uint32_t lt = lanemask_lt();
    TmaWindow win;
    win.ring = base + P::kWarpsOff + warp * P::kWarpStride;
    win.seq_ready = kTmaSeqBase;

    for (;;) {
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(&work->next_chunk, 1u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (chunk >= n_chunks) break;
        const uint64_t first = static_cast<uint64_t>(chunk) * chunk_syms;
        const uint64_t left = n - first;
        const uint32_t m = left < chunk_syms ? static_cast<uint32_t>(left) : chunk_syms;
        const uint64_t off = offsets[chunk];
        const uint64_t end = offsets[chunk + 1] & ~static_cast<uint64_t>(15);
        const bool dir_bad = off + kHeaderBytes > end || end > blob_size || end - off > (1u << 30);   // the directory is not trusted
        if (__shfl_sync(0xffffffffu, dir_bad ? 1u : 0u, 0)) {
            if (lane == 0) atomicOr(status, kStatStream);
            continue;
        }
        const uint32_t off_in = __shfl_sync(0xffffffffu, static_cast<uint32_t>(off % P::kUnit), 0);
        const uint32_t len = __shfl_sync(0xffffffffu, static_cast<uint32_t>(end - off), 0);
        const uint32_t seq0 = win.seq_ready;
        const uint32_t pos0 = seq0 * P::kUnit + off_in;
        win.end_pos = pos0 + len;
        win.src = reinterpret_cast<uint64_t>(blob) + (off - off_in) - static_cast<uint64_t>(seq0) * P::kUnit + lane * 16;
        win.cur = win.ring + (pos0 & (P::kRing - 1));
        win.limit = win.cur - off_in - P::kNeed;
        if (!(P::kAblate & kAblNoRefill)) {
            tma_issue<P>(win, seq0, lane);
            asm volatile("cp.async.commit_group;" ::: "memory");
            while (win.cur > win.limit) tma_advance<P>(win, lane, status);
        }
        __syncwarp();

        // RansDecInit x 32 (rans_byte.h:109-122): lane k's state is the k-th little-endian u32; the mirror covers a wrap
        uint32_t x = 0;
#pragma unroll
        for (int b = 3; b >= 0; b--) x = (x << 8) | lds_u8(win.cur + 4 * lane + b);
        win.cur += kHeaderBytes;

        uint8_t* o = out + first + lane;
        const uint32_t ring_end = win.ring + P::kRing;
        uint32_t todo = m >> 5;
        for (; todo >= P::kGroup; todo -= P::kGroup) {
            alias_dec_group<SB, P, LEAN>(win, x, ring_end, tab_lane, o, lt, lane, sb_rt, status);
            o += 32 * P::kGroup;
        }
        if (win.cur >= ring_end) {
            win.cur -= P::kRing;
            win.limit -= P::kRing;
        }
        if (!(P::kAblate & kAblNoRefill) && __any_sync(0xffffffffu, win.cur > win.limit)) tma_advance<P>(win, lane, status);
        for (; todo; todo--) {
            alias_dec_step_p<SB, P::kAblate, LEAN>(x, win.cur, tab_lane, o, lt, sb_rt, true);
            o += 32;
        }
        if (m & 31) alias_dec_step_p<SB, P::kAblate, LEAN>(x, win.cur, tab_lane, o, lt, sb_rt, lane < (m & 31));   // main_alias.cpp:399-404

        const uint32_t pos = win.cur + (win.seq_ready * P::kUnit - P::kNeed - win.limit);
        const bool bad = (pos != pos0 + len) || (x != kByteL);
        if (!P::kAblate && __any_sync(0xffffffffu, bad) && lane == 0) atomicOr(status, kStatStream);
        if (!(P::kAblate & kAblNoRefill)) {
            const uint32_t seq_end = (win.end_pos + P::kUnit - 1) / P::kUnit;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            win.seq_ready = win.seq_ready < seq_end ? win.seq_ready + 1 : seq_end;
        }
        __syncwarp();
    }
    if (lane == 0) {
        const uint32_t total = gridDim.x * P::kWarps;
        if (atomicAdd(&work->warps_done, 1u) == total - 1) {
            work->next_chunk = 0;
            work->warps_done = 0;
        }
    }
}


// ---------------------------------------------------------------------------
// K4: 32-way alias encode.  RansEncPutAlias (main_alias.cpp:241-250) = RansEncRenorm
// (rans_byte.h:62-74) + divide + alias_remap gather.
//
// Persistent kernel, one 32-warp CTA per SM, because the encoder's slot permutation
// alias_remap (u16, 2 << scale_bits bytes = 128 KiB at scale_bits 16) lives in
// shared memory next to the 8x replicated per-symbol table ({magic, freq, cum, shift}, one
// conflict-free LDS.128 per step) and the per-warp 512 B symbol stage + 512 B output ring
// (same staged, vectorised I/O as the word encoder).
// ---------------------------------------------------------------------------
constexpr int kAliasEncWarps = 32;
constexpr uint32_t kAliasEncFixedSmem = kAliasEncWarps * kEncWarpSmem + kEncTableBytes;   // 64 KiB + remap

__device__ __forceinline__ uint32_t lds_u16_ro(uint32_t addr)
{
    uint16_t v;
    asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}

struct AliasEncState {
    uint32_t x;        // rANS state
    uint32_t wpos;     // un-wrapped ring position of the next byte = 511 - bytes_emitted
    uint32_t flags;
};

// One encode step of the byte-renormalising coders for the whole warp.
//   ALIAS:  RansEncPutAlias (main_alias.cpp:241-250); table {magic, freq, cum, shift}
//   !ALIAS: RansEncPutSymbol (rans_byte.h:258-280);   table {x_max, rcp_freq, bias, cmpl_freq | rcp_shift << 16}
//           = RansEncSymbol (rans_byte.h:159-165), bit 31 of the last word marks a symbol outside the model
// FULL: all 32 lanes take part (active == true at compile time); only the ragged first step of a chunk does not.
template <bool ALIAS, bool FULL, uint32_t SB>
__device__ __forceinline__ void alias_enc_step(AliasEncState& st, uint32_t sym, uint32_t tab_lane, uint32_t remap, uint32_t ring,
                                               uint32_t gt, uint32_t sb_rt, bool active)
{
    const uint32_t sb = SB ? SB : sb_rt;
    bool n1 = false, n2 = false;
    uint4 e = make_uint4(0, 1, 0, 0);
    uint32_t x_max = 0;
    if (active) {
        e = lds_u128_ro(tab_lane + sym * (kEncReplicas * 16));
        st.flags |= e.w;
        x_max = ALIAS ? e.y << (31 - sb) : e.x;                   // ((L >> sb) << 8) * freq, rans_byte.h:64 / :197
    }
    if (FULL) {
        // the renormalisation below as one PTX sequence: two predicates feed the votes, the ranked stores and
        // the shifts (the compiler otherwise re-evaluates the comparisons and branches around the stores)
        asm volatile(
            "{\n\t"
            ".reg .pred p1, p2;\n\t"
            ".reg .b32 m1, m2, r1, r2, a, b, xs;\n\t"
            "shr.u32 xs, %0, 8;\n\t"
            "setp.ge.u32 p1, %0, %2;\n\t"
            "setp.ge.u32 p2, xs, %2;\n\t"
            "vote.sync.ballot.b32 m1, p1, 0xffffffff;\n\t"
            "vote.sync.ballot.b32 m2, p2, 0xffffffff;\n\t"
            "and.b32 r1, m1, %3;\n\t"
            "and.b32 r2, m2, %3;\n\t"
            "popc.b32 r1, r1;\n\t"
            "popc.b32 r2, r2;\n\t"
            "sub.u32 a, %1, r1;\n\t"
            "sub.u32 a, a, r2;\n\t"
            "sub.u32 b, a, 1;\n\t"
            "and.b32 a, a, %4;\n\t"
            "and.b32 b, b, %4;\n\t"
            "or.b32 a, a, %5;\n\t"
            "or.b32 b, b, %5;\n\t"
            "@p1 st.shared.u8 [a], %0;\n\t"
            "@p2 st.shared.u8 [b], xs;\n\t"
            "@p1 mov.b32 %0, xs;\n\t"
            "@p2 shr.u32 %0, %0, 8;\n\t"
            "popc.b32 m1, m1;\n\t"
            "popc.b32 m2, m2;\n\t"
            "sub.u32 %1, %1, m1;\n\t"
            "sub.u32 %1, %1, m2;\n\t"
            "}"
            : "+r"(st.x), "+r"(st.wpos)
            : "r"(x_max), "r"(gt), "n"(kEncRingBytes - 1), "r"(ring)
            : "memory");
    } else {
        if (active) {
            n1 = st.x >= x_max;                                   // :65 / :265
            n2 = (st.x >> 8) >= x_max;                            // second trip of the do/while, :67-70
        }
        const uint32_t m1 = __ballot_sync(0xffffffffu, n1);
        const uint32_t m2 = __ballot_sync(0xffffffffu, n2);
        if (n1) {
            // lanes are visited 31..0 (main_alias.cpp:365-370 generalised); each writes downwards
            const uint32_t pos = st.wpos - __popc(m1 & gt) - __popc(m2 & gt);
            sts_u8(ring | (pos & (kEncRingBytes - 1)), st.x);                          // :68
            if (n2) sts_u8(ring | ((pos - 1) & (kEncRingBytes - 1)), st.x >> 8);
            st.x >>= n2 ? 16 : 8;                                                      // :69
        }
        st.wpos -= __popc(m1) + __popc(m2);
    }
    if (active) {
        if (ALIAS) {
            // exact x / freq with the 33-bit round-up reciprocal 2^32 + magic, as in the word encoder; here
            // x < 2^31 (rans_byte.h:26-28), so x + mulhi(x, magic) < 2^32 and there is no carry to fold in
            const uint32_t q = funnel_shr_wrap(st.x + __umulhi(st.x, e.x), 0u, e.w);
            const uint32_t r = st.x - q * e.y;                                     // x % freq
            st.x = (q << sb) + lds_u16_ro(remap + 2u * (r + e.z));                 // main_alias.cpp:249
        } else {
            const uint32_t q = __umulhi(st.x, e.y) >> ((e.w >> 16) & 31u);         // rans_byte.h:278
            st.x = st.x + e.z + q * (e.w & 0xffffu);                               // :279
        }
    }
}

// Encode m symbols (one warp) as one 32-way byte-renormalised stream ending at slot_end (16-byte aligned);
// returns the stream size in bytes (warp-uniform).
template <bool ALIAS, uint32_t SB>
__device__ __forceinline__ uint32_t alias_encode_stream(const uint8_t* __restrict__ chunk_in, uint32_t m, uint32_t tab, uint32_t remap,
                                                        uint32_t wsm, uint32_t sb, uint8_t* __restrict__ slot_end,
                                                        uint32_t* __restrict__ status)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t stage = wsm, ring = wsm + kEncStageBytes;
    const uint32_t tab_lane = tab + (lane & (kEncReplicas - 1)) * 16;
    const uint32_t gt = lanemask_gt();

    AliasEncState st;
    st.x = kByteL;                              // RansEncInit, rans_byte.h:56-59
    st.wpos = kEncRingBytes - 1;
    st.flags = 0;
    uint32_t flushed = 0;
    const uint32_t steps = m >> 5, rem = m & 31;
    const uint32_t nblk = steps >> 4;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(chunk_in) & 15) == 0;

    uint4 parked = make_uint4(0, 0, 0, 0);
    if (nblk && vec_ok) parked = ldg_stream_u128(reinterpret_cast<const uint4*>(chunk_in + (nblk - 1) * kEncStageBytes) + lane);
    if (rem) {
        const bool active = lane < rem;
        const uint32_t s = active ? chunk_in[static_cast<uint64_t>(steps) * 32 + lane] : 0;
        alias_enc_step<ALIAS, false, SB>(st, s, tab_lane, remap, ring, gt, sb, active);
    }
    for (uint32_t g = steps; g > nblk * 16; g--) {
        const uint32_t s = chunk_in[static_cast<uint64_t>(g - 1) * 32 + lane];
        alias_enc_step<ALIAS, true, SB>(st, s, tab_lane, remap, ring, gt, sb, true);
        if (((g - 1) & 3) == 0) word_enc_flush(kEncRingBytes - 1 - st.wpos, flushed, ring, slot_end, lane);
    }
    word_enc_flush(kEncRingBytes - 1 - st.wpos, flushed, ring, slot_end, lane);

    for (uint32_t b = nblk; b-- > 0;) {
        __syncwarp();
        if (vec_ok) {
            sts_u128(stage + lane * 16, parked);
            if (b) parked = ldg_stream_u128(reinterpret_cast<const uint4*>(chunk_in + (b - 1) * kEncStageBytes) + lane);
        } else {
            const uint8_t* p = chunk_in + b * kEncStageBytes + lane;
#pragma unroll
            for (int j = 0; j < 16; j++) sts_u8(stage + j * 32 + lane, p[j * 32]);
        }
        __syncwarp();
#pragma unroll
        for (int grp = 3; grp >= 0; grp--) {
#pragma unroll
            for (int j = 3; j >= 0; j--)
                alias_enc_step<ALIAS, true, SB>(st, lds_u8(stage + (grp * 4 + j) * 32 + lane), tab_lane, remap, ring, gt, sb, true);
            if (kEncRingBytes - 1 - st.wpos - flushed >= 256) word_enc_flush(kEncRingBytes - 1 - st.wpos, flushed, ring, slot_end, lane);
        }
    }

    // RansEncFlush for lanes 31..0 (rans_byte.h:93-105): lane 31's most significant byte is the first byte
    // written (highest address); lane 0's least significant byte ends up first in memory
    word_enc_flush(kEncRingBytes - 1 - st.wpos, flushed, ring, slot_end, lane);     // < 16 bytes stay pending
    const uint32_t hpos = st.wpos - 4u * (31 - lane);
#pragma unroll
    for (int j = 0; j < 4; j++) sts_u8(ring | ((hpos - (3 - j)) & (kEncRingBytes - 1)), st.x >> (8 * j));
    st.wpos -= kHeaderBytes;
    const uint32_t produced = kEncRingBytes - 1 - st.wpos;
    word_enc_flush(produced, flushed, ring, slot_end, lane);
    const uint32_t left = produced - flushed;                        // < 16 bytes: head of the stream
    if (lane < left) {
        const uint32_t off = flushed + lane + 1;
        *(slot_end - off) = static_cast<uint8_t>(lds_u8(ring | ((0u - off) & (kEncRingBytes - 1))));
    }
    if (__any_sync(0xffffffffu, (st.flags & kEncBadSymbol) != 0) && lane == 0) atomicOr(status, kStatSymbol);
    return produced;
}

// !ALIAS: g_enc holds RansEncSymbol images {x_max, rcp_freq, bias, cmpl | shift << 16}, g_remap is unused
// Two modes (like the word encoder):
//   look == nullptr: chunks strided over the persistent CTAs, each into its own worst-case slot of `scratch`,
//                    sizes[] written for the tile scan + compaction that follow;
//   look != nullptr: fused -- chunk ids from an atomic counter, two scratch slots per warp, warp 0 of CTA 0 is the
//                    scanner, every worker places chunk k after encoding chunk k+1 (see word_kernels.cuh, K2f).
template <bool ALIAS, uint32_t SB>
__global__ void __launch_bounds__(kAliasEncWarps * 32, 1)
alias_encode_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t chunk_syms, uint32_t n_chunks, uint32_t sb_rt,
                    const AliasEncEntry* __restrict__ g_enc, const uint16_t* __restrict__ g_remap,
                    uint8_t* __restrict__ scratch, uint32_t slot_bytes, uint32_t* __restrict__ sizes,
                    uint64_t* __restrict__ look, uint32_t* __restrict__ counter, uint8_t* __restrict__ blob, uint64_t blob_cap,
                    uint64_t* __restrict__ offsets, uint32_t* __restrict__ status)
{
    const uint32_t sb = SB ? SB : sb_rt;
    extern __shared__ __align__(1024) uint8_t s_alias[];      // [32 x 1.0 KiB stage+ring][32 KiB table][remap]
    uint4* s_tab = reinterpret_cast<uint4*>(s_alias + kAliasEncWarps * kEncWarpSmem);
    uint4* s_remap = reinterpret_cast<uint4*>(s_alias + kAliasEncFixedSmem);
    for (uint32_t i = threadIdx.x; i < 256 * kEncReplicas; i += blockDim.x) {
        const AliasEncEntry e = g_enc[i / kEncReplicas];
        // a symbol the model does not contain: keep the step well defined (freq 1), flag it
        if (ALIAS) s_tab[i] = (e.shift & kEncBadSymbol) ? make_uint4(0, 1, 0, kEncBadSymbol) : make_uint4(e.magic, e.freq, e.cum, e.shift);
        else       s_tab[i] = make_uint4(e.magic, e.freq, e.cum, e.shift);          // the host already made bad entries safe
    }
    if (ALIAS) {
        const uint32_t remap_vecs = (2u << sb) / 16;
        for (uint32_t i = threadIdx.x; i < remap_vecs; i += blockDim.x) s_remap[i] = reinterpret_cast<const uint4*>(g_remap)[i];
    }
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t base = smem_addr(s_alias);
    const uint32_t tab = base + kAliasEncWarps * kEncWarpSmem, remap = base + kAliasEncFixedSmem, wsm = base + warp * kEncWarpSmem;
    if (!look) {
        for (uint32_t chunk = blockIdx.x * kAliasEncWarps + warp; chunk < n_chunks; chunk += gridDim.x * kAliasEncWarps) {
            const uint64_t first = static_cast<uint64_t>(chunk) * chunk_syms;
            const uint32_t m = static_cast<uint32_t>(min(static_cast<uint64_t>(chunk_syms), n - first));
            const uint32_t produced = alias_encode_stream<ALIAS, SB>(in + first, m, tab, remap, wsm, sb,
                                                                 scratch + static_cast<uint64_t>(chunk + 1) * slot_bytes, status);
            if (lane == 0) sizes[chunk] = produced;
        }
        return;
    }
    if (blockIdx.x == 0 && warp == 0) {
        fused_scanner(look, n_chunks, lane, status);
        return;
    }
    uint8_t* slots = scratch + (static_cast<uint64_t>(blockIdx.x) * kAliasEncWarps + warp) * 2 * slot_bytes;
    uint32_t pend_chunk = 0, pend_size = 0, parity = 0;
    bool pending = false;
    for (;;) {
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(counter, 1u);
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (chunk >= n_chunks) break;
        const uint64_t first = static_cast<uint64_t>(chunk) * chunk_syms;
        const uint32_t m = static_cast<uint32_t>(min(static_cast<uint64_t>(chunk_syms), n - first));
        const uint32_t produced = alias_encode_stream<ALIAS, SB>(in + first, m, tab, remap, wsm, sb,
                                                             slots + (parity + 1) * static_cast<uint64_t>(slot_bytes), status);
        if (lane == 0) st_relaxed_u64(look + chunk, kLookAgg | ((produced + 15u) & ~15u));
        __syncwarp();
        if (pending)
            fused_place(look, pend_chunk, n_chunks, slots + (2 - parity) * static_cast<uint64_t>(slot_bytes), pend_size, blob, blob_cap,
                        offsets, lane, status);
        pend_chunk = chunk;
        pend_size = produced;
        pending = true;
        parity ^= 1;
        __syncwarp();
    }
    if (pending)
        fused_place(look, pend_chunk, n_chunks, slots + (2 - parity) * static_cast<uint64_t>(slot_bytes), pend_size, blob, blob_cap,
                    offsets, lane, status);
}

// scale_bits the kernels are specialised for (main_alias.cpp:276 uses 16, main.cpp:136 / main64.cpp:136 use 14, the word
// coder's 12); every other value runs the generic instantiation (SB = 0).
template <bool ALIAS, uint32_t SB>
inline void configure_alias_pair()
{
    const uint32_t dec_smem = ALIAS ? kAliasDecSmem : kAliasDecWarps * kRingBytes + 1024 + (1u << 16);
    cudaFuncSetAttribute(alias_decode_kernel<ALIAS, SB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(alias_decode_kernel<ALIAS, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize, dec_smem);
    cudaFuncSetAttribute(alias_encode_kernel<ALIAS, SB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(alias_encode_kernel<ALIAS, SB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         kAliasEncFixedSmem + (ALIAS ? (2u << 16) : 0u));
}
// Build switch: 1 ships the repacked ("lean") bucket entry of alias_dec_step_p for the compile-time scale_bits
// (A/B in profiles/r2_alias_decode_lab.md); a runtime scale_bits always takes the plain entry.
#ifndef RB200_ALIAS_LEAN
#define RB200_ALIAS_LEAN 1
#endif
constexpr int kAliasLean = RB200_ALIAS_LEAN;
template <uint32_t SB>
inline void configure_alias_persist()
{
    auto k = alias_decode_persist_kernel<SB, AliasDecShip, kAliasLean>;
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, AliasDecShip::kSmemBytes);
}
inline void configure_alias_kernels()
{
    configure_alias_persist<0>(); configure_alias_persist<16>(); configure_alias_persist<14>(); configure_alias_persist<12>();
    configure_alias_pair<true, 0>(); configure_alias_pair<true, 16>(); configure_alias_pair<true, 14>(); configure_alias_pair<true, 12>();
    configure_alias_pair<false, 0>(); configure_alias_pair<false, 16>(); configure_alias_pair<false, 14>(); configure_alias_pair<false, 12>();
}

#define RB200_ALIAS_SB_DISPATCH(sb, CALL) \
    switch (sb) {                        \
    case 16: { constexpr uint32_t SB = 16; CALL; } break; \
    case 14: { constexpr uint32_t SB = 14; CALL; } break; \
    case 12: { constexpr uint32_t SB = 12; CALL; } break; \
    default: { constexpr uint32_t SB = 0; CALL; } break;  \
    }

// remap == nullptr selects the rans_byte cum2sym coder.  look == nullptr: split mode (sizes[] out, one slot per
// chunk); otherwise fused mode (two slots per resident warp; look/counter zeroed by the caller).
// sms = multiprocessors of the context's device (rb200_ctx keeps it; nothing process-wide is cached here).
inline int launch_alias_encode(cudaStream_t stream, uint32_t sms, const uint8_t* d_in, uint64_t n, uint32_t chunk_syms, uint32_t n_chunks,
                               uint32_t sb, const AliasEncEntry* enc, const uint16_t* remap, uint8_t* scratch, uint32_t slot,
                               uint32_t* sizes, uint64_t* look, uint32_t* counter, uint8_t* blob, uint64_t blob_cap, uint64_t* offsets,
                               uint32_t* status)
{
    uint32_t grid = (n_chunks + (look ? 1 : 0) + kAliasEncWarps - 1) / kAliasEncWarps;
    if (grid > sms) grid = sms;                       // persistent: one CTA per SM
    if (remap) {
        RB200_ALIAS_SB_DISPATCH(sb, (alias_encode_kernel<true, SB><<<grid, kAliasEncWarps * 32, kAliasEncFixedSmem + (2u << sb), stream>>>(
            d_in, n, chunk_syms, n_chunks, sb, enc, remap, scratch, slot, sizes, look, counter, blob, blob_cap, offsets, status)))
    } else {
        RB200_ALIAS_SB_DISPATCH(sb, (alias_encode_kernel<false, SB><<<grid, kAliasEncWarps * 32, kAliasEncFixedSmem, stream>>>(
            d_in, n, chunk_syms, n_chunks, sb, enc, nullptr, scratch, slot, sizes, look, counter, blob, blob_cap, offsets, status)))
    }
    return 0;
}
inline uint32_t alias_fused_slots(uint32_t sms) { return sms * kAliasEncWarps * 2; }

// work != nullptr: the persistent kernel (sms x 2 CTAs of 20 warps, chunk ids from the context's work counter)
inline int launch_alias_decode(cudaStream_t stream, const uint8_t* blob, uint64_t blob_size, const uint64_t* offsets, uint32_t sb,
                               const AliasDecEntry* dec, uint8_t* out, uint64_t n, uint32_t chunk_syms, uint32_t n_chunks,
                               uint32_t* status, DecodeWork* work = nullptr, uint32_t sms = 0)
{
    if (work) {
        using P = AliasDecShip;
        const uint32_t want = (n_chunks + P::kWarps - 1) / P::kWarps, full = sms * P::kMinBlocks;
        const uint32_t pgrid = want < full ? want : full;
        RB200_ALIAS_SB_DISPATCH(sb, (alias_decode_persist_kernel<SB, P, kAliasLean><<<pgrid, P::kWarps * 32, P::kSmemBytes, stream>>>(
            blob, blob_size, offsets, sb, dec, out, n, chunk_syms, n_chunks, work, status)))
        return 0;
    }
    const uint32_t grid = (n_chunks + kAliasDecWarps - 1) / kAliasDecWarps;
    RB200_ALIAS_SB_DISPATCH(sb, (alias_decode_kernel<true, SB><<<grid, kAliasDecWarps * 32, kAliasDecSmem, stream>>>(
        blob, blob_size, offsets, sb, dec, out, n, chunk_syms, n_chunks, status)))
    return 0;
}

// rans_byte cum2sym decode: table = cum2sym[1 << sb] followed by 256 x u32 {start | freq << 16}
inline int launch_byte_decode(cudaStream_t stream, const uint8_t* blob, uint64_t blob_size, const uint64_t* offsets, uint32_t sb,
                              const uint8_t* table, uint8_t* out, uint64_t n, uint32_t chunk_syms, uint32_t n_chunks, uint32_t* status)
{
    const uint32_t grid = (n_chunks + kAliasDecWarps - 1) / kAliasDecWarps;
    RB200_ALIAS_SB_DISPATCH(sb, (alias_decode_kernel<false, SB><<<grid, kAliasDecWarps * 32, kAliasDecWarps * kRingBytes + 1024 + (1u << sb), stream>>>(
        blob, blob_size, offsets, sb, table, out, n, chunk_syms, n_chunks, status)))
    return 0;
}

}  // namespace rb200
