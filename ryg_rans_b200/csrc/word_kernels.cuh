// word_kernels.cuh -- sm_100a kernels for the word coder (rans_word_sse41.h semantics):
// 32-bit state, 16-bit renormalisation, L = 1 << 16, scale_bits = 12.
//
// One warp = one chunk = one 32-way interleaved stream; lane k owns rANS state k.
// This is the warp-wide form of the reference's SSE4.1 design (README:27-30):
//   RansSimdDecSym    (rans_word_sse41.h:151-179)  -> one shared-memory gather + IMAD
//   RansSimdDecRenorm (rans_word_sse41.h:182-227)  -> __ballot_sync + popc(mask & lanemask_lt)
//                                                     replaces movemask + pshufb LUT
//   RansWordEncPut    (rans_word_sse41.h:81-93)    -> ballot + popc(mask & lanemask_gt),
//                                                     exact division by reciprocal multiply
#pragma once
#include "device_utils.cuh"
#include "tables.h"

namespace rb200 {

constexpr uint32_t kWordL = 1u << 16;       // RANS_WORD_L, rans_word_sse41.h:35
constexpr uint32_t kWordScaleBits = 12;     // RANS_WORD_SCALE_BITS, :37
constexpr uint32_t kWordSlots = 1u << kWordScaleBits;

// ---------------------------------------------------------------------------
// Per-warp stream window: a 1 KiB shared-memory ring over the compressed stream,
// refilled 512 B at a time with one coalesced 128-bit load per lane.  The load for
// the next unit is issued one refill ahead and parked in registers, so its HBM
// latency is hidden behind >= 8 decode steps.  The cursor never leaves the ring,
// which also bounds every read a corrupt stream could cause.
// ---------------------------------------------------------------------------
constexpr uint32_t kRingBytes = 1024;
constexpr uint32_t kUnitBytes = 512;

struct StreamWindow {
    uint32_t ring;          // shared-space address of this warp's ring, 1 KiB aligned
    uint32_t fill_end;      // low 32 bits of the absolute blob byte offset filled so far
    uint32_t next_vec;      // this lane's 16-byte vector of the next unit to fetch (index into the blob)
    uint32_t limit_vec;     // blob_size / 16 (the C-ABI rejects blobs of 64 GiB and more)
    const uint4* base;      // blob, warp-uniform
    uint4 parked;           // unit in flight

    // Past the end of the blob the last vector is fetched again instead: what lies beyond a stream's
    // own end is never consumed by a valid stream, and a corrupt one may read any in-bounds bytes.
    __device__ __forceinline__ uint4 fetch()
    {
        const uint4 v = ldg_stream_u128(base + min(next_vec, limit_vec - 1));
        next_vec += 32;
        return v;
    }

    // off = absolute byte offset of the stream start inside the blob
    __device__ __forceinline__ void open(const uint8_t* blob, uint64_t blob_size, uint64_t off, uint32_t ring_addr, uint32_t lane)
    {
        ring = ring_addr;
        base = reinterpret_cast<const uint4*>(blob);
        const uint32_t unit0 = static_cast<uint32_t>(off / kUnitBytes);
        next_vec = unit0 * 32 + lane;
        limit_vec = static_cast<uint32_t>(blob_size / 16);
        uint4 a = fetch();
        uint4 b = fetch();
        parked = fetch();
        sts_u128(ring + ((unit0 & 1) * kUnitBytes) + lane * 16, a);
        sts_u128(ring + (((unit0 + 1) & 1) * kUnitBytes) + lane * 16, b);
        fill_end = (unit0 + 2) * kUnitBytes;
        __syncwarp();
    }

    // Call before every group of steps that consumes at most 256 bytes.
    __device__ __forceinline__ void top_up(uint32_t cursor, uint32_t lane)
    {
        if (fill_end - cursor <= kUnitBytes) {          // warp-uniform
            __syncwarp();
            sts_u128(ring + (fill_end & kUnitBytes) + lane * 16, parked);
            fill_end += kUnitBytes;
            parked = fetch();
            __syncwarp();
        }
    }
};

// ---------------------------------------------------------------------------
// K1: 32-way word-coder decode
// ---------------------------------------------------------------------------
constexpr int kDecWarps = 8;

template <bool WIDE>
__device__ __forceinline__ uint32_t word_table_freq(uint32_t e)
{
    uint32_t f = e >> 20;
    if (WIDE) f = f ? f : kWordSlots;   // single-symbol model: freq 4096 stored as 0
    return f;
}

// one decode step for the whole warp: RansWordDecSym + RansWordDecRenorm
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// one decode step for the whole warp: RansWordDecSym + RansWordDecRenorm.
// tab = shared-space address of the packed table (kept live in a register by the caller).
template <bool WIDE>
__device__ __forceinline__ void word_dec_step(uint32_t& x, uint32_t& cursor, uint32_t tab, uint32_t ring,
                                              uint8_t* o, uint32_t lt, bool active)
{
    bool need = false;
    if (active) {
        const uint32_t e = lds_u32_ro(mad_u32(x & (kWordSlots - 1), 4u, tab));  // rans_word_sse41.h:126
        x = word_table_freq<WIDE>(e) * (x >> kWordScaleBits) + ((e >> 8) & 0xfffu);   // :129
        *o = static_cast<uint8_t>(e);                                       // :130
        need = x < kWordL;                                                  // :137
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, need);
    const uint32_t a = mad_u32(__popc(mask & lt), 2u, cursor);              // lane order within the step
    const uint32_t w = lds_u16(ring | (a & (kRingBytes - 1)));
    if (need) x = (x << 16) | w;                                            // :138
    cursor = mad_u32(__popc(mask), 2u, cursor);                             // :139
}

// Decode chunk `chunk` (one warp).  tab = shared address of the 4096-entry packed table,
// ring = shared address of this warp's 1 KiB stream window.
template <bool WIDE>
__device__ __forceinline__ void word_decode_chunk(const uint8_t* __restrict__ blob, uint64_t blob_size,
                                                  const uint64_t* __restrict__ offsets, uint32_t chunk, uint32_t tab,
                                                  uint32_t ring, uint8_t* __restrict__ chunk_out, uint32_t m,
                                                  uint32_t* __restrict__ status)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t off = offsets[chunk];
    const uint64_t end = offsets[chunk + 1] & ~static_cast<uint64_t>(15);
    if ((off & 1) || off + kHeaderBytes > end || end > blob_size) {      // directory is not trusted
        if (lane == 0) atomicOr(status, kStatStream);
        return;
    }

    StreamWindow win;
    win.open(blob, blob_size, off, ring, lane);
    uint32_t cursor = static_cast<uint32_t>(off);
    // RansWordDecInit x 32 (rans_word_sse41.h:109-120): lane k's state is the k-th u32
    uint32_t x = lds_u16(win.ring | ((cursor + 4 * lane) & (kRingBytes - 1)))
               | (lds_u16(win.ring | ((cursor + 4 * lane + 2) & (kRingBytes - 1))) << 16);
    cursor += kHeaderBytes;

    const uint32_t lt = lanemask_lt();
    uint8_t* o = chunk_out + lane;
    const uint32_t steps = m >> 5, rem = m & 31;
    uint32_t g = 0;
    for (; g + 8 <= steps; g += 8) {
        win.top_up(cursor, lane);
        word_dec_step<WIDE>(x, cursor, tab, win.ring, o, lt, true);
        word_dec_step<WIDE>(x, cursor, tab, win.ring, o + 32, lt, true);
        word_dec_step<WIDE>(x, cursor, tab, win.ring, o + 64, lt, true);
        word_dec_step<WIDE>(x, cursor, tab, win.ring, o + 96, lt, true);
        win.top_up(cursor, lane);
        word_dec_step<WIDE>(x, cursor, tab, win.ring, o + 128, lt, true);
        word_dec_step<WIDE>(x, cursor, tab, win.ring, o + 160, lt, true);
        word_dec_step<WIDE>(x, cursor, tab, win.ring, o + 192, lt, true);
        word_dec_step<WIDE>(x, cursor, tab, win.ring, o + 224, lt, true);
        o += 256;
    }
    if (g + 4 <= steps) {
        win.top_up(cursor, lane);
        word_dec_step<WIDE>(x, cursor, tab, win.ring, o, lt, true);
        word_dec_step<WIDE>(x, cursor, tab, win.ring, o + 32, lt, true);
        word_dec_step<WIDE>(x, cursor, tab, win.ring, o + 64, lt, true);
        word_dec_step<WIDE>(x, cursor, tab, win.ring, o + 96, lt, true);
        o += 128;
        g += 4;
    }
    win.top_up(cursor, lane);
    for (; g < steps; g++) {
        word_dec_step<WIDE>(x, cursor, tab, win.ring, o, lt, true);
        o += 32;
    }
    if (rem) word_dec_step<WIDE>(x, cursor, tab, win.ring, o, lt, lane < rem);   // main_simd.cpp:328-332

    // A valid stream is consumed exactly to its (aligned) end and leaves every lane
    // in the encoder's initial state.
    const bool bad = (cursor != static_cast<uint32_t>(end)) || (x != kWordL);
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(status, kStatStream);
}

template <bool WIDE>
__global__ void __launch_bounds__(kDecWarps * 32, 8)
word_decode_kernel(const uint8_t* __restrict__ blob, uint64_t blob_size, const uint64_t* __restrict__ offsets,
                   const uint32_t* __restrict__ g_table,   // 4096 packed slots, one model for the whole buffer
                   uint8_t* __restrict__ out, uint64_t n, uint32_t chunk_syms, uint32_t n_chunks,
                   uint32_t* __restrict__ status)
{
    __shared__ __align__(16) uint32_t s_tab[kWordSlots];
    __shared__ __align__(1024) uint8_t s_ring[kDecWarps][kRingBytes];

    for (uint32_t i = threadIdx.x; i < kWordSlots / 4; i += blockDim.x)
        reinterpret_cast<uint4*>(s_tab)[i] = reinterpret_cast<const uint4*>(g_table)[i];
    __syncthreads();

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t chunk = blockIdx.x * kDecWarps + warp;
    if (chunk >= n_chunks) return;
    const uint64_t first = static_cast<uint64_t>(chunk) * chunk_syms;
    const uint32_t m = static_cast<uint32_t>(min(static_cast<uint64_t>(chunk_syms), n - first));
    word_decode_chunk<WIDE>(blob, blob_size, offsets, chunk, smem_addr_pinned(s_tab), smem_addr(&s_ring[warp][0]), out + first, m, status);
}

// ---------------------------------------------------------------------------
// K2: 32-way word-coder encode into per-chunk worst-case slots of a scratch buffer
//
// Per warp: 512 B input stage (16 steps of symbols, filled by one coalesced 128-bit load
// per lane issued a block ahead) and a 512 B output ring (emitted words are placed with
// ballot/popc compaction, flushed to HBM as whole 16-byte vectors).  The per-symbol
// parameters come from ONE conflict-free LDS.128: the 256-entry table is replicated 8x so
// that the 8 lanes of a quarter-warp always hit different 16-byte bank groups.
// ---------------------------------------------------------------------------
// build-time switches of the encoders (A/B runs: nvcc -DNAME=value into a second .so, loaded with RB200_LIB)
#ifndef RB200_FUSED_SLOTS
#define RB200_FUSED_SLOTS 3            // scratch slots per worker warp of the fused word encoder: a chunk is placed SLOTS - 1 chunks late (2, 3, 4 measured: 1.155 / 1.126 / 1.163 ms per GiB)
#endif
constexpr uint32_t kFusedSlots = RB200_FUSED_SLOTS;
#ifndef RB200_ENC_UNIFORM_FLUSH
#define RB200_ENC_UNIFORM_FLUSH 1      // vote on the (warp-uniform) flush test so that ptxas emits a uniform branch
#endif
#ifndef RB200_ENC_KEEP_HINT
#define RB200_ENC_KEEP_HINT false      // evict_last hint on the fused encoder's scratch stores (see word_enc_flush)
#endif
#ifndef RB200_ENC_ASM_RENORM
#define RB200_ENC_ASM_RENORM 1
#endif
constexpr int kEncWarps = 16;
constexpr uint32_t kEncReplicas = 8;
constexpr uint32_t kEncStageBytes = 512;     // 16 steps x 32 symbols
constexpr uint32_t kEncRingBytes = 512;
constexpr uint32_t kEncTableBytes = 256 * kEncReplicas * 16;                 // 32 KiB
constexpr uint32_t kEncWarpSmem = kEncStageBytes + kEncRingBytes;            // 1 KiB per warp

// {magic, x_max | shift, start, 4096 - freq}
//   x_max = freq << 20 in 32-bit arithmetic (rans_word_sse41.h:85; wraps to 0 for freq 4096, as the
//   reference does); its low 20 bits are zero, so the 4-bit reciprocal shift rides in them:
//   x >= x_max  <=>  (x | 31) >= (x_max | shift), and the funnel shift only looks at the low 5 bits.
// R32: `e` comes from the 32-bit-reciprocal table (tables.h: enc32).  mulhi(x, 2^32 - 1) = x - 1 for the
// freq-1 entries, so their `start` carries the missing (4096 - 1).
template <bool R32>
__device__ __forceinline__ uint4 word_enc_expand(WordEncEntry e)
{
    const uint32_t freq = e.packed & 0x1fffu, start = (e.packed >> 13) & 0xfffu, shift = (e.packed >> 25) & 0xfu;
    const uint32_t fix = (R32 && freq == 1) ? kWordSlots - 1 : 0;
    return make_uint4(e.magic, (freq << 20) | shift, start + fix, kWordSlots - freq);    // freq 0 (bad symbol) <=> w == 4096
}

__device__ __forceinline__ uint32_t funnel_shr_wrap(uint32_t lo, uint32_t hi, uint32_t n)
{
    uint32_t r;
    asm("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(n));
    return r;
}

struct WordEncState {
    uint32_t x;        // rANS state
    uint32_t wpos;     // ring byte position (relative to slot end, un-wrapped) of the next word = 510 - 2 * words_emitted
    uint32_t flags;    // OR of the table's (4096 - freq) words: bit 12 set <=> a symbol with freq 0 was met
};

// RansWordEncPut for 32 lanes (rans_word_sse41.h:81-93)
// FULL: all 32 lanes take part (the caller passes active == true); the ragged first step of a chunk is the only
// one that does not.
template <bool R32, bool FULL>
__device__ __forceinline__ void word_enc_step(WordEncState& st, uint32_t sym, uint32_t tab_lane, uint32_t ring, uint32_t gt, bool active)
{
    bool need = false;
    uint4 e = make_uint4(0, 0, 0, 0);
    if (active) {
        e = lds_u128_ro(tab_lane + sym * (kEncReplicas * 16));
        st.flags |= e.w;
        need = (st.x | 31u) >= e.y;                                   // x >= ((L >> 12) << 16) * freq, :85
    }
#if RB200_ENC_ASM_RENORM
    // the same renormalisation as below with ONE predicate feeding the vote, the store and the shift (the
    // compiler otherwise evaluates the comparison twice and copies x before shifting it)
    if (FULL) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            ".reg .b32 m, r, a;\n\t"
            "setp.ge.u32 p, %2, %3;\n\t"
            "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
            "and.b32 r, m, %4;\n\t"
            "popc.b32 r, r;\n\t"
            "shl.b32 r, r, 1;\n\t"
            "sub.u32 a, %1, r;\n\t"
            "and.b32 a, a, %5;\n\t"
            "or.b32 a, a, %6;\n\t"
            "@p st.shared.u16 [a], %0;\n\t"
            "@p shr.u32 %0, %0, 16;\n\t"
            "popc.b32 m, m;\n\t"
            "shl.b32 m, m, 1;\n\t"
            "sub.u32 %1, %1, m;\n\t"
            "}"
            : "+r"(st.x), "+r"(st.wpos)
            : "r"(st.x | 31u), "r"(e.y), "r"(gt), "n"(kEncRingBytes - 1), "r"(ring)
            : "memory");
    } else
#endif
    {
    const uint32_t mask = __ballot_sync(0xffffffffu, need);
    if (need) {
        sts_u16(ring | ((st.wpos - 2u * __popc(mask & gt)) & (kEncRingBytes - 1)), st.x);   // :86-87, lanes 31..0 downwards
        st.x >>= 16;                                                                        // :88
    }
    st.wpos -= 2u * __popc(mask);
    }
    if (active) {
        uint32_t q;
        if (R32) {
            // x < freq << 20 here, where the 32-bit reciprocal is exact: q = mulhi(x, M32) >> shift
            q = funnel_shr_wrap(__umulhi(st.x, e.x), 0u, e.y);
        } else {
            // q = x / freq exactly for any x: M = 2^32 + magic = ceil(2^(32+shift) / freq), q = (x + mulhi(x, magic)) >> shift
            const uint32_t hi = __umulhi(st.x, e.x);
            const uint32_t lo = st.x + hi;
            q = funnel_shr_wrap(lo, lo < hi ? 1u : 0u, e.y);
        }
        st.x = st.x + e.z + q * e.w;                       // ((x / freq) << 12) + x % freq + start, :92
    }
}

// flush every complete 16-byte vector of the ring; *flushed = bytes already written below slot_end.
// KEEP: the destination is a scratch slot that is read back soon (fused paths): ask the L2 to hold on to it.
template <bool KEEP = false>
__device__ __forceinline__ void word_enc_flush(uint32_t produced, uint32_t& flushed, uint32_t ring, uint8_t* slot_end, uint32_t lane)
{
    const uint32_t nvec = (produced - flushed) >> 4;
    __syncwarp();
    if (lane < nvec) {
        const uint32_t v = (flushed >> 4) + lane + 1;                 // vector v ends 16 * (v - 1) bytes below slot_end
        const uint4 q = lds_u128(ring + ((0u - 16u * v) & (kEncRingBytes - 1)));
        if (KEEP) stg_hint_u128(reinterpret_cast<uint4*>(slot_end - 16ull * v), q, l2_policy_keep());
        else stg_stream_u128(reinterpret_cast<uint4*>(slot_end - 16ull * v), q);
    }
    flushed += nvec << 4;
    __syncwarp();
}

// Encode m symbols (one warp) as one 32-way stream that ENDS at slot_end (16-byte aligned) and grows
// downwards; returns the stream size in bytes (warp-uniform).
//   tab  = shared address of the 8x replicated uint4 table
//   wsm  = shared address of this warp's 1 KiB (stage, then ring; 512-byte aligned)
//   KEEP = the slot is scratch that is read back soon (see word_enc_flush)
template <bool R32, bool KEEP = false>
__device__ __forceinline__ uint32_t word_encode_stream(const uint8_t* __restrict__ chunk_in, uint32_t m, uint32_t tab, uint32_t wsm,
                                                       uint8_t* __restrict__ slot_end, uint32_t* __restrict__ status)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t stage = wsm, ring = wsm + kEncStageBytes;
    const uint32_t tab_lane = tab + (lane & (kEncReplicas - 1)) * 16;
    const uint32_t gt = lanemask_gt();

    WordEncState st;
    st.x = kWordL;                              // RansWordEncInit, rans_word_sse41.h:75-78
    st.wpos = kEncRingBytes - 2;
    st.flags = 0;
    uint32_t flushed = 0;
    const uint32_t steps = m >> 5, rem = m & 31;
    const uint32_t nblk = steps >> 4;           // full 16-step blocks, staged; the rest is read directly
    const bool vec_ok = (reinterpret_cast<uintptr_t>(chunk_in) & 15) == 0;

    // symbols are walked last to first (main_simd.cpp:294): ragged tail first
    uint4 parked = make_uint4(0, 0, 0, 0);
    if (nblk && vec_ok) parked = ldg_stream_u128(reinterpret_cast<const uint4*>(chunk_in + (nblk - 1) * kEncStageBytes) + lane);
    if (rem) {
        const bool active = lane < rem;
        const uint32_t s = active ? chunk_in[static_cast<uint64_t>(steps) * 32 + lane] : 0;
        word_enc_step<R32, false>(st, s, tab_lane, ring, gt, active);
    }
    for (uint32_t g = steps; g > nblk * 16; g--) {
        const uint32_t s = chunk_in[static_cast<uint64_t>(g - 1) * 32 + lane];
        word_enc_step<R32, true>(st, s, tab_lane, ring, gt, true);
        if (((g - 1) & 3) == 0) word_enc_flush<KEEP>(kEncRingBytes - 2 - st.wpos, flushed, ring, slot_end, lane);
    }
    word_enc_flush<KEEP>(kEncRingBytes - 2 - st.wpos, flushed, ring, slot_end, lane);

    for (uint32_t b = nblk; b-- > 0;) {
        __syncwarp();
        if (vec_ok) {
            sts_u128(stage + lane * 16, parked);
            if (b) parked = ldg_stream_u128(reinterpret_cast<const uint4*>(chunk_in + (b - 1) * kEncStageBytes) + lane);
        } else {                                 // unaligned input: byte loads into the stage
            const uint8_t* p = chunk_in + b * kEncStageBytes + lane;
#pragma unroll
            for (int j = 0; j < 16; j++) sts_u8(stage + j * 32 + lane, p[j * 32]);
        }
        __syncwarp();
#pragma unroll
        for (int grp = 3; grp >= 0; grp--) {
#pragma unroll
            for (int j = 3; j >= 0; j--)
                word_enc_step<R32, true>(st, lds_u8(stage + (grp * 4 + j) * 32 + lane), tab_lane, ring, gt, true);
            // <= 256 bytes per 4 steps; flushing whenever >= 256 are pending keeps the 512-byte ring safe.  The test is
            // warp-uniform; voting on it tells ptxas so (a uniform branch: no convergence check before the next ballot)
#if RB200_ENC_UNIFORM_FLUSH
            if (__any_sync(0xffffffffu, kEncRingBytes - 2 - st.wpos - flushed >= 256))
#else
            if (kEncRingBytes - 2 - st.wpos - flushed >= 256)
#endif
                word_enc_flush<KEEP>(kEncRingBytes - 2 - st.wpos, flushed, ring, slot_end, lane);
        }
    }

    // RansWordEncFlush for lanes 31..0 (main_simd.cpp:298-299): lane 31's hi word first (highest), lane 0's lo word last
    word_enc_flush<KEEP>(kEncRingBytes - 2 - st.wpos, flushed, ring, slot_end, lane);     // < 16 bytes stay pending
    const uint32_t hpos = st.wpos - 4u * (31 - lane);
    sts_u16(ring | (hpos & (kEncRingBytes - 1)), st.x >> 16);
    sts_u16(ring | ((hpos - 2) & (kEncRingBytes - 1)), st.x);
    st.wpos -= kHeaderBytes;
    const uint32_t produced = kEncRingBytes - 2 - st.wpos;           // total stream bytes
    word_enc_flush<KEEP>(produced, flushed, ring, slot_end, lane);
    const uint32_t left = produced - flushed;                        // < 16, even: head of the stream, not vector aligned
    if (2 * lane < left) {
        const uint32_t off = flushed + 2 * lane + 2;                 // bytes below slot_end
        *reinterpret_cast<uint16_t*>(slot_end - off) = static_cast<uint16_t>(lds_u16(ring | ((0u - off) & (kEncRingBytes - 1))));
    }
    if (__any_sync(0xffffffffu, (st.flags & kWordSlots) != 0) && lane == 0) atomicOr(status, kStatSymbol);
    return produced;
}

// chunk `chunk` into the end of its own worst-case slot of `scratch`; size to sizes[chunk]
template <bool R32>
__device__ __forceinline__ void word_encode_chunk(const uint8_t* __restrict__ chunk_in, uint32_t m, uint32_t chunk,
                                                  uint32_t tab, uint32_t wsm, uint8_t* __restrict__ scratch,
                                                  uint32_t slot_bytes, uint32_t* __restrict__ sizes, uint32_t* __restrict__ status)
{
    const uint32_t produced = word_encode_stream<R32>(chunk_in, m, tab, wsm, scratch + static_cast<uint64_t>(chunk + 1) * slot_bytes, status);
    if ((threadIdx.x & 31) == 0) sizes[chunk] = produced;
}

#ifndef RB200_ENC_MINBLOCKS
#define RB200_ENC_MINBLOCKS 3
#endif
template <bool R32>
__global__ void __launch_bounds__(kEncWarps * 32, RB200_ENC_MINBLOCKS)
word_encode_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t chunk_syms, uint32_t n_chunks,
                   const WordEncEntry* __restrict__ g_table, uint8_t* __restrict__ scratch, uint32_t slot_bytes,
                   uint32_t* __restrict__ sizes, uint32_t* __restrict__ status)
{
    extern __shared__ __align__(1024) uint8_t s_enc[];     // [32 KiB table][warps x 1 KiB]
    uint4* s_tab = reinterpret_cast<uint4*>(s_enc);
    for (uint32_t i = threadIdx.x; i < 256 * kEncReplicas; i += blockDim.x) s_tab[i] = word_enc_expand<R32>(g_table[i / kEncReplicas]);
    __syncthreads();

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t chunk = blockIdx.x * kEncWarps + warp;
    if (chunk >= n_chunks) return;
    const uint64_t first = static_cast<uint64_t>(chunk) * chunk_syms;
    const uint32_t m = static_cast<uint32_t>(min(static_cast<uint64_t>(chunk_syms), n - first));
    word_encode_chunk<R32>(in + first, m, chunk, smem_addr(s_enc), smem_addr(s_enc) + kEncTableBytes + warp * kEncWarpSmem, scratch,
                      slot_bytes, sizes, status);
}
constexpr uint32_t kEncSmemBytes = kEncTableBytes + kEncWarps * kEncWarpSmem;

// ---------------------------------------------------------------------------
// K2f: persistent encode with the directory and the compaction fused in.
//
// Every worker warp fetches chunk ids from an atomic counter (so chunks start in order), encodes the
// chunk into one of two scratch slots that belong to the WARP and publishes its padded size.  One
// scanner warp turns the published sizes into end offsets in chunk order.  A worker places chunk k
// (moves its stream to blob[E_k - size .. E_k) and writes the directory entry) only after it has
// encoded chunk k+1, so its E_k is ready and no warp polls.  One launch, no CTA barrier in the loop.
// (Two earlier variants -- per-CTA and per-warp decoupled look-back -- were correct but slower than the
// split path because of polling)
// ---------------------------------------------------------------------------
constexpr uint64_t kLookAgg = 1ull << 62, kLookPrefix = 2ull << 62, kLookValue = (1ull << 62) - 1;
constexpr uint64_t kLookStallNs = 8ull * 1000 * 1000 * 1000;   // a bug must not hang the GPU: after 8 s without progress give up and
                                                               // flag kStatStall (time, not iterations: the waits are legitimate and
                                                               // stretch by orders of magnitude under compute-sanitizer)

__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t* p)
{
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t* p, uint64_t v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint4 ldg_l2_u128(const uint4* p)      // L2-coherent: data this kernel wrote itself
{
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// move one finished stream (it ends at src_end, 16-byte aligned) to blob[end - size .. end)
__device__ __forceinline__ void place_stream(const uint8_t* src_end, uint32_t size, uint8_t* __restrict__ blob, uint64_t end, uint32_t lane)
{
    const uint32_t padded = (size + 15u) & ~15u;
    const uint32_t gap = padded - size;
    const uint8_t* src = src_end - size;
    uint8_t* dst_vec0 = blob + (end - padded);
    if (lane < 16) {
        uint8_t v = 0;
        if (lane >= gap) v = *reinterpret_cast<const volatile uint8_t*>(src + (lane - gap));
        dst_vec0[lane] = v;
    }
    const uint4* s4 = reinterpret_cast<const uint4*>(src + (16 - gap));
    uint4* d4 = reinterpret_cast<uint4*>(dst_vec0 + 16);
    const uint32_t nvec = padded / 16 - 1;
    uint32_t v = lane;
    for (; v + 96 < nvec; v += 128) {                      // four loads in flight per lane
        const uint4 a = ldg_l2_u128(s4 + v), b = ldg_l2_u128(s4 + v + 32), c = ldg_l2_u128(s4 + v + 64), d = ldg_l2_u128(s4 + v + 96);
        stg_stream_u128(d4 + v, a);
        stg_stream_u128(d4 + v + 32, b);
        stg_stream_u128(d4 + v + 64, c);
        stg_stream_u128(d4 + v + 96, d);
    }
    for (; v < nvec; v += 32) stg_stream_u128(d4 + v, ldg_l2_u128(s4 + v));
}

// The scanner: ONE warp of the grid walks the status array in chunk order and rewrites every published
// size (kLookAgg | padded size) into the inclusive end offset E_c (kLookPrefix | E_c).  Chunks are handed
// out in order, so the entry it waits for always belongs to a chunk some resident warp is encoding.
__device__ __forceinline__ void fused_scanner(uint64_t* look, uint32_t n_chunks, uint32_t lane, uint32_t* status)
{
    constexpr int kPer = 8;                 // consecutive entries per lane: 256 entries per trip, one warp scan per trip
    uint64_t run = 0, idle_since = 0;
    uint32_t pos = 0;
    while (pos < n_chunks) {
        const uint32_t base = pos + kPer * lane;
        uint64_t v[kPer];
#pragma unroll
        for (int j = 0; j < kPer; j++) v[j] = (base + j < n_chunks) ? ld_relaxed_u64(look + base + j) : 0;   // independent loads
        // r = length of this lane's leading run of published entries; local inclusive sums over that run
        uint32_t r = 0;
        uint64_t sum[kPer];
        uint64_t acc = 0;
        bool open = true;
#pragma unroll
        for (int j = 0; j < kPer; j++) {
            open = open && ((v[j] >> 62) == 1);
            if (open) { acc += v[j] & kLookValue; r = j + 1; }
            sum[j] = acc;
        }
        // the contiguous ready prefix of the whole window: all of lanes < f, the first r_f entries of lane f
        const uint32_t full = __ballot_sync(0xffffffffu, r == kPer);
        const uint32_t f = full == 0xffffffffu ? 32u : static_cast<uint32_t>(__ffs(~full)) - 1u;
        const uint32_t cnt = lane < f ? kPer : (lane == f ? r : 0u);
        const uint64_t mine = cnt ? sum[cnt - 1] : 0;
        uint64_t incl = mine;
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= static_cast<uint32_t>(d)) incl += t;
        }
        const uint64_t before = run + incl - mine;
#pragma unroll
        for (int j = 0; j < kPer; j++)
            if (static_cast<uint32_t>(j) < cnt) st_relaxed_u64(look + base + j, kLookPrefix | (before + sum[j]));
        run += __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t r_f = __shfl_sync(0xffffffffu, r, f & 31);
        const uint32_t done = f == 32 ? 32u * kPer : f * kPer + r_f;
        pos += done;
        if (done == 0) {
            const uint64_t now = global_timer_ns();
            if (!idle_since) idle_since = now;
            if (now - idle_since > kLookStallNs) {   // never hang the GPU: give up and flag
                if (lane == 0) atomicOr(status, kStatStall);
                return;
            }
            __nanosleep(100);
        } else {
            idle_since = 0;
        }
    }
}

// one warp: wait for the scanner to publish E_c of `chunk`, then move the stream and write the directory entry
__device__ __forceinline__ void fused_place(const uint64_t* look, uint32_t chunk, uint32_t n_chunks, const uint8_t* slot_end,
                                            uint32_t produced, uint8_t* __restrict__ blob, uint64_t blob_cap,
                                            uint64_t* __restrict__ offsets, uint32_t lane, uint32_t* status)
{
    // Warps do not progress evenly, so the scan front (the oldest unfinished chunk) can trail this warp by a chunk time
    // and more: the wait is real, and every poll costs ~12 issue slots that the encoding warps of the SM want.  Back off
    // exponentially (128 ns .. 4 us): round 1 polled every 100 ns and spent a fifth of the kernel's instructions here
    // (ncu source view: 5.3 M poll iterations per GiB, profiles/r2_encode_experiments.md).
    uint64_t v = 0, since = 0;
    uint32_t ns = 128;
    for (;;) {
        if (lane == 0) v = ld_relaxed_u64(look + chunk);
        v = __shfl_sync(0xffffffffu, v, 0);
        if ((v >> 62) == 2) break;
        if (ns >= 4096) {                                  // only once the back-off has topped out: keep the common path short
            const uint64_t now = global_timer_ns();
            if (!since) since = now;
            if (now - since > kLookStallNs) {
                if (lane == 0) atomicOr(status, kStatStall);
                return;
            }
        }
        __nanosleep(ns);
        if (ns < 4096) ns <<= 1;
    }
    const uint64_t end = v & kLookValue;                   // E_c
    if (lane == 0) {
        offsets[chunk] = end - produced;
        if (chunk == n_chunks - 1) offsets[n_chunks] = end;
    }
    if (end <= blob_cap) {
        place_stream(slot_end, produced, blob, end, lane);
    } else if (lane == 0) {
        atomicOr(status, kStatSpace);
    }
}

// Grid = resident capacity (SMs x RB200_ENC_MINBLOCKS CTAs).  Warp 0 of CTA 0 is the scanner, every other warp is
// a worker with TWO scratch slots: it encodes chunk k+1 into one slot before it places chunk k from the other,
// so by the time it asks for E_k the scanner has normally long passed k and nobody polls.
template <bool R32>
__global__ void __launch_bounds__(kEncWarps * 32, RB200_ENC_MINBLOCKS)
word_encode_fused_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t chunk_syms, uint32_t n_chunks,
                         const WordEncEntry* __restrict__ g_table, uint8_t* __restrict__ scratch, uint32_t slot_bytes,
                         uint64_t* __restrict__ look, uint32_t* __restrict__ counter, uint8_t* __restrict__ blob, uint64_t blob_cap,
                         uint64_t* __restrict__ offsets, uint32_t* __restrict__ status)
{
    extern __shared__ __align__(1024) uint8_t s_enc[];     // [32 KiB table][warps x 1 KiB]
    uint4* s_tab = reinterpret_cast<uint4*>(s_enc);
    for (uint32_t i = threadIdx.x; i < 256 * kEncReplicas; i += blockDim.x) s_tab[i] = word_enc_expand<R32>(g_table[i / kEncReplicas]);
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (blockIdx.x == 0 && warp == 0) {                    // no CTA barrier below: warps are independent from here on
        fused_scanner(look, n_chunks, lane, status);
        return;
    }
    const uint32_t tab = smem_addr(s_enc), wsm = tab + kEncTableBytes + warp * kEncWarpSmem;
    // this warp's kFusedSlots scratch slots: chunk j of this warp goes to slot j % kFusedSlots and is placed
    // kFusedSlots - 1 chunks later
    uint8_t* slots = scratch + (static_cast<uint64_t>(blockIdx.x) * kEncWarps + warp) * kFusedSlots * slot_bytes;

    // The previous chunk(s) are placed late on purpose.  Warps do not progress evenly (the issue arbiter favours high warp
    // ids), so the scan front -- the oldest unfinished chunk -- trails the fastest warps by about one chunk time: placing
    // 32 steps into the next chunk instead was measured at 2.39 ms per GiB against 1.15 ms, 580 M polling instructions
    // (profiles/r2_encode_experiments.md).
    uint32_t pend_chunk[kFusedSlots - 1], pend_size[kFusedSlots - 1];      // oldest first; registers (static indexing only)
    uint32_t n_pending = 0, cur = 0;                                        // cur = slot of the chunk being encoded
    for (;;) {
        uint32_t chunk = 0;
        if (lane == 0) chunk = atomicAdd(counter, 1u);     // chunks start in order
        chunk = __shfl_sync(0xffffffffu, chunk, 0);
        if (chunk >= n_chunks) break;
        const uint64_t first = static_cast<uint64_t>(chunk) * chunk_syms;
        const uint32_t m = static_cast<uint32_t>(min(static_cast<uint64_t>(chunk_syms), n - first));
        uint8_t* slot_end = slots + (cur + 1) * static_cast<uint64_t>(slot_bytes);
        const uint32_t produced = word_encode_stream<R32, RB200_ENC_KEEP_HINT>(in + first, m, tab, wsm, slot_end, status);
        if (lane == 0) st_relaxed_u64(look + chunk, kLookAgg | ((produced + 15u) & ~15u));
        __syncwarp();
        if (n_pending == kFusedSlots - 1) {                // every other slot is taken: place the oldest, it sits in slot cur + 1
            const uint32_t oldest = cur + 1 == kFusedSlots ? 0 : cur + 1;
            fused_place(look, pend_chunk[0], n_chunks, slots + (oldest + 1) * static_cast<uint64_t>(slot_bytes), pend_size[0], blob, blob_cap,
                        offsets, lane, status);
#pragma unroll
            for (int j = 0; j + 1 < kFusedSlots - 1; j++) {
                pend_chunk[j] = pend_chunk[j + 1];
                pend_size[j] = pend_size[j + 1];
            }
            n_pending--;
        }
#pragma unroll
        for (int j = 0; j < kFusedSlots - 1; j++)
            if (static_cast<uint32_t>(j) == n_pending) {
                pend_chunk[j] = chunk;
                pend_size[j] = produced;
            }
        n_pending++;
        cur = cur + 1 == kFusedSlots ? 0 : cur + 1;
        __syncwarp();
    }
    // drain, oldest first: the k-th pending chunk sits k slots behind the next free one
#pragma unroll
    for (int j = 0; j < kFusedSlots - 1; j++)
        if (static_cast<uint32_t>(j) < n_pending) {
            const uint32_t slot = (cur + kFusedSlots - n_pending + j) % kFusedSlots;
            fused_place(look, pend_chunk[j], n_chunks, slots + (slot + 1) * static_cast<uint64_t>(slot_bytes), pend_size[j], blob, blob_cap,
                        offsets, lane, status);
        }
}

// ---------------------------------------------------------------------------
// K6: directory + compaction of the per-chunk slots into the final blob
//
// offsets[c] = E_c - sizes[c], E_c = sum_{j<=c} round_up_16(sizes[j]); offsets[n_chunks] = E_last.
// Pass 1 (scan_tiles_kernel, one CTA per 4096 chunks): tile-local inclusive ends into offsets[],
// tile totals into tile_sums[].  Pass 2 (compact_kernel): every CTA adds the totals of the tiles
// before its own (a handful of values), finalises its directory entries and moves its streams.
// ---------------------------------------------------------------------------
constexpr uint32_t kScanTile = 4096;

__global__ void __launch_bounds__(1024)
scan_tiles_kernel(const uint32_t* __restrict__ sizes, uint32_t n_chunks, uint64_t* __restrict__ offsets,
                  uint64_t* __restrict__ tile_sums)
{
    __shared__ uint64_t s_warp[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t c0 = blockIdx.x * kScanTile + tid * 4;
    uint32_t pad[4];
    uint64_t sum = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        pad[j] = (c0 + j < n_chunks) ? ((sizes[c0 + j] + 15u) & ~15u) : 0u;
        sum += pad[j];
    }
    uint64_t incl = sum;
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= static_cast<uint32_t>(d)) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint64_t w = s_warp[lane];
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t v = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= static_cast<uint32_t>(d)) w += v;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    uint64_t run = incl - sum + (warp ? s_warp[warp - 1] : 0);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        run += pad[j];
        if (c0 + j < n_chunks) offsets[c0 + j] = run;
    }
    if (tid == 1023) tile_sums[blockIdx.x] = s_warp[31];
}

// for very large directories: turn tile_sums into exclusive prefixes once, so compact CTAs read one value
__global__ void __launch_bounds__(1024)
tile_prefix_kernel(uint64_t* __restrict__ tile_sums, uint32_t n_tiles)
{
    __shared__ uint64_t s_warp[32];
    __shared__ uint64_t s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_tiles; base += 1024) {
        const uint32_t i = base + tid;
        const uint64_t v = i < n_tiles ? tile_sums[i] : 0;
        uint64_t incl = v;
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= static_cast<uint32_t>(d)) incl += u;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint64_t w = s_warp[lane];
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t u = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= static_cast<uint32_t>(d)) w += u;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const uint64_t excl = s_carry + (warp ? s_warp[warp - 1] : 0) + incl - v;
        if (i < n_tiles) tile_sums[i] = excl;
        __syncthreads();
        if (tid == 0) s_carry += s_warp[31];
        __syncthreads();
    }
}

constexpr int kCopyWarps = 8;
constexpr uint32_t kTilePrefixThreshold = 256;   // above this many tiles, tile_prefix_kernel runs first

// one warp moves one chunk stream from the end of its scratch slot to blob[offsets[c]..E_c)
__global__ void __launch_bounds__(kCopyWarps * 32)
compact_kernel(const uint8_t* __restrict__ scratch, uint32_t slot_bytes, const uint32_t* __restrict__ sizes,
               uint64_t* __restrict__ offsets, const uint64_t* __restrict__ tile_sums, uint32_t tiles_prefixed,
               uint32_t n_chunks, uint8_t* __restrict__ blob, uint64_t blob_cap, uint32_t* __restrict__ status)
{
    __shared__ uint64_t s_base;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t chunk0 = blockIdx.x * kCopyWarps;
    if (warp == 0) {      // all chunks of this CTA live in one scan tile (kScanTile % kCopyWarps == 0)
        const uint32_t tile = chunk0 / kScanTile;
        uint64_t acc = 0;
        if (tiles_prefixed) {
            acc = tile_sums[tile];
        } else {
            for (uint32_t t = lane; t < tile; t += 32) acc += tile_sums[t];
            for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        }
        if (lane == 0) s_base = acc;
    }
    __syncthreads();
    const uint32_t chunk = chunk0 + warp;
    if (chunk >= n_chunks) return;
    const uint32_t size = sizes[chunk];
    const uint64_t end = s_base + offsets[chunk];          // E_c, multiple of 16 (tile-local value from pass 1)
    const uint64_t off = end - size;
    __syncwarp();
    if (lane == 0) {
        offsets[chunk] = off;
        if (chunk == n_chunks - 1) {
            offsets[n_chunks] = end;
            if (end > blob_cap) atomicOr(status, kStatSpace);
        }
    }
    if (end > blob_cap) return;
    const uint32_t padded = (size + 15u) & ~15u;
    const uint32_t gap = padded - size;                    // zero bytes in front of the stream, < 16
    const uint8_t* src = scratch + static_cast<uint64_t>(chunk + 1) * slot_bytes - size;
    uint8_t* dst_vec0 = blob + (end - padded);             // 16-byte aligned

    if (lane < 16) {                                       // first vector: gap zeros + stream head
        uint8_t v = 0;
        if (lane >= gap) v = src[lane - gap];
        dst_vec0[lane] = v;
    }
    const uint4* s4 = reinterpret_cast<const uint4*>(src + (16 - gap));   // 16-byte aligned by construction
    uint4* d4 = reinterpret_cast<uint4*>(dst_vec0 + 16);
    const uint32_t nvec = padded / 16 - 1;
    for (uint32_t v = lane; v < nvec; v += 32) stg_stream_u128(d4 + v, ldg_stream_u128(s4 + v));
}

}  // namespace rb200
