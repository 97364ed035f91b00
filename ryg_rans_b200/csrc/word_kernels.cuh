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
    const uint4* next;      // this lane's 16 B of the next unit to fetch
    const uint4* limit;     // end of blob (exclusive)
    uint4 parked;           // unit in flight

    __device__ __forceinline__ uint4 fetch()
    {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (next < limit) v = ldg_stream_u128(next);
        next += 32;
        return v;
    }

    // off = absolute byte offset of the stream start inside the blob
    __device__ __forceinline__ void open(const uint8_t* blob, uint64_t blob_size, uint64_t off, uint32_t ring_addr, uint32_t lane)
    {
        ring = ring_addr;
        const uint64_t unit0 = off / kUnitBytes;
        next = reinterpret_cast<const uint4*>(blob) + unit0 * 32 + lane;
        limit = reinterpret_cast<const uint4*>(blob) + blob_size / 16;
        uint4 a = fetch();
        uint4 b = fetch();
        parked = fetch();
        sts_u128(ring + (static_cast<uint32_t>(unit0 & 1) * kUnitBytes) + lane * 16, a);
        sts_u128(ring + (static_cast<uint32_t>((unit0 + 1) & 1) * kUnitBytes) + lane * 16, b);
        fill_end = static_cast<uint32_t>((unit0 + 2) * kUnitBytes);
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
template <bool WIDE>
__device__ __forceinline__ void word_dec_step(uint32_t& x, uint32_t& cursor, uint32_t tab /*shared addr*/, uint32_t ring,
                                              uint8_t* o, uint32_t lt, bool active)
{
    bool need = false;
    if (active) {
        const uint32_t e = lds_u32_ro(tab + 4u * (x & (kWordSlots - 1)));       // rans_word_sse41.h:126
        x = word_table_freq<WIDE>(e) * (x >> kWordScaleBits) + ((e >> 8) & 0xfffu);   // :129
        *o = static_cast<uint8_t>(e);                                       // :130
        need = x < kWordL;                                                  // :137
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, need);
    const uint32_t a = cursor + 2u * __popc(mask & lt);                     // lane order within the step
    const uint32_t w = lds_u16(ring | (a & (kRingBytes - 1)));
    if (need) x = (x << 16) | w;                                            // :138
    cursor += 2u * __popc(mask);                                            // :139
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
    for (; g + 4 <= steps; g += 4) {
        win.top_up(cursor, lane);
        word_dec_step<WIDE>(x, cursor, tab, win.ring, o, lt, true);
        word_dec_step<WIDE>(x, cursor, tab, win.ring, o + 32, lt, true);
        word_dec_step<WIDE>(x, cursor, tab, win.ring, o + 64, lt, true);
        word_dec_step<WIDE>(x, cursor, tab, win.ring, o + 96, lt, true);
        o += 128;
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
    word_decode_chunk<WIDE>(blob, blob_size, offsets, chunk, smem_addr(s_tab), smem_addr(&s_ring[warp][0]), out + first, m, status);
}

// ---------------------------------------------------------------------------
// K2: 32-way word-coder encode into per-chunk worst-case slots of a scratch buffer
// ---------------------------------------------------------------------------
constexpr int kEncWarps = 8;
constexpr uint32_t kEncReplicas = 16;   // copies of the 256-entry table: lane l reads copy l & 15 -> conflict-free LDS.64

__device__ __forceinline__ void word_enc_step(uint32_t& x, uint32_t& emitted, uint32_t& flags, uint32_t sym,
                                              uint32_t tab_lane /*shared addr of this lane's replica*/, uint16_t* slot_end,
                                              uint32_t gt, bool active)
{
    bool need = false;
    uint32_t magic = 0, freq = 1, start = 0, shift = 0;
    if (active) {
        const uint2 ent = lds_u64_ro(tab_lane + sym * (kEncReplicas * 8));
        magic = ent.x;
        freq = ent.y & 0x1fffu;
        start = (ent.y >> 13) & 0xfffu;
        shift = (ent.y >> 25) & 0xfu;
        flags |= ent.y;
        need = x >= (freq << 20);            // ((L >> 12) << 16) * freq in 32-bit arithmetic, rans_word_sse41.h:85
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, need);
    if (need) {
        slot_end[-static_cast<int64_t>(emitted + 1 + __popc(mask & gt))] = static_cast<uint16_t>(x);   // :86-87
        x >>= 16;                                                                                      // :88
    }
    emitted += __popc(mask);
    if (active) {
        // q = x / freq exactly: M = 2^32 + magic = ceil(2^(32+shift) / freq)
        const uint32_t q = static_cast<uint32_t>((static_cast<uint64_t>(x) + __umulhi(x, magic)) >> shift);
        x = x + start + q * (kWordSlots - freq);          // == ((x/freq) << 12) + x%freq + start, :92
    }
}

// Encode chunk `chunk` (one warp) into the END of its scratch slot; s_tab = shared address of
// the 16x replicated {magic, packed} table.
__device__ __forceinline__ void word_encode_chunk(const uint8_t* __restrict__ chunk_in, uint32_t m, uint32_t chunk,
                                                  uint32_t s_tab, uint8_t* __restrict__ scratch, uint32_t slot_bytes,
                                                  uint32_t* __restrict__ sizes, uint32_t* __restrict__ status)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint8_t* src = chunk_in + lane;
    uint16_t* slot_end = reinterpret_cast<uint16_t*>(scratch + static_cast<uint64_t>(chunk + 1) * slot_bytes);
    const uint32_t tab_lane = s_tab + (lane & (kEncReplicas - 1)) * 8;
    const uint32_t gt = lanemask_gt();

    uint32_t x = kWordL;                       // RansWordEncInit, rans_word_sse41.h:75-78
    uint32_t emitted = 0, flags = 0;
    const uint32_t steps = m >> 5, rem = m & 31;

    // symbols are walked last to first (main_simd.cpp:294): ragged tail step first
    if (rem) {
        const bool active = lane < rem;
        const uint32_t s = active ? src[static_cast<uint64_t>(steps) * 32] : 0;
        word_enc_step(x, emitted, flags, s, tab_lane, slot_end, gt, active);
    }
    uint32_t g = steps;
    for (; g >= 4; g -= 4) {
        const uint8_t* p = src + static_cast<uint64_t>(g - 4) * 32;
        const uint32_t s3 = p[96], s2 = p[64], s1 = p[32], s0 = p[0];
        word_enc_step(x, emitted, flags, s3, tab_lane, slot_end, gt, true);
        word_enc_step(x, emitted, flags, s2, tab_lane, slot_end, gt, true);
        word_enc_step(x, emitted, flags, s1, tab_lane, slot_end, gt, true);
        word_enc_step(x, emitted, flags, s0, tab_lane, slot_end, gt, true);
    }
    for (; g >= 1; g--) {
        const uint32_t s = src[static_cast<uint64_t>(g - 1) * 32];
        word_enc_step(x, emitted, flags, s, tab_lane, slot_end, gt, true);
    }

    // RansWordEncFlush for lanes 31..0 (main_simd.cpp:298-299): lane 0 ends up first in memory
    uint16_t* head = slot_end - emitted - 64;
    head[2 * lane] = static_cast<uint16_t>(x);
    head[2 * lane + 1] = static_cast<uint16_t>(x >> 16);
    if (lane == 0) sizes[chunk] = kHeaderBytes + 2u * emitted;
    if (__any_sync(0xffffffffu, (flags & kEncBadSymbol) != 0) && lane == 0) atomicOr(status, kStatSymbol);
}

__global__ void __launch_bounds__(kEncWarps * 32, 4)
word_encode_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t chunk_syms, uint32_t n_chunks,
                   const WordEncEntry* __restrict__ g_table, uint8_t* __restrict__ scratch, uint32_t slot_bytes,
                   uint32_t* __restrict__ sizes, uint32_t* __restrict__ status)
{
    __shared__ __align__(16) uint2 s_tab[256 * kEncReplicas];   // 32 KiB
    for (uint32_t i = threadIdx.x; i < 256 * kEncReplicas; i += blockDim.x) {
        const WordEncEntry e = g_table[i / kEncReplicas];
        s_tab[i] = make_uint2(e.magic, e.packed);
    }
    __syncthreads();

    const uint32_t chunk = blockIdx.x * kEncWarps + (threadIdx.x >> 5);
    if (chunk >= n_chunks) return;
    const uint64_t first = static_cast<uint64_t>(chunk) * chunk_syms;
    const uint32_t m = static_cast<uint32_t>(min(static_cast<uint64_t>(chunk_syms), n - first));
    word_encode_chunk(in + first, m, chunk, smem_addr(s_tab), scratch, slot_bytes, sizes, status);
}

// ---------------------------------------------------------------------------
// K6: directory scan + compaction of the per-chunk slots into the final blob
// ---------------------------------------------------------------------------

// offsets[c] = E_c - sizes[c], E_c = sum_{j<=c} round_up_16(sizes[j]); offsets[n_chunks] = E_last.
// One CTA; n_chunks is at most a few hundred thousand.
__global__ void __launch_bounds__(1024)
directory_scan_kernel(const uint32_t* __restrict__ sizes, uint32_t n_chunks, uint64_t* __restrict__ offsets,
                      uint64_t blob_cap, uint32_t* __restrict__ status)
{
    __shared__ uint64_t s_warp[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t per = (n_chunks + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = min(n_chunks, tid * per), hi = min(n_chunks, lo + per);
    uint64_t sum = 0;
    for (uint32_t c = lo; c < hi; c++) sum += (sizes[c] + 15u) & ~15u;
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
    for (uint32_t c = lo; c < hi; c++) {
        const uint32_t sz = sizes[c];
        run += (sz + 15u) & ~15u;
        offsets[c] = run - sz;
    }
    if (tid == blockDim.x - 1) {
        const uint64_t total = s_warp[31];
        offsets[n_chunks] = total;
        if (total > blob_cap) atomicOr(status, kStatSpace);
    }
}

constexpr int kCopyWarps = 8;

// one warp moves one chunk stream from the end of its scratch slot to blob[offsets[c]..E_c)
__global__ void __launch_bounds__(kCopyWarps * 32)
compact_kernel(const uint8_t* __restrict__ scratch, uint32_t slot_bytes, const uint32_t* __restrict__ sizes,
               const uint64_t* __restrict__ offsets, uint32_t n_chunks, uint8_t* __restrict__ blob, uint64_t blob_cap)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t chunk = blockIdx.x * kCopyWarps + (threadIdx.x >> 5);
    if (chunk >= n_chunks) return;
    const uint32_t size = sizes[chunk];
    const uint64_t off = offsets[chunk];
    const uint64_t end = off + size;                       // multiple of 16
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
