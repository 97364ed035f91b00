// rans64_kernels.cuh -- sm_100a kernels for the rans64.h coder as main64.cpp drives it:
// 64-bit state in [2^31, 2^63), 32-bit renormalisation units (never loops), cum2sym symbol lookup.
//
//   Rans64DecGet / cum2sym / Rans64DecAdvanceSymbolStep / Rans64DecRenorm   rans64.h:118-121, main64.cpp:202,
//                                                                           rans64.h:289-316
//   Rans64EncPutSymbol (reciprocal form) + Rans64EncFlush                    rans64.h:262-278, :96-103
//
// Same mapping as the other coders (one warp = one chunk = one 32-way stream, one ballot per step
// because a lane moves at most one u32 per step).  This is the "lets the GPU decode streams made by
// the reference's fastest CPU encoder" row of SURVEY 8(f); it reuses the word coder's stream window,
// symbol stage, output ring and flush, and is not tuned beyond that.
#pragma once
#include "device_utils.cuh"
#include "tables.h"
#include "word_kernels.cuh"

namespace rb200 {

constexpr uint64_t kRans64L = 1ull << 31;          // RANS64_L, rans64.h:59
constexpr uint32_t kRans64HeaderBytes = 256;       // 32 lanes x u64 (Rans64EncFlush x 32)
constexpr int kRans64Warps = 16;
constexpr uint32_t kRans64DecReplicas = 16;        // an LDS.64 is served 16 lanes at a time: one {start, freq} copy per lane of a half-warp
constexpr uint32_t kRans64DecSymBytes = 256 * kRans64DecReplicas * 8;      // 32 KiB

// tab = shared address of cum2sym[1 << sb]; syms_lane = this lane's replica of the 256 x {start, freq} table that follows it
// (16 replicas, entry s of replica r at (s * 16 + r) * 8: the 16 lanes of each LDS.64 phase hit 16 different 8-byte bank
// groups whatever their symbols are)
__device__ __forceinline__ void rans64_dec_step(uint64_t& x, uint32_t& cursor, uint32_t tab, uint32_t syms_lane, uint32_t ring, uint8_t* o,
                                                uint32_t lt, uint32_t sb, bool active)
{
    bool need = false;
    if (active) {
        const uint32_t cf = static_cast<uint32_t>(x) & ((1u << sb) - 1);               // rans64.h:120
        const uint32_t s = lds_u8_ro(tab + cf);                                        // main64.cpp:202
        const uint2 ds = lds_u64_ro(syms_lane + s * (kRans64DecReplicas * 8));         // Rans64DecSymbol {start, freq}
        x = static_cast<uint64_t>(ds.y) * (x >> sb) + cf - ds.x;                       // rans64.h:297
        *o = static_cast<uint8_t>(s);
        need = x < kRans64L;                                                           // :309
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, need);
    const uint32_t a = cursor + 4u * __popc(mask & lt);
    const uint32_t w = lds_u32(ring | (a & (kRingBytes - 1)));
    if (need) x = (x << 32) | w;                                                       // :310
    cursor += 4u * __popc(mask);
}

__global__ void __launch_bounds__(kRans64Warps * 32, 2)
rans64_decode_kernel(const uint8_t* __restrict__ blob, uint64_t blob_size, const uint64_t* __restrict__ offsets, uint32_t sb,
                     const uint8_t* __restrict__ g_tab,   // cum2sym[1 << sb] + 256 x {start, freq}
                     uint8_t* __restrict__ out, uint64_t n, uint32_t chunk_syms, uint32_t n_chunks, uint32_t* __restrict__ status)
{
    extern __shared__ __align__(1024) uint8_t s_r64[];        // [16 x 1 KiB rings][cum2sym][256 x 16 replicas x {start, freq}]
    uint4* s_tab = reinterpret_cast<uint4*>(s_r64 + kRans64Warps * kRingBytes);
    const uint32_t vecs = (1u << sb) / 16;
    for (uint32_t i = threadIdx.x; i < vecs; i += blockDim.x) s_tab[i] = reinterpret_cast<const uint4*>(g_tab)[i];
    {
        const uint2* g_syms = reinterpret_cast<const uint2*>(g_tab + (1u << sb));
        uint2* s_syms = reinterpret_cast<uint2*>(s_r64 + kRans64Warps * kRingBytes + (1u << sb));
        for (uint32_t i = threadIdx.x; i < 256 * kRans64DecReplicas; i += blockDim.x) s_syms[i] = g_syms[i / kRans64DecReplicas];
    }
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t chunk = blockIdx.x * kRans64Warps + warp;
    if (chunk >= n_chunks) return;
    const uint64_t first = static_cast<uint64_t>(chunk) * chunk_syms;
    const uint32_t m = static_cast<uint32_t>(min(static_cast<uint64_t>(chunk_syms), n - first));
    const uint64_t off = offsets[chunk];
    const uint64_t end = offsets[chunk + 1] & ~static_cast<uint64_t>(15);
    if ((off & 3) || off + kRans64HeaderBytes > end || end > blob_size) {
        if (lane == 0) atomicOr(status, kStatStream);
        return;
    }

    StreamWindow win;
    win.open(blob, blob_size, off, smem_addr(s_r64) + warp * kRingBytes, lane);
    uint32_t cursor = static_cast<uint32_t>(off);
    // Rans64DecInit x 32 (rans64.h:107-115)
    uint64_t x = lds_u32(win.ring | ((cursor + 8 * lane) & (kRingBytes - 1)));
    x |= static_cast<uint64_t>(lds_u32(win.ring | ((cursor + 8 * lane + 4) & (kRingBytes - 1)))) << 32;
    cursor += kRans64HeaderBytes;

    const uint32_t lt = lanemask_lt();
    const uint32_t tab = smem_addr(s_tab);
    const uint32_t syms_lane = tab + (1u << sb) + (lane & (kRans64DecReplicas - 1)) * 8;
    uint8_t* o = out + first + lane;
    const uint32_t steps = m >> 5, rem = m & 31;
    uint32_t g = 0;
    for (; g + 2 <= steps; g += 2) {                            // <= 128 bytes per step: top up every 2 steps
        win.top_up(cursor, lane);
        rans64_dec_step(x, cursor, tab, syms_lane, win.ring, o, lt, sb, true);
        rans64_dec_step(x, cursor, tab, syms_lane, win.ring, o + 32, lt, sb, true);
        o += 64;
    }
    win.top_up(cursor, lane);
    if (g < steps) {
        rans64_dec_step(x, cursor, tab, syms_lane, win.ring, o, lt, sb, true);
        o += 32;
    }
    if (rem) rans64_dec_step(x, cursor, tab, syms_lane, win.ring, o, lt, sb, lane < rem);

    const bool bad = (cursor != static_cast<uint32_t>(end)) || (x != kRans64L);
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(status, kStatStream);
}

struct Rans64EncState {
    uint64_t x;
    uint32_t wpos;     // un-wrapped ring position of the next u32 = 508 - 4 * words_emitted
    uint32_t flags;
};

// Rans64EncPutSymbol for 32 lanes (rans64.h:262-278).  tab_lane = this lane's replica of the 256-entry table
// {rcp_lo, rcp_hi, bias, cmpl_freq | rcp_shift << 20 | bad << 31}: one 16-byte Rans64EncSymbol image per symbol (freq
// is M - cmpl_freq), replicated 8x so that the 8 lanes of a quarter-warp hit 8 different 16-byte bank groups -- ONE
// conflict-free LDS.128 per step.  (Round 1 read the 32-byte image with two LDS.128 from an unreplicated table: ~28
// wavefronts per step, shared-memory data pipe 99 % busy, 3.8 ms per GiB.)
__device__ __forceinline__ void rans64_enc_step(Rans64EncState& st, uint32_t sym, uint32_t tab_lane, uint32_t ring, uint32_t gt, uint32_t sb,
                                                bool active)
{
    bool need = false;
    uint4 e = make_uint4(0, 0, 0, 0);
    if (active) {
        e = lds_u128_ro(tab_lane + sym * (kEncReplicas * 16));
        st.flags |= e.w;
        const uint32_t freq = (1u << sb) - (e.w & 0xfffffu);
        const uint64_t x_max = static_cast<uint64_t>(freq) << (63 - sb);               // ((L >> sb) << 32) * freq, :269
        need = st.x >= x_max;
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, need);
    if (need) {
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(ring | ((st.wpos - 4u * __popc(mask & gt)) & (kEncRingBytes - 1))),
                     "r"(static_cast<uint32_t>(st.x))
                     : "memory");                                                      // :271-272
        st.x >>= 32;                                                                   // :273
    }
    st.wpos -= 4u * __popc(mask);
    if (active) {
        const uint64_t rcp = static_cast<uint64_t>(e.x) | (static_cast<uint64_t>(e.y) << 32);
        const uint64_t q = __umul64hi(st.x, rcp) >> ((e.w >> 20) & 63u);               // :276
        st.x = st.x + e.z + q * (e.w & 0xfffffu);                                      // :277
    }
}

__global__ void __launch_bounds__(kRans64Warps * 32, 2)
rans64_encode_kernel(const uint8_t* __restrict__ in, uint64_t n, uint32_t chunk_syms, uint32_t n_chunks, uint32_t sb,
                     const uint4* __restrict__ g_enc,      // 256 x 2 x uint4
                     uint8_t* __restrict__ scratch, uint32_t slot_bytes, uint32_t* __restrict__ sizes, uint32_t* __restrict__ status)
{
    extern __shared__ __align__(1024) uint8_t s_r64e[];      // [16 x 1 KiB stage + ring][32 KiB table, 8 replicas]
    uint4* s_tab = reinterpret_cast<uint4*>(s_r64e + kRans64Warps * kEncWarpSmem);
    for (uint32_t i = threadIdx.x; i < 256 * kEncReplicas; i += blockDim.x) {
        const uint4 a = g_enc[2 * (i / kEncReplicas)], b = g_enc[2 * (i / kEncReplicas) + 1];   // {rcp_lo, rcp_hi, freq, bias}, {cmpl, shift | bad, -, -}
        s_tab[i] = make_uint4(a.x, a.y, a.w, (b.x & 0xfffffu) | ((b.y & 63u) << 20) | (b.y & kEncBadSymbol));
    }
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t chunk = blockIdx.x * kRans64Warps + warp;
    if (chunk >= n_chunks) return;
    const uint64_t first = static_cast<uint64_t>(chunk) * chunk_syms;
    const uint32_t m = static_cast<uint32_t>(min(static_cast<uint64_t>(chunk_syms), n - first));
    const uint8_t* chunk_in = in + first;
    const uint32_t wsm = smem_addr(s_r64e) + warp * kEncWarpSmem;
    const uint32_t stage = wsm, ring = wsm + kEncStageBytes, tab = smem_addr(s_tab) + (lane & (kEncReplicas - 1)) * 16;
    uint8_t* slot_end = scratch + static_cast<uint64_t>(chunk + 1) * slot_bytes;
    const uint32_t gt = lanemask_gt();

    Rans64EncState st;
    st.x = kRans64L;                            // Rans64EncInit, rans64.h:65-68
    st.wpos = kEncRingBytes - 4;
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
        rans64_enc_step(st, s, tab, ring, gt, sb, active);
    }
    for (uint32_t g = steps; g > nblk * 16; g--) {
        const uint32_t s = chunk_in[static_cast<uint64_t>(g - 1) * 32 + lane];
        rans64_enc_step(st, s, tab, ring, gt, sb, true);
        if (((g - 1) & 1) == 0) word_enc_flush(kEncRingBytes - 4 - st.wpos, flushed, ring, slot_end, lane);
    }
    word_enc_flush(kEncRingBytes - 4 - st.wpos, flushed, ring, slot_end, lane);

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
        for (int pr = 7; pr >= 0; pr--) {                    // <= 128 bytes per step: look at the ring every 2 steps
            rans64_enc_step(st, lds_u8(stage + (pr * 2 + 1) * 32 + lane), tab, ring, gt, sb, true);
            rans64_enc_step(st, lds_u8(stage + (pr * 2) * 32 + lane), tab, ring, gt, sb, true);
            if (kEncRingBytes - 4 - st.wpos - flushed >= 256) word_enc_flush(kEncRingBytes - 4 - st.wpos, flushed, ring, slot_end, lane);
        }
    }

    // Rans64EncFlush for lanes 31..0 (rans64.h:96-103): 256 bytes, in two halves so the 512-byte ring never overflows
    word_enc_flush(kEncRingBytes - 4 - st.wpos, flushed, ring, slot_end, lane);         // < 16 bytes stay pending
#pragma unroll
    for (int half = 1; half >= 0; half--) {
        if ((lane >> 4) == static_cast<uint32_t>(half)) {
            const uint32_t hpos = st.wpos - 8u * (15 - (lane & 15));
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(ring | (hpos & (kEncRingBytes - 1))), "r"(static_cast<uint32_t>(st.x >> 32)) : "memory");
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(ring | ((hpos - 4) & (kEncRingBytes - 1))), "r"(static_cast<uint32_t>(st.x)) : "memory");
        }
        st.wpos -= 128;
        word_enc_flush(kEncRingBytes - 4 - st.wpos, flushed, ring, slot_end, lane);
    }
    const uint32_t produced = kEncRingBytes - 4 - st.wpos;
    const uint32_t left = produced - flushed;                  // < 16 bytes, multiple of 4: head of the stream
    if (4 * lane < left) {
        const uint32_t off = flushed + 4 * lane + 4;
        *reinterpret_cast<uint32_t*>(slot_end - off) = lds_u32(ring | ((0u - off) & (kEncRingBytes - 1)));
    }
    if (lane == 0) sizes[chunk] = produced;
    if (__any_sync(0xffffffffu, (st.flags & kEncBadSymbol) != 0) && lane == 0) atomicOr(status, kStatSymbol);
}

inline void configure_rans64_kernels()
{
    cudaFuncSetAttribute(rans64_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRans64Warps * kRingBytes + kRans64DecSymBytes + (1u << 16));
    cudaFuncSetAttribute(rans64_decode_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(rans64_encode_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(rans64_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRans64Warps * kEncWarpSmem + kEncTableBytes);
}

inline void launch_rans64_decode(cudaStream_t stream, const uint8_t* blob, uint64_t blob_size, const uint64_t* offsets, uint32_t sb,
                                 const uint8_t* table, uint8_t* out, uint64_t n, uint32_t chunk_syms, uint32_t n_chunks, uint32_t* status)
{
    const uint32_t grid = (n_chunks + kRans64Warps - 1) / kRans64Warps;
    rans64_decode_kernel<<<grid, kRans64Warps * 32, kRans64Warps * kRingBytes + kRans64DecSymBytes + (1u << sb), stream>>>(
        blob, blob_size, offsets, sb, table, out, n, chunk_syms, n_chunks, status);
}

inline void launch_rans64_encode(cudaStream_t stream, const uint8_t* d_in, uint64_t n, uint32_t chunk_syms, uint32_t n_chunks, uint32_t sb,
                                 const uint4* enc, uint8_t* scratch, uint32_t slot, uint32_t* sizes, uint32_t* status)
{
    const uint32_t grid = (n_chunks + kRans64Warps - 1) / kRans64Warps;
    rans64_encode_kernel<<<grid, kRans64Warps * 32, kRans64Warps * kEncWarpSmem + kEncTableBytes, stream>>>(d_in, n, chunk_syms, n_chunks, sb, enc,
                                                                                                  scratch, slot, sizes, status);
}

}  // namespace rb200
