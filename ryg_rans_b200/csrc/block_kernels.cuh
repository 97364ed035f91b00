// block_kernels.cuh -- device model construction and the per-block-model variants of the
// word-coder kernels (BASELINE config 5: many independent blocks, each with its own
// 256-entry frequency table, scale_bits 12).
//
//   histogram_kernel       SymbolStats::count_freqs   (main.cpp:59-66) on the device
//   block_model_kernel     count_freqs + calc_cum_freqs + normalize_freqs (main.cpp:59-129)
//                          per block, bit-exact, one warp running the order-dependent
//                          "steal a slot" loop with a warp-wide arg-min
//   block_decode_kernel    RansWordTablesInitSymbol (rans_word_sse41.h:64-72) in shared
//                          memory, then the 32-way decode of each of the block's chunks
//   block_encode_kernel    encoder reciprocal table in shared memory, then 32-way encode
#pragma once
#include "device_utils.cuh"
#include "tables.h"
#include "word_kernels.cuh"

namespace rb200 {

// ---------------------------------------------------------------------------
// K7: byte histogram
// ---------------------------------------------------------------------------
constexpr int kHistWarps = 8;

__global__ void __launch_bounds__(kHistWarps * 32)
histogram_kernel(const uint8_t* __restrict__ in, uint64_t n, unsigned long long* __restrict__ counts)
{
    __shared__ uint32_t s_h[kHistWarps][256];
    for (uint32_t i = threadIdx.x; i < kHistWarps * 256; i += blockDim.x) (&s_h[0][0])[i] = 0;
    __syncthreads();
    uint32_t* h = s_h[threadIdx.x >> 5];

    // unaligned head (< 16 bytes) and tail (< 16 bytes) go to block 0, the 16-byte body is grid-strided
    const uint64_t head = min(n, static_cast<uint64_t>((16 - (reinterpret_cast<uintptr_t>(in) & 15)) & 15));
    const uint64_t nvec = (n - head) / 16;
    const uint4* body = reinterpret_cast<const uint4*>(in + head);
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
    for (uint64_t v = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        const uint4 q = ldg_stream_u128(body + v);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            atomicAdd(&h[w[j] & 0xff], 1u);
            atomicAdd(&h[(w[j] >> 8) & 0xff], 1u);
            atomicAdd(&h[(w[j] >> 16) & 0xff], 1u);
            atomicAdd(&h[w[j] >> 24], 1u);
        }
    }
    if (blockIdx.x == 0) {
        const uint64_t tail_lo = head + nvec * 16;
        for (uint64_t i = threadIdx.x; i < head; i += blockDim.x) atomicAdd(&h[in[i]], 1u);
        for (uint64_t i = tail_lo + threadIdx.x; i < n; i += blockDim.x) atomicAdd(&h[in[i]], 1u);
    }
    __syncthreads();
    for (uint32_t s = threadIdx.x; s < 256; s += blockDim.x) {
        unsigned long long t = 0;
#pragma unroll
        for (int w = 0; w < kHistWarps; w++) t += s_h[w][s];
        if (t) atomicAdd(&counts[s], t);
    }
}

inline void launch_histogram(cudaStream_t stream, const uint8_t* d_in, uint64_t n, unsigned long long* d_counts)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint64_t want = (n / 16 + kHistWarps * 32 - 1) / (kHistWarps * 32);
    // keep each block's per-bin count below 2^32: >= 1 block per 2^31 bytes is plenty
    uint64_t grid = static_cast<uint64_t>(sms) * 8;
    if (want < grid) grid = want ? want : 1;
    histogram_kernel<<<static_cast<uint32_t>(grid), kHistWarps * 32, 0, stream>>>(d_in, n, d_counts);
}

// ---------------------------------------------------------------------------
// per-block model: histogram + normalize_freqs(4096), one CTA per block
// ---------------------------------------------------------------------------
constexpr int kModelWarps = 8;

__global__ void __launch_bounds__(kModelWarps * 32)
block_model_kernel(const uint8_t* __restrict__ in, uint32_t block_size, uint16_t* __restrict__ block_freqs,
                   uint32_t* __restrict__ status)
{
    // one private histogram per warp (more copies per warp were measured slower: the copies share banks)
    __shared__ __align__(16) uint32_t s_h[kModelWarps][256];
    __shared__ uint32_t s_cnt[256];
    __shared__ uint32_t s_cum[257];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t* blk = in + static_cast<uint64_t>(blockIdx.x) * block_size;

    for (uint32_t i = tid; i < kModelWarps * 256; i += blockDim.x) (&s_h[0][0])[i] = 0;
    __syncthreads();
    {   // count_freqs, main.cpp:59-66
        uint32_t* h = s_h[warp];
        const uint32_t head = min(block_size, static_cast<uint32_t>((16 - (reinterpret_cast<uintptr_t>(blk) & 15)) & 15));
        const uint32_t nvec = (block_size - head) / 16;
        const uint4* body = reinterpret_cast<const uint4*>(blk + head);
        for (uint32_t v = tid; v < nvec; v += blockDim.x) {
            const uint4 q = ldg_stream_u128(body + v);
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                atomicAdd(&h[w[j] & 0xff], 1u);
                atomicAdd(&h[(w[j] >> 8) & 0xff], 1u);
                atomicAdd(&h[(w[j] >> 16) & 0xff], 1u);
                atomicAdd(&h[w[j] >> 24], 1u);
            }
        }
        for (uint32_t i = tid; i < head; i += blockDim.x) atomicAdd(&h[blk[i]], 1u);
        for (uint32_t i = head + nvec * 16 + tid; i < block_size; i += blockDim.x) atomicAdd(&h[blk[i]], 1u);
    }
    __syncthreads();
    for (uint32_t s = tid; s < 256; s += blockDim.x) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < kModelWarps; w++) t += s_h[w][s];
        s_cnt[s] = t;
    }
    __syncthreads();
    if (warp != 0) return;

    // calc_cum_freqs (main.cpp:68-73): lane l owns symbols 8l .. 8l+7
    uint32_t run = 0, local[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { local[j] = run; run += s_cnt[8 * lane + j]; }
    uint32_t incl = run;
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= static_cast<uint32_t>(d)) incl += v;
    }
    const uint32_t base = incl - run;
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);      // == block_size
    // resample, main.cpp:83-84 (cum[0] stays 0)
#pragma unroll
    for (int j = 0; j < 8; j++)
        s_cum[8 * lane + j] = static_cast<uint32_t>((static_cast<uint64_t>(kWordSlots) * (base + local[j])) / total);
    if (lane == 31) s_cum[256] = kWordSlots;                        // 4096 * total / total
    __syncwarp();

    // main.cpp:90-116, symbols in order.  Moving the boundaries between donor and taker by one changes exactly
    // two widths (donor - 1, taker + 1), so the loop runs on widths; the arg-min over 256 widths is warp-wide
    // with the key (width << 8 | symbol): lowest width first, lowest symbol on ties, as the reference's scan.
    uint32_t* s_w = s_h[0];                                         // reuse: 256 widths
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t t = 8 * lane + j;
        s_w[t] = s_cum[t + 1] - s_cum[t];
    }
    __syncwarp();
    bool failed = false;
    for (uint32_t s = 0; s < 256; s++) {
        if (s_cnt[s] == 0 || s_w[s] != 0) continue;                 // warp-uniform
        uint32_t key = 0xffffffffu;
        const uint4 lo = *reinterpret_cast<const uint4*>(&s_w[8 * lane]);
        const uint4 hi = *reinterpret_cast<const uint4*>(&s_w[8 * lane + 4]);
        const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (w[j] > 1) key = min(key, (w[j] << 8) | (8 * lane + j));
        key = __reduce_min_sync(0xffffffffu, key);
        if (key == 0xffffffffu) { failed = true; break; }           // main.cpp:104
        __syncwarp();
        if (lane == 0) {
            s_w[key & 0xffu] -= 1;                                  // main.cpp:107-113 in terms of widths
            s_w[s] += 1;
        }
        __syncwarp();
    }
    if (failed && lane == 0) atomicOr(status, kStatStream);
    uint16_t* dst = block_freqs + static_cast<uint64_t>(blockIdx.x) * 256;
#pragma unroll
    for (int j = 0; j < 8; j++) dst[8 * lane + j] = static_cast<uint16_t>(s_w[8 * lane + j]);     // main.cpp:127
}

inline void launch_block_models(cudaStream_t stream, const uint8_t* d_in, uint32_t n_blocks, uint32_t block_size,
                                uint16_t* d_freqs, uint32_t* status)
{
    block_model_kernel<<<n_blocks, kModelWarps * 32, 0, stream>>>(d_in, block_size, d_freqs, status);
}

// ---------------------------------------------------------------------------
// K5: per-block decode / encode.  One CTA per block, one warp per chunk of the block.
// ---------------------------------------------------------------------------
constexpr int kMaxBlockWarps = 32;

// cum[s] for the block's u16 frequencies; returns false if they do not sum to 4096
__device__ __forceinline__ bool block_prefix(const uint16_t* __restrict__ freqs, uint32_t* s_cum, uint32_t* s_bad)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) *s_bad = 0;
    __syncthreads();
    if (tid < 32) {
        uint32_t run = 0, local[8];
#pragma unroll
        for (int j = 0; j < 8; j++) { local[j] = run; run += freqs[8 * lane + j]; }
        uint32_t incl = run;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= static_cast<uint32_t>(d)) incl += v;
        }
#pragma unroll
        for (int j = 0; j < 8; j++) s_cum[8 * lane + j] = incl - run + local[j];
        if (lane == 31) {
            s_cum[256] = incl;
            if (incl != kWordSlots) *s_bad = 1;
        }
    }
    __syncthreads();
    return *s_bad == 0;
}

__global__ void __launch_bounds__(kMaxBlockWarps * 32)
block_decode_kernel(const uint8_t* __restrict__ blob, uint64_t blob_size, const uint64_t* __restrict__ offsets,
                    const uint16_t* __restrict__ block_freqs, uint32_t block_size, uint32_t chunk_syms,
                    uint8_t* __restrict__ out, uint32_t* __restrict__ status)
{
    extern __shared__ __align__(1024) uint8_t s_dyn[];           // [warps][1 KiB] rings
    __shared__ __align__(16) uint32_t s_tab[kWordSlots];
    __shared__ uint32_t s_cum[257];
    __shared__ uint32_t s_flag[2];                               // [0] bad model, [1] wide (a freq of 4096)

    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint16_t* freqs = block_freqs + static_cast<uint64_t>(blockIdx.x) * 256;
    if (tid == 0) s_flag[1] = 0;
    const bool ok = block_prefix(freqs, s_cum, &s_flag[0]);
    if (!ok) {
        if (tid == 0) atomicOr(status, kStatStream);
        return;
    }
    // RansWordTablesInitSymbol for every slot (rans_word_sse41.h:64-72).  Each thread owns a run of
    // consecutive slots: one binary search over the cumulative table for the first, then it walks
    // forward, so the work is balanced however skewed the model is.
    {
        const uint32_t per = kWordSlots / blockDim.x;            // blockDim is a power of two between 128 and 1024
        uint32_t slot = tid * per;
        uint32_t lo = 0, hi = 256;                               // largest s with cum[s] <= slot
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (s_cum[mid] <= slot) lo = mid; else hi = mid;
        }
        for (uint32_t j = 0; j < per; j++, slot++) {
            while (slot >= s_cum[lo + 1]) lo++;                  // also steps over zero-width symbols
            const uint32_t f = s_cum[lo + 1] - s_cum[lo];
            if (f == kWordSlots) s_flag[1] = 1;
            s_tab[slot] = ((f & 0xfffu) << 20) | ((slot - s_cum[lo]) << 8) | lo;
        }
    }
    __syncthreads();

    const uint32_t per_block = block_size / chunk_syms;
    if (warp >= per_block) return;
    const uint32_t chunk = blockIdx.x * per_block + warp;
    uint8_t* dst = out + static_cast<uint64_t>(blockIdx.x) * block_size + static_cast<uint64_t>(warp) * chunk_syms;
    const uint32_t ring = smem_addr(s_dyn) + warp * kRingBytes;
    if (s_flag[1])
        word_decode_chunk<true>(blob, blob_size, offsets, chunk, smem_addr_pinned(s_tab), ring, dst, chunk_syms, status);
    else
        word_decode_chunk<false>(blob, blob_size, offsets, chunk, smem_addr_pinned(s_tab), ring, dst, chunk_syms, status);
}

__global__ void __launch_bounds__(kMaxBlockWarps * 32)
block_encode_kernel(const uint8_t* __restrict__ in, uint32_t block_size, const uint16_t* __restrict__ block_freqs,
                    uint32_t chunk_syms, uint8_t* __restrict__ scratch, uint32_t slot_bytes, uint32_t* __restrict__ sizes,
                    uint32_t* __restrict__ status)
{
    extern __shared__ __align__(1024) uint8_t s_enc[];          // [32 KiB table][warps x 1 KiB stage + ring]
    __shared__ uint32_t s_cum[257];
    __shared__ uint32_t s_flag[1];
    uint4* s_tab = reinterpret_cast<uint4*>(s_enc);
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    const uint16_t* freqs = block_freqs + static_cast<uint64_t>(blockIdx.x) * 256;
    const uint32_t per_block = block_size / chunk_syms;
    const bool ok = block_prefix(freqs, s_cum, &s_flag[0]);
    // 32-bit reciprocal where it is exact for every symbol of this block's model (tables.h: enc32), else the 33-bit one
    bool exact32 = true;
    for (uint32_t s = tid; s < 256; s += blockDim.x) {
        const uint32_t f = ok ? s_cum[s + 1] - s_cum[s] : 0;
        if (f > 1) {
            uint32_t sh = 0;
            while ((1u << sh) < f) sh++;
            const uint64_t pow = 1ull << (31 + sh), M32 = (pow + f - 1) / f;
            if (M32 >> 32 || ((static_cast<uint64_t>(f) << 20) - 1) * (M32 * f - pow) >= pow) exact32 = false;
        }
    }
    const bool r32 = __syncthreads_and(exact32) != 0;
    for (uint32_t s = tid; s < 256; s += blockDim.x) {
        const uint32_t f = ok ? s_cum[s + 1] - s_cum[s] : 0;
        WordEncEntry e = {0u, kEncBadSymbol};
        if (f) {
            uint32_t sh = 0;
            while ((1u << sh) < f) sh++;
            if (!r32) {
                const uint64_t M = ((1ull << (32 + sh)) + f - 1) / f;      // in [2^32, 2^33): keep the low word
                e.magic = static_cast<uint32_t>(M);
                e.packed = f | (s_cum[s] << 13) | (sh << 25);
            } else if (f == 1) {
                e.magic = 0xffffffffu;
                e.packed = f | (s_cum[s] << 13);
            } else {
                e.magic = static_cast<uint32_t>(((1ull << (31 + sh)) + f - 1) / f);
                e.packed = f | (s_cum[s] << 13) | ((sh - 1) << 25);
            }
        }
        const uint4 x = r32 ? word_enc_expand<true>(e) : word_enc_expand<false>(e);
#pragma unroll
        for (uint32_t r = 0; r < kEncReplicas; r++) s_tab[s * kEncReplicas + r] = x;
    }
    __syncthreads();
    if (warp >= per_block) return;
    const uint32_t chunk = blockIdx.x * per_block + warp;
    const uint8_t* src = in + static_cast<uint64_t>(blockIdx.x) * block_size + static_cast<uint64_t>(warp) * chunk_syms;
    const uint32_t wsm = smem_addr(s_enc) + kEncTableBytes + warp * kEncWarpSmem;
    if (r32)
        word_encode_chunk<true>(src, chunk_syms, chunk, smem_addr(s_enc), wsm, scratch, slot_bytes, sizes, status);
    else
        word_encode_chunk<false>(src, chunk_syms, chunk, smem_addr(s_enc), wsm, scratch, slot_bytes, sizes, status);
}

inline uint32_t block_threads(uint32_t block_size, uint32_t chunk_syms)
{
    uint32_t warps = 4;                // table construction wants a few warps even for 1-chunk blocks;
    while (warps < block_size / chunk_syms) warps <<= 1;   // power of two so 4096 slots split evenly
    return warps * 32;
}

inline void configure_block_kernels()
{
    cudaFuncSetAttribute(block_decode_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    // ~17.4 KiB static (table, cumulative, flag) + 1 KiB of ring per warp: above 48 KiB from 31 warps on, so opt in
    cudaFuncSetAttribute(block_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxBlockWarps * kRingBytes);
    cudaFuncSetAttribute(block_encode_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(block_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kEncTableBytes + kMaxBlockWarps * kEncWarpSmem);
}

inline void launch_block_encode(cudaStream_t stream, const uint8_t* d_in, uint32_t n_blocks, uint32_t block_size,
                                const uint16_t* d_freqs, uint32_t chunk_syms, uint8_t* scratch, uint32_t slot,
                                uint32_t* sizes, uint32_t* status)
{
    const uint32_t threads = block_threads(block_size, chunk_syms);
    block_encode_kernel<<<n_blocks, threads, kEncTableBytes + (threads / 32) * kEncWarpSmem, stream>>>(d_in, block_size, d_freqs, chunk_syms,
                                                                                                       scratch, slot, sizes, status);
}

inline void launch_block_decode(cudaStream_t stream, const uint8_t* blob, uint64_t blob_size, const uint64_t* offsets,
                                const uint16_t* d_freqs, uint32_t n_blocks, uint32_t block_size, uint32_t chunk_syms,
                                uint8_t* out, uint32_t* status)
{
    const uint32_t threads = block_threads(block_size, chunk_syms);
    block_decode_kernel<<<n_blocks, threads, (threads / 32) * kRingBytes, stream>>>(blob, blob_size, offsets, d_freqs, block_size,
                                                                                    chunk_syms, out, status);
}

}  // namespace rb200
