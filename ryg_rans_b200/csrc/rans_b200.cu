// rans_b200.cu -- C-ABI (include/rans_b200.h) over the sm_100a rANS kernels.
//
// Host code is C++ like the reference's drivers; it owns no global state: everything
// lives in rb200_ctx (device, stream, grow-only workspaces) and rb200_model (device
// tables).  Data pointers are either host memory (the call stages them through the
// context's device buffers and is synchronous, like the reference's loops) or device
// memory (the call only enqueues kernels on the context's stream).
#include "rans_b200.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "alias_kernels.cuh"
#include "block_kernels.cuh"
#include "rans64_kernels.cuh"
#include "tables.h"
#include "word_kernels.cuh"
#include "word_decode_tma.cuh"

using namespace rb200;

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct rb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint32_t* d_status = nullptr;     // kStat* bits, OR-ed by kernels
    uint32_t* h_status = nullptr;     // pinned mirror
    DecodeWork* d_work = nullptr;     // persistent decoder's chunk counter; zero between launches (self-resetting)
    int sms = 0;                      // multiprocessors of `device`
    uint64_t launches = 0;
    std::string last_error;
    // encode workspaces
    DevBuf scratch, sizes;
    // staging for RB200_MEM_HOST calls
    DevBuf st_in, st_blob, st_offsets, st_out, st_aux;
    // host-mode slice pipeline: H2D on s_in, kernels on `stream`, D2H on s_out
    cudaStream_t s_in = nullptr, s_out = nullptr;
    static constexpr int kMaxSlices = 256;
    cudaEvent_t ev_in[kMaxSlices] = {}, ev_done[kMaxSlices] = {};
    cudaEvent_t ev_start = nullptr, ev_out = nullptr;
    uint64_t* h_slice_total = nullptr;   // pinned, kMaxSlices entries
    uint64_t* h_dir = nullptr;           // pinned staging for directories (caller's `offsets` may be pageable)
    size_t h_dir_cap = 0;                // entries
};

struct rb200_model {
    rb200_ctx* ctx = nullptr;
    int coder = 0;
    uint32_t scale_bits = 0;
    int wide = 0;
    uint32_t* d_word_dec = nullptr;          // 4096 x u32
    WordEncEntry* d_word_enc = nullptr;      // 256; the 32-bit-reciprocal table when enc_r32 (tables.h)
    int enc_r32 = 0;
    AliasDecEntry* d_alias_dec = nullptr;    // 256 x 16 B
    AliasEncEntry* d_alias_enc = nullptr;    // 256
    uint16_t* d_alias_remap = nullptr;       // 1 << scale_bits
    uint8_t* d_byte_dec = nullptr;           // cum2sym[1 << scale_bits] + 256 x u32
    AliasEncEntry* d_byte_enc = nullptr;     // 256 RansEncSymbol images
    uint8_t* d_r64_dec = nullptr;            // cum2sym[1 << scale_bits] + 256 x {start, freq}
    uint4* d_r64_enc = nullptr;              // 256 x 32 B Rans64EncSymbol images
};

namespace {

int cuda_fail(rb200_ctx* ctx, cudaError_t e, const char* what)
{
    if (ctx) ctx->last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return RB200_E_CUDA;
}

#define RB_CUDA(ctx, call)                                              \
    do {                                                                \
        cudaError_t e_ = (call);                                        \
        if (e_ != cudaSuccess) return cuda_fail((ctx), e_, #call);      \
    } while (0)

int reserve(rb200_ctx* ctx, DevBuf& b, size_t bytes)
{
    if (bytes <= b.cap) return RB200_OK;
    if (b.p) {
        RB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        RB_CUDA(ctx, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        b.p = nullptr;
        cuda_fail(ctx, e, "cudaMalloc(workspace)");
        return RB200_E_NOMEM;
    }
    b.cap = want;
    return RB200_OK;
}

void release(DevBuf& b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

inline size_t round16(size_t v) { return (v + 15) & ~static_cast<size_t>(15); }
constexpr size_t kBoundFixed = 512;   // headers (128 B; 256 B for rans64) + one extra unit per lane
inline uint32_t slot_bytes_for(uint32_t chunk_syms) { return static_cast<uint32_t>(round16(kBoundFixed + 2ull * chunk_syms)); }

int status_to_code(uint32_t bits)
{
    if (bits & kStatSymbol) return RB200_E_SYMBOL;
    if (bits & kStatSpace) return RB200_E_SPACE;
    if (bits & kStatStream) return RB200_E_STREAM;
    if (bits & kStatStall) return RB200_E_STALL;
    return RB200_OK;
}

int check_launch(rb200_ctx* ctx, const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(ctx, e, what);
    ctx->launches++;
    return RB200_OK;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

constexpr uint32_t kMinChunk = 32;
constexpr uint32_t kMaxChunk = 1u << 24;

bool chunk_ok(uint32_t chunk_syms) { return chunk_syms >= kMinChunk && chunk_syms <= kMaxChunk; }

}  // namespace

extern "C" int rb200_version(void) { return RB200_VERSION; }

extern "C" const char* rb200_strerror(int code)
{
    switch (code) {
    case RB200_OK: return "ok";
    case RB200_E_ARG: return "bad argument";
    case RB200_E_MODEL: return "invalid model";
    case RB200_E_SPACE: return "output buffer too small";
    case RB200_E_STREAM: return "corrupt or truncated stream";
    case RB200_E_CUDA: return "CUDA error";
    case RB200_E_NOMEM: return "out of device memory";
    case RB200_E_SYMBOL: return "symbol with zero model frequency";
    case RB200_E_NCCL: return "NCCL error";
    case RB200_E_STALL: return "a bounded wait inside a kernel expired";
    }
    return "unknown";
}

extern "C" int rb200_ctx_create(rb200_ctx** out, int device, void* stream)
{
    if (!out) return RB200_E_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return RB200_E_CUDA;
    rb200_ctx* ctx = new (std::nothrow) rb200_ctx;
    if (!ctx) return RB200_E_NOMEM;
    ctx->device = device;
    ctx->stream = static_cast<cudaStream_t>(stream);
    DeviceGuard g(device);
    cudaError_t e = cudaMalloc(&ctx->d_status, sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_status, 0, sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_work, sizeof(DecodeWork));
    if (e == cudaSuccess) e = cudaMemset(ctx->d_work, 0, sizeof(DecodeWork));
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sms, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_status, sizeof(uint32_t));
    if (e != cudaSuccess || ctx->sms <= 0) {
        if (ctx->d_status) cudaFree(ctx->d_status);
        if (ctx->d_work) cudaFree(ctx->d_work);
        delete ctx;
        return RB200_E_CUDA;
    }
    *ctx->h_status = 0;
    e = cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_out, cudaEventDisableTiming);
    for (int i = 0; i < rb200_ctx::kMaxSlices && e == cudaSuccess; i++) {
        e = cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaMallocHost(&ctx->h_slice_total, rb200_ctx::kMaxSlices * sizeof(uint64_t));
    if (e != cudaSuccess) {
        rb200_ctx_destroy(ctx);
        return RB200_E_CUDA;
    }
    // the decoders want the large shared-memory carve-out (tables + per-warp rings)
    cudaFuncSetAttribute(word_decode_tma_kernel<DecShip, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(word_decode_tma_kernel<DecShip, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DecShip::kSmemBytes);
    cudaFuncSetAttribute(word_decode_tma_kernel<DecShip, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(word_decode_tma_kernel<DecShip, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DecShip::kSmemBytes);
    cudaFuncSetAttribute(word_decode_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(word_decode_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(word_encode_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(word_encode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEncSmemBytes);
    cudaFuncSetAttribute(word_encode_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(word_encode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEncSmemBytes);
    cudaFuncSetAttribute(word_encode_fused_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(word_encode_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEncSmemBytes);
    cudaFuncSetAttribute(word_encode_fused_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(word_encode_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEncSmemBytes);
    configure_alias_kernels();
    configure_block_kernels();
    configure_rans64_kernels();
    cudaGetLastError();
    *out = ctx;
    return RB200_OK;
}

extern "C" void rb200_ctx_destroy(rb200_ctx* ctx)
{
    if (!ctx) return;
    DeviceGuard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    release(ctx->scratch); release(ctx->sizes);
    release(ctx->st_in); release(ctx->st_blob); release(ctx->st_offsets); release(ctx->st_out); release(ctx->st_aux);
    if (ctx->d_status) cudaFree(ctx->d_status);
    if (ctx->d_work) cudaFree(ctx->d_work);
    if (ctx->h_status) cudaFreeHost(ctx->h_status);
    if (ctx->h_slice_total) cudaFreeHost(ctx->h_slice_total);
    if (ctx->h_dir) cudaFreeHost(ctx->h_dir);
    for (int i = 0; i < rb200_ctx::kMaxSlices; i++) {
        if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
        if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
    }
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    if (ctx->ev_out) cudaEventDestroy(ctx->ev_out);
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    delete ctx;
}

extern "C" int rb200_ctx_set_stream(rb200_ctx* ctx, void* stream)
{
    if (!ctx) return RB200_E_ARG;
    ctx->stream = static_cast<cudaStream_t>(stream);
    return RB200_OK;
}

extern "C" const char* rb200_last_cuda_error(const rb200_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }
extern "C" uint64_t rb200_launch_count(const rb200_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int rb200_sync(rb200_ctx* ctx)
{
    if (!ctx) return RB200_E_ARG;
    DeviceGuard g(ctx->device);
    RB_CUDA(ctx, cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    RB_CUDA(ctx, cudaMemsetAsync(ctx->d_status, 0, sizeof(uint32_t), ctx->stream));
    RB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*ctx->h_status) RB_CUDA(ctx, cudaMemsetAsync(ctx->d_work, 0, sizeof(DecodeWork), ctx->stream));   // a kernel may have bailed out
    return status_to_code(*ctx->h_status);
}

// ---------------------------------------------------------------------------
// models
// ---------------------------------------------------------------------------

extern "C" void rb200_model_destroy(rb200_model* m)
{
    if (!m) return;
    DeviceGuard g(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    cudaFree(m->d_word_dec); cudaFree(m->d_word_enc);
    cudaFree(m->d_alias_dec); cudaFree(m->d_alias_enc); cudaFree(m->d_alias_remap);
    cudaFree(m->d_byte_dec); cudaFree(m->d_byte_enc);
    cudaFree(m->d_r64_dec); cudaFree(m->d_r64_enc);
    delete m;
}

namespace {
// RB200_WORD_RECIPROCAL=33 forces the any-x (33-bit) reciprocal even where the 32-bit one is exact (testing)
bool word_r32_allowed()
{
    static const bool allowed = [] {
        const char* e = std::getenv("RB200_WORD_RECIPROCAL");
        return !(e && std::strcmp(e, "33") == 0);
    }();
    return allowed;
}
}  // namespace

extern "C" int rb200_model_create(rb200_ctx* ctx, int coder, uint32_t scale_bits, const uint32_t freqs[256], rb200_model** out)
{
    if (!ctx || !freqs || !out) return RB200_E_ARG;
    *out = nullptr;
    DeviceGuard g(ctx->device);
    rb200_model* m = new (std::nothrow) rb200_model;
    if (!m) return RB200_E_NOMEM;
    m->ctx = ctx;
    m->coder = coder;
    m->scale_bits = scale_bits;
    int rc = RB200_OK;
    cudaError_t e = cudaSuccess;
    if (coder == RB200_CODER_WORD) {
        if (scale_bits != kWordScaleBits) { delete m; return RB200_E_ARG; }
        WordDeviceTables* t = new (std::nothrow) WordDeviceTables;
        if (!t) { delete m; return RB200_E_NOMEM; }
        rc = build_word_device_tables(freqs, *t);
        if (rc == RB200_OK) {
            m->wide = t->wide;
            m->enc_r32 = t->enc32_ok && word_r32_allowed();
            e = cudaMalloc(&m->d_word_dec, sizeof t->dec);
            if (e == cudaSuccess) e = cudaMalloc(&m->d_word_enc, sizeof t->enc);
            if (e == cudaSuccess) e = cudaMemcpy(m->d_word_dec, t->dec, sizeof t->dec, cudaMemcpyHostToDevice);
            if (e == cudaSuccess)
                e = cudaMemcpy(m->d_word_enc, m->enc_r32 ? t->enc32 : t->enc, sizeof t->enc, cudaMemcpyHostToDevice);
        }
        delete t;
    } else if (coder == RB200_CODER_ALIAS) {
        AliasDeviceTables* t = new (std::nothrow) AliasDeviceTables;
        if (!t) { delete m; return RB200_E_NOMEM; }
        rc = build_alias_device_tables(freqs, scale_bits, *t);
        if (rc == RB200_OK) {
            const size_t remap_bytes = t->remap.size() * sizeof(uint16_t);
            e = cudaMalloc(&m->d_alias_dec, sizeof t->dec);
            if (e == cudaSuccess) e = cudaMalloc(&m->d_alias_enc, sizeof t->enc);
            if (e == cudaSuccess) e = cudaMalloc(&m->d_alias_remap, remap_bytes);
            if (e == cudaSuccess) e = cudaMemcpy(m->d_alias_dec, t->dec, sizeof t->dec, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemcpy(m->d_alias_enc, t->enc, sizeof t->enc, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemcpy(m->d_alias_remap, t->remap.data(), remap_bytes, cudaMemcpyHostToDevice);
        }
        delete t;
    } else if (coder == RB200_CODER_BYTE) {
        ByteDeviceTables* t = new (std::nothrow) ByteDeviceTables;
        if (!t) { delete m; return RB200_E_NOMEM; }
        rc = build_byte_device_tables(freqs, scale_bits, *t);
        if (rc == RB200_OK) {
            e = cudaMalloc(&m->d_byte_dec, t->dec.size());
            if (e == cudaSuccess) e = cudaMalloc(&m->d_byte_enc, sizeof t->enc);
            if (e == cudaSuccess) e = cudaMemcpy(m->d_byte_dec, t->dec.data(), t->dec.size(), cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemcpy(m->d_byte_enc, t->enc, sizeof t->enc, cudaMemcpyHostToDevice);
        }
        delete t;
    } else if (coder == RB200_CODER_RANS64) {
        Rans64DeviceTables* t = new (std::nothrow) Rans64DeviceTables;
        if (!t) { delete m; return RB200_E_NOMEM; }
        rc = build_rans64_device_tables(freqs, scale_bits, *t);
        if (rc == RB200_OK) {
            e = cudaMalloc(&m->d_r64_dec, t->dec.size());
            if (e == cudaSuccess) e = cudaMalloc(&m->d_r64_enc, sizeof t->enc);
            if (e == cudaSuccess) e = cudaMemcpy(m->d_r64_dec, t->dec.data(), t->dec.size(), cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = cudaMemcpy(m->d_r64_enc, t->enc, sizeof t->enc, cudaMemcpyHostToDevice);
        }
        delete t;
    } else {
        delete m;
        return RB200_E_ARG;
    }
    if (rc != RB200_OK || e != cudaSuccess) {
        if (e != cudaSuccess) rc = cuda_fail(ctx, e, "model upload");
        rb200_model_destroy(m);
        return rc;
    }
    *out = m;
    return RB200_OK;
}

// ---------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------

extern "C" size_t rb200_chunk_count(size_t n, uint32_t chunk_syms)
{
    if (!chunk_syms) return 0;
    return n / chunk_syms + (n % chunk_syms != 0);      // no wrap-around for any n
}

// A chunk stream never exceeds 512 + 2 * symbols bytes, whatever the coder: the word coder emits
// <= one u16 per symbol (rans_word_sse41.h:85-89) after a 128-byte header; the byte coders emit
// <= two bytes per symbol at scale_bits <= 16 (state < 2^31, x_max >= 2^15; rans_byte.h:64-70);
// rans64 emits one u32 per 32 bits of information, <= scale_bits/8 <= 2 bytes per symbol plus at
// most one extra word per lane, after a 256-byte header.  Streams are padded to multiples of 16.
extern "C" size_t rb200_encode_bound(size_t n, uint32_t chunk_syms)
{
    if (!chunk_syms) return 0;
    const size_t full = n / chunk_syms, tail = n % chunk_syms;
    size_t b = full * round16(kBoundFixed + 2ull * chunk_syms);
    if (tail) b += round16(kBoundFixed + 2ull * tail);
    return b;
}

// ---------------------------------------------------------------------------
// device-resident hot path
// ---------------------------------------------------------------------------

namespace {


// directory + compaction after any encode kernel: sizes[] -> offsets[] + blob
int finish_encode(rb200_ctx* ctx, uint8_t* scratch, uint32_t slot, uint32_t* sizes, uint64_t* tile_sums, uint32_t n_chunks,
                  uint8_t* d_blob, size_t blob_cap, uint64_t* d_offsets)
{
    if (!n_chunks) {
        RB_CUDA(ctx, cudaMemsetAsync(d_offsets, 0, sizeof(uint64_t), ctx->stream));
        return RB200_OK;
    }
    const uint32_t tiles = (n_chunks + kScanTile - 1) / kScanTile;
    scan_tiles_kernel<<<tiles, 1024, 0, ctx->stream>>>(sizes, n_chunks, d_offsets, tile_sums);
    int rc = check_launch(ctx, "scan_tiles_kernel");
    if (rc != RB200_OK) return rc;
    const uint32_t prefixed = tiles > kTilePrefixThreshold ? 1u : 0u;
    if (prefixed) {
        tile_prefix_kernel<<<1, 1024, 0, ctx->stream>>>(tile_sums, tiles);
        rc = check_launch(ctx, "tile_prefix_kernel");
        if (rc != RB200_OK) return rc;
    }
    const uint32_t grid = (n_chunks + kCopyWarps - 1) / kCopyWarps;
    compact_kernel<<<grid, kCopyWarps * 32, 0, ctx->stream>>>(scratch, slot, sizes, d_offsets, tile_sums, prefixed, n_chunks, d_blob,
                                                               blob_cap, ctx->d_status);
    return check_launch(ctx, "compact_kernel");
}

int reserve_encode_workspace(rb200_ctx* ctx, uint32_t n_chunks, uint32_t slot, uint8_t** scratch, uint32_t** sizes, uint64_t** tile_sums)
{
    int rc = reserve(ctx, ctx->scratch, static_cast<size_t>(n_chunks) * slot + 16);
    if (rc != RB200_OK) return rc;
    const size_t sizes_bytes = round16((static_cast<size_t>(n_chunks) + 1) * sizeof(uint32_t));
    const size_t tiles = (static_cast<size_t>(n_chunks) + kScanTile - 1) / kScanTile + 1;
    rc = reserve(ctx, ctx->sizes, sizes_bytes + tiles * sizeof(uint64_t));
    if (rc != RB200_OK) return rc;
    *scratch = static_cast<uint8_t*>(ctx->scratch.p);
    *sizes = static_cast<uint32_t*>(ctx->sizes.p);
    *tile_sums = reinterpret_cast<uint64_t*>(static_cast<uint8_t*>(ctx->sizes.p) + sizes_bytes);
    return RB200_OK;
}

// Word-coder encode has two paths:
//   fused: ONE persistent launch -- workers encode, a scanner warp turns published sizes into end offsets
//          in chunk order, workers place chunk k after encoding chunk k+1 (word_kernels.cuh, K2f);
//   split: three launches -- encode into per-chunk worst-case slots, tile scan, compaction.
// fused is ~10 % faster per GiB on B200 (1.43 vs 1.58 ms) as long as the scanner keeps up, i.e. for
// chunks of >= 4096 symbols; below that the split path is used.  RB200_ENCODE_PATH=split|fused forces one.
bool use_fused_encode(uint32_t chunk_syms)
{
    static const int forced = [] {           // read once; thread-safe static initialisation
        const char* e = std::getenv("RB200_ENCODE_PATH");
        return !e ? -1 : (std::strcmp(e, "fused") == 0 ? 1 : (std::strcmp(e, "split") == 0 ? 0 : -1));
    }();
    // two worst-case slots per resident warp: keep that scratch below ~2 GB (chunks of <= 64 Ki symbols)
    if (chunk_syms > 65536) return false;
    if (forced >= 0) return forced == 1;
    return chunk_syms >= 4096;
}

// RB200_DECODE_PATH=classic selects the round-1 decoder (one CTA per 8 chunks, LDG -> STS window) for A/B runs;
// the default is the persistent kernel of word_decode_tma.cuh.
bool use_persistent_decode()
{
    static const bool persistent = [] {
        const char* e = std::getenv("RB200_DECODE_PATH");
        return !(e && std::strcmp(e, "classic") == 0);
    }();
    return persistent;
}

uint32_t fused_word_grid(rb200_ctx* ctx, uint32_t n_chunks)
{
    const uint32_t want = (n_chunks + 1 + kEncWarps - 1) / kEncWarps;       // + 1: one warp of the grid is the scanner
    const uint32_t grid = static_cast<uint32_t>(ctx->sms) * RB200_ENC_MINBLOCKS;
    return grid > want ? want : grid;
}

int encode_word_fused(rb200_ctx* ctx, const rb200_model* model, const uint8_t* d_in, size_t n, uint32_t chunk_syms, uint32_t n_chunks,
                      uint8_t* d_blob, size_t blob_cap, uint64_t* d_offsets)
{
    const uint32_t slot = slot_bytes_for(chunk_syms);
    const uint32_t grid = fused_word_grid(ctx, n_chunks);
    int rc = reserve(ctx, ctx->scratch, static_cast<size_t>(grid) * kEncWarps * kFusedSlots * slot + 16);  // kFusedSlots per resident warp
    if (rc == RB200_OK) rc = reserve(ctx, ctx->sizes, 16 + static_cast<size_t>(n_chunks) * sizeof(uint64_t));
    if (rc != RB200_OK) return rc;
    uint32_t* counter = static_cast<uint32_t*>(ctx->sizes.p);
    uint64_t* look = reinterpret_cast<uint64_t*>(static_cast<uint8_t*>(ctx->sizes.p) + 16);
    RB_CUDA(ctx, cudaMemsetAsync(ctx->sizes.p, 0, 16 + static_cast<size_t>(n_chunks) * sizeof(uint64_t), ctx->stream));
    auto kernel = model->enc_r32 ? word_encode_fused_kernel<true> : word_encode_fused_kernel<false>;
    kernel<<<grid, kEncWarps * 32, kEncSmemBytes, ctx->stream>>>(d_in, n, chunk_syms, n_chunks, model->d_word_enc,
                                                                 static_cast<uint8_t*>(ctx->scratch.p), slot, look, counter, d_blob,
                                                                 blob_cap, d_offsets, ctx->d_status);
    return check_launch(ctx, "word_encode_fused_kernel");
}

// Grows the encode workspaces to what a call over `n_chunks` chunks needs (the host pipeline does this once
// for its largest slice so that no slice has to stall on a reallocation).
int reserve_encode(rb200_ctx* ctx, const rb200_model* model, uint32_t n_chunks, uint32_t chunk_syms)
{
    if (!n_chunks) return RB200_OK;
    const uint32_t slot = slot_bytes_for(chunk_syms);
    if (use_fused_encode(chunk_syms) && model->coder != RB200_CODER_RANS64) {
        const size_t slots = model->coder == RB200_CODER_WORD ? static_cast<size_t>(fused_word_grid(ctx, n_chunks)) * kEncWarps * kFusedSlots
                                                              : static_cast<size_t>(alias_fused_slots(static_cast<uint32_t>(ctx->sms)));
        int rc = reserve(ctx, ctx->scratch, slots * slot + 16);
        if (rc == RB200_OK) rc = reserve(ctx, ctx->sizes, 16 + static_cast<size_t>(n_chunks) * sizeof(uint64_t));
        return rc;
    }
    uint8_t* scratch; uint32_t* sizes; uint64_t* tile_sums;
    return reserve_encode_workspace(ctx, n_chunks, slot, &scratch, &sizes, &tile_sums);
}

int encode_device(rb200_ctx* ctx, const rb200_model* model, const uint8_t* d_in, size_t n, uint32_t chunk_syms,
                  uint8_t* d_blob, size_t blob_cap, uint64_t* d_offsets)
{
    const size_t n_chunks_sz = rb200_chunk_count(n, chunk_syms);
    if (n_chunks_sz >= (1ull << 31)) return RB200_E_ARG;
    const uint32_t n_chunks = static_cast<uint32_t>(n_chunks_sz);
    const bool fused = n_chunks && use_fused_encode(chunk_syms);
    if (fused && model->coder == RB200_CODER_WORD)
        return encode_word_fused(ctx, model, d_in, n, chunk_syms, n_chunks, d_blob, blob_cap, d_offsets);
    const uint32_t slot = slot_bytes_for(chunk_syms);
    if (fused && (model->coder == RB200_CODER_ALIAS || model->coder == RB200_CODER_BYTE)) {
        int rc = reserve(ctx, ctx->scratch, static_cast<size_t>(alias_fused_slots(static_cast<uint32_t>(ctx->sms))) * slot + 16);
        if (rc == RB200_OK) rc = reserve(ctx, ctx->sizes, 16 + static_cast<size_t>(n_chunks) * sizeof(uint64_t));
        if (rc != RB200_OK) return rc;
        uint32_t* counter = static_cast<uint32_t*>(ctx->sizes.p);
        uint64_t* look = reinterpret_cast<uint64_t*>(static_cast<uint8_t*>(ctx->sizes.p) + 16);
        RB_CUDA(ctx, cudaMemsetAsync(ctx->sizes.p, 0, 16 + static_cast<size_t>(n_chunks) * sizeof(uint64_t), ctx->stream));
        const bool alias = model->coder == RB200_CODER_ALIAS;
        launch_alias_encode(ctx->stream, static_cast<uint32_t>(ctx->sms), d_in, n, chunk_syms, n_chunks, model->scale_bits, alias ? model->d_alias_enc : model->d_byte_enc,
                            alias ? model->d_alias_remap : nullptr, static_cast<uint8_t*>(ctx->scratch.p), slot, nullptr, look, counter,
                            d_blob, blob_cap, d_offsets, ctx->d_status);
        return check_launch(ctx, alias ? "alias_encode_kernel(fused)" : "byte_encode_kernel(fused)");
    }
    uint8_t* scratch; uint32_t* sizes; uint64_t* tile_sums;
    int rc = reserve_encode_workspace(ctx, n_chunks, slot, &scratch, &sizes, &tile_sums);
    if (rc != RB200_OK) return rc;
    if (n_chunks) {
        if (model->coder == RB200_CODER_WORD) {
            const uint32_t grid = (n_chunks + kEncWarps - 1) / kEncWarps;
            auto kernel = model->enc_r32 ? word_encode_kernel<true> : word_encode_kernel<false>;
            kernel<<<grid, kEncWarps * 32, kEncSmemBytes, ctx->stream>>>(d_in, n, chunk_syms, n_chunks, model->d_word_enc, scratch, slot,
                                                                         sizes, ctx->d_status);
            rc = check_launch(ctx, "word_encode_kernel");
        } else if (model->coder == RB200_CODER_ALIAS) {
            rc = launch_alias_encode(ctx->stream, static_cast<uint32_t>(ctx->sms), d_in, n, chunk_syms, n_chunks, model->scale_bits, model->d_alias_enc,
                                     model->d_alias_remap, scratch, slot, sizes, nullptr, nullptr, nullptr, 0, nullptr, ctx->d_status);
            if (rc == RB200_OK) rc = check_launch(ctx, "alias_encode_kernel");
        } else if (model->coder == RB200_CODER_BYTE) {
            rc = launch_alias_encode(ctx->stream, static_cast<uint32_t>(ctx->sms), d_in, n, chunk_syms, n_chunks, model->scale_bits, model->d_byte_enc, nullptr, scratch,
                                     slot, sizes, nullptr, nullptr, nullptr, 0, nullptr, ctx->d_status);
            if (rc == RB200_OK) rc = check_launch(ctx, "byte_encode_kernel");
        } else {
            launch_rans64_encode(ctx->stream, d_in, n, chunk_syms, n_chunks, model->scale_bits, model->d_r64_enc, scratch, slot, sizes,
                                 ctx->d_status);
            rc = check_launch(ctx, "rans64_encode_kernel");
        }
        if (rc != RB200_OK) return rc;
    }
    return finish_encode(ctx, scratch, slot, sizes, tile_sums, n_chunks, d_blob, blob_cap, d_offsets);
}

int decode_device(rb200_ctx* ctx, const rb200_model* model, const uint8_t* d_blob, size_t blob_size, const uint64_t* d_offsets,
                  uint32_t chunk_syms, uint8_t* d_out, size_t n)
{
    const size_t n_chunks_sz = rb200_chunk_count(n, chunk_syms);
    if (n_chunks_sz >= (1ull << 31)) return RB200_E_ARG;
    const uint32_t n_chunks = static_cast<uint32_t>(n_chunks_sz);
    if (!n_chunks) return RB200_OK;
    if (model->coder == RB200_CODER_WORD) {
        if (use_persistent_decode()) {
            // persistent grid: 2 CTAs of 32 warps per SM, chunks handed out by an atomic counter (word_decode_tma.cuh)
            const uint32_t want = (n_chunks + DecShip::kWarps - 1) / DecShip::kWarps;
            const uint32_t full = static_cast<uint32_t>(ctx->sms) * DecShip::kMinBlocks;
            const uint32_t grid = want < full ? want : full;
            auto kernel = model->wide ? word_decode_tma_kernel<DecShip, true> : word_decode_tma_kernel<DecShip, false>;
            kernel<<<grid, DecShip::kWarps * 32, DecShip::kSmemBytes, ctx->stream>>>(d_blob, blob_size, d_offsets, model->d_word_dec, d_out, n,
                                                                                     chunk_syms, n_chunks, ctx->d_work, ctx->d_status, 0);
            return check_launch(ctx, "word_decode_tma_kernel");
        }
        const uint32_t grid = (n_chunks + kDecWarps - 1) / kDecWarps;
        if (model->wide)
            word_decode_kernel<true><<<grid, kDecWarps * 32, 0, ctx->stream>>>(d_blob, blob_size, d_offsets, model->d_word_dec, d_out,
                                                                                n, chunk_syms, n_chunks, ctx->d_status);
        else
            word_decode_kernel<false><<<grid, kDecWarps * 32, 0, ctx->stream>>>(d_blob, blob_size, d_offsets, model->d_word_dec, d_out,
                                                                                 n, chunk_syms, n_chunks, ctx->d_status);
        return check_launch(ctx, "word_decode_kernel");
    }
    if (model->coder == RB200_CODER_RANS64) {
        launch_rans64_decode(ctx->stream, d_blob, blob_size, d_offsets, model->scale_bits, model->d_r64_dec, d_out, n, chunk_syms,
                             n_chunks, ctx->d_status);
        return check_launch(ctx, "rans64_decode_kernel");
    }
    if (model->coder == RB200_CODER_BYTE) {
        launch_byte_decode(ctx->stream, d_blob, blob_size, d_offsets, model->scale_bits, model->d_byte_dec, d_out, n, chunk_syms,
                           n_chunks, ctx->d_status);
        return check_launch(ctx, "byte_decode_kernel");
    }
    int rc = launch_alias_decode(ctx->stream, d_blob, blob_size, d_offsets, model->scale_bits, model->d_alias_dec, d_out, n,
                                 chunk_syms, n_chunks, ctx->d_status, use_persistent_decode() ? ctx->d_work : nullptr,
                                 static_cast<uint32_t>(ctx->sms));
    if (rc == RB200_OK) rc = check_launch(ctx, "alias_decode_kernel");
    return rc;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

namespace {

int reserve_dir(rb200_ctx* ctx, size_t entries)
{
    if (entries <= ctx->h_dir_cap) return RB200_OK;
    if (ctx->h_dir) {
        RB_CUDA(ctx, cudaStreamSynchronize(ctx->s_in));
        RB_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));
        RB_CUDA(ctx, cudaFreeHost(ctx->h_dir));
        ctx->h_dir = nullptr;
        ctx->h_dir_cap = 0;
    }
    const size_t want = entries + entries / 8 + 64;
    cudaError_t e = cudaMallocHost(&ctx->h_dir, want * sizeof(uint64_t));
    if (e != cudaSuccess) {
        ctx->h_dir = nullptr;
        cuda_fail(ctx, e, "cudaMallocHost(directory staging)");
        return RB200_E_NOMEM;
    }
    ctx->h_dir_cap = want;
    return RB200_OK;
}

// An error exit of a host pipeline must not leave copies in flight: they target the caller's buffers and the context's
// staging, which a later reserve() may free.  Also clears whatever the kernels flagged.
int fail_pipeline(rb200_ctx* ctx, int rc)
{
    cudaStreamSynchronize(ctx->s_in);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->s_out);
    cudaMemsetAsync(ctx->d_status, 0, sizeof(uint32_t), ctx->stream);
    cudaMemsetAsync(ctx->d_work, 0, sizeof(DecodeWork), ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    return rc;
}

// One pipeline slice of a host-mode call: chunks [c0, c0 + cnt) = symbols [lo, lo + len).
struct Slice {
    size_t c0, cnt, lo, len;
};

// Slice plan for the host pipeline.  Steady-state slices are `RB200_SLICE_MIB` (default 32 MiB) of symbols:
// big enough that PCIe time dwarfs launch overhead and the per-slice host wake-up, small enough to overlap.
// The first and last three slices ramp (1/8, 1/4, 1/2 of that) so that the pipeline fill (first H2D) and
// drain (last D2H), which nothing overlaps, are short.  Returns the number of slices (<= kMaxSlices).
size_t plan_slices(size_t n, uint32_t chunk_syms, Slice* out)
{
    static const size_t tuned = [] {        // tuning knob for the host pipeline (1..1024 MiB)
        const char* e = std::getenv("RB200_SLICE_MIB");
        const long v = e ? std::atol(e) : 0;
        return (v >= 1 && v <= 1024) ? static_cast<size_t>(v) << 20 : static_cast<size_t>(32u << 20);
    }();
    const size_t total = rb200_chunk_count(n, chunk_syms);
    size_t tc = (tuned + chunk_syms - 1) / chunk_syms;                // chunks per steady-state slice
    if (tc == 0) tc = 1;
    const size_t room = rb200_ctx::kMaxSlices - 8;
    if (total / tc >= room) tc = total / room + 1;
    size_t ramp[3] = {tc / 8, tc / 4, tc / 2};
    const bool ramped = total >= 4 * tc && ramp[0] > 0;
    size_t k = 0, c = 0;
    auto push = [&](size_t cnt) {
        if (cnt == 0) return;
        if (cnt > total - c) cnt = total - c;
        const size_t lo = c * chunk_syms;
        const size_t len = (n - lo < cnt * chunk_syms) ? n - lo : cnt * chunk_syms;
        out[k++] = Slice{c, cnt, lo, len};
        c += cnt;
    };
    const size_t tail = ramped ? ramp[0] + ramp[1] + ramp[2] : 0;
    if (ramped)
        for (int r = 0; r < 3; r++) push(ramp[r]);
    while (total - c > tail) push((total - c - tail < tc) ? total - c - tail : tc);
    if (ramped)
        for (int r = 2; r >= 0; r--) push(ramp[r]);
    return k;
}

// Host buffers -> blob, overlapped: slice i+1 is copied in while slice i is encoded and slice i-1 is
// copied out.  Every slice is its own container in a staging region; because containers end 16-byte
// aligned they concatenate, and the directory entries only need the running base added.  Directories are
// staged in pinned memory (an async copy to a pageable `offsets` would block the issuing thread).
int encode_host(rb200_ctx* ctx, const rb200_model* model, const uint8_t* in, size_t n, uint32_t chunk_syms, uint8_t* blob,
                size_t blob_cap, uint64_t* offsets, size_t* blob_size)
{
    const size_t n_chunks = rb200_chunk_count(n, chunk_syms);
    if (n == 0) {
        offsets[0] = 0;
        if (blob_size) *blob_size = 0;
        return RB200_OK;
    }
    Slice sl[rb200_ctx::kMaxSlices];
    const size_t n_slices = plan_slices(n, chunk_syms, sl);
    size_t max_cnt = 0;
    for (size_t i = 0; i < n_slices; i++) max_cnt = sl[i].cnt > max_cnt ? sl[i].cnt : max_cnt;
    // slice i's container is built at bound(symbols before it); its directory at c0 + i (cnt + 1 entries)
    int rc = reserve(ctx, ctx->st_in, n + 16);
    if (rc == RB200_OK) rc = reserve(ctx, ctx->st_blob, n_chunks * static_cast<size_t>(slot_bytes_for(chunk_syms)) + 16);
    if (rc == RB200_OK) rc = reserve(ctx, ctx->st_offsets, (n_chunks + n_slices) * sizeof(uint64_t));
    if (rc == RB200_OK) rc = reserve_dir(ctx, n_chunks + n_slices);
    if (rc == RB200_OK) rc = reserve_encode(ctx, model, static_cast<uint32_t>(max_cnt), chunk_syms);   // no regrowth mid-pipeline
    if (rc != RB200_OK) return rc;
    uint8_t* d_in = static_cast<uint8_t*>(ctx->st_in.p);
    uint8_t* d_blob = static_cast<uint8_t*>(ctx->st_blob.p);
    uint64_t* d_off = static_cast<uint64_t*>(ctx->st_offsets.p);
    const size_t slot = slot_bytes_for(chunk_syms);

    RB_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));       // staging buffers may still be in use upstream
    RB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_start, 0));
    RB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_start, 0));

    auto issue = [&](size_t i) -> int {      // H2D + kernels for slice i
        const Slice& s = sl[i];
        RB_CUDA(ctx, cudaMemcpyAsync(d_in + s.lo, in + s.lo, s.len, cudaMemcpyHostToDevice, ctx->s_in));
        RB_CUDA(ctx, cudaEventRecord(ctx->ev_in[i], ctx->s_in));
        RB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[i], 0));
        uint64_t* so = d_off + s.c0 + i;
        int r = encode_device(ctx, model, d_in + s.lo, s.len, chunk_syms, d_blob + s.c0 * slot, s.cnt * slot, so);
        if (r != RB200_OK) return r;
        RB_CUDA(ctx, cudaMemcpyAsync(&ctx->h_slice_total[i], so + s.cnt, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        RB_CUDA(ctx, cudaEventRecord(ctx->ev_done[i], ctx->stream));
        return RB200_OK;
    };

    rc = issue(0);
    if (rc != RB200_OK) return fail_pipeline(ctx, rc);
    size_t base = 0;
    bool overflow = false;
    for (size_t i = 0; i < n_slices; i++) {
        if (i + 1 < n_slices) {
            rc = issue(i + 1);
            if (rc != RB200_OK) return fail_pipeline(ctx, rc);
        }
        RB_CUDA(ctx, cudaEventSynchronize(ctx->ev_done[i]));          // slice i's size is now on the host
        const size_t total = static_cast<size_t>(ctx->h_slice_total[i]);
        const Slice& s = sl[i];
        if (base + total > blob_cap) {
            overflow = true;
        } else if (!overflow) {
            RB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_done[i], 0));
            RB_CUDA(ctx, cudaMemcpyAsync(blob + base, d_blob + s.c0 * slot, total, cudaMemcpyDeviceToHost, ctx->s_out));
            RB_CUDA(ctx, cudaMemcpyAsync(ctx->h_dir + s.c0 + i, d_off + s.c0 + i, s.cnt * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                                         ctx->s_out));
        }
        ctx->h_slice_total[i] = base;                                 // reuse as "base of slice i"
        base += total;
    }
    RB_CUDA(ctx, cudaEventRecord(ctx->ev_out, ctx->s_out));
    RB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_out, 0));
    rc = rb200_sync(ctx);
    if (rc != RB200_OK) return rc;
    if (overflow) return RB200_E_SPACE;
    for (size_t i = 0; i < n_slices; i++) {                           // rebase the directory into the caller's array
        const uint64_t add = ctx->h_slice_total[i];
        const uint64_t* src = ctx->h_dir + sl[i].c0 + i;
        uint64_t* dst = offsets + sl[i].c0;
        for (size_t c = 0; c < sl[i].cnt; c++) dst[c] = src[c] + add;
    }
    offsets[n_chunks] = base;
    if (blob_size) *blob_size = base;
    return RB200_OK;
}

// blob -> host buffers, overlapped the same way; no host synchronisation until the end because the
// directory (and therefore every slice's byte range) is already on the host.
int decode_host(rb200_ctx* ctx, const rb200_model* model, const uint8_t* blob, size_t blob_size, const uint64_t* offsets,
                uint32_t chunk_syms, uint8_t* out, size_t n)
{
    const size_t n_chunks = rb200_chunk_count(n, chunk_syms);
    if (n == 0) return RB200_OK;
    if (offsets[n_chunks] != blob_size) return RB200_E_STREAM;
    Slice sl[rb200_ctx::kMaxSlices];
    const size_t n_slices = plan_slices(n, chunk_syms, sl);
    int rc = reserve(ctx, ctx->st_blob, blob_size + 16);
    if (rc == RB200_OK) rc = reserve(ctx, ctx->st_offsets, (n_chunks + 1) * sizeof(uint64_t));
    if (rc == RB200_OK) rc = reserve(ctx, ctx->st_out, n + 16);
    if (rc == RB200_OK) rc = reserve_dir(ctx, n_chunks + 1);
    if (rc != RB200_OK) return rc;
    uint8_t* d_blob = static_cast<uint8_t*>(ctx->st_blob.p);
    uint64_t* d_off = static_cast<uint64_t*>(ctx->st_offsets.p);
    uint8_t* d_out = static_cast<uint8_t*>(ctx->st_out.p);

    RB_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
    RB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_start, 0));
    RB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_start, 0));
    for (size_t c = 0; c < n_chunks; c++) {                          // the copies below trust these
        if (offsets[c] > offsets[c + 1] || offsets[c + 1] > blob_size) return RB200_E_STREAM;
        ctx->h_dir[c] = offsets[c];
    }
    ctx->h_dir[n_chunks] = offsets[n_chunks];
    RB_CUDA(ctx, cudaMemcpyAsync(d_off, ctx->h_dir, (n_chunks + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->s_in));
    for (size_t i = 0; i < n_slices; i++) {
        const Slice& s = sl[i];
        const size_t c1 = s.c0 + s.cnt;
        // bytes of chunks [c0, c1): from the (aligned) end of chunk c0-1 to the end of chunk c1-1
        const size_t b0 = s.c0 ? static_cast<size_t>(offsets[s.c0] & ~15ull) : 0;
        const size_t b1 = static_cast<size_t>(offsets[c1] & ~15ull);
        if (b1 > b0) RB_CUDA(ctx, cudaMemcpyAsync(d_blob + b0, blob + b0, b1 - b0, cudaMemcpyHostToDevice, ctx->s_in));
        RB_CUDA(ctx, cudaEventRecord(ctx->ev_in[i], ctx->s_in));
        RB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[i], 0));
        rc = decode_device(ctx, model, d_blob, blob_size, d_off + s.c0, chunk_syms, d_out + s.lo, s.len);
        if (rc != RB200_OK) return fail_pipeline(ctx, rc);
        RB_CUDA(ctx, cudaEventRecord(ctx->ev_done[i], ctx->stream));
        RB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_done[i], 0));
        RB_CUDA(ctx, cudaMemcpyAsync(out + s.lo, d_out + s.lo, s.len, cudaMemcpyDeviceToHost, ctx->s_out));
    }
    RB_CUDA(ctx, cudaEventRecord(ctx->ev_out, ctx->s_out));
    RB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_out, 0));
    return rb200_sync(ctx);
}

}  // namespace

extern "C" int rb200_encode(rb200_ctx* ctx, const rb200_model* model, const uint8_t* in, size_t n, uint32_t chunk_syms,
                            uint8_t* blob, size_t blob_cap, uint64_t* offsets, size_t* blob_size, int mem_kind)
{
    if (!ctx || !model || model->ctx != ctx || !blob || !offsets || (!in && n) || !chunk_ok(chunk_syms)) return RB200_E_ARG;
    DeviceGuard g(ctx->device);
    if (mem_kind == RB200_MEM_DEVICE) {
        if (!aligned16(blob) || (reinterpret_cast<uintptr_t>(offsets) & 7)) return RB200_E_ARG;
        return encode_device(ctx, model, in, n, chunk_syms, blob, blob_cap, offsets);
    }
    if (mem_kind != RB200_MEM_HOST) return RB200_E_ARG;
    return encode_host(ctx, model, in, n, chunk_syms, blob, blob_cap, offsets, blob_size);
}

extern "C" int rb200_decode(rb200_ctx* ctx, const rb200_model* model, const uint8_t* blob, size_t blob_size,
                            const uint64_t* offsets, uint32_t chunk_syms, uint8_t* out, size_t n, int mem_kind)
{
    if (!ctx || !model || model->ctx != ctx || !offsets || (!out && n) || (!blob && blob_size) || !chunk_ok(chunk_syms))
        return RB200_E_ARG;
    if (blob_size & 15) return RB200_E_ARG;     // container invariant: the blob ends on a 16-byte boundary
    if (blob_size >> 36) return RB200_E_ARG;    // the decoders index the blob in 16-byte vectors with 32 bits
    DeviceGuard g(ctx->device);
    if (mem_kind == RB200_MEM_DEVICE) {
        if (!aligned16(blob) || (reinterpret_cast<uintptr_t>(offsets) & 7)) return RB200_E_ARG;
        return decode_device(ctx, model, blob, blob_size, offsets, chunk_syms, out, n);
    }
    if (mem_kind != RB200_MEM_HOST) return RB200_E_ARG;
    return decode_host(ctx, model, blob, blob_size, offsets, chunk_syms, out, n);
}

// ---------------------------------------------------------------------------
// histogram + per-block models: block_kernels.cuh
// ---------------------------------------------------------------------------

extern "C" int rb200_histogram(rb200_ctx* ctx, const uint8_t* in, size_t n, uint64_t counts[256], int mem_kind)
{
    if (!ctx || !counts || (!in && n)) return RB200_E_ARG;
    DeviceGuard g(ctx->device);
    int rc = reserve(ctx, ctx->st_aux, 256 * sizeof(unsigned long long));
    if (rc != RB200_OK) return rc;
    const uint8_t* d_in = in;
    if (mem_kind == RB200_MEM_HOST) {
        rc = reserve(ctx, ctx->st_in, n + 16);
        if (rc != RB200_OK) return rc;
        if (n) RB_CUDA(ctx, cudaMemcpyAsync(ctx->st_in.p, in, n, cudaMemcpyHostToDevice, ctx->stream));
        d_in = static_cast<const uint8_t*>(ctx->st_in.p);
    } else if (mem_kind != RB200_MEM_DEVICE) {
        return RB200_E_ARG;
    }
    unsigned long long* d_counts = static_cast<unsigned long long*>(ctx->st_aux.p);
    RB_CUDA(ctx, cudaMemsetAsync(d_counts, 0, 256 * sizeof(unsigned long long), ctx->stream));
    if (n) {
        launch_histogram(ctx->stream, d_in, n, d_counts);
        rc = check_launch(ctx, "histogram_kernel");
        if (rc != RB200_OK) return rc;
    }
    RB_CUDA(ctx, cudaMemcpyAsync(counts, d_counts, 256 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    RB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return RB200_OK;
}

extern "C" int rb200_model_from_data(rb200_ctx* ctx, int coder, uint32_t scale_bits, const uint8_t* data, size_t n, int mem_kind,
                                     uint32_t freqs_out[256], rb200_model** out)
{
    if (!ctx || !out || (!data && n) || scale_bits < 8 || scale_bits > 16) return RB200_E_ARG;
    if (n >> 32) return RB200_E_ARG;                 // the reference counts in 32 bits (main.cpp:52)
    if (n == 0) return RB200_E_MODEL;                // normalize_freqs divides by the total
    uint64_t counts[256];
    int rc = rb200_histogram(ctx, data, n, counts, mem_kind);
    if (rc != RB200_OK) return rc;
    uint32_t freqs[256], cum[257];
    for (int s = 0; s < 256; s++) freqs[s] = static_cast<uint32_t>(counts[s]);
    rc = rb200_normalize_freqs(freqs, cum, 1u << scale_bits);
    if (rc != RB200_OK) return rc;
    rc = rb200_model_create(ctx, coder, scale_bits, freqs, out);
    if (rc == RB200_OK && freqs_out) std::memcpy(freqs_out, freqs, sizeof freqs);
    return rc;
}

extern "C" int rb200_blocks_build_models(rb200_ctx* ctx, const uint8_t* in, uint32_t n_blocks, uint32_t block_size,
                                         uint16_t* block_freqs, int mem_kind)
{
    if (!ctx || !in || !block_freqs || !n_blocks || !block_size) return RB200_E_ARG;
    DeviceGuard g(ctx->device);
    const size_t n = static_cast<size_t>(n_blocks) * block_size;
    const size_t fbytes = static_cast<size_t>(n_blocks) * 256 * sizeof(uint16_t);
    const uint8_t* d_in = in;
    uint16_t* d_freqs = block_freqs;
    int rc;
    if (mem_kind == RB200_MEM_HOST) {
        rc = reserve(ctx, ctx->st_in, n + 16);
        if (rc == RB200_OK) rc = reserve(ctx, ctx->st_aux, fbytes);
        if (rc != RB200_OK) return rc;
        RB_CUDA(ctx, cudaMemcpyAsync(ctx->st_in.p, in, n, cudaMemcpyHostToDevice, ctx->stream));
        d_in = static_cast<const uint8_t*>(ctx->st_in.p);
        d_freqs = static_cast<uint16_t*>(ctx->st_aux.p);
    } else if (mem_kind != RB200_MEM_DEVICE) {
        return RB200_E_ARG;
    }
    launch_block_models(ctx->stream, d_in, n_blocks, block_size, d_freqs, ctx->d_status);
    rc = check_launch(ctx, "block_model_kernel");
    if (rc != RB200_OK) return rc;
    if (mem_kind == RB200_MEM_HOST) {
        RB_CUDA(ctx, cudaMemcpyAsync(block_freqs, d_freqs, fbytes, cudaMemcpyDeviceToHost, ctx->stream));
        return rb200_sync(ctx);
    }
    return RB200_OK;
}

namespace {

// Per-block encode.  Default: ONE persistent launch (block_encode_fused_kernel: [model,] tables, encode, directory,
// placement); build_models makes the kernel run count_freqs + normalize_freqs per block first and write d_freqs.
// RB200_ENCODE_PATH=split keeps the round-1 sequence (encode into per-chunk slots, tile scan, compaction).
int blocks_encode_device(rb200_ctx* ctx, const uint8_t* d_in, uint32_t n_blocks, uint32_t block_size, uint16_t* d_freqs,
                         bool build_models, uint32_t chunk_syms, uint8_t* d_blob, size_t blob_cap, uint64_t* d_offsets)
{
    const uint32_t per_block = block_size / chunk_syms;
    const uint64_t n_chunks64 = static_cast<uint64_t>(n_blocks) * per_block;
    if (n_chunks64 >= (1ull << 31)) return RB200_E_ARG;
    const uint32_t n_chunks = static_cast<uint32_t>(n_chunks64);
    const uint32_t slot = slot_bytes_for(chunk_syms);
    if (use_fused_encode(chunk_syms) || build_models) {
        const uint32_t threads = block_threads(block_size, chunk_syms);
        const uint32_t grid = block_fused_grid(static_cast<uint32_t>(ctx->sms), n_blocks, threads, build_models);
        int rc = reserve(ctx, ctx->scratch, static_cast<size_t>(grid) * (threads / 32) * 2 * slot + 16);   // two slots per resident warp
        if (rc == RB200_OK) rc = reserve(ctx, ctx->sizes, 16 + static_cast<size_t>(n_chunks) * sizeof(uint64_t));
        if (rc != RB200_OK) return rc;
        uint32_t* counter = static_cast<uint32_t*>(ctx->sizes.p);
        uint64_t* look = reinterpret_cast<uint64_t*>(static_cast<uint8_t*>(ctx->sizes.p) + 16);
        RB_CUDA(ctx, cudaMemsetAsync(ctx->sizes.p, 0, 16 + static_cast<size_t>(n_chunks) * sizeof(uint64_t), ctx->stream));
        launch_block_encode_fused(ctx->stream, grid, threads, build_models, d_in, n_blocks, block_size, d_freqs, chunk_syms,
                                  static_cast<uint8_t*>(ctx->scratch.p), slot, look, counter, d_blob, blob_cap, d_offsets, ctx->d_status);
        return check_launch(ctx, "block_encode_fused_kernel");
    }
    uint8_t* scratch; uint32_t* sizes; uint64_t* tile_sums;
    int rc = reserve_encode_workspace(ctx, n_chunks, slot, &scratch, &sizes, &tile_sums);
    if (rc != RB200_OK) return rc;
    launch_block_encode(ctx->stream, d_in, n_blocks, block_size, d_freqs, chunk_syms, scratch, slot, sizes, ctx->d_status);
    rc = check_launch(ctx, "block_encode_kernel");
    if (rc != RB200_OK) return rc;
    return finish_encode(ctx, scratch, slot, sizes, tile_sums, n_chunks, d_blob, blob_cap, d_offsets);
}

// one CTA per block, one warp per chunk of the block: at most 32 chunks per block
bool blocks_geometry_ok(uint32_t n_blocks, uint32_t block_size, uint32_t chunk_syms)
{
    return n_blocks && block_size && chunk_ok(chunk_syms) && block_size % chunk_syms == 0 && block_size / chunk_syms <= 32;
}


// Host buffers -> blob for the per-block path, overlapped like encode_host: slices of whole blocks; slice i+1 is
// copied in (with its frequency tables when the caller supplies them) while slice i is encoded and slice i-1 is copied
// out (blob, directory, and the frequency tables when the kernel built them).
int blocks_encode_host(rb200_ctx* ctx, const uint8_t* in, uint32_t n_blocks, uint32_t block_size, uint16_t* block_freqs, bool build,
                       uint32_t chunk_syms, uint8_t* blob, size_t blob_cap, uint64_t* offsets, size_t* blob_size)
{
    const uint32_t per_block = block_size / chunk_syms;
    const size_t n = static_cast<size_t>(n_blocks) * block_size;
    const size_t n_chunks = static_cast<size_t>(n_blocks) * per_block;
    Slice sl[rb200_ctx::kMaxSlices];                         // planned in units of blocks: c0 / cnt count blocks here
    const size_t n_slices = plan_slices(n, block_size, sl);
    const size_t slot = slot_bytes_for(chunk_syms);
    const size_t fbytes = static_cast<size_t>(n_blocks) * 256 * sizeof(uint16_t);
    int rc = reserve(ctx, ctx->st_in, n + 16);
    if (rc == RB200_OK) rc = reserve(ctx, ctx->st_blob, n_chunks * slot + 16);
    if (rc == RB200_OK) rc = reserve(ctx, ctx->st_offsets, (n_chunks + n_slices) * sizeof(uint64_t));
    if (rc == RB200_OK) rc = reserve(ctx, ctx->st_aux, fbytes);
    if (rc == RB200_OK) rc = reserve_dir(ctx, n_chunks + n_slices);
    if (rc != RB200_OK) return rc;
    uint8_t* d_in = static_cast<uint8_t*>(ctx->st_in.p);
    uint8_t* d_blob = static_cast<uint8_t*>(ctx->st_blob.p);
    uint64_t* d_off = static_cast<uint64_t*>(ctx->st_offsets.p);
    uint16_t* d_freqs = static_cast<uint16_t*>(ctx->st_aux.p);

    RB_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
    RB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_start, 0));
    RB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_start, 0));

    auto issue = [&](size_t i) -> int {
        const Slice& s = sl[i];
        const size_t c0 = s.c0 * per_block, cnt = s.cnt * per_block;          // chunks of this slice
        RB_CUDA(ctx, cudaMemcpyAsync(d_in + s.lo, in + s.lo, s.len, cudaMemcpyHostToDevice, ctx->s_in));
        if (!build)
            RB_CUDA(ctx, cudaMemcpyAsync(d_freqs + s.c0 * 256, block_freqs + s.c0 * 256, s.cnt * 256 * sizeof(uint16_t),
                                         cudaMemcpyHostToDevice, ctx->s_in));
        RB_CUDA(ctx, cudaEventRecord(ctx->ev_in[i], ctx->s_in));
        RB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[i], 0));
        uint64_t* so = d_off + c0 + i;
        int r = blocks_encode_device(ctx, d_in + s.lo, static_cast<uint32_t>(s.cnt), block_size, d_freqs + s.c0 * 256, build, chunk_syms,
                                     d_blob + c0 * slot, cnt * slot, so);
        if (r != RB200_OK) return r;
        RB_CUDA(ctx, cudaMemcpyAsync(&ctx->h_slice_total[i], so + cnt, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
        RB_CUDA(ctx, cudaEventRecord(ctx->ev_done[i], ctx->stream));
        return RB200_OK;
    };
    rc = issue(0);
    if (rc != RB200_OK) return fail_pipeline(ctx, rc);
    size_t base = 0;
    bool overflow = false;
    for (size_t i = 0; i < n_slices; i++) {
        if (i + 1 < n_slices) {
            rc = issue(i + 1);
            if (rc != RB200_OK) return fail_pipeline(ctx, rc);
        }
        RB_CUDA(ctx, cudaEventSynchronize(ctx->ev_done[i]));
        const size_t total = static_cast<size_t>(ctx->h_slice_total[i]);
        const Slice& s = sl[i];
        const size_t c0 = s.c0 * per_block, cnt = s.cnt * per_block;
        if (base + total > blob_cap) {
            overflow = true;
        } else if (!overflow) {
            RB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_done[i], 0));
            RB_CUDA(ctx, cudaMemcpyAsync(blob + base, d_blob + c0 * slot, total, cudaMemcpyDeviceToHost, ctx->s_out));
            RB_CUDA(ctx, cudaMemcpyAsync(ctx->h_dir + c0 + i, d_off + c0 + i, cnt * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->s_out));
            if (build)
                RB_CUDA(ctx, cudaMemcpyAsync(block_freqs + s.c0 * 256, d_freqs + s.c0 * 256, s.cnt * 256 * sizeof(uint16_t),
                                             cudaMemcpyDeviceToHost, ctx->s_out));
        }
        ctx->h_slice_total[i] = base;
        base += total;
    }
    RB_CUDA(ctx, cudaEventRecord(ctx->ev_out, ctx->s_out));
    RB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_out, 0));
    rc = rb200_sync(ctx);
    if (rc != RB200_OK) return rc;
    if (overflow) return RB200_E_SPACE;
    for (size_t i = 0; i < n_slices; i++) {                  // rebase the directory into the caller's array
        const uint64_t add = ctx->h_slice_total[i];
        const size_t c0 = sl[i].c0 * per_block, cnt = sl[i].cnt * per_block;
        const uint64_t* src = ctx->h_dir + c0 + i;
        uint64_t* dst = offsets + c0;
        for (size_t c = 0; c < cnt; c++) dst[c] = src[c] + add;
    }
    offsets[n_chunks] = base;
    if (blob_size) *blob_size = base;
    return RB200_OK;
}

int blocks_decode_host(rb200_ctx* ctx, const uint8_t* blob, size_t blob_size, const uint64_t* offsets, const uint16_t* block_freqs,
                       uint32_t n_blocks, uint32_t block_size, uint32_t chunk_syms, uint8_t* out)
{
    const uint32_t per_block = block_size / chunk_syms;
    const size_t n = static_cast<size_t>(n_blocks) * block_size;
    const size_t n_chunks = static_cast<size_t>(n_blocks) * per_block;
    if (offsets[n_chunks] != blob_size) return RB200_E_STREAM;
    Slice sl[rb200_ctx::kMaxSlices];                         // units of blocks
    const size_t n_slices = plan_slices(n, block_size, sl);
    const size_t fbytes = static_cast<size_t>(n_blocks) * 256 * sizeof(uint16_t);
    int rc = reserve(ctx, ctx->st_blob, blob_size + 16);
    if (rc == RB200_OK) rc = reserve(ctx, ctx->st_offsets, (n_chunks + 1) * sizeof(uint64_t));
    if (rc == RB200_OK) rc = reserve(ctx, ctx->st_out, n + 16);
    if (rc == RB200_OK) rc = reserve(ctx, ctx->st_aux, fbytes);
    if (rc == RB200_OK) rc = reserve_dir(ctx, n_chunks + 1);
    if (rc != RB200_OK) return rc;
    uint8_t* d_blob = static_cast<uint8_t*>(ctx->st_blob.p);
    uint64_t* d_off = static_cast<uint64_t*>(ctx->st_offsets.p);
    uint8_t* d_out = static_cast<uint8_t*>(ctx->st_out.p);
    uint16_t* d_freqs = static_cast<uint16_t*>(ctx->st_aux.p);
    for (size_t c = 0; c < n_chunks; c++) {                  // the copies below trust these
        if (offsets[c] > offsets[c + 1] || offsets[c + 1] > blob_size) return RB200_E_STREAM;
        ctx->h_dir[c] = offsets[c];
    }
    ctx->h_dir[n_chunks] = offsets[n_chunks];

    RB_CUDA(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
    RB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_start, 0));
    RB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_start, 0));
    RB_CUDA(ctx, cudaMemcpyAsync(d_off, ctx->h_dir, (n_chunks + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->s_in));
    for (size_t i = 0; i < n_slices; i++) {
        const Slice& s = sl[i];
        const size_t c0 = s.c0 * per_block, c1 = (s.c0 + s.cnt) * per_block;
        const size_t b0 = c0 ? static_cast<size_t>(offsets[c0] & ~15ull) : 0;
        const size_t b1 = static_cast<size_t>(offsets[c1] & ~15ull);
        if (b1 > b0) RB_CUDA(ctx, cudaMemcpyAsync(d_blob + b0, blob + b0, b1 - b0, cudaMemcpyHostToDevice, ctx->s_in));
        RB_CUDA(ctx, cudaMemcpyAsync(d_freqs + s.c0 * 256, block_freqs + s.c0 * 256, s.cnt * 256 * sizeof(uint16_t), cudaMemcpyHostToDevice,
                                     ctx->s_in));
        RB_CUDA(ctx, cudaEventRecord(ctx->ev_in[i], ctx->s_in));
        RB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[i], 0));
        launch_block_decode(ctx->stream, d_blob, blob_size, d_off + c0, d_freqs + s.c0 * 256, static_cast<uint32_t>(s.cnt), block_size,
                            chunk_syms, d_out + s.lo, ctx->d_status);
        rc = check_launch(ctx, "block_decode_kernel");
        if (rc != RB200_OK) return fail_pipeline(ctx, rc);
        RB_CUDA(ctx, cudaEventRecord(ctx->ev_done[i], ctx->stream));
        RB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_done[i], 0));
        RB_CUDA(ctx, cudaMemcpyAsync(out + s.lo, d_out + s.lo, s.len, cudaMemcpyDeviceToHost, ctx->s_out));
    }
    RB_CUDA(ctx, cudaEventRecord(ctx->ev_out, ctx->s_out));
    RB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_out, 0));
    return rb200_sync(ctx);
}

int blocks_encode_entry(rb200_ctx* ctx, const uint8_t* in, uint32_t n_blocks, uint32_t block_size, uint16_t* block_freqs, bool build,
                        uint32_t chunk_syms, uint8_t* blob, size_t blob_cap, uint64_t* offsets, size_t* blob_size, int mem_kind)
{
    if (!ctx || !in || !block_freqs || !blob || !offsets || !blocks_geometry_ok(n_blocks, block_size, chunk_syms)) return RB200_E_ARG;
    DeviceGuard g(ctx->device);
    if (mem_kind == RB200_MEM_DEVICE) {
        if (!aligned16(blob) || (reinterpret_cast<uintptr_t>(offsets) & 7)) return RB200_E_ARG;
        return blocks_encode_device(ctx, in, n_blocks, block_size, block_freqs, build, chunk_syms, blob, blob_cap, offsets);
    }
    if (mem_kind != RB200_MEM_HOST) return RB200_E_ARG;
    return blocks_encode_host(ctx, in, n_blocks, block_size, block_freqs, build, chunk_syms, blob, blob_cap, offsets, blob_size);
}

}  // namespace

extern "C" int rb200_blocks_encode(rb200_ctx* ctx, const uint8_t* in, uint32_t n_blocks, uint32_t block_size,
                                   const uint16_t* block_freqs, uint32_t chunk_syms, uint8_t* blob, size_t blob_cap,
                                   uint64_t* offsets, size_t* blob_size, int mem_kind)
{
    return blocks_encode_entry(ctx, in, n_blocks, block_size, const_cast<uint16_t*>(block_freqs), false, chunk_syms, blob, blob_cap, offsets,
                               blob_size, mem_kind);
}

extern "C" int rb200_blocks_model_encode(rb200_ctx* ctx, const uint8_t* in, uint32_t n_blocks, uint32_t block_size,
                                         uint16_t* block_freqs, uint32_t chunk_syms, uint8_t* blob, size_t blob_cap,
                                         uint64_t* offsets, size_t* blob_size, int mem_kind)
{
    return blocks_encode_entry(ctx, in, n_blocks, block_size, block_freqs, true, chunk_syms, blob, blob_cap, offsets, blob_size, mem_kind);
}

extern "C" int rb200_blocks_decode(rb200_ctx* ctx, const uint8_t* blob, size_t blob_size, const uint64_t* offsets,
                                   const uint16_t* block_freqs, uint32_t n_blocks, uint32_t block_size, uint32_t chunk_syms,
                                   uint8_t* out, int mem_kind)
{
    if (!ctx || !blob || !offsets || !block_freqs || !out || !blocks_geometry_ok(n_blocks, block_size, chunk_syms)) return RB200_E_ARG;
    if (blob_size & 15) return RB200_E_ARG;
    if (blob_size >> 36) return RB200_E_ARG;
    DeviceGuard g(ctx->device);
    const size_t n_chunks = static_cast<size_t>(n_blocks) * (block_size / chunk_syms);
    if (n_chunks >= (1ull << 31)) return RB200_E_ARG;
    if (mem_kind == RB200_MEM_DEVICE) {
        if (!aligned16(blob) || (reinterpret_cast<uintptr_t>(offsets) & 7)) return RB200_E_ARG;
        launch_block_decode(ctx->stream, blob, blob_size, offsets, block_freqs, n_blocks, block_size, chunk_syms, out, ctx->d_status);
        return check_launch(ctx, "block_decode_kernel");
    }
    if (mem_kind != RB200_MEM_HOST) return RB200_E_ARG;
    return blocks_decode_host(ctx, blob, blob_size, offsets, block_freqs, n_blocks, block_size, chunk_syms, out);
}

// ---------------------------------------------------------------------------
// multi-GPU: gathering the shards' blobs over NCCL
// ---------------------------------------------------------------------------
#include "gather_nccl.cuh"
