/*
 * rans_b200.h -- bulk C-ABI of the B200-native interleaved rANS coder.
 *
 * The reference (rygorous/ryg_rans) is header-only: its hot path is the per-symbol
 * step functions of rans_byte.h / rans64.h / rans_word_sse41.h inlined into the
 * driver loops of main.cpp / main64.cpp / main_simd.cpp / main_alias.cpp.  It has
 * no FFI.  This header is the boundary a replacement .so exports so that those
 * driver loops can be swapped for one bulk call each; every entry point names the
 * reference code it stands in for.  Plain C types only (no CUDA, no torch types):
 * the CUDA stream is passed as an opaque void* (a cudaStream_t).
 *
 * Container ("chunked N=32 streams"):
 *   the symbol buffer is cut into chunks of `chunk_syms` symbols.  Chunk c is one
 *   independent 32-way interleaved rANS stream laid out EXACTLY as the reference
 *   drivers lay out their N-way streams (main_simd.cpp:287-300 with 8 -> 32;
 *   main_alias.cpp:353-373 with 2 -> 32): 32 little-endian u32 final states, lane 0
 *   first, then the renormalisation units in decode order.  Stream c occupies
 *   blob[offsets[c] .. E_c) where E_c = offsets[c+1] & ~15 is 16-byte aligned
 *   (streams are END-aligned; the <16 bytes between E_c and offsets[c+1] are zero),
 *   and offsets[n_chunks] = blob size.  Any reference decoder primitive can decode
 *   one chunk stream given blob + offsets[c].
 */
#ifndef RANS_B200_H
#define RANS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB200_VERSION 2

/* status codes (the reference only has assert(); SURVEY section 5 asks for codes) */
#define RB200_OK          0
#define RB200_E_ARG      -1   /* bad argument (null, misaligned, size mismatch)         */
#define RB200_E_MODEL    -2   /* frequencies do not form a valid model                  */
#define RB200_E_SPACE    -3   /* output buffer too small                                */
#define RB200_E_STREAM   -4   /* corrupt / truncated stream detected while decoding     */
#define RB200_E_CUDA     -5   /* CUDA runtime error, see rb200_last_cuda_error          */
#define RB200_E_NOMEM    -6
#define RB200_E_SYMBOL   -7   /* encoder met a symbol whose model frequency is 0        */
#define RB200_E_NCCL     -8   /* NCCL missing or an NCCL call failed, see rb200_last_cuda_error */
#define RB200_E_STALL    -9   /* a bounded wait inside a kernel (8 s without progress) expired: a bug or a wedged
                                 device, never bad input -- the kernels give up instead of hanging the GPU            */

/* which memory the data pointers of a bulk call live in */
#define RB200_MEM_HOST    0   /* host pointers; the call copies H2D/D2H and is synchronous */
#define RB200_MEM_DEVICE  1   /* device pointers; the call only enqueues on the stream     */

/* coder families (one per reference header / driver) */
#define RB200_CODER_WORD    0 /* rans_word_sse41.h: 32-bit state, u16 renorm, scale_bits = 12          */
#define RB200_CODER_BYTE    1 /* rans_byte.h + cum2sym lookup as in main.cpp, scale_bits 8..16           */
#define RB200_CODER_ALIAS   2 /* rans_byte.h state machine + main_alias.cpp alias tables, scale_bits 8..16 */
#define RB200_CODER_RANS64  3 /* rans64.h: 64-bit state, u32 renorm, cum2sym lookup as in main64.cpp      */

#define RB200_LANES 32        /* interleave width of one chunk stream = one warp */

typedef struct rb200_ctx rb200_ctx;
typedef struct rb200_model rb200_model;

/* ---------------------------------------------------------------- host-side model */

/* SymbolStats::count_freqs, main.cpp:59-66 (host loop; the device version is
 * rb200_histogram). */
int rb200_count_freqs(const uint8_t* in, size_t n, uint32_t freqs[256]);

/* SymbolStats::calc_cum_freqs + normalize_freqs, main.cpp:68-129.  In: raw counts.
 * Out: freqs summing to target_total (a power of two >= 256) and cum_freqs[257].
 * Returns RB200_E_MODEL where the reference would assert. */
int rb200_normalize_freqs(uint32_t freqs[256], uint32_t cum_freqs[257], uint32_t target_total);

/* RansWordTablesInitSymbol for all 256 symbols, rans_word_sse41.h:64-72 /
 * main_simd.cpp:141-143.  slots[i] = freq | bias << 16 (RansWordSlot.u32). */
int rb200_word_tables_build(const uint32_t freqs[256], const uint32_t cum_freqs[257],
                            uint32_t slots[4096], uint8_t slot2sym[4096]);

/* SymbolStats::make_alias_table, main_alias.cpp:147-237.  alias_remap needs
 * cum_freqs[256] entries. */
int rb200_alias_tables_build(const uint32_t freqs[256], const uint32_t cum_freqs[257],
                             uint32_t divider[256], uint32_t slot_adjust[512],
                             uint32_t slot_freqs[512], uint8_t sym_id[512], uint32_t* alias_remap);

/* ---------------------------------------------------------------- context */

/* One context per (device, stream, calling thread).  stream = cudaStream_t or NULL. */
int rb200_ctx_create(rb200_ctx** out, int device, void* stream);
void rb200_ctx_destroy(rb200_ctx* ctx);
int rb200_ctx_set_stream(rb200_ctx* ctx, void* stream);
/* Wait for the stream, then report-and-clear what the kernels flagged
 * (RB200_OK, RB200_E_STREAM or RB200_E_SYMBOL). */
int rb200_sync(rb200_ctx* ctx);
const char* rb200_last_cuda_error(const rb200_ctx* ctx);
const char* rb200_strerror(int code);
int rb200_version(void);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t rb200_launch_count(const rb200_ctx* ctx);

/* ---------------------------------------------------------------- device model */

/* Upload the coding tables for one model.  freqs must already be normalised to
 * 1 << scale_bits (WORD: scale_bits must be 12, rans_word_sse41.h:37; BYTE / ALIAS / RANS64:
 * 8..16 -- main.cpp:136 and main64.cpp:136 use 14, main_alias.cpp:276 uses 16; BYTE needs every
 * freq < 65536 like the reference's 16-bit RansDecSymbol).  Replaces the table set-up at
 * main.cpp:145-162 / main64.cpp:145-164 / main_simd.cpp:141-143 / main_alias.cpp:279-282. */
int rb200_model_create(rb200_ctx* ctx, int coder, uint32_t scale_bits,
                       const uint32_t freqs[256], rb200_model** out);
void rb200_model_destroy(rb200_model* m);

/* ---------------------------------------------------------------- geometry */

size_t rb200_chunk_count(size_t n, uint32_t chunk_syms);
/* Worst-case blob size for n symbols: replaces the reference's guesses
 * (32 MB at main.cpp:150; n + n/8 + 128 at main_simd.cpp:145, not a true bound). */
size_t rb200_encode_bound(size_t n, uint32_t chunk_syms);

/* ---------------------------------------------------------------- the hot path */

/* Replaces the N-way encode loop + flush (main_simd.cpp:287-300 / main_alias.cpp:353-373).
 * in[n] symbols -> blob (<= blob_cap bytes) + offsets[n_chunks + 1].
 * HOST mode: synchronous; *blob_size receives offsets[n_chunks].
 * DEVICE mode: in/blob/offsets are device pointers (blob and offsets 16-byte aligned),
 * work is enqueued on the context stream, blob_size may be NULL (read offsets[n_chunks]). */
int rb200_encode(rb200_ctx* ctx, const rb200_model* model,
                 const uint8_t* in, size_t n, uint32_t chunk_syms,
                 uint8_t* blob, size_t blob_cap, uint64_t* offsets, size_t* blob_size, int mem_kind);

/* Replaces the N-way decode loop (main_simd.cpp:313-332 / main_alias.cpp:386-405).
 * blob/offsets as produced by rb200_encode (or by the reference primitives driven
 * with N = 32) -> out[n].  Reads are bounded by blob_size; a stream that does not
 * end where its chunk ends is reported as RB200_E_STREAM by rb200_sync (DEVICE
 * mode) or by the call itself (HOST mode). */
int rb200_decode(rb200_ctx* ctx, const rb200_model* model,
                 const uint8_t* blob, size_t blob_size, const uint64_t* offsets,
                 uint32_t chunk_syms, uint8_t* out, size_t n, int mem_kind);

/* Device SymbolStats::count_freqs (main.cpp:59-66) over a device or host buffer;
 * counts[256] is written to HOST memory, the call synchronises. */
int rb200_histogram(rb200_ctx* ctx, const uint8_t* in, size_t n, uint64_t counts[256], int mem_kind);

/* The whole model set-up of the drivers in one call, with the pass over the data on the GPU
 * (SURVEY 8f.1): stats.count_freqs (main.cpp:59-66; rb200_histogram's kernel), then
 * stats.normalize_freqs(1 << scale_bits) (main.cpp:75-129; 256 counters, on the host, the
 * reference's algorithm verbatim), then the table build of rb200_model_create.  Replaces
 * main_simd.cpp:131-143 / main.cpp:138-162 / main_alias.cpp:276-282.  freqs_out[256] (may be
 * NULL) receives the normalised frequencies -- the model a decoder needs.  n must be below
 * 2^32 like the reference's 32-bit histogram (RB200_E_ARG otherwise); an empty input or one
 * the reference would assert on is RB200_E_MODEL.  The call synchronises. */
int rb200_model_from_data(rb200_ctx* ctx, int coder, uint32_t scale_bits, const uint8_t* data, size_t n,
                          int mem_kind, uint32_t freqs_out[256], rb200_model** out);

/* ---------------------------------------------------------------- per-block models */

/* BASELINE config 5: n_blocks independent blocks of block_size symbols, each with
 * its own order-0 model at scale_bits 12 (word coder).  block_freqs is
 * [n_blocks][256] u16, normalised to 4096 per block (what RansWordTablesInitSymbol
 * would be fed, rans_word_sse41.h:64-72); the 4096-slot decode table of each block
 * is built in shared memory inside the kernel.  Each block is cut into chunks of
 * chunk_syms symbols (block_size % chunk_syms == 0, and at most 32 chunks per block: one CTA per
 * block, one warp per chunk -- RB200_E_ARG otherwise); chunk index = block *
 * (block_size / chunk_syms) + chunk-in-block; the container is the same as above. */
int rb200_blocks_build_models(rb200_ctx* ctx, const uint8_t* in, uint32_t n_blocks, uint32_t block_size,
                              uint16_t* block_freqs, int mem_kind);
int rb200_blocks_encode(rb200_ctx* ctx, const uint8_t* in, uint32_t n_blocks, uint32_t block_size,
                        const uint16_t* block_freqs, uint32_t chunk_syms,
                        uint8_t* blob, size_t blob_cap, uint64_t* offsets, size_t* blob_size, int mem_kind);
int rb200_blocks_decode(rb200_ctx* ctx, const uint8_t* blob, size_t blob_size, const uint64_t* offsets,
                        const uint16_t* block_freqs, uint32_t n_blocks, uint32_t block_size,
                        uint32_t chunk_syms, uint8_t* out, int mem_kind);
/* rb200_blocks_build_models + rb200_blocks_encode in ONE launch: every block's model is built (count_freqs +
 * normalize_freqs(4096), main.cpp:59-129, bit-exact) by the CTA that then encodes the block, and returned in
 * block_freqs[n_blocks][256] -- what a decoder needs next to the blob.  A block has at most 32 chunks
 * (block_size / chunk_syms <= 32: one warp per chunk, one CTA per block); HOST mode pipelines slices of whole blocks
 * over three streams like rb200_encode. */
int rb200_blocks_model_encode(rb200_ctx* ctx, const uint8_t* in, uint32_t n_blocks, uint32_t block_size,
                              uint16_t* block_freqs, uint32_t chunk_syms,
                              uint8_t* blob, size_t blob_cap, uint64_t* offsets, size_t* blob_size, int mem_kind);

/* ---------------------------------------------------------------- multi-GPU: gathering the shards' blobs (NCCL) */

/* Shards are independent: every rank encodes / decodes its own contiguous range of symbols with no
 * communication.  The one exchange step is collecting the compressed blobs + directories on one rank, which these
 * calls do over NCCL / NVLink (ncclAllGather of the sizes, then grouped ncclSend / ncclRecv of the payloads straight
 * into their final position; the directories are rebased on the root).  The concatenation is itself a valid container
 * because every shard's blob ends 16-byte aligned.  The reference analogue is the driver keeping `rans_begin` per
 * buffer (main_simd.cpp:287-300); it has no multi-device story.  libnccl.so.2 is loaded on first use (dlopen), so
 * single-GPU users need no NCCL.
 *
 *   rank 0: rb200_comm_unique_id(id)  -> hand the 128 bytes to every rank (file, socket, MPI, torch.distributed ...)
 *   all   : rb200_comm_create(ctx, id, rank, world, &comm)             (collective: ncclCommInitRank)
 *   all   : rb200_gather_plan(comm, blob_size, n_chunks, totals)       (collective, synchronises: sizes of every shard)
 *   all   : rb200_gather_blobs(comm, root, d_blob, d_offsets, d_out_blob, out_cap, d_out_offsets)
 *           enqueued on the context's stream; only `root` needs the out buffers: totals[0] bytes of blob and
 *           totals[1] + 1 directory entries.  May be repeated while the planned sizes stay valid. */
typedef struct rb200_comm rb200_comm;
#define RB200_NCCL_ID_BYTES 128
int rb200_comm_unique_id(uint8_t id[RB200_NCCL_ID_BYTES]);
int rb200_comm_create(rb200_ctx* ctx, const uint8_t id[RB200_NCCL_ID_BYTES], int rank, int world, rb200_comm** out);
void rb200_comm_destroy(rb200_comm* comm);
int rb200_gather_plan(rb200_comm* comm, uint64_t blob_size, uint64_t n_chunks, uint64_t totals[2]);
int rb200_gather_blobs(rb200_comm* comm, int root, const uint8_t* d_blob, const uint64_t* d_offsets,
                       uint8_t* d_out_blob, uint64_t out_cap, uint64_t* d_out_offsets);

/* ---------------------------------------------------------------- wire format (host only) */

/* Self-describing envelope around (freqs, offsets, blob): versioned 64-byte header (coder, scale_bits,
 * lanes, chunk_syms, n, n_chunks, blob size), the model, the directory, the blob; CRC-32 over header +
 * model + directory always, over the blob when RB200_CONTAINER_CRC_BLOB is given.  The reference has
 * nothing comparable (it keeps a pointer, main.cpp:182-188).  rb200_container_open validates and
 * returns pointers INTO buf (which must be 8-byte aligned; the blob then sits 16-byte aligned). */
#define RB200_CONTAINER_CRC_BLOB 1u

typedef struct rb200_container_info {
    uint32_t coder, scale_bits, chunk_syms, flags;
    uint64_t n_symbols, n_chunks, blob_bytes;
    const uint32_t* freqs;      /* 256 */
    const uint64_t* offsets;    /* n_chunks + 1 */
    const uint8_t* blob;
} rb200_container_info;

size_t rb200_container_size(size_t n_chunks, size_t blob_bytes);
int rb200_container_pack(int coder, uint32_t scale_bits, uint32_t chunk_syms, size_t n, const uint32_t freqs[256],
                         const uint64_t* offsets, const uint8_t* blob, size_t blob_bytes, uint32_t flags,
                         uint8_t* out, size_t out_cap, size_t* out_size);
int rb200_container_open(const uint8_t* buf, size_t size, rb200_container_info* info);

#ifdef __cplusplus
}
#endif
#endif /* RANS_B200_H */
