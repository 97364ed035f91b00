// gather_nccl.cuh -- the path's only exchange step: every rank's compressed blob + directory to one rank
// over NCCL.  Included by rans_b200.cu (needs rb200_ctx).  libnccl.so.2 is bound at run time with dlopen/dlsym: the
// library has no link-time NCCL dependency, and inside a process that already loaded an NCCL (PyTorch bundles one) the
// same instance is used.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>
#include <vector>

namespace rb200 {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

inline const NcclApi& nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        auto sym = [&](const char* name) { return dlsym(h, name); };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
        api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.Send && api.Recv && api.GroupStart &&
                 api.GroupEnd && api.GetErrorString;
    });
    return api;
}

// directory entries of one shard, moved into the gathered directory: + the bytes of the shards before it
__global__ void rebase_directory_kernel(const uint64_t* __restrict__ src, uint64_t* __restrict__ dst, uint64_t n, uint64_t add)
{
    const uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
    if (i < n) dst[i] = src[i] + add;
}
__global__ void set_u64_kernel(uint64_t* p, uint64_t v) { *p = v; }

}  // namespace rb200

struct rb200_comm {
    rb200_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    uint64_t* d_sizes = nullptr;            // [world + 1][2]: slot `world` holds this rank's pair
    uint64_t* h_sizes = nullptr;            // pinned mirror
    std::vector<uint64_t> blob_bytes, chunks;   // the current plan
    bool planned = false;
};

namespace {
int nccl_fail(rb200_ctx* ctx, ncclResult_t r, const char* what)
{
    const rb200::NcclApi& api = rb200::nccl_api();
    if (ctx) ctx->last_error = std::string(what) + ": " + (api.ok ? api.GetErrorString(r) : "libnccl.so.2 not found");
    return RB200_E_NCCL;
}
#define RB_NCCL(ctx, call)                                              \
    do {                                                                \
        ncclResult_t r_ = (call);                                       \
        if (r_ != ncclSuccess) return nccl_fail((ctx), r_, #call);      \
    } while (0)
}  // namespace

extern "C" int rb200_comm_unique_id(uint8_t id[RB200_NCCL_ID_BYTES])
{
    static_assert(sizeof(ncclUniqueId) == RB200_NCCL_ID_BYTES, "ncclUniqueId is 128 bytes");
    const rb200::NcclApi& api = rb200::nccl_api();
    if (!id) return RB200_E_ARG;
    if (!api.ok) return RB200_E_NCCL;
    ncclUniqueId u;
    if (api.GetUniqueId(&u) != ncclSuccess) return RB200_E_NCCL;
    std::memcpy(id, &u, sizeof u);
    return RB200_OK;
}

extern "C" int rb200_comm_create(rb200_ctx* ctx, const uint8_t id[RB200_NCCL_ID_BYTES], int rank, int world, rb200_comm** out)
{
    if (!ctx || !id || !out || world < 1 || rank < 0 || rank >= world) return RB200_E_ARG;
    *out = nullptr;
    const rb200::NcclApi& api = rb200::nccl_api();
    if (!api.ok) return nccl_fail(ctx, ncclSystemError, "dlopen(libnccl.so.2)");
    DeviceGuard g(ctx->device);
    rb200_comm* c = new (std::nothrow) rb200_comm;
    if (!c) return RB200_E_NOMEM;
    c->ctx = ctx;
    c->rank = rank;
    c->world = world;
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof u);
    ncclResult_t r = api.CommInitRank(&c->comm, world, u, rank);
    if (r != ncclSuccess) {
        delete c;
        return nccl_fail(ctx, r, "ncclCommInitRank");
    }
    cudaError_t e = cudaMalloc(&c->d_sizes, (static_cast<size_t>(world) + 1) * 2 * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMallocHost(&c->h_sizes, (static_cast<size_t>(world) + 1) * 2 * sizeof(uint64_t));
    if (e != cudaSuccess) {
        rb200_comm_destroy(c);
        return cuda_fail(ctx, e, "rb200_comm_create");
    }
    c->blob_bytes.assign(world, 0);
    c->chunks.assign(world, 0);
    *out = c;
    return RB200_OK;
}

extern "C" void rb200_comm_destroy(rb200_comm* c)
{
    if (!c) return;
    DeviceGuard g(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    if (c->comm) rb200::nccl_api().CommDestroy(c->comm);
    if (c->d_sizes) cudaFree(c->d_sizes);
    if (c->h_sizes) cudaFreeHost(c->h_sizes);
    delete c;
}

extern "C" int rb200_gather_plan(rb200_comm* c, uint64_t blob_size, uint64_t n_chunks, uint64_t totals[2])
{
    if (!c || !totals || (blob_size & 15)) return RB200_E_ARG;
    rb200_ctx* ctx = c->ctx;
    const rb200::NcclApi& api = rb200::nccl_api();
    DeviceGuard g(ctx->device);
    uint64_t* mine = c->h_sizes + 2 * static_cast<size_t>(c->world);
    mine[0] = blob_size;
    mine[1] = n_chunks;
    RB_CUDA(ctx, cudaMemcpyAsync(c->d_sizes + 2 * static_cast<size_t>(c->world), mine, 2 * sizeof(uint64_t), cudaMemcpyHostToDevice,
                                 ctx->stream));
    RB_NCCL(ctx, api.AllGather(c->d_sizes + 2 * static_cast<size_t>(c->world), c->d_sizes, 2, ncclUint64, c->comm, ctx->stream));
    RB_CUDA(ctx, cudaMemcpyAsync(c->h_sizes, c->d_sizes, static_cast<size_t>(c->world) * 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                                 ctx->stream));
    RB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    totals[0] = totals[1] = 0;
    for (int r = 0; r < c->world; r++) {
        c->blob_bytes[r] = c->h_sizes[2 * r];
        c->chunks[r] = c->h_sizes[2 * r + 1];
        if (c->blob_bytes[r] & 15) return RB200_E_STREAM;      // every shard must end 16-byte aligned to concatenate
        totals[0] += c->blob_bytes[r];
        totals[1] += c->chunks[r];
    }
    c->planned = true;
    return RB200_OK;
}

extern "C" int rb200_gather_blobs(rb200_comm* c, int root, const uint8_t* d_blob, const uint64_t* d_offsets, uint8_t* d_out_blob,
                                  uint64_t out_cap, uint64_t* d_out_offsets)
{
    if (!c || !c->planned || root < 0 || root >= c->world) return RB200_E_ARG;
    rb200_ctx* ctx = c->ctx;
    const rb200::NcclApi& api = rb200::nccl_api();
    const uint64_t my_bytes = c->blob_bytes[c->rank], my_chunks = c->chunks[c->rank];
    if ((my_bytes && !d_blob) || (my_chunks && !d_offsets)) return RB200_E_ARG;
    DeviceGuard g(ctx->device);
    if (c->rank != root) {
        // payload and directory go straight to their final position on the root; exact sizes, no padding
        RB_NCCL(ctx, api.GroupStart());
        if (my_bytes) RB_NCCL(ctx, api.Send(d_blob, my_bytes, ncclUint8, root, c->comm, ctx->stream));
        if (my_chunks) RB_NCCL(ctx, api.Send(d_offsets, my_chunks, ncclUint64, root, c->comm, ctx->stream));
        RB_NCCL(ctx, api.GroupEnd());
        return RB200_OK;
    }
    uint64_t total_bytes = 0, total_chunks = 0;
    for (int r = 0; r < c->world; r++) {
        total_bytes += c->blob_bytes[r];
        total_chunks += c->chunks[r];
    }
    if (!d_out_offsets || (total_bytes && !d_out_blob)) return RB200_E_ARG;
    if (total_bytes > out_cap) return RB200_E_SPACE;
    RB_NCCL(ctx, api.GroupStart());
    uint64_t bbase = 0, cbase = 0;
    for (int r = 0; r < c->world; r++) {
        if (r != root) {
            if (c->blob_bytes[r]) RB_NCCL(ctx, api.Recv(d_out_blob + bbase, c->blob_bytes[r], ncclUint8, r, c->comm, ctx->stream));
            if (c->chunks[r]) RB_NCCL(ctx, api.Recv(d_out_offsets + cbase, c->chunks[r], ncclUint64, r, c->comm, ctx->stream));
        }
        bbase += c->blob_bytes[r];
        cbase += c->chunks[r];
    }
    RB_NCCL(ctx, api.GroupEnd());
    // the root's own shard, and every directory rebased by the bytes of the shards before it
    bbase = cbase = 0;
    for (int r = 0; r < c->world; r++) {
        const uint64_t nb = c->blob_bytes[r], nc = c->chunks[r];
        if (r == root && nb) RB_CUDA(ctx, cudaMemcpyAsync(d_out_blob + bbase, d_blob, nb, cudaMemcpyDeviceToDevice, ctx->stream));
        if (nc) {
            const uint64_t* src = r == root ? d_offsets : d_out_offsets + cbase;
            rb200::rebase_directory_kernel<<<static_cast<unsigned>((nc + 255) / 256), 256, 0, ctx->stream>>>(src, d_out_offsets + cbase, nc,
                                                                                                               bbase);
            int rc = check_launch(ctx, "rebase_directory_kernel");
            if (rc != RB200_OK) return rc;
        }
        bbase += nb;
        cbase += nc;
    }
    rb200::set_u64_kernel<<<1, 1, 0, ctx->stream>>>(d_out_offsets + total_chunks, total_bytes);
    return check_launch(ctx, "set_u64_kernel");
}
