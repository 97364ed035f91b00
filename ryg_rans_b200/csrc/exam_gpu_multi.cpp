// exam_gpu_multi.cpp -- the sharded form of the reference's driver loop (main_simd.cpp:287-332), one process per GPU.
//
// Every rank owns one contiguous shard of the symbols (cut on chunk boundaries), builds its own order-0 model,
// encodes its shard with rb200_encode (device buffers) and takes part in rb200_gather_blobs; rank 0 ends up with
// the concatenated container, decodes every shard from it with that shard's model and compares with the input.
// No communication on the data path: the gather is the only exchange.  Host code stays C++.
//
//   exam_gpu_multi WORLD [word|alias] [chunk_syms] [total_symbols]      parent: spawns WORLD ranks of itself
//   exam_gpu_multi --rank R --world W --idfile PATH ...                 one rank (also usable under any launcher)
//
// The 128-byte NCCL id travels from rank 0 to the others through a file (any side channel would do).
#include <cuda_runtime.h>
#include <spawn.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "rans_b200.h"

extern char** environ;

static void die(int rank, const char* what, int rc, rb200_ctx* ctx)
{
    std::fprintf(stderr, "[rank %d] Error: %s: %s %s\n", rank, what, rb200_strerror(rc), ctx ? rb200_last_cuda_error(ctx) : "");
    std::exit(1);
}
#define RBX(call)                                        \
    do {                                                 \
        int rc_ = (call);                                \
        if (rc_ != RB200_OK) die(rank, #call, rc_, ctx); \
    } while (0)
#define CUX(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            std::fprintf(stderr, "[rank %d] Error: %s: %s\n", rank, #call, cudaGetErrorString(e_));       \
            std::exit(1);                                                                                \
        }                                                                                                \
    } while (0)

// the same seeded skewed bytes for every process (splitmix64; AND of two uniform bytes), position-addressable
static uint8_t synth_byte(uint64_t i)
{
    uint64_t z = 0x5EED0000ull + (i + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return static_cast<uint8_t>((z & 0xff) & ((z >> 8) & 0xff));
}

static int run_rank(int rank, int world, const std::string& idfile, bool alias, uint32_t chunk, uint64_t total)
{
    rb200_ctx* ctx = nullptr;
    int n_dev = 0;
    CUX(cudaGetDeviceCount(&n_dev));
    if (n_dev < world) {
        std::fprintf(stderr, "[rank %d] Error: %d GPUs visible, %d needed\n", rank, n_dev, world);
        return 1;
    }
    CUX(cudaSetDevice(rank));
    RBX(rb200_ctx_create(&ctx, rank, nullptr));

    // the id: rank 0 creates it, the others wait for the file
    uint8_t id[RB200_NCCL_ID_BYTES];
    if (rank == 0) {
        RBX(rb200_comm_unique_id(id));
        const std::string tmp = idfile + ".tmp";
        FILE* f = std::fopen(tmp.c_str(), "wb");
        if (!f || std::fwrite(id, 1, sizeof id, f) != sizeof id) { std::fprintf(stderr, "cannot write %s\n", tmp.c_str()); return 1; }
        std::fclose(f);
        std::rename(tmp.c_str(), idfile.c_str());
    } else {
        FILE* f = nullptr;
        for (int tries = 0; tries < 6000 && !(f = std::fopen(idfile.c_str(), "rb")); tries++)
            std::this_thread::sleep_for(std::chrono::milliseconds(10));
        if (!f || std::fread(id, 1, sizeof id, f) != sizeof id) { std::fprintf(stderr, "[rank %d] no NCCL id in %s\n", rank, idfile.c_str()); return 1; }
        std::fclose(f);
    }
    rb200_comm* comm = nullptr;
    RBX(rb200_comm_create(ctx, id, rank, world, &comm));

    // shard [lo, hi): whole chunks, so that the concatenated containers form one container
    const uint64_t n_chunks_total = (total + chunk - 1) / chunk;
    const uint64_t per = (n_chunks_total + world - 1) / world;
    auto bound = [&](int r) { uint64_t v = static_cast<uint64_t>(r) * per * chunk; return v < total ? v : total; };
    const uint64_t lo = bound(rank), hi = bound(rank + 1), n = hi - lo;
    std::vector<uint8_t> h_in(n);
    for (uint64_t i = 0; i < n; i++) h_in[i] = synth_byte(lo + i);

    const int coder = alias ? RB200_CODER_ALIAS : RB200_CODER_WORD;
    const uint32_t scale_bits = alias ? 16 : 12;
    uint8_t *d_in = nullptr, *d_blob = nullptr;
    uint64_t* d_off = nullptr;
    const size_t n_chunks = rb200_chunk_count(n, chunk), cap = rb200_encode_bound(n, chunk);
    CUX(cudaMalloc(&d_in, n + 16));
    CUX(cudaMalloc(&d_blob, cap + 16));
    CUX(cudaMalloc(&d_off, (n_chunks + 1) * sizeof(uint64_t)));
    CUX(cudaMemcpy(d_in, h_in.data(), n, cudaMemcpyHostToDevice));
    rb200_model* model = nullptr;
    uint32_t freqs[256];
    RBX(rb200_model_from_data(ctx, coder, scale_bits, d_in, n, RB200_MEM_DEVICE, freqs, &model));
    RBX(rb200_encode(ctx, model, d_in, n, chunk, d_blob, cap, d_off, nullptr, RB200_MEM_DEVICE));
    RBX(rb200_sync(ctx));
    uint64_t blob_size = 0;
    CUX(cudaMemcpy(&blob_size, d_off + n_chunks, sizeof blob_size, cudaMemcpyDeviceToHost));

    uint64_t totals[2];
    RBX(rb200_gather_plan(comm, blob_size, n_chunks, totals));
    uint8_t* d_all = nullptr;
    uint64_t* d_all_off = nullptr;
    if (rank == 0) {
        CUX(cudaMalloc(&d_all, totals[0] + 16));
        CUX(cudaMalloc(&d_all_off, (totals[1] + 1) * sizeof(uint64_t)));
    }
    RBX(rb200_gather_blobs(comm, 0, d_blob, d_off, d_all, totals[0], d_all_off));
    RBX(rb200_sync(ctx));
    std::printf("[rank %d] shard [%llu, %llu): %llu symbols -> %llu bytes in %llu chunks\n", rank, (unsigned long long)lo, (unsigned long long)hi,
                (unsigned long long)n, (unsigned long long)blob_size, (unsigned long long)n_chunks);

    int status = 0;
    if (rank == 0) {
        // decode every shard out of the gathered container: shard r's chunks are [c0, c0 + k) of the global directory;
        // its model is rebuilt from the same bytes (a real application ships the 256 frequencies with the shard)
        std::vector<uint64_t> h_dir(totals[1] + 1);
        CUX(cudaMemcpy(h_dir.data(), d_all_off, h_dir.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        if (h_dir[totals[1]] != totals[0]) { std::fprintf(stderr, "gathered directory does not end at the gathered size\n"); status = 1; }
        uint64_t c0 = 0;
        for (int r = 0; r < world && !status; r++) {
            const uint64_t rlo = bound(r), rhi = bound(r + 1), rn = rhi - rlo;
            const size_t rc_n = rb200_chunk_count(rn, chunk);
            std::vector<uint8_t> want(rn), got(rn);
            for (uint64_t i = 0; i < rn; i++) want[i] = synth_byte(rlo + i);
            uint8_t *d_w = nullptr, *d_o = nullptr;
            CUX(cudaMalloc(&d_w, rn + 16));
            CUX(cudaMalloc(&d_o, rn + 16));
            CUX(cudaMemcpy(d_w, want.data(), rn, cudaMemcpyHostToDevice));
            rb200_model* m = nullptr;
            RBX(rb200_model_from_data(ctx, coder, scale_bits, d_w, rn, RB200_MEM_DEVICE, nullptr, &m));
            RBX(rb200_decode(ctx, m, d_all, totals[0], d_all_off + c0, chunk, d_o, rn, RB200_MEM_DEVICE));
            RBX(rb200_sync(ctx));
            CUX(cudaMemcpy(got.data(), d_o, rn, cudaMemcpyDeviceToHost));
            if (std::memcmp(got.data(), want.data(), rn) != 0) { std::fprintf(stderr, "shard %d: decoded bytes differ\n", r); status = 1; }
            rb200_model_destroy(m);
            cudaFree(d_w);
            cudaFree(d_o);
            c0 += rc_n;
        }
        if (!status)
            std::printf("gathered %llu bytes / %llu chunks from %d ranks: decode ok!\n", (unsigned long long)totals[0],
                        (unsigned long long)totals[1], world);
    }
    rb200_model_destroy(model);
    rb200_comm_destroy(comm);
    rb200_ctx_destroy(ctx);
    return status;
}

int main(int argc, char** argv)
{
    int rank = -1, world = 0;
    std::string idfile;
    std::vector<std::string> rest;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "--rank" && i + 1 < argc) rank = std::atoi(argv[++i]);
        else if (a == "--world" && i + 1 < argc) world = std::atoi(argv[++i]);
        else if (a == "--idfile" && i + 1 < argc) idfile = argv[++i];
        else rest.push_back(a);
    }
    if (rank < 0 && !rest.empty()) {      // parent form: first positional argument is the world size
        world = std::atoi(rest[0].c_str());
        rest.erase(rest.begin());
    }
    const bool alias = !rest.empty() && rest[0] == "alias";
    const uint32_t chunk = rest.size() > 1 ? static_cast<uint32_t>(std::atoi(rest[1].c_str())) : 8192;
    const uint64_t total = rest.size() > 2 ? std::strtoull(rest[2].c_str(), nullptr, 0) : (64ull << 20);
    if (world < 1 || chunk < 32 || chunk % 32) {
        std::fprintf(stderr, "usage: exam_gpu_multi WORLD [word|alias] [chunk_syms] [total_symbols]\n");
        return 2;
    }
    if (rank >= 0) return run_rank(rank, world, idfile, alias, chunk, total);

    // parent: one child process per GPU (no CUDA in this process)
    char tmpl[] = "/tmp/rb200_nccl_id_XXXXXX";
    const int fd = mkstemp(tmpl);
    if (fd >= 0) close(fd);
    unlink(tmpl);
    std::vector<pid_t> pids;
    for (int r = 0; r < world; r++) {
        std::vector<std::string> args = {argv[0], "--rank", std::to_string(r), "--world", std::to_string(world), "--idfile", tmpl,
                                         alias ? "alias" : "word", std::to_string(chunk), std::to_string(total)};
        std::vector<char*> av;
        for (auto& s : args) av.push_back(const_cast<char*>(s.c_str()));
        av.push_back(nullptr);
        pid_t pid = 0;
        if (posix_spawn(&pid, argv[0], nullptr, nullptr, av.data(), environ) != 0) { std::perror("posix_spawn"); return 1; }
        pids.push_back(pid);
    }
    int bad = 0;
    for (pid_t pid : pids) {
        int st = 0;
        waitpid(pid, &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) bad++;
    }
    unlink(tmpl);
    if (bad) std::fprintf(stderr, "Error: %d of %d ranks failed\n", bad, world);
    return bad ? 1 : 0;
}
