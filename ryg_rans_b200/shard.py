"""Multi-GPU plumbing for the rANS hot path: shards are independent, so the
only exchange step is gathering the per-rank compressed blobs (+ their directories) after
encode.  Works on any torch.distributed backend: NCCL over NVLink on the GPU box, gloo in
the CPU tests.  No compute happens here.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_total, world, rank, chunk_syms):
    """Contiguous shard [lo, hi) for `rank`; boundaries fall on chunk boundaries so that the
    concatenation of the per-rank containers is itself a valid container."""
    n_chunks = (n_total + chunk_syms - 1) // chunk_syms
    per = (n_chunks + world - 1) // world
    lo = min(n_total, rank * per * chunk_syms)
    hi = min(n_total, (rank + 1) * per * chunk_syms)
    return lo, hi


def gather_blobs(blob, offsets, dst=0, group=None):
    """blob: 1-D uint8 tensor holding this rank's container (size = offsets[-1], a multiple
    of 16); offsets: 1-D int64 tensor [n_chunks + 1].  Returns on `dst` the concatenated
    container and its global directory (offsets shifted by the preceding blob sizes, a valid
    rb200 directory because every blob ends 16-byte aligned); elsewhere (None, None).

    One all_gather of the sizes, then exact-size point-to-point transfers straight into their
    final position in the destination buffers (no padding, no staging copies)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = blob.device
    mine = torch.tensor([blob.numel(), offsets.numel()], dtype=torch.int64, device=dev)
    allsz = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(allsz, mine, group=group)
    sizes = [int(t[0]) for t in allsz]
    n_offs = [int(t[1]) for t in allsz]
    assert all(s % 16 == 0 for s in sizes), "container invariant: every blob ends on a 16-byte boundary"
    blob = blob.contiguous()
    offsets = offsets.contiguous()
    if rank != dst:
        ops = []
        if sizes[rank]:
            ops.append(dist.P2POp(dist.isend, blob, dst, group))
        ops.append(dist.P2POp(dist.isend, offsets, dst, group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        return None, None
    gblob = torch.empty(sum(sizes), dtype=torch.uint8, device=dev)
    gdir = torch.empty(sum(n - 1 for n in n_offs) + 1, dtype=torch.int64, device=dev)
    tmp_dirs = [torch.empty(n, dtype=torch.int64, device=dev) for n in n_offs]
    ops, base = [], 0
    for r in range(world):
        if r == dst:
            gblob[base:base + sizes[r]] = blob
            tmp_dirs[r].copy_(offsets)
        else:
            if sizes[r]:
                ops.append(dist.P2POp(dist.irecv, gblob[base:base + sizes[r]], r, group))
            ops.append(dist.P2POp(dist.irecv, tmp_dirs[r], r, group))
        base += sizes[r]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    base, pos = 0, 0
    for r in range(world):
        k = n_offs[r] - 1
        gdir[pos:pos + k] = tmp_dirs[r][:k] + base
        pos += k
        base += sizes[r]
    gdir[pos] = base
    return gblob, gdir
