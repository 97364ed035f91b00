"""Multi-GPU plumbing for the rANS hot path: shards are independent, so the
only exchange step is gathering the per-rank compressed blobs (+ their directories) after
encode.  The product path is the C-ABI (rb200_comm_create / rb200_gather_plan /
rb200_gather_blobs: NCCL send/recv straight into the final position, include/rans_b200.h);
`NcclGather` below is its ctypes view for bench.py and the tests, with torch.distributed only
carrying the 128-byte NCCL id from rank 0 to the others.  `gather_blobs` is the same exchange
written against torch.distributed, kept for the gloo CPU tests of the host-side logic (shard
boundaries, directory rebasing).  No compute happens here.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


class NcclGather:
    """rb200_comm over the ranks of the default process group (one process per GPU)."""

    def __init__(self, ctx, rank, world, dist_mod=None):
        d = dist_mod or dist
        self.ctx, self.lib, self.rank, self.world = ctx, ctx.lib, rank, world
        ident = np.zeros(128, np.uint8)
        if rank == 0:
            self.lib.check(self.lib.dll.rb200_comm_unique_id(ident.ctypes.data), ctx.h)
        t = torch.from_numpy(ident)
        if d.get_backend() == "nccl":
            t = t.cuda()
        d.broadcast(t, src=0)                  # the side channel for the id: any transport would do
        ident = t.cpu().numpy()
        h = C.c_void_p()
        self.lib.check(self.lib.dll.rb200_comm_create(ctx.h, ident.ctypes.data, rank, world, C.byref(h)), ctx.h)
        self.h = h

    def total_bytes(self, blob_size, n_chunks):
        """Collective: (gathered blob bytes, gathered chunk count), valid on every rank."""
        totals = np.zeros(2, np.uint64)
        self.lib.check(self.lib.dll.rb200_gather_plan(self.h, blob_size, n_chunks, totals.ctypes.data_as(C.POINTER(C.c_uint64))),
                       self.ctx.h)
        return int(totals[0]), int(totals[1])

    def gather(self, blob_ptr, blob_size, offsets_ptr, n_chunks, out_blob_ptr, out_cap, out_offsets_ptr, root=0):
        self.lib.check(self.lib.dll.rb200_gather_blobs(self.h, root, blob_ptr, offsets_ptr, out_blob_ptr, out_cap, out_offsets_ptr),
                       self.ctx.h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.dll.rb200_comm_destroy(self.h)
            self.h = None


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
    rank = dist.get_rank(group)                       # group-local; `dst` is a group rank too
    peer = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)   # P2POp wants global ranks
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
            ops.append(dist.P2POp(dist.isend, blob, peer(dst), group))
        ops.append(dist.P2POp(dist.isend, offsets, peer(dst), group))
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
                ops.append(dist.P2POp(dist.irecv, gblob[base:base + sizes[r]], peer(r), group))
            ops.append(dist.P2POp(dist.irecv, tmp_dirs[r], peer(r), group))
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
