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
    rb200 directory because every blob ends 16-byte aligned); elsewhere (None, None)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = blob.device
    mine = torch.tensor([blob.numel(), offsets.numel()], dtype=torch.int64, device=dev)
    allsz = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(allsz, mine, group=group)
    sizes = [int(t[0]) for t in allsz]
    n_offs = [int(t[1]) for t in allsz]
    assert all(s % 16 == 0 for s in sizes), "container invariant: every blob ends on a 16-byte boundary"
    max_b, max_o = max(sizes), max(n_offs)
    pad_b = torch.zeros(max_b, dtype=torch.uint8, device=dev)
    pad_b[:blob.numel()] = blob
    pad_o = torch.zeros(max_o, dtype=torch.int64, device=dev)
    pad_o[:offsets.numel()] = offsets
    if rank == dst:
        gb = [torch.empty(max_b, dtype=torch.uint8, device=dev) for _ in range(world)]
        go = [torch.empty(max_o, dtype=torch.int64, device=dev) for _ in range(world)]
    else:
        gb = go = None
    dist.gather(pad_b, gb, dst=dst, group=group)
    dist.gather(pad_o, go, dst=dst, group=group)
    if rank != dst:
        return None, None
    base = 0
    blobs, dirs = [], []
    for r in range(world):
        blobs.append(gb[r][:sizes[r]])
        dirs.append(go[r][:n_offs[r] - 1] + base)
        base += sizes[r]
    dirs.append(torch.tensor([base], dtype=torch.int64, device=dev))
    return torch.cat(blobs), torch.cat(dirs)
