"""N>1 host logic on CPU: world_size-2 gloo run of the shard/gather plumbing.
The per-rank 'encode' here is the ORACLE (this is a test), so the gathered container can be
checked end to end without a GPU: it must decode, with the global directory, to the full input."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, chunk, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from ryg_rans_b200.shard import gather_blobs, shard_bounds
    orc = oracle.Oracle()
    rng = np.random.default_rng(5)
    p = 1.0 / np.arange(1, 257) ** 1.1
    data = rng.choice(256, n_total, p=p / p.sum()).astype(np.uint8)     # same on every rank
    freqs, cum = orc.model(data, 12)
    lo, hi = shard_bounds(n_total, world, rank, chunk)
    blob, offs = orc.chunked_encode(oracle.CODER_WORD, data[lo:hi], freqs, cum, chunk)
    gblob, gdir = gather_blobs(torch.from_numpy(blob), torch.from_numpy(offs.astype(np.int64)), dst=0)
    if rank == 0:
        out = orc.chunked_decode(oracle.CODER_WORD, gblob.numpy(), gdir.numpy().astype(np.uint64), n_total, freqs, cum, chunk)
        whole, woffs = orc.chunked_encode(oracle.CODER_WORD, data, freqs, cum, chunk)
        q.put((bool(np.array_equal(out, data)), bool(np.array_equal(gblob.numpy(), whole)),
               bool(np.array_equal(gdir.numpy().astype(np.uint64), woffs))))
    else:
        assert gblob is None and gdir is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total,chunk", [(100000, 4096), (4096 * 3 + 17, 4096), (5000, 8192)])
def test_two_rank_gather_reassembles_the_container(n_total, chunk):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, chunk, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == (True, True, True), res


def test_shard_bounds_cover_everything():
    from ryg_rans_b200.shard import shard_bounds
    for n, w, c in [(0, 4, 64), (1, 8, 64), (1000, 3, 64), (1 << 20, 8, 4096), (4097, 2, 4096)]:
        pieces = [shard_bounds(n, w, r, c) for r in range(w)]
        assert pieces[0][0] == 0 and pieces[-1][1] == n
        for (a, b), (c2, d) in zip(pieces, pieces[1:]):
            assert b == c2 and a <= b
        assert all(lo % c == 0 or lo == n for lo, _ in pieces)
