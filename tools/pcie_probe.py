"""Host-path probe: what bounds `e2e`?  (run on the GPU box: python tools/pcie_probe.py)

Measures pinned H2D / D2H / bidirectional copy bandwidth for buffers first-touched on each NUMA
node, then the host-pointer encode and decode calls separately, for several pipeline slice sizes
(RB200_SLICE_MIB is read once per process, so each setting runs in a child process).
"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def node_cpus():
    nodes = {}
    base = "/sys/devices/system/node"
    for d in sorted(os.listdir(base)):
        if d.startswith("node") and d[4:].isdigit():
            cpus = set()
            for part in open(f"{base}/{d}/cpulist").read().strip().split(","):
                if "-" in part:
                    a, b = part.split("-")
                    cpus |= set(range(int(a), int(b) + 1))
                elif part:
                    cpus.add(int(part))
            nodes[int(d[4:])] = cpus
    return nodes


def gpu_node(idx=0):
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(idx)],
                             capture_output=True, text=True).stdout.strip().lower()
        bus = bus[4:] if len(bus) > 12 else bus       # 00000000:1B:00.0 -> 0000:1b:00.0
        return int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
    except Exception as e:  # noqa: BLE001
        return f"unknown ({e})"


def bw(fn, nbytes, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


def copies():
    n = 1 << 30
    dev = torch.device("cuda:0")
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s2 = torch.cuda.Stream()
    all_cpus = os.sched_getaffinity(0)
    out = {"gpu0_numa_node": gpu_node(0)}
    for node, cpus in node_cpus().items():
        cpus = cpus & all_cpus
        if not cpus:
            continue
        os.sched_setaffinity(0, cpus)
        h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_a.fill_(1); h_b.fill_(2)

        def h2d():
            d_a.copy_(h_a, non_blocking=True)

        def d2h():
            h_b.copy_(d_b, non_blocking=True)

        def both():
            d_a.copy_(h_a, non_blocking=True)
            with torch.cuda.stream(s2):
                h_b.copy_(d_b, non_blocking=True)

        r = {"h2d_GBs": round(bw(h2d, n), 1), "d2h_GBs": round(bw(d2h, n), 1), "bidir_each_GBs": round(bw(both, n), 1)}
        # each direction's own rate while the other one is busy (the slower one is what a pipeline sees):
        # keep the other direction running for longer than the timed copies
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        torch.cuda.synchronize()
        with torch.cuda.stream(s2):
            for _ in range(8):
                h_b.copy_(d_b, non_blocking=True)
        time.sleep(0.02)
        ev[0].record()
        for _ in range(3):
            d_a.copy_(h_a, non_blocking=True)
        ev[1].record()
        torch.cuda.synchronize()
        r["h2d_while_d2h_GBs"] = round(3 * n / ev[0].elapsed_time(ev[1]) / 1e6, 1)
        for _ in range(8):
            d_a.copy_(h_a, non_blocking=True)
        time.sleep(0.02)
        with torch.cuda.stream(s2):
            ev[2].record()
            for _ in range(3):
                h_b.copy_(d_b, non_blocking=True)
            ev[3].record()
        torch.cuda.synchronize()
        r["d2h_while_h2d_GBs"] = round(3 * n / ev[2].elapsed_time(ev[3]) / 1e6, 1)
        out[f"node{node}"] = r
        del h_a, h_b
    os.sched_setaffinity(0, all_cpus)
    return out


def calls(pin_node):
    import ryg_rans_b200 as rb
    if pin_node >= 0:
        os.sched_setaffinity(0, node_cpus()[pin_node] & os.sched_getaffinity(0))
    n, chunk = 1 << 30, 8192
    rng = np.random.default_rng(1)
    data = torch.from_numpy(rng.integers(0, 256, n, dtype=np.uint8))
    ctx = rb.Context(0)
    stats = rb.SymbolStats().count_freqs(data[: 1 << 24].numpy())
    stats.normalize_freqs(1 << 12)
    model = rb.Model(ctx, rb.CODER_WORD, 12, stats.freqs)
    lib = ctx.lib
    cap = lib.dll.rb200_encode_bound(n, chunk)
    n_chunks = lib.dll.rb200_chunk_count(n, chunk)
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_in.copy_(data)
    h_blob = torch.empty(cap, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    res = {}
    for offs_kind in ("pageable", "pinned"):
        if offs_kind == "pinned":
            t_off = torch.zeros(n_chunks + 1, dtype=torch.int64).pin_memory()
            off_ptr = t_off.data_ptr()
        else:
            a_off = np.zeros(n_chunks + 1, np.uint64)
            off_ptr = a_off.ctypes.data
        size = C.c_size_t(0)

        def enc():
            lib.check(lib.dll.rb200_encode(ctx.h, model.h, h_in.data_ptr(), n, chunk, h_blob.data_ptr(), cap, off_ptr,
                                           C.byref(size), rb.MEM_HOST), ctx.h)

        def dec():
            lib.check(lib.dll.rb200_decode(ctx.h, model.h, h_blob.data_ptr(), size.value, off_ptr, chunk, h_out.data_ptr(), n,
                                           rb.MEM_HOST), ctx.h)
        enc(); dec()
        te, td = [], []
        for _ in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter(); enc(); torch.cuda.synchronize(); te.append(time.perf_counter() - t0)
            t0 = time.perf_counter(); dec(); torch.cuda.synchronize(); td.append(time.perf_counter() - t0)
        assert torch.equal(h_out, h_in)
        res[offs_kind] = {"encode_ms": round(min(te) * 1e3, 2), "decode_ms": round(min(td) * 1e3, 2),
                          "round_trip_gsym_s": round(n / (min(te) + min(td)) / 1e9, 2), "blob_bytes": size.value}
    return res


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "calls":
        print(json.dumps(calls(int(sys.argv[2]))))
        sys.exit(0)
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
    print(json.dumps({"copies": copies()}))
    nodes = sorted(node_cpus())
    if "--copies-only" in sys.argv:
        sys.exit(0)
    for node in [-1] + nodes:
        for mib in (8, 16, 32, 64):
            if node != -1 and mib not in (16, 32):
                continue
            env = dict(os.environ, RB200_SLICE_MIB=str(mib))
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "calls", str(node)], env=env, capture_output=True, text=True)
            print(json.dumps({"pin_node": node, "slice_mib": mib, "result": (r.stdout.strip().splitlines() or [r.stderr[-400:]])[-1]}))
