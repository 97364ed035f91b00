#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 rANS hot path (BASELINE.json metric).

Step      = one bit-exact round trip of the hot path over one batch: encode the
            batch (rb200_encode: encode kernel + directory scan + compaction), then
            decode it (rb200_decode: one kernel), device-resident, through the C-ABI.
Workload  = BASELINE.json configs[1]: 1 GiB i.i.d. uniform bytes, one static
            256-symbol model (scale_bits 12), word coder, 32-way interleaved chunks.
            Per-GPU work is fixed (weak scaling): every rank round-trips its own shard.
value     = symbols round-tripped per second over all ranks (Gsymbols/s, 1 symbol = 1 byte)
e2e       = the same round trip through the host-pointer C-ABI calls (pinned host
            buffers, H2D/D2H inside the timed region).
roofline  = the decode kernel (the north-star kernel) against measured HBM copy bandwidth;
            the encode call is reported beside it.
--impl reference times the reference's own CPU coders (oracle/_ref, built from the
reference sources) on all host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gsymbols/s decode+encode (bit-exact round-trip)"
UNIT = "Gsymbols/s"
WORKLOADS = {
    # name: (coder, scale_bits, generator)
    "uniform_1GiB_word32": ("word", 12, "uniform"),
    "zipf1.1_1GiB_alias32": ("alias", 16, "zipf"),
    "text_1GiB_word32": ("word", 12, "text"),            # BASELINE configs[3] = this at --gpus 8 (8 x 1 GiB shards)
    "blocks_64KiB_word32": ("blocks", 12, "blocks"),     # BASELINE configs[4]: 64 KiB blocks, one model per block
}
BLOCK_SIZE = 65536


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="uniform_1GiB_word32", choices=sorted(WORKLOADS))
    ap.add_argument("--size", type=int, default=1 << 30, help="symbols per GPU")
    ap.add_argument("--chunk", type=int, default=int(os.environ.get("RB200_CHUNK", 8192)), help="symbols per 32-way chunk stream")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=256 << 20, help="bytes of the workload the CPU baseline is timed on")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------ synthetic data

def text_probs():
    """Order-0 distribution of the reference's test file book1 (82 symbols, 4.527 bit/symbol),
    from the committed fixture tests/golden/book1_hist.json (made by tests/golden/make_book1_hist.py)."""
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "golden", "book1_hist.json")) as f:
        counts = np.asarray(json.load(f)["counts"], dtype=np.float64)
    return counts / counts.sum()


def synth_torch(kind, n, seed, device):
    """Seeded synthetic symbols generated on the device (so 1 GiB need not cross PCIe)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(0x5EED0000 + seed)
    if kind == "uniform":
        return torch.randint(0, 256, (n,), dtype=torch.uint8, device=device, generator=g)
    if kind == "blocks":
        # every 64 KiB block has its own distribution: Zipf ranks pushed through a per-block permutation
        nb = n // BLOCK_SIZE
        p = 1.0 / torch.arange(1, 257, dtype=torch.float64) ** 1.3
        cdf = torch.cumsum(p / p.sum(), 0).to(device=device, dtype=torch.float32)
        out = torch.empty(n, dtype=torch.uint8, device=device)
        perms = torch.argsort(torch.rand(nb, 256, device=device, generator=g), dim=1).to(torch.uint8)
        step_blocks = 1024
        for b0 in range(0, nb, step_blocks):
            b1 = min(nb, b0 + step_blocks)
            u = torch.rand((b1 - b0) * BLOCK_SIZE, device=device, generator=g)
            ranks = torch.searchsorted(cdf, u).clamp_(max=255).view(b1 - b0, BLOCK_SIZE)
            out[b0 * BLOCK_SIZE:b1 * BLOCK_SIZE] = torch.gather(perms[b0:b1], 1, ranks).reshape(-1)
        return out
    if kind == "zipf":
        p = 1.0 / torch.arange(1, 257, dtype=torch.float64) ** 1.1
    elif kind == "text":
        p = torch.from_numpy(text_probs())
    else:
        raise ValueError(kind)
    cdf = torch.cumsum(p / p.sum(), 0).to(device=device, dtype=torch.float32)
    out = torch.empty(n, dtype=torch.uint8, device=device)
    step = 1 << 26
    for lo in range(0, n, step):
        m = min(step, n - lo)
        u = torch.rand(m, device=device, generator=g)
        out[lo:lo + m] = torch.searchsorted(cdf, u).clamp_(max=255).to(torch.uint8)
    return out


def synth_numpy(kind, n, seed):
    rng = np.random.default_rng(0x5EED0000 + seed)
    if kind == "uniform":
        return rng.integers(0, 256, n, dtype=np.uint8)
    if kind == "zipf":
        p = 1.0 / np.arange(1, 257) ** 1.1
    else:
        p = text_probs()
    cdf = np.cumsum(p / p.sum())
    out = np.empty(n, np.uint8)
    step = 1 << 24
    for lo in range(0, n, step):
        m = min(step, n - lo)
        out[lo:lo + m] = np.minimum(np.searchsorted(cdf, rng.random(m)), 255).astype(np.uint8)
    return out


# ------------------------------------------------------------------ clocks

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic(workload, chunk, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json), or None when no capture exists for this workload/chunk."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        if t["workload"] == workload and t["chunk_syms"] == chunk:
            return t["dram_bytes_per_launch"].get(kernel)
    except Exception:
        pass
    return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------ reference arm / cpu baseline

def cpu_reference_run(kind, coder, scale_bits, nbytes, runs, threads):
    """Time the reference's own CPU coders (oracle/_ref) on `nbytes` of the workload."""
    import oracle
    data = synth_numpy(kind, nbytes, seed=1)
    if oracle.Reference.available():
        ref = oracle.Reference()
        which = "alias" if coder == "alias" else "simd"
        r = ref.cpu_baseline(which, data, threads, runs=runs, scale_bits=scale_bits if coder == "alias" else None)
        r64 = ref.cpu_baseline("rans64", data, threads, runs=max(1, runs - 1), scale_bits=14)
        if not (r["ok"] and r64["ok"]):
            raise RuntimeError("reference CPU round trip failed")
        kind_s = "reference"
        paths = {
            ("main_alias.cpp 2-way alias" if coder == "alias" else "main_simd.cpp 8-way scalar enc + SSE4.1 dec"): r,
            "main64.cpp rans64 2-way": r64,
        }
    else:   # the C port (oracle/rans_oracle.c), single thread
        orc = oracle.Oracle()
        cid = oracle.CODER_ALIAS if coder == "alias" else oracle.CODER_WORD
        freqs, cum = orc.model(data, scale_bits)
        nl = 2 if coder == "alias" else 8
        t0 = time.perf_counter()
        stream = orc.encode(cid, data, freqs, cum, nl, scale_bits)
        t1 = time.perf_counter()
        dec, _ = orc.decode(cid, stream, data.size, freqs, cum, nl, scale_bits)
        t2 = time.perf_counter()
        assert np.array_equal(dec, data)
        kind_s, threads = "port", 1
        paths = {"oracle/rans_oracle.c port": {"enc_s": t1 - t0, "dec_s": t2 - t1, "bytes": stream.size, "ok": True}}
    best_name, best = min(paths.items(), key=lambda kv: kv[1]["enc_s"] + kv[1]["dec_s"])
    rt = nbytes / (best["enc_s"] + best["dec_s"]) / 1e9
    detail = {k: {"encode_gsym_s": round(nbytes / v["enc_s"] / 1e9, 4), "decode_gsym_s": round(nbytes / v["dec_s"] / 1e9, 4),
                  "compressed_bytes": v["bytes"]} for k, v in paths.items()}
    if kind_s == "reference" and threads > 1:      # single-thread figures beside the all-core ones
        small = data[:min(nbytes, 32 << 20)]
        st = {("main_alias.cpp 2-way alias" if coder == "alias" else "main_simd.cpp 8-way scalar enc + SSE4.1 dec"):
              ref.cpu_baseline("alias" if coder == "alias" else "simd", small, 1, runs=2, scale_bits=scale_bits if coder == "alias" else None),
              "main64.cpp rans64 2-way": ref.cpu_baseline("rans64", small, 1, runs=2, scale_bits=14)}
        for k, v in st.items():
            detail[k]["single_thread_encode_gsym_s"] = round(small.size / v["enc_s"] / 1e9, 4)
            detail[k]["single_thread_decode_gsym_s"] = round(small.size / v["dec_s"] / 1e9, 4)
    return {"value": rt, "unit": UNIT, "cores": threads, "kind": kind_s,
            "sample": f"{nbytes >> 20} MiB of the {kind} workload, one contiguous slice per thread, best of {runs}; "
                      f"round trip = encode + decode; fastest path: {best_name}",
            "paths": detail}


def run_reference(args, rank, world):
    if rank != 0:
        return
    coder, sb, kind = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    sample = min(args.cpu_sample, args.size)
    t0 = time.perf_counter()
    base = None
    for _ in range(max(1, min(args.warmup, 1))):
        base = cpu_reference_run(kind, coder, sb, sample, runs=1, threads=threads)
    vals = []
    for _ in range(max(1, min(args.steps, 3))):
        base = cpu_reference_run(kind, coder, sb, sample, runs=1, threads=threads)
        vals.append(base["value"])
    v = float(np.median(vals))
    base["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sample / (v * 1e9), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": args.workload, "symbols_per_step": sample, "note": "CPU, host cores only; bounded sample"},
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(time.perf_counter() - t0, 2),
    }))


# ------------------------------------------------------------------ our arm

def run_ours(args, rank, local_rank, world):
    import torch
    import ryg_rans_b200 as rb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the rANS hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    coder_name, sb, kind = WORKLOADS[args.workload]
    coder = rb.CODER_ALIAS if coder_name == "alias" else rb.CODER_WORD
    n, chunk = args.size, args.chunk
    if coder_name == "blocks" and BLOCK_SIZE % chunk:
        raise SystemExit("--chunk must divide the 64 KiB block size")
    stream = torch.cuda.current_stream()
    ctx = rb.Context(local_rank, stream.cuda_stream)

    blocks = coder_name == "blocks"
    if blocks:
        n = (n // BLOCK_SIZE) * BLOCK_SIZE if n >= BLOCK_SIZE else BLOCK_SIZE
        if args.size == 1 << 30:
            n = 8192 * BLOCK_SIZE                      # 64 Ki blocks over 8 GPUs = 8192 blocks (512 MiB) per GPU
    data = synth_torch(kind, n, seed=rank, device=dev)
    n_chunks = ctx.chunk_count(n, chunk)
    cap = ctx.encode_bound(n, chunk)
    blob = torch.empty(cap, dtype=torch.uint8, device=dev)
    offsets = torch.zeros(n_chunks + 1, dtype=torch.int64, device=dev)
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    model_bytes = 0
    if blocks:
        n_blocks = n // BLOCK_SIZE
        bfreqs = torch.zeros(n_blocks * 256, dtype=torch.int16, device=dev)
        model_bytes = 512 * n_blocks                  # the per-block frequency tables travel with the blob

        def enc():                                    # per-block models are part of encoding a block
            ctx.blocks_build_models_device(data.data_ptr(), n_blocks, BLOCK_SIZE, bfreqs.data_ptr())
            ctx.blocks_encode_device(data.data_ptr(), n_blocks, BLOCK_SIZE, bfreqs.data_ptr(), chunk, blob.data_ptr(), cap,
                                     offsets.data_ptr())

        def dec(blob_size):
            ctx.blocks_decode_device(blob.data_ptr(), blob_size, offsets.data_ptr(), bfreqs.data_ptr(), n_blocks, BLOCK_SIZE, chunk,
                                     out.data_ptr())
    else:
        # model: device histogram -> host normalisation (reference order-dependent code stays on the host)
        counts = ctx.histogram_device(data.data_ptr(), n)
        st = rb.SymbolStats()
        st.freqs[:] = counts.astype(np.uint32)
        st.normalize_freqs(1 << sb)
        model = ctx.model(coder, sb, st.freqs)

        def enc():
            ctx.encode_device(model, data.data_ptr(), n, chunk, blob.data_ptr(), cap, offsets.data_ptr())

        def dec(blob_size):
            ctx.decode_device(model, blob.data_ptr(), blob_size, offsets.data_ptr(), chunk, out.data_ptr(), n)

    # warm-up + bit-exact verification (outside the timed region)
    enc()
    ctx.sync()
    blob_size = int(offsets[-1].item())
    for _ in range(max(args.warmup, 3)):
        enc()
        dec(blob_size)
    ctx.sync()
    if not torch.equal(out, data):
        raise SystemExit("bench.py: round trip is NOT bit-exact")
    out.zero_()

    launches0 = ctx.launches
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(args.steps):
        ev[k][0].record()
        enc()
        ev[k][1].record()
        dec(blob_size)
        ev[k][2].record()
    t_end.record()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    clocks = sampler.stop() if sampler else None
    launches = ctx.launches - launches0
    ctx.sync()
    if not torch.equal(out, data):
        raise SystemExit("bench.py: timed round trip is NOT bit-exact")

    total_ms = t_start.elapsed_time(t_end)
    enc_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    dec_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    tt = torch.tensor([total_ms, enc_ms, dec_ms], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, enc_ms_max, dec_ms_max = tt.tolist()

    # ---- NCCL gather of the compressed blobs + directories, timed separately
    gather_ms = None
    if dist:
        from ryg_rans_b200.shard import gather_blobs
        g0 = torch.cuda.Event(enable_timing=True); g1 = torch.cuda.Event(enable_timing=True)
        gblob, gdir = gather_blobs(blob[:blob_size], offsets, dst=0)       # warm-up: NCCL channel set-up, allocations
        del gblob, gdir
        dist.barrier(); torch.cuda.synchronize()
        g0.record()
        gblob, gdir = gather_blobs(blob[:blob_size], offsets, dst=0)
        g1.record(); torch.cuda.synchronize()
        gt = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(gt, op=dist.ReduceOp.MAX)
        gather_ms = gt.item()
        if rank == 0:
            assert gblob.numel() % 16 == 0 and int(gdir[-1]) == gblob.numel()
        del gblob, gdir

    # ---- e2e: host pointers through the C-ABI, copies inside the timed region
    e2e = None
    if args.e2e_steps > 0 and not blocks:
        # first-touch the pinned buffers on the GPU's own NUMA node (what `numactl --cpunodebind` would do for a
        # caller); with 8 ranks the copies otherwise cross the socket interconnect
        numa_node, saved_affinity = bind_to_gpu_numa_node(local_rank)
        h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_in.copy_(data)
        h_blob = torch.empty(cap, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_off = np.zeros(n_chunks + 1, np.uint64)
        import ctypes as C
        lib = ctx.lib
        size = C.c_size_t(0)

        def e2e_step():
            lib.check(lib.dll.rb200_encode(ctx.h, model.h, h_in.data_ptr(), n, chunk, h_blob.data_ptr(), cap, h_off.ctypes.data,
                                           C.byref(size), rb.MEM_HOST), ctx.h)
            lib.check(lib.dll.rb200_decode(ctx.h, model.h, h_blob.data_ptr(), size.value, h_off.ctypes.data, chunk,
                                           h_out.data_ptr(), n, rb.MEM_HOST), ctx.h)
        e2e_step()     # warm-up (allocates staging)
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if not torch.equal(h_out, h_in):
            raise SystemExit("bench.py: e2e round trip is NOT bit-exact")
        et = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist:
            dist.all_reduce(et, op=dist.ReduceOp.MAX)
        if saved_affinity is not None:
            os.sched_setaffinity(0, saved_affinity)
        e2e = {"value": world * n * args.e2e_steps / et.item() / 1e9, "unit": UNIT, "host_numa_node": numa_node,
               "h2d_bytes_per_step": int(n + size.value + 8 * (n_chunks + 1)), "d2h_bytes_per_step": int(size.value + n + 8 * (n_chunks + 1)),
               "steps": args.e2e_steps, "note": "rb200_encode + rb200_decode with RB200_MEM_HOST on pinned buffers, wall clock",
               "bound": "PCIe: each call streams its input in and its output back concurrently; with both directions busy the "
                        "slower one gets 43-47 GB/s on this box (tools/pcie_probe.py, profiles/r1_pcie_copies.log)"}

    if rank != 0:
        if dist:
            dist.barrier()
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    algo_bytes = n + blob_size + model_bytes         # (1 + c) bytes per symbol (+ 512 B per block model)
    dec_gbs = algo_bytes / (dec_ms_max * 1e-3) / 1e9
    enc_gbs = algo_bytes / (enc_ms_max * 1e-3) / 1e9
    value = world * n * args.steps / (total_ms * 1e-3) / 1e9
    per_step = launches // max(args.steps, 1)          # launches of one encode call + one decode call
    if blocks:
        enc_kernels, enc_step = "block_model + block_encode + directory_scan + compact", "per-block models (1 launch) + encode (3 launches)"
    elif per_step == 2:
        enc_kernels, enc_step = "fused encode (encode + directory scan + placement in one persistent launch)", "encode (1 fused launch)"
    else:
        enc_kernels, enc_step = "encode + directory_scan + compact", "encode (%d launches)" % (per_step - 1)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": args.workload, "symbols_per_gpu": n, "chunk_syms": chunk, "lanes": 32, "coder": coder_name,
                   "scale_bits": sb, "compressed_bytes_per_symbol": blob_size / n,
                   "l2": "inputs (1 GiB symbols + ~1 GiB blob) exceed the 126 MB L2; no flush needed",
                   "step": enc_step + " + decode (1 launch), device-resident"},
        "decode_gsym_s": world * n / (dec_ms_max * 1e-3) / 1e9, "encode_gsym_s": world * n / (enc_ms_max * 1e-3) / 1e9,
        "decode_ms": dec_ms_max, "encode_ms": enc_ms_max,
        "roofline": {"kernel": {"word": "word_decode_kernel", "alias": "alias_decode_kernel", "blocks": "block_decode_kernel"}[coder_name],
                     "bound": "hbm",
                     "achieved": dec_gbs, "peak": peak, "unit": "GB/s", "frac": dec_gbs / peak,
                     "traffic": ncu_traffic(args.workload, chunk, "word_decode_kernel") if n == 1 << 30 else None,
                     "algorithmic_bytes_per_launch": algo_bytes, "peak_source": peak_src},
        "roofline_encode_call": {"kernels": enc_kernels, "bound": "hbm", "achieved": enc_gbs, "peak": peak,
                                 "unit": "GB/s", "frac": enc_gbs / peak, "algorithmic_bytes_per_call": algo_bytes},
        "gpu_launches": launches, "clocks": clocks, "e2e": e2e,
    }
    if gather_ms is not None:
        line["nccl_blob_gather_ms"] = gather_ms
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_reference_run("zipf" if blocks else kind, "word" if blocks else coder_name, sb,
                                                     min(args.cpu_sample, n), runs=2, threads=os.cpu_count() or 1)
        except Exception as e:  # the baseline must not take the GPU number down with it
            line["cpu_baseline"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def bind_to_gpu_numa_node(dev_index):
    """Restrict this process to the CPUs of the NUMA node GPU `dev_index` hangs off.  Returns (node, previous
    affinity) or (None, None) when the topology cannot be read; the caller restores the affinity."""
    try:
        import torch
        props = torch.cuda.get_device_properties(dev_index)
        if hasattr(props, "pci_bus_id"):
            bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        else:
            bdf = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(dev_index)],
                                 capture_output=True, text=True, timeout=20).stdout.strip().lower()[-12:]
        path = "/sys/bus/pci/devices/%s/numa_node" % bdf
        node = int(open(path).read())
        if node < 0:
            return None, None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus |= set(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        saved = os.sched_getaffinity(0)
        cpus &= saved
        if not cpus:
            return None, None
        os.sched_setaffinity(0, cpus)
        return node, saved
    except Exception:  # noqa: BLE001 -- topology files are optional
        return None, None


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
