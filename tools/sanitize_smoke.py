"""A small tour of every kernel for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Sizes are tiny (the tools slow kernels down ~100x) but cross every ring wrap, ragged tail, multi-slice host pipeline
(RB200_SLICE_MIB=1), persistent-grid loop and the corrupt-stream path.  Results are checked against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("RB200_SLICE_MIB", "1")
import oracle  # noqa: E402
import ryg_rans_b200 as rb  # noqa: E402


def main():
    rng = np.random.default_rng(5)
    p = 1.0 / np.arange(1, 257) ** 1.1
    orc = oracle.Oracle()
    ctx = rb.Context(0)
    for n in (1, 33, 40_000, 1_300_003):
        for kind in ("zipf", "uniform"):
            data = rng.choice(256, n, p=p / p.sum()).astype(np.uint8) if kind == "zipf" else rng.integers(0, 256, n, dtype=np.uint8)
            for coder, ocoder, sb in ((rb.CODER_WORD, oracle.CODER_WORD, 12), (rb.CODER_ALIAS, oracle.CODER_ALIAS, 16),
                                      (rb.CODER_ALIAS, oracle.CODER_ALIAS, 11), (rb.CODER_BYTE, oracle.CODER_BYTE, 14),
                                      (rb.CODER_RANS64, oracle.CODER_RANS64, 14)):
                f, c = orc.model(data, sb)
                m = ctx.model(coder, sb, f)
                for chunk in ((32, 4096) if n < 100_000 else (4096, 65536)):
                    blob, offs = ctx.encode_host(m, data, chunk)
                    ob, oo = orc.chunked_encode(ocoder, data, f, c, chunk, scale_bits=sb)
                    assert np.array_equal(offs, oo) and np.array_equal(blob, ob), (n, kind, coder, sb, chunk)
                    assert np.array_equal(ctx.decode_host(m, blob, offs, n, chunk), data)
                    if n == 40_000 and chunk == 4096:
                        bad = blob.copy()
                        bad[int(offs[1]) + 130:int(offs[1]) + 300] ^= 0x33
                        try:
                            ctx.decode_host(m, bad, offs, n, chunk)
                            raise SystemExit("corruption not reported")
                        except rb.RansError:
                            pass
                        assert np.array_equal(ctx.decode_host(m, blob, offs, n, chunk), data)
                m.close()
    # per-block path: fused model+encode, given-model encode, decode
    nb, bs, chunk = 23, 65536, 8192
    blocks = [rng.permutation(256).astype(np.uint8)[rng.choice(256, bs, p=(q := 1.0 / np.arange(1, 257) ** (0.7 + 0.1 * (b % 7))) / q.sum())] for b in range(nb)]
    data = np.concatenate(blocks)
    blob, offs, freqs = ctx.blocks_model_encode_host(data, nb, bs, chunk)
    assert np.array_equal(freqs, ctx.blocks_build_models(data, nb, bs))
    b2, o2 = ctx.blocks_encode_host(data, nb, bs, freqs, chunk)
    assert np.array_equal(b2, blob) and np.array_equal(o2, offs)
    assert np.array_equal(ctx.blocks_decode_host(blob, offs, freqs, nb, bs, chunk), data)
    st = rb.SymbolStats().count_freqs(data)
    assert np.array_equal(ctx.histogram(data), st.freqs.astype(np.uint64))
    print("sanitize smoke ok:", ctx.launches, "launches")


if __name__ == "__main__":
    main()
