#!/usr/bin/env python
"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Runs only where /root/reference exists (it drives oracle/_ref/libryg_ref.so, which is
the reference's own headers and driver code compiled in place).  The fixtures are small
seeded inputs, the reference's normalised model for them, and the reference's N-way
streams; the GPU box (which has no /root/reference) checks the oracle and the CUDA path
against these files.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from conftest import _gen  # noqa: E402

CASES = [
    # name, kind, n, seed
    ("zipf_3001", "zipf", 3001, 11),
    ("text_4096", "text", 4096, 12),
    ("uniform_2500", "uniform", 2500, 13),
    ("two_1000", "two", 1000, 14),
    ("const_777", "const", 777, 15),
    ("skew_5000", "skew", 5000, 16),
]
CODERS = [("word", oracle.CODER_WORD, 12, (1, 2, 8, 32)), ("byte", oracle.CODER_BYTE, 14, (1, 2, 32)),
          ("alias", oracle.CODER_ALIAS, 16, (1, 2, 32)), ("rans64", oracle.CODER_RANS64, 14, (1, 2, 32))]


def main():
    ref = oracle.Reference()
    arrays, index = {}, {}
    for name, kind, n, seed in CASES:
        data = _gen(kind, n, seed)
        arrays[f"{name}/data"] = data
        index[name] = {"kind": kind, "n": n, "seed": seed, "sha256": hashlib.sha256(data.tobytes()).hexdigest(), "streams": {}}
        for cname, cid, sb, lanes in CODERS:
            freqs, cum = ref.model(data, sb)
            arrays[f"{name}/{cname}/freqs"] = freqs
            for nl in lanes:
                s = ref.encode(cid, data, freqs, cum, nl, sb)
                dec, used = ref.decode(cid, s, n, freqs, cum, nl, sb)
                assert np.array_equal(dec, data) and used == s.size
                arrays[f"{name}/{cname}/N{nl}"] = s
                index[name]["streams"][f"{cname}/N{nl}"] = {"bytes": int(s.size), "sha256": hashlib.sha256(s.tobytes()).hexdigest()}
        # alias tables for the 16-bit model
        freqs, cum = ref.model(data, 16)
        for k, v in zip(("divider", "slot_adjust", "slot_freqs", "sym_id", "remap"), ref.alias_build(freqs, cum)):
            if k == "remap":     # 256 KiB each: keep only the hash
                index[name]["alias_remap_sha256"] = hashlib.sha256(v.tobytes()).hexdigest()
            else:
                arrays[f"{name}/alias_tables/{k}"] = v
    # book1 known answers (sizes are README:48,62,82,96,110; hashes pin the full streams)
    book = np.fromfile("/root/reference/book1", dtype=np.uint8)
    bk = {"n": int(book.size), "sha256": hashlib.sha256(book.tobytes()).hexdigest(), "streams": {}}
    for cname, cid, sb, lanes in CODERS:
        freqs, cum = ref.model(book, sb)
        for nl in lanes:
            s = ref.encode(cid, book, freqs, cum, nl, sb)
            bk["streams"][f"{cname}/N{nl}"] = {"bytes": int(s.size), "sha256": hashlib.sha256(s.tobytes()).hexdigest()}
    index["book1"] = bk
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **arrays)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(index, f, indent=1, sort_keys=True)
    print("wrote", len(arrays), "arrays;", os.path.getsize(os.path.join(HERE, "golden.npz")), "bytes")


if __name__ == "__main__":
    main()
