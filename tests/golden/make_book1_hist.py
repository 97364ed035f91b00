#!/usr/bin/env python
"""Byte histogram of the reference's test file (book1), for the text-like synthetic workload.

SURVEY 8(d) C4 asks for i.i.d. draws from book1's order-0 distribution; the GPU box has no
/root/reference, so the 256 counts are committed as a fixture.  Run where /root/reference exists:

    python tests/golden/make_book1_hist.py
"""
import hashlib
import json
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    raw = open("/root/reference/book1", "rb").read()
    counts = np.bincount(np.frombuffer(raw, np.uint8), minlength=256)
    p = counts[counts > 0] / counts.sum()
    out = {"source": "book1 (Calgary corpus), as shipped with rygorous/ryg_rans", "bytes": len(raw),
           "sha256": hashlib.sha256(raw).hexdigest(), "distinct_symbols": int((counts > 0).sum()),
           "entropy_bits_per_symbol": round(float(-(p * np.log2(p)).sum()), 4), "counts": [int(c) for c in counts]}
    with open(os.path.join(HERE, "book1_hist.json"), "w") as f:
        json.dump(out, f)
        f.write("\n")
    print({k: v for k, v in out.items() if k != "counts"})


if __name__ == "__main__":
    main()
