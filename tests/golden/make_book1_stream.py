#!/usr/bin/env python
"""Golden vector: the REFERENCE's own 32-way word-coder stream of its test file book1 (435 702 bytes).

Made with the reference's primitives (oracle/_ref: RansWordEncPut / RansWordEncFlush driven as main_simd.cpp:287-300
does, with N = 32 and the model of normalize_freqs(4096)).  The GPU box has no /root/reference, so the stream and the
256 frequencies are committed; the GPU test decodes it (the bytes must hash to book1's SHA-256) and re-encodes the
result as one chunk (must reproduce this stream byte for byte).  Run where /root/reference exists:

    python tests/golden/make_book1_stream.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402


def main():
    book = np.fromfile("/root/reference/book1", dtype=np.uint8)
    ref = oracle.Reference()
    freqs, cum = ref.model(book, 12)
    stream = ref.encode(oracle.CODER_WORD, book, freqs, cum, 32)
    dec, used = ref.decode(oracle.CODER_WORD, stream, book.size, freqs, cum, 32)
    assert np.array_equal(dec, book) and used == stream.size
    assert stream.size == 435702, stream.size                 # the size SURVEY 8(a) measured for N = 32
    np.savez(os.path.join(HERE, "book1_n32.npz"), stream=stream, freqs=freqs.astype(np.uint32), n=np.int64(book.size),
             book_sha256=np.frombuffer(hashlib.sha256(book.tobytes()).digest(), np.uint8),
             stream_sha256=np.frombuffer(hashlib.sha256(stream.tobytes()).digest(), np.uint8))
    print("book1", book.size, "->", stream.size, "bytes;", hashlib.sha256(stream.tobytes()).hexdigest())


if __name__ == "__main__":
    main()
