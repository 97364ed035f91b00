// Test shim around include/rans_word_avx2.h: decodes chunk streams / whole containers of the word coder on the CPU.
#include "rans_word_avx2.h"
extern "C" int avx2_decode_chunk(const uint8_t* stream, size_t bytes, const uint32_t* freqs, const uint32_t* cum, uint8_t* out, size_t m, int via_word_tables)
{
    static RansWord32Tables t32;
    if (via_word_tables) {
        static RansWordTables t;
        memset(&t, 0, sizeof t);
        for (int s = 0; s < 256; s++) if (freqs[s]) RansWordTablesInitSymbol(&t, (uint8_t)s, cum[s], freqs[s]);
        RansWord32TablesInit(&t32, &t);
    } else {
        RansWord32TablesReset(&t32);
        for (int s = 0; s < 256; s++) if (freqs[s]) RansWord32TablesInitSymbol(&t32, (uint8_t)s, cum[s], freqs[s]);
    }
    return RansWord32DecodeChunk(stream, bytes, &t32, out, m);
}

extern "C" int avx2_decode_container(const uint8_t* blob, const uint64_t* offs, size_t n_chunks, size_t chunk, size_t n,
                                     const uint32_t* freqs, const uint32_t* cum, uint8_t* out)
{
    static RansWord32Tables t32;
    RansWord32TablesReset(&t32);
    for (int s = 0; s < 256; s++) if (freqs[s]) RansWord32TablesInitSymbol(&t32, (uint8_t)s, cum[s], freqs[s]);
    for (size_t c = 0; c < n_chunks; c++) {
        const size_t lo = offs[c], end = offs[c + 1] & ~(uint64_t)15;
        const size_t m = (c + 1) * chunk <= n ? chunk : n - c * chunk;
        if (RansWord32DecodeChunk(blob + lo, end - lo, &t32, out + c * chunk, m)) return -1 - (int)c;
    }
    return 0;
}

extern "C" long decode_chunks(const uint8_t* blob, const uint64_t* offs, size_t n_chunks, uint32_t chunk, size_t n,
                              const uint32_t* freqs, const uint32_t* cum, uint8_t* out)
{
    static RansWord32Tables t32;
    RansWord32TablesReset(&t32);
    for (int s = 0; s < 256; s++) if (freqs[s]) RansWord32TablesInitSymbol(&t32, (uint8_t)s, cum[s], freqs[s]);
    return RansWord32DecodeChunks(blob, offs, 0, n_chunks, chunk, n, &t32, out);
}
