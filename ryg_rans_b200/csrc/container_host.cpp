// container_host.cpp -- self-describing wire format around (model, directory, blob)  [SURVEY 8f.3].
//
// The reference defines no container: its drivers keep a pointer and the sizes in local variables
// (main.cpp:182-188).  The bulk API's natural output (blob + offsets + the caller's freqs) is enough
// for in-process use; this file adds the versioned, checksummed envelope needed to ship it:
//
//   [ 64-byte header ][ freqs: 256 x u32 ][ offsets: (n_chunks + 1) x u64 ][ pad to 16 ][ blob ]
//
// Host-only, no GPU work.  All integers little-endian.  CRC-32 (IEEE, reflected) over header
// (with the crc fields zeroed) + freqs + offsets always; over the blob when RB200_CONTAINER_CRC_BLOB
// is set (1 GiB takes ~1 s on one core, so it is optional).
#include "rans_b200.h"

#include <cstring>

namespace {

constexpr uint32_t kMagic = 0x43324252u;   // "RB2C"
constexpr uint16_t kVersion = 1;

struct Header {                 // 64 bytes
    uint32_t magic;
    uint16_t version;
    uint8_t coder;
    uint8_t scale_bits;
    uint32_t lanes;
    uint32_t chunk_syms;
    uint64_t n_symbols;
    uint64_t n_chunks;
    uint64_t blob_bytes;
    uint32_t flags;
    uint32_t meta_crc;          // header (crc fields zero) + freqs + offsets
    uint32_t blob_crc;          // valid when flags & RB200_CONTAINER_CRC_BLOB
    uint32_t reserved[3];
};
static_assert(sizeof(Header) == 64, "container header must stay 64 bytes");

struct CrcTable {
    uint32_t t[8][256];
    CrcTable()
    {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
            t[0][i] = c;
        }
        for (uint32_t i = 0; i < 256; i++)
            for (int k = 1; k < 8; k++) t[k][i] = (t[k - 1][i] >> 8) ^ t[0][t[k - 1][i] & 0xff];
    }
};

uint32_t crc32_update(uint32_t crc, const void* data, size_t len)     // slicing-by-8
{
    static const CrcTable table;                                      // built once, thread-safe (C++11 static)
    const uint32_t (*T)[256] = table.t;
    const uint8_t* p = static_cast<const uint8_t*>(data);
    crc = ~crc;
    while (len >= 8) {
        uint32_t a, b;
        std::memcpy(&a, p, 4);
        std::memcpy(&b, p + 4, 4);
        a ^= crc;
        crc = T[7][a & 0xff] ^ T[6][(a >> 8) & 0xff] ^ T[5][(a >> 16) & 0xff] ^ T[4][a >> 24]
            ^ T[3][b & 0xff] ^ T[2][(b >> 8) & 0xff] ^ T[1][(b >> 16) & 0xff] ^ T[0][b >> 24];
        p += 8;
        len -= 8;
    }
    while (len--) crc = (crc >> 8) ^ T[0][(crc ^ *p++) & 0xff];
    return ~crc;
}

size_t pad16(size_t v) { return (v + 15) & ~static_cast<size_t>(15); }
size_t blob_offset(uint64_t n_chunks) { return pad16(sizeof(Header) + 256 * sizeof(uint32_t) + (n_chunks + 1) * sizeof(uint64_t)); }

uint32_t meta_crc(const Header& h, const uint32_t* freqs, const uint64_t* offsets)
{
    Header z = h;
    z.meta_crc = 0;
    z.blob_crc = 0;
    uint32_t c = crc32_update(0, &z, sizeof z);
    c = crc32_update(c, freqs, 256 * sizeof(uint32_t));
    return crc32_update(c, offsets, (h.n_chunks + 1) * sizeof(uint64_t));
}

// what rb200_model_create would accept: a known coder, its scale_bits range, frequencies that sum to 1 << scale_bits
bool model_fields_ok(uint32_t coder, uint32_t scale_bits, const uint32_t* freqs)
{
    if (coder > RB200_CODER_RANS64 || scale_bits < 8 || scale_bits > 16) return false;
    if (coder == RB200_CODER_WORD && scale_bits != 12) return false;
    uint64_t sum = 0;
    for (int s = 0; s < 256; s++) sum += freqs[s];
    return sum == (1ull << scale_bits);
}

}  // namespace

extern "C" size_t rb200_container_size(size_t n_chunks, size_t blob_bytes) { return blob_offset(n_chunks) + blob_bytes; }

extern "C" int rb200_container_pack(int coder, uint32_t scale_bits, uint32_t chunk_syms, size_t n, const uint32_t freqs[256],
                                    const uint64_t* offsets, const uint8_t* blob, size_t blob_bytes, uint32_t flags, uint8_t* out,
                                    size_t out_cap, size_t* out_size)
{
    if (!freqs || !offsets || (!blob && blob_bytes) || !out || !chunk_syms) return RB200_E_ARG;
    if (coder < 0 || !model_fields_ok(static_cast<uint32_t>(coder), scale_bits, freqs)) return RB200_E_MODEL;   // never write what open rejects
    const size_t n_chunks = rb200_chunk_count(n, chunk_syms);
    if (offsets[n_chunks] != blob_bytes || (blob_bytes & 15)) return RB200_E_ARG;
    const size_t total = rb200_container_size(n_chunks, blob_bytes);
    if (total > out_cap) return RB200_E_SPACE;
    Header h{};
    h.magic = kMagic;
    h.version = kVersion;
    h.coder = static_cast<uint8_t>(coder);
    h.scale_bits = static_cast<uint8_t>(scale_bits);
    h.lanes = RB200_LANES;
    h.chunk_syms = chunk_syms;
    h.n_symbols = n;
    h.n_chunks = n_chunks;
    h.blob_bytes = blob_bytes;
    h.flags = flags & RB200_CONTAINER_CRC_BLOB;
    h.meta_crc = meta_crc(h, freqs, offsets);
    if (h.flags & RB200_CONTAINER_CRC_BLOB) h.blob_crc = crc32_update(0, blob, blob_bytes);
    uint8_t* p = out;
    std::memcpy(p, &h, sizeof h);
    p += sizeof h;
    std::memcpy(p, freqs, 256 * sizeof(uint32_t));
    p += 256 * sizeof(uint32_t);
    std::memcpy(p, offsets, (n_chunks + 1) * sizeof(uint64_t));
    p += (n_chunks + 1) * sizeof(uint64_t);
    const size_t boff = blob_offset(n_chunks);
    std::memset(p, 0, out + boff - p);
    if (blob_bytes) std::memcpy(out + boff, blob, blob_bytes);
    if (out_size) *out_size = total;
    return RB200_OK;
}

extern "C" int rb200_container_open(const uint8_t* buf, size_t size, rb200_container_info* info)
{
    if (!buf || !info) return RB200_E_ARG;
    if (size < sizeof(Header)) return RB200_E_STREAM;
    Header h;
    std::memcpy(&h, buf, sizeof h);
    if (h.magic != kMagic || h.version != kVersion || h.lanes != RB200_LANES || !h.chunk_syms) return RB200_E_STREAM;
    if (h.n_chunks != h.n_symbols / h.chunk_syms + (h.n_symbols % h.chunk_syms != 0) || h.n_chunks >= (1ull << 31) ||
        (h.blob_bytes & 15) || h.coder > RB200_CODER_RANS64 || h.scale_bits < 8 || h.scale_bits > 16)
        return RB200_E_STREAM;
    const size_t boff = blob_offset(h.n_chunks);
    if (boff > size || h.blob_bytes > size - boff) return RB200_E_STREAM;
    // freqs/offsets sit at 4- and 8-byte aligned positions when buf itself is 8-byte aligned (callers map or malloc it)
    const uint32_t* freqs = reinterpret_cast<const uint32_t*>(buf + sizeof(Header));
    const uint64_t* offsets = reinterpret_cast<const uint64_t*>(buf + sizeof(Header) + 256 * sizeof(uint32_t));
    if (reinterpret_cast<uintptr_t>(buf) & 7) return RB200_E_ARG;
    if (meta_crc(h, freqs, offsets) != h.meta_crc) return RB200_E_STREAM;
    if (!model_fields_ok(h.coder, h.scale_bits, freqs)) return RB200_E_STREAM;        // CRC-valid but inconsistent
    if (offsets[h.n_chunks] != h.blob_bytes) return RB200_E_STREAM;
    if (h.n_chunks && offsets[0] >= 16) return RB200_E_STREAM;                        // the first stream starts inside the first vector
    for (uint64_t c = 0; c < h.n_chunks; c++)
        if (offsets[c] > offsets[c + 1]) return RB200_E_STREAM;
    if ((h.flags & RB200_CONTAINER_CRC_BLOB) && crc32_update(0, buf + boff, h.blob_bytes) != h.blob_crc) return RB200_E_STREAM;
    info->coder = h.coder;
    info->scale_bits = h.scale_bits;
    info->chunk_syms = h.chunk_syms;
    info->flags = h.flags;
    info->n_symbols = h.n_symbols;
    info->n_chunks = h.n_chunks;
    info->blob_bytes = h.blob_bytes;
    info->freqs = freqs;
    info->offsets = offsets;
    info->blob = buf + boff;
    return RB200_OK;
}
