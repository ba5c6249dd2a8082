// rfx_newline.cuh -- the per-chunk masks of the K1 newline scan (rfx_fastq.cu).  A header of its own so that tests/hostemu can run
// the same arithmetic on the host against a byte-by-byte scan (no CUDA built-ins in here).
#pragma once
#include <stdint.h>
#ifndef RFX_NEWLINE_HOSTEMU
#include <cuda_runtime.h>
#define RFX_NL_FN __device__ __forceinline__
#else
#define RFX_NL_FN static inline
#endif

namespace rfx {

// ------------------------------------------------------------------------------------------
// newline scan: element = one 64-byte aligned chunk of the text
// ------------------------------------------------------------------------------------------
struct TextView {
    const uint8_t* aligned;  // text pointer rounded down to 64 bytes
    uint32_t delta;          // text - aligned
    uint64_t len;
};

// 0x80 in every byte of x that equals c (exact: no carry crosses a byte), then one bit per byte
RFX_NL_FN uint32_t bytes_equal(uint32_t x, uint32_t c4) {
    x ^= c4;
    return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x | 0x7f7f7f7fu);
}
RFX_NL_FN uint32_t movemask4(uint32_t hi) { return (((hi >> 7) * 0x00204081u) >> 21) & 0xfu; }
RFX_NL_FN uint32_t mask16(const uint4& v, uint32_t c4) {
    return movemask4(bytes_equal(v.x, c4)) | (movemask4(bytes_equal(v.y, c4)) << 4) | (movemask4(bytes_equal(v.z, c4)) << 8) |
           (movemask4(bytes_equal(v.w, c4)) << 12);
}
// bytes of the chunk that belong to the text [0, len)
RFX_NL_FN uint64_t chunk_valid(const TextView& tv, uint64_t chunk) {
    const int64_t p0 = (int64_t)(chunk * 64) - (int64_t)tv.delta;  // text position of byte 0 of the chunk
    uint64_t valid = ~0ull;
    if (p0 < 0) valid &= ~0ull << (uint32_t)(-p0);
    const int64_t over = p0 + 64 - (int64_t)tv.len;
    if (over > 0) valid &= over >= 64 ? 0ull : (~0ull >> (uint32_t)over);
    return valid;
}

// newline positions of one 64-byte aligned chunk as a 64-bit mask
RFX_NL_FN uint64_t newline_mask64(const TextView& tv, uint64_t chunk) {
    const uint4* p = reinterpret_cast<const uint4*>(tv.aligned + chunk * 64);
    const uint4 v0 = p[0], v1 = p[1], v2 = p[2], v3 = p[3];
    const uint64_t mask = (uint64_t)mask16(v0, 0x0a0a0a0au) | ((uint64_t)mask16(v1, 0x0a0a0a0au) << 16) | ((uint64_t)mask16(v2, 0x0a0a0a0au) << 32) |
                          ((uint64_t)mask16(v3, 0x0a0a0a0au) << 48);
    return mask & chunk_valid(tv, chunk);
}

RFX_NL_FN int lowest_bit(uint64_t m) {
#ifndef RFX_NEWLINE_HOSTEMU
    return __ffsll((long long)m) - 1;
#else
    return __builtin_ctzll(m);
#endif
}

// The same chunk as two masks: the newlines, and those of them behind which an '@' follows (the first byte of the next line).
// Pass 1 of the newline scan stores both (16 bytes per 64 bytes of text), so that pass 2 never reads the text again.  A chunk
// holds less than one newline on average: the byte behind each is looked at on its own (an L1 hit, the chunk was just loaded;
// for bit 63 it is byte 0 of the next chunk) instead of building a second 64-bit mask from the whole chunk.
RFX_NL_FN void chunk_masks(const TextView& tv, uint64_t chunk, uint64_t& nl_out, uint64_t& nl_at_out) {
    const uint64_t nl = newline_mask64(tv, chunk);
    const int64_t p0 = (int64_t)(chunk * 64) - (int64_t)tv.delta;  // text position of byte 0 of the chunk
    const uint8_t* bytes = tv.aligned + chunk * 64;
    uint64_t nl_at = 0;
    for (uint64_t m = nl; m; m &= m - 1) {
        const int j = lowest_bit(m);
        if (p0 + j + 1 < (int64_t)tv.len && bytes[j + 1] == '@') nl_at |= 1ull << j;
    }
    nl_out = nl;
    nl_at_out = nl_at;
}

}  // namespace rfx
