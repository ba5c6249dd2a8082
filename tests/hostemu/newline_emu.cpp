// newline_emu.cpp -- TEST-ONLY: the per-chunk newline / newline-then-'@' masks of reflexiv_b200/csrc/rfx_newline.cuh (K1) run on the
// host against a byte-by-byte scan, on random texts at every alignment with garbage around them.  Not part of libreflexiv_cuda.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#define RFX_NEWLINE_HOSTEMU 1
struct uint4 { uint32_t x, y, z, w; };
#include "../../reflexiv_b200/csrc/rfx_newline.cuh"

using namespace rfx;

extern "C" {
// returns the number of texts on which the masks disagree with the naive scan (0 = all good); *n_lines_out = lines seen
int64_t emu_newline_masks(uint32_t seed, int64_t n_texts, int64_t max_len, int64_t* n_lines_out) {
    srand(seed);
    int64_t bad = 0, lines = 0;
    static const char alpha[] = "\n\n@@ACGT\r+N@";
    for (int64_t it = 0; it < n_texts; it++) {
        const size_t len = (size_t)(rand() % (max_len + 1));
        const int delta = rand() % 64;
        std::vector<uint8_t> buf(64 * 4 + len + 256, 'x');
        uint8_t* base = (uint8_t*)(((uintptr_t)buf.data() + 63) & ~(uintptr_t)63);
        uint8_t* text = base + delta;
        for (size_t i = 0; i < len; i++) text[i] = (uint8_t)alpha[rand() % 12];
        for (int i = 0; i < delta; i++) base[i] = (rand() & 1) ? '\n' : '@';          // bytes in front of the text ...
        for (size_t i = len; i < len + 128; i++) text[i] = (rand() & 1) ? '\n' : '@';  // ... and behind it must not count
        TextView tv;
        tv.len = len; tv.delta = (uint32_t)delta; tv.aligned = base;
        const uint64_t n_chunks = (len + delta + 63) / 64;
        std::vector<uint64_t> ls, ls2;
        std::vector<uint8_t> la, la2;
        bool ok = true;
        for (uint64_t c = 0; c < n_chunks; c++) {
            uint64_t nl, nla;
            chunk_masks(tv, c, nl, nla);
            if (nl != newline_mask64(tv, c)) ok = false;
            const int64_t p0 = (int64_t)(c * 64) - delta;
            for (uint64_t m = nl; m; m &= m - 1) {
                const int j = __builtin_ffsll((long long)m) - 1;
                ls.push_back((uint64_t)(p0 + j) + 1);
                la.push_back((uint8_t)((nla >> j) & 1));
            }
        }
        for (size_t i = 0; i < len; i++)
            if (text[i] == '\n') { ls2.push_back(i + 1); la2.push_back((i + 1 < len && text[i + 1] == '@') ? 1 : 0); }
        lines += (int64_t)ls2.size();
        if (!ok || ls != ls2 || la != la2) bad++;
    }
    if (n_lines_out) *n_lines_out = lines;
    return bad;
}
}
