// stitch_emu.cpp -- TEST-ONLY: runs the kernels of reflexiv_b200/csrc/rfx_stitch_kernels.cuh on the host, thread by thread, in the
// order rfx_stitch.cu launches them, so that their logic is checked against oracle/stitch_oracle.c on a machine without a GPU.
// A shim stands in for the CUDA built-ins (one "thread" runs at a time: races are not modelled, the order-free results are).
// Not part of libreflexiv_cuda; nothing in reflexiv_b200/ uses it.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#define RFX_STITCH_HOSTEMU 1
#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(x)
struct EmuDim { unsigned x = 1, y = 1, z = 1; };
static EmuDim blockIdx, threadIdx, gridDim, blockDim;
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long v) { unsigned long long o = *p; if (o == cmp) *p = v; return o; }
static inline uint32_t atomicCAS(uint32_t* p, uint32_t cmp, uint32_t v) { uint32_t o = *p; if (o == cmp) *p = v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
static inline uint32_t atomicOr(uint32_t* p, uint32_t v) { uint32_t o = *p; *p |= v; return o; }
static inline void __syncwarp() {}
namespace rfx {
struct U64x3 { uint64_t a, b, c; };
static const uint32_t NONE32 = 0xffffffffu;
}  // namespace rfx

#include "../../reflexiv_b200/csrc/rfx_stitch_kernels.cuh"

using namespace rfx;
using namespace rfx::stitch;

template <class F> static void launch(unsigned grid, unsigned block, F f) {
    gridDim.x = grid; blockDim.x = block;
    for (unsigned b = 0; b < grid; b++)
        for (unsigned t = 0; t < block; t++) { blockIdx.x = b; threadIdx.x = t; f(); }
}

extern "C" {

struct EmuStitchOut {
    uint64_t n_contigs;
    uint64_t* off;
    char* bases;
    int32_t *left, *right;
    uint64_t stats[6];  // probes, fragments, after pass 1, joined on both sides, stitched records, rings
};

void emu_stitch_free(EmuStitchOut* o) { free(o->off); free(o->bases); free(o->left); free(o->right); memset(o, 0, sizeof(*o)); }

// `chunk_reads`: the reads are scanned in pieces of that many (as rfx_push_fastq does chunk by chunk); hit_cap0: first capacity of the
// per-chunk fragment list (small values exercise the overflow re-run)
int emu_stitch(uint64_t n, const uint64_t* off, const char* bases, const int32_t* cl, const int32_t* cr, const uint8_t* text, uint64_t n_reads,
               const uint64_t* starts, const uint32_t* lens, int k, int min_contig, uint64_t chunk_reads, uint64_t hit_cap0, EmuStitchOut* out) {
    memset(out, 0, sizeof(*out));
    if (k < 2 || k > 31) return -1;
    const unsigned G = 3, B = 64;
    // ---- stage_stitch_begin ----
    uint64_t cap = 1024;
    while (cap < 4 * n + 16) cap <<= 1;
    std::vector<uint64_t> keys(cap, ~0ull), firstk(n + 1, 0);
    std::vector<uint32_t> vals(cap, NONE32), bloom(ST_BLOOM_BITS / 32, 0);
    // the scan reads the text in aligned 4-byte words: give it the slack the device buffer has
    uint64_t text_len = 0;
    for (uint64_t r = 0; r < n_reads; r++) if (starts[r] + lens[r] > text_len) text_len = starts[r] + lens[r];
    std::vector<uint8_t> padded(text_len + 16, 0);
    if (text_len) memcpy(padded.data() + 8, text, text_len);
    text = padded.data() + 8;
    unsigned long long ctr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (n) {
        launch(G, B, [&] { probe_first_kernel(off, bases, n, k, firstk.data()); });
        launch(G, B, [&] { probe_insert_kernel(off, bases, cl, cr, n, k, firstk.data(), keys.data(), vals.data(), cap - 1, bloom.data(), &ctr[2]); });
        launch(G, B, [&] { count_keys_kernel(keys.data(), cap, &ctr[3]); });
    }
    out->stats[0] = ctr[3];
    // ---- stitch_scan_reads, chunk by chunk ----
    std::vector<Frag> frags;
    std::vector<uint8_t> codes;
    std::vector<uint32_t> elen(n_reads ? n_reads : 1);
    for (uint64_t r = 0; r < n_reads; r++) elen[r] = ((int64_t)lens[r] - (k - 1) <= 1) ? 0u : lens[r];  // effective_read_len(len, k - 1, 0, 0)
    if (chunk_reads == 0) chunk_reads = n_reads ? n_reads : 1;
    std::vector<Hit> hits(hit_cap0 ? hit_cap0 : 1);
    for (uint64_t r0 = 0; r0 < n_reads && ctr[3]; r0 += chunk_reads) {
        const uint64_t nr = n_reads - r0 < chunk_reads ? n_reads - r0 : chunk_reads;
        for (int attempt = 0;; attempt++) {
            ctr[0] = ctr[1] = 0;
            launch(G, B, [&] { stitch_scan_kernel(text, starts + r0, elen.data() + r0, nr, k, keys.data(), vals.data(), cap - 1, bloom.data(), hits.data(), hits.size(), ctr); });
            if (ctr[0] <= hits.size()) break;
            if (attempt) return -2;
            hits.resize(ctr[0]);
        }
        if (!ctr[0]) continue;
        const uint64_t f0 = frags.size(), c0 = codes.size();
        frags.resize(f0 + ctr[0]);
        codes.resize(c0 + ctr[1] + 16);
        launch(G, B, [&] { stitch_copy_kernel(text, hits.data(), ctr[0], k, keys.data(), vals.data(), cap - 1, codes.data(), c0, frags.data(), f0); });
        codes.resize(c0 + ctr[1]);
    }
    // ---- stage_stitch_finish ----
    const uint64_t nf = frags.size();
    out->stats[1] = nf;
    if (n == 0) return 0;
    std::vector<uint32_t> nxt(n), prv(n);
    std::vector<uint8_t> role(n, 0);
    std::vector<uint64_t> out_len(n), slot(n);
    std::vector<int32_t> out_right(n);
    launch(G, B, [&] { fill_u32_kernel(nxt.data(), n, NONE32); });
    launch(G, B, [&] { fill_u32_kernel(prv.data(), n, NONE32); });
    codes.resize(codes.size() + 16);
    if (nf) {
        launch(G, B, [&] { frag_pick_kernel(frags.data(), codes.data(), nf, 0, nullptr, nxt.data()); });
        launch(G, B, [&] { frag_pick_kernel(frags.data(), codes.data(), nf, 1, nxt.data(), prv.data()); });
        launch(G, B, [&] { count_picked_kernel(frags.data(), nf, nxt.data(), prv.data(), &ctr[4]); });
    }
    launch(G, B, [&] { chain_heads_kernel(n, nxt.data(), prv.data(), frags.data(), role.data()); });
    launch(G, B, [&] { ring_heads_kernel(n, nxt.data(), prv.data(), frags.data(), firstk.data(), role.data(), &ctr[7]); });
    launch(G, B, [&] { chain_sizes_kernel(n, k, min_contig, off, cl, cr, nxt.data(), prv.data(), frags.data(), role.data(), out_len.data(), out_right.data(), &ctr[6]); });
    // the transform-scan of rfx_scan.cuh, element by element
    KeepIn in{out_len.data()};
    U64x3 tot{0, 0, 0};
    for (uint64_t c = 0; c < n; c++) { const U64x3 v = in(c); tot.a += v.a; tot.b += v.b; }
    out->off = (uint64_t*)malloc((tot.a + 1) * sizeof(uint64_t));
    out->left = (int32_t*)malloc((tot.a + 1) * sizeof(int32_t));
    out->right = (int32_t*)malloc((tot.a + 1) * sizeof(int32_t));
    out->bases = (char*)malloc(tot.b + 16);
    KeepOut ko{cl, out_right.data(), out->off, slot.data(), out->left, out->right};
    U64x3 acc{0, 0, 0};
    for (uint64_t c = 0; c < n; c++) { const U64x3 v = in(c); ko(c, acc, v); acc.a += v.a; acc.b += v.b; }
    out->off[tot.a] = tot.b;
    launch(5, 32, [&] { chain_gather_kernel(n, k, off, bases, nxt.data(), prv.data(), frags.data(), codes.data(), slot.data(), out->off, out->bases); });
    out->n_contigs = tot.a;
    out->stats[2] = ctr[4]; out->stats[3] = ctr[5]; out->stats[4] = ctr[6]; out->stats[5] = ctr[7];
    return 0;
}
}
