// rfx_fastq.cu -- K1: FASTQ line split, the reference's line filters, 2-bit encoder.
//
// Replaces (paths relative to /root/reference/src/main/java/uni/bielefeld/cmg/reflexiv/pipeline/):
//   spark.read().text + DSFastqFilterWithQual        ReflexivDSMain.java:188-195, 4048-4072
//   DSFastqFilterOnlySeq                              ReflexivDataFrameCounter.java:243-289
//   nucleotideValue + the clip / minimum-length rule  ReflexivDataFrameCounter.java:471, 513-525
//
// Device data flow (all in HBM):
//   text bytes --newline scan--> line_start[] --filter scan--> seq_eff[] (read? effective length) --sum scan--> read table
//   (source offset, effective length, word offset) --warp-cooperative encoder--> packed 2-bit reads.
#include "rfx_internal.h"
#include "rfx_scan.cuh"

#include "rfx_newline.cuh"

namespace rfx {

struct NewlineMaskIn {  // pass 1: counts the newlines of a chunk and leaves its masks behind
    TextView tv;
    ulonglong2* masks;
    __device__ __forceinline__ uint64_t operator()(uint64_t chunk) const {
        uint64_t nl, nl_at;
        chunk_masks(tv, chunk, nl, nl_at);
        masks[chunk] = make_ulonglong2(nl, nl_at);
        return __popcll(nl);
    }
};
struct MaskCountIn {  // pass 2
    const ulonglong2* masks;
    __device__ __forceinline__ uint64_t operator()(uint64_t chunk) const { return __popcll(masks[chunk].x); }
};
struct MaskOut {
    TextView tv;
    const ulonglong2* masks;
    uint64_t* line_start;
    uint8_t* line_at;
    __device__ __forceinline__ void operator()(uint64_t chunk, uint64_t excl, uint64_t v) const {
        if (!v) return;
        const ulonglong2 m = masks[chunk];
        uint64_t mask = m.x;
        const int64_t p0 = (int64_t)(chunk * 64) - (int64_t)tv.delta;
        uint64_t idx = excl + 1;
        while (mask) {
            const int j = __ffsll((long long)mask) - 1;
            mask &= mask - 1;
            line_start[idx] = (uint64_t)(p0 + j) + 1;  // the line after this newline starts here
            line_at[idx] = (uint8_t)((m.y >> j) & 1ull);
            idx++;
        }
    }
};

struct NewlineIn {
    TextView tv;
    __device__ __forceinline__ uint64_t operator()(uint64_t chunk) const { return __popcll(newline_mask64(tv, chunk)); }
};
struct NewlineOut {
    TextView tv;
    const uint8_t* text;
    uint64_t* line_start;  // line_start[0] = 0 is written by finish_lines_kernel
    uint8_t* line_at;      // 1 if the line starts with '@'
    __device__ __forceinline__ void operator()(uint64_t chunk, uint64_t excl, uint64_t v) const {
        if (!v) return;
        uint64_t mask = newline_mask64(tv, chunk);
        const int64_t p0 = (int64_t)(chunk * 64) - (int64_t)tv.delta;
        uint64_t idx = excl + 1;
        while (mask) {
            const int j = __ffsll((long long)mask) - 1;
            mask &= mask - 1;
            const uint64_t next = (uint64_t)(p0 + j) + 1;  // the line after this newline starts here
            line_start[idx] = next;
            line_at[idx] = (next < tv.len && text[next] == '@') ? 1 : 0;
            idx++;
        }
    }
};

__global__ void finish_lines_kernel(const uint8_t* text, uint64_t len, const uint64_t* n_newlines, uint64_t* line_start, uint8_t* line_at,
                                    uint64_t* out /* [0] = n_lines */) {
    uint64_t n = *n_newlines;
    line_start[0] = 0;
    line_at[0] = (len > 0 && text[0] == '@') ? 1 : 0;
    if (len > 0 && text[len - 1] != '\n') {  // final line without terminator is still a record
        line_start[n + 1] = len + 1;
        n += 1;
    }
    out[0] = n;
}

// ------------------------------------------------------------------------------------------
// line filters
// ------------------------------------------------------------------------------------------
struct Lines {
    const uint8_t* text;
    const uint64_t* start;  // n_lines + 1
    uint64_t n_lines;
    uint32_t gap;  // 1: lines separated by '\n' (FASTQ text); 0: back-to-back reads (rfx_push_reads)
    __device__ __forceinline__ uint64_t len(uint64_t i) const {
        uint64_t s = start[i], e = start[i + 1] - gap;
        uint64_t l = e - s;
        if (gap && l > 0 && text[e - 1] == '\r') l--;
        return l;
    }
};

// DSFastqFilterWithQual as a scan of state-transition functions.  State = lineMark in {0..4},
// a function is 5 x 3 bits.  '@' line: 0->1 1->1 2->3 3->4 4->1 ; other line: 0->0 1->2 2->3 3->4 4->4
// (branch order of ReflexivDSMain.java:4051-4071: lineMark 2 and 3 consume any line blindly).
constexpr uint32_t FN_AT = 1u | (1u << 3) | (3u << 6) | (4u << 9) | (1u << 12);
constexpr uint32_t FN_OTHER = 0u | (2u << 3) | (3u << 6) | (4u << 9) | (4u << 12);
constexpr uint32_t FN_IDENT = 0u | (1u << 3) | (2u << 6) | (3u << 9) | (4u << 12);

struct OpCompose {  // (f then g)
    __device__ __forceinline__ uint32_t operator()(uint32_t f, uint32_t g) const {
        uint32_t r = 0;
#pragma unroll
        for (int s = 0; s < 5; s++) {
            uint32_t fs = (f >> (3 * s)) & 7u;
            r |= ((g >> (3 * fs)) & 7u) << (3 * s);
        }
        return r;
    }
};

struct LineFnIn {
    const uint8_t* line_at;  // written next to line_start by the newline pass: no random text access here
    __device__ __forceinline__ uint32_t operator()(uint64_t i) const { return line_at[i] ? FN_AT : FN_OTHER; }
};
// Per line: NOT_A_READ, or the effective length of the read (after clipping and the minimum-length rule).  Written once
// by the filter pass, so the two passes of the read-table scan load 4 coalesced bytes per line instead of chasing
// line_start[] into the text (the '\r' check) twice.
constexpr uint32_t NOT_A_READ = 0xffffffffu;

struct LineFnOut {
    uint64_t n_lines;  // lines known to exist from the start of this chunk on (the chunk's own, +2 if more text follows)
    uint32_t* seq_eff;
    const unsigned long long* state_in;  // lineMark at the start of the chunk
    Lines L;
    int k, fc, ec;
    __device__ __forceinline__ void operator()(uint64_t i, uint32_t excl, uint32_t fn) const {
        const uint32_t state_before = (excl >> (3u * (uint32_t)*state_in)) & 7u;  // prefix function applied to the carried lineMark
        // the sequence line is the one that moves lineMark 1 -> 2; the unit is emitted only when the
        // two following lines exist (lineMark 3 -> 4)
        const bool is_read = state_before == 1u && fn == FN_OTHER && i + 2 < n_lines;
        seq_eff[i] = is_read ? effective_read_len((int64_t)L.len(i), k, fc, ec) : NOT_A_READ;
    }
};

__global__ void fq_state_update_kernel(const uint32_t* total_fn, unsigned long long* state) {
    *state = (*total_fn >> (3u * (uint32_t)*state)) & 7u;
}

// The common case of the state machine above: a chunk whose lines are whole 4-line records (header '@', sequence not '@',
// two lines taken blindly), starting at whatever phase the carried lineMark says.  Then "which lines are reads" is i mod 4
// and the scan of transition functions (two passes over the per-line arrays) is not needed.  The check is exact: any
// line that breaks the layout sends the chunk down the general path.
__global__ void fq_regular_check_kernel(const uint8_t* __restrict__ line_at, unsigned long long* dstat) {
    const uint64_t n_lines = dstat[DS_FQ_NLINES];
    const uint32_t s0 = (uint32_t)dstat[DS_FQ_STATE];
    const uint32_t hdr = (s0 == 0u || s0 == 4u) ? 0u : (s0 == 1u ? 3u : (s0 == 2u ? 2u : 1u)), seq = (hdr + 1u) & 3u;
    bool bad = false;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_lines; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t ph = (uint32_t)i & 3u;
        const uint8_t at = line_at[i];
        bad |= (ph == hdr && !at) || (ph == seq && at);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicExch(&dstat[DS_FQ_IRREGULAR], 1ull);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        dstat[DS_FQ_SEQPOS] = seq;
        // lineMark after the last line: header -> 1, sequence -> 2, then 3, 4
        const uint32_t last = n_lines ? ((uint32_t)(n_lines - 1) + 4u - hdr) & 3u : 0u;
        dstat[DS_FQ_NEXT] = n_lines ? last + 1u : s0;
    }
}
__global__ void fq_regular_lines_kernel(Lines L, uint64_t n_known, uint32_t seq_pos, uint32_t* __restrict__ seq_eff, int k, int fc, int ec, unsigned long long* dstat) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < L.n_lines; i += (uint64_t)gridDim.x * blockDim.x)
        seq_eff[i] = (((uint32_t)i & 3u) == seq_pos && i + 2 < n_known) ? effective_read_len((int64_t)L.len(i), k, fc, ec) : NOT_A_READ;
    if (blockIdx.x == 0 && threadIdx.x == 0) dstat[DS_FQ_STATE] = dstat[DS_FQ_NEXT];
}

__device__ __forceinline__ bool is_atcgn(uint8_t a) { return a == 'A' || a == 'T' || a == 'C' || a == 'G' || a == 'N'; }

__global__ void flag_lines_kernel(Lines L, int mode, uint32_t* seq_eff, int k, int fc, int ec) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < L.n_lines; i += (uint64_t)gridDim.x * blockDim.x) {
        bool f = true;
        const uint64_t l = L.len(i);
        if (mode == RFX_FASTQ_COUNTER) {  // ReflexivDataFrameCounter.java:247-266
            const uint8_t* s = L.text + L.start[i];
            f = l > 20 && s[0] != '@' && s[0] != '+' && is_atcgn(s[0]) && is_atcgn(s[4]) && is_atcgn(s[9]) && is_atcgn(s[14]) && is_atcgn(s[19]);
        }
        seq_eff[i] = f ? effective_read_len((int64_t)l, k, fc, ec) : NOT_A_READ;
    }
}

// ------------------------------------------------------------------------------------------
// read table
// ------------------------------------------------------------------------------------------
// The scan element is one 64-bit word, reads in the high half and packed words in the low half (both < 2^32 per call:
// up to 137 G bases of text at once); bases and k-mer instances are only needed as totals and are summed on the side.
struct ReadIn {
    Lines L;
    const uint32_t* seq_eff;  // nullptr: every line is a read, lengths from line_start[]
    int k, fc, ec;
    __device__ __forceinline__ uint32_t eff(uint64_t i) const {
        return seq_eff ? seq_eff[i] : effective_read_len((int64_t)L.len(i), k, fc, ec);
    }
    __device__ __forceinline__ uint64_t src(uint64_t i) const { return L.start[i] + (uint64_t)fc; }
    __device__ __forceinline__ uint64_t operator()(uint64_t i) const {
        const uint32_t e = eff(i);
        if (e == NOT_A_READ) return 0ull;
        return (1ull << 32) | (uint64_t)((e + 31u) >> 5);
    }
};
// Whole 4-line records (fq_regular_check_kernel said so): element r is the sequence line 4 r + seq_pos -- a quarter of the elements
// of the scan over lines, and no per-line array of read flags in between.  A unit whose quality line is not among the lines known
// to exist is no read yet (DSFastqFilterWithQual returns the unit at its fourth line, ReflexivDSMain.java:4056-4059).
struct RegularReadIn {
    Lines L;
    uint64_t n_known;
    uint32_t seq_pos;
    int k, fc, ec;
    __device__ __forceinline__ uint64_t line(uint64_t r) const { return 4 * r + seq_pos; }
    __device__ __forceinline__ uint32_t eff(uint64_t r) const {
        const uint64_t i = line(r);
        return (i < L.n_lines && i + 2 < n_known) ? effective_read_len((int64_t)L.len(i), k, fc, ec) : NOT_A_READ;
    }
    __device__ __forceinline__ uint64_t src(uint64_t r) const { return L.start[line(r)] + (uint64_t)fc; }
    __device__ __forceinline__ uint64_t operator()(uint64_t r) const {
        const uint32_t e = eff(r);
        if (e == NOT_A_READ) return 0ull;
        return (1ull << 32) | (uint64_t)((e + 31u) >> 5);
    }
};
template <class In> struct ReadOutT {
    In in;
    uint64_t read_base, word_base;
    uint64_t* rd_src;
    uint32_t* rd_len;
    uint64_t* rd_woff;
    __device__ __forceinline__ void operator()(uint64_t i, uint64_t excl, uint64_t v) const {
        if (!v) return;
        const uint32_t e = in.eff(i);
        const uint64_t r = excl >> 32;
        rd_src[r] = in.src(i);
        rd_len[read_base + r] = e;
        rd_woff[read_base + r] = word_base + (excl & 0xffffffffull);
    }
};
// bases kept and k-mer instances of the appended reads: one pass over their lengths, one atomic pair per block
__global__ void __launch_bounds__(256) read_totals_kernel(const uint32_t* __restrict__ rd_len, uint64_t n, int k, unsigned long long* totals) {
    unsigned long long b = 0, inst = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t e = rd_len[i];
        b += e;
        inst += e ? (unsigned long long)(e - (uint32_t)k + 1u) : 0ull;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { b += __shfl_xor_sync(0xffffffffu, b, d); inst += __shfl_xor_sync(0xffffffffu, inst, d); }
    __shared__ unsigned long long sb[8], si[8];
    if ((threadIdx.x & 31) == 0) { sb[threadIdx.x >> 5] = b; si[threadIdx.x >> 5] = inst; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) { b += sb[w]; inst += si[w]; }
        if (b) atomicAdd(&totals[0], b);
        if (inst) atomicAdd(&totals[1], inst);
    }
}

// ------------------------------------------------------------------------------------------
// warp-cooperative 2-bit encoder.  Half a warp owns one read: lane l loads the l-th 16-byte ALIGNED chunk of the text the
// read lies in (one 128-bit load per lane, 16 bytes in flight per thread instead of 4), takes the four words of its
// neighbour's chunk by shuffle and cuts its own 16 bases out of the eight words at the read's byte offset (word rotation
// by selects, then funnel shifts).  Four SIMD-in-register compares turn 4 ASCII bases into one byte; two lanes make one
// 64-bit word.  A round covers 224 bases (14 producing lanes + 1 that only feeds its neighbour).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t encode4(uint32_t w) {  // 4 ASCII bases (first in the low byte) -> 8 bits, first base in bits 7..6
    const uint32_t mA = __vcmpeq4(w, 0x41414141u), mC = __vcmpeq4(w, 0x43434343u), mG = __vcmpeq4(w, 0x47474747u);
    const uint32_t codes = (~(mA | mC) & 0x02020202u) | (~(mA | mG) & 0x01010101u);
    return (codes * 0x40100401u) >> 24;
}
struct EncRead {
    uint32_t elen;
    uintptr_t a0;   // address of the first base
    uint64_t* dst;  // first packed word
};
__device__ __forceinline__ EncRead enc_read(const uint8_t* text, const uint64_t* rd_src, const uint32_t* rd_len, const uint64_t* rd_woff, uint64_t r, uint64_t n_new,
                                            uint64_t read_base, uint64_t* packed) {
    EncRead e{0u, (uintptr_t)text, packed};
    if (r < n_new) {
        e.elen = rd_len[read_base + r];
        e.a0 = (uintptr_t)(text + rd_src[r]);
        e.dst = packed + rd_woff[read_base + r];
    }
    return e;
}
__device__ __forceinline__ uint4 enc_load(const EncRead& e, int l, uint32_t round) {
    const uintptr_t addr = (e.a0 & ~(uintptr_t)15) + 16u * (uintptr_t)((uint32_t)l + 14u * round);
    return (l < 15 && addr < e.a0 + e.elen) ? *reinterpret_cast<const uint4*>(addr) : make_uint4(0u, 0u, 0u, 0u);
}
__device__ __forceinline__ void enc_round(const EncRead& e, uint4 w, int l, uint32_t round) {
    const uint32_t mis = (uint32_t)(e.a0 & 15), s = mis >> 2, sh = (mis & 3u) * 8u;
    uint32_t W[8] = {w.x, w.y, w.z, w.w, 0u, 0u, 0u, 0u};
    W[4] = __shfl_down_sync(0xffffffffu, w.x, 1, 16);
    W[5] = __shfl_down_sync(0xffffffffu, w.y, 1, 16);
    W[6] = __shfl_down_sync(0xffffffffu, w.z, 1, 16);
    W[7] = __shfl_down_sync(0xffffffffu, w.w, 1, 16);
    if (s & 1u) {  // W[t] <- W[t + s], t = 0..4
#pragma unroll
        for (int t = 0; t < 7; t++) W[t] = W[t + 1];
    }
    if (s & 2u) {
#pragma unroll
        for (int t = 0; t < 5; t++) W[t] = W[t + 2];
    }
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) bits |= encode4(__funnelshift_r(W[j], W[j + 1], sh)) << (24 - 8 * j);
    const uint32_t b0 = 16u * ((uint32_t)l + 14u * round);  // first base of this lane's window
    const uint32_t nv = b0 < e.elen ? min(16u, e.elen - b0) : 0u;
    bits = nv == 0u ? 0u : (bits & (0xffffffffu << (2u * (16u - nv))));  // clear the bases past the end of the read
    const uint32_t partner = __shfl_down_sync(0xffffffffu, bits, 1, 16);
    if (!(l & 1) && l < 14 && b0 < e.elen) e.dst[b0 >> 5] = ((uint64_t)bits << 32) | partner;
}
__global__ void __launch_bounds__(256) encode_reads_kernel(const uint8_t* __restrict__ text, const uint64_t* __restrict__ rd_src,
                                                            const uint32_t* __restrict__ rd_len, const uint64_t* __restrict__ rd_woff,
                                                            uint64_t n_new, uint64_t read_base, uint64_t* __restrict__ packed) {
    const int lane = threadIdx.x & 31, l = lane & 15, half = lane >> 4;
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    // four reads per warp and step, two per half: both table entries, then both text chunks are in flight before either is used
    for (uint64_t r0 = warp0 * 4; r0 < n_new; r0 += n_warps * 4) {
        const EncRead A = enc_read(text, rd_src, rd_len, rd_woff, r0 + half, n_new, read_base, packed);
        const EncRead B = enc_read(text, rd_src, rd_len, rd_woff, r0 + 2 + half, n_new, read_base, packed);
        const uint4 wa = enc_load(A, l, 0), wb = enc_load(B, l, 0);
        uint32_t rounds = (max(A.elen, B.elen) + 223u) / 224u;
        rounds = max(rounds, __shfl_xor_sync(0xffffffffu, rounds, 16));  // warp-uniform trip count: the shuffles need every lane
        enc_round(A, wa, l, 0);
        enc_round(B, wb, l, 0);
        for (uint32_t round = 1; round < rounds; round++) {  // reads longer than 224 bases
            const uint4 xa = enc_load(A, l, round), xb = enc_load(B, l, round);
            enc_round(A, xa, l, round);
            enc_round(B, xb, l, round);
        }
    }
}

__global__ void offsets_to_u64_kernel(const uint64_t* in, uint64_t n, uint64_t* out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = in[i];
}

static unsigned grid_for(uint64_t n, int block, unsigned cap = 0) {
    if (!cap) cap = sm_count() * 16;
    uint64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    return (unsigned)(g > cap ? cap : g);
}

static inline int parse_k(const Ctx* c) { return c->st_active ? c->k - 1 : c->k; }
static inline int parse_fc(const Ctx* c) { return c->st_active ? 0 : c->prm.front_clip; }
static inline int parse_ec(const Ctx* c) { return c->st_active ? 0 : c->prm.end_clip; }

// Builds the read table for `n_lines` lines and appends the packed reads to the context.
template <class In> static int append_reads_from(Ctx* c, const uint8_t* d_text, In in, uint64_t n_elems) {
    cudaStream_t st = c->stream;
    typedef ReadOutT<In> ReadOut;
    ScanPlan<uint64_t> plan;
    RFX_TRY(devbuf_reserve(c, c->scan_ws, ScanPlan<uint64_t>::workspace_elems(n_elems) * sizeof(uint64_t)));
    plan.bind(n_elems, c->scan_ws.as<uint64_t>());
    scan_prepare(plan, in, OpAddU64{}, (uint64_t)0, st);
    c->launches += 2 * plan.levels;
    uint64_t tot = 0;
    RFX_CUDA(c, cudaMemcpyAsync(&tot, plan.total, sizeof(tot), cudaMemcpyDeviceToHost, st));
    RFX_CUDA(c, cudaStreamSynchronize(st));
    const uint64_t n_new = tot >> 32, w_new = tot & 0xffffffffull;
    if (n_new == 0) return RFX_OK;
    DevBuf& rd_src = c->rd_src;
    RFX_TRY(devbuf_reserve(c, rd_src, n_new * sizeof(uint64_t)));
    if (c->st_active) {
        RFX_TRY(devbuf_reserve(c, c->st_len, n_new * sizeof(uint32_t)));
        RFX_TRY(devbuf_reserve(c, c->st_woff, n_new * sizeof(uint64_t)));
        ReadOut out{in, 0, 0, rd_src.as<uint64_t>(), c->st_len.as<uint32_t>(), c->st_woff.as<uint64_t>()};
        scan_apply(plan, in, out, OpAddU64{}, (uint64_t)0, st);
        c->launches += 1;
        return stitch_scan_reads(c, d_text, rd_src.as<uint64_t>(), c->st_len.as<uint32_t>(), n_new);
    }
    int rc = RFX_OK;
    do {
        if ((rc = devbuf_reserve(c, c->rd_len, (c->n_reads + n_new) * sizeof(uint32_t), true)) != RFX_OK) break;
        if ((rc = devbuf_reserve(c, c->rd_woff, (c->n_reads + n_new) * sizeof(uint64_t), true)) != RFX_OK) break;
        if ((rc = devbuf_reserve(c, c->packed, (c->n_words + w_new + 8) * sizeof(uint64_t), true)) != RFX_OK) break;
        unsigned long long* totals = c->dstat.as<unsigned long long>() + DS_READ_TOTALS;
        cudaMemsetAsync(totals, 0, 2 * sizeof(uint64_t), st);
        ReadOut out{in, c->n_reads, c->n_words, rd_src.as<uint64_t>(), c->rd_len.as<uint32_t>(), c->rd_woff.as<uint64_t>()};
        scan_apply(plan, in, out, OpAddU64{}, (uint64_t)0, st);
        read_totals_kernel<<<grid_for(n_new, 256, sm_count() * 2), 256, 0, st>>>(c->rd_len.as<uint32_t>() + c->n_reads, n_new, c->k, totals);
        encode_reads_kernel<<<grid_for(n_new * 8, 256, sm_count() * 32), 256, 0, st>>>(d_text, rd_src.as<uint64_t>(), c->rd_len.as<uint32_t>(),
                                                                                c->rd_woff.as<uint64_t>(), n_new, c->n_reads,
                                                                                c->packed.as<uint64_t>());
        // padding words after the last read: packed_window() may look one word ahead
        cudaMemsetAsync(c->packed.as<uint64_t>() + c->n_words + w_new, 0, 8 * sizeof(uint64_t), st);
        c->launches += 2;
        uint64_t ht[2] = {0, 0};
        cudaMemcpyAsync(ht, totals, sizeof(ht), cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { rc = ctx_fail(c, RFX_E_CUDA, "encode failed: %s", cudaGetErrorString(e)); break; }
        c->n_reads += n_new;
        c->n_words += w_new;
        c->n_bases += ht[0];
        c->n_instances += ht[1];
    } while (0);
    return rc;
}

// While the stitch stage is open (rfx_stitch.cu) the reads are the UNclipped sequence lines of DSLowCoverageReadDetection
// (ReflexivDSMain.java:1467-1475: read = units[1], skipped when readLength - (k-1) <= 1) and are scanned, not stored.
static int append_reads(Ctx* c, const uint8_t* d_text, Lines L, const uint32_t* seq_flag) {
    if (L.n_lines && (L.start == nullptr)) return ctx_fail(c, RFX_E_INVALID, "append_reads: no line table");
    return append_reads_from(c, d_text, ReadIn{L, seq_flag, parse_k(c), parse_fc(c), parse_ec(c)}, L.n_lines);
}
static int append_reads_regular(Ctx* c, const uint8_t* d_text, Lines L, uint64_t n_known, uint32_t seq_pos) {
    const uint64_t n_rec = L.n_lines > seq_pos ? (L.n_lines - seq_pos + 3) / 4 : 0;
    return append_reads_from(c, d_text, RegularReadIn{L, n_known, seq_pos, parse_k(c), parse_fc(c), parse_ec(c)}, n_rec);
}
__global__ void fq_state_next_kernel(unsigned long long* dstat) { dstat[DS_FQ_STATE] = dstat[DS_FQ_NEXT]; }

// One chunk of FASTQ text (whole lines).  `first_chunk` resets the carried lineMark; `more_follows` promises that at
// least two more lines come after this chunk, so a unit that starts in its last lines is known to complete.
int stage_parse_fastq(Ctx* c, const uint8_t* d_text, size_t len, bool first_chunk, bool more_follows) {
    cudaStream_t st = c->stream;
    unsigned long long* fq_state = c->dstat.as<unsigned long long>() + DS_FQ_STATE;
    if (first_chunk) RFX_CUDA(c, cudaMemsetAsync(fq_state, 0, sizeof(uint64_t), st));
    if (len == 0) return RFX_OK;
    stage_begin(c);
    TextView tv;
    tv.len = len;
    tv.delta = (uint32_t)((uintptr_t)d_text & 63);
    tv.aligned = d_text - tv.delta;
    const uint64_t n_chunks = (len + tv.delta + 63) / 64;

    // 1. newline scan
    ScanPlan<uint64_t> nl;
    RFX_TRY(devbuf_reserve(c, c->scan_ws, ScanPlan<uint64_t>::workspace_elems(n_chunks) * sizeof(uint64_t)));
    nl.bind(n_chunks, c->scan_ws.as<uint64_t>());
    // pass 1 leaves the newline / '@' masks of every chunk behind (a quarter of the text), pass 2 works from them
    const bool use_masks = !getenv("RFX_NEWLINE_REREAD");  // (measured alternative: pass 2 reads the text again)
    if (use_masks) {
        RFX_TRY(devbuf_reserve(c, c->nl_masks, (n_chunks + 1) * sizeof(ulonglong2)));
        scan_prepare(nl, NewlineMaskIn{tv, c->nl_masks.as<ulonglong2>()}, OpAddU64{}, (uint64_t)0, st);
    } else
    scan_prepare(nl, NewlineIn{tv}, OpAddU64{}, (uint64_t)0, st);
    c->launches += 2 * nl.levels;
    uint64_t n_newlines = 0;
    RFX_CUDA(c, cudaMemcpyAsync(&n_newlines, nl.total, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    RFX_CUDA(c, cudaStreamSynchronize(st));
    RFX_TRY(devbuf_reserve(c, c->line_start, (n_newlines + 2) * sizeof(uint64_t)));
    RFX_TRY(devbuf_reserve(c, c->line_at, n_newlines + 2));
    uint64_t* ls = c->line_start.as<uint64_t>();
    uint8_t* lat = c->line_at.as<uint8_t>();
    if (use_masks) scan_apply(nl, MaskCountIn{c->nl_masks.as<ulonglong2>()}, MaskOut{tv, c->nl_masks.as<ulonglong2>(), ls, lat}, OpAddU64{}, (uint64_t)0, st);
    else scan_apply(nl, NewlineIn{tv}, NewlineOut{tv, d_text, ls, lat}, OpAddU64{}, (uint64_t)0, st);
    finish_lines_kernel<<<1, 1, 0, st>>>(d_text, len, nl.total, ls, lat, c->dstat.as<uint64_t>() + DS_FQ_NLINES);
    c->launches += 2;
    const int fmode = c->st_active ? RFX_FASTQ_RUN : c->prm.fastq_mode;  // the stitch stage always reads units (ReflexivDSMain.java:601-606)
    const bool try_regular = fmode == RFX_FASTQ_RUN && !getenv("RFX_FASTQ_GENERAL");  // (tests force the general path)
    if (try_regular) {
        RFX_CUDA(c, cudaMemsetAsync(c->dstat.as<uint64_t>() + DS_FQ_IRREGULAR, 0, sizeof(uint64_t), st));
        fq_regular_check_kernel<<<sm_count() * 8, 256, 0, st>>>(lat, c->dstat.as<unsigned long long>());
        c->launches++;
    }
    uint64_t fq[4] = {0, 0, 0, 0};  // lines, irregular?, position of the sequence lines, lineMark behind the chunk
    RFX_CUDA(c, cudaMemcpyAsync(fq, c->dstat.as<uint64_t>() + DS_FQ_NLINES, (try_regular ? 4 : 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    RFX_CUDA(c, cudaStreamSynchronize(st));
    const uint64_t n_lines = fq[0];
    if (n_lines == 0) { c->ms[0] += stage_end(c); return RFX_OK; }

    // 2. which lines are reads
    Lines L{d_text, ls, n_lines, 1u};
    if (try_regular && !fq[1] && !getenv("RFX_FASTQ_LINE_FLAGS")) {  // (measured alternative: per-line flags + the scan over lines)
        fq_state_next_kernel<<<1, 1, 0, st>>>(c->dstat.as<unsigned long long>());
        c->launches++;
        const int rc = append_reads_regular(c, d_text, L, n_lines + (more_follows ? 2u : 0u), (uint32_t)fq[2]);
        c->ms[0] += stage_end(c);
        return rc;
    }
    RFX_TRY(devbuf_reserve(c, c->seq_flag, n_lines * sizeof(uint32_t)));
    const uint32_t* flags = c->seq_flag.as<uint32_t>();
    if (try_regular && !fq[1]) {
        fq_regular_lines_kernel<<<grid_for(n_lines, 256), 256, 0, st>>>(L, n_lines + (more_follows ? 2u : 0u), (uint32_t)fq[2], c->seq_flag.as<uint32_t>(), parse_k(c),
                                                                       parse_fc(c), parse_ec(c), c->dstat.as<unsigned long long>());
        c->launches++;
    } else if (fmode == RFX_FASTQ_RUN) {
        ScanPlan<uint32_t> fs;
        RFX_TRY(devbuf_reserve(c, c->scan_ws, ScanPlan<uint32_t>::workspace_elems(n_lines) * sizeof(uint32_t)));
        fs.bind(n_lines, c->scan_ws.as<uint32_t>());
        scan_prepare(fs, LineFnIn{lat}, OpCompose{}, FN_IDENT, st);
        scan_apply(fs, LineFnIn{lat}, LineFnOut{n_lines + (more_follows ? 2u : 0u), c->seq_flag.as<uint32_t>(), fq_state, L, parse_k(c), parse_fc(c), parse_ec(c)},
                   OpCompose{}, FN_IDENT, st);
        fq_state_update_kernel<<<1, 1, 0, st>>>(fs.total, fq_state);
        c->launches += 2 * fs.levels + 1;
    } else if (fmode == RFX_FASTQ_COUNTER) {
        flag_lines_kernel<<<grid_for(n_lines, 256), 256, 0, st>>>(L, RFX_FASTQ_COUNTER, c->seq_flag.as<uint32_t>(), parse_k(c), parse_fc(c), parse_ec(c));
        c->launches += 1;
    } else {
        flags = nullptr;
    }
    // 3. read table + encode
    int rc = append_reads(c, d_text, L, flags);
    c->ms[0] += stage_end(c);
    return rc;
}

int stage_push_reads(Ctx* c, const uint8_t* h_bases, const uint64_t* h_offsets, uint64_t n_reads) {
    if (n_reads == 0) return RFX_OK;
    cudaStream_t st = c->stream;
    const uint64_t total = h_offsets[n_reads];
    RFX_TRY(devbuf_reserve(c, c->text, total + 64));
    RFX_TRY(devbuf_reserve(c, c->line_start, (n_reads + 1) * sizeof(uint64_t)));
    RFX_CUDA(c, cudaMemcpyAsync(c->text.p, h_bases, total, cudaMemcpyHostToDevice, st));
    RFX_CUDA(c, cudaMemsetAsync(c->text.as<uint8_t>() + total, 0, 64, st));
    RFX_CUDA(c, cudaMemcpyAsync(c->line_start.p, h_offsets, (n_reads + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    stage_begin(c);
    Lines L{c->text.as<uint8_t>(), c->line_start.as<uint64_t>(), n_reads, 0u};
    int rc = append_reads(c, c->text.as<uint8_t>(), L, nullptr);
    c->ms[0] += stage_end(c);
    return rc;
}

}  // namespace rfx
