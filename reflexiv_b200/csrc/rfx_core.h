// rfx_core.h -- arithmetic shared by every kernel of libreflexiv_cuda.
//
// Everything here is __host__ __device__ so the same bit manipulation that runs
// inside the sm_100a kernels can be exercised by the host-side logic tests
// (tests/hostemu) without a GPU.  The kernels themselves live in the .cu files.
//
// Reference citations are relative to
//   /root/reference/src/main/java/uni/bielefeld/cmg/reflexiv/pipeline/
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define RFX_HD __host__ __device__ __forceinline__
#else
#define RFX_HD inline
#endif

namespace rfx {

typedef unsigned __int128 u128;

// ---------------------------------------------------------------------------------------------
// hashing (murmur3 finalisers; bijective, so a hash of a canonical m-mer is as good as the m-mer)
// ---------------------------------------------------------------------------------------------
RFX_HD uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}
RFX_HD uint64_t fmix64(uint64_t h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33;
    return h;
}
RFX_HD uint64_t key_hash(uint64_t k) { return fmix64(k); }
RFX_HD uint64_t key_hash(u128 k) { return fmix64((uint64_t)k ^ fmix64((uint64_t)(k >> 64) + 0x9e3779b97f4a7c15ULL)); }

// nucleotideValue (ReflexivDataFrameCounter.java:513-525): A=0 C=1 G=2, anything else = 3.
RFX_HD uint32_t base_code(uint32_t c) { return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : 3u; }

RFX_HD uint32_t brev32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
    return (x >> 16) | (x << 16);
#endif
}
RFX_HD uint64_t brev64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __brevll(x);
#else
    return ((uint64_t)brev32((uint32_t)x) << 32) | brev32((uint32_t)(x >> 32));
#endif
}

// Reverse complement of a right-aligned 2-bit string of nb bases.
RFX_HD uint64_t revcomp(uint64_t x, int nb) {
    uint64_t y = brev64(x);                                                        // bases reversed, bit pairs swapped
    y = ((y >> 1) & 0x5555555555555555ULL) | ((y & 0x5555555555555555ULL) << 1);   // un-swap the pairs
    return (~y) >> (64 - 2 * nb);                                                  // complement = xor 3
}
RFX_HD u128 revcomp(u128 x, int nb) {
    uint64_t hi = (uint64_t)(x >> 64), lo = (uint64_t)x;
    uint64_t rh = brev64(lo), rl = brev64(hi);
    rh = ((rh >> 1) & 0x5555555555555555ULL) | ((rh & 0x5555555555555555ULL) << 1);
    rl = ((rl >> 1) & 0x5555555555555555ULL) | ((rl & 0x5555555555555555ULL) << 1);
    u128 y = ((u128)(~rh) << 64) | (uint64_t)(~rl);
    return y >> (128 - 2 * nb);
}

template <class KT> RFX_HD KT mask_bases(int nb) {
    return (nb * 2 >= (int)(8 * sizeof(KT))) ? ~(KT)0 : (((KT)1 << (2 * nb)) - 1);
}

// ---------------------------------------------------------------------------------------------
// packed reads: 32 bases per 64-bit word, first base in the two most significant bits
// ---------------------------------------------------------------------------------------------
RFX_HD uint32_t packed_base(const uint64_t* w, uint64_t j) { return (uint32_t)(w[j >> 5] >> (62 - 2 * (j & 31))) & 3u; }

// 32 bases starting at base j (may read w[(j>>5)+1]; buffers carry padding words)
RFX_HD uint64_t packed_window(const uint64_t* w, uint64_t j) {
    uint64_t idx = j >> 5;
    uint32_t s = 2 * (uint32_t)(j & 31);
    uint64_t hi = w[idx];
    if (s == 0) return hi;
    return (hi << s) | (w[idx + 1] >> (64 - s));
}

// ---------------------------------------------------------------------------------------------
// Does the reference extract k-mers from this read, and how many?
//   k <= 31: ReflexivDataFrameCounter.java:471 / ReflexivDSMain.java:3968
//            "readLength - k - endClip <= 1 || frontClip > readLength" -> skipped
//   k  > 31: ReflexivDataFrameCounter64.java:410 "readLength - k - endClip + 1 <= 0"
// The k-mers come from bases [frontClip, readLength - endClip).
// ---------------------------------------------------------------------------------------------
RFX_HD uint32_t effective_read_len(int64_t len, int k, int front_clip, int end_clip) {
    if (front_clip > len) return 0;
    if (k <= 31) { if (len - k - end_clip <= 1) return 0; }
    else { if (len - k - end_clip + 1 <= 0) return 0; }
    int64_t e = len - end_clip - front_clip;
    return e >= k ? (uint32_t)e : 0u;
}

// ---------------------------------------------------------------------------------------------
// super-k-mer records
//   A record is RECW 64-bit words.  Bits 63..48 of word 0 hold n_k (k-mers in the record, >= 1);
//   the n_k + k - 1 bases follow as one MSB-first 2-bit stream starting at bit 47 of word 0.
//   Capacity (RECW*64 - 16) / 2 bases: 56 for RECW = 2 (k <= 31), 120 for RECW = 4 (k <= 63).
// ---------------------------------------------------------------------------------------------
RFX_HD int rec_words_for_k(int k) { return k <= 31 ? 2 : 4; }
RFX_HD int rec_max_bases(int recw) { return (recw * 64 - 16) / 2; }
RFX_HD int rec_max_kmers(int recw, int k) { return rec_max_bases(recw) - k + 1; }

// `k` is needed to zero the bits behind the last base, so that equal super-k-mers are equal bit patterns
// (the counting kernel de-duplicates whole records before it touches k-mers).
template <int RECW> RFX_HD void rec_build(const uint64_t* read_words, uint32_t base_pos, uint32_t n_k, int k, uint64_t* out) {
    out[0] = ((uint64_t)n_k << 48) | (packed_window(read_words, base_pos) >> 16);
#pragma unroll
    for (int i = 1; i < RECW; i++) out[i] = packed_window(read_words, (uint64_t)base_pos + 32u * i - 8u);
    const uint32_t used = 16u + 2u * (n_k + (uint32_t)k - 1u);  // header + bases, in bits from the top of word 0
#pragma unroll
    for (int i = 0; i < RECW; i++) {
        const uint32_t lo = 64u * i;
        if (used <= lo) out[i] = 0;
        else if (used < lo + 64u) out[i] &= ~0ull << (lo + 64u - used);
    }
}

// Visit every canonical k-mer of a record.  Same rolling update as the reference extractor
// (ReflexivDataFrameCounter.java:483-506): fwd shifts the base in at the bottom, rc at the top,
// the smaller of the two is the canonical key (forward on a tie).
template <class KT, int RECW, class F> RFX_HD void rec_foreach_kmer(const uint64_t* rec, int k, F&& f) {
    uint64_t r[RECW];
#pragma unroll
    for (int i = 0; i < RECW; i++) r[i] = rec[i];
    const uint32_t n_k = (uint32_t)(r[0] >> 48);
    const uint32_t nb = n_k + (uint32_t)k - 1u;
    const KT msk = mask_bases<KT>(k);
    const int top = 2 * (k - 1);
    KT fwd = 0, rc = 0;
    uint64_t cur = r[0] << 16;
    uint32_t avail = 24;
    for (uint32_t t = 0; t < nb; t++) {
        if (avail == 0) {
            cur = r[1];
#pragma unroll
            for (int i = 1; i + 1 < RECW; i++) r[i] = r[i + 1];
            avail = 32;
        }
        KT v = (KT)(cur >> 62);
        cur <<= 2; avail--;
        fwd = ((fwd << 2) | v) & msk;
        rc = (rc >> 2) | ((v ^ (KT)3) << top);
        if (t + 1 >= (uint32_t)k) f(fwd < rc ? fwd : rc);
    }
}

// ---------------------------------------------------------------------------------------------
// minimiser binning of one read (pass 1).  One thread walks one read.
//   m-mer hash      h(j) = mmer_hash(min(mmer, revcomp(mmer)))         (m <= 16)
//   k-mer minimiser      = min over its w = k - m + 1 m-mers of h      (strand symmetric)
//   bin                  = fmix32(minimiser ^ salt) * n_bins >> 32
// Consecutive k-mers of equal bin form one super-k-mer run; runs are cut at max_nk k-mers.
// The sliding minimum is van Herk / Gil-Werman on the fly: h is kept for two blocks of w
// positions in `ring` (2*w entries, stride `rs` so lanes of a warp interleave in shared memory);
// at the end of each block its suffix minima are computed in place; the minimum of a window
// that straddles blocks b-1 | b is min(suffix[b-1][i], prefix[b][j]).  The schedule depends on
// the position only, so the 32 reads of a warp stay converged.
// ---------------------------------------------------------------------------------------------
struct BinParams {
    int k;          // k-mer length
    int m;          // minimiser length (<= 16, <= k)
    int w;          // k - m + 1
    uint32_t n_bins;
    uint32_t max_nk;  // k-mers per record
};

// Order of the m-mers inside a k-mer: two multiplies and a fold are enough to make "smallest hash" an arbitrary,
// strand-symmetric choice; the bin is re-mixed separately below.
RFX_HD uint32_t mmer_hash(uint32_t canon_mmer) {
    uint32_t h = (canon_mmer ^ 0x3c6ef372u) * 0x9E3779B1u;
    h ^= h >> 15;
    return h * 0x85ebca6bu;
}

RFX_HD uint32_t bin_of_minimizer(uint32_t hmin, uint32_t n_bins) {
    return (uint32_t)(((uint64_t)fmix32(hmin ^ 0x9e3779b9u) * n_bins) >> 32);
}

// emit(bin, first_kmer_index, n_kmers)
template <class Emit> RFX_HD void bin_scan_read(const uint64_t* rd, uint32_t len, const BinParams& P, uint32_t* ring,
                                               uint32_t rs, Emit&& emit) {
    const int k = P.k, m = P.m, w = P.w;
    if (len < (uint32_t)k) return;
    const uint32_t mmask = (m >= 16) ? 0xffffffffu : ((1u << (2 * m)) - 1u);
    const int mtop = 2 * (m - 1);
    uint32_t mf = 0, mr = 0;
    uint64_t cur = 0;
    uint32_t e = 0;
    for (; e + 1 < (uint32_t)m; e++) {  // the first m-1 bases complete no m-mer
        if ((e & 31u) == 0) cur = rd[e >> 5];
        const uint32_t v = (uint32_t)(cur >> 62);
        cur <<= 2;
        mf = ((mf << 2) | v) & mmask;
        mr = (mr >> 2) | ((v ^ 3u) << mtop);
    }
    uint32_t* blk_cur = ring;                       // block being filled
    uint32_t* blk_prev = ring + (uint32_t)w * rs;   // previous block, already turned into suffix minima
    int pib = 0;                                    // position in block of the m-mer being produced
    uint32_t pmin = 0xffffffffu;                    // prefix minimum of the current block
    uint32_t prev_h = 0, run_bin = 0, run_start = 0;
    const uint32_t n_mmers = len - (uint32_t)m + 1u;
    for (uint32_t j = 0; j < n_mmers; j++, e++) {
        if ((e & 31u) == 0) cur = rd[e >> 5];
        const uint32_t v = (uint32_t)(cur >> 62);
        cur <<= 2;
        mf = ((mf << 2) | v) & mmask;
        mr = (mr >> 2) | ((v ^ 3u) << mtop);
        const uint32_t h = mmer_hash(mf < mr ? mf : mr);
        blk_cur[(uint32_t)pib * rs] = h;
        pmin = h < pmin ? h : pmin;
        if (j + 1 >= (uint32_t)w) {
            // k-mer i = j - w + 1 is complete: its m-mers are i .. j
            const uint32_t i = j + 1 - (uint32_t)w;
            uint32_t hmin = pmin;
            if (pib != w - 1) {  // window starts inside the previous block at offset pib + 1
                const uint32_t sfx = blk_prev[(uint32_t)(pib + 1) * rs];
                hmin = sfx < hmin ? sfx : hmin;
            }
            if (i == 0) {
                prev_h = hmin; run_bin = bin_of_minimizer(hmin, P.n_bins);
            } else {
                uint32_t bin = run_bin;
                if (hmin != prev_h) { prev_h = hmin; bin = bin_of_minimizer(hmin, P.n_bins); }  // same minimiser, same bin
                if (bin != run_bin || i - run_start == P.max_nk) {
                    emit(run_bin, run_start, i - run_start);
                    run_bin = bin; run_start = i;
                }
            }
        }
        if (++pib == w) {
            // block finished: turn its h values into suffix minima, swap the two halves of the ring
            uint32_t sm = 0xffffffffu;
            for (int t = w - 1; t >= 0; t--) {
                const uint32_t x = blk_cur[(uint32_t)t * rs];
                sm = x < sm ? x : sm;
                blk_cur[(uint32_t)t * rs] = sm;
            }
            uint32_t* tmp = blk_cur; blk_cur = blk_prev; blk_prev = tmp;
            pib = 0; pmin = 0xffffffffu;
        }
    }
    emit(run_bin, run_start, len - (uint32_t)k + 1u - run_start);
}

// The k-mer with index `off` of a record (forward orientation, right aligned), cut straight out of the record's bit
// stream: what the counting kernel's warp-wide expansion does per lane (rfx_count.cu: expand_warp).  Static indices
// only, so the words stay in registers.
template <class KT, int RECW> RFX_HD KT rec_kmer_at(const uint64_t (&w)[RECW], uint32_t off, int k) {
    const uint32_t b = 16u + 2u * off;  // first bit of the k-mer in the bit stream (bit 0 = top bit of word 0)
    if (RECW == 2 && sizeof(KT) == 8) {
        uint64_t hi;
        if (b < 64u) hi = (w[0] << b) | ((w[1] >> 1) >> (63u - b));
        else hi = w[1] << (b - 64u);
        return (KT)(hi >> (64 - 2 * k));
    }
    // 128 bits starting at bit b of the stream w[0] w[1] .. w[RECW-1] 0 0
    const uint32_t wi = b >> 6, sh = b & 63u;
    const uint64_t w2 = RECW > 2 ? w[RECW > 2 ? 2 : 0] : 0ull, w3 = RECW > 3 ? w[RECW > 3 ? 3 : 0] : 0ull;
    const uint64_t a0 = wi == 0 ? w[0] : wi == 1 ? w[1] : wi == 2 ? w2 : w3;
    const uint64_t a1 = wi == 0 ? w[1] : wi == 1 ? w2 : wi == 2 ? w3 : 0ull;
    const uint64_t a2 = wi == 0 ? w2 : wi == 1 ? w3 : 0ull;
    const uint64_t hi = (a0 << sh) | ((a1 >> 1) >> (63u - sh));
    const uint64_t lo = (a1 << sh) | ((a2 >> 1) >> (63u - sh));
    const u128 v = (((u128)hi << 64) | lo) >> (128 - 2 * k);
    return (KT)v;
}

// Hashing for the counting tables.  Every 32-bit word goes through a 32 x 32 -> 64-bit multiply whose halves are folded
// together ("mum"): the high half carries the word's top bits down, so -- unlike a plain (word * odd) ^ ... chain,
// where a difference in the top bits of one word can cancel a difference in the top bits of another -- related k-mers
// (a substitution here, another 16 bases further on) do not collide systematically.  Only speed depends on the
// quality of these hashes: equal hashes are always confirmed against the full record / key before anything is counted.
RFX_HD uint32_t mum32(uint32_t x, uint32_t c) {
    const uint64_t p = (uint64_t)x * c;
    return (uint32_t)p ^ (uint32_t)(p >> 32);
}
RFX_HD uint32_t hash_lane_a(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    uint32_t x = mum32(a ^ 0x9E3779B9u, 0x9E3779B1u) ^ b;
    x = mum32(x, 0x85EBCA77u) ^ c;
    x = mum32(x, 0xC2B2AE3Du) ^ d;
    return mum32(x, 0x27D4EB2Fu);
}
RFX_HD uint32_t hash_lane_b(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    uint32_t x = mum32(a ^ 0x7F4A7C15u, 0x2C1B3C6Du) ^ b;
    x = mum32(x, 0x7FEB352Du) ^ c;
    x = mum32(x, 0x846CA68Bu) ^ d;
    return mum32(x, 0x165667B1u);
}
// k <= 31: table slot from the result; sub-class bits (24) from a second product of the same mixed word
RFX_HD uint32_t narrow_hash(uint64_t key, uint32_t& cls) {
    const uint32_t x = mum32((uint32_t)key ^ 0x9E3779B9u, 0x9E3779B1u) ^ (uint32_t)(key >> 32);
    const uint64_t p = (uint64_t)x * 0x85EBCA77u;
    cls = ((uint32_t)(p >> 32) * 0x27D4EB2Fu) >> 8;
    return (uint32_t)p ^ (uint32_t)(p >> 32);
}

// Bin of a record = bin of its first k-mer (every k-mer of a record shares it).  Used by the receiving
// side of a sharded run to re-group records that arrive as per-sender slices.
template <int RECW> RFX_HD uint32_t rec_first_bin(const uint64_t* rec, const BinParams& P) {
    uint64_t r[RECW];
#pragma unroll
    for (int i = 0; i < RECW; i++) r[i] = rec[i];
    const int m = P.m;
    const uint32_t mmask = (m >= 16) ? 0xffffffffu : ((1u << (2 * m)) - 1u);
    const int mtop = 2 * (m - 1);
    uint32_t mf = 0, mr = 0, hmin = 0xffffffffu;
    uint64_t cur = r[0] << 16;
    uint32_t avail = 24;
    for (int t = 0; t < P.k; t++) {
        if (avail == 0) {
            cur = r[1];
#pragma unroll
            for (int i = 1; i + 1 < RECW; i++) r[i] = r[i + 1];
            avail = 32;
        }
        const uint32_t v = (uint32_t)(cur >> 62);
        cur <<= 2; avail--;
        mf = ((mf << 2) | v) & mmask;
        mr = (mr >> 2) | ((v ^ 3u) << mtop);
        if (t + 1 >= m) {
            const uint32_t h = mmer_hash(mf < mr ? mf : mr);
            hmin = h < hmin ? h : hmin;
        }
    }
    return bin_of_minimizer(hmin, P.n_bins);
}

// ---------------------------------------------------------------------------------------------
// fork filters (A7 / A8), evaluated per (k-1)-mer group from the four candidate neighbours.
// ---------------------------------------------------------------------------------------------
struct ForkResult {
    int winner;    // base (0..3) of the surviving candidate, -1 if the group is empty
    int32_t flag;  // -1-coverage (clean / error fork) or k-1 (real fork won)
};

// Right fork filter over the k-mers that share a (k-1)-prefix; cnt[b] = coverage of prefix+b (0 = absent),
// dup[b] = the k-mer is its own reverse complement, so the reference holds the row twice
// (DSKmerReverseComplementLong, ReflexivDSMain.java:3864-3865).  Rows are visited in ascending
// last-base order (canonical resolution of the arrival-order dependence).
//   E != 0: DSFilterForkSubKmerWithErrorCorrection, ReflexivDSMain.java:3431-3483
//   E == 0: DSFilterForkSubKmer,                    ReflexivDSMain.java:3375-3417
RFX_HD ForkResult right_fork(const uint32_t cnt[4], const bool dup[4], int E, int sub) {
    ForkResult r; r.winner = -1; r.flag = 0;
    int32_t cc = 0;
    for (int b = 0; b < 4; b++) {
        if (!cnt[b]) continue;
        const int reps = dup[b] ? 2 : 1;
        for (int rep = 0; rep < reps; rep++) {
            const int32_t cx = (int32_t)cnt[b];
            if (r.winner < 0) { r.winner = b; cc = cx; r.flag = E ? -1 - cx : -1; }
            else if (cx > cc) {
                r.flag = (E && cc <= E && cx >= 2 * cc) ? -1 - cx : sub;
                r.winner = b; cc = cx;
            } else if (cx == cc) {
                if (b > r.winner) r.winner = b;
                r.flag = sub;
            } else {
                r.flag = (E && cx <= E && cc >= 2 * cx) ? -1 - cc : sub;
            }
        }
    }
    return r;
}

// Left fork filter over the survivors of the right filter that share a (k-1)-suffix; rows visited in
// ascending first-base order.  On equal coverage the arriving row always wins
// (ReflexivDSMain.java:3573-3578 compares 4|base with 1); in the "loser is an error" branch the stored
// row keeps whatever flag it had (ReflexivDSMain.java:3591-3596).
//   E != 0: DSFilterForkReflectedSubKmerWithErrorCorrection, ReflexivDSMain.java:3550-3616
//   E == 0: DSFilterForkReflectedSubKmer,                    ReflexivDSMain.java:3489-3541
RFX_HD ForkResult left_fork(const uint32_t cnt[4], int E, int sub) {
    ForkResult r; r.winner = -1; r.flag = 0;
    int32_t H = 0;
    for (int b = 0; b < 4; b++) {
        if (!cnt[b]) continue;
        const int32_t cx = (int32_t)cnt[b];
        if (r.winner < 0) { r.winner = b; H = cx; r.flag = E ? -1 - cx : -1; }
        else if (cx > H) {
            r.flag = (E && H <= E && cx >= 2 * H) ? -1 - cx : sub;
            H = cx; r.winner = b;
        } else if (cx == H) {
            r.winner = b; r.flag = sub;
        } else {
            if (!(E && cx <= E && H >= 2 * cx)) r.flag = sub;
        }
    }
    return r;
}

// ---------------------------------------------------------------------------------------------
// Fork filters of the "left and right sorting" stage that writes Count_<k>_sorted
// (ReflexivDSKmerLeftAndRightSorting.java:105-243).  Same scan as A7 / A8, other flags:
// a clean end is -1, a fork winner is X = (largest k of the k-mer list) + 3, coverages saturate at
// 30000 (buildingAlongFromThreeInt, :596-622), an error needs `fold` times the coverage
// (param.minRepeatFold) and coverage 1 is always an error.  Two quirks of the reference are kept:
// when a weaker, non-error row arrives the stored row takes over the ARRIVING row's coverage in the
// right filter (:504-518: the attribute is built before the stored row is re-read) and the arriving
// row's right flag in the left filter (:790-806).  Candidates are visited in ascending base order,
// the same canonical resolution of Spark's arrival order as above.
// ---------------------------------------------------------------------------------------------
struct SortedFork {
    int winner;           // base (0..3) of the surviving candidate, -1 if the group is empty
    int32_t left, right;  // right filter: coverage carried on, right flag; left filter: left flag, right flag
};
RFX_HD int32_t sorted_cov(uint32_t count) { return count >= 30000u ? 30000 : (int32_t)count; }

// DSFilterForkSubKmerWithErrorCorrection, ReflexivDSKmerLeftAndRightSorting.java:432-537
RFX_HD SortedFork sorted_right_fork(const uint32_t cnt[4], const bool dup[4], int E, double fold, int X) {
    SortedFork r; r.winner = -1; r.left = 0; r.right = 0;
    for (int b = 0; b < 4; b++) {
        if (!cnt[b]) continue;
        const int reps = dup[b] ? 2 : 1;
        for (int rep = 0; rep < reps; rep++) {
            const int32_t c = sorted_cov(cnt[b]);
            if (r.winner < 0) { r.winner = b; r.left = c; r.right = -1; continue; }
            const int32_t h = r.left;
            if (c > h) {
                r.right = (h <= E && (double)c >= fold * (double)h) ? -1 : (h == 1 ? -1 : X);
                r.winner = b; r.left = c;
            } else if (c == h) {
                if (b > r.winner) r.winner = b;
                r.right = h == 1 ? -1 : X;
            } else if (c <= E && (double)h >= fold * (double)c) {
                r.right = -1;
            } else {
                r.left = c;
                r.right = c == 1 ? -1 : X;
            }
        }
    }
    return r;
}

// DSFilterForkReflectedSubKmerWithErrorCorrection, ReflexivDSKmerLeftAndRightSorting.java:700-818.
// cov[a] = coverage the right filter left on the candidate with first base a (0 = absent), rfl[a] = its right flag.
RFX_HD SortedFork sorted_left_fork(const int32_t cov[4], const int32_t rfl[4], int E, double fold, int X) {
    SortedFork r; r.winner = -1; r.left = 0; r.right = 0;
    int32_t H = 0;  // HighCoverLastCoverage
    for (int a = 0; a < 4; a++) {
        if (!cov[a]) continue;
        const int32_t c = cov[a];
        if (r.winner < 0) { r.winner = a; H = c; r.left = -1; r.right = rfl[a]; continue; }
        if (c > H) {
            r.left = (H <= E && (double)c >= fold * (double)H) ? -1 : X;
            r.right = rfl[a]; r.winner = a; H = c;
        } else if (c == H) {
            r.winner = a; r.left = X;  // the larger first base wins (:746-756), which is the arriving one
            r.right = H == 1 ? -1 : rfl[a];
        } else if (c <= E && (double)H >= fold * (double)c) {
            r.left = -1;
        } else {
            r.left = X;
            r.right = c == 1 ? -1 : rfl[a];
        }
    }
    return r;
}

// A junction X -> Y is merged when the reflected record's right flag and the forward record's left
// flag are both negative or both non-negative (ReflexivDSMain.java:3069-3075, 1809-1815).  Mixed-sign
// ("budget") junctions stay open: canonical resolution, DESIGN.md.
RFX_HD bool junction_joins(int32_t x_right, int32_t y_left) { return (x_right < 0) == (y_left < 0); }

}  // namespace rfx
