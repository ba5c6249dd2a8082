// rfx_count.cu -- K3 + K4: per-bin k-mer counting in shared memory, coverage-filter compaction.
//
// Replaces groupBy("value").count() + filter(count >= min && count <= max)
//   (ReflexivDataFrameCounter.java:198-210, ReflexivDataFrameCounter64.java:200-212, ReflexivDSMain.java:207-216).
//
// One CTA owns one minimiser bin at a time.  Its threads stream the bin's super-k-mer records from
// HBM (one coalesced 16/32-byte record per thread), unroll them into canonical k-mers with the same
// rolling update as the reference extractor, and count them in an open-addressing table that lives
// in shared memory (64-bit atomicCAS claims a slot, 32-bit atomicAdd counts).  Only rows that pass the
// coverage filter ever reach HBM again.  A bin whose distinct k-mers do not fit the table is re-run
// in 2, 4, ... sub-classes selected by independent hash bits, so the result is exact for any input.
//
// k > 31 (128-bit keys): the slot is claimed with a 32-bit CAS on a tag word, the 128-bit key is then
// published, and a second pass over the records verifies every k-mer against the published keys
// while counting, so two different k-mers can never be merged (details at insert_wide()).
#include <stdlib.h>
#include <string.h>

#include "rfx_internal.h"

namespace rfx {

constexpr int CNT_THREADS = 256;
constexpr int CNT_STACK = 48;

struct CountArgs {
    const uint64_t* records;
    // bin b = concatenation over segments s of records[seg_base[s] + off_s[b] - off_s[0] .. seg_base[s] + off_s[b+1] - off_s[0])
    // with off_s = seg_off + s * (n_bins + 1).  A local partition is one segment with base 0.
    const uint64_t* seg_off;
    const uint64_t* seg_base;
    int n_seg;
    uint32_t n_bins;
    int k;
    uint32_t min_count, max_count;
    void* out_keys;
    uint32_t* out_counts;
    unsigned long long* dstat;
    unsigned long long out_cap;
};

// ---------------------------------------------------------------------------------------------
// k <= 31
// ---------------------------------------------------------------------------------------------
// Cheap 2 x 32-bit mix of a <= 62-bit key: slot bits from `h`, sub-class bits from an independent second product.
__device__ __forceinline__ uint32_t narrow_hash(uint64_t key) {
    uint32_t h = (uint32_t)key * 0x9E3779B1u ^ (uint32_t)(key >> 32) * 0x85EBCA77u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    return h;
}
__device__ __forceinline__ uint32_t narrow_class(uint64_t key, uint32_t h) { return ((h ^ (uint32_t)(key >> 32)) * 0x27D4EB2Fu) >> 8; }

// 64-bit hash of a whole 16-byte record, built from 32-bit multiplies.  Only speed depends on its quality: equal
// hashes are confirmed against the stored record before anything is counted.
__device__ __forceinline__ uint64_t record_hash(uint64_t w0, uint64_t w1) {
    const uint32_t a = (uint32_t)w0, b = (uint32_t)(w0 >> 32), c = (uint32_t)w1, d = (uint32_t)(w1 >> 32);
    uint32_t h1 = a * 0x9E3779B1u ^ b * 0x85EBCA77u ^ c * 0xC2B2AE3Du ^ d * 0x27D4EB2Fu;
    uint32_t h2 = a * 0x165667B1u ^ b * 0xD3A2646Cu ^ c * 0xFD7046C5u ^ d * 0xB55A4F09u;
    h1 ^= h1 >> 15; h1 *= 0x2C1B3C6Du; h1 ^= h1 >> 13;
    h2 ^= h2 >> 16; h2 *= 0x7FEB352Du; h2 ^= h2 >> 15;
    const uint64_t h = ((uint64_t)h2 << 32) | h1;
    return h == ~0ull ? 0x5bd1e9955bd1e995ull : h;
}

// adds `mult` occurrences of one canonical k-mer to the shared-memory table
template <int CAP>
__device__ __forceinline__ void insert_narrow(unsigned long long* keys, uint32_t* cnts, uint32_t* n_distinct, volatile uint32_t* overflow,
                                              uint64_t key, uint32_t h, uint32_t mult) {
    uint32_t slot = h & (CAP - 1);
    for (int probe = 0; probe < CAP; probe++) {
        unsigned long long cur = keys[slot];
        if (cur != key) {
            if (cur != ~0ull) { slot = (slot + 1) & (CAP - 1); continue; }
            cur = atomicCAS(&keys[slot], ~0ull, (unsigned long long)key);
            if (cur == ~0ull) {
                if (atomicAdd(n_distinct, 1u) >= (uint32_t)(CAP * 3 / 4)) *overflow = 1u;
            } else if (cur != key) { slot = (slot + 1) & (CAP - 1); continue; }
        }
        atomicAdd(&cnts[slot], mult);
        return;
    }
    *overflow = 1u;
}

struct NarrowTable {
    unsigned long long* keys;
    uint32_t* cnts;
    uint32_t* n_distinct;
    volatile uint32_t* overflow;
    uint32_t depth, cval;
};

template <int CAP> struct DirectInsert {  // rare path: a record whose 64-bit hash collided with a different record
    NarrowTable T;
    RFX_HD void operator()(uint64_t key) const {
#if defined(__CUDA_ARCH__)
        const uint32_t h = narrow_hash(key);
        if (T.depth == 0 || (narrow_class(key, h) >> (24u - T.depth)) == T.cval) insert_narrow<CAP>(T.keys, T.cnts, T.n_distinct, T.overflow, key, h, 1u);
#else
        (void)key;
#endif
    }
};

// One CTA per minimiser bin.  The bin's records are taken in chunks of RCAP*3/4:
//   A1  every thread hashes its records whole (16 bytes) into a shared-memory record table: identical super-k-mers --
//       the same genome window seen by many reads -- collapse into one entry with a multiplicity;
//   A2  every thread checks that the entry it counted into really holds its record (a 64-bit hash collision sends the
//       record down a direct path instead), so the collapse is exact;
//   B   the distinct records are expanded: a warp takes 32 table slots, prefix-sums their k-mer counts and walks the
//       flattened (record, k-mer) space 32 k-mers per step -- 5-step shuffle binary search for the source record, record
//       words fetched by shuffle, the k-mer cut straight out of the 2-bit stream, brev-based reverse complement -- and
//       adds the record's multiplicity to the canonical k-mer's slot of the k-mer table (64-bit atomicCAS + atomicAdd).
// At 100x coverage phase B sees about one k-mer in eight; the rest of the instances cost one record-level insert per
// ~10 k-mers.  K4: only rows inside the coverage bounds leave shared memory (block-scan compaction).
template <int RECW, int CAP, int RCAP, int NT>
__global__ void __launch_bounds__(NT) count_bins_narrow_kernel(CountArgs A) {
    static_assert(RECW == 2, "k <= 31 uses 16-byte records");
    constexpr int CHUNK = RCAP * 3 / 4;
    static_assert(CHUNK % NT == 0, "chunk must be a whole number of records per thread");
    constexpr int PER_THREAD = CHUNK / NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
    ulonglong2* rrec = reinterpret_cast<ulonglong2*>(keys + CAP);
    unsigned long long* rhash = reinterpret_cast<unsigned long long*>(rrec + RCAP);
    uint32_t* cnts = reinterpret_cast<uint32_t*>(rhash + RCAP);
    uint32_t* rmult = cnts + CAP;
    uint16_t* ulist = reinterpret_cast<uint16_t*>(rmult + RCAP);  // slots of the distinct records of the chunk, in claim order
    __shared__ uint32_t s_distinct, s_overflow, s_sp, s_nuniq;
    __shared__ uint32_t s_stack_val[CNT_STACK], s_stack_depth[CNT_STACK];
    __shared__ uint32_t s_warp_tot[NT / 32];
    __shared__ unsigned long long s_out_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = A.k;
    const int kshift = 64 - 2 * k;

    for (uint32_t bin = blockIdx.x; bin < A.n_bins; bin += gridDim.x) {
        __syncthreads();  // every thread has left the previous bin's class loop before the stack is re-armed
        if (tid == 0) { s_sp = 1; s_stack_val[0] = 0; s_stack_depth[0] = 0; }
        __syncthreads();
        while (true) {
            __syncthreads();
            if (s_sp == 0) break;
            const uint32_t depth = s_stack_depth[s_sp - 1], cval = s_stack_val[s_sp - 1];
            __syncthreads();
            if (tid == 0) { s_sp--; s_distinct = 0; s_overflow = 0; }
            for (int i = tid; i < CAP; i += NT) { keys[i] = ~0ull; cnts[i] = 0; }
            const uint32_t cshift = 24u - depth;  // class = top `depth` bits of a 24-bit second hash
            const NarrowTable T{keys, cnts, &s_distinct, &s_overflow, depth, cval};
            for (int seg = 0; seg < A.n_seg; seg++) {
            const uint64_t* so = A.seg_off + (size_t)seg * (A.n_bins + 1);
            const uint64_t beg = A.seg_base[seg] + so[bin] - so[0], end = A.seg_base[seg] + so[bin + 1] - so[0];
            for (uint64_t cbeg = beg; cbeg < end; cbeg += CHUNK) {
                for (int i = tid; i < RCAP; i += NT) { rhash[i] = ~0ull; rmult[i] = 0; }
                if (tid == 0) s_nuniq = 0;
                __syncthreads();
                if (s_overflow) break;
                const uint64_t cend = cbeg + CHUNK < end ? cbeg + CHUNK : end;
                // ---- A1: collapse identical records ----
                uint32_t myslot[PER_THREAD];
#pragma unroll
                for (int j = 0; j < PER_THREAD; j++) {
                    const uint64_t r = cbeg + (uint64_t)j * NT + tid;
                    myslot[j] = 0xffffffffu;
                    if (r < cend) {
                        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(A.records + r * RECW);
                        const uint64_t h = record_hash(v.x, v.y);
                        uint32_t slot = (uint32_t)(h >> 20) & (RCAP - 1);
                        while (true) {  // at most CHUNK < RCAP entries: an empty slot always exists
                            unsigned long long cur = rhash[slot];
                            if (cur == ~0ull) {
                                cur = atomicCAS(&rhash[slot], ~0ull, (unsigned long long)h);
                                if (cur == ~0ull) { rrec[slot] = v; ulist[atomicAdd(&s_nuniq, 1u)] = (uint16_t)slot; break; }
                            }
                            if (cur == h) break;
                            slot = (slot + 1) & (RCAP - 1);
                        }
                        atomicAdd(&rmult[slot], 1u);
                        myslot[j] = slot;
                    }
                }
                __syncthreads();
                // ---- A2: confirm (records are re-read: L1/L2 hits) ----
#pragma unroll
                for (int j = 0; j < PER_THREAD; j++) {
                    if (myslot[j] != 0xffffffffu) {
                        const uint64_t r = cbeg + (uint64_t)j * NT + tid;
                        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(A.records + r * RECW);
                        const ulonglong2 t = rrec[myslot[j]];
                        if (t.x != v.x || t.y != v.y) {
                            atomicSub(&rmult[myslot[j]], 1u);
                            uint64_t rec[RECW] = {v.x, v.y};
                            rec_foreach_kmer<uint64_t, RECW>(rec, k, DirectInsert<CAP>{T});
                        }
                    }
                }
                __syncthreads();
                // ---- B: expand the distinct records ----
                const int n_uniq = (int)s_nuniq;
                for (int ubase = warp * 32; ubase < n_uniq; ubase += NT) {
                    if (__any_sync(0xffffffffu, *(volatile uint32_t*)&s_overflow)) break;  // warp-uniform: shuffles follow
                    const int u = ubase + lane;
                    uint32_t mult = 0;
                    ulonglong2 v = make_ulonglong2(0ull, 0ull);
                    if (u < n_uniq) { const int slot = ulist[u]; mult = rmult[slot]; if (mult) v = rrec[slot]; }
                    const uint32_t nk = mult ? (uint32_t)(v.x >> 48) : 0u;
                    uint32_t pi = nk;  // inclusive prefix of k-mer counts over the warp's 32 table slots
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, pi, d); if (lane >= d) pi += o; }
                    const uint32_t total = __shfl_sync(0xffffffffu, pi, 31);
                    for (uint32_t t0 = 0; t0 < total; t0 += 32) {
                        const uint32_t t = t0 + lane;
                        uint32_t s = 0;  // number of slots whose inclusive prefix is <= t  ==  source slot of k-mer t
#pragma unroll
                        for (int step = 16; step; step >>= 1) {
                            const uint32_t pv = __shfl_sync(0xffffffffu, pi, (s + step - 1) & 31);
                            if (pv <= t) s += step;
                        }
                        s &= 31;
                        const uint64_t w0 = __shfl_sync(0xffffffffu, v.x, s), w1 = __shfl_sync(0xffffffffu, v.y, s);
                        const uint32_t pis = __shfl_sync(0xffffffffu, pi, s);
                        const uint32_t ms = __shfl_sync(0xffffffffu, mult, s);
                        if (t < total) {
                            const uint32_t off = t - (pis - (uint32_t)(w0 >> 48));  // k-mer index inside the record
                            const uint32_t b = 16u + 2u * off;                      // first bit of the k-mer in the 128-bit stream
                            uint64_t hi;
                            if (b < 64u) hi = (w0 << b) | (w1 >> (64u - b));
                            else hi = w1 << (b - 64u);
                            const uint64_t fwd = hi >> kshift;
                            const uint64_t rc = revcomp(fwd, k);
                            const uint64_t key = fwd < rc ? fwd : rc;
                            const uint32_t h = narrow_hash(key);
                            if (depth == 0 || (narrow_class(key, h) >> cshift) == cval)
                                insert_narrow<CAP>(keys, cnts, &s_distinct, &s_overflow, key, h, ms);
                        }
                    }
                }
                __syncthreads();
            }
            }
            __syncthreads();
            if (s_overflow) {
                // too many distinct k-mers for the table: split this class in two by the next hash bit
                if (tid == 0) {
                    if (depth >= 22 || s_sp + 2 > CNT_STACK) {
                        atomicExch(&A.dstat[DS_OVERFLOW], 2ull);
                    } else {
                        s_stack_val[s_sp] = cval << 1; s_stack_depth[s_sp] = depth + 1; s_sp++;
                        s_stack_val[s_sp] = (cval << 1) | 1u; s_stack_depth[s_sp] = depth + 1; s_sp++;
                        atomicAdd(&A.dstat[DS_SPLITS], 1ull);
                    }
                }
                continue;
            }
            // K4: coverage filter + compaction.  Block-wide exclusive scan of per-thread survivor counts.
            uint32_t mine = 0, inst = 0;
            for (int i = tid; i < CAP; i += NT) {
                const uint32_t cn = cnts[i];
                inst += cn;
                mine += (cn >= A.min_count && cn <= A.max_count && keys[i] != ~0ull) ? 1u : 0u;
            }
            uint32_t incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) inst += __shfl_xor_sync(0xffffffffu, inst, d);
            if (lane == 31) s_warp_tot[warp] = incl;
            __syncthreads();
            if (tid == 0) {
                uint32_t tot = 0;
                for (int w = 0; w < NT / 32; w++) { uint32_t t = s_warp_tot[w]; s_warp_tot[w] = tot; tot += t; }
                s_out_base = tot ? atomicAdd(&A.dstat[DS_OUT_CURSOR], (unsigned long long)tot) : 0ull;
                atomicAdd(&A.dstat[DS_DISTINCT], (unsigned long long)s_distinct);
                if (s_out_base + tot > A.out_cap) atomicExch(&A.dstat[DS_OVERFLOW], 1ull);
            }
            if (lane == 0 && inst) atomicAdd(&A.dstat[DS_INSTANCES], (unsigned long long)inst);
            __syncthreads();
            unsigned long long o = s_out_base + s_warp_tot[warp] + (incl - mine);
            if (o + mine <= A.out_cap) {
                uint64_t* ok = reinterpret_cast<uint64_t*>(A.out_keys);
                for (int i = tid; i < CAP; i += NT) {
                    const uint32_t cn = cnts[i];
                    if (cn >= A.min_count && cn <= A.max_count && keys[i] != ~0ull) { ok[o] = keys[i]; A.out_counts[o] = cn; o++; }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k in 32..63: 128-bit keys
//   tags[slot]  : 0 = empty, else 1 | (31-bit fingerprint << 1)     (claimed with a 32-bit CAS)
//   keys[slot]  : the 128-bit key, written by the claiming thread right after the CAS
// Pass A inserts (claims + publishes keys; a thread that meets an equal tag assumes "same key").
// After a barrier every key is visible.  Pass B walks the records again and, for every k-mer, probes
// until it finds the slot whose full key matches, then counts.  If pass A merged two different keys
// that share slot sequence and fingerprint, pass B does not find the second key: it flags overflow
// and the class is split (different class hash bits separate the pair), so counts stay exact.
// ---------------------------------------------------------------------------------------------
template <int CAP> struct InsertWide {
    uint32_t* tags;
    u128* keys;
    uint32_t* n_distinct;
    volatile uint32_t* overflow;
    uint32_t cls_mask, cls_val;
    RFX_HD void operator()(u128 key) const {
#if defined(__CUDA_ARCH__)
        const uint64_t h = key_hash(key);
        if (((uint32_t)(h >> 43) & cls_mask) != cls_val) return;  // class = hash bits 43..63
        const uint32_t tag = 1u | ((uint32_t)(h >> 12) << 1);  // fingerprint = hash bits 12..42
        uint32_t slot = (uint32_t)h & (CAP - 1);
        for (int probe = 0; probe < CAP; probe++) {
            uint32_t cur = tags[slot];
            if (cur == 0u) cur = atomicCAS(&tags[slot], 0u, tag);
            if (cur == 0u) {
                keys[slot] = key;
                if (atomicAdd(n_distinct, 1u) >= (uint32_t)(CAP * 3 / 4)) *overflow = 1u;
                return;
            }
            if (cur == tag) return;  // presumed equal; verified in pass B
            slot = (slot + 1) & (CAP - 1);
        }
        *overflow = 1u;
#else
        (void)key;
#endif
    }
};

template <int CAP> struct CountWide {
    const uint32_t* tags;
    const u128* keys;
    uint32_t* cnts;
    volatile uint32_t* overflow;
    uint32_t cls_mask, cls_val;
    RFX_HD void operator()(u128 key) const {
#if defined(__CUDA_ARCH__)
        const uint64_t h = key_hash(key);
        if (((uint32_t)(h >> 43) & cls_mask) != cls_val) return;  // class = hash bits 43..63
        uint32_t slot = (uint32_t)h & (CAP - 1);
        for (int probe = 0; probe < CAP; probe++) {
            if (tags[slot] == 0u) break;
            if (keys[slot] == key) { atomicAdd(&cnts[slot], 1u); return; }
            slot = (slot + 1) & (CAP - 1);
        }
        *overflow = 1u;  // key was swallowed by a fingerprint collision in pass A
#else
        (void)key;
#endif
    }
};

template <int RECW, int CAP>
__global__ void __launch_bounds__(CNT_THREADS) count_bins_wide_kernel(CountArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u128* keys = reinterpret_cast<u128*>(smem_raw);
    uint32_t* tags = reinterpret_cast<uint32_t*>(keys + CAP);
    uint32_t* cnts = tags + CAP;
    __shared__ uint32_t s_distinct, s_overflow, s_sp;
    __shared__ uint32_t s_stack_val[CNT_STACK], s_stack_depth[CNT_STACK];
    __shared__ uint32_t s_warp_tot[CNT_THREADS / 32];
    __shared__ unsigned long long s_out_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (uint32_t bin = blockIdx.x; bin < A.n_bins; bin += gridDim.x) {
        __syncthreads();  // every thread has left the previous bin's class loop before the stack is re-armed
        if (tid == 0) { s_sp = 1; s_stack_val[0] = 0; s_stack_depth[0] = 0; }
        __syncthreads();
        while (true) {
            __syncthreads();
            if (s_sp == 0) break;
            const uint32_t depth = s_stack_depth[s_sp - 1], cval = s_stack_val[s_sp - 1];
            __syncthreads();
            if (tid == 0) { s_sp--; s_distinct = 0; s_overflow = 0; }
            for (int i = tid; i < CAP; i += CNT_THREADS) { tags[i] = 0u; cnts[i] = 0u; }
            __syncthreads();
            const uint32_t cmask = (1u << depth) - 1u;
            InsertWide<CAP> ins{tags, keys, &s_distinct, &s_overflow, cmask, cval};
            for (int seg = 0; seg < A.n_seg; seg++) {
            const uint64_t* so = A.seg_off + (size_t)seg * (A.n_bins + 1);
            const uint64_t beg = A.seg_base[seg] + so[bin] - so[0], end = A.seg_base[seg] + so[bin + 1] - so[0];
            for (uint64_t r = beg + tid; r < end; r += CNT_THREADS) {
                if (*(volatile uint32_t*)&s_overflow) break;
                uint64_t rec[RECW];
#pragma unroll
                for (int i = 0; i < RECW; i += 2) {
                    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(A.records + r * RECW + i);
                    rec[i] = v.x; rec[i + 1] = v.y;
                }
                rec_foreach_kmer<u128, RECW>(rec, A.k, ins);
            }
            }
            __syncthreads();
            if (!s_overflow) {
                CountWide<CAP> cnt{tags, keys, cnts, &s_overflow, cmask, cval};
                for (int seg = 0; seg < A.n_seg; seg++) {
                const uint64_t* so = A.seg_off + (size_t)seg * (A.n_bins + 1);
                const uint64_t beg = A.seg_base[seg] + so[bin] - so[0], end = A.seg_base[seg] + so[bin + 1] - so[0];
                for (uint64_t r = beg + tid; r < end; r += CNT_THREADS) {
                    if (*(volatile uint32_t*)&s_overflow) break;
                    uint64_t rec[RECW];
#pragma unroll
                    for (int i = 0; i < RECW; i += 2) {
                        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(A.records + r * RECW + i);
                        rec[i] = v.x; rec[i + 1] = v.y;
                    }
                    rec_foreach_kmer<u128, RECW>(rec, A.k, cnt);
                }
                }
                __syncthreads();
            }
            if (s_overflow) {
                if (tid == 0) {
                    if (depth >= 20 || s_sp + 2 > CNT_STACK) {
                        atomicExch(&A.dstat[DS_OVERFLOW], 2ull);
                    } else {
                        s_stack_val[s_sp] = cval; s_stack_depth[s_sp] = depth + 1; s_sp++;
                        s_stack_val[s_sp] = cval | (1u << depth); s_stack_depth[s_sp] = depth + 1; s_sp++;
                        atomicAdd(&A.dstat[DS_SPLITS], 1ull);
                    }
                }
                continue;
            }
            uint32_t mine = 0, inst = 0;
            for (int i = tid; i < CAP; i += CNT_THREADS) {
                const uint32_t cn = cnts[i];
                inst += cn;
                mine += (cn >= A.min_count && cn <= A.max_count && tags[i] != 0u) ? 1u : 0u;
            }
            uint32_t incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) inst += __shfl_xor_sync(0xffffffffu, inst, d);
            if (lane == 31) s_warp_tot[warp] = incl;
            __syncthreads();
            if (tid == 0) {
                uint32_t tot = 0;
                for (int w = 0; w < CNT_THREADS / 32; w++) { uint32_t t = s_warp_tot[w]; s_warp_tot[w] = tot; tot += t; }
                s_out_base = tot ? atomicAdd(&A.dstat[DS_OUT_CURSOR], (unsigned long long)tot) : 0ull;
                atomicAdd(&A.dstat[DS_DISTINCT], (unsigned long long)s_distinct);
                if (s_out_base + tot > A.out_cap) atomicExch(&A.dstat[DS_OVERFLOW], 1ull);
            }
            if (lane == 0 && inst) atomicAdd(&A.dstat[DS_INSTANCES], (unsigned long long)inst);
            __syncthreads();
            unsigned long long o = s_out_base + s_warp_tot[warp] + (incl - mine);
            if (o + mine <= A.out_cap) {
                u128* ok = reinterpret_cast<u128*>(A.out_keys);
                for (int i = tid; i < CAP; i += CNT_THREADS) {
                    const uint32_t cn = cnts[i];
                    if (cn >= A.min_count && cn <= A.max_count && tags[i] != 0u) { ok[o] = keys[i]; A.out_counts[o] = cn; o++; }
                }
            }
        }
    }
}

constexpr int CAP_NARROW = 2048;   // k-mer table: 2048 * (8 + 4) B = 24 KB
constexpr int RCAP_NARROW = 1024;  // record table: 1024 * (16 + 8 + 4 + 2) B = 30 KB   -> 54 KB / CTA, 4 CTAs / SM
constexpr int CAP_WIDE = 4096;    // 4096 * (16 + 4 + 4) B = 96 KB -> 2 CTAs / SM

int stage_count(Ctx* c) {
    cudaStream_t st = c->stream;
    if (!c->have_records) return ctx_fail(c, RFX_E_STATE, "rfx_count: no records (push reads first)");
    stage_begin(c);
    // effective coverage bounds (A4)
    uint32_t minc = (uint32_t)(c->prm.min_kmer_coverage < 0 ? 0 : c->prm.min_kmer_coverage);
    uint32_t maxc = (uint32_t)c->prm.max_kmer_coverage;
    if (c->prm.counter_mode) {
        if (c->prm.min_kmer_coverage <= 1) minc = 0;
        if (c->prm.max_kmer_coverage >= 10000000) maxc = 0xffffffffu;
    }
    if (minc < 1) minc = 1;
    // capacity of the filtered table
    uint64_t cap = c->prm.table_capacity > 0 ? (uint64_t)c->prm.table_capacity : 0;
    if (!cap) {
        // every surviving row needs >= minc instances; keep a floor for tiny inputs
        uint64_t inst = c->n_instances ? c->n_instances : c->n_records * c->max_nk;
        cap = inst / minc + 1024;
        const uint64_t soft = 1ull << 31;  // 2 G rows (24-40 GB): beyond this the caller must say so
        if (cap > soft) cap = soft;
    }
    const size_t ksz = c->wide ? sizeof(u128) : sizeof(uint64_t);
    RFX_TRY(devbuf_reserve(c, c->keys, cap * ksz));
    RFX_TRY(devbuf_reserve(c, c->counts, cap * sizeof(uint32_t)));
    c->table_cap = cap;
    RFX_CUDA(c, cudaMemsetAsync(c->dstat.p, 0, DS_NSLOTS * sizeof(uint64_t), st));
    CountArgs A;
    const bool segmented = c->shard_id >= 0 && c->n_seg > 0;
    if (!segmented) {  // one segment: the locally partitioned (or re-binned) records
        RFX_TRY(devbuf_reserve(c, c->seg_base, 64 * sizeof(uint64_t)));
        RFX_CUDA(c, cudaMemsetAsync(c->seg_base.p, 0, sizeof(uint64_t), st));
    }
    A.records = segmented ? c->rx_records.as<uint64_t>() : c->records.as<uint64_t>();
    A.seg_off = segmented ? c->seg_off.as<uint64_t>() : c->bin_off.as<uint64_t>();
    A.seg_base = c->seg_base.as<uint64_t>();
    A.n_seg = segmented ? c->n_seg : 1;
    A.n_bins = c->n_bins;
    A.k = c->k;
    A.min_count = minc; A.max_count = maxc;
    A.out_keys = c->keys.p; A.out_counts = c->counts.as<uint32_t>();
    A.dstat = c->dstat.as<unsigned long long>();
    A.out_cap = cap;
    if (c->n_records) {
        cudaEventRecord(c->evk[4], st);
        if (!c->wide) {
            const size_t smem = (size_t)CAP_NARROW * 12 + (size_t)RCAP_NARROW * 30;
            unsigned grid = c->n_bins < 148u * 4u * 8u ? c->n_bins : 148u * 4u * 8u;
            const char* variant = getenv("RFX_COUNT_VARIANT");  // tuning knob: "big" = 512 threads, 2048-entry record table
            if (variant && !strcmp(variant, "big")) {
                const size_t smem2 = (size_t)CAP_NARROW * 12 + (size_t)2048 * 30;
                RFX_CUDA(c, cudaFuncSetAttribute(count_bins_narrow_kernel<2, CAP_NARROW, 2048, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
                count_bins_narrow_kernel<2, CAP_NARROW, 2048, 512><<<grid, 512, smem2, st>>>(A);
            } else {
                RFX_CUDA(c, cudaFuncSetAttribute(count_bins_narrow_kernel<2, CAP_NARROW, RCAP_NARROW, CNT_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                count_bins_narrow_kernel<2, CAP_NARROW, RCAP_NARROW, CNT_THREADS><<<grid, CNT_THREADS, smem, st>>>(A);
            }
        } else {
            const size_t smem = (size_t)CAP_WIDE * 24;
            RFX_CUDA(c, cudaFuncSetAttribute(count_bins_wide_kernel<4, CAP_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            unsigned grid = c->n_bins < 148u * 2u * 8u ? c->n_bins : 148u * 2u * 8u;
            count_bins_wide_kernel<4, CAP_WIDE><<<grid, CNT_THREADS, smem, st>>>(A);
        }
        cudaEventRecord(c->evk[5], st);
        c->launches++;
    }
    uint64_t h[DS_NSLOTS];
    RFX_CUDA(c, cudaMemcpyAsync(h, c->dstat.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return ctx_fail(c, RFX_E_CUDA, "count kernel failed: %s", cudaGetErrorString(e));
    c->ms[2] += stage_end(c);
    c->ms_kernel[2] = 0;
    if (c->n_records) cudaEventElapsedTime(&c->ms_kernel[2], c->evk[4], c->evk[5]);
    if (h[DS_OVERFLOW] == 2) return ctx_fail(c, RFX_E_CAPACITY, "a counting bin could not be split further");
    if (h[DS_OVERFLOW] == 1 || h[DS_OUT_CURSOR] > cap)
        return ctx_fail(c, RFX_E_CAPACITY, "filtered table needs %llu rows, capacity %llu: raise table_capacity",
                        (unsigned long long)h[DS_OUT_CURSOR], (unsigned long long)cap);
    c->n_rows = h[DS_OUT_CURSOR];
    c->n_distinct = h[DS_DISTINCT];
    c->n_bin_splits = h[DS_SPLITS];
    if (c->n_instances == 0) c->n_instances = h[DS_INSTANCES];
    else if (h[DS_INSTANCES] != c->n_instances)
        return ctx_fail(c, RFX_E_STATE, "internal: counted %llu k-mer instances, extracted %llu", (unsigned long long)h[DS_INSTANCES],
                        (unsigned long long)c->n_instances);
    c->have_counts = true;
    c->have_contigs = false;
    return RFX_OK;
}

}  // namespace rfx
