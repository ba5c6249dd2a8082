// rfx_shard_graph.cuh -- the graph stages (K5, K6, K7) of a sharded run: every rank keeps ITS rows of the filtered table
// (the rows whose minimiser bin it owns, exactly as rfx_count_sharded left them), the index over them and all per-node
// state; whatever a node needs from another rank it reads -- or writes -- in that rank's HBM over NVLink / NVSwitch.
// Included at the end of rfx_graph.cu (same translation unit: Graph<KT>, the index build and the rule functions are shared).
//
// Replaces, across GPUs, the global sort("k-1") shuffles of ReflexivDSMain.java:232, 244 (fork filters) and :261-326 (one
// or two per pass of the extension loop).
//
//   ids        a node is  gid = rank << 29 | (2 * row + strand)  with `row` local to its rank; links, ancestors and
//              splitter ids all use that form, so following a chain across GPUs is following a pointer.
//   owner      of a k-mer = owner of the bin its canonical minimiser hashes to (rfx_core.h: bin_of_minimizer), the very
//              map the counting stage shards by.  A neighbour shares k-1 bases with the node that asks for it, so the
//              asking node knows the neighbour's minimiser -- and owner -- after hashing ONE more m-mer.
//   probes     every rank holds a copy of every rank's presence bits (16 bits per row, pulled once per run: 2 B per
//              row instead of the 12 B per row of the table).  Three of four probes ask for a k-mer that does not
//              exist and end there; about 90 % of the rest stay on the rank (same minimiser); what is left walks the
//              owner's index, keys and counts in place over NVLink (counted: rfx_shard_stats_t.n_remote_probes).
//   chains     level-1 splitters = heads, a 1-in-64 sample and every node whose predecessor lives on another rank, so a
//              level-1 segment never leaves its GPU and the walk that stamps it is purely local.  Level-2 splitters = heads
//              and 1 in 8 of the level-1 splitters; they walk the level-1 list (one 8-byte peer read and one 8-byte peer
//              write per hop).  Only the level-2 list (about 1 % of the nodes) goes through pointer jumping, each round
//              one small kernel + one cross-GPU barrier.
//   contigs    stay with the owner of their head: the tail tells the head its length and right flag, the owner sizes
//              and lays out its contigs, every node writes its base into the owner's buffer.
//   cycles     a closed path has no head: the ranks notice (a node without ancestor, or an ancestor that is no head)
//              and rank 0 pulls the shard tables and runs the single-GPU stages on the whole table (rare).
#pragma once

#include "rfx_shard.h"

namespace rfx {

constexpr uint32_t GID_MASK = 0x1fffffffu;
__host__ __device__ __forceinline__ uint32_t gid_make(int r, uint32_t l) { return ((uint32_t)r << 29) | l; }
__host__ __device__ __forceinline__ int gid_rank(uint32_t g) { return (int)(g >> 29); }
__host__ __device__ __forceinline__ uint32_t gid_loc(uint32_t g) { return g & GID_MASK; }

// what a rank needs to know of a rank (itself included); every address as seen from THIS device
struct SGPeer {
    const void* keys;
    const uint32_t* counts;
    const uint32_t* ht;
    const uint32_t* bloom;  // the LOCAL copy of that rank's presence bits
    uint8_t* alive;
    int32_t *lflag, *rflag, *eff_l, *eff_r;
    uint32_t *succ, *pred, *spl_id;
    uint64_t* l1_nl;   // per level-1 splitter: next level-1 splitter (gid form: rank, index) | nodes of the segment << 32
    uint64_t* l1_loc;  // per level-1 splitter: (level-2 splitter that walked over it, nodes in front of it)
    uint32_t* l2_of;   // per level-1 splitter: its level-2 index or NONE
    uint64_t* l2_up[2];
    uint32_t* l2_node;
    uint64_t* l2_fin;  // per level-2 splitter: (head node, nodes in front of it)
    uint32_t* chain_len;
    int32_t* tail_rf;
    uint32_t* ctg_idx;
    uint64_t* ctg_off;
    char* ctg_bases;
    uint64_t bloom_mask;
    uint32_t ht_cap, pad;
};
struct SGView {
    SGPeer p[RFX_MAX_RANKS];
    int me, world, k, m;
    uint32_t B, bps;
};

// slots of the published block (ShardCtl::pub, from PUB_GRAPH on): arena offsets first, then values
enum {
    GP_KEYS, GP_COUNTS, GP_HT, GP_BLOOM, GP_ALIVE, GP_LFLAG, GP_RFLAG, GP_EFFL, GP_EFFR, GP_SUCC, GP_PRED, GP_SPLID, GP_L1NL, GP_L1LOC, GP_L2OF, GP_L2UP0,
    GP_L2UP1, GP_L2NODE, GP_L2FIN, GP_CHAINLEN, GP_TAILRF, GP_CTGIDX, GP_CTGOFF, GP_CTGBASES, GP_NPTR,
    GP_NROWS = GP_NPTR, GP_HTCAP, GP_BLOOMMASK, GP_HOST,  // GP_HOST: 6 host values, then up to 6 device values
    GP_NHOST = 6, GP_DEV = GP_HOST + GP_NHOST, GP_NDEV = 6, GP_END = GP_DEV + GP_NDEV
};
static_assert(PUB_GRAPH + GP_END <= RFX_PUB_SLOTS, "published block too small");

struct GShard {
    DevBuf bloom_all, l1_nl, l1_loc, l1_fin, l1_dst, l2_of, l2_l1, l2_node, l2_up[2], l2_fin, tail_rf, pubsrc, all_keys, all_counts;
    const uint32_t* bloom_local[RFX_MAX_RANKS] = {nullptr};  // local copies of the peers' presence bits (this run)
    uint64_t n_remote = 0, n_l1 = 0, n_l2 = 0, n_rows_global = 0, n_oriented_global = 0, n_contigs_global = 0, n_bases_global = 0;
    int fell_back = 0;
};

// ---- neighbour lookup ------------------------------------------------------------------------------------------------
// gid of the oriented k-mer Z (minimiser hash hmin), NONE32 if its canonical form is in no rank's table
template <class KT> __device__ __forceinline__ uint32_t sg_find(const SGView& V, KT Z, uint32_t hmin, uint32_t* cnt, unsigned long long* dstat) {
    const KT zc = revcomp(Z, V.k);
    const bool fwd = !(zc < Z);
    const KT canon = fwd ? Z : zc;
    const int r = (int)(bin_of_minimizer(hmin, V.B) / V.bps);
    const SGPeer& P = V.p[r];
    const uint64_t kh = key_hash(canon);
    const uint64_t bit = (kh >> 13) & P.bloom_mask;
    if (!((P.bloom[bit >> 5] >> (bit & 31u)) & 1u)) return NONE32;
    if (r != V.me) atomicAdd(&dstat[DS_REMOTE], 1ull);
    const KT* keys = reinterpret_cast<const KT*>(P.keys);
    const uint32_t cap = P.ht_cap;
    uint32_t slot = (uint32_t)(((uint64_t)(uint32_t)(kh >> 20) * cap) >> 32);
    uint32_t v;
    while (true) {
        v = P.ht[slot];
        if (v == NONE32) return NONE32;
        if (keys[v] == canon) break;
        slot = slot + 1 == cap ? 0u : slot + 1;
    }
    *cnt = P.counts[v];
    return gid_make(r, 2u * v + (fwd ? 0u : 1u));
}
template <class KT> __device__ __forceinline__ KT sg_oriented(const SGView& V, uint32_t g) {
    const uint32_t l = gid_loc(g);
    const KT key = reinterpret_cast<const KT*>(V.p[gid_rank(g)].keys)[l >> 1];
    return (l & 1u) ? revcomp(key, V.k) : key;
}

// rows must sit on the rank that owns their minimiser bin, or neighbours would look for them elsewhere
template <class KT> __global__ void sg_check_owner_kernel(const __grid_constant__ SGView V, uint64_t n_rows, unsigned long long* dstat) {
    const KT* keys = reinterpret_cast<const KT*>(V.p[V.me].keys);
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t a, b;
        kmer_minima<KT>(keys[r], V.k, V.m, a, b);
        uint32_t h = a < b ? a : b;
        if (V.m == V.k) h = mm_hash_m((uint32_t)keys[r], V.m);  // one m-mer, in neither the prefix nor the suffix part
        if ((int)(bin_of_minimizer(h, V.B) / V.bps) != V.me) atomicExch(&dstat[DS_GRAPH_ERR], 3ull);
    }
}

// ---- K5: A7, A8, links (rules: rfx_core.h; the single-GPU kernels of rfx_graph.cu with peer-aware probes) ----------------
template <class KT> __global__ void sg_right_filter_kernel(const __grid_constant__ SGView V, int E, uint64_t n, unsigned long long* dstat) {
    const SGPeer& Me = V.p[V.me];
    const KT* keys = reinterpret_cast<const KT*>(Me.keys);
    for (uint64_t oid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; oid < n; oid += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t row = (uint32_t)(oid >> 1);
        const KT key = keys[row];
        const KT rc = revcomp(key, V.k);
        if ((oid & 1u) && rc == key) { Me.alive[oid] = 0; Me.rflag[oid] = 0; continue; }  // palindrome: one node, not two
        const KT X = (oid & 1u) ? rc : key;
        const KT prefix = X >> 2;
        const uint32_t myb = (uint32_t)X & 3u;
        uint32_t pre_min, suf_min;
        kmer_minima<KT>(X, V.k, V.m, pre_min, suf_min);
        uint32_t cnt[4];
        bool dup[4];
#pragma unroll
        for (uint32_t b = 0; b < 4; b++) {
            if (b == myb) { cnt[b] = Me.counts[row]; dup[b] = (rc == key); }
            else {
                const KT Z = (prefix << 2) | (KT)b;
                const uint32_t hl = last_mm_of<KT>(prefix, b, V.m);
                uint32_t cz = 0;
                const uint32_t g = sg_find<KT>(V, Z, hl < pre_min ? hl : pre_min, &cz, dstat);
                cnt[b] = g == NONE32 ? 0u : cz;
                dup[b] = (Z == revcomp(Z, V.k));
            }
        }
        const ForkResult res = right_fork(cnt, dup, E, V.k - 1);
        Me.alive[oid] = (uint8_t)(((res.winner == (int)myb) ? 1 : 0) | (res.flag < 0 ? 4 : 0));
        Me.rflag[oid] = res.flag;
    }
}

template <class KT> __global__ void sg_left_filter_kernel(const __grid_constant__ SGView V, int E, uint64_t n, unsigned long long* dstat) {
    const SGPeer& Me = V.p[V.me];
    const int top = 2 * (V.k - 1);
    const KT sufmask = mask_bases<KT>(V.k - 1);
    for (uint64_t oid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; oid < n; oid += (uint64_t)gridDim.x * blockDim.x) {
        Me.lflag[oid] = 0;
        if (!(Me.alive[oid] & 1)) continue;
        const KT X = sg_oriented<KT>(V, gid_make(V.me, (uint32_t)oid));
        const KT suffix = X & sufmask;
        const uint32_t mya = (uint32_t)(X >> top) & 3u;
        uint32_t pre_min, suf_min;
        kmer_minima<KT>(X, V.k, V.m, pre_min, suf_min);
        uint32_t cnt[4];
#pragma unroll
        for (uint32_t a = 0; a < 4; a++) {
            if (a == mya) cnt[a] = Me.counts[oid >> 1];
            else {
                uint32_t cz = 0;
                const uint32_t hf = first_mm_of<KT>(suffix, a, V.k, V.m);
                const uint32_t g = sg_find<KT>(V, ((KT)a << top) | suffix, hf < suf_min ? hf : suf_min, &cz, dstat);
                cnt[a] = (g != NONE32 && (V.p[gid_rank(g)].alive[gid_loc(g)] & 1)) ? cz : 0u;
            }
        }
        const ForkResult res = left_fork(cnt, E, V.k - 1);
        if (res.winner == (int)mya) { Me.alive[oid] = (uint8_t)((Me.alive[oid] & 4) | 3 | (res.flag < 0 ? 8 : 0)); Me.lflag[oid] = res.flag; }
    }
}

// Raw links over every junction (which of them hold is decided by the budget walks, as on one GPU).  The successor's
// pred[] may be in a peer's memory: one remote atomic per junction that crosses ranks.
template <class KT> __global__ void sg_link_kernel(const __grid_constant__ SGView V, uint64_t n, unsigned long long* dstat) {
    const SGPeer& Me = V.p[V.me];
    const KT sufmask = mask_bases<KT>(V.k - 1);
    for (uint64_t oid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; oid < n; oid += (uint64_t)gridDim.x * blockDim.x) {
        Me.eff_l[oid] = Me.lflag[oid];
        Me.eff_r[oid] = Me.rflag[oid];
        if (!(Me.alive[oid] & 2)) continue;
        const uint32_t self = gid_make(V.me, (uint32_t)oid);
        const KT X = sg_oriented<KT>(V, self);
        const KT suffix = X & sufmask;
        uint32_t pre_min, suf_min;
        kmer_minima<KT>(X, V.k, V.m, pre_min, suf_min);
        uint32_t next = NONE32;
        int n_cand = 0;
#pragma unroll
        for (uint32_t b = 0; b < 4; b++) {
            uint32_t cz;
            const uint32_t hl = last_mm_of<KT>(suffix, b, V.m);
            const uint32_t g = sg_find<KT>(V, (suffix << 2) | (KT)b, hl < suf_min ? hl : suf_min, &cz, dstat);
            if (g != NONE32 && (V.p[gid_rank(g)].alive[gid_loc(g)] & 2)) { next = g; n_cand++; }
        }
        if (n_cand > 1) { atomicExch(&dstat[DS_GRAPH_ERR], 1ull); continue; }
        if (Me.lflag[oid] >= 0 || Me.rflag[oid] >= 0) atomicAdd(&dstat[DS_FLAGGED], 1ull);
        if (next != NONE32) {
            const bool joins = junction_joins(Me.rflag[oid], V.p[gid_rank(next)].lflag[gid_loc(next)]);
            if (!joins) atomicAdd(&dstat[DS_BUDGET], 1ull);
            if (next == self) {
                if (joins) atomicAdd(&dstat[DS_CYCLES], 1ull);  // 1-cycle: a record never merges with itself
            } else {
                Me.succ[oid] = next;
                if (atomicExch(&V.p[gid_rank(next)].pred[gid_loc(next)], self) != NONE32) atomicExch(&dstat[DS_GRAPH_ERR], 2ull);
            }
        }
    }
}

// ---- budget walks across ranks (rfx_graph.cu: budget_walk_kernel, every array read through its owner) ----------------------
struct SGWalkArrays {  // DIR 0: bud = rflag, face = lflag, nxt = succ, prv = pred, eff = eff_r;  DIR 1: mirrored
    __device__ __forceinline__ static int32_t bud(const SGView& V, int DIR, uint32_t g) { return (DIR == 0 ? V.p[gid_rank(g)].rflag : V.p[gid_rank(g)].lflag)[gid_loc(g)]; }
    __device__ __forceinline__ static int32_t face(const SGView& V, int DIR, uint32_t g) { return (DIR == 0 ? V.p[gid_rank(g)].lflag : V.p[gid_rank(g)].rflag)[gid_loc(g)]; }
    __device__ __forceinline__ static uint32_t nxt(const SGView& V, int DIR, uint32_t g) { return (DIR == 0 ? V.p[gid_rank(g)].succ : V.p[gid_rank(g)].pred)[gid_loc(g)]; }
    __device__ __forceinline__ static uint32_t prv(const SGView& V, int DIR, uint32_t g) { return (DIR == 0 ? V.p[gid_rank(g)].pred : V.p[gid_rank(g)].succ)[gid_loc(g)]; }
};
template <class KT, int DIR> __global__ void sg_budget_walk_kernel(const __grid_constant__ SGView V, uint64_t n, int bmax, unsigned long long* dstat) {
    typedef SGWalkArrays W;
    const SGPeer& Me = V.p[V.me];
    for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (uint64_t)gridDim.x * blockDim.x) {
        if (!(Me.alive[x] & 2)) continue;
        const uint32_t s = gid_make(V.me, (uint32_t)x);
        if (W::bud(V, DIR, s) < 0) continue;
        if (W::nxt(V, DIR, s) == NONE32) continue;  // nothing downstream to absorb
        uint32_t stop = s;
        bool fresh = true;
        if (W::face(V, DIR, s) < 0 && W::prv(V, DIR, s) != NONE32) {
            uint32_t u = W::prv(V, DIR, s), start = NONE32;
            int gap = 0;
            while (true) {
                if (u == s) break;  // closed path without a fixed point: smallest fork winner starts
                gap = W::bud(V, DIR, u) < 0 ? gap + 1 : 0;
                if (W::face(V, DIR, u) >= 0 || W::prv(V, DIR, u) == NONE32 || gap >= bmax) { start = u; break; }
                u = W::prv(V, DIR, u);
            }
            if (start == NONE32) {
                uint32_t mn = s;
                KT mk = sg_oriented<KT>(V, s);
                for (u = W::prv(V, DIR, s); u != s; u = W::prv(V, DIR, u))
                    if (W::bud(V, DIR, u) >= 0) { const KT ku = sg_oriented<KT>(V, u); if (ku < mk) { mk = ku; mn = u; } }
                start = mn;
                stop = mn;
            }
            if (start != s) {
                int32_t Eb = W::bud(V, DIR, start);
                for (uint32_t v = W::nxt(V, DIR, start); v != s; v = W::nxt(V, DIR, v)) Eb = (W::face(V, DIR, v) < 0 && Eb >= 1) ? Eb - 1 : W::bud(V, DIR, v);
                fresh = !(Eb >= 1);  // face[s] < 0 here
            }
        }
        if (!fresh) continue;  // absorbed: the walk that takes it writes its flag
        int32_t rem = W::bud(V, DIR, s);
        uint32_t cur = s;
        unsigned long long taken = 0;
        while (rem >= 1) {
            const uint32_t z = W::nxt(V, DIR, cur);
            if (z == NONE32 || z == stop || z == s || W::face(V, DIR, z) >= 0) break;
            rem--;
            (DIR == 0 ? V.p[gid_rank(z)].eff_r : V.p[gid_rank(z)].eff_l)[gid_loc(z)] = rem;
            const uint32_t j = DIR == 0 ? cur : z;  // the junction's left node
            alive_or(V.p[gid_rank(j)].alive, gid_loc(j), DIR == 0 ? 16u : 32u);
            taken++;
            cur = z;
        }
        if (taken) atomicAdd(&dstat[DS_ABSORBED], taken);
    }
}
__global__ void sg_junction_finalize_kernel(const __grid_constant__ SGView V, uint64_t n) {
    const SGPeer& Me = V.p[V.me];
    for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (uint64_t)gridDim.x * blockDim.x) {
        if (!(Me.alive[x] & 2)) continue;
        const uint32_t y = Me.succ[x];
        if (y == NONE32) continue;
        if ((Me.alive[x] & 48) || junction_joins(Me.eff_r[x], V.p[gid_rank(y)].eff_l[gid_loc(y)])) continue;
        Me.succ[x] = NONE32;
        V.p[gid_rank(y)].pred[gid_loc(y)] = NONE32;
    }
}

// ---- K6: two levels of splitters ----------------------------------------------------------------------------------------
__device__ __forceinline__ bool sg_l1_sample(uint32_t x) { return (fmix32(x ^ 0xa5a5a5a5u) & 63u) == 0u; }
__device__ __forceinline__ bool sg_l2_sample(uint32_t l1_gid) { return (fmix32(l1_gid ^ 0x3c6ef372u) & 7u) == 0u; }

__global__ void sg_select_kernel(const __grid_constant__ SGView V, uint64_t n, uint32_t* __restrict__ spl_node, uint32_t* __restrict__ l2_l1, unsigned long long* dstat) {
    const SGPeer& Me = V.p[V.me];
    for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t id = NONE32;
        if (Me.alive[x] & 2) {
            const uint32_t p = Me.pred[x];
            if (p == NONE32 || gid_rank(p) != V.me || sg_l1_sample((uint32_t)x)) {
                id = (uint32_t)atomicAdd(&dstat[DS_NSPL], 1ull);
                spl_node[id] = (uint32_t)x;
                uint32_t j = NONE32;
                if (p == NONE32 || sg_l2_sample(gid_make(V.me, id))) {
                    j = (uint32_t)atomicAdd(&dstat[DS_NL2], 1ull);
                    l2_l1[j] = id;
                    Me.l2_node[j] = (uint32_t)x;
                    Me.l2_up[0][j] = ad_pack(gid_make(V.me, j), 0u);  // a non-head is overwritten by the level-2 walk that reaches it
                }
                Me.l2_of[id] = j;
            }
        }
        Me.spl_id[x] = id;
    }
}
// level-1 walk: purely local; the segment ends in front of the next level-1 splitter (on this rank or the first node on another)
__global__ void sg_l1_walk_kernel(const __grid_constant__ SGView V, uint64_t m1, const uint32_t* __restrict__ spl_node, uint64_t* __restrict__ loc) {
    const SGPeer& Me = V.p[V.me];
    for (uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; id < m1; id += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t x = spl_node[id];
        uint32_t off = 0, next = NONE32;
        loc[x] = ad_pack((uint32_t)id, 0u);
        uint32_t y = Me.succ[x];
        while (y != NONE32) {
            if (gid_rank(y) != V.me) { next = gid_make(gid_rank(y), V.p[gid_rank(y)].spl_id[gid_loc(y)]); break; }
            const uint32_t yl = gid_loc(y);
            const uint32_t s = Me.spl_id[yl];
            if (s != NONE32) { next = gid_make(V.me, s); break; }
            off++;
            loc[yl] = ad_pack((uint32_t)id, off);
            y = Me.succ[yl];
        }
        Me.l1_nl[id] = ad_pack(next, off + 1u);
    }
}
// level-2 walk over the level-1 list: stamps (owner, nodes in front) on every level-1 splitter it passes, ends at the next level-2 splitter
__global__ void sg_l2_walk_kernel(const __grid_constant__ SGView V, uint64_t m2, const uint32_t* __restrict__ l2_l1) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m2; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t self = gid_make(V.me, (uint32_t)j);
        uint32_t a = gid_make(V.me, l2_l1[j]);
        uint32_t dist = 0;
        V.p[V.me].l1_loc[gid_loc(a)] = ad_pack(self, 0u);
        while (true) {
            const uint64_t nl = V.p[gid_rank(a)].l1_nl[gid_loc(a)];
            const uint32_t nx = (uint32_t)nl;
            dist += (uint32_t)(nl >> 32);
            if (nx == NONE32) break;
            const SGPeer& Q = V.p[gid_rank(nx)];
            const uint32_t t = Q.l2_of[gid_loc(nx)];
            if (t != NONE32) { Q.l2_up[0][t] = ad_pack(self, dist); break; }
            Q.l1_loc[gid_loc(nx)] = ad_pack(self, dist);
            a = nx;
        }
    }
}
// one round of pointer jumping over the level-2 list (Jacobi: everybody reads buffer `in`, writes its own entries of `out`)
__global__ void sg_l2_jump_kernel(const __grid_constant__ SGView V, uint64_t m2, int in) {
    const SGPeer& Me = V.p[V.me];
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m2; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t mine = Me.l2_up[in][j];
        const uint32_t a = (uint32_t)mine;
        if (a == gid_make(V.me, (uint32_t)j)) { Me.l2_up[in ^ 1][j] = mine; continue; }
        const uint64_t up = V.p[gid_rank(a)].l2_up[in][gid_loc(a)];
        Me.l2_up[in ^ 1][j] = ad_pack((uint32_t)up, (uint32_t)(mine >> 32) + (uint32_t)(up >> 32));
    }
}
__global__ void sg_l2_fin_kernel(const __grid_constant__ SGView V, uint64_t m2, int res, unsigned long long* dstat) {
    const SGPeer& Me = V.p[V.me];
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < m2; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t v = Me.l2_up[res][j];
        const uint32_t a = (uint32_t)v;
        const uint32_t head = gid_make(gid_rank(a), V.p[gid_rank(a)].l2_node[gid_loc(a)]);
        if (V.p[gid_rank(head)].pred[gid_loc(head)] != NONE32) atomicExch(&dstat[DS_SG_CYCLE], 1ull);  // the chain's first splitter is no head
        Me.l2_fin[j] = ad_pack(head, (uint32_t)(v >> 32));
    }
}
__global__ void sg_l1_fin_kernel(const __grid_constant__ SGView V, uint64_t m1, uint64_t* __restrict__ l1_fin, unsigned long long* dstat) {
    const SGPeer& Me = V.p[V.me];
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m1; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t l = Me.l1_loc[i];
        if (l == ~0ull) { atomicExch(&dstat[DS_SG_CYCLE], 1ull); l1_fin[i] = ~0ull; continue; }  // no level-2 walk came by: closed path
        const uint32_t o = (uint32_t)l;
        const uint64_t f = V.p[gid_rank(o)].l2_fin[gid_loc(o)];
        l1_fin[i] = ad_pack((uint32_t)f, (uint32_t)(f >> 32) + (uint32_t)(l >> 32));
    }
}
__global__ void sg_node_fin_kernel(const __grid_constant__ SGView V, uint64_t n, const uint64_t* __restrict__ loc, const uint64_t* __restrict__ l1_fin, uint64_t* __restrict__ ad,
                                   unsigned long long* dstat) {
    const SGPeer& Me = V.p[V.me];
    for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (uint64_t)gridDim.x * blockDim.x) {
        if (!(Me.alive[x] & 2)) continue;
        const uint64_t l = loc[x];
        if (l == ~0ull) { atomicExch(&dstat[DS_SG_CYCLE], 1ull); continue; }  // no level-1 walk came by: closed path inside this rank
        const uint64_t f = l1_fin[(uint32_t)l];
        if (f == ~0ull) continue;
        const uint32_t d = (uint32_t)(f >> 32) + (uint32_t)(l >> 32);
        ad[x] = ad_pack((uint32_t)f, d);
        if (Me.succ[x] == NONE32) {  // tail: tell the head (wherever it lives) how long its chain is and what its right flag is
            const uint32_t h = (uint32_t)f;
            V.p[gid_rank(h)].chain_len[gid_loc(h)] = d + 1u;
            V.p[gid_rank(h)].tail_rf[gid_loc(h)] = Me.eff_r[x];
        }
    }
}

// ---- K7 -------------------------------------------------------------------------------------------------------------------
struct SContigIn {
    const uint8_t* alive;
    const uint32_t* pred;
    const uint32_t* chain_len;
    const int32_t* lflag;    // effective left flag of the head
    const int32_t* tail_rf;  // effective right flag of the chain's tail, stored at the head
    int k, min_contig;
    __device__ __forceinline__ U64x3 operator()(uint64_t x) const {
        if (!(alive[x] & 2)) return U64x3{0, 0, 0};
        if (pred[x] != NONE32) return U64x3{0, 0, 1};
        const uint64_t len = (uint64_t)chain_len[x] + (uint64_t)k - 1;
        const bool keep = !(lflag[x] <= -10000000 && tail_rf[x] <= -10000000) && len >= (uint64_t)min_contig;  // DSKmerToContig, ReflexivDSMain.java:749-754
        return keep ? U64x3{1, len, 1} : U64x3{0, 0, 1};
    }
};
struct SContigOut {
    const int32_t* lflag;
    const int32_t* tail_rf;
    uint32_t* ctg_idx;
    uint64_t* ctg_off;
    int32_t* ctg_left;
    int32_t* ctg_right;
    __device__ __forceinline__ void operator()(uint64_t x, U64x3 excl, U64x3 v) const {
        ctg_idx[x] = v.a ? (uint32_t)excl.a : NONE32;
        if (v.a) {
            ctg_off[excl.a] = excl.b;
            ctg_left[excl.a] = lflag[x];
            ctg_right[excl.a] = tail_rf[x];
        }
    }
};
// where the bases of a level-1 segment go: address of its first node's last base inside the head owner's buffer (0: contig not kept)
__global__ void sg_l1_dst_kernel(const __grid_constant__ SGView V, uint64_t m1, const uint64_t* __restrict__ l1_fin, unsigned long long* __restrict__ l1_dst) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m1; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t f = l1_fin[i];
        unsigned long long d = 0;
        if (f != ~0ull) {
            const uint32_t h = (uint32_t)f;
            const SGPeer& Q = V.p[gid_rank(h)];
            const uint32_t ci = Q.ctg_idx[gid_loc(h)];
            if (ci != NONE32) d = (unsigned long long)(Q.ctg_bases + Q.ctg_off[ci] + (uint64_t)(V.k - 1) + (uint32_t)(f >> 32));
        }
        l1_dst[i] = d;
    }
}
template <class KT>
__global__ void sg_gather_kernel(const __grid_constant__ SGView V, uint64_t n, const uint64_t* __restrict__ loc, const unsigned long long* __restrict__ l1_dst) {
    const SGPeer& Me = V.p[V.me];
    for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (uint64_t)gridDim.x * blockDim.x) {
        if (!(Me.alive[x] & 2)) continue;
        const uint64_t l = loc[x];
        if (l == ~0ull) continue;
        const unsigned long long d = l1_dst[(uint32_t)l];
        if (!d) continue;
        const KT X = sg_oriented<KT>(V, gid_make(V.me, (uint32_t)x));
        char* dst = reinterpret_cast<char*>(d) + (uint32_t)(l >> 32);
        *dst = "ACGT"[(uint32_t)X & 3u];
        if (Me.pred[x] == NONE32)  // head: its contig lives on this rank; the first k-1 bases come from it as well
            for (int j = 0; j < V.k - 1; j++) dst[j - (V.k - 1)] = "ACGT"[(uint32_t)(X >> (2 * (V.k - 1 - j))) & 3u];
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------------
// publish the arena offsets of this rank's graph buffers + values; rebuild the view of everybody's buffers
static int sg_publish(Ctx* c, SGView& V, const unsigned long long host_vals[GP_NHOST], int n_dev, const int* dev_slots, unsigned long long* all) {
    GShard* gs = c->gshard;
    unsigned long long mine[GP_END];
    auto off_of = [&](const void* p) -> unsigned long long { return p ? (unsigned long long)((const uint8_t*)p - c->arena) : ~0ull; };
    const void* ptrs[GP_NPTR] = {c->keys.p, c->counts.p, c->ht.p, c->g_bloom.p, c->alive.p, c->lflag.p, c->rflag.p, c->eff_l.p, c->eff_r.p, c->succ.p, c->pred.p, c->spl_id.p,
                                 gs->l1_nl.p, gs->l1_loc.p, gs->l2_of.p, gs->l2_up[0].p, gs->l2_up[1].p, gs->l2_node.p, gs->l2_fin.p, c->chain_len.p, gs->tail_rf.p,
                                 c->ctg_idx.p, c->ctg_off.p, c->ctg_bases.p};
    for (int i = 0; i < GP_NPTR; i++) mine[i] = off_of(ptrs[i]);
    mine[GP_NROWS] = c->n_rows;
    mine[GP_HTCAP] = c->ht_cap;
    mine[GP_BLOOMMASK] = c->g_bloom_mask;
    for (int i = 0; i < GP_NHOST; i++) mine[GP_HOST + i] = host_vals ? host_vals[i] : 0ull;
    // device values: gathered into a staging block right behind the host values by one small copy each
    for (int i = 0; i < GP_NDEV; i++) mine[GP_DEV + i] = 0ull;
    RFX_TRY(shard_exchange(c, PUB_GRAPH, GP_END, mine, all, n_dev, dev_slots, PUB_GRAPH + GP_DEV));
    V.me = c->sh_rank; V.world = c->sh_world; V.k = c->k; V.m = c->m;
    V.B = c->n_bins; V.bps = c->n_bins / (uint32_t)c->sh_world;
    for (int r = 0; r < RFX_MAX_RANKS; r++) {
        SGPeer& P = V.p[r];
        memset(&P, 0, sizeof(P));
        if (r >= c->sh_world) continue;
        const unsigned long long* v = all + (size_t)r * RFX_PUB_SLOTS + PUB_GRAPH;
        uint8_t* base = c->peer_base[r];
        auto at = [&](int slot) -> uint8_t* { return v[slot] == ~0ull ? nullptr : base + v[slot]; };
        P.keys = at(GP_KEYS); P.counts = (const uint32_t*)at(GP_COUNTS); P.ht = (const uint32_t*)at(GP_HT); P.bloom = (const uint32_t*)at(GP_BLOOM);
        P.alive = at(GP_ALIVE); P.lflag = (int32_t*)at(GP_LFLAG); P.rflag = (int32_t*)at(GP_RFLAG); P.eff_l = (int32_t*)at(GP_EFFL); P.eff_r = (int32_t*)at(GP_EFFR);
        P.succ = (uint32_t*)at(GP_SUCC); P.pred = (uint32_t*)at(GP_PRED); P.spl_id = (uint32_t*)at(GP_SPLID);
        P.l1_nl = (uint64_t*)at(GP_L1NL); P.l1_loc = (uint64_t*)at(GP_L1LOC); P.l2_of = (uint32_t*)at(GP_L2OF);
        P.l2_up[0] = (uint64_t*)at(GP_L2UP0); P.l2_up[1] = (uint64_t*)at(GP_L2UP1); P.l2_node = (uint32_t*)at(GP_L2NODE); P.l2_fin = (uint64_t*)at(GP_L2FIN);
        P.chain_len = (uint32_t*)at(GP_CHAINLEN); P.tail_rf = (int32_t*)at(GP_TAILRF); P.ctg_idx = (uint32_t*)at(GP_CTGIDX); P.ctg_off = (uint64_t*)at(GP_CTGOFF);
        P.ctg_bases = (char*)at(GP_CTGBASES);
        if (gs->bloom_local[r]) P.bloom = gs->bloom_local[r];
        P.bloom_mask = v[GP_BLOOMMASK];
        P.ht_cap = (uint32_t)v[GP_HTCAP];
    }
    return RFX_OK;
}
static inline unsigned long long sg_sum(const unsigned long long* all, int world, int slot) {
    unsigned long long s = 0;
    for (int r = 0; r < world; r++) s += all[(size_t)r * RFX_PUB_SLOTS + PUB_GRAPH + slot];
    return s;
}

// a closed path somewhere: rank 0 pulls every shard table and runs the single-GPU stages over the whole table
template <class KT> static int sg_fallback_whole_table(Ctx* c, const SGView& V, const unsigned long long* all) {
    GShard* gs = c->gshard;
    cudaStream_t st = c->stream;
    gs->fell_back = 1;
    int rc = RFX_OK;
    if (c->sh_rank == 0) {
        uint64_t tot = 0;
        for (int r = 0; r < c->sh_world; r++) tot += all[(size_t)r * RFX_PUB_SLOTS + PUB_GRAPH + GP_NROWS];
        RFX_TRY(devbuf_reserve(c, gs->all_keys, (tot + 1) * sizeof(KT)));
        RFX_TRY(devbuf_reserve(c, gs->all_counts, (tot + 1) * sizeof(uint32_t)));
        uint64_t pos = 0;
        for (int r = 0; r < c->sh_world; r++) {
            const uint64_t nr = all[(size_t)r * RFX_PUB_SLOTS + PUB_GRAPH + GP_NROWS];
            if (nr) {
                RFX_CUDA(c, cudaMemcpyAsync(gs->all_keys.as<KT>() + pos, V.p[r].keys, nr * sizeof(KT), cudaMemcpyDefault, st));
                RFX_CUDA(c, cudaMemcpyAsync(gs->all_counts.as<uint32_t>() + pos, V.p[r].counts, nr * sizeof(uint32_t), cudaMemcpyDefault, st));
            }
            pos += nr;
        }
        DevBuf own_keys = c->keys, own_counts = c->counts;
        const uint64_t own_rows = c->n_rows;
        c->keys = gs->all_keys; c->counts = gs->all_counts; c->n_rows = tot;
        rc = graph_impl<KT>(c);
        gs->all_keys = c->keys; gs->all_counts = c->counts;
        c->keys = own_keys; c->counts = own_counts; c->n_rows = own_rows;
        c->have_contigs = false;  // alive / flags describe the whole table, not this rank's rows: only the contigs are kept
    } else {
        RFX_TRY(devbuf_reserve(c, c->ctg_off, sizeof(uint64_t)));
        RFX_CUDA(c, cudaMemsetAsync(c->ctg_off.p, 0, sizeof(uint64_t), st));
        c->n_contigs = c->n_contig_bases = c->n_oriented = c->n_budget = c->n_budget_adm = c->n_cycles = 0;
    }
    return rc;
}

template <class KT> static int sharded_graph_impl(Ctx* c) {
    cudaStream_t st = c->stream;
    if (!c->gshard) c->gshard = new GShard();
    GShard* gs = c->gshard;
    const int world = c->sh_world;
    const uint64_t n_rows = c->n_rows, n = 2 * n_rows, nn = n ? n : 1;
    unsigned long long* dstat = c->dstat.as<unsigned long long>();
    unsigned long long all[RFX_MAX_RANKS * RFX_PUB_SLOTS];
    if (n >= GID_MASK) return ctx_fail(c, RFX_E_CAPACITY, "sharded assembly: more than 2^28 rows on one rank");
    if (!c->n_bins || c->n_bins % (uint32_t)world) return ctx_fail(c, RFX_E_STATE, "sharded assembly needs the table of rfx_count_sharded (rows sharded by minimiser bin)");
    RFX_CUDA(c, cudaMemsetAsync(c->dstat.p, 0, DS_NSLOTS * sizeof(uint64_t), st));
    c->n_oriented = c->n_budget = c->n_budget_adm = c->n_cycles = c->n_contigs = c->n_contig_bases = 0;
    c->have_sorted = false; c->have_contigs = false;
    c->ms_comm = 0;
    gs->fell_back = 0;
    for (int r = 0; r < RFX_MAX_RANKS; r++) gs->bloom_local[r] = nullptr;

    // ---- K5 ----
    stage_begin(c);
    RFX_TRY(devbuf_reserve(c, c->rflag, nn * sizeof(int32_t)));
    RFX_TRY(devbuf_reserve(c, c->lflag, nn * sizeof(int32_t)));
    RFX_TRY(devbuf_reserve(c, c->eff_l, nn * sizeof(int32_t)));
    RFX_TRY(devbuf_reserve(c, c->eff_r, nn * sizeof(int32_t)));
    RFX_TRY(devbuf_reserve(c, c->alive, nn + 8));
    RFX_TRY(devbuf_reserve(c, c->succ, nn * sizeof(uint32_t)));
    RFX_TRY(devbuf_reserve(c, c->pred, nn * sizeof(uint32_t)));
    RFX_TRY(devbuf_reserve(c, c->spl_id, nn * sizeof(uint32_t)));
    RFX_TRY(devbuf_reserve(c, c->spl_node, nn * sizeof(uint32_t)));
    RFX_TRY(devbuf_reserve(c, c->loc, nn * sizeof(uint64_t)));
    RFX_TRY(devbuf_reserve(c, c->ad[0], nn * sizeof(uint64_t)));
    RFX_TRY(devbuf_reserve(c, c->chain_len, nn * sizeof(uint32_t)));
    RFX_TRY(devbuf_reserve(c, c->ctg_idx, nn * sizeof(uint32_t)));
    RFX_TRY(devbuf_reserve(c, gs->tail_rf, nn * sizeof(int32_t)));
    RFX_TRY(devbuf_reserve(c, gs->l1_nl, nn * sizeof(uint64_t)));
    RFX_TRY(devbuf_reserve(c, gs->l1_loc, nn * sizeof(uint64_t)));
    RFX_TRY(devbuf_reserve(c, gs->l1_fin, nn * sizeof(uint64_t)));
    RFX_TRY(devbuf_reserve(c, gs->l1_dst, nn * sizeof(uint64_t)));
    RFX_TRY(devbuf_reserve(c, gs->l2_of, nn * sizeof(uint32_t)));
    RFX_TRY(devbuf_reserve(c, gs->l2_l1, nn * sizeof(uint32_t)));
    RFX_TRY(devbuf_reserve(c, gs->l2_node, nn * sizeof(uint32_t)));
    for (int i = 0; i < 2; i++) RFX_TRY(devbuf_reserve(c, gs->l2_up[i], nn * sizeof(uint64_t)));
    RFX_TRY(devbuf_reserve(c, gs->l2_fin, nn * sizeof(uint64_t)));
    RFX_TRY(devbuf_reserve(c, c->ctg_off, sizeof(uint64_t)));
    RFX_TRY(devbuf_reserve(c, c->ctg_bases, 16));
    // the own index: one region, load factor <= 1/4, presence bits in front of it (>= 16 per row)
    uint64_t bits = 1024;
    while (bits < 16 * n_rows) bits <<= 1;
    const uint64_t slots = 4 * n_rows + 2;
    RFX_TRY(devbuf_reserve(c, c->ht, slots * sizeof(uint32_t)));
    RFX_TRY(devbuf_reserve(c, c->g_bloom, bits / 8));
    c->ht_cap = slots; c->g_bins = 1; c->g_m = c->m; c->g_bloom_mask = bits - 1;
    RFX_CUDA(c, cudaMemsetAsync(c->ht.p, 0xff, slots * sizeof(uint32_t), st));
    RFX_CUDA(c, cudaMemsetAsync(c->g_bloom.p, 0, bits / 8, st));
    if (n_rows) {
        Graph<KT> G = make_graph<KT>(c);
        ht_build_kernel<KT><<<grid_n(n_rows), 256, 0, st>>>(G, nullptr);
        c->launches++;
    }
    cudaMemsetAsync(c->succ.p, 0xff, nn * sizeof(uint32_t), st);
    cudaMemsetAsync(c->pred.p, 0xff, nn * sizeof(uint32_t), st);
    cudaMemsetAsync(c->chain_len.p, 0, nn * sizeof(uint32_t), st);
    cudaMemsetAsync(gs->tail_rf.p, 0, nn * sizeof(int32_t), st);
    cudaMemsetAsync(c->loc.p, 0xff, nn * sizeof(uint64_t), st);
    cudaMemsetAsync(gs->l1_loc.p, 0xff, nn * sizeof(uint64_t), st);
    cudaMemsetAsync(c->alive.p, 0, nn + 8, st);

    SGView V;
    RFX_TRY(sg_publish(c, V, nullptr, 0, nullptr, all));  // [barrier] everybody's index and presence bits are complete
    gs->n_rows_global = sg_sum(all, world, GP_NROWS);
    {   // presence bits of the peers -> local copies
        uint64_t total = 0;
        for (int r = 0; r < world; r++) if (r != c->sh_rank) total += (V.p[r].bloom_mask + 1) / 8;
        RFX_TRY(devbuf_reserve(c, gs->bloom_all, total + 256));
        uint64_t pos = 0;
        for (int r = 0; r < world; r++) {
            if (r == c->sh_rank) continue;
            const uint64_t nb = (V.p[r].bloom_mask + 1) / 8;
            RFX_CUDA(c, cudaMemcpyAsync(gs->bloom_all.as<uint8_t>() + pos, V.p[r].bloom, nb, cudaMemcpyDefault, st));
            gs->bloom_local[r] = reinterpret_cast<const uint32_t*>(gs->bloom_all.as<uint8_t>() + pos);
            V.p[r].bloom = gs->bloom_local[r];
            pos += nb;
        }
    }
    const int E = c->prm.min_error_coverage;
    if (n_rows) sg_check_owner_kernel<KT><<<grid_n(n_rows), 256, 0, st>>>(V, n_rows, dstat);
    if (n) sg_right_filter_kernel<KT><<<grid_n(n), 256, 0, st>>>(V, E, n, dstat);
    RFX_TRY(shard_barrier(c));
    if (n) sg_left_filter_kernel<KT><<<grid_n(n), 256, 0, st>>>(V, E, n, dstat);
    RFX_TRY(shard_barrier(c));
    if (n) sg_link_kernel<KT><<<grid_n(n), 256, 0, st>>>(V, n, dstat);
    c->launches += 4;
    {
        const int slots_dev[3] = {DS_FLAGGED, DS_GRAPH_ERR, DS_REMOTE};
        RFX_TRY(sg_publish(c, V, nullptr, 3, slots_dev, all));  // [barrier] every pred[] has its remote writes
    }
    const unsigned long long flagged = sg_sum(all, world, GP_DEV + 0), gerr = sg_sum(all, world, GP_DEV + 1);
    gs->n_remote = all[(size_t)c->sh_rank * RFX_PUB_SLOTS + PUB_GRAPH + GP_DEV + 2];
    if (gerr) return ctx_fail(c, RFX_E_GRAPH, "sharded fork filters: a (k-1)-mer with degree > 1, or a row on a rank that does not own its minimiser bin (codes add up to %llu)", gerr);
    if (flagged) {
        if (n) {
            sg_budget_walk_kernel<KT, 0><<<grid_n(n), 256, 0, st>>>(V, n, c->k - 1, dstat);
            sg_budget_walk_kernel<KT, 1><<<grid_n(n), 256, 0, st>>>(V, n, c->k - 1, dstat);
        }
        RFX_TRY(shard_barrier(c));
        if (n) sg_junction_finalize_kernel<<<grid_n(n), 256, 0, st>>>(V, n);
        c->launches += 3;
        // the barrier inside the next exchange orders the cuts before the splitter selection of the peers
    }

    // ---- K6 ----
    RFX_TRY(shard_barrier(c));
    RFX_TRY(shard_check(c, "sharded fork filters"));
    c->ms[3] += stage_end(c);
    stage_begin(c);
    if (n) sg_select_kernel<<<grid_n(n), 256, 0, st>>>(V, n, c->spl_node.as<uint32_t>(), gs->l2_l1.as<uint32_t>(), dstat);
    c->launches++;
    {
        const int slots_dev[2] = {DS_NSPL, DS_NL2};
        RFX_TRY(sg_publish(c, V, nullptr, 2, slots_dev, all));  // [barrier] spl_id / l2_of of every rank are final
    }
    const uint64_t m1 = all[(size_t)c->sh_rank * RFX_PUB_SLOTS + PUB_GRAPH + GP_DEV + 0];
    const uint64_t m2 = all[(size_t)c->sh_rank * RFX_PUB_SLOTS + PUB_GRAPH + GP_DEV + 1];
    const uint64_t m2_all = sg_sum(all, world, GP_DEV + 1);
    gs->n_l1 = m1; gs->n_l2 = m2;
    int rounds = 1;
    while ((1ull << rounds) < m2_all + 1) rounds++;
    rounds += 1;
    if (m1) sg_l1_walk_kernel<<<grid_n(m1), 256, 0, st>>>(V, m1, c->spl_node.as<uint32_t>(), c->loc.as<uint64_t>());
    RFX_TRY(shard_barrier(c));
    if (m2) {
        uint64_t g2 = (m2 + 63) / 64;
        if (g2 > sm_count() * 32u) g2 = sm_count() * 32u;
        sg_l2_walk_kernel<<<(unsigned)g2, 64, 0, st>>>(V, m2, gs->l2_l1.as<uint32_t>());
    }
    RFX_TRY(shard_barrier(c));
    c->launches += 2;
    int cur = 0;
    for (int r = 0; r < rounds; r++) {
        if (m2) sg_l2_jump_kernel<<<grid_n(m2), 256, 0, st>>>(V, m2, cur);
        RFX_TRY(shard_barrier(c));
        c->launches++;
        cur ^= 1;
    }
    if (m2) sg_l2_fin_kernel<<<grid_n(m2), 256, 0, st>>>(V, m2, cur, dstat);
    RFX_TRY(shard_barrier(c));
    if (m1) sg_l1_fin_kernel<<<grid_n(m1), 256, 0, st>>>(V, m1, gs->l1_fin.as<uint64_t>(), dstat);
    if (n) sg_node_fin_kernel<<<grid_n(n), 256, 0, st>>>(V, n, c->loc.as<uint64_t>(), gs->l1_fin.as<uint64_t>(), c->ad[0].as<uint64_t>(), dstat);
    c->launches += 3;
    {
        const int slots_dev[5] = {DS_SG_CYCLE, DS_BUDGET, DS_ABSORBED, DS_CYCLES, DS_XBAR_ERR};
        RFX_TRY(sg_publish(c, V, nullptr, 5, slots_dev, all));  // [barrier] every head knows its chain's length and right flag
    }
    c->ms[4] += stage_end(c);
    const unsigned long long any_cycle = sg_sum(all, world, GP_DEV + 0);
    c->n_budget = all[(size_t)c->sh_rank * RFX_PUB_SLOTS + PUB_GRAPH + GP_DEV + 1];
    c->n_budget_adm = all[(size_t)c->sh_rank * RFX_PUB_SLOTS + PUB_GRAPH + GP_DEV + 2];
    c->n_cycles = all[(size_t)c->sh_rank * RFX_PUB_SLOTS + PUB_GRAPH + GP_DEV + 3];
    if (any_cycle) {
        stage_begin(c);
        const int frc = sg_fallback_whole_table<KT>(c, V, all);
        // nobody may touch its table before rank 0 has read it
        RFX_TRY(shard_barrier(c));
        RFX_TRY(shard_check(c, "sharded assembly (whole-table fallback)"));
        c->ms[5] += stage_end(c);
        RFX_TRY(frc);
        c->have_contigs = true;
        unsigned long long hv[GP_NHOST] = {c->n_contigs, c->n_contig_bases, c->n_oriented, 0, 0, 0};
        RFX_TRY(sg_publish(c, V, hv, 0, nullptr, all));
        gs->n_contigs_global = sg_sum(all, world, GP_HOST + 0);
        gs->n_bases_global = sg_sum(all, world, GP_HOST + 1);
        gs->n_oriented_global = sg_sum(all, world, GP_HOST + 2);
        return RFX_OK;
    }

    // ---- K7 ----
    stage_begin(c);
    ScanPlan<U64x3> plan;
    RFX_TRY(devbuf_reserve(c, c->scan_ws, ScanPlan<U64x3>::workspace_elems(nn) * sizeof(U64x3)));
    plan.bind(n, c->scan_ws.as<U64x3>());
    SContigIn in{c->alive.as<uint8_t>(), c->pred.as<uint32_t>(), c->chain_len.as<uint32_t>(), c->eff_l.as<int32_t>(), gs->tail_rf.as<int32_t>(), c->k, c->prm.min_contig};
    U64x3 tot{0, 0, 0};
    if (n) {
        scan_prepare(plan, in, OpAddU64x3{}, U64x3{0, 0, 0}, st);
        c->launches += 2 * plan.levels;
        RFX_CUDA(c, cudaMemcpyAsync(&tot, plan.total, sizeof(tot), cudaMemcpyDeviceToHost, st));
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return ctx_fail(c, RFX_E_CUDA, "sharded contig scan failed: %s", cudaGetErrorString(e));
    }
    RFX_TRY(devbuf_reserve(c, c->ctg_off, (tot.a + 1) * sizeof(uint64_t)));
    RFX_TRY(devbuf_reserve(c, c->ctg_left, (tot.a + 1) * sizeof(int32_t)));
    RFX_TRY(devbuf_reserve(c, c->ctg_right, (tot.a + 1) * sizeof(int32_t)));
    RFX_TRY(devbuf_reserve(c, c->ctg_bases, tot.b + 16));
    if (n) {
        SContigOut out{c->eff_l.as<int32_t>(), gs->tail_rf.as<int32_t>(), c->ctg_idx.as<uint32_t>(), c->ctg_off.as<uint64_t>(), c->ctg_left.as<int32_t>(), c->ctg_right.as<int32_t>()};
        scan_apply(plan, in, out, OpAddU64x3{}, U64x3{0, 0, 0}, st);
        set_u64_kernel<<<1, 1, 0, st>>>(c->ctg_off.as<uint64_t>() + tot.a, plan.total);
        c->launches += 2;
    } else {
        RFX_CUDA(c, cudaMemsetAsync(c->ctg_off.p, 0, sizeof(uint64_t), st));
    }
    c->n_contigs = tot.a; c->n_contig_bases = tot.b; c->n_oriented = tot.c;
    {
        unsigned long long hv[GP_NHOST] = {tot.a, tot.b, tot.c, 0, 0, 0};
        RFX_TRY(sg_publish(c, V, hv, 0, nullptr, all));  // [barrier] contig tables laid out, buffers published
    }
    gs->n_contigs_global = sg_sum(all, world, GP_HOST + 0);
    gs->n_bases_global = sg_sum(all, world, GP_HOST + 1);
    gs->n_oriented_global = sg_sum(all, world, GP_HOST + 2);
    if (m1) sg_l1_dst_kernel<<<grid_n(m1), 256, 0, st>>>(V, m1, gs->l1_fin.as<uint64_t>(), gs->l1_dst.as<unsigned long long>());
    if (n) sg_gather_kernel<KT><<<grid_n(n), 256, 0, st>>>(V, n, c->loc.as<uint64_t>(), gs->l1_dst.as<unsigned long long>());
    c->launches += 2;
    RFX_TRY(shard_barrier(c));  // every base of the own contigs has arrived
    RFX_TRY(shard_check(c, "sharded assembly"));
    c->ms[5] += stage_end(c);
    c->have_contigs = true;
    return RFX_OK;
}

int stage_assemble_sharded(Ctx* c) {
    if (c->sh_world < 1 || !c->peer_base[c->sh_rank]) return ctx_fail(c, RFX_E_STATE, "rfx_assemble_sharded: call rfx_shard_init / rfx_shard_connect first");
    if (!c->have_counts) return ctx_fail(c, RFX_E_STATE, "rfx_assemble_sharded: no count table (call rfx_count_sharded first)");
    if (!c->prm.bubble) return ctx_fail(c, RFX_E_UNSUPPORTED, "-bubble: undefined in the reference (see rfx_assemble)");
    if (c->k < 2) return ctx_fail(c, RFX_E_INVALID, "assembly needs k >= 2");
    return c->wide ? sharded_graph_impl<u128>(c) : sharded_graph_impl<uint64_t>(c);
}

void shard_graph_release(Ctx* c) {
    delete c->gshard;
    c->gshard = nullptr;
}

void shard_graph_stats(Ctx* c, rfx_shard_stats_t* out) {
    if (!c->gshard) return;
    const GShard* gs = c->gshard;
    out->n_rows_global = gs->n_rows_global;
    out->n_oriented_global = gs->n_oriented_global;
    out->n_contigs_global = gs->n_contigs_global;
    out->n_contig_bases_global = gs->n_bases_global;
    out->n_remote_probes = gs->n_remote;
    out->n_l1_splitters = gs->n_l1;
    out->n_l2_splitters = gs->n_l2;
    out->fell_back = gs->fell_back;
}

}  // namespace rfx
