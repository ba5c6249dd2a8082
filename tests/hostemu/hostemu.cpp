// hostemu.cpp -- TEST-ONLY host harness around reflexiv_b200/csrc/rfx_core.h.
//
// The arithmetic of the CUDA kernels (2-bit packing, minimiser binning, super-k-mer records, rolling
// canonical k-mers, fork-filter rules) is written once as __host__ __device__ functions.  This file
// drives those same functions from plain loops so the logic can be checked against the oracle on a
// machine without a GPU.  It is not part of libreflexiv_cuda and nothing in reflexiv_b200/ uses it.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <unordered_map>
#include <vector>

#include "../../reflexiv_b200/csrc/rfx_core.h"

using namespace rfx;

extern "C" {

// packs reads exactly as encode_reads_kernel lays them out; returns total words
int64_t emu_pack_reads(const uint8_t* text, const uint64_t* starts, const uint32_t* lens, int64_t n, int k, int fc, int ec,
                       uint32_t* elen_out, uint64_t* woff_out, uint64_t* words_out /* may be NULL to size */) {
    uint64_t w = 0;
    for (int64_t r = 0; r < n; r++) {
        uint32_t e = effective_read_len(lens[r], k, fc, ec);
        elen_out[r] = e;
        woff_out[r] = w;
        uint32_t nw = (e + 31) >> 5;
        if (words_out) {
            for (uint32_t j = 0; j < nw; j++) words_out[w + j] = 0;
            for (uint32_t i = 0; i < e; i++) {
                uint64_t c = base_code(text[starts[r] + fc + i]);
                words_out[w + (i >> 5)] |= c << (62 - 2 * (i & 31));
            }
        }
        w += nw;
    }
    return (int64_t)w;
}

struct EmitVec {
    const uint64_t* rd;
    int recw;
    int k;
    std::vector<uint64_t>* recs;
    std::vector<uint32_t>* bins;
    void operator()(uint32_t bin, uint32_t first, uint32_t nk) const {
        uint64_t rec[4] = {0, 0, 0, 0};
        if (recw == 2) rec_build<2>(rd, first, nk, k, rec);
        else rec_build<4>(rd, first, nk, k, rec);
        for (int i = 0; i < recw; i++) recs->push_back(rec[i]);
        bins->push_back(bin);
    }
};

static std::vector<uint64_t> g_recs;
static std::vector<uint32_t> g_bins;

// runs bin_scan_read over every read; results fetched with emu_partition_fetch
int64_t emu_partition(const uint64_t* words, int64_t n_words_padded, const uint32_t* elen, const uint64_t* woff, int64_t n, int k, int m,
                      uint32_t n_bins) {
    (void)n_words_padded;
    g_recs.clear(); g_bins.clear();
    BinParams P;
    P.k = k; P.m = m; P.w = k - m + 1; P.n_bins = n_bins;
    const int recw = rec_words_for_k(k);
    P.max_nk = (uint32_t)rec_max_kmers(recw, k);
    std::vector<uint32_t> ring(2 * P.w);
    for (int64_t r = 0; r < n; r++) {
        if (elen[r] < (uint32_t)k) continue;
        bin_scan_read(words + woff[r], elen[r], P, ring.data(), 1u, EmitVec{words + woff[r], recw, k, &g_recs, &g_bins});
    }
    return (int64_t)g_bins.size();
}
void emu_partition_fetch(uint64_t* recs, uint32_t* bins) {
    memcpy(recs, g_recs.data(), g_recs.size() * 8);
    memcpy(bins, g_bins.data(), g_bins.size() * 4);
}

struct MapSink {
    std::map<u128, uint32_t>* m;
    void operator()(uint64_t key) const { (*m)[(u128)key]++; }
    void operator()(u128 key) const { (*m)[key]++; }
};

static std::map<u128, uint32_t> g_counts;

// counts the k-mers of the records (any order); checks on the way that every record's k-mers share its bin
int64_t emu_count_records(const uint64_t* recs, const uint32_t* bins, int64_t n_rec, int k, int m, uint32_t n_bins, int64_t* bin_mismatch) {
    g_counts.clear();
    const int recw = rec_words_for_k(k);
    BinParams P;
    P.k = k; P.m = m; P.w = k - m + 1; P.n_bins = n_bins; P.max_nk = (uint32_t)rec_max_kmers(recw, k);
    int64_t bad = 0;
    for (int64_t r = 0; r < n_rec; r++) {
        const uint64_t* rec = recs + r * recw;
        MapSink s{&g_counts};
        if (recw == 2) {
            if (k <= 31) rec_foreach_kmer<uint64_t, 2>(rec, k, s); else rec_foreach_kmer<u128, 2>(rec, k, s);
            if (bins && rec_first_bin<2>(rec, P) != bins[r]) bad++;
        } else {
            rec_foreach_kmer<u128, 4>(rec, k, s);
            if (bins && rec_first_bin<4>(rec, P) != bins[r]) bad++;
        }
    }
    if (bin_mismatch) *bin_mismatch = bad;
    return (int64_t)g_counts.size();
}
void emu_count_fetch(uint64_t* hi, uint64_t* lo, uint32_t* cnt) {
    size_t i = 0;
    for (auto& kv : g_counts) { hi[i] = (uint64_t)(kv.first >> 64); lo[i] = (uint64_t)kv.first; cnt[i] = kv.second; i++; }
}

// bin of every k-mer position of one read, brute force (no sliding window): reference for bin_scan_read
int64_t emu_bins_bruteforce(const uint64_t* rd, uint32_t len, int k, int m, uint32_t n_bins, uint32_t* out) {
    if (len < (uint32_t)k) return 0;
    const uint32_t mmask = m >= 16 ? 0xffffffffu : ((1u << (2 * m)) - 1u);
    for (uint32_t i = 0; i + k <= len; i++) {
        uint32_t hmin = 0xffffffffu;
        for (uint32_t j = i; j + m <= i + k; j++) {
            uint32_t mf = 0;
            for (int t = 0; t < m; t++) mf = ((mf << 2) | packed_base(rd, j + t)) & mmask;
            uint32_t mr = (uint32_t)revcomp((uint64_t)mf, m);
            uint32_t h = mmer_hash(mf < mr ? mf : mr);
            hmin = h < hmin ? h : hmin;
        }
        out[i] = bin_of_minimizer(hmin, n_bins);
    }
    return len - k + 1;
}

}  // extern "C"

// the GPU formulation of A6-A8 + links, driven from an std::map instead of the device hash index
static std::vector<u128> g_okeys;
static std::vector<int32_t> g_oleft, g_oright;

template <class KT> static int64_t fork_impl(const std::map<u128, uint32_t>& tab, int k, int E) {
    g_okeys.clear(); g_oleft.clear(); g_oright.clear();
    auto cnt_of = [&](KT Z) -> uint32_t {
        KT zc = revcomp(Z, k);
        auto it = tab.find((u128)(zc < Z ? zc : Z));
        return it == tab.end() ? 0u : it->second;
    };
    std::map<u128, int32_t> rflag;  // survivors of the right filter
    for (auto& kv : tab) {
        KT key = (KT)kv.first, rc = revcomp(key, k);
        for (int strand = 0; strand < 2; strand++) {
            if (strand && rc == key) continue;
            KT X = strand ? rc : key;
            KT prefix = X >> 2;
            uint32_t myb = (uint32_t)X & 3u, cnt[4];
            bool dup[4];
            for (uint32_t b = 0; b < 4; b++) {
                KT Z = (prefix << 2) | (KT)b;
                cnt[b] = b == myb ? kv.second : cnt_of(Z);
                dup[b] = (Z == revcomp(Z, k));
            }
            ForkResult res = right_fork(cnt, dup, E, k - 1);
            if (res.winner == (int)myb) rflag[(u128)X] = res.flag;
        }
    }
    const int top = 2 * (k - 1);
    const KT sufmask = mask_bases<KT>(k - 1);
    for (auto& kv : rflag) {
        KT X = (KT)kv.first;
        KT suffix = X & sufmask;
        uint32_t mya = (uint32_t)(X >> top) & 3u, cnt[4];
        for (uint32_t a = 0; a < 4; a++) {
            KT Z = ((KT)a << top) | suffix;
            cnt[a] = rflag.count((u128)Z) ? cnt_of(Z) : 0u;
        }
        ForkResult res = left_fork(cnt, E, k - 1);
        if (res.winner == (int)mya) { g_okeys.push_back((u128)X); g_oleft.push_back(res.flag); g_oright.push_back(kv.second); }
    }
    return (int64_t)g_okeys.size();
}

// Count_<k>_sorted: the per-group formulation of sorted_right_kernel / sorted_left_kernel (rfx_graph.cu)
template <class KT> static int64_t sorted_impl(const std::map<u128, uint32_t>& tab, int k, int E, double fold, int X) {
    g_okeys.clear(); g_oleft.clear(); g_oright.clear();
    auto cnt_of = [&](KT Z) -> uint32_t {
        KT zc = revcomp(Z, k);
        auto it = tab.find((u128)(zc < Z ? zc : Z));
        return it == tab.end() ? 0u : it->second;
    };
    std::map<u128, std::pair<int32_t, int32_t>> rsurv;  // survivors of the right filter: (coverage carried on, right flag)
    for (auto& kv : tab) {
        KT key = (KT)kv.first, rc = revcomp(key, k);
        for (int strand = 0; strand < 2; strand++) {
            if (strand && rc == key) continue;
            KT Xk = strand ? rc : key;
            KT prefix = Xk >> 2;
            uint32_t myb = (uint32_t)Xk & 3u, cnt[4];
            bool dup[4];
            for (uint32_t b = 0; b < 4; b++) {
                KT Z = (prefix << 2) | (KT)b;
                cnt[b] = b == myb ? kv.second : cnt_of(Z);
                dup[b] = (Z == revcomp(Z, k));
            }
            SortedFork res = sorted_right_fork(cnt, dup, E, fold, X);
            if (res.winner == (int)myb) rsurv[(u128)Xk] = std::make_pair(res.left, res.right);
        }
    }
    const int top = 2 * (k - 1);
    const KT sufmask = mask_bases<KT>(k - 1);
    for (auto& kv : rsurv) {
        KT Xk = (KT)kv.first;
        KT suffix = Xk & sufmask;
        uint32_t mya = (uint32_t)(Xk >> top) & 3u;
        int32_t cov[4], rfl[4];
        for (uint32_t a = 0; a < 4; a++) {
            KT Z = ((KT)a << top) | suffix;
            auto it = rsurv.find((u128)Z);
            cov[a] = it == rsurv.end() ? 0 : it->second.first;
            rfl[a] = it == rsurv.end() ? 0 : it->second.second;
        }
        SortedFork res = sorted_left_fork(cov, rfl, E, fold, X);
        if (res.winner == (int)mya) { g_okeys.push_back((u128)Xk); g_oleft.push_back(res.left); g_oright.push_back(res.right); }
    }
    return (int64_t)g_okeys.size();
}

extern "C" {

int64_t emu_fork_filter(const uint64_t* hi, const uint64_t* lo, const uint32_t* cnt, int64_t n, int k, int E) {
    std::map<u128, uint32_t> tab;
    for (int64_t i = 0; i < n; i++) tab[((u128)hi[i] << 64) | lo[i]] = cnt[i];
    return k <= 31 ? fork_impl<uint64_t>(tab, k, E) : fork_impl<u128>(tab, k, E);
}
int64_t emu_sorted_filter(const uint64_t* hi, const uint64_t* lo, const uint32_t* cnt, int64_t n, int k, int E, double fold, int X) {
    std::map<u128, uint32_t> tab;
    for (int64_t i = 0; i < n; i++) tab[((u128)hi[i] << 64) | lo[i]] = cnt[i];
    return k <= 31 ? sorted_impl<uint64_t>(tab, k, E, fold, X) : sorted_impl<u128>(tab, k, E, fold, X);
}
void emu_fork_fetch(uint64_t* hi, uint64_t* lo, int32_t* left, int32_t* right) {
    for (size_t i = 0; i < g_okeys.size(); i++) {
        hi[i] = (uint64_t)(g_okeys[i] >> 64); lo[i] = (uint64_t)g_okeys[i]; left[i] = g_oleft[i]; right[i] = g_oright[i];
    }
}

// rec_kmer_at (the counting kernel's per-lane cut) against the rolling extraction, k-mer by k-mer; returns mismatches
struct VecSink {
    std::vector<u128>* v;
    void operator()(uint64_t key) const { v->push_back((u128)key); }
    void operator()(u128 key) const { v->push_back(key); }
};
int64_t emu_check_kmer_at(const uint64_t* recs, int64_t n_rec, int k) {
    const int recw = rec_words_for_k(k);
    int64_t bad = 0;
    std::vector<u128> ref;
    for (int64_t r = 0; r < n_rec; r++) {
        const uint64_t* rec = recs + r * recw;
        ref.clear();
        const uint32_t nk = (uint32_t)(rec[0] >> 48);
        if (recw == 2) {
            rec_foreach_kmer<uint64_t, 2>(rec, k, VecSink{&ref});
            const uint64_t w[2] = {rec[0], rec[1]};
            for (uint32_t off = 0; off < nk; off++) {
                const uint64_t f = rec_kmer_at<uint64_t, 2>(w, off, k), rc = revcomp(f, k);
                if (off >= ref.size() || ref[off] != (u128)(f < rc ? f : rc)) bad++;
            }
        } else {
            rec_foreach_kmer<u128, 4>(rec, k, VecSink{&ref});
            const uint64_t w[4] = {rec[0], rec[1], rec[2], rec[3]};
            for (uint32_t off = 0; off < nk; off++) {
                const u128 f = rec_kmer_at<u128, 4>(w, off, k), rc = revcomp(f, k);
                if (off >= ref.size() || ref[off] != (f < rc ? f : rc)) bad++;
            }
        }
        if (ref.size() != nk) bad++;
    }
    return bad;
}
// the table hashes: out[0] = lane a, out[1] = lane b of the 128-bit key, out[2] = narrow_hash(lo), out[3] = its class bits
void emu_table_hashes(uint64_t hi, uint64_t lo, uint32_t* out) {
    out[0] = hash_lane_a((uint32_t)hi, (uint32_t)(hi >> 32), (uint32_t)lo, (uint32_t)(lo >> 32));
    out[1] = hash_lane_b((uint32_t)hi, (uint32_t)(hi >> 32), (uint32_t)lo, (uint32_t)(lo >> 32));
    uint32_t cls = 0;
    out[2] = narrow_hash(lo, cls);
    out[3] = cls;
}

uint64_t emu_revcomp64(uint64_t x, int nb) { return revcomp(x, nb); }
void emu_revcomp128(uint64_t hi, uint64_t lo, int nb, uint64_t* ohi, uint64_t* olo) {
    u128 r = revcomp(((u128)hi << 64) | lo, nb);
    *ohi = (uint64_t)(r >> 64); *olo = (uint64_t)r;
}

}  // extern "C"
