/*
 * reflexiv_oracle.c -- CPU restatement of Reflexiv's k-mer counting and
 * reflexible contig extension.  TEST INFRASTRUCTURE ONLY (see the header).
 *
 * Citations are relative to
 *   /root/reference/src/main/java/uni/bielefeld/cmg/reflexiv/
 * ("DSMain" = pipeline/ReflexivDSMain.java, "Counter" =
 * pipeline/ReflexivDataFrameCounter.java, "Counter64" =
 * pipeline/ReflexivDataFrameCounter64.java).
 *
 * Where the reference's result depends on the arrival order Spark's shuffle
 * happens to deliver (fork groups with >= 3 members, equal-count ties in the
 * left fork filter, cycles, budget flags), this file fixes ONE realisable
 * order and says so at the spot ("CANONICAL ORDER").  libreflexiv_cuda matches
 * the same choices; DESIGN.md lists them.
 */
#define _GNU_SOURCE
#include "reflexiv_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

static inline u128 mk128(uint64_t hi, uint64_t lo) { return ((u128)hi << 64) | lo; }
static inline u128 mask_bases(int nb) { return nb >= 64 ? ~(u128)0 : (((u128)1 << (2 * nb)) - 1); }

/* nucleotideValue: A=0 C=1 G=2, everything else (T, N, lower case ...) = 3.
 * Counter:513-525, DSMain:4010-4022. */
static inline unsigned base_code(unsigned char c) {
    return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : 3u;
}

static u128 revcomp(u128 x, int k) {
    /* DSKmerReverseComplementLong, DSMain:3857-3863 */
    u128 r = 0;
    for (int i = 0; i < k; i++) {
        r = (r << 2) | ((x & 3) ^ 3);
        x >>= 2;
    }
    return r;
}

void orc_free(void *p) { free(p); }

/* ---- host threads for the sort-based stages ------------------------------ */
/* The reference runs every sort("k-1") as a range shuffle over all executor
 * cores and every mapPartitions over one partition per core.  With
 * orc_set_threads(T > 1) the fork filters sort with T threads and the
 * extension passes (ORC_ASM_REFSIM) sort with T threads and scan T range
 * partitions concurrently -- a partition never splits a run of equal keys,
 * exactly like Spark's RangePartitioner.  T = 1 (default) is the
 * deterministic single-partition order the tests use. */
static int g_threads = 1;
void orc_set_threads(int n) { g_threads = n > 1 ? n : 1; }

/* qsort on T chunks in parallel, then pairwise merges (parallel across pairs) */
static void par_qsort(void *base, size_t n, size_t size, int (*cmp)(const void *, const void *)) {
    int T = g_threads;
    if (T <= 1 || n < 65536) { qsort(base, n, size, cmp); return; }
    while ((size_t)T * 4096 > n) T /= 2;
    size_t *cut = (size_t *)malloc((size_t)(T + 1) * sizeof(size_t));
    for (int t = 0; t <= T; t++) cut[t] = n * (size_t)t / (size_t)T;
    char *a = (char *)base;
#pragma omp parallel for schedule(dynamic, 1) num_threads(T)
    for (int t = 0; t < T; t++) qsort(a + cut[t] * size, cut[t + 1] - cut[t], size, cmp);
    char *tmp = (char *)malloc(n * size);
    char *src = a, *dst = tmp;
    for (int width = 1; width < T; width *= 2) {
        int n_pairs = (T + 2 * width - 1) / (2 * width);
#pragma omp parallel for schedule(dynamic, 1) num_threads(T)
        for (int pi = 0; pi < n_pairs; pi++) {
            int lo = pi * 2 * width, mid = lo + width < T ? lo + width : T, hi = lo + 2 * width < T ? lo + 2 * width : T;
            size_t i = cut[lo], ie = cut[mid], j = cut[mid], je = cut[hi], o = cut[lo];
            while (i < ie && j < je) {
                if (cmp(src + j * size, src + i * size) < 0) { memcpy(dst + o * size, src + j * size, size); j++; }
                else { memcpy(dst + o * size, src + i * size, size); i++; }
                o++;
            }
            if (i < ie) memcpy(dst + o * size, src + i * size, (ie - i) * size);
            if (j < je) memcpy(dst + o * size, src + j * size, (je - j) * size);
        }
        char *sw = src; src = dst; dst = sw;
    }
    if (src != a) memcpy(a, src, n * size);
    free(tmp); free(cut);
}

/* ------------------------------------------------------------------------ */
/* A1 / A1': FASTQ line filters                                              */
/* ------------------------------------------------------------------------ */

static int is_atcgn(char a) { /* checkSeq, Counter:270-289 */
    return a == 'A' || a == 'T' || a == 'C' || a == 'G' || a == 'N';
}

int64_t orc_fastq_reads(const char *txt, size_t n, int mode, uint64_t **starts_out, uint32_t **lens_out) {
    size_t cap = 1024, cnt = 0;
    uint64_t *starts = (uint64_t *)malloc(cap * sizeof(uint64_t));
    uint32_t *lens = (uint32_t *)malloc(cap * sizeof(uint32_t));
    int line_mark = 0;         /* DSFastqFilterWithQual.lineMark, DSMain:4050 */
    uint64_t pend_start = 0;   /* the sequence line of the unit being assembled */
    uint32_t pend_len = 0;
    size_t pos = 0;
    while (pos < n) {
        /* spark.read().text(): one record per line, line terminator stripped
         * (Hadoop LineRecordReader: \n, \r\n).  A final line without a
         * terminator is still a record. */
        const char *nl = (const char *)memchr(txt + pos, '\n', n - pos);
        size_t end = nl ? (size_t)(nl - txt) : n;
        size_t len = end - pos;
        if (len > 0 && txt[pos + len - 1] == '\r') len--;
        const char *s = txt + pos;
        int emit = 0;
        uint64_t e_start = 0;
        uint32_t e_len = 0;
        if (mode == ORC_FASTQ_RUN) {
            /* DSMain:4051-4071 -- branch order matters */
            if (line_mark == 2) {
                line_mark = 3;
            } else if (line_mark == 3) {
                line_mark = 4;
                emit = 1; e_start = pend_start; e_len = pend_len; /* units[1], DSMain:3964-3965 */
            } else if (len > 0 && s[0] == '@') {
                line_mark = 1;
            } else if (line_mark == 1) {
                line_mark = 2;
                pend_start = pos; pend_len = (uint32_t)len;
            }
        } else if (mode == ORC_FASTQ_COUNTER) {
            /* DSFastqFilterOnlySeq, Counter:247-266 */
            if (len > 20 && s[0] != '@' && s[0] != '+' && is_atcgn(s[0]) && is_atcgn(s[4]) &&
                is_atcgn(s[9]) && is_atcgn(s[14]) && is_atcgn(s[19])) {
                emit = 1; e_start = pos; e_len = (uint32_t)len;
            }
        } else {
            emit = 1; e_start = pos; e_len = (uint32_t)len;
        }
        if (emit) {
            if (cnt == cap) {
                cap *= 2;
                starts = (uint64_t *)realloc(starts, cap * sizeof(uint64_t));
                lens = (uint32_t *)realloc(lens, cap * sizeof(uint32_t));
            }
            starts[cnt] = e_start; lens[cnt] = e_len; cnt++;
        }
        pos = end + 1;
    }
    *starts_out = starts;
    *lens_out = lens;
    return (int64_t)cnt;
}

/* ------------------------------------------------------------------------ */
/* A2 + A3 + A4: canonical k-mer extraction, counting, coverage filter       */
/* ------------------------------------------------------------------------ */

/* Does the reference keep this read?  k<=31: Counter:471 / DSMain:3968
 * ("readLength - k - endClip <= 1" drops reads of length k and k+1);
 * k>31: Counter64:410 ("readLength - k - endClip + 1 <= 0"). */
static inline int read_is_kept(int64_t len, int k, int front_clip, int end_clip) {
    if (front_clip > len) return 0;
    if (k <= 31) return !(len - k - end_clip <= 1);
    return !(len - k - end_clip + 1 <= 0);
}

static inline int64_t read_kmer_count(int64_t len, int k, int front_clip, int end_clip) {
    if (!read_is_kept(len, k, front_clip, end_clip)) return 0;
    int64_t m = len - end_clip - front_clip - k + 1;
    return m > 0 ? m : 0;
}

/* splitmix64 finaliser: only used to spread keys over CPU partitions. */
static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}

static void radix_sort_u64(uint64_t *a, uint64_t *tmp, size_t n, int key_bits) {
    for (int shift = 0; shift < key_bits; shift += 8) {
        size_t hist[256] = {0};
        for (size_t i = 0; i < n; i++) hist[(a[i] >> shift) & 255]++;
        size_t sum = 0, single = 0;
        for (int b = 0; b < 256; b++) { size_t c = hist[b]; if (c == n) single = 1; hist[b] = sum; sum += c; }
        if (single) continue;
        for (size_t i = 0; i < n; i++) tmp[hist[(a[i] >> shift) & 255]++] = a[i];
        memcpy(a, tmp, n * sizeof(uint64_t));
    }
}

static void radix_sort_u128(u128 *a, u128 *tmp, size_t n, int key_bits) {
    for (int shift = 0; shift < key_bits; shift += 8) {
        size_t hist[256] = {0};
        for (size_t i = 0; i < n; i++) hist[(unsigned)(a[i] >> shift) & 255]++;
        size_t sum = 0, single = 0;
        for (int b = 0; b < 256; b++) { size_t c = hist[b]; if (c == n) single = 1; hist[b] = sum; sum += c; }
        if (single) continue;
        for (size_t i = 0; i < n; i++) tmp[hist[(unsigned)(a[i] >> shift) & 255]++] = a[i];
        memcpy(a, tmp, n * sizeof(u128));
    }
}

#define ORC_PARTS 256

typedef struct { u128 key; uint32_t count; } kc_t;

static int kc_cmp(const void *a, const void *b) {
    u128 x = ((const kc_t *)a)->key, y = ((const kc_t *)b)->key;
    return x < y ? -1 : x > y ? 1 : 0;
}

/* The rolling extractor of Counter:477-508 (k<=31) and, for k>31, the
 * multi-slot one of Counter64:417-650 restated on one 128-bit integer:
 *   fwd = ((fwd << 2) | v) & mask ; rc = (rc >> 2) | ((v ^ 3) << 2(k-1))
 * emit min(fwd, rc) once k bases are in (Counter:501-506; Counter64:652-687
 * compares base by base = numeric order of equal-length 2-bit strings,
 * forward on a tie).  Calls sink(key) for every instance. */
#define EXTRACT_READ(seq, len, k, fc, ec, KT, SINK)                                   \
    do {                                                                              \
        KT fwd = 0, rc = 0;                                                           \
        const KT msk = (KT)mask_bases(k);                                             \
        for (int64_t i_ = (fc); i_ < (int64_t)(len) - (ec); i_++) {                   \
            KT v = (KT)base_code((unsigned char)(seq)[i_]);                           \
            fwd = ((fwd << 2) | v) & msk;                                             \
            rc = (rc >> 2) | ((v ^ 3) << (2 * ((k)-1)));                              \
            if (i_ - (fc) >= (k)-1) { KT key_ = fwd < rc ? fwd : rc; SINK(key_); }    \
        }                                                                             \
    } while (0)

int64_t orc_count_kmers(const char *txt, const uint64_t *starts, const uint32_t *lens, int64_t n_reads,
                        int k, int front_clip, int end_clip, int64_t min_count, int64_t max_count,
                        int n_threads, uint64_t **keys_hi, uint64_t **keys_lo, uint32_t **counts,
                        int64_t *n_instances, int64_t *n_distinct) {
    if (k < 1 || k > 63) return -1;
    if (n_threads < 1) n_threads = 1;
#ifndef _OPENMP
    n_threads = 1;
#endif
    const int wide = k > 31;
    const size_t ksz = wide ? sizeof(u128) : sizeof(uint64_t);

    /* Pass 1: per (thread, partition) instance counts -> exact buffers.  This
     * mirrors the reference's shape: map side emits every instance, a hash
     * shuffle groups equal keys, a per-partition aggregate counts them
     * (groupBy("value").count(), Counter:198-200). */
    size_t *tp_count = (size_t *)calloc((size_t)n_threads * ORC_PARTS, sizeof(size_t));
#define SINK_COUNT64(key) my[mix64(key) >> 56]++
#define SINK_COUNT128(key) my[mix64((uint64_t)(key) ^ mix64((uint64_t)((key) >> 64))) >> 56]++
#pragma omp parallel for schedule(static, 1) num_threads(n_threads)
    for (int t = 0; t < n_threads; t++) {
        size_t *my = tp_count + (size_t)t * ORC_PARTS;
        int64_t lo = n_reads * t / n_threads, hi = n_reads * (t + 1) / n_threads;
        for (int64_t r = lo; r < hi; r++) {
            if (!read_is_kept(lens[r], k, front_clip, end_clip)) continue;
            const char *seq = txt + starts[r];
            if (!wide) EXTRACT_READ(seq, lens[r], k, front_clip, end_clip, uint64_t, SINK_COUNT64);
            else EXTRACT_READ(seq, lens[r], k, front_clip, end_clip, u128, SINK_COUNT128);
        }
    }
    /* partition-major layout: part p = [thread 0 | thread 1 | ...] */
    size_t *tp_off = (size_t *)malloc(((size_t)n_threads * ORC_PARTS + 1) * sizeof(size_t));
    size_t part_off[ORC_PARTS + 1];
    size_t total = 0;
    for (int p = 0; p < ORC_PARTS; p++) {
        part_off[p] = total;
        for (int t = 0; t < n_threads; t++) {
            tp_off[(size_t)t * ORC_PARTS + p] = total;
            total += tp_count[(size_t)t * ORC_PARTS + p];
        }
    }
    part_off[ORC_PARTS] = total;
    *n_instances = (int64_t)total;

    char *buf = (char *)malloc((total ? total : 1) * ksz);
    char *tmp = (char *)malloc((total ? total : 1) * ksz);
#define SINK_PUT64(key) ((uint64_t *)buf)[my[mix64(key) >> 56]++] = (key)
#define SINK_PUT128(key) ((u128 *)buf)[my[mix64((uint64_t)(key) ^ mix64((uint64_t)((key) >> 64))) >> 56]++] = (key)
#pragma omp parallel for schedule(static, 1) num_threads(n_threads)
    for (int t = 0; t < n_threads; t++) {
        size_t *my = tp_off + (size_t)t * ORC_PARTS;
        int64_t lo = n_reads * t / n_threads, hi = n_reads * (t + 1) / n_threads;
        for (int64_t r = lo; r < hi; r++) {
            if (!read_is_kept(lens[r], k, front_clip, end_clip)) continue;
            const char *seq = txt + starts[r];
            if (!wide) EXTRACT_READ(seq, lens[r], k, front_clip, end_clip, uint64_t, SINK_PUT64);
            else EXTRACT_READ(seq, lens[r], k, front_clip, end_clip, u128, SINK_PUT128);
        }
    }
    free(tp_count);
    free(tp_off);

    /* Pass 2: per-partition sort + run-length count + coverage filter (A4). */
    kc_t *part_out[ORC_PARTS];
    size_t part_n[ORC_PARTS];
    size_t part_distinct[ORC_PARTS];
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (int p = 0; p < ORC_PARTS; p++) {
        size_t lo = part_off[p], n = part_off[p + 1] - lo;
        part_out[p] = NULL; part_n[p] = 0; part_distinct[p] = 0;
        if (n == 0) continue;
        size_t cap = 1024, m = 0, distinct = 0;
        kc_t *out = (kc_t *)malloc(cap * sizeof(kc_t));
        if (!wide) {
            uint64_t *a = (uint64_t *)buf + lo;
            radix_sort_u64(a, (uint64_t *)tmp + lo, n, 2 * k);
            for (size_t i = 0; i < n;) {
                size_t j = i + 1;
                while (j < n && a[j] == a[i]) j++;
                int64_t c = (int64_t)(j - i);
                distinct++;
                if (c >= min_count && c <= max_count) {
                    if (m == cap) { cap *= 2; out = (kc_t *)realloc(out, cap * sizeof(kc_t)); }
                    out[m].key = a[i]; out[m].count = (uint32_t)c; m++;
                }
                i = j;
            }
        } else {
            u128 *a = (u128 *)buf + lo;
            radix_sort_u128(a, (u128 *)tmp + lo, n, 2 * k);
            for (size_t i = 0; i < n;) {
                size_t j = i + 1;
                while (j < n && a[j] == a[i]) j++;
                int64_t c = (int64_t)(j - i);
                distinct++;
                if (c >= min_count && c <= max_count) {
                    if (m == cap) { cap *= 2; out = (kc_t *)realloc(out, cap * sizeof(kc_t)); }
                    out[m].key = a[i]; out[m].count = (uint32_t)c; m++;
                }
                i = j;
            }
        }
        part_out[p] = out; part_n[p] = m; part_distinct[p] = distinct;
    }
    free(buf);
    free(tmp);

    size_t m_total = 0, d_total = 0;
    for (int p = 0; p < ORC_PARTS; p++) { m_total += part_n[p]; d_total += part_distinct[p]; }
    *n_distinct = (int64_t)d_total;
    kc_t *all = (kc_t *)malloc((m_total ? m_total : 1) * sizeof(kc_t));
    size_t w = 0;
    for (int p = 0; p < ORC_PARTS; p++) {
        if (part_n[p]) memcpy(all + w, part_out[p], part_n[p] * sizeof(kc_t));
        w += part_n[p];
        free(part_out[p]);
    }
    /* Row order of the reference's CSV is partition dependent (A5); the
     * oracle returns rows sorted by key so tables compare directly. */
    qsort(all, m_total, sizeof(kc_t), kc_cmp);
    *keys_hi = (uint64_t *)malloc((m_total ? m_total : 1) * sizeof(uint64_t));
    *keys_lo = (uint64_t *)malloc((m_total ? m_total : 1) * sizeof(uint64_t));
    *counts = (uint32_t *)malloc((m_total ? m_total : 1) * sizeof(uint32_t));
    for (size_t i = 0; i < m_total; i++) {
        (*keys_hi)[i] = (uint64_t)(all[i].key >> 64);
        (*keys_lo)[i] = (uint64_t)all[i].key;
        (*counts)[i] = all[i].count;
    }
    free(all);
    return (int64_t)m_total;
}

/* ------------------------------------------------------------------------ */
/* A6 + A7 + A8: both orientations, right fork filter, left fork filter      */
/* ------------------------------------------------------------------------ */

typedef struct {
    u128 key;      /* oriented k-mer */
    int32_t left;  /* count until the left filter overwrites it, then flag */
    int32_t right; /* count until the right filter overwrites it, then flag */
} okmer_t;

static int ok_cmp_key(const void *a, const void *b) {
    u128 x = ((const okmer_t *)a)->key, y = ((const okmer_t *)b)->key;
    return x < y ? -1 : x > y ? 1 : 0;
}

static int g_rot_k; /* qsort has no context argument; single-threaded use only */
static inline u128 rot_key(u128 key, int k) {
    /* sort key of the reflected record: (k-1)-suffix, then first base.
     * DSReflectedSubKmerExtractionFromForward, DSMain:3661-3685, followed by
     * sort("k-1"), DSMain:244.  CANONICAL ORDER: rows that share a suffix are
     * visited in ascending first-base order. */
    return ((key & mask_bases(k - 1)) << 2) | (key >> (2 * (k - 1)));
}
static int ok_cmp_rot(const void *a, const void *b) {
    u128 x = rot_key(((const okmer_t *)a)->key, g_rot_k), y = rot_key(((const okmer_t *)b)->key, g_rot_k);
    return x < y ? -1 : x > y ? 1 : 0;
}

int64_t orc_fork_filter(const uint64_t *keys_hi, const uint64_t *keys_lo, const uint32_t *counts,
                        int64_t n, int k, int min_error_cov, uint64_t **o_hi, uint64_t **o_lo,
                        int32_t **o_left, int32_t **o_right, int64_t *stats) {
    if (k < 2 || k > 63) return -1;
    const int E = min_error_cov;
    const int sub = k - 1; /* param.subKmerSize */
    int64_t st[8] = {0};
    /* A6: DSKmerReverseComplementLong (DSMain:3849-3869) emits the k-mer and
     * its reverse complement (twice the same row for a palindrome, even k),
     * DSForwardSubKmerExtraction (DSMain:3625-3644) sets left = right = count. */
    okmer_t *a = (okmer_t *)malloc((size_t)(n > 0 ? 2 * n : 1) * sizeof(okmer_t));
    for (int64_t i = 0; i < n; i++) {
        u128 key = mk128(keys_hi[i], keys_lo[i]);
        a[2 * i].key = key;
        a[2 * i + 1].key = revcomp(key, k);
        a[2 * i].left = a[2 * i].right = a[2 * i + 1].left = a[2 * i + 1].right = (int32_t)counts[i];
    }
    int64_t m = 2 * n;
    /* sort("k-1") on the (k-1)-prefix, DSMain:232.  CANONICAL ORDER: rows that
     * share a prefix are visited in ascending last-base order (full key order). */
    par_qsort(a, (size_t)m, sizeof(okmer_t), ok_cmp_key);

    /* A7 right fork filter: DSFilterForkSubKmerWithErrorCorrection
     * (DSMain:3431-3483) when minErrorCoverage != 0, DSFilterForkSubKmer
     * (DSMain:3375-3417) when it is 0 (DSMain:233-239). */
    int64_t w = 0;
    for (int64_t i = 0; i < m;) {
        int64_t j = i + 1;
        u128 prefix = a[i].key >> 2;
        while (j < m && (a[j].key >> 2) == prefix) j++;
        okmer_t cur = a[i];
        cur.right = E ? -1 - cur.left : -1;
        if (j - i >= 2) st[0]++;
        if (j - i >= 3) st[1]++;
        for (int64_t x = i + 1; x < j; x++) {
            int cx = a[x].left, cc = cur.left;
            if (cx > cc) {
                int flag = sub;
                if (E && cc <= E && cx >= 2 * cc) flag = -1 - cx;
                cur = a[x]; cur.right = flag;
            } else if (cx == cc) {
                st[2]++;
                if ((a[x].key & 3) > (cur.key & 3)) cur = a[x];
                cur.right = sub;
            } else {
                if (E && cx <= E && cc >= 2 * cx) cur.right = -1 - cc;
                else cur.right = sub;
            }
        }
        a[w++] = cur;
        i = j;
    }
    m = w;

    /* A8: re-key on the suffix (DSMain:3661-3685), sort, left fork filter
     * DSFilterForkReflectedSubKmerWithErrorCorrection (DSMain:3550-3616) /
     * DSFilterForkReflectedSubKmer (DSMain:3489-3541). */
    g_rot_k = k;
    par_qsort(a, (size_t)m, sizeof(okmer_t), ok_cmp_rot);
    const u128 sufmask = mask_bases(k - 1);
    w = 0;
    for (int64_t i = 0; i < m;) {
        int64_t j = i + 1;
        u128 suffix = a[i].key & sufmask;
        while (j < m && (a[j].key & sufmask) == suffix) j++;
        okmer_t cur = a[i];
        int H = cur.left; /* HighCoverLastCoverage */
        cur.left = E ? -1 - H : -1;
        if (j - i >= 2) st[3]++;
        if (j - i >= 3) st[4]++;
        for (int64_t x = i + 1; x < j; x++) {
            int cx = a[x].left;
            if (cx > H) {
                int flag = sub;
                if (E && H <= E && cx >= 2 * H) flag = -1 - cx;
                H = cx;
                cur = a[x]; cur.left = flag;
            } else if (cx == H) {
                st[5]++;
                /* DSMain:3573-3578: the stored row's extension (4|base) is
                 * shifted one base too far, so "4|base" is compared with 1
                 * and the arriving row always wins. */
                uint64_t ext_x = 4u | (uint64_t)(a[x].key >> (2 * (k - 1)));
                uint64_t ext_c = 4u | (uint64_t)(cur.key >> (2 * (k - 1)));
                uint64_t first_x = ext_x >> (2 * (1 - 1));
                uint64_t first_c = ext_c >> (2 * 1);
                if (first_x > first_c) cur = a[x];
                cur.left = sub;
            } else {
                if (E && cx <= E && H >= 2 * cx) { /* stored row re-emitted unchanged, DSMain:3591-3596 */ }
                else cur.left = sub;
            }
        }
        a[w++] = cur;
        i = j;
    }
    m = w;
    par_qsort(a, (size_t)m, sizeof(okmer_t), ok_cmp_key);
    *o_hi = (uint64_t *)malloc((size_t)(m ? m : 1) * sizeof(uint64_t));
    *o_lo = (uint64_t *)malloc((size_t)(m ? m : 1) * sizeof(uint64_t));
    *o_left = (int32_t *)malloc((size_t)(m ? m : 1) * sizeof(int32_t));
    *o_right = (int32_t *)malloc((size_t)(m ? m : 1) * sizeof(int32_t));
    for (int64_t i = 0; i < m; i++) {
        (*o_hi)[i] = (uint64_t)(a[i].key >> 64);
        (*o_lo)[i] = (uint64_t)a[i].key;
        (*o_left)[i] = a[i].left;
        (*o_right)[i] = a[i].right;
        if (a[i].left >= 0) st[6]++;
        if (a[i].right >= 0) st[7]++;
    }
    free(a);
    if (stats) memcpy(stats, st, sizeof(st));
    return m;
}

/* ------------------------------------------------------------------------ */
/* Count_<k>_sorted, the "left and right sorting" stage          */
/* pipeline/ReflexivDSKmerLeftAndRightSorting.java:105-243 (LRS below)         */
/* ------------------------------------------------------------------------ */

/* buildingAlongFromThreeInt, LRS:596-622: marker in the top two bits, the two
 * covers as non-negative 32-bit halves (a negative value v is stored as
 * 30000 - v, magnitudes saturate at 30000). */
static int64_t lrs_build(int marker, int left, int right) {
    int64_t info = (int64_t)((uint64_t)marker << 62);
    if (left >= 30000) left = 30000; else if (left <= -30000) left = 30000 - (-30000); else if (left < 0) left = 30000 - left;
    if (right >= 30000) right = 30000; else if (right <= -30000) right = 30000 - (-30000); else if (right < 0) right = 30000 - right;
    info |= (int64_t)left << 32;
    info |= (int64_t)right;
    return info;
}
static int lrs_marker(int64_t a) { return (int)((uint64_t)a >> 62); }            /* getReflexivMarker, LRS:569-572 */
static int lrs_left(int64_t a) {                                                   /* getLeftMarker, LRS:574-584 */
    int v = (int)((uint64_t)a >> 32);
    v &= ~(3 << 30);
    if (v > 30000) v = 30000 - v;
    return v;
}
static int lrs_right(int64_t a) {                                                  /* getRightMarker, LRS:586-594 */
    int v = (int)a;
    if (v > 30000) v = 30000 - v;
    return v;
}

typedef struct { u128 key; int64_t attr; } lrs_t;
static int lrs_cmp_key(const void *a, const void *b) {
    u128 x = ((const lrs_t *)a)->key, y = ((const lrs_t *)b)->key;
    return x < y ? -1 : x > y ? 1 : 0;
}
static int lrs_cmp_rot(const void *a, const void *b) {
    u128 x = rot_key(((const lrs_t *)a)->key, g_rot_k), y = rot_key(((const lrs_t *)b)->key, g_rot_k);
    return x < y ? -1 : x > y ? 1 : 0;
}

/* Input: the (k-mer, count) rows of a Count_<k> table.  Output: the rows of Count_<k>_sorted as (oriented k-mer,
 * left, right), sorted by key; the reference prints them as `KMER,1|left|right` (LRS:249-274).  Rows above
 * max_cov are dropped (LRS:186-193: the lower bound is commented out there).  X = max_kmer_size + 3 with
 * max_kmer_size = param.kmerListInt[last] (LRS:445).  E == 0 selects DSFilterForkSubKmer, which reads a five-column
 * row out of a three-column dataset and cannot run in the reference: rejected.  k-mers with (k-1) % 31 == 0 lose
 * their last base in DSForwardSubKmerExtraction (LRS:930-936 takes the wrong block): rejected. */
int64_t orc_sorted_rows(const uint64_t *keys_hi, const uint64_t *keys_lo, const uint32_t *counts, int64_t n, int k,
                        int min_error_cov, double min_repeat_fold, int max_kmer_size, int64_t max_cov,
                        uint64_t **o_hi, uint64_t **o_lo, int32_t **o_left, int32_t **o_right) {
    if (k < 2 || k > 63 || min_error_cov == 0 || (k - 1) % 31 == 0) return -1;
    const int E = min_error_cov, X = max_kmer_size + 3;
    const double F = min_repeat_fold;
    lrs_t *a = (lrs_t *)malloc((size_t)(n > 0 ? 2 * n : 1) * sizeof(lrs_t));
    int64_t m = 0;
    for (int64_t i = 0; i < n; i++) {
        if ((int64_t)counts[i] > max_cov) continue;
        /* DSKmerReverseComplement, LRS:1583-1632: the k-mer, then its reverse complement, same cover;
         * DSForwardSubKmerExtraction, LRS:915-978: attribute (1, cover, cover) */
        u128 key = mk128(keys_hi[i], keys_lo[i]);
        a[m].key = key; a[m].attr = lrs_build(1, (int)counts[i], (int)counts[i]); m++;
        a[m].key = revcomp(key, k); a[m].attr = a[m - 1].attr; m++;
    }
    /* sort("k-1"), LRS:202; canonical order inside a group: ascending last base */
    qsort(a, (size_t)m, sizeof(lrs_t), lrs_cmp_key);
    /* DSFilterForkSubKmerWithErrorCorrection, LRS:432-537 */
    int64_t w = 0;
    for (int64_t i = 0; i < m;) {
        int64_t j = i + 1;
        u128 prefix = a[i].key >> 2;
        while (j < m && (a[j].key >> 2) == prefix) j++;
        lrs_t cur = a[i];
        cur.attr = lrs_build(lrs_marker(a[i].attr), lrs_left(a[i].attr), -1);
        for (int64_t x = i + 1; x < j; x++) {
            int marker = lrs_marker(a[x].attr), leftMarker = lrs_left(a[x].attr);
            int highest = lrs_left(cur.attr);
            if (leftMarker > highest) {
                int64_t attr;
                if (highest <= E && leftMarker >= F * highest) attr = lrs_build(marker, leftMarker, -1);
                else if (highest == 1) attr = lrs_build(marker, leftMarker, -1);
                else attr = lrs_build(marker, leftMarker, max_kmer_size + 3);
                cur = a[x]; cur.attr = attr;
            } else if (leftMarker == highest) {
                /* extension = last base left-aligned + end marker: the larger base wins (LRS:473) */
                if ((a[x].key & 3) > (cur.key & 3)) {
                    cur = a[x];
                    cur.attr = lrs_build(marker, leftMarker, highest == 1 ? -1 : X);
                } else {
                    cur.attr = lrs_build(lrs_marker(cur.attr), lrs_left(cur.attr), highest == 1 ? -1 : X);
                }
            } else {
                if (leftMarker <= E && highest >= F * leftMarker) {
                    cur.attr = lrs_build(lrs_marker(cur.attr), lrs_left(cur.attr), -1);
                } else {
                    /* LRS:504-518: the attribute is built from the ARRIVING row's markers before the stored row is
                     * read back, so the stored k-mer continues with the weaker row's coverage */
                    cur.attr = lrs_build(marker, leftMarker, leftMarker == 1 ? -1 : X);
                }
            }
        }
        a[w++] = cur;
        i = j;
    }
    m = w;
    /* DSReflectedSubKmerExtractionFromForward, LRS:1109-1178: key on the suffix, marker 2; sort("k-1"), LRS:217 */
    for (int64_t i = 0; i < m; i++) a[i].attr = lrs_build(2, lrs_left(a[i].attr), lrs_right(a[i].attr));
    g_rot_k = k;
    qsort(a, (size_t)m, sizeof(lrs_t), lrs_cmp_rot);
    /* DSFilterForkReflectedSubKmerWithErrorCorrection, LRS:700-818 */
    const u128 sufmask = mask_bases(k - 1);
    w = 0;
    for (int64_t i = 0; i < m;) {
        int64_t j = i + 1;
        u128 suffix = a[i].key & sufmask;
        while (j < m && (a[j].key & sufmask) == suffix) j++;
        lrs_t cur = a[i];
        int HC = lrs_left(a[i].attr); /* HighCoverLastCoverage */
        cur.attr = lrs_build(lrs_marker(a[i].attr), -1, lrs_right(a[i].attr));
        for (int64_t x = i + 1; x < j; x++) {
            int marker = lrs_marker(a[x].attr), leftMarker = lrs_left(a[x].attr), rightMarker = lrs_right(a[x].attr);
            if (leftMarker > HC) {
                int64_t attr;
                if (HC <= E && leftMarker >= F * HC) attr = lrs_build(marker, -1, rightMarker);
                else attr = lrs_build(marker, X, rightMarker);
                HC = leftMarker;
                cur = a[x]; cur.attr = attr;
            } else if (leftMarker == HC) {
                unsigned fx = (unsigned)(a[x].key >> (2 * (k - 1))) & 3u, fc = (unsigned)(cur.key >> (2 * (k - 1))) & 3u;
                if (((fx << 2) | 1u) > ((fc << 2) | 1u)) { /* first base + end marker, LRS:746-751 */
                    cur = a[x];
                    cur.attr = lrs_build(marker, X, HC == 1 ? -1 : rightMarker);
                } else {
                    cur.attr = lrs_build(lrs_marker(cur.attr), X, lrs_right(cur.attr));
                }
            } else {
                if (leftMarker <= E && HC >= F * leftMarker) {
                    cur.attr = lrs_build(lrs_marker(cur.attr), -1, lrs_right(cur.attr));
                } else {
                    /* LRS:790-806: built from the arriving row's right marker, then stored on the kept k-mer */
                    cur.attr = lrs_build(marker, X, leftMarker == 1 ? -1 : rightMarker);
                }
            }
        }
        a[w++] = cur;
        i = j;
    }
    m = w;
    /* DSSubKmerToFullKmer, LRS:1321-1346: full k-mer, marker 1 */
    qsort(a, (size_t)m, sizeof(lrs_t), lrs_cmp_key);
    *o_hi = (uint64_t *)malloc((size_t)(m ? m : 1) * sizeof(uint64_t));
    *o_lo = (uint64_t *)malloc((size_t)(m ? m : 1) * sizeof(uint64_t));
    *o_left = (int32_t *)malloc((size_t)(m ? m : 1) * sizeof(int32_t));
    *o_right = (int32_t *)malloc((size_t)(m ? m : 1) * sizeof(int32_t));
    for (int64_t i = 0; i < m; i++) {
        (*o_hi)[i] = (uint64_t)(a[i].key >> 64);
        (*o_lo)[i] = (uint64_t)a[i].key;
        (*o_left)[i] = lrs_left(a[i].attr);
        (*o_right)[i] = lrs_right(a[i].attr);
    }
    free(a);
    return m;
}

/* ------------------------------------------------------------------------ */
/* A9 + A10: extension                                                       */
/* ------------------------------------------------------------------------ */

void orc_contigs_free(orc_contigs *c) {
    free(c->offsets); free(c->bases); free(c->left); free(c->right);
    memset(c, 0, sizeof(*c));
}

typedef struct {
    size_t n, cap_n, nb, cap_b;
    orc_contigs *c;
} cbuild_t;

static void cb_init(cbuild_t *b, orc_contigs *c) {
    memset(c, 0, sizeof(*c));
    b->n = 0; b->cap_n = 16; b->nb = 0; b->cap_b = 4096; b->c = c;
    c->offsets = (uint64_t *)malloc((b->cap_n + 1) * sizeof(uint64_t));
    c->left = (int32_t *)malloc(b->cap_n * sizeof(int32_t));
    c->right = (int32_t *)malloc(b->cap_n * sizeof(int32_t));
    c->bases = (char *)malloc(b->cap_b);
    c->offsets[0] = 0;
}

static char *cb_begin(cbuild_t *b, size_t len, int32_t left, int32_t right) {
    orc_contigs *c = b->c;
    if (b->n == b->cap_n) {
        b->cap_n *= 2;
        c->offsets = (uint64_t *)realloc(c->offsets, (b->cap_n + 1) * sizeof(uint64_t));
        c->left = (int32_t *)realloc(c->left, b->cap_n * sizeof(int32_t));
        c->right = (int32_t *)realloc(c->right, b->cap_n * sizeof(int32_t));
    }
    while (b->nb + len > b->cap_b) { b->cap_b *= 2; c->bases = (char *)realloc(c->bases, b->cap_b); }
    char *dst = c->bases + b->nb;
    c->left[b->n] = left; c->right[b->n] = right;
    b->nb += len; b->n++;
    c->offsets[b->n] = b->nb;
    c->n_contigs = (int64_t)b->n;
    return dst;
}

/* DSKmerToContig, DSMain:743-771: drop if both flags <= -10^7, keep if
 * length >= minContig. */
static inline int contig_kept(int64_t len, int32_t left, int32_t right, int min_contig) {
    if (left <= -10000000 && right <= -10000000) return 0;
    return len >= min_contig;
}

static const char ACGT[4] = {'A', 'C', 'G', 'T'};

/* lower_bound on sorted keys */
static int64_t lb_key(const u128 *keys, int64_t n, u128 x) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (keys[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}

/* ---- junctions and budget flags ------------------------------------------
 * A junction X -> Y (suffix(X) == prefix(Y)) is joined by the scan of
 * DSExtendReflexivKmer*.call (DSMain:3069-3075 / 1809-1815) when the forward
 * record's left flag and the reflected record's right flag are both negative
 * or both non-negative (clauses 1 and 2).  A fork winner's flag k-1 facing a
 * clean end is a BUDGET (clauses 3 / 4, DSMain:3077-3084, reflexivExtend
 * :3237-3325): the flagged fragment absorbs the clean neighbour while the
 * neighbour's extension is no longer than the budget; the budget shrinks by
 * the bases absorbed and OVERWRITES the flag of the new outer end.  How much a
 * fork winner absorbs therefore depends on how long its neighbour already is
 * when Spark happens to co-present the pair -- anything from 0 to k-1 k-mers.
 *
 * CANONICAL ORDER: flagged ends grow one k-mer at a time before
 * any clean-clean join is made, the walker further upstream in its own
 * direction first.  Along a path this is the recurrence
 *     E(v) = (face(v) < 0 && E(prev(v)) >= 1) ? E(prev(v)) - 1 : budget(v)
 * evaluated once along succ (budget = right flag, face = left flag) and once
 * along pred (budget = left flag, face = right flag): a walk absorbs at most
 * k-1 k-mers, ends in front of a non-negative facing flag (clause 2 joins
 * there), and an absorbed fork winner loses its own budget exactly like the
 * overwritten flag in reflexivExtend.  A junction is joined when a walk went
 * through it or when the two effective flags have the same sign.
 * assemble_scheduled() below applies the reference's four clauses literally
 * under that schedule; tests check that both give the same contigs. */
static inline int junction_joins(int32_t x_right, int32_t y_left) {
    return (y_left < 0 && x_right < 0) || (y_left >= 0 && x_right >= 0);
}

/* start of the recurrence on a closed raw path L[0..m-1] (direction order): a node that cannot be absorbed
 * (facing flag >= 0), else the node behind >= bmax consecutive budget-less nodes (no walk survives them), else --
 * every walk may wrap around -- CANONICAL ORDER: the fork winner with the smallest oriented k-mer starts fresh.
 * Returns -1 when no node of the loop carries a budget. */
static int64_t loop_start(const int64_t *L, int64_t m, const int32_t *bud, const int32_t *face, const u128 *keys, int bmax) {
    int64_t n_cand = 0;
    for (int64_t i = 0; i < m; i++) if (bud[L[i]] >= 0) n_cand++;
    if (!n_cand) return -1;
    for (int64_t i = 0; i < m; i++) if (face[L[i]] >= 0) return i;
    int64_t run = 0;
    for (int64_t i = 0; i < 2 * m; i++) {
        if (bud[L[i % m]] < 0) { if (++run >= bmax) return (i + 1) % m; }
        else run = 0;
    }
    int64_t best = -1;
    for (int64_t i = 0; i < m; i++)
        if (bud[L[i]] >= 0 && (best < 0 || keys[L[i]] < keys[L[best]])) best = i;
    return best;
}

/* one direction of the recurrence.  nxt / prv: raw links along the walking direction (self loops removed);
 * eff[] comes in as a copy of bud[]; bit: 1 = walks along succ, 2 = along pred (jmark is indexed by the LEFT node
 * of the junction in sequence order). */
static void budget_scan(int bit, int64_t n, const int64_t *nxt, const int64_t *prv, const int32_t *bud, const int32_t *face,
                        const u128 *keys, int bmax, int32_t *eff, uint8_t *jmark, int64_t *absorbed) {
    char *seen = (char *)calloc((size_t)(n ? n : 1), 1);
    int64_t *L = (int64_t *)malloc((size_t)(n ? n : 1) * sizeof(int64_t));
    for (int64_t h = 0; h < n; h++) {
        if (prv[h] >= 0) continue;
        seen[h] = 1;
        for (int64_t p = h, v = nxt[h]; v >= 0; p = v, v = nxt[v]) {
            seen[v] = 1;
            if (face[v] < 0 && eff[p] >= 1) { eff[v] = eff[p] - 1; jmark[bit == 1 ? p : v] |= (uint8_t)bit; (*absorbed)++; }
        }
    }
    for (int64_t v0 = 0; v0 < n; v0++) {
        if (seen[v0]) continue;
        int64_t m = 0;
        for (int64_t v = v0; !seen[v]; v = nxt[v]) { seen[v] = 1; L[m++] = v; }
        const int64_t st = loop_start(L, m, bud, face, keys, bmax);
        if (st < 0) continue;
        for (int64_t i = 1; i < m; i++) {
            const int64_t p = L[(st + i - 1) % m], v = L[(st + i) % m];
            if (face[v] < 0 && eff[p] >= 1) { eff[v] = eff[p] - 1; jmark[bit == 1 ? p : v] |= (uint8_t)bit; (*absorbed)++; }
        }
    }
    free(seen); free(L);
}

/* raw neighbour links of the filter output (every (k-1)-mer has in- and out-degree <= 1); a k-mer whose suffix is
 * its own prefix never meets itself (one record): no link, reported through self_loop[]. */
static int raw_links(const u128 *keys, int64_t n, int k, int64_t *succ, int64_t *pred, char *self_loop) {
    const u128 sufmask = mask_bases(k - 1);
    for (int64_t i = 0; i < n; i++) { pred[i] = -1; self_loop[i] = 0; }
    int bad = 0;
    for (int64_t i = 0; i < n; i++) {
        u128 lo = (keys[i] & sufmask) << 2;
        int64_t p = lb_key(keys, n, lo);
        succ[i] = -1;
        if (p < n && (keys[p] >> 2) == (lo >> 2)) {
            if (p + 1 < n && (keys[p + 1] >> 2) == (lo >> 2)) bad = 1; /* out-degree > 1: not a filter output */
            if (p == i) self_loop[i] = 1; else succ[i] = p;
        }
    }
    for (int64_t i = 0; i < n && !bad; i++) {
        if (succ[i] >= 0) {
            if (pred[succ[i]] >= 0) bad = 1; /* in-degree > 1 */
            pred[succ[i]] = i;
        }
    }
    return bad;
}

static int assemble_canonical(const u128 *keys, const int32_t *left, const int32_t *right, int64_t n,
                              int k, int min_contig, orc_contigs *out) {
    const size_t nn = (size_t)(n ? n : 1);
    int64_t *succ = (int64_t *)malloc(nn * sizeof(int64_t));
    int64_t *pred = (int64_t *)malloc(nn * sizeof(int64_t));
    char *self_loop = (char *)malloc(nn);
    if (raw_links(keys, n, k, succ, pred, self_loop)) { free(succ); free(pred); free(self_loop); return -2; }
    /* budget walks */
    int32_t *effL = (int32_t *)malloc(nn * sizeof(int32_t)), *effR = (int32_t *)malloc(nn * sizeof(int32_t));
    uint8_t *jmark = (uint8_t *)calloc(nn, 1);
    memcpy(effL, left, (size_t)n * sizeof(int32_t));
    memcpy(effR, right, (size_t)n * sizeof(int32_t));
    int64_t absorbed = 0, budget = 0, self_cycles = 0;
    budget_scan(1, n, succ, pred, right, left, keys, k - 1, effR, jmark, &absorbed);
    budget_scan(2, n, pred, succ, left, right, keys, k - 1, effL, jmark, &absorbed);
    /* keep only joining junctions */
    for (int64_t i = 0; i < n; i++) {
        if (self_loop[i]) { if (junction_joins(right[i], left[i])) self_cycles++; else budget++; }
        if (succ[i] < 0) continue;
        if (!junction_joins(right[i], left[succ[i]])) budget++; /* mixed signs as the filters left them */
        if (!(jmark[i] || junction_joins(effR[i], effL[succ[i]]))) { pred[succ[i]] = -1; succ[i] = -1; }
    }
    cbuild_t cb; cb_init(&cb, out);
    out->n_budget_junctions = budget;
    out->n_budget_admissible = absorbed; /* k-mers absorbed by budget walks (clause 3 / 4 merges) */
    out->n_cycles = self_cycles;
    char *seen = (char *)calloc(nn, 1);
    /* chains, in ascending order of head key */
    for (int64_t h = 0; h < n; h++) {
        if (pred[h] >= 0) continue;
        int64_t cnt = 0, t = h;
        for (int64_t x = h; x >= 0; x = succ[x]) { seen[x] = 1; cnt++; t = x; }
        int64_t len = cnt + k - 1;
        if (!contig_kept(len, effL[h], effR[t], min_contig)) continue;
        char *dst = cb_begin(&cb, (size_t)len, effL[h], effR[t]);
        for (int b = 0; b < k; b++) dst[b] = ACGT[(unsigned)(keys[h] >> (2 * (k - 1 - b))) & 3];
        int64_t pos = k;
        for (int64_t x = succ[h]; x >= 0; x = succ[x]) dst[pos++] = ACGT[(unsigned)keys[x] & 3];
    }
    /* cycles: every junction joins.  The reference ends with one record whose
     * two ends are the same (k-1)-mer; where it is cut is arrival-order
     * dependent.  CANONICAL ORDER: start at the smallest k-mer of the cycle. */
    for (int64_t s = 0; s < n; s++) {
        if (seen[s]) continue;
        /* s is the smallest index (= smallest key) of its cycle because we scan ascending */
        int64_t cnt = 0;
        for (int64_t x = s; !seen[x]; x = succ[x]) { seen[x] = 1; cnt++; }
        out->n_cycles++;
        int64_t len = cnt + k - 1;
        int64_t t = pred[s];
        if (!contig_kept(len, effL[s], effR[t], min_contig)) continue;
        char *dst = cb_begin(&cb, (size_t)len, effL[s], effR[t]);
        for (int b = 0; b < k; b++) dst[b] = ACGT[(unsigned)(keys[s] >> (2 * (k - 1 - b))) & 3];
        int64_t pos = k;
        for (int64_t x = succ[s]; x != s; x = succ[x]) dst[pos++] = ACGT[(unsigned)keys[x] & 3];
    }
    free(seen); free(succ); free(pred); free(self_loop); free(effL); free(effR); free(jmark);
    return 0;
}

/* ---- the same schedule, clause by clause ------------------------------------
 * Fragments are runs of the raw path; merge_frag() is the decision of
 * DSExtendReflexivKmer.call (DSMain:3070-3089) and the flag rule of
 * reflexivExtend (DSMain:3265-3279) on a (reflected R, forward F) pair.
 * Phase 1: every flagged right end absorbs single clean k-mers, paths walked
 * downstream; phase 2: the same for flagged left ends, walked upstream;
 * phase 3: every junction the clauses admit, to the fixed point. */
typedef struct { int64_t tail; int32_t left, right; int64_t n; } frag_t;

static int merge_frag(frag_t *fr, int64_t *head_of_tail, char *is_tail, int64_t a, int64_t b, int allow_same_sign) {
    frag_t *R = &fr[a], *F = &fr[b];
    int32_t bubble;
    if (F->left < 0 && R->right < 0) { if (!allow_same_sign) return 0; bubble = -1; }
    else if (F->left >= 0 && R->right >= 0) { if (!allow_same_sign) return 0; bubble = -1; }
    else if (F->left >= 0 && F->left - R->n >= 0) bubble = (int32_t)(F->left - R->n);
    else if (R->right >= 0 && R->right - F->n >= 0) bubble = (int32_t)(R->right - F->n);
    else return 0;
    int32_t nl, nr;
    if (bubble < 0) { nl = R->left; nr = F->right; }
    else if (F->left > 0) { nl = bubble; nr = F->right; }
    else { nl = R->left; nr = bubble; }
    is_tail[R->tail] = 0;
    R->left = nl; R->right = nr; R->n += F->n; R->tail = F->tail;
    head_of_tail[R->tail] = a;
    return 1;
}

static int assemble_scheduled(const u128 *keys, const int32_t *left, const int32_t *right, int64_t n,
                              int k, int min_contig, orc_contigs *out) {
    const size_t nn = (size_t)(n ? n : 1);
    int64_t *succ = (int64_t *)malloc(nn * sizeof(int64_t));
    int64_t *pred = (int64_t *)malloc(nn * sizeof(int64_t));
    char *self_loop = (char *)malloc(nn);
    if (raw_links(keys, n, k, succ, pred, self_loop)) { free(succ); free(pred); free(self_loop); return -2; }
    frag_t *fr = (frag_t *)malloc(nn * sizeof(frag_t));          /* indexed by head node */
    int64_t *head_of_tail = (int64_t *)malloc(nn * sizeof(int64_t));
    char *is_head = (char *)malloc(nn), *is_tail = (char *)malloc(nn);
    for (int64_t i = 0; i < n; i++) { is_tail[i] = 1; fr[i].tail = i; fr[i].left = left[i]; fr[i].right = right[i]; fr[i].n = 1; head_of_tail[i] = i; is_head[i] = 1; }
    /* closed raw paths are walked from the canonical start of each direction (loop_start): cut_R / cut_L mark the
     * junction IN FRONT of that start as not to be crossed by a walk of that direction */
    char *seen = (char *)calloc(nn, 1), *stop_R = (char *)calloc(nn, 1), *stop_L = (char *)calloc(nn, 1);
    int64_t *L = (int64_t *)malloc(nn * sizeof(int64_t));
    int64_t *startR = (int64_t *)malloc(nn * sizeof(int64_t)), *startL = (int64_t *)malloc(nn * sizeof(int64_t));
    int64_t n_sR = 0, n_sL = 0;
    for (int64_t h = 0; h < n; h++) if (pred[h] < 0) { startR[n_sR++] = h; for (int64_t v = h; v >= 0; v = succ[v]) seen[v] = 1; }
    for (int64_t t = 0; t < n; t++) if (succ[t] < 0) startL[n_sL++] = t;
    for (int64_t v0 = 0; v0 < n; v0++) {
        if (seen[v0]) continue;
        int64_t m = 0;
        for (int64_t v = v0; !seen[v]; v = succ[v]) { seen[v] = 1; L[m++] = v; }
        int64_t st = loop_start(L, m, right, left, keys, k - 1);
        if (st >= 0) { startR[n_sR++] = L[st]; stop_R[L[st]] = 1; }
        /* the same loop in upstream order */
        for (int64_t i = 0; i < m / 2; i++) { int64_t t = L[i]; L[i] = L[m - 1 - i]; L[m - 1 - i] = t; }
        st = loop_start(L, m, left, right, keys, k - 1);
        if (st >= 0) { startL[n_sL++] = L[st]; stop_L[L[st]] = 1; }
    }
    /* phase 1: right budgets, downstream */
    for (int64_t s = 0; s < n_sR; s++) {
        int64_t a = startR[s];
        for (;;) {
            const int64_t z = succ[fr[a].tail];
            if (z < 0 || stop_R[z]) break;
            if (fr[a].right >= 0 && fr[z].left < 0 && fr[z].n == 1 && merge_frag(fr, head_of_tail, is_tail, a, z, 0)) { is_head[z] = 0; continue; }
            a = z;
        }
    }
    /* phase 2: left budgets, upstream.  b is always a fragment head */
    for (int64_t s = 0; s < n_sL; s++) {
        int64_t x0 = startL[s];
        while (!is_tail[x0]) x0 = succ[x0];  /* the fragment that holds the start node */
        int64_t b = head_of_tail[x0];
        const int closed = stop_L[startL[s]];
        for (;;) {
            const int64_t zt = pred[b];
            if (zt < 0 || (closed && zt == x0)) break;
            const int64_t a = head_of_tail[zt];
            if (a == b) break;
            if (fr[b].left >= 0 && fr[a].right < 0 && fr[a].n == 1 && merge_frag(fr, head_of_tail, is_tail, a, b, 0)) { is_head[b] = 0; b = a; continue; }
            b = a;
        }
    }
    /* phase 3: whatever the clauses admit, to the fixed point */
    for (int changed = 1; changed;) {
        changed = 0;
        for (int64_t a = 0; a < n; a++) {
            if (!is_head[a]) continue;
            for (;;) {
                const int64_t z = succ[fr[a].tail];
                if (z < 0 || z == a || !is_head[z]) break;
                if (!merge_frag(fr, head_of_tail, is_tail, a, z, 1)) break;
                is_head[z] = 0; changed = 1;
            }
        }
    }
    cbuild_t cb; cb_init(&cb, out);
    for (int64_t h = 0; h < n; h++) {
        if (!is_head[h]) continue;
        int64_t s = h;
        if (succ[fr[h].tail] == h && junction_joins(fr[h].right, fr[h].left)) { /* closed: rotate to the smallest k-mer like the canonical form (flags are those of the cut the schedule made) */
            out->n_cycles++;
            for (int64_t x = succ[h]; x != h; x = succ[x]) if (keys[x] < keys[s]) s = x;
        }
        const int64_t len = fr[h].n + k - 1;
        if (!contig_kept(len, fr[h].left, fr[h].right, min_contig)) continue;
        char *dst = cb_begin(&cb, (size_t)len, fr[h].left, fr[h].right);
        for (int b = 0; b < k; b++) dst[b] = ACGT[(unsigned)(keys[s] >> (2 * (k - 1 - b))) & 3];
        int64_t pos = k, x = succ[s];
        for (int64_t i = 1; i < fr[h].n; i++, x = succ[x]) dst[pos++] = ACGT[(unsigned)keys[x] & 3];
    }
    for (int64_t i = 0; i < n; i++) if (self_loop[i] && junction_joins(right[i], left[i])) out->n_cycles++;
    free(succ); free(pred); free(self_loop); free(fr); free(head_of_tail); free(is_head); free(is_tail); free(seen); free(stop_R); free(stop_L);
    free(L); free(startR); free(startL);
    return 0;
}

/* ---- pass-by-pass simulation of the reference's extension loop ---------- */

typedef struct {
    u128 sub;      /* sort key: first (marker 1) or last (marker 2) k-1 bases */
    int marker;    /* 1 forward, 2 reflected */
    int32_t left, right;
    uint32_t len;  /* bases in seq = (k-1) + |extension| */
    uint8_t *seq;  /* base codes */
} rrec_t;

static u128 seq_sub(const rrec_t *r, int k) {
    u128 s = 0;
    const uint8_t *p = r->marker == 1 ? r->seq : r->seq + (r->len - (uint32_t)(k - 1));
    for (int i = 0; i < k - 1; i++) s = (s << 2) | p[i];
    return s;
}

typedef struct { rrec_t *v; size_t n, cap; } rvec_t;
static void rv_push(rvec_t *v, rrec_t r) {
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 1024; v->v = (rrec_t *)realloc(v->v, v->cap * sizeof(rrec_t)); }
    v->v[v->n++] = r;
}

/* singleKmerRandomizer, DSMain:3153-3226 / 3705-3788: the toggle decides on
 * which end the record is keyed for the next sort, then flips. */
static void randomize(rvec_t *out, rrec_t r, int *toggle, int k) {
    if (r.marker != *toggle) { r.marker = *toggle; r.sub = seq_sub(&r, k); }
    rv_push(out, r);
    *toggle = (*toggle == 1) ? 2 : 1;
}

/* reflexivExtend, DSMain:3237-3325 (and 2077-2514): merged = R.ext + sub + F.ext */
static void extend(rvec_t *out, rrec_t F, rrec_t R, int bubble, int *toggle, int k) {
    rrec_t m;
    m.len = R.len + F.len - (uint32_t)(k - 1);
    m.seq = (uint8_t *)malloc(m.len);
    memcpy(m.seq, R.seq, R.len);
    memcpy(m.seq + R.len, F.seq + (k - 1), F.len - (uint32_t)(k - 1));
    if (bubble < 0) { m.left = R.left; m.right = F.right; }
    else if (F.left > 0) { m.left = bubble; m.right = F.right; }
    else { m.left = R.left; m.right = bubble; }
    m.marker = *toggle;
    m.sub = seq_sub(&m, k);
    free(F.seq); free(R.seq);
    rv_push(out, m);
    *toggle = (*toggle == 2) ? 1 : 2;
}

static int rrec_cmp(const void *a, const void *b) {
    const rrec_t *x = (const rrec_t *)a, *y = (const rrec_t *)b;
    if (x->sub != y->sub) return x->sub < y->sub ? -1 : 1;
    /* Spark's sort is not stable; ties (one forward + one reflected record on
     * the same key) merge identically in either order.  Fix forward-first. */
    return x->marker - y->marker;
}

/* One mapPartitions(DSExtendReflexivKmer*) pass over one sorted partition,
 * DSMain:3040-3147 / 1776-1888. */
static void extension_pass(rvec_t *in, rvec_t *out, int k) {
    int toggle = 2; /* randomReflexivMarker, fresh per task */
    int have = 0;
    rrec_t held;
    memset(&held, 0, sizeof(held));
    for (size_t i = 0; i < in->n; i++) {
        rrec_t s = in->v[i];
        if (!have) { held = s; have = 1; continue; } /* lineMarker==1 or tmp list empty */
        if (s.sub == held.sub) {
            if (s.marker != held.marker) {
                rrec_t F = s.marker == 1 ? s : held, R = s.marker == 1 ? held : s;
                int f_ext = (int)F.len - (k - 1), r_ext = (int)R.len - (k - 1);
                if (F.left < 0 && R.right < 0) { extend(out, F, R, -1, &toggle, k); have = 0; }
                else if (F.left >= 0 && R.right >= 0) { extend(out, F, R, -1, &toggle, k); have = 0; }
                else if (F.left >= 0 && F.left - r_ext >= 0) { extend(out, F, R, F.left - r_ext, &toggle, k); have = 0; }
                else if (R.right >= 0 && R.right - f_ext >= 0) { extend(out, F, R, R.right - f_ext, &toggle, k); have = 0; }
                else randomize(out, s, &toggle, k);
            } else {
                randomize(out, s, &toggle, k);
            }
        } else {
            randomize(out, held, &toggle, k); /* tmpKmerRandomizer */
            held = s;
        }
    }
    if (have) randomize(out, held, &toggle, k);
    in->n = 0;
}

/* sort("k-1") + mapPartitions(DSExtendReflexivKmer*): the sorted records are cut into range partitions at key
 * boundaries and every partition is scanned by its own task (own toggle, own output), DSMain:261-326. */
static void sort_and_pass(rvec_t *in, rvec_t *out, int k) {
    par_qsort(in->v, in->n, sizeof(rrec_t), rrec_cmp);
    int T = g_threads;
    if (T <= 1 || in->n < 65536) { extension_pass(in, out, k); return; }
    size_t *cut = (size_t *)malloc((size_t)(T + 1) * sizeof(size_t));
    cut[0] = 0;
    for (int t = 1; t < T; t++) {
        size_t c = in->n * (size_t)t / (size_t)T;
        if (c < cut[t - 1]) c = cut[t - 1];
        while (c > 0 && c < in->n && in->v[c].sub == in->v[c - 1].sub) c++; /* equal keys stay in one partition */
        cut[t] = c;
    }
    cut[T] = in->n;
    rvec_t *parts = (rvec_t *)calloc((size_t)T, sizeof(rvec_t));
#pragma omp parallel for schedule(dynamic, 1) num_threads(T)
    for (int t = 0; t < T; t++) {
        rvec_t view = {in->v + cut[t], cut[t + 1] - cut[t], 0};
        extension_pass(&view, &parts[t], k);
    }
    size_t total = out->n;
    for (int t = 0; t < T; t++) total += parts[t].n;
    if (total > out->cap) { out->cap = total; out->v = (rrec_t *)realloc(out->v, out->cap * sizeof(rrec_t)); }
    for (int t = 0; t < T; t++) {
        if (parts[t].n) memcpy(out->v + out->n, parts[t].v, parts[t].n * sizeof(rrec_t));
        out->n += parts[t].n;
        free(parts[t].v);
    }
    free(parts); free(cut);
    in->n = 0;
}

static int assemble_refsim(const u128 *keys, const int32_t *left, const int32_t *right, int64_t n,
                           int k, int min_contig, int min_iter, int max_iter, orc_contigs *out) {
    rvec_t a = {0}, b = {0};
    /* Output of the left fork filter: reflected records (suffix, 2, 4|firstBase),
     * in suffix order (DSMain:3661-3685, 244). */
    okmer_t *o = (okmer_t *)malloc((size_t)(n ? n : 1) * sizeof(okmer_t));
    for (int64_t i = 0; i < n; i++) { o[i].key = keys[i]; o[i].left = left[i]; o[i].right = right[i]; }
    g_rot_k = k;
    par_qsort(o, (size_t)n, sizeof(okmer_t), ok_cmp_rot);
    /* DSkmerRandomReflection, DSMain:3688-3792 */
    {
        const int T = (g_threads > 1 && n >= 65536) ? g_threads : 1; /* one task per partition, each with its own toggle */
        a.cap = (size_t)(n ? n : 1);
        a.v = (rrec_t *)malloc(a.cap * sizeof(rrec_t));
        a.n = (size_t)n;
#pragma omp parallel for schedule(static, 1) num_threads(T)
        for (int t = 0; t < T; t++) {
            int toggle = 2;
            for (int64_t i = n * t / T; i < n * (t + 1) / T; i++) {
                rrec_t r;
                r.len = (uint32_t)k;
                r.seq = (uint8_t *)malloc((size_t)k);
                for (int j = 0; j < k; j++) r.seq[j] = (uint8_t)((o[i].key >> (2 * (k - 1 - j))) & 3);
                r.marker = 2; r.left = o[i].left; r.right = o[i].right;
                r.sub = seq_sub(&r, k);
                if (r.marker != toggle) { r.marker = toggle; r.sub = seq_sub(&r, k); } /* randomize(), in place */
                a.v[i] = r;
                toggle = (toggle == 1) ? 2 : 1;
            }
        }
    }
    free(o);
    /* DSMain:261-326 */
    int iterations = 0;
    int64_t passes = 0;
    sort_and_pass(&a, &b, k); passes++;
    for (int i = 1; i < 4; i++) {
        iterations++;
        sort_and_pass(&b, &a, k); passes++;
        rvec_t t = a; a = b; b = t;
    }
    iterations++;
    sort_and_pass(&b, &a, k); passes++; /* DSExtendReflexivKmerToArrayFirstTime */
    int64_t contig_number = 0;
    while (iterations <= max_iter) {
        iterations++;
        if (iterations >= min_iter && iterations % 3 == 0) {
            int64_t cur = (int64_t)a.n;
            if (contig_number == cur) break;
            contig_number = cur;
        }
        sort_and_pass(&a, &b, k); passes++;
        rvec_t t = a; a = b; b = t;
    }
    /* DSBinaryReflexivKmerArrayToString + DSKmerToContig, DSMain:855-900, 743-771 */
    cbuild_t cb; cb_init(&cb, out);
    out->n_passes = passes;
    for (size_t i = 0; i < a.n; i++) {
        rrec_t *r = &a.v[i];
        if (contig_kept(r->len, r->left, r->right, min_contig)) {
            char *dst = cb_begin(&cb, r->len, r->left, r->right);
            for (uint32_t j = 0; j < r->len; j++) dst[j] = ACGT[r->seq[j]];
        }
        free(r->seq);
    }
    free(a.v); free(b.v);
    return 0;
}

int orc_assemble(const uint64_t *o_hi, const uint64_t *o_lo, const int32_t *o_left, const int32_t *o_right,
                 int64_t n, int k, int min_contig, int mode, int min_iter, int max_iter, orc_contigs *out) {
    if (k < 2 || k > 63) return -1;
    u128 *keys = (u128 *)malloc((size_t)(n ? n : 1) * sizeof(u128));
    for (int64_t i = 0; i < n; i++) keys[i] = mk128(o_hi[i], o_lo[i]);
    for (int64_t i = 1; i < n; i++) if (!(keys[i - 1] < keys[i])) { free(keys); return -3; } /* sorted, unique */
    int rc = mode == ORC_ASM_REFSIM      ? assemble_refsim(keys, o_left, o_right, n, k, min_contig, min_iter, max_iter, out)
             : mode == ORC_ASM_SCHEDULED ? assemble_scheduled(keys, o_left, o_right, n, k, min_contig, out)
                                         : assemble_canonical(keys, o_left, o_right, n, k, min_contig, out);
    free(keys);
    return rc;
}
