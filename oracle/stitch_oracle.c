/*
 * stitch_oracle.c -- CPU restatement of `reflexiv run -stitch`, the low-coverage read rescue of
 * pipeline/ReflexivDSMain.java:585-672.  TEST INFRASTRUCTURE ONLY (see reflexiv_oracle.h).
 *
 * PARITY UNPINNED: the reference holds no vector for this branch (it only exists in assemblyFromKmer(), needs -kmerc
 * AND -fastq, and docs/ never shows a run of it); the restatement follows the Java line by line and is checked against
 * hand-built cases (tests/test_stitch.py), nothing more.
 *
 * What the branch does ("DSMain" = pipeline/ReflexivDSMain.java):
 *   S1  DSLowCoverageSubKmerExtraction, DSMain:1211-1268: every contig record of >= 61 bases whose left flag is in
 *       [-5,-1] (a clean end, -1 - coverage) gives a probe (first (k-1)-mer, direction 1 "right extendable", contig id);
 *       a right flag in [-5,-1] gives (last (k-1)-mer, direction 0 "left extendable", contig id).
 *       SubKmerProbRowToHash, DSMain:109-118: Hashtable.put, the last put of a key stays.
 *   S2  DSLowCoverageReadDetection, DSMain:1448-1612: every FASTQ unit (DSFastqFilterWithQual) with
 *       readLength - (k-1) > 1 is scanned forward and as its reverse-complement STRING (complementary(): A<->T, C<->G,
 *       lower case folded, U = T, anything else 'N'; nucleotideValue() then maps everything that is not A/C/G to 3).
 *       left = position of the FIRST direction-0 hit, its contig is remembered; right = position of the LAST
 *       direction-1 hit whose contig differs from the remembered one.  left < right gives the fragment
 *       read[left-(k-1)+1 .. right], a record with both flags -10000000.
 *   S3  DSFilterRepeatLowCoverageFragment twice (DSMain:629-638, 922-1010): sorted by "k-1", the first record of every
 *       run of equal keys stays; the survivors are re-keyed alternately (forward / reflected) between the two passes.
 *   S4  union with the contigs, sort + DSExtendReflexivKmerToArrayLoop until the record count stands still
 *       (DSMain:640-670): clean end meets clean end (clause 1, DSMain:1811-1814), so contig A + fragment + contig B
 *       become one record.
 *   S5  DSKmerToContig, DSMain:743-771 (length >= minContig, not both flags <= -10^7).
 *
 * CANONICAL ORDER (Spark's arrival order decides these in the reference; libreflexiv_cuda matches the same choices):
 *   - contig ids: the reference numbers the rows of each partition from 1 (ids collide across partitions); here every
 *     contig has its own id (the one-partition behaviour).
 *   - probes with the same (k-1)-mer: the probe of the contig whose first k-mer is largest stays, direction 0 over
 *     direction 1 inside one contig (= the last put when the records arrive sorted by their first k-mer, forward form).
 *   - S3: pass 1 keeps, of the fragments with the same first (k-1)-mer (= leaving the same contig), the shortest, ties
 *     by the smaller 2-bit sequence.  Pass 2 only ever removes survivors that pass 1 re-keyed to their LAST (k-1)-mer,
 *     and pass 1 re-keys every second record of a partition, never the first: with every survivor the first of its
 *     partition (the usual case: few fragments over -partitionredu 200 range partitions) nothing is re-keyed and pass 2
 *     removes nothing.  That is the order fixed here.
 *   - S4: a fragment joins the contig it leaves and, of the fragments that end on the same contig, the smallest one
 *     (same comparison) joins that contig; the others stay records that end with their fragment, exactly what the
 *     extension loop leaves of a run of three (one pair merges, DSMain:1808-1830, the third record is re-emitted).
 *     Other records that happen to share a junction (k-1)-mer are left alone.  A closed ring of contigs and fragments is
 *     opened in front of the contig with the smallest first k-mer and keeps the closing fragment at its end (a record
 *     never meets itself).
 * NOT REPRODUCED: reflexivKmerExtractionFromLowCoverageFragment (DSMain:1541-1562) builds a fragment whose extension is a multiple
 *     of 31 bases long with an empty first block that lacks the length marker (firstBlock = 0: the marker line only fires at
 *     i - subKmerSize == firstBlock - 1), so the reference's own length arithmetic is off by one base on that record.  Here such
 *     a fragment is an ordinary one.
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <string.h>

#include "reflexiv_oracle.h"

static inline unsigned nv(unsigned char c) { /* nucleotideValue, DSMain:1597-1609 */
    return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : 3u;
}
static inline unsigned char complementary(unsigned char a) { /* DSMain:1527-1539 */
    if (a == 'A' || a == 'a') return 'T';
    if (a == 'T' || a == 't' || a == 'U' || a == 'u') return 'A';
    if (a == 'C' || a == 'c') return 'G';
    if (a == 'G' || a == 'g') return 'C';
    return 'N';
}

typedef struct {
    uint64_t key;      /* (k-1)-mer */
    uint64_t first;    /* first k-mer of the contig */
    int32_t dir, ctg;
} probe_t;

static int cmp_probe(const void *a, const void *b) {
    const probe_t *x = (const probe_t *)a, *y = (const probe_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    if (x->first != y->first) return x->first < y->first ? -1 : 1;
    if (x->ctg != y->ctg) return x->ctg < y->ctg ? -1 : 1; /* (never needed on an assembly: every k-mer starts one contig) */
    /* direction 0 is put after direction 1 of the same contig (DSMain:1231-1264) */
    return (x->dir == 0) - (y->dir == 0);
}

static const probe_t *probe_find(const probe_t *p, int64_t n, uint64_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (p[mid].key < key) lo = mid + 1; else hi = mid; }
    return (lo < n && p[lo].key == key) ? &p[lo] : NULL;
}

typedef struct {
    uint8_t *codes;    /* 2-bit codes, one per byte */
    int32_t len, left_ctg, right_ctg, alive;
    uint64_t first_key, last_key;
} frag_t;

static int frag_less(const frag_t *a, const frag_t *b) { /* CANONICAL ORDER: shortest, then smallest sequence */
    if (a->len != b->len) return a->len < b->len;
    return memcmp(a->codes, b->codes, (size_t)a->len) < 0;
}

typedef struct { frag_t *v; int64_t n, cap; } fragvec_t;

/* one orientation of one read: DSMain:1484-1543 (forward), :1549-1607 (reverse complement) */
static void scan_codes(const uint8_t *x, int L, int k, const probe_t *pt, int64_t np, fragvec_t *out) {
    const int sk = k - 1;
    const uint64_t mask = sk >= 32 ? ~0ull : ((1ull << (2 * sk)) - 1);
    int64_t probed_ctg = -1;
    int left = -1, right = -1;
    uint64_t w = 0;
    for (int i = 0; i < L; i++) {
        w = ((w << 2) | x[i]) & mask;
        if (i < sk - 1) continue;
        const probe_t *p = probe_find(pt, np, w);
        if (!p) continue;
        if (p->dir == 0) {
            if (left == -1) { probed_ctg = p->ctg; left = i; }
        } else {
            if (probed_ctg != p->ctg) right = i;
        }
    }
    if (left >= 0 && right >= 0 && left < right) {
        if (out->n == out->cap) { out->cap = out->cap ? out->cap * 2 : 64; out->v = (frag_t *)realloc(out->v, (size_t)out->cap * sizeof(frag_t)); }
        frag_t *f = &out->v[out->n++];
        const int s = left - sk + 1;
        f->len = right + 1 - s;
        f->codes = (uint8_t *)malloc((size_t)f->len);
        memcpy(f->codes, x + s, (size_t)f->len);
        f->alive = 1;
        uint64_t a = 0, b = 0;
        for (int j = 0; j < sk; j++) { a = (a << 2) | f->codes[j]; b = (b << 2) | f->codes[f->len - sk + j]; }
        f->first_key = a; f->last_key = b;
        /* the contigs whose probes cut the fragment */
        f->left_ctg = probe_find(pt, np, a)->ctg;
        f->right_ctg = probe_find(pt, np, b)->ctg;
    }
}

static void out_push(orc_contigs *o, int64_t *cap_n, int64_t *cap_b, int64_t len, int32_t left, int32_t right, char **dst) {
    if (o->n_contigs == *cap_n) {
        *cap_n *= 2;
        o->offsets = (uint64_t *)realloc(o->offsets, (size_t)(*cap_n + 1) * sizeof(uint64_t));
        o->left = (int32_t *)realloc(o->left, (size_t)*cap_n * sizeof(int32_t));
        o->right = (int32_t *)realloc(o->right, (size_t)*cap_n * sizeof(int32_t));
    }
    const int64_t at = (int64_t)o->offsets[o->n_contigs];
    while (at + len > *cap_b) { *cap_b = *cap_b ? *cap_b * 2 : 4096; o->bases = (char *)realloc(o->bases, (size_t)*cap_b); }
    o->left[o->n_contigs] = left; o->right[o->n_contigs] = right;
    o->n_contigs++;
    o->offsets[o->n_contigs] = (uint64_t)(at + len);
    *dst = o->bases + at;
}

/* stats: [0] probes in the table, [1] fragments cut from reads, [2] after pass 1, [3] joined on both sides,
 *        [4] stitched records (chains of >= 2 contigs), [5] rings */
int orc_stitch(int64_t n_ctg, const uint64_t *off, const char *bases, const int32_t *cl, const int32_t *cr,
               const char *txt, const uint64_t *starts, const uint32_t *lens, int64_t n_reads,
               int k, int min_contig, orc_contigs *out, int64_t *stats) {
    memset(out, 0, sizeof(*out));
    for (int i = 0; i < 6; i++) stats[i] = 0;
    if (k < 2 || k > 63) return -1;
    /* k > 31 is ReflexivDSMain64: its DSLowCoverageReadDetection (DSMain64:1562-1600) looks a Long up in a
     * Hashtable<List<Long>, Integer> (SubKmerProbRowToHash, DSMain64:119-131); a Long never equals a List, no read is
     * ever cut, the contigs come out as they went in.  Restated as: no probes. */
    const int no_probe_can_match = k > 31;
    const int sk = k - 1;

    /* ---- S1 ---- */
    probe_t *pr = (probe_t *)malloc((size_t)(2 * n_ctg + 1) * sizeof(probe_t));
    uint64_t *firstk = (uint64_t *)calloc((size_t)n_ctg + 1, sizeof(uint64_t));
    int64_t np = 0;
    for (int64_t c = 0; c < n_ctg && !no_probe_can_match; c++) {
        const char *s = bases + off[c];
        const int64_t len = (int64_t)(off[c + 1] - off[c]);
        if (len >= k) { uint64_t f = 0; for (int j = 0; j < k; j++) f = (f << 2) | nv((unsigned char)s[j]); firstk[c] = f; }
        if (len < 61) continue; /* DSMain:1229 */
        uint64_t a = 0, b = 0;
        for (int j = 0; j < sk; j++) { a = (a << 2) | nv((unsigned char)s[j]); b = (b << 2) | nv((unsigned char)s[len - sk + j]); }
        if (cl[c] >= -5 && cl[c] < 0) pr[np++] = (probe_t){a, firstk[c], 1, (int32_t)c};
        if (cr[c] >= -5 && cr[c] < 0) pr[np++] = (probe_t){b, firstk[c], 0, (int32_t)c};
    }
    qsort(pr, (size_t)np, sizeof(probe_t), cmp_probe);
    int64_t nu = 0;
    for (int64_t i = 0; i < np; i++) { /* last put stays */
        if (i + 1 < np && pr[i + 1].key == pr[i].key) continue;
        pr[nu++] = pr[i];
    }
    np = nu;
    stats[0] = np;

    /* ---- S2 ---- */
    fragvec_t fv = {NULL, 0, 0};
    uint8_t *x = NULL;
    size_t xcap = 0;
    for (int64_t r = 0; r < n_reads && np > 0; r++) {
        const int L = (int)lens[r];
        if (L - sk <= 1) continue; /* DSMain:1473 */
        if ((size_t)L > xcap) { xcap = (size_t)L * 2; x = (uint8_t *)realloc(x, xcap); }
        const unsigned char *rd = (const unsigned char *)txt + starts[r];
        for (int i = 0; i < L; i++) x[i] = (uint8_t)nv(rd[i]);
        scan_codes(x, L, k, pr, np, &fv);
        for (int i = 0; i < L; i++) x[i] = (uint8_t)nv(complementary(rd[L - 1 - i]));
        scan_codes(x, L, k, pr, np, &fv);
    }
    free(x);
    stats[1] = fv.n;

    /* ---- S3 + S4: which fragment leaves a contig, which one arrives ---- */
    int64_t *nxt = (int64_t *)malloc((size_t)(n_ctg + 1) * sizeof(int64_t));  /* fragment leaving the contig */
    int64_t *prv = (int64_t *)malloc((size_t)(n_ctg + 1) * sizeof(int64_t));  /* fragment arriving at it */
    uint8_t *seen = (uint8_t *)calloc((size_t)n_ctg + 1, 1);
    for (int64_t c = 0; c < n_ctg; c++) nxt[c] = prv[c] = -1;
    /* pass 1 of DSFilterRepeatLowCoverageFragment: fragments with the same first (k-1)-mer = fragments leaving the same
     * contig; one stays (CANONICAL ORDER: the shortest, then the smallest sequence) */
    for (int64_t i = 0; i < fv.n; i++) {
        const int32_t a = fv.v[i].left_ctg;
        if (nxt[a] < 0 || frag_less(&fv.v[i], &fv.v[nxt[a]])) nxt[a] = i;
    }
    for (int64_t i = 0; i < fv.n; i++) { fv.v[i].alive = nxt[fv.v[i].left_ctg] == i; stats[2] += fv.v[i].alive; }
    /* pass 2 removes nothing under the canonical order (see the header).  Survivors that end on the same contig: the
     * extension joins one of them to it (CANONICAL ORDER: the same comparison), the others stay records of their own that
     * end with the fragment (right flag -10000000). */
    for (int64_t i = 0; i < fv.n; i++) {
        if (!fv.v[i].alive) continue;
        const int32_t b = fv.v[i].right_ctg;
        if (prv[b] < 0 || frag_less(&fv.v[i], &fv.v[prv[b]])) prv[b] = i;
    }
    for (int64_t i = 0; i < fv.n; i++) stats[3] += fv.v[i].alive && prv[fv.v[i].right_ctg] == i;
#define ATTACHED(f) (prv[fv.v[f].right_ctg] == (f))
    int64_t cap_n = 0, cap_b = 0;
    cap_n = 16;
    out->offsets = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(cap_n + 1));
    out->left = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap_n);
    out->right = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap_n);
    out->offsets[0] = 0;
    /* rings: no member without an arriving fragment; opened at the smallest first k-mer */
    for (int64_t c = 0; c < n_ctg; c++) {
        if (prv[c] < 0 || nxt[c] < 0 || seen[c]) continue;
        int64_t cur = c, best = c, steps = 0;
        int ring = 0;
        while (nxt[cur] >= 0 && ATTACHED(nxt[cur]) && steps <= n_ctg) {
            cur = fv.v[nxt[cur]].right_ctg; steps++;
            if (cur == c) { ring = 1; break; }
            if (firstk[cur] < firstk[best] || (firstk[cur] == firstk[best] && cur < best)) best = cur;
        }
        if (!ring) continue;
        cur = c;
        do { seen[cur] = 2; cur = fv.v[nxt[cur]].right_ctg; } while (cur != c);
        seen[best] = 3; /* the ring's head */
        stats[5]++;
    }
    for (int64_t c = 0; c < n_ctg; c++) {
        const int64_t len0 = (int64_t)(off[c + 1] - off[c]);
        const int head = (seen[c] == 3) || (seen[c] == 0 && prv[c] < 0);
        if (!head) continue; /* inner member of a chain: written by its head */
        /* length and right flag of the record */
        int64_t len = len0, cur = c;
        int32_t right = cr[c];
        while (nxt[cur] >= 0) {
            const frag_t *f = &fv.v[nxt[cur]];
            len += f->len - sk;
            right = -10000000;
            if (!ATTACHED(nxt[cur])) break; /* another fragment won the contig this one ends on */
            cur = f->right_ctg;
            if (cur == c) break; /* ring closed: the closing fragment is the end of the record */
            len += (int64_t)(off[cur + 1] - off[cur]) - sk;
            right = cr[cur];
        }
        if (nxt[c] >= 0) stats[4]++;
        const int32_t left = cl[c];
        if (left <= -10000000 && right <= -10000000) continue; /* DSMain:749 */
        if (len < min_contig) continue;
        char *dst;
        out_push(out, &cap_n, &cap_b, len, left, right, &dst);
        memcpy(dst, bases + off[c], (size_t)len0); dst += len0;
        cur = c;
        while (nxt[cur] >= 0) {
            const frag_t *f = &fv.v[nxt[cur]];
            for (int j = sk; j < f->len; j++) *dst++ = "ACGT"[f->codes[j]];
            if (!ATTACHED(nxt[cur])) break;
            cur = f->right_ctg;
            if (cur == c) break;
            const int64_t l2 = (int64_t)(off[cur + 1] - off[cur]);
            memcpy(dst, bases + off[cur] + sk, (size_t)(l2 - sk)); dst += l2 - sk;
        }
    }
    for (int64_t i = 0; i < fv.n; i++) free(fv.v[i].codes);
    free(fv.v); free(nxt); free(prv); free(seen); free(pr); free(firstk);
    return 0;
}
