/*
 * rfx_synth.c -- deterministic synthetic genomes and paired FASTQ (host side, OpenMP).
 *
 * Implements the generator SURVEY.md section 8(d) specifies for BASELINE.json's synthetic configs:
 * iid-uniform genome; fragments of fixed length with uniform start and a coin-flip strand; read 1 =
 * first L bases of the fragment, read 2 = reverse complement of its last L bases; optional per-base
 * substitution errors; FASTQ text "@r<pair>/<mate>\n<seq>\n+\n<L x 'I'>\n" (headers start with '@',
 * quality 'I' fails DSFastqFilterOnlySeq's ATCGN test, so both reference line filters agree).
 * Every random draw is a SplitMix64 hash of (seed, index), so output is independent of thread count.
 */
#include <stdint.h>
#include <string.h>


/* Not part of libreflexiv_cuda: the workload generator of bench.py and the tests (librfx_synth.so, plain C). */
int64_t rfx_synth_genome(uint8_t* out, int64_t n_bases, uint64_t seed);
int64_t rfx_synth_fastq(const uint8_t* genome, int64_t genome_len, int64_t first_pair, int64_t n_pairs, int32_t read_len, int32_t frag_len,
                        double error_rate, uint64_t seed_reads, uint64_t seed_errors, uint8_t* out, int64_t cap);

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ULL;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
    return x ^ (x >> 31);
}

int64_t rfx_synth_genome(uint8_t* out, int64_t n_bases, uint64_t seed) {
    if (!out || n_bases < 0) return -1;
    const uint64_t s = splitmix64(seed);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_bases; i++) out[i] = (uint8_t)"ACGT"[splitmix64(s + (uint64_t)i) >> 62];
    return n_bases;
}

static int n_digits(int64_t v) {
    int d = 1;
    while (v >= 10) { v /= 10; d++; }
    return d;
}

/* total decimal digits of 0 .. n-1 */
static int64_t digits_below(int64_t n) {
    int64_t total = 0, lo = 0, hi = 10;
    int d = 1;
    while (lo < n) {
        int64_t top = hi < n ? hi : n;
        total += (top - lo) * d;
        lo = hi; hi *= 10; d++;
    }
    return total;
}

static inline uint8_t comp(uint8_t c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A'; }

int64_t rfx_synth_fastq(const uint8_t* genome, int64_t genome_len, int64_t first_pair, int64_t n_pairs, int32_t read_len,
                        int32_t frag_len, double error_rate, uint64_t seed_reads, uint64_t seed_errors, uint8_t* out, int64_t cap) {
    if (n_pairs < 0 || read_len < 1 || frag_len < read_len || genome_len < frag_len || first_pair < 0) return -1;
    const int64_t fixed = 2LL * read_len + 9;
    const int64_t d0 = digits_below(first_pair);
    const int64_t mate_bytes = n_pairs * fixed + digits_below(first_pair + n_pairs) - d0;
    const int64_t need = 2 * mate_bytes;
    if (!out) return need;
    if (!genome || cap < need) return -1;
    const uint64_t sr = splitmix64(seed_reads), se = splitmix64(seed_errors);
    const uint64_t thresh = error_rate <= 0 ? 0 : (uint64_t)(error_rate * 9007199254740992.0); /* 2^53 */
    const uint64_t span = (uint64_t)(genome_len - frag_len + 1);
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < n_pairs; q++) {
        const int64_t p = first_pair + q;
        const uint64_t start = splitmix64(sr + 2 * (uint64_t)p) % span;
        const int strand = (int)(splitmix64(sr + 2 * (uint64_t)p + 1) & 1);
        const int64_t off = q * fixed + digits_below(p) - d0;
        for (int mate = 1; mate <= 2; mate++) {
            uint8_t* w = out + (mate == 1 ? 0 : mate_bytes) + off;
            *w++ = '@'; *w++ = 'r';
            const int nd = n_digits(p);
            int64_t v = p;
            for (int i = nd - 1; i >= 0; i--) { w[i] = (uint8_t)('0' + v % 10); v /= 10; }
            w += nd;
            *w++ = '/'; *w++ = (uint8_t)('0' + mate); *w++ = '\n';
            /* forward copy of the fragment head, or reverse complement of its tail */
            const int fwd = (mate == 1) == (strand == 0);
            const uint8_t* g = genome + start;
            for (int i = 0; i < read_len; i++) {
                uint8_t c = fwd ? g[i] : comp(g[frag_len - 1 - i]);
                if (thresh) {
                    const uint64_t u = splitmix64(se + ((uint64_t)(2 * p + (mate - 1)) << 12) + (uint64_t)i);
                    if ((u >> 11) < thresh) {
                        const int code = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3;
                        c = (uint8_t)"ACGT"[(code + 1 + (int)(u % 3)) & 3];
                    }
                }
                w[i] = c;
            }
            w += read_len;
            *w++ = '\n'; *w++ = '+'; *w++ = '\n';
            memset(w, 'I', (size_t)read_len);
            w += read_len;
            *w++ = '\n';
        }
    }
    return need;
}
