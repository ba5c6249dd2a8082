/*
 * reflexiv_oracle.h -- CPU restatement of the Reflexiv k-mer counting + contig
 * extension path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check
 * in __graft_entry__.py and the cpu_baseline / --impl reference legs of
 * bench.py may load it.  The product path (libreflexiv_cuda) never calls it
 * and has no CPU fallback.
 *
 * Parity status: the reference has no golden vectors in its tests
 * (src/test/.../ReflexivMainTest.java is vacuous).  The oracle is pinned on
 * the one documented known answer, docs/example.html:303-343 (example/ FASTQ
 * pair, -kmer 31 -cover 3 -> one 4558 bp contig per strand, first 1200 bases
 * printed), see tests/test_oracle_golden.py.  The reference itself (JVM +
 * Spark) cannot run in this image, so everything beyond that vector is
 * "restatement, pinned on one documented vector".
 *
 * All file:line citations are relative to
 * /root/reference/src/main/java/uni/bielefeld/cmg/reflexiv/.
 */
#ifndef REFLEXIV_ORACLE_H
#define REFLEXIV_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* FASTQ line filters ------------------------------------------------------ */
enum {
    ORC_FASTQ_RUN = 0,     /* 4-line state machine, pipeline/ReflexivDSMain.java:4048-4072 */
    ORC_FASTQ_COUNTER = 1, /* stateless heuristic, pipeline/ReflexivDataFrameCounter.java:243-289 */
    ORC_FASTQ_LINE = 2     /* -infmt line: every line is a read, ReflexivDataFrameCounter.java:184 */
};

/* Returns number of reads; *starts / *lens are malloc'ed (caller frees with orc_free). */
int64_t orc_fastq_reads(const char *txt, size_t n, int mode, uint64_t **starts, uint32_t **lens);

/* k-mer counting (A2 + A3 + A4) ------------------------------------------- */
/* Keys are right-aligned 2k-bit integers (A=0 C=1 G=2 other=3, first base most
 * significant) split into hi/lo 64-bit halves; k <= 63.
 * min_count / max_count are applied as given (callers translate the
 * reference's "only if cover > 1" rules).  Output sorted ascending by key.
 * n_threads > 1 uses OpenMP over a hash-partitioned table (CPU baseline).  */
int64_t orc_count_kmers(const char *txt, const uint64_t *starts, const uint32_t *lens, int64_t n_reads,
                        int k, int front_clip, int end_clip, int64_t min_count, int64_t max_count,
                        int n_threads, uint64_t **keys_hi, uint64_t **keys_lo, uint32_t **counts,
                        int64_t *n_instances, int64_t *n_distinct);

/* Fork filters (A6 + A7 + A8) --------------------------------------------- */
/* Input: filtered canonical table (sorted or not).  Output: oriented k-mers
 * that survive both fork filters with their (left,right) flags, sorted by key. */
int64_t orc_fork_filter(const uint64_t *keys_hi, const uint64_t *keys_lo, const uint32_t *counts,
                        int64_t n, int k, int min_error_cov, uint64_t **o_hi, uint64_t **o_lo,
                        int32_t **o_left, int32_t **o_right, int64_t *stats /* [8] */);

/* Count_<k>_sorted ------------------------------------------- */
/* pipeline/ReflexivDSKmerLeftAndRightSorting.java:105-243: both orientations of every row with count <= max_cov
 * through that class's two fork filters (flags -1 / max_kmer_size + 3, coverages saturating at 30000).  Output
 * sorted by key; the reference writes the rows as `KMER,1|left|right`.  Returns -1 outside the reference's working
 * domain (min_error_cov == 0, (k-1) % 31 == 0, k > 63). */
int64_t orc_sorted_rows(const uint64_t *keys_hi, const uint64_t *keys_lo, const uint32_t *counts, int64_t n, int k,
                        int min_error_cov, double min_repeat_fold, int max_kmer_size, int64_t max_cov,
                        uint64_t **o_hi, uint64_t **o_lo, int32_t **o_left, int32_t **o_right);

/* Extension (A9 + A10) ---------------------------------------------------- */
enum {
    ORC_ASM_CANONICAL = 0, /* fixed point under the canonical schedule: budget walks (closed form), then maximal chains */
    ORC_ASM_REFSIM = 1,    /* pass-by-pass simulation of sort + DSExtendReflexivKmer* with the reference's own toggle order */
    ORC_ASM_SCHEDULED = 2  /* the reference's four merge clauses applied literally under the canonical schedule */
};

typedef struct {
    int64_t n_contigs;
    uint64_t *offsets; /* n_contigs + 1 */
    char *bases;       /* concatenated ACGT */
    int32_t *left;
    int32_t *right;
    int64_t n_passes;          /* refsim only */
    int64_t n_budget_junctions;/* junctions whose two flags have different signs as the fork filters left them */
    int64_t n_budget_admissible;/* k-mers absorbed by budget walks (clause 3 / 4 merges, DSMain:3077-3084) */
    int64_t n_cycles;
} orc_contigs;

int orc_assemble(const uint64_t *o_hi, const uint64_t *o_lo, const int32_t *o_left, const int32_t *o_right,
                 int64_t n, int k, int min_contig, int mode, int min_iter, int max_iter, orc_contigs *out);
void orc_contigs_free(orc_contigs *c);

/* -stitch low-coverage read rescue ------------------------------ */
/* pipeline/ReflexivDSMain.java:585-672 with DSLowCoverageSubKmerExtraction (:1211-1268), DSLowCoverageReadDetection
 * (:1448-1612) and DSFilterRepeatLowCoverageFragment (:922-1010); for k > 31 (ReflexivDSMain64) the reference's probe
 * lookup can never match and the contigs come back unchanged (stitch_oracle.c).  PARITY UNPINNED, see
 * stitch_oracle.c.  Input: ALL contig records of the extension (assemble with min_contig 0) and the reads the `run`
 * FASTQ filter keeps (orc_fastq_reads, UNclipped).  Output: the contig set after stitching, filtered like A10.
 * stats[6]: probes, fragments cut, after pass 1, joined on both sides, stitched records, rings. */
int orc_stitch(int64_t n_contigs, const uint64_t *offsets, const char *bases, const int32_t *left, const int32_t *right,
               const char *txt, const uint64_t *starts, const uint32_t *lens, int64_t n_reads,
               int k, int min_contig, orc_contigs *out, int64_t *stats);

void orc_free(void *p);

/* Host threads for the sort-based stages (fork filters, ORC_ASM_REFSIM passes): T-thread sorts and one scan task per
 * range partition, as Spark local[T] runs them.  Default 1: one partition, the deterministic order the tests pin. */
void orc_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
