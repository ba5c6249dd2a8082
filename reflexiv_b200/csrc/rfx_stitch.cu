// rfx_stitch.cu -- `reflexiv run -stitch`: the low-coverage read rescue of ReflexivDSMain.java:585-672; for k > 31 the
// reference's own branch can never cut a read (see stage_stitch_begin) and the stage is the plain assembly.
//
//   reference (pipeline/ReflexivDSMain.java)                          here
//   DSLowCoverageSubKmerExtraction :1211-1268 + collect +            probe_first_kernel / probe_insert_kernel: an open-addressing
//     SubKmerProbRowToHash :109-118 + broadcast :596                   table of contig-end (k-1)-mers in HBM (it stays in the L2)
//   second spark.read().text + DSFastqFilterWithQual :601-606         the K1 line scan + state machine of rfx_fastq.cu, unclipped
//   DSLowCoverageReadDetection :1448-1612                             stitch_scan_kernel: one thread per (read, strand) rolls the
//                                                                       (k-1)-mer over the TEXT (the reverse strand goes through
//                                                                       complementary() as in the reference: 'N' stays 'N' -> 3)
//   reflexivKmerExtractionFromLowCoverageFragment :1541-1595          stitch_copy_kernel: fragment bases as 2-bit codes
//   DSFilterRepeatLowCoverageFragment x 2 :629-638, :922-1010         frag_pick_kernel (one fragment leaves a contig, one arrives)
//   union + sort + DSExtendReflexivKmerToArrayLoop loop :640-670      chains over CONTIGS: heads, rings, lengths, one gather
//   DSKmerToContig :743-771                                           the same keep rule, then the contig arrays are replaced
//
// Spark's arrival order decides four things in the reference; the order fixed here is the one oracle/stitch_oracle.c states
// (probe collisions, which fragment of a run stays, which of several fragments joins a contig, where a ring is opened).
#include "rfx_internal.h"
#include "rfx_scan.cuh"
#include "rfx_stitch_kernels.cuh"

namespace rfx {

using namespace stitch;

namespace {

unsigned grid_for_n(uint64_t n, unsigned per_block = 256) {
    uint64_t g = (n + per_block - 1) / per_block;
    if (g < 1) g = 1;
    const uint64_t cap = (uint64_t)sm_count() * 16u;
    return (unsigned)(g > cap ? cap : g);
}

// device counters of the stitch stage (Ctx::st_ctr): [0] hits of the chunk, [1] codes of the chunk, [2] probes put, [3] keys in
// the table, [4] after pass 1, [5] joined on both sides, [6] stitched records, [7] rings
enum { SC_HITS = 0, SC_CODES = 1, SC_PUT = 2, SC_KEYS = 3, SC_PASS1 = 4, SC_BOTH = 5, SC_STITCHED = 6, SC_RINGS = 7, SC_N = 8 };

}  // namespace

int stage_stitch_begin(Ctx* c) {
    // k > 31 (ReflexivDSMain64.java:715-790): DSLowCoverageReadDetection rolls the read's (k-1)-mer into ONE Long and asks a
    // Hashtable<List<Long>, Integer> for it (:1562-1600 against SubKmerProbRowToHash :119-131) -- a Long never equals a List, so
    // no read is ever cut there and the branch leaves the contigs as they are.  Same here: the stage opens with an empty probe table.
    const bool no_probe_can_match = c->k > 31;
    if (c->arena) return ctx_fail(c, RFX_E_UNSUPPORTED, "-stitch runs on one GPU (assemble the shard tables with rfx_load_counts_device + rfx_assemble first)");
    if (!c->have_counts) return ctx_fail(c, RFX_E_STATE, "rfx_stitch_begin: no count table (call rfx_count or rfx_load_counts first)");
    // every record of the extension takes part, whatever its length (the minContig rule is applied to the stitched set)
    const int32_t min_contig = c->prm.min_contig;
    c->prm.min_contig = 0;
    const int rc = stage_graph(c);
    c->prm.min_contig = min_contig;
    RFX_TRY(rc);
    cudaStream_t st = c->stream;
    stage_begin(c);
    const uint64_t n = c->n_contigs;
    uint64_t cap = 1024;
    while (cap < 4 * n + 16) cap <<= 1;  // at most 2 probes per contig: load <= 0.5
    RFX_TRY(devbuf_reserve(c, c->st_keys, cap * sizeof(uint64_t)));
    RFX_TRY(devbuf_reserve(c, c->st_vals, cap * sizeof(uint32_t)));
    RFX_TRY(devbuf_reserve(c, c->st_firstk, (n + 1) * sizeof(uint64_t)));
    RFX_TRY(devbuf_reserve(c, c->st_ctr, SC_N * sizeof(uint64_t)));
    RFX_TRY(devbuf_reserve(c, c->st_bloom, ST_BLOOM_BITS / 8));
    RFX_CUDA(c, cudaMemsetAsync(c->st_bloom.p, 0, ST_BLOOM_BITS / 8, st));
    c->st_cap = cap;
    RFX_CUDA(c, cudaMemsetAsync(c->st_keys.p, 0xff, cap * sizeof(uint64_t), st));
    RFX_CUDA(c, cudaMemsetAsync(c->st_vals.p, 0xff, cap * sizeof(uint32_t), st));
    RFX_CUDA(c, cudaMemsetAsync(c->st_ctr.p, 0, SC_N * sizeof(uint64_t), st));
    unsigned long long* ctr = c->st_ctr.as<unsigned long long>();
    if (n && !no_probe_can_match) {
        probe_first_kernel<<<grid_for_n(n), 256, 0, st>>>(c->ctg_off.as<uint64_t>(), c->ctg_bases.as<char>(), n, c->k, c->st_firstk.as<uint64_t>());
        probe_insert_kernel<<<grid_for_n(n), 256, 0, st>>>(c->ctg_off.as<uint64_t>(), c->ctg_bases.as<char>(), c->ctg_left.as<int32_t>(), c->ctg_right.as<int32_t>(), n,
                                                          c->k, c->st_firstk.as<uint64_t>(), c->st_keys.as<uint64_t>(), c->st_vals.as<uint32_t>(), cap - 1, c->st_bloom.as<uint32_t>(), ctr + SC_PUT);
        count_keys_kernel<<<grid_for_n(cap), 256, 0, st>>>(c->st_keys.as<uint64_t>(), cap, ctr + SC_KEYS);
        c->launches += 3;
    }
    uint64_t h[SC_N];
    RFX_CUDA(c, cudaMemcpyAsync(h, c->st_ctr.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    RFX_CUDA(c, cudaStreamSynchronize(st));
    c->ms_stitch = stage_end(c);
    c->st_stat[0] = h[SC_KEYS];
    for (int i = 1; i < 6; i++) c->st_stat[i] = 0;
    c->st_nfrag = 0; c->st_ncodes = 0; c->st_reads = 0;
    c->st_active = true;
    return RFX_OK;
}

// Called by the K1 read-table pass (rfx_fastq.cu) for the reads of one chunk of text while the stitch stage is open.
int stitch_scan_reads(Ctx* c, const uint8_t* d_text, const uint64_t* rd_src, const uint32_t* rd_len, uint64_t n_reads) {
    c->st_reads += n_reads;
    if (n_reads == 0 || c->st_stat[0] == 0) return RFX_OK;  // no probe: no read can be cut
    cudaStream_t st = c->stream;
    unsigned long long* ctr = c->st_ctr.as<unsigned long long>();
    uint64_t hit_cap = c->st_hits.cap / sizeof(Hit);
    if (hit_cap < 4096) { RFX_TRY(devbuf_reserve(c, c->st_hits, 4096 * sizeof(Hit))); hit_cap = c->st_hits.cap / sizeof(Hit); }
    uint64_t h[2];
    for (int attempt = 0;; attempt++) {
        RFX_CUDA(c, cudaMemsetAsync(ctr, 0, 2 * sizeof(uint64_t), st));
        stitch_scan_kernel<<<grid_for_n(n_reads), 256, 0, st>>>(d_text, rd_src, rd_len, n_reads, c->k, c->st_keys.as<uint64_t>(), c->st_vals.as<uint32_t>(), c->st_cap - 1,
                                                                  c->st_bloom.as<uint32_t>(), c->st_hits.as<Hit>(), hit_cap, ctr);
        c->launches++;
        RFX_CUDA(c, cudaMemcpyAsync(h, ctr, sizeof(h), cudaMemcpyDeviceToHost, st));
        RFX_CUDA(c, cudaStreamSynchronize(st));
        if (h[0] <= hit_cap) break;
        if (attempt) return ctx_fail(c, RFX_E_STATE, "internal: the fragment list overflowed twice");
        RFX_TRY(devbuf_reserve(c, c->st_hits, h[0] * sizeof(Hit)));  // (at most 2 per read) and once more with room for all
        hit_cap = c->st_hits.cap / sizeof(Hit);
    }
    if (h[0] == 0) return RFX_OK;
    RFX_TRY(devbuf_reserve(c, c->st_frags, (c->st_nfrag + h[0]) * sizeof(Frag), true));
    RFX_TRY(devbuf_reserve(c, c->st_codes, c->st_ncodes + h[1] + 16, true));
    stitch_copy_kernel<<<grid_for_n(h[0] * 32), 256, 0, st>>>(d_text, c->st_hits.as<Hit>(), h[0], c->k, c->st_keys.as<uint64_t>(), c->st_vals.as<uint32_t>(), c->st_cap - 1,
                                                            c->st_codes.as<uint8_t>(), c->st_ncodes, c->st_frags.as<Frag>(), c->st_nfrag);
    c->launches++;
    RFX_CUDA(c, cudaStreamSynchronize(st));  // the text buffer may be reused by the caller's next chunk
    c->st_nfrag += h[0];
    c->st_ncodes += h[1];
    return RFX_OK;
}

int stage_stitch_finish(Ctx* c) {
    if (!c->st_active) return ctx_fail(c, RFX_E_STATE, "rfx_stitch_finish: no stitch stage open (call rfx_stitch_begin first)");
    c->st_active = false;
    cudaStream_t st = c->stream;
    const uint64_t n = c->n_contigs, nf = c->st_nfrag;
    c->st_stat[1] = nf;
    if (nf >= 0xffffffffull) return ctx_fail(c, RFX_E_CAPACITY, "more than 2^32 fragments");
    if (n == 0) return RFX_OK;
    stage_begin(c);
    unsigned long long* ctr = c->st_ctr.as<unsigned long long>();
    RFX_TRY(devbuf_reserve(c, c->st_nxt, n * sizeof(uint32_t)));
    RFX_TRY(devbuf_reserve(c, c->st_prv, n * sizeof(uint32_t)));
    RFX_TRY(devbuf_reserve(c, c->st_role, n));
    RFX_TRY(devbuf_reserve(c, c->st_outlen, n * sizeof(uint64_t)));
    RFX_TRY(devbuf_reserve(c, c->st_outright, n * sizeof(int32_t)));
    RFX_TRY(devbuf_reserve(c, c->st_slot, n * sizeof(uint64_t)));
    uint32_t *nxt = c->st_nxt.as<uint32_t>(), *prv = c->st_prv.as<uint32_t>();
    uint8_t* role = c->st_role.as<uint8_t>();
    const Frag* f = c->st_frags.as<Frag>();
    const uint8_t* codes = c->st_codes.as<uint8_t>();
    fill_u32_kernel<<<grid_for_n(n), 256, 0, st>>>(nxt, n, NONE32);
    fill_u32_kernel<<<grid_for_n(n), 256, 0, st>>>(prv, n, NONE32);
    RFX_CUDA(c, cudaMemsetAsync(role, 0, n, st));
    c->launches += 2;
    if (nf) {
        frag_pick_kernel<<<grid_for_n(nf), 256, 0, st>>>(f, codes, nf, 0, nullptr, nxt);
        frag_pick_kernel<<<grid_for_n(nf), 256, 0, st>>>(f, codes, nf, 1, nxt, prv);
        count_picked_kernel<<<grid_for_n(nf), 256, 0, st>>>(f, nf, nxt, prv, ctr + SC_PASS1);
        c->launches += 3;
    }
    chain_heads_kernel<<<grid_for_n(n), 256, 0, st>>>(n, nxt, prv, f, role);
    ring_heads_kernel<<<grid_for_n(n), 256, 0, st>>>(n, nxt, prv, f, c->st_firstk.as<uint64_t>(), role, ctr + SC_RINGS);
    chain_sizes_kernel<<<grid_for_n(n), 256, 0, st>>>(n, c->k, c->prm.min_contig, c->ctg_off.as<uint64_t>(), c->ctg_left.as<int32_t>(), c->ctg_right.as<int32_t>(), nxt, prv, f,
                                                     role, c->st_outlen.as<uint64_t>(), c->st_outright.as<int32_t>(), ctr + SC_STITCHED);
    c->launches += 3;
    ScanPlan<U64x3> plan;
    RFX_TRY(devbuf_reserve(c, c->scan_ws, ScanPlan<U64x3>::workspace_elems(n) * sizeof(U64x3)));
    plan.bind(n, c->scan_ws.as<U64x3>());
    KeepIn in{c->st_outlen.as<uint64_t>()};
    scan_prepare(plan, in, OpAddU64x3{}, U64x3{0, 0, 0}, st);
    c->launches += 2 * plan.levels;
    U64x3 tot;
    uint64_t h[SC_N];
    RFX_CUDA(c, cudaMemcpyAsync(&tot, plan.total, sizeof(tot), cudaMemcpyDeviceToHost, st));
    RFX_CUDA(c, cudaMemcpyAsync(h, ctr, sizeof(h), cudaMemcpyDeviceToHost, st));
    RFX_CUDA(c, cudaStreamSynchronize(st));
    DevBuf n_off, n_left, n_right, n_bases;
    int rc = RFX_OK;
    do {
        if ((rc = devbuf_reserve(c, n_off, (tot.a + 1) * sizeof(uint64_t))) != RFX_OK) break;
        if ((rc = devbuf_reserve(c, n_left, (tot.a + 1) * sizeof(int32_t))) != RFX_OK) break;
        if ((rc = devbuf_reserve(c, n_right, (tot.a + 1) * sizeof(int32_t))) != RFX_OK) break;
        if ((rc = devbuf_reserve(c, n_bases, tot.b + 16)) != RFX_OK) break;
        KeepOut out{c->ctg_left.as<int32_t>(), c->st_outright.as<int32_t>(), n_off.as<uint64_t>(), c->st_slot.as<uint64_t>(), n_left.as<int32_t>(), n_right.as<int32_t>()};
        scan_apply(plan, in, out, OpAddU64x3{}, U64x3{0, 0, 0}, st);
        set_end_kernel<<<1, 1, 0, st>>>(n_off.as<uint64_t>() + tot.a, plan.total);
        unsigned blocks = (unsigned)(n < (uint64_t)sm_count() * 8u ? (n ? n : 1) : (uint64_t)sm_count() * 8u);
        chain_gather_kernel<<<blocks, 256, 0, st>>>(n, c->k, c->ctg_off.as<uint64_t>(), c->ctg_bases.as<char>(), nxt, prv, f, codes, c->st_slot.as<uint64_t>(),
                                                    n_off.as<uint64_t>(), n_bases.as<char>());
        c->launches += 3;
        const cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { rc = ctx_fail(c, RFX_E_CUDA, "stitch gather failed: %s", cudaGetErrorString(e)); break; }
    } while (0);
    if (rc != RFX_OK) { devbuf_free(n_off); devbuf_free(n_left); devbuf_free(n_right); devbuf_free(n_bases); return rc; }
    // the stitched set replaces the contigs of the extension
    devbuf_free(c->ctg_off); devbuf_free(c->ctg_left); devbuf_free(c->ctg_right); devbuf_free(c->ctg_bases);
    c->ctg_off = n_off; c->ctg_left = n_left; c->ctg_right = n_right; c->ctg_bases = n_bases;
    c->n_contigs = tot.a;
    c->n_contig_bases = tot.b;
    c->ms_stitch += stage_end(c);
    c->st_stat[2] = h[SC_PASS1]; c->st_stat[3] = h[SC_BOTH]; c->st_stat[4] = h[SC_STITCHED]; c->st_stat[5] = h[SC_RINGS];
    return RFX_OK;
}

}  // namespace rfx
