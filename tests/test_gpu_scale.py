"""BASELINE-size inputs.  (1) Size-independent properties on RFX_SCALE_TEST=<million reads> (default 4: 1.3 GB of text, a few
seconds; 16 crosses 4 GiB of text and tens of millions of lines).  (2) The oracle itself, on all host threads, at the full
size of configs[1] and on a scaled configs[3]: count table, fork-filter survivors, contigs and flags bit for bit.
RFX_SCALE_TEST=0 skips the file."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
MREADS = float(os.environ.get("RFX_SCALE_TEST", "4"))


@pytest.mark.skipif(MREADS <= 0, reason="RFX_SCALE_TEST=0")
def test_large_input_properties(orc):
    import reflexiv_b200 as R
    from workload import synth
    n_pairs = int(MREADS * 1e6 / 2)
    G = int(n_pairs * 2 * 150 / 100)          # 100x coverage
    g = synth.genome(G)
    txt = synth.fastq(g, n_pairs)
    assert MREADS < 14 or len(txt) > 2**32     # the point of the test: offsets beyond 32 bits
    results = []
    for bin_target in (0, 40_000):
        with R.ReflexivContext(R.DefaultParam(kmerSize=31, minKmerCoverage=1), bin_target_kmers=bin_target) as ctx:
            ctx.push_fastq(txt)
            st = ctx.count()
            keys, cnt = ctx.counts()
            assert st["n_reads"] == 2 * n_pairs and st["n_instances"] == 2 * n_pairs * 120
            assert int(cnt.astype(np.int64).sum()) == st["n_instances"]
            order = np.argsort(keys[:, 0])
            keys, cnt = keys[order, 0], cnt[order]
            assert np.all(keys[1:] > keys[:-1])            # no duplicate rows
            results.append((keys, cnt))
            if bin_target == 0:
                st = ctx.assemble()
                contigs = ctx.contigs_raw()
                assert st["n_contigs"] >= 2 and st["n_contig_bases"] > 1.98 * (G - 1000)
    assert np.array_equal(results[0][0], results[1][0]) and np.array_equal(results[0][1], results[1][1])  # binning does not matter
    # every k-mer of the genome is in the table (error-free reads at 100x), checked on a sample
    keys = results[0][0]
    rng = np.random.default_rng(0)
    gs = bytes(g)
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    for p in rng.integers(0, G - 31, 2000):
        v = 0
        for ch in gs[p:p + 31]:
            v = (v << 2) | code[ch]
        rc = 0
        x = v
        for _ in range(31):
            rc = (rc << 2) | ((x & 3) ^ 3)
            x >>= 2
        c = min(v, rc)
        i = np.searchsorted(keys, np.uint64(c))
        assert i < len(keys) and int(keys[i]) == c


def _sorted_table(keys, cnt, k):
    """(keys as (hi, lo) uint64 columns of the right-aligned 2k-bit value, counts) sorted by key.  The C ABI hands out the
    reference's key layout: one word for k <= 31, else (leading bases, last k % 32 bases) -- pipeline.keys_to_int."""
    k2 = keys.reshape(len(cnt), -1)
    if k <= 31:
        hi, lo = np.zeros(len(cnt), np.uint64), k2[:, 0]
    else:
        sh = np.uint64(2 * (k % 32))
        hi = k2[:, 0] >> (np.uint64(64) - sh)
        lo = (k2[:, 0] << sh) | k2[:, 1]
    order = np.lexsort((lo, hi))
    return hi[order], lo[order], cnt[order]


def _full_size_against_oracle(orc, txt, k, cover, min_contig):
    """The library against the oracle on a BASELINE-size input: the whole count table row by row, the fork-filter survivors
    with both flags, the contigs with their header flags, the statistics.  The oracle runs on all host threads."""
    import reflexiv_b200 as R
    threads = os.cpu_count() or 1
    orc.set_threads(threads)
    try:
        starts, lens = orc.fastq_reads(txt, orc.FASTQ_RUN)
        c = orc.count_kmers(txt, starts, lens, k, 0, 0, cover, 10_000_000, threads)
        f = orc.fork_filter(c["keys_hi"], c["keys_lo"], c["counts"], k, 8)
        a = orc.assemble(f["keys_hi"], f["keys_lo"], f["left"], f["right"], k, min_contig, orc.ASM_CANONICAL)
    finally:
        orc.set_threads(1)
    with R.ReflexivContext(R.DefaultParam(kmerSize=k, minKmerCoverage=cover, minContig=min_contig)) as ctx:
        ctx.push_fastq(txt)
        st = ctx.count()
        hi, lo, cnt = _sorted_table(*ctx.counts(), k)
        assert st["n_instances"] == c["n_instances"] and st["n_rows"] == len(c["counts"])
        assert np.array_equal(hi, c["keys_hi"]) and np.array_equal(lo, c["keys_lo"]) and np.array_equal(cnt, c["counts"])
        st = ctx.assemble()
        ohi, olo, ole, ori = ctx.oriented()
        order = np.lexsort((olo, ohi))
        assert np.array_equal(ohi[order], f["keys_hi"]) and np.array_equal(olo[order], f["keys_lo"])
        assert np.array_equal(ole[order], f["left"]) and np.array_equal(ori[order], f["right"])
        got = sorted((s, l, r) for s, l, r in ctx.contigs())
    exp = sorted(zip(a["contigs"], a["left"].tolist(), a["right"].tolist()))
    assert len(got) == len(exp)
    assert got == exp
    assert (st["n_budget_junctions"], st["n_budget_admissible"], st["n_cycles"]) == (a["n_budget_junctions"], a["n_budget_admissible"], a["n_cycles"])
    return st, a


@pytest.mark.skipif(MREADS <= 0, reason="RFX_SCALE_TEST=0")
def test_config2_in_full_against_the_oracle(orc):
    """BASELINE configs[1] at its full size: 4.6 Mbp genome, 3 066 668 x 150 bp reads at 100x, k = 31, cover 2 -- the input of
    the bench -- bit for bit against the oracle (about 10 s of host time on 16 threads)."""
    import bench
    wl = bench.Workload(2, 1)
    txt = wl.text(0)
    st, a = _full_size_against_oracle(orc, txt, 31, 2, 500)
    assert st["n_rows"] > 4_590_000 and st["n_contigs"] == 2 and a["n_budget_junctions"] == 0


@pytest.mark.skipif(MREADS <= 0, reason="RFX_SCALE_TEST=0")
def test_config4_scaled_against_the_oracle(orc):
    """BASELINE configs[3] (k = 61, 1 % substitution errors, cover 2, 30x) on a 5 Mbp genome instead of 100 Mbp: two-word keys,
    noisy reads (most distinct k-mers are singletons that the coverage filter drops), real forks where errors collide."""
    from workload import synth
    G = 5_000_000
    g = synth.genome(G, 4242)
    g[1_000_000:1_003_000] = g[3_000_000:3_003_000]  # a 3 kb repeat: fork winners and budget junctions at this size as well
    txt = synth.fastq(g, synth.n_pairs_for(G, 30.0, 150), read_len=150, frag_len=400, error_rate=0.01)
    st, a = _full_size_against_oracle(orc, txt, 61, 2, 500)
    assert st["n_rows"] > 4_000_000 and st["n_contigs"] >= 2
