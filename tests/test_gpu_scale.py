"""Large-input checks: size-independent properties only, the oracle is too slow here.  RFX_SCALE_TEST=<million reads>
sets the size: the default `pytest -m gpu` run uses 4 (1.3 GB of text, a few seconds); 16 crosses 4 GiB of text and
tens of millions of lines; RFX_SCALE_TEST=0 skips."""
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
