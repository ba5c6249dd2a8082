"""Small end-to-end runs for compute-sanitizer (profiles/sanitize.sh): the example data, a split-forcing noisy input, k = 61,
and two ranks sharing the device through the sharded calls.  Results are checked against the oracle so that a silent
corruption would show as well."""
import gzip
import os
import sys

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import reflexiv_b200 as R  # noqa: E402
from oracle import orc  # noqa: E402
from reflexiv_b200 import sharded  # noqa: E402
from workload import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
which = sys.argv[1] if len(sys.argv) > 1 else "all"


def check(txt, k, cover, label, world=1, **ctx_kw):
    ref = orc.run_pipeline(txt, k=k, cover=cover, min_contig=100)
    exp = sorted(ref["asm"]["contigs"])
    if world == 1:
        with R.ReflexivContext(R.DefaultParam(kmerSize=k, minKmerCoverage=cover, minContig=100), **ctx_kw) as ctx:
            ctx.push_fastq(txt)
            st = ctx.count()
            ctx.assemble()
            got = sorted(s for s, _, _ in ctx.contigs())
            rows = st["n_rows"]
    else:
        ctxs = [R.ReflexivContext(R.DefaultParam(kmerSize=k, minKmerCoverage=cover, minContig=100), **ctx_kw) for _ in range(world)]
        grp = sharded.LocalRanks(ctxs, arena_bytes=512 << 20)
        cut = txt.find(b"\n@", len(txt) // 2) + 1
        parts = [txt[:cut], txt[cut:]]
        out = [None] * world

        def body(r, c):
            c.push_fastq(parts[r])
            st = c.count_sharded()
            c.assemble_sharded()
            out[r] = (st["n_rows"], [s for s, _, _ in c.contigs()])
        grp.run(body)
        for c in ctxs:
            c.close()
        rows = sum(o[0] for o in out)
        got = sorted(s for o in out for s in o[1])
    ok = rows == len(ref["counts"]["counts"]) and got == exp
    print(f"{label}: rows {rows}, contigs {len(got)} -> {'matches the oracle' if ok else 'MISMATCH'}", flush=True)
    if not ok:
        sys.exit(2)


example = b"".join(gzip.open(os.path.join(GOLDEN, f)).read() for f in ("paired_dat1.fq.gz", "paired_dat2.fq.gz"))
g = synth.genome(30_000, 7)
g[12000:12800] = g[3000:3800]
noisy = bytes(synth.fastq(g, 3000, read_len=150, frag_len=400, error_rate=0.01, seed_reads=5, seed_errors=6))
if which in ("all", "example"):
    check(example, 31, 3, "example k=31 cover 3")
if which in ("all", "split"):
    os.environ["RFX_COUNT_VARIANT"] = "small"
    check(noisy, 31, 1, "noisy reads, bins of 200000 k-mers on the small table (every bin splits)", bin_target_kmers=200_000)
    del os.environ["RFX_COUNT_VARIANT"]
if which in ("all", "wide"):
    check(noisy, 61, 2, "noisy reads k=61 (two-word keys)")
if which in ("all", "sharded"):
    check(noisy, 31, 2, "two ranks sharing the device (requests / answers, splitter levels, packed contig words)", world=2)
