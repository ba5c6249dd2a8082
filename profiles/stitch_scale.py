"""-stitch at the read count of configs[1]: a 46 Mbp genome at 10x (3.07 M x 150 bp reads, k = 31, cover 3), where the k-mer coverage
drops below -cover here and there -- the case the branch exists for.  Device-resident text, CUDA-event stage timers of the library.
usage: python profiles/stitch_scale.py [genome_bp] [coverage] > profiles/r2_stitch_scale.json"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
import numpy as np
import torch

import reflexiv_b200 as R
from workload import synth

glen = int(sys.argv[1]) if len(sys.argv) > 1 else 46_000_000
cov = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
g = synth.genome(glen, seed=7)
txt = synth.fastq(g, synth.n_pairs_for(glen, cov, 150), read_len=150, frag_len=400)
n_bytes = txt.size
d_text = torch.empty(n_bytes + 64, dtype=torch.uint8, device="cuda:0")
d_text[:n_bytes] = torch.from_numpy(np.asarray(txt)).cuda()
torch.cuda.synchronize()
out = {"workload": f"{glen} bp genome, {cov:g}x, 150 bp paired reads, k=31, cover 3, -mincontig 500", "fastq_bytes": int(n_bytes), "runs": []}
with R.ReflexivContext(R.DefaultParam(kmerSize=31, minKmerCoverage=3), device=0) as ctx:
    for it in range(3):
        ctx.reset()
        ctx.push_fastq_device(d_text.data_ptr(), n_bytes)
        st = ctx.count()
        t0 = time.perf_counter()
        ctx.assemble()
        plain = ctx.stats()
        t1 = time.perf_counter()
        parse0 = plain["ms_parse"]
        ctx.stitch_begin()
        t2 = time.perf_counter()
        ctx.push_fastq_device(d_text.data_ptr(), n_bytes)
        t3 = time.perf_counter()
        ss = ctx.stitch_finish()
        t4 = time.perf_counter()
        after = ctx.stats()
        out["runs"].append({
            "n_reads": plain["n_reads"], "contigs_plain_ge500": plain["n_contigs"], "bases_plain": plain["n_contig_bases"],
            "contigs_stitched_ge500": after["n_contigs"], "bases_stitched": after["n_contig_bases"], "stitch": ss,
            "host_ms": {"assemble": (t1 - t0) * 1e3, "stitch_begin (assembly with every record + probe table)": (t2 - t1) * 1e3,
                        "read scan (K1 on the text + stitch_scan_kernel)": (t3 - t2) * 1e3, "stitch_finish": (t4 - t3) * 1e3},
            "ms_parse_scan_pass": after["ms_parse"] - parse0})
print(json.dumps(out))
