"""Profiling driver (not a test): config-2-size input, ONE rank through the sharded calls, then the plain calls."""
import os, sys
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import reflexiv_b200 as R
from reflexiv_b200 import sharded
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
wl = bench.Workload(2, 1)
txt = wl.text(0)
ctx = R.ReflexivContext(R.DefaultParam(kmerSize=31), device=0)
grp = sharded.LocalRanks([ctx], arena_bytes=24 << 30)
for i in range(reps):
    ctx.reset(); ctx.push_fastq(txt); ctx.count_sharded(); st = ctx.assemble_sharded()
print("sharded x1:", {k: round(v, 3) for k, v in st.items() if k.startswith("ms_")}, st["n_contigs"], file=sys.stderr)
ctx.close()
ctx = R.ReflexivContext(R.DefaultParam(kmerSize=31), device=0)
for i in range(reps):
    ctx.reset(); ctx.push_fastq(txt); ctx.count(); st = ctx.assemble()
print("plain:", {k: round(v, 3) for k, v in st.items() if k.startswith("ms_")}, st["n_contigs"], file=sys.stderr)
ctx.close()
