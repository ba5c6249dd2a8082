"""Debug driver (not a test): N ranks on device 0, timestamps per phase."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import reflexiv_b200 as R
from reflexiv_b200 import sharded
from test_multi_gpu import _genome_reads, _split
world = int(sys.argv[1]); k = int(sys.argv[2]); err = float(sys.argv[3]); reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
txt = _genome_reads(k, err)
parts = _split(txt, world)
ctxs = [R.ReflexivContext(R.DefaultParam(kmerSize=k, minKmerCoverage=1 if k < 30 else 2, minContig=100), device=0) for _ in range(world)]
grp = sharded.LocalRanks(ctxs, arena_bytes=1 << 30)
t0 = time.time()
def body(rank, ctx):
    def log(m): print(f"[{time.time()-t0:7.3f}] rank {rank}: {m}", flush=True)
    ctx.reset(); ctx.push_fastq(parts[rank]); log("pushed")
    try:
        st = ctx.count_sharded(); log(f"counted rows={st['n_rows']} bins={st['n_bins']}")
        st2 = ctx.assemble_sharded(); log(f"assembled contigs={st2['n_contigs']} {ctx.shard_stats()}")
    except Exception as e:
        log(f"FAILED {e}"); raise
for i in range(reps):
    try:
        grp.run(body)
    except Exception as e:
        print("run failed:", e); break
for c in ctxs: c.close()
