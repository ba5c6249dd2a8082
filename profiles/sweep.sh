#!/bin/bash
# usage: sweep.sh "<args1>" "<args2>" ...   runs bench.py with each arg set and prints a one-line summary
for args in "$@"; do
  env $ENVX python bench.py --steps 5 --warmup 3 --no-cpu-baseline $args > gpurun_out/sweep.json 2> gpurun_out/sweep.err || { echo "FAILED: $args"; tail -5 gpurun_out/sweep.err; continue; }
  python -c "
import json,sys; d=json.load(open('gpurun_out/sweep.json')); print(sys.argv[1], '| ms', round(d['ms_per_step'],3), {k[3:]:round(v,2) for k,v in d['stage_ms'].items()}, {k[10:]:round(v,3) for k,v in d['roofline']['kernel_ms'].items()}, 'frac', round(d['roofline']['frac'],4), {k:d['result'][k] for k in ('n_distinct','n_rows','n_records','n_bins','n_bin_splits','n_contigs')}, 'e2e', round(d['e2e']['ms_per_step'],2))" "$args"
done
