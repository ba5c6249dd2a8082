#!/bin/bash
# Single-GPU bench lines of round 2 (run on a GPU box from the repo root; outputs under gpurun_out/, copied to profiles/).
set -u
run() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r2_line_$name.json 2> gpurun_out/r2_line_$name.err || tail -3 gpurun_out/r2_line_$name.err; }
run config2 --steps 20 --warmup 3
run config2_trimmed --steps 10 --warmup 3 --trim-to 100 --no-cpu-baseline
run config4_scaled --config 4 --scale 0.1 --steps 5 --warmup 3 --no-cpu-baseline
run config3_full --config 3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e
for f in gpurun_out/r2_line_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d["roofline"]["denominators"]
    print(sys.argv[1].split("r2_line_")[1], "ms/step %.3f" % d["ms_per_step"], "value %.3g" % d["value"], "frac kernel %.3f stage %.3f fastq %.3f" % (r["frac_kernel_events"], r["frac_stage_timers"], r["frac_from_device_fastq"]), d["stage_ms"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
