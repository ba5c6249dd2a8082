#!/bin/bash
# compute-sanitizer over the hot path.  Run on a GPU box from the repo root:
#   bash profiles/sanitize.sh > gpurun_out/sanitize.log 2>&1 ; the summary lines go to profiles/r2_sanitizer.txt
set -u
for tool in memcheck racecheck; do
  for what in example split wide sharded; do
    echo "=== compute-sanitizer --tool $tool : $what"
    timeout 900 /usr/local/cuda/bin/compute-sanitizer --tool $tool --print-limit 5 python profiles/sanitize_driver.py $what 2>&1 | grep -E "matches the oracle|MISMATCH|ERROR SUMMARY|RACECHECK SUMMARY|rror|hazard|Invalid|=========     at|rfx::|not found|No such" | head -30
  done
done
