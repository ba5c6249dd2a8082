#!/bin/bash
# one point of the weak-scaling curve: bash profiles/scale_r2.sh N   (run under gpurun --gpus N)
N=$1
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_scale_$N.json 2> gpurun_out/r2_scale_$N.err
else
  RFX_SHARD_PROF=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_scale_$N.json 2> gpurun_out/r2_scale_$N.err
fi
grep "rfx_assemble_sharded" gpurun_out/r2_scale_$N.err | tail -1
python - <<PY
import json
d = json.load(open("gpurun_out/r2_scale_$N.json"))
print("N=$N ms/step %.3f value %.4g e2e %.4g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), d["stage_ms"], d.get("parity", {}).get("contig_set_equal"), d.get("shard"))
PY
