#!/bin/bash
# One GPU-box visit for a batch of experiments (everything lands in gpurun_out/<tag>_*):  tools/gpu_batch.sh TAG
TAG=${1:-b}
OUT=gpurun_out
cd "$(dirname "$0")/.."
mkdir -p $OUT
bash tools/gpu_quick.sh $TAG nosweep
for m in 0 1; do
RTB_EARLY_OUTPUT=$m timeout 300 python bench.py --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_early$m.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench_early$m.json").read().strip().splitlines()[-1])
print("cfg4 early=$m", d["ms_per_step"], d["parity_sha_ok"], "e2e", d["e2e"]["ms_per_step"], "float", d["e2e_float"]["ms_per_step"], "pipe", d.get("e2e_pipelined",{}).get("ms_per_step"))
PY
done
