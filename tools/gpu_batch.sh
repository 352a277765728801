#!/bin/bash
# One GPU-box visit for a batch of experiments (everything lands in gpurun_out/<tag>_*):  tools/gpu_batch.sh TAG
TAG=${1:-b}
OUT=gpurun_out
cd "$(dirname "$0")/.."
mkdir -p $OUT
bash tools/gpu_quick.sh $TAG
echo "--- without early output"
RTB_EARLY_OUTPUT=0 timeout 300 python bench.py --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_noearly.json 2> $OUT/${TAG}_bench_noearly.err
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench_noearly.json").read().strip().splitlines()[-1])
print("cfg4 no early", d["ms_per_step"], d["parity_sha_ok"], "e2e", d["e2e"]["ms_per_step"], "pipe", d.get("e2e_pipelined",{}).get("ms_per_step"))
PY
for cfg in cfgD_dragon_1080 cfg5_shotgun_2160; do
for m in 0 1; do
RTB_EARLY_OUTPUT=$m timeout 300 python bench.py --warmup 3 --steps 50 --no-cpu-baseline --scene $cfg > $OUT/${TAG}_bench_early${m}_$cfg.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench_early${m}_$cfg.json").read().strip().splitlines()[-1])
print("$cfg early=$m", d["ms_per_step"], d["parity_sha_ok"], "e2e", d["e2e"]["ms_per_step"], "pipe", d.get("e2e_pipelined",{}).get("ms_per_step"))
PY
done
done
RTB_EARLY_OUTPUT=1 timeout 300 python bench.py --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_early1.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench_early1.json").read().strip().splitlines()[-1])
print("cfg4 early=1", d["ms_per_step"], d["parity_sha_ok"], "e2e", d["e2e"]["ms_per_step"], "pipe", d.get("e2e_pipelined",{}).get("ms_per_step"))
PY
