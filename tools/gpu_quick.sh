#!/bin/bash
# Short GPU-box visit: parity tests (optional), default bench line, frame times of the other configs, camera sweep.
#   tools/gpu_quick.sh TAG [notests] [nosweep]
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
if [[ " $* " != *" notests "* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $OUT/${TAG}_pytest.log
  cat $OUT/${TAG}_pytest.log
fi
timeout 300 python bench.py --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print("cfg4", d["ms_per_step"], d["parity_sha_ok"], "e2e", d["e2e"]["ms_per_step"], "pipe", d.get("e2e_pipelined",{}).get("ms_per_step"), d["kernel_ms_per_step"], d["rays"]["traced"])
PY
for cfg in cfgD_dragon_1080 cfg5_shotgun_2160 cfg3_reflective_refractive_1080 cfg1_simple_shapes_256; do
  timeout 300 python bench.py --warmup 3 --steps 50 --no-cpu-baseline --scene $cfg > $OUT/${TAG}_bench_$cfg.json 2> $OUT/${TAG}_bench_$cfg.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_bench_$cfg.json").read().strip().splitlines()[-1])
    print("$cfg", d["ms_per_step"], d["parity_sha_ok"], "e2e", d["e2e"]["ms_per_step"], d["kernel_ms_per_step"])
except Exception as e:
    print("$cfg failed", e)
PY
done
if [[ " $* " != *" nosweep "* ]]; then
  timeout 200 python tools/camera_sweep.py cfg4_shotgun_1080 120 > $OUT/${TAG}_sweep.log 2>&1; cat $OUT/${TAG}_sweep.log
fi
