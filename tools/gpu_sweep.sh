#!/bin/bash
# tools/gpu_sweep.sh TAG "<nvcc extra flags 1>" "<flags 2>" ... : rebuild librtb_cuda.so on the GPU box per flag set and bench
TAG=$1; shift
mkdir -p gpurun_out
for FL in "$@"; do
  RTB_NVCC_EXTRA="$FL" python -c "from rendering_b200 import build; build.build_cuda(True)" > /dev/null 2>&1
  for SC in cfg4_shotgun_1080 cfgD_dragon_1080; do
    python bench.py --steps 30 --warmup 3 --scene $SC --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$FL', '$SC', 'ms', round(d['ms_per_step'], 4), 'e2e_ms', round(d['e2e']['ms_per_step'], 4), {k: round(v, 4) for k, v in d['kernel_ms_per_step'].items() if v})
" | tee -a gpurun_out/${TAG}_sweep.txt
  done
done
