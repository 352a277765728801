#!/bin/bash
# tools/gpu_sweep.sh TAG "<env assignments> | <nvcc extra flags>" ... : per configuration rebuild librtb_cuda.so on the
# GPU box with the flags, export the env assignments, check parity on two small scenes and bench cfg4 + dragon.
TAG=$1; shift
mkdir -p gpurun_out
for CFG in "$@"; do
  ENVS="${CFG%%|*}"; FL="${CFG#*|}"
  RTB_NVCC_EXTRA="$FL" python -c "from rendering_b200 import build; build.build_cuda(True)" > /dev/null 2>&1
  ( export $ENVS
    python -m pytest tests/test_gpu_parity.py -x -q -k "small_configs or trace_and_cast" 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_sweep.txt
    for SC in cfg4_shotgun_1080 cfgD_dragon_1080; do
      python bench.py --steps 30 --warmup 3 --scene $SC --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('[$CFG]', '$SC', 'ms', round(d['ms_per_step'], 4), 'e2e_ms', round(d['e2e']['ms_per_step'], 4), {k: round(v, 4) for k, v in d['kernel_ms_per_step'].items() if v})
" | tee -a gpurun_out/${TAG}_sweep.txt
    done )
done
