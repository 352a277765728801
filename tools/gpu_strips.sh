#!/bin/bash
# tools/gpu_strips.sh TAG N "strip sizes": strong-scaling bench at N GPUs for several strip heights
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
for S in $@; do
for SC in cfg4_shotgun_1080 cfgD_dragon_1080; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 100 --warmup 3 --scene $SC --strip-rows $S --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=$N strip=$S', '$SC', 'ms', round(d['ms_per_step'], 4), 'Mrays/s', round(d['value'],1), 'e2e_ms', round(d['e2e']['ms_per_step'], 4), {k: round(v, 4) for k, v in d['kernel_ms_per_step'].items() if v})
" | tee -a gpurun_out/${TAG}_strips.txt
done; done
