#!/bin/bash
# tools/gpu_scale.sh TAG N : strong-scaling bench lines at N GPUs (cfg4, dragon, 4K)
TAG=$1; N=$2
OUT=gpurun_out; mkdir -p $OUT
for SC in cfg4_shotgun_1080 cfgD_dragon_1080 cfg5_shotgun_2160; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 100 --warmup 3 --scene $SC --no-cpu-baseline \
    > $OUT/${TAG}_n${N}_$SC.json 2> $OUT/${TAG}_n${N}_$SC.err
python -c "
import json
d = json.loads(open('$OUT/${TAG}_n${N}_$SC.json').read().strip().splitlines()[-1])
print('N=$N', '$SC', 'ms', round(d['ms_per_step'], 4), 'Mrays/s', round(d['value'], 1), 'e2e_ms', round(d['e2e']['ms_per_step'], 4), d['config']['partition'][:70])
" || tail -5 $OUT/${TAG}_n${N}_$SC.err
done
