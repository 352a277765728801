#!/bin/bash
# tools/gpu_scale_all.sh TAG : on an 8-GPU box — strong-scaling bench lines: cfg4 at N = 2, 4, 8; the 4K frame and the dragon at N = 8
TAG=$1
OUT=gpurun_out; mkdir -p $OUT
cd "$(dirname "$0")/.."
run() { # N scene
  local N=$1 SC=$2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 60 --warmup 3 --scene $SC --no-cpu-baseline \
      > $OUT/${TAG}_n${N}_$SC.json 2> $OUT/${TAG}_n${N}_$SC.err
  python -c "
import json
d = json.loads(open('$OUT/${TAG}_n${N}_$SC.json').read().strip().splitlines()[-1])
print('N=$N', '$SC', 'ms', round(d['ms_per_step'], 4), 'Mrays/s', round(d['value'], 1), 'e2e_ms', round(d['e2e']['ms_per_step'], 4), 'parity', d['parity_sha_ok'], 'per rank', [round(x, 3) for x in d['per_rank_render_ms']])
" || tail -5 $OUT/${TAG}_n${N}_$SC.err
}
run 8 cfg4_shotgun_1080
run 4 cfg4_shotgun_1080
run 2 cfg4_shotgun_1080
run 8 cfg5_shotgun_2160
run 8 cfgD_dragon_1080
