#!/bin/bash
# tools/gpu_scale_ab.sh TAG : on an 8-GPU box — cfg4 at N = 8 with the adaptive tile size vs 256-ray tiles, twice each
TAG=$1
OUT=gpurun_out; mkdir -p $OUT
cd "$(dirname "$0")/.."
run() { # label env
  local L=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 60 --warmup 3 --scene cfg4_shotgun_1080 --no-cpu-baseline \
      > $OUT/${TAG}_$L.json 2> $OUT/${TAG}_$L.err
  python -c "
import json
d = json.loads(open('$OUT/${TAG}_$L.json').read().strip().splitlines()[-1])
print('$L', 'ms', round(d['ms_per_step'], 4), 'e2e_ms', round(d['e2e']['ms_per_step'], 4), 'per rank', [round(x, 3) for x in d['per_rank_render_ms']], {k: round(v, 4) for k, v in d['kernel_ms_per_step'].items()})
" || tail -5 $OUT/${TAG}_$L.err
}
run adaptive_1 X=1
run r256_1 RTB_TILE_RAYS=256
run adaptive_2 X=1
run r256_2 RTB_TILE_RAYS=256
