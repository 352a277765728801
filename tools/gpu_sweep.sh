#!/bin/bash
# experiment sweep on the GPU box: hand-over pool threshold on the main configs (tile pipeline, device ms, no L2 flush)
cd "$(dirname "$0")/.."
CFGS="cfg3_reflective_refractive_1080 cfg4_shotgun_1080 cfgD_dragon_1080 cfg5_shotgun_2160"
for B in 0 4 8 12 16; do
  RTB_POOL_BELOW=$B RTB_AB_TAG=sweep_pool$B timeout 300 python tools/gpu_ab.py --tile-only $CFGS 2>&1 | grep -E "SUMMARY|Error|error"
done
