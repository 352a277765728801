#!/bin/bash
# One GPU-box visit for a batch of experiments (everything lands in gpurun_out/<tag>_*):  tools/gpu_batch.sh TAG
TAG=${1:-b}
OUT=gpurun_out
cd "$(dirname "$0")/.."
mkdir -p $OUT
bash tools/gpu_quick.sh $TAG
echo "--- strips on one GPU: adaptive tile size vs 256"
timeout 300 python tools/gpu_strips.py cfg4_shotgun_1080 16 2>&1 | grep "==" | sed "s/^/adaptive /" | tee $OUT/${TAG}_strips.log
RTB_TILE_RAYS=256 timeout 300 python tools/gpu_strips.py cfg4_shotgun_1080 16 2>&1 | grep "==" | sed "s/^/R=256 /" | tee -a $OUT/${TAG}_strips.log
RTB_TILE_RAYS=128 timeout 300 python tools/gpu_strips.py cfg4_shotgun_1080 16 2>&1 | grep "==" | sed "s/^/R=128 /" | tee -a $OUT/${TAG}_strips.log
