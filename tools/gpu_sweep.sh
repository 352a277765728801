#!/bin/bash
# experiment sweep on the GPU box: A/B of one switch / build variant on the main configs (tile pipeline, device ms, no L2 flush)
cd "$(dirname "$0")/.."
CFGS="cfg2_smooth_shading_1024 cfg4_shotgun_1080 cfgD_dragon_1080 cfg5_shotgun_2160"
unset RTB_CUDA_LIB
RTB_AB_TAG=sweep_default timeout 300 python tools/gpu_ab.py --tile-only $CFGS 2>&1 | grep -E "SUMMARY|Error|error"
RTB_NO_COVER=1 RTB_AB_TAG=sweep_nocover timeout 300 python tools/gpu_ab.py --tile-only $CFGS 2>&1 | grep -E "SUMMARY|Error|error"
for lib in "$@"; do
  RTB_CUDA_LIB=$PWD/build_variants/librtb_cuda_$lib.so RTB_AB_TAG=sweep_$lib timeout 300 python tools/gpu_ab.py --tile-only $CFGS 2>&1 | grep -E "SUMMARY|Error|error"
done
