#!/bin/bash
# A/B of experiment builds on the GPU box: tools/gpu_variants.sh name1 name2 ...   (build_variants/librtb_cuda_<name>.so; "main" = the in-tree library)
cd "$(dirname "$0")/.."
CFGS=${RTB_AB_CFGS:-"cfg2_smooth_shading_1024 cfg3_reflective_refractive_1080 cfg4_shotgun_1080 cfgD_dragon_1080 cfg5_shotgun_2160"}
for rep in 1 2; do
for v in "$@"; do
  if [ "$v" = main ]; then unset RTB_CUDA_LIB; else export RTB_CUDA_LIB=$PWD/build_variants/librtb_cuda_$v.so; fi
  RTB_AB_TAG=var_${v}_$rep timeout 300 python tools/gpu_ab.py --tile-only $CFGS 2>&1 | grep -E "SUMMARY|Error|error"
done
done
