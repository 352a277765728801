#!/bin/bash
# experiment sweep on the GPU box: tile-pipeline variants (group width / CTAs per SM / tile size) on the main configs
cd "$(dirname "$0")/.."
CFGS="cfg1_simple_shapes_256 cfg3_reflective_refractive_1080 cfg4_shotgun_1080 cfgD_dragon_1080"
for lib in default t256b3 t128b8 t128b6; do
  for R in 0 256 512 1024; do
    if [ $lib = default ]; then unset RTB_CUDA_LIB; else export RTB_CUDA_LIB=$PWD/build_variants/librtb_cuda_$lib.so; fi
    RTB_TILE_RAYS=$R RTB_AB_TAG=sweep_${lib}_R$R timeout 300 python tools/gpu_ab.py --tile-only $CFGS 2>&1 | grep SUMMARY
  done
done
