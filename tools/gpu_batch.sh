#!/bin/bash
TAG=${1:-b}
OUT=gpurun_out
cd "$(dirname "$0")/.."
mkdir -p $OUT
RTB_AB_CFGS="cfg3_reflective_refractive_1080 cfg4_shotgun_1080 cfgD_dragon_1080 cfg5_shotgun_2160" bash tools/gpu_variants.sh main i4l4 i16l4 i8l2 i8l8 i8l12 i12l6 2>&1 | tee $OUT/${TAG}_variants.log
