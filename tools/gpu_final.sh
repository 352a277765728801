#!/bin/bash
# tools/gpu_final.sh TAG : the numbers README.md / DESIGN.md quote — parity tests, every config on one GPU, the reference arm,
# the ncu launch list of the bench command and full captures of the two k_tile launches of a frame (cfg4, dragon, 4K).
TAG=$1
OUT=gpurun_out; mkdir -p $OUT
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader > $OUT/${TAG}_gpu.txt; nproc >> $OUT/${TAG}_gpu.txt; grep -m1 'model name' /proc/cpuinfo >> $OUT/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; tail -2 $OUT/${TAG}_pytest.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -c 600 $OUT/${TAG}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
for SC in cfg1_simple_shapes_256 cfg2_smooth_shading_1024 cfg3_reflective_refractive_1080 cfg5_shotgun_2160 cfgD_dragon_1080; do
  timeout 600 python bench.py --steps 100 --warmup 3 --scene $SC --no-cpu-baseline > $OUT/${TAG}_bench_$SC.json 2> $OUT/${TAG}_bench_$SC.err
done
timeout 300 python tools/camera_sweep.py cfg4_shotgun_1080 120 > $OUT/${TAG}_sweep.log 2>&1
timeout 300 python tools/camera_sweep.py cfgD_dragon_1080 120 >> $OUT/${TAG}_sweep.log 2>&1
timeout 300 python tools/camera_sweep.py cfg3_reflective_refractive_1080 120 >> $OUT/${TAG}_sweep.log 2>&1
cat $OUT/${TAG}_sweep.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_bench_under_ncu.log 2>&1
for SC in cfg4_shotgun_1080 cfgD_dragon_1080 cfg5_shotgun_2160; do
  S=${SC%%_*}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile -s 6 -c 2 -f -o $OUT/${TAG}_tile_$S python tools/render_loop.py $SC 5 > $OUT/${TAG}_ncu_$S.log 2>&1
  tail -1 $OUT/${TAG}_ncu_$S.log
done
echo done
