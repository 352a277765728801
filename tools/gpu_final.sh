#!/bin/bash
# tools/gpu_final.sh TAG : the numbers README.md / DESIGN.md quote — all configs on one GPU, reference arm, ncu launch list + full captures
TAG=$1
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader > $OUT/${TAG}_gpu.txt; nproc >> $OUT/${TAG}_gpu.txt; grep -m1 'model name' /proc/cpuinfo >> $OUT/${TAG}_gpu.txt
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; tail -2 $OUT/${TAG}_pytest.log
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
for SC in cfg1_simple_shapes_256 cfg2_smooth_shading_1024 cfg3_reflective_refractive_1080 cfg5_shotgun_2160 cfgD_dragon_1080; do
  python bench.py --steps 100 --warmup 3 --scene $SC > $OUT/${TAG}_bench_$SC.json 2> $OUT/${TAG}_bench_$SC.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_walk -s 4 -c 4 -f -o $OUT/${TAG}_walk_cfg4 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_cfg4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_walk -s 4 -c 4 -f -o $OUT/${TAG}_walk_dragon python bench.py --steps 1 --warmup 1 --no-cpu-baseline --scene cfgD_dragon_1080 > $OUT/${TAG}_ncu_dragon.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_walk -s 4 -c 4 -f -o $OUT/${TAG}_walk_cfg5 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --scene cfg5_shotgun_2160 > $OUT/${TAG}_ncu_cfg5.log 2>&1
echo done
