#!/bin/bash
# One GPU-box visit: parity tests, bench (ours + reference arm), per-config check, ncu launch list and one
# full capture of the traversal kernels.  Everything lands in gpurun_out/<tag>_*.
#   tools/gpu_round.sh TAG [notests] [noncu]
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt; grep -m1 'model name' /proc/cpuinfo >> $OUT/${TAG}_gpu.txt
if [[ " $* " != *" notests "* ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -5 $OUT/${TAG}_pytest.log
fi
timeout 600 python bench.py --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 3000 $OUT/${TAG}_bench.json
timeout 600 python bench.py --steps 100 --warmup 3 --scene cfgD_dragon_1080 --no-cpu-baseline > $OUT/${TAG}_bench_dragon.json 2> $OUT/${TAG}_bench_dragon.err
tail -c 1500 $OUT/${TAG}_bench_dragon.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
tail -c 1000 $OUT/${TAG}_bench_ref.json
timeout 900 python tools/gpu_check.py cfg1_simple_shapes_256 cfg2_smooth_shading_1024 cfg3_reflective_refractive_1080 cfg4_shotgun_1080 cfgD_dragon_1080 --no-ref > $OUT/${TAG}_check.log 2>&1
cp $OUT/check.json $OUT/${TAG}_check.json 2>/dev/null
if [[ " $* " != *" noncu "* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_bench_under_ncu.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walk -s 8 -c 4 -f -o $OUT/${TAG}_walk \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walk -s 2 -c 2 -f -o $OUT/${TAG}_walk_dragon \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --scene cfgD_dragon_1080 > $OUT/${TAG}_ncu_full_dragon.log 2>&1
fi
echo done
