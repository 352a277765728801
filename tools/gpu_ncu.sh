#!/bin/bash
# tools/gpu_ncu.sh TAG : ncu launch list of a short bench run + full captures of the four k_walk launches of one frame
# (pass-1 closest, pass-1 shadow, SSAA closest, SSAA shadow) for cfg4 and the dragon.
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walk -s 4 -c 4 -f -o $OUT/${TAG}_walk_cfg4 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_walk -s 4 -c 4 -f -o $OUT/${TAG}_walk_dragon \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --scene cfgD_dragon_1080 > $OUT/${TAG}_ncu_full_dragon.log 2>&1
ls -la $OUT | tail -5
