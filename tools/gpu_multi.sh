#!/bin/bash
# tools/gpu_multi.sh TAG N : bench.py on N GPUs of one box (torchrun, NCCL), ours + reference arm, plus N=1 for comparison
TAG=$1; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpus.txt
python bench.py --gpus 1 --steps 100 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_n1.json 2> $OUT/${TAG}_n1.err
tail -c 600 $OUT/${TAG}_n1.json; echo
for SC in cfg4_shotgun_1080 cfgD_dragon_1080; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 100 --warmup 3 --scene $SC \
    > $OUT/${TAG}_n${N}_$SC.json 2> $OUT/${TAG}_n${N}_$SC.err
tail -c 1500 $OUT/${TAG}_n${N}_$SC.json; echo; tail -5 $OUT/${TAG}_n${N}_$SC.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 2 --warmup 1 \
    > $OUT/${TAG}_n${N}_ref.json 2> $OUT/${TAG}_n${N}_ref.err
tail -c 400 $OUT/${TAG}_n${N}_ref.json
