#!/bin/bash
# 8-GPU evidence (under gpurun --gpus 8): config 3 strong + weak scaling, PHAT-scale config 4, and
# BASELINE config 5 (4,096 pixels).  Usage: scripts/gpu_n8.sh <tag>
set -u
TAG=${1:-r1}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 bench.py --gpus 8 --steps 100 --warmup 10 > $OUT/${TAG}_cfg3_n8.json 2> $OUT/${TAG}_cfg3_n8.err; echo "cfg3 n8 rc=$?"
timeout 600 $TR --master-port 29542 scripts/bench_pixels.py --steps 30 > $OUT/${TAG}_cfg5_n8.json 2> $OUT/${TAG}_cfg5_n8.err; echo "cfg5 n8 rc=$?"
timeout 900 $TR --master-port 29543 bench.py --gpus 8 --workload cfg4 --steps 30 --warmup 5 > $OUT/${TAG}_cfg4_n8.json 2> $OUT/${TAG}_cfg4_n8.err; echo "cfg4 n8 rc=$?"
tail -c 400 $OUT/${TAG}_cfg3_n8.err $OUT/${TAG}_cfg5_n8.err $OUT/${TAG}_cfg4_n8.err
