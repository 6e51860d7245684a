#!/bin/bash
# config 3 on N GPUs (under gpurun --gpus 8): N = 2, 4, 8 back to back.  Usage: scripts/gpu_n8_cfg3.sh <tag>
set -u
TAG=${1:-r1}; OUT=gpurun_out; mkdir -p $OUT
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29550 + N)) \
      bench.py --gpus $N --steps 100 --warmup 10 > $OUT/${TAG}_cfg3_n$N.json 2> $OUT/${TAG}_cfg3_n$N.err; echo "cfg3 n$N rc=$?"
done
