#!/bin/bash
# ncu --set full captures of the HBM-side kernels (under gpurun): K1 expand + K2 gather (table_grid),
# K1e elementwise + K1b (elementwise).  Usage: scripts/gpu_ncu_modes.sh <tag>
set -u
TAG=${1:-r1}; OUT=gpurun_out; mkdir -p $OUT
for m in table_grid elementwise; do
  SHORT="python bench.py --mode $m --steps 3 --warmup 3 --no-cpu-baseline --no-clustered"
  timeout 600 $SHORT > $OUT/${TAG}_${m}_plain.log 2>&1 || { echo "plain $m failed"; continue; }
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mb_k1|mb_k2_stars" -s 12 -c 3 \
      -f -o $OUT/${TAG}_${m} $SHORT > $OUT/${TAG}_${m}_ncu.log 2>&1
  echo "ncu $m rc=$?"
done
