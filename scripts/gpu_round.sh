#!/bin/bash
# Runs on the GPU box (under gpurun): tests, the headline bench, then the ncu evidence.
# Usage: scripts/gpu_round.sh [tag]     -> everything lands in gpurun_out/<tag>_*
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/${TAG}_pytest.log
timeout 900 python bench.py --all-modes > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
tail -c 6000 $OUT/${TAG}_bench.json
SHORT="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-clustered"
timeout 600 $SHORT > $OUT/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/${TAG}_launches.csv $SHORT > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mb_k2_stars -s 4 -c 2 \
    -f -o $OUT/${TAG}_k2 $SHORT > $OUT/${TAG}_ncu_k2.log 2>&1
echo "ncu k2 rc=$?"
ls -la $OUT | tail -20
