#!/bin/bash
# Quick GPU check (under gpurun): smoke, GPU tests, short bench of the modes given as arguments.
# Usage: scripts/gpu_quick.sh <tag> [mode ...]
set -u
TAG=${1:-q}; shift || true
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -7 $OUT/${TAG}_smoke.log
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/${TAG}_pytest.log
for m in "$@"; do
  timeout 600 python bench.py --mode $m --steps 50 --warmup 5 --no-cpu-baseline --no-clustered > $OUT/${TAG}_bench_$m.json 2> $OUT/${TAG}_bench_$m.err; echo "bench $m rc=$?"; tail -3 $OUT/${TAG}_bench_$m.err
  python -c "
import json;d=json.load(open('$OUT/${TAG}_bench_$m.json'));print(d['config']['engine_mode'], round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['kernel_ms'], d['roofline']['frac'])"
done
