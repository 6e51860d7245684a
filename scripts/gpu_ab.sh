#!/bin/bash
# A/B of K2's prologue variants on one box (under gpurun).  Usage: scripts/gpu_ab.sh <tag>
set -u
TAG=${1:-ab}; OUT=gpurun_out; mkdir -p $OUT
run() {  # name, env..., then workload
  local name=$1; shift
  local wl=$1; shift
  env "$@" timeout 600 python bench.py --workload $wl --steps 400 --warmup 20 --no-cpu-baseline --no-clustered --no-sustained > $OUT/${TAG}_${name}_${wl}.json 2> $OUT/${TAG}_${name}_${wl}.err
  python - $OUT/${TAG}_${name}_${wl}.json $name $wl <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
p = d.get("k2_phases_us", {})
print("%-14s %-12s %8.2f us/step  k2=%.2f  e2e=%.2f us  tables=%.2f bins=%.2f loop_done=%.2f reduced=%.2f end=%.2f" % (
    sys.argv[2], sys.argv[3], 1e3 * d["ms_per_step"], 1e3 * d["kernel_ms"]["k2_stars"], d["e2e"]["ms_per_step"] * 1e3,
    p.get("tables_built", 0), p.get("nstars_bins_done", 0), p.get("star_loop_done", 0), p.get("block_reduced", 0), p.get("end", 0)))
PY
}
mkdir -p build_variants
python -c "from megabeast_b200 import build; build.build(force=True, defines=['MB_RING=0'], out='build_variants/lib_ring0.so')"
python -c "from megabeast_b200 import build; build.build(force=True, defines=['MB_RING=1'], out='build_variants/lib_ring1.so')"
VARIANTS=${2:-"both_tma:MB_X=0 both_cpasync:MB_NO_TMA=1 only_cpasync:MEGABEAST_B200_LIB=$PWD/build_variants/lib_ring0.so only_tma:MEGABEAST_B200_LIB=$PWD/build_variants/lib_ring1.so"}
for rep in 1 2; do
for wl in cfg3_eighth cfg3; do
  for v in $VARIANTS; do
    run ${v%%:*} $wl ${v#*:}
  done
done
done
