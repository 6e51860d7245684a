#!/bin/bash
# A/B of an environment switch on ONE box (under gpurun): scripts/ab_env.sh VAR  -> runs with VAR unset and VAR=1,
# three interleaved rounds, config 3 and its 1/8 shard.
set -u
VAR=$1
OUT=gpurun_out; mkdir -p $OUT
B="python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-clustered --no-sustained"
for round in 1 2 3; do
  for v in unset 1; do
    if [ $v = unset ]; then unset $VAR; else export $VAR=1; fi
    for wl in ${WLS:-cfg3_eighth cfg3}; do
      timeout 300 $B --workload $wl > $OUT/ab.json 2> $OUT/ab.err || tail -3 $OUT/ab.err
      python - $OUT/ab.json "$VAR=$v" $wl $round <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d["k2_phases_us"]
print("round %s %-16s %-12s us/step %.2f  loop_done %.2f end %.2f" % (sys.argv[4], sys.argv[2], sys.argv[3], 1e3 * d["ms_per_step"], k["star_loop_done"], k["end"]))
PY
    done
  done
done
