#!/bin/bash
# Round-2 GPU evidence (under gpurun [--gpus N]).  Usage: scripts/gpu_r2.sh <tag> [steps...]
#   steps: smoke tests bench eighth finedust n2 n8 cfg4n1 ncu modes   (default: smoke tests bench eighth)
# Everything lands in gpurun_out/<tag>_*.
set -u
TAG=${1:-r2}; shift || true
STEPS=${*:-smoke tests bench eighth}
OUT=gpurun_out; mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
has() { [[ " $STEPS " == *" $1 "* ]]; }
summ() { python - "$1" <<'EOF'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print("  no json:", e); sys.exit(0)
e = d.get("engine", {})
print("  N=%s %s value=%.0f ms/step=%.4f e2e=%.0f kernel_ms=%s" % (d["n_gpus"], e.get("engine_mode"), d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 4) for k, v in d["kernel_ms"].items()}))
print("  split=%s outer=%s inner=%s entries=%s runs=%s code_bytes=%s ingest_s=%s" % (e.get("split_inner_mask"), e.get("outer_rows"), e.get("inner_rows"), e.get("entries"), e.get("runs"), e.get("code_bytes"), e.get("ingest_s")))
for k in ("k2_phases_us", "parity", "sustained", "cpu_baseline", "clustered_indices", "weak_scaling", "secondary", "secondary_cfg5", "onchip"):
    if k in d and d[k] is not None:
        print("  %s: %s" % (k, json.dumps(d[k])[:700]))
EOF
}
if has smoke; then
  python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -8 $OUT/${TAG}_smoke.log
fi
if has tests; then
  timeout 1500 python -m pytest tests -q -m gpu --maxfail=25 -rs > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$? (gpus: $NG)"; tail -40 $OUT/${TAG}_pytest.log
fi
if has bench; then
  timeout 900 python bench.py --steps 100 --warmup 10 --all-modes > $OUT/${TAG}_bench_cfg3_n1.json 2> $OUT/${TAG}_bench_cfg3_n1.err; echo "bench cfg3 rc=$?"; tail -3 $OUT/${TAG}_bench_cfg3_n1.err
  summ $OUT/${TAG}_bench_cfg3_n1.json
fi
if has eighth; then
  for wl in cfg3_eighth cfg3_half; do
    timeout 600 python bench.py --workload $wl --steps 200 --warmup 10 --no-cpu-baseline --no-clustered > $OUT/${TAG}_bench_${wl}_n1.json 2> $OUT/${TAG}_bench_${wl}_n1.err; echo "bench $wl rc=$?"; tail -3 $OUT/${TAG}_bench_${wl}_n1.err
    summ $OUT/${TAG}_bench_${wl}_n1.json
  done
fi
if has launch; then
  # where do the ~7 us between the CUDA events and the kernel's own first/last stamps go?
  MB_NO_PDL=1 timeout 600 python bench.py --workload cfg3_eighth --steps 200 --warmup 10 --no-cpu-baseline --no-clustered --no-sustained > $OUT/${TAG}_bench_eighth_nopdl.json 2> $OUT/${TAG}_bench_eighth_nopdl.err; echo "eighth no-PDL rc=$?"
  summ $OUT/${TAG}_bench_eighth_nopdl.json
  mkdir -p build_variants
  python -c "from megabeast_b200 import build; build.build(force=True, defines=['MB_PHI_INLINE=64'], out='build_variants/lib_phi64.so')" && \
  MEGABEAST_B200_LIB=$PWD/build_variants/lib_phi64.so timeout 600 python bench.py --workload cfg3_eighth --steps 200 --warmup 10 --no-cpu-baseline --no-clustered --no-sustained > $OUT/${TAG}_bench_eighth_phi64.json 2> $OUT/${TAG}_bench_eighth_phi64.err; echo "eighth small-params rc=$?"
  summ $OUT/${TAG}_bench_eighth_phi64.json
fi
if has finedust; then
  timeout 900 python bench.py --workload fine_dust --steps 50 --warmup 5 --cpu-sample 8 --no-clustered > $OUT/${TAG}_bench_fine_dust_n1.json 2> $OUT/${TAG}_bench_fine_dust_n1.err; echo "bench fine_dust rc=$?"; tail -3 $OUT/${TAG}_bench_fine_dust_n1.err
  summ $OUT/${TAG}_bench_fine_dust_n1.json
  timeout 900 python bench.py --workload fine_dust --index-mode clustered --steps 50 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_fine_dust_clustered_n1.json 2> $OUT/${TAG}_bench_fine_dust_clustered_n1.err; echo "bench fine_dust clustered rc=$?"
  summ $OUT/${TAG}_bench_fine_dust_clustered_n1.json
  timeout 900 python bench.py --workload fine_dust --mode table_unique --steps 50 --warmup 5 --no-cpu-baseline --no-clustered > $OUT/${TAG}_bench_fine_dust_tu_n1.json 2> $OUT/${TAG}_bench_fine_dust_tu_n1.err; echo "bench fine_dust table_unique rc=$?"
  summ $OUT/${TAG}_bench_fine_dust_tu_n1.json
  timeout 900 python bench.py --workload fine_dust --index-mode clustered --mode table_unique --steps 50 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_fine_dust_tu_clustered_n1.json 2> $OUT/${TAG}_bench_fine_dust_tu_clustered_n1.err; echo "bench fine_dust table_unique clustered rc=$?"
  summ $OUT/${TAG}_bench_fine_dust_tu_clustered_n1.json
fi
if has cfg4n1; then
  timeout 1200 python bench.py --workload cfg4 --steps 30 --warmup 5 --cpu-sample 8 --no-clustered > $OUT/${TAG}_bench_cfg4_n1.json 2> $OUT/${TAG}_bench_cfg4_n1.err; echo "bench cfg4 n1 rc=$?"; tail -3 $OUT/${TAG}_bench_cfg4_n1.err
  summ $OUT/${TAG}_bench_cfg4_n1.json
fi
if has n2 && [ "$NG" -ge 2 ]; then
  for n in 2 4; do
    [ "$NG" -ge $n ] || continue
    timeout 900 $TR --nproc-per-node $n --master-port $((29540 + n)) bench.py --gpus $n --steps 100 --warmup 10 > $OUT/${TAG}_bench_cfg3_n$n.json 2> $OUT/${TAG}_bench_cfg3_n$n.err; echo "bench cfg3 n$n rc=$?"; tail -3 $OUT/${TAG}_bench_cfg3_n$n.err
    summ $OUT/${TAG}_bench_cfg3_n$n.json
  done
fi
if has n2both && [ "$NG" -ge 2 ]; then
  # the two secondary blocks of the 8-GPU line (configs 4 and 5), exercised on 2 GPUs
  timeout 1500 $TR --nproc-per-node 2 --master-port 29546 bench.py --gpus 2 --steps 50 --warmup 10 --secondary both --no-sustained > $OUT/${TAG}_bench_cfg3_n2_both.json 2> $OUT/${TAG}_bench_cfg3_n2_both.err; echo "bench cfg3 n2 + cfg4 + cfg5 rc=$?"; tail -3 $OUT/${TAG}_bench_cfg3_n2_both.err
  summ $OUT/${TAG}_bench_cfg3_n2_both.json
fi
if has n8 && [ "$NG" -ge 8 ]; then
  timeout 1500 $TR --nproc-per-node 8 --master-port 29548 bench.py --gpus 8 --steps 100 --warmup 10 > $OUT/${TAG}_bench_cfg3_n8.json 2> $OUT/${TAG}_bench_cfg3_n8.err; echo "bench cfg3 n8 rc=$?"; tail -3 $OUT/${TAG}_bench_cfg3_n8.err
  summ $OUT/${TAG}_bench_cfg3_n8.json
fi
if has pixels; then
  if [ "$NG" -ge 2 ]; then
    timeout 900 $TR --nproc-per-node $NG --master-port 29551 scripts/bench_pixels.py --steps 30 --check 16 > $OUT/${TAG}_bench_cfg5_n$NG.json 2> $OUT/${TAG}_bench_cfg5_n$NG.err; echo "cfg5 pixels n$NG rc=$?"; tail -2 $OUT/${TAG}_bench_cfg5_n$NG.err
  else
    timeout 900 python scripts/bench_pixels.py --steps 30 --check 16 > $OUT/${TAG}_bench_cfg5_n1.json 2> $OUT/${TAG}_bench_cfg5_n1.err; echo "cfg5 pixels n1 rc=$?"; tail -2 $OUT/${TAG}_bench_cfg5_n1.err
  fi
  tail -c 1500 $OUT/${TAG}_bench_cfg5_n*.json
fi
if has ncu; then
  SHORT="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-clustered --no-sustained"
  timeout 600 $SHORT > $OUT/${TAG}_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $OUT/${TAG}_launches.csv $SHORT > $OUT/${TAG}_ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:mb_k2_stars -s 4 -c 2 \
      -f -o $OUT/${TAG}_k2 $SHORT > $OUT/${TAG}_ncu_k2.log 2>&1
  echo "ncu k2 rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:mb_k2_stars -s 4 -c 2 \
      -f -o $OUT/${TAG}_k2_eighth $SHORT --workload cfg3_eighth > $OUT/${TAG}_ncu_k2_eighth.log 2>&1
  echo "ncu k2 eighth rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:mb_k2_stars -s 4 -c 2 \
      -f -o $OUT/${TAG}_k2_fine_dust $SHORT --workload fine_dust > $OUT/${TAG}_ncu_k2_fine_dust.log 2>&1
  echo "ncu k2 fine_dust rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mb_k1_elementwise|mb_k2_stars" -s 8 -c 4 \
      -f -o $OUT/${TAG}_elementwise $SHORT --mode elementwise > $OUT/${TAG}_ncu_elementwise.log 2>&1
  echo "ncu elementwise rc=$?"
fi
ls -la $OUT | tail -25
