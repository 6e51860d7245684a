T0=$(date +%s); python bench.py --impl reference > gpurun_out/r2h_ref.json 2> gpurun_out/r2h_ref.err; echo "ref rc=$? $(( $(date +%s) - T0 )) s"
T0=$(date +%s); python bench.py > gpurun_out/r2h_ours.json 2> gpurun_out/r2h_ours.err; echo "ours rc=$? $(( $(date +%s) - T0 )) s"
python - <<PY
import json
r=json.loads(open("gpurun_out/r2h_ref.json").read().strip().splitlines()[-1])
o=json.loads(open("gpurun_out/r2h_ours.json").read().strip().splitlines()[-1])
print("ref", r["value"], r["unit"], r["cpu_baseline"]["cores"], "same config", r["config"]==o["config"])
print("ours", round(o["value"]), round(o["e2e"]["value"]), o["ms_per_step"], o["gpu_launches"], o["roofline"]["frac"], o["clocks"])
print("ratio e2e", o["e2e"]["value"]/r["value"])
PY
