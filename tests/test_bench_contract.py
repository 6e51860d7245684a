"""
bench.py's driver contract, checked without a GPU on the reference arm: exactly ONE line on
stdout, valid JSON, the keys the driver reads.  (The GPU arm prints the same keys plus the
device-side ones; profiles/r2_bench_cfg3_n1.json and profiles/r2_bench_cfg3_n8.json are such lines.)
"""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_reference_arm_prints_one_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--workload", "tiny", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["value"] > 0 and line["unit"] == "evals/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_committed_gpu_bench_line_has_the_contract_keys():
    with open(os.path.join(ROOT, "profiles", "r2_bench_cfg3_n1.json")) as f:
        line = json.loads(f.read().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches",
                "roofline", "cpu_baseline"):
        assert key in line, key
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(line["roofline"])
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(line["cpu_baseline"])
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(line["e2e"])
    assert line["gpu_launches"] == line["steps"]          # one launch per step
    assert set(line["config"]) == {"workload", "index_mode"}   # byte-identical in both arms
    assert line["parity"]["walkers_checked"] >= 8 and line["parity"]["max_rel_err_e2e_path"] < 1e-9
    assert line["sustained"]["seconds"] >= 2.0


def test_both_arms_print_the_same_config():
    """The driver compares the `config` objects of the two arms: they come from one function."""
    import argparse
    import bench

    a = argparse.Namespace(workload="cfg3", index_mode="uniform")
    cfg = {"S": 100000, "K": 50}
    assert bench.workload_config(a, cfg, 1000400, 64) == bench.workload_config(a, cfg, 1000400, 64)
    with open(os.path.join(ROOT, "profiles", "r2_bench_cfg3_n1.json")) as f:
        ours = json.loads(f.read().strip().splitlines()[-1])["config"]
    assert ours == bench.workload_config(a, cfg, 1000400, 64)


def test_committed_eight_gpu_line_carries_parity_and_config4():
    with open(os.path.join(ROOT, "profiles", "r2_bench_cfg3_n8.json")) as f:
        line = json.loads(f.read().strip().splitlines()[-1])
    assert line["n_gpus"] == 8 and line["scaling"] == "strong"
    p = line["parity"]
    assert p["walkers_checked"] >= 8 and p["max_rel_err_device_path"] < 1e-9 and p["p2p_equals_nccl_bitwise"] is True
    s = line["secondary"]
    assert s["workload"].startswith("cfg4") and s["parity"]["max_rel_err_device_path"] < 1e-9
    assert s["parity"]["walkers_checked"] >= 4
    assert line["kernel_ms"]["k2_stars"] <= line["ms_per_step"] * 1.15   # per-kernel time consistent with the step
    if "secondary_cfg5" in line:   # lines written after the pixel block was added to bench.py
        s5 = line["secondary_cfg5"]
        assert s5["workload"].startswith("cfg5") and s5["parity"]["max_rel_err"] < 1e-9
        assert s5["pixel_walkers_per_step"] == 4096 * 45 and s5["e2e"]["h2d_bytes_per_step"] > 0
