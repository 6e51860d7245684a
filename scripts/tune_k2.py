#!/usr/bin/env python
"""
Kernel-tuning helper.  `build` (CPU box) compiles K2 variants into build_variants/*.so;
`run` (GPU box) times each variant's kernels on a workload in a fresh process.

    python scripts/tune_k2.py build
    python scripts/tune_k2.py run [cfg3] [uniform]     # prints one JSON line per variant
    python scripts/tune_k2.py one <lib> <workload> <index_mode>
"""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "build_variants")
VARIANTS = {
    "st2": ["MB_K2_STAGES=2"],
    "st3": ["MB_K2_STAGES=3"],
    "st4": ["MB_K2_STAGES=4"],
    "st6": ["MB_K2_STAGES=6"],
}


def build():
    from megabeast_b200 import build as mb_build

    os.makedirs(VDIR, exist_ok=True)
    for name, defs in VARIANTS.items():
        out = os.path.join(VDIR, f"lib_{name}.so")
        mb_build.build(force=True, defines=defs, out=out)
        print("built", out)


def one(lib, workload, index_mode):
    os.environ["MEGABEAST_B200_LIB"] = lib
    import copy

    import numpy as np
    import torch

    from megabeast_b200 import MB_Model, synth
    from megabeast_b200.parallel import ShardedLikelihood

    settings, moddata, stars, cfg = synth.make_problem(workload, mode=index_mode)
    model = MB_Model(copy.deepcopy(settings), table_precision=os.environ.get("MB_TABLE_PRECISION", "fp64"))
    cfg = dict(cfg, W=int(os.environ.get("MB_WALKERS", cfg["W"])))
    phis = synth.make_walkers(model, cfg["W"], seed=3)
    sh = ShardedLikelihood(model, stars, moddata, device=0)
    dev = torch.device("cuda", 0)
    d_phi = torch.from_numpy(np.ascontiguousarray(phis[:, sh.perm])).to(dev)
    d_rows = torch.empty((4, len(phis)), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    if os.environ.get("MB_NO_FLUSH"):
        flush = torch.empty(16, dtype=torch.uint8, device=dev)
    for _ in range(5):
        flush.zero_()
        sh.lnlike_device(d_phi, d_rows)
    torch.cuda.synchronize()
    n = 50
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        sh.lnlike_device(d_phi, d_rows)
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    sh.engine.set_profiling(True)
    kms = {}
    for _ in range(20):
        flush.zero_()
        sh.lnlike_device(d_phi, d_rows)
        torch.cuda.synchronize()
        for k, v in sh.engine.last_kernel_ms().items():
            kms[k] = kms.get(k, 0.0) + v / 20
    print(json.dumps({"lib": os.path.basename(lib), "step_ms": tot / n,
                      "evals_per_s": cfg["W"] * n / (tot * 1e-3),
                      "kernel_ms": {k: round(v, 5) for k, v in kms.items()},
                      "checksum": float((d_rows[0] + d_rows[1]).sum().item())}), flush=True)


def run(workload="cfg3", index_mode="uniform"):
    libs = sorted(os.path.join(VDIR, f) for f in os.listdir(VDIR) if f.endswith(".so"))
    for lib in libs:
        subprocess.run([sys.executable, __file__, "one", lib, workload, index_mode], check=False)


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else "run"
    if cmd == "build":
        build()
    elif cmd == "one":
        one(*sys.argv[2:5])
    else:
        run(*sys.argv[2:4])
