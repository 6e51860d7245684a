#!/usr/bin/env python
"""GPU micro-check: fixed per-launch cost of K0/K2 on tiny problems (where does the time go?)."""
import copy
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from megabeast_b200 import MB_Model, synth  # noqa: E402
from megabeast_b200.parallel import ShardedLikelihood  # noqa: E402


def run(name, W, mask_all=False, mode="auto"):
    settings, moddata, stars, cfg = synth.make_problem(name)
    if mask_all:
        stars["vals"][:] = -np.inf
    model = MB_Model(copy.deepcopy(settings), mode=mode)
    phis = synth.make_walkers(model, W, seed=3)
    sh = ShardedLikelihood(model, stars, moddata, device=0, mode=mode)
    dev = torch.device("cuda", 0)
    d_phi = torch.from_numpy(np.ascontiguousarray(phis[:, sh.perm])).to(dev)
    d_rows = torch.empty((4, W), dtype=torch.float64, device=dev)
    for _ in range(5):
        sh.lnlike_device(d_phi, d_rows)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    a.record()
    for _ in range(n):
        sh.lnlike_device(d_phi, d_rows)
    b.record()
    torch.cuda.synchronize()
    back2back = a.elapsed_time(b) / n
    sh.engine.set_profiling(True)
    kms = {}
    for _ in range(20):
        sh.lnlike_device(d_phi, d_rows)
        torch.cuda.synchronize()
        for k, v in sh.engine.last_kernel_ms().items():
            kms[k] = kms.get(k, 0.0) + v / 20
    info = sh.engine.info()
    print(json.dumps({"cfg": name, "W": W, "mask_all": mask_all, "mode": info["mode"], "entries": info["n_entries"],
                      "back_to_back_step_us": round(1e3 * back2back, 2),
                      "kernel_us": {k: round(1e3 * v, 2) for k, v in kms.items()}}), flush=True)
    sh.close()


if __name__ == "__main__":
    run("tiny", 8)
    run("tiny", 8, mask_all=True)
    run("tiny", 64)
    run("tiny", 64, mask_all=True)
    run("tiny", 128, mask_all=True)
    run("cfg1", 64)
    run("cfg1", 64, mode="table_unique")
    run("cfg1", 32)
