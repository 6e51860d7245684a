#!/usr/bin/env python
"""Where does K2's time go?  Per-block phase stamps for a workload (GPU box)."""
import copy
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from megabeast_b200 import MB_Model, synth  # noqa: E402
from megabeast_b200.parallel import ShardedLikelihood  # noqa: E402

for name in sys.argv[1:] or ["cfg3", "cfg3_eighth"]:
    settings, moddata, stars, cfg = synth.make_problem(name)
    model = MB_Model(copy.deepcopy(settings))
    phis = synth.make_walkers(model, cfg["W"], seed=3)
    sh = ShardedLikelihood(model, stars, moddata, device=0)
    dev = torch.device("cuda", 0)
    d_phi = torch.from_numpy(np.ascontiguousarray(phis)).to(dev)
    d_rows = torch.empty((4, len(phis)), dtype=torch.float64, device=dev)
    for _ in range(5):
        sh.lnlike_device(d_phi, d_rows)
    torch.cuda.synchronize()
    sh.engine.set_profiling(True)
    acc = []
    for _ in range(10):
        sh.lnlike_device(d_phi, d_rows)
        torch.cuda.synchronize()
        acc.append(sh.engine.last_k2_timeline())
    t = np.mean(acc, axis=0)
    kms = sh.engine.last_kernel_ms()
    names = ["start", "tables", "poisson", "stars", "blockred", "end"]
    print(f"== {name}: entries {sh.engine.info()['n_entries']}, blocks {len(t)}, k2 event ms {kms['k2_stars']:.4f}")
    nw = min(len(phis), len(t))
    pois = t[:nw]
    rest = t[nw:]
    for label, grp in ((f"blocks 0-{nw - 1} (Poisson)", pois), (f"blocks {nw}+", rest)):
        if len(grp):
            print(f"  {label:22s} " + "  ".join(f"{n}={np.mean(grp[:, k]):6.2f}" for k, n in enumerate(names[:5])))
    print(f"  max over blocks        " + "  ".join(f"{n}={np.max(t[:, k]):6.2f}" for k, n in enumerate(names[:5]))
          + f"  end={np.max(t[:, 7]):6.2f}")
    sh.close()
