#!/usr/bin/env python
"""Where does K2's time go?  Per-block phase stamps for a workload (GPU box), L2 flushed between the
launches like bench.py: distribution over the blocks of every stamp and of the star-loop duration."""
import copy
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from megabeast_b200 import MB_Model, synth  # noqa: E402
from megabeast_b200.parallel import ShardedLikelihood  # noqa: E402

NAMES = ["start", "tables", "bins", "stars", "blockred", "sent", "recv", "end"]


def q(x):
    return " ".join(f"{v:6.2f}" for v in np.percentile(x, [0, 10, 50, 90, 100]))


for name in sys.argv[1:] or ["cfg3", "cfg3_eighth"]:
    settings, moddata, stars, cfg = synth.make_problem(name)
    model = MB_Model(copy.deepcopy(settings))
    phis = synth.make_walkers(model, cfg["W"], seed=3)
    sh = ShardedLikelihood(model, stars, moddata, device=0)
    dev = torch.device("cuda", 0)
    d_phi = torch.from_numpy(np.ascontiguousarray(phis)).to(dev)
    d_rows = torch.empty((4, len(phis)), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(5):
        sh.lnlike_device(d_phi, d_rows)
    torch.cuda.synchronize()
    sh.engine.set_profiling(True)
    acc = []
    for _ in range(20):
        if not os.environ.get("NO_FLUSH"):
            flush.zero_()
        torch.cuda.synchronize()
        sh.lnlike_device(d_phi, d_rows)
        torch.cuda.synchronize()
        acc.append(sh.engine.last_k2_timeline())
    acc = np.asarray(acc)              # [launch][block][point]
    t = np.mean(acc, axis=0)
    kms = sh.engine.last_kernel_ms()
    print(f"== {name}: entries {sh.engine.info()['n_entries']}, blocks {t.shape[0]}, k2 event ms {kms['k2_stars']:.4f}")
    print("  stamp (us since the first block started): min p10 median p90 max over blocks, mean over 20 launches")
    for k in range(5):
        print(f"  {NAMES[k]:9s} {q(t[:, k])}")
    print(f"  star loop per block (stars - bins)   {q(t[:, 3] - t[:, 2])}")
    print(f"  build per block (tables - start)     {q(t[:, 1] - t[:, 0])}")
    one = acc[-1]
    print(f"  one launch: stars stamp {q(one[:, 3])};  loop duration {q(one[:, 3] - one[:, 2])}")
    late = np.argsort(t[:, 3])[-5:]
    print("  latest blocks (id: start tables bins stars):", "; ".join(f"{b}: " + " ".join(f"{t[b, k]:.2f}" for k in range(4)) for b in late))
    print(f"  last-block stamps: sent {np.max(t[:, 5]):.2f} recv {np.max(t[:, 6]):.2f} end {np.max(t[:, 7]):.2f}")
    sh.close()
