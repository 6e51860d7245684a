#!/usr/bin/env python
"""Per-WARP timing of K2's star loop (measurement build -DMB_WARP_STAMPS; GPU box):
    python -c "from megabeast_b200 import build; build.build(force=True, defines=['MB_WARP_STAMPS'], out='build_variants/lib_wstamps.so')"
    MEGABEAST_B200_LIB=$PWD/build_variants/lib_wstamps.so python scripts/k2_warp_stamps.py cfg3_eighth
When does each warp enter and leave the loop, how long is its segment, which warps hold the block up?"""
import copy
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from megabeast_b200 import MB_Model, _cabi, synth  # noqa: E402
from megabeast_b200.parallel import ShardedLikelihood  # noqa: E402

NP = 8 + 24 * 4


def q(x):
    return " ".join(f"{v:7.2f}" for v in np.percentile(x, [0, 10, 50, 90, 100]))


for name in sys.argv[1:] or ["cfg3_eighth"]:
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
    runs = []
    for _ in range(10):
        flush.zero_()
        torch.cuda.synchronize()
        sh.lnlike_device(d_phi, d_rows)
        torch.cuda.synchronize()
        us = np.empty((256, NP))
        n = sh.engine.lib.mb_last_k2_timeline(sh.engine.handles[0], _cabi.dptr(us), 256)
        runs.append(us[:n].copy())
    t = runs[-1]                       # one launch (stamps of different launches do not average per warp)
    blk, w = t[:, :8], t[:, 8:].reshape(len(t), 24, 4)
    enter, leave = w[:, :, 1], w[:, :, 2]
    length = w[:, :, 3] - blk[:, 0:1]
    dur = leave - enter
    print(f"== {name}: {len(t)} blocks; one launch")
    print(f"  tables stamp per block          {q(blk[:, 1])}")
    print(f"  warp enters loop - block tables {q((enter - blk[:, 1:2]).ravel())}")
    print(f"  star parts + bins (enter - build left) {q((enter - w[:, :, 0]).ravel())}")
    print(f"  warp loop duration              {q(dur.ravel())}")
    print(f"  segment entries                 {q(length.ravel())}")
    print(f"  ns per entry per warp           {q((1e3 * dur / np.maximum(length, 1)).ravel())}")
    print(f"  block: last leave - median leave {q(leave.max(1) - np.median(leave, 1))}")
    print(f"  block: last leave - first leave  {q(leave.max(1) - leave.min(1))}")
    print(f"  block: stars stamp - last leave  {q(blk[:, 3] - leave.max(1))}")
    late_entry = (enter - blk[:, 1:2]) > 1.0
    print(f"  warps entering > 1 us after the tables (took a bin): {late_entry.sum()} of {late_entry.size}; "
          f"their delay {q((enter - blk[:, 1:2])[late_entry]) if late_entry.any() else ''}")
    last = leave.argmax(1)
    rows = np.arange(len(t))
    print(f"  the block's LAST warp: took a bin in {late_entry[rows, last].mean() * 100:.0f} % of the blocks; "
          f"its segment - block mean segment {q(length[rows, last] - length.mean(1))}")
    corr = np.corrcoef(length.ravel(), dur.ravel())[0, 1]
    print(f"  corr(segment entries, duration) = {corr:.2f}")
    b = int(np.argmax(blk[:, 3]))
    print(f"  slowest block {b}: enter {np.round(enter[b] - blk[b, 1], 2).tolist()}")
    print(f"                    leave {np.round(leave[b] - blk[b, 1], 2).tolist()}")
    print(f"                    entries {length[b].astype(int).tolist()}")
    allr = np.asarray([r[:, 3] - r[:, 1] for r in runs])
    print(f"  (block loop time, stars - tables, 10 launches: {q(allr.ravel())})")
    sh.close()
