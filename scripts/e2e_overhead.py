#!/usr/bin/env python
"""Host-side cost of one batched lnprob call (GPU box): wall time per call of the public API, of
the single ctypes call underneath it, and the device time of the same launches."""
import copy
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from megabeast_b200 import MB_Model, lnprob, synth  # noqa: E402


def wall(fn, n=2000):
    for _ in range(50):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return 1e6 * (time.perf_counter() - t0) / n


for name, W in (("tiny", 64), ("cfg1", 64), ("cfg3", 64)):
    settings, moddata, stars, cfg = synth.make_problem(name)
    model = MB_Model(copy.deepcopy(settings))
    phis = synth.make_walkers(model, W, seed=3)
    lnprob(phis, model, stars, moddata)
    eng = model._engine_for(stars, moddata)[0]
    lo, hi = model._box()
    n = 2000 if name != "cfg3" else 300
    t_api = wall(lambda: lnprob(phis, model, stars, moddata), n)
    t_c = wall(lambda: eng.lnprob_box(phis, lo, hi), n)
    eng.set_profiling(True)
    eng.lnprob_box(phis, lo, hi)
    kms = eng.last_kernel_ms()
    eng.set_profiling(False)
    print(f"{name}: lnprob() {t_api:.1f} us/call, ctypes mb_lnprob_box {t_c:.1f} us/call, "
          f"kernels (events) {1e3 * sum(kms.values()):.1f} us  {kms}")
