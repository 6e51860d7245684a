#!/usr/bin/env python
"""Wall time of the reference's driver (fit_ensemble: 5*ndim walkers, stretch move, two half-ensemble
lnprob calls per step) on the GPU path.  Usage: python scripts/fit_timing.py [cfg1 cfg3] (GPU box)."""
import copy
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from megabeast_b200 import MB_Model, fit_ensemble, synth  # noqa: E402

for name in sys.argv[1:] or ["cfg1", "cfg3"]:
    settings, moddata, stars, cfg = synth.make_problem(name, mode="clustered" if name == "cfg1" else "uniform")
    model = MB_Model(copy.deepcopy(settings))
    fit_ensemble(model, stars, moddata, nsteps=5, seed=1)  # builds the resident data, warms up
    n = 200
    t0 = time.perf_counter()
    best = fit_ensemble(model, stars, moddata, nsteps=n, seed=2)
    dt = time.perf_counter() - t0
    nw = 5 * len(best)
    print(f"{name}: fit_ensemble {n} steps x {nw} walkers on {cfg['S']} stars: {dt:.3f} s "
          f"= {1e3 * dt / n:.3f} ms per step, {n * nw / dt:.0f} lnprob evaluations per second")
