"""Random small problems for the differential tests (reference vs oracles on CPU, CUDA path vs oracle on GPU)."""
import copy

import numpy as np

from megabeast_b200 import synth
from oracle import lnprob_port as port


def random_problem(rng):
    """A small random problem: grid axes, model variant, sparse star data with masking, a few walkers (one of
    them possibly outside the prior box), sometimes zeroed completeness rows or an all-zero star."""
    axes = tuple(int(rng.integers(lo, hi)) for lo, hi in ((2, 7), (1, 4), (2, 5), (2, 8), (1, 5), (1, 4)))
    settings = synth.docs_example_settings()
    variant = int(rng.integers(0, 8))
    if variant == 1:
        settings.stellar_model["logA"]["prior"]["name"] = "fixed"
    elif variant == 2:
        settings.fd_model["Rv"]["prior"]["name"] = "fixed"
    elif variant == 3:
        settings.fd_model["fA"]["prior"] = {"name": "flat", "minmax": [[0.05, 1.5], [0.05, 0.5]]}
        settings.fd_model["fA"]["varinit"] = [0.8, 0.25]
    elif variant == 4:
        settings.fd_model["Av"] = {
            "name": "two_lognormal", "varnames": ["mean1", "mean2", "sigma1", "sigma2", "N1_to_N2"],
            "varinit": [0.4, 2.0, 0.3, 0.4, 2.5],
            "prior": {"name": "flat", "minmax": [[0.005, 5.0], [0.005, 5.0], [0.05, 1.0], [0.05, 1.0], [0.1, 10.0]]}}
    elif variant == 5:
        settings.fd_model["Av"]["prior"]["name"] = "fixed"
    elif variant == 6:   # SFH interpolated between 3 to 7 nodes instead of a step function
        n = int(rng.integers(3, 8))
        settings.stellar_model["x"] = list(np.linspace(6.0, 10.0, n))
        settings.stellar_model["logA"] = {
            "name": "bins_interp", "varnames": ["values"], "varinit": [list(rng.uniform(0.5e-8, 3e-8, n))],
            "prior": {"name": "flat", "minmax": [[0.0] * n, [1e-3] * n]}}
    elif variant == 7:   # a Salpeter IMF behind the mass multipliers (fixed, as upstream requires)
        settings.stellar_model["M_ini"]["name"] = "salpeter"
    moddata = synth.make_grid(axes, seed=int(rng.integers(1 << 20)))
    S, K = int(rng.integers(1, 30)), int(rng.integers(1, 9))
    stars = synth.make_star_data(moddata, axes, S, K, seed=int(rng.integers(1 << 20)),
                                 mode=str(rng.choice(["uniform", "clustered"])),
                                 mask_frac=float(rng.choice([0.0, 0.02, 0.3])))
    if rng.random() < 0.2:
        moddata["completeness"] = moddata["completeness"] * (rng.random(moddata["Av"].size) > 0.5)
    if rng.random() < 0.2:
        stars["vals"][:, rng.integers(S)] = 0.0
    pm = port.PortModel(copy.deepcopy(settings))
    phis = np.array([np.asarray(pm.start_params()[1], float)] + list(synth.make_walkers(pm, 2, seed=int(rng.integers(1 << 20)))))
    if rng.random() < 0.3:
        phis[1][rng.integers(phis.shape[1])] *= -1.0
    return settings, moddata, stars, phis
