"""
Synthetic BEAST-shaped inputs for the ensemble likelihood.

Produces exactly the two dictionaries the reference passes to ``lnprob`` on every call
(``megabeast/fit_single.py:37-60``):

* ``beast_moddata``  -- 1-D float64 columns of length G: ``Av, Rv, f_A, fA, M_ini, logA, Z,
  distance, prior_weight, grid_weight, completeness`` (``fit_single.py:37-57``; ``fA``
  is added beside ``f_A`` because ``singlepop_dust_model.py:147`` looks the column up
  under the *parameter* name),
* ``star_lnpgriddata`` -- ``indxs`` (K, S) int64 and ``vals`` (K, S) float64 after the
  ``get_likelihoods`` transform ``exp(lnp)/prior_weight[idx]`` (``helpers.py:35-41``),
  padded with ``-inf``.

and the docs-example model settings (``docs/megabeast/settings.rst:34-104``).
"""
import copy

import numpy as np

__all__ = [
    "CONFIGS",
    "docs_example_settings",
    "Settings",
    "make_grid",
    "make_lnp",
    "make_star_data",
    "make_walkers",
    "make_problem",
    "make_pixel_problem",
]

# axis sizes in the order (logA, Z, M_ini, Av, Rv, fA); SURVEY.md section 8(d)
CONFIGS = {
    "tiny": dict(axes=(9, 2, 5, 6, 3, 1), S=64, K=12, W=8),
    "small": dict(axes=(21, 3, 9, 8, 3, 1), S=500, K=20, W=16),
    "cfg1": dict(axes=(41, 3, 27, 10, 3, 1), S=1_000, K=50, W=45),
    "cfg2": dict(axes=(41, 4, 61, 10, 3, 1), S=10_000, K=100, W=32),
    "cfg3": dict(axes=(41, 4, 61, 20, 5, 1), S=100_000, K=50, W=64),
    "cfg4": dict(axes=(41, 4, 61, 20, 10, 1), S=1_000_000, K=50, W=128),
    # config 3 with 1/2, 1/4, 1/8 of the stars: what one GPU holds when it is sharded 2/4/8 ways
    "cfg3_half": dict(axes=(41, 4, 61, 20, 5, 1), S=50_000, K=50, W=64),
    "cfg3_quarter": dict(axes=(41, 4, 61, 20, 5, 1), S=25_000, K=50, W=64),
    "cfg3_eighth": dict(axes=(41, 4, 61, 20, 5, 1), S=12_500, K=50, W=64),
    # production-like dust axes (A(V) 0-10 in steps of 0.1, 9 R(V), 5 f_A: 4,500 dust classes, too many
    # for shared memory -> `table_unique`, the class table gathered from L2)
    "fine_dust": dict(axes=(41, 4, 10, 100, 9, 5), S=100_000, K=50, W=64),
    # small-G versions of production dust axes (parity tests): 900 and 1,800 dust classes
    "dust900_small": dict(axes=(9, 2, 3, 100, 9, 1), S=400, K=24, W=16),
    "dust1800_small": dict(axes=(9, 2, 2, 200, 9, 1), S=400, K=24, W=16),
    # config 4's table shape (200 dust classes x 128 walkers = one 512-thread pass) on a small grid
    "cfg4_shape_small": dict(axes=(41, 4, 3, 20, 10, 1), S=2_000, K=50, W=128),
    # one GPU's share of the PHAT-scale config 4 on 8 GPUs
    "cfg4_eighth": dict(axes=(41, 4, 61, 20, 10, 1), S=125_000, K=50, W=128),
}

_STELLAR = {
    "name": "bins_histo",
    "x": [6.0, 7.0, 8.0, 9.0, 10.0],
    "logA": {
        "name": "bins_histo",
        "varnames": ["values"],
        "varinit": [[1e-8, 1e-8, 1e-8, 1e-8, 1e-8]],
        "prior": {
            "name": "flat",
            "minmax": [[0.0, 0.0, 0.0, 0.0, 0.0], [1e-3, 1e-3, 1e-3, 1e-3, 1e-3]],
        },
    },
    "M_ini": {
        "name": "kroupa",
        "varnames": ["slope"],
        "varinit": [2.35],
        "prior": {"name": "fixed", "minmax": [[2.0, 3.0]]},
    },
}

_DUST = {
    "Av": {
        "name": "lognormal",
        "varnames": ["mean", "sigma"],
        "varinit": [0.5, 0.25],
        "prior": {"name": "flat", "minmax": [[0.005, 5.0], [0.05, 1.0]]},
    },
    "Rv": {
        "name": "lognormal",
        "varnames": ["mean", "sigma"],
        "varinit": [4.0, 0.25],
        "prior": {"name": "flat", "minmax": [[2.0, 6.0], [0.05, 1.0]]},
    },
    "fA": {
        "name": "lognormal",
        "varnames": ["mean", "sigma"],
        "varinit": [1.0, 0.25],
        "prior": {"name": "fixed", "minmax": [[0.0, 1.0], [0.05, 0.5]]},
    },
}


class Settings:
    """The two attributes ``MB_Model.__init__`` reads from an ``mbsettings`` object
    (``singlepop_dust_model.py:22-23``)."""

    def __init__(self, stellar_model, fd_model, beast_base=None):
        self.stellar_model = stellar_model
        self.fd_model = fd_model
        self.beast_base = beast_base


def docs_example_settings():
    """Fresh deep copies of the docs-example model dictionaries (``settings.rst:34-104``
    minus the ``amr_model``/``dist_model`` blocks the code never reads)."""
    return Settings(copy.deepcopy(_STELLAR), copy.deepcopy(_DUST))


def _axis_values(axes):
    nA, nZ, nM, nAv, nRv, nfA = axes
    return (
        np.linspace(6.0, 10.0, nA),
        np.linspace(0.0004, 0.03, nZ) if nZ > 1 else np.array([0.008]),
        np.logspace(np.log10(0.2), np.log10(60.0), nM),
        np.linspace(0.01, 10.0, nAv),
        np.linspace(2.0, 6.0, nRv) if nRv > 1 else np.array([3.1]),
        np.linspace(0.25, 1.0, nfA) if nfA > 1 else np.array([1.0]),
    )


def make_grid(axes, seed=1):
    """Cartesian BEAST-like physics grid, C-order with logA slowest."""
    rng = np.random.default_rng(seed)
    vals = _axis_values(axes)
    mesh = np.meshgrid(*vals, indexing="ij")
    cols = [np.ascontiguousarray(m.reshape(-1)) for m in mesh]
    G = cols[0].size
    pw = rng.lognormal(0.0, 0.5, G)
    gw = rng.lognormal(0.0, 0.5, G)
    moddata = {
        "logA": cols[0],
        "Z": cols[1],
        "M_ini": cols[2],
        "Av": cols[3],
        "Rv": cols[4],
        "f_A": cols[5],
        "fA": cols[5],
        "distance": np.full(G, 10.0 ** (18.5 / 5.0 + 1.0)),
        "prior_weight": pw / pw.mean(),
        "grid_weight": gw / gw.mean(),
        "completeness": rng.random(G),
    }
    return moddata


def _first_k_unique(cands, K, rng):
    """From each row of ``cands`` keep K distinct entries in random order."""
    S, M = cands.shape
    order = np.argsort(cands, axis=1, kind="stable")
    srt = np.take_along_axis(cands, order, axis=1)
    dup_sorted = np.zeros((S, M), dtype=bool)
    dup_sorted[:, 1:] = srt[:, 1:] == srt[:, :-1]
    dup = np.empty_like(dup_sorted)
    np.put_along_axis(dup, order, dup_sorted, axis=1)
    keys = rng.random((S, M))
    keys[dup] = 2.0
    pick = np.argsort(keys, axis=1)[:, :K]
    out = np.take_along_axis(cands, pick, axis=1)
    short = (~dup).sum(axis=1) < K
    return out, short


def make_lnp(moddata, axes, S, K, mode="uniform", seed=2, mask_frac=0.02):
    """Sparse log-posteriors in the layout ``read_lnp_data`` returns: ``vals`` (K, S) of
    lnp with -inf padding (max per star = 0) and ``indxs`` (K, S) int64."""
    rng = np.random.default_rng(seed)
    G = moddata["Av"].size
    if mode == "uniform":
        cands = rng.integers(0, G, size=(S, K + 8), dtype=np.int64)
        idx, short = _first_k_unique(cands, K, rng)
        for s in np.nonzero(short)[0]:
            idx[s] = rng.choice(G, size=K, replace=False)
    elif mode == "clustered":
        dims = np.array(axes, dtype=np.int64)
        centre = np.stack([rng.integers(0, d, size=S) for d in dims], axis=1)
        M = 4 * K
        off = rng.integers(-2, 3, size=(S, M, len(dims)))
        coords = np.clip(centre[:, None, :] + off, 0, dims - 1)
        strides = np.array(
            [int(np.prod(dims[i + 1:])) for i in range(len(dims))], dtype=np.int64
        )
        cands = (coords * strides).sum(axis=2)
        idx, short = _first_k_unique(cands, K, rng)
        for s in np.nonzero(short)[0]:
            have = np.unique(cands[s])
            extra = np.setdiff1d(rng.choice(G, size=4 * K, replace=False), have)
            idx[s] = np.concatenate([have, extra])[:K]
    else:
        raise ValueError("mode must be 'uniform' or 'clustered'")
    lnp = -rng.exponential(3.0, size=(S, K))
    np.clip(lnp, -10.0, None, out=lnp)
    lnp[:, 0] = 0.0
    # a few stars have trailing padded rows (arbitrary index, -inf value)
    nmask = int(round(mask_frac * S))
    if nmask:
        who = rng.choice(S, size=nmask, replace=False)
        npad = rng.integers(1, min(10, K - 1) + 1, size=nmask)
        for s, n in zip(who, npad):
            lnp[s, K - n:] = -np.inf
            idx[s, K - n:] = rng.integers(0, G, size=n)
    return {
        "vals": np.ascontiguousarray(lnp.T),
        "indxs": np.ascontiguousarray(idx.T),
    }


def make_star_data(moddata, axes, S, K, mode="uniform", seed=2, mask_frac=0.02):
    """``star_lnpgriddata`` as ``get_likelihoods`` leaves it (``helpers.py:35-41``):
    linear ``exp(lnp)/prior_weight[idx]`` where finite, ``-inf`` elsewhere."""
    lnp = make_lnp(moddata, axes, S, K, mode=mode, seed=seed, mask_frac=mask_frac)
    vals, indxs = lnp["vals"], lnp["indxs"]
    fin = np.isfinite(vals)
    out = np.full(vals.shape, -np.inf)
    out[fin] = np.exp(vals[fin]) / moddata["prior_weight"][indxs[fin]]
    return {"vals": out, "indxs": indxs}


def make_walkers(model, W, seed=3):
    """``start * (1 + 0.1 randn)`` as in ``singlepop_dust_model.py:366``, redrawn until every
    walker lies inside the flat-prior box so each one costs a full evaluation."""
    rng = np.random.default_rng(seed)
    start = np.asarray(model.start_params()[1], dtype=float)
    out = np.empty((W, start.size))
    n = 0
    while n < W:
        cand = start * (1.0 + 0.1 * rng.standard_normal(start.size))
        if model.lnprior(cand) == 0.0:
            out[n] = cand
            n += 1
    return out


def make_pixel_problem(n_pixels=64, mean_stars=50, axes=(21, 3, 9, 8, 3, 1), K=20, mode="clustered",
                       seeds=(1, 2, 4)):
    """BASELINE config 5 in miniature: ``n_pixels`` sky pixels with Poisson(mean_stars) stars each
    over ONE shared grid.  Returns (settings, beast_moddata, star_lnpgriddata, star_offsets)."""
    rng = np.random.default_rng(seeds[2])
    counts = rng.poisson(mean_stars, size=n_pixels)
    offsets = np.zeros(n_pixels + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    moddata = make_grid(axes, seed=seeds[0])
    stars = make_star_data(moddata, axes, int(offsets[-1]), K, mode=mode, seed=seeds[1])
    return docs_example_settings(), moddata, stars, offsets


def make_problem(name="cfg1", mode="uniform", seeds=(1, 2, 3)):
    """(settings, beast_moddata, star_lnpgriddata, config dict) for a named config."""
    cfg = CONFIGS[name]
    moddata = make_grid(cfg["axes"], seed=seeds[0])
    stars = make_star_data(moddata, cfg["axes"], cfg["S"], cfg["K"], mode=mode, seed=seeds[1])
    return docs_example_settings(), moddata, stars, dict(cfg, name=name, mode=mode)
