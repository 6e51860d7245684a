"""
Host-side prior-model callables with the interface ``MB_Model`` expects from
``beast.physicsmodel.priormodel`` (``/root/reference/megabeast/singlepop_dust_model.py:6-8,
55-67,146-148``): built from a parameter dictionary, keeps a reference to it, evaluates
``ndarray -> ndarray`` with whatever values the dictionary holds at call time.

They are NOT on the likelihood hot path -- the CUDA kernels (``prior_eval`` in mb_kernels.cuh) evaluate the
same formulas on the device.  The host versions serve three cold uses: the IMF integral of
``precompute_mass_multipliers`` (once per fit), the ``physics_model[...]["model"]`` entries
other code may call, and the construction-time cross-check of the device formulas.

If the real ``beast`` package is importable its classes are used instead
(``resolve_prior_classes``), so the drop-in follows the installed BEAST; model names the device
does not know analytically are then evaluated by that callable on the parameter's distinct grid
values and shipped as a table (``MB_TABULATED``).

Formula anchors inside the reference tree: lognormal ``megabeast/old/ensemble_model.py:13-47``,
mixture weights ``:122-123``; names ``docs/megabeast/settings.rst:39,47,78``.  ``beast`` itself
is not available in the build environment, so these are restatements of its documented
behaviour.
"""
import numpy as np

__all__ = [
    "PriorModel",
    "PriorDustModel",
    "PriorAgeModel",
    "PriorMassModel",
    "compute_bin_boundaries",
    "resolve_prior_classes",
]


def _lognormal(x, peak, sigma, scale=1.0):
    y = np.zeros(np.shape(x), dtype=float)
    good = x > 0
    xs = x[good]
    mu = np.log(peak) + sigma ** 2
    y[good] = scale * (1.0 / np.sqrt(2.0 * np.pi)) / (xs * sigma) * np.exp(
        -0.5 * ((np.log(xs) - mu) / sigma) ** 2
    )
    return y


def _flat(m, x):
    return np.full(np.shape(x), float(m.get("amp", 1.0)))


def _lognormal_model(m, x):
    return _lognormal(x, m["mean"], m["sigma"])


def _two_lognormal_model(m, x):
    frac2 = 1.0 / (m["N1_to_N2"] + 1.0)
    return _lognormal(x, m["mean1"], m["sigma1"], 1.0 - frac2) + _lognormal(
        x, m["mean2"], m["sigma2"], frac2
    )


def _bins_histo(m, x):
    nodes = np.asarray(m["x"], dtype=float)
    heights = np.asarray(m["values"], dtype=float)
    if heights.size == nodes.size - 1:
        heights = np.concatenate([heights, [0.0]])
    if x.size and (x.max() > nodes.max() or x.min() < nodes.min()):
        raise ValueError("requested x outside of model x range")
    # left-constant steps; the last node carries the last height
    pos = np.minimum(np.searchsorted(nodes, x, side="right") - 1, nodes.size - 1)
    return heights[pos]


def _bins_interp(m, x):
    return np.interp(x, np.asarray(m["x"], dtype=float), np.asarray(m["values"], dtype=float))


def _exponential_age(m, x):
    return np.exp(-1.0 * (10.0 ** x) / (float(m["tau"]) * 1e9))


def _kroupa(m, x):
    lo, hi = 0.08, 0.5
    s_lo, s_mid, s_hi = -0.3, -1.3, -2.3
    k_mid = hi ** (s_hi - s_mid)
    k_lo = k_mid * lo ** (s_mid - s_lo)
    return np.where(x >= hi, x ** s_hi, np.where(x >= lo, k_mid * x ** s_mid, k_lo * x ** s_lo))


def _salpeter(m, x):
    return x ** (-float(m.get("slope", 2.35)))


_TABLE = {
    "flat": _flat,
    "lognormal": _lognormal_model,
    "two_lognormal": _two_lognormal_model,
    "bins_histo": _bins_histo,
    "bins_interp": _bins_interp,
    "exponential": _exponential_age,
    "kroupa": _kroupa,
    "salpeter": _salpeter,
}


class PriorModel:
    names = tuple(_TABLE)

    def __init__(self, model):
        if model["name"] not in self.names:
            raise NotImplementedError(f"{model['name']} is not an allowed model")
        self.model = model

    def __call__(self, x):
        return _TABLE[self.model["name"]](self.model, np.asarray(x, dtype=float))


class PriorDustModel(PriorModel):
    names = ("flat", "lognormal", "two_lognormal")


class PriorAgeModel(PriorModel):
    names = ("flat", "bins_histo", "bins_interp", "exponential")


class PriorMassModel(PriorModel):
    names = ("kroupa", "salpeter", "flat")


def compute_bin_boundaries(tab):
    """Edges between consecutive tabulated values, outer edges mirrored: ``len(tab)+1`` values
    (the contract ``megabeast/helpers.py:78-79,95`` relies on)."""
    tab = np.asarray(tab, dtype=float)
    if tab.size == 1:
        return np.array([tab[0] - 0.5, tab[0] + 0.5])
    mid = 0.5 * (tab[1:] + tab[:-1])
    return np.concatenate([[tab[0] - (mid[0] - tab[0])], mid, [tab[-1] + (tab[-1] - mid[-1])]])


def resolve_prior_classes():
    """(PriorAgeModel, PriorMassModel, PriorDustModel, compute_bin_boundaries, source): BEAST's
    own when installed, else the restatements above."""
    try:
        from beast.physicsmodel.priormodel import PriorAgeModel as A
        from beast.physicsmodel.priormodel import PriorDustModel as D
        from beast.physicsmodel.priormodel import PriorMassModel as M
        from beast.physicsmodel.grid_weights_stars import compute_bin_boundaries as cbb

        return A, M, D, cbb, "beast"
    except Exception:
        return PriorAgeModel, PriorMassModel, PriorDustModel, compute_bin_boundaries, "megabeast_b200"
