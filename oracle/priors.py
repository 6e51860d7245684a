"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
may import it, and only as the checker / the CPU arm that is timed beside the GPU.

CPU restatement of the *third-party* prior-model arithmetic that megaBEAST's hot
path calls but does not contain.

Source status
-------------
The arithmetic lives in the ``beast`` package (github.com/BEAST-Fitting/beast),
which the reference pins only to a moving branch (``setup.cfg:27``
``beast@git+https://github.com/BEAST-Fitting/beast.git@master``) and which is NOT
installed in this container and not vendored under ``/root/reference``.
==> **parity for these functions is UNPINNED**: they are a restatement of beast's
published behaviour (``beast.physicsmodel.priormodel``,
``beast.physicsmodel.priormodel_functions``, ``beast.physicsmodel.grid_weights_stars``)
anchored on what *is* in the reference tree:

* lognormal / two-lognormal formula: ``megabeast/old/ensemble_model.py:13-47`` and
  ``:50-81`` (a copy of beast's ``_lognorm``/``_two_lognorm`` kept there "temporary for
  development"); mixture weights ``N1 = 1 - 1/(r+1)``, ``N2 = 1/(r+1)``
  (``old/ensemble_model.py:122-123``).
* model names and variable names: ``docs/megabeast/settings.rst:34-104``
  (``bins_histo``/``values``/``x``, ``kroupa``/``slope``, ``lognormal``/``mean``/``sigma``).
* call contract: a model object is built from the parameter dict, keeps a reference
  to that dict, and maps ``ndarray -> ndarray`` of the same shape
  (``megabeast/singlepop_dust_model.py:55-67,146-148``; ``megabeast/helpers.py:71,135``).
* ``compute_bin_boundaries`` must return ``len(x)+1`` edges
  (``megabeast/helpers.py:78-79,95``).

Everything the megaBEAST code itself does (the part that IS in ``/root/reference``)
is pinned separately by executing the reference -- see ``oracle/lnprob_port.py`` and
``tests/golden/make_golden.py``.
"""
import numpy as np

__all__ = [
    "lognormal_density",
    "two_lognormal_density",
    "step_histogram",
    "kroupa_imf",
    "salpeter_imf",
    "bin_boundaries",
    "evaluate_prior_model",
    "DUST_MODELS",
    "AGE_MODELS",
    "MASS_MODELS",
]

DUST_MODELS = ("flat", "lognormal", "two_lognormal")
AGE_MODELS = ("flat", "bins_histo", "bins_interp", "exponential")
MASS_MODELS = ("kroupa", "salpeter", "flat")

_INV_SQRT_2PI = 1.0 / np.sqrt(2.0 * np.pi)


def lognormal_density(x, mode, sigma, amp=1.0):
    """
    Lognormal in x whose maximum sits at ``mode``.

    Anchor: ``megabeast/old/ensemble_model.py:13-47`` -- ``mu = ln(mode) + sigma**2``;
    value ``amp / (x sigma sqrt(2 pi)) * exp(-0.5 ((ln x - mu)/sigma)**2)`` for x > 0
    and exactly 0 for x <= 0.  Operation order follows that anchor so the float64
    result is the same number.
    """
    x = np.asarray(x, dtype=float)
    out = np.zeros(x.shape, dtype=float)
    pos = x > 0
    xp = x[pos]
    mu = np.log(mode) + sigma ** 2
    norm = _INV_SQRT_2PI / (xp * sigma)
    out[pos] = amp * norm * np.exp(-0.5 * ((np.log(xp) - mu) / sigma) ** 2)
    return out


def two_lognormal_density(x, mode1, mode2, sigma1, sigma2, n1_to_n2):
    """
    Two-component lognormal mixture, ``megabeast/old/ensemble_model.py:50-81`` with the
    weights of ``:122-123``.
    """
    n1 = 1.0 - 1.0 / (n1_to_n2 + 1.0)
    n2 = 1.0 / (n1_to_n2 + 1.0)
    return lognormal_density(x, mode1, sigma1, n1) + lognormal_density(
        x, mode2, sigma2, n2
    )


def step_histogram(x, nodes, heights):
    """
    ``bins_histo``: left-constant steps through (nodes, heights) -- the behaviour of
    ``scipy.interpolate.interp1d(nodes, heights, kind="zero")``: for
    ``nodes[i] <= x < nodes[i+1]`` the value is ``heights[i]``; ``x == nodes[-1]`` gives
    ``heights[-1]``.  A request outside ``[min(nodes), max(nodes)]`` is an error (beast
    raises ``ValueError``); one fewer height than nodes means a trailing 0.0.
    """
    x = np.asarray(x, dtype=float)
    nodes = np.asarray(nodes, dtype=float)
    heights = np.asarray(heights, dtype=float)
    if len(heights) == len(nodes) - 1:
        heights = np.append(heights, 0.0)
    if len(heights) != len(nodes):
        raise ValueError("bins_histo: len(values) must be len(x) or len(x)-1")
    if x.size and (np.any(x > nodes.max()) or np.any(x < nodes.min())):
        raise ValueError("requested x outside of model x range")
    idx = np.searchsorted(nodes, x, side="right") - 1
    idx = np.clip(idx, 0, len(nodes) - 1)
    return heights[idx]


def kroupa_imf(m):
    """
    Kroupa broken power law dN/dM: exponents -0.3 (M < 0.08), -1.3 (0.08 <= M < 0.5),
    -2.3 (M >= 0.5), continuous across the breaks, equal to ``M**-2.3`` above 0.5.
    (The ``slope`` variable of ``settings.rst:49-50`` is not used by this form.)
    """
    m = np.asarray(m, dtype=float)
    b1, b2 = 0.08, 0.5
    a0, a1, a2 = -0.3, -1.3, -2.3
    c1 = b2 ** a2 / b2 ** a1  # continuity at 0.5
    c0 = c1 * b1 ** a1 / b1 ** a0  # continuity at 0.08
    out = np.where(m >= b2, m ** a2, np.where(m >= b1, c1 * m ** a1, c0 * m ** a0))
    return out


def salpeter_imf(m, slope=2.35):
    m = np.asarray(m, dtype=float)
    return m ** (-slope)


def bin_boundaries(tab):
    """
    ``compute_bin_boundaries``: edges halfway between consecutive entries; the two outer
    edges mirror the adjacent half-spacing, so ``len(tab)+1`` edges come back
    (required by ``megabeast/helpers.py:78-79,95``).
    """
    tab = np.asarray(tab, dtype=float)
    if len(tab) == 1:
        # a single entry has no spacing to mirror; unit-wide bin in the tabulated unit
        return np.array([tab[0] - 0.5, tab[0] + 0.5])
    half = 0.5 * np.diff(tab)
    edges = np.empty(len(tab) + 1, dtype=float)
    edges[1:-1] = tab[:-1] + half
    edges[0] = tab[0] - half[0]
    edges[-1] = tab[-1] + half[-1]
    return edges


def evaluate_prior_model(model, x):
    """
    Dispatch on ``model["name"]`` the way ``beast.physicsmodel.priormodel.PriorModel``
    does, reading the variables from the (possibly just mutated) dict
    (``megabeast/singlepop_dust_model.py:136-148`` writes phi into that dict before
    every call).
    """
    name = model["name"]
    x = np.asarray(x, dtype=float)
    if name == "flat":
        return np.full(x.shape, float(model.get("amp", 1.0)))
    if name == "lognormal":
        return lognormal_density(x, model["mean"], model["sigma"])
    if name == "two_lognormal":
        return two_lognormal_density(
            x,
            model["mean1"],
            model["mean2"],
            model["sigma1"],
            model["sigma2"],
            model["N1_to_N2"],
        )
    if name == "bins_histo":
        return step_histogram(x, model["x"], model["values"])
    if name == "bins_interp":
        return np.interp(x, np.asarray(model["x"], float), np.asarray(model["values"], float))
    if name == "exponential":
        # age prior: e-folding time tau in Gyr, x is log10(age/yr)
        return np.exp(-1.0 * (10.0 ** x) / (float(model["tau"]) * 1e9))
    if name == "kroupa":
        return kroupa_imf(x)
    if name == "salpeter":
        return salpeter_imf(x, float(model.get("slope", 2.35)))
    raise NotImplementedError(f"{name} is not an allowed model")
