"""
Host-side companions of the ensemble likelihood, same names and argument meaning as
``/root/reference/megabeast/helpers.py``:

``get_likelihoods``              (:15-43)   one-time ingest, the exp/divide runs on the GPU
``precompute_mass_multipliers``  (:46-102)  once per fit, cold, stays on the host
``get_predicted_num_stars``      (:105-164) per call in the reference; here the device computes
                                            it inside the likelihood (kernels K1b + K3) and this
                                            host form exists for callers that want the number
"""
import numpy as np

from . import _cabi
from .priormodel import resolve_prior_classes

__all__ = [
    "get_likelihoods",
    "likelihoods_from_lnp",
    "precompute_mass_multipliers",
    "get_predicted_num_stars",
]


def likelihoods_from_lnp(lnpdata, beast_model_data, device=0):
    """``vals = exp(lnp) / prior_weight[indxs]`` on the finite entries (helpers.py:33-41), in
    place on ``lnpdata["vals"]``, computed by the CUDA kernel ``mb_ingest_lnp``."""
    lib = _cabi.load()
    vals = lnpdata["vals"]
    if not (isinstance(vals, np.ndarray) and vals.dtype == np.float64 and vals.flags.c_contiguous):
        vals = np.ascontiguousarray(vals, dtype=np.float64)
    indxs = np.ascontiguousarray(lnpdata["indxs"], dtype=np.int64)
    pw = np.ascontiguousarray(beast_model_data["prior_weight"], dtype=np.float64)
    K, S = indxs.shape
    _cabi.check(
        lib.mb_likelihoods_from_lnp(
            device, K, S, indxs.ctypes.data_as(_cabi.c_int64_p), _cabi.dptr(vals), pw.size,
            _cabi.dptr(pw),
        )
    )
    lnpdata["vals"] = vals
    return lnpdata


def get_likelihoods(ppdf_file, beast_model_data, device=0):
    """Read the saved sparse BEAST posteriors and divide out the BEAST prior weights
    (helpers.py:15-43).  The reader is BEAST's ``read_lnp_data``; without BEAST installed an
    ``.npz`` with ``vals``/``indxs`` of the same layout is accepted."""
    try:
        from beast.tools.read_beast_data import read_lnp_data

        lnpdata = read_lnp_data(ppdf_file)
    except ImportError:
        with np.load(ppdf_file) as f:
            vals = np.array(f["vals"], dtype=np.float64)
            lnpdata = {"vals": vals - vals.max(axis=0, keepdims=True), "indxs": np.array(f["indxs"])}
    return likelihoods_from_lnp(lnpdata, beast_model_data, device=device)


def _trapz(y, x):
    # a numpy scalar like np.trapz's: on a grid with a single M_ini value the reference gets 0/0 = nan
    # (helpers.py:72-73,92), not a ZeroDivisionError
    return np.float64(np.sum(np.diff(x) * (y[1:] + y[:-1]) / 2.0))


def precompute_mass_multipliers(bphysparams, physmodmass):
    """
    Value to multiply the SFR by to get the mass formed per (age, Z) bin, plus the bin
    coordinates and the mean stellar mass (helpers.py:46-102).  Same outputs as the reference;
    the per-bin mass range comes from one grouped min/max pass instead of a boolean mask per bin.
    """
    _, _, _, compute_bin_boundaries, _ = resolve_prior_classes()
    mini = np.asarray(bphysparams["M_ini"], dtype=float)
    loga = np.asarray(bphysparams["logA"], dtype=float)
    zmet = np.asarray(bphysparams["Z"], dtype=float)
    masspts = np.logspace(np.log10(mini.min()), np.log10(mini.max()), 100)
    massprior = np.asarray(physmodmass(masspts), dtype=float)
    totmass = _trapz(massprior, masspts)
    with np.errstate(divide="ignore", invalid="ignore"):
        avemass = _trapz(masspts * massprior, masspts) / totmass

    ages, ia = np.unique(loga, return_inverse=True)
    mets, iz = np.unique(zmet, return_inverse=True)
    widths = np.diff(10 ** compute_bin_boundaries(ages))
    nb = ages.size * mets.size
    b = ia * mets.size + iz
    lo = np.full(nb, np.inf)
    hi = np.full(nb, -np.inf)
    np.minimum.at(lo, b, mini)
    np.maximum.at(hi, b, mini)
    massmult = np.zeros((ages.size, mets.size))
    for i in range(ages.size):
        for j in range(mets.size):
            k = i * mets.size + j
            if not np.isfinite(lo[k]):
                raise ValueError("min() arg is an empty sequence")  # what the reference hits
            sel = (masspts >= lo[k]) & (masspts <= hi[k])
            with np.errstate(divide="ignore", invalid="ignore"):
                massmult[i, j] = widths[i] * _trapz(massprior[sel], masspts[sel]) / totmass
    return {"massmult": massmult, "ages": ages, "Zs": mets, "avemass": avemass}


def get_predicted_num_stars(massmult_info, bphysparams, bphysmod, physmodage):
    """Expected number of detected stars for a physics model given on the grid
    (helpers.py:105-164), by grouped sums over the (age, Z) bins."""
    with np.errstate(divide="ignore", invalid="ignore"):
        w = bphysmod * bphysparams["grid_weight"]
        w = w / np.sum(w)
    ages, mets = massmult_info["ages"], massmult_info["Zs"]
    ageprior = np.asarray(physmodage(ages), dtype=float)
    ia = np.searchsorted(ages, bphysparams["logA"])
    iz = np.searchsorted(mets, bphysparams["Z"])
    b = ia * len(mets) + iz
    nb = len(ages) * len(mets)
    s = np.bincount(b, weights=w, minlength=nb)
    c = np.bincount(b, weights=w * bphysparams["completeness"], minlength=nb)
    nexp = (ageprior[:, None] * massmult_info["massmult"] / len(mets) / massmult_info["avemass"]).ravel()
    ok = s > 0
    return float(np.sum(nexp[ok] * c[ok] / s[ok]))
