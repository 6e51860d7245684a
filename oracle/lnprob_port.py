"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by ``megabeast_b200``; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs use it,
as the checker and as the CPU arm timed beside the GPU.

numpy restatement ("port") of the reference's ensemble likelihood

    /root/reference/megabeast/singlepop_dust_model.py   MB_Model (:15-274), lnprob (:277-304)
    /root/reference/megabeast/helpers.py                precompute_mass_multipliers (:46-102)
                                                        get_predicted_num_stars (:105-164)
                                                        get_likelihoods (:15-43)

Pinning status
--------------
* The reference has NO test, golden vector or fixture for this path
  (``megabeast/tests/basic_test.py:4-7`` only builds an empty ``mbsettings``), so the
  port is pinned against *outputs of the reference itself executed in the build
  container*: ``tests/golden/make_golden.py`` imports the unmodified reference modules
  (``oracle/reference_loader.py``) and freezes inputs + outputs as ``tests/golden/*.npz``;
  ``tests/test_oracle.py`` checks this port against every one of them, and against the
  live reference whenever ``/root/reference`` is present.
* The prior-model arithmetic the reference imports from ``beast`` is NOT in the
  reference tree; it is restated in ``oracle/priors.py`` -- parity UNPINNED for that part.

Two evaluation strategies, same numbers (tests hold them to 1e-13 relative):

``loop=True``   follows the reference's control flow literally -- one Python iteration per
                star with five fancy-index gathers (:194-223) and two boolean masks over
                the whole grid per (age, Z) bin (helpers.py:143-162).  This is the CPU
                arm ``bench.py`` times: it costs what the reference costs.
``loop=False``  the same sums written as whole-array numpy expressions, used by the
                parity tests at sizes where the loop would take minutes.
"""
import numpy as np
from scipy.special import gammaln

from . import priors

__all__ = [
    "PortModel",
    "lnprob",
    "mass_multipliers",
    "predicted_num_stars",
    "likelihoods_from_lnp",
    "gathered_indices",
]

PARAM_ORDER = ("logA", "M_ini", "Av", "Rv", "fA")  # singlepop_dust_model.py:29


def _copy_model_dict(d):
    out = {}
    for k, v in d.items():
        out[k] = list(v) if isinstance(v, (list, tuple)) and k in ("values",) else v
    return out


class PortModel:
    """
    State an ``MB_Model`` holds after ``__init__`` (singlepop_dust_model.py:21-80), kept
    as plain dicts; evaluation never mutates it (the reference writes phi into the dicts,
    :136-143 -- the port passes phi along instead and applies it to a scratch copy).
    """

    def __init__(self, settings):
        star, dust = settings.stellar_model, settings.fd_model
        self.params = list(PARAM_ORDER)
        self.physics_model = {}
        for cparam in self.params:
            if cparam in star:
                src = star[cparam]
            elif cparam in dust:
                src = dust[cparam]
            else:
                raise ValueError("requested parameter not in mbsetting file")  # :40
            mod = {"name": src["name"], "varnames": src["varnames"], "prior": src["prior"]}
            for vname, vinit in zip(src["varnames"], src["varinit"]):
                mod[vname] = list(vinit) if isinstance(vinit, (list, tuple)) else vinit
            if cparam in star:
                if cparam == "logA":
                    mod["x"] = list(star["x"])  # :51
                    mod["nsubvars"] = len(mod["x"])  # :52-54
                # M_ini: no "nsubvars" key -- the reference never sets one (:58-62)
            else:
                mod["nsubvars"] = 1  # :64
            self.physics_model[cparam] = mod
        self.compute_N_stars = self.physics_model["logA"]["prior"]["name"] != "fixed"  # :71-74
        self.massmultipliers = None

    # -- parameter vector layout (:82-108, :131-143) ------------------------------------
    def active(self):
        return [p for p in self.params if self.physics_model[p]["prior"]["name"] != "fixed"]

    def start_params(self):
        names, values = [], []
        for p in self.params:
            mod = self.physics_model[p]
            if mod["prior"]["name"] == "fixed":
                continue
            for vname in mod["varnames"]:
                cur = np.atleast_1d(mod[vname])
                if len(cur) > 1:
                    for j, v in enumerate(cur):
                        names.append(f"{p}_{vname}{j + 1}")
                        values.append(v)
                else:
                    names.append(f"{p}_{vname}")
                    values.append(mod[vname])
        return names, values

    def with_phi(self, phi):
        """Per-parameter model dicts with phi written in, in the order of :133-143."""
        out, k = {}, 0
        for p in self.params:
            mod = _copy_model_dict(self.physics_model[p])
            if mod["prior"]["name"] != "fixed":
                for vname in mod["varnames"]:
                    if mod["nsubvars"] > 1:  # KeyError for M_ini, as in the reference
                        vals = list(mod["values"])  # key name hard-coded at :139
                        for j in range(mod["nsubvars"]):
                            vals[j] = phi[k]
                            k += 1
                        mod["values"] = vals
                    else:
                        mod[vname] = phi[k]
                        k += 1
            out[p] = mod
        return out

    # -- lnprior (:231-274) ----------------------------------------------------------------
    def lnprior(self, phi):
        k = 0
        for p in self.params:
            mod = self.physics_model[p]
            prior = mod["prior"]
            if prior["name"] == "fixed":
                continue
            if prior["name"] != "flat":
                raise ValueError(f"{prior['name']} prior not supported")  # :273
            for i, _ in enumerate(mod["varnames"]):
                if mod["nsubvars"] > 1:
                    lo, hi = prior["minmax"][0], prior["minmax"][1]  # :260-262
                    for j in range(len(np.atleast_1d(prior["minmax"][i]))):
                        if phi[k] < lo[j] or phi[k] > hi[j]:
                            return -np.inf
                        k += 1
                else:
                    lo, hi = prior["minmax"][i][0], prior["minmax"][i][1]  # :267-268
                    if phi[k] < lo or phi[k] > hi:
                        return -np.inf
                    k += 1
        return 0.0

    # -- phase A: prior weights over the grid (:131-148) ---------------------------------
    def physmod(self, phi, moddata):
        mods = self.with_phi(phi)
        out = np.full(len(moddata["Av"]), 1.0)
        for p in self.params:
            if mods[p]["prior"]["name"] != "fixed":
                out *= priors.evaluate_prior_model(mods[p], moddata[p])
        return out, mods

    # -- lnlike (:110-229) -------------------------------------------------------------------
    def lnlike(self, phi, stars, moddata, loop=True):
        physmod, mods = self.physmod(phi, moddata)
        indxs, vals = stars["indxs"], stars["vals"]
        n_stars = indxs.shape[1]  # :157

        if self.compute_N_stars:  # :159-186
            if self.massmultipliers is None:
                self.massmultipliers = mass_multipliers(moddata, mods["M_ini"])
            pred = predicted_num_stars(
                self.massmultipliers, moddata, physmod, mods["logA"], loop=loop
            )
            with np.errstate(divide="ignore", invalid="ignore"):
                total = n_stars * np.log(pred) - pred - gammaln(n_stars + 1)
        else:
            total = 0.0

        pw, gw, comp = moddata["prior_weight"], moddata["grid_weight"], moddata["completeness"]
        if loop:
            for i in range(n_stars):  # :194-223
                keep = np.isfinite(vals[:, i])
                rows = indxs[:, i][keep]
                star = np.sum(vals[:, i][keep] * physmod[rows] * pw[rows] * gw[rows] * comp[rows])
                if not np.isfinite(star):
                    raise ValueError("invidual integrated star prob is not finite")  # :217
                if star != 0.0:
                    total += np.log(star)
        else:
            star = star_sums(vals, indxs, physmod, pw, gw, comp)
            if not np.all(np.isfinite(star)):
                raise ValueError("invidual integrated star prob is not finite")
            nz = star != 0.0
            # sequential accumulation order of the reference differs from np.sum at the
            # 1e-16 level only
            total = total + np.sum(np.log(star[nz]))

        if total == 0.0:  # :225-227
            print(total)
            raise SystemExit
        return total


def star_sums(vals, indxs, physmod, pw, gw, comp):
    """Per-star ``sum_k vals*physmod[idx]*prior_weight[idx]*grid_weight[idx]*completeness[idx]``
    over the finite entries (:196-213) as one whole-array expression."""
    keep = np.isfinite(vals)
    rows = np.where(keep, indxs, 0)
    with np.errstate(invalid="ignore"):
        terms = np.where(keep, vals, 0.0) * physmod[rows] * pw[rows] * gw[rows] * comp[rows]
    terms[~keep] = 0.0
    return terms.sum(axis=0)


def gathered_indices(stars):
    """The grid rows each star actually dereferences (:196-198), star-major, as a flat
    int64 array plus per-star offsets -- the quantity the GPU path must match bit for bit."""
    keep = np.isfinite(stars["vals"])
    keepT = keep.T
    flat = stars["indxs"].T[keepT].astype(np.int64)
    offs = np.zeros(keepT.shape[0] + 1, dtype=np.int64)
    np.cumsum(keepT.sum(axis=1), out=offs[1:])
    return flat, offs


def lnprob(phi, model, stars, moddata, loop=True):
    """:277-304 -- lnlike is not evaluated for a walker outside the prior box."""
    lp = model.lnprior(phi)
    if not np.isfinite(lp):
        return -np.inf
    return lp + model.lnlike(phi, stars, moddata, loop=loop)


def _trapezoid(y, x):
    # same arithmetic as np.trapz (helpers.py:72,73,92): sum(dx * (y[1:] + y[:-1]) / 2)
    # (a numpy scalar, as np.trapz returns: 0/0 on a grid with a single M_ini value is nan there, not an exception)
    y, x = np.asarray(y, float), np.asarray(x, float)
    return np.float64(np.sum(np.diff(x) * (y[1:] + y[:-1]) / 2.0))


def mass_multipliers(moddata, mass_model):
    """helpers.py:46-102 -- once per fit while the IMF is fixed (singlepop_dust_model.py:161-170)."""
    mini, loga, zmet = moddata["M_ini"], moddata["logA"], moddata["Z"]
    lo, hi = min(mini), max(mini)
    mpts = np.logspace(np.log10(lo), np.log10(hi), 100)
    imf = priors.evaluate_prior_model(mass_model, mpts)
    tot = _trapezoid(imf, mpts)
    with np.errstate(divide="ignore", invalid="ignore"):
        avemass = _trapezoid(mpts * imf, mpts) / tot
    ages = np.unique(loga)
    widths = np.diff(10 ** priors.bin_boundaries(ages))
    mets = np.unique(zmet)
    mult = np.zeros((len(ages), len(mets)))
    for i, a in enumerate(ages):
        for j, z in enumerate(mets):
            sel = (loga == a) & (zmet == z)
            mlo, mhi = min(mini[sel]), max(mini[sel])
            inr = (mpts >= mlo) & (mpts <= mhi)
            with np.errstate(divide="ignore", invalid="ignore"):
                mult[i, j] = widths[i] * _trapezoid(imf[inr], mpts[inr]) / tot
    return {"massmult": mult, "ages": ages, "Zs": mets, "avemass": avemass}


def predicted_num_stars(massinfo, moddata, physmod, age_model, loop=True):
    """helpers.py:105-164."""
    with np.errstate(divide="ignore", invalid="ignore"):
        w = physmod * moddata["grid_weight"]
        w = w / np.sum(w)
    ages, mets = massinfo["ages"], massinfo["Zs"]
    ageprior = priors.evaluate_prior_model(age_model, ages)
    metprior = np.full(len(mets), 1.0)
    metprior /= np.sum(metprior)
    comp = moddata["completeness"]
    total = 0.0
    if loop:
        for i, a in enumerate(ages):
            for j, z in enumerate(mets):
                sel = (moddata["logA"] == a) & (moddata["Z"] == z)
                nexp = ageprior[i] * metprior[j] * massinfo["massmult"][i, j] / massinfo["avemass"]
                cw = w[sel]
                s = np.sum(cw)
                if s > 0:
                    cw = cw * (nexp / s)
                    cw = cw * comp[sel]
                    total += np.sum(cw)
        return total
    ia = np.searchsorted(ages, moddata["logA"])
    iz = np.searchsorted(mets, moddata["Z"])
    b = ia * len(mets) + iz
    nb = len(ages) * len(mets)
    s = np.bincount(b, weights=w, minlength=nb)
    c = np.bincount(b, weights=w * comp, minlength=nb)
    nexp = (ageprior[:, None] * metprior[None, :] * massinfo["massmult"] / massinfo["avemass"]).ravel()
    ok = s > 0
    return float(np.sum(nexp[ok] * c[ok] / s[ok]))


def likelihoods_from_lnp(lnpdata, moddata):
    """helpers.py:33-41 -- log posteriors to linear likelihood x grid weight."""
    vals = np.array(lnpdata["vals"], dtype=float)
    indxs = lnpdata["indxs"]
    for i in range(indxs.shape[1]):
        keep = np.isfinite(vals[:, i])
        vals[keep, i] = np.exp(vals[keep, i]) / moddata["prior_weight"][indxs[:, i][keep]]
    return {"vals": vals, "indxs": indxs}
