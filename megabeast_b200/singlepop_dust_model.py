"""
Drop-in for ``/root/reference/megabeast/singlepop_dust_model.py``: same class, functions, call
signatures, return values and exceptions -- the ensemble likelihood itself runs on the GPU.

    MB_Model(params)                                             (:21-80)
    MB_Model.start_params() -> (names, values)                   (:82-108)
    MB_Model.lnlike(phi, star_lnpgriddata, beast_moddata)        (:110-229)
    MB_Model.lnprior(phi)                                        (:231-274)
    lnprob(phi, megabeast_model, star_lnpgriddata, beast_moddata) (:277-304)
    fit_ensemble(megabeast_model, star_lnpgriddata, beast_moddata) (:325-383)

Additive extension (not a signature change): ``phi`` may be a ``(W, ndim)`` array, in which
case ``lnlike``/``lnprior``/``lnprob`` return ``W`` values from ONE device evaluation -- what
``emcee.EnsembleSampler(..., vectorize=True)`` feeds.

The data dictionaries arrive as arguments on every call, exactly as in the reference; their
device-resident form is built on first use and cached on the identity of the two dictionaries
(the reference passes the same objects on every call, :368-370).

There is no CPU fallback: without the CUDA library or a device, ``lnlike`` raises.
"""
import numpy as np

from . import _cabi
from .engine import LikelihoodEngine, ModelSpec
from .helpers import precompute_mass_multipliers
from .priormodel import resolve_prior_classes
from .sampler import StretchMoveSampler

__all__ = ["MB_Model", "lnprob", "fit_ensemble"]

_NONFINITE_MSG = "invidual integrated star prob is not finite"  # sic, :217


class MB_Model:
    """
    MegaBEAST model that provides member functions to compute the likelihood and priors for a
    specific physical model.

    Parameters
    ----------
    params : object with ``stellar_model`` and ``fd_model`` dictionaries (an ``mbsettings``)
    devices : list of CUDA ordinals to shard the stars over (default: device 0)
    mode : "auto" | "factor" | "table_unique" | "table_grid"  
    table_precision : "fp64" (default; 1e-9 parity with the reference) | "fp32" (32-bit factor
        tables on chip -- each factor rounded to 21 significant bits -- sums still fp64; about
        8 % faster, measured error 7e-9 relative on lnlike at 100k stars, stated tolerance 1e-7)
    verbose : print the three setup lines the reference prints (:31-33)
    """

    def __init__(self, params, devices=None, mode="auto", verbose=False, table_precision="fp64",
                 split_inner=None):
        self.star_model = params.stellar_model
        self.dust_model = params.fd_model
        age_cls, mass_cls, dust_cls, _, self.prior_source = resolve_prior_classes()

        self.params = ["logA", "M_ini", "Av", "Rv", "fA"]
        self.physics_model = {}
        if verbose:
            print(self.params)
            print(self.star_model.keys())
            print(self.dust_model.keys())
        for cparam in self.params:
            if cparam in self.star_model.keys():
                src, stellar = self.star_model[cparam], True
            elif cparam in self.dust_model.keys():
                src, stellar = self.dust_model[cparam], False
            else:
                raise ValueError("requested parameter not in mbsetting file")
            mod = {"name": src["name"], "varnames": src["varnames"], "prior": src["prior"]}
            for vname, vinit in zip(src["varnames"], src["varinit"]):
                mod[vname] = vinit
            self.physics_model[cparam] = mod
            if stellar:
                if cparam == "logA":
                    mod["x"] = self.star_model["x"]
                    mod["nsubvars"] = len(mod["x"])
                    mod["model"] = age_cls(mod)
                elif cparam == "M_ini":
                    mod["model"] = mass_cls(mod)  # no "nsubvars": the reference sets none
            else:
                mod["nsubvars"] = 1
                mod["model"] = dust_cls(mod)

        # N stars is only modelled when the SFH is being fit (:71-74)
        self.compute_N_stars = self.physics_model["logA"]["prior"]["name"] != "fixed"
        self.compute_massmult = True
        self.massmultipliers = None

        self.devices = devices
        self.mode = mode
        self.table_precision = table_precision
        self.split_inner = split_inner  # FACTOR mode: dust parameters of the inner table (None = automatic)
        self._engines = {}  # identity of the data dictionaries -> (engine, strong refs)
        # the model structure (which priors are free, their boxes) is fixed at construction
        self._layout_cache = None
        self._box_cache = None
        self._store_targets = None
        self._last_data = None  # (star dict, grid dict, indxs, vals, prior_weight, engine entry)

    # ------------------------------------------------------------------------------------------
    def start_params(self):
        """Names and start values of every variable whose prior is not fixed (:82-108)."""
        names, values = [], []
        for ckey, mod in self.physics_model.items():
            if mod["prior"]["name"] == "fixed":
                continue
            for vname in mod["varnames"]:
                cur = mod[vname]
                if len(np.atleast_1d(cur)) > 1:
                    for ll, cval in enumerate(cur):
                        names.append(f"{ckey}_{vname}{ll + 1}")
                        values.append(cval)
                else:
                    names.append(f"{ckey}_{vname}")
                    values.append(cur)
        return (names, values)

    # -- phi bookkeeping ---------------------------------------------------------------------------
    def _free(self):
        return [p for p in self.params if self.physics_model[p]["prior"]["name"] != "fixed"]

    def _layout(self):
        """[(param, varname, first column in phi, count)] in the reference's order (:133-143)."""
        if self._layout_cache is not None:
            return self._layout_cache
        out, k = [], 0
        for p in self._free():
            mod = self.physics_model[p]
            for vname in mod["varnames"]:
                n = mod["nsubvars"] if mod["nsubvars"] > 1 else 1  # KeyError for M_ini, as upstream
                # a vector variable is always written to ["values"] (:139 hard-codes the key)
                out.append((p, "values" if n > 1 else vname, k, n))
                k += n
        self._layout_cache = (out, k)
        return out, k

    def _store_phi(self, phi):
        """The reference leaves the last evaluated phi in ``physics_model`` (:136-143)."""
        targets = self._store_targets
        if targets is None:  # (container, key) per phi column, resolved once
            targets = []
            for p, vname, k, n in self._layout()[0]:
                mod = self.physics_model[p]
                if n > 1:
                    targets.extend((mod[vname], j) for j in range(n))
                else:
                    targets.append((mod, vname))
            self._store_targets = targets
        for (box, key), v in zip(targets, phi.tolist() if hasattr(phi, "tolist") else phi):
            box[key] = v

    def _spec(self):
        """Device model description + the column permutation from reference phi order."""
        kinds, nodes, perm = {}, {}, []
        layout, ndim = self._layout()
        where = {(p, v): (k, n) for p, v, k, n in layout}
        for p in _cabi.PARAMS:
            mod = self.physics_model[p]
            if mod["prior"]["name"] == "fixed":
                kinds[p] = "fixed"
                continue
            name = mod["name"]
            want = _cabi.KIND_VARNAMES.get(name)
            if want is None or sorted(want) != sorted(mod["varnames"]):
                kinds[p] = "tabulated"  # evaluated by the host callable on the class values
                continue
            if name in ("bins_histo", "bins_interp"):
                if p != "logA" or len(mod["x"]) != len(np.atleast_1d(mod["values"])):
                    kinds[p] = "tabulated"
                    continue
                nodes[p] = list(mod["x"])
            kinds[p] = name
            for v in want:
                k, n = where[(p, v)]
                perm.extend(range(k, k + n))
        return ModelSpec(kinds, nodes), np.asarray(perm, dtype=np.intp), ndim

    # -- device-resident data ------------------------------------------------------------------------
    @staticmethod
    def _data_key(star_lnpgriddata, beast_moddata):
        """Identity of the data a device copy was made from: the two dictionaries, and for every
        array that went to the device its object id, data pointer and shape.
        In-place changes of the VALUES are caught by the library's fingerprint check
        (``mb_watch_host``, see ``_engine_for``)."""
        key = [id(star_lnpgriddata), id(beast_moddata)]
        for d, names in ((star_lnpgriddata, ("indxs", "vals")),
                         (beast_moddata, ("prior_weight", "grid_weight", "completeness"))):
            for name in names:
                a = d[name]
                ai = getattr(a, "__array_interface__", None)
                key.append((id(a), ai["data"][0], ai["shape"]) if ai is not None else id(a))
        return tuple(key)

    def _engine_for(self, star_lnpgriddata, beast_moddata):
        key = self._data_key(star_lnpgriddata, beast_moddata)
        hit = self._engines.get(key)
        if hit is not None:
            return hit[0], hit[1]
        spec, perm, ndim = self._spec()
        mm = None
        if self.compute_N_stars:
            self.massmultipliers = precompute_mass_multipliers(
                beast_moddata, self.physics_model["M_ini"]["model"]
            )
            if self.physics_model["M_ini"]["prior"]["name"] == "fixed":
                self.compute_massmult = False
            mm = self.massmultipliers
        try:
            engine = LikelihoodEngine(spec, star_lnpgriddata, beast_moddata, massmult=mm,
                                      devices=self.devices, mode=self.mode,
                                      table_precision=self.table_precision, split_inner=self.split_inner)
        except _cabi.MegabeastB200Error as exc:
            msg = str(exc)
            if "outside of model x range" in msg:
                raise ValueError("requested x outside of model x range") from None
            if "out of bounds" in msg:
                raise IndexError(msg) from None
            raise
        for old in list(self._engines.values()):  # one resident data set at a time
            old[0].close()
        self._last_data = None
        # let the library notice in-place changes of the arrays the device copy was made from
        watch = [star_lnpgriddata["vals"], star_lnpgriddata["indxs"], beast_moddata["prior_weight"],
                 beast_moddata["grid_weight"], beast_moddata["completeness"]]
        watch += [beast_moddata[p] for p in _cabi.PARAMS if spec.kinds[p] != "fixed"]
        engine.watch(watch)
        if self._identity_perm(perm, ndim):
            perm = None  # device order == reference order (the usual case): no column shuffle
        self._engines = {key: (engine, (spec, perm, ndim), (star_lnpgriddata, beast_moddata))}
        return engine, (spec, perm, ndim)

    def _tables(self, engine, spec, phis):
        """Host-evaluated prior tables for parameters whose model the device does not know."""
        tabs = {}
        for p in spec.tabulated:
            x = engine.class_values(p)
            tab = np.empty((len(phis), x.size))
            for w, phi in enumerate(phis):
                self._store_phi(phi)
                tab[w] = self.physics_model[p]["model"](x)
            tabs[p] = tab
        return tabs

    # ------------------------------------------------------------------------------------------
    def lnlike(self, phi, star_lnpgriddata, beast_moddata):
        """
        log(likelihood) of the ensemble parameters (:110-229).

        phi : (ndim,) -> numpy float64 scalar;  (W, ndim) -> ndarray of W values.
        Raises ``ValueError`` when a star's integrated probability is not finite (:216-217) and
        ``SystemExit`` after printing when the total is exactly 0.0 (:225-227).
        """
        phi = np.asarray(phi, dtype=np.float64)
        single = phi.ndim == 1
        phis = phi[None, :] if single else phi
        engine, (spec, perm, ndim) = self._engine_for(star_lnpgriddata, beast_moddata)
        if phis.shape[1] < ndim:
            raise IndexError(f"index {phis.shape[1]} is out of bounds for axis 0 with size {phis.shape[1]}")
        tabs = self._tables(engine, spec, phis) if spec.tabulated else None
        self._store_phi(phis[-1])
        if perm is None and phis.shape[1] != ndim:
            phis = phis[:, :ndim]
        try:
            vals, status = engine.lnlike(phis if perm is None else phis[:, perm], tabs)
        except _cabi.MegabeastB200Error as exc:
            if "MB_STALE" not in str(exc):
                raise
            self._drop_engines()  # the arrays were changed in place: new device copy, same call again
            return self.lnlike(phi, star_lnpgriddata, beast_moddata)
        if status.any():
            if np.any(status & _cabi.STATUS_PEER_TIMEOUT):
                raise RuntimeError("a peer GPU did not deliver its star shard (timeout)")
            if np.any(status & _cabi.STATUS_NONFINITE_STAR):
                raise ValueError(_NONFINITE_MSG)
            print(0.0)
            raise SystemExit
        return vals[0] if single else vals.copy()

    def _drop_engines(self):
        for old in list(self._engines.values()):
            old[0].close()
        self._engines = {}
        self._last_data = None

    @staticmethod
    def _identity_perm(perm, ndim):
        return perm.size == ndim and np.array_equal(perm, np.arange(ndim))

    def _box(self):
        """(low[ndim], high[ndim]) of the flat priors, indexed exactly as :255-271 index them."""
        if self._box_cache is not None:
            return self._box_cache
        lows, highs = [], []
        for cparam, mod in self.physics_model.items():
            cprior = mod["prior"]
            cname = cprior["name"]
            if cname == "fixed":
                continue
            if cname != "flat":
                raise ValueError(f"{cname} prior not supported")
            for i, _ in enumerate(mod["varnames"]):
                if mod["nsubvars"] > 1:
                    for j in range(len(np.atleast_1d(cprior["minmax"][i]))):
                        lows.append(cprior["minmax"][0][j])
                        highs.append(cprior["minmax"][1][j])
                else:
                    lows.append(cprior["minmax"][i][0])
                    highs.append(cprior["minmax"][i][1])
        self._box_cache = (np.asarray(lows, dtype=np.float64), np.asarray(highs, dtype=np.float64))
        return self._box_cache

    def _lnprob_batch(self, phis, star_lnpgriddata, beast_moddata):
        """Batched lnprob in one C call (box check, compaction of the in-box walkers, likelihood,
        scatter).  Returns None when the general path is needed (tabulated models, several
        devices, permuted variables)."""
        last = self._last_data
        if (last is not None and last[0] is star_lnpgriddata and last[1] is beast_moddata
                and last[2] is star_lnpgriddata["indxs"] and last[3] is star_lnpgriddata["vals"]
                and last[4] is beast_moddata["prior_weight"]):
            hit = last[5]  # same objects as on the previous call (the sampler's steady state)
        else:
            hit = self._engines.get(self._data_key(star_lnpgriddata, beast_moddata))
            if hit is None:  # first call on this data: the general path builds the device data if needed
                return None
            self._last_data = (star_lnpgriddata, beast_moddata, star_lnpgriddata["indxs"],
                               star_lnpgriddata["vals"], beast_moddata["prior_weight"], hit)
        engine, (spec, perm, ndim) = hit[0], hit[1]
        if perm is not None or spec.tabulated or len(engine.handles) != 1 or phis.shape[1] != ndim:
            return None
        lows, highs = self._box()
        try:
            vals, status, last = engine.lnprob_box(phis, lows, highs)
        except _cabi.MegabeastB200Error as exc:
            if "MB_STALE" not in str(exc):
                raise
            self._drop_engines()
            return None  # the general path rebuilds the device copy
        if last >= 0:
            self._store_phi(phis[last])
        if status.tobytes() != engine.zero_status(len(status)):  # (cheaper than status.any() on 64 ints)
            if np.any(status & _cabi.STATUS_PEER_TIMEOUT):
                raise RuntimeError("a peer GPU did not deliver its star shard (timeout)")
            if np.any(status & _cabi.STATUS_NONFINITE_STAR):
                raise ValueError(_NONFINITE_MSG)
            print(0.0)
            raise SystemExit
        return vals.copy()

    def lnprior(self, phi):
        """0 inside the flat-prior box, -inf outside (:231-274).  A 1-D phi gives the
        reference's Python float; a 2-D phi gives one value per row from one array compare."""
        phi = np.asarray(phi, dtype=np.float64)
        lows, highs = self._box()
        n = lows.size
        if phi.shape[-1] < n:
            raise IndexError(f"index {phi.shape[-1]} is out of bounds for axis 0 with size {phi.shape[-1]}")
        outside = np.any((phi[..., :n] < lows) | (phi[..., :n] > highs), axis=-1)
        if phi.ndim == 1:
            return -np.inf if outside else 0.0
        return np.where(outside, -np.inf, 0.0)

    def engine_info(self):
        """Facts about the resident device data (mode, classes, entries, bytes) or None."""
        for engine, _, _ in self._engines.values():
            return engine.info()
        return None


def lnprob(phi, megabeast_model, star_lnpgriddata, beast_moddata):
    """
    log(probability) of the ensemble parameters (:277-304): the prior, and the likelihood only
    for walkers inside the prior box.  A 2-D ``phi`` returns one value per row.
    """
    phi = np.asarray(phi, dtype=np.float64)
    if phi.ndim == 1:
        ln_prior = megabeast_model.lnprior(phi)
        if not np.isfinite(ln_prior):
            return -np.inf
        return ln_prior + megabeast_model.lnlike(phi, star_lnpgriddata, beast_moddata)
    fast = getattr(megabeast_model, "_lnprob_batch", None)
    if fast is not None:
        out = fast(phi, star_lnpgriddata, beast_moddata)
        if out is not None:
            return out
    ln_prior = np.asarray(megabeast_model.lnprior(phi), dtype=np.float64)
    inside = np.isfinite(ln_prior)
    if inside.all():  # the usual case late in a chain: no masking, no copies
        return ln_prior + megabeast_model.lnlike(phi, star_lnpgriddata, beast_moddata)
    out = np.full(phi.shape[0], -np.inf)
    if inside.any():
        out[inside] = ln_prior[inside] + megabeast_model.lnlike(
            phi[inside], star_lnpgriddata, beast_moddata
        )
    return out


def _get_best_fit_params(sampler):
    """Sample with the highest lnprob in an emcee-style sampler (:307-322)."""
    lnp = np.asarray(sampler.lnprobability)
    k, t = np.unravel_index(np.nanargmax(lnp), lnp.shape)
    return sampler.chain[k, t, :]


def fit_ensemble(megabeast_model, star_lnpgriddata, beast_moddata, nsteps=300, nwalkers=None,
                 seed=None, progress=False):
    """
    Run the MegaBEAST on a single set of BEAST results (:325-383): ``5 * ndim`` walkers started
    in a 10 % ball around ``start_params`` (:364-366), ``nsteps`` ensemble-sampler steps, best
    sample returned.  Every half-ensemble proposal is ONE batched GPU evaluation: with emcee
    installed through ``EnsembleSampler(vectorize=True)``, otherwise through the in-repo
    stretch-move sampler (same move, same ``chain``/``lnprobability`` attributes).
    """
    sparams = np.asarray(megabeast_model.start_params()[1], dtype=np.float64)
    ndim = len(sparams)
    nwalkers = 5 * ndim if nwalkers is None else nwalkers
    rng = np.random.default_rng(seed) if seed is not None else np.random
    noise = rng.standard_normal((nwalkers, ndim)) if seed is not None else np.random.randn(nwalkers, ndim)
    pos = sparams * (1.0 + 1e-1 * noise)
    args = (megabeast_model, star_lnpgriddata, beast_moddata)
    try:
        import emcee

        if not hasattr(emcee, "moves"):
            raise ImportError("emcee too old or a stand-in")
        sampler = emcee.EnsembleSampler(nwalkers, ndim, lnprob, args=args, vectorize=True)
        sampler.run_mcmc(pos, nsteps, progress=progress)
        # emcee 3 stores (nsteps, nwalkers[, ndim]); the reference reads the walker-major views
        sampler = _EmceeView(sampler)
    except ImportError:
        sampler = StretchMoveSampler(nwalkers, ndim, lnprob, args=args, seed=seed)
        sampler.run_mcmc(pos, nsteps, progress=progress)
    return _get_best_fit_params(sampler)


class _EmceeView:
    def __init__(self, sampler):
        self.chain = np.swapaxes(sampler.get_chain(), 0, 1)
        self.lnprobability = np.swapaxes(sampler.get_log_prob(), 0, 1)
