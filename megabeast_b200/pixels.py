"""
Batched independent pixel fits (SURVEY.md section 8(f) "next #3").

The old megaBEAST image driver (``/root/reference/megabeast/old/megabeast_fit.py:97-182``) looped
serially over sky pixels -- one lnp file, one ensemble fit and one best-fit vector per pixel
(:144-161) -- against a single shared BEAST grid (:119-130).  Here all pixels are ONE device
problem: every pixel keeps its own stars, its own walkers and its own Poisson term, and every
likelihood call evaluates all walkers of all pixels in one launch (``mb_lnlike_pixels``).

    pe = PixelEnsemble(mbparams, star_lnpgriddata, beast_moddata, star_offsets)
    pe.lnprob(phis)            # phis (P, W, ndim) -> (P, W)
    best = pe.fit(nsteps=300)  # (P, ndim): the stretch move run on all pixels at once

``star_offsets`` (length P+1) says which columns of the (n_lnps, n_stars) arrays belong to which
pixel.  Pixels are independent: to use several GPUs give each rank a subset of the pixels.
"""
import numpy as np

from . import _cabi
from .engine import LikelihoodEngine
from .helpers import precompute_mass_multipliers
from .singlepop_dust_model import MB_Model

__all__ = ["PixelEnsemble"]


class PixelEnsemble:
    def __init__(self, params, star_lnpgriddata, beast_moddata, star_offsets, device=0, model=None):
        """``model``: an existing ``MB_Model`` to use instead of building one from ``params`` (e.g. one
        whose prior callables were replaced)."""
        self.model = model if model is not None else MB_Model(params, devices=[device])
        self.spec, perm, self.ndim = self.model._spec()
        self.perm = None if self.model._identity_perm(perm, self.ndim) else perm
        self.star_offsets = np.ascontiguousarray(star_offsets, dtype=np.int64)
        self.n_pixels = self.star_offsets.size - 1
        self._park = None
        mm = None
        if self.model.compute_N_stars:
            mm = precompute_mass_multipliers(beast_moddata, self.model.physics_model["M_ini"]["model"])
            self.model.massmultipliers, self.model.compute_massmult = mm, False
        self.engine = LikelihoodEngine(self.spec, star_lnpgriddata, beast_moddata, massmult=mm,
                                       devices=[device], mode="factor", pixel_offsets=self.star_offsets)

    def close(self):
        self.engine.close()

    def lnlike(self, phis):
        """(P, W, ndim) -> (P, W); raises like ``MB_Model.lnlike`` if any walker of any pixel hits a
        non-finite star or an exactly-zero total."""
        phis = np.asarray(phis, dtype=np.float64)
        vals, status = self.engine.lnlike_pixels(self._device_phi(phis), self._tables(phis))
        self._raise(status)
        return vals

    def _device_phi(self, phis):
        if self.perm is None:
            return phis if phis.shape[-1] == self.ndim else phis[..., :self.ndim]
        return phis[..., self.perm]

    def _tables(self, phis):
        """Prior models the device does not know: the host callable evaluated on the parameter's
        class values for every (pixel, walker) -> tables [P*W][n_classes] (singlepop_dust_model.py:
        55-67,146-148: whatever callable the installed BEAST provides)."""
        if not self.spec.tabulated:
            return None
        P, W, _ = phis.shape
        return self.model._tables(self.engine, self.spec, phis.reshape(P * W, -1))

    @staticmethod
    def _raise(status):
        if status.any():
            if np.any(status & _cabi.STATUS_NONFINITE_STAR):
                raise ValueError("invidual integrated star prob is not finite")
            print(0.0)
            raise SystemExit

    def lnprob(self, phis):
        """lnprior + lnlike per walker; -inf outside the prior box (those walkers' likelihood
        values and status flags are discarded, as the reference never evaluates them)."""
        phis = np.asarray(phis, dtype=np.float64)
        P, W, _ = phis.shape
        lows, highs = self.model._box()
        if self.perm is None and not self.spec.tabulated and phis.shape[-1] == self.ndim == lows.size:
            # the box test, the parking of out-of-box walkers and the -inf all happen inside one C call
            if self._park is None:
                park = np.asarray(self.model.start_params()[1], dtype=np.float64)
                inbox = park.size == lows.size and np.all((park >= lows) & (park <= highs))
                self._park = park if inbox else 0.5 * (lows + highs)
            vals, status = self.engine.lnprob_pixels_box(phis, lows, highs, self._park)
            self._raise(status)
            return vals
        lp = self.model.lnprior(phis.reshape(P * W, -1)).reshape(P, W)
        inside = np.isfinite(lp)
        # out-of-box walkers still occupy a column: park them on an in-box position of their pixel
        safe = phis
        if not inside.all():
            safe = phis.copy()
            start = np.asarray(self.model.start_params()[1], dtype=np.float64)
            safe[~inside] = start
        vals, status = self.engine.lnlike_pixels(self._device_phi(safe), self._tables(safe))
        self._raise(np.where(inside, status, 0))
        return np.where(inside, lp + vals, -np.inf)

    def fit(self, nsteps=300, nwalkers=None, seed=None, a=2.0):
        """The reference's ensemble fit (singlepop_dust_model.py:364-383) for every pixel at
        once: 5*ndim walkers in a 10 % ball around the start parameters, stretch moves on the two
        half-ensembles, best sample per pixel.  Returns (best[P, ndim], best_lnprob[P])."""
        rng = np.random.default_rng(seed)
        start = np.asarray(self.model.start_params()[1], dtype=np.float64)
        ndim = start.size
        W = 5 * ndim if nwalkers is None else nwalkers
        P = self.n_pixels
        pos = start * (1.0 + 0.1 * rng.standard_normal((P, W, ndim)))
        lp = self.lnprob(pos)
        best, best_lp = pos[:, 0].copy(), np.full(P, -np.inf)
        half = W // 2
        halves = (np.arange(half), np.arange(half, W))
        for _ in range(nsteps):
            for s in (0, 1):
                me, other = halves[s], halves[1 - s]
                z = ((a - 1.0) * rng.random((P, me.size)) + 1.0) ** 2 / a
                pick = other[rng.integers(other.size, size=(P, me.size))]
                partner = np.take_along_axis(pos, pick[..., None], axis=1)
                prop = partner - (partner - pos[:, me]) * z[..., None]
                lp_new = self.lnprob(prop)
                with np.errstate(invalid="ignore"):
                    lnq = (ndim - 1.0) * np.log(z) + lp_new - lp[:, me]
                take = np.log(rng.random((P, me.size))) < lnq
                cur = pos[:, me]
                cur[take] = prop[take]
                pos[:, me] = cur
                cl = lp[:, me]
                cl[take] = lp_new[take]
                lp[:, me] = cl
            k = np.argmax(lp, axis=1)
            top = lp[np.arange(P), k]
            better = top > best_lp
            best[better] = pos[np.arange(P), k][better]
            best_lp[better] = top[better]
        return best, best_lp
