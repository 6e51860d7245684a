"""
Oracle-only stand-in for ``emcee`` (absent from the container).  The reference imports it
at module top (``megabeast/singlepop_dust_model.py:4``) and uses it only inside
``fit_ensemble`` (``:368-372``, ``:313-320``): ``EnsembleSampler(nwalkers, ndim, fn,
args=...)``, ``run_mcmc(pos, nsteps, progress=...)``, ``.lnprobability`` (nwalkers, nsteps)
and ``.chain`` (nwalkers, nsteps, ndim).  This is a plain Goodman & Weare stretch move
(a = 2, two half-ensembles per step) so the reference driver can run end to end in
tests; it makes no claim to reproduce emcee's random stream.
"""
import numpy as np

__version__ = "0.0-standin"


class EnsembleSampler:
    def __init__(self, nwalkers, ndim, log_prob_fn, args=(), kwargs=None, a=2.0, seed=None):
        self.nwalkers, self.ndim, self.fn = nwalkers, ndim, log_prob_fn
        self.args, self.kwargs, self.a = tuple(args), dict(kwargs or {}), a
        self.rng = np.random.default_rng(seed)
        self.chain = np.empty((nwalkers, 0, ndim))
        self.lnprobability = np.empty((nwalkers, 0))

    def _lp(self, p):
        return np.array([self.fn(row, *self.args, **self.kwargs) for row in p], dtype=float)

    def run_mcmc(self, pos, nsteps, progress=False):
        p = np.array(pos, dtype=float)
        lp = self._lp(p)
        chain = np.empty((self.nwalkers, nsteps, self.ndim))
        lnp = np.empty((self.nwalkers, nsteps))
        half = self.nwalkers // 2
        sets = [np.arange(0, half), np.arange(half, self.nwalkers)]
        for t in range(nsteps):
            for s in (0, 1):
                me, other = sets[s], sets[1 - s]
                zz = ((self.a - 1.0) * self.rng.random(len(me)) + 1.0) ** 2 / self.a
                partner = p[other[self.rng.integers(len(other), size=len(me))]]
                prop = partner - (partner - p[me]) * zz[:, None]
                lp_new = self._lp(prop)
                lnq = (self.ndim - 1.0) * np.log(zz) + lp_new - lp[me]
                acc = np.log(self.rng.random(len(me))) < lnq
                p[me[acc]] = prop[acc]
                lp[me[acc]] = lp_new[acc]
            chain[:, t, :] = p
            lnp[:, t] = lp
        self.chain, self.lnprobability = chain, lnp
        return p, lp, None

    def get_chain(self):
        return np.swapaxes(self.chain, 0, 1)
