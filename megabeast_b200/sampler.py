"""
Minimal affine-invariant ensemble sampler (Goodman & Weare stretch move, a = 2) used by
``fit_ensemble`` when ``emcee`` is not installed.  It exposes what the reference reads from
an emcee sampler (``/root/reference/megabeast/singlepop_dust_model.py:313-320,368-372``):
``run_mcmc(pos, nsteps, progress=)``, ``.chain`` (nwalkers, nsteps, ndim) and
``.lnprobability`` (nwalkers, nsteps).

Unlike the reference's call pattern (one ``lnprob`` call per walker), the log-probability
function is called with a whole half-ensemble ``(n, ndim)`` at a time, so each proposal round
is one batched GPU evaluation.
"""
import numpy as np

__all__ = ["StretchMoveSampler"]


class StretchMoveSampler:
    def __init__(self, nwalkers, ndim, log_prob_fn, args=(), kwargs=None, a=2.0, seed=None):
        if nwalkers < 2 * ndim:
            raise ValueError("need at least 2 * ndim walkers")  # the reference uses 5 * ndim (:364)
        self.nwalkers, self.ndim = nwalkers, ndim
        self.log_prob_fn, self.args, self.kwargs = log_prob_fn, tuple(args), dict(kwargs or {})
        self.a = float(a)
        self.rng = np.random.default_rng(seed)
        self.chain = np.empty((nwalkers, 0, ndim))
        self.lnprobability = np.empty((nwalkers, 0))
        self.naccepted = np.zeros(nwalkers, dtype=np.int64)
        self.n_calls = 0

    def _lp(self, p):
        self.n_calls += 1
        out = np.asarray(self.log_prob_fn(p, *self.args, **self.kwargs), dtype=np.float64)
        if np.any(np.isnan(out)):
            raise ValueError("Probability function returned NaN")
        return out

    def run_mcmc(self, pos, nsteps, progress=False):
        p = np.array(pos, dtype=np.float64)
        if p.shape != (self.nwalkers, self.ndim):
            raise ValueError("initial positions must be (nwalkers, ndim)")
        lp = self._lp(p)
        chain = np.empty((self.nwalkers, nsteps, self.ndim))
        lnp = np.empty((self.nwalkers, nsteps))
        half = self.nwalkers // 2
        halves = (np.arange(half), np.arange(half, self.nwalkers))
        for t in range(nsteps):
            for s in (0, 1):
                me, other = halves[s], halves[1 - s]
                z = ((self.a - 1.0) * self.rng.random(me.size) + 1.0) ** 2 / self.a
                partner = p[other[self.rng.integers(other.size, size=me.size)]]
                prop = partner - (partner - p[me]) * z[:, None]
                lp_new = self._lp(prop)
                with np.errstate(invalid="ignore"):
                    lnq = (self.ndim - 1.0) * np.log(z) + lp_new - lp[me]
                take = np.log(self.rng.random(me.size)) < lnq
                p[me[take]] = prop[take]
                lp[me[take]] = lp_new[take]
                self.naccepted[me[take]] += 1
            chain[:, t, :] = p
            lnp[:, t] = lp
            if progress and (t + 1) % max(1, nsteps // 10) == 0:
                print(f"step {t + 1}/{nsteps}  max lnprob {lp.max():.6g}")
        self.chain, self.lnprobability = chain, lnp
        return p, lp, None

    def get_chain(self):
        return np.swapaxes(self.chain, 0, 1)

    def get_log_prob(self):
        return self.lnprobability.T

    @property
    def acceptance_fraction(self):
        n = max(1, self.chain.shape[1])
        return self.naccepted / n
