"""
megabeast_b200 -- B200-native ensemble likelihood behind megaBEAST's own Python API.

    from megabeast_b200 import MB_Model, lnprob, fit_ensemble     # as megabeast.singlepop_dust_model
    from megabeast_b200.helpers import get_likelihoods, precompute_mass_multipliers

Importing the package does not touch CUDA; the library is loaded (and, if stale, rebuilt with
nvcc) on the first likelihood evaluation.  There is no CPU fallback.
"""
from .singlepop_dust_model import MB_Model, fit_ensemble, lnprob  # noqa: F401

__all__ = ["MB_Model", "lnprob", "fit_ensemble"]
__version__ = "0.1.0"
