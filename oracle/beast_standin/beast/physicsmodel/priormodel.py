"""
Oracle-only stand-in for ``beast.physicsmodel.priormodel`` (absent from the container;
restated in ``oracle/priors.py``).  Call contract per
``megabeast/singlepop_dust_model.py:55-67,146-148``: built from the parameter dict,
keeps a reference to it, reads the current values at call time.
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..", "..")))
from oracle import priors as _p  # noqa: E402

__all__ = ["PriorModel", "PriorDustModel", "PriorAgeModel", "PriorMassModel"]


class PriorModel:
    allowed = None

    def __init__(self, model):
        if self.allowed is not None and model["name"] not in self.allowed:
            raise NotImplementedError(f"{model['name']} is not an allowed model")
        self.model = model

    def __call__(self, x):
        return _p.evaluate_prior_model(self.model, x)


class PriorDustModel(PriorModel):
    allowed = _p.DUST_MODELS


class PriorAgeModel(PriorModel):
    allowed = _p.AGE_MODELS


class PriorMassModel(PriorModel):
    allowed = _p.MASS_MODELS
