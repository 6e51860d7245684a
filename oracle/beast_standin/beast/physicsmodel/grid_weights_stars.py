"""Oracle-only stand-in for ``beast.physicsmodel.grid_weights_stars`` (see ../../README.md)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..", "..")))
from oracle.priors import bin_boundaries as compute_bin_boundaries  # noqa: E402,F401
