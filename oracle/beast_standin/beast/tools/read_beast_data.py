"""
Oracle-only stand-in for ``beast.tools.read_beast_data.read_lnp_data``.  h5py is not
installed, so the "file" is an ``.npz`` holding the two arrays the real reader returns:
``vals`` (K, S) float64 log-posteriors padded with -inf and ``indxs`` (K, S) integer grid
rows (``megabeast/helpers.py:31-41`` is the only consumer).
"""
import numpy as np


def read_lnp_data(filename, nstars=None, shift_lnp=True):
    with np.load(filename) as f:
        vals = np.array(f["vals"], dtype=float)
        indxs = np.array(f["indxs"])
    if nstars is not None:
        vals, indxs = vals[:, :nstars], indxs[:, :nstars]
    if shift_lnp:
        vals = vals - np.max(vals, axis=0, keepdims=True)
    return {"vals": vals, "indxs": indxs}
