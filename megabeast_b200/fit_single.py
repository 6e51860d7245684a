"""
Command-line driver with the flow of ``/root/reference/megabeast/fit_single.py:12-76``:

    python -m megabeast_b200.fit_single settings.txt

reads the settings, builds the ``MB_Model``, loads the nine grid columns, the completeness and
the sparse likelihoods of the BEAST run named by ``beast_base``, prints the start lnprob, runs
``fit_ensemble`` and prints the best fit -- every likelihood evaluation on the GPU.

Inputs: the BEAST files ``<base>_seds_trim.grid.hd5``, ``<base>_noisemodel.grid.hd5`` and
``<base>_lnp.hd5`` (needs ``beast`` and ``h5py``, as upstream), or -- for machines without
them -- ``<base>_moddata.npz`` (the ``beast_moddata`` columns) and ``<base>_lnp.npz``
(``vals`` log-posteriors and ``indxs``, both (n_lnps, n_stars)).
"""
import argparse
import os

import numpy as np

from .helpers import get_likelihoods
from .mbsettings import mbsettings
from .singlepop_dust_model import MB_Model, fit_ensemble, lnprob

GRID_COLUMNS = ["Av", "Rv", "f_A", "M_ini", "logA", "Z", "distance", "prior_weight", "grid_weight"]


def read_beast_moddata(base):
    """``beast_moddata`` as ``fit_single.py:37-57`` builds it."""
    npz = base + "_moddata.npz"
    if os.path.isfile(npz):
        with np.load(npz) as f:
            return {k: np.array(f[k]) for k in f.files}
    try:
        import h5py
        from beast.physicsmodel.grid import SEDGrid
    except ImportError as exc:
        raise ImportError(f"reading {base}_seds_trim.grid.hd5 needs beast and h5py ({exc})") from exc
    moddata = {}
    sgrid = SEDGrid(base + "_seds_trim.grid.hd5", backend="disk")
    for cparam in GRID_COLUMNS:
        moddata[cparam] = np.asarray(sgrid.grid[cparam])
    # the maximum completeness over the bands stands in for the per-model completeness (:52-57)
    with h5py.File(base + "_noisemodel.grid.hd5", "r") as obs_hdf:
        moddata["completeness"] = np.max(obs_hdf["completeness"], axis=1)
    return moddata


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("settings_file", help="Name of the file that contains the MegaBEAST settings")
    parser.add_argument("--nsteps", type=int, default=300, help="ensemble sampler steps (upstream: 300)")
    parser.add_argument("--seed", type=int, default=None)
    parser.add_argument("--devices", type=int, nargs="*", default=None, help="CUDA ordinals to shard the stars over")
    args = parser.parse_args(argv)

    mbparams = mbsettings(args.settings_file)
    megabeast_model = MB_Model(mbparams, devices=args.devices, verbose=True)

    beast_moddata = read_beast_moddata(mbparams.beast_base)
    if "fA" not in beast_moddata and "f_A" in beast_moddata:
        beast_moddata["fA"] = beast_moddata["f_A"]  # the model looks columns up by parameter name (:147)
    lnpfile = mbparams.beast_base + "_lnp.hd5"
    if not os.path.isfile(lnpfile):
        lnpfile = mbparams.beast_base + "_lnp.npz"
    star_lnpgriddata = get_likelihoods(lnpfile, beast_moddata)

    print("starting parameters")
    sparams = megabeast_model.start_params()
    print(sparams[0])
    print(sparams[1], lnprob(sparams[1], megabeast_model, star_lnpgriddata, beast_moddata))

    bestparams = fit_ensemble(megabeast_model, star_lnpgriddata, beast_moddata, nsteps=args.nsteps,
                              seed=args.seed)

    np.set_printoptions(precision=3)
    print("final parameters")
    print(bestparams, lnprob(bestparams, megabeast_model, star_lnpgriddata, beast_moddata))
    return bestparams


if __name__ == "__main__":
    main()
