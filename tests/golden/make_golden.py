"""
Freeze golden vectors for the ensemble-likelihood hot path by EXECUTING THE UNMODIFIED
REFERENCE (``/root/reference/megabeast/singlepop_dust_model.py`` ``MB_Model`` / ``lnprob`` and
``helpers.py``) on seeded synthetic inputs.

Runs only in the build container (needs ``/root/reference``); the ``.npz`` files it writes are
committed so the oracle and the CUDA path can be checked anywhere.  ``beast``/``emcee`` are
absent here, so the reference runs with the oracle's stand-ins on ``sys.path``
(``oracle/reference_loader.py``; prior arithmetic restated in ``oracle/priors.py``).

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz

Each file holds: the settings dictionaries (JSON), every ``beast_moddata`` column, the
``star_lnpgriddata`` arrays, the walker positions ``phis`` (W, ndim) and, per walker,
``expect`` (the reference's ``lnprob``) and ``kind``: 0 = value returned, 1 = the reference
raised ``ValueError`` (non-finite star, singlepop_dust_model.py:216-217), 2 = it printed and
called ``exit()`` (total exactly 0.0, :225-227).
"""
import contextlib
import copy
import io
import json
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

from megabeast_b200 import synth  # noqa: E402
from oracle.reference_loader import load_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
COLS = ("Av", "Rv", "f_A", "fA", "M_ini", "logA", "Z", "distance", "prior_weight",
        "grid_weight", "completeness")


def run_reference(spdm, settings, moddata, stars, phis):
    with contextlib.redirect_stdout(io.StringIO()):
        model = spdm.MB_Model(copy.deepcopy(settings))
    expect = np.full(len(phis), np.nan)
    kind = np.zeros(len(phis), dtype=np.int32)
    for i, phi in enumerate(phis):
        try:
            with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
                expect[i] = spdm.lnprob(phi, model, stars, moddata)
        except ValueError:
            kind[i] = 1
        except SystemExit:
            kind[i] = 2
    return expect, kind


def save(name, settings, moddata, stars, phis, expect, kind, note):
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(
        path,
        settings_json=json.dumps(
            {"stellar_model": settings.stellar_model, "fd_model": settings.fd_model}
        ),
        note=note,
        indxs=stars["indxs"],
        vals=stars["vals"],
        phis=np.asarray(phis, dtype=np.float64),
        expect=expect,
        kind=kind,
        **{c: moddata[c] for c in COLS},
    )
    print(f"{name:28s} W={len(phis):3d} kinds={np.bincount(kind, minlength=3)} "
          f"{os.path.getsize(path) / 1024:.0f} KiB  {note}")


def walkers(settings, W, seed, extra=()):
    from oracle.lnprob_port import PortModel

    pm = PortModel(copy.deepcopy(settings))
    start = np.asarray(pm.start_params()[1], dtype=float)
    phis = [start] + list(synth.make_walkers(pm, W - 1, seed=seed)) + list(extra)
    return np.array(phis)


def round1_cases(spdm):

    # 1. docs-example model, uniform and clustered sparse indices
    for mode in ("uniform", "clustered"):
        settings, mod, stars, cfg = synth.make_problem("small", mode=mode)
        phis = walkers(settings, 12, seed=3)
        e, k = run_reference(spdm, settings, mod, stars, phis)
        save(f"docs_small_{mode}", settings, mod, stars, phis, e, k,
             "docs-example model (settings.rst:34-104), Poisson term on")

    # 2. SFH fixed -> no Poisson term (singlepop_dust_model.py:71-74,188-189)
    settings, mod, stars, cfg = synth.make_problem("small")
    settings.stellar_model["logA"]["prior"]["name"] = "fixed"
    phis = walkers(settings, 8, seed=4)
    e, k = run_reference(spdm, settings, mod, stars, phis)
    save("sfh_fixed", settings, mod, stars, phis, e, k, "logA prior fixed: 4 parameters, no N-stars term")

    # 3. empty SFH bins (helpers.py:156 skip), all-zero SFH (-> -inf), out-of-box walkers (-> -inf)
    settings, mod, stars, cfg = synth.make_problem("small")
    base = walkers(settings, 4, seed=5)
    zero_bin = base[1].copy(); zero_bin[2] = 0.0
    two_zero = base[2].copy(); two_zero[0] = 0.0; two_zero[4] = 0.0
    all_zero = base[3].copy(); all_zero[:5] = 0.0
    oob_hi = base[1].copy(); oob_hi[5] = 7.5          # Av mean above 5.0
    oob_lo = base[1].copy(); oob_lo[0] = -1e-9        # negative SFH height
    oob_edge = base[1].copy(); oob_edge[8] = 1.0      # exactly on the upper edge: allowed
    phis = np.array(list(base) + [zero_bin, two_zero, all_zero, oob_hi, oob_lo, oob_edge])
    e, k = run_reference(spdm, settings, mod, stars, phis)
    save("sfh_zero_bins_and_box", settings, mod, stars, phis, e, k,
         "zeroed SFH bins, all-zero SFH, walkers outside / on the prior box")

    # 4. masked entries, fully masked star, exact-zero star
    settings, mod, stars, cfg = synth.make_problem("tiny")
    rng = np.random.default_rng(11)
    vals, indxs = stars["vals"], stars["indxs"]
    vals[:, 3] = -np.inf                       # every entry masked -> sum 0.0 -> skipped
    vals[:, 7] = 0.0                           # finite zeros -> sum 0.0 -> skipped
    vals[rng.random(vals.shape) < 0.15] = -np.inf
    vals[2, 11] = np.inf                       # +inf and nan are dropped by isfinite too
    vals[5, 12] = np.nan
    indxs[~np.isfinite(vals)] = np.iinfo(np.int64).max // 3   # padded rows must never be read
    phis = walkers(settings, 6, seed=6)
    e, k = run_reference(spdm, settings, mod, stars, phis)
    save("masked_and_zero_stars", settings, mod, stars, phis, e, k,
         "15% -inf padding with wild indices, +inf/nan entries, an all-masked and an all-zero star")

    # 5. non-finite star -> ValueError
    settings, mod, stars, cfg = synth.make_problem("tiny")
    mod["completeness"] = mod["completeness"].copy()
    mod["completeness"][stars["indxs"][0, 5]] = np.inf
    phis = walkers(settings, 3, seed=7)
    e, k = run_reference(spdm, settings, mod, stars, phis)
    save("nonfinite_star", settings, mod, stars, phis, e, k,
         "completeness=+inf at a gathered row: ValueError from every in-box walker")

    # 6. total exactly 0.0 -> exit(): SFH fixed and every star sums to 0
    settings, mod, stars, cfg = synth.make_problem("tiny")
    settings.stellar_model["logA"]["prior"]["name"] = "fixed"
    stars["vals"][:] = 0.0
    phis = walkers(settings, 2, seed=8)
    e, k = run_reference(spdm, settings, mod, stars, phis)
    save("total_exactly_zero", settings, mod, stars, phis, e, k,
         "SFH fixed and all stars zero: the reference prints and exit()s")

    # 7. fA free as well (needs an 'fA' column, singlepop_dust_model.py:147), more fA values
    axes = (9, 2, 5, 6, 3, 3)
    settings = synth.docs_example_settings()
    settings.fd_model["fA"]["prior"] = {"name": "flat", "minmax": [[0.05, 1.5], [0.05, 0.5]]}
    settings.fd_model["fA"]["varinit"] = [0.8, 0.25]
    mod = synth.make_grid(axes, seed=21)
    stars = synth.make_star_data(mod, axes, 80, 14, seed=22)
    phis = walkers(settings, 8, seed=9)
    e, k = run_reference(spdm, settings, mod, stars, phis)
    save("fa_free", settings, mod, stars, phis, e, k, "logA, Av, Rv and fA all free: 11 parameters")

    # 8. other model names: two_lognormal A(V), bins_interp SFH with 6 nodes, flat R(V)
    axes = (13, 2, 5, 7, 3, 1)
    settings = synth.docs_example_settings()
    settings.stellar_model["x"] = [6.0, 6.8, 7.6, 8.4, 9.2, 10.0]
    settings.stellar_model["logA"] = {
        "name": "bins_interp", "varnames": ["values"],
        "varinit": [[2e-8, 1e-8, 3e-8, 1e-8, 2e-8, 1e-8]],
        "prior": {"name": "flat", "minmax": [[0.0] * 6, [1e-3] * 6]},
    }
    settings.fd_model["Av"] = {
        "name": "two_lognormal",
        "varnames": ["mean1", "mean2", "sigma1", "sigma2", "N1_to_N2"],
        "varinit": [0.4, 2.0, 0.3, 0.4, 2.5],
        "prior": {"name": "flat",
                  "minmax": [[0.005, 5.0], [0.005, 5.0], [0.05, 1.0], [0.05, 1.0], [0.1, 10.0]]},
    }
    settings.fd_model["Rv"]["prior"]["name"] = "fixed"
    mod = synth.make_grid(axes, seed=31)
    stars = synth.make_star_data(mod, axes, 90, 16, seed=32, mode="clustered")
    phis = walkers(settings, 8, seed=10)
    e, k = run_reference(spdm, settings, mod, stars, phis)
    save("two_lognormal_bins_interp", settings, mod, stars, phis, e, k,
         "two_lognormal A(V), bins_interp SFH (6 nodes), R(V) fixed: 11 parameters")

    # 9. bins_histo with one fewer height than nodes is not reachable through MB_Model
    #    (nsubvars = len(x), :52-54), so it is covered at the priors level only.

    # 10. BASELINE config 1 (1,000 stars x 50 points, 99,630-row grid): a few walkers
    settings, mod, stars, cfg = synth.make_problem("cfg1")
    phis = walkers(settings, 5, seed=3)
    e, k = run_reference(spdm, settings, mod, stars, phis)
    path = os.path.join(OUT, "cfg1_expect.npz")
    np.savez_compressed(path, phis=phis, expect=e, kind=k,
                        note="inputs regenerate from synth.make_problem('cfg1') seeds (1,2,3)")
    print(f"{'cfg1_expect':28s} W={len(phis):3d} (inputs are re-generated, only phis+expect stored)")


def round2_cases(spdm, only=None):
    """Cases added in round 2 (the files of the earlier cases are not rewritten by `--only`)."""
    def want(name):
        return only is None or name in only

    # 11. production dust axes in miniature: 100 A(V) x 9 R(V) = 900 free dust classes (what a real BEAST
    #     grid has; the device splits its factor tables for these), uniform-random indices
    if want("production_dust_axes"):
        axes = (4, 1, 2, 100, 9, 1)
        settings = synth.docs_example_settings()
        mod = synth.make_grid(axes, seed=41)
        stars = synth.make_star_data(mod, axes, 60, 12, seed=42)
        phis = walkers(settings, 8, seed=11)
        e, k = run_reference(spdm, settings, mod, stars, phis)
        save("production_dust_axes", settings, mod, stars, phis, e, k,
             "100 A(V) x 9 R(V) = 900 free dust classes on a 7,200-row grid")

    # 12. a grid that is NOT Cartesian: every row has its own A(V) and R(V) value (nothing to class-code);
    #     the reference does not care (singlepop_dust_model.py:146-148 evaluates the models on the columns)
    if want("non_cartesian_grid"):
        axes = (5, 2, 3, 6, 3, 1)
        settings = synth.docs_example_settings()
        mod = synth.make_grid(axes, seed=51)
        rng = np.random.default_rng(53)
        for key in ("Av", "Rv"):
            mod[key] = mod[key] * (1.0 + 1e-3 * rng.standard_normal(mod[key].size))
        stars = synth.make_star_data(mod, axes, 70, 10, seed=52, mode="clustered")
        phis = walkers(settings, 8, seed=12)
        e, k = run_reference(spdm, settings, mod, stars, phis)
        save("non_cartesian_grid", settings, mod, stars, phis, e, k,
             "every grid row its own A(V), R(V) value: no class coding possible")

    # 13. what numpy indexing accepts and a BEAST file would not contain: the same grid row twice within a
    #     star (the reference simply adds both, :207-213) and negative indices (counted from the end)
    if want("duplicate_and_negative_indices"):
        settings, mod, stars, cfg = synth.make_problem("tiny")
        G = mod["Av"].size
        stars["indxs"][1, :] = stars["indxs"][0, :]
        stars["indxs"][2, ::2] -= G
        phis = walkers(settings, 4, seed=13)
        e, k = run_reference(spdm, settings, mod, stars, phis)
        save("duplicate_and_negative_indices", settings, mod, stars, phis, e, k,
             "every star holds one grid row twice; half the stars index one point from the end")

    # 14. no stars at all: only the N-stars term is left, 0 ln(pred) - pred - lgamma(1) (:182-186)
    if want("no_stars"):
        settings, mod, stars, cfg = synth.make_problem("tiny")
        stars = {"indxs": stars["indxs"][:, :0].copy(), "vals": stars["vals"][:, :0].copy()}
        phis = walkers(settings, 3, seed=15)
        e, k = run_reference(spdm, settings, mod, stars, phis)
        save("no_stars", settings, mod, stars, phis, e, k, "star arrays of shape (K, 0)")

    # 15. completeness 0 everywhere: every star sums to 0 (skipped) and the predicted number is 0 -> -inf
    if want("zero_completeness"):
        settings, mod, stars, cfg = synth.make_problem("tiny")
        mod["completeness"] = np.zeros_like(mod["completeness"])
        phis = walkers(settings, 3, seed=16)
        e, k = run_reference(spdm, settings, mod, stars, phis)
        save("zero_completeness", settings, mod, stars, phis, e, k, "completeness column all zero: ln(0) in the N-stars term")

    # 16. a grid with ONE initial mass: the 100-point mass grid collapses, every IMF integral is 0 and the
    #     mass multipliers are 0/0 -- the reference returns nan for every walker (helpers.py:72-73,92)
    if want("single_mass_grid"):
        axes = (5, 2, 1, 6, 3, 1)
        settings = synth.docs_example_settings()
        mod = synth.make_grid(axes, seed=61)
        stars = synth.make_star_data(mod, axes, 40, 8, seed=62)
        phis = walkers(settings, 3, seed=14)
        e, k = run_reference(spdm, settings, mod, stars, phis)
        save("single_mass_grid", settings, mod, stars, phis, e, k, "one M_ini value in the whole grid: nan")

    # 17. (age, Z) bins without grid weight (helpers.py:156 skips them) -- one bin, and a whole age
    if want("weightless_bins"):
        settings, mod, stars, cfg = synth.make_problem("tiny")
        mod["grid_weight"] = mod["grid_weight"].copy()
        ages, mets = np.unique(mod["logA"]), np.unique(mod["Z"])
        mod["grid_weight"][(mod["logA"] == ages[1]) & (mod["Z"] == mets[0])] = 0.0
        mod["grid_weight"][mod["logA"] == ages[3]] = 0.0
        phis = walkers(settings, 4, seed=17)
        e, k = run_reference(spdm, settings, mod, stars, phis)
        save("weightless_bins", settings, mod, stars, phis, e, k, "grid_weight 0 in one (age, Z) bin and in a whole age")

    # 18. a star whose sum is negative: np.log gives nan (a warning, no exception) and the total is nan (:218-223)
    if want("negative_star"):
        settings, mod, stars, cfg = synth.make_problem("tiny")
        fin = np.isfinite(stars["vals"][:, 4])
        stars["vals"][fin, 4] = -np.abs(stars["vals"][fin, 4])
        phis = walkers(settings, 3, seed=18)
        e, k = run_reference(spdm, settings, mod, stars, phis)
        save("negative_star", settings, mod, stars, phis, e, k, "one star with negative likelihood values: nan, not an exception")


def main():
    """`python tests/golden/make_golden.py` rewrites everything; `--only a,b` writes just those round-2 cases."""
    spdm, _ = load_reference()
    if "--only" in sys.argv:
        round2_cases(spdm, sys.argv[sys.argv.index("--only") + 1].split(","))
        return
    round1_cases(spdm)
    round2_cases(spdm)


if __name__ == "__main__":
    main()
