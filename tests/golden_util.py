"""Load the golden vectors frozen from the reference by tests/golden/make_golden.py."""
import glob
import json
import os

import numpy as np

from megabeast_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
COLS = ("Av", "Rv", "f_A", "fA", "M_ini", "logA", "Z", "distance", "prior_weight",
        "grid_weight", "completeness")


def golden_names():
    return sorted(
        os.path.basename(p)[:-4]
        for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
        if not p.endswith("cfg1_expect.npz")
    )


def class_codable(name):
    """False for the golden whose grid gives every row its own A(V) / R(V) value: only the elementwise
    evaluation can take it; the class-coded modes must refuse it."""
    return name != "non_cartesian_grid"


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as f:
        sj = json.loads(str(f["settings_json"]))
        settings = synth.Settings(sj["stellar_model"], sj["fd_model"])
        moddata = {c: np.array(f[c]) for c in COLS}
        stars = {"indxs": np.array(f["indxs"]), "vals": np.array(f["vals"])}
        return settings, moddata, stars, np.array(f["phis"]), np.array(f["expect"]), np.array(f["kind"])


def load_cfg1_expect():
    with np.load(os.path.join(GOLDEN_DIR, "cfg1_expect.npz")) as f:
        return np.array(f["phis"]), np.array(f["expect"]), np.array(f["kind"])


def assert_close(got, expect, rtol=1e-9):
    got, expect = np.asarray(got, float), np.asarray(expect, float)
    fin = np.isfinite(expect)
    assert np.array_equal(got[~fin], expect[~fin], equal_nan=True), (got, expect)
    if fin.any():
        err = np.max(np.abs(got[fin] - expect[fin]) / np.abs(expect[fin]))
        assert err <= rtol, f"relative error {err:.3e} > {rtol}"
