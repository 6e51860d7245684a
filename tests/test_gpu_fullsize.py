"""
BASELINE config 3 at FULL size (100,000 stars x 50 points, 1,000,400-row grid, 64 walkers) through
size-independent properties, plus the oracle on a few walkers:

* repeatability: the same call twice is bitwise identical (integer reduction over stars);
* walker independence: a walker's value does not depend on which batch it is evaluated in;
* shard additivity: the STARSUM rows of two half shards add up to the full job's;
* star-order invariance: permuting the stars changes nothing beyond fixed-point rounding;
* the C oracle (reference algorithm, materialised physmod) on 6 of the 64 walkers: 1e-9 relative.
"""
import copy

import numpy as np
import pytest

from megabeast_b200 import MB_Model, lnprob, synth
from megabeast_b200.engine import LikelihoodEngine
from oracle import c_oracle
from oracle import lnprob_port as port

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg3():
    settings, moddata, stars, cfg = synth.make_problem("cfg3")
    model = MB_Model(copy.deepcopy(settings))
    phis = synth.make_walkers(model, cfg["W"], seed=3)
    vals = lnprob(phis, model, stars, moddata)
    return settings, moddata, stars, model, phis, vals


def test_fullsize_repeatable_and_walker_independent(cfg3):
    settings, moddata, stars, model, phis, vals = cfg3
    assert np.all(np.isfinite(vals))
    again = lnprob(phis, model, stars, moddata)
    assert np.array_equal(again, vals)
    # the same walkers in other batch compositions / widths (NW = 1, 2, 4 kernels)
    one = np.array([lnprob(phis[i], model, stars, moddata) for i in (0, 17, 63)])
    assert np.allclose(one, vals[[0, 17, 63]], rtol=1e-13, atol=0)
    half = lnprob(phis[10:30], model, stars, moddata)
    assert np.allclose(half, vals[10:30], rtol=1e-13, atol=0)
    wide = lnprob(np.concatenate([phis, phis[::-1]]), model, stars, moddata)
    assert np.allclose(wide[:64], vals, rtol=1e-13, atol=0) and np.allclose(wide[64:], vals[::-1], rtol=1e-13, atol=0)


def test_fullsize_shard_additivity_and_star_order(cfg3):
    settings, moddata, stars, model, phis, vals = cfg3
    spec, perm, ndim = model._spec()
    S = stars["indxs"].shape[1]
    rows = []
    for lo, hi in ((0, S // 3), (S // 3, S)):
        eng = LikelihoodEngine(spec, stars, moddata, massmult=model.massmultipliers, star_range=(lo, hi),
                               n_stars_total=S)
        rows.append(eng.lnlike_rows(phis).copy())
        eng.close()
    # STARSUM words of the two shards (exact sums) + the (shared) Poisson row
    total = ((rows[0][0] + rows[1][0]) + (rows[0][1] + rows[1][1])) + rows[0][3]
    assert np.array_equal(rows[0][3], rows[1][3])
    assert np.allclose(total, vals, rtol=1e-13, atol=0)
    order = np.random.default_rng(0).permutation(S)
    shuffled = {"indxs": np.ascontiguousarray(stars["indxs"][:, order]),
                "vals": np.ascontiguousarray(stars["vals"][:, order])}
    model2 = MB_Model(copy.deepcopy(settings))
    assert np.allclose(lnprob(phis, model2, shuffled, moddata), vals, rtol=1e-13, atol=0)


def test_fullsize_against_c_oracle(cfg3):
    settings, moddata, stars, model, phis, vals = cfg3
    pm = port.PortModel(copy.deepcopy(settings))
    pick = [0, 1, 13, 31, 32, 63]
    want = c_oracle.lnprob_batch(pm, phis[pick], stars, moddata)
    assert np.max(np.abs(vals[pick] - want) / np.abs(want)) < 1e-9
