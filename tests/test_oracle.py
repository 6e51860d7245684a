"""
The oracle (oracle/lnprob_port.py numpy port, oracle/mb_oracle.c plain-C restatement) against
(a) the golden vectors frozen from the reference, and (b) the live reference when
/root/reference is present (build container only).
"""
import contextlib
import copy
import io

import numpy as np
import pytest

from golden_util import assert_close, golden_names, load_cfg1_expect, load_golden
from megabeast_b200 import synth
from oracle import c_oracle, priors
from oracle import lnprob_port as port
from fuzz_util import random_problem as _random_problem
from oracle.reference_loader import load_reference, reference_available


def _port_eval(settings, moddata, stars, phis, loop):
    pm = port.PortModel(copy.deepcopy(settings))
    out = np.full(len(phis), np.nan)
    kind = np.zeros(len(phis), np.int32)
    for i, phi in enumerate(phis):
        try:
            with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
                out[i] = port.lnprob(phi, pm, stars, moddata, loop=loop)
        except ValueError:
            kind[i] = 1
        except SystemExit:
            kind[i] = 2
    return out, kind


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("loop", [True, False])
def test_port_matches_golden(name, loop):
    settings, moddata, stars, phis, expect, kind = load_golden(name)
    got, gkind = _port_eval(settings, moddata, stars, phis, loop)
    assert np.array_equal(gkind, kind)
    ok = kind == 0
    # the loop form repeats the reference's operation order: bit-identical
    assert_close(got[ok], expect[ok], rtol=0.0 if loop else 1e-13)


@pytest.mark.parametrize("name", golden_names())
def test_c_oracle_matches_golden(name):
    settings, moddata, stars, phis, expect, kind = load_golden(name)
    pm = port.PortModel(copy.deepcopy(settings))
    inside = np.array([np.isfinite(pm.lnprior(p)) for p in phis])
    got = np.full(len(phis), -np.inf)
    v, st = c_oracle.lnlike_batch(pm, phis[inside], stars, moddata)
    got[inside] = v
    gkind = np.zeros(len(phis), np.int32)
    gkind[inside] = np.where(st & 1, 1, np.where(st & 2, 2, 0))
    assert np.array_equal(gkind, kind)
    ok = kind == 0
    assert_close(got[ok], expect[ok], rtol=1e-12)


def test_cfg1_expect():
    phis, expect, kind = load_cfg1_expect()
    settings, moddata, stars, cfg = synth.make_problem("cfg1")
    pm = port.PortModel(settings)
    got = c_oracle.lnprob_batch(pm, phis, stars, moddata)
    assert_close(got, expect, rtol=1e-12)
    got1 = port.lnprob(phis[1], pm, stars, moddata, loop=True)
    assert got1 == expect[1]


def test_gathered_indices_layout():
    settings, moddata, stars, phis, expect, kind = load_golden("masked_and_zero_stars")
    flat, offs = port.gathered_indices(stars)
    K, S = stars["indxs"].shape
    assert offs[0] == 0 and offs[-1] == len(flat) == np.isfinite(stars["vals"]).sum()
    for s in (0, 3, 7, S - 1):
        keep = np.isfinite(stars["vals"][:, s])
        assert np.array_equal(flat[offs[s]:offs[s + 1]], stars["indxs"][:, s][keep])
    assert offs[4] - offs[3] == 0  # the fully masked star gathers nothing


def test_priors_restatement_anchors():
    # lognormal peaks at `mean` (old/ensemble_model.py:13-47: mu = ln(max_pos) + sigma^2)
    x = np.linspace(0.01, 5, 20001)
    y = priors.lognormal_density(x, 0.7, 0.3)
    assert abs(x[np.argmax(y)] - 0.7) < 1e-3
    assert priors.lognormal_density(np.array([0.0, -1.0]), 0.7, 0.3).tolist() == [0.0, 0.0]
    # bins_histo: left-constant steps (checked against scipy interp1d kind="zero")
    from scipy.interpolate import interp1d

    nodes, h = [6.0, 7.0, 8.0, 9.0, 10.0], [1.0, 2.0, 3.0, 4.0, 5.0]
    xs = np.array([6.0, 6.5, 7.0, 9.99, 10.0])
    assert np.array_equal(priors.step_histogram(xs, nodes, h), interp1d(nodes, h, kind="zero")(xs))
    assert priors.step_histogram(np.array([9.5]), nodes, h[:-1] )[0] == 4.0
    assert priors.step_histogram(np.array([10.0]), nodes, h[:-1])[0] == 0.0
    with pytest.raises(ValueError):
        priors.step_histogram(np.array([10.1]), nodes, h)
    # bin boundaries: len+1 edges, midpoints inside (helpers.py:78-79,95)
    e = priors.bin_boundaries(np.array([6.0, 6.1, 6.3]))
    assert np.allclose(e, [5.95, 6.05, 6.2, 6.4])
    # kroupa: continuous at both breaks
    for b in (0.08, 0.5):
        assert np.isclose(priors.kroupa_imf(b * (1 - 1e-12)), priors.kroupa_imf(b), rtol=1e-9)


@pytest.mark.skipif(not reference_available(), reason="/root/reference not on this machine")
@pytest.mark.parametrize("cfg", ["tiny", "small"])
@pytest.mark.parametrize("mode", ["uniform", "clustered"])
def test_port_matches_live_reference(cfg, mode):
    spdm, helpers = load_reference()
    settings, moddata, stars, _ = synth.make_problem(cfg, mode=mode, seeds=(5, 6, 7))
    with contextlib.redirect_stdout(io.StringIO()):
        ref = spdm.MB_Model(copy.deepcopy(settings))
    pm = port.PortModel(copy.deepcopy(settings))
    assert ref.start_params() == tuple(pm.start_params())
    phis = synth.make_walkers(pm, 5, seed=9)
    for phi in phis:
        with contextlib.redirect_stdout(io.StringIO()):
            r = spdm.lnprob(phi, ref, stars, moddata)
        assert port.lnprob(phi, pm, stars, moddata, loop=True) == r
        assert abs(port.lnprob(phi, pm, stars, moddata, loop=False) - r) <= 1e-13 * abs(r)
    # helpers pieces on their own
    mm_ref = helpers.precompute_mass_multipliers(moddata, ref.physics_model["M_ini"]["model"])
    mm = port.mass_multipliers(moddata, pm.physics_model["M_ini"])
    assert np.array_equal(mm_ref["massmult"], mm["massmult"]) and mm_ref["avemass"] == mm["avemass"]
    # get_likelihoods transform (helpers.py:33-41)
    lnp = synth.make_lnp(moddata, synth.CONFIGS[cfg]["axes"], 40, 10, seed=3)
    a = port.likelihoods_from_lnp(lnp, moddata)
    fin = np.isfinite(lnp["vals"])
    assert np.array_equal(a["vals"][fin], np.exp(lnp["vals"][fin]) / moddata["prior_weight"][lnp["indxs"][fin]])


@pytest.mark.skipif(not reference_available(), reason="/root/reference not on this machine")
def test_oracles_against_live_reference_on_random_problems():
    """Differential check of both oracles against the UNMODIFIED reference executed here, on 40 random small
    problems (model variants, masking, zero rows, out-of-box walkers): values to 1e-11, exceptions alike."""
    spdm, _ = load_reference()
    rng = np.random.default_rng(2024)
    compared = 0
    for _ in range(40):
        try:
            settings, moddata, stars, phis = _random_problem(rng)
        except ValueError:   # (fewer distinct grid rows than sparse points asked for)
            continue
        with contextlib.redirect_stdout(io.StringIO()):
            ref = spdm.MB_Model(copy.deepcopy(settings))
        expect, kind = np.full(len(phis), np.nan), np.zeros(len(phis), int)
        for i, phi in enumerate(phis):
            try:
                with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
                    expect[i] = spdm.lnprob(phi, ref, stars, moddata)
            except ValueError:
                kind[i] = 1
            except SystemExit:
                kind[i] = 2
        pm = port.PortModel(copy.deepcopy(settings))
        pkind = []
        for i, phi in enumerate(phis):
            try:
                with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
                    v = port.lnprob(phi, pm, stars, moddata, loop=True)
                pkind.append(0)
                assert v == expect[i] or (np.isnan(v) and np.isnan(expect[i]))
            except ValueError:
                pkind.append(1)
            except SystemExit:
                pkind.append(2)
        assert pkind == list(kind)
        ok = kind == 0
        if ok.any():
            got = c_oracle.lnprob_batch(port.PortModel(copy.deepcopy(settings)), phis[ok], stars, moddata)
            assert_close(got, expect[ok], 1e-11)
        compared += 1
    assert compared >= 30
