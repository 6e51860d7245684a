"""
FACTOR mode with the factor tables SPLIT (outer = age classes x some dust parameters, inner = the
other dust parameters): the path production BEAST grids take, whose dust classes (100-200 A(V)
values x 9 R(V)) do not fit shared memory as one product table.  Also: what the device holds,
decoded (``mb_device_entries``), the config-4 table shape, and totals near zero.

Tolerance 1e-9 relative against the oracle everywhere; decoded class codes and folded weights
bit-exact.  Reference: /root/reference/megabeast/singlepop_dust_model.py:131-148 (product over
the parameters), :194-229 (per-star gather, sum, log, sum).
"""
import copy

import numpy as np
import pytest

from golden_util import assert_close, class_codable, golden_names, load_golden
from megabeast_b200 import MB_Model, lnprob, synth
from megabeast_b200.engine import LikelihoodEngine
from oracle import c_oracle
from oracle import lnprob_port as port

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _free_dust(settings):
    return [p for p in ("Av", "Rv", "fA") if settings.fd_model[p]["prior"]["name"] != "fixed"]


def _eval_one_by_one(model, phis, stars, moddata):
    out = np.full(len(phis), np.nan)
    kind = np.zeros(len(phis), np.int32)
    for i, phi in enumerate(phis):
        try:
            out[i] = lnprob(phi, model, stars, moddata)
        except ValueError:
            kind[i] = 1
        except SystemExit:
            kind[i] = 2
    return out, kind


@pytest.mark.parametrize("name", [n for n in golden_names() if class_codable(n)])
def test_goldens_with_every_split(name, capsys):
    """Every golden vector (frozen from the reference) under every possible table split."""
    settings, moddata, stars, phis, expect, kind = load_golden(name)
    free = _free_dust(settings)
    if MB_Model(copy.deepcopy(settings))._spec()[0].tabulated:
        free = [p for p in free]  # tabulated parameters split like any other (K0 path)
    splits = [tuple(p for j, p in enumerate(free) if (m >> j) & 1) for m in range(1, 1 << len(free))]
    assert splits or not free
    ran = 0
    for split in splits:
        model = MB_Model(copy.deepcopy(settings), mode="factor", split_inner=split)
        try:
            got, gkind = _eval_one_by_one(model, phis[:1], stars, moddata)
        except Exception as exc:   # a forced split whose tables are too large for shared memory is refused
            assert "do not fit shared memory" in str(exc) and name == "production_dust_axes", (split, exc)
            continue
        ran += 1
        got, gkind = _eval_one_by_one(model, phis, stars, moddata)
        assert np.array_equal(gkind, kind), split
        ok = kind == 0
        assert_close(got[ok], expect[ok], RTOL)
        info = model.engine_info()
        want_mask = sum(1 << ("Av", "Rv", "fA").index(p) for p in split)
        assert info["split_inner_mask"] == want_mask
        if not kind.any():  # the batch evaluates all walkers in one launch
            assert_close(lnprob(phis, model, stars, moddata), expect, RTOL)
    assert ran >= 1 or not splits


@pytest.mark.parametrize("cfg,mode,W", [("small", "uniform", 7), ("cfg1", "clustered", 45),
                                         ("cfg1", "uniform", 64), ("cfg1", "uniform", 130)])
@pytest.mark.parametrize("split", [("Av",), ("Rv",)])
def test_split_vs_c_oracle(cfg, mode, W, split):
    settings, moddata, stars, _ = synth.make_problem(cfg, mode=mode, seeds=(61, 62, 63))
    pm = port.PortModel(copy.deepcopy(settings))
    phis = synth.make_walkers(pm, W, seed=63)
    expect = c_oracle.lnprob_batch(pm, phis, stars, moddata)
    model = MB_Model(copy.deepcopy(settings), mode="factor", split_inner=split)
    assert_close(lnprob(phis, model, stars, moddata), expect, RTOL)
    info = model.engine_info()
    assert info["n_outer_rows"] > info["n_age_classes"]  # a dust parameter really sits outside
    assert abs(lnprob(phis[3], model, stars, moddata) - expect[3]) <= RTOL * abs(expect[3])


@pytest.mark.parametrize("cfg,ndust", [("dust900_small", 900), ("dust1800_small", 1800)])
def test_production_dust_axes(cfg, ndust):
    """900 / 1,800 dust classes: the product table does not fit shared memory, AUTO must keep the
    tables separate (A(V) inside, age x R(V) outside) and still match the oracle on >= 8 walkers;
    the `table_unique` fallback (class table in global memory) must match as well."""
    settings, moddata, stars, c = synth.make_problem(cfg, mode="uniform", seeds=(71, 72, 73))
    pm = port.PortModel(copy.deepcopy(settings))
    phis = synth.make_walkers(pm, 24, seed=73)
    expect = c_oracle.lnprob_batch(pm, phis, stars, moddata)
    model = MB_Model(copy.deepcopy(settings))
    got = lnprob(phis, model, stars, moddata)
    info = model.engine_info()
    assert info["mode"] == "factor" and info["n_dust_classes"] == ndust
    assert info["split_inner_mask"] == 1 and info["n_inner_rows"] == ndust // 9
    assert info["n_outer_rows"] == info["n_age_classes"] * 9
    assert info["launches_last_eval"] == 1
    assert_close(got, expect, RTOL)
    for W in (1, 33, 70):  # 32-, 64- and 128-walker pass shapes
        ph = synth.make_walkers(pm, W, seed=74 + W)
        assert_close(lnprob(ph, model, stars, moddata), c_oracle.lnprob_batch(pm, ph, stars, moddata), RTOL)
    tu = MB_Model(copy.deepcopy(settings), mode="table_unique")
    assert_close(lnprob(phis, tu, stars, moddata), expect, RTOL)
    assert tu.engine_info()["mode"] == "table_unique"
    # clustered (BEAST-like) indices on the same grid
    _, mod_c, stars_c, _ = synth.make_problem(cfg, mode="clustered", seeds=(71, 75, 73))
    assert_close(lnprob(phis, MB_Model(copy.deepcopy(settings)), stars_c, mod_c),
                 c_oracle.lnprob_batch(pm, phis, stars_c, mod_c), RTOL)


def test_cfg4_table_shape_single_pass_of_128_walkers():
    """Config 4's plan: 200 dust classes x 128 walkers = 210 KB of tables, ONE pass of the 512-thread
    kernel (two-chunk entry ring, reduction slabs over the dead tables, N-stars scratch in what is
    left).  All 128 walkers against the oracle."""
    settings, moddata, stars, cfg = synth.make_problem("cfg4_shape_small", seeds=(81, 82, 83))
    pm = port.PortModel(copy.deepcopy(settings))
    phis = synth.make_walkers(pm, 128, seed=83)
    expect = c_oracle.lnprob_batch(pm, phis, stars, moddata)
    model = MB_Model(copy.deepcopy(settings))
    got = lnprob(phis, model, stars, moddata)
    info = model.engine_info()
    assert info["mode"] == "factor" and info["n_dust_classes"] == 200 and info["split_inner_mask"] == 3
    assert info["walkers_per_pass"] == 128
    assert_close(got, expect, RTOL)
    assert_close(lnprob(phis[:100], model, stars, moddata), expect[:100], RTOL)


def _class_index(values, col, histo):
    if histo:  # interval of a left-constant step function, last node -> last height
        return np.minimum(np.searchsorted(values, col, side="right") - 1, len(values) - 1)
    idx = np.searchsorted(values, col)
    assert np.array_equal(values[idx], col)
    return idx


@pytest.mark.parametrize("mode,split", [("factor", None), ("factor", ("Av",)), ("factor", ("Rv",)),
                                        ("table_unique", None)])
def test_device_entries_decode_bit_exact(mode, split):
    """What the DEVICE holds: the class codes of every entry, decoded, must be the classes of the
    grid rows the reference gathers (indxs[:, s][isfinite(vals[:, s])], :196-198), and the folded
    weights vals*prior_weight*grid_weight*completeness bit for bit."""
    settings, moddata, stars, phis, expect, kind = load_golden("masked_and_zero_stars")
    model = MB_Model(copy.deepcopy(settings))
    spec, perm, ndim = model._spec()
    eng = LikelihoodEngine(spec, stars, moddata, mode=mode, keep_indices=True, merge_duplicates=False,
                           split_inner=split)
    cls, weight, offs = eng.device_entries()
    flat, eoffs = port.gathered_indices(stars)
    assert np.array_equal(offs, eoffs)
    want_cls = np.zeros((flat.size, 4), dtype=np.int64)
    for j, p in enumerate(("logA", "Av", "Rv", "fA")):
        if spec.kinds[p] == "fixed":
            continue
        want_cls[:, j] = _class_index(eng.class_values(p), moddata[p][flat], spec.kinds[p] == "bins_histo")
    vals, indxs = stars["vals"], stars["indxs"]
    S = indxs.shape[1]
    for s in range(S):
        fin = np.isfinite(vals[:, s])
        r = indxs[:, s][fin]
        w = vals[:, s][fin] * moddata["prior_weight"][r] * moddata["grid_weight"][r] * moddata["completeness"][r]
        assert np.array_equal(r, flat[eoffs[s]:eoffs[s + 1]])
        want = sorted(zip(map(tuple, want_cls[eoffs[s]:eoffs[s + 1]]), w))
        got = sorted(zip(map(tuple, cls[offs[s]:offs[s + 1]].astype(np.int64)), weight[offs[s]:offs[s + 1]]))
        assert got == want, s
    eng.close()


def test_totals_near_zero_keep_relative_accuracy():
    """|sum ln L| << 1: a one-word 2^-32 fixed-point sum has an absolute floor of ~1e-10 per segment
    and would miss 1e-9 RELATIVE here; the two-word sum (2^-24 + 2^-68) does not."""
    settings, moddata, stars, _ = synth.make_problem("small", seeds=(91, 92, 93))
    settings.stellar_model["logA"]["prior"]["name"] = "fixed"  # no N-stars term: lnlike = sum ln L only
    pm = port.PortModel(copy.deepcopy(settings))
    S = 7
    sub = {"indxs": np.ascontiguousarray(stars["indxs"][:, :S]), "vals": np.ascontiguousarray(stars["vals"][:, :S])}
    phis = synth.make_walkers(pm, 9, seed=93)
    phis = phis[0] * (1.0 + 1e-4 * np.random.default_rng(1).standard_normal(phis.shape))
    physmod, _ = pm.physmod(phis[0], moddata)
    L = port.star_sums(sub["vals"], sub["indxs"], physmod, moddata["prior_weight"], moddata["grid_weight"],
                       moddata["completeness"])
    delta = 1e-3 * np.array([1, -1, 1, -1, 1, -1, 0.5])
    sub["vals"] = sub["vals"] / L * np.exp(delta)      # every star's L close to 1 for walker 0
    expect = np.array([port.lnprob(p, pm, sub, moddata, loop=False) for p in phis])
    assert np.all(np.abs(expect) < 0.5) and np.all(expect != 0.0)
    for mode in ("factor", "table_unique", "table_grid", "elementwise"):
        model = MB_Model(copy.deepcopy(settings), mode=mode)
        assert_close(lnprob(phis, model, sub, moddata), expect, RTOL)
        assert abs(lnprob(phis[2], model, sub, moddata) - expect[2]) <= RTOL * abs(expect[2])


def _host_lognormal(x, mode, sigma):
    x = np.asarray(x, dtype=float)
    y = np.zeros_like(x)
    g = x > 0
    mu = np.log(mode) + sigma ** 2
    y[g] = (1.0 / np.sqrt(2.0 * np.pi)) / (x[g] * sigma) * np.exp(-0.5 * ((np.log(x[g]) - mu) / sigma) ** 2)
    return y


def test_pixel_batch_with_a_host_tabulated_model():
    """A prior model the device does not know (an opaque host callable, as an installed BEAST with a
    new model name would provide: singlepop_dust_model.py:55-67,146-148) inside a pixel batch: every
    (pixel, walker) table is evaluated on the class values and shipped; same numbers as the analytic
    model."""
    from megabeast_b200.pixels import PixelEnsemble

    settings, moddata, stars, offs = synth.make_pixel_problem(n_pixels=9, mean_stars=25, seeds=(3, 4, 7))
    pe = PixelEnsemble(copy.deepcopy(settings), stars, moddata, offs)
    pm = port.PortModel(copy.deepcopy(settings))
    phis = np.stack([synth.make_walkers(pm, 12, seed=300 + p) for p in range(len(offs) - 1)])
    want = pe.lnprob(phis)
    pe.close()
    s2 = copy.deepcopy(settings)
    model = MB_Model(s2, devices=[0])
    av = model.physics_model["Av"]
    av["name"] = "some_future_beast_model"
    av["model"] = lambda x, av=av: _host_lognormal(x, av["mean"], av["sigma"])
    pe2 = PixelEnsemble(None, stars, moddata, offs, model=model)
    assert pe2.spec.tabulated == ["Av"]
    got = pe2.lnprob(phis)
    assert_close(got, want, 1e-12)
    pe2.close()


def test_host_tabulated_model_on_a_non_cartesian_grid():
    """The last combination that used to be refused: a prior model the device does not know (host callable)
    on a grid whose dust columns cannot be class-coded (ELEMENTWISE mode).  Every grid row is then its own
    class: the callable is evaluated on the whole column per walker, as the reference does
    (singlepop_dust_model.py:146-148), and the [W][G] table rides along."""
    settings, moddata, stars, _ = synth.make_problem("small", seeds=(21, 22, 23))
    rng = np.random.default_rng(5)
    moddata = dict(moddata)
    for key in ("Av", "Rv"):
        moddata[key] = moddata[key] * (1.0 + 1e-3 * rng.standard_normal(moddata[key].size))
    pm = port.PortModel(copy.deepcopy(settings))
    for W in (1, 9, 70, 140):     # one-walker, one pass, 128-wide pass, two passes
        phis = synth.make_walkers(pm, W, seed=24 + W)
        want = c_oracle.lnprob_batch(pm, phis, stars, moddata)
        model = MB_Model(copy.deepcopy(settings))
        av = model.physics_model["Av"]
        av["name"] = "some_future_beast_model"
        av["model"] = lambda x, av=av: _host_lognormal(x, av["mean"], av["sigma"])
        got = lnprob(phis, model, stars, moddata)
        info = model.engine_info()
        assert info["mode"] == "elementwise" and info["n_classes"][1] == moddata["Av"].size   # (logA, Av, Rv, fA)
        assert_close(got, want, RTOL)


def test_pixel_batch_prior_box_in_the_c_call():
    """``PixelEnsemble.lnprob`` applies the flat prior box inside one C call (``mb_lnprob_pixels_box``):
    same values, bit for bit, as lnprior + lnlike composed on the host (singlepop_dust_model.py:277-304);
    walkers outside the box are -inf and whatever their parked column computes is discarded -- also when
    a whole pixel is outside, and at the box edges (<= and >= are inside, :255-271)."""
    from megabeast_b200.pixels import PixelEnsemble

    settings, moddata, stars, offs = synth.make_pixel_problem(n_pixels=11, mean_stars=30, seeds=(3, 4, 8))
    pe = PixelEnsemble(copy.deepcopy(settings), stars, moddata, offs)
    pm = port.PortModel(copy.deepcopy(settings))
    P, W = len(offs) - 1, 20
    phis = np.stack([synth.make_walkers(pm, W, seed=400 + p) for p in range(P)])
    lows, highs = pe.model._box()
    phis[1, 3, 0] = lows[0] - 1e-9
    phis[1, 4, 2] = highs[2] + 1.0
    phis[4, :, 7] = -50.0                      # every walker of pixel 4
    phis[8, 0, 6] = highs[6]                   # on the edge: inside
    phis[8, 1, 6] = lows[6]
    got = pe.lnprob(phis)
    outside = np.zeros((P, W), bool)
    outside[1, 3] = outside[1, 4] = True
    outside[4, :] = True
    assert np.all(got[outside] == -np.inf) and np.all(np.isfinite(got[~outside]))
    # the host composition of the same calls
    lp = pe.model.lnprior(phis.reshape(P * W, -1)).reshape(P, W)
    assert np.array_equal(np.isfinite(lp), ~outside)
    safe = phis.copy()
    safe[outside] = np.asarray(pe.model.start_params()[1], float)
    vals, status = pe.engine.lnlike_pixels(safe)
    assert np.array_equal(got, np.where(outside, -np.inf, lp + vals))
    pe.close()


def test_in_place_change_of_the_data_is_noticed():
    """The data dictionaries arrive on every call (singlepop_dust_model.py:110) but live on the device:
    an array changed IN PLACE must not be evaluated from the stale copy."""
    settings, moddata, stars, _ = synth.make_problem("small", seeds=(95, 96, 97))
    stars = {"indxs": stars["indxs"].copy(), "vals": stars["vals"].copy()}
    pm = port.PortModel(copy.deepcopy(settings))
    phis = synth.make_walkers(pm, 5, seed=97)
    model = MB_Model(copy.deepcopy(settings))
    assert_close(lnprob(phis, model, stars, moddata), c_oracle.lnprob_batch(pm, phis, stars, moddata), RTOL)
    stars["vals"] *= 3.0                                   # same objects, same pointers, new numbers
    assert_close(lnprob(phis, model, stars, moddata), c_oracle.lnprob_batch(pm, phis, stars, moddata), RTOL)
    moddata["completeness"] *= 0.5
    assert_close(lnprob(phis[0], model, stars, moddata), c_oracle.lnprob_batch(pm, phis[:1], stars, moddata)[0], RTOL)
    moddata["completeness"] *= 2.0
    # a NEW array object under the same key is noticed as well (key: ids, data pointers, shapes)
    stars["vals"] = stars["vals"] / 3.0
    assert_close(lnprob(phis, model, stars, moddata), c_oracle.lnprob_batch(pm, phis, stars, moddata), RTOL)


@pytest.mark.parametrize("cfg,mode,split,merge", [
    ("small", "factor", None, True), ("small", "factor", None, False), ("cfg1", "factor", ("Av",), True),
    ("cfg1", "table_unique", None, True), ("cfg2", "factor", None, True), ("cfg1", "table_grid", None, True),
    ("cfg1", "elementwise", None, False)])
def test_device_build_equals_host_build(cfg, mode, split, merge, monkeypatch):
    """The entry stream is built on the device (mask, fold, class look-up, per-star bitonic sort, merge,
    flags: mb_build_entries); the host loop it replaces is kept behind MB_HOST_BUILD=1.  Both must leave
    bit-identical data on the device, and the same likelihood."""
    settings, moddata, stars, _ = synth.make_problem(cfg, mode="clustered", seeds=(101, 102, 103))
    model = MB_Model(copy.deepcopy(settings))
    spec, perm, ndim = model._spec()
    mm = port.mass_multipliers(moddata, {"name": "kroupa"})
    phis = synth.make_walkers(port.PortModel(copy.deepcopy(settings)), 9, seed=103)[:, perm]
    out = {}
    for which in ("device", "host"):
        if which == "host":
            monkeypatch.setenv("MB_HOST_BUILD", "1")
        else:
            monkeypatch.delenv("MB_HOST_BUILD", raising=False)
        eng = LikelihoodEngine(spec, stars, moddata, massmult=mm, mode=mode, keep_indices=True,
                               merge_duplicates=merge, split_inner=split)
        info = eng.info()
        if mode in ("factor", "table_unique"):
            data = eng.device_entries()
        else:
            data = eng.gathered_indices()
        out[which] = (data, {k: info[k] for k in ("n_entries", "n_entries_raw", "n_runs", "code_bytes")},
                      eng.lnlike(phis)[0].copy())
        eng.close()
    assert out["device"][1] == out["host"][1]
    for a, b in zip(out["device"][0], out["host"][0]):
        assert np.array_equal(a, b)
    assert np.array_equal(out["device"][2], out["host"][2])


def test_device_build_reports_bad_indices_like_numpy():
    settings, moddata, stars, _ = synth.make_problem("small", seeds=(105, 106, 107))
    stars = {"indxs": stars["indxs"].copy(), "vals": stars["vals"].copy()}
    G = len(moddata["Av"])
    stars["indxs"][3, 17] = G + 5           # a finite point with an index past the grid
    model = MB_Model(copy.deepcopy(settings))
    phis = synth.make_walkers(port.PortModel(copy.deepcopy(settings)), 3, seed=107)
    with pytest.raises(IndexError, match=f"index {G + 5} is out of bounds for axis 0 with size {G}"):
        lnprob(phis, model, stars, moddata)
    stars["indxs"][3, 17] = -1              # numpy's negative indexing: the last row
    want = c_oracle.lnprob_batch(port.PortModel(copy.deepcopy(settings)), phis, stars, moddata)
    assert_close(lnprob(phis, MB_Model(copy.deepcopy(settings)), stars, moddata), want, RTOL)


def test_elementwise_one_exp_path_and_its_underflow_fallback():
    """K1e evaluates the lognormal dust factors of a row with ONE exp (summed exponents) and falls back to
    factor-by-factor evaluation where the sum is below -700.  Walkers with the narrowest allowed A(V)
    sigma drive most rows into that regime: their factors underflow to exactly 0 and the stars whose
    points all vanish are skipped like in the reference (:218-221).

    (Walkers are chosen so that no star's integrated probability is SUBNORMAL: there the reference's own
    value is a handful of units of 5e-324 -- rounding noise of its multiplication order, ln L = -744 +- 0.7
    per star -- and no implementation that folds the static weights can reproduce it)"""
    settings, moddata, stars, _ = synth.make_problem("cfg1", mode="clustered", seeds=(111, 112, 113))
    rng = np.random.default_rng(9)
    moddata = dict(moddata)
    for key in ("Av", "Rv"):  # a non-Cartesian grid: ELEMENTWISE is the only mode that takes it
        moddata[key] = moddata[key] * (1.0 + 1e-3 * rng.standard_normal(moddata[key].size))
    pm = port.PortModel(copy.deepcopy(settings))
    base = synth.make_walkers(pm, 40, seed=113)
    phis = []
    for i, sig in ((3, 0.05), (10, 0.05), (11, 0.05), (5, 0.055), (15, 0.055), (0, 0.08), (7, 0.08), (1, 0.1), (2, 0.25)):
        p = base[i].copy()
        p[6] = sig
        phis.append(p)
    phis = np.array(phis)
    n_zero = 0
    for p in phis:
        assert pm.lnprior(p) == 0.0
        physmod, _ = pm.physmod(p, moddata)
        L = port.star_sums(stars["vals"], stars["indxs"], physmod, moddata["prior_weight"], moddata["grid_weight"],
                           moddata["completeness"])
        assert not np.any((L > 0) & (L < 1e-290))   # the precondition stated above
        n_zero += int((L == 0).sum())
    assert n_zero > 1000                            # many vanished stars are part of the test
    expect = c_oracle.lnprob_batch(pm, phis, stars, moddata)
    model = MB_Model(copy.deepcopy(settings))
    got = lnprob(phis, model, stars, moddata)
    assert model.engine_info()["mode"] == "elementwise"
    assert_close(got, expect, RTOL)
