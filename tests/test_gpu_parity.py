"""
Parity of the CUDA path (through MB_Model / the C ABI) against the golden vectors frozen from
the reference and against the oracle on seeded synthetic inputs.  Tolerance: 1e-9 relative on
every walker's lnprob (BASELINE.json north_star), fp64 throughout; -inf and the exception
kinds must match exactly; gathered indices bit-exact.
"""
import copy

import numpy as np
import pytest

from golden_util import assert_close, class_codable, golden_names, load_cfg1_expect, load_golden
from megabeast_b200 import MB_Model, lnprob, synth
from megabeast_b200.engine import LikelihoodEngine
from oracle import c_oracle
from oracle import lnprob_port as port

pytestmark = pytest.mark.gpu
MODES = ["factor", "table_unique", "table_grid", "elementwise"]
RTOL = 1e-9


def _eval(model, phis, stars, moddata, batch):
    out = np.full(len(phis), np.nan)
    kind = np.zeros(len(phis), np.int32)
    if batch:
        try:
            return lnprob(phis, model, stars, moddata), kind
        except ValueError:
            return out, np.full(len(phis), 1, np.int32)
        except SystemExit:
            return out, np.full(len(phis), 2, np.int32)
    for i, phi in enumerate(phis):
        try:
            out[i] = lnprob(phi, model, stars, moddata)
        except ValueError:
            kind[i] = 1
        except SystemExit:
            kind[i] = 2
    return out, kind


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("mode", MODES)
def test_golden_one_walker_at_a_time(name, mode, capsys):
    settings, moddata, stars, phis, expect, kind = load_golden(name)
    model = MB_Model(copy.deepcopy(settings), mode=mode)
    if mode != "elementwise" and not class_codable(name):
        with pytest.raises(Exception, match="distinct values|classes|factor-table rows"):
            lnprob(phis[0], model, stars, moddata)
        return
    got, gkind = _eval(model, phis, stars, moddata, batch=False)
    assert np.array_equal(gkind, kind)
    ok = kind == 0
    assert_close(got[ok], expect[ok], RTOL)
    assert model.engine_info()["mode"] == mode


@pytest.mark.parametrize("name", golden_names())
def test_golden_batched(name, capsys):
    settings, moddata, stars, phis, expect, kind = load_golden(name)
    model = MB_Model(copy.deepcopy(settings))
    got, gkind = _eval(model, phis, stars, moddata, batch=True)
    if kind.any():  # a batch raises as soon as any walker would
        assert gkind[0] == kind.max()
    else:
        assert_close(got, expect, RTOL)


def test_return_types_and_side_effects():
    settings, moddata, stars, phis, expect, kind = load_golden("docs_small_uniform")
    model = MB_Model(copy.deepcopy(settings))
    v = model.lnlike(phis[1], stars, moddata)
    assert isinstance(v, np.float64)
    # the reference leaves the last phi in physics_model (:136-143) -> start_params echoes it
    assert np.array_equal(np.asarray(model.start_params()[1]), phis[1])
    assert model.compute_massmult is False and model.massmultipliers is not None
    assert lnprob(phis[1] * 0 - 1.0, model, stars, moddata) == -np.inf


@pytest.mark.parametrize("mode", ["factor", "table_grid", "elementwise"])
def test_gathered_indices_bit_exact(mode):
    settings, moddata, stars, phis, expect, kind = load_golden("masked_and_zero_stars")
    model = MB_Model(copy.deepcopy(settings))
    spec, perm, ndim = model._spec()
    eng = LikelihoodEngine(spec, stars, moddata, mode=mode, keep_indices=True, merge_duplicates=False)
    rows, offs = eng.gathered_indices()
    flat, eoffs = port.gathered_indices(stars)
    assert np.array_equal(offs, eoffs)
    for s in range(len(offs) - 1):
        assert np.array_equal(np.sort(rows[offs[s]:offs[s + 1]]), np.sort(flat[eoffs[s]:eoffs[s + 1]]))
    eng.close()


@pytest.mark.parametrize("cfg,mode,W", [("small", "uniform", 7), ("small", "clustered", 33),
                                         ("cfg1", "uniform", 45), ("cfg1", "clustered", 64),
                                         ("cfg1", "uniform", 130), ("cfg2", "uniform", 32)])
@pytest.mark.parametrize("engine_mode", MODES)
def test_synthetic_vs_c_oracle(cfg, mode, W, engine_mode):
    settings, moddata, stars, _ = synth.make_problem(cfg, mode=mode, seeds=(11, 12, 13))
    pm = port.PortModel(copy.deepcopy(settings))
    phis = synth.make_walkers(pm, W, seed=13)
    expect = c_oracle.lnprob_batch(pm, phis, stars, moddata)
    model = MB_Model(copy.deepcopy(settings), mode=engine_mode)
    got = lnprob(phis, model, stars, moddata)
    assert_close(got, expect, RTOL)
    # merged and unmerged entry streams agree; single-walker calls agree with the batch
    spec, perm, ndim = model._spec()
    eng = LikelihoodEngine(spec, stars, moddata, massmult=model.massmultipliers, mode=engine_mode,
                           merge_duplicates=False)
    v, st = eng.lnlike(phis[:, perm])
    assert not st.any()
    assert_close(v, expect, RTOL)
    eng.close()
    assert abs(lnprob(phis[3], model, stars, moddata) - expect[3]) <= RTOL * abs(expect[3])


def _jittered(moddata, rng, frac=1e-3):
    """A NON-Cartesian grid: every row gets its own Av and Rv value (no small set of classes)."""
    out = dict(moddata)
    for key in ("Av", "Rv"):
        out[key] = moddata[key] * (1.0 + frac * rng.standard_normal(moddata[key].size))
    return out


@pytest.mark.parametrize("cfg,W", [("small", 16), ("cfg1", 45), ("cfg1", 130)])
def test_non_cartesian_grid_auto_elementwise(cfg, W):
    """Columns with ~G distinct values cannot be class-coded: AUTO must pick the elementwise
    evaluation over the grid's physics columns (singlepop_dust_model.py:131-148 literally) and
    still match the oracle; the class-coded modes must refuse such a grid."""
    settings, moddata, stars, _ = synth.make_problem(cfg, seeds=(21, 22, 23))
    moddata = _jittered(moddata, np.random.default_rng(5))
    assert np.unique(moddata["Av"]).size > 4096
    pm = port.PortModel(copy.deepcopy(settings))
    phis = synth.make_walkers(pm, W, seed=23)
    expect = c_oracle.lnprob_batch(pm, phis, stars, moddata)
    model = MB_Model(copy.deepcopy(settings))
    got = lnprob(phis, model, stars, moddata)
    assert model.engine_info()["mode"] == "elementwise"
    assert_close(got, expect, RTOL)
    assert abs(lnprob(phis[2], model, stars, moddata) - expect[2]) <= RTOL * abs(expect[2])
    with pytest.raises(Exception, match="distinct values"):
        lnprob(phis, MB_Model(copy.deepcopy(settings), mode="factor"), stars, moddata)


def test_elementwise_matches_class_coded_on_cartesian_grid():
    settings, moddata, stars, _ = synth.make_problem("cfg1", mode="clustered", seeds=(41, 42, 43))
    phis = synth.make_walkers(MB_Model(copy.deepcopy(settings)), 64, seed=43)
    a = lnprob(phis, MB_Model(copy.deepcopy(settings), mode="factor"), stars, moddata)
    b = lnprob(phis, MB_Model(copy.deepcopy(settings), mode="elementwise"), stars, moddata)
    assert_close(b, a, RTOL)


def test_many_walkers_run_as_several_passes():
    """300 walkers = passes of 128 + 128 + 44: every pass builds its own tables inside K2."""
    settings, moddata, stars, _ = synth.make_problem("small", seeds=(51, 52, 53))
    pm = port.PortModel(copy.deepcopy(settings))
    phis = synth.make_walkers(pm, 300, seed=53)
    expect = c_oracle.lnprob_batch(pm, phis, stars, moddata)
    for mode in ("factor", "table_unique", "elementwise"):
        assert_close(lnprob(phis, MB_Model(copy.deepcopy(settings), mode=mode), stars, moddata), expect, RTOL)


def test_phi_too_large_for_the_kernel_parameters():
    """11 free variables x 128 walkers do not fit the 1,152 doubles K2 carries in its parameters:
    the host path then keeps the separate K0 launch; results are the same as for a small batch."""
    settings, moddata, stars, phis0, expect0, kind = load_golden("two_lognormal_bins_interp")
    pm = port.PortModel(copy.deepcopy(settings))
    phis = synth.make_walkers(pm, 128, seed=7)
    expect = c_oracle.lnprob_batch(pm, phis, stars, moddata)
    model = MB_Model(copy.deepcopy(settings), mode="factor")
    assert_close(lnprob(phis, model, stars, moddata), expect, RTOL)
    assert model.engine_info()["launches_last_eval"] == 2          # K0 + K2
    assert_close(lnprob(phis[:64], model, stars, moddata), expect[:64], RTOL)
    assert model.engine_info()["launches_last_eval"] == 1          # tables built inside K2


def test_cfg1_golden_expect():
    phis, expect, kind = load_cfg1_expect()
    settings, moddata, stars, _ = synth.make_problem("cfg1")
    model = MB_Model(settings)
    assert_close(lnprob(phis, model, stars, moddata), expect, RTOL)


def test_ingest_matches_port():
    settings, moddata, stars, _ = synth.make_problem("small")
    lnp = synth.make_lnp(moddata, synth.CONFIGS["small"]["axes"], 300, 20, seed=5)
    from megabeast_b200.helpers import likelihoods_from_lnp

    want = port.likelihoods_from_lnp(lnp, moddata)
    got = likelihoods_from_lnp({"vals": lnp["vals"].copy(), "indxs": lnp["indxs"]}, moddata)
    fin = np.isfinite(want["vals"])
    assert np.array_equal(np.isfinite(got["vals"]), fin)
    assert_close(got["vals"][fin], want["vals"][fin], 1e-15)


def test_tabulated_model_matches_analytic():
    """A model name the device does not know goes through the host callable + table path."""
    settings, moddata, stars, phis, expect, kind = load_golden("docs_small_uniform")
    model = MB_Model(copy.deepcopy(settings))
    spec, perm, ndim = model._spec()
    spec.kinds["Av"] = "tabulated"
    eng = LikelihoodEngine(spec, stars, moddata, massmult=port.mass_multipliers(moddata, {"name": "kroupa"}))
    x = eng.class_values("Av")
    from oracle import priors

    tab = np.array([priors.lognormal_density(x, p[5], p[6]) for p in phis])
    keep = [i for i in range(ndim) if i not in (5, 6)]
    v, st = eng.lnlike(phis[:, keep], {"Av": tab})
    assert_close(v, expect, RTOL)
    eng.close()


def test_two_devices_or_two_shards_sum_to_whole():
    """Star shards add up: two handles (same device if only one is visible) == one handle."""
    import torch

    settings, moddata, stars, _ = synth.make_problem("cfg1", seeds=(21, 22, 23))
    pm = port.PortModel(copy.deepcopy(settings))
    phis = synth.make_walkers(pm, 16, seed=23)
    expect = c_oracle.lnprob_batch(pm, phis, stars, moddata)
    devs = [0, 1] if torch.cuda.device_count() > 1 else [0, 0]
    model = MB_Model(copy.deepcopy(settings), devices=devs)
    assert_close(lnprob(phis, model, stars, moddata), expect, RTOL)


def test_config1_drivers_scipy_and_ensemble_sampler():
    """BASELINE config 1: the reference's two drivers (scipy minimize on -lnprob, pattern at
    singlepop_dust_model.py:348-362; the ensemble sampler of fit_ensemble :364-383) run unchanged
    on the GPU lnprob and follow the same trajectory as on the oracle."""
    from scipy.optimize import minimize

    from megabeast_b200 import fit_ensemble
    from megabeast_b200.sampler import StretchMoveSampler

    settings, moddata, stars, _ = synth.make_problem("cfg1")
    model = MB_Model(copy.deepcopy(settings))
    pm = port.PortModel(copy.deepcopy(settings))
    x0 = np.asarray(model.start_params()[1], dtype=float)

    r_gpu = minimize(lambda p: -lnprob(p, model, stars, moddata), x0, method="Nelder-Mead",
                     options={"maxfev": 50, "xatol": 1e-12, "fatol": 1e-12})
    r_cpu = minimize(lambda p: -port.lnprob(p, pm, stars, moddata, loop=False), x0, method="Nelder-Mead",
                     options={"maxfev": 50, "xatol": 1e-12, "fatol": 1e-12})
    assert r_gpu.nfev == r_cpu.nfev
    assert np.allclose(r_gpu.x, r_cpu.x, rtol=1e-9, atol=0)
    assert abs(r_gpu.fun - r_cpu.fun) <= 1e-9 * abs(r_cpu.fun)

    model2 = MB_Model(copy.deepcopy(settings))
    best = fit_ensemble(model2, stars, moddata, nsteps=6, seed=5)
    ndim = len(x0)
    rng = np.random.default_rng(5)
    pos = x0 * (1.0 + 1e-1 * rng.standard_normal((5 * ndim, ndim)))
    ref = StretchMoveSampler(5 * ndim, ndim, lambda p: c_oracle.lnprob_batch(pm, p, stars, moddata), seed=5)
    ref.run_mcmc(pos, 6)
    k, t = np.unravel_index(np.nanargmax(ref.lnprobability), ref.lnprobability.shape)
    assert np.allclose(best, ref.chain[k, t], rtol=1e-9, atol=0)
    assert np.isfinite(lnprob(best, model2, stars, moddata))


def test_fit_single_cli_end_to_end(tmp_path, capsys):
    """The reference's only entry point (fit_single.py:12-76) on npz inputs: settings file ->
    MB_Model -> GPU ingest of the lnp file -> start lnprob -> short fit -> final lnprob."""
    from megabeast_b200 import fit_single

    settings, moddata, _, cfg = synth.make_problem("cfg1")
    lnp = synth.make_lnp(moddata, cfg["axes"], 400, 30, seed=9)
    base = str(tmp_path / "sim")
    np.savez(base + "_moddata.npz", **{k: v for k, v in moddata.items() if k != "fA"})
    np.savez(base + "_lnp.npz", vals=lnp["vals"], indxs=lnp["indxs"])
    text = (f"beast_base = {base!r}\n"
            f"stellar_model = {settings.stellar_model!r}\n"
            f"fd_model = {settings.fd_model!r}\n")
    sfile = tmp_path / "settings.txt"
    sfile.write_text(text)
    best = fit_single.main([str(sfile), "--nsteps", "4", "--seed", "1"])
    out = capsys.readouterr().out
    assert "starting parameters" in out and "final parameters" in out
    # the start lnprob printed by the driver equals the oracle's on the same ingested data
    pm = port.PortModel(copy.deepcopy(settings))
    stars = port.likelihoods_from_lnp({"vals": lnp["vals"] - lnp["vals"].max(axis=0), "indxs": lnp["indxs"]}, moddata)
    want = port.lnprob(np.asarray(pm.start_params()[1], float), pm, stars, moddata, loop=False)
    lines = out.splitlines()
    line = lines[[i for i, ln in enumerate(lines) if "logA_values1" in ln][0] + 1]
    got = float(line.rsplit("]", 1)[1])
    assert abs(got - want) <= 1e-9 * abs(want)
    assert best.shape == (9,)


@pytest.mark.parametrize("W", [9, 45, 100])
def test_pixel_batch_matches_per_pixel_oracle(W):
    """BASELINE config 5 (batched independent pixel fits): every pixel's walkers against the oracle
    run on that pixel's stars alone (its own star count in the Poisson term)."""
    from megabeast_b200.pixels import PixelEnsemble

    settings, moddata, stars, offs = synth.make_pixel_problem(n_pixels=37, mean_stars=40, seeds=(3, 4, 5))
    offs[6] = offs[5]  # an empty pixel: only the Poisson term with S = 0
    pe = PixelEnsemble(copy.deepcopy(settings), stars, moddata, offs)
    pm = port.PortModel(copy.deepcopy(settings))
    P = len(offs) - 1
    phis = np.stack([synth.make_walkers(pm, W, seed=100 + p) for p in range(P)])
    phis[2, 1, 5] = 99.0  # outside the box: -inf, flags of that column ignored
    got = pe.lnprob(phis)
    assert got.shape == (P, W) and got[2, 1] == -np.inf
    for p in range(P):
        sub = {"indxs": np.ascontiguousarray(stars["indxs"][:, offs[p]:offs[p + 1]]),
               "vals": np.ascontiguousarray(stars["vals"][:, offs[p]:offs[p + 1]])}
        pmp = port.PortModel(copy.deepcopy(settings))
        if sub["indxs"].shape[1] == 0:
            # no stars: lnlike = 0*ln(pred) - pred - gammaln(1) = -pred
            physmod, mods = pmp.physmod(phis[p, 0], moddata)
            mm = port.mass_multipliers(moddata, pmp.physics_model["M_ini"])
            pred = port.predicted_num_stars(mm, moddata, physmod, mods["logA"], loop=False)
            assert abs(got[p, 0] + pred) <= 1e-9 * pred
            continue
        want = c_oracle.lnprob_batch(pmp, phis[p], sub, moddata)
        assert_close(got[p], want, RTOL)
    pe.close()


def test_pixel_batch_fit_runs_all_pixels_at_once():
    from megabeast_b200.pixels import PixelEnsemble

    settings, moddata, stars, offs = synth.make_pixel_problem(n_pixels=16, mean_stars=30, seeds=(3, 4, 6))
    pe = PixelEnsemble(copy.deepcopy(settings), stars, moddata, offs)
    best, best_lp = pe.fit(nsteps=5, seed=2)
    assert best.shape == (16, 9) and np.all(np.isfinite(best_lp))
    # the reported best lnprob is reproducible from the reported parameters
    again = pe.lnprob(best[:, None, :])[:, 0]
    assert_close(again, best_lp, 1e-12)
    pe.close()


def test_fp32_table_mode_stated_tolerance():
    """Optional mode: 32-bit factor tables on chip (21 significant bits), sums in fp64.  Stated
    tolerance 1e-7 relative on lnprob (the fp64 default is held to 1e-9 everywhere else)."""
    settings, moddata, stars, _ = synth.make_problem("cfg1", seeds=(41, 42, 43))
    pm = port.PortModel(copy.deepcopy(settings))
    for W in (7, 45, 100):
        phis = synth.make_walkers(pm, W, seed=43)
        want = c_oracle.lnprob_batch(pm, phis, stars, moddata)
        model = MB_Model(copy.deepcopy(settings), table_precision="fp32")
        got = lnprob(phis, model, stars, moddata)
        err = np.max(np.abs(got - want) / np.abs(want))
        assert err < 1e-7, err
        exact = lnprob(phis, MB_Model(copy.deepcopy(settings)), stars, moddata)
        assert np.max(np.abs(exact - want) / np.abs(want)) < RTOL
