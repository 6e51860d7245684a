"""
CPU-side checks (no GPU compute): the C-ABI library builds, loads and exports every symbol
include/megabeast_b200.h declares; the host mirror of the reference interface (MB_Model,
lnprior, start_params, phi layout, error behaviour) matches the oracle; the product fails
loudly without a device.
"""
import copy
import ctypes
import os
import re

import numpy as np
import pytest

from golden_util import golden_names, load_golden
from megabeast_b200 import MB_Model, _cabi, lnprob, synth
from megabeast_b200.engine import ModelSpec, shard_bounds
from megabeast_b200.helpers import get_predicted_num_stars, precompute_mass_multipliers
from megabeast_b200.priormodel import PriorAgeModel, PriorDustModel, PriorMassModel, compute_bin_boundaries
from megabeast_b200.sampler import StretchMoveSampler
from oracle import lnprob_port as port
from oracle import priors

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    lib = _cabi.load()
    header = open(os.path.join(ROOT, "include", "megabeast_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(mb_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)
    assert lib.mb_abi_version() == _cabi.MB_ABI_VERSION


def test_enum_maps_match_the_header():
    """The ctypes mirror's mode / prior-kind / status numbers are the header's."""
    header = open(os.path.join(ROOT, "include", "megabeast_b200.h")).read()
    modes = dict(re.findall(r"MB_MODE_([A-Z_]+) = (\d+)", header))
    assert {k.lower(): int(v) for k, v in modes.items()} == _cabi.MODE
    kinds = dict(re.findall(r"^  MB_([A-Z_]+) = (\d+)", header, flags=re.M))
    kinds = {k.lower(): int(v) for k, v in kinds.items() if not k.startswith(("MODE_", "OUT_", "P_"))}
    assert kinds == _cabi.KIND
    status = {k: int(v) for k, v in re.findall(r"#define MB_STATUS_([A-Z_]+) (\d+)", header)}
    assert status["NONFINITE_STAR"] == _cabi.STATUS_NONFINITE_STAR
    assert status["TOTAL_ZERO"] == _cabi.STATUS_TOTAL_ZERO
    assert status["PEER_TIMEOUT"] == _cabi.STATUS_PEER_TIMEOUT


def test_struct_layouts_match_header_sizes():
    # sizes computed by hand from the header's field lists (LP64)
    assert ctypes.sizeof(_cabi.ModelDesc) == 4 * 4 + 4 * 4 + 4 * 64 * 8
    assert ctypes.sizeof(_cabi.GridDesc) == 8 + 4 * 8 + 5 * 8
    assert ctypes.sizeof(_cabi.StarsDesc) == 7 * 8
    assert ctypes.sizeof(_cabi.MassMultDesc) == 8 + 3 * 8 + 8
    assert ctypes.sizeof(_cabi.Options) == 16 * 4
    assert ctypes.sizeof(_cabi.Info) == 5 * 4 + 4 * 4 + 6 * 4 + 2 * 0 + 4 + 6 * 8 + 8 * 8


def test_combine_is_host_arithmetic():
    lib = _cabi.load()
    # STARSUM, STARSUM_LO, NONFINITE, POISSON
    rows = np.array([[-10.0, 0.0, 5.0], [-0.25, 0.0, 0.125], [0.0, 0.0, 2.0], [-1.25, 0.0, 0.875]])
    out = np.empty(3)
    st = np.empty(3, np.int32)
    lib.mb_combine(_cabi.dptr(np.ascontiguousarray(rows)), 3, _cabi.dptr(out), st.ctypes.data_as(_cabi.c_int32_p))
    assert out.tolist() == [-11.5, 0.0, 6.0]
    assert st.tolist() == [0, _cabi.STATUS_TOTAL_ZERO, _cabi.STATUS_NONFINITE_STAR]


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device failure path")
def test_no_cpu_fallback():
    settings, moddata, stars, phis, expect, kind = load_golden("docs_small_uniform")
    model = MB_Model(copy.deepcopy(settings))
    with pytest.raises(_cabi.MegabeastB200Error, match="no CPU fallback"):
        lnprob(phis[0], model, stars, moddata)
    # outside the prior box the likelihood is never touched (:302-303): no device needed
    assert lnprob(phis[0] * 0 - 1, model, stars, moddata) == -np.inf


@pytest.mark.parametrize("name", golden_names())
def test_host_mirror_matches_port(name):
    settings, moddata, stars, phis, expect, kind = load_golden(name)
    m = MB_Model(copy.deepcopy(settings))
    pm = port.PortModel(copy.deepcopy(settings))
    assert m.start_params() == tuple(pm.start_params())
    assert np.array_equal(m.lnprior(phis), np.array([pm.lnprior(p) for p in phis]))
    for p in phis:
        assert m.lnprior(p) == pm.lnprior(p)
    spec, perm, ndim = m._spec()
    assert ndim == phis.shape[1] == spec.ndim
    assert sorted(perm.tolist()) == list(range(ndim))
    assert m.compute_N_stars == pm.compute_N_stars
    m._store_phi(phis[1])
    assert np.array_equal(np.asarray(m.start_params()[1], float), phis[1])


def test_error_behaviour_of_the_reference_is_kept():
    s = synth.docs_example_settings()
    del s.fd_model["Rv"]
    with pytest.raises(ValueError, match="requested parameter not in mbsetting file"):
        MB_Model(s)
    s = synth.docs_example_settings()
    s.fd_model["Av"]["prior"]["name"] = "gaussian"
    with pytest.raises(ValueError, match="gaussian prior not supported"):
        MB_Model(s).lnprior(np.zeros(9))
    s = synth.docs_example_settings()
    s.stellar_model["M_ini"]["prior"]["name"] = "flat"  # free IMF: KeyError 'nsubvars' upstream (:257)
    with pytest.raises(KeyError, match="nsubvars"):
        MB_Model(s).lnprior(np.zeros(10))


def test_model_spec_and_unknown_models_go_tabulated():
    s = synth.docs_example_settings()
    s.fd_model["Av"]["name"] = "flat"
    s.fd_model["Av"]["varnames"], s.fd_model["Av"]["varinit"] = [], []
    s.fd_model["Av"]["prior"] = {"name": "flat", "minmax": []}
    spec, perm, ndim = MB_Model(s)._spec()
    assert spec.kinds["Av"] == "flat" and ndim == 7
    spec = ModelSpec({"logA": "bins_histo", "Av": "tabulated"}, {"logA": [6, 7, 8]})
    assert spec.ndim == 3 and spec.tabulated == ["Av"]
    c = spec.to_c()
    assert list(c.kind) == [4, 7, 0, 0] and c.nnodes[0] == 3
    with pytest.raises(ValueError):
        ModelSpec({"Av": "nope"})


def test_shard_bounds_partition():
    for S, n in [(100000, 8), (7, 3), (5, 8), (0, 2)]:
        b = [shard_bounds(S, n, i) for i in range(n)]
        assert b[0][0] == 0 and b[-1][1] == S
        assert all(b[i][1] == b[i + 1][0] for i in range(n - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1


def test_host_prior_models_agree_with_oracle_restatement():
    x = np.array([0.0, 0.01, 0.3, 1.0, 2.5, 7.0])
    assert np.allclose(PriorDustModel({"name": "lognormal", "mean": 0.7, "sigma": 0.3})(x),
                       priors.lognormal_density(x, 0.7, 0.3), rtol=1e-15)
    m2 = {"name": "two_lognormal", "mean1": 0.4, "mean2": 2.0, "sigma1": 0.3, "sigma2": 0.4, "N1_to_N2": 2.5}
    assert np.allclose(PriorDustModel(m2)(x), priors.evaluate_prior_model(m2, x), rtol=1e-15)
    age = {"name": "bins_histo", "x": [6.0, 7.0, 8.0, 9.0, 10.0], "values": [1.0, 2.0, 3.0, 4.0, 5.0]}
    xs = np.array([6.0, 6.5, 7.0, 9.99, 10.0])
    assert np.array_equal(PriorAgeModel(age)(xs), priors.step_histogram(xs, age["x"], age["values"]))
    with pytest.raises(ValueError, match="outside of model x range"):
        PriorAgeModel(age)(np.array([10.5]))
    ms = np.logspace(-1.2, 2, 50)
    assert np.allclose(PriorMassModel({"name": "kroupa"})(ms), priors.kroupa_imf(ms), rtol=1e-14)
    assert np.allclose(compute_bin_boundaries([6.0, 6.1, 6.3]), priors.bin_boundaries([6.0, 6.1, 6.3]))
    with pytest.raises(NotImplementedError):
        PriorDustModel({"name": "kroupa"})


def test_host_helpers_match_port():
    settings, moddata, stars, _ = synth.make_problem("small")
    mm = precompute_mass_multipliers(moddata, PriorMassModel({"name": "kroupa"}))
    ref = port.mass_multipliers(moddata, {"name": "kroupa"})
    assert np.allclose(mm["massmult"], ref["massmult"], rtol=1e-14) and np.isclose(mm["avemass"], ref["avemass"], rtol=1e-15)
    assert np.array_equal(mm["ages"], ref["ages"]) and np.array_equal(mm["Zs"], ref["Zs"])
    pm = port.PortModel(settings)
    phi = synth.make_walkers(pm, 1, seed=2)[0]
    physmod, mods = pm.physmod(phi, moddata)
    want = port.predicted_num_stars(ref, moddata, physmod, mods["logA"], loop=True)
    got = get_predicted_num_stars(mm, moddata, physmod, PriorAgeModel(mods["logA"]))
    assert abs(got - want) <= 1e-13 * abs(want)


def test_stretch_move_sampler_on_a_gaussian():
    def lp(p):
        return -0.5 * np.sum(p * p, axis=-1)

    s = StretchMoveSampler(20, 3, lp, seed=1)
    p0 = np.random.default_rng(0).standard_normal((20, 3))
    s.run_mcmc(p0, 400)
    assert s.chain.shape == (20, 400, 3) and s.lnprobability.shape == (20, 400)
    tail = s.chain[:, 200:, :].reshape(-1, 3)
    assert np.all(np.abs(tail.mean(axis=0)) < 0.25) and np.all(np.abs(tail.std(axis=0) - 1.0) < 0.25)
    assert s.n_calls == 1 + 2 * 400  # one batched call per half-ensemble proposal
    assert 0.2 < s.acceptance_fraction.mean() < 0.9


def test_mbsettings_reads_the_docs_example(tmp_path):
    from megabeast_b200.mbsettings import mbsettings

    text = '''
# BEAST results files
beast_base = "beast_sim/scylla_sim/scylla_sim"   # comment after a value

# stellar population model
stellar_model = {
    "name": "bins_histo",
    "x": [6.0, 7.0, 8.0, 9.0, 10.0],  # units are log(years)
    "logA": {"name": "bins_histo", "varnames": ["values"],
             "varinit": [[1e-8, 1e-8, 1e-8, 1e-8, 1e-8]],
             "prior": {"name": "flat",
                       "minmax": [[0.0, 0.0, 0.0, 0.0, 0.0], [1e-3, 1e-3, 1e-3, 1e-3, 1e-3]]}},
    "M_ini": {"name": "kroupa", "varnames": ["slope"], "varinit": [2.35],
              "prior": {"name": "fixed", "minmax": [[2.0, 3.0]]}},
}
import numpy as np
fd_model = {
    "Av": {"name": "lognormal", "varnames": ["mean", "sigma"], "varinit": [0.5, 0.25],
           "prior": {"name": "flat", "minmax": [[0.005, 5.0], [0.05, 1.0]]}},
    "Rv": {"name": "lognormal", "varnames": ["mean", "sigma"], "varinit": [4.0, 0.25],
           "prior": {"name": "flat", "minmax": [[2.0, 6.0], [0.05, 1.0]]}},
    "fA": {"name": "lognormal", "varnames": ["mean", "sigma"], "varinit": [1.0, 0.25],
           "prior": {"name": "fixed", "minmax": [[0.0, 1.0], [0.05, 0.5]]}},
}
'''
    path = tmp_path / "settings.txt"
    path.write_text(text)
    s = mbsettings(str(path))
    assert s.beast_base == "beast_sim/scylla_sim/scylla_sim"
    assert not hasattr(s, "np")
    m = MB_Model(s)
    ref = MB_Model(synth.docs_example_settings())
    assert m.start_params() == ref.start_params()
    assert isinstance(mbsettings(), mbsettings)  # the reference's one live test (tests/basic_test.py:4-7)
