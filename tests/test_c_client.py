"""
The drop-in boundary is a C ABI: a plain C program (tests/c_client/cabi_client.c; gcc, no Python,
no torch, no CUDA headers) links libmegabeast_b200.so and evaluates the likelihood.

* CPU box: it compiles and links against every symbol it uses, and -- there being no CUDA device
  and no CPU fallback -- fails loudly with the library's message.
* GPU box: its output equals the oracle's lnlike (1e-9 relative).
"""
import copy
import os
import subprocess

import numpy as np
import pytest

from megabeast_b200 import _cabi, synth
from oracle import lnprob_port as port

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _build_client(tmp_path):
    _cabi.load()  # makes sure the library is built
    libdir = os.path.join(ROOT, "megabeast_b200")
    exe = str(tmp_path / "cabi_client")
    cmd = ["gcc", "-O1", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_client", "cabi_client.c"), "-o", exe,
           "-L", libdir, "-lmegabeast_b200", f"-Wl,-rpath,{libdir}"]
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run(cmd, check=True, env=env, capture_output=True, text=True)
    return exe


def _write_problem(path, W=5):
    settings, moddata, stars, cfg = synth.make_problem("tiny")
    pm = port.PortModel(copy.deepcopy(settings))
    phis = synth.make_walkers(pm, W, seed=3)
    mm = port.mass_multipliers(moddata, pm.physics_model["M_ini"])  # once per fit, a host constant
    nodes = np.asarray(settings.stellar_model["x"], dtype=np.float64)
    K, S = stars["indxs"].shape
    G = moddata["Av"].size
    ages, mets = np.asarray(mm["ages"], float), np.asarray(mm["Zs"], float)
    with open(path, "wb") as f:
        np.array([G, K, S, W, phis.shape[1], nodes.size, ages.size, mets.size], dtype=np.int64).tofile(f)
        np.array([mm["avemass"]], dtype=np.float64).tofile(f)
        for a in (nodes, moddata["logA"], moddata["Z"], moddata["Av"], moddata["Rv"], moddata["prior_weight"],
                  moddata["grid_weight"], moddata["completeness"], ages, mets, mm["massmult"], phis, stars["vals"]):
            np.ascontiguousarray(a, dtype=np.float64).tofile(f)
        np.ascontiguousarray(stars["indxs"], dtype=np.int64).tofile(f)
    return pm, phis, stars, moddata


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_c_client_links_and_fails_loudly_without_a_device(tmp_path):
    exe = _build_client(tmp_path)
    _write_problem(str(tmp_path / "p.bin"))
    if _has_gpu():
        pytest.skip("a CUDA device is present: covered by the gpu test")
    res = subprocess.run([exe, str(tmp_path / "p.bin")], capture_output=True, text=True)
    assert res.returncode == 3
    assert "no CPU fallback" in res.stderr


@pytest.mark.gpu
def test_c_client_matches_oracle(tmp_path):
    exe = _build_client(tmp_path)
    pm, phis, stars, moddata = _write_problem(str(tmp_path / "p.bin"))
    res = subprocess.run([exe, str(tmp_path / "p.bin")], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    rows = [ln.split() for ln in res.stdout.strip().splitlines()]
    got = np.array([float(r[0]) for r in rows])
    assert all(int(r[1]) == 0 for r in rows)
    want = np.array([pm.lnlike(p, stars, moddata, loop=False) for p in phis])
    assert np.max(np.abs(got - want) / np.abs(want)) < 1e-9
