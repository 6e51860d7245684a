"""
The CUDA path against the C oracle on random small problems: grid shapes down to one class per axis, model
variants (SFH fixed, R(V) fixed, A(V) fixed, f_A free, two_lognormal A(V)), 0-30 % masking, zeroed completeness
rows, all-zero stars, out-of-box walkers, one star, one sparse point -- every engine mode.  Values to 1e-9
relative, exceptions alike (singlepop_dust_model.py:216-217,225-227).
"""
import contextlib
import copy
import io

import numpy as np
import pytest

from fuzz_util import random_problem
from golden_util import assert_close
from megabeast_b200 import MB_Model, lnprob, synth
from oracle import c_oracle
from oracle import lnprob_port as port

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["auto", "factor", "table_unique", "table_grid", "elementwise"])
def test_random_small_problems_match_the_oracle(mode):
    rng = np.random.default_rng(4242)
    compared = 0
    for _ in range(40):
        try:
            settings, moddata, stars, phis = random_problem(rng)
        except ValueError:   # (fewer distinct grid rows than sparse points asked for)
            continue
        pm = port.PortModel(copy.deepcopy(settings))
        want, wkind = np.full(len(phis), np.nan), np.zeros(len(phis), int)
        for i, phi in enumerate(phis):   # the port one walker at a time: values and exceptions
            try:
                with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
                    want[i] = port.lnprob(phi, pm, stars, moddata, loop=False)
            except ValueError:
                wkind[i] = 1
            except SystemExit:
                wkind[i] = 2
        model = MB_Model(copy.deepcopy(settings), mode=mode)
        got, gkind = np.full(len(phis), np.nan), np.zeros(len(phis), int)
        for i, phi in enumerate(phis):
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    got[i] = lnprob(phi, model, stars, moddata)
            except ValueError:
                gkind[i] = 1
            except SystemExit:
                gkind[i] = 2
        assert np.array_equal(gkind, wkind), (settings.fd_model, settings.stellar_model["logA"]["prior"], gkind, wkind)
        ok = wkind == 0
        assert_close(got[ok], want[ok], 1e-9)
        if ok.all():   # and as one batch, against the C oracle -- also wide batches (64- and 128-walker kernels, two passes)
            W = int(rng.choice([3, 3, 40, 70, 130]))
            wide = phis if W == 3 else np.concatenate([phis, synth.make_walkers(pm, W - 3, seed=int(rng.integers(1 << 20)))])
            expect = c_oracle.lnprob_batch(port.PortModel(copy.deepcopy(settings)), wide, stars, moddata)
            if np.all(np.isfinite(expect)):   # (a batch raises as soon as any of its walkers would)
                assert_close(lnprob(wide, model, stars, moddata), expect, 1e-9)
        compared += 1
    assert compared >= 30


def test_random_pixel_batches_match_the_per_pixel_oracle():
    """Pixel batches on the same random problems: the stars cut into random pixels (empty ones included),
    every model variant, 3 to 70 walkers per pixel; each pixel against the oracle run on its stars alone."""
    from megabeast_b200.pixels import PixelEnsemble

    rng = np.random.default_rng(777)
    compared = 0
    for _ in range(30):
        try:
            settings, moddata, stars, phis = random_problem(rng)
        except ValueError:
            continue
        S = stars["indxs"].shape[1]
        P = int(rng.integers(1, 6))
        offs = np.sort(np.concatenate([[0, S], rng.integers(0, S + 1, size=P - 1)])).astype(np.int64)
        pm = port.PortModel(copy.deepcopy(settings))
        W = int(rng.choice([3, 9, 45, 70]))
        walkers = np.stack([synth.make_walkers(pm, W, seed=int(rng.integers(1 << 20))) for _ in range(P)])
        want = np.empty((P, W))
        raises = False
        for p in range(P):
            sub = {"indxs": np.ascontiguousarray(stars["indxs"][:, offs[p]:offs[p + 1]]),
                   "vals": np.ascontiguousarray(stars["vals"][:, offs[p]:offs[p + 1]])}
            for w in range(W):
                try:
                    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
                        want[p, w] = port.lnprob(walkers[p, w], port.PortModel(copy.deepcopy(settings)), sub, moddata, loop=False)
                except (ValueError, SystemExit):
                    raises = True
        pe = PixelEnsemble(copy.deepcopy(settings), stars, moddata, offs)
        if raises:   # the batch raises as soon as any walker of any pixel would
            with pytest.raises((ValueError, SystemExit)), contextlib.redirect_stdout(io.StringIO()):
                pe.lnprob(walkers)
        else:
            assert_close(pe.lnprob(walkers), want, 1e-9)
            compared += 1
        pe.close()
    assert compared >= 15
