"""
N > 1 host logic on CPU: two gloo ranks, stars sharded, the shard-additive result rows
all-reduced with megabeast_b200.parallel.allreduce_rows, combined with combine_rows.  The
per-shard partial rows come from the oracle here (no GPU in this test): what is being checked
is the sharding, the row layout (STARSUM, NONFINITE summed; POISSON shared) and the combine.
"""
import copy
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from megabeast_b200 import synth  # noqa: E402
from megabeast_b200.engine import shard_bounds  # noqa: E402
from megabeast_b200.parallel import allreduce_rows, combine_rows, dist_info  # noqa: E402
from oracle import lnprob_port as port  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _partial_rows(pm, phis, stars, moddata, lo, hi):
    """[4][W] rows of one shard, computed by the oracle (STARSUM split into a multiple of 2^-24 and the
    remainder, as the kernel delivers it)."""
    rows = np.zeros((4, len(phis)))
    sub = {"indxs": stars["indxs"][:, lo:hi], "vals": stars["vals"][:, lo:hi]}
    S = stars["indxs"].shape[1]
    for w, phi in enumerate(phis):
        physmod, mods = pm.physmod(phi, moddata)
        star = port.star_sums(sub["vals"], sub["indxs"], physmod, moddata["prior_weight"],
                              moddata["grid_weight"], moddata["completeness"])
        rows[2, w] = np.count_nonzero(~np.isfinite(star))
        good = np.isfinite(star) & (star != 0)
        total = np.sum(np.log(star[good]))
        rows[0, w] = np.round(total * 2.0 ** 24) / 2.0 ** 24
        rows[1, w] = total - rows[0, w]
        from scipy.special import gammaln

        pred = port.predicted_num_stars(pm.massmultipliers, moddata, physmod, mods["logA"], loop=False)
        rows[3, w] = S * np.log(pred) - pred - gammaln(S + 1)
    return rows


def _worker(rank, world, port_no, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert dist_info() == (rank, world)
        settings, moddata, stars, _ = synth.make_problem("small", seeds=(3, 4, 5))
        pm = port.PortModel(copy.deepcopy(settings))
        pm.massmultipliers = port.mass_multipliers(moddata, pm.physics_model["M_ini"])
        phis = synth.make_walkers(pm, 6, seed=5)
        S = stars["indxs"].shape[1]
        lo, hi = shard_bounds(S, world, rank)
        rows = torch.from_numpy(_partial_rows(pm, phis, stars, moddata, lo, hi))
        allreduce_rows(rows)
        vals, status = combine_rows(rows.numpy())
        want = np.array([port.lnprob(p, pm, stars, moddata, loop=False) for p in phis])
        q.put((rank, float(np.max(np.abs(vals - want) / np.abs(want))), int(status.max())))
    finally:
        dist.destroy_process_group()


def test_two_rank_star_sharding_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, st in res:
        assert err < 1e-12, (rank, err)
        assert st == 0


def _worker_tab(rank, world, port_no, q):
    """Tabulated prior models across ranks: every rank evaluates the same host callable on the same
    class values, so all ranks hold identical tables; the shard rows are added with
    reduce_host_rows (the path ShardedLikelihood takes for models the device does not know)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from megabeast_b200 import MB_Model
        from megabeast_b200.parallel import reduce_host_rows

        settings, moddata, stars, _ = synth.make_problem("small", seeds=(3, 4, 5))
        pm = port.PortModel(copy.deepcopy(settings))
        pm.massmultipliers = port.mass_multipliers(moddata, pm.physics_model["M_ini"])
        phis = synth.make_walkers(pm, 5, seed=6)
        model = MB_Model(copy.deepcopy(settings))
        av = model.physics_model["Av"]
        av["name"] = "some_future_beast_model"   # unknown to the device -> tabulated
        from oracle import priors

        av["model"] = lambda x, av=av: priors.lognormal_density(np.asarray(x, float), av["mean"], av["sigma"])
        spec, perm, ndim = model._spec()
        assert spec.tabulated == ["Av"] and perm.size == ndim - 2

        class FakeEngine:  # class values as mb_class_values reports them: the sorted distinct column values
            def class_values(self, p):
                return np.unique(moddata[p])

        tabs = model._tables(FakeEngine(), spec, phis)
        # the table IS the analytic prior on the class values, identical on every rank
        x = np.unique(moddata["Av"])
        want_tab = np.array([priors.lognormal_density(x, p[5], p[6]) for p in phis])
        assert np.allclose(tabs["Av"], want_tab, rtol=1e-15)
        t = torch.from_numpy(tabs["Av"].copy())
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert np.array_equal(t.numpy(), tabs["Av"])
        S = stars["indxs"].shape[1]
        lo, hi = shard_bounds(S, world, rank)
        rows = reduce_host_rows(_partial_rows(pm, phis, stars, moddata, lo, hi))
        vals, status = combine_rows(rows)
        want = np.array([port.lnprob(p, pm, stars, moddata, loop=False) for p in phis])
        q.put((rank, float(np.max(np.abs(vals - want) / np.abs(want))), int(status.max())))
    finally:
        dist.destroy_process_group()


def test_two_rank_tabulated_models_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker_tab, args=(r, world, port_no, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, st in res:
        assert err < 1e-12, (rank, err)
        assert st == 0
