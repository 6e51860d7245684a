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
    """[3][W] rows of one shard, computed by the oracle."""
    rows = np.zeros((3, len(phis)))
    sub = {"indxs": stars["indxs"][:, lo:hi], "vals": stars["vals"][:, lo:hi]}
    S = stars["indxs"].shape[1]
    for w, phi in enumerate(phis):
        physmod, mods = pm.physmod(phi, moddata)
        star = port.star_sums(sub["vals"], sub["indxs"], physmod, moddata["prior_weight"],
                              moddata["grid_weight"], moddata["completeness"])
        rows[1, w] = np.count_nonzero(~np.isfinite(star))
        good = np.isfinite(star) & (star != 0)
        rows[0, w] = np.sum(np.log(star[good]))
        from scipy.special import gammaln

        pred = port.predicted_num_stars(pm.massmultipliers, moddata, physmod, mods["logA"], loop=False)
        rows[2, w] = S * np.log(pred) - pred - gammaln(S + 1)
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
