"""
Two real GPUs, one process each: star-sharded likelihood with the shard reduction (a) fused into
the kernel over NVLink peer memory and (b) as one NCCL all-reduce.  Skipped with fewer than two
devices.  Both must reproduce the oracle (1e-9 relative) and agree with each other bit for bit:
the kernel delivers the shard totals as two exact words (multiples of 2^-24 and 2^-68), so adding
them as integers inside the kernel or as doubles in NCCL is the same sum, rounded once.
Also: a prior model the device does not know (host callable, tabulated on the class values by every
rank) through both reductions.
"""
import copy
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _host_lognormal(x, mode, sigma):
    x = np.asarray(x, dtype=float)
    y = np.zeros_like(x)
    g = x > 0
    mu = np.log(mode) + sigma ** 2
    y[g] = (1.0 / np.sqrt(2.0 * np.pi)) / (x[g] * sigma) * np.exp(-0.5 * ((np.log(x[g]) - mu) / sigma) ** 2)
    return y


def _worker(rank, world, port_no, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from megabeast_b200 import MB_Model, synth
        from megabeast_b200.parallel import ShardedLikelihood

        settings, moddata, stars, _ = synth.make_problem("cfg1", seeds=(31, 32, 33))
        model = MB_Model(copy.deepcopy(settings))
        phis = synth.make_walkers(model, 45, seed=33)
        phis[3, 5] = 99.0  # one walker outside the box: -inf, likelihood not evaluated
        out = {}
        for mode in ("p2p", "nccl"):
            sh = ShardedLikelihood(MB_Model(copy.deepcopy(settings)), stars, moddata, device=rank, reduce=mode,
                                   check_sync=True)
            assert sh.reduce == mode
            # ranks that were handed different phi are caught before anything is launched
            try:
                sh.lnprob(phis * (1.0 + 1e-6 * rank))
                out[mode + "_desync_caught"] = world == 1
            except RuntimeError as exc:
                out[mode + "_desync_caught"] = "different phi" in str(exc)
            for _ in range(3):  # several launches: sequence numbers / parities advance
                vals = sh.lnprob(phis)
            out[mode] = np.array(vals)
            out[mode + "_one"] = float(sh.lnprob(phis[0]))
            sh.close()
            # the same model with the A(V) prior as an opaque host callable (-> MB_TABULATED)
            mt = MB_Model(copy.deepcopy(settings))
            av = mt.physics_model["Av"]
            av["name"] = "some_future_beast_model"
            av["model"] = lambda x, av=av: _host_lognormal(x, av["mean"], av["sigma"])
            sht = ShardedLikelihood(mt, stars, moddata, device=rank, reduce=mode)
            assert sht.spec.tabulated == ["Av"] and sht.reduce == mode
            out[mode + "_tab"] = np.array(sht.lnprob(phis))
            sht.close()
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_star_sharding_p2p_and_nccl(world):
    from megabeast_b200 import synth
    from oracle import c_oracle
    from oracle import lnprob_port as port

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    settings, moddata, stars, _ = synth.make_problem("cfg1", seeds=(31, 32, 33))
    pm = port.PortModel(copy.deepcopy(settings))
    phis = synth.make_walkers(pm, 45, seed=33)
    phis[3, 5] = 99.0
    want = c_oracle.lnprob_batch(pm, phis, stars, moddata)
    assert want[3] == -np.inf
    fin = np.isfinite(want)
    for rank in range(world):
        for mode in ("p2p", "nccl"):
            got = res[rank][mode]
            assert np.array_equal(got[~fin], want[~fin])
            assert np.max(np.abs(got[fin] - want[fin]) / np.abs(want[fin])) < 1e-9
            assert abs(res[rank][mode + "_one"] - want[0]) < 1e-9 * abs(want[0])
            assert res[rank][mode + "_desync_caught"] is True
            tab = res[rank][mode + "_tab"]
            assert np.array_equal(tab[~fin], want[~fin])
            assert np.max(np.abs(tab[fin] - want[fin]) / np.abs(want[fin])) < 1e-9
        assert np.array_equal(res[rank]["p2p"], res[0]["p2p"])  # every rank holds the same totals
    assert np.array_equal(res[0]["p2p"], res[0]["nccl"])
