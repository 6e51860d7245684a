#!/usr/bin/env python
"""
BASELINE config 5 (batched independent pixel fits) on one GPU: P pixels x ~Poisson(mean) stars,
K sparse points, shared 1M-row grid, W = 5*ndim = 45 walkers per pixel, ALL pixels and walkers in
one launch.  Default P = 512 = one GPU's share of the 4,096-pixel job on 8 GPUs (pixels are
independent: ranks take disjoint pixel subsets, no collective).  Under torchrun every rank owns its
own `--pixels` pixels (different seeds), the timed region is bracketed by barriers, the time is the
max over ranks, and rank 0 prints ONE JSON line for the whole job (weak scaling by construction).

    python scripts/bench_pixels.py                                  # one GPU, 512 pixels
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 scripts/bench_pixels.py                 # BASELINE config 5: 4,096 pixels
"""
import argparse
import copy
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from megabeast_b200 import synth  # noqa: E402
from megabeast_b200.pixels import PixelEnsemble  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pixels", type=int, default=512)
    ap.add_argument("--mean-stars", type=int, default=500)
    ap.add_argument("--K", type=int, default=50)
    ap.add_argument("--index-mode", default="uniform")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--check", type=int, default=3, help="pixels checked against the oracle")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    axes = synth.CONFIGS["cfg3"]["axes"]
    # the grid is shared (seed 1); every rank draws its own pixels and stars
    settings, moddata, stars, offs = synth.make_pixel_problem(
        n_pixels=args.pixels, mean_stars=args.mean_stars, axes=axes, K=args.K, mode=args.index_mode,
        seeds=(1, 2 + 10 * rank, 4 + 10 * rank))
    t0 = time.perf_counter()
    pe = PixelEnsemble(copy.deepcopy(settings), stars, moddata, offs, device=local)
    t_ingest = time.perf_counter() - t0
    P, W, ndim = args.pixels, 45, pe.ndim
    rng = np.random.default_rng(3)
    start = np.asarray(pe.model.start_params()[1], dtype=float)
    phis = start * (1.0 + 0.05 * rng.standard_normal((P, W, ndim)))
    phis = np.clip(phis, *pe.model._box())
    dev = torch.device("cuda", local)
    d_phi = torch.from_numpy(np.ascontiguousarray(phis)).to(dev)
    d_out = torch.empty((P, 4, W), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(5):
        pe.engine.lnlike_pixels_device(d_phi.data_ptr(), W, ndim, d_out.data_ptr(), stream)
    torch.cuda.synchronize()
    tot = 0.0
    barrier()
    for _ in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        pe.engine.lnlike_pixels_device(d_phi.data_ptr(), W, ndim, d_out.data_ptr(), stream)
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    barrier()
    if world > 1:
        t = torch.tensor([tot], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot = float(t.item())
    pe.engine.set_profiling(True)
    pe.engine.lnlike_pixels_device(d_phi.data_ptr(), W, ndim, d_out.data_ptr(), stream)
    torch.cuda.synchronize()
    kms = pe.engine.last_kernel_ms()
    pe.engine.set_profiling(False)
    barrier()
    for _ in range(2):  # first call: staging buffers are allocated
        vals = pe.lnprob(phis)
    barrier()
    calls = []
    for _ in range(20):
        t0 = time.perf_counter()
        vals = pe.lnprob(phis)
        calls.append(time.perf_counter() - t0)
    e2e = float(np.mean(calls))
    e2e_median = float(np.median(calls))
    # the C call alone (box test, copies, launch, combine), without the Python wrapper
    lows, highs = pe.model._box()
    park = np.asarray(pe.model.start_params()[1], dtype=float)
    ccalls = []
    for _ in range(20):
        t0 = time.perf_counter()
        pe.engine.lnprob_pixels_box(phis, lows, highs, park)
        ccalls.append(time.perf_counter() - t0)
    barrier()
    if world > 1:
        t = torch.tensor([e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = float(t.item())
    # parity on a few pixels
    from oracle import c_oracle
    from oracle import lnprob_port as port

    err = 0.0
    for p in range(min(args.check, P)):
        sub = {"indxs": np.ascontiguousarray(stars["indxs"][:, offs[p]:offs[p + 1]]),
               "vals": np.ascontiguousarray(stars["vals"][:, offs[p]:offs[p + 1]])}
        want = c_oracle.lnprob_batch(port.PortModel(copy.deepcopy(settings)), phis[p], sub, moddata)
        err = max(err, float(np.max(np.abs(vals[p] - want) / np.abs(want))))
    info = pe.engine.info()
    step_ms = tot / args.steps
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    P = P * world  # whole job
    print(json.dumps({
        "metric": "pixel-ensemble lnprob evals/sec (all walkers of all pixels per launch)",
        "value": P * W / (step_ms * 1e-3), "unit": "evals/s", "ms_per_step": step_ms, "n_gpus": world,
        "scaling": "weak", "timing": "CUDA events per launch, max over ranks of the summed step times",
        "config": {"pixels": P, "pixels_per_gpu": P // world, "stars_per_gpu": int(offs[-1]), "K": args.K, "walkers_per_pixel": W,
                   "grid_rows": info["G"], "entries": info["n_entries"], "entries_raw": info["n_entries_raw"],
                   "index_mode": args.index_mode, "ingest_s": round(t_ingest, 2),
                   "l2": "flushed between steps"},
        "kernel_ms": kms,
        "e2e": {"value": P * W / e2e, "ms_per_step": 1e3 * e2e, "ms_median_rank0": 1e3 * e2e_median,
                "ms_c_call_median_rank0": 1e3 * float(np.median(ccalls)), "calls": 20,
                "api": "PixelEnsemble.lnprob(phis[P,W,ndim]) with host arrays"},
        "fits_per_second_at_600_evals_per_walker": P / (step_ms * 1e-3 * 601 * 1.0),
        "parity": {"pixels_checked": min(args.check, P), "max_rel_err": err, "tolerance": 1e-9, "rank": 0},
    }), flush=True)


if __name__ == "__main__":
    import contextlib

    sys.stdout.flush()
    _real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)  # NCCL's version banner goes to stderr; the JSON line stays alone on stdout
    with contextlib.redirect_stdout(_real):
        main()
    _real.flush()
