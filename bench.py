#!/usr/bin/env python
"""
bench.py -- ensemble lnprob evaluations per second on BASELINE.json's headline workload
(config 3: 100,000 stars x 50 sparse points, 1,000,400-row grid, 64 walkers; docs-example model).

    python bench.py [--gpus N] [--steps K] [--warmup W]             # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [...]   # the reference's CPU algorithm

A "step" is one pass of the hot path over one batch of W = 64 walkers (one emcee proposal
round).  N > 1 is launched by torchrun, one rank per GPU: the 100,000 stars are sharded over
the ranks (grid replicated), one NCCL all-reduce of 2*W doubles per step -> "strong" scaling of
the fixed-size job BASELINE.json names.  Rank 0 prints ONE JSON line.

Keys beyond the base contract:
  value        device-timed (CUDA events on the launching stream), phi resident in HBM
  e2e          the same metric through the public API with HOST buffers: lnprob(phis, model,
               star_lnpgriddata, beast_moddata) -- prior box on the host, pinned H2D of phi,
               kernels, (all-reduce,) D2H of the result rows, exceptions checked
  roofline     dominant kernel: algorithmic bytes per launch / its CUDA-event duration vs the
               measured HBM copy bandwidth (MEASURED_PEAKS.json).  See DESIGN.md section 5 for why
               the factor-mode hot kernel is bounded by shared-memory bandwidth instead, and
               `onchip` for that ceiling.
  cpu_baseline the oracle's numpy port (the reference's algorithm, one core) on a bounded
               sample of the same workload, timed on this box; its values double as a full-size
               parity check of the GPU path (`parity`).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ensemble lnprob evals/sec (100k stars x 1M grid, 64 walkers)"
UNIT = "evals/s"
L2_FLUSH_BYTES = 256 << 20


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--index-mode", default="uniform", choices=["uniform", "clustered"])
    ap.add_argument("--mode", default="auto", choices=["auto", "factor", "table_unique", "table_grid", "elementwise"])
    ap.add_argument("--walkers", type=int, default=0, help="override the config's walker count")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=4, help="walkers the CPU baseline evaluates")
    ap.add_argument("--no-flush", action="store_true", help="keep L2 warm between steps (not the headline)")
    ap.add_argument("--all-modes", action="store_true", help="also time the other engine modes")
    ap.add_argument("--no-clustered", action="store_true",
                    help="skip the extra measurement with clustered (BEAST-like) sparse indices")
    ap.add_argument("--table-precision", default="fp64", choices=["fp64", "fp32"],
                    help="fp32: 32-bit factor tables on chip, stated tolerance 1e-7 (not the headline)")
    ap.add_argument("--reduce", default="auto", choices=["auto", "p2p", "nccl"],
                    help="N > 1: how the star-shard partials are added")
    return ap.parse_args()


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def metric_name(args, cfg, G, W):
    if args.workload == "cfg3" and W == 64:
        return METRIC
    return f"ensemble lnprob evals/sec ({cfg['S']} stars x {G} grid, {W} walkers)"


def make_workload(args):
    import copy

    from megabeast_b200 import synth

    settings, moddata, stars, cfg = synth.make_problem(args.workload, mode=args.index_mode)
    W = args.walkers or cfg["W"]
    return copy.deepcopy(settings), moddata, stars, cfg, W


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's algorithm (numpy, per-star Python loop) on all host cores
# ---------------------------------------------------------------------------------------------------
_REF = {}


def _ref_eval(phi):
    from oracle import lnprob_port as port

    return port.lnprob(phi, _REF["pm"], _REF["stars"], _REF["mod"], loop=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the CPU arm is a one-process job; the other ranks have nothing to do
    import multiprocessing as mp

    import numpy as np

    from megabeast_b200 import synth
    from oracle import lnprob_port as port

    settings, moddata, stars, cfg, W = make_workload(args)
    pm = port.PortModel(settings)
    phis = synth.make_walkers(pm, W, seed=3)
    pm.massmultipliers = port.mass_multipliers(moddata, pm.physics_model["M_ini"])  # once per fit
    _REF.update(pm=pm, stars=stars, mod=moddata)
    cores = host_cores()
    per_step = min(cores, W)
    budget_s = 240.0
    ctx = mp.get_context("fork")
    with ctx.Pool(per_step) as pool:
        t0 = time.perf_counter()
        pool.map(_ref_eval, list(phis[:per_step]))
        t_first = time.perf_counter() - t0
        warm = max(0, min(args.warmup, int(0.2 * budget_s / t_first)) - 1)
        for _ in range(warm):
            pool.map(_ref_eval, list(phis[:per_step]))
        steps = max(1, min(args.steps, int(0.75 * budget_s / t_first)))
        t0 = time.perf_counter()
        for i in range(steps):
            lo = (i * per_step) % W
            sel = [phis[(lo + j) % W] for j in range(per_step)]
            pool.map(_ref_eval, sel)
        dt = time.perf_counter() - t0
    value = per_step * steps / dt
    sample = (f"each step = {per_step} of {W} walkers x all {cfg['S']} stars, one walker per core; "
              f"{steps} of the requested {args.steps} steps run to stay inside {budget_s:.0f} s")
    line = {
        "impl": "reference",
        "metric": metric_name(args, cfg, len(moddata["Av"]), W), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm + 1, "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: {cfg['S']} stars x {cfg['K']} sparse points, "
                               f"{len(moddata['Av'])}-row grid, {W} walkers, docs-example model",
                   "index_mode": args.index_mode,
                   "what": "oracle/lnprob_port.py loop=True: numpy restatement of the reference's "
                           "lnprob (per-star Python loop, per-bin grid masks), walker-parallel "
                           "over the host cores with multiprocessing"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": per_step, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.err = index, [], False, None
        self.active = False

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self.stop_flag:
                if self.active:
                    try:
                        reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:
                        reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), reasons,
                                         nv.nvmlDeviceGetPowerUsage(h) / 1e3))
                time.sleep(0.004)
        except Exception as exc:  # NVML missing: report, do not fail the bench
            self.err = repr(exc)

    def summary(self):
        if self.err or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": getattr(self, "max_mhz", None), "reasons": [],
                    "note": self.err or "no samples"}
        import statistics

        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
                 0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
                 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        seen = 0
        for _, r, _ in self.samples:
            seen |= r
        reasons = [n for b, n in names.items() if seen & b and n != "gpu_idle"]
        return {"sm_mhz": statistics.median(s[0] for s in self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.samples),
                "power_w_max": max(s[2] for s in self.samples)}


def run_ours(args):
    import copy

    import numpy as np
    import torch
    import torch.distributed as dist

    from megabeast_b200 import MB_Model, lnprob, synth
    from megabeast_b200.parallel import ShardedLikelihood, combine_rows

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; megabeast_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    settings, moddata, stars, cfg, W = make_workload(args)
    model = MB_Model(copy.deepcopy(settings), devices=[local], mode=args.mode,
                     table_precision=args.table_precision)
    phis = synth.make_walkers(model, W, seed=3)
    t0 = time.perf_counter()
    sharded = ShardedLikelihood(model, stars, moddata, device=local, mode=args.mode, reduce=args.reduce)
    t_ingest = time.perf_counter() - t0
    info = sharded.engine.info()
    ndim = sharded.ndim
    d_phi = torch.from_numpy(np.ascontiguousarray(phis[:, sharded.perm])).to(dev)
    d_rows = torch.empty((3, W), dtype=torch.float64, device=dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    K = args.steps

    def flush_l2():
        if not args.no_flush:
            flush.zero_()

    # ---- device-timed value ------------------------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        flush_l2()
        sharded.lnlike_device(d_phi, d_rows)
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    sampler.active = True
    t_wall = time.perf_counter()
    for i in range(K):
        flush_l2()
        ev[i][0].record()
        sharded.lnlike_device(d_phi, d_rows)
        ev[i][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = max_over_ranks(sum(step_ms))
    launches = sharded.engine.info()["launches_last_eval"] * K
    rows_dev = d_rows.cpu().numpy()
    vals_dev, status_dev = combine_rows(rows_dev)

    # ---- end to end through the public API, host buffers ----------------------------------------------
    if world == 1:
        def e2e_call():
            return lnprob(phis, model, stars, moddata)
    else:
        def e2e_call():
            return sharded.lnprob(phis)
    for _ in range(max(3, args.warmup)):
        flush_l2()
        e2e_vals = e2e_call()
    barrier()
    e2e_s = 0.0
    for i in range(K):
        flush_l2()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_vals = e2e_call()
        e2e_s += time.perf_counter() - t0
    barrier()
    sampler.active = False
    sampler.stop_flag = True
    sampler.join(timeout=2)
    e2e_s = max_over_ranks(e2e_s)

    # ---- weak scaling (N > 1): every rank holds a FULL 100k-star shard of an N x 100k-star job --------
    weak = None
    if world > 1:
        shw = ShardedLikelihood(MB_Model(copy.deepcopy(settings), devices=[local], mode=args.mode), stars, moddata,
                                device=local, mode=args.mode, reduce=args.reduce,
                                star_range=(0, cfg["S"]), n_stars_total=world * cfg["S"])
        d_rows_w = torch.empty((3, W), dtype=torch.float64, device=dev)
        for _ in range(max(3, args.warmup)):
            flush_l2()
            shw.lnlike_device(d_phi, d_rows_w)
        evw = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        barrier()
        for i in range(K):
            flush_l2()
            evw[i][0].record()
            shw.lnlike_device(d_phi, d_rows_w)
            evw[i][1].record()
        barrier()
        weak_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in evw))
        # every rank holds the same stars: the job's star sum must be N x the single-copy one
        ratio = (d_rows_w[0] / (world * d_rows[0])).cpu().numpy()
        weak = {"stars_per_gpu": cfg["S"], "stars_total": world * cfg["S"],
                "value": world * W * K / (weak_ms * 1e-3), "unit": "100k-star-equivalent evals/s",
                "job_evals_per_s": W * K / (weak_ms * 1e-3), "ms_per_step": weak_ms / K,
                "reduce": shw.reduce,
                "check_starsum_vs_n_copies_max_rel_err": float(np.max(np.abs(ratio - 1.0))),
                "note": "per-GPU work fixed (each rank a full config-3 star shard), one reduction of "
                        f"{2 * W} partials per step; value counts an evaluation over N x 100k stars as N evaluations"}
        shw.close()

    # ---- per-kernel times (CUDA events inside the library, same stream) ------------------------------
    eng = sharded.engine
    eng.set_profiling(True)
    kms = {}
    nprof = min(K, 50)
    for _ in range(nprof):
        flush_l2()
        sharded.lnlike_device(d_phi, d_rows)
        torch.cuda.synchronize()
        for k, v in eng.last_kernel_ms().items():
            kms[k] = kms.get(k, 0.0) + v / nprof
    eng.set_profiling(False)
    dom = max(kms, key=kms.get)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel --------------------------------------------------------------
    peak, peak_src = measured_peaks()
    Wpad = 32 * (1 if W <= 32 else 2 if W <= 64 else 4)
    n_ent = info["n_entries"]
    nA, nB = info["n_age_classes"], info["n_dust_classes"]
    mode = info["mode"]
    by_row = mode in ("table_grid", "elementwise")  # physmod materialised over every grid row
    U = len(moddata["Av"]) if by_row else nA * nB
    nfree = sum(1 for p in ("logA", "Av", "Rv", "fA") if sharded.spec.kinds[p] != "fixed")
    touched = None
    if by_row:
        lo, hi = sharded.star_range
        sub_i, sub_v = stars["indxs"][:, lo:hi], stars["vals"][:, lo:hi]
        touched = int(np.unique(sub_i[np.isfinite(sub_v)]).size)
    alg = {
        "k0_factors": 8 * W * ndim + 8 * (nA + nB) * Wpad,
        # table_grid: row codes in, table out; elementwise: the free columns + grid_weight and
        # completeness (N-stars partial sums) in, table out
        "k1_expand": (4 * U if mode == "table_grid" else (8 * nfree + 16) * U if mode == "elementwise" else 0)
                     + 8 * U * Wpad,
        "k2_stars": (8 + info["code_bytes"]) * n_ent + 16 * info["n_bins"] * nB + (8 * (nA + nB) * Wpad if mode == "factor"
                                  else 8 * Wpad * (touched if touched is not None else U)),
    }
    dur_s = kms[dom] * 1e-3
    achieved = alg[dom] / dur_s / 1e9
    traffic, traffic_src = None, "no ncu capture for this kernel/workload/mode in profiles/ncu_traffic.json"
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            hit = json.load(f).get(f"{dom}|{args.workload}|{mode}|{world}")
        if hit and args.index_mode == "uniform":
            traffic, traffic_src = hit["bytes"], hit["source"]
    except Exception:
        pass
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg[dom], "avg_launch_ms": kms[dom],
                "traffic_source": traffic_src,
                "note": "this kernel is bounded by the shared-memory/L1 data pipe, not HBM: see `onchip` and DESIGN.md section 5"
                        if mode == "factor" else
                        "K1e is bound by the fp64 exp of the lognormal priors (one per parameter, row and walker), not by HBM"
                        if (mode == "elementwise" and dom == "k1_expand") else "HBM-bound: materialised table written / gathered"}
    clocks = sampler.summary()
    sm_mhz = clocks.get("sm_mhz") or 0
    onchip = None
    if mode == "factor" and sm_mhz:
        # shared-memory wavefronts K2 must issue (measured in lds_mix.cu): per
        # entry 2*NW for the Wpad*8-byte factor row + 1.5 for the broadcast of weight and code
        nw = Wpad // 32
        bcast = 1.5 if info["code_bytes"] == 4 else 1.25  # four codes: one 16-byte or one 8-byte broadcast
        wavefronts = n_ent * (2.0 * nw + bcast)
        peak_wf = info["sm_count"] * sm_mhz * 1e6 / 1e9
        ach = wavefronts / (kms["k2_stars"] * 1e-3) / 1e9
        onchip = {"bound": "smem", "wavefronts_per_launch": wavefronts, "achieved": ach, "peak": peak_wf,
                  "unit": "Gwavefront/s", "frac": ach / peak_wf,
                  "note": f"model count ({2 * nw} factor + {bcast} broadcast wavefronts per entry); peak = 1 "
                          f"wavefront/clk/SM x {info['sm_count']} SMs x {sm_mhz:.0f} MHz; ncu's measured "
                          "l1tex data-pipe utilisation for the same kernel is in profiles/"}
    # the north_star's own yardstick: the materialised-physmod design's compulsory bytes per step
    G = len(moddata["Av"])
    P = sum(1 for p in ("logA", "Av", "Rv", "fA") if sharded.spec.kinds[p] != "fixed")
    raw = info["n_entries_raw"]
    survey_bytes = G * (8 * P + 18) + 8 * G * W + 12 * raw + 8 * W * min(G, raw) + 8 * W
    step_s = total_ms * 1e-3 / K
    survey = {"bytes_per_step": survey_bytes, "achieved": survey_bytes / step_s / 1e9, "peak": peak,
              "frac": survey_bytes / step_s / 1e9 / peak, "unit": "GB/s",
              "note": "SURVEY.md 8(d) compulsory bytes of the materialised-physmod design / this "
                      "run's step time: >1 means the step beats that design's HBM floor, because "
                      "the factor path never writes or gathers physmod (its real DRAM traffic is "
                      "in roofline/profiles)"}

    line = {
        "metric": metric_name(args, cfg, G, W), "value": W * K / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": K, "warmup": max(3, args.warmup), "ms_per_step": total_ms / K,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64" if args.table_precision == "fp64" else "f64 sums, 32-bit (21 significant bits) factor tables",
        "data": "synthetic",
        "config": {
            "workload": f"{args.workload}: {cfg['S']} stars x {cfg['K']} sparse points, {G}-row grid, "
                        f"{W} walkers, docs-example model (9 parameters)",
            "index_mode": args.index_mode, "engine_mode": mode,
            "sharding": f"stars split over {world} rank(s), grid replicated; shard reduction: "
                        + {"none": "none (1 GPU)", "nccl": f"one NCCL all-reduce of {2 * W} doubles per step",
                           "p2p": "fused into K2's tail over NVLink peer memory (no collective launch)"}[sharded.reduce],
            "l2": "kept warm" if args.no_flush else f"flushed between steps ({L2_FLUSH_BYTES >> 20} MiB memset, outside the timed events)",
            "entries": n_ent, "entries_raw": raw, "age_classes": nA, "dust_classes": nB,
            "bins": info["n_bins"], "ingest_s": round(t_ingest, 3),
        },
        "clocks": clocks,
        "e2e": {"value": W * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * W * ndim,
                "d2h_bytes_per_step": 8 * 3 * W, "ms_per_step": 1e3 * e2e_s / K,
                "api": "megabeast_b200.lnprob(phis[W,ndim], MB_Model, star_lnpgriddata, beast_moddata)"
                       if world == 1 else "megabeast_b200.parallel.ShardedLikelihood.lnprob(phis[W,ndim])"},
        "gpu_launches": launches,
        "kernel_ms": kms,
        "roofline": roofline,
        "onchip": onchip,
        "north_star_equivalent": survey,
        "wall_ms_per_step_incl_flush": 1e3 * t_wall / K,
    }
    if weak is not None:
        line["weak_scaling"] = weak

    # ---- CPU baseline + full-size parity ----------------------------------------------------------------
    if world == 1 and not args.no_cpu_baseline:
        from oracle import lnprob_port as port

        pm = port.PortModel(copy.deepcopy(settings))
        pm.massmultipliers = port.mass_multipliers(moddata, pm.physics_model["M_ini"])
        n = max(1, min(args.cpu_sample, W))
        port.lnprob(phis[0], pm, stars, moddata, loop=True)  # warm-up call
        t0 = time.perf_counter()
        cpu_vals = np.array([port.lnprob(p, pm, stars, moddata, loop=True) for p in phis[:n]])
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {
            "value": n / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n} of {W} walkers x all {cfg['S']} stars; oracle/lnprob_port.py loop=True "
                      f"(numpy restatement of the reference's lnprob, per-star Python loop); {dt:.1f} s"}
        rel = np.abs(vals_dev[:n] - cpu_vals) / np.abs(cpu_vals)
        rel2 = np.abs(np.asarray(e2e_vals)[:n] - cpu_vals) / np.abs(cpu_vals)
        line["parity"] = {"walkers_checked": n, "max_rel_err_device_path": float(rel.max()),
                          "max_rel_err_e2e_path": float(rel2.max()), "tolerance": 1e-9,
                          "status_flags": int(status_dev.max())}
    if world == 1 and args.index_mode == "uniform" and not args.no_clustered:
        # the headline uses uniform-random sparse indices (worst case: nothing to merge, no locality);
        # real BEAST lnp files hold the K best models around each star's best fit -- same job, same
        # kernels, with the top-K points clustered in grid-axis space
        t0 = time.perf_counter()
        _, mod_c, stars_c, _ = synth.make_problem(args.workload, mode="clustered")
        mdl = MB_Model(copy.deepcopy(settings), devices=[local], mode=args.mode, table_precision=args.table_precision)
        shc = ShardedLikelihood(mdl, stars_c, mod_c, device=local, mode=args.mode)
        t_setup = time.perf_counter() - t0
        for _ in range(5):
            flush_l2()
            shc.lnlike_device(d_phi, d_rows)
        torch.cuda.synchronize()
        tot, nrep = 0.0, min(K, 100)
        for _ in range(nrep):
            flush_l2()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            shc.lnlike_device(d_phi, d_rows)
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        lnprob(phis, mdl, stars_c, mod_c)  # builds the model's own resident copy of this data set
        e2e_c = 0.0
        for _ in range(nrep):
            flush_l2()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            lnprob(phis, mdl, stars_c, mod_c)
            e2e_c += time.perf_counter() - t0
        ci = shc.engine.info()
        line["clustered_indices"] = {
            "value": W * nrep / (tot * 1e-3), "unit": UNIT, "ms_per_step": tot / nrep,
            "e2e_value": W * nrep / e2e_c, "entries": ci["n_entries"], "entries_raw": ci["n_entries_raw"], "setup_s": round(t_setup, 1),
            "note": "same workload with each star's sparse points clustered around a centre in grid-axis "
                    "space (BEAST-like): duplicates of a star's class codes merge, fewer entries to stream"}
        shc.close()
    if args.all_modes and world == 1:
        others = {}
        for m in ("factor", "table_unique", "table_grid", "elementwise"):
            if m == mode:
                continue
            mdl = MB_Model(copy.deepcopy(settings), devices=[local], mode=m)
            sh = ShardedLikelihood(mdl, stars, moddata, device=local, mode=m)
            for _ in range(5):
                flush_l2()
                sh.lnlike_device(d_phi, d_rows)
            torch.cuda.synchronize()
            tot = 0.0
            nrep = min(K, 50)
            for _ in range(nrep):
                flush_l2()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                sh.lnlike_device(d_phi, d_rows)
                b.record()
                torch.cuda.synchronize()
                tot += a.elapsed_time(b)
            sh.engine.set_profiling(True)
            flush_l2()
            sh.lnlike_device(d_phi, d_rows)
            torch.cuda.synchronize()
            others[m] = {"value": W * nrep / (tot * 1e-3), "ms_per_step": tot / nrep,
                         "kernel_ms": sh.engine.last_kernel_ms()}
            sh.close()
        line["other_modes"] = others
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    # ONE JSON line on stdout: libraries that write to file descriptor 1 themselves (NCCL prints its
    # version banner there) are sent to stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(real_stdout, "w")
    try:
        import contextlib

        with contextlib.redirect_stdout(out):
            if args.impl == "reference":
                run_reference(args)
            else:
                run_ours(args)
    finally:
        out.flush()


if __name__ == "__main__":
    main()
