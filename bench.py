#!/usr/bin/env python
"""
bench.py -- ensemble lnprob evaluations per second on BASELINE.json's headline workload
(config 3: 100,000 stars x 50 sparse points, 1,000,400-row grid, 64 walkers; docs-example model).

    python bench.py [--gpus N] [--steps K] [--warmup W]             # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [...]   # the reference's CPU algorithm

A "step" is one pass of the hot path over one batch of W = 64 walkers (one emcee proposal
round).  N > 1 is launched by torchrun, one rank per GPU: the 100,000 stars are sharded over
the ranks (grid replicated), the shard totals are added inside the likelihood kernel over NVLink
peer memory (or by one NCCL all-reduce of 2*W doubles) -> "strong" scaling of the fixed-size job
BASELINE.json names.  Rank 0 prints ONE JSON line.

Keys beyond the base contract:
  value        device-timed (CUDA events on the launching stream), phi resident in HBM; at N > 1 the
               ranks are aligned by a device-side rendezvous (a tiny NCCL all-reduce enqueued after the
               L2 flush, outside the timed events) so that the flush's skew is not billed to the step
  e2e          the same metric through the public API with HOST buffers: lnprob(phis, model,
               star_lnpgriddata, beast_moddata) -- prior box on the host, phi to the device,
               kernels, (shard reduction,) result rows back, exceptions checked
  roofline     dominant kernel: algorithmic bytes per launch / its CUDA-event duration vs the
               measured HBM copy bandwidth (MEASURED_PEAKS.json).  See DESIGN.md section 5 for why
               the factor-mode hot kernel is bounded by shared-memory bandwidth instead, and
               `onchip` for that ceiling.
  sustained    >= 2 s of back-to-back launches (L2 warm), clocks and power sampled under load
  cpu_baseline the oracle's numpy port (the reference's algorithm) on a bounded sample of the same
               workload, timed on this box's host cores; its values double as a full-size parity check
               of the GPU path (`parity`, >= 8 walkers; at N > 1 against the C oracle, plus p2p == nccl)
  secondary    (N = 8) BASELINE config 4 -- 1,000,000 stars x 50 points, 2,000,800-row grid, 128 walkers --
               sharded over the same ranks: device-timed, end to end, parity
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
    ap.add_argument("--split-inner", default="", help="FACTOR mode: dust parameters of the inner table, e.g. Av or Av,Rv")
    ap.add_argument("--walkers", type=int, default=0, help="override the config's walker count")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=24, help="walkers the CPU baseline / parity check evaluates")
    ap.add_argument("--no-flush", action="store_true", help="keep L2 warm between steps (not the headline)")
    ap.add_argument("--all-modes", action="store_true", help="also time the other engine modes")
    ap.add_argument("--no-clustered", action="store_true",
                    help="skip the extra measurement with clustered (BEAST-like) sparse indices")
    ap.add_argument("--no-sustained", action="store_true", help="skip the 2 s back-to-back block")
    ap.add_argument("--secondary", default="auto", choices=["auto", "cfg4", "cfg5", "both", "none"],
                    help="N > 1: also measure BASELINE config 4 (stars sharded over the ranks) and config 5 (pixel "
                         "batches, pixels split over the ranks); auto: both at N = 8")
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


def workload_config(args, cfg, G, W):
    """The `config` object -- byte-identical in both arms (the driver compares them)."""
    return {"workload": f"{args.workload}: {cfg['S']} stars x {cfg['K']} sparse points, {G}-row grid, "
                        f"{W} walkers, docs-example model",
            "index_mode": args.index_mode}


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
    G = len(moddata["Av"])
    sample = (f"each step = {per_step} of {W} walkers x all {cfg['S']} stars, one walker per core; "
              f"{steps} of the requested {args.steps} steps run to stay inside {budget_s:.0f} s")
    line = {
        "impl": "reference",
        "metric": metric_name(args, cfg, G, W), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm + 1, "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, cfg, G, W),
        "what": "oracle/lnprob_port.py loop=True: numpy restatement of the reference's lnprob (per-star "
                "Python loop, per-bin grid masks), walker-parallel over the host cores with multiprocessing",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": per_step, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons / power of one GPU through NVML while a tagged phase runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.err = index, [], False, None
        self.phase = None

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self.stop_flag:
                phase = self.phase
                if phase is not None:
                    try:
                        reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:
                        reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), reasons,
                                         nv.nvmlDeviceGetPowerUsage(h) / 1e3, phase))
                time.sleep(0.004)
        except Exception as exc:  # NVML missing: report, do not fail the bench
            self.err = repr(exc)

    def summary(self, phase=None):
        sel = [s for s in self.samples if phase is None or s[3] == phase]
        if self.err or not sel:
            return {"sm_mhz": None, "sm_max_mhz": getattr(self, "max_mhz", None), "reasons": [],
                    "note": self.err or "no samples"}
        import statistics

        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
                 0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
                 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        seen = 0
        for _, r, _, _ in sel:
            seen |= r
        reasons = [n for b, n in names.items() if seen & b and n != "gpu_idle"]
        return {"sm_mhz": statistics.median(s[0] for s in sel), "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(sel),
                "power_w_max": max(s[2] for s in sel), "power_w_median": statistics.median(s[2] for s in sel)}


_CPU = {}


def _cpu_eval(phi):
    from oracle import lnprob_port as port

    return port.lnprob(phi, _CPU["pm"], _CPU["stars"], _CPU["mod"], loop=True)


def cpu_baseline_and_values(settings, moddata, stars, phis, n):
    """The numpy port (reference algorithm, per-star Python loop) on n walkers, one per host core."""
    import copy
    import multiprocessing as mp

    import numpy as np

    from oracle import lnprob_port as port

    pm = port.PortModel(copy.deepcopy(settings))
    pm.massmultipliers = port.mass_multipliers(moddata, pm.physics_model["M_ini"])
    _CPU.update(pm=pm, stars=stars, mod=moddata)
    procs = max(1, min(host_cores(), n))
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        pool.map(_cpu_eval, [phis[0]] * procs)  # warm-up call in every worker
        t0 = time.perf_counter()
        vals = np.array(pool.map(_cpu_eval, list(phis[:n]), chunksize=1))
        dt = time.perf_counter() - t0
    return vals, dt, procs


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
    split = tuple(p for p in args.split_inner.split(",") if p) or None
    settings, moddata, stars, cfg, W = make_workload(args)
    model = MB_Model(copy.deepcopy(settings), devices=[local], mode=args.mode,
                     table_precision=args.table_precision, split_inner=split)
    phis = synth.make_walkers(model, W, seed=3)
    # CPU baseline first: its worker processes are forked before this process holds a CUDA context
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        # (a whole number of rounds of one walker per host core)
        n_cpu = max(1, min(W, -(-args.cpu_sample // host_cores()) * host_cores()))
        cpu = cpu_baseline_and_values(settings, moddata, stars, phis, n_cpu)
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

    t0 = time.perf_counter()
    sharded = ShardedLikelihood(model, stars, moddata, device=local, mode=args.mode, reduce=args.reduce)
    t_ingest = time.perf_counter() - t0
    info = sharded.engine.info()
    ndim = sharded.ndim
    d_phi = torch.from_numpy(np.ascontiguousarray(phis[:, sharded.perm])).to(dev)
    d_rows = torch.empty((4, W), dtype=torch.float64, device=dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    K = args.steps

    def flush_l2():
        if not args.no_flush:
            flush.zero_()

    def time_device(sh, rows, nsteps, nwarm):
        for _ in range(nwarm):
            flush_l2()
            sh.rendezvous()
            sh.lnlike_device(d_phi, rows)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
        barrier()
        t_w = time.perf_counter()
        for i in range(nsteps):
            flush_l2()
            sh.rendezvous()
            evs[i][0].record()
            sh.lnlike_device(d_phi, rows)
            evs[i][1].record()
        barrier()
        t_w = time.perf_counter() - t_w
        return max_over_ranks(sum(a.elapsed_time(b) for a, b in evs)), t_w

    def time_e2e(call, nsteps, nwarm, sh=None):
        for _ in range(nwarm):
            flush_l2()
            out = call()
        barrier()
        tot = 0.0
        for i in range(nsteps):
            flush_l2()
            # N > 1: a sampler's ranks leave step n together (the in-kernel exchange releases them within a
            # microsecond); the flush must not de-align them, so they meet on the device before the clock starts
            if sh is not None and world > 1:
                sh.rendezvous()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = call()
            tot += time.perf_counter() - t0
        barrier()
        return max_over_ranks(tot), out

    # ---- device-timed value ------------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:   # (one NVML poller per job: the ranks share the host's cores)
        sampler.start()
    sampler.phase = "timed"
    if world > 1:
        # every rank's MAIN thread on its own host core where there are enough of them (the helper threads of
        # CUDA / NCCL exist by now and keep their masks): the end-to-end path is a chain of microsecond-scale
        # host steps on every rank, and a step ends when the slowest rank's shard arrives
        try:
            cores = sorted(os.sched_getaffinity(0))
            if len(cores) >= world:
                os.sched_setaffinity(0, {cores[local * (len(cores) // world)]})
        except (AttributeError, OSError):
            pass
    total_ms, t_wall = time_device(sharded, d_rows, K, max(3, args.warmup))
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
    e2e_s, e2e_vals = time_e2e(e2e_call, K, max(3, args.warmup), sharded)
    sampler.phase = None

    # ---- sustained: back-to-back launches for >= 2 s, clocks and power sampled under load ---------------
    sustained = None
    if not args.no_sustained:
        barrier()
        step_s = total_ms * 1e-3 / K
        n_sus = int(min(200000, max(200, 2.2 / max(step_s * 0.8, 1e-6))))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(20):
            sharded.lnlike_device(d_phi, d_rows)
        barrier()
        sampler.phase = "sustained"
        a.record()
        for _ in range(n_sus):
            sharded.lnlike_device(d_phi, d_rows)
        b.record()
        barrier()
        sampler.phase = None
        sus_ms = max_over_ranks(a.elapsed_time(b))
        sustained = {"value": W * n_sus / (sus_ms * 1e-3), "unit": UNIT, "seconds": sus_ms * 1e-3, "steps": n_sus,
                     "ms_per_step": sus_ms / n_sus, "l2": "warm (no flush between steps: launches back to back)",
                     "clocks": sampler.summary("sustained")}

    # ---- weak scaling (N > 1): every rank holds a FULL 100k-star shard of an N x 100k-star job --------
    weak = None
    if world > 1:
        shw = ShardedLikelihood(MB_Model(copy.deepcopy(settings), devices=[local], mode=args.mode), stars, moddata,
                                device=local, mode=args.mode, reduce=args.reduce,
                                star_range=(0, cfg["S"]), n_stars_total=world * cfg["S"])
        d_rows_w = torch.empty((4, W), dtype=torch.float64, device=dev)
        weak_ms, _ = time_device(shw, d_rows_w, K, max(3, args.warmup))
        # every rank holds the same stars: the job's star sum must be N x the single-copy one
        ratio = ((d_rows_w[0] + d_rows_w[1]) / (world * (d_rows[0] + d_rows[1]))).cpu().numpy()
        weak = {"stars_per_gpu": cfg["S"], "stars_total": world * cfg["S"],
                "value": world * W * K / (weak_ms * 1e-3), "unit": "100k-star-equivalent evals/s",
                "job_evals_per_s": W * K / (weak_ms * 1e-3), "ms_per_step": weak_ms / K,
                "reduce": shw.reduce,
                "check_starsum_vs_n_copies_max_rel_err": float(np.max(np.abs(ratio - 1.0))),
                "note": "per-GPU work fixed (each rank a full config-3 star shard), one reduction of "
                        f"{3 * W} partials per step; value counts an evaluation over N x 100k stars as N evaluations"}
        shw.close()

    # ---- per-kernel times (CUDA events inside the library, same stream) ------------------------------
    # N > 1: the ranks are aligned before every profiled launch, so a kernel's time holds its own work
    # plus the wait for the slowest peer of THAT launch, never the skew of an earlier step
    eng = sharded.engine
    eng.set_profiling(True)
    kms = {}
    nprof = min(K, 50)
    phases = None
    for _ in range(nprof):
        flush_l2()
        barrier()
        sharded.rendezvous()
        sharded.lnlike_device(d_phi, d_rows)
        torch.cuda.synchronize()
        for k, v in eng.last_kernel_ms().items():
            kms[k] = kms.get(k, 0.0) + v / nprof
        tl = eng.last_k2_timeline()
        if len(tl):
            ph = np.array([tl[:, k].max() for k in range(1, 8)])
            phases = ph / nprof if phases is None else phases + ph / nprof
    eng.set_profiling(False)
    if world > 1:
        t = torch.tensor([kms[k] for k in sorted(kms)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kms = dict(zip(sorted(kms), [float(x) for x in t.cpu()]))
    dom = max(kms, key=kms.get)

    # ---- N > 1: parity of the reduced values, and p2p == nccl bit for bit --------------------------------
    parity = None
    if world > 1:
        n = max(1, min(args.cpu_sample, W))
        other = "nccl" if sharded.reduce == "p2p" else "p2p"
        bitwise = None
        try:
            sh2 = ShardedLikelihood(MB_Model(copy.deepcopy(settings), devices=[local], mode=args.mode), stars, moddata,
                                    device=local, mode=args.mode, reduce=other)
            rows2 = torch.empty((4, W), dtype=torch.float64, device=dev)
            sh2.lnlike_device(d_phi, rows2)
            torch.cuda.synchronize()
            bitwise = bool(torch.equal(rows2[0] + rows2[1], d_rows[0] + d_rows[1]) and torch.equal(rows2[2:], d_rows[2:]))
            sh2.close()
        except Exception as exc:  # e.g. no peer access on this box: say so, do not fail the bench
            bitwise = f"unavailable: {exc!r}"
        if rank == 0:
            from oracle import c_oracle
            from oracle import lnprob_port as port

            pm = port.PortModel(copy.deepcopy(settings))
            t0 = time.perf_counter()
            want = c_oracle.lnprob_batch(pm, phis[:n], stars, moddata)
            t_or = time.perf_counter() - t0
            rel = np.abs(vals_dev[:n] - want) / np.abs(want)
            rel2 = np.abs(np.asarray(e2e_vals)[:n] - want) / np.abs(want)
            parity = {"walkers_checked": n, "against": "oracle/mb_oracle.c (reference algorithm, all stars, "
                      f"OpenMP over walkers; {t_or:.1f} s)",
                      "max_rel_err_device_path": float(rel.max()), "max_rel_err_e2e_path": float(rel2.max()),
                      "tolerance": 1e-9, "status_flags": int(status_dev.max()),
                      f"{sharded.reduce}_equals_{other}_bitwise": bitwise}
        barrier()

    # ---- secondary (N = 8): BASELINE config 4, sharded over the same ranks ---------------------------------
    secondary = secondary5 = None
    auto_secondary = args.secondary == "auto" and world == 8 and args.workload == "cfg3"
    if world > 1 and (args.secondary in ("cfg4", "both") or auto_secondary):
        secondary = run_secondary_cfg4(args, world, rank, local, dev, barrier, max_over_ranks, flush_l2)
    if world > 1 and (args.secondary in ("cfg5", "both") or auto_secondary):
        secondary5 = run_secondary_cfg5(args, world, rank, local, dev, barrier, max_over_ranks, flush_l2)

    if rank != 0:
        sampler.stop_flag = True
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel --------------------------------------------------------------
    peak, peak_src = measured_peaks()
    Wpad = 32 * (1 if W <= 32 else 2 if W <= 64 else 4)
    n_ent = info["n_entries"]
    nA, nB = info["n_age_classes"], info["n_dust_classes"]
    nO, nI = info["n_outer_rows"], info["n_inner_rows"]
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
        "k2_stars": (8 + info["code_bytes"]) * n_ent + 16 * info["n_bins"] * nB + (8 * (nO + nI) * Wpad if mode == "factor"
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
                        "K1e is bound by the fp64 exp of the lognormal priors, not by HBM"
                        if (mode == "elementwise" and dom == "k1_expand") else "HBM-bound: materialised table written / gathered"}
    clocks = sampler.summary("timed")
    sampler.stop_flag = True
    sm_mhz = clocks.get("sm_mhz") or 0
    onchip = None
    if mode == "factor" and sm_mhz:
        # shared-memory wavefronts K2 must issue (measured in lds_mix.cu): per
        # entry 2*NW for the Wpad*8-byte inner-factor row + 1.5 (1.25 with 16-bit codes) for the broadcast
        # of weight and code, per run 2*NW for the outer-factor row
        nw = Wpad // 32
        bcast = 1.5 if info["code_bytes"] == 4 else 1.25  # four codes: one 16-byte or one 8-byte broadcast
        wavefronts = n_ent * (2.0 * nw + bcast) + info["n_runs"] * 2.0 * nw
        peak_wf = info["sm_count"] * sm_mhz * 1e6 / 1e9
        ach = wavefronts / (kms["k2_stars"] * 1e-3) / 1e9
        onchip = {"bound": "smem", "wavefronts_per_launch": wavefronts, "achieved": ach, "peak": peak_wf,
                  "unit": "Gwavefront/s", "frac": ach / peak_wf,
                  "note": f"model count ({2 * nw} inner-factor + {bcast} broadcast wavefronts per entry, {2 * nw} "
                          f"outer-factor wavefronts per run); peak = 1 wavefront/clk/SM x {info['sm_count']} SMs x "
                          f"{sm_mhz:.0f} MHz; ncu's measured l1tex data-pipe utilisation for the same kernel is in profiles/"}
    # the north_star's own yardstick: the materialised-physmod design's compulsory bytes per step
    G = len(moddata["Av"])
    raw = info["n_entries_raw"]
    survey_bytes = G * (8 * nfree + 18) + 8 * G * W + 12 * raw + 8 * W * min(G, raw) + 8 * W
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
        "config": workload_config(args, cfg, G, W),
        "engine": {
            "engine_mode": mode, "n_free_parameters": ndim,
            "sharding": f"stars split over {world} rank(s), grid replicated; shard reduction: "
                        + {"none": "none (1 GPU)", "nccl": f"one NCCL all-reduce of {3 * W} doubles per step",
                           "p2p": "fused into K2's tail over NVLink peer memory (no collective launch)"}[sharded.reduce],
            "l2": "kept warm" if args.no_flush else f"flushed between steps ({L2_FLUSH_BYTES >> 20} MiB memset, outside the timed events)",
            "rank_alignment": "none (1 GPU)" if world == 1 else
                              "device-side rendezvous (mb_peer_rendezvous: flags over NVLink peer memory; a 4-byte NCCL all-reduce "
                              "when the shards are reduced by NCCL) enqueued after the flush, before the start event",
            "entries": n_ent, "entries_raw": raw, "runs": info["n_runs"], "age_classes": nA, "dust_classes": nB,
            "outer_rows": nO, "inner_rows": nI, "split_inner_mask": info["split_inner_mask"],
            "code_bytes": info["code_bytes"], "bins": info["n_bins"], "ingest_s": round(sharded.t_create, 3),
            "peer_connect_s": round(sharded.t_connect, 3), "setup_s": round(t_ingest, 3),
        },
        "clocks": clocks,
        "e2e": {"value": W * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * W * ndim,
                "d2h_bytes_per_step": 8 * 4 * W, "ms_per_step": 1e3 * e2e_s / K,
                "api": "megabeast_b200.lnprob(phis[W,ndim], MB_Model, star_lnpgriddata, beast_moddata)"
                       if world == 1 else "megabeast_b200.parallel.ShardedLikelihood.lnprob(phis[W,ndim])"},
        "gpu_launches": launches,
        "kernel_ms": kms,
        "roofline": roofline,
        "onchip": onchip,
        "north_star_equivalent": survey,
        "wall_ms_per_step_incl_flush": 1e3 * t_wall / K,
    }
    if phases is not None:
        line["k2_phases_us"] = dict(zip(("tables_built", "nstars_bins_done", "star_loop_done", "block_reduced",
                                         "last_block_totals_poisson_sent", "peers_received", "end"),
                                        [round(float(x), 2) for x in phases]))
        line["k2_phases_us"]["note"] = ("%globaltimer stamps inside K2, microseconds since its first block started, "
                                        "max over blocks, mean over the profiled launches (rank 0)")
    if sustained is not None:
        line["sustained"] = sustained
    if weak is not None:
        line["weak_scaling"] = weak
    if parity is not None:
        line["parity"] = parity
    if secondary is not None:
        line["secondary"] = secondary
    if secondary5 is not None:
        line["secondary_cfg5"] = secondary5

    # ---- CPU baseline + full-size parity ----------------------------------------------------------------
    if cpu is not None:
        n = n_cpu
        cpu_vals, dt, procs = cpu
        line["cpu_baseline"] = {
            "value": n / dt, "unit": UNIT, "cores": procs, "kind": "port",
            "sample": f"{n} of {W} walkers x all {cfg['S']} stars, one walker per process on {procs} host cores; "
                      f"oracle/lnprob_port.py loop=True (numpy restatement of the reference's lnprob, per-star "
                      f"Python loop); {dt:.1f} s"}
        rel = np.abs(vals_dev[:n] - cpu_vals) / np.abs(cpu_vals)
        rel2 = np.abs(np.asarray(e2e_vals)[:n] - cpu_vals) / np.abs(cpu_vals)
        line["parity"] = {"walkers_checked": n, "against": "oracle/lnprob_port.py loop=True (numpy port of the reference)",
                          "max_rel_err_device_path": float(rel.max()),
                          "max_rel_err_e2e_path": float(rel2.max()), "tolerance": 1e-9,
                          "status_flags": int(status_dev.max())}
    if world == 1 and args.index_mode == "uniform" and not args.no_clustered:
        # the headline uses uniform-random sparse indices (worst case: nothing to merge, no locality);
        # real BEAST lnp files hold the K best models around each star's best fit -- same job, same
        # kernels, with the top-K points clustered in grid-axis space
        t0 = time.perf_counter()
        _, mod_c, stars_c, _ = synth.make_problem(args.workload, mode="clustered")
        mdl = MB_Model(copy.deepcopy(settings), devices=[local], mode=args.mode, table_precision=args.table_precision,
                       split_inner=split)
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
            "e2e_value": W * nrep / e2e_c, "entries": ci["n_entries"], "entries_raw": ci["n_entries_raw"],
            "runs": ci["n_runs"], "setup_s": round(t_setup, 1),
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


def run_secondary_cfg4(args, world, rank, local, dev, barrier, max_over_ranks, flush_l2):
    """BASELINE config 4 (PHAT scale: 1,000,000 stars x 50 points, 2,000,800-row grid, 128 walkers), star-sharded
    over the ranks.  Every rank synthesises only ITS shard of stars (seed = 2 + 1000 * rank; the grid, seed 1,
    is replicated), so the job is the concatenation of the shards.  Parity: every rank runs the C oracle on its
    own shard (sum of ln L over its stars), the shard sums are added, rank 0 adds the N-stars term of the numpy port."""
    import copy

    import numpy as np
    import torch
    import torch.distributed as dist

    from megabeast_b200 import MB_Model, synth
    from megabeast_b200.engine import shard_bounds
    from megabeast_b200.parallel import ShardedLikelihood, combine_rows
    from oracle import c_oracle
    from oracle import lnprob_port as port

    c4 = synth.CONFIGS["cfg4"]
    S, K4, W4 = c4["S"], c4["K"], c4["W"]
    t0 = time.perf_counter()
    moddata = synth.make_grid(c4["axes"], seed=1)
    lo, hi = shard_bounds(S, world, rank)
    stars = synth.make_star_data(moddata, c4["axes"], hi - lo, K4, mode=args.index_mode, seed=2 + 1000 * rank)
    settings = synth.docs_example_settings()
    model = MB_Model(copy.deepcopy(settings), devices=[local], mode=args.mode)
    phis = synth.make_walkers(model, W4, seed=3)
    t_synth = time.perf_counter() - t0
    t0 = time.perf_counter()
    sh = ShardedLikelihood(model, stars, moddata, device=local, mode=args.mode, reduce=args.reduce,
                           star_range=(0, hi - lo), n_stars_total=S)
    t_ingest = time.perf_counter() - t0
    d_phi = torch.from_numpy(np.ascontiguousarray(phis[:, sh.perm])).to(dev)
    d_rows = torch.empty((4, W4), dtype=torch.float64, device=dev)
    nsteps, nwarm = max(10, min(args.steps, 50)), 5
    for _ in range(nwarm):
        flush_l2()
        sh.rendezvous()
        sh.lnlike_device(d_phi, d_rows)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
    barrier()
    for i in range(nsteps):
        flush_l2()
        sh.rendezvous()
        evs[i][0].record()
        sh.lnlike_device(d_phi, d_rows)
        evs[i][1].record()
    barrier()
    dev_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))
    vals_dev, status_dev = combine_rows(d_rows.cpu().numpy())
    for _ in range(nwarm):
        flush_l2()
        vals_e2e = sh.lnprob(phis)
    e2e = 0.0
    for _ in range(nsteps):
        flush_l2()
        sh.rendezvous()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        vals_e2e = sh.lnprob(phis)
        e2e += time.perf_counter() - t0
    barrier()
    e2e = max_over_ranks(e2e)
    info = sh.engine.info()
    # parity on n walkers: shard star sums from the C oracle (no N-stars term), added over the ranks
    n = max(4, min(args.cpu_sample, W4))
    pm = port.PortModel(copy.deepcopy(settings))
    pm.compute_N_stars = False
    t0 = time.perf_counter()
    part, st = c_oracle.lnlike_batch(pm, phis[:n], stars, moddata)
    t_or = time.perf_counter() - t0
    t = torch.from_numpy(np.ascontiguousarray(part)).to(dev)
    dist.all_reduce(t)
    starsum = t.cpu().numpy()
    out = None
    if rank == 0:
        from scipy.special import gammaln

        pm2 = port.PortModel(copy.deepcopy(settings))
        mm = port.mass_multipliers(moddata, pm2.physics_model["M_ini"])
        want = np.empty(n)
        for i in range(n):
            physmod, mods = pm2.physmod(phis[i], moddata)
            pred = port.predicted_num_stars(mm, moddata, physmod, mods["logA"], loop=False)
            want[i] = S * np.log(pred) - pred - gammaln(S + 1) + starsum[i]
        rel = np.abs(vals_dev[:n] - want) / np.abs(want)
        rel2 = np.abs(np.asarray(vals_e2e)[:n] - want) / np.abs(want)
        out = {
            "workload": f"cfg4: {S} stars x {K4} sparse points, {len(moddata['Av'])}-row grid, {W4} walkers, "
                        f"docs-example model; stars sharded over {world} ranks ({hi - lo} on rank 0), grid replicated",
            "index_mode": args.index_mode,
            "metric": f"ensemble lnprob evals/sec ({S} stars x {len(moddata['Av'])} grid, {W4} walkers)",
            "value": W4 * nsteps / (dev_ms * 1e-3), "unit": UNIT, "ms_per_step": dev_ms / nsteps, "steps": nsteps,
            "scaling": "strong",
            "e2e": {"value": W4 * nsteps / e2e, "unit": UNIT, "ms_per_step": 1e3 * e2e / nsteps,
                    "h2d_bytes_per_step": 8 * W4 * sh.ndim, "d2h_bytes_per_step": 8 * 4 * W4},
            "engine": {"engine_mode": info["mode"], "reduce": sh.reduce, "entries_rank0": info["n_entries"],
                       "dust_classes": info["n_dust_classes"], "walkers_per_pass": info["walkers_per_pass"],
                       "launches_per_step": info["launches_last_eval"], "synth_s": round(t_synth, 1),
                       "ingest_s": round(sh.t_create, 3), "peer_connect_s": round(sh.t_connect, 3)},
            "parity": {"walkers_checked": n, "max_rel_err_device_path": float(rel.max()),
                       "max_rel_err_e2e_path": float(rel2.max()), "tolerance": 1e-9,
                       "status_flags": int(status_dev.max()),
                       "against": "oracle/mb_oracle.c on every rank's own shard (sum of ln L), added over the ranks, plus "
                                  f"the numpy port's N-stars term on rank 0 ({t_or:.1f} s per rank)"},
            "n1_reference": "profiles/r2_bench_cfg4_n1.json (same job on 1 GPU, builder-run)",
        }
    sh.close()
    barrier()
    return out


def run_secondary_cfg5(args, world, rank, local, dev, barrier, max_over_ranks, flush_l2):
    """BASELINE config 5 (batched independent spatial-pixel fits: 4,096 pixels x ~500 stars each over a shared
    1M-row grid, all walkers of all pixels in one launch across the GPUs): every rank owns 4,096 / N pixels
    (its own seeds; the grid, seed 1, is shared), no collective on the data path -- pixels are independent.
    Device-timed (CUDA events per launch, max over ranks), end to end through PixelEnsemble.lnprob with
    host arrays, parity of 8 pixels per rank against the per-pixel oracle."""
    import copy

    import numpy as np
    import torch

    from megabeast_b200 import synth
    from megabeast_b200.pixels import PixelEnsemble
    from oracle import c_oracle
    from oracle import lnprob_port as port

    P = 4096 // world
    axes = synth.CONFIGS["cfg3"]["axes"]
    t0 = time.perf_counter()
    settings, moddata, stars, offs = synth.make_pixel_problem(
        n_pixels=P, mean_stars=500, axes=axes, K=50, mode=args.index_mode, seeds=(1, 2 + 10 * rank, 4 + 10 * rank))
    t_synth = time.perf_counter() - t0
    t0 = time.perf_counter()
    pe = PixelEnsemble(copy.deepcopy(settings), stars, moddata, offs, device=local)
    t_ingest = time.perf_counter() - t0
    W, ndim = 45, pe.ndim
    rng = np.random.default_rng(3)
    start = np.asarray(pe.model.start_params()[1], dtype=float)
    phis = np.clip(start * (1.0 + 0.05 * rng.standard_normal((P, W, ndim))), *pe.model._box())
    d_phi = torch.from_numpy(np.ascontiguousarray(phis)).to(dev)
    d_out = torch.empty((P, 4, W), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    nsteps = max(10, min(args.steps, 30))
    for _ in range(5):
        flush_l2()
        pe.engine.lnlike_pixels_device(d_phi.data_ptr(), W, ndim, d_out.data_ptr(), stream)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
    barrier()
    for i in range(nsteps):
        flush_l2()
        evs[i][0].record()
        pe.engine.lnlike_pixels_device(d_phi.data_ptr(), W, ndim, d_out.data_ptr(), stream)
        evs[i][1].record()
    barrier()
    dev_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))
    vals = pe.lnprob(phis)
    barrier()
    e2e = 0.0
    for _ in range(5):
        flush_l2()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        vals = pe.lnprob(phis)
        e2e += time.perf_counter() - t0
    barrier()
    e2e = max_over_ranks(e2e) / 5
    err = 0.0
    ncheck = 8
    for p in range(ncheck):
        sub = {"indxs": np.ascontiguousarray(stars["indxs"][:, offs[p]:offs[p + 1]]),
               "vals": np.ascontiguousarray(stars["vals"][:, offs[p]:offs[p + 1]])}
        want = c_oracle.lnprob_batch(port.PortModel(copy.deepcopy(settings)), phis[p], sub, moddata)
        err = max(err, float(np.max(np.abs(vals[p] - want) / np.abs(want))))
    err = max_over_ranks(err)
    info = pe.engine.info()
    pe.close()
    barrier()
    if rank != 0:
        return None
    step_ms = dev_ms / nsteps
    return {
        "workload": f"cfg5: {P * world} pixels x ~500 stars x 50 sparse points, shared {info['G']}-row grid, {W} walkers per "
                    f"pixel, docs-example model; {P} pixels per rank, all walkers of all pixels in one launch per rank",
        "index_mode": args.index_mode,
        "metric": "pixel-ensemble lnprob evals/sec (one eval = one walker of one pixel over the pixel's stars)",
        "value": P * world * W / (step_ms * 1e-3), "unit": UNIT, "ms_per_step": step_ms, "steps": nsteps,
        "scaling": "weak", "pixel_walkers_per_step": P * world * W,
        "e2e": {"value": P * world * W / e2e, "unit": UNIT, "ms_per_step": 1e3 * e2e,
                "h2d_bytes_per_step": 8 * P * W * ndim, "d2h_bytes_per_step": 8 * 4 * P * W,
                "api": "megabeast_b200.pixels.PixelEnsemble.lnprob(phis[P,W,ndim]) with host arrays, per rank"},
        "engine": {"entries_rank0": info["n_entries"], "stars_rank0": int(offs[-1]), "synth_s": round(t_synth, 1),
                   "ingest_s": round(t_ingest, 2)},
        "parity": {"pixels_checked_per_rank": ncheck, "walkers_per_pixel": W, "max_rel_err": err, "tolerance": 1e-9,
                   "against": "oracle/mb_oracle.c on each checked pixel's stars alone (its own star count in the Poisson term); "
                              "max over the ranks"},
    }


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
