"""
Star-sharded ensemble likelihood across processes: one process per GPU (``torchrun``), every
rank owns a contiguous shard of stars plus a replicated grid, evaluates all W walkers on its
shard, and ONE reduction of 3*W values (partial sums of ln(star prob) in two exact words + non-finite counts) per
step completes the likelihood (BASELINE.json north_star item 4).  The
reduction runs either as one NCCL all-reduce after the kernels ("nccl") or -- the default when
the GPUs can map each other's memory -- INSIDE the likelihood kernel: the last block of every
rank's K2 stores its shard totals into all peers over NVLink and adds what the peers stored
("p2p", see mb_peer_connect in include/megabeast_b200.h).

torch is used for what it is good at here -- device buffers, the current stream and
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests); the likelihood itself is the C ABI.
"""
import numpy as np

from . import _cabi
from .engine import LikelihoodEngine, shard_bounds

__all__ = ["dist_info", "allreduce_rows", "reduce_host_rows", "combine_rows", "ShardedLikelihood"]


def dist_info(group=None):
    """(rank, world_size) of the default process group, (0, 1) when there is none."""
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(group), dist.get_world_size(group)
    except ImportError:
        pass
    return 0, 1


def allreduce_rows(rows, group=None):
    """Sum the shard-additive rows (STARSUM, STARSUM_LO, NONFINITE = the first 3*W values) of a
    [4][W] tensor over all ranks, in place; the POISSON row is identical everywhere and is left
    alone.  The two STARSUM words are exact multiples of 2^-24 and 2^-68, so this floating-point sum
    is exact.  Works for CUDA tensors (NCCL) and CPU tensors (gloo)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        W = rows.shape[1]
        dist.all_reduce(rows.view(-1)[: _cabi.MB_OUT_ADDITIVE_ROWS * W], op=dist.ReduceOp.SUM, group=group)
    return rows


def reduce_host_rows(rows, group=None, device=None):
    """Sum host rows [4][W] (numpy) over the ranks: STARSUM, STARSUM_LO and NONFINITE add, POISSON is shared.
    gloo groups reduce the CPU tensor directly, NCCL groups go through ``device``."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rows
    t = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float64))
    if dist.get_backend(group) == "nccl":
        t = t.to(device)
    allreduce_rows(t, group)
    return t.cpu().numpy()


def combine_rows(rows):
    """(lnlike[W], status[W]) from reduced rows -- the arithmetic of ``mb_combine``."""
    rows = np.asarray(rows, dtype=np.float64)
    total = rows[3] + (rows[0] + rows[1])  # POISSON + (STARSUM + STARSUM_LO), as mb_combine
    status = np.where(rows[2] > 0.0, _cabi.STATUS_NONFINITE_STAR, 0).astype(np.int32)
    status |= np.where(rows[2] < 0.0, _cabi.STATUS_PEER_TIMEOUT, 0).astype(np.int32)
    status |= np.where(total == 0.0, _cabi.STATUS_TOTAL_ZERO, 0).astype(np.int32)
    return total, status


class ShardedLikelihood:
    """
    Per-rank evaluator.  ``model`` is an ``MB_Model``; the data dictionaries are the full ones
    (every rank holds them, only its shard goes to its GPU).

    ``lnlike_device(d_phi)``  phi already on the device -> reduced rows on the device, all
                              asynchronous on torch's current stream (K0..K3 + one all-reduce)
    ``lnprob(phis_host)``     the user-facing call: prior box on the host, pinned H2D of phi,
                              kernels, all-reduce, D2H of the rows, exceptions as the reference
    """

    def __init__(self, model, star_lnpgriddata, beast_moddata, device=None, mode="auto", group=None,
                 reduce="auto", star_range=None, n_stars_total=None, check_sync=False):
        """``check_sync``: every ``lnprob`` call first verifies (one small all-reduce) that all ranks
        were handed the same phi.  The fused NVLink reduction needs every rank to issue the same
        sequence of launches -- identical phi, hence identical in-box compaction -- and a rank that
        diverges would leave the others waiting for its shard until the peer timeout.  Off by default
        (it costs a collective per call); switch it on while wiring up a new driver."""
        import torch

        from .helpers import precompute_mass_multipliers

        self.torch = torch
        self.model = model
        self.group = group
        self.rank, self.world = dist_info(group)
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.spec, self.perm, self.ndim = model._spec()
        mm = None
        if model.compute_N_stars:
            model.massmultipliers = precompute_mass_multipliers(
                beast_moddata, model.physics_model["M_ini"]["model"])
            model.compute_massmult = False
            mm = model.massmultipliers
        S = np.asarray(star_lnpgriddata["indxs"]).shape[1]
        # default: the arrays hold the whole job and this rank takes its contiguous share; a rank
        # whose arrays hold only its own stars passes star_range=(0, S_local) and the job's total
        self.star_range = shard_bounds(S, self.world, self.rank) if star_range is None else tuple(star_range)
        import time as _time

        t0 = _time.perf_counter()
        self.engine = LikelihoodEngine(self.spec, star_lnpgriddata, beast_moddata, massmult=mm,
                                       devices=[self.device], mode=mode,
                                       star_range=self.star_range,
                                       n_stars_total=S if n_stars_total is None else int(n_stars_total),
                                       table_precision=getattr(model, "table_precision", "fp64"),
                                       split_inner=getattr(model, "split_inner", None))
        # how the shard partials are added: "p2p" = inside the likelihood kernel over NVLink peer
        # memory (mb_peer_connect), "nccl" = one all-reduce after it; "auto" tries p2p first
        self.t_create = _time.perf_counter() - t0   # ingest: class coding + entry build + uploads
        t0 = _time.perf_counter()
        self.reduce = "none" if self.world == 1 else self._setup_reduce(reduce)
        self.t_connect = _time.perf_counter() - t0  # peer mapping (CUDA IPC + peer access), once per process group
        self.check_sync = bool(check_sync)
        self._cap = 0
        self._ensure(64)
        # device variable order == reference order (the usual case): lnprob can be one C call
        self._fast_ok = (not self.spec.tabulated and self.perm.size == self.ndim
                         and np.array_equal(self.perm, np.arange(self.ndim)))

    def _setup_reduce(self, want):
        import torch.distributed as dist

        if want not in ("auto", "p2p", "nccl"):
            raise ValueError("reduce must be 'auto', 'p2p' or 'nccl'")
        if want == "nccl":
            return "nccl"
        t = self.torch
        ok = 1
        try:
            tokens = [None] * self.world
            dist.all_gather_object(tokens, self.engine.peer_export(), group=self.group)
            self.engine.peer_connect(self.rank, self.world, tokens)
        except Exception as exc:  # no peer access / IPC: every rank must fall back together
            ok, self._p2p_error = 0, repr(exc)
        flag = t.tensor([ok], dtype=t.int32, device=t.device("cuda", self.device))
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 1:
            return "p2p"
        if ok:
            self.engine.peer_disconnect()
        if want == "p2p":
            raise RuntimeError("peer-memory reduction unavailable: " + getattr(self, "_p2p_error", "a peer failed"))
        return "nccl"

    def _ensure(self, W):
        if W <= self._cap:
            return
        t = self.torch
        dev = t.device("cuda", self.device)
        self._cap = W
        self.d_phi = t.empty((W, max(1, self.ndim)), dtype=t.float64, device=dev)
        self.d_rows = t.empty((_cabi.MB_OUT_ROWS, W), dtype=t.float64, device=dev)
        self.h_phi = t.empty((W, max(1, self.ndim)), dtype=t.float64).pin_memory()
        self.h_rows = t.empty((_cabi.MB_OUT_ROWS, W), dtype=t.float64).pin_memory()

    def lnlike_device(self, d_phi, d_rows=None):
        t = self.torch
        W = d_phi.shape[0]
        if d_rows is None:
            self._ensure(W)
            d_rows = self.d_rows if W == self._cap else self.d_rows.view(-1)[: 4 * W].view(4, W)
        stream = t.cuda.current_stream(self.device).cuda_stream
        self.engine.lnlike_device(d_phi.data_ptr(), W, self.ndim, d_rows.data_ptr(), stream)
        if self.reduce == "nccl":
            allreduce_rows(d_rows, self.group)
        return d_rows  # "p2p": the kernel already summed the shards over NVLink

    def rendezvous(self):
        """Align the ranks on the current stream ahead of a timed launch: over the NVLink exchange
        buffers when the shards are reduced there, else with a tiny NCCL all-reduce."""
        if self.world == 1:
            return
        t = self.torch
        if self.reduce == "p2p":
            self.engine.peer_rendezvous(t.cuda.current_stream(self.device).cuda_stream)
        else:
            import torch.distributed as dist

            if not hasattr(self, "_tick"):
                self._tick = t.zeros(1, dtype=t.float32, device=t.device("cuda", self.device))
            dist.all_reduce(self._tick, group=self.group)

    def lnlike_rows(self, phis, tabs=None):
        """Host phi (W, ndim in device order) -> reduced host rows [4][W].  ``tabs``: host-evaluated
        tables of tabulated prior models (every rank evaluates the same callable on the same class
        values, so all ranks pass identical tables)."""
        if self.reduce in ("p2p", "none"):
            # one C call: phi and the rows go through mapped pinned memory, the kernels run on the
            # handle's own stream and (p2p) add the shards over NVLink themselves -- no torch ops
            return self.engine.lnlike_rows(phis, tabs)
        if tabs:
            # tabulated models go through the host form of the library (it uploads the tables); the
            # shard partials are then added by the process group
            rows = self.engine.lnlike_rows(phis, tabs)
            return reduce_host_rows(rows, self.group, self.torch.device("cuda", self.device))
        t = self.torch
        phis = np.ascontiguousarray(phis, dtype=np.float64)
        W = phis.shape[0]
        self._ensure(W)
        h_phi = self.h_phi[:W]
        h_phi.numpy()[...] = phis.reshape(W, -1) if self.ndim else 0.0
        d_phi = self.d_phi[:W]
        d_phi.copy_(h_phi, non_blocking=True)
        d_rows = self.d_rows.view(-1)[: 4 * W].view(4, W)
        self.lnlike_device(d_phi, d_rows)
        h_rows = self.h_rows.view(-1)[: 4 * W].view(4, W)
        h_rows.copy_(d_rows, non_blocking=True)
        t.cuda.current_stream(self.device).synchronize()
        return h_rows.numpy()

    def _assert_same_phi(self, phis):
        import torch.distributed as dist

        t = self.torch
        # (sum, sum of squares, count) of phi: min and max over the ranks must agree
        a = np.ascontiguousarray(phis, dtype=np.float64)
        sig = t.tensor([float(a.sum()), float((a * a).sum()), float(a.size)], dtype=t.float64)
        if dist.get_backend(self.group) == "nccl":
            sig = sig.to(t.device("cuda", self.device))
        lo, hi = sig.clone(), sig.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
        if not t.equal(lo, hi):
            raise RuntimeError("ShardedLikelihood.lnprob: the ranks were called with different phi")

    def lnprob(self, phi):
        phi = np.asarray(phi, dtype=np.float64)
        single = phi.ndim == 1
        phis = phi[None, :] if single else phi
        if self.check_sync and self.world > 1:
            self._assert_same_phi(phis)
        if self.reduce in ("p2p", "none") and self._fast_ok and phis.shape[1] == self.ndim:
            # one C call per rank (mb_lnprob_box): prior box, likelihood of the in-box walkers on
            # this rank's shard, shard totals added inside the kernel over NVLink, combine.  Every
            # rank holds the same phi, so every rank launches (or skips) the same kernels.
            lo, hi = self.model._box()
            vals, status, _ = self.engine.lnprob_box(phis, lo, hi)
            out = vals.copy()
        else:
            lp = np.asarray(self.model.lnprior(phis), dtype=np.float64)
            out = np.full(len(phis), -np.inf)
            inside = np.isfinite(lp)
            status = np.zeros(1, dtype=np.int32)
            if inside.any():
                tabs = self.model._tables(self.engine, self.spec, phis[inside]) if self.spec.tabulated else None
                rows = self.lnlike_rows(phis[inside][:, self.perm], tabs)
                vals, status = combine_rows(rows)
                out[inside] = lp[inside] + vals
        if status.any():
            if np.any(status & _cabi.STATUS_PEER_TIMEOUT):
                raise RuntimeError("a peer GPU did not deliver its star shard (timeout)")
            if np.any(status & _cabi.STATUS_NONFINITE_STAR):
                raise ValueError("invidual integrated star prob is not finite")
            if np.any(status & _cabi.STATUS_TOTAL_ZERO):
                print(0.0)
                raise SystemExit
        return out[0] if single else out

    def close(self):
        self.engine.close()
