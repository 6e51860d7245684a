"""
``LikelihoodEngine`` -- owns the device-resident form of one (grid, stars, model) triple and
evaluates the ensemble likelihood for batches of walkers through the C ABI.

One ``mb_handle`` per CUDA device; with several devices the stars are split into contiguous
shards (grid replicated), every device evaluates all walkers on its shard, and the per-walker
partials are added.  Inside one process that sum is done on the host
after the asynchronous ``mb_lnlike_begin``/``mb_lnlike_end`` pair; across processes
(``torchrun``) it is one NCCL all-reduce, see ``parallel.py``.
"""
import ctypes
import os

import numpy as np

from . import _cabi

__all__ = ["ModelSpec", "LikelihoodEngine", "shard_bounds"]


def shard_bounds(n_stars, n_shards, i):
    """Contiguous, balanced star range of shard ``i`` (first ``n_stars % n_shards`` shards get
    one extra star)."""
    q, r = divmod(int(n_stars), int(n_shards))
    lo = i * q + min(i, r)
    return lo, lo + q + (1 if i < r else 0)


class ModelSpec:
    """What the device needs to know about the free prior models: per parameter (logA, Av, Rv,
    fA) a kind name from ``_cabi.KIND`` and, for binned models, the node coordinates."""

    def __init__(self, kinds, nodes=None):
        self.kinds = dict(kinds)
        self.nodes = dict(nodes or {})
        for p in _cabi.PARAMS:
            self.kinds.setdefault(p, "fixed")
            if self.kinds[p] not in _cabi.KIND:
                raise ValueError(f"unknown prior model kind {self.kinds[p]!r}")

    def nvars(self, p):
        kind = self.kinds[p]
        if kind in ("bins_histo", "bins_interp"):
            return len(self.nodes[p])
        return len(_cabi.KIND_VARNAMES.get(kind, ()))

    @property
    def ndim(self):
        return sum(self.nvars(p) for p in _cabi.PARAMS)

    @property
    def tabulated(self):
        return [p for p in _cabi.PARAMS if self.kinds[p] == "tabulated"]

    def to_c(self):
        m = _cabi.ModelDesc()
        for i, p in enumerate(_cabi.PARAMS):
            m.kind[i] = _cabi.KIND[self.kinds[p]]
            if self.kinds[p] in ("bins_histo", "bins_interp"):
                x = [float(v) for v in self.nodes[p]]
                if len(x) > _cabi.MB_MAX_NODES:
                    raise ValueError(f"{p}: more than {_cabi.MB_MAX_NODES} nodes")
                m.nnodes[i] = len(x)
                for j, v in enumerate(x):
                    m.nodes[i][j] = v
        return m


class LikelihoodEngine:
    def __init__(self, spec, star_lnpgriddata, beast_moddata, massmult=None, devices=None,
                 mode="auto", star_range=None, n_stars_total=None, keep_indices=False,
                 merge_duplicates=True, segments_per_warp=0, pixel_offsets=None, table_precision="fp64",
                 split_inner=None):
        """``split_inner``: FACTOR mode only -- the dust parameters kept in the inner factor table
        (e.g. ``("Av",)``; the others join the age classes in the outer table).  ``None`` lets
        ``mb_create`` choose (all of them when the product table fits shared memory)."""
        self.lib = _cabi.load()
        self._mb_lnlike = self.lib.mb_lnlike
        self._bufs = {}
        self._zero_status = {}
        self.spec = spec
        self.handles = []
        self.devices = list(devices) if devices is not None else [0]
        if not self.devices:
            raise ValueError("at least one device is required")

        keep = []  # host arrays that must outlive mb_create

        def arr(a, dtype):
            a = np.ascontiguousarray(a, dtype=dtype)
            keep.append(a)
            return a

        grid = _cabi.GridDesc()
        grid.G = len(beast_moddata["prior_weight"])
        for i, p in enumerate(_cabi.PARAMS):
            if spec.kinds[p] != "fixed":
                if p not in beast_moddata:
                    raise KeyError(p)  # singlepop_dust_model.py:147 looks the column up by name
                grid.col[i] = _cabi.dptr(arr(beast_moddata[p], np.float64))
        grid.prior_weight = _cabi.dptr(arr(beast_moddata["prior_weight"], np.float64))
        grid.grid_weight = _cabi.dptr(arr(beast_moddata["grid_weight"], np.float64))
        grid.completeness = _cabi.dptr(arr(beast_moddata["completeness"], np.float64))
        mm = None
        if massmult is not None:
            grid.logA = _cabi.dptr(arr(beast_moddata["logA"], np.float64))
            grid.Z = _cabi.dptr(arr(beast_moddata["Z"], np.float64))
            mm = _cabi.MassMultDesc()
            ages = arr(massmult["ages"], np.float64)
            mets = arr(massmult["Zs"], np.float64)
            mult = arr(massmult["massmult"], np.float64)
            mm.n_ages, mm.n_mets = ages.size, mets.size
            mm.ages, mm.mets, mm.massmult = _cabi.dptr(ages), _cabi.dptr(mets), _cabi.dptr(mult)
            mm.avemass = float(massmult["avemass"])

        indxs = np.asarray(star_lnpgriddata["indxs"])
        if not np.issubdtype(indxs.dtype, np.integer):
            raise IndexError("arrays used as indices must be of integer (or boolean) type")
        indxs = arr(indxs, np.int64)
        vals = arr(star_lnpgriddata["vals"], np.float64)
        if indxs.shape != vals.shape or indxs.ndim != 2:
            raise ValueError("indxs and vals must both have shape (n_lnps, n_stars)")
        K, S = indxs.shape
        lo, hi = star_range if star_range is not None else (0, S)
        self.star_range = (int(lo), int(hi))
        self.n_stars_total = int(n_stars_total) if n_stars_total is not None else S
        model_c = spec.to_c()
        self.n_pixels = 0
        if pixel_offsets is not None:
            if len(self.devices) != 1:
                raise ValueError("a pixel batch lives on one device; give each rank its own pixels")
            pixel_offsets = arr(pixel_offsets, np.int64)
            self.n_pixels = pixel_offsets.size - 1

        try:
            for i, dev in enumerate(self.devices):
                a, b = shard_bounds(hi - lo, len(self.devices), i)
                st = _cabi.StarsDesc()
                st.K, st.S = K, S
                st.indxs = indxs.ctypes.data_as(_cabi.c_int64_p)
                st.vals = _cabi.dptr(vals)
                st.star_begin, st.star_end = lo + a, lo + b
                st.n_stars_total = self.n_stars_total
                opt = _cabi.Options()
                opt.device = int(dev)
                opt.mode = _cabi.MODE[mode]
                opt.keep_indices = int(bool(keep_indices))
                opt.merge_duplicates = int(bool(merge_duplicates))
                opt.segments_per_warp = int(os.environ.get("MB_SEGMENTS_PER_WARP", segments_per_warp))
                opt.tail_percent = int(os.environ.get("MB_TAIL_PERCENT", 0))
                opt.table_precision = {"fp64": 0, "fp32": 1}[table_precision]
                opt.split_inner_mask = sum(1 << (_cabi.PARAMS.index(p) - 1) for p in (split_inner or ()))
                h = ctypes.c_void_p()
                if pixel_offsets is not None:
                    _cabi.check(
                        self.lib.mb_create_pixels(
                            ctypes.byref(grid), ctypes.byref(st), ctypes.byref(model_c),
                            ctypes.byref(mm) if mm is not None else None, ctypes.byref(opt),
                            self.n_pixels, pixel_offsets.ctypes.data_as(_cabi.c_int64_p), ctypes.byref(h),
                        )
                    )
                else:
                    _cabi.check(
                        self.lib.mb_create(
                            ctypes.byref(grid), ctypes.byref(st), ctypes.byref(model_c),
                            ctypes.byref(mm) if mm is not None else None, ctypes.byref(opt),
                            ctypes.byref(h),
                        )
                    )
                self.handles.append(h)
        except Exception:
            self.close()
            raise
        del keep

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        for h in self.handles:
            self.lib.mb_destroy(h)
        self.handles = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- introspection ---------------------------------------------------------------------------
    def info(self, i=0):
        out = _cabi.Info()
        _cabi.check(self.lib.mb_get_info(self.handles[i], ctypes.byref(out)))
        d = {name: getattr(out, name) for name, _ in _cabi.Info._fields_ if name != "reserved"}
        d["n_classes"] = list(out.n_classes)
        d["mode"] = _cabi.MODE_NAME[out.mode]
        return d

    def class_values(self, param, i=0):
        p = _cabi.PARAMS.index(param)
        n = self.lib.mb_class_values(self.handles[i], p, None, 0)
        out = np.empty(max(n, 0))
        self.lib.mb_class_values(self.handles[i], p, _cabi.dptr(out), n)
        return out

    def gathered_indices(self, i=0):
        info = self.info(i)
        rows = np.empty(max(1, info["n_entries_raw"]), dtype=np.int64)
        offs = np.empty(info["n_stars_local"] + 1, dtype=np.int64)
        _cabi.check(
            self.lib.mb_gathered_indices(
                self.handles[i], rows.ctypes.data_as(_cabi.c_int64_p), rows.size,
                offs.ctypes.data_as(_cabi.c_int64_p),
            )
        )
        return rows[: offs[-1]], offs

    def watch(self, arrays):
        """Guard against in-place changes of the host arrays the device copy was made from
        (``mb_watch_host``): every later evaluation first compares a fingerprint of each array and
        fails with ``MB_STALE`` when one changed.  The engine keeps the arrays alive."""
        self._watched = [a for a in arrays if isinstance(a, np.ndarray) and a.flags.c_contiguous][:16]
        for h in self.handles:
            _cabi.check(self.lib.mb_watch_host(h, None, 0))
            for a in self._watched:
                _cabi.check(self.lib.mb_watch_host(h, ctypes.c_void_p(a.ctypes.data), a.nbytes))

    def device_entries(self, i=0):
        """Parity hook (``mb_device_entries``): the class-coded entry stream as the DEVICE holds it.
        Returns (cls[n, 4] class index per parameter, weight[n], offsets[S_local + 1])."""
        info = self.info(i)
        n = max(1, info["n_entries"])
        cls = np.zeros((n, _cabi.MB_NPARAM), dtype=np.int32)
        weight = np.zeros(n)
        offs = np.zeros(info["n_stars_local"] + 1, dtype=np.int64)
        _cabi.check(self.lib.mb_device_entries(self.handles[i], n, cls.ctypes.data_as(_cabi.c_int32_p),
                                               _cabi.dptr(weight), offs.ctypes.data_as(_cabi.c_int64_p)))
        return cls[: offs[-1]], weight[: offs[-1]], offs

    # -- cross-GPU reduction fused into the likelihood kernel (one process per GPU) -------------
    def peer_export(self, i=0):
        """Opaque token of this handle's exchange buffer, to be all-gathered across ranks."""
        token = ctypes.create_string_buffer(_cabi.MB_PEER_HANDLE_BYTES)
        _cabi.check(self.lib.mb_peer_export(self.handles[i], token))
        return token.raw

    def peer_connect(self, rank, world, tokens, i=0):
        """Map every rank's exchange buffer (tokens in rank order); afterwards result rows come
        back already summed over all ranks."""
        blob = b"".join(tokens)
        if len(blob) != world * _cabi.MB_PEER_HANDLE_BYTES:
            raise ValueError("need one token per rank")
        _cabi.check(self.lib.mb_peer_connect(self.handles[i], rank, world, blob))

    def peer_rendezvous(self, stream=0, i=0):
        """Device-side rendezvous of the connected ranks on ``stream`` (measurement aid)."""
        _cabi.check(self.lib.mb_peer_rendezvous(self.handles[i], ctypes.c_void_p(stream)))

    def peer_disconnect(self, i=0):
        _cabi.check(self.lib.mb_peer_disconnect(self.handles[i]))

    def set_profiling(self, on=True):
        for h in self.handles:
            _cabi.check(self.lib.mb_set_profiling(h, int(on)))

    def last_kernel_ms(self, i=0):
        ms = (ctypes.c_float * _cabi.MB_NKERNELS)()
        _cabi.check(self.lib.mb_last_kernel_ms(self.handles[i], ctypes.byref(ms)))
        return dict(zip(_cabi.KERNEL_NAMES, list(ms)))

    def last_k2_timeline(self, i=0):
        """Per-block phase stamps of the last K2 launch (profiling on): array [blocks][8] in
        microseconds -- start, tables staged, N-stars bins done, star loop done, block reduced, and for
        the last block: Poisson term done + shard totals sent, peers' totals received, end (-1 elsewhere)."""
        cap = 1024
        us = np.empty((cap, 8))
        n = self.lib.mb_last_k2_timeline(self.handles[i], _cabi.dptr(us), cap)
        if n < 0:
            _cabi.check(1)
        return us[:n]

    # -- evaluation ------------------------------------------------------------------------------
    def _tabs(self, tabs, W):
        t = _cabi.TabArray()
        keep = []
        for i, p in enumerate(_cabi.PARAMS):
            if self.spec.kinds[p] == "tabulated":
                a = np.ascontiguousarray(tabs[p], dtype=np.float64)
                if a.shape[0] != W:
                    raise ValueError(f"table for {p} must have one row per walker")
                keep.append(a)
                t[i] = _cabi.dptr(a)
        return t, keep

    def lnlike_rows(self, phis, tabs=None):
        """Raw result rows [4][W] (sum over this engine's stars of ln(star prob) in two exact words;
        count of non-finite stars; Poisson term), summed over the engine's devices."""
        phis = np.ascontiguousarray(phis, dtype=np.float64)
        if phis.ndim != 2:
            raise ValueError("phis must be (W, ndim)")
        W, ndim = phis.shape
        t, keep = self._tabs(tabs or {}, W)
        for h in self.handles:
            _cabi.check(self.lib.mb_lnlike_begin(h, _cabi.dptr(phis), W, ndim, t))
        total = None
        for h in self.handles:
            rows = np.empty((_cabi.MB_OUT_ROWS, W))
            _cabi.check(self.lib.mb_lnlike_end(h, _cabi.dptr(rows)))
            if total is None:
                total = rows
            else:
                # STARSUM, STARSUM_LO and NONFINITE add over shards (the two STARSUM words are exact
                # multiples of 2^-24 / 2^-68: no rounding here), POISSON is shared
                total[:_cabi.MB_OUT_ADDITIVE_ROWS] += rows[:_cabi.MB_OUT_ADDITIVE_ROWS]
        return total

    def combine(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        W = rows.shape[1]
        out = np.empty(W)
        status = np.empty(W, dtype=np.int32)
        self.lib.mb_combine(_cabi.dptr(rows), W, _cabi.dptr(out), status.ctypes.data_as(_cabi.c_int32_p))
        return out, status

    def lnlike(self, phis, tabs=None):
        """(lnlike[W], status[W]) for a batch of walkers.  The returned arrays are reused by the
        next call with the same W (copy them to keep them)."""
        if len(self.handles) == 1 and not self.spec.tabulated:
            if not (phis.dtype == np.float64 and phis.flags.c_contiguous):
                phis = np.ascontiguousarray(phis, dtype=np.float64)
            W, ndim = phis.shape
            buf = self._bufs.get(W)
            if buf is None:
                out, status = np.empty(W), np.empty(W, dtype=np.int32)
                buf = self._bufs[W] = (out, status, out.ctypes.data, status.ctypes.data)
            out, status, p_out, p_status = buf
            if self._mb_lnlike(self.handles[0], phis.ctypes.data, W, ndim, p_out, p_status):
                _cabi.check(1)
            return out, status
        return self.combine(self.lnlike_rows(phis, tabs))

    def lnlike_pixels(self, phis, tabs=None):
        """Pixel batch: phis [P][W][ndim] -> (lnlike[P][W], status[P][W]), one launch.  ``tabs``:
        host-evaluated tables {param: [P*W][n_classes]} for tabulated parameters."""
        phis = np.ascontiguousarray(phis, dtype=np.float64)
        if phis.ndim != 3 or phis.shape[0] != self.n_pixels:
            raise ValueError(f"phis must be ({self.n_pixels}, W, ndim)")
        P, W, ndim = phis.shape
        out = np.empty((P, W))
        status = np.empty((P, W), dtype=np.int32)
        t, keep = self._tabs(tabs or {}, P * W)
        _cabi.check(self.lib.mb_lnlike_pixels_tabulated(self.handles[0], phis.ctypes.data, W, ndim, t,
                                                        out.ctypes.data, status.ctypes.data))
        return out, status

    def lnprob_pixels_box(self, phis, lo, hi, park):
        """Pixel batch with the flat prior box applied in the C call (``mb_lnprob_pixels_box``): phis
        [P][W][ndim] -> (lnprob[P][W] with -inf outside [lo, hi], status[P][W]).  ``park``: an in-box
        position evaluated in place of the out-of-box walkers.  Analytic models."""
        phis = np.ascontiguousarray(phis, dtype=np.float64)
        if phis.ndim != 3 or phis.shape[0] != self.n_pixels:
            raise ValueError(f"phis must be ({self.n_pixels}, W, ndim)")
        P, W, ndim = phis.shape
        lo, hi, park = (np.ascontiguousarray(a, dtype=np.float64) for a in (lo, hi, park))
        if not (lo.size == hi.size == park.size == ndim):
            raise ValueError("lo, hi and park must have ndim entries")
        out = np.empty((P, W))
        status = np.empty((P, W), dtype=np.int32)
        _cabi.check(self.lib.mb_lnprob_pixels_box(self.handles[0], phis.ctypes.data, W, ndim, lo.ctypes.data,
                                                  hi.ctypes.data, park.ctypes.data, out.ctypes.data,
                                                  status.ctypes.data))
        return out, status

    def lnlike_pixels_device(self, d_phi, W, ndim, d_out, stream=0):
        """Device-pointer form: d_phi [P][W][ndim], d_out [P][4][W], asynchronous on ``stream``."""
        _cabi.check(self.lib.mb_lnlike_pixels_device(self.handles[0], ctypes.c_void_p(d_phi), W, ndim,
                                                     ctypes.c_void_p(d_out), ctypes.c_void_p(stream)))

    def zero_status(self, W):
        """The bytes of W all-clear status words (for a fast "anything flagged?" test)."""
        z = self._zero_status.get(W)
        if z is None:
            z = self._zero_status[W] = bytes(4 * W)
        return z

    def lnprob_box(self, phis, lo, hi):
        """(lnprob[W], status[W], last_inside) for flat box priors, one C call (``mb_lnprob_box``):
        -inf outside the box without evaluating the likelihood.  Single device, analytic models.
        The returned arrays are reused by the next call with the same W."""
        W, ndim = phis.shape
        buf = self._bufs.get((-W, ndim))
        if buf is None or buf[5] is not lo or buf[6] is not hi:
            out, status, last = np.empty(W), np.empty(W, dtype=np.int32), ctypes.c_int32(-1)
            stage = np.empty((W, ndim))  # phi is copied here: fetching an array's address costs more than the copy
            # everything that does not change between calls is converted to ctypes once
            buf = self._bufs[(-W, ndim)] = (out, status, last, self.lib.mb_lnprob_box,
                                            (ctypes.c_void_p(lo.ctypes.data), ctypes.c_void_p(hi.ctypes.data),
                                             ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(status.ctypes.data),
                                             ctypes.byref(last), ctypes.c_void_p(stage.ctypes.data)), lo, hi, stage)
        out, status, last, fn, (p_lo, p_hi, p_out, p_status, p_last, p_stage) = buf[:5]
        np.copyto(buf[7], phis)
        if fn(self.handles[0], p_stage, W, ndim, p_lo, p_hi, p_out, p_status, p_last):
            _cabi.check(1)
        return out, status, last.value

    def lnlike_device(self, d_phi, W, ndim, d_out, stream=0, i=0):
        """Asynchronous device-pointer form (``mb_lnlike_device``): ``d_phi`` [W][ndim] and
        ``d_out`` [4][W] are raw device addresses, ``stream`` a raw ``cudaStream_t``."""
        _cabi.check(
            self.lib.mb_lnlike_device(
                self.handles[i], ctypes.c_void_p(d_phi), W, ndim, ctypes.c_void_p(d_out),
                ctypes.c_void_p(stream),
            )
        )
