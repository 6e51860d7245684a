"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  ctypes front end of ``oracle/mb_oracle.c`` (the plain-C
restatement, OpenMP over walkers).  Used where the numpy port would take minutes: full-size
parity checks and the multi-threaded CPU arm of ``bench.py``.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import lnprob_port as port

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libmb_oracle.so")

NPARAM, MAX_NODES = 4, 64
C_PARAMS = ("logA", "Av", "Rv", "fA")
KIND = {"fixed": 0, "flat": 1, "lognormal": 2, "two_lognormal": 3, "bins_histo": 4,
        "bins_interp": 5, "exponential": 6}


class MboModel(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int32 * NPARAM),
        ("nnodes", ctypes.c_int32 * NPARAM),
        ("nodes", (ctypes.c_double * MAX_NODES) * NPARAM),
        ("fixed_tail", ctypes.c_double * NPARAM),
    ]


def build(force=False):
    src = os.path.join(_HERE, "mb_oracle.c")
    if force or not os.path.isfile(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "mb_oracle.c")
        # (the source is absent where only the built library travels; rebuild when it is there and newer)
        if not os.path.isfile(_LIB) or (os.path.isfile(src) and os.path.getmtime(_LIB) < os.path.getmtime(src)):
            build()
        _lib = ctypes.CDLL(_LIB)
        _lib.mbo_lnlike_batch.restype = ctypes.c_int
        assert _lib.mbo_sizeof_model() == ctypes.sizeof(MboModel)
    return _lib


def _model_struct(pm):
    m = MboModel()
    for p, name in enumerate(C_PARAMS):
        mod = pm.physics_model[name]
        if mod["prior"]["name"] == "fixed":
            m.kind[p] = KIND["fixed"]
            continue
        m.kind[p] = KIND[mod["name"]]
        if mod["name"] in ("bins_histo", "bins_interp"):
            x = list(mod["x"])
            m.nnodes[p] = len(x)
            for i, v in enumerate(x):
                m.nodes[p][i] = v
    if pm.physics_model["M_ini"]["prior"]["name"] != "fixed":
        raise KeyError("nsubvars")  # what the reference does (singlepop_dust_model.py:137)
    return m


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def lnlike_batch(pm, phis, stars, moddata, nthreads=0):
    """lnlike for every row of ``phis`` -> (values[W], status[W]); see mb_oracle.c."""
    phis = np.ascontiguousarray(np.atleast_2d(phis), dtype=np.float64)
    W, ndim = phis.shape
    m = _model_struct(pm)
    G = len(moddata["Av"])
    keep = []
    cols = (ctypes.POINTER(ctypes.c_double) * NPARAM)()
    for p, name in enumerate(C_PARAMS):
        if m.kind[p] != 0:
            a = np.ascontiguousarray(moddata[name], dtype=np.float64)
            keep.append(a)
            cols[p] = _dptr(a)
    pw = np.ascontiguousarray(moddata["prior_weight"], dtype=np.float64)
    gw = np.ascontiguousarray(moddata["grid_weight"], dtype=np.float64)
    comp = np.ascontiguousarray(moddata["completeness"], dtype=np.float64)
    if pm.compute_N_stars:
        if pm.massmultipliers is None:
            pm.massmultipliers = port.mass_multipliers(moddata, pm.physics_model["M_ini"])
        mi = pm.massmultipliers
        ages, mets = mi["ages"], mi["Zs"]
        ia = np.searchsorted(ages, moddata["logA"])
        iz = np.searchsorted(mets, moddata["Z"])
        bins = (ia * len(mets) + iz).astype(np.int32)
        nb = len(ages) * len(mets)
        age_of_bin = np.repeat(np.arange(len(ages)), len(mets)).astype(np.int32)
        coef = np.ascontiguousarray((mi["massmult"] / len(mets) / mi["avemass"]).ravel())
        ages = np.ascontiguousarray(ages, dtype=np.float64)
    else:
        bins = np.zeros(1, np.int32)
        nb, age_of_bin, coef, ages = 0, np.zeros(1, np.int32), np.zeros(1), np.zeros(1)
    indxs = np.ascontiguousarray(stars["indxs"], dtype=np.int64)
    vals = np.ascontiguousarray(stars["vals"], dtype=np.float64)
    K, S = indxs.shape
    out = np.empty(W)
    status = np.empty(W, np.int32)
    rc = lib().mbo_lnlike_batch(
        ctypes.c_int64(G), cols, _dptr(pw), _dptr(gw), _dptr(comp),
        bins.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.c_int32(nb),
        age_of_bin.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _dptr(coef),
        ctypes.c_int32(len(ages)), _dptr(ages), ctypes.byref(m),
        ctypes.c_int64(S), ctypes.c_int64(K),
        indxs.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _dptr(vals),
        ctypes.c_int32(W), ctypes.c_int32(ndim), _dptr(phis), _dptr(out),
        status.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), ctypes.c_int32(nthreads))
    if rc != 0:
        raise RuntimeError(f"mbo_lnlike_batch failed rc={rc}")
    return out, status


def lnprob_batch(pm, phis, stars, moddata, nthreads=0):
    """lnprior + lnlike, -inf (lnlike not evaluated) outside the box; raises like the
    reference on a non-finite star."""
    phis = np.atleast_2d(phis)
    lp = np.array([pm.lnprior(p) for p in phis])
    inside = np.isfinite(lp)
    out = np.full(len(phis), -np.inf)
    if inside.any():
        v, st = lnlike_batch(pm, phis[inside], stars, moddata, nthreads)
        if np.any(st & 1):
            raise ValueError("invidual integrated star prob is not finite")
        if np.any(st & 2):
            raise SystemExit
        out[inside] = lp[inside] + v
    return out
