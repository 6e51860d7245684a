"""
ctypes binding of ``libmegabeast_b200.so`` -- a field-for-field mirror of
``include/megabeast_b200.h``.  This is the whole Python<->CUDA boundary: plain pointers and
sizes, no torch types.  There is no CPU fallback: if the library is missing it is built with
nvcc, and if that is impossible the import of the engine fails loudly.
"""
import ctypes
import os

from . import build as _build

MB_ABI_VERSION = 2
MB_NPARAM = 4
MB_MAX_NODES = 64
MB_NKERNELS = 3
MB_OUT_ROWS = 4
MB_OUT_ADDITIVE_ROWS = 3  # STARSUM, STARSUM_LO, NONFINITE add over star shards; POISSON is shared
PARAMS = ("logA", "Av", "Rv", "fA")  # enum mb_param order

KIND = {
    "fixed": 0,
    "flat": 1,
    "lognormal": 2,
    "two_lognormal": 3,
    "bins_histo": 4,
    "bins_interp": 5,
    "exponential": 6,
    "tabulated": 7,
}
# variables each analytic model reads from phi, in order (beast priormodel names)
KIND_VARNAMES = {
    "flat": (),
    "lognormal": ("mean", "sigma"),
    "two_lognormal": ("mean1", "mean2", "sigma1", "sigma2", "N1_to_N2"),
    "bins_histo": ("values",),
    "bins_interp": ("values",),
    "exponential": ("tau",),
}
MODE = {"auto": 0, "factor": 1, "table_unique": 2, "table_grid": 3, "elementwise": 4}
MODE_NAME = {v: k for k, v in MODE.items()}
KERNEL_NAMES = ("k0_factors", "k1_expand", "k2_stars")

STATUS_NONFINITE_STAR = 1
STATUS_TOTAL_ZERO = 2
STATUS_PEER_TIMEOUT = 4
MB_MAX_PEERS = 16
MB_PEER_HANDLE_BYTES = 64

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int64_p = ctypes.POINTER(ctypes.c_int64)
c_int32_p = ctypes.POINTER(ctypes.c_int32)


class ModelDesc(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int32 * MB_NPARAM),
        ("nnodes", ctypes.c_int32 * MB_NPARAM),
        ("nodes", (ctypes.c_double * MB_MAX_NODES) * MB_NPARAM),
    ]


class GridDesc(ctypes.Structure):
    _fields_ = [
        ("G", ctypes.c_int64),
        ("col", c_double_p * MB_NPARAM),
        ("logA", c_double_p),
        ("Z", c_double_p),
        ("prior_weight", c_double_p),
        ("grid_weight", c_double_p),
        ("completeness", c_double_p),
    ]


class StarsDesc(ctypes.Structure):
    _fields_ = [
        ("K", ctypes.c_int64),
        ("S", ctypes.c_int64),
        ("indxs", c_int64_p),
        ("vals", c_double_p),
        ("star_begin", ctypes.c_int64),
        ("star_end", ctypes.c_int64),
        ("n_stars_total", ctypes.c_int64),
    ]


class MassMultDesc(ctypes.Structure):
    _fields_ = [
        ("n_ages", ctypes.c_int32),
        ("n_mets", ctypes.c_int32),
        ("ages", c_double_p),
        ("mets", c_double_p),
        ("massmult", c_double_p),
        ("avemass", ctypes.c_double),
    ]


class Options(ctypes.Structure):
    _fields_ = [
        ("device", ctypes.c_int32),
        ("mode", ctypes.c_int32),
        ("keep_indices", ctypes.c_int32),
        ("merge_duplicates", ctypes.c_int32),
        ("segments_per_warp", ctypes.c_int32),
        ("tail_percent", ctypes.c_int32),
        ("table_precision", ctypes.c_int32),
        ("split_inner_mask", ctypes.c_int32),
        ("reserved", ctypes.c_int32 * 8),
    ]


class Info(ctypes.Structure):
    _fields_ = [
        ("abi_version", ctypes.c_int32),
        ("device", ctypes.c_int32),
        ("mode", ctypes.c_int32),
        ("ndim", ctypes.c_int32),
        ("poisson", ctypes.c_int32),
        ("n_classes", ctypes.c_int32 * MB_NPARAM),
        ("n_age_classes", ctypes.c_int32),
        ("n_dust_classes", ctypes.c_int32),
        ("n_bins", ctypes.c_int32),
        ("sm_count", ctypes.c_int32),
        ("walkers_per_pass", ctypes.c_int32),
        ("launches_last_eval", ctypes.c_int32),
        ("G", ctypes.c_int64),
        ("n_stars_local", ctypes.c_int64),
        ("n_stars_total", ctypes.c_int64),
        ("n_entries", ctypes.c_int64),
        ("n_entries_raw", ctypes.c_int64),
        ("device_bytes", ctypes.c_int64),
        ("code_bytes", ctypes.c_int64),
        ("n_outer_rows", ctypes.c_int64),
        ("n_inner_rows", ctypes.c_int64),
        ("split_inner_mask", ctypes.c_int64),
        ("n_runs", ctypes.c_int64),
        ("reserved", ctypes.c_int64 * 3),
    ]


TabArray = c_double_p * MB_NPARAM

# every symbol include/megabeast_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "mb_abi_version": (ctypes.c_int, []),
    "mb_last_error": (ctypes.c_char_p, []),
    "mb_device_count": (ctypes.c_int, []),
    "mb_create": (
        ctypes.c_int,
        [ctypes.POINTER(GridDesc), ctypes.POINTER(StarsDesc), ctypes.POINTER(ModelDesc),
         ctypes.POINTER(MassMultDesc), ctypes.POINTER(Options), ctypes.POINTER(ctypes.c_void_p)],
    ),
    "mb_destroy": (None, [ctypes.c_void_p]),
    "mb_get_info": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(Info)]),
    # hot call: raw addresses (ints) instead of typed pointers keeps the per-call ctypes cost low
    "mb_lnlike": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "mb_lnprob_box": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32)],
    ),
    "mb_lnlike_tabulated": (
        ctypes.c_int,
        [ctypes.c_void_p, c_double_p, ctypes.c_int32, ctypes.c_int32, TabArray, c_double_p, c_int32_p],
    ),
    "mb_lnlike_begin": (
        ctypes.c_int,
        [ctypes.c_void_p, c_double_p, ctypes.c_int32, ctypes.c_int32, TabArray],
    ),
    "mb_lnlike_end": (ctypes.c_int, [ctypes.c_void_p, c_double_p]),
    "mb_combine": (None, [c_double_p, ctypes.c_int32, c_double_p, c_int32_p]),
    "mb_lnlike_device": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
         ctypes.c_void_p],
    ),
    "mb_create_pixels": (
        ctypes.c_int,
        [ctypes.POINTER(GridDesc), ctypes.POINTER(StarsDesc), ctypes.POINTER(ModelDesc),
         ctypes.POINTER(MassMultDesc), ctypes.POINTER(Options), ctypes.c_int64, c_int64_p,
         ctypes.POINTER(ctypes.c_void_p)],
    ),
    "mb_lnlike_pixels": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "mb_lnlike_pixels_device": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "mb_lnlike_pixels_tabulated": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, TabArray, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "mb_lnprob_pixels_box": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p,
         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p],
    ),
    "mb_pixel_count": (ctypes.c_int64, [ctypes.c_void_p]),
    "mb_watch_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]),
    "mb_peer_export": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "mb_peer_connect": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]),
    "mb_peer_disconnect": (ctypes.c_int, [ctypes.c_void_p]),
    "mb_peer_rendezvous": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "mb_class_values": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, c_double_p, ctypes.c_int32]),
    "mb_gathered_indices": (ctypes.c_int, [ctypes.c_void_p, c_int64_p, ctypes.c_int64, c_int64_p]),
    "mb_device_entries": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, c_int32_p, c_double_p, c_int64_p]),
    "mb_set_profiling": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32]),
    "mb_last_kernel_ms": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float * MB_NKERNELS)]),
    "mb_last_k2_timeline": (ctypes.c_int, [ctypes.c_void_p, c_double_p, ctypes.c_int32]),
    "mb_likelihoods_from_lnp": (
        ctypes.c_int,
        [ctypes.c_int32, ctypes.c_int64, ctypes.c_int64, c_int64_p, c_double_p, ctypes.c_int64,
         c_double_p],
    ),
}

_lib = None


class MegabeastB200Error(RuntimeError):
    """A C-ABI call failed (CUDA error, invalid argument, no device)."""


def lib_path():
    return _build.LIB


def load():
    """Load (building first if needed) the CUDA library; raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if _build.is_stale():
        try:
            _build.build()
        except Exception as exc:  # no nvcc and no prebuilt library: nothing to run on
            if not os.path.isfile(path):
                raise MegabeastB200Error(
                    f"libmegabeast_b200.so is missing and cannot be built ({exc}); "
                    "megabeast_b200 has no CPU fallback"
                ) from exc
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    if lib.mb_abi_version() != MB_ABI_VERSION:
        raise MegabeastB200Error("libmegabeast_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise MegabeastB200Error(load().mb_last_error().decode("utf-8", "replace"))


def dptr(a):
    return a.ctypes.data_as(c_double_p)
