"""
ORACLE -- TEST INFRASTRUCTURE ONLY.

Imports the *unmodified* reference modules ``megabeast.singlepop_dust_model`` and
``megabeast.helpers`` from ``/root/reference`` with the oracle's ``beast``/``emcee``
stand-ins shadowing the two absent third-party packages.

``/root/reference`` exists only in the build container, never on the GPU box, so the
only users are ``tests/golden/make_golden.py`` (which freezes golden vectors) and the
CPU tests that are skipped when the directory is missing.
"""
import importlib
import os
import sys
import warnings

REFERENCE_ROOT = os.environ.get("MEGABEAST_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_available():
    return os.path.isfile(
        os.path.join(REFERENCE_ROOT, "megabeast", "singlepop_dust_model.py")
    )


def load_reference():
    """Return (singlepop_dust_model, helpers) modules of the real reference."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    for p in (
        os.path.join(_HERE, "beast_standin"),
        os.path.join(_HERE, "emcee_standin"),
    ):
        if p not in sys.path:
            sys.path.insert(0, p)
    if REFERENCE_ROOT not in sys.path:
        # appended, not prepended: the stand-ins and this repo win every other name
        sys.path.append(REFERENCE_ROOT)
    # np.trapz (helpers.py:72,73,92) is deprecated in numpy 2.x but still present
    warnings.filterwarnings("ignore", message=".*trapz.*", category=DeprecationWarning)
    spdm = importlib.import_module("megabeast.singlepop_dust_model")
    helpers = importlib.import_module("megabeast.helpers")
    assert os.path.abspath(spdm.__file__).startswith(os.path.abspath(REFERENCE_ROOT))
    return spdm, helpers


class Settings:
    """Minimal object with the attributes ``MB_Model.__init__`` reads
    (``singlepop_dust_model.py:22-23``)."""

    def __init__(self, stellar_model, fd_model):
        self.stellar_model = stellar_model
        self.fd_model = fd_model
