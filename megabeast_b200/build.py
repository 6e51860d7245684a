"""
Build ``libmegabeast_b200.so`` (the C-ABI library of include/megabeast_b200.h) in-tree with
nvcc for sm_100a.  The built file is git-ignored but travels to the GPU box with the snapshot.

    python -m megabeast_b200.build [--force]
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(PKG, "csrc", "mb_api.cu")]
DEPS = SRC + [
    os.path.join(PKG, "csrc", "mb_kernels.cuh"),
    os.path.join(PKG, "..", "include", "megabeast_b200.h"),
]
# MEGABEAST_B200_LIB points the loader at an alternative build (kernel-tuning variants)
LIB = os.environ.get("MEGABEAST_B200_LIB") or os.path.join(PKG, "libmegabeast_b200.so")

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-shared",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libmegabeast_b200.so")


def is_stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False, defines=(), out=None):
    """Compile when missing or older than its sources; returns the library path."""
    out = out or LIB
    if not force and out == LIB and not is_stale():
        return LIB
    cmd = ([nvcc_path()] + NVCC_FLAGS + [f"-D{d}" for d in defines]
           + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + SRC)
    # the image's $CC/$CXX point at a gcc without the default search paths nvcc expects
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    res = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libmegabeast_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
