"""Builds ``libtcrisk.so`` (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m tropical_cyclone_risk_b200.build [--force]

-fmad=false is part of the arithmetic contract (include/tcr_libm.h): only explicit fma()
calls fuse, so the device rounds exactly like the float64 specification.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libtcrisk.so")

SOURCES = ["tcrisk.cu"]
DEPS = ["tcrisk.cu", "tcr_kernels.cuh", "tcr_device.cuh", "tcr_rhs_fast.cuh", "tcr_preproc.cuh",
        os.path.join(ROOT, "include", "tcrisk.h"), os.path.join(ROOT, "include", "tcr_libm.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libtcrisk.so cannot be built")
    return exe


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for d in DEPS:
        path = d if os.path.isabs(d) else os.path.join(CSRC, d)
        if os.path.getmtime(path) > t:
            return True
    return False


def build(force=False, verbose=False):
    """Compile the library if missing or older than its sources; returns its path."""
    if not force and not stale():
        return LIB
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
