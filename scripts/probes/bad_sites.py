#!/usr/bin/env python
"""Which operations of the straight-line RHS (csrc/tcr_rhs_fast.cuh) leave their common case, and how often.
Build the instrumented library first (here, nvcc cross-compiles):
    nvcc <build.py flags> -DTCR_DEBUG_BAD -o scripts/probes/libtcrisk_dbg.so tropical_cyclone_risk_b200/csrc/tcrisk.cu
then on the GPU box:  TCR_LIB_PATH=scripts/probes/libtcrisk_dbg.so python scripts/probes/bad_sites.py [basin years tracks]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("TCR_LIB_PATH", os.path.join(ROOT, "scripts", "probes", "libtcrisk_dbg.so"))
os.environ.setdefault("TCR_INTEG_VARIANT", "22")
import numpy as np                                                   # noqa: E402
from tropical_cyclone_risk_b200 import _lib, workload                # noqa: E402
from tropical_cyclone_risk_b200.engine import Engine                 # noqa: E402

SITES = ["state", "fs_index", "locate", "cos", "cholesky_pivot", "range", "range_fourier", "-", "polar", "-", "log", "exp", "range_mixing"]


def main():
    basin = sys.argv[1] if len(sys.argv) > 1 else "NA"
    years = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    tracks = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    wl = workload.Workload(basin, [2001 + i for i in range(years)])
    eng = Engine(wl.p, device=0)
    wl.upload(eng)
    lib = _lib.load()
    lib.tcr_debug_bad.argtypes = [C.c_void_p, C.c_int]
    out = (C.c_ulonglong * 33)()
    lib.tcr_debug_bad(out, 1)
    r = eng.run_years([12 * i for i in range(years)], [2001 + i for i in range(years)], 20260101, tracks)
    lib.tcr_debug_bad(out, 0)
    total = out[32]
    print("evaluations %d, storm-steps %d" % (total, sum(s["storm_steps"] for s in r["stats"])))
    for i, name in enumerate(SITES):
        print("  %-12s %10d  %.4f %%" % (name, out[i], 100.0 * out[i] / max(1, total)))
    eng.close()


if __name__ == "__main__":
    main()
