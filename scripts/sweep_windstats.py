"""Variant sweep of k_wind_stats on the B200 box (TCR_WS_VARIANT; see tcrisk.cu): GB/s per variant."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tropical_cyclone_risk_b200 import namelist as nl
from tropical_cyclone_risk_b200.engine import Engine
from tropical_cyclone_risk_b200.params import params_from_namelist

dev = torch.device("cuda:0")
eng = Engine(params_from_namelist(nl, "NA"), device=0)
n_pts = 721 * 1440
out = torch.empty((14, n_pts), dtype=torch.float64, device=dev)
res = {}
for name, spd, grouped, variants in (("single", 2, False, range(10, 15)), ("grouped", 4, True, range(0, 10))):
    n_time = 31 * spd
    ua = torch.randn((n_time, 2, n_pts), device=dev) * 8.0
    va = torch.randn((n_time, 2, n_pts), device=dev) * 6.0
    gs = np.arange(0, n_time + 1, spd if grouped else 1, dtype=np.int32)
    series = [ua.data_ptr(), va.data_ptr(), ua.data_ptr() + 4 * n_pts, va.data_ptr() + 4 * n_pts]
    moved = 16.0 * n_time * n_pts + 112.0 * n_pts
    for v in variants:
        os.environ["TCR_WS_VARIANT"] = str(v)
        try:
            for _ in range(2):
                eng.wind_stats_dev(n_time, n_pts, 2 * n_pts, series, gs, out.data_ptr())
            torch.cuda.synchronize()
            eng.set_timing(True)
            for _ in range(5):
                eng.wind_stats_dev(n_time, n_pts, 2 * n_pts, series, gs, out.data_ptr())
            ms, cnt = eng.kernel_times()["windstat"]
            eng.set_timing(False)
            res["%s_v%d" % (name, v)] = {"ms": ms / cnt, "GBps": moved / (ms / cnt * 1e-3) / 1e9}
        except Exception as e:                                   # e.g. shared-memory limit of a variant
            res["%s_v%d" % (name, v)] = {"error": str(e)[:100]}
            eng.set_timing(False)
    del ua, va
for k, v in res.items():
    print(k, json.dumps(v))
