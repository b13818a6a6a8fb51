"""A few launches of k_wind_stats on the 0.25-degree grid (ncu target).  usage: run_windstats_once.py [grouped]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tropical_cyclone_risk_b200 import namelist as nl
from tropical_cyclone_risk_b200.engine import Engine
from tropical_cyclone_risk_b200.params import params_from_namelist

grouped = len(sys.argv) > 1 and sys.argv[1] == "grouped"
dev = torch.device("cuda:0")
eng = Engine(params_from_namelist(nl, "NA"), device=0)
n_pts = 721 * 1440
spd = 4 if grouped else 2
n_time = 31 * spd
out = torch.empty((14, n_pts), dtype=torch.float64, device=dev)
ua = torch.randn((n_time, 2, n_pts), device=dev) * 8.0
va = torch.randn((n_time, 2, n_pts), device=dev) * 6.0
gs = np.arange(0, n_time + 1, spd if grouped else 1, dtype=np.int32)
series = [ua.data_ptr(), va.data_ptr(), ua.data_ptr() + 4 * n_pts, va.data_ptr() + 4 * n_pts]
for _ in range(3):
    eng.wind_stats_dev(n_time, n_pts, 2 * n_pts, series, gs, out.data_ptr())
torch.cuda.synchronize()
print("ok", float(out[0, 0]))
