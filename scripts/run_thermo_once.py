"""A few launches of k_thermo on the 0.25-degree grid, 28 levels (ncu target)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tropical_cyclone_risk_b200 import namelist as nl
from tropical_cyclone_risk_b200 import synth_thermo
from tropical_cyclone_risk_b200.engine import Engine
from tropical_cyclone_risk_b200.params import params_from_namelist

dev = torch.device("cuda:0")
eng = Engine(params_from_namelist(nl, "NA"), device=0)
eng.set_entropy_table(*synth_thermo.fixture_table())
n_pts, base = 721 * 1440, 8192
p, ta, hus, sst, psl = synth_thermo.soundings(base, seed=21)
reps = (n_pts + base - 1) // base
tile = lambda a: torch.from_numpy(np.ascontiguousarray(np.tile(a, reps)[..., :n_pts])).to(dev)
d_ta, d_hus, d_sst, d_psl = tile(ta), tile(hus), tile(sst), tile(psl)
out = torch.empty((3, n_pts), dtype=torch.float64, device=dev)
for _ in range(3):
    eng.thermo_month_dev(n_pts, p, d_ta.data_ptr(), d_hus.data_ptr(), d_sst.data_ptr(), d_psl.data_ptr(), 1.0, 13,
                         out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr())
torch.cuda.synchronize()
print("ok", float(out[0].mean()))
