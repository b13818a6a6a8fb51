#!/bin/bash
# compute-sanitizer (memcheck, then racecheck) over a small end-to-end slice of the hot path:
# uploads, both sampler kernels, explicit seeds through the integrator + post-processing, one small year
# (Fourier tables on the FP64 tensor cores), the return-period kernel, the pre-processing kernels.
OUT=gpurun_out/${1:-san}
mkdir -p $OUT
cat > /tmp/san_slice.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from tropical_cyclone_risk_b200.workload import Workload
from tropical_cyclone_risk_b200.engine import Engine
w = Workload("NA", [2001], months=[8, 9])
eng = Engine(w.p, device=0)
w.upload(eng)
rng = np.random.default_rng(0)
n = 3000
ym = rng.integers(0, 2, n).astype(np.int32); lon = rng.uniform(255, 365, n); lat = rng.uniform(-5, 65, n)
for v in (0, 5, 6):
    eng.set_interp_variant(v); eng.env_interp(ym, lon, lat)
eng.set_interp_variant(0)
n = 700
ym = rng.integers(0, 2, n).astype(np.int32)
o = eng.integrate(ym, rng.uniform(285, 345, n), rng.uniform(8, 32, n), 5 + rng.standard_normal(n), rng.uniform(.13, .32, n),
                  np.full(n, 1400.0), rng.random((n, 60)))
print("integrate ok", int(o["n_time"].sum()))
w2 = Workload("NA", [2002])
eng2 = Engine(w2.p, device=0); w2.upload(eng2)
r = eng2.run_years([0], [2002], 7, 6)
print("run_years ok", r["stats"][0]["storm_steps"])
print("poi", np.isfinite(eng2.poi_vmax(r["lon"][0], r["lat"][0], r["vmax"][0], 300.0, 25.0, 500.0)).sum())
r = eng2.run_years([0, 0], [2003, 2004], 9, 40)            # two years, several 256-attempt blocks each: the selection kernels
print("run_years (2 years x 40) ok", [s["attempts"] for s in r["stats"]])
# pre-processing kernels: ungrouped / grouped wind statistics (aligned and ragged rows), thermodynamics
from tropical_cyclone_risk_b200 import synth_thermo
for shape in ((19, 64), (7, 33)):
    ua = rng.normal(0, 8, (24, 2) + shape).astype(np.float32); va = rng.normal(0, 6, (24, 2) + shape).astype(np.float32)
    ua[3, 0, 2, 5] = np.nan
    eng2.wind_stats(ua, va, 0, 1, np.arange(25, dtype=np.int32))
    eng2.wind_stats(ua, va, 0, 1, np.arange(0, 25, 4, dtype=np.int32))
ua = rng.normal(0, 8, (744, 2, 3, 40)).astype(np.float32); va = rng.normal(0, 6, (744, 2, 3, 40)).astype(np.float32)
eng2.wind_stats(ua, va, 0, 1, np.arange(745, dtype=np.int32))          # hourly month, ungrouped: the streaming kernel
eng2.set_entropy_table(*synth_thermo.fixture_table())
p, ta, hus, sst, psl = synth_thermo.soundings(700, seed=2)
print("thermo", float(np.nanmean(eng2.thermo_month(p, ta, hus, sst, psl, 1.0, 13)[0])))
eng.close(); eng2.close()
PY
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_slice.py > $OUT/sanitizer_$tool.log 2>&1
  echo "exit $?" >> $OUT/sanitizer_$tool.log
  tail -6 $OUT/sanitizer_$tool.log
done
