#!/bin/bash
# round 2, visit X: reversible thermodynamics on the GPU, compute-sanitizer over the Fourier-ring path of the integrator and
# k_thermo<true>, and an A/B of the workspace budget (fewer, larger waves at configs[2])
TAG=${1:-r04x}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 600 python -m pytest tests -m gpu -x -q -k "thermo or ring" > $OUT/pytest_gpu_thermo_ring.log 2>&1 ) 2>&1 | grep real; tail -4 $OUT/pytest_gpu_thermo_ring.log | cut -c1-600
cat > /tmp/san_ring.py <<'PY'
import os, sys, types, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
os.environ["TCR_FTAB_RING"] = "1"
from tropical_cyclone_risk_b200.workload import Workload
from tropical_cyclone_risk_b200.engine import Engine
from tropical_cyclone_risk_b200 import namelist as nl, synth_thermo
w = Workload("NA", [2002])
eng = Engine(w.p, device=0); w.upload(eng)
r = eng.run_years([0], [2002], 7, 6)
print("ring run_years ok", r["stats"][0]["storm_steps"], int(np.sum(~np.isnan(r["lon"]))))
eng.set_tuning(max_wave=2048, max_slots=600, oversub_permille=1100)
r = eng.run_years([0, 0], [2003, 2004], 9, 12)
print("ring run_years (tiny waves) ok", [s["attempts"] for s in r["stats"]])
eng.close()
nl900 = types.SimpleNamespace(**{k: getattr(nl, k) for k in dir(nl) if not k.startswith("__")}); nl900.output_interval_s = 900
w9 = Workload("WP", [2006], namelist=nl900)
e9 = Engine(w9.p, device=0); w9.upload(e9)
r = e9.run_years([0], [2006], 3, 4)
print("ring 900 s ok", r["stats"][0]["storm_steps"])
g = np.load("tests/golden/ref_thermo_rev.npz")
e9.set_entropy_table_reversible(g["table_p"], g["table_s"], g["table_rt"], g["table_T"])
p, ta, hus, sst, psl = synth_thermo.soundings(700, seed=2)
print("thermo rev", float(np.nanmean(e9.thermo_month(p, ta, hus, sst, psl, 1.0, 13, select_thermo=2)[0])))
e9.close()
PY
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  ( time timeout 700 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_ring.py > $OUT/sanitizer_ring_$tool.log 2>&1 ) 2>&1 | grep real
  echo "exit $?" >> $OUT/sanitizer_ring_$tool.log
  tail -7 $OUT/sanitizer_ring_$tool.log | cut -c1-300
done
run() {  # name, env, args
  ( time env $2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-interp $3 > $OUT/bench_$1.json 2> $OUT/bench_$1.err ) 2>&1 | grep real
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$1.json"))
    print("$1: value %.3e e2e %.3e ms/step %.2f integrate avg %.3f ms share %.2f waves %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["roofline"]["share_of_step"],d["details"]["waves_per_step"]), {k:round(v,3) for k,v in d["details"]["kernel_share_of_step"].items()})
except Exception as e:
    print("$1 failed", e); print(open("$OUT/bench_$1.err").read()[-2000:])
PY
}
run cfg2_ws60 "TCR_WS_FRAC=0.6" ""
run cfg2_ws80 "TCR_WS_FRAC=0.8 TCR_WS_CAP_GB=150" ""
