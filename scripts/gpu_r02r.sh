#!/bin/bash
# round 2, visit R: new default integrator (variant 22) on configs[2] / configs[1], driver end to end on the reference's real static files
TAG=${1:-r02r}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== driver end to end (NA, 200 tracks)"
( time timeout 900 python scripts/run_driver_e2e.py NA --tracks 200 > $OUT/driver_e2e_NA.log 2>&1 ) 2>&1 | grep real; tail -3 $OUT/driver_e2e_NA.log | cut -c1-900
echo "== driver end to end (GL, 500 tracks)"
( time timeout 900 python scripts/run_driver_e2e.py GL --tracks 500 > $OUT/driver_e2e_GL.log 2>&1 ) 2>&1 | grep real; tail -2 $OUT/driver_e2e_GL.log | cut -c1-900
timeout 900 python bench.py --no-cpu > $OUT/bench.json 2> $OUT/bench.err; python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("bench: value %.3e e2e %.3e ms/step %.1f roofline.frac %.3f interp frac %.3f waves %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["frac"],d["roofline_interp"]["frac"],d["details"]["waves_per_step"]))
print({k:round(v,3) for k,v in d["details"]["kernel_share_of_step"].items()}, d["details"]["numa"])
PY
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log | cut -c1-300
