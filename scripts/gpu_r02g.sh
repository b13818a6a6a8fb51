#!/bin/bash
# round 2, visit G: within-year sharding tests, point-record layout A/B
TAG=${1:-r02g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 $OUT/pytest_gpu.log | cut -c1-400
run() {  # name, env, args
  env $2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-interp $3 > $OUT/bench_$1.json 2> $OUT/bench_$1.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$1.json"))
    print("$1: value %.3e e2e %.3e ms/step %.2f integrate avg %.3f ms share %.2f waves %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["roofline"]["share_of_step"],d["details"]["waves_per_step"]), {k:round(v,3) for k,v in d["details"]["kernel_share_of_step"].items()})
except Exception as e:
    print("$1 failed", e); print(open("$OUT/bench_$1.err").read()[-2000:])
PY
}
run cfg2_default "A=1" ""
run cfg2_point "A=1" "--integ-variant 26"
run cfg1_default "A=1" "--basin NA --years 10 --tracks 1000"
run cfg1_point "A=1" "--basin NA --years 10 --tracks 1000 --integ-variant 26"
run cfg5_default "A=1" "--basin WP --years 1 --tracks 50000 --interval 900"
run cfg5_point "A=1" "--basin WP --years 1 --tracks 50000 --interval 900 --integ-variant 26"
