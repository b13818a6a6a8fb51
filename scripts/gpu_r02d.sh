#!/bin/bash
# round 2, visit D: tests after the record-path default / ym-sorted queue / prefetch / shim; A/B of sort and prefetch at configs[2] and [1]
TAG=${1:-r02d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log | cut -c1-300
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
run cfg2_nosort "TCR_NO_SORT=1" ""
run cfg2_noprefetch "TCR_NO_PREFETCH=1" ""
run cfg2_neither "TCR_NO_SORT=1 TCR_NO_PREFETCH=1" ""
run cfg1_default "A=1" "--basin NA --years 10 --tracks 1000"
run cfg1_neither "TCR_NO_SORT=1 TCR_NO_PREFETCH=1" "--basin NA --years 10 --tracks 1000"
run cfg2_rec1 "A=1" "--integ-variant 22"
