#!/bin/bash
# parity tests, then the bench for each integrate-kernel variant
TAG=${1:-var}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 $OUT/pytest_gpu.log
for v in ${VARIANTS:-1 2 3 4 5 6}; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-interp --integ-variant $v > $OUT/bench_v$v.json 2> $OUT/bench_v$v.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_v$v.json"))
    print("variant $v value %.3e e2e %.3e ms/step %.2f integrate: avg %.3f ms share %.2f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["roofline"]["share_of_step"]), {k:round(v,3) for k,v in d["kernel_share_of_step"].items()})
except Exception as e:
    print("variant $v failed", e); print(open("$OUT/bench_v$v.err").read()[-2000:])
PY
done
