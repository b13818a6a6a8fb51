#!/bin/bash
# round 2, visit A: record-path variants of k_integrate (parity first, then the bench per variant)
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "integrate or run_year_na or env_interp" > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/pytest_gpu.log
for v in ${VARIANTS:-18 22 23 24 17 21}; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-interp --integ-variant $v > $OUT/bench_v$v.json 2> $OUT/bench_v$v.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_v$v.json"))
    print("variant $v value %.3e e2e %.3e ms/step %.2f integrate: avg %.3f ms share %.2f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["roofline"]["share_of_step"]), {k:round(v,3) for k,v in d["kernel_share_of_step"].items()})
except Exception as e:
    print("variant $v failed", e); print(open("$OUT/bench_v$v.err").read()[-2000:])
PY
done
