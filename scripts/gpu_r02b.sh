#!/bin/bash
# round 2, visit B: full GPU test suite, PARK / occupancy variants of k_integrate, first configs[2] run
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "^wide|^cuda|passed|failed|Error" $OUT/pytest_gpu.log | cut -c1-300 | tail -25
for v in ${VARIANTS:-18 25 26 27 28}; do
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
for v in ${VARIANTS2:-18 27}; do
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu --no-interp --basin GL --years 40 --tracks 5000 --integ-variant $v > $OUT/bench_cfg2_v$v.json 2> $OUT/bench_cfg2_v$v.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_cfg2_v$v.json"))
    print("cfg2 v$v value %.3e e2e %.3e ms/step %.2f integrate: avg %.3f ms share %.2f waves %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["roofline"]["share_of_step"],d["waves_per_step"]), {k:round(v,3) for k,v in d["kernel_share_of_step"].items()})
except Exception as e:
    print("cfg2 failed", e); print(open("$OUT/bench_cfg2_v$v.err").read()[-3000:])
PY
done
