#!/bin/bash
# quick visit: GPU tests + default bench summary (+ configs[1])   usage: bash scripts/gpu_quick.sh <tag> [pytest -k expr]
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log | cut -c1-300
for cfg in "cfg2:" "cfg1:--basin NA --years 10 --tracks 1000" "cfg3rank:--years 5 --tracks 20000"; do
name=${cfg%%:*}; args=${cfg#*:}
timeout 900 python bench.py --no-cpu --no-interp $args > $OUT/bench_$name.json 2> $OUT/bench_$name.err; python - <<PY
import json
d=json.load(open("$OUT/bench_$name.json"))
print("$name: value %.3e e2e %.3e ms/step %.1f integrate %.3f ms roofline.frac %.3f waves %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["roofline"]["frac"],d["details"]["waves_per_step"]))
print({k:round(v,3) for k,v in d["details"]["kernel_share_of_step"].items()})
PY
done
