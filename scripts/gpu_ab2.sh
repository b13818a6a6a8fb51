#!/bin/bash
# A/B of environment switches:  bash scripts/gpu_ab2.sh <tag> "<name> <env> <bench args>" ...
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for spec in "$@"; do
  set -- $spec
  name=$1; envs=$2; shift 2
  env $(echo $envs | tr ',' ' ') timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-interp "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$name.json"))
    print("$name: value %.3e e2e %.3e ms/step %.2f integrate avg %.3f ms share %.2f waves %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["roofline"]["share_of_step"],d["details"]["waves_per_step"]), {k:round(v,3) for k,v in d["details"]["kernel_share_of_step"].items()})
except Exception as e:
    print("$name failed", e); print(open("$OUT/bench_$name.err").read()[-2000:])
PY
done
