#!/bin/bash
# A/B of integrate variants on configs[1] and configs[2]:  bash scripts/gpu_ab.sh <tag> <variant> [<variant> ...]
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() {  # name, args
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-interp $2 > $OUT/bench_$1.json 2> $OUT/bench_$1.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$1.json"))
    print("$1: value %.3e e2e %.3e ms/step %.2f integrate avg %.3f ms share %.2f waves %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["roofline"]["avg_launch_ms"],d["roofline"]["share_of_step"],d["details"]["waves_per_step"]), {k:round(v,3) for k,v in d["details"]["kernel_share_of_step"].items()})
except Exception as e:
    print("$1 failed", e); print(open("$OUT/bench_$1.err").read()[-2000:])
PY
}
for v in "$@"; do
  run cfg1_v$v "--basin NA --years 10 --tracks 1000 --integ-variant $v"
done
if [ -z "$NO_CFG2" ]; then for v in "$@"; do run cfg2_v$v "--integ-variant $v"; done; fi
if [ -n "$PARITY" ]; then
  TCR_INTEG_VARIANT=$PARITY timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_v$PARITY.log 2>&1; echo "pytest(v$PARITY) exit $?"; tail -3 $OUT/pytest_gpu_v$PARITY.log | cut -c1-400
fi
if [ -n "$NCU" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 4 -c 1 -f -o $OUT/prof_integrate \
    python bench.py --basin NA --years 10 --tracks 1000 --steps 2 --warmup 3 --no-cpu --no-interp --integ-variant $NCU > $OUT/ncu_integrate.log 2>&1
python scripts/ncu_summary.py $OUT/prof_integrate.ncu-rep 50 > $OUT/prof_integrate_summary.txt 2>&1
fi
