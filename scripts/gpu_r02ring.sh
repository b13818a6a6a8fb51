#!/bin/bash
# round 2, visit "ring": Fourier ring inside the integrator (TCR_FTAB_RING=1) against full tables (=0), same box:
# GPU tests with the ring, then A/B on configs[1], configs[2] and a 900-s-output WP year; TCR_RING_CTA=0: requests served
# by the lane's own warp instead of the whole CTA
TAG=${1:-r04b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_ring.log 2>&1 ) 2>&1 | grep real; tail -4 $OUT/pytest_gpu_ring.log | cut -c1-600
( time TCR_RING_CTA=0 timeout 900 python -m pytest tests -m gpu -x -q -k "ring or run_year" > $OUT/pytest_gpu_ring_warp.log 2>&1 ) 2>&1 | grep real; tail -4 $OUT/pytest_gpu_ring_warp.log | cut -c1-600
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
run cfg1_ring_cta "TCR_FTAB_RING=1 TCR_RING_CTA=1" "--basin NA --years 10 --tracks 1000"
run cfg1_ring_warp "TCR_FTAB_RING=1 TCR_RING_CTA=0" "--basin NA --years 10 --tracks 1000"
run cfg1_tables "TCR_FTAB_RING=0" "--basin NA --years 10 --tracks 1000"
run cfg2_ring_cta "TCR_FTAB_RING=1 TCR_RING_CTA=1" ""
run cfg2_tables "TCR_FTAB_RING=0" ""
run wp900_ring_cta "TCR_FTAB_RING=1 TCR_RING_CTA=1" "--basin WP --years 1 --tracks 2000 --interval 900"
run wp900_tables "TCR_FTAB_RING=0" "--basin WP --years 1 --tracks 2000 --interval 900"
