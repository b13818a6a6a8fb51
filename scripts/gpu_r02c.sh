#!/bin/bash
# round 2, visit C: the default bench (configs[2]) end to end, ncu full capture of k_integrate at configs[1] and configs[2]
TAG=${1:-r02c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== default bench"
( time timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err ) 2>&1 | tail -3
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench.json"))
    print({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk!='details'}) for k,v in d.items() if k not in ('details','roofline_poi','roofline_windstats','roofline_thermo')})
    print(d['details']['kernel_share_of_step'], d['details']['waves_per_step'])
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench.err").read()[-3000:])
PY
echo "== ncu full: k_integrate at configs[1] (comparable with round 1)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 4 -c 1 -f -o $OUT/prof_integrate_cfg1 \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-interp --basin NA --years 10 --tracks 1000 > $OUT/ncu_integrate_cfg1.log 2>&1
python scripts/ncu_summary.py $OUT/prof_integrate_cfg1.ncu-rep 60 > $OUT/prof_integrate_cfg1_summary.txt 2>&1
head -22 $OUT/prof_integrate_cfg1_summary.txt
echo "== ncu full: k_integrate at configs[2] (3rd wave of the 4th step)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 32 -c 1 -f -o $OUT/prof_integrate_cfg2 \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-interp > $OUT/ncu_integrate_cfg2.log 2>&1
python scripts/ncu_summary.py $OUT/prof_integrate_cfg2.ncu-rep 40 > $OUT/prof_integrate_cfg2_summary.txt 2>&1
head -22 $OUT/prof_integrate_cfg2_summary.txt
rm -f $OUT/prof_integrate_cfg2.ncu-rep
ls -la $OUT
