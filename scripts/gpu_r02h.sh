#!/bin/bash
# round 2, visit H: the judged evidence of the round — GPU tests, smoke, both bench arms on the default
# (configs[2]) workload, ncu launch list of the same command, ncu --set full of the roofline kernels.
TAG=${1:-r02h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu"
( time timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1 ) 2>&1 | grep real; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log | cut -c1-300
echo "== smoke"
timeout 600 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log
tail -2 $OUT/smoke.log | cut -c1-300
echo "== bench (default = configs[2])"
( time timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err ) 2>&1 | grep real; echo "bench exit $?"
head -c 1800 $OUT/bench.json; echo; tail -3 $OUT/bench.err | cut -c1-400
( time timeout 900 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err ) 2>&1 | grep real
head -c 900 $OUT/bench_reference.json; echo
echo "== bench configs[1] (round-1 headline, for comparison)"
timeout 900 python bench.py --basin NA --years 10 --tracks 1000 --no-cpu > $OUT/bench_cfg1.json 2> $OUT/bench_cfg1.err
head -c 1200 $OUT/bench_cfg1.json; echo
if [ -z "$NO_NCU" ]; then
echo "== ncu launch list (default command, fewer steps)"
( time timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --interp-queries 4194304 > $OUT/ncu_bench.log 2>&1 ) 2>&1 | grep real
echo "== ncu full: k_integrate on configs[1] (one wave per step; comparable with r01)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 4 -c 1 -f -o $OUT/prof_integrate \
    python bench.py --basin NA --years 10 --tracks 1000 --steps 2 --warmup 3 --no-cpu --no-interp > $OUT/ncu_integrate.log 2>&1
echo "== ncu full: env_interp"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_env_interp -s 4 -c 1 -f -o $OUT/prof_interp \
    python bench.py --basin NA --steps 1 --warmup 1 --no-cpu --years 10 --tracks 20 > $OUT/ncu_interp.log 2>&1
echo "== ncu full: front/back-end kernels of the step"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fourier|k_select|k_postprocess|k_seed|k_gather|k_coef" -s 20 -c 6 -f -o $OUT/prof_others \
    python bench.py --basin NA --years 10 --tracks 1000 --steps 2 --warmup 3 --no-cpu --no-interp > $OUT/ncu_others.log 2>&1
echo "== ncu full: return-period kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_poi_vmax -s 3 -c 1 -f -o $OUT/prof_poi \
    python bench.py --basin NA --years 1 --tracks 100 --steps 1 --warmup 1 --no-cpu > $OUT/ncu_poi.log 2>&1
for r in integrate interp poi; do
    [ -f $OUT/prof_$r.ncu-rep ] && python scripts/ncu_summary.py $OUT/prof_$r.ncu-rep 40 > $OUT/prof_${r}_summary.txt 2>&1
done
[ -f $OUT/prof_others.ncu-rep ] && ncu -i $OUT/prof_others.ncu-rep --page raw --csv > $OUT/prof_others_raw.csv 2>/dev/null
rm -f $OUT/prof_others.ncu-rep $OUT/prof_interp.ncu-rep $OUT/prof_poi.ncu-rep
fi
ls -la $OUT
