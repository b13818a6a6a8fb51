#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list + full captures.
# Usage (from the repo root, on the B200 box):  bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
if [ -z "$ONLY_NCU" ]; then
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
echo "== smoke"
timeout 600 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log
tail -3 $OUT/smoke.log
echo "== bench"
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cat $OUT/bench.json; tail -5 $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
cat $OUT/bench_reference.json
fi
if [ -z "$NO_NCU" ]; then
echo "== ncu launch list (same command as the bench, fewer steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --interp-queries 4194304 > $OUT/ncu_bench.log 2>&1
echo "== ncu full: k_integrate (one-wave steps after the warm-up: launch 4)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 4 -c 1 -f -o $OUT/prof_integrate \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-interp > $OUT/ncu_integrate.log 2>&1
echo "== ncu full: env_interp"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_env_interp -s 4 -c 1 -f -o $OUT/prof_interp \
    python bench.py --steps 1 --warmup 1 --no-cpu --years 10 --tracks 20 > $OUT/ncu_interp.log 2>&1
echo "== ncu full: k_fourier_table, k_select, k_postprocess"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fourier_table|k_select|k_postprocess|k_seed|k_gather" -s 20 -c 5 -f -o $OUT/prof_others \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-interp > $OUT/ncu_others.log 2>&1
echo "== ncu full: pre-processing kernels (N3)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wind_stats -s 2 -c 1 -f -o $OUT/prof_windstats \
    python scripts/run_windstats_once.py > $OUT/ncu_windstats.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_thermo$ -s 2 -c 1 -f -o $OUT/prof_thermo \
    python scripts/run_thermo_once.py > $OUT/ncu_thermo.log 2>&1
# gpurun brings back at most 64 MiB: summarise every report here, keep only the three main .ncu-rep files
for r in integrate interp windstats thermo; do
    [ -f $OUT/prof_$r.ncu-rep ] && python scripts/ncu_summary.py $OUT/prof_$r.ncu-rep 40 > $OUT/prof_${r}_summary.txt 2>&1
done
[ -f $OUT/prof_others.ncu-rep ] && ncu -i $OUT/prof_others.ncu-rep --page raw --csv > $OUT/prof_others_raw.csv 2>/dev/null
rm -f $OUT/prof_others.ncu-rep $OUT/prof_interp.ncu-rep
fi
ls -la $OUT
