#!/bin/bash
# round 2, final visit: the judged evidence with the last kernels -- GPU tests, smoke, both bench arms on the default
# (configs[2]) workload, configs[1], the 900-s WP shapes (Fourier ring on by default there), ncu launch list of the default
# command, ncu --set full of k_integrate (configs[1]), DRAM counters of k_integrate over whole steps of configs[2]
TAG=${1:-r02final2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
echo "== pytest -m gpu"
( time timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1 ) 2>&1 | grep real; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log | cut -c1-300
echo "== smoke"
( time timeout 600 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1 ) 2>&1 | grep real; tail -2 $OUT/smoke.log | cut -c1-300
echo "== bench (default = configs[2])"
( time timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err ) 2>&1 | grep real; echo "bench exit $?"
head -c 1500 $OUT/bench.json; echo; tail -3 $OUT/bench.err | cut -c1-400
( time timeout 900 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err ) 2>&1 | grep real
head -c 600 $OUT/bench_reference.json; echo
echo "== bench configs[1]"
timeout 600 python bench.py --basin NA --years 10 --tracks 1000 --no-cpu --no-interp > $OUT/bench_cfg1.json 2> $OUT/bench_cfg1.err
head -c 700 $OUT/bench_cfg1.json; echo
echo "== bench WP 900 s (2000 tracks; ring default) and configs[4] shape on one GPU (50000 tracks)"
timeout 600 python bench.py --basin WP --years 1 --tracks 2000 --interval 900 --no-cpu --no-interp > $OUT/bench_wp900_2000.json 2> $OUT/bench_wp900_2000.err
head -c 500 $OUT/bench_wp900_2000.json; echo
( time timeout 900 python bench.py --basin WP --years 1 --tracks 50000 --interval 900 --steps 2 --warmup 3 --no-cpu --no-interp > $OUT/bench_cfg4shape_n1.json 2> $OUT/bench_cfg4shape_n1.err ) 2>&1 | grep real
head -c 500 $OUT/bench_cfg4shape_n1.json; echo; tail -2 $OUT/bench_cfg4shape_n1.err | cut -c1-300
echo "== ncu launch list (default command, fewer steps)"
( time timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --interp-queries 4194304 > $OUT/ncu_bench.log 2>&1 ) 2>&1 | grep real
echo "== ncu full: k_integrate on configs[1]"
( time timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 4 -c 1 -f -o $OUT/prof_integrate \
    python bench.py --basin NA --years 10 --tracks 1000 --steps 2 --warmup 3 --no-cpu --no-interp > $OUT/ncu_integrate.log 2>&1 ) 2>&1 | grep real
[ -f $OUT/prof_integrate.ncu-rep ] && python scripts/ncu_summary.py $OUT/prof_integrate.ncu-rep 40 > $OUT/prof_integrate_summary.txt 2>&1
echo "== ncu DRAM counters: k_integrate over whole steps of configs[2]"
( time timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_integrate -s 24 -c 16 --csv --log-file $OUT/integrate_cfg2_dram.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-interp > $OUT/ncu_integrate_cfg2.log 2>&1 ) 2>&1 | grep real
ls -la $OUT | head -40
