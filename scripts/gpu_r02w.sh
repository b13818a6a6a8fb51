#!/bin/bash
# round 2, visit W (8 GPUs): the scaling bench at N = 8 and N = 2 (configs[3] sharded by years), configs[4] (one WP year,
# 50 000 tracks, 900 s output) sharded WITHIN the year over 8 GPUs
TAG=${1:-r02w}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/gpus.txt
for N in 8 2; do
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-interp > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err ) 2>&1 | grep real
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_n$N.json"))
    print("N=$N: value %.3e e2e %.3e ms/step %.1f (e2e %.1f) gather %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],d["e2e"]["ms_per_step"],d["details"].get("gather")))
except Exception as e:
    print("N=$N failed", e); print(open("$OUT/bench_n$N.err").read()[-1500:])
PY
done
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 scripts/run_sharded_year.py --steps 2 --warmup 1 > $OUT/sharded_cfg4_n8.json 2> $OUT/sharded_cfg4_n8.err ) 2>&1 | grep real
tail -c 1500 $OUT/sharded_cfg4_n8.json; tail -3 $OUT/sharded_cfg4_n8.err | cut -c1-300
