#!/usr/bin/env python
"""Within-year sharding over real GPUs (SURVEY 8e mode 2; BASELINE configs[4] shape): one WP year, tracks_per_year
tracks, output_interval_s = 900, split over the ranks of torchrun by seed-attempt index, NCCL for the per-wave exchange
and for the write-out merge.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        scripts/run_sharded_year.py [--basin WP] [--tracks 50000] [--interval 900] [--steps 3]
    python scripts/run_sharded_year.py ...            # one GPU, no sharding: the reference point

Prints one JSON line from rank 0: storm-steps/s of the whole job (max-over-ranks device time), and a CRC of the merged
9-tuple -- identical for every N (the result does not depend on the number of GPUs)."""
import argparse
import json
import os
import sys
import types
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--basin", default="WP")
    ap.add_argument("--tracks", type=int, default=50000)
    ap.add_argument("--interval", type=int, default=900)
    ap.add_argument("--year", type=int, default=2001)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from tropical_cyclone_risk_b200 import gather as tgather
    from tropical_cyclone_risk_b200 import namelist as nl
    from tropical_cyclone_risk_b200.engine import Engine, PinnedPool
    from tropical_cyclone_risk_b200.pipeline import _Block
    from tropical_cyclone_risk_b200.workload import Workload
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    cfg = types.SimpleNamespace(**{k: getattr(nl, k) for k in dir(nl) if not k.startswith("__")})
    cfg.output_interval_s = a.interval
    wl = Workload(a.basin, [a.year], full_res=True, namelist=cfg, pinned_alloc=PinnedPool.empty)
    eng = Engine(wl.p, device=local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    wl.upload(eng)
    if world > 1:
        eng.set_shard(rank, world, tgather.dist_allreduce(dev))
    blk = _Block(torch, dev, 1, a.tracks, eng.n_steps, pinned=False)
    n_track_words = int(blk.offsets[6])
    ym_base, key = np.zeros(1, np.int32), np.asarray([a.year], np.int32)

    def step(i):
        st = eng.run_years_dev(ym_base, key, 20260101 + i, a.tracks, blk.dptr)
        if world > 1:                                             # write-out merge: integer sum of the bit patterns
            dist.all_reduce(blk.dev[:n_track_words].view(torch.int64), op=dist.ReduceOp.SUM)
        return st[0]

    for i in range(a.warmup):
        step(100 + i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    steps = 0
    for i in range(a.steps):
        steps += step(i)["storm_steps"]
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = step(0)                                                 # a fixed seed for the checksum
    torch.cuda.synchronize()
    v = torch.tensor([ms, float(steps), float(st["attempts"]), float(st["n_waves"])], dtype=torch.float64, device=dev)
    mx, sm = v.clone(), v.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    if rank == 0:
        host = blk.dev.cpu().numpy()
        views = blk.views_of(host)
        crc = 0
        for k in ("lon", "lat", "v", "m", "vmax", "env", "tc_month", "tc_basin", "n_seeds"):
            crc = zlib.crc32(np.ascontiguousarray(views[k]).tobytes(), crc)
        n_time = np.sum(~np.isnan(views["lon"][0]), axis=1)
        print(json.dumps({"metric": "storm-steps/sec (ensemble x timesteps)", "value": float(sm[1]) / (float(mx[0]) * 1e-3),
                          "n_gpus": world, "steps": a.steps, "ms_per_step": float(mx[0]) / a.steps,
                          "config": {"workload": "within-year sharding: %s basin, 1 year x %d tracks, %d output steps" % (a.basin, a.tracks, eng.n_steps)},
                          "attempts": int(st["attempts"]), "waves": int(st["n_waves"]), "kept_rows_complete": bool((n_time >= 1).all()),
                          "result_crc32": crc}))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
