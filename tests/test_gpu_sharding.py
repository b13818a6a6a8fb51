"""Within-year sharding (SURVEY 8e mode 2, tcr_set_shard): `world` engines run ONE year collectively -- every rank seeds
every attempt, integrates the attempts k with k % world == rank, and the ranks exchange kept flags / counted-seed
histograms once per wave -- and the merged result is bit-identical to the single-engine result, for any world.

The ranks here are threads of one process driving `world` handles on cuda:0, with an in-process all-reduce between
them, so the test runs on the single-GPU box of the round-end `pytest -m gpu`; the NCCL transport of the same callback
(gather.dist_allreduce) is exercised by scripts/run_sharded_year.py under torchrun on 2+ GPUs."""
import threading

import numpy as np
import pytest

from conftest import Case

pytestmark = pytest.mark.gpu


def _engine(case):
    from tropical_cyclone_risk_b200.engine import Engine
    e = Engine(case.p, device=0)
    e.upload_case(case.lon, case.lat, case.planes, case.static, case.mask_lon, case.mask_lat, case.mask_planes)
    return e


def run_world(case, world, ym_base, year_key, run_seed, n_tracks, tuning=None):
    import torch
    from tropical_cyclone_risk_b200 import gather
    dev = torch.device("cuda", 0)
    barrier = threading.Barrier(world)
    bufs = [None] * world
    calls = [0] * world

    def make_cb(rank):
        def cb(ptr, count, dtype, op, stream):
            torch.cuda.synchronize()
            bufs[rank] = gather.device_tensor(ptr, count, dtype, dev)
            calls[rank] += 1
            barrier.wait()
            if rank == 0:
                stack = torch.stack([b.to(torch.int64) for b in bufs])
                red = stack.min(0).values if op == 1 else stack.sum(0)
                for b in bufs:
                    b.copy_(red.to(b.dtype))
                torch.cuda.synchronize()
            barrier.wait()
        return cb

    engines = [_engine(case) for _ in range(world)]
    out, err = [None] * world, [None] * world

    def work(rank):
        try:
            e = engines[rank]
            if tuning:
                e.set_tuning(**tuning[rank % len(tuning)])
            e.set_shard(rank, world, make_cb(rank))
            out[rank] = e.run_years(ym_base, year_key, run_seed, n_tracks)
        except Exception as ex:                                   # a dead rank must not leave the others at the barrier
            err[rank] = ex
            barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(600)
    for e in engines:
        e.close()
    for ex in err:
        if ex is not None and not isinstance(ex, threading.BrokenBarrierError):
            raise ex
    assert all(o is not None for o in out), err
    assert len(set(calls)) == 1 and calls[0] > 0
    return out


def merge(outs):
    m = {}
    for k in ("lon", "lat", "v", "m", "vmax", "env"):
        m[k] = np.sum([o[k].view(np.int64) for o in outs], axis=0).view(np.float64)       # integer sum of the bit patterns
    for k in ("tc_month", "tc_basin", "n_seeds"):
        for o in outs[1:]:
            assert np.array_equal(o[k], outs[0][k], equal_nan=True), k                   # complete on every rank
        m[k] = outs[0][k]
    return m


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_year_equals_single_engine(na_year, world):
    one = _engine(na_year)
    try:
        want = one.run_years([0, 0], [2001, 2007], 31337, 40)
    finally:
        one.close()
    # ranks with DIFFERENT wave / slot limits: the waves are cut where the tightest rank cuts them
    tuning = [dict(max_wave=20000, oversub_permille=1300), dict(max_wave=1 << 40, max_slots=700)] if world == 2 else None
    outs = run_world(na_year, world, [0, 0], [2001, 2007], 31337, 40, tuning)
    got = merge(outs)
    for k in ("lon", "lat", "v", "m", "vmax", "env", "tc_month", "tc_basin", "n_seeds"):
        assert np.array_equal(got[k], want[k], equal_nan=True), k
    # every row came from exactly one rank, and the ranks shared the work
    owned = [(~np.all(o["lon"].view(np.int64) == 0, axis=-1)) for o in outs]
    assert np.array_equal(np.sum(owned, axis=0), np.ones_like(owned[0], dtype=int))
    assert all(ow.sum() > 0 for ow in owned)
    for y in range(2):
        for key in ("attempts", "counted_seeds", "n_kept"):                               # global counters
            assert all(o["stats"][y][key] == want["stats"][y][key] for o in outs), key
        for key in ("integrated", "storm_steps", "kept_steps", "rhs_evals"):               # shares: they add up
            assert sum(o["stats"][y][key] for o in outs) == want["stats"][y][key], key


def test_sharded_world1_is_the_plain_call(na_year):
    e = _engine(na_year)
    try:
        a = e.run_years([0], [2003], 5, 30)
        e.set_shard(0, 1)
        b = e.run_years([0], [2003], 5, 30)
    finally:
        e.close()
    for k in ("lon", "vmax", "env", "n_seeds"):
        assert np.array_equal(a[k], b[k], equal_nan=True)
