"""Return-period reduction: oracle pinned to the notebook's own NumPy formulas
(notebooks/sample_analysis.ipynb cells 13-17, restated verbatim below), count all-reduce over gloo,
and -- on the GPU -- bit-exact parity of tcr_poi_vmax / tcr_exceedance with the oracle."""
import os
import socket
import warnings

import numpy as np
import pytest

from oracle import tcr_oracle as orc

MIAMI = (-80.1918, 25.7617)             # cell 13


def notebook_haversine(lon1, lat1, lon2, lat2):
    """cell 13, verbatim arithmetic."""
    lon1, lat1, lon2, lat2 = map(np.deg2rad, (lon1, lat1, lon2, lat2))
    dlon = lon2 - lon1
    dlat = lat2 - lat1
    a = (np.square(np.sin(dlat / 2)) + np.cos(lat1) * np.cos(lat2) * np.square(np.sin(dlon / 2)))
    c = 2 * np.arcsin(np.sqrt(a))
    r_earth = 6378000
    return (r_earth / 1000.) * c


def synthetic_tracks(n, ns=361, seed=0):
    """Random walks that start east of the point and drift across it; NaN-padded tails, 0-360 longitudes."""
    rng = np.random.default_rng(seed)
    lon = 360 + MIAMI[0] + 6.0 + np.cumsum(rng.normal(-0.05, 0.08, (n, ns)), axis=1) + rng.normal(0, 2, (n, 1))
    lat = MIAMI[1] - 3.0 + np.cumsum(rng.normal(0.02, 0.06, (n, ns)), axis=1) + rng.normal(0, 2, (n, 1))
    v = rng.uniform(5, 85, (n, 1)) * np.exp(-0.5 * ((np.arange(ns) - rng.uniform(50, 250, (n, 1))) / 60.0) ** 2) + 10
    n_time = rng.integers(1, ns + 1, n)
    pad = np.arange(ns)[None, :] >= n_time[:, None]
    for a in (lon, lat, v):
        a[pad] = np.nan
    return lon, lat, v


def test_oracle_matches_notebook_formulas():
    lon, lat, v = synthetic_tracks(3000)
    d = notebook_haversine(MIAMI[0], MIAMI[1], lon, lat)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = np.nanmax(np.where(d <= 100, v, np.nan), axis=1)          # .where(dists <= 100).max(dim='time')
    got = orc.poi_vmax(lon, lat, v, MIAMI[0], MIAMI[1])
    assert np.array_equal(got, want, equal_nan=True)
    assert 0 < np.isnan(got).sum() < got.size                              # both outcomes occur
    bins = np.arange(10, 81, 5)
    assert np.array_equal(orc.exceedance(got, bins), [np.sum(want >= b) for b in bins])


def test_return_period_arithmetic():
    from tropical_cyclone_risk_b200 import analysis
    rp = analysis.return_period([10, 2, 0], 100)
    assert rp[0] == 10.0 and rp[1] == 50.0 and np.isinf(rp[2])


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tropical_cyclone_risk_b200 import analysis
        lon, lat, v = synthetic_tracks(1001, seed=3)
        rows = np.arange(rank, 1001, world)                                  # rows sharded over ranks
        local = orc.poi_vmax(lon[rows], lat[rows], v[rows], MIAMI[0], MIAMI[1])
        counts = analysis.all_reduce_counts(orc.exceedance(local, analysis.DEFAULT_BINS))
        q.put((rank, counts))
    finally:
        dist.destroy_process_group()


def test_counts_all_reduce_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    lon, lat, v = synthetic_tracks(1001, seed=3)
    want = orc.exceedance(orc.poi_vmax(lon, lat, v, MIAMI[0], MIAMI[1]), np.arange(10, 81, 5))
    assert np.array_equal(got[0], want) and np.array_equal(got[1], want)


@pytest.mark.gpu
def test_gpu_poi_vmax_and_exceedance_bit_exact(na_year):
    from tropical_cyclone_risk_b200 import analysis
    from tropical_cyclone_risk_b200.engine import Engine
    eng = Engine(na_year.p, device=0)
    try:
        for n, seed in ((1, 1), (33, 2), (5000, 3)):
            lon, lat, v = synthetic_tracks(n, seed=seed)
            got = analysis.vmax_at_poi(eng, lon, lat, v, MIAMI[0], MIAMI[1])
            want = orc.poi_vmax(lon, lat, v, MIAMI[0], MIAMI[1])
            assert np.array_equal(got, want, equal_nan=True), n
            assert np.array_equal(eng.exceedance(got, analysis.DEFAULT_BINS), orc.exceedance(want, analysis.DEFAULT_BINS))
        # a different radius / point, 1441-step rows, and an all-NaN row
        lon, lat, v = synthetic_tracks(257, ns=1441, seed=4)
        lon[5], lat[5], v[5] = np.nan, np.nan, np.nan
        got = eng.poi_vmax(lon, lat, v, 279.0, 24.0, radius_km=250.0)
        assert np.array_equal(got, orc.poi_vmax(lon, lat, v, 279.0, 24.0, radius_km=250.0), equal_nan=True)
        assert eng.poi_vmax(lon[:0], lat[:0], v[:0], 279.0, 24.0).shape == (0,)
    finally:
        eng.close()


@pytest.mark.gpu
def test_gpu_return_period_on_generated_tracks(na_year):
    """End of the chain the reference stops at: generate a year of tracks, reduce them at a point."""
    from conftest import Case  # noqa: F401
    from tropical_cyclone_risk_b200 import analysis
    from tropical_cyclone_risk_b200.engine import Engine
    eng = Engine(na_year.p, device=0)
    try:
        eng.upload_case(na_year.lon, na_year.lat, na_year.planes, na_year.static, na_year.mask_lon, na_year.mask_lat,
                        na_year.mask_planes)
        r = eng.run_years([0], [2001], 9, 400)
        got = analysis.vmax_at_poi(eng, r["lon"][0], r["lat"][0], r["vmax"][0], 300.0, 25.0, radius_km=300.0)
        want = orc.poi_vmax(r["lon"][0], r["lat"][0], r["vmax"][0], 300.0, 25.0, radius_km=300.0)
        assert np.array_equal(got, want, equal_nan=True) and (~np.isnan(got)).sum() > 5
        counts = analysis.exceedance_counts(eng, got, reduce_over_ranks=False)
        assert np.array_equal(counts, orc.exceedance(want, analysis.DEFAULT_BINS)) and counts[0] >= counts[-1]
    finally:
        eng.close()
