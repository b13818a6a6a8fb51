#!/usr/bin/env python
"""The reference's run.py flow end to end on a GPU, through tropical_cyclone_risk_b200.driver:

    python scripts/run_driver_e2e.py [BASIN] [--tracks N] [--out DIR]

  1. a synthetic ERA5-shaped input tree is written (NetCDF-3, packed int16 winds like classic ERA5 files): 2 x daily
     u / v at 250 / 850 hPa for one year on the 1-degree grid of scripts/download_era5.py, monthly T / q on the 28
     ERA5 levels, SST and surface pressure;
  2. driver.run(basin): wind statistics (tcr_wind_stats) and potential intensity / chi / rh (tcr_thermo_month) on the
     GPU -> env_wnd_*.nc / thermo_*.nc caches in the reference's schema -> run_downscaling on those caches and on the
     reference's REAL static files (bathymetry, land, Levitus mixed layer / stratification: baseline/_ref/intensity/data,
     staged from the reference checkout; basin masks derived from its land mask) -> tracks_<basin>_*.nc;
  3. the written track file is read back and checked against the reference's output schema (util/compute.py:244-268,
     notebooks/data/tracks_NA_era5_*.nc).
"""
import argparse
import datetime
import json
import os
import sys
import tempfile
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def write_nc(path, name, times_h, data, lat, lon, levels=None, packed=False, units=""):
    from scipy.io import netcdf_file
    with netcdf_file(path, "w", version=2) as f:
        f.createDimension("time", len(times_h)); f.createDimension("latitude", lat.size); f.createDimension("longitude", lon.size)
        t = f.createVariable("time", "i4", ("time",)); t.units = "hours since 1900-01-01 00:00:00.0"; t.calendar = "gregorian"
        t[:] = np.asarray(times_h, dtype=np.int32)
        f.createVariable("latitude", "f4", ("latitude",))[:] = lat
        f.createVariable("longitude", "f4", ("longitude",))[:] = lon
        dims = ("time", "latitude", "longitude")
        if levels is not None:
            f.createDimension("level", len(levels))
            lv = f.createVariable("level", "i4", ("level",)); lv.units = "millibars"; lv[:] = levels
            dims = ("time", "level", "latitude", "longitude")
        if packed:
            lo, hi = float(np.min(data)), float(np.max(data))
            sf = (hi - lo) / 65000.0
            ao = 0.5 * (hi + lo)
            v = f.createVariable(name, "i2", dims); v.scale_factor = sf; v.add_offset = ao; v._FillValue = np.int16(-32767); v.units = units
            v[:] = np.clip(np.rint((data - ao) / sf), -32766, 32767).astype(np.int16)
        else:
            v = f.createVariable(name, "f4", dims); v.units = units
            v[:] = data.astype(np.float32)


def build_tree(base, year):
    from tropical_cyclone_risk_b200 import synth_thermo
    os.makedirs(base, exist_ok=True)
    lat = np.linspace(90.0, -90.0, 181)                                   # ERA5 files run north to south
    lon = np.arange(0.0, 360.0, 1.0)
    lam, phi = np.deg2rad(lon)[None, :], np.deg2rad(lat)[:, None]
    h0 = (datetime.datetime(year, 1, 1) - datetime.datetime(1900, 1, 1)).total_seconds() / 3600.0
    n_days = 366 if year % 4 == 0 else 365
    th = h0 + 12 * np.arange(2 * (n_days + 5))                            # a few days into the next year
    rng = np.random.default_rng(year)
    doy = (th - h0) / 24.0
    season = np.cos(2 * np.pi * (doy - 228.0) / 365.0)[:, None, None]     # +1 in boreal late summer
    u850 = (-5.0 * np.cos(2.5 * phi) + 0.0 * lam)[None] + 0.0 * season
    u250 = u850 + 25.0 * np.sin(phi[None] - np.deg2rad(6.0) * season) ** 2 + 2.0 * np.sin(lam)[None]
    v250 = (2.0 * np.sin(2.0 * lam) * np.cos(phi))[None] + 0.0 * season
    v850 = (1.5 * np.cos(3.0 * lam) * np.cos(phi))[None] + 0.0 * season
    t0 = time.time()
    for name, upper, lower in (("u", u250, u850), ("v", v250, v850)):
        data = np.empty((th.size, 2, lat.size, lon.size), np.float32)
        data[:, 0] = upper + rng.normal(0.0, 5.0, (th.size, lat.size, lon.size))
        data[:, 1] = lower + 0.4 * (data[:, 0] - upper) + rng.normal(0.0, 3.0, (th.size, lat.size, lon.size))
        write_nc(os.path.join(base, "era5_%s_daily_%d.nc" % (name, year)), name, th, data, lat, lon, [250, 850], packed=True, units="m s**-1")
    # monthly thermodynamic inputs: 13 months (the 1st of each), 28 levels 70 .. 1000 hPa ascending like the ERA5 files
    tm = [h0 + 24 * (datetime.datetime(year + (m // 12), m % 12 + 1, 1) - datetime.datetime(year, 1, 1)).days for m in range(13)]
    p = synth_thermo.ERA5_LEVELS_HPA * 100.0
    z = 7500.0 * np.log(p[0] / p)[:, None, None]
    ta = np.empty((13, 28, lat.size, lon.size), np.float32)
    hus = np.empty_like(ta)
    sst = np.empty((13, lat.size, lon.size), np.float32)
    for k in range(13):
        s = np.cos(2 * np.pi * ((k % 12) + 1 - 8.5) / 12.0)
        sst_k = 300.5 - 30.0 * np.sin(phi - np.deg2rad(8.0 * s)) ** 2 + 0.8 * np.cos(2 * lam) + 0.0 * lam
        T0 = sst_k - 1.0
        ta_k = np.maximum(T0[None] - 6.5e-3 * z, 200.0) + rng.normal(0, 0.3, (28, lat.size, lon.size))
        rh = np.clip((0.8 + 0.08 * np.sin(3 * lam)[None]) * np.exp(-z / 6000.0), 0.01, 1.0)
        ta[k], hus[k], sst[k] = ta_k[::-1], (rh * synth_thermo.sat_q(ta_k, p[:, None, None]))[::-1], sst_k
    lv = (p / 100.0)[::-1].astype(int)
    write_nc(os.path.join(base, "era5_t_monthly.nc"), "t", tm, ta, lat, lon, lv, units="K")
    write_nc(os.path.join(base, "era5_q_monthly.nc"), "q", tm, hus, lat, lon, lv, units="kg kg**-1")
    write_nc(os.path.join(base, "era5_sst_monthly.nc"), "sst", tm, sst, lat, lon, units="K")
    write_nc(os.path.join(base, "era5_sp_monthly.nc"), "sp", tm, np.full(sst.shape, 101000.0) + 300.0 * np.cos(lam)[None], lat, lon, units="Pa")
    return time.time() - t0


def namelist_for(base, out, year, tracks):
    from tropical_cyclone_risk_b200 import namelist as nl
    cfg = types.SimpleNamespace(**{k: getattr(nl, k) for k in dir(nl) if not k.startswith("__")})
    cfg.base_directory, cfg.output_directory = base, out
    cfg.exp_name, cfg.exp_prefix, cfg.dataset_type = "e2e", "era5", "ERA5"
    cfg.var_keys = {'ERA5': {'sst': 'sst', 'mslp': 'sp', 'temp': 't', 'sp_hum': 'q', 'u': 'u', 'v': 'v',
                             'lvl': 'level', 'lon': 'longitude', 'lat': 'latitude'}}
    cfg.start_year, cfg.start_month, cfg.end_year, cfg.end_month = year, 1, year, 12
    cfg.tracks_per_year = tracks
    return cfg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("basin", nargs="?", default="NA")
    ap.add_argument("--tracks", type=int, default=200)
    ap.add_argument("--year", type=int, default=2001)
    ap.add_argument("--out", default=None)
    ap.add_argument("--reference-root", default=os.path.join(ROOT, "baseline", "_ref"))
    a = ap.parse_args()
    from tropical_cyclone_risk_b200 import driver, trackfile
    work = a.out or tempfile.mkdtemp(prefix="tcr_e2e_")
    base, out = os.path.join(work, "in"), os.path.join(work, "out")
    t_tree = build_tree(base, a.year)
    nl = namelist_for(base, out, a.year, a.tracks)
    driver.init_distributed()
    t0 = time.time()
    res = driver.run(a.basin, nl, a.reference_root)
    t_run = time.time() - t0
    if res is None:                                                      # not the writing rank
        return
    f = trackfile.read_tracks(res["fn_trk_out"])
    want = {"lon_trks", "lat_trks", "u250_trks", "v250_trks", "u850_trks", "v850_trks", "v_trks", "m_trks", "vmax_trks",
            "tc_month", "tc_basins", "tc_years", "seeds_per_month", "n_trk", "time", "year", "basin", "month"}
    assert set(f) == want, set(f) ^ want
    n = a.tracks
    assert f["lon_trks"].shape == (n, 361) and f["seeds_per_month"].shape == (1, 7, 12) and f["time"][-1] == 15 * 86400.0
    n_time = np.sum(~np.isnan(f["lon_trks"]), axis=1)
    assert n_time.min() >= 1 and (np.nanmax(f["vmax_trks"], axis=1) >= 18.0).all() and (np.nanmax(f["v_trks"], axis=1) >= 15.0).all()
    assert list(f["basin"]) == ["AU", "EP", "NA", "NI", "SI", "SP", "WP"] and set(f["tc_years"]) == {a.year}
    rep = dict(basin=a.basin, tracks=n, year=a.year, tree_s=round(t_tree, 1), run_s=round(t_run, 1), track_file=os.path.basename(res["fn_trk_out"]),
               mean_track_len=float(n_time.mean()), lmi_mean=float(np.nanmax(f["vmax_trks"], axis=1).mean()),
               seeds_per_year=float(f["seeds_per_month"].sum()), months=np.bincount(f["tc_month"].astype(int), minlength=13)[1:].tolist(),
               caches=sorted(os.listdir(out)))
    print("driver e2e ok:", json.dumps(rep))


if __name__ == "__main__":
    main()
