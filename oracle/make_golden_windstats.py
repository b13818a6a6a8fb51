"""Golden vectors of the monthly wind statistics: the reference's UNMODIFIED calc_wnd_stat
(track/env_wind.py:169-228) run over the NumPy-backed xarray stand-in of oracle/xr_shim.py.
Build container only (needs /root/reference):

    python oracle/make_golden_windstats.py      ->  tests/golden/ref_windstats.npz
"""
import datetime
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh                                  # noqa: E402
from oracle import xr_shim                                            # noqa: E402


def synth(n_time, levels, nlat, nlon, seed, nan_frac=0.0):
    rng = np.random.default_rng(seed)
    ua = (rng.normal(0, 8, (n_time, len(levels), nlat, nlon)) + 10 * np.cos(np.linspace(0, 3, nlat))[None, None, :, None]).astype(np.float32)
    va = (0.4 * ua + rng.normal(0, 5, ua.shape)).astype(np.float32)
    if nan_frac:
        ua[rng.random(ua.shape) < nan_frac] = np.nan
        va[rng.random(va.shape) < nan_frac] = np.nan
    return ua, va


def reference_stats(ref, ua, va, times, levels, units, dt):
    """calc_wnd_stat(ua, va, dt) of the reference on shim DataArrays; returns wnd_stats.data [14, lat, lon]."""
    lvl_key, lon_key, lat_key = (ref.env_wind.input.get_lvl_key(), ref.env_wind.input.get_lon_key(), ref.env_wind.input.get_lat_key())
    nlat, nlon = ua.shape[2:]
    coords = {"time": xr_shim.Coord(np.asarray(times, dtype="datetime64[ns]")), lvl_key: xr_shim.Coord(np.asarray(levels), units),
              lat_key: xr_shim.Coord(np.linspace(-10, 10, nlat)), lon_key: xr_shim.Coord(np.linspace(100, 130, nlon))}
    dims = ["time", lvl_key, lat_key, lon_key]
    A = xr_shim.DataArray(ua.astype(np.float64), dims, coords)
    B = xr_shim.DataArray(va.astype(np.float64), dims, coords)
    return np.asarray(ref.env_wind.calc_wnd_stat(A, B, dt).data)


CASES = {
    # the reference's own ERA5 input: 2 x daily samples (scripts/download_era5.py:135) -> never averaged per day
    "era5_2x_daily": dict(step_h=12, n=160, levels=[250, 850], units="hPa", nan_frac=0.0, month=(2001, 2)),
    # sampling interval longer than a day: the groupby("time.day") branch runs (one sample per day group)
    "five_daily": dict(step_h=120, n=40, levels=[85000, 50000, 25000], units="Pa", nan_frac=0.0, month=(2001, 3)),
    # missing samples: skipna reductions, pairwise-complete covariance
    "era5_with_nans": dict(step_h=12, n=130, levels=[1000, 850, 250], units="millibars", nan_frac=0.1, month=(2001, 1)),
}


def main():
    if not rh.available():
        raise SystemExit("reference tree not present")
    ref = rh.load_reference()
    xr_shim.install(ref.env_wind.xr)
    out = {}
    for name, c in CASES.items():
        t0 = datetime.datetime(2001, 1, 1)
        times = np.array([np.datetime64(t0 + datetime.timedelta(hours=c["step_h"] * k)) for k in range(c["n"])])
        ua, va = synth(c["n"], c["levels"], 6, 9, seed=len(name), nan_frac=c["nan_frac"])
        dt = datetime.datetime(c["month"][0], c["month"][1], 15)
        stats = reference_stats(ref, ua, va, times, c["levels"], c["units"], dt)
        out[name + "_ua"], out[name + "_va"] = ua, va
        out[name + "_times"] = times.astype("datetime64[s]").astype(np.int64)
        out[name + "_levels"] = np.asarray(c["levels"], dtype=np.float64)
        out[name + "_units"] = np.array(c["units"])
        out[name + "_month"] = np.asarray(c["month"])
        out[name + "_stats"] = stats
        print(name, stats.shape, "NaN stats:", int(np.isnan(stats).sum()))
    path = os.path.join(ROOT, "tests", "golden", "ref_windstats.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
