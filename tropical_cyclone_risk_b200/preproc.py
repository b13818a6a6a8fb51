"""Pre-processing stages on the device ("next" row N3): host-side mirror of the
reference's function names for the two stages whose cost scales with the record length.

    calc_wnd_stat(engine, ua, va, times, levels, dt)      track/env_wind.py:169-228
    wind_mean_vector_names(), wind_cov_matrix_names()     track/env_wind.py:22-42

The arrays are plain NumPy (time, level, lat, lon) float32 blocks -- what ``ua.data`` / ``va.data`` of the
reference's xarray objects hold; no copy is made for the level selection (level-slice pointers with a
time stride go straight to ``tcr_wind_stats``).
"""
import datetime

import numpy as np

STEERING_LEVELS = (250, 850)              # namelist.steering_levels; env_wind.py:180-184 hard-codes the same two


def wind_mean_vector_names(p_lvls=STEERING_LEVELS):
    """env_wind.py:22-26."""
    return ["%s%s_Mean" % (w, p) for p in p_lvls for w in ("ua", "va")]


def wind_cov_matrix_names(p_lvls=STEERING_LEVELS):
    """env_wind.py:31-42, flattened row by row (the order of the 10 second-moment statistics)."""
    names = ["%s%s" % (w, p) for p in p_lvls for w in ("ua", "va")]
    return [names[i] + "_Var" if i == j else names[i] + "_" + names[j] + "_cov"
            for i in range(len(names)) for j in range(i + 1)]


def _as_datetimes(times):
    t = np.asarray(times)
    if np.issubdtype(t.dtype, np.datetime64):
        return t.astype("datetime64[s]").tolist()             # util/input.py:128
    return list(t)


def month_samples(times, dt, group_sub_daily=False):
    """Sample selection and day grouping of calc_wnd_stat (env_wind.py:170-196).

    Returns (idx, group_start): idx = the samples inside dt's month (the reference's month_mask),
    group_start [n_groups + 1] = day groups over idx.  The reference groups by day-of-month only
    when ``dt_step < 0``, i.e. when the sampling interval is LONGER than a day (its comment says
    "less"; the code is what runs) -- in which case every day holds one sample -- so its own
    2 x daily ERA5 input (scripts/download_era5.py:135) is never averaged per day.  That literal
    behaviour is the default; group_sub_daily=True is the comment's intent (daily means first)."""
    ts = _as_datetimes(times)
    lo = datetime.datetime(dt.year, dt.month, 1)
    hi = datetime.datetime(dt.year + 1, 1, 1) if dt.month == 12 else datetime.datetime(dt.year, dt.month + 1, 1)
    idx = np.array([k for k, t in enumerate(ts) if lo <= t < hi], dtype=np.int64)
    if idx.size == 0:
        raise ValueError("no samples in %04d-%02d" % (dt.year, dt.month))
    if np.any(np.diff(idx) != 1):
        raise ValueError("the month's samples must be contiguous in time")
    days = np.array([ts[k].day for k in idx])
    if np.any(np.diff(days) < 0):
        raise ValueError("samples must be in time order")
    step_s = (ts[1] - ts[0]).total_seconds() if len(ts) > 1 else 86400.0
    if (86400.0 - step_s) < 0 or group_sub_daily:            # env_wind.py:187-188 (see docstring)
        starts = np.flatnonzero(np.r_[True, np.diff(days) != 0])
        group_start = np.r_[starts, idx.size].astype(np.int32)
    else:
        group_start = np.arange(idx.size + 1, dtype=np.int32)
    return idx, group_start


def level_index(levels, units, p_hpa):
    """env_wind.py:180-184: 250 / 850 in hPa or 25000 / 85000 in Pa; exact match like .sel()."""
    want = p_hpa if units in ("millibars", "hPa") else p_hpa * 100
    hit = np.flatnonzero(np.asarray(levels) == want)
    if hit.size != 1:
        raise KeyError("level %s not found" % want)
    return int(hit[0])


def calc_wnd_stat(engine, ua, va, times, levels, dt, level_units="hPa", group_sub_daily=False):
    """Monthly mean and covariance of the environmental winds (env_wind.py:169-228).

    ua, va: (time, level, lat, lon) float32, C-contiguous.  Returns wnd_stats [14, lat, lon] float64:
    wind_mean_vector_names() then wind_cov_matrix_names()."""
    ua = np.ascontiguousarray(ua, dtype=np.float32)
    va = np.ascontiguousarray(va, dtype=np.float32)
    if ua.shape != va.shape or ua.ndim != 4:
        raise ValueError("ua, va must be (time, level, lat, lon) arrays of one shape")
    idx, group_start = month_samples(times, dt, group_sub_daily)
    iu, il = level_index(levels, level_units, 250), level_index(levels, level_units, 850)
    t0, t1 = int(idx[0]), int(idx[-1]) + 1
    out = engine.wind_stats(ua[t0:t1], va[t0:t1], iu, il, group_start)
    return out.reshape((14,) + ua.shape[2:])
