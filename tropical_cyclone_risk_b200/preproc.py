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


# ---------------------------------------------------------------------------------------------
# thermodynamics: thermo/calc_thermo.py:24-72 (compute_thermo) on the device
# ---------------------------------------------------------------------------------------------
def load_entropy_table(path):
    """thermo/entropy_table.npz of a reference checkout (thermo.py:274-278): (p_look, s_look, T_lookup); for
    thermo/entropy_table_reversible.npz (thermo.py:279-284): (p_look, s_look, rt_look, T_lookup)."""
    with np.load(path) as t:
        keys = ("p", "s", "rt", "T") if "rt" in t.files else ("p", "s", "T")
        return tuple(np.array(t[k], dtype=np.float64) for k in keys)


def order_levels(levels, level_units, ta, hus, p_midlevel_pa):
    """calc_thermo.py:50-59: lowest model level first, pressures in Pa, index of the level nearest p_midlevel.
    ta, hus are (level, lat, lon) or (time, level, lat, lon); the level axis is flipped as a view."""
    lvl = np.array(levels, dtype=np.float64)
    axis = ta.ndim - 3
    if lvl[0] - lvl[1] < 0:                                            # calc_thermo.py:51
        lvl = lvl[::-1]
        ta, hus = np.flip(ta, axis=axis), np.flip(hus, axis=axis)
    p_env = lvl * 100.0 if level_units in ("millibars", "hPa") else lvl.copy()
    k_mid = int(np.argmin(np.abs(p_env - p_midlevel_pa)))              # .sel(method='nearest')
    return p_env, k_mid, ta, hus


def compute_thermo(engine, sst, psl, ta, hus, levels, namelist, level_units="hPa", sst_units="K",
                   sst_lon=None, sst_lat=None, lon=None, lat=None):
    """compute_thermo (thermo/calc_thermo.py:24-72) for every time sample.

    sst (time, lat_s, lon_s), psl (time, lat, lon), ta / hus (time, level, lat, lon).  When the SST grid
    differs from the atmospheric grid pass both axis pairs: the SST is regridded like
    mat.interp_2d_grid(nan_to_num(sst)) (calc_thermo.py:38-40).  Returns (vmax, chi, rh_mid), each
    (time, lat, lon) float64.  engine.set_entropy_table (namelist.select_thermo = 1) or
    engine.set_entropy_table_reversible (select_thermo = 2) must have been called.  select_interp = 1 (BFGS inversion of
    the entropy, thermo.py:223-233) exists only in the reference's scalar CAPE_PI, which nothing calls: CAPE_PI_vectorized,
    the function calc_thermo.py:61 runs, reads the look-up table whatever select_interp says (and fails with a NameError
    when it is not 2); it is refused here."""
    from . import fields
    if namelist.select_thermo not in (1, 2):
        raise ValueError("namelist.select_thermo must be 1 (pseudoadiabatic) or 2 (reversible)")
    if namelist.select_interp != 2:
        raise NotImplementedError("select_interp = 1: CAPE_PI_vectorized (thermo.py:266) has no computed-inversion branch either")
    sst, psl = np.asarray(sst), np.asarray(psl)
    p_env, k_mid, ta, hus = order_levels(levels, level_units, np.asarray(ta), np.asarray(hus), float(namelist.p_midlevel))
    n_time = psl.shape[0]
    out = [np.zeros(psl.shape) for _ in range(3)]                      # calc_thermo.py:33-35
    for i in range(n_time):
        s = np.nan_to_num(np.asarray(sst[i], dtype=np.float64))        # calc_thermo.py:39
        if sst_lon is not None:
            s = fields.regrid(sst_lon, sst_lat, s, lon, lat)
        if "C" in sst_units:                                           # calc_thermo.py:41-42
            s = s + 273.15
        v, c, r = engine.thermo_month(p_env, ta[i], hus[i], s, psl[i], namelist.Ck / namelist.Cd, k_mid, namelist.select_thermo)
        out[0][i], out[1][i], out[2][i] = v, c, r
    return tuple(out)



# ---------------------------------------------------------------------------------------------
# drivers over the whole record: months / time samples shard over torch.distributed ranks the way the
# reference spreads them over dask workers (env_wind.py:97-101, calc_thermo.py:88-93); no data-path
# collective, one all-gather of the finished statistics.
# ---------------------------------------------------------------------------------------------
def shard_items(n, rank, world):
    """Round-robin item indices of this rank."""
    return list(range(rank, n, world))


def gather_items(local, n, rank, world, device=None):
    """local: float64 [len(shard_items(n, rank, world)), ...] -> [n, ...] in item order on every rank
    (equal-sized slots, unused ones NaN; NCCL on GPUs, gloo on CPU)."""
    local = np.ascontiguousarray(local, dtype=np.float64)
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    n_slots = (n + world - 1) // world
    item_shape = local.shape[1:]
    block = np.full((n_slots,) + item_shape, np.nan)
    block[:local.shape[0]] = local
    t = torch.from_numpy(block.reshape(n_slots, -1))
    if device is not None:
        t = t.to(device, non_blocking=True)
    g = torch.empty((world * n_slots, t.shape[1]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(g, t)
    allb = g.cpu().numpy().reshape((world, n_slots) + item_shape)
    return np.stack([allb[i % world, i // world] for i in range(n)])


def _rank_world():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dev = None
            if dist.get_backend() == "nccl":
                import torch
                dev = torch.device("cuda", torch.cuda.current_device())
            return dist.get_rank(), dist.get_world_size(), dev
    except ImportError:
        pass
    return 0, 1, None


def gen_wind_mean_cov(engine, ua, va, times, levels, months, level_units="hPa", group_sub_daily=False):
    """gen_wind_mean_cov / wnd_stat_wrapper (env_wind.py:84-166) over a list of months (datetimes): returns
    wnd_stats [n_months, 14, lat, lon]; months are sharded over the ranks of torch.distributed."""
    rank, world, dev = _rank_world()
    mine = shard_items(len(months), rank, world)
    shape = (14,) + tuple(np.shape(ua)[2:])
    local = np.empty((len(mine),) + shape)
    for k, i in enumerate(mine):
        local[k] = calc_wnd_stat(engine, ua, va, times, levels, months[i], level_units, group_sub_daily)
    return gather_items(local, len(months), rank, world, dev)


def gen_thermo(engine, sst, psl, ta, hus, levels, namelist, **kw):
    """gen_thermo (calc_thermo.py:74-117) without the file I/O: compute_thermo on this rank's time samples,
    all-gathered; returns (vmax, chi, rh_mid), each (time, lat, lon)."""
    rank, world, dev = _rank_world()
    n = np.shape(psl)[0]
    mine = shard_items(n, rank, world)
    if mine:
        loc = compute_thermo(engine, np.asarray(sst)[mine], np.asarray(psl)[mine], np.asarray(ta)[mine], np.asarray(hus)[mine],
                             levels, namelist, **kw)
        local = np.stack(loc, axis=1)                                   # [n_mine, 3, lat, lon]
    else:
        local = np.empty((0, 3) + tuple(np.shape(psl)[1:]))
    g = gather_items(local, n, rank, world, dev)
    return g[:, 0], g[:, 1], g[:, 2]
