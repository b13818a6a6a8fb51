"""CPU oracle of the pre-processing kernels -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may import this module.

wind_stats(): NumPy restatement of calc_wnd_stat (track/env_wind.py:169-228).  The reference
runs on xarray, which is NOT installed in this container and whose version environment.yml does
not pin -- **parity unpinned** against xarray's own arithmetic (the reference's function body itself IS
pinned: it runs unmodified over the NumPy-backed stand-in of oracle/xr_shim.py, fixtures in
tests/golden/ref_windstats.npz).  What is restated is
xarray's published reduction semantics:
  * ``.groupby("time.day").mean(dim="time")``  -> per day: nanmean (skipna default for floats)
  * ``.mean(dim)``                             -> nanmean
  * ``.var(dim)``  (ddof = 0)                  -> numpy.nanvar: mean of (x - nanmean)^2 over valid x
  * ``xr.cov(a, b, dim)`` (ddof = 1)           -> xarray/computation.py::_cov_corr:
        valid = a.notnull() & b.notnull(); a, b = a.where(valid), b.where(valid)
        cov = ((a - a.mean()) * (b - b.mean())).sum(skipna=True, min_count=1) / (valid.sum() - ddof)
and it is pinned against numpy.mean / numpy.var / numpy.cov on NaN-free input and against pandas'
groupby-mean / var / pairwise-complete DataFrame.cov on input with missing samples (tests/test_preproc.py).  Arithmetic: float32 samples widened to float64, every reduction a
sequential float64 sum in time order (what NumPy does for axis-0 reductions of a C-contiguous
array), so the CUDA kernel can be compared bit for bit.
"""
import numpy as np

N_STATS = 14
# lower triangle, row by row (env_wind.py:30-42, 215): (i, j) with j <= i
PAIRS = [(i, j) for i in range(4) for j in range(i + 1)]


def _nansum0(x):
    """Sequential sum over axis 0 with NaNs replaced by zero (numpy.nansum semantics)."""
    return np.add.reduce(np.where(np.isnan(x), 0.0, x), axis=0)


def daily_means(x, group_start):
    """x [n_time, n_pts] float32 -> [n_groups, n_pts] float64, nanmean of each day's samples."""
    x = np.asarray(x, dtype=np.float64)
    out = np.empty((len(group_start) - 1, x.shape[1]))
    for g in range(len(group_start) - 1):
        blk = x[group_start[g]:group_start[g + 1]]
        cnt = np.add.reduce(~np.isnan(blk), axis=0)
        with np.errstate(invalid="ignore", divide="ignore"):
            out[g] = np.where(cnt > 0, _nansum0(blk) / cnt, np.nan)
    return out


def wind_stats(series, group_start):
    """series: four arrays [n_time, n_pts] float32 in the order ua250, va250, ua850, va850
    (env_wind.py:200-201); group_start [n_groups + 1].  Returns [14, n_pts] float64 in the
    reference's order: 4 means, then var / cov over the lower triangle row by row."""
    dm = [daily_means(s, group_start) for s in series]
    n_pts = dm[0].shape[1]
    out = np.empty((N_STATS, n_pts))
    with np.errstate(invalid="ignore", divide="ignore"):
        for k, (i, j) in enumerate(PAIRS):
            valid = ~np.isnan(dm[i]) & ~np.isnan(dm[j])
            cnt = np.add.reduce(valid, axis=0).astype(np.float64)
            a = np.where(valid, dm[i], np.nan)
            b = np.where(valid, dm[j], np.nan)
            mi = _nansum0(a) / cnt
            mj = _nansum0(b) / cnt
            acc = _nansum0((a - mi) * (b - mj))
            if i == j:
                out[i] = mi                                   # .mean(dim)        env_wind.py:207
                out[4 + k] = acc / cnt                        # .var(dim), ddof 0  env_wind.py:211
            else:
                out[4 + k] = np.where(cnt > 0, acc / (cnt - 1.0), np.nan)   # xr.cov, ddof 1  env_wind.py:213
    return out


# ---------------------------------------------------------------------------------------------
# thermodynamic pre-processing: ctypes front end of preproc_oracle.c (thermo/thermo.py restated)
# ---------------------------------------------------------------------------------------------
def thermo(p_env, ta, hus, sst, psl, table, cecd, k_mid):
    """vmax (potential intensity), chi, rh_mid for every column -- one time sample of compute_thermo
    (thermo/calc_thermo.py:60-69).  ta, hus [nlev, n] float32 (lowest level first), p_env [nlev] Pa,
    sst / psl [n]; table = (p_look, s_look, T_lookup) of thermo/entropy_table.npz (select_thermo = 1) or
    (p_look, s_look, rt_look, T_lookup) of thermo/entropy_table_reversible.npz (select_thermo = 2)."""
    if len(table) == 4:
        return _thermo_rev(p_env, ta, hus, sst, psl, table, cecd, k_mid)
    import ctypes as C
    from oracle import tcr_oracle
    lib = tcr_oracle.lib()
    p_env = np.ascontiguousarray(p_env, dtype=np.float64)
    ta = np.ascontiguousarray(ta, dtype=np.float32)
    hus = np.ascontiguousarray(hus, dtype=np.float32)
    sst = np.ascontiguousarray(sst, dtype=np.float64).reshape(-1)
    psl = np.ascontiguousarray(psl, dtype=np.float64).reshape(-1)
    pl, sl, tl = (np.ascontiguousarray(a, dtype=np.float64) for a in table)
    nlev, n = p_env.size, sst.size
    ta = ta.reshape(nlev, n)
    hus = hus.reshape(nlev, n)
    out = [np.empty(n) for _ in range(3)]
    vp = C.c_void_p
    lib.orc_thermo.restype = None
    lib.orc_thermo.argtypes = [C.c_int64, C.c_int, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, C.c_double, C.c_int,
                               C.c_double] + [vp] * 3
    lib.orc_thermo(n, nlev, p_env.ctypes.data, ta.ctypes.data, hus.ctypes.data, sst.ctypes.data, psl.ctypes.data,
                   pl.size, sl.size, pl.ctypes.data, sl.ctypes.data, tl.ctypes.data, float(cecd), int(k_mid),
                   float(p_env[k_mid]), *[o.ctypes.data for o in out])
    return tuple(out)


def _thermo_rev(p_env, ta, hus, sst, psl, table, cecd, k_mid):
    import ctypes as C
    from oracle import tcr_oracle
    lib = tcr_oracle.lib()
    p_env = np.ascontiguousarray(p_env, dtype=np.float64)
    sst = np.ascontiguousarray(sst, dtype=np.float64).reshape(-1)
    psl = np.ascontiguousarray(psl, dtype=np.float64).reshape(-1)
    nlev, n = p_env.size, sst.size
    ta = np.ascontiguousarray(ta, dtype=np.float32).reshape(nlev, n)
    hus = np.ascontiguousarray(hus, dtype=np.float32).reshape(nlev, n)
    pl, sl, rl, tl = (np.ascontiguousarray(a, dtype=np.float64) for a in table)
    assert tl.shape == (pl.size, sl.size, rl.size)
    out = [np.empty(n) for _ in range(3)]
    vp = C.c_void_p
    lib.orc_thermo_rev.restype = None
    lib.orc_thermo_rev.argtypes = [C.c_int64, C.c_int, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.c_double,
                                   C.c_int, C.c_double] + [vp] * 3
    lib.orc_thermo_rev(n, nlev, p_env.ctypes.data, ta.ctypes.data, hus.ctypes.data, sst.ctypes.data, psl.ctypes.data,
                       pl.size, sl.size, rl.size, pl.ctypes.data, sl.ctypes.data, rl.ctypes.data, tl.ctypes.data, float(cecd),
                       int(k_mid), float(p_env[k_mid]), *[o.ctypes.data for o in out])
    return tuple(out)


def entropy_lookup3(p, s, r, table):
    """T(p, s, rt) from the reversible inversion table the way scipy.interpolate.interpn(method='linear',
    bounds_error=False, fill_value=nan) evaluates it (test hook of the restatement in preproc_oracle.c)."""
    import ctypes as C
    from oracle import tcr_oracle
    lib = tcr_oracle.lib()
    p, s, r = (np.ascontiguousarray(a, dtype=np.float64).reshape(-1) for a in (p, s, r))
    pl, sl, rl, tl = (np.ascontiguousarray(a, dtype=np.float64) for a in table)
    out = np.empty(p.size)
    vp = C.c_void_p
    lib.orc_entropy_lookup3.restype = None
    lib.orc_entropy_lookup3.argtypes = [C.c_int64, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    lib.orc_entropy_lookup3(p.size, p.ctypes.data, s.ctypes.data, r.ctypes.data, pl.size, sl.size, rl.size,
                            pl.ctypes.data, sl.ctypes.data, rl.ctypes.data, tl.ctypes.data, out.ctypes.data)
    return out
