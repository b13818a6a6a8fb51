"""The reference's top-level driver on this stack: run.py, util/compute.py::compute_downscaling_inputs,
track/env_wind.py::gen_wind_mean_cov and thermo/calc_thermo.py::gen_thermo, with the reference's file discovery
(util/input.py:22-27), cache file names and schemas, but no xarray / dask / netCDF4:

    python -m tropical_cyclone_risk_b200.driver NA --namelist /path/to/namelist.py --reference-root /path/to/tropical_cyclone_risk

  1. compute_downscaling_inputs: the monthly wind statistics (tcr_wind_stats) and the thermodynamic fields
     (tcr_thermo_month) are computed on the GPU from the raw reanalysis / GCM files under namelist.base_directory
     and written as env_wnd_<prefix>_<dates>.nc / thermo_<prefix>_<dates>.nc -- skipped when the files exist, like
     the reference (env_wind.py:85-87, calc_thermo.py:80-81).  Written as NetCDF-3 64-bit offset (SciPy is the only
     NetCDF writer in this image); read back by refdata.ReferenceInputs, and by xarray's scipy engine.
  2. run_downscaling(basin) on those caches and the reference tree's static files.

Input files are read through refdata.open_variables (HDF5 via h5lite, NetCDF-3 via SciPy) with CF decoding
(scale_factor / add_offset / _FillValue / missing_value) the way xarray's default mask_and_scale does.
"""
import calendar
import datetime
import glob
import importlib.util
import os
import shutil
import sys
import time

import numpy as np

from . import preproc, refdata


# ---------------------------------------------------------------------------------------------
# files
# ---------------------------------------------------------------------------------------------
def glob_prefix(nl, var_prefix):
    """util/input.py:22-27."""
    fns = glob.glob('%s/**/*%s*.nc' % (nl.base_directory, nl.exp_prefix), recursive=True)
    fns_var = sorted([x for x in fns if '_%s_' % var_prefix in x])
    if len(fns_var) == 0:
        fns_var = sorted([x for x in fns if '%s_' % var_prefix in x])
    return fns_var


def get_bounding_times(nl):
    """util/input.py:135-139."""
    s_dt = datetime.datetime(nl.start_year, nl.start_month, 1)
    e_dt = datetime.datetime(nl.end_year, nl.end_month, calendar.monthrange(nl.end_year, nl.end_month)[1])
    return s_dt, e_dt


def get_env_wnd_fn(nl):
    """track/env_wind.py:13-17."""
    return '%s/env_wnd_%s_%d%02d_%d%02d.nc' % (nl.output_directory, nl.exp_prefix, nl.start_year, nl.start_month,
                                               nl.end_year, nl.end_month)


def get_fn_thermo(nl):
    """thermo/calc_thermo.py:17-21."""
    return '%s/thermo_%s_%d%02d_%d%02d.nc' % (nl.output_directory, nl.exp_prefix, nl.start_year, nl.start_month,
                                              nl.end_year, nl.end_month)


def cf_float_dtype(a, attrs):
    """The float type xarray's mask_and_scale decodes a packed variable to (xarray/coding/variables.py
    _choose_float_dtype): float32 only for float32 data and for integers of at most 16 bits WITHOUT an add_offset;
    an add_offset (classic packed ERA5) or a wider integer gives float64."""
    a = np.asarray(a)
    if a.dtype.kind == "f":
        return np.float32 if a.dtype.itemsize <= 4 else np.float64
    if a.dtype.itemsize <= 2 and attrs.get("add_offset") is None:
        return np.float32
    return np.float64


def decode_cf(a, attrs, dtype=None):
    """xarray's mask_and_scale: fill / missing values -> NaN, then raw * scale_factor + add_offset, in the float
    type xarray would pick (cf_float_dtype) unless `dtype` forces one."""
    a = np.asarray(a)
    if a.dtype.kind not in "iuf":
        return a
    dtype = np.dtype(dtype or cf_float_dtype(a, attrs)).type
    out = a.astype(dtype)
    for k in ("_FillValue", "missing_value"):
        if k in attrs and attrs[k] is not None:
            fv = np.asarray(attrs[k]).reshape(-1)
            if fv.size and not (fv.dtype.kind == "f" and np.isnan(fv[0])):
                out[a == fv[0].astype(a.dtype)] = np.nan
    sf, ao = attrs.get("scale_factor"), attrs.get("add_offset")
    if sf is not None:
        out = out * dtype(np.asarray(sf).reshape(-1)[0])
    if ao is not None:
        out = out + dtype(np.asarray(ao).reshape(-1)[0])
    return out


class _Source:
    """One raw input file: coordinates, decoded time axis, one data variable (time, [level,] lat, lon)."""

    def __init__(self, nl, path, var_key, with_levels):
        keys = nl.var_keys[nl.dataset_type]
        v = refdata.open_variables(path)
        self.path = path
        self.lon = np.asarray(v[keys['lon']][0], dtype=np.float64)
        self.lat = np.asarray(v[keys['lat']][0], dtype=np.float64)
        t, tattrs = v['time']
        self.times = refdata.decode_cf_time(t, tattrs.get('units'), tattrs.get('calendar', 'standard'))
        data, attrs = v[var_key]
        self.units = str(attrs.get('units', ''))
        # decoded the way xarray decodes (float64 for packed int16 + add_offset), then handed to the float32
        # kernels: the reference carries the float64 values into its statistics, so for packed inputs the planes
        # here differ from its own by the float32 rounding of each sample (~6e-8 relative); documented deviation
        self.data = np.asarray(decode_cf(data, attrs), dtype=np.float32)
        if with_levels:
            lv, lattrs = v[keys['lvl']]
            self.levels = np.asarray(lv, dtype=np.float64)
            self.level_units = str(lattrs.get('units', 'hPa'))


def _write_cache(path, times, lon, lat, variables):
    """(time, lat, lon) float64 variables + coordinates, time as days since 1900-01-01 (what xarray would decode).
    Written under a temporary name and moved into place: a reader never sees a half-written cache."""
    from scipy.io import netcdf_file
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    t0 = datetime.datetime(1900, 1, 1)
    tmp = "%s.tmp.%d" % (path, os.getpid())
    with netcdf_file(tmp, "w", version=2) as f:
        f.createDimension("time", len(times)); f.createDimension("lat", lat.size); f.createDimension("lon", lon.size)
        v = f.createVariable("time", "f8", ("time",))
        v.units = "days since 1900-01-01 00:00:00"
        v.calendar = "proleptic_gregorian"
        v[:] = [(t - t0).total_seconds() / 86400.0 for t in times]
        f.createVariable("lat", "f8", ("lat",))[:] = lat
        f.createVariable("lon", "f8", ("lon",))[:] = lon
        for name, data in variables.items():
            w = f.createVariable(name, "f8", ("time", "lat", "lon"))
            w._FillValue = np.nan
            w[:] = data
    os.replace(tmp, path)


def _barrier():
    """Every rank waits until rank 0 has written a cache file the next stage opens."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.barrier()
    except ImportError:
        pass


def _rank():
    rank, world, dev = preproc._rank_world()
    return rank, world, dev


# ---------------------------------------------------------------------------------------------
# gen_wind_mean_cov (track/env_wind.py:84-166)
# ---------------------------------------------------------------------------------------------
def month_stamps(nl, file_times):
    """The months one file contributes and the time stamps they are stored under (wnd_stat_wrapper,
    env_wind.py:138-150): the first stamp is max(start date, first sample) itself, the later ones are the 15th --
    a quirk of the reference that its later time interpolation inherits; reproduced, not corrected."""
    dt_start, dt_end = get_bounding_times(nl)
    dt_start = max([dt_start, file_times[0]])
    t_months = [dt_start]
    while t_months[-1] <= min([dt_end, file_times[-1]]):
        y, m = t_months[-1].year, t_months[-1].month
        t_months.append(datetime.datetime(y + 1, 1, 15) if m == 12 else datetime.datetime(y, m + 1, 15))
    return t_months[0:-1]


def _cache_exists(path):
    """os.path.exists as rank 0 sees it, agreed by every rank (a cache that appears between two ranks' checks
    must not split them into a computing and a skipping group: the gathers below are collectives)."""
    have = os.path.exists(path)
    try:
        import torch
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            flag = [have]
            dist.broadcast_object_list(flag, src=0)
            have = bool(flag[0])
    except ImportError:
        pass
    return have


def gen_wind_mean_cov(engine, nl, group_sub_daily=False):
    fn_out = get_env_wnd_fn(nl)
    if _cache_exists(fn_out):
        return fn_out
    keys = nl.var_keys[nl.dataset_type]
    fns_ua, fns_va = glob_prefix(nl, keys['u']), glob_prefix(nl, keys['v'])
    rank, world, dev = _rank()
    stamps, stats, lon, lat = [], [], None, None
    for fu, fv in zip(fns_ua, fns_va):                                 # env_wind.py:97-100
        su, sv = _Source(nl, fu, keys['u'], True), _Source(nl, fv, keys['v'], True)
        months = month_stamps(nl, su.times)
        if not months:
            continue
        lon, lat = su.lon, su.lat
        mine = preproc.shard_items(len(months), rank, world)
        local = np.empty((len(mine), 14) + su.data.shape[2:])
        for k, i in enumerate(mine):
            local[k] = preproc.calc_wnd_stat(engine, su.data, sv.data, su.times, su.levels, months[i], su.level_units,
                                             group_sub_daily)
        stats.append(preproc.gather_items(local, len(months), rank, world, dev))
        stamps += months
    if not stamps:
        raise RuntimeError("no wind files under %s match %r" % (nl.base_directory, nl.exp_prefix))
    if rank == 0:
        allst = np.concatenate(stats, axis=0)                          # [n_months, 14, lat, lon]
        names = preproc.wind_mean_vector_names() + preproc.wind_cov_matrix_names()
        _write_cache(fn_out, stamps, lon, lat, {n: allst[:, i] for i, n in enumerate(names)})
        print('Saved %s' % fn_out)
    _barrier()
    return fn_out


# ---------------------------------------------------------------------------------------------
# gen_thermo (thermo/calc_thermo.py:74-117)
# ---------------------------------------------------------------------------------------------
def _concat(sources):
    times = sum([s.times for s in sources], [])
    return times, np.concatenate([s.data for s in sources], axis=0)


def gen_thermo(engine, nl):
    fn_out = get_fn_thermo(nl)
    if _cache_exists(fn_out):
        return fn_out
    keys = nl.var_keys[nl.dataset_type]
    dt_start, dt_end = get_bounding_times(nl)
    load = lambda k, lv: [_Source(nl, p, keys[k], lv) for p in glob_prefix(nl, keys[k])]
    psl_s, sst_s, ta_s, hus_s = load('mslp', False), load('sst', False), load('temp', True), load('sp_hum', True)
    if not (psl_s and sst_s and ta_s and hus_s):
        raise RuntimeError("thermodynamic input files missing under %s" % nl.base_directory)
    t_psl, psl = _concat(psl_s)
    t_sst, sst = _concat(sst_s)
    t_ta, ta = _concat(ta_s)
    t_hus, hus = _concat(hus_s)
    sel = [i for i, t in enumerate(t_psl) if dt_start <= t <= dt_end]                  # calc_thermo.py:88-90
    if not sel:
        raise RuntimeError("no samples between %s and %s" % (dt_start, dt_end))

    def pick(times, data, what):
        idx = []
        for i in sel:
            if t_psl[i] not in times:
                raise RuntimeError("%s has no sample at %s" % (what, t_psl[i]))      # the reference assumes equal time axes (:76 TODO)
            idx.append(times.index(t_psl[i]))
        return data[idx]

    rank, world, dev = _rank()
    mine = [sel[i] for i in preproc.shard_items(len(sel), rank, world)]
    same_grid = (sst_s[0].lon.shape == ta_s[0].lon.shape and np.array_equal(sst_s[0].lon, ta_s[0].lon)
                 and np.array_equal(sst_s[0].lat, ta_s[0].lat))
    kw = {} if same_grid else dict(sst_lon=sst_s[0].lon, sst_lat=sst_s[0].lat, lon=ta_s[0].lon, lat=ta_s[0].lat)
    if mine:
        sub = lambda times, data, what: pick(times, data, what)[[sel.index(i) for i in mine]]
        loc = preproc.compute_thermo(engine, sub(t_sst, sst, "sst"), psl[mine], sub(t_ta, ta, "temperature"),
                                     sub(t_hus, hus, "specific humidity"), ta_s[0].levels, nl,
                                     level_units=ta_s[0].level_units, sst_units=sst_s[0].units, **kw)
        local = np.stack(loc, axis=1)
    else:
        local = np.empty((0, 3) + psl.shape[1:])
    g = preproc.gather_items(local, len(sel), rank, world, dev)
    if rank == 0:
        stamps = [datetime.datetime(t_psl[i].year, t_psl[i].month, 15) for i in sel]   # calc_thermo.py:98-101
        _write_cache(fn_out, stamps, psl_s[0].lon, psl_s[0].lat, dict(vmax=g[:, 0], chi=g[:, 1], rh_mid=g[:, 2]))
        print('Saved %s' % fn_out)
    _barrier()
    return fn_out


def compute_downscaling_inputs(engine, nl):
    """util/compute.py:24-35."""
    print('Computing monthly mean and variance of environmental wind...')
    s = time.time()
    gen_wind_mean_cov(engine, nl)
    print('Time Elapsed: %f s' % (time.time() - s))
    print('Computing thermodynamic variables...')
    s = time.time()
    gen_thermo(engine, nl)
    print('Time Elapsed: %f s' % (time.time() - s))


# ---------------------------------------------------------------------------------------------
# run.py
# ---------------------------------------------------------------------------------------------
def load_namelist(path):
    spec = importlib.util.spec_from_file_location("namelist", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_engine(nl, reference_root, device=0):
    """An engine for the pre-processing kernels, with the reference's entropy table loaded (thermo.py:274-278)."""
    from .engine import Engine
    from .params import params_from_namelist
    eng = Engine(params_from_namelist(nl, "GL"), device=device)
    if getattr(nl, "select_thermo", 1) == 2:                             # thermo.py:279-284
        eng.set_entropy_table_reversible(*preproc.load_entropy_table(os.path.join(reference_root, "thermo", "entropy_table_reversible.npz")))
    else:
        eng.set_entropy_table(*preproc.load_entropy_table(os.path.join(reference_root, "thermo", "entropy_table.npz")))
    return eng


def local_device():
    """This rank's GPU: LOCAL_RANK under torchrun (one process per GPU), else torch's current device."""
    import torch
    if "LOCAL_RANK" in os.environ:
        dev = int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(dev)
        return dev
    return torch.cuda.current_device()


def init_distributed():
    """Under torchrun: one process per GPU, NCCL over NVLink (the write-out all-gather of compute.run_downscaling)."""
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return False
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dev = local_device()
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    return True


def run(basin_id, nl, reference_root, namelist_path=None, engine=None, run_years_fn=None):
    """run.py:8-19.  `engine` / `run_years_fn` replace the CUDA engine in host-logic tests."""
    from . import compute
    f_base = '%s/%s/' % (nl.output_directory, nl.exp_name)
    rank, _, _ = _rank()
    if rank == 0:
        os.makedirs(f_base, exist_ok=True)
        print('Saving model output to %s' % f_base)
        if namelist_path:
            shutil.copyfile(namelist_path, '%s/namelist.py' % f_base)
    eng = engine if engine is not None else make_engine(nl, reference_root, device=local_device())
    try:
        compute_downscaling_inputs(eng, nl)
    finally:
        if engine is None:
            eng.close()
    compute.configure(namelist=nl, inputs=refdata.ReferenceInputs(reference_root, get_env_wnd_fn(nl), get_fn_thermo(nl)))
    print('Running tracks for basin %s...' % basin_id)
    return compute.run_downscaling(basin_id, run_years_fn=run_years_fn)


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="run.py of the reference on the B200 stack")
    ap.add_argument("basin")
    ap.add_argument("--namelist", required=True, help="the reference's namelist.py (or a copy)")
    ap.add_argument("--reference-root", default=None, help="reference checkout (static data, entropy table); default: the namelist's directory")
    a = ap.parse_args(argv)
    nl = load_namelist(a.namelist)
    init_distributed()
    run(a.basin, nl, a.reference_root or os.path.dirname(os.path.abspath(a.namelist)), a.namelist)


if __name__ == "__main__":
    main(sys.argv[1:])
