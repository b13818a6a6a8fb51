"""Input provider that reads the reference's own files -- without xarray / netCDF4 (h5lite.py reads the HDF5
flavour, scipy the NetCDF-3 one):

  * static inputs shipped in the reference tree: intensity/data/bathymetry.nc, land.nc (intensity/geo.py:9-33),
    mld_climatology.nc, strat_climatology.nc (intensity/ocean.py:11-60);
  * the basin masks the reference generates with scripts/generate_land_masks.py (boxes AND ocean on a 0.25-degree
    grid).  That script takes its land/sea flag from the `global_land_mask` package, which is not available here;
    the flag is taken from the reference's own intensity/data/land.nc instead (nearest 0.125-degree point), or from
    ready-made land/<basin>.nc files when the caller has them;
  * the two caches the reference's pre-processing writes: env_wnd_<prefix>_<dates>.nc (14 wind statistics,
    track/env_wind.py:104-158) and thermo_<prefix>_<dates>.nc (vmax, chi, rh_mid, thermo/calc_thermo.py:103-116),
    interpolated in time to the 15th of each month the way util/compute.py:107-112 and track/bam_track.py:88-91 do.

`compute.configure(inputs=ReferenceInputs(reference_root, env_wnd_file, thermo_file))` makes run_tracks /
run_downscaling work on exactly the data the reference would read.
"""
import datetime
import os
import re

import numpy as np

from . import fields, h5lite, layout, synth


# ---------------------------------------------------------------------------------------------
# file access: HDF5 (NetCDF-4) through h5lite, NetCDF-3 classic / 64-bit offset through scipy
# ---------------------------------------------------------------------------------------------
def open_variables(path, names=None):
    """{name: (ndarray, attrs dict)} of a NetCDF file's root group, whichever of the two container formats it is."""
    with open(path, "rb") as fh:
        magic = fh.read(8)
    out = {}
    if magic == h5lite.SIGNATURE:
        f = h5lite.File(path)
        for k in (names or f.keys()):
            o = f[k]
            if isinstance(o, h5lite.Dataset):
                out[k] = (o.read(), dict(o.attrs))
        return out
    if magic[:3] == b"CDF":
        from scipy.io import netcdf_file
        with netcdf_file(path, "r", mmap=False) as f:
            for k in (names or list(f.variables)):
                v = f.variables[k]
                attrs = {a: (b.decode() if isinstance(b, bytes) else b) for a, b in v._attributes.items()}
                a = np.array(v[:])
                if a.dtype.byteorder == ">":                               # NetCDF-3 is big-endian on disk
                    a = a.astype(a.dtype.newbyteorder("="))
                out[k] = (a, attrs)
        return out
    raise ValueError("%s is neither NetCDF-3 nor NetCDF-4/HDF5" % path)


_UNIT_SECONDS = {"second": 1.0, "seconds": 1.0, "minute": 60.0, "minutes": 60.0, "hour": 3600.0, "hours": 3600.0,
                 "day": 86400.0, "days": 86400.0}


def decode_cf_time(values, units, calendar="standard"):
    """CF time coordinate -> list of datetime.datetime ('<unit> since <date>[ time]'; standard / gregorian /
    proleptic_gregorian calendars -- what xarray writes for datetime64 data; util/input.py:123-133 converts the same
    way).  Non-standard calendars (noleap ...) are refused rather than mis-dated."""
    if calendar not in ("standard", "gregorian", "proleptic_gregorian"):
        raise NotImplementedError("calendar %r" % calendar)
    m = re.match(r"\s*(\w+)\s+since\s+(\d{1,4})-(\d{1,2})-(\d{1,2})(?:[T\s]+(\d{1,2}):(\d{1,2})(?::(\d{1,2})(?:\.\d*)?)?)?", str(units))
    if not m or m.group(1).lower() not in _UNIT_SECONDS:
        raise ValueError("cannot parse time units %r" % (units,))
    y, mo, d = int(m.group(2)), int(m.group(3)), int(m.group(4))
    hh, mi, ss = (int(g) if g else 0 for g in m.group(5, 6, 7))
    t0 = datetime.datetime(y, mo, d, hh, mi, ss)
    scale = _UNIT_SECONDS[m.group(1).lower()]
    return [t0 + datetime.timedelta(seconds=float(v) * scale) for v in np.asarray(values).reshape(-1)]


def time_weights(times, t):
    """Linear interpolation in time like xarray's .interp(time=t) (util/compute.py:109-112): (i0, i1, w1) with
    value = (1 - w1) x[i0] + w1 x[i1]; outside the record xarray gives NaN -- here the caller gets an error."""
    secs = np.array([(x - times[0]).total_seconds() for x in times])
    ts = (t - times[0]).total_seconds()
    if ts < secs[0] or ts > secs[-1]:
        raise ValueError("%s is outside the record %s .. %s" % (t, times[0], times[-1]))
    i1 = int(np.searchsorted(secs, ts, side="left"))
    if secs[i1] == ts:
        return i1, i1, 0.0
    i0 = i1 - 1
    return i0, i1, (ts - secs[i0]) / (secs[i1] - secs[i0])


# ---------------------------------------------------------------------------------------------
# static inputs
# ---------------------------------------------------------------------------------------------
def load_static(reference_root, mask_dir=None):
    """The dict fields.prepare_static / fields.mask_planes consume, from the reference's files."""
    data = os.path.join(reference_root, "intensity", "data")
    b = open_variables(os.path.join(data, "bathymetry.nc"), ("lon", "lat", "bathymetry"))
    l = open_variables(os.path.join(data, "land.nc"), ("lon", "lat", "land"))
    lon_l, lat_l = l["lon"][0].astype(np.float64), l["lat"][0].astype(np.float64)
    land = np.ascontiguousarray(l["land"][0], dtype=np.int8)
    if mask_dir and all(os.path.exists(os.path.join(mask_dir, "%s.nc" % k)) for k in layout.BASIN_IDS + ("GL",)):
        planes = {}
        for k in layout.BASIN_IDS + ("GL",):                          # util/compute.py:87-97
            v = open_variables(os.path.join(mask_dir, "%s.nc" % k), ("lon", "lat", "basin"))
            planes[k] = np.asarray(v["basin"][0]).astype(np.uint8)
            lon_m, lat_m = v["lon"][0].astype(np.float64), v["lat"][0].astype(np.float64)
        masks, gl = np.stack([planes[k] for k in layout.BASIN_IDS]), planes["GL"]
    else:
        def ocean(lon_m, lat_m):                                      # nearest point of the 0.125-degree land mask
            ix = np.clip(np.rint((lon_m - lon_l[0]) / (lon_l[1] - lon_l[0])).astype(int), 0, lon_l.size - 1)
            iy = np.clip(np.rint((lat_m - lat_l[0]) / (lat_l[1] - lat_l[0])).astype(int), 0, lat_l.size - 1)
            return land[np.ix_(iy, ix)] == 0
        lon_m, lat_m, masks, gl = synth.basin_masks(ocean)
    return dict(lon_b=b["lon"][0].astype(np.float64), lat_b=b["lat"][0].astype(np.float64),
                bathy=np.ascontiguousarray(b["bathymetry"][0], dtype=np.int16),
                lon_l=lon_l, lat_l=lat_l, land=land, lon_m=lon_m, lat_m=lat_m, masks=masks, mask_GL=gl)


def load_ocean_climatology(reference_root):
    """(lon [360], lat [180], mld [12][180][360], strat [12][180][360]): the Levitus climatologies with the wrap
    column dropped, as intensity/ocean.py:26,55 hands them to the setup loop (NaN over land, zero-filled later by
    util/compute.py:117-118)."""
    data = os.path.join(reference_root, "intensity", "data")
    m = open_variables(os.path.join(data, "mld_climatology.nc"), ("lon", "lat", "month", "mixed_layer"))
    s = open_variables(os.path.join(data, "strat_climatology.nc"), ("lon", "lat", "month", "strat"))
    lon, lat = m["lon"][0].astype(np.float64)[:-1], m["lat"][0].astype(np.float64)
    order = np.argsort(np.asarray(m["month"][0]).astype(int))
    mld = np.moveaxis(m["mixed_layer"][0][:, :-1, :], 2, 0)[order]
    strat = np.moveaxis(s["strat"][0][:, :-1, :], 2, 0)[order]
    return lon, lat, np.ascontiguousarray(mld), np.ascontiguousarray(strat)


# ---------------------------------------------------------------------------------------------
# the pre-processing caches
# ---------------------------------------------------------------------------------------------
class _Cache:
    """One cache file: its grid, its time axis and lazily read (time, lat, lon) variables."""

    def __init__(self, path, names):
        v = open_variables(path)
        missing = [n for n in names if n not in v]
        if missing:
            raise KeyError("%s lacks %s" % (path, ", ".join(missing)))
        self.vars = {n: v[n][0] for n in names}
        self.lon, self.lat = v["lon"][0].astype(np.float64), v["lat"][0].astype(np.float64)
        t, attrs = v["time"]
        if np.issubdtype(np.asarray(t).dtype, np.datetime64):
            self.times = np.asarray(t).astype("datetime64[s]").tolist()
        else:
            self.times = decode_cf_time(t, attrs.get("units"), attrs.get("calendar", "standard"))

    def at(self, name, t):
        i0, i1, w = time_weights(self.times, t)
        a = np.asarray(self.vars[name][i0], dtype=np.float64)
        if i0 == i1:
            return a
        b = np.asarray(self.vars[name][i1], dtype=np.float64)
        return a + w * (b - a)                                         # xarray / scipy interp1d: y0 + w (y1 - y0)


class ReferenceInputs:
    """compute.configure(inputs=ReferenceInputs(...)): static grids from the reference tree, monthly planes from
    the reference's env_wnd / thermo caches."""

    def __init__(self, reference_root, env_wnd_file=None, thermo_file=None, mask_dir=None):
        self.root = reference_root
        self.env_wnd_file, self.thermo_file = env_wnd_file, thermo_file
        self.mask_dir = mask_dir if mask_dir is not None else os.path.join(reference_root, "land")
        self._static = self._ocean = self._wnd = self._thermo = None

    def static(self):
        if self._static is None:
            self._static = load_static(self.root, self.mask_dir)
        return self._static

    def ocean(self):
        if self._ocean is None:
            self._ocean = load_ocean_climatology(self.root)
        return self._ocean

    def year_planes(self, namelist, bounds, year):
        """(lon_b, lat_b, planes [12][19][nlat][nlon] float32): the setup loop of run_tracks (util/compute.py:66-121)
        on the cache files."""
        if not (self.env_wnd_file and self.thermo_file):
            raise RuntimeError("ReferenceInputs needs the env_wnd_*.nc and thermo_*.nc cache files for the monthly tables")
        if self._wnd is None:
            self._wnd = _Cache(self.env_wnd_file, layout.FIELD_NAMES[:14])
            self._thermo = _Cache(self.thermo_file, ("vmax", "chi", "rh_mid"))
            if not (np.array_equal(self._wnd.lon, self._thermo.lon) and np.array_equal(self._wnd.lat, self._thermo.lat)):
                raise ValueError("the env_wnd and thermo caches are on different grids")
        olon, olat, mld, strat = self.ocean()
        months = []
        for k in range(12):
            t = datetime.datetime(year, k + 1, 15)                    # util/compute.py:108
            raw = {n: self._wnd.at(n, t) for n in layout.FIELD_NAMES[:14]}
            raw.update({n: self._thermo.at(n, t) for n in ("vmax", "chi", "rh_mid")})
            lon_b, lat_b, planes = fields.prepare_month(namelist, bounds, self._wnd.lon, self._wnd.lat, raw,
                                                        olon, olat, mld[k], strat[k])
            months.append(planes)
        return lon_b, lat_b, np.stack(months)
