"""Per-basin track file with the reference's schema (util/compute.py:244-268).

The reference writes through xarray -> netCDF4/HDF5, neither of which exists in this image; the
same variables, dimensions and coordinates are written here as NetCDF-3 64-bit-offset with
``scipy.io.netcdf_file`` (readable by xarray with ``engine="scipy"`` and by netCDF4).  Strings
become fixed-width char arrays (``tc_basins[n_trk][2]``, ``basin[basin][2]``); NaN is the fill.
"""
import numpy as np
from scipy.io import netcdf_file

TRACK_VARS = (("lon_trks", "tc_lon", None), ("lat_trks", "tc_lat", None),
              ("u250_trks", "tc_env_wnds", 0), ("v250_trks", "tc_env_wnds", 1),
              ("u850_trks", "tc_env_wnds", 2), ("v850_trks", "tc_env_wnds", 3),
              ("v_trks", "tc_v", None), ("m_trks", "tc_m", None), ("vmax_trks", "tc_vmax", None))


def write_tracks(path, out):
    """out: the dict built by compute.run_downscaling."""
    n_trk, n_time = out["tc_lon"].shape
    basin_ids = list(out["basin_ids"])
    with netcdf_file(path, "w", version=2) as f:
        f.createDimension("n_trk", n_trk)
        f.createDimension("time", n_time)
        f.createDimension("year", len(out["years"]))
        f.createDimension("basin", len(basin_ids))
        f.createDimension("month", 12)
        f.createDimension("strlen", 2)

        def var(name, dtype, dims, data, fill=None):
            v = f.createVariable(name, dtype, dims)
            if fill is not None:
                v._FillValue = fill
            v[:] = data
            return v

        var("n_trk", "i4", ("n_trk",), np.arange(n_trk, dtype=np.int32))
        var("time", "f8", ("time",), out["ts_output"])
        var("year", "i4", ("year",), np.asarray(out["years"], dtype=np.int32))
        var("basin", "S1", ("basin", "strlen"), np.array([list(b.ljust(2)) for b in basin_ids], dtype="S1"))
        var("month", "i4", ("month",), np.arange(1, 13, dtype=np.int32))
        for name, key, comp in TRACK_VARS:
            data = out[key] if comp is None else out[key][:, :, comp]
            var(name, "f8", ("n_trk", "time"), data, fill=np.nan)
        var("tc_month", "f8", ("n_trk",), out["tc_months"], fill=np.nan)
        var("tc_basins", "S1", ("n_trk", "strlen"),
            np.array([list(str(b).ljust(2)) for b in out["tc_basins"]], dtype="S1").reshape(n_trk, 2))
        var("tc_years", "i4", ("n_trk",), np.asarray(out["tc_years"], dtype=np.int32))
        var("seeds_per_month", "f8", ("year", "basin", "month"), out["n_seeds"], fill=np.nan)


def read_tracks(path):
    """Reader for both flavours: the NetCDF-3 files written above and the NetCDF-4 / HDF5 files the reference writes
    through xarray (util/compute.py:263; the samples under notebooks/data) -- the latter through h5lite.py, strings
    arriving as variable-length strings.  Returns {variable: ndarray} with `tc_basins` / `basin` as 'U2' arrays."""
    from . import refdata
    out = {k: v for k, (v, _) in refdata.open_variables(path).items()}
    for k in ("tc_basins", "basin"):
        a = out[k]
        if a.dtype.kind == "S" and a.ndim == 2:                       # char array [n][strlen]
            out[k] = np.array([b"".join(row).decode().strip() for row in a], dtype="U2")
        else:
            out[k] = np.array([str(x).strip() for x in a.reshape(-1)], dtype="U2")
    return out
