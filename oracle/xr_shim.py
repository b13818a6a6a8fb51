"""A NumPy-backed stand-in for the handful of xarray calls calc_wnd_stat makes (track/env_wind.py:169-228)
-- TEST INFRASTRUCTURE, build container only.

xarray is not installed here, so the reference's own function cannot run on real DataArrays.  This shim lets
the UNMODIFIED function body run: it supplies `.sel`, `.groupby("time.day").mean`, `.mean`, `.var`, `xr.cov`
and the `xr.DataArray` constructor with xarray's published semantics (skipna reductions through numpy's
nan-functions; `xr.cov` as xarray/computation.py::_cov_corr).  What is pinned that way is the reference's
control flow -- month mask, the day-grouping condition, level selection, the order of the 14 statistics and
the ddof of each -- not xarray's own arithmetic (see oracle/preproc_oracle.py)."""
import warnings

import numpy as np


class Coord:
    def __init__(self, values, units=None):
        self.values = np.asarray(values)
        self.data = self.values
        self.units = units

    def __getitem__(self, i):
        return Coord(self.values[i], self.units)

    def __sub__(self, other):
        return Coord(self.values - other.values)

    def __array__(self, dtype=None, copy=None):
        return self.values if dtype is None else self.values.astype(dtype)

    def __len__(self):
        return len(self.values)


class DataArray:
    def __init__(self, data=None, dims=None, coords=None, units=None):
        self.data = np.asarray(data)
        self.dims = list(dims)
        self.coords = {}
        for k, v in (coords or {}).items():
            self.coords[k] = v if isinstance(v, Coord) else Coord(v[1] if isinstance(v, tuple) else v)
        self.shape = self.data.shape

    # -- what calc_wnd_stat touches --------------------------------------------------------------
    def __getitem__(self, key):
        return self.coords[key]

    def __len__(self):
        return self.data.shape[0]

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)

    def _without(self, dim):
        return [d for d in self.dims if d != dim], {k: v for k, v in self.coords.items() if k != dim}

    def sel(self, indexers=None, **kw):
        kw = dict(indexers or {}, **kw)
        out = self
        for dim, sel in kw.items():
            ax = out.dims.index(dim)
            sel_arr = np.asarray(sel)
            if sel_arr.dtype == bool:                                     # ua.sel(time = month_mask)
                data = np.compress(sel_arr, out.data, axis=ax)
                coords = dict(out.coords)
                coords[dim] = Coord(out.coords[dim].values[sel_arr], out.coords[dim].units)
                out = DataArray(data, out.dims, coords)
            else:                                                         # .sel({lvl_key: p}): exact label
                hit = np.flatnonzero(out.coords[dim].values == sel)
                if hit.size != 1:
                    raise KeyError(sel)
                dims, coords = out._without(dim)
                out = DataArray(np.take(out.data, hit[0], axis=ax), dims, coords)
        return out

    def groupby(self, key):
        assert key == "time.day"
        return _GroupByDay(self)

    def _reduce(self, fn, dim):
        ax = self.dims.index(dim)
        dims, coords = self._without(dim)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return DataArray(fn(self.data, axis=ax), dims, coords)

    def mean(self, dim):
        return self._reduce(np.nanmean, dim)                              # skipna=True for float data

    def var(self, dim):
        return self._reduce(np.nanvar, dim)                               # ddof = 0


class _GroupByDay:
    def __init__(self, da):
        self.da = da

    def mean(self, dim):
        assert dim == "time"
        da = self.da
        ax = da.dims.index("time")
        days = da.coords["time"].values.astype("datetime64[D]")
        dom = (days - days.astype("datetime64[M]")).astype(int) + 1
        uniq = np.unique(dom)                                             # groupby sorts the group labels
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            parts = [np.nanmean(np.compress(dom == d, da.data, axis=ax), axis=ax) for d in uniq]
        dims = ["day"] + [d for d in da.dims if d != "time"]
        coords = {k: v for k, v in da.coords.items() if k != "time"}
        coords["day"] = Coord(uniq)
        data = np.stack(parts, axis=0)
        order = ["day"] + [d for d in da.dims if d != "time"]
        assert order == dims
        return DataArray(data, dims, coords)


def cov(da_a, da_b, dim=None, ddof=1):
    """xarray.cov -> computation._cov_corr(method='cov')."""
    ax = da_a.dims.index(dim)
    a, b = np.asarray(da_a.data, dtype=np.float64), np.asarray(da_b.data, dtype=np.float64)
    valid = ~np.isnan(a) & ~np.isnan(b)
    a, b = np.where(valid, a, np.nan), np.where(valid, b, np.nan)
    valid_count = valid.sum(axis=ax) - ddof
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        da = a - np.nanmean(a, axis=ax, keepdims=True)
        db = b - np.nanmean(b, axis=ax, keepdims=True)
        prod = da * db
        s = np.where(valid.sum(axis=ax) >= 1, np.nansum(prod, axis=ax), np.nan)      # sum(skipna=True, min_count=1)
        out = s / valid_count
    dims, coords = da_a._without(dim)
    return DataArray(out, dims, coords)


def install(xr_module):
    """Put the stand-ins on the (stub) xarray module the reference imported as `xr`."""
    xr_module.cov = cov
    xr_module.DataArray = DataArray
