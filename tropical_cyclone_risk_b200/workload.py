"""Synthetic ERA5-shaped inputs of one run basin, prepared exactly as the host side of the
reference's ``run_tracks`` prepares its real ones (util/compute.py:66-121): the unit the
benchmark, the smoke test and the parity tests all share.

No oracle and no CUDA in here -- only NumPy (synth.py generators + fields.py preparation).
"""
import numpy as np

from . import fields, params, synth
from . import namelist as default_namelist


class Workload:
    """Prepared inputs of `basin` for `years` (one table per (year, month)).

    Attributes: namelist, basin, p (TcrParams), bounds, lon, lat (basin-cropped axes),
    planes float32 [n_ym][19][nlat][nlon] (ym = year_slot*12 + month_slot), static (cropped
    bathymetry / land), mask_lon, mask_lat, mask_planes uint8 [8][nlat_m][nlon_m]."""

    def __init__(self, basin, years, months=range(1, 13), full_res=False, roughness=1.0,
                 zero_cov_over_land=False, namelist=None, pinned_alloc=None):
        nl = namelist or default_namelist
        self.namelist = nl
        self.basin = basin
        self.years = list(years)
        self.months = list(months)
        self.p = params.params_from_namelist(nl, basin)
        self.bounds = params.basin_bounds(nl, basin)
        lon, lat = synth.era5_axes()
        olon, olat = synth.ocean_axes()
        self.planes = None
        i = 0
        for y in self.years:
            for mth in self.months:
                raw = synth.synth_month_raw(y, mth, lon, lat, roughness, zero_cov_over_land)
                mld, strat = synth.synth_ocean(olon, olat, mth)
                self.lon, self.lat, pl = fields.prepare_month(nl, self.bounds, lon, lat, raw, olon, olat, mld, strat)
                if self.planes is None:
                    shape = (len(self.years) * len(self.months),) + pl.shape
                    self.planes = pinned_alloc(shape, np.float32) if pinned_alloc else np.empty(shape, np.float32)
                self.planes[i] = pl
                i += 1
        st = synth.synth_static(full_res=full_res)
        self.static_global = st
        self.static = fields.prepare_static(self.bounds, st)
        self.mask_lon, self.mask_lat, self.mask_planes = fields.crop_masks(st, basin, self.bounds, self.p)

    @property
    def n_ym(self):
        return self.planes.shape[0]

    def upload(self, eng, tables=True):
        """Static grids, masks and (optionally) every monthly table into an Engine."""
        eng.upload_static(self.static)
        eng.upload_masks(self.mask_lon, self.mask_lat, self.mask_planes)
        eng.alloc_tables(self.n_ym, self.lon, self.lat)
        if tables:
            self.upload_tables(eng)
        eng.synchronize()

    def upload_tables(self, eng):
        eng.upload_months(0, self.planes)
