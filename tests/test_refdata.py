"""Reading the reference's own files without xarray / netCDF4: the HDF5 reader (h5lite) on the files shipped
in the reference tree (build container only -- /root/reference does not travel), the NetCDF-3 path and the
cache -> planes provider on files written here."""
import datetime
import os

import numpy as np
import pytest

from oracle import ref_harness as rh

REF = rh.REF_ROOT
needs_ref = pytest.mark.skipif(not rh.available(), reason="reference tree not present")


@needs_ref
def test_h5lite_reads_the_static_inputs():
    from tropical_cyclone_risk_b200 import refdata
    st = refdata.load_static(REF, mask_dir="/nonexistent")
    assert st["bathy"].shape == (1350, 2700) and st["bathy"].dtype == np.int16        # intensity/geo.py:15
    assert st["land"].shape == (1440, 2880) and st["land"].dtype == np.int8           # intensity/geo.py:29
    assert (st["bathy"].min(), st["bathy"].max()) == (-10806, 6874)
    assert set(np.unique(st["land"])) == {0, 1}
    for ax in ("lon_b", "lat_b", "lon_l", "lat_l"):
        assert np.all(np.diff(st[ax]) > 0)
    # the planet: 29 % land by area, mean ocean depth 3.4 km; bathymetry and land mask agree on which is which
    w = np.cos(np.deg2rad(st["lat_l"]))[:, None]
    assert abs((st["land"] * w).sum() / (w.sum() * st["land"].shape[1]) - 0.2916) < 2e-3
    assert abs(st["bathy"][st["bathy"] < 0].mean() + 3437) < 5
    iy = np.searchsorted(st["lat_b"], st["lat_l"][::8]).clip(0, 1349)
    ix = np.searchsorted(st["lon_b"], st["lon_l"][::8]).clip(0, 2699)
    agree = ((st["bathy"][np.ix_(iy, ix)] >= 0) == (st["land"][::8, ::8] == 1)).mean()
    assert agree > 0.97
    # basin masks: boxes of scripts/generate_land_masks.py AND ocean (from land.nc)
    assert st["masks"].shape == (7, 721, 1440) and st["mask_GL"].shape == (721, 1440)
    na = st["masks"][2]
    assert na[:, :1020].sum() == 0 and 40000 < na.sum() < 80000                       # NA lives east of 255 E
    at = lambda lon, lat: (int(round((lat + 90) * 4)), int(round(lon * 4)))
    assert na[at(300.0, 20.0)] == 1 and na[at(279.0, 27.0)] == 0 and st["masks"][:, at(260.0, 40.0)[0], at(260.0, 40.0)[1]].sum() == 0
    assert st["masks"][6][at(140.0, 15.0)] == 1 and st["masks"][1][at(250.0, 15.0)] == 1     # WP, EP open ocean


@needs_ref
def test_h5lite_reads_the_ocean_climatologies():
    from tropical_cyclone_risk_b200 import refdata
    lon, lat, mld, strat = refdata.load_ocean_climatology(REF)
    assert lon.shape == (360,) and lat.shape == (180,) and mld.shape == (12, 180, 360) == strat.shape
    assert lon[0] == 0.0 and lon[-1] == 359.0 and np.all(np.diff(lat) > 0)
    assert np.isnan(mld).mean() > 0.2 and np.nanmin(mld) >= 0 and np.nanmax(mld) < 200.0      # Levitus MLD, NaN over land
    assert np.nanmin(strat) < 0 < np.nanmax(strat)                                              # negative stratification exists
    # northern winter mixed layers are deeper than summer ones in the North Atlantic
    box = (slice(None), slice(125, 145), slice(310, 340))
    assert np.nanmax(mld[1][box[1:]]) > 150 > 60 > np.nanmax(mld[7][box[1:]])


@needs_ref
def test_reference_sample_track_file_has_the_schema_we_write(tmp_path):
    """N2 parity against a REAL output file of the reference (notebooks/data, written by xarray through netCDF4:
    superblock 2, dense link storage in a fractal heap, variable-length strings)."""
    import types
    from tropical_cyclone_risk_b200 import compute, refdata, trackfile
    from tropical_cyclone_risk_b200 import namelist as nl
    from test_host_logic import _fake_run_years
    path = os.path.join(REF, "notebooks", "data", "tracks_NA_era5_197901_202312.nc")
    ref = trackfile.read_tracks(path)
    raw = refdata.open_variables(path)
    cfg = types.SimpleNamespace(**{k: getattr(nl, k) for k in dir(nl) if not k.startswith("__")})
    cfg.start_year, cfg.end_year, cfg.tracks_per_year = 2001, 2002, 4
    cfg.base_directory = cfg.output_directory = str(tmp_path)
    compute.configure(namelist=cfg)
    try:
        out = compute.run_downscaling("NA", write=True, run_years_fn=_fake_run_years)
    finally:
        compute.configure(namelist=nl)
    ours = trackfile.read_tracks(out["fn_trk"] if "fn_trk" in out else compute.get_fn_tracks(compute.TC_Basin("NA"), cfg))
    assert set(ours) == set(ref)                                                      # same variables ...
    for k in ref:                                                                     # ... of the same rank
        assert ours[k].ndim == ref[k].ndim, k
    n_trk, n_time = ref["lon_trks"].shape
    assert n_time == 361 and ref["time"][-1] == 1296000.0 and ref["seeds_per_month"].shape[1:] == (7, 12)
    assert list(ref["basin"]) == ["AU", "EP", "NA", "NI", "SI", "SP", "WP"] == list(ours["basin"])
    assert set(ref["tc_basins"]) == {"NA"} and ref["tc_years"].min() == 1979
    for k in ("lon_trks", "vmax_trks", "tc_month", "seeds_per_month", "time"):
        assert ref[k].dtype == np.float64 == ours[k].dtype, k
        assert "_FillValue" in raw[k][1]
    # every track is NaN-padded after its last sample, like ours
    valid = ~np.isnan(ref["lon_trks"])
    assert (np.diff(valid.astype(int), axis=1) <= 0).all() and valid[:, 0].all()


@needs_ref
def test_return_period_oracle_on_real_reference_tracks():
    """Notebook cells 13-17 on the notebook's own data (five sample files): oracle vs the NumPy formulas."""
    import warnings
    from oracle import tcr_oracle as orc
    from tropical_cyclone_risk_b200 import trackfile
    from test_analysis import MIAMI, notebook_haversine
    n_hit = 0
    for tag in ("", "_e0", "_e1"):
        t = trackfile.read_tracks(os.path.join(REF, "notebooks", "data", "tracks_NA_era5_197901_202312%s.nc" % tag))
        lon, lat, v = t["lon_trks"], t["lat_trks"], t["vmax_trks"]
        d = notebook_haversine(MIAMI[0], MIAMI[1], lon, lat)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = np.nanmax(np.where(d <= 100, v, np.nan), axis=1)
        got = orc.poi_vmax(lon, lat, v, MIAMI[0], MIAMI[1])
        assert np.array_equal(got, want, equal_nan=True)
        n_hit += int((~np.isnan(got)).sum())
    assert n_hit > 0


def test_cf_time_and_interpolation_weights():
    from tropical_cyclone_risk_b200 import refdata
    t = refdata.decode_cf_time([0, 31, 59.5], "days since 2001-01-15 00:00:00")
    assert t == [datetime.datetime(2001, 1, 15), datetime.datetime(2001, 2, 15), datetime.datetime(2001, 3, 15, 12)]
    assert refdata.decode_cf_time([6], "hours since 1900-01-01")[0] == datetime.datetime(1900, 1, 1, 6)
    with pytest.raises(NotImplementedError):
        refdata.decode_cf_time([0], "days since 2001-01-01", "noleap")
    times = [datetime.datetime(2001, m, 15) for m in (1, 2, 3)]
    assert refdata.time_weights(times, datetime.datetime(2001, 2, 15)) == (1, 1, 0.0)
    i0, i1, w = refdata.time_weights(times, datetime.datetime(2001, 1, 30, 12))
    assert (i0, i1) == (0, 1) and abs(w - 0.5) < 1e-12
    with pytest.raises(ValueError):
        refdata.time_weights(times, datetime.datetime(2001, 4, 1))


def test_reference_inputs_from_cache_files_written_here(tmp_path):
    """env_wnd / thermo caches in the reference's schema (NetCDF-3 here: SciPy is the only writer in the image),
    read back through ReferenceInputs: the planes equal those prepared directly from the same fields."""
    from scipy.io import netcdf_file
    from tropical_cyclone_risk_b200 import layout, params, refdata, synth
    from tropical_cyclone_risk_b200 import namelist as nl
    lon, lat = synth.era5_axes()
    year = 2003
    stamps = [datetime.datetime(year, m, 15) for m in range(1, 13)]
    days = np.array([(t - datetime.datetime(1979, 1, 1)).total_seconds() / 86400.0 for t in stamps])
    raws = [synth.synth_month_raw(year, m, lon, lat) for m in range(1, 13)]

    def write(path, names):
        with netcdf_file(path, "w", version=2) as f:
            f.createDimension("time", 12); f.createDimension("lat", lat.size); f.createDimension("lon", lon.size)
            v = f.createVariable("time", "f8", ("time",)); v.units = "days since 1979-01-01 00:00:00"; v.calendar = "proleptic_gregorian"; v[:] = days
            f.createVariable("lat", "f8", ("lat",))[:] = lat[::-1]                       # ERA5 files run north to south
            f.createVariable("lon", "f8", ("lon",))[:] = lon
            for n in names:
                f.createVariable(n, "f8", ("time", "lat", "lon"))[:] = np.stack([r[n][::-1] for r in raws])

    write(str(tmp_path / "env_wnd.nc"), layout.FIELD_NAMES[:14])
    write(str(tmp_path / "thermo.nc"), ("vmax", "chi", "rh_mid"))
    ri = refdata.ReferenceInputs("/nonexistent", str(tmp_path / "env_wnd.nc"), str(tmp_path / "thermo.nc"))
    olon, olat = synth.ocean_axes()
    oc = [synth.synth_ocean(olon, olat, m) for m in range(1, 13)]
    ri._ocean = (olon, olat, np.stack([o[0] for o in oc]), np.stack([o[1] for o in oc]))
    bounds = params.basin_bounds(nl, "NA")
    lon_b, lat_b, planes = ri.year_planes(nl, bounds, year)
    lon_w, lat_w, want = synth.prepared_year(nl, bounds, year)
    assert np.array_equal(lon_b, lon_w) and np.array_equal(lat_b, lat_w)
    assert planes.shape == want.shape == (12, layout.N_FIELDS, lat_b.size, lon_b.size)
    assert np.array_equal(planes, want)
    with pytest.raises(ValueError):
        ri.year_planes(nl, bounds, year + 1)                                              # outside the record


def test_h5lite_refuses_what_it_does_not_understand(tmp_path):
    from tropical_cyclone_risk_b200 import h5lite, refdata
    p = tmp_path / "x.nc"
    p.write_bytes(b"CDF\x01" + b"\0" * 64)
    with pytest.raises(h5lite.H5Error):
        h5lite.File(str(p))
    p.write_bytes(b"not a netcdf file at all")
    with pytest.raises(ValueError):
        refdata.open_variables(str(p))


@needs_ref
def test_oracle_year_on_the_reference_static_data():
    """The reference's real bathymetry, land mask, basin masks (from land.nc) and Levitus climatologies, under a
    synthetic atmosphere: a North Atlantic year forms its storms over open water in the tropical Atlantic."""
    from oracle import tcr_oracle as orc
    from tropical_cyclone_risk_b200 import fields, params, refdata, synth
    from tropical_cyclone_risk_b200 import namelist as nl
    ri = refdata.ReferenceInputs(REF)
    st = ri.static()
    bounds = params.basin_bounds(nl, "NA")
    stat = fields.prepare_static(bounds, st)
    mlon, mlat, m = fields.crop_masks(st, "NA", bounds)
    olon, olat, mld, strat = ri.ocean()
    lon, lat = synth.era5_axes()
    planes = []
    for month in range(1, 13):
        raw = synth.synth_month_raw(2001, month, lon, lat)
        lon_b, lat_b, pl = fields.prepare_month(nl, bounds, lon, lat, raw, olon, olat, mld[month - 1], strat[month - 1])
        planes.append(pl)
    env = orc.OracleEnv(lon_b, lat_b, np.stack(planes), stat)
    masks = orc.Masks(mlon, mlat, np.ascontiguousarray(m, dtype=np.uint8))
    r = orc.run_year(params.params_from_namelist(nl, "NA"), env, 0, masks, 20260101, 2001, 12, chunk=4096, n_threads=4)
    lon0, lat0 = r["lon"][:, 0], r["lat"][:, 0]
    assert not np.isnan(lon0).any() and np.nanmax(r["vmax"]) >= 18.0
    assert (lon0 > 262).all() and (lon0 < 358).all() and (lat0 > 3).all() and (lat0 < 45).all()
    # genesis points are over water according to the reference's own land mask
    ix = np.rint((lon0 - st["lon_l"][0]) / (st["lon_l"][1] - st["lon_l"][0])).astype(int)
    iy = np.rint((lat0 - st["lat_l"][0]) / (st["lat_l"][1] - st["lat_l"][0])).astype(int)
    assert (st["land"][iy, ix] == 0).all()


@needs_ref
def test_h5lite_attribute_span_matches_real_attribute_messages():
    """Dense attribute storage (objects with more than eight attributes) is walked with the same fractal-heap code as
    the dense links of the sample track files plus a per-message length computed from the message's own header; that
    length is checked here against every compact attribute message of two real files (both header generations)."""
    from tropical_cyclone_risk_b200 import h5lite
    n = 0
    for path in (os.path.join(REF, "intensity", "data", "mld_climatology.nc"),
                 os.path.join(REF, "notebooks", "data", "tracks_NA_era5_197901_202312.nc")):
        f = h5lite.File(path)
        for name in f.keys():
            for mtype, body in f._object_header(f._links[name]):
                if mtype != 0x0C:
                    continue
                probe = h5lite.File.__new__(h5lite.File)
                probe.buf, probe.osz, probe.lsz = bytes(body) + b"\0" * 16, 8, 8
                used = probe._attribute_span(0)[0]                      # DIMENSION_LIST (vlen of references) included
                assert 0 <= len(body) - used < 8, (path, name, len(body), used)
                n += 1
    assert n > 60
