"""Host-side preparation of one year of environment tables (the step right before the hot path).

Restates, in NumPy, what the reference does with xarray at the top of ``run_tracks``
(util/compute.py:66-121) and inside ``BetaAdvectionTrack._load_wnd_stat`` /
``Coupled_FAST.init_fields`` (track/bam_track.py:72-91, intensity/coupled_fast.py:217-225):
NaN policies, PI scaling, the chi transform, regridding of the ocean climatologies to the
thermo grid, and the basin crop of util/basins.py:57-75.  The result is the 19 float32 planes
per month that ``tcr_upload_month`` interleaves into the HBM record layout.

Precision policy: the planes are stored as float32 (source data are float32/int16/int8);
all interpolation weights and ODE arithmetic downstream are float64.
"""
import numpy as np

from . import layout


def crop_to_basin(lon, lat, field, bounds):
    """TC_Basin.transform_global_field (util/basins.py:57-75) for ascending axes.

    bounds = (lon_min, lat_min, lon_max, lat_max).  Returns (lon_b, lat_b, field_b)."""
    lon = np.asarray(lon, dtype=np.float64)
    lat = np.asarray(lat, dtype=np.float64)
    lon_min, lat_min, lon_max, lat_max = bounds
    if lon[0] >= -1e-5 and (lon_min < 0 or lon_max < 0):
        wrap = lon >= (180 - 1e-5)                                  # basins.py:92-98
        field = np.concatenate((field[..., wrap], field[..., ~wrap]), axis=-1)
        lon = np.hstack((lon[wrap] - 360, lon[~wrap]))
    elif (lon < 0).any() and lon_min >= 0:
        neg = lon < -1e-5                                           # basins.py:103-107
        field = np.concatenate((field[..., ~neg], field[..., neg]), axis=-1)
        lon = np.hstack((lon[~neg], lon[neg] + 360))
    lon_mask = (lon <= lon_max + 1e-5) & (lon >= lon_min - 1e-5)
    lat_mask = (lat >= lat_min - 1e-5) & (lat <= lat_max + 1e-5)
    return lon[lon_mask], lat[lat_mask], field[..., lat_mask, :][..., lon_mask]


def crop_index_maps(lon, lat, bounds):
    """crop_to_basin as index maps: (lon_b, lat_b, src_col, src_row) with field_b == field[src_row][:, src_col]
    (what tcr_prepare_month consumes).  Axes in storage order; latitude may be descending."""
    lon = np.asarray(lon, dtype=np.float64)
    lat = np.asarray(lat, dtype=np.float64)
    rows = np.arange(lat.size)
    if lat[0] > lat[1]:                                              # util/compute.py:80-84
        lat, rows = lat[::-1], rows[::-1]
    cols = np.arange(lon.size, dtype=np.float64)
    lon_b, lat_b, cols_b = crop_to_basin(lon, lat, np.broadcast_to(cols, (lat.size, lon.size)), bounds)
    _, _, rows_b = crop_to_basin(lon, lat, np.broadcast_to(rows[:, None].astype(np.float64), (lat.size, lon.size)), bounds)
    return lon_b, lat_b, cols_b[0].astype(np.int32), rows_b[:, 0].astype(np.int32)


def _locate(axis, x):
    """Clamped interval index and the two linear B-spline weights (FITPACK fpbspl, k=1)."""
    x = np.clip(x, axis[0], axis[-1])
    i = np.clip(np.searchsorted(axis, x, side="right") - 1, 0, axis.size - 2)
    f = 1.0 / (axis[i + 1] - axis[i])
    return i, f * (axis[i + 1] - x), f * (x - axis[i])


def bilinear(lon, lat, field, qlon, qlat):
    """RectBivariateSpline(lon, lat, field.T, kx=1, ky=1).ev(qlon, qlat): clamped bilinear
    (util/mat.py:142-153).  field is [lat, lon] with ascending axes."""
    lon = np.asarray(lon, dtype=np.float64)
    lat = np.asarray(lat, dtype=np.float64)
    f = np.asarray(field, dtype=np.float64)
    ix, wx0, wx1 = _locate(lon, np.asarray(qlon, dtype=np.float64))
    iy, wy0, wy1 = _locate(lat, np.asarray(qlat, dtype=np.float64))
    sp = f[iy, ix] * wx0 * wy0
    sp = sp + f[iy + 1, ix] * wx0 * wy1
    sp = sp + f[iy, ix + 1] * wx1 * wy0
    sp = sp + f[iy + 1, ix + 1] * wx1 * wy1
    return sp


def regrid(lon_src, lat_src, field, lon_dst, lat_dst):
    """mat.interp_2d_grid (util/mat.py:159-164)."""
    lat_src = np.asarray(lat_src, dtype=np.float64)
    if lat_src[1] - lat_src[0] < 0:                                  # util/mat.py:143-146
        lat_src = lat_src[::-1]
        field = field[::-1, :]
    LON, LAT = np.meshgrid(lon_dst, lat_dst)
    return bilinear(lon_src, lat_src, field, LON, LAT)


def prepare_month(namelist, bounds, lon, lat, raw, ocean_lon, ocean_lat, mld, strat):
    """One month of prepared, basin-cropped float32 planes.

    raw: dict with the 14 wind statistics (layout.FIELD_NAMES[:14]) and 'vmax', 'chi', 'rh_mid',
    each [nlat, nlon] on (lon, lat) (ascending).  mld/strat: [nlat_o, nlon_o] on the ocean grid.
    Returns (lon_b, lat_b, planes[19][nlat_b][nlon_b] float32)."""
    lon = np.asarray(lon, dtype=np.float64)
    lat = np.asarray(lat, dtype=np.float64)
    if lat[0] > lat[1]:                                              # util/compute.py:80-84
        lat = lat[::-1]
        raw = {k: v[::-1, :] for k, v in raw.items()}
    f64 = lambda a: np.asarray(a, dtype=np.float64)
    vpot = np.nan_to_num(f64(raw["vmax"]) * namelist.PI_reduc * np.sqrt(namelist.Ck / namelist.Cd))
    chi = f64(raw["chi"]).copy()
    chi[np.isnan(chi)] = 5                                           # util/compute.py:113
    chi = np.maximum(np.minimum(np.exp(np.log(chi + 1e-3) + namelist.log_chi_fac)
                                + namelist.chi_fac, 5), 1e-5)        # util/compute.py:115
    mld_g = regrid(ocean_lon, ocean_lat, np.nan_to_num(f64(mld)), lon, lat)       # :117
    strat_g = regrid(ocean_lon, ocean_lat, np.nan_to_num(f64(strat)), lon, lat)   # :118
    fields = [np.nan_to_num(f64(raw[name])) for name in layout.FIELD_NAMES[:14]]  # bam_track.py:74
    fields += [chi, vpot, mld_g, strat_g, f64(raw["rh_mid"])]
    stack = np.stack(fields)
    lon_b, lat_b, planes = crop_to_basin(lon, lat, stack, bounds)
    return lon_b, lat_b, np.ascontiguousarray(planes, dtype=np.float32)


def prepare_static(bounds, static):
    """Basin crop of bathymetry and land (intensity/geo.py:9-33) and the 8 mask planes
    (7 sorted basins + the run basin's own mask, util/compute.py:87-97)."""
    lon_b, lat_b, bathy = crop_to_basin(static["lon_b"], static["lat_b"], static["bathy"], bounds)
    lon_l, lat_l, land = crop_to_basin(static["lon_l"], static["lat_l"], static["land"], bounds)
    return dict(lon_b=lon_b, lat_b=lat_b, bathy=np.ascontiguousarray(bathy, dtype=np.int16),
                lon_l=lon_l, lat_l=lat_l, land=np.ascontiguousarray(land, dtype=np.int8))


def crop_masks(static, run_basin_id, bounds, p=None):
    """The 8 genesis mask planes on the part of their global grid a seed attempt can query.

    The reference samples the GLOBAL 0.25-degree masks (util/compute.py:87-97, 146, 156-157).  The first
    draw of an attempt is latitude-weighted over [lat_min, lat_max] of :140-141 -- which, through np.sign(-0.0)
    >= 0, reaches 45 N for the southern basins whose box ends at '0S' -- so the crop must cover that range as
    well as the basin box: a query outside a narrower crop would be clamped to its edge row and read ocean
    where the global mask reads 0.  Returns (lon_m, lat_m, planes uint8 [8][nlat][nlon])."""
    lat_lo, lat_hi = bounds[1], bounds[3]
    if p is not None:
        lat_lo, lat_hi = min(lat_lo, float(p.gen_lat_min)), max(lat_hi, float(p.gen_lat_max))
    else:
        lat_lo, lat_hi = min(lat_lo, -45.0), max(lat_hi, 45.0)
    wide = (bounds[0], lat_lo, bounds[2], lat_hi)
    mlon, mlat, m = crop_to_basin(static["lon_m"], static["lat_m"], mask_planes(static, run_basin_id), wide)
    return mlon, mlat, np.ascontiguousarray(m, dtype=np.uint8)


def mask_planes(static, run_basin_id):
    """uint8 [8][nlat_m][nlon_m]: layout.BASIN_IDS order then the run basin's mask."""
    if run_basin_id == "GL":
        own = static["mask_GL"]
    else:
        own = static["masks"][layout.BASIN_IDS.index(run_basin_id)]
    return np.ascontiguousarray(np.concatenate([static["masks"], own[None]], axis=0), dtype=np.uint8)
