"""Deterministic synthetic ERA5-shaped inputs (no network, no NetCDF reader in this image).

Shapes follow what the reference downloads and ships:
  * monthly wind statistics / thermodynamic fields: 1 deg, 181 lat x 360 lon
    (scripts/download_era5.py:58), 14 wind-stat + 3 thermo fields per month
    (track/env_wind.py:22-42, thermo/calc_thermo.py:103-116)
  * ocean climatologies: mixed-layer depth / stratification (intensity/ocean.py)
  * bathymetry int16 1350 x 2700, land int8 1440 x 2880 (intensity/data/*.nc)
  * basin masks 721 x 1440, 0.25 deg (scripts/generate_land_masks.py:24-110)

Every value is float32-representable so the fp32 HBM tables and the float64 CPU oracle
read identical numbers.  Roughness is fixed at 1 m/s grid-scale white noise on the mean
winds (the RK45 step count depends on it).
"""
import numpy as np

from . import layout

# (lon0, lon1, lat0, lat1) boxes of the synthetic continents, degrees east / north
_CONTINENTS = (
    (235.0, 282.0, 28.0, 72.0),     # "North America"
    (255.0, 275.0, 16.0, 30.0),     # "Mexico"
    (280.0, 325.0, -56.0, 9.0),     # "South America"
    (0.0, 48.0, -35.0, 36.0),       # "Africa" east of Greenwich
    (343.0, 360.0, 5.0, 33.0),      # "Africa" west of Greenwich
    (0.0, 140.0, 38.0, 76.0),       # "Eurasia"
    (68.0, 90.0, 8.0, 38.0),        # "India"
    (95.0, 122.0, 10.0, 38.0),      # "Indochina / China"
    (113.0, 153.0, -39.0, -12.0),   # "Australia"
)


def era5_axes(res=1.0):
    """Ascending lon (0 .. 360-res) and lat (-90 .. 90) axes of an ERA5-shaped grid."""
    nlon = int(round(360.0 / res))
    nlat = int(round(180.0 / res)) + 1
    lon = (np.arange(nlon) * res).astype(np.float64)
    lat = (-90.0 + np.arange(nlat) * res).astype(np.float64)
    return lon, lat


def land_fraction(lon, lat):
    """Boolean land mask [lat, lon] of the synthetic continents on any grid."""
    LON, LAT = np.meshgrid(np.asarray(lon, dtype=np.float64), np.asarray(lat, dtype=np.float64))
    land = np.zeros(LON.shape, dtype=bool)
    for (x0, x1, y0, y1) in _CONTINENTS:
        land |= (LON >= x0) & (LON <= x1) & (LAT >= y0) & (LAT <= y1)
    land |= LAT < -70.0
    return land


def _season(month):
    """+1 in boreal late summer, -1 in austral late summer."""
    return float(np.cos(2.0 * np.pi * (month - 8.5) / 12.0))


def synth_month_raw(year, month, lon, lat, roughness=1.0, zero_cov_over_land=False):
    """Raw (un-prepared) monthly fields on the global grid, dict name -> float32 [nlat, nlon].

    Keys: the 14 wind statistics of layout.FIELD_NAMES, plus 'vmax' (potential intensity,
    m/s, before PI_reduc), 'chi' (raw saturation deficit), 'rh_mid'.  The generator seed is
    1000*year + month.
    """
    rng = np.random.default_rng(1000 * int(year) + int(month))
    lam = np.deg2rad(lon)[None, :]
    phi = np.deg2rad(lat)[:, None]
    s = _season(month)
    shape = (lat.size, lon.size)

    def noise():
        return roughness * rng.standard_normal(shape)

    u850 = -5.0 * np.cos(2.5 * phi) + 0.0 * lam
    shear_u = 25.0 * np.sin(phi - np.deg2rad(6.0 * s)) ** 2 + 2.0 * np.sin(lam)
    out = {}
    out["ua250_Mean"] = u850 + shear_u + noise()
    out["va250_Mean"] = 2.0 * np.sin(2.0 * lam) * np.cos(phi) + noise()
    out["ua850_Mean"] = u850 + noise()
    out["va850_Mean"] = 1.5 * np.cos(3.0 * lam) * np.cos(phi) + noise()

    # covariance = A A^T with a smooth lower-triangular A (diag ~3 m/s, off-diag ~0.8)
    A = np.zeros((4, 4) + shape)
    for i in range(4):
        for j in range(i + 1):
            if i == j:
                A[i, j] = 3.0 + 0.6 * np.cos((i + 1) * lam + 0.3 * j) * np.cos(phi) \
                    + 0.8 * np.sin(phi) ** 2
            else:
                A[i, j] = 0.8 * np.sin((i + j + 1) * lam + 0.7 * i) * np.cos(2.0 * phi + 0.4 * j)
    zero = land_fraction(lon, lat) if zero_cov_over_land else None
    for i in range(4):
        for j in range(i + 1):
            cij = np.zeros(shape)
            for k in range(j + 1):
                cij = cij + A[i, k] * A[j, k]
            if zero is not None:
                cij = np.where(zero, 0.0, cij)
            out[layout.FIELD_NAMES[layout.CH_COV + layout.cov_index(i, j)]] = cij

    out["vmax"] = 70.0 * np.exp(-(((lat[:, None] - 15.0 * s) / 18.0) ** 2)) \
        * (1.0 + 0.05 * np.cos(2.0 * lam)) + 0.0 * lam
    out["chi"] = 0.8 + 0.2 * np.sin(2.0 * lam + 0.5 * s) * np.cos(3.0 * phi)
    out["rh_mid"] = 0.55 + 0.35 * np.cos(2.0 * phi) * np.sin(lam + 1.0 + 0.3 * s)
    return {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in out.items()}


def synth_ocean(lon, lat, month):
    """Mixed-layer depth (m) and sub-mixed-layer stratification (K/100 m) on the given grid.

    A small patch of negative stratification is kept on purpose: the real climatology has
    negative values and they drive dv/dt to the NaN -> 0 branch
    (intensity/coupled_fast.py:93,150)."""
    lam = np.deg2rad(lon)[None, :]
    phi = np.deg2rad(lat)[:, None]
    s = _season(month)
    mld = 40.0 + 20.0 * np.cos(lam) * np.cos(phi) + 5.0 * s * np.sin(phi)
    strat = 5.0 + 2.0 * np.sin(4.0 * phi) + 0.0 * lam
    LON, LAT = np.meshgrid(lon, lat)
    strat = np.where((LON > 300.0) & (LON < 306.0) & (LAT > 18.0) & (LAT < 22.0), -1.5, strat)
    return mld.astype(np.float32), strat.astype(np.float32)


def _dilate(mask, n):
    out = mask.copy()
    for _ in range(n):
        m = out.copy()
        m[1:, :] |= out[:-1, :]
        m[:-1, :] |= out[1:, :]
        m |= np.roll(out, 1, axis=1) | np.roll(out, -1, axis=1)
        out = m
    return out


def synth_static(full_res=True):
    """Bathymetry (int16, m), land (int8 0/1) with their own axes, and the 8 basin masks.

    full_res=True gives the shapes of the reference's bundled files (1350 x 2700 and
    1440 x 2880); False gives 4x coarser grids for quick CPU tests.
    Returns dict with lon_b, lat_b, bathy, lon_l, lat_l, land, lon_m, lat_m, masks
    (masks: uint8 [7][721][1440] in layout.BASIN_IDS order, plus 'GL')."""
    if full_res:
        nlat_b, nlon_b, nlat_l, nlon_l = 1350, 2700, 1440, 2880
    else:
        nlat_b, nlon_b, nlat_l, nlon_l = 338, 675, 360, 720
    db = 360.0 / nlon_b
    lon_b = ((np.arange(nlon_b) + 0.5) * db).astype(np.float32).astype(np.float64)
    lat_b = (-90.0 + (np.arange(nlat_b) + 0.5) * (180.0 / nlat_b)).astype(np.float32).astype(np.float64)
    dl = 360.0 / nlon_l
    lon_l = (np.arange(nlon_l) * dl).astype(np.float64)
    lat_l = (-90.0 + dl + np.arange(nlat_l) * dl).astype(np.float64)   # -89.875 .. 90 at 0.125 deg

    land_b = land_fraction(lon_b, lat_b)
    shelf = _dilate(land_b, max(1, int(round(0.7 / db)))) & ~land_b
    bathy = np.full(land_b.shape, -4000, dtype=np.int16)
    bathy[shelf] = -30
    bathy[land_b] = 300
    land = land_fraction(lon_l, lat_l).astype(np.int8)

    lon_m, lat_m, masks, gl = basin_masks(lambda lon, lat: ~land_fraction(lon, lat))
    return dict(lon_b=lon_b, lat_b=lat_b, bathy=bathy, lon_l=lon_l, lat_l=lat_l, land=land,
                lon_m=lon_m, lat_m=lat_m, masks=masks, mask_GL=gl)


def basin_masks(ocean_fn):
    """The basin masks of scripts/generate_land_masks.py:43-110 on its 0.25-degree grid (721 x 1440, 0..360 E):
    boxes AND ocean.  ocean_fn(lon_m, lat_m) -> bool [721][1440].  Returns (lon_m, lat_m, masks uint8 [7][721][1440]
    in layout.BASIN_IDS order, mask_GL uint8)."""
    lat_m = np.linspace(-90.0, 90.0, 721)
    lon_m = np.linspace(0.0, 360.0, 1441)[:-1]
    LON, LAT = np.meshgrid(lon_m, lat_m)
    ocean = np.asarray(ocean_fn(lon_m, lat_m), dtype=bool)
    na_box = np.zeros(LON.shape, dtype=bool)
    for la, lo in zip((0, 9, 10, 14, 18), (285, 278, 276, 271, 262)):
        na_box |= (LAT >= la) & (LON >= lo) & ocean
    ep_box = np.zeros(LON.shape, dtype=bool)
    for la, lo in zip((7.5, 8.8, 9, 10, 15, 18, 60), (295, 282, 277, 276.5, 276, 271, 262)):
        ep_box |= (LAT <= la) & (LON <= lo) & ocean
    box = lambda x0, x1, y0, y1: (LON >= x0) & (LON <= x1) & (LAT >= y0) & (LAT <= y1)
    m = {
        "NA": box(255, 360, 0, 60) & na_box,
        "EP": box(180, 290, 0, 60) & ep_box,
        "WP": box(100, 180, 0, 60) & ocean,
        "NI": box(30, 100, 0, 49) & ocean,
        "SI": box(10, 100, -45, 0) & ocean,
        "AU": box(100, 170, -45, 0) & ocean,
        "SP": box(170, 260, -45, 0) & ocean,
    }
    gl = ocean & (np.abs(LAT) <= 50)
    masks = np.stack([m[b] for b in layout.BASIN_IDS]).astype(np.uint8)
    return lon_m, lat_m, masks, gl.astype(np.uint8)


def ocean_axes():
    """Axes of the bundled Levitus-style climatologies after the wrap column is dropped
    (intensity/ocean.py:26,55): 180 lat x 360 lon, 1 deg, cell-centred latitudes."""
    lon = np.arange(360, dtype=np.float64)
    lat = -89.5 + np.arange(180, dtype=np.float64)
    return lon, lat


def prepared_year(namelist, bounds, year, lon=None, lat=None, roughness=1.0, zero_cov_over_land=False):
    """Twelve months of prepared float32 planes for one year on the basin crop.

    Returns (lon_b, lat_b, planes[12][19][nlat_b][nlon_b])."""
    from . import fields
    if lon is None:
        lon, lat = era5_axes(1.0)
    olon, olat = ocean_axes()
    months = []
    for month in range(1, 13):
        raw = synth_month_raw(year, month, lon, lat, roughness, zero_cov_over_land)
        mld, strat = synth_ocean(olon, olat, month)
        lon_b, lat_b, planes = fields.prepare_month(namelist, bounds, lon, lat, raw, olon, olat, mld, strat)
        months.append(planes)
    return lon_b, lat_b, np.stack(months)
