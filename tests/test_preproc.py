"""Pre-processing kernels.

Wind mean / covariance reduction (track/env_wind.py:169-228): the reference runs on xarray, which
this container does not have, so the oracle (oracle/preproc_oracle.py) restates xarray's published
reduction semantics and is pinned here against numpy.mean / numpy.var / numpy.cov; the CUDA kernel
(tcr_wind_stats) must match the oracle bit for bit."""
import datetime
import os

import numpy as np
import pytest

from oracle import preproc_oracle as po


def synth_winds(n_time, n_lvl, nlat, nlon, seed=0, nan_frac=0.0):
    """ERA5-like (time, level, lat, lon) float32 winds: smooth jets + day-to-day synoptic noise."""
    rng = np.random.default_rng(seed)
    lat = np.linspace(-90, 90, nlat)[None, None, :, None]
    lon = np.linspace(0, 360, nlon, endpoint=False)[None, None, None, :]
    lvl = np.arange(n_lvl)[None, :, None, None]
    t = np.arange(n_time)[:, None, None, None]
    ua = 12 * np.cos(np.deg2rad(3 * lat)) * (1 + 0.3 * lvl) + 3 * np.sin(np.deg2rad(lon) + 0.2 * t)
    va = 2 * np.sin(np.deg2rad(2 * lon) + 0.3 * t) * np.cos(np.deg2rad(lat)) + 0.5 * lvl
    ua = (ua + rng.normal(0, 4, (n_time, n_lvl, nlat, nlon))).astype(np.float32)
    va = (va + 0.5 * ua + rng.normal(0, 3, (n_time, n_lvl, nlat, nlon))).astype(np.float32)
    if nan_frac:
        ua[rng.random(ua.shape) < nan_frac] = np.nan
        va[rng.random(va.shape) < nan_frac] = np.nan
    return ua, va


def series_of(ua, va, iu, il):
    n_time = ua.shape[0]
    return [a[:, k].reshape(n_time, -1) for k in (iu, il) for a in (ua, va)]


def test_oracle_matches_numpy_on_nan_free_input():
    ua, va = synth_winds(62, 3, 13, 24, seed=1)
    s = series_of(ua, va, 0, 2)
    got = po.wind_stats(s, np.arange(63))
    x = np.stack([a.astype(np.float64) for a in s])                  # [4, n_time, n_pts]
    assert np.array_equal(got[:4], x.mean(axis=1))                    # .mean(dim)          env_wind.py:207
    k = 4
    for i in range(4):
        for j in range(i + 1):
            if i == j:
                assert np.array_equal(got[k], x[i].var(axis=0))       # .var(dim), ddof 0   env_wind.py:211
            else:
                want = np.array([np.cov(x[i][:, p], x[j][:, p], ddof=1)[0, 1] for p in range(x.shape[2])])
                np.testing.assert_allclose(got[k], want, rtol=1e-11, atol=1e-12)   # xr.cov, ddof 1   :213
            k += 1
    assert k == po.N_STATS


def test_oracle_daily_grouping_and_nan_policy():
    ua, va = synth_winds(20, 2, 3, 5, seed=2, nan_frac=0.15)
    s = series_of(ua, va, 0, 1)
    gs = np.arange(0, 21, 4)                                           # 5 days x 4 samples
    got = po.wind_stats(s, gs)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dm = [np.nanmean(a.astype(np.float64).reshape(5, 4, -1), axis=1) for a in s]
        np.testing.assert_allclose(got[0], np.nanmean(dm[0], axis=0), rtol=1e-14)
        np.testing.assert_allclose(got[4], np.nanvar(dm[0], axis=0), rtol=1e-13)
        # xr.cov: only the days on which BOTH variables are valid
        a, b = dm[1].copy(), dm[0].copy()
        both = ~np.isnan(a) & ~np.isnan(b)
        a[~both] = np.nan
        b[~both] = np.nan
        want = np.nansum((a - np.nanmean(a, 0)) * (b - np.nanmean(b, 0)), axis=0) / (both.sum(0) - 1)
    ok = both.sum(0) > 1
    np.testing.assert_allclose(got[5][ok], want[ok], rtol=1e-12)


def test_month_samples_literal_and_intent():
    from tropical_cyclone_risk_b200 import preproc
    t0 = datetime.datetime(2001, 8, 30)
    times = [t0 + datetime.timedelta(hours=12 * k) for k in range(80)]   # 2 x daily, Aug 30 .. Oct 8
    idx, gs = preproc.month_samples(times, datetime.datetime(2001, 9, 15))
    assert idx[0] == 4 and idx.size == 60                                 # September: 30 days x 2
    assert np.array_equal(gs, np.arange(61))                              # literal reference: no daily averaging
    idx, gs = preproc.month_samples(times, datetime.datetime(2001, 9, 15), group_sub_daily=True)
    assert np.array_equal(gs, np.arange(0, 61, 2))
    times5 = [datetime.datetime(2001, 1, 1) + datetime.timedelta(days=5 * k) for k in range(30)]
    idx, gs = preproc.month_samples(np.array(times5, dtype="datetime64[s]"), datetime.datetime(2001, 3, 15))
    assert idx.size == 6 and np.array_equal(gs, np.arange(7))            # one sample per day group
    with pytest.raises(ValueError):
        preproc.month_samples(times, datetime.datetime(2003, 1, 15))


def test_names_match_table_channels():
    from tropical_cyclone_risk_b200 import layout, preproc
    assert preproc.wind_mean_vector_names() + preproc.wind_cov_matrix_names() == list(layout.FIELD_NAMES[:14])
    assert preproc.level_index([1000, 850, 250], "hPa", 250) == 2
    assert preproc.level_index([100000, 85000, 25000], "Pa", 850) == 1


# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def engine():
    from tropical_cyclone_risk_b200 import namelist as nl
    from tropical_cyclone_risk_b200.engine import Engine
    from tropical_cyclone_risk_b200.params import params_from_namelist
    eng = Engine(params_from_namelist(nl, "NA"), device=0)
    yield eng
    eng.close()


WS_VARIANTS_SINGLE = [None, 10, 11, 12, 13, 14]
WS_VARIANTS_GROUPED = [None, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9]


def _with_variant(v, fn):
    old = os.environ.pop("TCR_WS_VARIANT", None)
    try:
        if v is not None:
            os.environ["TCR_WS_VARIANT"] = str(v)
        return fn()
    finally:
        os.environ.pop("TCR_WS_VARIANT", None)
        if old is not None:
            os.environ["TCR_WS_VARIANT"] = old


@pytest.mark.gpu
@pytest.mark.parametrize("shape,nan_frac", [((181, 360), 0.0), ((7, 143), 0.1), ((1, 3), 0.0), ((33, 128), 0.3)])
def test_gpu_wind_stats_ungrouped_bit_exact(engine, shape, nan_frac):
    """The reference's own case: 2 x daily samples, no daily averaging (62 single-sample groups);
    ragged / unaligned rows (n_pts = 1001, 3) take the register path."""
    ua, va = synth_winds(62, 3, shape[0], shape[1], seed=3, nan_frac=nan_frac)
    gs = np.arange(63, dtype=np.int32)
    want = po.wind_stats(series_of(ua, va, 2, 0), gs)
    for v in WS_VARIANTS_SINGLE:
        got = _with_variant(v, lambda: engine.wind_stats(ua, va, 2, 0, gs))
        assert np.array_equal(got, want, equal_nan=True), "variant %s" % v
    assert not np.isnan(want[:, 0]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("shape,nan_frac", [((91, 180), 0.0), ((5, 77), 0.2)])
def test_gpu_wind_stats_daily_groups_bit_exact(engine, shape, nan_frac):
    """Daily means first (4 x daily, ragged first and last day), every kernel variant."""
    ua, va = synth_winds(118, 2, shape[0], shape[1], seed=4, nan_frac=nan_frac)
    gs = np.r_[0, np.arange(2, 118, 4), 118].astype(np.int32)             # 2 + 29 x 4 (the last one 4) samples
    assert np.all(np.diff(gs) > 0)
    want = po.wind_stats(series_of(ua, va, 0, 1), gs)
    for v in WS_VARIANTS_GROUPED:
        got = _with_variant(v, lambda: engine.wind_stats(ua, va, 0, 1, gs))
        assert np.array_equal(got, want, equal_nan=True), "variant %s" % v


@pytest.mark.gpu
def test_gpu_calc_wnd_stat_mirror(engine):
    """calc_wnd_stat through the reference-shaped host function: month selection out of a longer record."""
    from tropical_cyclone_risk_b200 import preproc
    t0 = datetime.datetime(2001, 8, 30)
    times = [t0 + datetime.timedelta(hours=12 * k) for k in range(80)]
    ua, va = synth_winds(80, 4, 19, 36, seed=5)
    levels = [1000, 850, 500, 250]
    got = preproc.calc_wnd_stat(engine, ua, va, times, levels, datetime.datetime(2001, 9, 15))
    want = po.wind_stats(series_of(ua[4:64], va[4:64], 3, 1), np.arange(61))
    assert got.shape == (14, 19, 36)
    assert np.array_equal(got.reshape(14, -1), want)
    got = preproc.calc_wnd_stat(engine, ua, va, times, levels, datetime.datetime(2001, 9, 15), group_sub_daily=True)
    want = po.wind_stats(series_of(ua[4:64], va[4:64], 3, 1), np.arange(0, 61, 2))
    assert np.array_equal(got.reshape(14, -1), want)


@pytest.mark.gpu
def test_gpu_wind_stats_rejects_bad_groups(engine):
    from tropical_cyclone_risk_b200._lib import TcrError
    ua, va = synth_winds(8, 2, 4, 8)
    with pytest.raises(TcrError):
        engine.wind_stats(ua, va, 0, 1, np.array([0, 4, 4, 8], np.int32))      # empty day
    with pytest.raises(TcrError):
        engine.wind_stats(ua, va, 0, 1, np.array([0, 4, 7], np.int32))         # does not end at n_time


# ================================================================================================
# Thermodynamics (thermo/thermo.py, thermo/calc_thermo.py): the C oracle is pinned against golden
# vectors produced by the UNMODIFIED reference (oracle/make_golden_thermo.py); the CUDA kernel must
# match the oracle bit for bit and the reference to 1e-9.
# ================================================================================================
def _thermo_golden():
    from conftest import golden
    g = golden("ref_thermo.npz")
    return g, (g["table_p"], g["table_s"], g["table_T"])


def _assert_close_to_reference(got, want, rtol, atol, what):
    assert np.array_equal(np.isnan(got), np.isnan(want)), what
    ok = ~np.isnan(want)
    np.testing.assert_allclose(got[ok], want[ok], rtol=rtol, atol=atol, err_msg=what)


def test_thermo_oracle_matches_reference_golden():
    """2048 synthetic ERA5-shaped soundings incl. land (sst = 0 K), dry, super-saturated, isothermal and
    NaN-level columns: identical NaN pattern, PI within 1e-11 relative of the live reference."""
    g, table = _thermo_golden()
    assert int(np.frombuffer(table[2].tobytes(), dtype=np.uint32).sum() & 0xffffffff) == int(g["table_crc"])
    from tropical_cyclone_risk_b200 import synth_thermo
    assert all(np.array_equal(a, b) for a, b in zip(synth_thermo.fixture_table(), table))        # what bench / smoke load
    v, c, r = po.thermo(g["p"], g["ta"], g["hus"], g["sst"], g["psl"], table, float(g["cecd"]), int(g["k_mid"]))
    _assert_close_to_reference(v, g["vmax"], 1e-11, 1e-11, "vmax")
    _assert_close_to_reference(c, g["chi"], 1e-10, 1e-12, "chi")
    _assert_close_to_reference(r, g["rh_mid"], 1e-13, 0, "rh_mid")
    assert (g["vmax"] > 30).sum() > 500 and (g["vmax"] == 0).sum() > 50           # both regimes are covered


def test_thermo_oracle_against_live_reference_if_present():
    """Fresh soundings through the reference itself when /root/reference exists (build container only)."""
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("reference tree not present")
    import warnings
    from tropical_cyclone_risk_b200 import synth_thermo
    ref = rh.load_reference()
    _, table = _thermo_golden()
    p, ta, hus, sst, psl = synth_thermo.soundings(8192, seed=99)
    shp = (64, 128)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = ref.thermo.CAPE_PI_vectorized(sst.reshape(shp), psl.reshape(shp), p.copy(),
                                             ta.astype(np.float64).reshape((-1,) + shp), hus.astype(np.float64).reshape((-1,) + shp))
    v, _, _ = po.thermo(p, ta, hus, sst, psl, table, ref.namelist.Ck / ref.namelist.Cd, 13)
    _assert_close_to_reference(v, want.reshape(-1), 1e-11, 1e-11, "vmax")


# ---- namelist.select_thermo = 2 (reversible thermodynamics, three-dimensional inversion table) ----
def _thermo_rev_golden():
    from conftest import golden
    g = golden("ref_thermo_rev.npz")
    return g, (g["table_p"], g["table_s"], g["table_rt"], g["table_T"])


def test_thermo_reversible_oracle_matches_reference_golden():
    """The same kind of soundings through the unmodified reference with namelist.select_thermo = 2
    (oracle/make_golden_thermo.py --reversible): identical NaN / zero pattern, PI within 1e-10 relative."""
    g, table = _thermo_rev_golden()
    v, c, r = po.thermo(g["p"], g["ta"], g["hus"], g["sst"], g["psl"], table, float(g["cecd"]), int(g["k_mid"]))
    _assert_close_to_reference(v, g["vmax"], 1e-10, 1e-10, "vmax")
    assert np.array_equal(v == 0, g["vmax"] == 0)
    _assert_close_to_reference(c, g["chi"], 1e-10, 1e-12, "chi")
    _assert_close_to_reference(r, g["rh_mid"], 1e-13, 0, "rh_mid")
    assert (g["vmax"] > 30).sum() > 500 and (g["vmax"] == 0).sum() > 50
    # the two thermodynamics differ: this is not the pseudoadiabatic answer under another name
    g1, t1 = _thermo_golden()
    v1, _, _ = po.thermo(g["p"], g["ta"], g["hus"], g["sst"], g["psl"], t1, float(g["cecd"]), int(g["k_mid"]))
    assert np.nanmax(np.abs(v1 - v)) > 1.0


def test_entropy_lookup3_is_scipy_interpn_bit_for_bit():
    """thermo.py:343-353 reads the reversible table through scipy.interpolate.interpn(method='linear', bounds_error=False,
    fill_value=nan): the restatement (interval search, norm distances, corner order, weight products) against scipy itself on
    random points, exact nodes, both ends of every axis, points outside and NaNs."""
    from scipy.interpolate import interpn
    _, table = _thermo_rev_golden()
    pl, sl, rl, T = table
    rng = np.random.default_rng(3)
    n = 20000
    p = rng.uniform(pl[0] - 2000.0, pl[-1] + 2000.0, n)
    s = rng.uniform(sl[0] - 20.0, sl[-1] + 20.0, n)
    r = rng.uniform(-0.001, rl[-1] + 0.002, n)
    p[:40] = np.repeat(pl[[0, -1, 5, 17]], 10); s[:40:3] = sl[[0, -1, 9, 2, 30, 33, 1, 0, -1, 12, 4, 8, 20, 21]]
    r[1:40:4] = rl[[0, -1, 3, 24, 0, 11, 12, 1, -1, 7]]
    p[40], s[41], r[42] = np.nan, np.nan, np.nan
    p[43], s[44], r[45] = np.inf, -np.inf, np.inf
    want = interpn((pl, sl, rl), T, (p, s, r), method="linear", bounds_error=False, fill_value=np.nan)
    got = po.entropy_lookup3(p, s, r, table)
    assert np.array_equal(got, want, equal_nan=True)
    assert np.isnan(want).sum() > 1000 and (~np.isnan(want)).sum() > 10000


def test_thermo_reversible_oracle_against_live_reference_if_present():
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("reference tree not present")
    import tempfile
    import warnings
    from tropical_cyclone_risk_b200 import synth_thermo
    ref = rh.load_reference()
    nl = ref.namelist
    _, table = _thermo_rev_golden()
    p, ta, hus, sst, psl = synth_thermo.soundings(4096, seed=123)
    shp = (64, 64)
    saved = (nl.select_thermo, nl.src_directory)
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "thermo"))
        np.savez(os.path.join(tmp, "thermo", "entropy_table_reversible.npz"), p=table[0], s=table[1], rt=table[2], T=table[3])
        nl.select_thermo, nl.src_directory = 2, tmp
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                want = ref.thermo.CAPE_PI_vectorized(sst.reshape(shp), psl.reshape(shp), p.copy(), ta.astype(np.float64).reshape((-1,) + shp),
                                                     hus.astype(np.float64).reshape((-1,) + shp))
        finally:
            nl.select_thermo, nl.src_directory = saved
    v, _, _ = po.thermo(p, ta, hus, sst, psl, table, nl.Ck / nl.Cd, 13)
    _assert_close_to_reference(v, want.reshape(-1), 1e-10, 1e-10, "vmax")


def test_compute_thermo_refuses_what_the_reference_cannot_run():
    """select_interp = 1 has no branch in CAPE_PI_vectorized (thermo.py:266: the table is not loaded and f_lookup fails);
    select_thermo outside {1, 2} leaves s undefined in s_unsat (thermo.py:53-60)."""
    import types
    from tropical_cyclone_risk_b200 import preproc
    from tropical_cyclone_risk_b200 import namelist as nl
    z = np.zeros((1, 2, 2, 2), dtype=np.float32)
    bad = types.SimpleNamespace(**{k: getattr(nl, k) for k in dir(nl) if not k.startswith("__")})
    bad.select_interp = 1
    with pytest.raises(NotImplementedError):
        preproc.compute_thermo(None, z[:, 0], z[:, 0], z, z, [1000, 900], bad)
    bad.select_interp, bad.select_thermo = 2, 3
    with pytest.raises(ValueError):
        preproc.compute_thermo(None, z[:, 0], z[:, 0], z, z, [1000, 900], bad)


def test_order_levels():
    from tropical_cyclone_risk_b200 import preproc
    ta = np.arange(2 * 3 * 2 * 2, dtype=np.float32).reshape(2, 3, 2, 2)
    p_env, k_mid, ta2, _ = preproc.order_levels([250, 600, 1000], "hPa", ta, ta, 60000.0)
    assert np.array_equal(p_env, [100000.0, 60000.0, 25000.0]) and k_mid == 1
    assert np.array_equal(ta2[:, 0], ta[:, 2])
    p_env, k_mid, ta2, _ = preproc.order_levels([100000, 85000, 60000], "Pa", ta, ta, 60000.0)
    assert k_mid == 2 and ta2 is not None and np.array_equal(ta2, ta)


@pytest.mark.gpu
def test_gpu_thermo_bit_exact_vs_oracle_and_close_to_reference(engine):
    g, table = _thermo_golden()
    engine.set_entropy_table(*table)
    got = engine.thermo_month(g["p"], g["ta"], g["hus"], g["sst"], g["psl"], float(g["cecd"]), int(g["k_mid"]))
    want = po.thermo(g["p"], g["ta"], g["hus"], g["sst"], g["psl"], table, float(g["cecd"]), int(g["k_mid"]))
    for name, a, b in zip(("vmax", "chi", "rh_mid"), got, want):
        assert np.array_equal(a, b, equal_nan=True), name
    _assert_close_to_reference(got[0], g["vmax"], 1e-9, 1e-9, "vmax vs reference")
    _assert_close_to_reference(got[1], g["chi"], 1e-9, 1e-12, "chi vs reference")
    _assert_close_to_reference(got[2], g["rh_mid"], 1e-12, 0, "rh_mid vs reference")


@pytest.mark.gpu
@pytest.mark.parametrize("n,nlev_cut", [(1, 0), (129, 0), (5000, 6), (40000, 0)])
def test_gpu_thermo_other_shapes(engine, n, nlev_cut):
    """Ragged column counts, fewer levels (the top ones dropped), fresh soundings."""
    from tropical_cyclone_risk_b200 import synth_thermo
    _, table = _thermo_golden()
    engine.set_entropy_table(*table)
    p, ta, hus, sst, psl = synth_thermo.soundings(n, seed=n, edge_cases=n >= 64)
    if nlev_cut:
        p, ta, hus = p[:-nlev_cut], ta[:-nlev_cut], hus[:-nlev_cut]
    got = engine.thermo_month(p, ta, hus, sst, psl, 0.9, 11)
    want = po.thermo(p, ta, hus, sst, psl, table, 0.9, 11)
    for name, a, b in zip(("vmax", "chi", "rh_mid"), got, want):
        assert np.array_equal(a, b, equal_nan=True), name


@pytest.mark.gpu
def test_gpu_compute_thermo_mirror(engine):
    """compute_thermo through the reference-shaped host function: ascending hPa levels (flipped like
    calc_thermo.py:50-55), Celsius SST on its own coarser grid, two time samples."""
    from tropical_cyclone_risk_b200 import fields, preproc, synth_thermo
    from tropical_cyclone_risk_b200 import namelist as nl
    _, table = _thermo_golden()
    engine.set_entropy_table(*table)
    nlat, nlon = 12, 20
    p, ta, hus, sst, psl = synth_thermo.soundings(2 * nlat * nlon, seed=5, edge_cases=False)
    ta = ta.reshape(-1, 2, nlat, nlon).transpose(1, 0, 2, 3)
    hus = hus.reshape(-1, 2, nlat, nlon).transpose(1, 0, 2, 3)
    psl = psl.reshape(2, nlat, nlon)
    lat, lon = np.linspace(-30, 30, nlat), np.linspace(100, 160, nlon)
    slat, slon = np.linspace(-32, 32, 9), np.linspace(98, 162, 14)
    sst_c = (28.0 - 0.01 * slat[:, None] ** 2 + 0.02 * slon[None, :]).astype(np.float32)[None].repeat(2, 0)
    sst_c[0, 0, 0] = np.nan
    levels = (p / 100.0)[::-1]
    got = preproc.compute_thermo(engine, sst_c, psl, ta[:, ::-1], hus[:, ::-1], levels, nl, "hPa", "degC", slon, slat, lon, lat)
    for i in range(2):
        s = fields.regrid(slon, slat, np.nan_to_num(sst_c[i].astype(np.float64)), lon, lat) + 273.15
        want = po.thermo(p, ta[i], hus[i], s, psl[i], table, nl.Ck / nl.Cd, 13)
        for a, b in zip(got, want):
            assert np.array_equal(a[i].reshape(-1), b, equal_nan=True)
    assert got[0].shape == (2, nlat, nlon) and (got[0] > 0).any()


@pytest.mark.gpu
def test_gpu_thermo_reversible_bit_exact_vs_oracle_and_close_to_reference(engine):
    """namelist.select_thermo = 2: k_thermo<true> against the oracle bit for bit and against the unmodified reference."""
    g, table = _thermo_rev_golden()
    engine.set_entropy_table_reversible(*table)
    got = engine.thermo_month(g["p"], g["ta"], g["hus"], g["sst"], g["psl"], float(g["cecd"]), int(g["k_mid"]), select_thermo=2)
    want = po.thermo(g["p"], g["ta"], g["hus"], g["sst"], g["psl"], table, float(g["cecd"]), int(g["k_mid"]))
    for name, a, b in zip(("vmax", "chi", "rh_mid"), got, want):
        assert np.array_equal(a, b, equal_nan=True), name
    _assert_close_to_reference(got[0], g["vmax"], 1e-9, 1e-9, "vmax vs reference")
    _assert_close_to_reference(got[1], g["chi"], 1e-9, 1e-12, "chi vs reference")
    _assert_close_to_reference(got[2], g["rh_mid"], 1e-12, 0, "rh_mid vs reference")
    # the pseudoadiabatic table of the same handle is untouched by the reversible one
    g1, t1 = _thermo_golden()
    engine.set_entropy_table(*t1)
    a = engine.thermo_month(g1["p"], g1["ta"], g1["hus"], g1["sst"], g1["psl"], float(g1["cecd"]), int(g1["k_mid"]))
    b = po.thermo(g1["p"], g1["ta"], g1["hus"], g1["sst"], g1["psl"], t1, float(g1["cecd"]), int(g1["k_mid"]))
    assert np.array_equal(a[0], b[0], equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("n,nlev_cut", [(1, 0), (129, 0), (5000, 6), (40000, 0)])
def test_gpu_thermo_reversible_other_shapes(engine, n, nlev_cut):
    from tropical_cyclone_risk_b200 import synth_thermo
    _, table = _thermo_rev_golden()
    engine.set_entropy_table_reversible(*table)
    p, ta, hus, sst, psl = synth_thermo.soundings(n, seed=1000 + n, edge_cases=n >= 64)
    if nlev_cut:
        p, ta, hus = p[:-nlev_cut], ta[:-nlev_cut], hus[:-nlev_cut]
    got = engine.thermo_month(p, ta, hus, sst, psl, 0.9, 11, select_thermo=2)
    want = po.thermo(p, ta, hus, sst, psl, table, 0.9, 11)
    for name, a, b in zip(("vmax", "chi", "rh_mid"), got, want):
        assert np.array_equal(a, b, equal_nan=True), name


@pytest.mark.gpu
def test_gpu_compute_thermo_mirror_reversible(engine):
    """compute_thermo with a namelist that says select_thermo = 2."""
    import types
    from tropical_cyclone_risk_b200 import preproc, synth_thermo
    from tropical_cyclone_risk_b200 import namelist as nl
    nl2 = types.SimpleNamespace(**{k: getattr(nl, k) for k in dir(nl) if not k.startswith("__")})
    nl2.select_thermo = 2
    _, table = _thermo_rev_golden()
    engine.set_entropy_table_reversible(*table)
    nlat, nlon = 6, 10
    p, ta, hus, sst, psl = synth_thermo.soundings(nlat * nlon, seed=6, edge_cases=False)
    got = preproc.compute_thermo(engine, sst.reshape(1, nlat, nlon), psl.reshape(1, nlat, nlon), ta.reshape(1, -1, nlat, nlon),
                                 hus.reshape(1, -1, nlat, nlon), p, nl2, "Pa", "K")
    want = po.thermo(p, ta, hus, sst, psl, table, nl.Ck / nl.Cd, 13)
    for a, b in zip(got, want):
        assert np.array_equal(a[0].reshape(-1), b, equal_nan=True)


@pytest.mark.gpu
def test_gpu_thermo_rejects_bad_input(engine):
    from tropical_cyclone_risk_b200._lib import TcrError
    from tropical_cyclone_risk_b200 import synth_thermo
    _, table = _thermo_golden()
    engine.set_entropy_table(*table)
    p, ta, hus, sst, psl = synth_thermo.soundings(64, seed=1)
    with pytest.raises(TcrError):
        engine.thermo_month(p[::-1], ta, hus, sst, psl, 1.0, 13)             # top level first
    with pytest.raises(TcrError):
        engine.thermo_month(p, ta, hus, sst, psl, 1.0, 99)


# ------------------------------------------------------------------------------------------------
# multi-rank drivers: months / time samples sharded over ranks, one all-gather (gloo, world 2, CPU)
# ------------------------------------------------------------------------------------------------
class _OracleEngine:
    """Stands in for the CUDA engine in the gloo test (no GPU here): same methods, oracle arithmetic."""

    def wind_stats(self, ua, va, iu, il, gs):
        return po.wind_stats(series_of(ua, va, iu, il), gs)

    def thermo_month(self, p_env, ta, hus, sst, psl, cecd, k_mid, select_thermo=1):
        _, table = _thermo_golden()
        out = po.thermo(p_env, ta, hus, sst, psl, table, cecd, k_mid)
        return tuple(o.reshape(np.shape(sst)) for o in out)


def _driver_inputs():
    from tropical_cyclone_risk_b200 import synth_thermo
    t0 = datetime.datetime(2001, 1, 1)
    times = [t0 + datetime.timedelta(hours=12 * k) for k in range(2 * 151)]         # Jan .. May, 2 x daily
    ua, va = synth_winds(len(times), 2, 5, 8, seed=8)
    months = [datetime.datetime(2001, m, 15) for m in range(1, 6)]
    p, ta, hus, sst, psl = synth_thermo.soundings(3 * 4 * 6, seed=3, edge_cases=False)
    ta = ta.reshape(-1, 3, 4, 6).transpose(1, 0, 2, 3)
    hus = hus.reshape(-1, 3, 4, 6).transpose(1, 0, 2, 3)
    return times, ua, va, months, p, ta, hus, sst.reshape(3, 4, 6), psl.reshape(3, 4, 6)


def _run_drivers():
    from tropical_cyclone_risk_b200 import preproc
    from tropical_cyclone_risk_b200 import namelist as nl
    times, ua, va, months, p, ta, hus, sst, psl = _driver_inputs()
    eng = _OracleEngine()
    w = preproc.gen_wind_mean_cov(eng, ua, va, times, [250, 850], months)
    v, c, r = preproc.gen_thermo(eng, sst, psl, ta, hus, p / 100.0, nl)
    return w, v, c, r


def _driver_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank, _run_drivers()))
    finally:
        dist.destroy_process_group()


def test_preproc_drivers_gloo_world2_match_single_rank():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_driver_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = _run_drivers()
    assert want[0].shape == (5, 14, 5, 8) and want[1].shape == (3, 4, 6)
    for r in (0, 1):
        for a, b in zip(got[r], want):
            assert np.array_equal(a, b, equal_nan=True)


# ------------------------------------------------------------------------------------------------
# full-size runs (native 0.25-degree ERA5 grid): every output of a random sample of grid points is
# checked against the oracle run on just those points -- the kernels are point-wise, so the result
# of a point must not depend on the size or tiling of the launch.
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("spd,grouped", [(2, False), (4, True)])
def test_gpu_wind_stats_full_grid_sampled_oracle(engine, spd, grouped):
    nlat, nlon, n_days = 721, 1440, 31
    n_time, n_pts = n_days * spd, nlat * nlon
    rng = np.random.default_rng(17)
    ua = rng.normal(0, 8, (n_time, 2, n_pts)).astype(np.float32)
    va = (0.4 * ua + rng.normal(0, 5, (n_time, 2, n_pts))).astype(np.float32)
    ua[rng.integers(0, n_time, 50), 0, rng.integers(0, n_pts, 50)] = np.nan
    gs = np.arange(0, n_time + 1, spd if grouped else 1, dtype=np.int32)
    got = engine.wind_stats(ua, va, 0, 1, gs)
    pick = np.unique(np.r_[0, 1, 63, 64, n_pts - 1, n_pts - 65, rng.integers(0, n_pts, 6000),
                           np.flatnonzero(np.isnan(ua[:, 0]).any(axis=0))])
    want = po.wind_stats([a[:, k][:, pick] for k in (0, 1) for a in (ua, va)], gs)
    assert np.array_equal(got[:, pick], want, equal_nan=True)
    assert np.isfinite(got).all(axis=0).mean() > 0.999
    # a covariance matrix: symmetric by construction, so check positive variances and |corr| <= 1 everywhere
    ok = np.isfinite(got).all(axis=0)
    var_u, var_v, cov_uv = got[4][ok], got[6][ok], got[5][ok]
    assert (var_u > 0).all() and (var_v > 0).all()
    n = gs.size - 1
    assert (np.abs(cov_uv) * (n - 1) / n <= np.sqrt(var_u * var_v) * (1 + 1e-12)).all()      # ddof 1 vs ddof 0 (env_wind.py:211,213)


@pytest.mark.gpu
def test_gpu_thermo_full_grid_sampled_oracle(engine):
    from tropical_cyclone_risk_b200 import synth_thermo
    _, table = _thermo_golden()
    engine.set_entropy_table(*table)
    n_pts, base = 721 * 1440, 16384
    p, ta, hus, sst, psl = synth_thermo.soundings(base, seed=23)
    rng = np.random.default_rng(5)
    perm = rng.integers(0, base, n_pts)                                   # every column of the grid = one of the soundings
    got = engine.thermo_month(p, ta[:, perm], hus[:, perm], sst[perm], psl[perm], 1.0, 13)
    want = po.thermo(p, ta, hus, sst, psl, table, 1.0, 13)                # the oracle on the 16384 distinct soundings
    for name, a, b in zip(("vmax", "chi", "rh_mid"), got, want):
        assert np.array_equal(a, b[perm], equal_nan=True), name
    assert (got[0] >= 0).all() and (got[0] > 30).mean() > 0.2
    ok = ~np.isnan(got[1])
    assert (got[1][ok] >= 0).all() and (got[1][ok] <= 10).all() and (got[2] >= 1e-5).all() and (got[2] <= 1).all()


@pytest.mark.gpu
def test_gpu_wind_stats_long_ungrouped_record_falls_back_to_smaller_tiles(engine):
    """3-hourly samples, no daily averaging: 248 single-sample groups do not fit a 64-point tile (254 KB of
    shared memory) -- the library picks 32-point tiles; an hourly month (744 samples) does not fit a tile at all and is
    streamed from global memory (k_wind_stats_stream) -- the reference's ungrouped semantics for any record length,
    with and without missing samples."""
    ua, va = synth_winds(248, 2, 9, 40, seed=6)
    gs = np.arange(249, dtype=np.int32)
    got = engine.wind_stats(ua, va, 0, 1, gs)
    assert np.array_equal(got, po.wind_stats(series_of(ua, va, 0, 1), gs))
    for nan_frac in (0.0, 0.05):
        ua, va = synth_winds(744, 2, 5, 67, seed=6, nan_frac=nan_frac)
        gs = np.arange(745, dtype=np.int32)
        got = engine.wind_stats(ua, va, 0, 1, gs)
        assert np.array_equal(got, po.wind_stats(series_of(ua, va, 0, 1), gs), equal_nan=True)
    got = engine.wind_stats(ua, va, 0, 1, np.arange(0, 745, 24, dtype=np.int32))          # the same record as 31 daily groups
    assert np.array_equal(got, po.wind_stats(series_of(ua, va, 0, 1), np.arange(0, 745, 24)), equal_nan=True)


def test_wind_oracle_nan_policy_matches_pandas():
    """An independent implementation of the same published semantics: pandas' groupby-mean (skipna),
    var(ddof=0) and DataFrame.cov (pairwise-complete observations, ddof=1) -- what xarray's
    groupby("time.day").mean / .var / xr.cov compute -- on series with missing samples."""
    import pandas as pd
    ua, va = synth_winds(40, 2, 2, 3, seed=12, nan_frac=0.2)
    s = series_of(ua, va, 0, 1)
    gs = np.arange(0, 41, 4)
    got = po.wind_stats(s, gs)
    day = np.repeat(np.arange(10), 4)
    for pt in range(s[0].shape[1]):
        df = pd.DataFrame({k: a[:, pt].astype(np.float64) for k, a in enumerate(s)})
        dm = df.groupby(day).mean()                                   # daily means, NaN where a day has no valid sample
        np.testing.assert_allclose(got[:4, pt], dm.mean().to_numpy(), rtol=1e-13, equal_nan=True)
        cov = dm.cov(ddof=1).to_numpy()
        var0 = dm.var(ddof=0).to_numpy()
        k = 4
        for i in range(4):
            for j in range(i + 1):
                want = var0[i] if i == j else cov[i, j]
                np.testing.assert_allclose(got[k, pt], want, rtol=1e-11, atol=1e-12, equal_nan=True)
                k += 1


# ------------------------------------------------------------------------------------------------
# the reference's OWN calc_wnd_stat (unmodified function body, run over the NumPy-backed xarray
# stand-in of oracle/xr_shim.py by oracle/make_golden_windstats.py): pins the month mask, the
# day-grouping condition, the level selection, the order of the 14 statistics and their ddof.
# ------------------------------------------------------------------------------------------------
WINDSTAT_CASES = ["era5_2x_daily", "five_daily", "era5_with_nans"]


def _windstat_case(name):
    from conftest import golden
    g = golden("ref_windstats.npz")
    times = g[name + "_times"].astype("datetime64[s]")
    month = g[name + "_month"]
    return (g[name + "_ua"], g[name + "_va"], times, g[name + "_levels"], str(g[name + "_units"]),
            datetime.datetime(int(month[0]), int(month[1]), 15), g[name + "_stats"])


@pytest.mark.parametrize("name", WINDSTAT_CASES)
def test_calc_wnd_stat_mirror_matches_reference_function(name):
    from tropical_cyclone_risk_b200 import preproc
    ua, va, times, levels, units, dt, want = _windstat_case(name)
    got = preproc.calc_wnd_stat(_OracleEngine(), ua, va, times, levels, dt, level_units=units)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-13)


def test_reference_calc_wnd_stat_live_if_present():
    """Regenerates one fixture case through the live reference (build container only)."""
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("reference tree not present")
    from oracle import make_golden_windstats as mg
    from oracle import xr_shim
    ref = rh.load_reference()
    xr_shim.install(ref.env_wind.xr)
    ua, va, times, levels, units, dt, want = _windstat_case("era5_with_nans")
    got = mg.reference_stats(ref, ua, va, times.astype("datetime64[ns]"), list(levels.astype(int)), units, dt)
    assert np.array_equal(got, want, equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", WINDSTAT_CASES)
def test_gpu_calc_wnd_stat_matches_reference_function(engine, name):
    from tropical_cyclone_risk_b200 import preproc
    ua, va, times, levels, units, dt, want = _windstat_case(name)
    got = preproc.calc_wnd_stat(engine, ua, va, times, levels, dt, level_units=units)
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-13)
    # and bit for bit what the oracle gives through the same host function
    assert np.array_equal(got, preproc.calc_wnd_stat(_OracleEngine(), ua, va, times, levels, dt, level_units=units), equal_nan=True)
