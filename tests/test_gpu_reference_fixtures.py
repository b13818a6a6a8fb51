"""The CUDA path (through the C ABI) compared DIRECTLY with fixtures produced by the unmodified reference:

  * tests/golden/ref_wide_*.npz   510 storms, five basin-months (oracle/make_golden_wide.py); rules: tests/wide.py
  * tests/golden/ref_loop.npz     the reference's own seeding / acceptance loop lines, executed (oracle/make_golden_loop.py)
  * tests/golden/ref_prep.npz     the reference's own field-preparation lines, executed

The CPU oracle takes part only as the reporter of `n_clean` (where a storm first evaluates the exact land test inside
an all-land cell); every number compared comes from the GPU and from the reference."""
import numpy as np
import pytest

import wide
from conftest import Case, golden
from oracle import tcr_oracle as orc
from test_loop_golden import check_loop, check_seeding, seed_case

pytestmark = pytest.mark.gpu


def _engine(case):
    from tropical_cyclone_risk_b200.engine import Engine
    e = Engine(case.p, device=0)
    e.upload_case(case.lon, case.lat, case.planes, case.static, case.mask_lon, case.mask_lat, case.mask_planes)
    return e


@pytest.mark.parametrize("name", wide.WIDE_CASES)
def test_cuda_vs_wide_reference(name):
    g, case = wide.load_case(name)
    seeds = wide.seeds_of(g)
    eng = _engine(case)
    try:
        got = eng.integrate(*seeds)
    finally:
        eng.close()
    n_clean = orc.integrate_batch(case.p, case.env, *seeds, post_all=False, n_threads=8)["n_clean"]
    rep = wide.check(name, g, got, n_clean)
    assert rep["reference_stable"] >= 0.6 * rep["storms"] and rep["worst_stable_err"] < 1e-5


@pytest.mark.parametrize("basin", ["NA", "GL", "SI"])
def test_cuda_seeding_vs_reference_loop(basin, request):
    g = golden("ref_loop.npz")
    case = seed_case(basin, request)
    year, run_seed, n_att, _ = (int(x) for x in g["seed_%s_meta" % basin])
    eng = _engine(case)
    try:
        rec = eng.seed_attempts(0, year, run_seed, 0, n_att)
    finally:
        eng.close()
    n = check_seeding(g, basin, case, rec)
    print("cuda seeding[%s]: %d attempts, %d gen_track calls identical to the reference loop's" % (basin, n_att, n))


def test_cuda_year_vs_reference_loop(na_year):
    g = golden("ref_loop.npz")
    year, run_seed, n_tracks, _ = (int(x) for x in g["loop_NA_meta"])
    eng = _engine(na_year)
    try:
        r = eng.run_years([0], [year], run_seed, n_tracks)
    finally:
        eng.close()
    got = {k: r[k][0] for k in ("lon", "lat", "v", "m", "vmax", "env", "tc_month", "tc_basin", "n_seeds")}
    worst = check_loop(g, "NA", got, r["stats"][0]["attempts"])
    print("cuda loop[NA]: %d tracks after %d attempts, worst %.1e where the reference reproduces itself" % (
        n_tracks, r["stats"][0]["attempts"], worst))


def test_cuda_field_preparation_vs_reference_lines():
    """k_prepare_month against util/compute.py:76-84,101-121 executed over shims.  chi goes through tcr_libm's
    exp / log on the device (the reference: numpy's): at most one float32 ulp; the other planes are bit-identical."""
    from oracle import make_golden_loop as mgl
    from tropical_cyclone_risk_b200 import layout, params
    from tropical_cyclone_risk_b200 import namelist as nl
    from tropical_cyclone_risk_b200.engine import Engine
    g = golden("ref_prep.npz")
    lon, lat, olon, olat, raws, stack, mld, strat = mgl.prep_inputs(int(g["year"]))
    eng = Engine(params.params_from_namelist(nl, "GL"), device=0)
    try:
        for j, i in enumerate(g["months"]):
            raw = dict(raws[i], vmax=stack["vmax"][i], chi=stack["chi"][i], rh_mid=stack["rh_mid"][i])
            raw_desc = {k: np.ascontiguousarray(v[::-1, :]) for k, v in raw.items()}
            lon_b, lat_b, planes = eng.prepare_month(-1, nl, (0.0, -90.0, 360.0, 90.0), lon, lat[::-1].copy(), raw_desc, olon, olat,
                                                     np.ascontiguousarray(mld[:, :, i]), np.ascontiguousarray(strat[:, :, i]),
                                                     return_planes=True)
            assert np.array_equal(lat_b, lat) and np.array_equal(lon_b, lon)
            for key, ch in (("vpot", layout.CH_VPOT), ("mld", layout.CH_MLD), ("strat", layout.CH_STRAT), ("rh", layout.CH_RH)):
                assert np.array_equal(planes[ch], g["ref_" + key][j].astype(np.float32), equal_nan=True), (int(i), key)
            ref = g["ref_chi"][j].astype(np.float32)
            ulp = np.abs(planes[layout.CH_CHI].view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
            assert ulp.max() <= 1 and (ulp == 0).mean() > 0.99
    finally:
        eng.close()


# ---------------------------------------------------------------------------------------------
# the inner tier of the seam: the Coupled_FAST-compatible object (tropical_cyclone_risk_b200/coupled_fast.py)
# ---------------------------------------------------------------------------------------------
def _shim(monkeypatch):
    from oracle import ref_harness as rh                      # injected_phases only: nothing of the reference tree is read
    from tropical_cyclone_risk_b200 import compute, coupled_fast, fields, layout, synth
    from tropical_cyclone_risk_b200 import namelist as nl
    lon, lat = synth.era5_axes()
    olon, olat = synth.ocean_axes()
    raw = synth.synth_month_raw(2000, 9, lon, lat)
    mld, strat = synth.synth_ocean(olon, olat, 9)
    _, _, pg = fields.prepare_month(nl, (0.0, -90.0, 360.0, 90.0), lon, lat, raw, olon, olat, mld, strat)
    wnd = dict(lon=lon, lat=lat, **{name: pg[i] for i, name in enumerate(layout.FIELD_NAMES[:14])})
    monkeypatch.setattr(coupled_fast, "random_seed", lambda: None)
    import datetime
    fast = coupled_fast.Coupled_FAST(wnd, compute.TC_Basin("NA"), datetime.datetime(2000, 9, 15), nl.output_interval_s,
                                     nl.total_track_time_days * 86400, static=synth.synth_static(full_res=False), device=0)
    fast.init_fields(lon, lat, pg[layout.CH_CHI], pg[layout.CH_VPOT], pg[layout.CH_MLD], pg[layout.CH_STRAT])
    return fast, rh


def test_coupled_fast_shim_gen_track_vs_reference_fixture(monkeypatch, na_case):
    """fast.gen_track(clon, clat, v, m) -> .t / .y / .status / .nfev as the reference's run_tracks loop reads them
    (util/compute.py:176-203), against the outputs of the reference's own gen_track (tests/golden/ref_tracks.npz)."""
    fast, rh = _shim(monkeypatch)
    g = golden("ref_tracks.npz")
    try:
        assert fast.total_steps == 361 and fast.nWLvl == 4 and fast.t_s[-1] == 15 * 86400.0
        o = orc.integrate_batch(na_case.p, na_case.env, np.zeros(g["lon0"].size, np.int32), g["lon0"], g["lat0"], g["v0"], g["m0"],
                                g["h_bl"], g["phases"], post_all=False)
        n_none = n_res = 0
        for i in range(g["lon0"].size):
            fast.h_bl = float(g["h_bl"][i])
            with rh.injected_phases(g["phases"][i]):
                res = fast.gen_track(g["lon0"][i], g["lat0"][i], g["v0"][i], g["m0"][i])
            if g["status"][i] == 2:
                assert res is None                                        # ventilation pre-check -> None (coupled_fast.py:241-244)
                n_none += 1
                continue
            n_res += 1
            assert res.status == g["status"][i] and res.y.shape[0] == 4
            k = res.t.size
            # bit-identical to the batched C-ABI path (and so to the oracle), whose comparison with the reference's
            # numbers is test_cuda_vs_wide_reference / test_tracks_vs_reference_fixtures
            assert k == o["n_time"][i] and res.nfev == o["nfev"][i]
            assert np.array_equal(res.y.T, o["track"][i, :k]) and np.array_equal(res.t, fast.t_s[:k])
            kk = min(k, int(g["n_time"][i]), 24)                          # first day: before any storm turns chaotic
            err = np.abs(res.y.T[:kk] - g["track"][i, :kk]) / np.maximum(np.abs(g["track"][i, :kk]), 1e-3)
            assert err.max() < 1e-4
        assert n_none >= 1 and n_res >= 30
        # batched form: same results
        with rh.injected_phases(g["phases"][0]):
            one = fast.gen_track(g["lon0"][0], g["lat0"][0], g["v0"][0], g["m0"][0])
        many = fast.gen_tracks(g["lon0"][:5], g["lat0"][:5], g["v0"][:5], g["m0"][:5], h_bl=g["h_bl"][:5], phases=g["phases"][:5])
        fast.h_bl = float(g["h_bl"][0])
        assert (one is None) == (many[0] is None)
    finally:
        fast.close()


def test_coupled_fast_shim_dydt_env_winds_vpot_vs_reference_fixtures(monkeypatch):
    """fast.dydt, fast._env_winds (util/compute.py:201-202) and fast.f_vpot.ev (:162) against the reference's values."""
    fast, rh = _shim(monkeypatch)
    try:
        g = golden("ref_rhs.npz")
        worst = worst_w = 0.0
        for i in range(g["lon"].size):
            fast.h_bl = float(g["h_bl"][i])
            fast.Fs_phases = g["phases"][i].reshape(60)
            y = np.array([g["lon"][i], g["lat"][i], g["v"][i], g["m"][i]])
            dy = fast.dydt(g["t"][i], y)
            ref = g["dydt"][i]
            worst = max(worst, float((np.abs(dy - ref) / np.maximum(np.abs(ref), 1e-12 + 1e-6 * np.abs(ref).max())).max()))
            w = fast._env_winds(g["lon"][i], g["lat"][i], g["t"][i])
            rw = g["env_winds"][i]
            worst_w = max(worst_w, float((np.abs(w - rw) / np.maximum(np.abs(rw), 1e-6 + 1e-6 * np.abs(rw).max())).max()))
        assert worst < 1e-9 and worst_w < 1e-9, (worst, worst_w)
        b = golden("ref_bilinear.npz")
        v = fast.f_vpot.ev(b["lon"], b["lat"])
        ref = b["vals"][:, 15]
        assert np.max(np.abs(v - ref) / np.maximum(np.abs(ref), 1e-3 * np.abs(ref).max())) < 1e-12
        assert np.isscalar(fast.f_vpot.ev(300.0, 20.0)) or np.ndim(fast.f_vpot.ev(300.0, 20.0)) == 0
    finally:
        fast.close()
