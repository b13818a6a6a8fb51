"""The seeding / rejection loop and the per-month field preparation (row N1) against fixtures
produced by EXECUTING THE REFERENCE'S OWN SOURCE LINES (util/compute.py:123-209 and :76-84, 101-121) under an indexed
random stream / xarray shims -- oracle/make_golden_loop.py.  CPU side: the oracle and the host mirror; the CUDA
path is compared with the same fixtures in tests/test_gpu_parity.py."""
import numpy as np
import pytest

from conftest import Case, golden
from oracle import tcr_oracle as orc

SEED_CASES = {"NA": ("na_year", None), "GL": ("gl_year", None), "SI": (None, ("SI", 2002))}


def seed_case(basin, request):
    fixture, spec = SEED_CASES[basin]
    return request.getfixturevalue(fixture) if fixture else Case(spec[0], [spec[1]])


def check_seeding(g, basin, case, rec):
    """rec: per-attempt arrays code, basin, month, lon, lat, v0, m0 of the implementation under test."""
    year, run_seed, n_att, longest = (int(x) for x in g["seed_%s_meta" % basin])
    calls = g["seed_%s_calls" % basin]
    assert longest <= case.p.max_redraws
    assert rec["code"].size == n_att and not (rec["code"] == 3).any()
    passed = np.flatnonzero(rec["code"] == 2)
    # the reference called gen_track for exactly these attempts, in this order (util/compute.py:176)
    assert np.array_equal(passed, calls[:, 0].astype(np.int64))
    assert np.array_equal(rec["month"][passed], calls[:, 1].astype(np.int32))
    for j, key in ((2, "lon"), (3, "lat"), (4, "v0"), (5, "m0")):
        err = np.abs(rec[key][passed] - calls[:, j]) / np.maximum(np.abs(calls[:, j]), 1e-3)
        assert err.max() < 1e-12, (key, float(err.max()))
    hbl = np.asarray(case.p.atm_bl_depth)[rec["basin"][passed]]          # fast.h_bl = atm_bl_depth[basin] (:175)
    assert np.array_equal(hbl, calls[:, 6])
    counted = (rec["code"] == 1) | (rec["code"] == 2)                     # n_seeds[basin, month - 1] += 1 (:167)
    n_seeds = np.zeros((7, 12))
    np.add.at(n_seeds, (rec["basin"][counted], rec["month"][counted] - 1), 1)
    assert np.array_equal(n_seeds, g["seed_%s_n_seeds" % basin])
    return passed.size


@pytest.mark.parametrize("basin", ["NA", "GL", "SI"])
def test_oracle_seeding_vs_reference_loop(basin, request):
    g = golden("ref_loop.npz")
    case = seed_case(basin, request)
    year, run_seed, n_att, _ = (int(x) for x in g["seed_%s_meta" % basin])
    o = orc.run_attempts(case.p, case.env, 0, case.masks, run_seed, year, 0, n_att, want_tracks=False, n_threads=8)
    rec = dict(code=o["code"], basin=o["basin"], month=o["month"], lon=o["ic"][:, 0], lat=o["ic"][:, 1],
               v0=o["ic"][:, 2], m0=o["ic"][:, 3])
    n = check_seeding(g, basin, case, rec)
    print("seeding[%s]: %d attempts, %d gen_track calls identical to the reference loop's" % (basin, n_att, n))


def check_loop(g, basin, got, attempts):
    """got: the 9-tuple arrays of one year (lon, lat, v, m, vmax [n_tracks][ns], env, tc_month, tc_basin, n_seeds)."""
    year, run_seed, n_tracks, ref_attempts = (int(x) for x in g["loop_%s_meta" % basin])
    tag = "loop_%s_" % basin
    assert attempts == ref_attempts                                       # the loop stopped at the same attempt
    assert np.array_equal(got["n_seeds"], g[tag + "n_seeds"])
    assert np.array_equal(got["tc_month"], g[tag + "tc_month"])
    assert np.array_equal(got["tc_basin"], g[tag + "tc_basin"])
    # kept storms live long, and the reference's own 1-ulp twins of some of them spread beyond 1e-4 before they end
    # (`chaos`, a running maximum per sample): full bar where the reference reproduces itself, 4 x its own spread after
    chaos = g[tag + "chaos"].astype(np.float64)
    strict = 4.0 * np.nan_to_num(chaos, nan=0.0, posinf=np.inf) <= 1e-4
    worst = 0.0
    for key, ref in (("lon", "tc_lon"), ("lat", "tc_lat"), ("v", "tc_v"), ("m", "tc_m"), ("vmax", "tc_vmax"), ("env", "tc_env_wnds")):
        a, b = got[key], g[tag + ref]
        assert np.array_equal(np.isnan(a), np.isnan(b)), key                # same track lengths, NaN padded
        ok = ~np.isnan(b)
        err = np.where(ok, np.abs(a - b) / np.maximum(np.abs(b), 1.0 if key == "env" else 1e-3), 0.0)
        if key == "env":
            err = err.max(axis=-1)
            ok = ok.all(axis=-1)
        worst = max(worst, float(err[strict & ok].max()))
        assert err[strict & ok].max() < 1e-4, (key, float(err[strict & ok].max()))
        if key in ("lon", "lat", "v", "m"):
            loose = ok & ~strict & np.isfinite(chaos)
            assert np.all(err[loose] <= 4.0 * chaos[loose]), key
    if "attempt" in got:
        assert np.array_equal(got["attempt"], g[tag + "attempt"])            # the same attempts produced the kept storms
    return worst


def test_oracle_loop_vs_reference_loop(na_year):
    g = golden("ref_loop.npz")
    year, run_seed, n_tracks, _ = (int(x) for x in g["loop_NA_meta"])
    w = orc.run_year(na_year.p, na_year.env, 0, na_year.masks, run_seed, year, n_tracks, chunk=1024, n_threads=8)
    worst = check_loop(g, "NA", w, w["stats"]["attempts"])
    strict = 4.0 * np.nan_to_num(g["loop_NA_chaos"].astype(np.float64), nan=np.inf) <= 1e-4
    print("loop[NA]: %d tracks after %d attempts; n_seeds, months, basins, attempts identical; %d of %d samples where the "
          "reference reproduces itself to 1e-4: worst %.1e" % (n_tracks, w["stats"]["attempts"], int(strict.sum()),
                                                              int(np.isfinite(g["loop_NA_chaos"]).sum()), worst))


def test_host_field_preparation_vs_reference_lines():
    """fields.prepare_month (what feeds tcr_upload_months, and the checker of k_prepare_month) against
    util/compute.py:76-84,101-121 executed over shims: NaN policies, PI scaling, chi transform, ocean regrid,
    latitude flip."""
    from oracle import make_golden_loop as mgl
    from tropical_cyclone_risk_b200 import fields, layout
    from tropical_cyclone_risk_b200 import namelist as nl
    g = golden("ref_prep.npz")
    lon, lat, olon, olat, raws, stack, mld, strat = mgl.prep_inputs(int(g["year"]))
    for j, i in enumerate(g["months"]):
        raw = dict(raws[i], vmax=stack["vmax"][i], chi=stack["chi"][i], rh_mid=stack["rh_mid"][i])
        # stored latitude DESCENDING, as the reference's flip (compute.py:80-84) expects to undo
        raw_desc = {k: v[::-1, :] for k, v in raw.items()}
        lon_b, lat_b, planes = fields.prepare_month(nl, (0.0, -90.0, 360.0, 90.0), lon, lat[::-1], raw_desc, olon, olat,
                                                    mld[:, :, i], strat[:, :, i])
        assert np.array_equal(lat_b, lat) and np.array_equal(lon_b, lon)
        for key, ch in (("chi", layout.CH_CHI), ("vpot", layout.CH_VPOT), ("mld", layout.CH_MLD), ("strat", layout.CH_STRAT),
                        ("rh", layout.CH_RH)):
            ref = g["ref_" + key][j]
            assert np.array_equal(planes[ch], ref.astype(np.float32), equal_nan=True), (int(i), key)
