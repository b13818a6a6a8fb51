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
