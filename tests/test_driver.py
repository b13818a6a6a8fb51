"""run.py / compute_downscaling_inputs mirror (driver.py): file discovery, CF decoding, month stamps, cache files.
No GPU here: an engine stand-in with the oracle's arithmetic takes the place of the CUDA engine (the kernels themselves
are checked bit for bit in tests/test_preproc.py)."""
import datetime
import types

import numpy as np
import pytest
from scipy.io import netcdf_file

from oracle import preproc_oracle as po


class OracleEngine:
    def __init__(self, table):
        self.table = table

    def wind_stats(self, ua, va, iu, il, gs):
        n = ua.shape[0]
        return po.wind_stats([a[:, k].reshape(n, -1) for k in (iu, il) for a in (ua, va)], gs)

    def thermo_month(self, p_env, ta, hus, sst, psl, cecd, k_mid, select_thermo=1):
        out = po.thermo(p_env, ta, hus, sst, psl, self.table, cecd, k_mid)
        return tuple(o.reshape(np.shape(sst)) for o in out)


def _namelist(tmp_path):
    from tropical_cyclone_risk_b200 import namelist as nl
    cfg = types.SimpleNamespace(**{k: getattr(nl, k) for k in dir(nl) if not k.startswith("__")})
    cfg.base_directory = str(tmp_path / "in")
    cfg.output_directory = str(tmp_path / "out")
    cfg.exp_prefix, cfg.dataset_type = "era5", "ERA5"
    cfg.var_keys = {'ERA5': {'sst': 'sst', 'mslp': 'sp', 'temp': 't', 'sp_hum': 'q', 'u': 'u', 'v': 'v',
                             'lvl': 'level', 'lon': 'longitude', 'lat': 'latitude'}}
    cfg.start_year, cfg.start_month, cfg.end_year, cfg.end_month = 2001, 1, 2001, 4
    return cfg


LAT = np.linspace(40, -40, 9)                    # ERA5 files run north to south
LON = np.arange(0.0, 360.0, 30.0)
HOURS0 = (datetime.datetime(2001, 1, 1) - datetime.datetime(1900, 1, 1)).total_seconds() / 3600.0


def _write(path, name, times_h, data, levels=None, packed=False, units="m s**-1"):
    with netcdf_file(str(path), "w", version=2) as f:
        f.createDimension("time", len(times_h)); f.createDimension("latitude", LAT.size); f.createDimension("longitude", LON.size)
        t = f.createVariable("time", "i4", ("time",)); t.units = "hours since 1900-01-01 00:00:00.0"; t.calendar = "gregorian"
        t[:] = np.asarray(times_h, dtype=np.int32)
        f.createVariable("latitude", "f4", ("latitude",))[:] = LAT
        f.createVariable("longitude", "f4", ("longitude",))[:] = LON
        dims = ("time", "latitude", "longitude")
        if levels is not None:
            f.createDimension("level", len(levels))
            lv = f.createVariable("level", "i4", ("level",)); lv.units = "millibars"; lv[:] = levels
            dims = ("time", "level", "latitude", "longitude")
        if packed:                                                        # ERA5 style: int16 + scale_factor / add_offset / _FillValue
            sf, ao = 0.002, 3.0
            v = f.createVariable(name, "i2", dims); v.scale_factor = sf; v.add_offset = ao; v._FillValue = np.int16(-32767); v.units = units
            v[:] = np.clip(np.rint((data - ao) / sf), -32766, 32767).astype(np.int16)
        else:
            v = f.createVariable(name, "f4", dims); v.units = units
            v[:] = data.astype(np.float32)


@pytest.fixture()
def era5_tree(tmp_path):
    from tropical_cyclone_risk_b200 import synth_thermo
    cfg = _namelist(tmp_path)
    (tmp_path / "in" / "winds").mkdir(parents=True)
    rng = np.random.default_rng(4)
    n_days = 31 + 28 + 31 + 30 + 10                                       # Jan 1 .. May 10, 2 x daily
    th = HOURS0 + 12 * np.arange(2 * n_days)
    levels = [250, 850]
    u = rng.normal(3, 8, (th.size, 2, LAT.size, LON.size))
    v = 0.3 * u + rng.normal(0, 5, u.shape)
    _write(tmp_path / "in" / "winds" / "era5_u_daily_2001.nc", "u", th, u, levels, packed=True)
    _write(tmp_path / "in" / "winds" / "era5_v_daily_2001.nc", "v", th, v, levels, packed=True)
    # monthly thermodynamic inputs: Jan .. May, 28 levels 70 .. 1000 hPa (ascending pressure, like the ERA5 files)
    tm = [HOURS0 + 24 * d for d in (0, 31, 59, 90, 120)]
    n = LAT.size * LON.size
    p, ta, hus, sst, psl = synth_thermo.soundings(5 * n, seed=9, edge_cases=False)
    ta = ta.reshape(28, 5, LAT.size, LON.size).transpose(1, 0, 2, 3)[:, ::-1]
    hus = hus.reshape(28, 5, LAT.size, LON.size).transpose(1, 0, 2, 3)[:, ::-1]
    lv = (p / 100.0)[::-1].astype(int)
    _write(tmp_path / "in" / "era5_t_monthly.nc", "t", tm, ta, lv, units="K")
    _write(tmp_path / "in" / "era5_q_monthly.nc", "q", tm, hus, lv, units="kg kg**-1")
    _write(tmp_path / "in" / "era5_sst_monthly.nc", "sst", tm, sst.reshape(5, LAT.size, LON.size), units="K")
    _write(tmp_path / "in" / "era5_sp_monthly.nc", "sp", tm, psl.reshape(5, LAT.size, LON.size), units="Pa")
    return cfg, dict(th=th, p=p, ta=ta, hus=hus, sst=sst.reshape(5, LAT.size, LON.size), psl=psl.reshape(5, LAT.size, LON.size))


def test_glob_prefix_and_names(era5_tree):
    from tropical_cyclone_risk_b200 import driver
    cfg, _ = era5_tree
    assert [p.split("/")[-1] for p in driver.glob_prefix(cfg, "u")] == ["era5_u_daily_2001.nc"]
    assert [p.split("/")[-1] for p in driver.glob_prefix(cfg, "sst")] == ["era5_sst_monthly.nc"]
    assert driver.get_env_wnd_fn(cfg).endswith("out/env_wnd_era5_200101_200104.nc")
    assert driver.get_fn_thermo(cfg).endswith("out/thermo_era5_200101_200104.nc")
    assert driver.get_bounding_times(cfg) == (datetime.datetime(2001, 1, 1), datetime.datetime(2001, 4, 30))


def test_compute_downscaling_inputs_writes_the_reference_caches(era5_tree, capsys):
    from conftest import golden
    from tropical_cyclone_risk_b200 import driver, layout, refdata
    cfg, src = era5_tree
    g = golden("ref_thermo.npz")
    table = (g["table_p"], g["table_s"], g["table_T"])
    eng = OracleEngine(table)
    driver.compute_downscaling_inputs(eng, cfg)
    assert "Saved" in capsys.readouterr().out
    # ---- wind statistics: month stamps of wnd_stat_wrapper (first = the start date itself, then the 15th) ----
    w = refdata._Cache(driver.get_env_wnd_fn(cfg), layout.FIELD_NAMES[:14])
    assert w.times == [datetime.datetime(2001, 1, 1), datetime.datetime(2001, 2, 15), datetime.datetime(2001, 3, 15),
                       datetime.datetime(2001, 4, 15)]
    assert np.array_equal(w.lat, LAT.astype(np.float32).astype(np.float64))
    su = driver._Source(cfg, driver.glob_prefix(cfg, "u")[0], "u", True)
    sv = driver._Source(cfg, driver.glob_prefix(cfg, "v")[0], "v", True)
    assert su.data.dtype == np.float32 and su.level_units == "millibars"
    feb = [i for i, t in enumerate(su.times) if t.month == 2]
    want = po.wind_stats([a[feb][:, k].reshape(len(feb), -1) for k in (0, 1) for a in (su.data, sv.data)], np.arange(len(feb) + 1))
    for i, name in enumerate(layout.FIELD_NAMES[:14]):
        assert np.array_equal(np.asarray(w.vars[name][1]).reshape(-1), want[i]), name
    # ---- thermodynamics: every sample inside the namelist's period, stamped the 15th, levels re-ordered ----
    t = refdata._Cache(driver.get_fn_thermo(cfg), ("vmax", "chi", "rh_mid"))
    assert t.times == [datetime.datetime(2001, m, 15) for m in (1, 2, 3, 4)]
    k = 2
    ta_k = src["ta"][k, ::-1].astype(np.float32)                         # as the kernel sees it: lowest level first
    hus_k = src["hus"][k, ::-1].astype(np.float32)
    wv, wc, wr = po.thermo(src["p"], ta_k, hus_k, src["sst"][k].astype(np.float32).astype(np.float64),
                           src["psl"][k].astype(np.float32).astype(np.float64), table, cfg.Ck / cfg.Cd, 13)
    assert np.array_equal(np.asarray(t.vars["vmax"][k]).reshape(-1), wv)
    assert np.array_equal(np.asarray(t.vars["chi"][k]).reshape(-1), wc, equal_nan=True)
    assert np.array_equal(np.asarray(t.vars["rh_mid"][k]).reshape(-1), wr)
    # ---- a second call finds the caches and does nothing (env_wind.py:85-87, calc_thermo.py:80-81) ----
    driver.compute_downscaling_inputs(eng, cfg)
    assert "Saved" not in capsys.readouterr().out


def test_decode_cf_masks_and_scales():
    from tropical_cyclone_risk_b200 import driver
    raw = np.array([-32767, 0, 100], dtype=np.int16)
    out = driver.decode_cf(raw, {"scale_factor": 0.5, "add_offset": 10.0, "_FillValue": np.int16(-32767)})
    # xarray's _choose_float_dtype: int16 WITH an add_offset decodes to float64 (classic packed ERA5) ...
    assert np.isnan(out[0]) and out[1] == 10.0 and out[2] == 60.0 and out.dtype == np.float64
    # ... int16 with a scale factor only to float32, float32 data stay float32, wider integers go to float64
    assert driver.decode_cf(raw, {"scale_factor": 0.5}).dtype == np.float32
    assert driver.decode_cf(np.array([1, 2], np.int32), {"scale_factor": 0.5}).dtype == np.float64
    f = driver.decode_cf(np.array([1.5, 2.5], np.float32), {})
    assert np.array_equal(f, [1.5, 2.5]) and f.dtype == np.float32


def _rank_worker(rank, world, port, base, q):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pathlib
        from conftest import golden
        from tropical_cyclone_risk_b200 import driver
        cfg = _namelist(pathlib.Path(base))
        cfg.output_directory = os.path.join(base, "out_w2")
        g = golden("ref_thermo.npz")
        driver.compute_downscaling_inputs(OracleEngine((g["table_p"], g["table_s"], g["table_T"])), cfg)
        dist.barrier()
        q.put(rank)
    finally:
        dist.destroy_process_group()


def test_compute_downscaling_inputs_gloo_world2_matches_single_rank(era5_tree, tmp_path):
    """Months / time samples sharded over two ranks (gloo), rank 0 writes: the same cache files as one rank."""
    import socket
    import torch.multiprocessing as mp
    from conftest import golden
    from tropical_cyclone_risk_b200 import driver, layout, refdata
    cfg, _ = era5_tree
    g = golden("ref_thermo.npz")
    driver.compute_downscaling_inputs(OracleEngine((g["table_p"], g["table_s"], g["table_T"])), cfg)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    assert sorted(q.get(timeout=180) for _ in range(2)) == [0, 1]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    cfg2 = _namelist(tmp_path)
    cfg2.output_directory = str(tmp_path / "out_w2")
    for fn, names in ((driver.get_env_wnd_fn, layout.FIELD_NAMES[:14]), (driver.get_fn_thermo, ("vmax", "chi", "rh_mid"))):
        a, b = refdata._Cache(fn(cfg), names), refdata._Cache(fn(cfg2), names)
        assert a.times == b.times
        for n in names:
            assert np.array_equal(a.vars[n], b.vars[n], equal_nan=True), n


def _run_worker(rank, world, port, base, q):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pathlib
        from conftest import golden
        from oracle import ref_harness as rh
        from test_host_logic import _fake_run_years
        from tropical_cyclone_risk_b200 import driver
        cfg = _namelist(pathlib.Path(base))
        cfg.output_directory = os.path.join(base, "out_run2")
        cfg.exp_name, cfg.tracks_per_year = "w2", 3
        g = golden("ref_thermo.npz")
        out = driver.run("NA", cfg, rh.REF_ROOT, engine=OracleEngine((g["table_p"], g["table_s"], g["table_T"])),
                         run_years_fn=_fake_run_years)
        q.put((rank, out["tc_lon"].shape, out.get("fn_trk_out")))
    finally:
        dist.destroy_process_group()


def test_driver_run_gloo_world2(era5_tree, tmp_path):
    """driver.run under two ranks (gloo): rank 0 writes the caches atomically, every rank waits for them behind a
    barrier before run_downscaling opens them, the years are sharded and gathered, rank 0 alone writes the track
    file.  (Build container only: the static inputs come from the reference tree.)"""
    import os
    import socket
    import torch.multiprocessing as mp
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("reference tree not present")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_run_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got[0][1] == got[1][1] == (3, 361)                          # 2001 only (start/end year of the test namelist) x 3 tracks
    assert got[0][2] is not None and got[1][2] is None                 # rank 0 wrote the track file, rank 1 did not
    assert os.path.exists(got[0][2])
    out = str(tmp_path / "out_run2")
    assert not [f for f in os.listdir(out) if ".tmp." in f]            # no half-written cache left behind


def test_the_reference_namelist_is_accepted_unchanged():
    """namelist.py of a reference checkout drives everything here as it is (build container only)."""
    import os
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("reference tree not present")
    from tropical_cyclone_risk_b200 import driver, params
    nl = driver.load_namelist(os.path.join(rh.REF_ROOT, "namelist.py"))
    p = params.params_from_namelist(nl, "NA")
    assert p.n_steps == 361 and p.dt_track == 3600.0
    assert params.basin_bounds(nl, "NA") == (260.0, 0.0, 360.0, 60.0)
    assert driver.get_env_wnd_fn(nl).endswith("env_wnd_era5_201601_202112.nc")
    assert driver.get_fn_thermo(nl).endswith("thermo_era5_201601_202112.nc")
    assert nl.var_keys[nl.dataset_type]["mslp"] == "sp" and nl.select_thermo == 1 and nl.select_interp == 2


def test_make_engine_loads_the_table_the_namelist_selects(tmp_path, monkeypatch):
    """driver.make_engine: thermo/entropy_table.npz for select_thermo = 1, thermo/entropy_table_reversible.npz for 2
    (thermo.py:274-284), through preproc.load_entropy_table."""
    import types
    from tropical_cyclone_risk_b200 import driver, engine, preproc
    from tropical_cyclone_risk_b200 import namelist as nl
    root = tmp_path / "ref"
    (root / "thermo").mkdir(parents=True)
    p, s, rt = np.linspace(2500.0, 105000.0, 5), np.linspace(2300.0, 3600.0, 4), np.linspace(0.0, 0.04, 3)
    np.savez(root / "thermo" / "entropy_table.npz", p=p, s=s, T=np.arange(20.0).reshape(5, 4))
    np.savez(root / "thermo" / "entropy_table_reversible.npz", p=p, s=s, rt=rt, T=np.arange(60.0).reshape(5, 4, 3))
    t2 = preproc.load_entropy_table(str(root / "thermo" / "entropy_table.npz"))
    t3 = preproc.load_entropy_table(str(root / "thermo" / "entropy_table_reversible.npz"))
    assert [a.shape for a in t2] == [(5,), (4,), (5, 4)] and [a.shape for a in t3] == [(5,), (4,), (3,), (5, 4, 3)]
    calls = []

    class FakeEngine:
        def __init__(self, params, device=0):
            calls.append(("create", device))

        def set_entropy_table(self, *table):
            calls.append(("pseudo", [np.shape(a) for a in table]))

        def set_entropy_table_reversible(self, *table):
            calls.append(("reversible", [np.shape(a) for a in table]))

    monkeypatch.setattr(engine, "Engine", FakeEngine)
    for sel, want in ((1, "pseudo"), (2, "reversible")):
        nl2 = types.SimpleNamespace(**{k: getattr(nl, k) for k in dir(nl) if not k.startswith("__")})
        nl2.select_thermo = sel
        del calls[:]
        driver.make_engine(nl2, str(root), device=3)
        assert calls[0] == ("create", 3) and calls[1][0] == want and len(calls) == 2
