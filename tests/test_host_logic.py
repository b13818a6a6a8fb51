"""Host-side logic that needs no GPU: the C-ABI library loads and exports every declared symbol,
the product never touches the oracle, year sharding + the write-out all-gather over gloo
(world_size 2), the track-file schema, the namelist -> params mapping."""
import os
import re
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_abi_exports_every_declared_symbol():
    import ctypes
    from tropical_cyclone_risk_b200 import _lib, build
    so = ctypes.CDLL(build.build())
    hdr = open(os.path.join(ROOT, "include", "tcrisk.h")).read()
    declared = set(re.findall(r"^\s*(?:int64_t|int|const char\*)\s+(tcr_\w+)\s*\(", hdr, flags=re.M))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(so, name) is not None
    assert so.tcr_version() >= 100


def test_params_struct_matches_header():
    """ctypes mirror and the C struct agree on size (checked against nvcc's layout via sizeof export)."""
    import ctypes
    from tropical_cyclone_risk_b200 import params
    # 8-byte members only + 4 int32 -> no padding surprises; the numbers are those of include/tcrisk.h
    n_doubles = 8 + 2 * 5 + 4 + 1 + 4 + 2 + 1 + 7 + 7 + 5 + 4 + 15
    assert ctypes.sizeof(params.TcrParams) == n_doubles * 8 + 16
    assert ctypes.sizeof(params.TcrYearStats) == 9 * 8 + 8 + 8


def test_product_never_imports_oracle():
    """The CUDA product path must not import, link or dlopen anything under oracle/."""
    pkg = os.path.join(ROOT, "tropical_cyclone_risk_b200")
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|liborc|tcr_oracle|oracle/", re.M)
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not bad.search(text), fn


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tropical_cyclone_risk_b200 import compute
    with pytest.raises(RuntimeError):
        compute.run_tracks(2001, 4, compute.TC_Basin("NA"))


def test_basin_and_params():
    from tropical_cyclone_risk_b200 import compute, params
    from tropical_cyclone_risk_b200 import namelist as nl
    b = compute.TC_Basin("NA")
    assert b.get_bounds() == (260.0, 0.0, 360.0, 60.0)
    assert b.in_basin(300.0, 20.0, 1) and not b.in_basin(359.5, 20.0, 1)
    with pytest.raises(ValueError):
        compute.TC_Basin("XX")
    p = params.params_from_namelist(nl, "SI")
    assert (p.gen_lat_min, p.gen_lat_max) == (-45.0, 45.0)        # '0S' -> -0.0 counts as >= 0
    p = params.params_from_namelist(nl, "NA")
    assert p.n_steps == 361 and (p.gen_lat_min, p.gen_lat_max) == (3.0, 45.0)


def test_shard_years():
    from tropical_cyclone_risk_b200.compute import shard_years
    years = list(range(1979, 1990))
    got = sorted(y for r in range(4) for y in shard_years(years, r, 4))
    assert got == years
    assert shard_years(years, 1, 4) == [1980, 1984, 1988]


def _fake_run_years(years, n_tracks, b):
    """Deterministic stand-in for the GPU call: arrays are functions of the year only."""
    ns = 361
    ny = len(years)
    out = {}
    for k, shp in (("lon", (ny, n_tracks, ns)), ("lat", (ny, n_tracks, ns)), ("v", (ny, n_tracks, ns)),
                   ("m", (ny, n_tracks, ns)), ("vmax", (ny, n_tracks, ns)), ("env", (ny, n_tracks, ns, 4)),
                   ("tc_month", (ny, n_tracks)), ("n_seeds", (ny, 7, 12))):
        a = np.empty(shp)
        for i, y in enumerate(years):
            rng = np.random.default_rng(1000 * y + sum(map(ord, k)))        # process-independent
            a[i] = rng.random(shp[1:])
            if k in ("lon", "lat", "v", "m", "vmax"):
                a[i, :, 200:] = np.nan
        out[k] = a
    out["tc_basin"] = np.stack([np.full(n_tracks, y % 7, np.int32) for y in years]) if ny else np.zeros((0, n_tracks), np.int32)
    return out


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import types
        from tropical_cyclone_risk_b200 import compute
        from tropical_cyclone_risk_b200 import namelist as nl
        cfg = types.SimpleNamespace(**{k: getattr(nl, k) for k in dir(nl) if not k.startswith("__")})
        cfg.start_year, cfg.end_year, cfg.tracks_per_year = 2001, 2005, 3          # 5 years over 2 ranks: ragged
        compute.configure(namelist=cfg)
        out = compute.run_downscaling("NA", write=False, run_years_fn=_fake_run_years)
        q.put((rank, {k: v for k, v in out.items() if isinstance(v, np.ndarray)}))
    finally:
        dist.destroy_process_group()


def test_run_downscaling_gloo_world2_matches_single_rank():
    import types
    import torch.multiprocessing as mp
    from tropical_cyclone_risk_b200 import compute
    from tropical_cyclone_risk_b200 import namelist as nl
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    cfg = types.SimpleNamespace(**{k: getattr(nl, k) for k in dir(nl) if not k.startswith("__")})
    cfg.start_year, cfg.end_year, cfg.tracks_per_year = 2001, 2005, 3
    compute.configure(namelist=cfg)
    try:
        want = compute.run_downscaling("NA", write=False, run_years_fn=_fake_run_years)
    finally:
        compute.configure(namelist=nl)
    for r in (0, 1):
        for k, v in got[r].items():
            assert np.array_equal(v, want[k], equal_nan=(v.dtype.kind == "f")), (r, k)
    assert want["tc_lon"].shape == (15, 361) and list(want["tc_years"][:4]) == [2001, 2001, 2001, 2002]
    assert want["tc_basins"].dtype == np.dtype("U2") and want["n_seeds"].shape == (5, 7, 12)


def test_trackfile_schema_roundtrip(tmp_path):
    """Variables / dims of the reference's output file (util/compute.py:250-262)."""
    import types
    from tropical_cyclone_risk_b200 import compute, trackfile
    from tropical_cyclone_risk_b200 import namelist as nl
    cfg = types.SimpleNamespace(**{k: getattr(nl, k) for k in dir(nl) if not k.startswith("__")})
    cfg.start_year, cfg.end_year, cfg.tracks_per_year = 2001, 2002, 4
    cfg.base_directory = cfg.output_directory = str(tmp_path)
    compute.configure(namelist=cfg)
    try:
        out = compute.run_downscaling("NA", write=True, run_years_fn=_fake_run_years)
        out2 = compute.run_downscaling("NA", write=True, run_years_fn=_fake_run_years)
    finally:
        compute.configure(namelist=nl)
    assert out["fn_trk_out"].endswith("tracks_NA_synthetic_200101_200212.nc")
    assert out2["fn_trk_out"].endswith("_e0.nc")                    # never overwrites (compute.py:52-58)
    f = trackfile.read_tracks(out["fn_trk_out"])
    want_vars = {"lon_trks", "lat_trks", "u250_trks", "v250_trks", "u850_trks", "v850_trks", "v_trks", "m_trks",
                 "vmax_trks", "tc_month", "tc_basins", "tc_years", "seeds_per_month", "n_trk", "time", "year",
                 "basin", "month"}
    assert set(f) == want_vars
    assert np.array_equal(f["lon_trks"], out["tc_lon"], equal_nan=True)
    assert np.array_equal(f["u850_trks"], out["tc_env_wnds"][:, :, 2], equal_nan=True)
    assert list(f["basin"]) == ["AU", "EP", "NA", "NI", "SI", "SP", "WP"]
    assert list(f["tc_basins"]) == list(out["tc_basins"])
    assert f["seeds_per_month"].shape == (2, 7, 12) and f["time"][-1] == 15 * 86400.0


def _merge_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tropical_cyclone_risk_b200 import gather
        rng = np.random.default_rng(3)
        full = rng.normal(size=(6, 50))
        full[:, 40:] = np.nan                                       # NaN padding
        full[1, 3], full[4, 7] = -0.0, 0.0                          # signed zeros survive an integer sum, not a float one
        block = np.zeros_like(full)
        block[rank::world] = full[rank::world]                      # rows this rank "integrated"; the others stay zero bits
        t = torch.from_numpy(block.reshape(-1).copy())
        gather.merge_sharded_block(t)
        q.put((rank, t.numpy().reshape(full.shape).view(np.int64).tolist(), full.view(np.int64).tolist()))
    finally:
        dist.destroy_process_group()


def test_merge_sharded_block_is_exact_gloo_world2():
    """The write-out merge of within-year sharding: integer sum of float64 bit patterns over ranks (gloo, world 2)."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_merge_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, merged, full in got:
        assert merged == full
