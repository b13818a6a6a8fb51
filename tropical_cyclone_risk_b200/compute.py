"""Drop-in for the reference's per-year ensemble loop (util/compute.py): same names, argument
meaning and return contract, with the hot loop executed by libtcrisk.so on the GPU.

    run_tracks(year, n_tracks, b)  -> (tc_lon, tc_lat, tc_v, tc_m, tc_vmax, tc_env_wnds,
                                       tc_month, tc_basin, n_seeds)          util/compute.py:64,210
    run_downscaling(basin_id)      -> dict of the concatenated arrays the reference hands to
                                      xarray (util/compute.py:216-268), gathered over ranks
    get_fn_tracks / fn_tracks_duplicates                                    util/compute.py:40-58

What differs from the reference, by design:
  * the reference forks one dask *process* per year (compute.py:224-230); here every rank
    (one process per GPU, torch.distributed) integrates its share of the years in ONE
    `tcr_run_years` call and the finished tracks meet in a single all-gather at write-out;
  * random numbers are an indexed Philox stream keyed by (run_seed, year) instead of the
    wall-clock-seeded global MT19937 (track/bam_track.py:37-42), so results are reproducible
    and independent of the number of GPUs;
  * the monthly environment tables come from an *input provider* (`configure(inputs=...)`): the default is
    the synthetic ERA5-shaped generator; `refdata.ReferenceInputs` reads the reference's own static files and
    its env_wnd / thermo caches (NetCDF-4 through the built-in HDF5 reader `h5lite`, NetCDF-3 through SciPy).

There is no CPU fallback: without a CUDA device `run_tracks` raises.
"""
import os

import numpy as np

from . import layout, params
from . import namelist as default_namelist


class TC_Basin:
    """util/basins.py:11-50 (the parts the hot path uses)."""

    def __init__(self, basin_id, namelist=None):
        nl = namelist or default_namelist
        if basin_id.upper() not in nl.basin_bounds.keys():
            raise ValueError('Basin ID is not valid. See list of valid basins.')
        self.basin_id = basin_id
        self.basin_bounds = nl.basin_bounds[basin_id]

    def get_bounds(self):
        return tuple(params.parse_bound(s) for s in self.basin_bounds)

    def in_basin(self, clon, clat, dx):
        lon_min, lat_min, lon_max, lat_max = self.get_bounds()
        return ((lon_min + dx) < clon < (lon_max - dx) and (lat_min + dx) < clat < (lat_max - dx))


# ---------------------------------------------------------------------------------------------
# input providers
# ---------------------------------------------------------------------------------------------
class SyntheticInputs:
    """Prepared monthly planes from the deterministic ERA5-shaped generator (synth.py)."""

    def __init__(self, full_res=True, roughness=1.0):
        self.full_res = full_res
        self.roughness = roughness
        self._static = None

    def static(self):
        from . import synth
        if self._static is None:
            self._static = synth.synth_static(full_res=self.full_res)
        return self._static

    def year_planes(self, namelist, bounds, year):
        """(lon_b, lat_b, planes[12][19][nlat][nlon] float32) for one year."""
        from . import synth
        return synth.prepared_year(namelist, bounds, year, roughness=self.roughness)


class _Session:
    """One engine per (device, basin), tables for the years uploaded so far."""

    def __init__(self):
        self.namelist = default_namelist
        self.inputs = SyntheticInputs()
        self.run_seed = 20260101
        self.engines = {}

    def engine(self, basin_id, device):
        key = (basin_id, device)
        if key not in self.engines:
            from . import fields
            from .engine import Engine
            nl = self.namelist
            p = params.params_from_namelist(nl, basin_id)
            bounds = params.basin_bounds(nl, basin_id)
            eng = Engine(p, device=device)
            st = self.inputs.static()
            eng.upload_static(fields.prepare_static(bounds, st))
            mlon, mlat, m = fields.crop_masks(st, basin_id, bounds, p)
            eng.upload_masks(mlon, mlat, m)
            self.engines[key] = (eng, bounds)
        return self.engines[key]

    def close(self):
        for eng, _ in self.engines.values():
            eng.close()
        self.engines = {}


_session = _Session()


def configure(namelist=None, inputs=None, run_seed=None):
    """Replace the namelist module, the input provider and/or the run seed (closes open engines)."""
    _session.close()
    if namelist is not None:
        _session.namelist = namelist
    if inputs is not None:
        _session.inputs = inputs
    if run_seed is not None:
        _session.run_seed = int(run_seed)


def shutdown():
    _session.close()


def _current_device():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("tropical_cyclone_risk_b200: no CUDA device; the track generator has no CPU fallback")
    return torch.cuda.current_device()


def _basin_labels(idx):
    ids = np.array(layout.BASIN_IDS, dtype='U2')
    out = np.full(idx.shape, "", dtype='U2')
    ok = idx >= 0
    out[ok] = ids[idx[ok]]
    return out


def run_years(years, n_tracks, b, device=None, engine=None, bounds=None):
    """`run_tracks` for several years in one launch sequence: dict of year-major arrays
    (lon, lat, v, m, vmax [ny][n_tracks][n_steps], env [...][4], tc_month, tc_basin (index),
    n_seeds [ny][7][12], stats)."""
    years = [int(y) for y in years]
    if engine is None:
        device = _current_device() if device is None else device
        engine, bounds = _session.engine(b.basin_id, device)
    lon = lat = None
    planes = []
    for y in years:
        lon, lat, pl = _session.inputs.year_planes(_session.namelist, bounds, y)
        planes.append(pl)
    engine.alloc_tables(12 * len(years), lon, lat)
    for i, pl in enumerate(planes):
        engine.upload_months(12 * i, pl)
    ym_base = np.arange(len(years), dtype=np.int32) * 12
    return engine.run_years(ym_base, np.asarray(years, dtype=np.int32), _session.run_seed, n_tracks)


def run_tracks(year, n_tracks, b):
    """Generates `n_tracks` tropical cyclone tracks in basin `b` in the year (util/compute.py:64).
    Returns the reference's 9-tuple (util/compute.py:210)."""
    r = run_years([year], n_tracks, b)
    return (r["lon"][0], r["lat"][0], r["v"][0], r["m"][0], r["vmax"][0], r["env"][0],
            r["tc_month"][0], _basin_labels(r["tc_basin"][0]), r["n_seeds"][0])


# ---------------------------------------------------------------------------------------------
# multi-GPU: years sharded over ranks, one all-gather of the finished tracks at write-out
# ---------------------------------------------------------------------------------------------
def shard_years(years, rank, world):
    """Round-robin whole years over ranks -- the reference's one-process-per-year scheme
    (util/compute.py:224-230) with ranks in place of dask workers."""
    return [y for i, y in enumerate(years) if i % world == rank]


_FIELDS = ("lon", "lat", "v", "m", "vmax", "env", "tc_month", "tc_basin", "n_seeds")


def _pack(local, n_slots, n_tracks, n_steps):
    """Year-major flat float64 block of `n_slots` year slots (unused slots NaN)."""
    per_year = n_tracks * n_steps * 9 + n_tracks * 2 + 84
    buf = np.full((n_slots, per_year), np.nan)
    ny = 0 if local is None else local["lon"].shape[0]
    for y in range(ny):
        parts = [local[k][y].reshape(-1) for k in ("lon", "lat", "v", "m", "vmax", "env", "tc_month")]
        parts.append(local["tc_basin"][y].astype(np.float64))
        parts.append(local["n_seeds"][y].reshape(-1))
        buf[y] = np.concatenate(parts)
    return buf


def _unpack(row, n_tracks, n_steps):
    T, S = n_tracks, n_steps
    out, o = {}, 0
    for k, n, shp in (("lon", T * S, (T, S)), ("lat", T * S, (T, S)), ("v", T * S, (T, S)), ("m", T * S, (T, S)),
                      ("vmax", T * S, (T, S)), ("env", T * S * 4, (T, S, 4)), ("tc_month", T, (T,)),
                      ("tc_basin", T, (T,)), ("n_seeds", 84, (7, 12))):
        out[k] = row[o:o + n].reshape(shp)
        o += n
    out["tc_basin"] = out["tc_basin"].astype(np.int32)
    return out


def gather_years(local, years, n_tracks, n_steps, rank, world, device=None):
    """All-gather the per-rank year blocks and return them in `years` order on every rank.

    local: this rank's dict from run_years for shard_years(years, rank, world) (None if the rank
    owns no year).  Uses torch.distributed's default group when world > 1 (NCCL on GPUs, gloo
    on CPU); one collective, equal-sized slots (ceil(len(years)/world) year slots per rank)."""
    n_slots = (len(years) + world - 1) // world
    block = _pack(local, n_slots, n_tracks, n_steps)
    if world > 1:
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(block)
        if device is not None:
            t = t.to(device, non_blocking=True)
        g = torch.empty((world * t.shape[0], t.shape[1]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(g, t)
        allb = g.cpu().numpy().reshape(world, n_slots, -1)
    else:
        allb = block[None]
    per_year = [_unpack(allb[i % world, i // world], n_tracks, n_steps) for i in range(len(years))]
    return {k: np.stack([p[k] for p in per_year]) for k in _FIELDS}


def run_downscaling_device(years, n_tracks, b, rank, world, device):
    """The CUDA write-out path of run_downscaling: this rank's years are integrated into ONE device-resident block
    (equal-sized year slots, unused ones NaN), the blocks of all ranks are copied into rank 0's gather buffer by the
    copy engines over NVLink (gather.PeerGather, CUDA IPC; `world` device-to-device copies, no collective kernel), and
    rank 0 alone brings the gathered block to the host -- one device->host copy for the whole job.  Returns the
    year-ordered dict of arrays on rank 0 and None on the other ranks."""
    import torch
    from . import gather as tgather
    from .pipeline import _Block
    if world > 1 and len(years) < world and os.environ.get("TCR_SHARD", "auto") != "years":
        return _run_downscaling_within_year(years, n_tracks, b, rank, world, device)
    mine = shard_years(years, rank, world)
    n_slots = (len(years) + world - 1) // world
    engine, bounds = _session.engine(b.basin_id, device.index)
    stream = torch.cuda.current_stream(device)
    engine.set_stream(stream.cuda_stream)
    blk = _Block(torch, device, n_slots, n_tracks, engine.n_steps, pinned=False)
    blk.dev.fill_(float("nan"))
    stats = []
    if mine:
        lon = lat = None
        planes = []
        for y in mine:
            lon, lat, pl = _session.inputs.year_planes(_session.namelist, bounds, y)
            planes.append(pl)
        engine.alloc_tables(12 * len(mine), lon, lat)
        for i, pl in enumerate(planes):
            engine.upload_months(12 * i, pl)
        # the rank's years fill the first len(mine) year slots of the block (a prefix of every section)
        sub = _Block(torch, device, len(mine), n_tracks, engine.n_steps, pinned=False) if len(mine) != n_slots else blk
        stats = engine.run_years_dev(np.arange(len(mine), dtype=np.int32) * 12, np.asarray(mine, dtype=np.int32),
                                     _session.run_seed, n_tracks, sub.dptr)
        if sub is not blk:
            blk.copy_prefix_from(sub)
    if world > 1:
        pg = tgather.PeerGather(blk.dev.numel(), torch.float64, device, depth=1, dst="root")
        pg.push(blk.dev, 0)
        pg.finish()
        if rank != 0:
            return None
        host = pg.buffer(0).cpu().numpy()                       # the single device->host copy of the job
    else:
        host = blk.dev.cpu().numpy()[None]
    per_rank = [blk.views_of(host[r]) for r in range(world)]
    per_year = [{k: per_rank[i % world][k][i // world] for k in _FIELDS} for i in range(len(years))]
    out = {k: np.stack([p[k] for p in per_year]) for k in _FIELDS}
    out["stats"] = stats
    return out


def _run_downscaling_within_year(years, n_tracks, b, rank, world, device):
    """Fewer years than GPUs (BASELINE configs[4]: one WP year of 50 000 tracks on 8 GPUs): every rank runs EVERY year,
    integrating the seed attempts k with k % world == rank (Engine.set_shard; one exchange of kept flags and counted-seed
    histograms per wave over NCCL), so each rank holds the rows of the storms it integrated and zero bits elsewhere;
    the blocks merge on rank 0 by an integer sum of the float64 bit patterns (dist.reduce), then one device->host copy."""
    import torch
    import torch.distributed as dist
    from . import gather as tgather
    from .pipeline import _Block
    engine, bounds = _session.engine(b.basin_id, device.index)
    stream = torch.cuda.current_stream(device)
    engine.set_stream(stream.cuda_stream)
    engine.set_shard(rank, world, tgather.dist_allreduce(device))
    try:
        lon = lat = None
        planes = []
        for y in years:
            lon, lat, pl = _session.inputs.year_planes(_session.namelist, bounds, y)
            planes.append(pl)
        engine.alloc_tables(12 * len(years), lon, lat)
        for i, pl in enumerate(planes):
            engine.upload_months(12 * i, pl)
        blk = _Block(torch, device, len(years), n_tracks, engine.n_steps, pinned=False)
        stats = engine.run_years_dev(np.arange(len(years), dtype=np.int32) * 12, np.asarray(years, dtype=np.int32),
                                     _session.run_seed, n_tracks, blk.dptr)
    finally:
        engine.set_shard(0, 1)
    n_track_words = int(blk.offsets[6])                               # lon, lat, v, m, vmax, env: the sharded sections
    dist.reduce(blk.dev[:n_track_words].view(torch.int64), dst=0, op=dist.ReduceOp.SUM)
    if rank != 0:
        return None
    v = blk.views_of(blk.dev.cpu().numpy())
    out = {k: np.array(v[k]) for k in _FIELDS}
    out["stats"] = stats
    return out


def get_fn_tracks(b, namelist=None):
    """util/compute.py:40-47"""
    nl = namelist or _session.namelist
    return '%s/%s/tracks_%s_%s_%d%02d_%d%02d.nc' % (nl.output_directory, nl.exp_name, b.basin_id, nl.exp_prefix,
                                                    nl.start_year, nl.start_month, nl.end_year, nl.end_month)


def fn_tracks_duplicates(fn_trk):
    """util/compute.py:52-58 (incl. its rstrip('.nc') character-set behaviour)."""
    f_int = 0
    fn_trk_out = fn_trk
    while os.path.exists(fn_trk_out):
        fn_trk_out = fn_trk.rstrip('.nc') + '_e%d.nc' % f_int
        f_int += 1
    return fn_trk_out


def run_downscaling(basin_id, write=True, run_years_fn=None):
    """Runs the downscaling model in basin `basin_id` according to the namelist
    (util/compute.py:216-270): every year start_year..end_year, tracks_per_year tracks each.

    Under torch.distributed (one process per GPU) the years are sharded over the ranks; the finished tracks meet in
    rank 0's memory through run_downscaling_device (device-resident, one device->host copy), rank 0 writes the track
    file and returns the result, the other ranks return None.  `run_years_fn(years, n_tracks, b)` replaces the GPU call
    in host-logic tests (NumPy blocks, gloo all-gather, every rank returns the full result)."""
    nl = _session.namelist
    n_tracks = nl.tracks_per_year
    b = TC_Basin(basin_id, nl)
    years = list(range(nl.start_year, nl.end_year + 1))
    rank, world, device = 0, 1, None
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(), dist.get_world_size()
            if dist.get_backend() == "nccl":
                import torch
                device = torch.device("cuda", torch.cuda.current_device())
    except ImportError:
        pass
    n_steps = int(nl.total_track_time_days * 24 * 60 * 60 / nl.output_interval_s) + 1
    if run_years_fn is None:
        import torch
        if device is None:
            device = torch.device("cuda", _current_device())
        g = run_downscaling_device(years, n_tracks, b, rank, world, device)
        if g is None:                                              # not the writing rank: nothing to hold on the host
            return None
    else:                                                          # host-logic tests: stand-in for the GPU call, gloo
        mine = shard_years(years, rank, world)
        local = run_years_fn(mine, n_tracks, b) if mine else None
        g = gather_years(local, years, n_tracks, n_steps, rank, world, device)
    ny = len(years)
    out = dict(
        tc_lon=g["lon"].reshape(ny * n_tracks, n_steps), tc_lat=g["lat"].reshape(ny * n_tracks, n_steps),
        tc_v=g["v"].reshape(ny * n_tracks, n_steps), tc_m=g["m"].reshape(ny * n_tracks, n_steps),
        tc_vmax=g["vmax"].reshape(ny * n_tracks, n_steps), tc_env_wnds=g["env"].reshape(ny * n_tracks, n_steps, 4),
        tc_months=g["tc_month"].reshape(-1), tc_basins=_basin_labels(g["tc_basin"].reshape(-1)),
        tc_years=np.repeat(np.asarray(years, dtype=np.int64), n_tracks),
        n_seeds=g["n_seeds"], ts_output=np.linspace(0, nl.total_track_time_days * 24 * 60 * 60, n_steps),
        years=np.asarray(years, dtype=np.int64), basin_ids=list(layout.BASIN_IDS))
    if write and rank == 0:
        from . import trackfile
        os.makedirs('%s/%s' % (nl.base_directory, nl.exp_name), exist_ok=True)
        fn_trk_out = fn_tracks_duplicates(get_fn_tracks(b, nl))
        trackfile.write_tracks(fn_trk_out, out)
        out["fn_trk_out"] = fn_trk_out
        print('Saved %s' % fn_trk_out)
    return out
