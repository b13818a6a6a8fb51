"""The INNER tier of the reference's seam on this stack: an object with the interface of the reference's
``Coupled_FAST`` (intensity/coupled_fast.py:18-267; base class track/bam_track.py:44-150) whose work runs on the GPU
through libtcrisk.so, so that the reference's own ``run_tracks`` loop (util/compute.py:123-209) can drive it unchanged:

    fast = Coupled_FAST(fn_wnd_stat, basin, dt_start, dt_s, total_time_s)       coupled_fast.py:19
    fast.init_fields(lon, lat, chi, vpot, mld, strat)                           coupled_fast.py:217
    fast.h_bl = ...                                                             util/compute.py:175
    res = fast.gen_track(clon, clat, v, m)      -> None | result with .t .y .status .nfev      coupled_fast.py:229
    fast._env_winds(lon, lat, t) -> [4]                                         track/bam_track.py:116
    fast.dydt(t, y) -> [4]                                                      coupled_fast.py:196
    fast.f_vpot.ev(lon, lat), fast.t_s, fast.dt_track, fast.total_steps, fast.nWLvl

One object = one month of fields = one table slot of its own Engine.  The outer tier (`compute.run_tracks`, all storms
of a year in one launch sequence) is the fast path; this tier costs one launch sequence per call and exists so that code
written against the reference's class keeps working.  `gen_tracks` is the batched form of `gen_track`.

What the reference reads from files inside its constructor is passed in instead (this image has no xarray):
  * the 14 monthly wind statistics of `fn_wnd_stat` at `dt_start`: a path to an env_wnd_*.nc cache (read with
    refdata), or a mapping {'lon', 'lat', <layout.FIELD_NAMES[:14]>: [lat, lon]} on the global grid;
  * bathymetry / land (intensity/geo.py:9-33): `static=` a dict as fields.prepare_static consumes, default the
    session's input provider (compute.configure).
Random phases come from ``np.random.rand`` exactly like gen_f (bam_track.py:27), after ``random_seed()`` (bam_track.py:37-42).
"""
import time

import numpy as np

from . import fields, layout, params
from . import namelist as default_namelist


def random_seed():
    """track/bam_track.py:37-42: reseed numpy's global generator from the wall clock."""
    t = int(time.time() * 1000.0)
    np.random.seed(((t & 0xff000000) >> 24) + ((t & 0x00ff0000) >> 8) + ((t & 0x0000ff00) << 8) + ((t & 0x000000ff) << 24))


class OdeResult:
    """The fields of scipy's OdeResult that callers of gen_track read (util/compute.py:178-203)."""

    def __init__(self, t, y, status, nfev):
        self.t, self.y, self.status, self.nfev = t, y, int(status), int(nfev)
        self.success = self.status >= 0
        self.message = {0: "The solver successfully reached the end of the integration interval.",
                        1: "A termination event occurred.", -1: "Required step size is less than spacing between numbers."}[self.status]
        self.t_events = None


class _Field:
    """RectBivariateSpline(kx=1, ky=1)-like sampler of one channel of the month table (`f_vpot.ev`, util/compute.py:162)."""

    def __init__(self, owner, channel):
        self._owner, self._ch = owner, channel

    def ev(self, lon, lat):
        lon, lat = np.asarray(lon, dtype=np.float64), np.asarray(lat, dtype=np.float64)
        out = self._owner._engine.env_interp(np.zeros(lon.size, np.int32), lon.reshape(-1), lat.reshape(-1))[:, self._ch]
        return out.reshape(lon.shape) if lon.shape else out[0]


class Coupled_FAST:
    def __init__(self, fn_wnd_stat, basin, dt_start, dt_s, total_time_s, static=None, namelist=None, device=None):
        nl = namelist or default_namelist
        self.namelist = nl
        self.basin = basin                                    # a TC_Basin (compute.TC_Basin or the reference's)
        self.dt_start = dt_start
        self.datetime_start = dt_start
        self.dt_track = dt_s                                  # bam_track.py:52-55
        self.total_time = total_time_s
        self.total_steps = int(self.total_time / self.dt_track) + 1
        self.t_s = np.linspace(0, self.total_time, self.total_steps)
        self.T_Fs = nl.T_days * 24 * 60 * 60
        self.u_beta, self.v_beta = nl.u_beta, nl.v_beta
        self.nLvl = len(nl.steering_levels)
        self.nWLvl = self.nLvl * 2
        self.Ck = nl.Ck
        self.h_bl = 1400.0                                    # coupled_fast.py:24; run_tracks overwrites it per storm
        self.epsilon, self.kappa = 0.33, 0.1
        self.beta = 1 - self.epsilon - self.kappa
        self.debug = False
        self._wnd = self._load_wnd_stat(fn_wnd_stat, dt_start)
        self._static = static
        self._device = device
        self._engine = None
        self.Fs_phases = None                                 # phases of the last gen_track (the reference keeps self.Fs)

    # -- what the reference reads from files ------------------------------------------------------
    @staticmethod
    def _load_wnd_stat(src, dt_start):
        """bam_track.py:76-91: the 14 wind statistics, interpolated in time to dt_start."""
        if isinstance(src, (str, bytes)):
            from . import refdata
            c = refdata._Cache(src, layout.FIELD_NAMES[:14])
            planes = [c.at(name, dt_start) for name in layout.FIELD_NAMES[:14]]
            return dict(lon=c.lon, lat=c.lat, planes=np.stack(planes))
        return dict(lon=np.asarray(src["lon"], dtype=np.float64), lat=np.asarray(src["lat"], dtype=np.float64),
                    planes=np.stack([np.asarray(src[name], dtype=np.float64) for name in layout.FIELD_NAMES[:14]]))

    def _bounds(self):
        return tuple(float(x) for x in self.basin.get_bounds())

    # -- Coupled_FAST.init_fields (coupled_fast.py:217-225) -----------------------------------------
    def init_fields(self, lon, lat, chi, vpot, mld, strat):
        """Global [lat, lon] fields of the month, already prepared by the caller (util/compute.py:107-121): cropped to
        the basin here like the reference does (transform_global_field), stored as float32 table records."""
        import torch
        from .engine import Engine
        lon, lat = np.asarray(lon, dtype=np.float64), np.asarray(lat, dtype=np.float64)
        w = self._wnd
        if w["lon"].shape != lon.shape or not (np.array_equal(w["lon"], lon) and np.array_equal(w["lat"], lat)):
            raise ValueError("wind statistics and thermodynamic fields must share one global grid")
        stack = np.concatenate([np.nan_to_num(w["planes"]),                       # bam_track.py:74
                                np.stack([np.asarray(f, dtype=np.float64) for f in (chi, vpot, mld, strat)]),
                                np.zeros((1,) + w["planes"].shape[1:])])           # rh_mid: sampled by the caller, not here
        bounds = self._bounds()
        lon_b, lat_b, planes = fields.crop_to_basin(lon, lat, stack, bounds)
        if self._engine is None:
            basin_id = getattr(self.basin, "basin_id", "GL")
            p = params.params_from_namelist(_with_interval(self.namelist, self.dt_track, self.total_time), basin_id)
            if tuple(p.basin_bounds) != bounds:
                for i in range(4):
                    p.basin_bounds[i] = bounds[i]
            dev = self._device if self._device is not None else torch.cuda.current_device()
            self._engine = Engine(p, device=dev)
            st = self._static
            if st is None:
                from . import compute
                st = compute._session.inputs.static()
            self._engine.upload_static(fields.prepare_static(bounds, st))
        self._engine.alloc_tables(1, lon_b, lat_b)
        self._engine.upload_months(0, np.ascontiguousarray(planes[None], dtype=np.float32))
        self._engine.synchronize()
        self.f_chi, self.f_vpot = _Field(self, layout.CH_CHI), _Field(self, layout.CH_VPOT)
        self.f_mld, self.f_strat = _Field(self, layout.CH_MLD), _Field(self, layout.CH_STRAT)
        self.f_bath, self.f_land = _Field(self, layout.OUT_BATHY), _Field(self, layout.OUT_LAND)

    # -- gen_f phases (bam_track.py:23-31, 111-113) ------------------------------------------------
    def _draw_phases(self):
        return np.stack([np.random.rand(15, 1) for _ in range(self.nWLvl)]).reshape(60)

    # -- Coupled_FAST.gen_track (coupled_fast.py:229-267) -----------------------------------------
    def gen_track(self, clon, clat, v, m=None):
        if m is None:
            raise NotImplementedError("gen_track without an initial m (coupled_fast.py:153-171) is dead code for run.py")
        random_seed()                                                               # :231
        self.Fs_phases = self._draw_phases()                                        # :234
        r = self._engine.integrate([0], [clon], [clat], [v], [m], [self.h_bl], self.Fs_phases[None])
        return self._result(r, 0)

    def gen_tracks(self, clon, clat, v, m, h_bl=None, phases=None):
        """Batched gen_track: arrays of seeds -> list of results (None where the reference returns None)."""
        n = len(clon)
        if phases is None:
            random_seed()
            phases = np.stack([self._draw_phases() for _ in range(n)])
        hbl = np.full(n, self.h_bl) if h_bl is None else np.asarray(h_bl, dtype=np.float64)
        r = self._engine.integrate(np.zeros(n, np.int32), clon, clat, v, m, hbl, phases)
        return [self._result(r, i) for i in range(n)]

    def _result(self, r, i):
        if r["status"][i] == 2:                                                     # ventilation pre-check, :238-244
            return None
        k = int(r["n_time"][i])
        return OdeResult(self.t_s[:k].copy(), np.ascontiguousarray(r["track"][i, :k].T), r["status"][i], r["nfev"][i])

    # -- BetaAdvectionTrack._env_winds (bam_track.py:116-128), Coupled_FAST.dydt (coupled_fast.py:196-207) ----
    def _eval(self, t, y):
        if self.Fs_phases is None:
            raise RuntimeError("no Fourier series yet: call gen_track first (the reference sets self.Fs there)")
        return self._engine.rhs_eval([0], [t], np.asarray(y, dtype=np.float64)[None], [self.h_bl], self.Fs_phases[None])

    def _env_winds(self, clon, clat, ts):
        return self._eval(ts, [clon, clat, 0.0, 0.0])[1][0]

    def dydt(self, t, y):
        return self._eval(t, y)[0][0]

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None


def _with_interval(nl, dt_s, total_time_s):
    """The namelist with the constructor's output interval / track length (the reference passes them as arguments)."""
    import types
    cfg = types.SimpleNamespace(**{k: getattr(nl, k) for k in dir(nl) if not k.startswith("__")})
    cfg.output_interval_s = dt_s
    cfg.total_track_time_days = total_time_s / 86400.0
    return cfg
