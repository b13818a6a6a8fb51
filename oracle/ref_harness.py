"""Run the UNMODIFIED reference hot-path modules from /root/reference -- build container only.

TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box: nothing in the gpu tests,
smoke() or bench.py imports this module.  It is used by oracle/make_golden.py (committed
fixtures under tests/golden/) and by the optional CPU test that cross-checks the C oracle
against the live reference when the tree is present.

Recipe: xarray/dask/cftime are not installed, so they are stubbed in
sys.modules; Coupled_FAST is constructed with object.__new__ and the attributes its
__init__ chain would set (track/bam_track.py:51-69, intensity/coupled_fast.py:23-32), while
init_fields / gen_track / dydt / _env_winds / axi_to_max_wind are the reference's own code.
"""
import datetime
import os
import sys
import types
import contextlib

import numpy as np

REF_ROOT = os.environ.get("TCR_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "intensity", "coupled_fast.py"))


class Ref(types.SimpleNamespace):
    pass


_REF = None


def load_reference():
    global _REF
    if _REF is not None:
        return _REF
    for name in ("xarray", "dask", "cftime"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    saved = list(sys.path)
    sys.path.insert(0, REF_ROOT)
    try:
        shadow = {k: sys.modules.pop(k) for k in list(sys.modules)
                  if k in ("namelist", "util", "track", "intensity", "wind", "thermo")}
        import namelist                                       # noqa: E402
        from intensity import coupled_fast                    # noqa: E402
        from track import bam_track, env_wind                 # noqa: E402
        from util import basins, mat, sphere                  # noqa: E402
        from wind import tc_wind                              # noqa: E402
        from thermo import thermo as thermo_mod               # noqa: E402
        sys.modules.update(shadow)
    finally:
        sys.path[:] = saved
    bam_track.random_seed = lambda: None                      # kill wall-clock reseeding (bam_track.py:37)
    _REF = Ref(namelist=namelist, coupled_fast=coupled_fast, bam_track=bam_track, env_wind=env_wind,
               basins=basins, mat=mat, sphere=sphere, tc_wind=tc_wind, thermo=thermo_mod)
    return _REF


def build_fast(ref, basin_id, lon, lat, planes, static, h_bl=1400.0):
    """One month's Coupled_FAST from GLOBAL prepared planes [19][nlat][nlon] (float32 values) and
    the global static dict (synth.synth_static).  Cropping is done by the reference's own
    TC_Basin.transform_global_field."""
    nl = ref.namelist
    b = ref.basins.TC_Basin(basin_id)
    f = object.__new__(ref.coupled_fast.Coupled_FAST)
    f.dt_track = nl.output_interval_s
    f.total_time = nl.total_track_time_days * 24 * 60 * 60
    f.total_steps = int(f.total_time / f.dt_track) + 1
    f.t_s = np.linspace(0, f.total_time, f.total_steps)
    f.T_Fs = nl.T_days * 24 * 60 * 60
    f.u_beta, f.v_beta = nl.u_beta, nl.v_beta
    f.nLvl, f.nWLvl, f.basin = 2, 4, b
    f.var_names = ref.env_wind.wind_mean_vector_names()
    f.u_Mean_idxs = np.array([0, 2])
    f.v_Mean_idxs = np.array([1, 3])
    f.datetime_start = datetime.datetime(2000, 9, 15)
    f.dt_start = None
    f.Ck, f.h_bl = nl.Ck, h_bl
    f.epsilon, f.kappa = 0.33, 0.1
    f.beta = 1 - f.epsilon - f.kappa
    f.debug = False
    P = np.asarray(planes, dtype=np.float64)

    def fld(lon_g, lat_g, X):                                 # == _interp_basin_field, bam_track.py:72-74
        lon_b, lat_b, X_b = b.transform_global_field(lon_g, lat_g, X)
        return ref.mat.interp2_fx(lon_b, lat_b, np.nan_to_num(X_b))

    f.wnd_Mean_Fxs = [fld(lon, lat, P[i]) for i in range(4)]
    f.wnd_Cov_Fxs = [[fld(lon, lat, P[4 + i * (i + 1) // 2 + j]) if j <= i else "" for j in range(4)]
                     for i in range(4)]
    f.init_fields(lon, lat, P[14], P[15], P[16], P[17])       # genuine coupled_fast.py:217-225
    f.f_bath = fld(static["lon_b"], static["lat_b"], np.asarray(static["bathy"], dtype=np.float64))
    f.f_land = fld(static["lon_l"], static["lat_l"], np.asarray(static["land"], dtype=np.float64))
    f.m_init_fx = ref.mat.interp2_fx(lon, lat, P[18])         # util/compute.py:114
    return f


@contextlib.contextmanager
def injected_phases(phases):
    """gen_f draws np.random.rand(N, 1) once per series (bam_track.py:27): feed it our phases."""
    it = iter(np.asarray(phases, dtype=np.float64).reshape(4, 15))
    orig = np.random.rand

    def fake(*shape):
        return next(it).reshape(shape)

    np.random.rand = fake
    try:
        yield
    finally:
        np.random.rand = orig


def gen_track(ref, f, lon0, lat0, v0, m0, phases, h_bl, post=True):
    """Reference gen_track + the per-candidate post-processing of run_tracks
    (util/compute.py:178-206), computed for every storm that returned a result
    (post=False: the integration only)."""
    nl = ref.namelist
    f.h_bl = h_bl
    with injected_phases(phases):
        res = f.gen_track(lon0, lat0, v0, m0)
    if res is None:
        return dict(status=2, nfev=0, n_time=0)
    out = dict(status=int(res.status), nfev=int(res.nfev), n_time=int(res.t.size),
               t=res.t.copy(), y=res.y.copy())
    if not post:
        return out
    lon_t, lat_t, v_t = res.y[0], res.y[1], res.y[2]
    v_2d = np.interp(2 * 24 * 60 * 60, res.t, v_t.flatten())
    is_tc = bool(np.logical_and(np.any(v_t >= nl.seed_v_threshold_ms), v_2d >= nl.seed_v_2d_threshold_ms))
    env = np.array([f._env_winds(lon_t[i], lat_t[i], f.t_s[i]) for i in range(lon_t.size)])
    with np.errstate(all="ignore"):
        vmax = ref.tc_wind.axi_to_max_wind(lon_t, lat_t, f.dt_track, v_t, env)
        vmax = np.asarray(vmax, dtype=np.float64).reshape(-1)
        kept = bool(is_tc and (np.nanmax(vmax) >= nl.seed_vmax_threshold_ms)) if np.any(~np.isnan(vmax)) else False
    out.update(env=env, vmax=vmax, flags=(1 if is_tc else 0) | (2 if kept else 0))
    return out
