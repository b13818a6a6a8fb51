"""ctypes front end of the CPU oracle (oracle/liborc.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does (tests/test_no_oracle_in_product.py).

The ordered-selection driver at the bottom restates the reference's sequential
``while nt < n_tracks`` loop (util/compute.py:134-209) over an indexed, counter-based seed
stream.
"""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

N_FIELDS = 19
N_INTERP_OUT = 21
N_PHASES = 60

c_double_p = C.POINTER(C.c_double)


class OrcGrid(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("x", c_double_p), ("y", c_double_p)]


class OrcEnv(C.Structure):
    _fields_ = [("g", OrcGrid), ("f", c_double_p * N_FIELDS),
                ("gb", OrcGrid), ("bathy", c_double_p),
                ("gl", OrcGrid), ("land", c_double_p)]


def build():
    """Compile liborc.so if missing or stale (gcc only; seconds)."""
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in ("tcr_oracle.c", "preproc_oracle.c")]
    hdrs = [os.path.join(_HERE, "..", "include", h) for h in ("tcrisk.h", "tcr_libm.h")]
    newest = max(os.path.getmtime(f) for f in srcs + hdrs)
    if (not os.path.exists(so)) or os.path.getmtime(so) < newest:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_bilinear.restype = C.c_double
        _LIB.orc_postprocess.restype = C.c_uint32
        if not _LIB.orc_cpu_has_fma():
            raise RuntimeError("liborc.so is built with -mfma but this CPU has no FMA unit")
    return _LIB


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _grid(lon, lat):
    g = OrcGrid()
    g.nx, g.ny, g.x, g.y = lon.size, lat.size, _dp(lon), _dp(lat)
    return g


class OracleEnv:
    """float64 copies of the prepared tables in the oracle's struct layout.

    planes: float32/float64 [n_ym][19][nlat][nlon]; st: dict from fields.prepare_static."""

    def __init__(self, lon, lat, planes, st):
        self.lon = np.ascontiguousarray(lon, dtype=np.float64)
        self.lat = np.ascontiguousarray(lat, dtype=np.float64)
        self.planes = np.ascontiguousarray(planes, dtype=np.float64)
        self.n_ym = self.planes.shape[0]
        self.lon_b = np.ascontiguousarray(st["lon_b"], dtype=np.float64)
        self.lat_b = np.ascontiguousarray(st["lat_b"], dtype=np.float64)
        self.bathy = np.ascontiguousarray(st["bathy"], dtype=np.float64)
        self.lon_l = np.ascontiguousarray(st["lon_l"], dtype=np.float64)
        self.lat_l = np.ascontiguousarray(st["lat_l"], dtype=np.float64)
        self.land = np.ascontiguousarray(st["land"], dtype=np.float64)
        self.envs = (OrcEnv * self.n_ym)()
        for i in range(self.n_ym):
            e = self.envs[i]
            e.g = _grid(self.lon, self.lat)
            for c in range(N_FIELDS):
                e.f[c] = _dp(self.planes[i, c])
            e.gb = _grid(self.lon_b, self.lat_b)
            e.bathy = _dp(self.bathy)
            e.gl = _grid(self.lon_l, self.lat_l)
            e.land = _dp(self.land)

    def env_ptr(self, ym0=0):
        return C.cast(C.byref(self.envs, ym0 * C.sizeof(OrcEnv)), C.POINTER(OrcEnv))


def time_axis(p):
    t = np.empty(p.n_steps)
    lib().orc_time_axis(C.byref(p), _dp(t))
    return t


def env_interp(env, ym, lon, lat):
    ym = np.ascontiguousarray(ym, dtype=np.int32)
    lon = np.ascontiguousarray(lon, dtype=np.float64)
    lat = np.ascontiguousarray(lat, dtype=np.float64)
    out = np.empty((ym.size, N_INTERP_OUT))
    lib().orc_env_interp(env.env_ptr(), C.c_int64(ym.size), ym.ctypes.data_as(C.POINTER(C.c_int32)),
                         _dp(lon), _dp(lat), _dp(out))
    return out


def gen_f(p, phases):
    """The storm's Fourier table self.Fs [4][n_steps] in the oracle's (spec) arithmetic."""
    phases = np.ascontiguousarray(phases, dtype=np.float64).reshape(-1)
    fs = np.empty((4, p.n_steps))
    lib().orc_gen_f(C.byref(p), _dp(phases), _dp(fs))
    return fs


def gen_f_direct(phases, T, t_s):
    """gen_f exactly as track/bam_track.py:23-31 writes it (direct sin per harmonic, glibc)."""
    phases = np.ascontiguousarray(phases, dtype=np.float64).reshape(-1)
    t_s = np.ascontiguousarray(t_s, dtype=np.float64)
    fs = np.empty((4, t_s.size))
    lib().orc_gen_f_direct(_dp(phases), C.c_double(T), _dp(t_s), C.c_int(t_s.size), _dp(fs))
    return fs


def fourier_coef(p, phases):
    phases = np.ascontiguousarray(phases, dtype=np.float64).reshape(-1)
    coef = np.empty(N_PHASES * 2)
    lib().orc_fourier_coef(C.byref(p), _dp(phases), _dp(coef))
    return coef


def dydt_at(p, env, ym, phases, h_bl, t, y):
    coef = fourier_coef(p, phases)
    y = np.ascontiguousarray(y, dtype=np.float64)
    dy = np.empty(4)
    lib().orc_dydt_at(C.byref(p), env.env_ptr(ym), _dp(coef), C.c_double(h_bl), C.c_double(t), _dp(y), _dp(dy))
    return dy


def libm_eval(fn, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    lib().orc_libm_eval(C.c_int(fn), C.c_int64(x.size), _dp(x), _dp(y))
    return y


def _chunks(n, n_threads):
    n_chunks = max(1, min(n, n_threads * 4)) if n_threads > 1 else 1
    edges = np.linspace(0, n, n_chunks + 1).astype(np.int64)
    return [(int(a), int(b)) for a, b in zip(edges[:-1], edges[1:]) if b > a]


def integrate_batch(p, env, ym, lon0, lat0, v0, m0, h_bl, phases, post_all=True, n_threads=1):
    """gen_track + post-processing for explicit seeds; dict of arrays shaped like tcr_integrate."""
    n = len(lon0)
    ns = p.n_steps
    a = lambda x, dt=np.float64: np.ascontiguousarray(x, dtype=dt)
    ym, lon0, lat0, v0, m0, h_bl = a(ym, np.int32), a(lon0), a(lat0), a(v0), a(m0), a(h_bl)
    phases = a(phases).reshape(n, N_PHASES)
    out = dict(track=np.empty((n, ns, 4)), env=np.empty((n, ns, 4)), vmax=np.empty((n, ns)),
               n_time=np.zeros(n, np.int32), status=np.zeros(n, np.int32),
               nfev=np.zeros(n, np.int32), flags=np.zeros(n, np.uint32), n_clean=np.zeros(n, np.int32))
    L = lib()
    ip = C.POINTER(C.c_int32)

    def run(ab):
        lo, hi = ab
        L.orc_integrate_batch(
            C.byref(p), env.env_ptr(), C.c_int64(hi - lo),
            ym[lo:hi].ctypes.data_as(ip), _dp(lon0[lo:hi]), _dp(lat0[lo:hi]), _dp(v0[lo:hi]),
            _dp(m0[lo:hi]), _dp(h_bl[lo:hi]), _dp(phases[lo:hi]),
            _dp(out["track"][lo:hi]), _dp(out["env"][lo:hi]), _dp(out["vmax"][lo:hi]),
            out["n_time"][lo:hi].ctypes.data_as(ip), out["status"][lo:hi].ctypes.data_as(ip),
            out["nfev"][lo:hi].ctypes.data_as(ip),
            out["flags"][lo:hi].ctypes.data_as(C.POINTER(C.c_uint32)),
            out["n_clean"][lo:hi].ctypes.data_as(ip), C.c_int(1 if post_all else 0))

    if n_threads > 1:
        with ThreadPoolExecutor(n_threads) as ex:
            list(ex.map(run, _chunks(n, n_threads)))
    else:
        run((0, n))
    return out


class Masks:
    def __init__(self, lon_m, lat_m, planes_u8):
        self.lon = np.ascontiguousarray(lon_m, dtype=np.float64)
        self.lat = np.ascontiguousarray(lat_m, dtype=np.float64)
        self.planes = np.ascontiguousarray(planes_u8, dtype=np.float64)
        self.grid = _grid(self.lon, self.lat)


def run_attempts(p, env, ym_base, masks, run_seed, year_key, k0, n, want_tracks=True, n_threads=1):
    """Seed attempts [k0, k0+n) of one year: seeding, integration, post-processing."""
    ns = p.n_steps
    out = dict(code=np.zeros(n, np.int32), basin=np.zeros(n, np.int32), month=np.zeros(n, np.int32),
               n_time=np.zeros(n, np.int32), status=np.zeros(n, np.int32), nfev=np.zeros(n, np.int32),
               flags=np.zeros(n, np.uint32), n_clean=np.zeros(n, np.int32), ic=np.zeros((n, 4)))
    if want_tracks:
        out.update(track=np.empty((n, ns, 4)), env=np.empty((n, ns, 4)), vmax=np.empty((n, ns)))
    L = lib()
    ip = C.POINTER(C.c_int32)
    null = C.cast(None, c_double_p)

    def run(ab):
        lo, hi = ab
        L.orc_run_attempts(
            C.byref(p), env.env_ptr(ym_base), C.byref(masks.grid), _dp(masks.planes),
            C.c_uint32(run_seed), C.c_int32(year_key), C.c_int64(k0 + lo), C.c_int64(hi - lo),
            out["code"][lo:hi].ctypes.data_as(ip), out["basin"][lo:hi].ctypes.data_as(ip),
            out["month"][lo:hi].ctypes.data_as(ip), out["n_time"][lo:hi].ctypes.data_as(ip),
            out["status"][lo:hi].ctypes.data_as(ip), out["nfev"][lo:hi].ctypes.data_as(ip),
            out["flags"][lo:hi].ctypes.data_as(C.POINTER(C.c_uint32)),
            out["n_clean"][lo:hi].ctypes.data_as(ip), _dp(out["ic"][lo:hi]),
            _dp(out["track"][lo:hi]) if want_tracks else null,
            _dp(out["env"][lo:hi]) if want_tracks else null,
            _dp(out["vmax"][lo:hi]) if want_tracks else null)

    if n_threads > 1:
        with ThreadPoolExecutor(n_threads) as ex:
            list(ex.map(run, _chunks(n, n_threads)))
    else:
        run((0, n))
    return out


def poi_vmax(lon, lat, vmax, poi_lon, poi_lat, radius_km=100.0, r_earth_m=6378000.0):
    """notebooks/sample_analysis.ipynb cell 15: per-track maximum of vmax within radius_km of a point."""
    lon, lat, vmax = (np.ascontiguousarray(x, dtype=np.float64) for x in (lon, lat, vmax))
    n_steps = lon.shape[-1]
    out = np.empty(lon.size // n_steps)
    lib().orc_poi_vmax(C.c_int64(out.size), C.c_int(n_steps), _dp(lon), _dp(lat), _dp(vmax), C.c_double(poi_lon),
                       C.c_double(poi_lat), C.c_double(radius_km), C.c_double(r_earth_m), _dp(out))
    return out.reshape(lon.shape[:-1])


def exceedance(v, bins):
    """cell 17: number of tracks whose vmax_at_poi reaches each bin."""
    v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1)
    bins = np.ascontiguousarray(bins, dtype=np.float64)
    counts = np.zeros(bins.size, np.int64)
    lib().orc_exceedance(C.c_int64(v.size), _dp(v), C.c_int(bins.size), _dp(bins), counts.ctypes.data_as(C.POINTER(C.c_int64)))
    return counts


def phases_for(run_seed, year_key, k):
    ph = np.empty(N_PHASES)
    lib().orc_phases(C.c_uint32(run_seed), C.c_int32(year_key), C.c_int64(k), _dp(ph))
    return ph


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32(c, k, o)
    return list(o)


def run_year(p, env, ym_base, masks, run_seed, year_key, n_tracks, chunk=4096, n_threads=1,
             max_attempts=50_000_000):
    """Oracle of run_tracks(year, n_tracks, b) (util/compute.py:64-210) on the indexed seed
    stream: walk attempts 0,1,2,... in order, keep the first n_tracks storms that pass the TC
    and vmax criteria; n_seeds counts every counted attempt up to and including the attempt
    that produced the last kept storm (the reference increments before integrating, :167)."""
    ns = p.n_steps
    res = dict(lon=np.full((n_tracks, ns), np.nan), lat=np.full((n_tracks, ns), np.nan),
               v=np.full((n_tracks, ns), np.nan), m=np.full((n_tracks, ns), np.nan),
               vmax=np.full((n_tracks, ns), np.nan), env=np.full((n_tracks, ns, 4), np.nan),
               tc_month=np.full(n_tracks, np.nan), tc_basin=np.full(n_tracks, -1, np.int32),
               n_seeds=np.zeros((7, 12)), attempt=np.full(n_tracks, -1, np.int64))
    stats = dict(attempts=0, counted_seeds=0, integrated=0, storm_steps=0, kept_steps=0, rhs_evals=0)
    nt, k0 = 0, 0
    while nt < n_tracks and k0 < max_attempts:
        o = run_attempts(p, env, ym_base, masks, run_seed, year_key, k0, chunk, True, n_threads)
        kept = np.nonzero((o["flags"] & 2) != 0)[0]
        if nt + kept.size >= n_tracks:
            last = int(kept[n_tracks - nt - 1])
            kept = kept[: n_tracks - nt]
        else:
            last = chunk - 1
        sl = slice(0, last + 1)
        counted = o["code"][sl] >= 1
        counted &= o["code"][sl] <= 2
        np.add.at(res["n_seeds"], (o["basin"][sl][counted], o["month"][sl][counted] - 1), 1)
        integ = o["code"][sl] == 2
        stats["attempts"] = k0 + last + 1
        stats["counted_seeds"] += int(counted.sum())
        stats["integrated"] += int(integ.sum())
        stats["storm_steps"] += int(o["n_time"][sl][integ].sum())
        stats["rhs_evals"] += int(o["nfev"][sl][integ].sum())
        for q in kept:
            n_time = int(o["n_time"][q])
            res["lon"][nt, :n_time] = o["track"][q, :n_time, 0]
            res["lat"][nt, :n_time] = o["track"][q, :n_time, 1]
            res["v"][nt, :n_time] = o["track"][q, :n_time, 2]
            res["m"][nt, :n_time] = o["track"][q, :n_time, 3]
            res["vmax"][nt, :n_time] = o["vmax"][q, :n_time]
            res["env"][nt, :n_time] = o["env"][q, :n_time]
            res["tc_month"][nt] = o["month"][q]
            res["tc_basin"][nt] = o["basin"][q]
            res["attempt"][nt] = k0 + q
            stats["kept_steps"] += n_time
            nt += 1
        k0 += chunk
    res["stats"] = stats
    res["n_kept"] = nt
    return res
