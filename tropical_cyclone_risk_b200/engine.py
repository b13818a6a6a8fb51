"""Host-side owner of one ``tcr_handle`` (one per GPU): uploads the prepared tables and calls the
hot-path entry points of libtcrisk.so.  All arithmetic happens in the CUDA library; this class
only marshals NumPy arrays (or raw device pointers) across the C ABI of include/tcrisk.h.
"""
import ctypes as C
import weakref

import numpy as np

from . import _lib, layout
from .params import TcrParams, TcrYearStats

N_OUT = layout.N_INTERP_OUT


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def _arr(x, dtype):
    return np.ascontiguousarray(x, dtype=dtype)


class PinnedPool:
    """NumPy views of page-locked host memory (tcr_host_alloc)."""

    @staticmethod
    def empty(shape, dtype):
        lib = _lib.load()
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        _lib.check(lib.tcr_host_alloc(max(n, 1), C.byref(p)))
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        a = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        weakref.finalize(buf, lib.tcr_host_free, p)
        return a


class Engine:
    def __init__(self, params, device=0):
        if not isinstance(params, TcrParams):
            raise TypeError("params must be a TcrParams (params.params_from_namelist)")
        self.lib = _lib.load()
        self.p = params
        self.n_steps = int(params.n_steps)
        self._h = C.c_void_p()
        _lib.check(self.lib.tcr_create(int(device), C.byref(params), C.byref(self._h)))
        self.device = int(device)
        self.n_ym = 0
        self.grid = None

    # -- life cycle ---------------------------------------------------------------------------
    def close(self):
        if self._h:
            self.lib.tcr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_stream(self, cuda_stream_ptr):
        _lib.check(self.lib.tcr_set_stream(self._h, C.c_void_p(int(cuda_stream_ptr))))

    def synchronize(self):
        _lib.check(self.lib.tcr_synchronize(self._h))

    def set_tuning(self, integ_variant=0, max_wave=0, max_slots=0, oversub_permille=0):
        """0 keeps a knob: integrate-kernel register variant (1: 256 thr x 1 CTA/SM, 2: 128 x 3,
        3: 128 x 4, 4: 160 x 2), seed attempts / integrated storms per wave, wave over-subscription
        (x1000).  Results never depend on these."""
        _lib.check(self.lib.tcr_set_tuning(self._h, integ_variant, max_wave, max_slots, oversub_permille))

    @property
    def fourier_ring_nodes(self):
        """0: run_years tabulates full Fourier tables ahead of the integrator; n: rings of n nodes filled inside it."""
        n = self.lib.tcr_fourier_ring_nodes(self._h)
        if n < 0:
            _lib.check(n)
        return int(n)

    def set_workspace_budget(self, fraction_of_free=0.6, cap_bytes=96 << 30):
        """Memory the per-wave workspace of run_years may take (tcr_set_workspace_budget): fewer, larger waves with more."""
        _lib.check(self.lib.tcr_set_workspace_budget(self._h, float(fraction_of_free), int(cap_bytes)))

    def set_shard(self, rank, world, allreduce=None):
        """Within-year sharding (tcr_set_shard): run_years becomes collective over `world` engines; this one integrates
        the attempts k with k % world == rank.  allreduce(ptr, count, dtype, op, stream) -> None performs the in-place
        all-reduce of a device buffer (dtype 0 u8 / 1 i32 / 2 i64, op 0 sum / 1 min); see gather.dist_allreduce."""
        if allreduce is None:
            cb = _lib.ALLREDUCE_FN()
        else:
            def tramp(user, ptr, count, dtype, op, stream):
                try:
                    allreduce(int(ptr), int(count), int(dtype), int(op), int(stream or 0))
                    return 0
                except Exception:                                  # never let an exception cross the C boundary
                    import traceback
                    traceback.print_exc()
                    return -1
            cb = _lib.ALLREDUCE_FN(tramp)
        self._allreduce_cb = cb                                    # keep the trampoline alive as long as the handle uses it
        _lib.check(self.lib.tcr_set_shard(self._h, int(rank), int(world), cb, None))

    def set_interp_variant(self, variant):
        _lib.check(self.lib.tcr_set_interp_variant(self._h, int(variant)))

    @property
    def launch_count(self):
        return int(self.lib.tcr_launch_count(self._h))

    KERNEL_CLASSES = ("env_interp", "integrate", "postprocess", "seed", "coef", "select", "gather", "build", "ftable",
                      "poi", "windstat", "thermo")

    def set_timing(self, enable=True):
        """CUDA-event accounting of every kernel class on the handle's stream (resets the totals)."""
        _lib.check(self.lib.tcr_set_timing(self._h, 1 if enable else 0))

    def kernel_times(self):
        """{class: (milliseconds, launches)} accumulated since set_timing()."""
        out = {}
        for i, name in enumerate(self.KERNEL_CLASSES):
            ms, n = C.c_double(), C.c_int64()
            _lib.check(self.lib.tcr_kernel_time(self._h, i, C.byref(ms), C.byref(n)))
            out[name] = (ms.value, n.value)
        return out

    # -- uploads ------------------------------------------------------------------------------
    def upload_static(self, st):
        """st: dict from fields.prepare_static (basin-cropped bathymetry / land + axes)."""
        lat_b, lon_b = _arr(st["lat_b"], np.float64), _arr(st["lon_b"], np.float64)
        lat_l, lon_l = _arr(st["lat_l"], np.float64), _arr(st["lon_l"], np.float64)
        bathy, land = _arr(st["bathy"], np.int16), _arr(st["land"], np.int8)
        assert bathy.shape == (lat_b.size, lon_b.size) and land.shape == (lat_l.size, lon_l.size)
        _lib.check(self.lib.tcr_upload_static(self._h, lat_b.size, lon_b.size, _ptr(lat_b), _ptr(lon_b), _ptr(bathy),
                                              lat_l.size, lon_l.size, _ptr(lat_l), _ptr(lon_l), _ptr(land)))

    def upload_masks(self, lon_m, lat_m, planes_u8):
        lon_m, lat_m = _arr(lon_m, np.float64), _arr(lat_m, np.float64)
        m = _arr(planes_u8, np.uint8)
        assert m.shape == (8, lat_m.size, lon_m.size)
        _lib.check(self.lib.tcr_upload_masks(self._h, lat_m.size, lon_m.size, _ptr(lat_m), _ptr(lon_m), _ptr(m)))

    def alloc_tables(self, n_ym, lon, lat):
        lon, lat = _arr(lon, np.float64), _arr(lat, np.float64)
        _lib.check(self.lib.tcr_alloc_tables(self._h, int(n_ym), lat.size, lon.size, _ptr(lat), _ptr(lon)))
        self.n_ym = int(n_ym)
        self.grid = (lat.size, lon.size)

    def upload_month(self, ym, planes):
        """planes: float32 [19][nlat][nlon] (layout.FIELD_NAMES order), host memory."""
        planes = _arr(planes, np.float32)
        assert planes.shape == (layout.N_FIELDS,) + self.grid, (planes.shape, self.grid)
        ptrs = (C.c_void_p * layout.N_FIELDS)(*[planes[i].ctypes.data for i in range(layout.N_FIELDS)])
        _lib.check(self.lib.tcr_upload_month(self._h, int(ym), ptrs))

    def upload_months(self, ym0, planes):
        """planes: float32 [n][19][nlat][nlon], C-contiguous host memory (pinned for full PCIe speed):
        one copy + one launch for all n months."""
        if not (isinstance(planes, np.ndarray) and planes.dtype == np.float32 and planes.flags.c_contiguous):
            planes = _arr(planes, np.float32)
        assert planes.shape[1:] == (layout.N_FIELDS,) + self.grid, (planes.shape, self.grid)
        _lib.check(self.lib.tcr_upload_months(self._h, int(ym0), int(planes.shape[0]), _ptr(planes)))

    def prepare_month(self, ym, namelist, bounds, lon, lat, raw, ocean_lon, ocean_lat, mld, strat, return_planes=False):
        """Device-side fields.prepare_month (util/compute.py:107-121 + bam_track.py:72-74) for month slot `ym`
        (ym < 0: prepare only).  raw: dict of the 14 wind statistics + 'vmax', 'chi', 'rh_mid' on the global
        grid (lon, lat in storage order); mld / strat on the ascending ocean grid.  Returns the prepared planes
        [19][nlat_b][nlon_b] float32 when return_planes."""
        from . import fields
        from .params import TcrPrepSpec
        lon, lat = _arr(lon, np.float64), _arr(lat, np.float64)
        olon, olat = _arr(ocean_lon, np.float64), _arr(ocean_lat, np.float64)
        lon_b, lat_b, src_col, src_row = fields.crop_index_maps(lon, lat, bounds)
        names = list(layout.FIELD_NAMES[:14]) + ["vmax", "chi", "rh_mid"]
        stack = np.ascontiguousarray(np.stack([np.asarray(raw[n], dtype=np.float32) for n in names]))
        ocean = np.ascontiguousarray(np.stack([np.asarray(mld, dtype=np.float32), np.asarray(strat, dtype=np.float32)]))
        sp = TcrPrepSpec()
        sp.nlat_g, sp.nlon_g = lat.size, lon.size
        sp.nlat_o, sp.nlon_o = olat.size, olon.size
        sp.nlat_b, sp.nlon_b = lat_b.size, lon_b.size
        sp.pi_reduc = float(namelist.PI_reduc)
        sp.sqrt_ck_cd = float(np.sqrt(namelist.Ck / namelist.Cd))
        sp.log_chi_fac, sp.chi_fac = float(namelist.log_chi_fac), float(namelist.chi_fac)
        out = np.empty((layout.N_FIELDS, lat_b.size, lon_b.size), np.float32) if return_planes else None
        _lib.check(self.lib.tcr_prepare_month(
            self._h, int(ym), C.byref(sp), _ptr(stack), _ptr(ocean), _ptr(lon), _ptr(lat), _ptr(olon), _ptr(olat),
            _ptr(src_col), _ptr(src_row), _ptr(out) if return_planes else C.c_void_p()))
        return (lon_b, lat_b, out) if return_planes else (lon_b, lat_b)

    def upload_month_dev(self, ym, d_planes_ptr):
        _lib.check(self.lib.tcr_upload_month_dev(self._h, int(ym), C.c_void_p(int(d_planes_ptr))))

    def upload_case(self, lon, lat, planes, static, mask_lon=None, mask_lat=None, mask_planes=None):
        """Everything one basin needs: planes [n_ym][19][nlat][nlon]."""
        self.upload_static(static)
        if mask_planes is not None:
            self.upload_masks(mask_lon, mask_lat, mask_planes)
        self.alloc_tables(planes.shape[0], lon, lat)
        self.upload_months(0, planes)
        self.synchronize()

    # -- hot path -----------------------------------------------------------------------------
    def env_interp(self, ym, lon, lat):
        """RectBivariateSpline(kx=1,ky=1).ev of all 19 monthly fields + bathymetry + land."""
        ym, lon, lat = _arr(ym, np.int32), _arr(lon, np.float64), _arr(lat, np.float64)
        out = np.empty((ym.size, N_OUT))
        _lib.check(self.lib.tcr_env_interp(self._h, ym.size, _ptr(ym), _ptr(lon), _ptr(lat), _ptr(out), 0))
        return out

    def env_interp_dev(self, n, d_ym, d_lon, d_lat, d_out):
        """Device pointers (ints); asynchronous on the handle's stream."""
        vp = C.c_void_p
        _lib.check(self.lib.tcr_env_interp(self._h, int(n), vp(d_ym), vp(d_lon), vp(d_lat), vp(d_out), 1))

    def integrate(self, ym, lon0, lat0, v0, m0, h_bl, phases):
        """Coupled_FAST.gen_track + post-processing for explicit seeds."""
        n = len(lon0)
        ns = self.n_steps
        ym = _arr(ym, np.int32)
        lon0, lat0, v0, m0, h_bl = (_arr(x, np.float64) for x in (lon0, lat0, v0, m0, h_bl))
        phases = _arr(phases, np.float64).reshape(n, 60)
        out = dict(track=np.empty((n, ns, 4)), env=np.empty((n, ns, 4)), vmax=np.empty((n, ns)),
                   n_time=np.zeros(n, np.int32), status=np.zeros(n, np.int32), nfev=np.zeros(n, np.int32),
                   flags=np.zeros(n, np.uint32))
        _lib.check(self.lib.tcr_integrate(
            self._h, n, _ptr(ym), _ptr(lon0), _ptr(lat0), _ptr(v0), _ptr(m0), _ptr(h_bl), _ptr(phases),
            _ptr(out["track"]), _ptr(out["env"]), _ptr(out["vmax"]), _ptr(out["n_time"]), _ptr(out["status"]),
            _ptr(out["nfev"]), _ptr(out["flags"]), 0))
        return out

    def rhs_eval(self, ym, t, y, h_bl, phases):
        """Coupled_FAST.dydt(t, y) and _env_winds(y[0], y[1], t) for n independent states: (dydt [n][4], env [n][4])."""
        ym, t, h_bl = _arr(ym, np.int32), _arr(t, np.float64), _arr(h_bl, np.float64)
        n = t.size
        y = _arr(y, np.float64).reshape(n, 4)
        phases = _arr(phases, np.float64).reshape(n, 60)
        dydt, env = np.empty((n, 4)), np.empty((n, 4))
        _lib.check(self.lib.tcr_rhs_eval(self._h, n, _ptr(ym), _ptr(t), _ptr(y), _ptr(h_bl), _ptr(phases), _ptr(dydt), _ptr(env)))
        return dydt, env

    def seed_attempts(self, ym_base, year_key, run_seed, k0, n):
        out = dict(code=np.zeros(n, np.int32), basin=np.zeros(n, np.int32), month=np.zeros(n, np.int32),
                   lon=np.zeros(n), lat=np.zeros(n), v0=np.zeros(n), m0=np.zeros(n), pi_gen=np.zeros(n))
        _lib.check(self.lib.tcr_seed_attempts(
            self._h, int(ym_base), int(year_key), int(run_seed), int(k0), int(n),
            _ptr(out["code"]), _ptr(out["basin"]), _ptr(out["month"]), _ptr(out["lon"]), _ptr(out["lat"]),
            _ptr(out["v0"]), _ptr(out["m0"]), _ptr(out["pi_gen"])))
        return out

    def alloc_results(self, n_years, n_tracks, pinned=True):
        """The 9-tuple's arrays for n_years years (util/compute.py:126-133), optionally page-locked."""
        ns = self.n_steps
        mk = PinnedPool.empty if pinned else (lambda shape, dtype: np.empty(shape, dtype))
        shp = (n_years, n_tracks, ns)
        return dict(lon=mk(shp, np.float64), lat=mk(shp, np.float64), v=mk(shp, np.float64), m=mk(shp, np.float64),
                    vmax=mk(shp, np.float64), env=mk(shp + (4,), np.float64),
                    tc_month=mk((n_years, n_tracks), np.float64), tc_basin=mk((n_years, n_tracks), np.int32),
                    n_seeds=mk((n_years, 7, 12), np.float64))

    def run_years(self, ym_base, year_key, run_seed, n_tracks, out=None, pinned=False):
        """run_tracks(year, n_tracks, b) for several years at once (util/compute.py:64-210)."""
        ym_base, year_key = _arr(ym_base, np.int32), _arr(year_key, np.int32)
        ny = ym_base.size
        if out is None:
            out = self.alloc_results(ny, n_tracks, pinned=pinned)
        stats = (TcrYearStats * ny)()
        _lib.check(self.lib.tcr_run_years(
            self._h, ny, _ptr(ym_base), _ptr(year_key), int(run_seed), int(n_tracks),
            _ptr(out["lon"]), _ptr(out["lat"]), _ptr(out["v"]), _ptr(out["m"]), _ptr(out["vmax"]), _ptr(out["env"]),
            _ptr(out["tc_month"]), _ptr(out["tc_basin"]), _ptr(out["n_seeds"]), stats, 0))
        out["stats"] = [{f: getattr(s, f) for f, _ in TcrYearStats._fields_} for s in stats]
        return out

    def run_years_dev(self, ym_base, year_key, run_seed, n_tracks, dptr):
        """Same with device result pointers (dict of ints); returns the per-year stats."""
        ym_base, year_key = _arr(ym_base, np.int32), _arr(year_key, np.int32)
        ny = ym_base.size
        stats = (TcrYearStats * ny)()
        vp = C.c_void_p
        _lib.check(self.lib.tcr_run_years(
            self._h, ny, _ptr(ym_base), _ptr(year_key), int(run_seed), int(n_tracks),
            vp(dptr["lon"]), vp(dptr["lat"]), vp(dptr["v"]), vp(dptr["m"]), vp(dptr["vmax"]), vp(dptr["env"]),
            vp(dptr["tc_month"]), vp(dptr["tc_basin"]), vp(dptr["n_seeds"]), stats, 1))
        return [{f: getattr(s, f) for f, _ in TcrYearStats._fields_} for s in stats]

    # -- return-period reduction (notebooks/sample_analysis.ipynb cells 13-17) -----------------------
    # -- pre-processing -----------------------------------------------------------
    def wind_stats(self, ua, va, i_upper, i_lower, group_start):
        """calc_wnd_stat (track/env_wind.py:169-228) on the month's samples ua, va
        [n_time, n_level, ...grid] float32 (C-contiguous; the two steering levels are passed as
        level-slice pointers, no copy).  Returns [14, n_pts] float64."""
        ua, va = _arr(ua, np.float32), _arr(va, np.float32)
        gs = _arr(group_start, np.int32)
        n_time, n_lvl = ua.shape[0], ua.shape[1]
        n_pts = int(np.prod(ua.shape[2:]))
        out = np.empty((14, n_pts))
        base_u, base_v = ua.ctypes.data, va.ctypes.data
        vp = C.c_void_p
        _lib.check(self.lib.tcr_wind_stats(
            self._h, n_time, n_pts, n_lvl * n_pts,
            vp(base_u + 4 * n_pts * i_upper), vp(base_v + 4 * n_pts * i_upper),
            vp(base_u + 4 * n_pts * i_lower), vp(base_v + 4 * n_pts * i_lower),
            gs.size - 1, _ptr(gs), _ptr(out), 0))
        return out

    def wind_stats_dev(self, n_time, n_pts, t_stride, d_series, group_start, d_out):
        """Same on device pointers: d_series = four device addresses (ua250, va250, ua850, va850)."""
        gs = _arr(group_start, np.int32)
        vp = C.c_void_p
        _lib.check(self.lib.tcr_wind_stats(self._h, int(n_time), int(n_pts), int(t_stride), vp(d_series[0]), vp(d_series[1]),
                                           vp(d_series[2]), vp(d_series[3]), gs.size - 1, _ptr(gs), vp(d_out), 1))

    def set_entropy_table(self, p_look, s_look, T_lookup):
        """The entropy inversion table of thermo/entropy_table.npz (thermo.py:274-278)."""
        p_look, s_look, T_lookup = (_arr(a, np.float64) for a in (p_look, s_look, T_lookup))
        if T_lookup.shape != (p_look.size, s_look.size):
            raise ValueError("T_lookup must be [len(p), len(s)]")
        _lib.check(self.lib.tcr_set_entropy_table(self._h, p_look.size, s_look.size, _ptr(p_look), _ptr(s_look), _ptr(T_lookup)))

    def set_entropy_table_reversible(self, p_look, s_look, rt_look, T_lookup):
        """The three-dimensional inversion table of thermo/entropy_table_reversible.npz (thermo.py:279-284), for
        namelist.select_thermo = 2."""
        p_look, s_look, rt_look, T_lookup = (_arr(a, np.float64) for a in (p_look, s_look, rt_look, T_lookup))
        if T_lookup.shape != (p_look.size, s_look.size, rt_look.size):
            raise ValueError("T_lookup must be [len(p), len(s), len(rt)]")
        _lib.check(self.lib.tcr_set_entropy_table_reversible(self._h, p_look.size, s_look.size, rt_look.size, _ptr(p_look),
                                                             _ptr(s_look), _ptr(rt_look), _ptr(T_lookup)))

    def thermo_month(self, p_env, ta, hus, sst, psl, ck_over_cd, k_mid, select_thermo=1):
        """vmax, chi, rh_mid of one time sample (thermo/calc_thermo.py:60-69); ta, hus [nlev, ...grid] float32,
        lowest model level first; returns three float64 arrays shaped like sst.  select_thermo as in the namelist:
        1 pseudoadiabatic (set_entropy_table), 2 reversible (set_entropy_table_reversible)."""
        if select_thermo not in (1, 2):
            raise ValueError("select_thermo must be 1 or 2 (namelist.py:59)")
        fn = self.lib.tcr_thermo_month if select_thermo == 1 else self.lib.tcr_thermo_month_reversible
        p_env = _arr(p_env, np.float64)
        ta, hus = _arr(ta, np.float32), _arr(hus, np.float32)
        sst, psl = _arr(sst, np.float64), _arr(psl, np.float64)
        n = sst.size
        if ta.shape[0] != p_env.size or ta.size != p_env.size * n or hus.shape != ta.shape or psl.size != n:
            raise ValueError("inconsistent shapes")
        out = [np.empty(sst.shape) for _ in range(3)]
        _lib.check(fn(self._h, n, p_env.size, _ptr(p_env), _ptr(ta), _ptr(hus), _ptr(sst), _ptr(psl),
                      float(ck_over_cd), int(k_mid), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), 0))
        return tuple(out)

    def thermo_month_dev(self, n_pts, p_env, d_ta, d_hus, d_sst, d_psl, ck_over_cd, k_mid, d_vmax, d_chi, d_rh, select_thermo=1):
        p_env = _arr(p_env, np.float64)
        vp = C.c_void_p
        fn = self.lib.tcr_thermo_month if select_thermo == 1 else self.lib.tcr_thermo_month_reversible
        _lib.check(fn(self._h, int(n_pts), p_env.size, _ptr(p_env), vp(d_ta), vp(d_hus), vp(d_sst), vp(d_psl),
                      float(ck_over_cd), int(k_mid), vp(d_vmax), vp(d_chi), vp(d_rh), 1))

    def poi_vmax(self, lon, lat, vmax, poi_lon, poi_lat, radius_km=100.0, r_earth_m=6378000.0):
        """Per-track maximum of vmax while within radius_km of (poi_lon, poi_lat); NaN if never."""
        lon, lat, vmax = (_arr(x, np.float64) for x in (lon, lat, vmax))
        ns = lon.shape[-1]
        out = np.empty(lon.size // ns)
        _lib.check(self.lib.tcr_poi_vmax(self._h, out.size, ns, _ptr(lon), _ptr(lat), _ptr(vmax), float(poi_lon),
                                         float(poi_lat), float(radius_km), float(r_earth_m), _ptr(out), 0))
        return out.reshape(lon.shape[:-1])

    def poi_vmax_dev(self, n_rows, n_steps, d_lon, d_lat, d_vmax, poi_lon, poi_lat, d_out, radius_km=100.0,
                     r_earth_m=6378000.0):
        vp = C.c_void_p
        _lib.check(self.lib.tcr_poi_vmax(self._h, int(n_rows), int(n_steps), vp(d_lon), vp(d_lat), vp(d_vmax), float(poi_lon),
                                         float(poi_lat), float(radius_km), float(r_earth_m), vp(d_out), 1))

    def exceedance(self, v, bins, on_device_ptr=None, n=None):
        """counts[b] = #(v >= bins[b]); v a host array, or a device pointer with its length n."""
        bins = _arr(bins, np.float64)
        counts = np.zeros(bins.size, np.int64)
        if on_device_ptr is None:
            v = _arr(v, np.float64).reshape(-1)
            _lib.check(self.lib.tcr_exceedance(self._h, v.size, _ptr(v), bins.size, _ptr(bins), _ptr(counts), 0))
        else:
            _lib.check(self.lib.tcr_exceedance(self._h, int(n), C.c_void_p(int(on_device_ptr)), bins.size, _ptr(bins),
                                               _ptr(counts), 1))
        return counts
