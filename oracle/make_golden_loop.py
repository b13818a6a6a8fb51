"""Reference-derived pins for the two parts of run_tracks that round 1 only restated:

  a1  the seeding / rejection loop and the sequential acceptance       util/compute.py:123-209
  N1  the per-month field preparation                                  util/compute.py:76-84, 101-121

TEST INFRASTRUCTURE -- build container only (needs /root/reference).  Both are pinned by EXECUTING THE REFERENCE'S
OWN SOURCE LINES (read from /root/reference/util/compute.py at run time, never copied into this repo) inside a
namespace that supplies what the surrounding function would have set up with xarray:

    python oracle/make_golden_loop.py        ->  tests/golden/ref_loop.npz, tests/golden/ref_prep.npz

a1.  Lines 123-209 (output arrays, `while nt < n_tracks`, `while not seed_passed`, gen_track, TC criteria, env-wind
recompute, vmax test) run unmodified.  `np.random` inside that namespace is an indexed stream: every pass of the
inner loop body is attempt k = 0, 1, 2, ... of (run_seed, year) and its draws are the Philox draws the build assigns
to attempt k (draw block 0 = first (lon, lat), 3.. = ocean-point redraws, 1 = (month, low-latitude test), 2 =
Box-Muller pair of v_init; stream 1 = the 60 Fourier phases of gen_f).  cpl_fast[] are genuine Coupled_FAST objects
on the synthetic fields (oracle/ref_harness.build_fast), f_b / f_basins the reference's own mat.interp2_fx of the
mask planes.  Two runs:
  * `seed`: gen_track replaced by a recorder that returns None -- 30 000 attempts of pure seeding logic; every
    gen_track call (attempt, month, lon, lat, v_init, m_init, h_bl) and the n_seeds table are stored;
  * `loop`: the whole loop with the genuine gen_track until n_tracks storms are kept: the 9-tuple is stored.

N1.  Lines 76-84 and 101-121 run over shim objects for the three xarray idioms they use (`da * scalar`,
`.reindex({'lat': ...})`, `.interp(time=...).data`, `mld['lon']`, `mld[:, :, i]`); Coupled_FAST is replaced by a
recorder of the init_fields arguments.  Input latitude is stored DESCENDING so that the flip of :80-84 runs.
"""
import datetime
import os
import sys
import textwrap
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh                                   # noqa: E402
from oracle import tcr_oracle as orc                                   # noqa: E402
from tropical_cyclone_risk_b200 import fields, layout, synth           # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
GLOBAL_BOUNDS = (0.0, -90.0, 360.0, 90.0)
COMPUTE_PY = os.path.join(rh.REF_ROOT, "util", "compute.py")


def source_lines(first, last):
    """Lines first..last (1-based, inclusive) of the reference's util/compute.py, dedented."""
    with open(COMPUTE_PY) as f:
        lines = f.readlines()
    return textwrap.dedent("".join(lines[first - 1:last]))


def check_anchors():
    """The line numbers above are those of the pinned commit; fail loudly if the file differs."""
    with open(COMPUTE_PY) as f:
        lines = f.readlines()
    assert lines[75].strip().startswith("vpot = ds['vmax'] * namelist.PI_reduc"), lines[75]
    assert lines[79].strip().startswith("if (lat[0] - lat[1]) > 0:"), lines[79]
    assert lines[100].strip() == "cpl_fast = [0] * 12", lines[100]
    assert "init_fields(lon, lat, chi_month, vpot_month, mld_month, strat_month)" in lines[120], lines[120]
    assert lines[122].strip() == "# Output vectors.", lines[122]
    assert lines[133].strip() == "while nt < n_tracks:", lines[133]
    assert lines[209].strip().startswith("return((tc_lon"), lines[209]


# ---------------------------------------------------------------------------------------------
# a1: the indexed random stream behind np.random
# ---------------------------------------------------------------------------------------------
class StopRun(Exception):
    pass


class IndexedRandom:
    """np.random as the loop sees it: a state machine over the call pattern of util/compute.py:144-172."""

    def __init__(self, run_seed, year_key, max_attempts):
        self.run_seed, self.year_key, self.max_attempts = run_seed, year_key, max_attempts
        self.k, self.state, self.redraws, self.max_redraws = -1, "start", 0, 0

    def _draw(self, blk, stream=0):
        u = np.empty(2)
        import ctypes as C
        orc.lib().orc_draw2(C.c_uint32(self.run_seed), C.c_int32(self.year_key), C.c_int64(self.k), C.c_uint32(blk),
                            C.c_uint32(stream), u.ctypes.data_as(C.POINTER(C.c_double)))
        return u

    def uniform(self, lo, hi, size):
        assert size == 1
        if self.state == "start":                                # :144 first draw of a new attempt
            self.k += 1
            if self.k >= self.max_attempts:
                raise StopRun()
            self.u, self.redraws, self.state, u = self._draw(0), 0, "lat", None
            u = self.u[0]
        elif self.state == "lat":                                # :145
            u, self.state = self.u[1], "body"
        elif self.state == "body":                               # :147 redraw longitude
            self.u = self._draw(3 + self.redraws)
            self.redraws += 1
            self.max_redraws = max(self.max_redraws, self.redraws)
            u, self.state = self.u[0], "relat"
        elif self.state == "relat":                              # :148 redraw latitude (uniform in latitude)
            u, self.state = self.u[1], "body"
        elif self.state == "lowlat":                             # :165
            u, self.state = self.m[1], "start"
        else:
            raise AssertionError(self.state)
        return np.array([lo + (hi - lo) * u])                    # numpy's own formula for uniform(lo, hi)

    def randint(self, lo, hi):                                   # :151
        assert self.state == "body" and (lo, hi) == (1, 13)
        self.m, self.state = self._draw(1), "lowlat"
        return lo + int(np.floor(self.m[0] * (hi - lo)))

    def randn(self, n):                                          # :172 (only after a passed attempt)
        assert n == 1 and self.state == "start"
        u = self._draw(2)
        return np.array([np.sqrt(-2.0 * np.log(1.0 - u[0])) * np.cos(2.0 * np.pi * u[1])])


class NumpyWithStream:
    """`np` of the executed lines: numpy, except for `.random`."""

    def __init__(self, stream):
        self.random = stream

    def __getattr__(self, name):
        return getattr(np, name)


def year_objects(ref, basin_id, year):
    """cpl_fast[12], m_init_fx[12], f_b, f_basins, basin_ids, b_bounds on the synthetic fields."""
    lon, lat = synth.era5_axes()
    olon, olat = synth.ocean_axes()
    st = synth.synth_static(full_res=False)
    cpl_fast, m_init_fx = [], []
    for month in range(1, 13):
        raw = synth.synth_month_raw(year, month, lon, lat)
        mld, strat = synth.synth_ocean(olon, olat, month)
        _, _, pg = fields.prepare_month(ref.namelist, GLOBAL_BOUNDS, lon, lat, raw, olon, olat, mld, strat)
        f = rh.build_fast(ref, basin_id, lon, lat, pg, st)
        cpl_fast.append(f)
        m_init_fx.append(f.m_init_fx)
    basin_ids = np.array(sorted([k for k in ref.namelist.basin_bounds if k != 'GL']))
    assert tuple(basin_ids) == layout.BASIN_IDS
    masks = fields.mask_planes(st, basin_id).astype(np.float64)
    f_basins = {bid: ref.mat.interp2_fx(st["lon_m"], st["lat_m"], masks[i]) for i, bid in enumerate(basin_ids)}
    f_b = ref.mat.interp2_fx(st["lon_m"], st["lat_m"], masks[7])
    b = ref.basins.TC_Basin(basin_id)
    return cpl_fast, m_init_fx, f_b, f_basins, basin_ids, b.get_bounds()


def run_loop(ref, basin_id, year, run_seed, n_tracks, max_attempts, integrate):
    cpl_fast, m_init_fx, f_b, f_basins, basin_ids, b_bounds = year_objects(ref, basin_id, year)
    stream = IndexedRandom(run_seed, year, max_attempts)
    calls, results = [], []

    for month0, f in enumerate(cpl_fast):
        def gen_track(clon, clat, v, m, _f=f, _month=month0 + 1):
            calls.append((stream.k, _month, float(clon), float(clat), float(v), float(m), float(_f.h_bl)))
            if not integrate:
                return None
            with rh.injected_phases(orc.phases_for(run_seed, year, stream.k)):
                res = type(_f).gen_track(_f, clon, clat, v, m)
            results.append(None if res is None else res.y.copy())
            return res
        f.gen_track = gen_track

    ns = dict(np=NumpyWithStream(stream), namelist=ref.namelist, tc_wind=ref.tc_wind, cpl_fast=cpl_fast,
              m_init_fx=m_init_fx, f_b=f_b, f_basins=f_basins, basin_ids=basin_ids, b_bounds=b_bounds,
              n_tracks=n_tracks, n_seeds=np.zeros((len(basin_ids), 12)))
    code = compile(source_lines(123, 209), COMPUTE_PY + ":123-209", "exec")
    try:
        with np.errstate(all="ignore"):
            exec(code, ns)
    except StopRun:
        pass
    ns["_results"], ns["_cpl_fast"] = results, cpl_fast
    return ns, np.array(calls, dtype=np.float64).reshape(-1, 7), stream


def kept_envelopes(ref, ns, calls, run_seed, year):
    """For every kept track: the attempt that produced it and the reference's own sensitivity -- the running maximum,
    per output sample, of the relative spread of four reference runs whose genesis point was moved by one ulp."""
    rel = lambda a, b: np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-3), axis=-1)
    n_tracks, n_steps = ns["tc_lon"].shape
    attempt = np.full(n_tracks, -1, np.int64)
    chaos = np.full((n_tracks, n_steps), np.nan, np.float32)
    for r in range(n_tracks):
        n = int(np.sum(~np.isnan(ns["tc_lon"][r])))
        hit = [j for j, y in enumerate(ns["_results"]) if y is not None and y.shape[1] == n and np.array_equal(y[0], ns["tc_lon"][r, :n])]
        assert len(hit) == 1
        k, month, lon0, lat0, v0, m0, hbl = calls[hit[0]]
        attempt[r] = int(k)
        f = ns["_cpl_fast"][int(month) - 1]
        f.__dict__.pop("gen_track", None)                        # back to the class's own method
        base = ns["_results"][hit[0]].T
        c = np.zeros(n)
        for dlon, dlat in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            x = np.nextafter(lon0, dlon * 1e9) if dlon else lon0
            y = np.nextafter(lat0, dlat * 1e9) if dlat else lat0
            q = rh.gen_track(ref, f, x, y, v0, m0, orc.phases_for(run_seed, year, int(k)), hbl, post=False)
            kk = min(n, int(q["n_time"]))
            c[:kk] = np.maximum(c[:kk], rel(q["y"].T[:kk], base[:kk]))
            c[kk:] = np.inf
        chaos[r, :n] = np.maximum.accumulate(c)
    return attempt, chaos


# ---------------------------------------------------------------------------------------------
# N1: shims for the xarray idioms of lines 76-84 and 101-121
# ---------------------------------------------------------------------------------------------
class TimeField:
    """A (time, lat, lon) DataArray stand-in: `* scalar`, `.reindex({'lat': ...})`, `.interp(time=t).data`."""

    def __init__(self, times, lat, data):
        self.times, self.lat, self.values = list(times), np.asarray(lat), np.asarray(data, dtype=np.float64)

    def __mul__(self, s):
        return TimeField(self.times, self.lat, self.values * s)

    def reindex(self, idx):
        new_lat = np.asarray(idx['lat'])
        order = [int(np.flatnonzero(self.lat == v)[0]) for v in new_lat]
        return TimeField(self.times, new_lat, self.values[:, order, :])

    def interp(self, time):
        # the fixtures sample exactly on the 15th of each month, where time interpolation is the identity; what
        # xarray / scipy do between samples is not part of this pin (refdata.py restates it, tests/test_refdata.py)
        j = self.times.index(time)
        return types.SimpleNamespace(data=np.array(self.values[j]))


class OceanClim:
    """mld / strat as ocean.mld_climatology returns them: dims (lat, lon, month), `['lon']`, `['lat']`, `[:, :, i]`."""

    def __init__(self, lon, lat, data):
        self.coords = {"lon": np.asarray(lon), "lat": np.asarray(lat)}
        self.values = np.asarray(data)

    def __getitem__(self, key):
        return self.coords[key] if isinstance(key, str) else self.values[key]


class RecordingFast:
    def __init__(self, fn_wnd_stat, b, dt, dt_s, total_time_s):
        self.ctor = (dt, dt_s, total_time_s)

    def init_fields(self, lon, lat, chi, vpot, mld, strat):
        self.args = dict(lon=np.array(lon), lat=np.array(lat), chi=np.array(chi), vpot=np.array(vpot),
                         mld=np.array(mld), strat=np.array(strat))


PREP_RES = 4.0            # a coarse global grid keeps the fixture small; the lines under test are grid-agnostic
PREP_KEEP = (0, 6, 11)    # months whose outputs are stored


def prep_inputs(year):
    """Deterministic inputs of the N1 pin (tests regenerate them instead of reading them from the fixture):
    raw thermo stacks [12][nlat][nlon] float32 with 2 % NaN holes on the ASCENDING grid, ocean climatologies
    [nlat_o][nlon_o][12] float32 with NaNs."""
    lon, lat = synth.era5_axes(PREP_RES)
    olon = np.arange(0.0, 360.0, 2.0)
    olat = -89.0 + 2.0 * np.arange(90.0)
    raws = [synth.synth_month_raw(year, m, lon, lat) for m in range(1, 13)]
    rng = np.random.default_rng(99)
    stack = {k: np.stack([r[k] for r in raws]).astype(np.float32) for k in ("vmax", "chi", "rh_mid")}
    for k in stack:                                              # the NaN policies must have something to act on
        stack[k][rng.random(stack[k].shape) < 0.02] = np.nan
    oc = [synth.synth_ocean(olon, olat, m) for m in range(1, 13)]
    mld = np.stack([o[0] for o in oc], axis=2).astype(np.float32)
    strat = np.stack([o[1] for o in oc], axis=2).astype(np.float32)
    mld[rng.random(mld.shape) < 0.05] = np.nan                   # the Levitus files are NaN over land
    strat[np.isnan(mld)] = np.nan
    return lon, lat, olon, olat, raws, stack, mld, strat


class RecordingMat:
    """util.mat, remembering the fields handed to interp2_fx (compute.py:114: the rh_mid planes of m_init_fx)."""

    def __init__(self, mat):
        self._mat, self.fields = mat, []

    def interp2_fx(self, lon, lat, X):
        self.fields.append(np.array(X))
        return self._mat.interp2_fx(lon, lat, np.nan_to_num(X))   # the spline object itself is not used by the pin

    def __getattr__(self, name):
        return getattr(self._mat, name)


def run_prep(ref, year):
    lon, lat, olon, olat, _, stack, mld, strat = prep_inputs(year)
    times = [datetime.datetime(year, m, 15) for m in range(1, 13)]
    lat_desc = lat[::-1].copy()
    ds = {k: TimeField(times, lat_desc, v[:, ::-1, :]) for k, v in stack.items()}
    ns = dict(np=np, namelist=ref.namelist, datetime=datetime, mat=RecordingMat(ref.mat), ds=ds, lon=lon, lat=lat_desc, year=year,
              mld=OceanClim(olon, olat, mld), strat=OceanClim(olon, olat, strat), b=None, basin_ids=list(layout.BASIN_IDS),
              input=types.SimpleNamespace(convert_from_datetime=lambda ds_, dts: dts),
              env_wind=types.SimpleNamespace(get_env_wnd_fn=lambda: "unused"),
              xr=types.SimpleNamespace(open_dataset=lambda fn: None),
              coupled_fast=types.SimpleNamespace(Coupled_FAST=RecordingFast))
    exec(compile(source_lines(76, 84), COMPUTE_PY + ":76-84", "exec"), ns)
    exec(compile(source_lines(101, 121), COMPUTE_PY + ":101-121", "exec"), ns)
    rec = ns["cpl_fast"]
    assert np.array_equal(rec[0].args["lat"], lat) and np.array_equal(rec[0].args["lon"], lon)     # flipped to ascending
    out = dict(year=year, months=np.array(PREP_KEEP), T_s=np.float64(rec[0].ctor[2]), dt_s=np.float64(rec[0].ctor[1]))
    for key in ("chi", "vpot", "mld", "strat"):
        out["ref_" + key] = np.stack([rec[i].args[key] for i in PREP_KEEP])
    out["ref_rh"] = np.stack([ns["mat"].fields[i] for i in PREP_KEEP])       # rh_mid as handed to m_init_fx (:114)
    return out


def main():
    if not rh.available():
        raise SystemExit("reference tree not present; fixtures can only be generated in the build container")
    check_anchors()
    ref = rh.load_reference()
    os.makedirs(GOLDEN, exist_ok=True)

    # ---- a1, seeding only
    out = {}
    for basin_id, year, run_seed, n_att in (("NA", 2001, 777, 30000), ("GL", 2002, 4242, 30000), ("SI", 2002, 31, 12000)):
        ns, calls, stream = run_loop(ref, basin_id, year, run_seed, n_tracks=4, max_attempts=n_att, integrate=False)
        tag = "seed_%s_" % basin_id
        out[tag + "calls"] = calls
        out[tag + "n_seeds"] = ns["n_seeds"]
        out[tag + "meta"] = np.array([year, run_seed, n_att, stream.max_redraws])
        print("seed %s: %d attempts, %d gen_track calls, %d counted seeds, longest redraw chain %d" % (
            basin_id, n_att, calls.shape[0], int(ns["n_seeds"].sum()), stream.max_redraws))

    # ---- a1, the whole loop with the genuine gen_track
    for basin_id, year, run_seed, n_tracks in (("NA", 2001, 20260101, 6),):
        ns, calls, stream = run_loop(ref, basin_id, year, run_seed, n_tracks=n_tracks, max_attempts=10 ** 9, integrate=True)
        tag = "loop_%s_" % basin_id
        for key in ("tc_lon", "tc_lat", "tc_v", "tc_m", "tc_vmax", "tc_env_wnds", "tc_month", "n_seeds"):
            out[tag + key] = np.asarray(ns[key])
        out[tag + "tc_basin"] = np.array([layout.BASIN_IDS.index(x) for x in ns["tc_basin"]], dtype=np.int32)
        out[tag + "calls"] = calls
        out[tag + "attempt"], out[tag + "chaos"] = kept_envelopes(ref, ns, calls, run_seed, year)
        out[tag + "meta"] = np.array([year, run_seed, n_tracks, stream.k + 1])
        print("loop %s: %d tracks after %d attempts, %d gen_track calls, months %s" % (
            basin_id, n_tracks, stream.k + 1, calls.shape[0], ns["tc_month"].tolist()))
    np.savez_compressed(os.path.join(GOLDEN, "ref_loop.npz"), **out)

    # ---- N1
    prep = run_prep(ref, 2001)
    np.savez_compressed(os.path.join(GOLDEN, "ref_prep.npz"), **prep)
    for f in ("ref_loop.npz", "ref_prep.npz"):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)), "bytes")


if __name__ == "__main__":
    main()
