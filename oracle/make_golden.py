"""Generate tests/golden/*.npz by running the UNMODIFIED reference hot-path modules.

TEST INFRASTRUCTURE -- runs only in the build container (needs /root/reference; see
oracle/ref_harness.py for how the reference is imported).  The fixtures it writes are
committed; tests on the GPU box read the fixtures, never the reference.

    python oracle/make_golden.py            # rewrites tests/golden/ref_*.npz

What is pinned (reference file:line of the code that produced each array):
  ref_bilinear.npz   RectBivariateSpline(kx=1,ky=1).ev through util/mat.py:142-153
  ref_fourier.npz    gen_f table, track/bam_track.py:23-31
  ref_rhs.npz        Coupled_FAST.dydt / _env_winds, intensity/coupled_fast.py:196-207,
                     track/bam_track.py:116-128
  ref_tracks.npz     Coupled_FAST.gen_track (scipy solve_ivp RK45), coupled_fast.py:229-267,
                     + env-wind recompute and axi_to_max_wind (util/compute.py:201-203,
                     wind/tc_wind.py:6-21) and the TC criteria (util/compute.py:185-189,205)
Inputs are the deterministic synthetic fields of tropical_cyclone_risk_b200.synth (seeded);
a checksum of the prepared planes is stored so drift of the generator is detected.
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh                                   # noqa: E402
from tropical_cyclone_risk_b200 import fields, params, synth           # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
BASIN, YEAR, MONTH = "NA", 2000, 9
GLOBAL_BOUNDS = (0.0, -90.0, 360.0, 90.0)


def setup(ref, basin=BASIN, year=YEAR, month=MONTH):
    lon, lat = synth.era5_axes()
    olon, olat = synth.ocean_axes()
    raw = synth.synth_month_raw(year, month, lon, lat)
    mld, strat = synth.synth_ocean(olon, olat, month)
    _, _, planes_g = fields.prepare_month(ref.namelist, GLOBAL_BOUNDS, lon, lat, raw, olon, olat, mld, strat)
    st = synth.synth_static(full_res=False)
    fast = rh.build_fast(ref, basin, lon, lat, planes_g, st)
    return lon, lat, planes_g, st, fast


def seeds(n, rng):
    """Genesis points biased to the warm, open-ocean part of the NA box plus a few edge cases."""
    lon0 = rng.uniform(285.0, 345.0, n)
    lat0 = rng.uniform(8.0, 32.0, n)
    v0 = 5.0 + rng.standard_normal(n)
    m0 = rng.uniform(0.13, 0.32, n)
    ph = rng.random((n, 60))
    return lon0, lat0, v0, m0, ph


def main():
    if not rh.available():
        raise SystemExit("reference tree not present; fixtures can only be generated in the build container")
    os.makedirs(GOLDEN, exist_ok=True)
    ref = rh.load_reference()
    lon, lat, planes_g, st, fast = setup(ref)
    crc = zlib.crc32(np.ascontiguousarray(planes_g).tobytes())
    rng = np.random.default_rng(20260101)

    # ---- bilinear: every monthly field + bathymetry + land at random points, incl. outside the box
    nq = 4000
    qlon = rng.uniform(255.0, 365.0, nq)
    qlat = rng.uniform(-5.0, 65.0, nq)
    qlon[:8] = [260.0, 359.0, 300.0, 300.5, 259.0, 361.0, 310.25, 280.0]
    qlat[:8] = [0.0, 60.0, 30.0, 30.5, -1.0, 61.0, 59.999, 0.5]
    vals = np.empty((nq, 21))
    for i in range(4):
        vals[:, i] = fast.wnd_Mean_Fxs[i].ev(qlon, qlat)
    for i in range(4):
        for j in range(i + 1):
            vals[:, 4 + i * (i + 1) // 2 + j] = fast.wnd_Cov_Fxs[i][j].ev(qlon, qlat)
    vals[:, 14] = fast.f_chi.ev(qlon, qlat)
    vals[:, 15] = fast.f_vpot.ev(qlon, qlat)
    vals[:, 16] = fast.f_mld.ev(qlon, qlat)
    vals[:, 17] = fast.f_strat.ev(qlon, qlat)
    vals[:, 18] = fast.m_init_fx.ev(qlon, qlat)
    vals[:, 19] = fast.f_bath.ev(qlon, qlat)
    vals[:, 20] = fast.f_land.ev(qlon, qlat)
    np.savez_compressed(os.path.join(GOLDEN, "ref_bilinear.npz"), lon=qlon, lat=qlat, vals=vals, planes_crc=crc)

    # ---- Fourier table of gen_f with injected phases
    nf = 6
    phases = rng.random((nf, 60))
    tabs = np.empty((nf, 4, fast.total_steps))
    for i in range(nf):
        with rh.injected_phases(phases[i]):
            tabs[i] = fast.gen_synthetic_f()
    np.savez_compressed(os.path.join(GOLDEN, "ref_fourier.npz"), phases=phases, table=tabs, t_s=fast.t_s)

    # ---- single RHS evaluations and env winds
    from scipy.interpolate import interp1d
    nr = 300
    lon0, lat0, v0, m0, ph = seeds(nr, rng)
    lat0[:6] = [81.0, -3.0, 1.5, 25.0, 25.0, 25.0]
    lon0[3:6] = [270.0, 303.0, 330.0]                        # over land, negative-strat patch, open ocean
    v0 = np.abs(v0) * rng.uniform(0.5, 8.0, nr)
    tq = rng.uniform(0.0, fast.total_time, nr)
    tq[:3] = [0.0, fast.total_time, 3600.0]
    hbl = rng.choice([1400.0, 1800.0, 2000.0], nr)
    dy = np.empty((nr, 4))
    ew = np.empty((nr, 4))
    for i in range(nr):
        with rh.injected_phases(ph[i]):
            fast.Fs = fast.gen_synthetic_f()
        fast.Fs_i = interp1d(fast.t_s, fast.Fs, axis=1)
        fast.h_bl = hbl[i]
        with np.errstate(all="ignore"):
            dy[i] = fast.dydt(tq[i], np.array([lon0[i], lat0[i], v0[i], m0[i]]))
            ew[i] = fast._env_winds(lon0[i], lat0[i], tq[i])
    np.savez_compressed(os.path.join(GOLDEN, "ref_rhs.npz"), lon=lon0, lat=lat0, v=v0, m=m0, t=tq, h_bl=hbl,
                        phases=ph, dydt=dy, env_winds=ew, planes_crc=crc)

    # ---- whole tracks
    nt = 48
    lon0, lat0, v0, m0, ph = seeds(nt, rng)
    v0[:4] = [3.9, 4.0, 12.0, 20.0]                         # event at t0 (g == 0), strong seeds
    lat0[4], lon0[4] = 1.0, 320.0                            # |lat| <= 2 at genesis
    lon0[5] = 358.5                                          # outside the shrunk basin box at genesis
    hbl = np.full(nt, 1400.0)
    hbl[1::3] = 1800.0
    ns = fast.total_steps
    track = np.full((nt, ns, 4), np.nan)
    env = np.full((nt, ns, 4), np.nan)
    vmax = np.full((nt, ns), np.nan)
    n_time = np.zeros(nt, np.int32)
    status = np.zeros(nt, np.int32)
    nfev = np.zeros(nt, np.int32)
    flags = np.zeros(nt, np.uint32)
    for i in range(nt):
        r = rh.gen_track(ref, fast, lon0[i], lat0[i], v0[i], m0[i], ph[i], hbl[i])
        status[i], nfev[i], n_time[i] = r["status"], r["nfev"], r["n_time"]
        if r["n_time"] > 0:
            k = r["n_time"]
            assert np.array_equal(r["t"], fast.t_s[:k])
            track[i, :k] = r["y"].T
            env[i, :k] = r["env"]
            vmax[i, :k] = r["vmax"]
            flags[i] = r["flags"]
    np.savez_compressed(os.path.join(GOLDEN, "ref_tracks.npz"), lon0=lon0, lat0=lat0, v0=v0, m0=m0, phases=ph,
                        h_bl=hbl, track=track, env=env, vmax=vmax, n_time=n_time, status=status, nfev=nfev,
                        flags=flags, planes_crc=crc, basin=BASIN, year=YEAR, month=MONTH)
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))
    print("status", np.bincount(status + 1), "n_time", n_time.tolist())


if __name__ == "__main__":
    main()
