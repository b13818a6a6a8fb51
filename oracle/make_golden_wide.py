"""Wide reference fixtures: tests/golden/ref_wide_<case>.npz, produced by the UNMODIFIED reference.

TEST INFRASTRUCTURE -- runs only in the build container (needs /root/reference; the reference is
imported the way oracle/ref_harness.py documents).  The fixtures are committed; the GPU box reads
the fixtures, never the reference.

    python oracle/make_golden_wide.py [case ...]

Round-1's ref_tracks.npz pins one basin-month (NA, September, 48 storms, 8-32 N).  These cases
widen the pin to what the reference code actually branches on:

  gl_feb      GL basin (global grid, no crop), February: SOUTHERN-hemisphere genesis (sign(lat) v_beta,
              bam_track.py:141), SI / AU / SP boundary-layer depths 1600 / 1800 / 2000 m
  gl_sep      GL basin, September: both hemispheres, storms near the 0 / 360 E seam of the global grid
              (edge clamping of RectBivariateSpline at lon > 359) and polewards of 45 degrees
  na_jul      NA, July, all five boundary-layer depths of namelist.atm_bl_depth
  na_oct      NA, October (weak potential intensity: many short-lived storms and event terminations)
  wp_aug_900  WP, August, output_interval_s = 900 -> 1441 samples per track (BASELINE configs[4])

Per case, for every storm, Coupled_FAST.gen_track (intensity/coupled_fast.py:229-267: scipy solve_ivp RK45,
dense output, terminal event) + the per-candidate post-processing of run_tracks (util/compute.py:178-206) are
run FIVE times by the reference: once at the genesis point and once with the genesis longitude / latitude
moved by one ulp in each direction.  The spread of the four perturbed twins around the base run is the
reference's OWN sensitivity to rounding (its "chaos envelope"): stored per output sample as the running maximum
of the relative spread, it is what tests may excuse -- reference against reference, not oracle against oracle.

Stored (ragged over the emitted samples, float32 -- the parity bar is 1e-4 relative): track (lon, lat, v, m),
env winds, vmax, chaos; per storm: genesis inputs (float64), status, nfev, n_time, flags, and the nfev /
n_time range of the perturbed twins.
"""
import os
import sys
import time
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh                                   # noqa: E402
from tropical_cyclone_risk_b200 import fields, synth                   # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
GLOBAL_BOUNDS = (0.0, -90.0, 360.0, 90.0)
HBL = np.array([1400.0, 1500.0, 1600.0, 1800.0, 2000.0])               # namelist.atm_bl_depth values

#         name        basin year month interval  n   genesis boxes (lon0, lon1, lat0, lat1), weights
CASES = {
    "gl_feb": ("GL", 2003, 2, 3600, 130, [((45.0, 105.0, -28.0, -6.0), 1), ((150.0, 250.0, -28.0, -6.0), 2)]),
    "gl_sep": ("GL", 2003, 9, 3600, 120, [((125.0, 230.0, 6.0, 34.0), 2), ((325.0, 358.5, 8.0, 30.0), 1),
                                          ((1.0, 359.0, -20.0, -5.0), 1), ((150.0, 220.0, 40.0, 52.0), 1)]),
    "na_jul": ("NA", 2002, 7, 3600, 110, [((283.0, 347.0, 7.0, 34.0), 1)]),
    "na_oct": ("NA", 2002, 10, 3600, 90, [((283.0, 347.0, 7.0, 34.0), 1)]),
    "wp_aug_900": ("WP", 2004, 8, 900, 60, [((124.0, 176.0, 7.0, 32.0), 1)]),
}


def seeds(case, rng):
    _, _, _, _, n, boxes = CASES[case]
    w = np.array([b[1] for b in boxes], dtype=float)
    which = rng.choice(len(boxes), n, p=w / w.sum())
    lon0, lat0 = np.empty(n), np.empty(n)
    for i, k in enumerate(which):
        x0, x1, y0, y1 = boxes[k][0]
        lon0[i], lat0[i] = rng.uniform(x0, x1), rng.uniform(y0, y1)
    v0 = 5.0 + rng.standard_normal(n)
    m0 = rng.uniform(0.13, 0.32, n)
    ph = rng.random((n, 60))
    hbl = HBL[rng.integers(0, HBL.size, n)]
    # a few strong seeds so that every case has kept storms
    v0[::9] += 6.0
    return lon0, lat0, v0, m0, ph, hbl


def rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-3), axis=-1)


def run_case(ref, case):
    basin, year, month, interval, n, _ = CASES[case]
    ref.namelist.output_interval_s = interval
    lon, lat = synth.era5_axes()
    olon, olat = synth.ocean_axes()
    raw = synth.synth_month_raw(year, month, lon, lat)
    mld, strat = synth.synth_ocean(olon, olat, month)
    _, _, planes_g = fields.prepare_month(ref.namelist, GLOBAL_BOUNDS, lon, lat, raw, olon, olat, mld, strat)
    st = synth.synth_static(full_res=False)
    fast = rh.build_fast(ref, basin, lon, lat, planes_g, st)
    crc = zlib.crc32(np.ascontiguousarray(planes_g).tobytes())
    rng = np.random.default_rng(zlib.crc32(case.encode()))
    lon0, lat0, v0, m0, ph, hbl = seeds(case, rng)
    status = np.zeros(n, np.int32); nfev = np.zeros(n, np.int32); n_time = np.zeros(n, np.int32)
    flags = np.zeros(n, np.uint32)
    nfev_lo = np.zeros(n, np.int32); nfev_hi = np.zeros(n, np.int32)
    nt_lo = np.zeros(n, np.int32); nt_hi = np.zeros(n, np.int32)
    status_same = np.ones(n, bool)
    trk, env, vmx, chaos = [], [], [], []
    t0 = time.time()
    for i in range(n):
        r = rh.gen_track(ref, fast, lon0[i], lat0[i], v0[i], m0[i], ph[i], hbl[i])
        status[i], nfev[i], n_time[i] = r["status"], r["nfev"], r["n_time"]
        twins = []
        for dlon, dlat in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            x = np.nextafter(lon0[i], dlon * 1e9) if dlon else lon0[i]
            y = np.nextafter(lat0[i], dlat * 1e9) if dlat else lat0[i]
            twins.append(rh.gen_track(ref, fast, x, y, v0[i], m0[i], ph[i], hbl[i], post=False))
        nfev_lo[i] = min(q["nfev"] for q in twins); nfev_hi[i] = max(q["nfev"] for q in twins)
        nt_lo[i] = min(q["n_time"] for q in twins); nt_hi[i] = max(q["n_time"] for q in twins)
        status_same[i] = all(q["status"] == r["status"] for q in twins)
        k = int(r["n_time"])
        if k == 0:
            continue
        flags[i] = r["flags"]
        assert np.array_equal(r["t"], fast.t_s[:k])
        y = r["y"].T
        c = np.zeros(k)
        for q in twins:
            kk = min(k, int(q["n_time"]))
            if kk:
                c[:kk] = np.maximum(c[:kk], rel(q["y"].T[:kk], y[:kk]))
            c[kk:] = np.inf
        trk.append(y); env.append(r["env"]); vmx.append(r["vmax"]); chaos.append(np.maximum.accumulate(c))
    off = np.concatenate([[0], np.cumsum(n_time)]).astype(np.int64)
    f32 = lambda parts, shape: (np.concatenate(parts).astype(np.float32) if parts else np.zeros(shape, np.float32))
    out = dict(basin=basin, year=year, month=month, interval=interval, planes_crc=crc,
               lon0=lon0, lat0=lat0, v0=v0, m0=m0, phases=ph, h_bl=hbl,
               status=status, nfev=nfev, n_time=n_time, flags=flags, off=off,
               nfev_twin_lo=nfev_lo, nfev_twin_hi=nfev_hi, n_time_twin_lo=nt_lo, n_time_twin_hi=nt_hi,
               status_twin_same=status_same,
               track=f32(trk, (0, 4)), env=f32(env, (0, 4)), vmax=f32(vmx, (0,)), chaos=f32(chaos, (0,)))
    fn = os.path.join(GOLDEN, "ref_wide_%s.npz" % case)
    np.savez_compressed(fn, **out)
    ch = np.array([c.max() if c.size else 0.0 for c in chaos])
    print("%-11s %3d storms in %5.1f s: status %s, kept %d, is_tc %d, samples %d, sensitive (self-spread > 1e-6): %d, "
          "twin nfev differs: %d, %d bytes" % (
              case, n, time.time() - t0, np.bincount(status + 1, minlength=4).tolist(), int((flags & 2).astype(bool).sum()),
              int((flags & 1).astype(bool).sum()), int(off[-1]), int((ch > 1e-6).sum()),
              int(((nfev_lo != nfev) | (nfev_hi != nfev)).sum()), os.path.getsize(fn)))


def main():
    if not rh.available():
        raise SystemExit("reference tree not present; fixtures can only be generated in the build container")
    ref = rh.load_reference()
    interval0 = ref.namelist.output_interval_s
    try:
        for case in (sys.argv[1:] or CASES):
            run_case(ref, case)
    finally:
        ref.namelist.output_interval_s = interval0


if __name__ == "__main__":
    main()
