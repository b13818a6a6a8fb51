"""Synthetic ERA5-shaped soundings for the thermodynamic pre-processing kernel:
the 28 pressure levels scripts/download_era5.py:80-84 requests, lowest level first as
calc_thermo.py:50-55 arranges them, float32 temperature / specific humidity like the ERA5 files."""
import os

import numpy as np

ERA5_LEVELS_HPA = np.array([1000, 975, 950, 925, 900, 875, 850, 825, 800, 775, 750, 700, 650, 600, 550, 500, 450, 400,
                            350, 300, 250, 225, 200, 175, 150, 125, 100, 70], dtype=np.float64)


def sat_q(T, p):
    """Bolton saturation specific humidity (only used to shape the synthetic humidity profile)."""
    es = 610.94 * np.exp(17.625 * (T - 273.15) / (T - 273.15 + 243.04))
    rs = 0.622 * es / np.maximum(p - es, 1.0)
    return rs / (1 + rs)


def soundings(n, seed=0, edge_cases=True):
    """n columns: (p_env [nlev] Pa, ta [nlev, n] f32, hus [nlev, n] f32, sst [n] f64 K, psl [n] f64 Pa)."""
    rng = np.random.default_rng(seed)
    p = ERA5_LEVELS_HPA * 100.0
    T0 = rng.uniform(272.0, 303.0, n)                                    # near-surface air temperature
    gamma = rng.uniform(5.0, 7.5, n) * 1e-3                              # K/m
    z = 7500.0 * np.log(p[0] / p)[:, None]                               # scale-height altitude
    T_trop = rng.uniform(195.0, 215.0, n)
    ta = np.maximum(T0[None, :] - gamma[None, :] * z, T_trop[None, :]) + rng.normal(0, 0.4, (p.size, n))
    rh = np.clip(rng.uniform(0.55, 0.95, n)[None, :] * np.exp(-z / rng.uniform(3000, 9000, n)[None, :]) +
                 rng.normal(0, 0.03, (p.size, n)), 0.01, 1.0)
    hus = rh * sat_q(ta, p[:, None])
    sst = T0 + rng.uniform(-1.5, 3.5, n)
    psl = rng.uniform(98000.0, 103500.0, n)
    if edge_cases and n >= 64:
        sst[0:8] = 0.0                      # land: nan_to_num(sst) of a Kelvin field (calc_thermo.py:40-42)
        hus[:, 8:12] = 0.0                  # bone-dry column: rh = 0, lambertw(0, -1) = -inf
        hus[0, 12:16] *= 3.0                # super-saturated near-surface parcel (LCL below the surface)
        sst[16:20] = T0[16:20] - 8.0        # sea much colder than the air: no saturated CAPE
        ta[:, 20:24] = 250.0                # isothermal column
        psl[24:28] = 87000.0                # surface pressure below the first levels
        ta[3, 28:30] = np.nan               # missing level
        hus[5, 30:32] = np.nan
    ta = ta.astype(np.float32)
    hus = hus.astype(np.float32)
    sst = sst.astype(np.float32).astype(np.float64)
    psl = psl.astype(np.float32).astype(np.float64)
    return p, ta, hus, sst, psl


def fixture_table():
    """(p_look, s_look, T_lookup): the entropy inversion table the thermodynamic fixtures were generated with
    (tests/golden/ref_thermo.npz, generated in the build container from the reference's
    thermo/entropy_table.npz) -- for benches, smoke runs and tests on machines without a reference checkout.
    Production callers pass their own table (preproc.load_entropy_table)."""
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_thermo.npz")
    with np.load(path) as g:
        return np.array(g["table_p"]), np.array(g["table_s"]), np.array(g["table_T"])
