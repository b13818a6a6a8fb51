"""``tcr_params`` POD mirror (include/tcrisk.h) and its construction from a namelist.

The reference reads module-level ``namelist`` globals at call time in every hot-path
function (e.g. intensity/coupled_fast.py:23-24,185-191; track/bam_track.py:56-59;
util/compute.py:140-175).  Here they are read once, on the host, into one struct that
crosses the C ABI.
"""
import ctypes as C
import math

import numpy as np

from . import layout

N_BASINS = len(layout.BASIN_IDS)


class TcrParams(C.Structure):
    _fields_ = [
        ("dt_track", C.c_double), ("total_time", C.c_double), ("T_Fs", C.c_double),
        ("max_step", C.c_double), ("rtol", C.c_double), ("atol", C.c_double),
        ("u_beta", C.c_double), ("v_beta", C.c_double),
        ("steering_coefs", C.c_double * 2),
        ("y_alpha", C.c_double * 2), ("m_alpha", C.c_double * 2),
        ("alpha_max", C.c_double * 2), ("alpha_min", C.c_double * 2),
        ("Ck", C.c_double), ("epsilon", C.c_double), ("kappa", C.c_double), ("beta", C.c_double),
        ("earth_R", C.c_double),
        ("basin_bounds", C.c_double * 4),
        ("gen_lat_min", C.c_double), ("gen_lat_max", C.c_double),
        ("lat_vort_fac", C.c_double),
        ("lat_vort_power", C.c_double * N_BASINS),
        ("atm_bl_depth", C.c_double * N_BASINS),
        ("seed_v_init", C.c_double), ("seed_v_2d_thresh", C.c_double),
        ("seed_v_thresh", C.c_double), ("seed_vmax_thresh", C.c_double),
        ("pi_gen_min", C.c_double),
        ("minit_amp", C.c_double), ("minit_center", C.c_double),
        ("minit_slope", C.c_double), ("minit_offset", C.c_double),
        ("fourier_amp", C.c_double * 15),
        ("n_steps", C.c_int32), ("coupled_track", C.c_int32),
        ("max_redraws", C.c_int32), ("reserved", C.c_int32),
    ]


class TcrYearStats(C.Structure):
    _fields_ = [
        ("attempts", C.c_int64), ("counted_seeds", C.c_int64), ("integrated", C.c_int64),
        ("storm_steps", C.c_int64), ("kept_steps", C.c_int64), ("rhs_evals", C.c_int64),
        ("wasted_integrated", C.c_int64), ("wasted_steps", C.c_int64), ("wasted_rhs_evals", C.c_int64),
        ("n_kept", C.c_int32), ("n_waves", C.c_int32), ("redraw_exhausted", C.c_int64),
    ]


class TcrPrepSpec(C.Structure):
    _fields_ = [("nlat_g", C.c_int32), ("nlon_g", C.c_int32), ("nlat_o", C.c_int32), ("nlon_o", C.c_int32),
                ("nlat_b", C.c_int32), ("nlon_b", C.c_int32),
                ("pi_reduc", C.c_double), ("sqrt_ck_cd", C.c_double), ("log_chi_fac", C.c_double), ("chi_fac", C.c_double)]


def parse_bound(text):
    """'45S' -> -45.0, '0S' -> -0.0 (sign preserved, as TC_Basin._adj_bnd, util/basins.py:23-27)."""
    x = float(text[:-1])
    if text[-1] in ("W", "S"):
        x *= -1
    return x


def basin_bounds(namelist, basin_id):
    """(lon_min, lat_min, lon_max, lat_max), util/basins.py:42-50."""
    if basin_id.upper() not in namelist.basin_bounds:
        raise ValueError("Basin ID is not valid. See list of valid basins.")
    return tuple(parse_bound(s) for s in namelist.basin_bounds[basin_id])


def _np_sign_ge0(x):
    # np.sign(-0.0) is -0.0 and (-0.0 >= 0) is True: '0S' counts as non-negative
    # (util/compute.py:140-141; SURVEY.md section 7 "SH-basin genesis range")
    return not (x < 0)


def _check_minit(namelist):
    """f_mInit is a lambda in the reference namelist (namelist.py:94); the device evaluates the
    same logistic from four scalars, so refuse a namelist whose f_mInit is something else."""
    amp, center, slope, offset = 0.20, 0.55, 10.0, 0.125
    f = getattr(namelist, "f_mInit", None)
    if f is not None:
        for rh in (0.0, 0.2, 0.55, 0.7, 1.0):
            want = amp / (1 + math.exp(-(rh - center) * slope)) + offset
            if abs(float(f(rh)) - want) > 1e-12:
                raise NotImplementedError(
                    "namelist.f_mInit differs from 0.20/(1+exp(-(rh-0.55)*10))+0.125; "
                    "only that logistic form is implemented on the device")
    return amp, center, slope, offset


def params_from_namelist(namelist, basin_id, max_redraws=64):
    """Build the POD parameter block for one run basin from a reference-style namelist module."""
    if list(namelist.steering_levels) != [250, 850]:
        raise NotImplementedError("only the two-level (250/850 hPa) steering configuration is built")
    p = TcrParams()
    p.dt_track = float(namelist.output_interval_s)
    p.total_time = float(namelist.total_track_time_days * 24 * 60 * 60)
    p.n_steps = int(p.total_time / p.dt_track) + 1            # track/bam_track.py:54
    p.T_Fs = float(namelist.T_days * 24 * 60 * 60)
    p.max_step = 86400.0                                       # intensity/coupled_fast.py:266
    p.rtol, p.atol = 1e-3, 1e-6                                # scipy solve_ivp defaults
    p.u_beta, p.v_beta = float(namelist.u_beta), float(namelist.v_beta)
    for i in range(2):
        p.steering_coefs[i] = float(namelist.steering_coefs[i])
        p.y_alpha[i] = float(namelist.y_alpha[i])
        p.m_alpha[i] = float(namelist.m_alpha[i])
        p.alpha_max[i] = float(namelist.alpha_max[i])
        p.alpha_min[i] = float(namelist.alpha_min[i])
    p.coupled_track = 1 if namelist.coupled_track else 0
    p.Ck = float(namelist.Ck)
    p.epsilon, p.kappa = 0.33, 0.1                             # intensity/coupled_fast.py:25-26
    p.beta = 1 - p.epsilon - p.kappa                           # :27 (0.5700000000000001 in binary64)
    p.earth_R = 6.3781 * (10 ** 6)                             # util/constants.py:7
    b = basin_bounds(namelist, basin_id)
    for i in range(4):
        p.basin_bounds[i] = b[i]
    p.gen_lat_min = 3.0 if _np_sign_ge0(b[1]) else -45.0       # util/compute.py:140
    p.gen_lat_max = 45.0 if _np_sign_ge0(b[3]) else -3.0       # util/compute.py:141
    p.lat_vort_fac = float(namelist.lat_vort_fac)
    for i, bid in enumerate(layout.BASIN_IDS):
        p.lat_vort_power[i] = float(namelist.lat_vort_power[bid])
        p.atm_bl_depth[i] = float(namelist.atm_bl_depth[bid])
    p.seed_v_init = float(namelist.seed_v_init_ms)
    p.seed_v_2d_thresh = float(namelist.seed_v_2d_threshold_ms)
    p.seed_v_thresh = float(namelist.seed_v_threshold_ms)
    p.seed_vmax_thresh = float(namelist.seed_vmax_threshold_ms)
    p.pi_gen_min = 35.0                                        # util/compute.py:168
    p.minit_amp, p.minit_center, p.minit_slope, p.minit_offset = _check_minit(namelist)
    n = np.linspace(1, 15, 15)                                 # track/bam_track.py:26-29
    amp = np.sqrt(2 / np.sum(np.power(n, -3))) * np.power(n, -1.5)
    for i in range(15):
        p.fourier_amp[i] = float(amp[i])
    p.max_redraws = int(max_redraws)
    return p
