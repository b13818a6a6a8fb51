"""Default configuration surface ("namelist").

The reference configures everything through module-level globals of a
``namelist.py`` that every module imports (reference namelist.py:8-120).  This
module keeps the same global *names* and default *values* for the globals the
track-generation hot path reads, so a user's own ``namelist.py`` can be passed
to :func:`tropical_cyclone_risk_b200.params.params_from_namelist` unchanged
(``import namelist; params_from_namelist(namelist, basin_id)``).

Only the hot-path subset is defaulted here; file-system and pre-processing
settings (``var_keys``, ``select_thermo`` ...) are out of scope.
"""
import math
import os

# --- where things go (reference namelist.py:9-17) ---------------------------
src_directory = os.path.dirname(os.path.abspath(__file__))
base_directory = os.path.join(os.getcwd(), "data", "synthetic")
output_directory = base_directory
exp_name = "test"
exp_prefix = "synthetic"
dataset_type = "ERA5"

# --- process-level parallelism of the reference; here: ranks == GPUs --------
n_procs = 16

# --- period and output sampling (reference namelist.py:40-50) ---------------
start_year, start_month = 2016, 1
end_year, end_month = 2021, 12
output_interval_s = 3600
total_track_time_days = 15
tracks_per_year = 20

# --- thermodynamic scaling (reference namelist.py:55-60) --------------------
p_midlevel = 60000                    # Pa, mid-level of the saturation deficit
PI_reduc = 0.80
Ck = 1.2e-3
Cd = 1.2e-3
select_thermo = 1                     # 1 pseudoadiabatic, 2 reversible (k_thermo<false> / k_thermo<true>)
select_interp = 2                     # 2 entropy look-up table (the only mode CAPE_PI_vectorized, thermo.py:266, has)

# --- track / intensity constants (reference namelist.py:70-94) --------------
steering_levels = [250, 850]
steering_coefs = [0.2, 0.8]
coupled_track = True
y_alpha = [0.17, 0.83]
m_alpha = [0.0025, -0.0025]
alpha_max = [0.41, 0.78]
alpha_min = [0.22, 0.59]
u_beta = -1.0
v_beta = 2.5
T_days = 20
seed_v_init_ms = 5
seed_v_2d_threshold_ms = 6.5
seed_v_threshold_ms = 15
seed_vmax_threshold_ms = 18
atm_bl_depth = dict(NA=1400.0, EP=1400.0, WP=1800.0, AU=1800.0,
                    SI=1600.0, SP=2000.0, NI=1500.0)
log_chi_fac = 0.5
chi_fac = 1.3
lat_vort_fac = 2
lat_vort_power = dict(NA=6, EP=6, WP=3.5, AU=6, SI=3, SP=7, NI=2.5)


def f_mInit(rh):
    """Initial inner-core moisture from mid-level RH (reference namelist.py:94)."""
    return 0.20 / (1 + math.exp(-(rh - 0.55) * 10)) + 0.125


# --- basin boxes: [LL lon, LL lat, UR lon, UR lat] (reference namelist.py:112-119)
basin_bounds = dict(
    EP=["180E", "0N", "290E", "60N"],
    NA=["260E", "0N", "360E", "60N"],
    NI=["30E", "0N", "100E", "50N"],
    SI=["20E", "45S", "100E", "0S"],
    AU=["100E", "45S", "180E", "0S"],
    SP=["180E", "45S", "250E", "0S"],
    WP=["100E", "0N", "180E", "60N"],
    GL=["0E", "90S", "360E", "90N"],
)
