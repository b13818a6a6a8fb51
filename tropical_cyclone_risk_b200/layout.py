"""Channel layout of the HBM-resident monthly environment tables.

The caller uploads one set of N_CH float32 planes ``[N_CH][nlat][nlon]`` per (year, month) on the basin-cropped,
ascending ERA5-shaped grid; on the device they become CELL RECORDS ``rec[ym][iy][ix][20]`` of float4 -- for each grid
cell the four corner values of every channel (+1 pad), 320 B, one aligned read per bilinear look-up of all channels
(k_build_month).

Channel order follows the reference's own vector orders:
  * means  ``[ua250, va250, ua850, va850]``        (track/env_wind.py:22-26)
  * lower-triangular covariance, row-major         (track/env_wind.py:31-42)
  * chi, vpot, mld, strat                          (intensity/coupled_fast.py:217-225)
  * rh_mid (seeding only, util/compute.py:114,173)
"""

N_CH = 20            # 19 used + 1 pad -> 80 B / point = 5 x float4

CH_MEAN = 0          # 4 channels
CH_COV = 4           # 10 channels: c00 c10 c11 c20 c21 c22 c30 c31 c32 c33
CH_CHI = 14
CH_VPOT = 15
CH_MLD = 16
CH_STRAT = 17
CH_RH = 18
CH_PAD = 19

N_FIELDS = 19        # channels a caller supplies to tcr_upload_month

# outputs of tcr_env_interp: the 19 interpolated channels followed by
# bathymetry and land (sampled on their own static grids) -> 21 doubles
N_INTERP_OUT = 21
OUT_BATHY = 19
OUT_LAND = 20

FIELD_NAMES = (
    "ua250_Mean", "va250_Mean", "ua850_Mean", "va850_Mean",
    "ua250_Var",
    "va250_ua250_cov", "va250_Var",
    "ua850_ua250_cov", "ua850_va250_cov", "ua850_Var",
    "va850_ua250_cov", "va850_va250_cov", "va850_ua850_cov", "va850_Var",
    "chi", "vpot", "mld", "strat", "rh_mid",
)

BASIN_IDS = ("AU", "EP", "NA", "NI", "SI", "SP", "WP")   # sorted, util/compute.py:87


def cov_index(i, j):
    """Channel offset (from CH_COV) of covariance element (i, j), j <= i."""
    if j > i:
        i, j = j, i
    return i * (i + 1) // 2 + j
