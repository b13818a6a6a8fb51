/*
 * tcrisk.h -- C ABI of the B200-native synthetic tropical-cyclone ensemble
 * integrator (libtcrisk.so).
 *
 * The reference (linjonathan/tropical_cyclone_risk) has no FFI: its seam for the
 * hot path is Python-level:
 *
 *   outer:  run_tracks(year, n_tracks, b) -> 9-tuple          util/compute.py:64,210
 *   inner:  Coupled_FAST(...).init_fields / .gen_track        intensity/coupled_fast.py:19,217,229
 *           BetaAdvectionTrack._env_winds                     track/bam_track.py:116
 *           RectBivariateSpline(kx=1,ky=1).ev                 util/mat.py:142-153
 *
 * Every entry point below names the reference interface it replaces.  Plain C,
 * caller-allocated buffers, int status return (0 = ok, <0 = error; text via
 * tcr_last_error()).  No exceptions, no torch types.  One handle per device;
 * calls on one handle must be serialised by the caller.  All kernels and copies
 * are issued on the handle's stream (tcr_set_stream; stream 0 until set -- prefer a dedicated
 * non-blocking stream: the legacy default stream serialises against NCCL's).
 *
 * Pointer arguments named h_* are HOST pointers; d_* are DEVICE pointers; the
 * remaining data pointers are host pointers unless `on_device` says otherwise.
 */
#ifndef TCRISK_H
#define TCRISK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TCR_N_CH          20   /* channels per grid point in HBM (19 used + pad)        */
#define TCR_N_FIELDS      19   /* fields supplied per month (layout.py FIELD_NAMES)     */
#define TCR_N_INTERP_OUT  21   /* tcr_env_interp outputs: 19 fields, bathymetry, land   */
#define TCR_N_BASINS       7   /* sorted basin ids AU EP NA NI SI SP WP (compute.py:87) */
#define TCR_N_MASKS        8   /* 7 basin masks + the run basin's own mask (f_b)        */
#define TCR_N_SERIES       4   /* Fourier series per storm = nWLvl (bam_track.py:60)    */
#define TCR_N_HARM        15   /* harmonics per series (bam_track.py:112)               */
#define TCR_N_PHASES      60

/* storm status codes (scipy solve_ivp status + the reference's `None`) */
#define TCR_STATUS_FINISHED   0   /* reached total_time                                  */
#define TCR_STATUS_EVENT      1   /* terminal event tc_dissipates (coupled_fast.py:246)  */
#define TCR_STATUS_FAILED    -1   /* step size underflow (scipy rk.py _step_impl)        */
#define TCR_STATUS_VENT       2   /* gen_track returned None (coupled_fast.py:241-244)   */

/* Safety net absent from the reference: scipy's RK45 loops forever when the step size
 * becomes NaN (a NaN genesis point) and crawls at 10 ulp(t) per step if error control
 * collapses.  A storm whose step size is NaN, or that needs more than this many RK attempts,
 * ends with TCR_STATUS_FAILED.  (Storms on real fields take 10-40 attempts.)                 */
#define TCR_MAX_RK_ATTEMPTS 20000

/* storm flag bits */
#define TCR_FLAG_IS_TC   1u   /* util/compute.py:185-189 */
#define TCR_FLAG_KEPT    2u   /* util/compute.py:205     */

/* All physics / configuration constants the hot path reads from `namelist`
 * (reference namelist.py:40-119) plus the fixed constants of coupled_fast.py:23-27,
 * bam_track.py:52-60, constants.py:7.  Filled by params_from_namelist() in Python. */
typedef struct tcr_params {
    double dt_track;            /* namelist.output_interval_s               (bam_track.py:52)  */
    double total_time;          /* total_track_time_days*86400              (bam_track.py:53)  */
    double T_Fs;                /* namelist.T_days*86400                    (bam_track.py:56)  */
    double max_step;            /* 86400                                    (coupled_fast.py:266) */
    double rtol, atol;          /* 1e-3, 1e-6 (scipy solve_ivp defaults)                        */
    double u_beta, v_beta;      /* namelist.py:77-78                                            */
    double steering_coefs[2];   /* used when coupled_track == 0             (coupled_fast.py:191) */
    double y_alpha[2], m_alpha[2], alpha_max[2], alpha_min[2];      /* namelist.py:73-76       */
    double Ck;                  /* namelist.py:57                                               */
    double epsilon, kappa, beta;/* 0.33, 0.1, 1-0.33-0.1                    (coupled_fast.py:25-27) */
    double earth_R;             /* 6.3781e6                                 (constants.py:7)    */
    double basin_bounds[4];     /* lon_min, lat_min, lon_max, lat_max       (basins.py:42-50)   */
    double gen_lat_min, gen_lat_max;   /* 3|-45, 45|-3                      (compute.py:140-141) */
    double lat_vort_fac;        /* namelist.py:89                                               */
    double lat_vort_power[TCR_N_BASINS];   /* namelist.py:90-92, sorted basin order             */
    double atm_bl_depth[TCR_N_BASINS];     /* namelist.py:85-86, sorted basin order             */
    double seed_v_init;         /* namelist.py:80                                               */
    double seed_v_2d_thresh;    /* namelist.py:81                                               */
    double seed_v_thresh;       /* namelist.py:82                                               */
    double seed_vmax_thresh;    /* namelist.py:83                                               */
    double pi_gen_min;          /* 35 m/s                                   (compute.py:168)    */
    double minit_amp, minit_center, minit_slope, minit_offset;  /* f_mInit  (namelist.py:94)    */
    double fourier_amp[TCR_N_HARM];   /* sqrt(2/sum n^-3) * n^-1.5, n=1..15  (bam_track.py:28-29) */
    int32_t n_steps;            /* int(total_time/dt_track)+1               (bam_track.py:54)   */
    int32_t coupled_track;      /* namelist.py:72                                               */
    int32_t max_redraws;        /* bound on the ocean-point redraw loop     (compute.py:146-148) */
    int32_t reserved;
} tcr_params;

/* Work counters of one tcr_run_years call (all per year; arrays of n_years). */
typedef struct tcr_year_stats {
    int64_t attempts;          /* seed attempts consumed, i.e. i*+1 (compute.py:136 iterations)  */
    int64_t counted_seeds;     /* sum of n_seeds                                                  */
    int64_t integrated;        /* gen_track calls with attempt index <= i*                       */
    int64_t storm_steps;       /* emitted samples of those calls (sum len(res.t))                */
    int64_t kept_steps;        /* samples of the n_tracks kept storms                            */
    int64_t rhs_evals;         /* dydt evaluations of those calls                                */
    int64_t wasted_integrated; /* gen_track calls beyond i* (over-shoot of the last wave)        */
    int64_t wasted_steps;      /* their samples                                                  */
    int64_t wasted_rhs_evals;  /* their dydt evaluations                                         */
    int32_t n_kept;            /* == n_tracks on success                                         */
    int32_t n_waves;
    int64_t redraw_exhausted;  /* attempts whose ocean-point redraw chain (compute.py:146-148, unbounded in the reference)
                                  hit tcr_params.max_redraws and were dropped: expected 0 (max_redraws = 64)          */
} tcr_year_stats;

typedef struct tcr_handle tcr_handle;

/* ---- life cycle ----------------------------------------------------------------------- */
/* replaces: Coupled_FAST.__init__ constants + namelist reads (coupled_fast.py:19-32)       */
int tcr_create(int device, const tcr_params* p, tcr_handle** out);
int tcr_destroy(tcr_handle* h);
const char* tcr_last_error(void);
/* issue all subsequent work of `h` on this cudaStream_t (pass torch's current stream)     */
int tcr_set_stream(tcr_handle* h, void* cuda_stream);
int tcr_synchronize(tcr_handle* h);
int tcr_version(void);

/* ---- static fields -------------------------------------------------------------------- */
/* replaces: geo.read_bathy / geo.read_land (intensity/geo.py:9-33): basin-cropped,
 * ascending axes; int16 bathymetry [nlat_b][nlon_b], int8 land [nlat_l][nlon_l].          */
int tcr_upload_static(tcr_handle* h,
                      int nlat_b, int nlon_b, const double* lat_b, const double* lon_b, const int16_t* bathy,
                      int nlat_l, int nlon_l, const double* lat_l, const double* lon_l, const int8_t* land);
/* replaces: f_basins[...] / f_b = mat.interp2_fx(land/<id>.nc) (util/compute.py:87-97):
 * masks uint8 [TCR_N_MASKS][nlat_m][nlon_m]; planes 0..6 sorted basin ids, plane 7 = run basin */
int tcr_upload_masks(tcr_handle* h, int nlat_m, int nlon_m, const double* lat_m, const double* lon_m,
                     const uint8_t* masks);

/* ---- monthly environment tables ------------------------------------------------------- */
/* replaces: BetaAdvectionTrack._load_wnd_stat + Coupled_FAST.init_fields for one month
 * (track/bam_track.py:76-91, intensity/coupled_fast.py:217-225).  All fields share one
 * basin-cropped ascending grid.  tcr_alloc_tables sizes the HBM arena for n_ym months
 * (index ym = year_slot*12 + month-1); tcr_upload_month interleaves 19 float32 planes
 * [nlat][nlon] into the [nlat][nlon][20] record layout on the device.                      */
int tcr_alloc_tables(tcr_handle* h, int n_ym, int nlat, int nlon, const double* lat, const double* lon);
int tcr_upload_month(tcr_handle* h, int ym, const float* const* fields /*[TCR_N_FIELDS]*/);
/* n_months consecutive months from one contiguous host block [n_months][19][nlat][nlon]: one
 * host->device copy and one table-building launch (a year, or all years of a batch, at once) */
int tcr_upload_months(tcr_handle* h, int ym0, int n_months, const float* planes);
/* same, source planes already in HBM as one contiguous [19][nlat][nlon] float32 block       */
int tcr_upload_month_dev(tcr_handle* h, int ym, const float* d_planes);

/* ---- per-month field preparation on the device (SURVEY 8f "next" row N1) ---------------- */
/* replaces: the per-month body of run_tracks' setup loop (util/compute.py:107-121: NaN policy,
 * PI scaling :76, chi transform :113-115, ocean regrid :117-118 via util/mat.py:159-164) and
 * BetaAdvectionTrack._interp_basin_field (track/bam_track.py:72-74: basin crop of
 * util/basins.py:57-75 + nan_to_num).  raw [17][nlat_g][nlon_g] float32 = the 14 wind statistics
 * (channel order of the tables), vmax, chi, rh_mid on the GLOBAL grid (any row order: src_row maps
 * output rows to storage rows, lat_g is in storage order); ocean [2][nlat_o][nlon_o] = mld, strat on
 * their own ascending grid; src_col / src_row = the basin crop (with longitude re-wrapping) as
 * index maps.  Writes month `ym` of the tables (ym < 0: do not upload) and, if planes_out is
 * non-NULL, returns the 19 prepared float32 planes [19][nlat_b][nlon_b] to the host.          */
typedef struct tcr_prep_spec {
    int32_t nlat_g, nlon_g, nlat_o, nlon_o, nlat_b, nlon_b;
    double pi_reduc, sqrt_ck_cd;  /* vpot = vmax * PI_reduc * sqrt(Ck/Cd), left to right (compute.py:76) */
    double log_chi_fac, chi_fac;  /* namelist.py                          (compute.py:115) */
} tcr_prep_spec;
int tcr_prepare_month(tcr_handle* h, int ym, const tcr_prep_spec* spec, const float* raw, const float* ocean,
                      const double* lon_g, const double* lat_g, const double* lon_o, const double* lat_o,
                      const int32_t* src_col, const int32_t* src_row, float* planes_out);

/* ---- stand-alone bilinear sampler (roofline kernel) ----------------------------------- */
/* replaces: RectBivariateSpline(kx=1,ky=1).ev on every field (util/mat.py:142-153,
 * bam_track.py:93-108, coupled_fast.py:35-58,125-126).  out[n][21] float64.                */
int tcr_env_interp(tcr_handle* h, int64_t n, const int32_t* ym, const double* lon, const double* lat,
                   double* out, int on_device);

/* ---- integrate given seeds ------------------------------------------------------------ */
/* replaces: Coupled_FAST.gen_track (coupled_fast.py:229-267) + the per-candidate post-
 * processing of run_tracks (compute.py:178-206: TC criteria, env-wind recompute, vmax).
 * phases[n][4][15] are the uniform random phases of gen_f (bam_track.py:27).
 * Outputs (any may be NULL): track [n][n_steps][4] = lon,lat,v,m ; env [n][n_steps][4];
 * vmax [n][n_steps] (NaN padded past n_time); n_time, status, nfev, flags [n].             */
int tcr_integrate(tcr_handle* h, int64_t n,
                  const int32_t* ym, const double* lon0, const double* lat0,
                  const double* v0, const double* m0, const double* h_bl, const double* phases,
                  double* track, double* env, double* vmax,
                  int32_t* n_time, int32_t* status, int32_t* nfev, uint32_t* flags,
                  int on_device);

/* ---- single evaluations of the right-hand side (inner tier of the seam) ------------------ */
/* replaces: Coupled_FAST.dydt(t, y) (intensity/coupled_fast.py:196-207) and BetaAdvectionTrack._env_winds(lon, lat,
 * t) (track/bam_track.py:116-128; called per output point by util/compute.py:201-202) for n independent states:
 * y [n][4] = lon, lat, v, m; phases [n][4][15] the storm's gen_f phases (self.Fs); dydt [n][4]; env_winds [n][4] =
 * _env_winds(y[0], y[1], t) (zeros on a NaN argument or a non-positive-definite covariance).  Host pointers.      */
int tcr_rhs_eval(tcr_handle* h, int64_t n, const int32_t* ym, const double* t, const double* y, const double* h_bl,
                 const double* phases, double* dydt, double* env_winds);

/* ---- whole years: seeding + integration + ordered selection ---------------------------- */
/* replaces: run_tracks(year, n_tracks, b) for n_years years at once (util/compute.py:64-210).
 * Year y uses tables ym = ym_base[y] .. ym_base[y]+11 and Philox key (run_seed, year_key[y]).
 * Seed attempts are indexed 0,1,2,... per year; rank r of `world` integrates the attempts it
 * owns and the first n_tracks kept storms in attempt order are returned.
 * Outputs per year y (row-major, year-major): lon/lat/v/m/vmax [n_years][n_tracks][n_steps],
 * env [n_years][n_tracks][n_steps][4], tc_month [n_years][n_tracks] (double, 1..12),
 * tc_basin [n_years][n_tracks] (index into sorted ids), n_seeds [n_years][7][12] (double).
 * Data outputs are host or device pointers per `on_device`; stats is always a host pointer. */
int tcr_run_years(tcr_handle* h, int n_years, const int32_t* ym_base, const int32_t* year_key,
                  uint32_t run_seed, int n_tracks,
                  double* lon, double* lat, double* v, double* m, double* vmax, double* env,
                  double* tc_month, int32_t* tc_basin, double* n_seeds,
                  tcr_year_stats* stats, int on_device);

/* ---- within-year sharding (partition mode 2) --------------------------------------- */
/* replaces: nothing in the reference (it gives a year to ONE dask process, util/compute.py:224-230); needed when one
 * (basin, year) is larger than a GPU should take (BASELINE configs[4]: 50 000 tracks / year).  After tcr_set_shard(h,
 * rank, world, fn, user) every tcr_run_years call on `h` is COLLECTIVE over `world` handles (one per GPU, all called
 * with the same arguments): every rank seeds every attempt (cheap), integrates the attempts k with k % world == rank,
 * and once per wave the ranks exchange, through `fn`, the per-attempt kept flags and the counted-seed histograms -- so
 * that all of them select the same first n_tracks survivors in attempt order: the result does not depend on world.
 * Each rank then holds the rows whose storms it integrated; every other row of its lon / lat / v / m / vmax / env
 * arrays is all-zero bits (merge = integer sum over ranks); tc_month, tc_basin, n_seeds are complete on every rank.
 * Of tcr_year_stats, attempts / counted_seeds / n_kept / n_waves are global, the other counters this rank's share.
 * fn(user, d_buf, count, dtype, op, cuda_stream): in-place all-reduce of `count` elements at device pointer d_buf over
 * the group, stream-ordered on `cuda_stream`; dtype 0 = uint8, 1 = int32, 2 = int64; op 0 = sum, 1 = min; 0 on success. */
typedef int (*tcr_allreduce_fn)(void* user, void* d_buf, int64_t count, int dtype, int op, void* cuda_stream);
int tcr_set_shard(tcr_handle* h, int rank, int world, tcr_allreduce_fn fn, void* user);

/* seeding only: evaluate attempts [k0, k0+n) of one year -> per-attempt records (test hook
 * for util/compute.py:136-175).  code: 0 = not a seed, 1 = counted (PI <= 35), 2 = passed,
 * 3 = redraw bound hit.  Host pointers.                                                     */
int tcr_seed_attempts(tcr_handle* h, int ym_base, int32_t year_key, uint32_t run_seed,
                      int64_t k0, int64_t n,
                      int32_t* code, int32_t* basin, int32_t* month,
                      double* lon, double* lat, double* v0, double* m0, double* pi_gen);

/* tuning knobs (0 keeps the default): variant of the integrate kernel = threads x CTAs/SM
 * [K: stage derivatives in shared memory] [L: CTA-lockstep RHS evaluations]:
 * 1 = 256x1, 2 = 128x3, 3 = 128x4 K, 4 = 160x2, 5 = 128x3 K, 6 = 192x2 K, 7 = 192x2 K L,
 * 8 = 256x1 L, 9 = 288x1 K L, 10 = 384x1 K L, 11 = 512x1 K L, 12 = 128x3 K L, 13 = 224x1 L,
 * 14-16 = 192x2 K with lockstep re-alignment at slots {0,2,4} / {0,3} / {0}, 17 = 256x2 KK L,
 * 18 = 192x2 KK L (default: every stage vector in shared memory, 168 registers, no spills),
 * 19 = 192x3 KK L, 20 = 224x2 KK L, 21 = 384x1 KK L  (KK: K0..K6 and y_new in shared memory);
 * upper bounds on the seed attempts and on the integrated storms of one wave; wave
 * over-subscription factor (x1000).  Results never depend on these (ordered selection, bit-exact
 * kernels); only speed and the amount of discarded work do.                                  */
int tcr_set_tuning(tcr_handle* h, int integ_variant, int64_t max_wave_cands, int64_t max_wave_slots,
                   int oversub_permille);
/* How tcr_run_years tabulates the storms' Fourier series for this handle's output grid: 0 = full tables ahead of the
 * integrator (k_fourier_table_mma), n > 0 = rings of n nodes per track-pool row that k_integrate fills on demand
 * (the default where the grid is at least four rings long, i.e. on 900-s output; TCR_FTAB_RING=0 / 1 at tcr_create
 * forces either).  Results never depend on it.                                                                   */
int tcr_fourier_ring_nodes(tcr_handle* h);
/* memory the per-wave workspace of tcr_run_years may take: a fraction of the device memory that is free when the
 * workspace is sized (default 0.6) and an absolute cap (default 96 GiB).  A larger budget means fewer, larger waves
 * (configs[2]: 9 waves at 0.6, 7 at 0.8: +1 %); what it leaves must hold whatever the caller allocates afterwards.
 * Results never depend on it.                                                                                    */
int tcr_set_workspace_budget(tcr_handle* h, double fraction_of_free, int64_t cap_bytes);
/* number of kernels launched by this handle so far (bench.py's gpu_launches)               */
int64_t tcr_launch_count(tcr_handle* h);
/* device-time accounting: with timing enabled every launch of a kernel class is bracketed by
 * CUDA events on the handle's stream; tcr_kernel_time synchronises the stream and returns the
 * accumulated milliseconds and launch count of one class since tcr_set_timing was last called
 * (the reference prints time.time() deltas instead, util/compute.py:26-35,229,270)            */
#define TCR_K_ENV_INTERP   0
#define TCR_K_INTEGRATE    1
#define TCR_K_POSTPROCESS  2
#define TCR_K_SEED         3
#define TCR_K_COEF         4
#define TCR_K_SELECT       5
#define TCR_K_GATHER       6
#define TCR_K_BUILD        7
#define TCR_K_FTABLE       8
#define TCR_K_POI          9
#define TCR_K_WINDSTAT     10
#define TCR_K_THERMO       11
#define TCR_N_KERNEL_CLASSES 12
int tcr_set_timing(tcr_handle* h, int enable);
int tcr_kernel_time(tcr_handle* h, int kernel_class, double* ms, int64_t* launches);
/* tcr_env_interp implementation: 0 = per-lane LDG.128 gathers, 1 = TMA bulk copies
 * (cp.async.bulk) of the cell records into shared memory behind an mbarrier pipeline        */
int tcr_set_interp_variant(tcr_handle* h, int variant);
/* ---- return-period reduction over finished tracks (SURVEY 8f "next" row N4) ------------- */
/* replaces: notebooks/sample_analysis.ipynb cell 15 -- dists = haversine(clon, clat, lon_trks,
 * lat_trks); vmax_at_poi = vmax_trks.where(dists <= radius_km).max(dim='time').  lon/lat/vmax
 * [n_rows][n_steps] (NaN padded), out [n_rows] (NaN where the track never comes within the
 * radius).  r_earth_m = 6378000 in the notebook.  Host or device pointers per on_device.      */
int tcr_poi_vmax(tcr_handle* h, int64_t n_rows, int n_steps, const double* lon, const double* lat, const double* vmax,
                 double poi_lon, double poi_lat, double radius_km, double r_earth_m, double* out, int on_device);
/* replaces: cell 17 -- exceedance_count[b] = sum(vmax_at_poi >= vmax_bins[b]); n_bins <= 64;
 * v host or device per on_device, bins and counts host                                        */
int tcr_exceedance(tcr_handle* h, int64_t n, const double* v, int n_bins, const double* bins, int64_t* counts, int on_device);

/* ---- monthly wind mean / covariance reduction (SURVEY 8f "next" row N3, first half) ------ */
/* replaces: calc_wnd_stat (track/env_wind.py:169-228) for one month: the samples of the month
 * (ua, va at the upper = 250 hPa and lower = 850 hPa steering level, float32, sample t of grid
 * point p at x[t * t_stride + p] -- so a (time, level, lat, lon) array is passed by level-slice
 * pointers with t_stride = n_level * n_pts, no copy) are averaged per day
 * (groupby("time.day").mean, :188-190; day g = samples [group_start[g], group_start[g+1]); daily
 * or coarser data: one sample per group), then reduced to the 14 statistics in the reference's
 * order (:206-216): ua250, va250, ua850, va850 means; then the lower triangle row by row with
 * .var (ddof 0, :211) on the diagonal and xr.cov (ddof 1, :213) off it.  NaNs are skipped the way
 * xarray's skipna reductions do.  out [14][n_pts] float64.  Pointers host or device per on_device. */
int tcr_wind_stats(tcr_handle* h, int n_time, int64_t n_pts, int64_t t_stride,
                   const float* ua_upper, const float* va_upper, const float* ua_lower, const float* va_lower,
                   int n_groups, const int32_t* group_start /* host, [n_groups + 1] */, double* out, int on_device);

/* ---- potential intensity, saturation deficit, mid-level humidity (second half) ---- */
/* replaces: one time sample of compute_thermo (thermo/calc_thermo.py:60-69) --
 *   vmax = thermo.CAPE_PI_vectorized(sst, psl, lvl, ta, hus)              (thermo/thermo.py:266-412)
 *   chi = clip(thermo.sat_deficit(sst, psl, ta_mid, p_mid, hus_mid), 0, 10)   (thermo.py:92-104)
 *   rh_mid = thermo.conv_q_to_rh(ta_mid, hus_mid, p_mid)                      (thermo.py:42-47)
 * for namelist.select_thermo = 1, select_interp = 2.  tcr_set_entropy_table uploads the look-up table
 * the reference loads from thermo/entropy_table.npz (p [np] Pa ascending, s [ns] ascending, T [np][ns];
 * thermo.py:274-278, 317).  ta, hus [nlev][n_pts] float32 with the lowest model level first
 * (calc_thermo.py:50-55), p_env [nlev] in Pa (host), sst [n_pts] K already on the atmospheric grid
 * (calc_thermo.py:38-42), psl [n_pts] Pa, k_mid = the level nearest namelist.p_midlevel; outputs
 * float64 [n_pts].  Field pointers host or device per on_device.                                  */
int tcr_set_entropy_table(tcr_handle* h, int np, int ns, const double* p_look, const double* s_look, const double* T_lookup);
int tcr_thermo_month(tcr_handle* h, int64_t n_pts, int nlev, const double* p_env, const float* ta, const float* hus,
                     const double* sst, const double* psl, double ck_over_cd, int k_mid,
                     double* vmax, double* chi, double* rh_mid, int on_device);
/* the same three fields for namelist.select_thermo = 2 (reversible thermodynamics: thermo.py:56-60, 71-75, 132-133):
 * the inversion table is the three-dimensional one of thermo/entropy_table_reversible.npz (p [np] Pa, s [ns],
 * rt [nrt] total water, T [np][ns][nrt]; thermo.py:279-284), read the way scipy.interpolate.interpn(method='linear',
 * bounds_error=False, fill_value=nan) reads it (thermo.py:343-353): a level, entropy or water content outside its axis
 * gives a NaN parcel temperature.  select_interp = 1 (BFGS inversion) exists only in the reference's unused scalar
 * CAPE_PI (thermo.py:144-263); CAPE_PI_vectorized, the function calc_thermo.py:61 calls, has no such branch.        */
int tcr_set_entropy_table_reversible(tcr_handle* h, int np, int ns, int nrt, const double* p_look, const double* s_look,
                                     const double* rt_look, const double* T_lookup);
int tcr_thermo_month_reversible(tcr_handle* h, int64_t n_pts, int nlev, const double* p_env, const float* ta, const float* hus,
                                const double* sst, const double* psl, double ck_over_cd, int k_mid,
                                double* vmax, double* chi, double* rh_mid, int on_device);

/* page-locked host memory for the caller's input planes / result arrays (the reference's
 * NumPy arrays of util/compute.py:126-133 become views of this block): makes the host<->device
 * copies of tcr_upload_month / tcr_run_years run at PCIe speed                               */
int tcr_host_alloc(size_t bytes, void** out);
int tcr_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* TCRISK_H */
