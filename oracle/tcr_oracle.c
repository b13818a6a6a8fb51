/*
 * tcr_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C float64 restatement of the reference's per-storm hot path
 * (linjonathan/tropical_cyclone_risk @ 5268fdb).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the CUDA product path never does.
 *
 * Parity status: the reference ships no tests or golden vectors for this path
 * ("parity unpinned" by the reference itself).  This restatement is pinned instead
 * against outputs of the reference's own modules run in the build container
 * (oracle/make_golden.py -> tests/golden/ ; tests/test_oracle_vs_golden.py).
 *
 * Third-party arithmetic restated here (un-vendored in the reference, versions as
 * installed in the build container: SciPy 1.18.1, NumPy 2.3.5):
 *   scipy.integrate.solve_ivp / RK45   (call site intensity/coupled_fast.py:264)
 *   scipy RectBivariateSpline(kx=ky=1).ev -> FITPACK bispev/fpbisp/fpbspl
 *   scipy.interpolate.interp1d (linear) (call site intensity/coupled_fast.py:235)
 *   numpy.linalg.cholesky -> LAPACK dpotrf (call site track/bam_track.py:123)
 *
 * Arithmetic contract.  The reference's adaptive integration is chaotic at the rounding
 * level (DESIGN.md "Chaos floor"), so this file is also the BIT-LEVEL SPECIFICATION the CUDA
 * path is tested against: every floating-point operation below is a correctly rounded IEEE
 * binary64 add/mul/div/sqrt or an explicit fma(), transcendentals come from
 * include/tcr_libm.h, and the build disables contraction (gcc -ffp-contract=off).  Where
 * that differs from the op order of the reference's BLAS/libm (fused sums, cos evaluated
 * once, Fourier series by angle addition, m*m*m for m**3), the difference is a few ulp per
 * RHS and is below the reference's own reproducibility across NumPy/SciPy builds.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/tcrisk.h"
#include "../include/tcr_libm.h"

/* ------------------------------------------------------------------------------------ */
/* fields                                                                                */
/* ------------------------------------------------------------------------------------ */
typedef struct orc_grid {
    int nx, ny;              /* nx = lon points, ny = lat points */
    const double* x;         /* lon axis, ascending */
    const double* y;         /* lat axis, ascending */
} orc_grid;

typedef struct orc_env {
    orc_grid g;                              /* monthly table grid                           */
    const double* f[TCR_N_FIELDS];           /* each [ny][nx] (lat-major), float64           */
    orc_grid gb; const double* bathy;        /* [ny][nx]                                     */
    orc_grid gl; const double* land;         /* [ny][nx]                                     */
} orc_env;

/* channel indices, same as tropical_cyclone_risk_b200/layout.py */
enum { CH_MEAN = 0, CH_COV = 4, CH_CHI = 14, CH_VPOT = 15, CH_MLD = 16, CH_STRAT = 17, CH_RH = 18 };

/* FITPACK fpbisp interval search + fpbspl weights for degree 1 with clamped argument
 * (scipy RectBivariateSpline.ev; built at util/mat.py:152, coupled_fast.py:219-225,
 * geo.py:19,33).  Interval is left-closed; the last interval is closed on the right. */
static void orc_locate(const double* ax, int n, double arg, int* i0, double* w0, double* w1)
{
    double a = arg;
    if (a < ax[0]) a = ax[0];
    if (a > ax[n - 1]) a = ax[n - 1];
    int lo = 0, hi = n - 1;               /* largest i with ax[i] <= a, capped at n-2 */
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (ax[mid] <= a) lo = mid; else hi = mid;
    }
    if (lo > n - 2) lo = n - 2;
    double f = 1.0 / (ax[lo + 1] - ax[lo]);
    *i0 = lo;
    *w0 = f * (ax[lo + 1] - a);
    *w1 = f * (a - ax[lo]);
}

/* FITPACK's own summation: sp = sum_{i1} sum_{j1} c * wx(i1) * wy(j1), x (lon) outer,
 * y (lat) inner, left-to-right unfused products -- bit-identical to scipy 1.18.1's .ev
 * (verified on 20 000 random points).  Used where the reference tests a sampled value
 * for exact equality (land == 1), for the genesis masks and for the public orc_bilinear. */
static double orc_bilin_fitpack(const double* d, int nx, int ix, int iy,
                                double wx0, double wx1, double wy0, double wy1)
{
    const double* r0 = d + (size_t)iy * nx + ix;
    const double* r1 = r0 + nx;
    double sp = 0.0;
    sp = sp + r0[0] * wx0 * wy0;
    sp = sp + r1[0] * wx0 * wy1;
    sp = sp + r0[1] * wx1 * wy0;
    sp = sp + r1[1] * wx1 * wy1;
    return sp;
}

/* spec form used for every field inside the integrator: four weight products, fused sum */
static double orc_bilin_fused(const double* d, int nx, int ix, int iy,
                              double w00, double w01, double w10, double w11)
{
    const double* r0 = d + (size_t)iy * nx + ix;
    const double* r1 = r0 + nx;
    return fma(r1[1], w11, fma(r0[1], w10, fma(r1[0], w01, r0[0] * w00)));
}

double orc_bilinear(const double* xax, int nx, const double* yax, int ny, const double* d,
                    double x, double y)
{
    int ix, iy; double wx0, wx1, wy0, wy1;
    orc_locate(xax, nx, x, &ix, &wx0, &wx1);
    orc_locate(yax, ny, y, &iy, &wy0, &wy1);
    return orc_bilin_fitpack(d, nx, ix, iy, wx0, wx1, wy0, wy1);
}

/* one located point on a grid: cell index + the four weight products
 * (w01 = lon corner 0, lat corner 1) */
typedef struct orc_cell { int ix, iy; double w00, w01, w10, w11, wx0, wx1, wy0, wy1; } orc_cell;

static void orc_cell_at(const orc_grid* g, double lon, double lat, orc_cell* c)
{
    orc_locate(g->x, g->nx, lon, &c->ix, &c->wx0, &c->wx1);
    orc_locate(g->y, g->ny, lat, &c->iy, &c->wy0, &c->wy1);
    c->w00 = c->wx0 * c->wy0; c->w01 = c->wx0 * c->wy1;
    c->w10 = c->wx1 * c->wy0; c->w11 = c->wx1 * c->wy1;
}

static double orc_field(const orc_env* e, const orc_cell* c, int ch)
{
    return orc_bilin_fused(e->f[ch], e->g.nx, c->ix, c->iy, c->w00, c->w01, c->w10, c->w11);
}

/* f_land.ev (exact FITPACK order) + whether all four corners are land: the value is then 1
 * or 1-ulp and the reference's `== 1` test (coupled_fast.py:38) is rounding noise */
static double orc_land(const orc_env* e, double lon, double lat, int* all_land)
{
    orc_cell c;
    orc_cell_at(&e->gl, lon, lat, &c);
    const double* r0 = e->land + (size_t)c.iy * e->gl.nx + c.ix;
    const double* r1 = r0 + e->gl.nx;
    *all_land = (r0[0] == 1.0 && r0[1] == 1.0 && r1[0] == 1.0 && r1[1] == 1.0);
    return orc_bilin_fitpack(e->land, e->gl.nx, c.ix, c.iy, c.wx0, c.wx1, c.wy0, c.wy1);
}

static double orc_bathy(const orc_env* e, double lon, double lat)
{
    orc_cell c;
    orc_cell_at(&e->gb, lon, lat, &c);
    return orc_bilin_fused(e->bathy, e->gb.nx, c.ix, c.iy, c.w00, c.w01, c.w10, c.w11);
}

/* all 19 monthly fields + bathymetry + land at n points: the oracle of tcr_env_interp */
void orc_env_interp(const orc_env* envs, int64_t n, const int32_t* ym, const double* lon,
                    const double* lat, double* out)
{
    for (int64_t q = 0; q < n; ++q) {
        const orc_env* e = envs + ym[q];
        double* o = out + q * TCR_N_INTERP_OUT;
        orc_cell c; int al;
        orc_cell_at(&e->g, lon[q], lat[q], &c);
        for (int ch = 0; ch < TCR_N_FIELDS; ++ch) o[ch] = orc_field(e, &c, ch);
        o[19] = orc_bathy(e, lon[q], lat[q]);
        o[20] = orc_land(e, lon[q], lat[q], &al);
    }
}

/* ------------------------------------------------------------------------------------ */
/* random-phase Fourier series: gen_f, track/bam_track.py:23-31                          */
/* ------------------------------------------------------------------------------------ */
/* reference-faithful table (direct sin per harmonic, glibc libm): only used to pin the
 * angle-addition form below against the reference, never inside the integrator */
void orc_gen_f_direct(const double* phases /*[4][15]*/, double T, const double* t_s, int n, double* fs /*[4][n]*/)
{
    double norm = 0.0;
    for (int k = 1; k <= TCR_N_HARM; ++k) norm += pow((double)k, -3.0);
    double amp = sqrt(2.0 / norm);
    for (int i = 0; i < TCR_N_SERIES; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0.0;
            for (int k = 1; k <= TCR_N_HARM; ++k) {
                double arg = 2.0 * TCR_PI * (((double)k * t_s[j]) / T + phases[i * TCR_N_HARM + k - 1]);
                s += pow((double)k, -1.5) * sin(arg);
            }
            fs[(size_t)i * n + j] = amp * s;
        }
}

/* spec form: a_n sin(n th + 2 pi x) = A sin(n th) + B cos(n th), A = a_n cos(2 pi x),
 * B = a_n sin(2 pi x); coef[i][n] = {A, B} */
void orc_fourier_coef(const tcr_params* p, const double* phases /*[4][15]*/, double* coef /*[4][15][2]*/)
{
    for (int i = 0; i < TCR_N_SERIES; ++i)
        for (int k = 0; k < TCR_N_HARM; ++k) {
            double s, c;
            tcr_sincos2pi(phases[i * TCR_N_HARM + k], &s, &c);
            coef[(i * TCR_N_HARM + k) * 2 + 0] = p->fourier_amp[k] * c;
            coef[(i * TCR_N_HARM + k) * 2 + 1] = p->fourier_amp[k] * s;
        }
}

/* F_i at table node time tn: harmonics by complex rotation from (sin th, cos th) */
static void orc_fourier_node(const tcr_params* p, const double* coef, double tn, double F[4])
{
    double s1, c1;
    tcr_sincos2pi(tn / p->T_Fs, &s1, &c1);
    double sn = s1, cn = c1;
    F[0] = F[1] = F[2] = F[3] = 0.0;
    for (int k = 0; k < TCR_N_HARM; ++k) {
        for (int i = 0; i < TCR_N_SERIES; ++i) {
            const double* ab = coef + (i * TCR_N_HARM + k) * 2;
            F[i] = fma(ab[0], sn, F[i]);
            F[i] = fma(ab[1], cn, F[i]);
        }
        double sn1 = fma(sn, c1, cn * s1);
        double cn1 = fma(cn, c1, -(sn * s1));
        sn = sn1; cn = cn1;
    }
}

/* table node j of np.linspace(0, T, n): j*step, last node == T exactly */
static double orc_node_time(const tcr_params* p, int j)
{
    if (j >= p->n_steps - 1) return p->total_time;
    return (double)j * (p->total_time / (double)(p->n_steps - 1));
}

void orc_time_axis(const tcr_params* p, double* t_s)
{
    for (int j = 0; j < p->n_steps; ++j) t_s[j] = orc_node_time(p, j);
}

/* the table self.Fs (coupled_fast.py:234) in spec arithmetic */
void orc_gen_f(const tcr_params* p, const double* phases, double* fs /*[4][n_steps]*/)
{
    double coef[TCR_N_PHASES * 2], F[4];
    orc_fourier_coef(p, phases, coef);
    for (int j = 0; j < p->n_steps; ++j) {
        orc_fourier_node(p, coef, orc_node_time(p, j), F);
        for (int i = 0; i < 4; ++i) fs[(size_t)i * p->n_steps + j] = F[i];
    }
}

/* scipy interp1d(kind='linear') at a scalar: idx = searchsorted(t_s, t, 'left') clipped to
 * [1, n-1]; slope*(t - x_lo) + y_lo (coupled_fast.py:235, evaluated at bam_track.py:127) */
static void orc_fs_at(const tcr_params* p, const double* coef, double t, double F[4])
{
    int n = p->n_steps;
    int lo = 0, hi = n;                  /* first index with node_time(idx) >= t */
    while (lo < hi) { int mid = (lo + hi) >> 1; if (orc_node_time(p, mid) < t) lo = mid + 1; else hi = mid; }
    int idx = lo; if (idx < 1) idx = 1; if (idx > n - 1) idx = n - 1;
    double x_lo = orc_node_time(p, idx - 1), x_hi = orc_node_time(p, idx);
    double Flo[4], Fhi[4];
    orc_fourier_node(p, coef, x_lo, Flo);
    orc_fourier_node(p, coef, x_hi, Fhi);
    for (int i = 0; i < TCR_N_SERIES; ++i) {
        double slope = (Fhi[i] - Flo[i]) / (x_hi - x_lo);
        F[i] = slope * (t - x_lo) + Flo[i];
    }
}

/* ------------------------------------------------------------------------------------ */
/* storm context                                                                         */
/* ------------------------------------------------------------------------------------ */
typedef struct orc_storm {
    const tcr_params* p;
    const orc_env* e;
    const double* coef;    /* [4][15][2] Fourier coefficients */
    double h_bl;
    int nfev;
    int land_ambig;        /* an RHS / pre-check evaluation fell inside an all-land cell, where the
                              reference's exact `f_land.ev(...) == 1` test (coupled_fast.py:38) is
                              decided by the last rounding bit of the bilinear sum               */
} orc_storm;

/* optional trace of RHS evaluation times (debug/test hook) */
static double* orc_trace_buf = 0; static int orc_trace_cap = 0; static int orc_trace_n = 0;
void orc_set_trace(double* buf, int cap) { orc_trace_buf = buf; orc_trace_cap = cap; orc_trace_n = 0; }
int orc_get_trace_n(void) { return orc_trace_n; }

/* lower Cholesky, OpenBLAS dpotf2 operation order (a_jj - dot, scale by reciprocal);
 * returns 0 on a non-positive / NaN pivot (numpy LinAlgError, bam_track.py:124) */
static int orc_chol4(const double a[4][4], double L[4][4])
{
    memset(L, 0, sizeof(double) * 16);
    for (int j = 0; j < 4; ++j) {
        double s = 0.0;
        for (int k = 0; k < j; ++k) s = s + L[j][k] * L[j][k];
        s = a[j][j] - s;
        if (!(s > 0.0)) return 0;
        double d = sqrt(s);
        L[j][j] = d;
        double r = 1.0 / d;
        for (int i = j + 1; i < 4; ++i) {
            double t = 0.0;
            for (int k = 0; k < j; ++k) t = t + L[i][k] * L[j][k];
            L[i][j] = (a[i][j] - t) * r;
        }
    }
    return 1;
}

/* BetaAdvectionTrack._env_winds + interp_wnd_mean_cov, track/bam_track.py:93-128 */
static void orc_env_winds_cell(const orc_storm* s, const orc_cell* c, double t, double w[4])
{
    double mean[4], cov[4][4], L[4][4], F[4];
    w[0] = w[1] = w[2] = w[3] = 0.0;
    for (int i = 0; i < 4; ++i) {
        mean[i] = orc_field(s->e, c, CH_MEAN + i);
        for (int j = 0; j <= i; ++j)
            cov[i][j] = orc_field(s->e, c, CH_COV + i * (i + 1) / 2 + j);
    }
    for (int i = 0; i < 4; ++i) for (int j = i; j < 4; ++j) cov[i][j] = cov[j][i];
    if (!orc_chol4(cov, L)) return;
    orc_fs_at(s->p, s->coef, t, F);
    for (int i = 0; i < 4; ++i) {
        double acc = 0.0;
        for (int j = 0; j <= i; ++j) acc = acc + L[i][j] * F[j];
        w[i] = mean[i] + acc;
    }
}

static void orc_env_winds(const orc_storm* s, double lon, double lat, double t, double w[4])
{
    w[0] = w[1] = w[2] = w[3] = 0.0;
    if (tcr_isnan(lon) || tcr_isnan(t)) return;
    orc_cell c;
    orc_cell_at(&s->e->g, lon, lat, &c);
    orc_env_winds_cell(s, &c, t, w);
}

/* Coupled_FAST._calc_steering_coefs, intensity/coupled_fast.py:183-192 */
static void orc_steering(const tcr_params* p, double v, double a[2])
{
    if (p->coupled_track) {
        int nan = 0;
        for (int i = 0; i < 2; ++i) {
            double fx = (v * 1.94384) * p->m_alpha[i] + p->y_alpha[i];
            double mn = tcr_isnan(fx) ? fx : (fx < p->alpha_max[i] ? fx : p->alpha_max[i]);
            double mx = tcr_isnan(mn) ? mn : (mn > p->alpha_min[i] ? mn : p->alpha_min[i]);
            a[i] = mx;
            if (tcr_isnan(mx)) nan = 1;
        }
        if (nan) { a[0] = p->y_alpha[0]; a[1] = p->y_alpha[1]; }
    } else {
        a[0] = p->steering_coefs[0]; a[1] = p->steering_coefs[1];
    }
}

static double orc_sign(double x) { return (double)((x > 0.0) - (x < 0.0)); }

/* Coupled_FAST.dydt, intensity/coupled_fast.py:196-207 with _step_bam_track
 * (bam_track.py:131-144), _dvdt (:141-150), _calc_alpha (:65-81), _dmdt (:175-180) */
static void orc_dydt(orc_storm* s, double t, const double y[4], double dy[4])
{
    const tcr_params* p = s->p;
    const double lon = y[0], lat = y[1], v = y[2], m = y[3];
    double a[2], w[4], vb0, vb1;
    if (orc_trace_buf && orc_trace_n < orc_trace_cap) orc_trace_buf[orc_trace_n++] = t;
    s->nfev++;
    orc_steering(p, v, a);
    orc_cell c;
    orc_cell_at(&s->e->g, lon, lat, &c);
    double coslat = tcr_cos(lat * TCR_DEG2RAD);
    w[0] = w[1] = w[2] = w[3] = 0.0;
    if (fabs(lat) >= 80.0) {
        vb0 = vb1 = 0.0;
    } else {
        if (!(tcr_isnan(lon) || tcr_isnan(t))) orc_env_winds_cell(s, &c, t, w);
        double v_beta_sgn = orc_sign(lat) * p->v_beta;
        vb0 = (w[0] * a[0] + w[2] * a[1]) + p->u_beta * coslat;
        vb1 = (w[1] * a[0] + w[3] * a[1]) + v_beta_sgn * coslat;
    }
    dy[0] = vb0 / p->earth_R * 180.0 / TCR_PI / coslat;
    dy[1] = vb1 / p->earth_R * 180.0 / TCR_PI;

    /* _get_current_vpot: exact land == 1 test (coupled_fast.py:38,54-58) */
    int all_land;
    double land = orc_land(s->e, lon, lat, &all_land);
    if (all_land) s->land_ambig = 1;
    double v_pot = (land == 1.0) ? 0.0 : orc_field(s->e, &c, CH_VPOT);
    /* _calc_alpha */
    double h_m = orc_field(s->e, &c, CH_MLD);
    double t_strat = orc_field(s->e, &c, CH_STRAT);
    double u_T = sqrt(vb0 * vb0 + vb1 * vb1);
    double bathy = orc_bathy(s->e, lon, lat);
    double alpha;
    if (bathy >= 0.0 || -h_m <= bathy || t_strat == 0.0) {
        alpha = 1.0;
    } else {
        double z = 0.01 * tcr_pow(t_strat, -0.4) * h_m * u_T * v_pot / v;
        double zc = z;                       /* np.clip keeps NaN */
        if (zc < 0.0) zc = 0.0;
        if (zc > 100.0) zc = 100.0;
        alpha = 1.0 - 0.87 * tcr_exp(-zc);
    }
    double gamma = p->epsilon + alpha * p->kappa;
    double m3 = m * m * m;
    double dvdt = 0.5 * p->Ck / s->h_bl * (alpha * p->beta * (v_pot * v_pot) * m3 - (1.0 - gamma * m3) * (v * v));
    dy[2] = tcr_isnan(dvdt) ? 0.0 : dvdt;
    /* _dmdt: venti = S * chi */
    double chi = orc_field(s->e, &c, CH_CHI);
    double su = w[0] - w[2], sv = w[1] - w[3];
    double S = sqrt(su * su + sv * sv);
    double venti = S * chi;
    dy[3] = 0.5 * p->Ck / s->h_bl * ((1.0 - m) * v - venti * m);
}

/* tc_dissipates, intensity/coupled_fast.py:246-256 with TC_Basin.in_basin (basins.py:32-37) */
static double orc_event(const tcr_params* p, const double y[4])
{
    const double* b = p->basin_bounds;
    int in_basin = ((b[0] + 1.0) < y[0] && y[0] < (b[2] - 1.0) && (b[1] + 1.0) < y[1] && y[1] < (b[3] - 1.0));
    if (!in_basin) return 0.0;
    if (fabs(y[1]) <= 2.0) return 0.0;
    double d = y[2] - 4.0;
    if (tcr_isnan(d)) return d;              /* np.maximum propagates NaN */
    return d > 0.0 ? d : 0.0;
}

/* ------------------------------------------------------------------------------------ */
/* SciPy RK45 (Dormand-Prince 5(4), Shampine dense output): scipy/integrate/_ivp/rk.py    */
/* ------------------------------------------------------------------------------------ */
static const double RK_C[6] = {0.0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0};
static const double RK_A[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
static const double RK_B[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
static const double RK_E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};
static const double RK_P[7][4] = {
    {1, -8048581381.0 / 2820520608, 8663915743.0 / 2820520608, -12715105075.0 / 11282082432},
    {0, 0, 0, 0},
    {0, 131558114200.0 / 32700410799, -68118460800.0 / 10900136933, 87487479700.0 / 32700410799},
    {0, -1754552775.0 / 470086768, 14199869525.0 / 1410260304, -10690763975.0 / 1880347072},
    {0, 127303824393.0 / 49829197408, -318862633887.0 / 49829197408, 701980252875.0 / 199316789632},
    {0, -282668133.0 / 205662961, 2019193451.0 / 616988883, -1453857185.0 / 822651844},
    {0, 40617522.0 / 29380423, -110615467.0 / 29380423, 69997945.0 / 29380423}};

/* norm(x) = ||x||_2 / sqrt(4) (scipy/integrate/_ivp/common.py norm) */
static double orc_rms4(const double x[4])
{
    double s = x[0] * x[0];
    s = fma(x[1], x[1], s);
    s = fma(x[2], x[2], s);
    s = fma(x[3], x[3], s);
    return sqrt(s) / 2.0;
}

/* sum_j coef[j] * K[j][i] over the non-zero coefficients, fused, ascending j */
static double orc_comb(double K[7][4], int i, const double* coef, int n)
{
    int j = 0;
    while (coef[j] == 0.0) ++j;
    double acc = K[j][i] * coef[j];
    for (++j; j < n; ++j) if (coef[j] != 0.0) acc = fma(K[j][i], coef[j], acc);
    return acc;
}

/* Hairer initial step: scipy/integrate/_ivp/common.py select_initial_step */
static double orc_initial_step(orc_storm* s, double t0, const double y0[4], double t_bound,
                               double max_step, const double f0[4], double rtol, double atol)
{
    double interval = fabs(t_bound - t0);
    if (interval == 0.0) return 0.0;
    double scale[4], a[4], b[4];
    for (int i = 0; i < 4; ++i) { scale[i] = atol + fabs(y0[i]) * rtol; a[i] = y0[i] / scale[i]; b[i] = f0[i] / scale[i]; }
    double d0 = orc_rms4(a), d1 = orc_rms4(b), h0;
    if (d0 < 1e-5 || d1 < 1e-5) h0 = 1e-6; else h0 = 0.01 * d0 / d1;
    if (interval < h0) h0 = interval;
    double y1[4], f1[4], c[4];
    for (int i = 0; i < 4; ++i) y1[i] = fma(h0, f0[i], y0[i]);
    orc_dydt(s, t0 + h0, y1, f1);
    for (int i = 0; i < 4; ++i) c[i] = (f1[i] - f0[i]) / scale[i];
    double d2 = orc_rms4(c) / h0, h1;
    if (d1 <= 1e-15 && d2 <= 1e-15) { h1 = h0 * 1e-3; if (h1 < 1e-6) h1 = 1e-6; }
    else h1 = tcr_pow(0.01 / (d2 > d1 ? d2 : d1), 0.2);      /* python max(d1, d2) */
    double h = 100.0 * h0;
    if (h1 < h) h = h1;
    if (interval < h) h = interval;
    if (max_step < h) h = max_step;
    return h;
}

/* Coupled_FAST.gen_track, intensity/coupled_fast.py:229-267, including scipy's solve_ivp
 * driver loop (ivp.py: event handling with a terminal event whose value is >= 0 everywhere,
 * t_eval sampling through the dense output).  coef = the storm's Fourier coefficients.
 * Returns the status; track is [n_steps][4].
 * n_clean: samples emitted before the first land-ambiguous evaluation (== n_time if none). */
int orc_gen_track(const tcr_params* p, const orc_env* e, const double* coef,
                  double lon0, double lat0, double v0, double m0, double h_bl,
                  double* track, int32_t* n_time, int32_t* nfev_out, int32_t* n_clean_out)
{
    orc_storm s = {p, e, coef, h_bl, 0, 0};
    int n_clean = -1;
    *n_time = 0; *n_clean_out = 0;
    /* ventilation pre-check (coupled_fast.py:238-244); not counted in nfev */
    {
        double w[4];
        orc_cell c;
        orc_cell_at(&e->g, lon0, lat0, &c);
        orc_env_winds(&s, lon0, lat0, 0.0, w);
        double su = w[0] - w[2], sv = w[1] - w[3];
        double S = sqrt(su * su + sv * sv);
        int all_land;
        double land = orc_land(e, lon0, lat0, &all_land);
        if (all_land) s.land_ambig = 1;
        double vpot = (land == 1.0) ? 0.0 : orc_field(e, &c, CH_VPOT);
        double chi = orc_field(e, &c, CH_CHI);
        if (vpot > 0.0) {
            double vent_index = S * chi / vpot;
            if (vent_index >= 1.0) { *nfev_out = 0; return TCR_STATUS_VENT; }
        }
    }
    const double t_bound = p->total_time, rtol = p->rtol, atol = p->atol, max_step = p->max_step;
    const int n_eval = p->n_steps;
    double t = 0.0, y[4] = {lon0, lat0, v0, m0}, f[4], K[7][4];
    orc_dydt(&s, t, y, f);
    double h_abs = orc_initial_step(&s, t, y, t_bound, max_step, f, rtol, atol);
    double g = orc_event(p, y);
    int status = 100, t_eval_i = 0, n_attempts = 0;

    while (status == 100) {
        /* ---- RungeKutta._step_impl ---- */
        double min_step = 10.0 * (tcr_bits2d(tcr_d2bits(t) + 1) - t);    /* 10*|nextafter(t, inf) - t| */
        if (h_abs > max_step) h_abs = max_step; else if (h_abs < min_step) h_abs = min_step;
        int accepted = 0, rejected = 0, failed = 0;
        double h = 0.0, t_new = t, y_new[4], f_new[4];
        while (!accepted) {
            if (h_abs < min_step || tcr_isnan(h_abs) || ++n_attempts > TCR_MAX_RK_ATTEMPTS) { failed = 1; break; }
            h = h_abs;
            t_new = t + h;
            if (t_new - t_bound > 0.0) t_new = t_bound;
            h = t_new - t;
            h_abs = fabs(h);
            /* rk_step */
            memcpy(K[0], f, sizeof f);
            for (int st = 1; st < 6; ++st) {
                double ys[4];
                for (int i = 0; i < 4; ++i) ys[i] = fma(orc_comb(K, i, RK_A[st], st), h, y[i]);
                orc_dydt(&s, t + RK_C[st] * h, ys, K[st]);
            }
            for (int i = 0; i < 4; ++i) y_new[i] = fma(h, orc_comb(K, i, RK_B, 6), y[i]);
            orc_dydt(&s, t + h, y_new, f_new);
            memcpy(K[6], f_new, sizeof f_new);
            double en[4];
            for (int i = 0; i < 4; ++i) {
                double ay = fabs(y[i]), an = fabs(y_new[i]);
                double mx = (tcr_isnan(ay) || tcr_isnan(an)) ? NAN : (ay > an ? ay : an);
                double scale = atol + mx * rtol;
                en[i] = (orc_comb(K, i, RK_E, 7) * h) / scale;
            }
            double err = orc_rms4(en);
            if (err < 1.0) {
                double factor;
                if (err == 0.0) factor = 10.0;
                else { factor = 0.9 * tcr_pow(err, -0.2); if (10.0 < factor) factor = 10.0; }
                if (rejected && 1.0 < factor) factor = 1.0;   /* min(1, factor) */
                h_abs *= factor;
                accepted = 1;
            } else {
                double fac = 0.9 * tcr_pow(err, -0.2);
                if (!(fac > 0.2)) fac = 0.2;                 /* python max(MIN_FACTOR, nan) == MIN_FACTOR */
                h_abs *= fac;
                rejected = 1;
            }
        }
        if (s.land_ambig && n_clean < 0) n_clean = t_eval_i;
        if (failed) { status = TCR_STATUS_FAILED; break; }
        double t_old = t, y_old[4];
        memcpy(y_old, y, sizeof y);
        t = t_new; memcpy(y, y_new, sizeof y); memcpy(f, f_new, sizeof f);
        if (t - t_bound >= 0.0) status = TCR_STATUS_FINISHED;      /* solver.status == 'finished' */

        /* ---- events (ivp.py find_active_events, direction 0, g >= 0 everywhere) ---- */
        double g_new = orc_event(p, y);
        int active = ((g <= 0.0) && (g_new >= 0.0)) || ((g >= 0.0) && (g_new <= 0.0));
        double t_emit = t;
        if (active) {
            /* brentq on [t_old, t]: f(t_old) == 0 returns t_old, else f(t) == 0 returns t */
            status = TCR_STATUS_EVENT;
            if (g == 0.0) t_emit = t_old;
        }
        g = g_new;

        /* ---- t_eval sampling: searchsorted(t_eval, t, side='right') ---- */
        int i_new = t_eval_i;
        while (i_new < n_eval && orc_node_time(p, i_new) <= t_emit) ++i_new;
        if (i_new > t_eval_i) {
            double Q[4][4];
            for (int i = 0; i < 4; ++i) for (int c = 0; c < 4; ++c) {
                double col[7];
                for (int j = 0; j < 7; ++j) col[j] = RK_P[j][c];
                Q[i][c] = orc_comb(K, i, col, 7);
            }
            double hd = t - t_old;
            for (int k = t_eval_i; k < i_new; ++k) {
                double x = (orc_node_time(p, k) - t_old) / hd;
                double p2 = x * x, p3 = p2 * x, p4 = p3 * x;       /* np.cumprod */
                for (int i = 0; i < 4; ++i) {
                    double acc = Q[i][0] * x;
                    acc = fma(Q[i][1], p2, acc);
                    acc = fma(Q[i][2], p3, acc);
                    acc = fma(Q[i][3], p4, acc);
                    track[(size_t)k * 4 + i] = fma(hd, acc, y_old[i]);
                }
            }
            t_eval_i = i_new;
        }
    }
    *n_time = t_eval_i;
    *n_clean_out = (n_clean < 0) ? t_eval_i : n_clean;
    *nfev_out = s.nfev;
    return status;
}

/* ------------------------------------------------------------------------------------ */
/* post-processing of one candidate: util/compute.py:178-206, wind/tc_wind.py:6-21,       */
/* util/sphere.py:15-30,58-83                                                             */
/* ------------------------------------------------------------------------------------ */
static double orc_haversine_km(const tcr_params* p, double lon1, double lat1, double lon2, double lat2)
{
    lon1 = lon1 * TCR_DEG2RAD; lat1 = lat1 * TCR_DEG2RAD;
    lon2 = lon2 * TCR_DEG2RAD; lat2 = lat2 * TCR_DEG2RAD;
    double dlon = lon2 - lon1, dlat = lat2 - lat1;
    double sa = tcr_sin(dlat / 2), sb = tcr_sin(dlon / 2);
    double a = sa * sa + tcr_cos(lat1) * tcr_cos(lat2) * (sb * sb);
    double c = 2.0 * tcr_asin(sqrt(a));
    return (p->earth_R / 1000.0) * c;
}

/* TC criteria: any(v >= thresh) and np.interp(2 d, t, v) >= thresh_2d (compute.py:187-189) */
static uint32_t orc_is_tc(const tcr_params* p, const double* track, int n_time)
{
    int any = 0;
    for (int k = 0; k < n_time; ++k) if (track[(size_t)k * 4 + 2] >= p->seed_v_thresh) any = 1;
    double t2 = 2.0 * 24 * 60 * 60, v2d;
    if (t2 >= orc_node_time(p, n_time - 1)) v2d = track[(size_t)(n_time - 1) * 4 + 2];
    else {
        int j = 0;
        while (j + 1 < n_time && orc_node_time(p, j + 1) <= t2) ++j;        /* t_s[j] <= t2 < t_s[j+1] */
        double x0 = orc_node_time(p, j), x1 = orc_node_time(p, j + 1);
        double slope = (track[(size_t)(j + 1) * 4 + 2] - track[(size_t)j * 4 + 2]) / (x1 - x0);
        v2d = slope * (t2 - x0) + track[(size_t)j * 4 + 2];
    }
    return (any && v2d >= p->seed_v_2d_thresh) ? TCR_FLAG_IS_TC : 0u;
}

/* returns flag bits; env [n_time][4], vmax [n_time] */
uint32_t orc_postprocess(const tcr_params* p, const orc_env* e, const double* coef,
                         const double* track, int n_time, double* env, double* vmax)
{
    orc_storm s = {p, e, coef, 0.0, 0, 0};
    if (n_time <= 0) return 0;
    uint32_t flags = orc_is_tc(p, track, n_time);

    for (int k = 0; k < n_time; ++k)
        orc_env_winds(&s, track[(size_t)k * 4 + 0], track[(size_t)k * 4 + 1], orc_node_time(p, k), env + (size_t)k * 4);

    double best = 0.0; int have = 0;
    for (int k = 0; k < n_time; ++k) {
        double ut, vt;
        if (n_time <= 1) { ut = vt = NAN; }
        else {
            #define LON(i) track[(size_t)(i) * 4 + 0]
            #define LAT(i) track[(size_t)(i) * 4 + 1]
            double lon_m = (k == 0) ? 2 * LON(0) - LON(1) : LON(k - 1);
            double lat_m = (k == 0) ? 2 * LAT(0) - LAT(1) : LAT(k - 1);
            double lon_p = (k == n_time - 1) ? 2 * LON(n_time - 1) - LON(n_time - 2) : LON(k + 1);
            double lat_p = (k == n_time - 1) ? 2 * LAT(n_time - 1) - LAT(n_time - 2) : LAT(k + 1);
            double dlon = 0.5 * (orc_sign(lon_p - lon_m) * orc_haversine_km(p, lon_p, LAT(k), lon_m, LAT(k)));
            double dlat = 0.5 * (orc_sign(lat_p - lat_m) * orc_haversine_km(p, LON(k), lat_p, LON(k), lat_m));
            ut = dlon * 1000.0 / p->dt_track;
            vt = dlat * 1000.0 / p->dt_track;
            #undef LON
            #undef LAT
        }
        double lat = track[(size_t)k * 4 + 1], v = track[(size_t)k * 4 + 2];
        const double* w = env + (size_t)k * 4;
        double G = 0.8 + 0.35 * (1.0 + tcr_tanh((lat - 35.0) / 10.0));
        if (1.0 < G) G = 1.0;
        double u_shr = w[0] - w[2], v_shr = w[1] - w[3];
        double U = G * ut + 0.1 * u_shr * v / 15.0;
        double V = G * vt + 0.1 * v_shr * v / 15.0;
        double mag_inc = sqrt(U * U + V * V);
        double vm;
        if (mag_inc == 0.0) {
            vm = fabs(v);                     /* theta = atan2(-0, +-0): increment vanishes */
        } else {
            double mag_fac = (v * 0.50) / mag_inc;
            if (!tcr_isnan(mag_fac) && 1.0 < mag_fac) mag_fac = 1.0;
            /* theta = atan2(-U, V): -sin(theta) = U/|inc|, cos(theta) = V/|inc| */
            double ug = v * (U / mag_inc) + U * mag_fac;
            double vg = v * (V / mag_inc) + V * mag_fac;
            vm = sqrt(ug * ug + vg * vg);
        }
        vmax[k] = vm;
        if (!tcr_isnan(vm)) { if (!have || vm > best) best = vm; have = 1; }
    }
    if ((flags & TCR_FLAG_IS_TC) && have && best >= p->seed_vmax_thresh) flags |= TCR_FLAG_KEPT;
    return flags;
}

/* ------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. 2011), the counter-based generator that replaces the      */
/* wall-clock-seeded MT19937 stream of the reference (bam_track.py:37-42)                 */
/* ------------------------------------------------------------------------------------ */
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static double orc_u53(uint32_t hi, uint32_t lo)
{
    return (double)((((uint64_t)(hi >> 5)) << 26) | (uint64_t)(lo >> 6)) * (1.0 / 9007199254740992.0);
}

/* two uniform doubles of draw block `blk` in stream `stream` of attempt k */
void orc_draw2(uint32_t run_seed, int32_t year_key, int64_t k, uint32_t blk, uint32_t stream, double u[2])
{
    uint32_t ctr[4] = {(uint32_t)((uint64_t)k & 0xffffffffu), (uint32_t)((uint64_t)k >> 32), blk, stream};
    uint32_t key[2] = {run_seed, (uint32_t)year_key}, o[4];
    orc_philox4x32(ctr, key, o);
    u[0] = orc_u53(o[0], o[1]);
    u[1] = orc_u53(o[2], o[3]);
}

void orc_phases(uint32_t run_seed, int32_t year_key, int64_t k, double* phases /*[60]*/)
{
    for (uint32_t b = 0; b < TCR_N_PHASES / 2; ++b) orc_draw2(run_seed, year_key, k, b, 1u, phases + 2 * b);
}

static double orc_mask(const orc_grid* gm, const double* plane, const orc_cell* c)
{
    return orc_bilin_fitpack(plane, gm->nx, c->ix, c->iy, c->wx0, c->wx1, c->wy0, c->wy1);
}

/* masks: float64 [8][ny][nx] on grid gm.  One seed attempt = one pass of the reference's
 * `while not seed_passed` body (util/compute.py:136-175).  envs = the year's 12 months.
 * Draw slots (stream 0): block 0 = (lon, lat), 1 = (month, low-latitude test),
 * 2 = Box-Muller pair for v_init, 3.. = redraw pairs.  Returns 0 not a seed, 1 counted
 * (PI <= 35), 2 passed, 3 redraw bound hit. */
int orc_seed_attempt(const tcr_params* p, const orc_env* envs /*[12]*/, const orc_grid* gm,
                     const double* masks, uint32_t run_seed, int32_t year_key, int64_t k,
                     int32_t* basin, int32_t* month, double* lon, double* lat,
                     double* v0, double* m0, double* pi_gen)
{
    const double* b = p->basin_bounds;
    const size_t plane = (size_t)gm->nx * gm->ny;
    double u[2];
    orc_draw2(run_seed, year_key, k, 0, 0, u);
    double y_min = tcr_sin(TCR_DEG2RAD * p->gen_lat_min), y_max = tcr_sin(TCR_DEG2RAD * p->gen_lat_max);
    double gen_lon = b[0] + (b[2] - b[0]) * u[0];
    double gen_lat = tcr_asin(y_min + (y_max - y_min) * u[1]) * 180.0 / TCR_PI;
    int redraw = 0, exhausted = 0;
    orc_cell c;
    for (;;) {
        orc_cell_at(gm, gen_lon, gen_lat, &c);
        if (!(orc_mask(gm, masks + 7 * plane, &c) < 1e-2)) break;
        if (redraw >= p->max_redraws) { exhausted = 1; break; }
        orc_draw2(run_seed, year_key, k, 3 + (uint32_t)redraw, 0, u);
        gen_lon = b[0] + (b[2] - b[0]) * u[0];
        gen_lat = b[1] + (b[3] - b[1]) * u[1];
        ++redraw;
    }
    orc_draw2(run_seed, year_key, k, 1, 0, u);
    int mon = 1 + (int)floor(u[0] * 12.0);
    double r_lowlat = u[1];
    double best = -INFINITY; int bi = 0;
    for (int i = 0; i < TCR_N_BASINS; ++i) {
        double val = orc_mask(gm, masks + i * plane, &c);
        if (val > best) { best = val; bi = i; }
    }
    const orc_env* e = envs + (mon - 1);
    orc_cell ce;
    orc_cell_at(&e->g, gen_lon, gen_lat, &ce);
    double pi = orc_field(e, &ce, CH_VPOT);
    double q = (fabs(gen_lat) - p->lat_vort_fac) / 12.0;
    if (q < 0.0) q = 0.0;
    if (q > 1.0) q = 1.0;
    double prob = tcr_pow(q, p->lat_vort_power[bi]);
    *basin = bi; *month = mon; *lon = gen_lon; *lat = gen_lat; *pi_gen = pi;
    orc_draw2(run_seed, year_key, k, 2, 0, u);
    double bs, bc;
    tcr_sincos2pi(u[1], &bs, &bc);
    double randn = sqrt(-2.0 * tcr_log(1.0 - u[0])) * bc;
    *v0 = p->seed_v_init + randn;
    double rh = orc_field(e, &ce, CH_RH);
    double mi = p->minit_amp / (1.0 + tcr_exp(-(rh - p->minit_center) * p->minit_slope)) + p->minit_offset;
    *m0 = mi > 0.0 ? mi : 0.0;
    if (exhausted) return 3;
    if (best > 1e-3 && r_lowlat < prob) return pi > p->pi_gen_min ? 2 : 1;
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* batch drivers (timed by bench.py as the CPU baseline; callers thread over chunks)      */
/* ------------------------------------------------------------------------------------ */
static void orc_fill_nan(double* a, size_t n) { for (size_t k = 0; k < n; ++k) a[k] = NAN; }

/* gen_track + post-processing for n explicit seeds; outputs as tcr_integrate */
void orc_integrate_batch(const tcr_params* p, const orc_env* envs, int64_t n,
                         const int32_t* ym, const double* lon0, const double* lat0,
                         const double* v0, const double* m0, const double* h_bl, const double* phases,
                         double* track, double* env, double* vmax,
                         int32_t* n_time, int32_t* status, int32_t* nfev, uint32_t* flags,
                         int32_t* n_clean, int post_all)
{
    const int ns = p->n_steps;
    double coef[TCR_N_PHASES * 2];
    for (int64_t q = 0; q < n; ++q) {
        double* tr = track + (size_t)q * ns * 4;
        double* ev = env + (size_t)q * ns * 4;
        double* vm = vmax + (size_t)q * ns;
        orc_fill_nan(tr, (size_t)ns * 4); orc_fill_nan(ev, (size_t)ns * 4); orc_fill_nan(vm, ns);
        orc_fourier_coef(p, phases + (size_t)q * TCR_N_PHASES, coef);
        status[q] = orc_gen_track(p, envs + ym[q], coef, lon0[q], lat0[q], v0[q], m0[q], h_bl[q],
                                  tr, &n_time[q], &nfev[q], &n_clean[q]);
        flags[q] = 0;
        if (status[q] != TCR_STATUS_VENT && n_time[q] > 0) {
            /* the reference only post-processes is_tc candidates (compute.py:191); post_all
             * computes env winds / vmax for every storm so tests can compare them all */
            flags[q] = orc_is_tc(p, tr, n_time[q]);
            if (post_all || (flags[q] & TCR_FLAG_IS_TC))
                flags[q] = orc_postprocess(p, envs + ym[q], coef, tr, n_time[q], ev, vm);
        }
    }
}

/* attempts [k0, k0+n) of one year: seeding + integration + post-processing, one record per
 * attempt (code 0..3 as orc_seed_attempt; integrated iff code == 2).  Used by the ordered-
 * selection oracle in oracle/tcr_oracle.py and by the CPU baseline timing.
 * track may be NULL (timing mode: only counters are returned).                             */
void orc_run_attempts(const tcr_params* p, const orc_env* envs /*[12]*/, const orc_grid* gm,
                      const double* masks, uint32_t run_seed, int32_t year_key, int64_t k0, int64_t n,
                      int32_t* code, int32_t* basin, int32_t* month, int32_t* n_time, int32_t* status,
                      int32_t* nfev, uint32_t* flags, int32_t* n_clean, double* ic /*[n][4] lon lat v0 m0*/,
                      double* track /*[n][ns][4] or NULL*/, double* env, double* vmax)
{
    const int ns = p->n_steps;
    double coef[TCR_N_PHASES * 2], ph[TCR_N_PHASES];
    double* ltr = (double*)malloc(sizeof(double) * 4 * ns);
    double* lev = (double*)malloc(sizeof(double) * 4 * ns);
    double* lvm = (double*)malloc(sizeof(double) * ns);
    for (int64_t q = 0; q < n; ++q) {
        double lon, lat, v0, m0, pi;
        code[q] = orc_seed_attempt(p, envs, gm, masks, run_seed, year_key, k0 + q,
                                   &basin[q], &month[q], &lon, &lat, &v0, &m0, &pi);
        ic[q * 4 + 0] = lon; ic[q * 4 + 1] = lat; ic[q * 4 + 2] = v0; ic[q * 4 + 3] = m0;
        n_time[q] = 0; status[q] = 0; nfev[q] = 0; flags[q] = 0; n_clean[q] = 0;
        double* tr = track ? track + (size_t)q * ns * 4 : ltr;
        double* ev = track ? env + (size_t)q * ns * 4 : lev;
        double* vm = track ? vmax + (size_t)q * ns : lvm;
        if (track) { orc_fill_nan(tr, (size_t)ns * 4); orc_fill_nan(ev, (size_t)ns * 4); orc_fill_nan(vm, ns); }
        if (code[q] != 2) continue;
        orc_phases(run_seed, year_key, k0 + q, ph);
        orc_fourier_coef(p, ph, coef);
        status[q] = orc_gen_track(p, envs + (month[q] - 1), coef, lon, lat, v0, m0,
                                  p->atm_bl_depth[basin[q]], tr, &n_time[q], &nfev[q], &n_clean[q]);
        if (status[q] != TCR_STATUS_VENT && n_time[q] > 0) {
            /* env winds + vmax only for TC candidates, as compute.py:191-205 */
            flags[q] = orc_is_tc(p, tr, n_time[q]);
            if (flags[q] & TCR_FLAG_IS_TC)
                flags[q] = orc_postprocess(p, envs + (month[q] - 1), coef, tr, n_time[q], ev, vm);
        }
    }
    free(ltr); free(lev); free(lvm);
}

/* ------------------------------------------------------------------------------------ */
/* return-period reduction: notebooks/sample_analysis.ipynb cells 13-17      */
/* ------------------------------------------------------------------------------------ */
/* cell 15: dists = haversine(clon, clat, lon_trks, lat_trks);
 *          vmax_at_poi = vmax_trks.where(dists <= radius).max(dim='time')   (NaN if never within) */
void orc_poi_vmax(int64_t n_rows, int n_steps, const double* lon, const double* lat, const double* vmax,
                  double poi_lon, double poi_lat, double radius_km, double r_earth_m, double* out)
{
    const double r_km = r_earth_m / 1000.0;
    for (int64_t r = 0; r < n_rows; ++r) {
        double best = NAN;
        for (int k = 0; k < n_steps; ++k) {
            const size_t i = (size_t)r * n_steps + k;
            double d = tcr_haversine_r(r_km, poi_lon, poi_lat, lon[i], lat[i]);
            if (d <= radius_km) {
                double v = vmax[i];
                if (!tcr_isnan(v) && (tcr_isnan(best) || v > best)) best = v;
            }
        }
        out[r] = best;
    }
}

/* cell 17: exceedance_count[i] = sum(vmax_at_poi >= vmax_bins[i]) */
void orc_exceedance(int64_t n, const double* v, int n_bins, const double* bins, int64_t* counts)
{
    for (int b = 0; b < n_bins; ++b) {
        int64_t c = 0;
        for (int64_t i = 0; i < n; ++i) if (v[i] >= bins[b]) ++c;
        counts[b] = c;
    }
}

/* ------------------------------------------------------------------------------------ */
/* unit-test hooks                                                                       */
/* ------------------------------------------------------------------------------------ */
void orc_dydt_at(const tcr_params* p, const orc_env* e, const double* coef,
                 double h_bl, double t, const double* y, double* dy)
{
    orc_storm s = {p, e, coef, h_bl, 0, 0};
    orc_dydt(&s, t, y, dy);
}

void orc_env_winds_at(const tcr_params* p, const orc_env* e, const double* coef,
                      double lon, double lat, double t, double* w)
{
    orc_storm s = {p, e, coef, 0.0, 0, 0};
    orc_env_winds(&s, lon, lat, t, w);
}

/* fn: 0 exp, 1 log, 2 sin, 3 cos, 4 asin, 5 tanh, 6 sin(2 pi x), 7 cos(2 pi x) */
void orc_libm_eval(int fn, int64_t n, const double* x, double* y)
{
    for (int64_t i = 0; i < n; ++i) {
        double s, c;
        switch (fn) {
        case 0: y[i] = tcr_exp(x[i]); break;
        case 1: y[i] = tcr_log(x[i]); break;
        case 2: y[i] = tcr_sin(x[i]); break;
        case 3: y[i] = tcr_cos(x[i]); break;
        case 4: y[i] = tcr_asin(x[i]); break;
        case 5: y[i] = tcr_tanh(x[i]); break;
        case 6: tcr_sincos2pi(x[i], &s, &c); y[i] = s; break;
        case 7: tcr_sincos2pi(x[i], &s, &c); y[i] = c; break;
        default: y[i] = NAN;
        }
    }
}

int orc_cpu_has_fma(void)
{
#if defined(__x86_64__)
    return __builtin_cpu_supports("fma") ? 1 : 0;
#else
    return 1;
#endif
}
