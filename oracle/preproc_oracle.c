/*
 * preproc_oracle.c -- CPU restatement of the reference's thermodynamic pre-processing for one
 * grid column (TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline legs may link or call this).
 *
 * Follows thermo/thermo.py of the reference step by step, with the profile arrays written out
 * per level exactly as CAPE_PI_vectorized builds them (thermo.py:266-412), for the namelist
 * defaults select_thermo = 1 (pseudoadiabatic), select_interp = 2 (entropy look-up table), and, further down,
 * for select_thermo = 2 (reversible; orc_thermo_rev):
 *   sat_thermo        thermo.py:29-39      conv_q_to_rh  thermo.py:42-47
 *   s_unsat / s_sat   thermo.py:50-76      sat_deficit   thermo.py:92-104
 *   get_LCL           thermo.py:107-127    (Romps 2017; scipy.special.lambertw branch -1)
 *   calc_T_rho        thermo.py:130-135
 * as called from thermo/calc_thermo.py:60-69.  Pinned against the UNMODIFIED reference run in the
 * build container (oracle/make_golden.py -> tests/golden/ref_thermo.npz).
 *
 * Third-party arithmetic on this path: scipy.special.lambertw(z, -1) (un-pinned SciPy; 1.18.1 here) --
 * restated as its published algorithm for real z in [-1/e, 0): initial guess log(-z), Halley steps
 * until |dw| <= 1e-8 |w| (the routine's default tol); and RectBivariateSpline(kx=1, ky=1).ev =
 * FITPACK clamped bilinear (same restatement as tcr_oracle.c).  Transcendentals through
 * include/tcr_libm.h so that the CUDA kernel can be compared bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include "../include/tcr_libm.h"

#define TH_RD 287.04
#define TH_RV 461.5
#define TH_CP (718 + 287.04)          /* util/constants.py:12-13 */
#define TH_EPS (TH_RD / TH_RV)
#define TH_L0 2.555e6
#define TH_TTRIP 273.16

typedef struct th_table { int np, ns; const double* p; const double* s; const double* T; } th_table;

/* thermo.py:29-39 (Bolton); a NaN temperature gives es = 0 (the reference's mask) */
static void th_sat(double T, double p, double* es, double* rs)
{
    double e = 0.0;
    if (T == T) {
        double Tc = T - 273;
        double x = (17.625 * Tc) / (Tc + 243.04);
        if (!(x <= 10)) x = (x != x) ? x : 10;            /* np.minimum(., 10) propagates NaN */
        e = 610.94 * tcr_exp(x);
    }
    *es = e;
    *rs = TH_RD / TH_RV * e / (p - e);
}

static double th_max(double a, double b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }   /* np.maximum */
static double th_min(double a, double b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }   /* np.minimum */

/* thermo.py:50-62, select_thermo == 1 */
static double th_s_unsat(double T, double p, double r)
{
    double es, rs;
    th_sat(T, p, &es, &rs);
    double rh = th_max(r / rs * (1 + rs / TH_EPS) / (1 + r / TH_EPS), 0);
    return TH_CP * tcr_log(T) - TH_RD * tcr_log(p - es * rh) + TH_L0 * r / T - r * TH_RV * tcr_log(rh);
}

/* thermo.py:66-76, select_thermo == 1 */
static double th_s_sat(double T, double p)
{
    double es, rs;
    th_sat(T, p, &es, &rs);
    T = th_max(T, 1e-4);
    return TH_CP * tcr_log(T) - TH_RD * tcr_log(th_max(p - es, 1e-4)) + TH_L0 * rs / T;
}

/* scipy.special.lambertw(z, k=-1).real for real z: the real branch on [-1/e, 0) */
static double th_lambertw_m1(double z)
{
    if (z != z) return NAN;
    if (z == 0.0) return -INFINITY;
    if (!(z < 0.0) || z < -0.36787944117144233) return NAN;        /* complex result: outside this restatement */
    double w = tcr_log(-z);
    for (int i = 0; i < 100; ++i) {
        double ew = tcr_exp(w);
        double wew = w * ew;
        double wewz = wew - z;
        double wn = w - wewz / (wew + ew - (w + 2) * wewz / (2 * w + 2));
        if (fabs(wn - w) <= 1e-8 * fabs(wn)) return wn;
        w = wn;
    }
    return NAN;
}

/* thermo.py:107-127 */
static double th_lcl(double p, double T, double r, double rh)
{
    const double E0v = 2.3740e6, cvv = 1418, cvl = 4119, cpv = cvv + TH_RV;
    double q = r / (1 + r);
    double Rm = (1 - q) * TH_RD + q * TH_RV;
    double cpm = (1 - q) * TH_CP + q * cpv;
    double a = cpm / Rm + (cvl - cpv) / TH_RV;
    double b = -(E0v - (cvv - cvl) * TH_TTRIP) / (TH_RV * T);
    double c = b / a;
    double T_lcl = c * T / th_lambertw_m1(tcr_pow(rh, 1 / a) * c * tcr_exp(c));
    return p * tcr_pow(T_lcl / T, cpm / Rm);
}

static double th_T_rho(double T, double rv) { return T * (1 + rv / TH_EPS) / (1 + rv); }   /* thermo.py:130-135 */

/* FITPACK clamped bilinear, degree 1 (see tcr_oracle.c: orc_locate / orc_bilin_fitpack) */
static void th_locate(const double* ax, int n, double arg, int* i0, double* w0, double* w1)
{
    double a = arg;
    if (a < ax[0]) a = ax[0];
    if (a > ax[n - 1]) a = ax[n - 1];
    int lo = 0, hi = n - 1;
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

static double th_lookup(const th_table* t, double p, double s)
{
    int ip, is; double wp0, wp1, ws0, ws1;
    th_locate(t->p, t->np, p, &ip, &wp0, &wp1);
    th_locate(t->s, t->ns, s, &is, &ws0, &ws1);
    const double* r0 = t->T + (size_t)ip * t->ns + is;
    const double* r1 = r0 + t->ns;
    double sp = 0.0;
    sp = sp + r0[0] * wp0 * ws0;
    sp = sp + r0[1] * wp0 * ws1;
    sp = sp + r1[0] * wp1 * ws0;
    sp = sp + r1[1] * wp1 * ws1;
    return sp;
}

/* CAPE_PI_vectorized for one column (thermo.py:266-412); T_env, r_env strided by `stride` */
static double th_pi_column(const th_table* tab, double cecd, double sst, double p_surf, int nlev, const double* p_env,
                           const double* dlnp, const float* T_env, const float* r_env, int64_t stride, double* w /* [7][nlev] scratch */)
{
    double* Te = w; double* Tre = w + nlev; double* Ta = w + 2 * nlev; double* ra = w + 3 * nlev;
    double* Tra = w + 4 * nlev; double* Trs = w + 5 * nlev; double* Ts = w + 6 * nlev;
    const double T_ns = (double)T_env[0], r_ns = (double)r_env[0], p_ns = p_env[0];             /* :289-291 */
    double ess, rs;
    th_sat(sst, p_surf, &ess, &rs);                                                              /* :293 */
    const double rh = r_ns / rs * (1 + rs / TH_EPS) / (1 + r_ns / TH_EPS);                       /* :295 */
    const double s_ns = th_s_unsat(T_ns, p_ns, r_ns);                                            /* :298 */
    const double ss = th_s_sat(sst, p_surf);                                                     /* :300 */
    for (int k = 0; k < nlev; ++k) {
        Te[k] = (double)T_env[k * stride];
        Tre[k] = th_T_rho(Te[k], (double)r_env[k * stride]);                                     /* :304 */
    }
    const double pLCL = th_lcl(p_ns, T_ns, r_ns, rh);                                            /* :315 */
    int icond = nlev - 1;                                                                        /* :321-324 */
    for (int k = 0; k < nlev; ++k) if (pLCL > p_env[k]) { icond = k; break; }
    for (int k = 0; k < nlev; ++k) {
        Ta[k] = T_ns * tcr_pow(p_env[k] / p_ns, TH_RD / TH_CP);                                  /* :328 */
        ra[k] = r_ns;                                                                            /* :330 */
    }
    for (int k = icond; k < nlev; ++k) {                                                         /* :333-340 */
        double es_;
        Ta[k] = th_lookup(tab, p_env[k], s_ns);
        th_sat(Ta[k], p_env[k], &es_, &ra[k]);
    }
    for (int k = 0; k < nlev; ++k) {
        double es_, rsp;
        Ts[k] = th_lookup(tab, p_env[k], ss);                                                    /* :342 */
        th_sat(Ts[k], p_env[k], &es_, &rsp);                                                     /* :355 */
        Tra[k] = th_T_rho(Ta[k], ra[k]);                                                         /* :357 */
        Trs[k] = th_T_rho(Ts[k], rsp);                                                           /* :358 */
    }
    int a_out = nlev - 1, s_out = nlev - 1;                                                      /* :361-362: last level with T_rho_parcel >= T_rho_env */
    for (int k = nlev - 1; k >= 0; --k) if (Tra[k] >= Tre[k]) { a_out = k; break; }
    for (int k = nlev - 1; k >= 0; --k) if (Trs[k] >= Tre[k]) { s_out = k; break; }
    double T_out_s = NAN, add_a = 0.0, add_s = 0.0;                                              /* :364-369 */
    if (s_out < nlev - 1) {                                                                      /* :372-383 */
        int k = s_out;
        double dT1 = Trs[k] - Tre[k], dT2 = Trs[k + 1] - Tre[k + 1];
        double p_out = (p_env[k] * dT2 - p_env[k + 1] * dT1) / (dT2 - dT1);
        T_out_s = (Te[k] * (p_out - p_env[k + 1]) + Te[k + 1] * (p_env[k] - p_out)) / (p_env[k] - p_env[k + 1]);
        add_s = TH_RD * dT1 * (p_env[k] - p_out) / (p_env[k] + p_out);
    }
    if (a_out < nlev - 1) {                                                                      /* :385-396 */
        int k = a_out;
        double dT1 = Tra[k] - Tre[k], dT2 = Tra[k + 1] - Tre[k + 1];
        double p_out = (p_env[k] * dT2 - p_env[k + 1] * dT1) / (dT2 - dT1);
        add_a = TH_RD * dT1 * (p_env[k] - p_out) / (p_env[k] + p_out);
    }
    double cape = 0.0, capes = 0.0;                                                              /* :398-404 */
    for (int k = 0; k < nlev; ++k) {
        if (k <= a_out) cape += TH_RD * (Tra[k] - Tre[k]) * -dlnp[k];
        if (k <= s_out) capes += TH_RD * (Trs[k] - Tre[k]) * -dlnp[k];
    }
    cape += add_a;                                                                               /* :405-406 */
    capes += add_s;
    cape = th_max(cape, 0);                                                                      /* :408-409 */
    if (cape != cape) cape = 0;
    double cape_diff = capes - cape;
    double pi = sqrt(th_max(cecd * (sst / T_out_s) * cape_diff, 0));                             /* :411 */
    if (pi != pi) pi = 0;                                                                        /* :412 */
    return pi;
}

/* ============================================================================================== */
/* select_thermo = 2 (reversible thermodynamics), select_interp = 2: the other branch of every function */
/* above -- thermo.py:56-60, 71-75, 132-133 -- and the three-dimensional inversion table               */
/* thermo/entropy_table_reversible.npz (p, s, rt) read through scipy.interpolate.interpn               */
/* (thermo.py:279-284, 343-353).  Pinned against the unmodified reference with                          */
/* namelist.select_thermo = 2 (oracle/make_golden_thermo.py -> tests/golden/ref_thermo_rev.npz).        */
/* ============================================================================================== */
#define TH_CPV 1870.0                  /* util/constants.py:14-17 */
#define TH_CL 4190.0
#define TH_LV 2.5e6

typedef struct th_table3 { int np, ns, nr; const double* p; const double* s; const double* r; const double* T; } th_table3;

/* thermo.py:50-60, select_thermo == 2 */
static double th_s_unsat_rev(double T, double p, double r, double r_t)
{
    double es, rs;
    th_sat(T, p, &es, &rs);
    double rh = th_max(r / rs * (1 + rs / TH_EPS) / (1 + r / TH_EPS), 0);
    double L = TH_LV - (TH_CPV - TH_CL) * (273.15 - T);
    return (TH_CP + TH_CL * r_t) * tcr_log(T) - TH_RD * tcr_log(p - es * rh) + L * r / T - r * TH_RV * tcr_log(rh);
}

/* thermo.py:64-76, select_thermo == 2 (L uses the clamped T, as the reference's statement order has it) */
static double th_s_sat_rev(double T, double p, double r_t)
{
    double es, rs;
    th_sat(T, p, &es, &rs);
    T = th_max(T, 1e-4);
    double L = TH_LV - (TH_CPV - TH_CL) * (273.15 - T);
    return (TH_CP + r_t * TH_CL) * tcr_log(T) - TH_RD * tcr_log(th_max(p - es, 1e-4)) + L * rs / T;
}

static double th_T_rho_rev(double T, double rv, double rt) { return T * (1 + rv / TH_EPS) / (1 + rt); }   /* thermo.py:132-133 */

/* scipy.interpolate._rgi_cython.find_indices for one axis: the interval i with g[i] <= x < g[i+1] (the last one closed),
 * clipped to [0, n-2]; norm distance y = (x - g[i]) / (g[i+1] - g[i]).  Returns 0 when x is NaN or outside the axis
 * (RegularGridInterpolator(bounds_error=False, fill_value=nan): the result is NaN).                                   */
static int th_rgi_cell(const double* g, int n, double x, int* i0, double* y)
{
    if (!(x >= g[0] && x <= g[n - 1])) return 0;
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (g[mid] <= x) lo = mid; else hi = mid;
    }
    if (lo > n - 2) lo = n - 2;
    *i0 = lo;
    *y = (x - g[lo]) / (g[lo + 1] - g[lo]);
    return 1;
}

/* RegularGridInterpolator._evaluate_linear in three dimensions: the eight corners in itertools.product order (first axis
 * slowest, lower corner first), weight = ((1 * w_p) * w_s) * w_rt with w = 1 - y at the lower and y at the upper node,
 * value = value + T[corner] * weight starting from 0.                                                                  */
static double th_lookup3(const th_table3* t, double p, double s, double r)
{
    int ip, is, ir; double yp, ys, yr;
    if (!th_rgi_cell(t->p, t->np, p, &ip, &yp) || !th_rgi_cell(t->s, t->ns, s, &is, &ys) || !th_rgi_cell(t->r, t->nr, r, &ir, &yr))
        return NAN;
    const double wp[2] = {1 - yp, yp}, ws[2] = {1 - ys, ys}, wr[2] = {1 - yr, yr};
    double value = 0.0;
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
            for (int c = 0; c < 2; ++c) {
                double weight = wp[a] * ws[b] * wr[c];
                value = value + t->T[((size_t)(ip + a) * t->ns + (is + b)) * t->nr + (ir + c)] * weight;
            }
    return value;
}

/* test hook: th_lookup3 at n points (tests/test_preproc.py compares it bit for bit with scipy.interpolate.interpn) */
void orc_entropy_lookup3(int64_t n, const double* p, const double* s, const double* r, int np, int ns, int nr,
                         const double* p_look, const double* s_look, const double* r_look, const double* T_look, double* out)
{
    th_table3 tab = {np, ns, nr, p_look, s_look, r_look, T_look};
    for (int64_t i = 0; i < n; ++i) out[i] = th_lookup3(&tab, p[i], s[i], r[i]);
}

/* CAPE_PI_vectorized for one column with select_thermo == 2 (thermo.py:266-412) */
static double th_pi_column_rev(const th_table3* tab, double cecd, double sst, double p_surf, int nlev, const double* p_env,
                               const double* dlnp, const float* T_env, const float* r_env, int64_t stride, double* w /* [7][nlev] scratch */)
{
    double* Te = w; double* Tre = w + nlev; double* Ta = w + 2 * nlev; double* ra = w + 3 * nlev;
    double* Tra = w + 4 * nlev; double* Trs = w + 5 * nlev; double* Ts = w + 6 * nlev;
    const double T_ns = (double)T_env[0], r_ns = (double)r_env[0], p_ns = p_env[0];             /* :289-291 */
    double ess, rs;
    th_sat(sst, p_surf, &ess, &rs);                                                              /* :293 */
    const double rh = r_ns / rs * (1 + rs / TH_EPS) / (1 + r_ns / TH_EPS);                       /* :295 */
    const double s_ns = th_s_unsat_rev(T_ns, p_ns, r_ns, r_ns);                                  /* :298 */
    const double ss = th_s_sat_rev(sst, p_surf, rs);                                             /* :300 */
    for (int k = 0; k < nlev; ++k) {
        Te[k] = (double)T_env[k * stride];
        Tre[k] = th_T_rho_rev(Te[k], (double)r_env[k * stride], (double)r_env[k * stride]);      /* :304 */
    }
    const double pLCL = th_lcl(p_ns, T_ns, r_ns, rh);                                            /* :315 */
    int icond = nlev - 1;                                                                        /* :321-324 */
    for (int k = 0; k < nlev; ++k) if (pLCL > p_env[k]) { icond = k; break; }
    for (int k = 0; k < nlev; ++k) {
        Ta[k] = T_ns * tcr_pow(p_env[k] / p_ns, TH_RD / TH_CP);                                  /* :328 */
        ra[k] = r_ns;                                                                            /* :330 */
    }
    for (int k = icond; k < nlev; ++k) {                                                         /* :343-351 */
        double es_;
        Ta[k] = th_lookup3(tab, p_env[k], s_ns, r_ns);
        th_sat(Ta[k], p_env[k], &es_, &ra[k]);
    }
    for (int k = 0; k < nlev; ++k) {
        double es_, rsp;
        Ts[k] = th_lookup3(tab, p_env[k], ss, rs);                                               /* :353 */
        th_sat(Ts[k], p_env[k], &es_, &rsp);                                                     /* :355 */
        Tra[k] = th_T_rho_rev(Ta[k], ra[k], r_ns);                                               /* :357 */
        Trs[k] = th_T_rho_rev(Ts[k], rsp, rs);                                                   /* :358 */
    }
    int a_out = nlev - 1, s_out = nlev - 1;                                                      /* :361-362 */
    for (int k = nlev - 1; k >= 0; --k) if (Tra[k] >= Tre[k]) { a_out = k; break; }
    for (int k = nlev - 1; k >= 0; --k) if (Trs[k] >= Tre[k]) { s_out = k; break; }
    double T_out_s = NAN, add_a = 0.0, add_s = 0.0;                                              /* :364-369 */
    if (s_out < nlev - 1) {                                                                      /* :372-383 */
        int k = s_out;
        double dT1 = Trs[k] - Tre[k], dT2 = Trs[k + 1] - Tre[k + 1];
        double p_out = (p_env[k] * dT2 - p_env[k + 1] * dT1) / (dT2 - dT1);
        T_out_s = (Te[k] * (p_out - p_env[k + 1]) + Te[k + 1] * (p_env[k] - p_out)) / (p_env[k] - p_env[k + 1]);
        add_s = TH_RD * dT1 * (p_env[k] - p_out) / (p_env[k] + p_out);
    }
    if (a_out < nlev - 1) {                                                                      /* :385-396 */
        int k = a_out;
        double dT1 = Tra[k] - Tre[k], dT2 = Tra[k + 1] - Tre[k + 1];
        double p_out = (p_env[k] * dT2 - p_env[k + 1] * dT1) / (dT2 - dT1);
        add_a = TH_RD * dT1 * (p_env[k] - p_out) / (p_env[k] + p_out);
    }
    double cape = 0.0, capes = 0.0;                                                              /* :398-404 */
    for (int k = 0; k < nlev; ++k) {
        if (k <= a_out) cape += TH_RD * (Tra[k] - Tre[k]) * -dlnp[k];
        if (k <= s_out) capes += TH_RD * (Trs[k] - Tre[k]) * -dlnp[k];
    }
    cape += add_a;                                                                               /* :405-406 */
    capes += add_s;
    cape = th_max(cape, 0);                                                                      /* :408-409 */
    if (cape != cape) cape = 0;
    double cape_diff = capes - cape;
    double pi = sqrt(th_max(cecd * (sst / T_out_s) * cape_diff, 0));                             /* :411 */
    if (pi != pi) pi = 0;                                                                        /* :412 */
    return pi;
}

/* orc_thermo with namelist.select_thermo = 2: table = (p [np], s [ns], rt [nr], T [np][ns][nr]) */
void orc_thermo_rev(int64_t n_pts, int nlev, const double* p_env, const float* ta, const float* hus, const double* sst,
                    const double* psl, int np, int ns, int nr, const double* p_look, const double* s_look, const double* r_look,
                    const double* T_look, double cecd, int k_mid, double p_mid, double* vmax, double* chi, double* rh_mid)
{
    th_table3 tab = {np, ns, nr, p_look, s_look, r_look, T_look};
    double* dlnp = (double*)malloc(sizeof(double) * (size_t)nlev * 9);
    double* lnp = dlnp + nlev;
    double* w = dlnp + 2 * nlev;
    for (int k = 0; k < nlev; ++k) lnp[k] = tcr_log(p_env[k]);                                   /* thermo.py:302-303 */
    for (int k = 0; k + 1 < nlev; ++k) dlnp[k] = lnp[k + 1] - lnp[k];
    dlnp[nlev - 1] = (2 * lnp[nlev - 1] - lnp[nlev - 2]) - lnp[nlev - 1];
    for (int64_t c = 0; c < n_pts; ++c) {
        vmax[c] = th_pi_column_rev(&tab, cecd, sst[c], psl[c], nlev, p_env, dlnp, ta + c, hus + c, n_pts, w);
        const double Tm = (double)ta[(size_t)k_mid * n_pts + c], qm = (double)hus[(size_t)k_mid * n_pts + c];
        /* sat_deficit (thermo.py:92-104): all three entropies carry the MID-LEVEL mixing ratio as total water */
        double sp = th_s_unsat_rev(Tm, p_mid, qm, qm);
        double sps = th_s_sat_rev(Tm, p_mid, qm);
        double spss = th_s_sat_rev(sst[c], psl[c], qm);
        chi[c] = th_min(th_max((sps - sp) / (spss - sps), 0), 10);
        double es, rs;
        th_sat(Tm, p_mid, &es, &rs);                                                             /* conv_q_to_rh, thermo.py:42-47 */
        double qs = rs / (1 + rs);
        rh_mid[c] = th_min(th_max(qm / qs, 1e-5), 1);
    }
    free(dlnp);
}

/* One time sample of compute_thermo (calc_thermo.py:60-69): vmax, chi, rh_mid for n_pts columns.
 * ta, hus [nlev][n_pts] float32, lowest model level (highest pressure) first; p_env in Pa.        */
void orc_thermo(int64_t n_pts, int nlev, const double* p_env, const float* ta, const float* hus, const double* sst,
                const double* psl, int np, int ns, const double* p_look, const double* s_look, const double* T_look,
                double cecd, int k_mid, double p_mid, double* vmax, double* chi, double* rh_mid)
{
    th_table tab = {np, ns, p_look, s_look, T_look};
    double* dlnp = (double*)malloc(sizeof(double) * (size_t)nlev * 9);
    double* lnp = dlnp + nlev;
    double* w = dlnp + 2 * nlev;
    for (int k = 0; k < nlev; ++k) lnp[k] = tcr_log(p_env[k]);                                   /* thermo.py:302-303 */
    for (int k = 0; k + 1 < nlev; ++k) dlnp[k] = lnp[k + 1] - lnp[k];
    dlnp[nlev - 1] = (2 * lnp[nlev - 1] - lnp[nlev - 2]) - lnp[nlev - 1];
    for (int64_t c = 0; c < n_pts; ++c) {
        vmax[c] = th_pi_column(&tab, cecd, sst[c], psl[c], nlev, p_env, dlnp, ta + c, hus + c, n_pts, w);
        const double Tm = (double)ta[(size_t)k_mid * n_pts + c], qm = (double)hus[(size_t)k_mid * n_pts + c];
        /* sat_deficit(sst, psl, T_mid, p_mid, q_mid), thermo.py:92-104; clipped to [0, 10] (calc_thermo.py:68) */
        double sp = th_s_unsat(Tm, p_mid, qm);
        double sps = th_s_sat(Tm, p_mid);
        double spss = th_s_sat(sst[c], psl[c]);
        chi[c] = th_min(th_max((sps - sp) / (spss - sps), 0), 10);
        /* conv_q_to_rh, thermo.py:42-47 */
        double es, rs;
        th_sat(Tm, p_mid, &es, &rs);
        double qs = rs / (1 + rs);
        rh_mid[c] = th_min(th_max(qm / qs, 1e-5), 1);
    }
    free(dlnp);
}
