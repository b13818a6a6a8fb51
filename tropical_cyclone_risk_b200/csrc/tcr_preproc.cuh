/*
 * tcr_preproc.cuh -- the two pre-processing stages whose cost scales with the record length
 * (SURVEY 8f "next" row N3): the monthly wind mean / covariance reduction of
 * track/env_wind.py:169-228 (calc_wnd_stat) and the per-column potential-intensity /
 * saturation-deficit computation of thermo/thermo.py:266-412, 92-104, 39-45
 * (CAPE_PI_vectorized, sat_deficit, conv_q_to_rh as called from thermo/calc_thermo.py:60-69).
 *
 * Arithmetic contract: float32 samples are widened to float64 exactly; every sum runs in time /
 * level order in float64 without contraction (nvcc -fmad=false), transcendentals through
 * include/tcr_libm.h -- the CUDA results are bit-identical to the CPU restatement the tests check them against.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "../../include/tcr_libm.h"

/* ======================================================================================== */
/* k_wind_stats: calc_wnd_stat (track/env_wind.py:169-228)                                    */
/* ======================================================================================== */
/* Per grid point, from the sub-daily samples of (ua250, va250, ua850, va850):
 *   daily means        ua.sel(time=month_mask).groupby("time.day").mean(dim='time')  (:188-190)
 *   monthly means      month_wnds[i].mean(dim=t_unit)                                  (:207)
 *   variances, ddof 0  month_wnds[i].var(dim=t_unit)                                   (:211)
 *   covariances, ddof 1  xr.cov(month_wnds[i], month_wnds[j], dim=t_unit)               (:213)
 * in the reference's output order: 4 means, then the lower triangle row by row (:215-216).
 * NaN samples are skipped the way xarray's skipna reductions skip them (nanmean / nanvar per
 * variable; xr.cov on the days where both variables are valid).
 *
 * Mapping.  A CTA owns P consecutive grid points (64 by default) for the whole month and reads every
 * sample of theirs exactly once.
 *   ungrouped input (k_wind_stats_single -- the reference's own case): the samples go global -> shared
 *     as 16-byte cp.async copies, all in flight at once, and stay float32 in shared memory;
 *   grouped input (k_wind_stats): warp (v, k) streams variable v for the days g = k, k + KSPLIT, ...,
 *     each lane VEC consecutive points (one 128*VEC-byte row segment per warp load, U independent loads
 *     in flight per lane, ld.global.cs) and leaves the float64 daily means in shared memory.
 * After one barrier two threads per point read its day series [day][variable][point] back from shared
 * memory twice (sums, then demeaned products) and write the 14 statistics with coalesced stores.
 * HBM traffic = the algorithmic minimum: 16 B per (sample, point) read once, 112 B per point written.
 * The kernel needs ~2.2 fp64-pipe operations (DADD / DMUL / F2F.F64.F32) per byte against a machine
 * balance of ~2.9: it meets the fp64 pipe and HBM at about the same time.       */
struct WindStatArgs {
    const float* src[4];        /* variable v, sample t, point p at src[v][t * t_stride + p]        */
    int64_t t_stride;           /* elements between consecutive samples                            */
    int64_t n_pts;
    int n_time, n_groups;
    const int32_t* gstart;      /* [n_groups + 1] samples of day g = [gstart[g], gstart[g+1])      */
    double* out;                /* [14][n_pts]                                                     */
};

template <int VEC> struct WsVec;
template <> struct WsVec<1> { typedef float T; };
template <> struct WsVec<2> { typedef float2 T; };
template <> struct WsVec<4> { typedef float4 T; };

template <int VEC>
__device__ __forceinline__ void ws_load(const float* p, bool vec_ok, int64_t p0, int64_t n_pts, float (&x)[VEC])
{
    if (VEC > 1 && vec_ok) {
        typename WsVec<VEC>::T q = __ldcs(reinterpret_cast<const typename WsVec<VEC>::T*>(p + p0));
        const float* qf = reinterpret_cast<const float*>(&q);
#pragma unroll
        for (int e = 0; e < VEC; ++e) x[e] = qf[e];
    } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) x[e] = __ldcs(p + min(p0 + e, n_pts - 1));   /* tail / unaligned rows */
    }
}

/* one (i, j) moment pair of one point with the NaN policy spelled out -- the arithmetic every path must reproduce */
template <class DM>
__device__ __noinline__ void ws_pair_general(const DM* a, const DM* b, size_t gstride, int n_groups, bool diag, double* mean_i, double* m2)
{
    double si = 0.0, sj = 0.0;
    int cnt = 0;
    for (int g = 0; g < n_groups; ++g) {
        const double xa = (double)a[g * gstride], xb = (double)b[g * gstride];
        const bool ok = (xa == xa) && (xb == xb);
        si += ok ? xa : 0.0;
        sj += ok ? xb : 0.0;
        cnt += ok ? 1 : 0;
    }
    const double mi = si / (double)cnt, mj = sj / (double)cnt;
    double acc = 0.0;
    for (int g = 0; g < n_groups; ++g) {
        const double xa = (double)a[g * gstride], xb = (double)b[g * gstride];
        const bool ok = (xa == xa) && (xb == xb);
        const double da = xa - mi, db = xb - mj;
        const double pr = da * db;
        acc += ok ? pr : 0.0;
    }
    *mean_i = mi;
    *m2 = diag ? acc / (double)cnt                                 /* .var(): ddof = 0 */
               : (cnt ? acc / (double)(cnt - 1) : NAN);            /* xr.cov: ddof = 1, min_count = 1 */
}

/* ---- phase 2: the four means and ten second moments of every point of the CTA.
 * TWO threads per point (warp-uniform halves, P a multiple of 32): half 0 sums variables 0-1 and forms the
 * moment pairs k = 0..4 = (0,0) (1,0) (1,1) (2,0) (2,1); half 1 sums variables 2-3 and forms k = 5..9 =
 * (2,2) (3,0) (3,1) (3,2) (3,3).  The four sums meet in shared memory (sx) so that both halves know the
 * means and whether the point has NaNs.  A point without NaNs takes the straight-line path -- the same
 * float64 operations in the same order as ws_pair_general performs pair by pair, which handles the
 * points that do have NaNs.  The split doubles the warps per byte of shared memory (the occupancy limit)
 * for 1.2 x the arithmetic. ---- */
template <class DM, int P, int NT>
__device__ __forceinline__ void ws_moments(const DM* dm, double* sx /* [4][P] */, int n_groups, int64_t cta0, int64_t n_pts, double* out)
{
    static_assert(P % 32 == 0 && NT >= 2 * P, "two warp-uniform halves per point");
    constexpr size_t GS = 4 * P;
    const int half = threadIdx.x / P, pt = threadIdx.x - half * P;
    const bool worker = threadIdx.x < 2 * P;
    const DM* x = dm + pt;
    if (worker) {
        const DM* xa = x + (2 * half) * P;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll 4
        for (int g = 0; g < n_groups; ++g) { s0 += (double)xa[g * GS]; s1 += (double)xa[g * GS + P]; }
        sx[(2 * half) * P + pt] = s0;
        sx[(2 * half + 1) * P + pt] = s1;
    }
    __syncthreads();
    if (!worker) return;
    const int64_t p = cta0 + pt;
    if (p >= n_pts) return;
    double s[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) s[v] = sx[v * P + pt];
    const int k0 = half * 5;
    double m[4], m2[5];
    if ((s[0] != s[0]) || (s[1] != s[1]) || (s[2] != s[2]) || (s[3] != s[3])) {
        int k = 0;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j <= i; ++j, ++k) {
                if (k < k0 || k >= k0 + 5) continue;
                double mi;
                ws_pair_general<DM>(x + i * P, x + j * P, GS, n_groups, i == j, &mi, &m2[k - k0]);
                if (i == j) m[i] = mi;
            }
    } else {
        const double n = (double)n_groups, n1 = (double)(n_groups - 1);
#pragma unroll
        for (int v = 0; v < 4; ++v) m[v] = s[v] / n;
#pragma unroll
        for (int k = 0; k < 5; ++k) m2[k] = 0.0;
        if (half == 0) {
#pragma unroll 2
            for (int g = 0; g < n_groups; ++g) {
                const double d0 = (double)x[g * GS] - m[0], d1 = (double)x[g * GS + P] - m[1], d2 = (double)x[g * GS + 2 * P] - m[2];
                const double p00 = d0 * d0, p10 = d1 * d0, p11 = d1 * d1, p20 = d2 * d0, p21 = d2 * d1;
                m2[0] += p00; m2[1] += p10; m2[2] += p11; m2[3] += p20; m2[4] += p21;
            }
            m2[0] /= n; m2[1] /= n1; m2[2] /= n; m2[3] /= n1; m2[4] /= n1;
        } else {
#pragma unroll 2
            for (int g = 0; g < n_groups; ++g) {
                const double d0 = (double)x[g * GS] - m[0], d1 = (double)x[g * GS + P] - m[1], d2 = (double)x[g * GS + 2 * P] - m[2],
                             d3 = (double)x[g * GS + 3 * P] - m[3];
                const double p22 = d2 * d2, p30 = d3 * d0, p31 = d3 * d1, p32 = d3 * d2, p33 = d3 * d3;
                m2[0] += p22; m2[1] += p30; m2[2] += p31; m2[3] += p32; m2[4] += p33;
            }
            m2[0] /= n; m2[1] /= n1; m2[2] /= n1; m2[3] /= n1; m2[4] /= n;
        }
    }
    __stcs(out + (size_t)(2 * half) * n_pts + p, m[2 * half]);
    __stcs(out + (size_t)(2 * half + 1) * n_pts + p, m[2 * half + 1]);
#pragma unroll
    for (int k = 0; k < 5; ++k) __stcs(out + (size_t)(4 + k0 + k) * n_pts + p, m2[k]);
}

/* Ungrouped input: every group is one sample (the reference's own 2 x daily ERA5 input never groups,
 * see tcrisk.h) -- the "daily means" are the float32 samples themselves and stay float32 in shared memory. */
template <int P, int NT>
__global__ void __launch_bounds__(NT) k_wind_stats_single(const WindStatArgs A)
{
    extern __shared__ __align__(16) unsigned char ws_raw[];
    double* const sx = reinterpret_cast<double*>(ws_raw);          /* [4][P] sums of pass 1 */
    float* const dm = reinterpret_cast<float*>(sx + 4 * P);        /* [n_time][4][P] */
    const int64_t cta0 = (int64_t)blockIdx.x * P;
    /* ---- phase 1: the CTA's samples, global -> shared with 16-byte asynchronous copies (LDGSTS),
     *      all of them in flight at once; ragged or unaligned rows go through registers ---- */
    constexpr int QPR = P / 4;                                     /* 16-byte chunks per row */
    constexpr int RPI = NT / QPR;                                  /* rows per iteration of the CTA */
    static_assert(NT % QPR == 0, "whole rows per iteration");
    bool al = (A.t_stride % 4) == 0 && (cta0 + P <= A.n_pts);
#pragma unroll
    for (int v = 0; v < 4; ++v) al = al && ((reinterpret_cast<uintptr_t>(A.src[v]) & 15) == 0);
    const int q = threadIdx.x % QPR, r0 = threadIdx.x / QPR;
    const int n_rows = A.n_groups * 4;
    if (al) {
        for (int row = r0; row < n_rows; row += RPI) {
            const float* src = A.src[row & 3] + (int64_t)(row >> 2) * A.t_stride + cta0 + q * 4;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dm + (size_t)row * P + q * 4)),
                         "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
        for (int row = r0; row < n_rows; row += RPI) {
            const float* src = A.src[row & 3] + (int64_t)(row >> 2) * A.t_stride;
#pragma unroll
            for (int e = 0; e < 4; ++e) dm[(size_t)row * P + q * 4 + e] = __ldcs(src + min(cta0 + q * 4 + e, A.n_pts - 1));
        }
    }
    __syncthreads();
    ws_moments<float, P, NT>(dm, sx, A.n_groups, cta0, A.n_pts, A.out);
}

/* Ungrouped records too long for a shared-memory month tile (an hourly month: 744 samples): the same sums and demeaned
 * products in the same order, one thread per grid point, the series read from global memory twice (coalesced across
 * the points of a warp; 2 x the HBM traffic of the tiled kernels).  ws_moments is the specification of the arithmetic. */
__global__ void __launch_bounds__(256) k_wind_stats_stream(const WindStatArgs A)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= A.n_pts) return;
    const int n_groups = A.n_groups;
    const size_t gs = (size_t)A.t_stride;
    const float* x[4] = {A.src[0] + p, A.src[1] + p, A.src[2] + p, A.src[3] + p};
    double s[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
    for (int g = 0; g < n_groups; ++g) {
#pragma unroll
        for (int v = 0; v < 4; ++v) s[v] += (double)__ldg(x[v] + g * gs);
    }
    double m[4], m2[10];
    if ((s[0] != s[0]) || (s[1] != s[1]) || (s[2] != s[2]) || (s[3] != s[3])) {
        int k = 0;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j <= i; ++j, ++k) {
                double mi;
                ws_pair_general<float>(x[i], x[j], gs, n_groups, i == j, &mi, &m2[k]);
                if (i == j) m[i] = mi;
            }
    } else {
        const double n = (double)n_groups, n1 = (double)(n_groups - 1);
#pragma unroll
        for (int v = 0; v < 4; ++v) m[v] = s[v] / n;
#pragma unroll
        for (int k = 0; k < 10; ++k) m2[k] = 0.0;
#pragma unroll 2
        for (int g = 0; g < n_groups; ++g) {
            const double d0 = (double)__ldg(x[0] + g * gs) - m[0], d1 = (double)__ldg(x[1] + g * gs) - m[1],
                         d2 = (double)__ldg(x[2] + g * gs) - m[2], d3 = (double)__ldg(x[3] + g * gs) - m[3];
            const double p00 = d0 * d0, p10 = d1 * d0, p11 = d1 * d1, p20 = d2 * d0, p21 = d2 * d1;
            const double p22 = d2 * d2, p30 = d3 * d0, p31 = d3 * d1, p32 = d3 * d2, p33 = d3 * d3;
            m2[0] += p00; m2[1] += p10; m2[2] += p11; m2[3] += p20; m2[4] += p21;
            m2[5] += p22; m2[6] += p30; m2[7] += p31; m2[8] += p32; m2[9] += p33;
        }
        m2[0] /= n; m2[1] /= n1; m2[2] /= n; m2[3] /= n1; m2[4] /= n1;
        m2[5] /= n; m2[6] /= n1; m2[7] /= n1; m2[8] /= n1; m2[9] /= n;
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) __stcs(A.out + (size_t)v * A.n_pts + p, m[v]);
#pragma unroll
    for (int k = 0; k < 10; ++k) __stcs(A.out + (size_t)(4 + k) * A.n_pts + p, m2[k]);
}

/* Grouped input: daily means first. */
template <int VEC, int KSPLIT, int U>
__global__ void __launch_bounds__(128 * KSPLIT) k_wind_stats(const WindStatArgs A)
{
    constexpr int P = 32 * VEC, NT = 128 * KSPLIT;
    extern __shared__ __align__(16) unsigned char ws_raw[];
    double* const sx = reinterpret_cast<double*>(ws_raw);          /* [4][P] sums of pass 1 */
    double* const ws_dm = sx + 4 * P;                              /* [n_groups][4][P] daily means */
    const int64_t cta0 = (int64_t)blockIdx.x * P;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int v = warp & 3, k = warp >> 2;
    const int64_t p0 = cta0 + (int64_t)lane * VEC;
    const float* src = A.src[v];
    /* vector loads need the whole VEC-group inside the row and a 4*VEC-byte aligned address in every row */
    const bool vec_ok = (p0 + VEC <= A.n_pts) && ((A.t_stride % VEC) == 0) &&
                        ((reinterpret_cast<uintptr_t>(src) & (sizeof(float) * VEC - 1)) == 0);
    const int64_t pc = min(p0, A.n_pts - 1);

    /* ---- phase 1: daily means (groupby("time.day").mean, skipna) ---- */
    for (int g = k; g < A.n_groups; g += KSPLIT) {
        const int t0 = __ldg(A.gstart + g), t1 = __ldg(A.gstart + g + 1);
        double s[VEC];
        int c[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) { s[e] = 0.0; c[e] = 0; }
        for (int t = t0; t < t1; t += U) {
            float x[U][VEC];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int tt = min(t + u, t1 - 1);                 /* the surplus loads repeat the last sample (L1 hit) */
                ws_load<VEC>(src + (int64_t)tt * A.t_stride, vec_ok, vec_ok ? p0 : pc, A.n_pts, x[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (t + u < t1) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        const double d = (double)x[u][e];
                        const bool ok = (__float_as_uint(x[u][e]) & 0x7fffffffu) <= 0x7f800000u;   /* not NaN, on the integer pipe */
                        s[e] += ok ? d : 0.0;
                        c[e] += ok ? 1 : 0;
                    }
                }
            }
        }
#pragma unroll
        for (int e = 0; e < VEC; ++e)
        {
            /* sum / count: a power-of-two count (2, 4, 8 samples a day) divides exactly by multiplication */
            const int ce = c[e];
            double mean;
            if (ce == 0) mean = NAN;
            else if ((ce & (ce - 1)) == 0) mean = s[e] * __hiloint2double((1023 - (__ffs(ce) - 1)) << 20, 0);
            else mean = s[e] / (double)ce;
            ws_dm[((size_t)g * 4 + v) * P + lane * VEC + e] = mean;
        }
    }
    __syncthreads();
    ws_moments<double, P, NT>(ws_dm, sx, A.n_groups, cta0, A.n_pts, A.out);
}

/* ======================================================================================== */
/* k_thermo: potential intensity, saturation deficit and mid-level relative humidity of every   */
/* grid column -- one time sample of compute_thermo (thermo/calc_thermo.py:60-69):              */
/*   vmax   = thermo.CAPE_PI_vectorized(sst, psl, p_env, ta, hus)          thermo.py:266-412    */
/*   chi    = clip(thermo.sat_deficit(sst, psl, T_mid, p_mid, q_mid), 0, 10) thermo.py:92-104   */
/*   rh_mid = thermo.conv_q_to_rh(T_mid, q_mid, p_mid)                     thermo.py:42-47      */
/* for the namelist defaults select_thermo = 1 (pseudoadiabatic), select_interp = 2 (entropy     */
/* look-up table through RectBivariateSpline(kx=1, ky=1).ev, i.e. FITPACK clamped bilinear).     */
/* ======================================================================================== */
/* The reference builds seven (level, lat, lon) profile arrays and then scans them for the levels
 * of neutral buoyancy.  Here one thread owns one column and walks its levels ONCE, bottom up:
 * the parcel profiles exist only as the current level's scalars; "the last level where the parcel is
 * at least as dense-warm as the environment" (thermo.py:361-362) is tracked by snapshotting, at every
 * level that qualifies, the running CAPE sum and the two temperatures the outflow interpolation
 * (thermo.py:372-396) needs, and completing the snapshot with the next level's values one iteration
 * later.  Everything that depends on the level only (log-pressure increments, the dry-adiabat factor,
 * the pressure-axis cell and weights of the look-up table) is computed once by k_thermo_levels and read
 * from shared memory; the entropy-axis cell of each parcel is located once per column.  Loads of
 * ta / hus are coalesced across the warp's 32 columns (one 128-byte line per level and variable):
 * 8 B per (level, column) + 16 B per column read, 24 B per column written.  The kernel is bound by the
 * float64 pipe (two exp, ~8 divisions per level), not by HBM.                                       */
#define TH_RD 287.04                      /* util/constants.py:10-13, 17-18 */
#define TH_RV 461.5
#define TH_CP (718 + 287.04)
#define TH_EPS (TH_RD / TH_RV)
#define TH_L0 2.555e6
#define TH_TTRIP 273.16
#define TH_MAX_LEVELS 64

struct ThLevel {            /* per pressure level, column independent */
    double p;               /* Pa                                                        */
    double mdlnp;           /* -dlnp[k]                    (thermo.py:302-303, 401-404)  */
    double dry;             /* (p / p_ns)^(Rd/cp)          (thermo.py:328)               */
    double wp0, wp1;        /* table weights on the pressure axis                        */
    int ip, pad;            /* table cell on the pressure axis; reversible table: pad = 1 when p lies inside the axis */
};

struct ThermoArgs {
    int64_t n_pts;
    int nlev, k_mid;
    const float* ta; const float* hus;        /* [nlev][n_pts], lowest model level first  */
    const double* sst; const double* psl;     /* [n_pts]                                  */
    const double* p_env;                      /* [nlev] Pa                                */
    ThLevel* lev;                             /* [nlev]                                   */
    int np, ns;
    const double* p_look; const double* s_look; const double* T_look;    /* entropy table [np][ns] */
    int nr; const double* r_look;     /* select_thermo = 2: total-water axis, T_look [np][ns][nr] (0 / NULL: pseudoadiabatic) */
    double cecd, p_mid;
    double* vmax; double* chi; double* rh_mid;
};

/* FITPACK fpbisp interval + degree-1 fpbspl weights with the argument clamped to the axis */
__device__ __forceinline__ void th_locate(const double* __restrict__ ax, int n, double arg, int& i0, double& w0, double& w1)
{
    double a = arg;
    if (a < __ldg(ax)) a = __ldg(ax);
    if (a > __ldg(ax + n - 1)) a = __ldg(ax + n - 1);
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(ax + mid) <= a) lo = mid; else hi = mid;
    }
    if (lo > n - 2) lo = n - 2;
    const double x0 = __ldg(ax + lo), x1 = __ldg(ax + lo + 1);
    const double f = 1.0 / (x1 - x0);
    i0 = lo;
    w0 = f * (x1 - a);
    w1 = f * (a - x0);
}

/* scipy.interpolate._rgi_cython.find_indices for one axis (the reversible table is read through interpn, thermo.py:343-353):
 * interval i with g[i] <= x < g[i+1] (the last one closed), norm distance y = (x - g[i]) / (g[i+1] - g[i]); false when x is
 * NaN or outside the axis (bounds_error=False, fill_value=nan) */
__device__ __forceinline__ bool th_rgi_cell(const double* __restrict__ g, int n, double x, int& i0, double& y)
{
    i0 = 0; y = 0.0;
    if (!(x >= __ldg(g) && x <= __ldg(g + n - 1))) return false;
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(g + mid) <= x) lo = mid; else hi = mid;
    }
    if (lo > n - 2) lo = n - 2;
    const double x0 = __ldg(g + lo), x1 = __ldg(g + lo + 1);
    i0 = lo;
    y = (x - x0) / (x1 - x0);
    return true;
}

__global__ void k_thermo_levels(const ThermoArgs A)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= A.nlev) return;
    const double p = A.p_env[k];
    const double lnp = tcr_log(p);
    double dlnp;
    if (k + 1 < A.nlev) dlnp = tcr_log(A.p_env[k + 1]) - lnp;
    else dlnp = (2 * lnp - tcr_log(A.p_env[k - 1])) - lnp;          /* np.diff(lnp, append = 2 lnp[-1] - lnp[-2]) */
    ThLevel L;
    L.p = p;
    L.mdlnp = -dlnp;
    L.dry = tcr_pow(p / A.p_env[0], TH_RD / TH_CP);
    L.pad = 0;
    if (A.nr > 0) {
        double y;
        L.pad = th_rgi_cell(A.p_look, A.np, p, L.ip, y) ? 1 : 0;
        L.wp0 = 1 - y; L.wp1 = y;
    } else {
        th_locate(A.p_look, A.np, p, L.ip, L.wp0, L.wp1);
    }
    A.lev[k] = L;
}

__device__ __forceinline__ double th_fmax(double a, double b) { return (a != a || b != b) ? NAN : (a > b ? a : b); }   /* np.maximum */
__device__ __forceinline__ double th_fmin(double a, double b) { return (a != a || b != b) ? NAN : (a < b ? a : b); }   /* np.minimum */

/* sat_thermo (thermo.py:29-39): Bolton; a NaN temperature gives es = 0 */
__device__ __forceinline__ void th_sat(double T, double p, double& es, double& rs)
{
    double e = 0.0;
    if (T == T) {
        const double Tc = T - 273;
        double x = (17.625 * Tc) / (Tc + 243.04);
        if (!(x <= 10)) x = (x != x) ? x : 10;
        e = 610.94 * tcr_exp(x);
    }
    es = e;
    rs = TH_RD / TH_RV * e / (p - e);
}

__device__ __forceinline__ double th_t_rho(double T, double rv) { return T * (1 + rv / TH_EPS) / (1 + rv); }

__device__ __forceinline__ double th_s_unsat(double T, double p, double r)
{
    double es, rs;
    th_sat(T, p, es, rs);
    const double rh = th_fmax(r / rs * (1 + rs / TH_EPS) / (1 + r / TH_EPS), 0);
    return TH_CP * tcr_log(T) - TH_RD * tcr_log(p - es * rh) + TH_L0 * r / T - r * TH_RV * tcr_log(rh);
}

__device__ __forceinline__ double th_s_sat(double T, double p)
{
    double es, rs;
    th_sat(T, p, es, rs);
    T = th_fmax(T, 1e-4);
    return TH_CP * tcr_log(T) - TH_RD * tcr_log(th_fmax(p - es, 1e-4)) + TH_L0 * rs / T;
}

/* select_thermo = 2 (reversible): thermo.py:56-60, 71-75, 132-133; util/constants.py:14-17 */
#define TH_CPV 1870.0
#define TH_CL 4190.0
#define TH_LV 2.5e6
__device__ __forceinline__ double th_s_unsat_rev(double T, double p, double r, double r_t)
{
    double es, rs;
    th_sat(T, p, es, rs);
    const double rh = th_fmax(r / rs * (1 + rs / TH_EPS) / (1 + r / TH_EPS), 0);
    const double L = TH_LV - (TH_CPV - TH_CL) * (273.15 - T);
    return (TH_CP + TH_CL * r_t) * tcr_log(T) - TH_RD * tcr_log(p - es * rh) + L * r / T - r * TH_RV * tcr_log(rh);
}

__device__ __forceinline__ double th_s_sat_rev(double T, double p, double r_t)
{
    double es, rs;
    th_sat(T, p, es, rs);
    T = th_fmax(T, 1e-4);
    const double L = TH_LV - (TH_CPV - TH_CL) * (273.15 - T);
    return (TH_CP + r_t * TH_CL) * tcr_log(T) - TH_RD * tcr_log(th_fmax(p - es, 1e-4)) + L * rs / T;
}

__device__ __forceinline__ double th_t_rho_rev(double T, double rv, double rt) { return T * (1 + rv / TH_EPS) / (1 + rt); }

/* RegularGridInterpolator._evaluate_linear over (p, s, rt): eight corners in itertools.product order, weight =
 * (w_p * w_s) * w_rt, value = value + T[corner] * weight from 0 */
__device__ __forceinline__ double th_lookup3(const double* __restrict__ T, int ns, int nr, int ip, double wp0, double wp1,
                                             int is, double ys, int ir, double yr)
{
    const double wp[2] = {wp0, wp1}, ws[2] = {1 - ys, ys}, wr[2] = {1 - yr, yr};
    double value = 0.0;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const double weight = wp[a] * ws[b] * wr[c];
                value = value + __ldg(T + ((size_t)(ip + a) * ns + (is + b)) * nr + (ir + c)) * weight;
            }
    return value;
}

/* real branch -1 of the Lambert W function on [-1/e, 0): scipy.special.lambertw(z, -1) as get_LCL uses it
 * (thermo.py:124): first guess log(-z), Halley steps to the routine's default tolerance 1e-8 */
__device__ __forceinline__ double th_lambertw_m1(double z)
{
    if (z != z) return NAN;
    if (z == 0.0) return -INFINITY;
    if (!(z < 0.0) || z < -0.36787944117144233) return NAN;
    double w = tcr_log(-z);
    for (int i = 0; i < 100; ++i) {
        const double ew = tcr_exp(w);
        const double wew = w * ew;
        const double wewz = wew - z;
        const double wn = w - wewz / (wew + ew - (w + 2) * wewz / (2 * w + 2));
        if (fabs(wn - w) <= 1e-8 * fabs(wn)) return wn;
        w = wn;
    }
    return NAN;
}

/* one parcel's bookkeeping of "the last level with T_rho_parcel >= T_rho_env" */
struct ThParcel {
    int out;                 /* that level, -1: none so far              */
    double cape, cape_at;    /* running sum; its value at level `out`    */
    double dT1, dT2, Te1, Te2;
    __device__ __forceinline__ void init() { out = -1; cape = cape_at = 0.0; dT1 = dT2 = Te1 = Te2 = 0.0; }
    __device__ __forceinline__ void level(int k, double Trp, double Tre, double Te, double mdlnp)
    {
        const double dT = Trp - Tre;
        if (out == k - 1 && k > 0) { dT2 = dT; Te2 = Te; }             /* completes the snapshot of level k-1 */
        cape += TH_RD * dT * mdlnp;
        if (Trp >= Tre) { out = k; dT1 = dT; Te1 = Te; cape_at = cape; }
    }
};

template <bool REV>
__global__ void __launch_bounds__(128) k_thermo(const ThermoArgs A)
{
    __shared__ ThLevel lev[TH_MAX_LEVELS];
    for (int i = threadIdx.x; i < A.nlev * (int)(sizeof(ThLevel) / 8); i += blockDim.x)
        reinterpret_cast<double*>(lev)[i] = reinterpret_cast<const double*>(A.lev)[i];
    __syncthreads();
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= A.n_pts) return;
    const int nlev = A.nlev, ns = A.ns;
    const double sst = A.sst[c], p_surf = A.psl[c];
    const double T_ns = (double)__ldcs(A.ta + c), r_ns = (double)__ldcs(A.hus + c), p_ns = lev[0].p;

    double ess, rs;
    th_sat(sst, p_surf, ess, rs);                                                       /* thermo.py:293 */
    const double rh = r_ns / rs * (1 + rs / TH_EPS) / (1 + r_ns / TH_EPS);              /* :295 */
    const double s_ns = REV ? th_s_unsat_rev(T_ns, p_ns, r_ns, r_ns) : th_s_unsat(T_ns, p_ns, r_ns);   /* :298 */
    const double ss = REV ? th_s_sat_rev(sst, p_surf, rs) : th_s_sat(sst, p_surf);      /* :300 */
    /* get_LCL (thermo.py:107-127) */
    double pLCL;
    {
        const double E0v = 2.3740e6, cvv = 1418, cvl = 4119, cpv = cvv + TH_RV;
        const double q = r_ns / (1 + r_ns);
        const double Rm = (1 - q) * TH_RD + q * TH_RV;
        const double cpm = (1 - q) * TH_CP + q * cpv;
        const double a = cpm / Rm + (cvl - cpv) / TH_RV;
        const double b = -(E0v - (cvv - cvl) * TH_TTRIP) / (TH_RV * T_ns);
        const double cc = b / a;
        const double T_lcl = cc * T_ns / th_lambertw_m1(tcr_pow(rh, 1 / a) * cc * tcr_exp(cc));
        pLCL = p_ns * tcr_pow(T_lcl / T_ns, cpm / Rm);
    }
    /* entropy-axis cells of the two parcels */
    int is_a, is_s;
    double wa0, wa1, wsa0, wsa1;
    int ir_a = 0, ir_s = 0;                    /* reversible table: total-water cells (r_ns for the boundary-layer parcel, rs for the */
    double yr_a = 0.0, yr_s = 0.0;             /* saturated one, thermo.py:346-353), norm distances instead of FITPACK weights         */
    bool ok_a = true, ok_s = true;
    if constexpr (REV) {
        ok_a = th_rgi_cell(A.s_look, ns, s_ns, is_a, wa1);
        ok_a = th_rgi_cell(A.r_look, A.nr, r_ns, ir_a, yr_a) && ok_a;
        ok_s = th_rgi_cell(A.s_look, ns, ss, is_s, wsa1);
        ok_s = th_rgi_cell(A.r_look, A.nr, rs, ir_s, yr_s) && ok_s;
        wa0 = wsa0 = 0.0;
    } else {
        th_locate(A.s_look, ns, s_ns, is_a, wa0, wa1);
        th_locate(A.s_look, ns, ss, is_s, wsa0, wsa1);
    }

    ThParcel pa, ps;
    pa.init(); ps.init();
    bool cond = false;                         /* at or above the first level with pLCL > p (the last level counts, :322) */
    double Tm = 0.0, qm = 0.0;
    float t_next = __ldcs(A.ta + c), q_next = __ldcs(A.hus + c);
    for (int k = 0; k < nlev; ++k) {
        const float t_cur = t_next, q_cur = q_next;
        if (k + 1 < nlev) {                                                              /* next level's line in flight during this level's arithmetic */
            t_next = __ldcs(A.ta + (size_t)(k + 1) * A.n_pts + c);
            q_next = __ldcs(A.hus + (size_t)(k + 1) * A.n_pts + c);
        }
        const ThLevel& L = lev[k];
        const double Te = (double)t_cur, re = (double)q_cur;
        if (k == A.k_mid) { Tm = Te; qm = re; }
        const double Tre = th_t_rho(Te, re);                                             /* :304 */
        cond = cond || (pLCL > L.p) || (k == nlev - 1);
        const double* t0 = A.T_look + (size_t)L.ip * ns;
        double Ta, ra;
        if (REV && cond) {                                                               /* :343-351 */
            Ta = (L.pad && ok_a) ? th_lookup3(A.T_look, ns, A.nr, L.ip, L.wp0, L.wp1, is_a, wa1, ir_a, yr_a) : NAN;
            double es_;
            th_sat(Ta, L.p, es_, ra);
        } else if (cond) {                                                               /* :333-340: moist adiabat from the table */
            const double* r0 = t0 + is_a;
            const double* r1 = r0 + ns;
            double sp = 0.0;
            sp = sp + __ldg(r0) * L.wp0 * wa0;
            sp = sp + __ldg(r0 + 1) * L.wp0 * wa1;
            sp = sp + __ldg(r1) * L.wp1 * wa0;
            sp = sp + __ldg(r1 + 1) * L.wp1 * wa1;
            Ta = sp;
            double es_;
            th_sat(Ta, L.p, es_, ra);
        } else {                                                                         /* :328-330: dry adiabat, constant mixing ratio */
            Ta = T_ns * L.dry;
            ra = r_ns;
        }
        double Ts, rsp;
        if constexpr (REV) {                                                             /* :353 */
            Ts = (L.pad && ok_s) ? th_lookup3(A.T_look, ns, A.nr, L.ip, L.wp0, L.wp1, is_s, wsa1, ir_s, yr_s) : NAN;
            double es_;
            th_sat(Ts, L.p, es_, rsp);
        } else {
            const double* r0 = t0 + is_s;                                                /* :342 */
            const double* r1 = r0 + ns;
            double sp = 0.0;
            sp = sp + __ldg(r0) * L.wp0 * wsa0;
            sp = sp + __ldg(r0 + 1) * L.wp0 * wsa1;
            sp = sp + __ldg(r1) * L.wp1 * wsa0;
            sp = sp + __ldg(r1 + 1) * L.wp1 * wsa1;
            Ts = sp;
            double es_;
            th_sat(Ts, L.p, es_, rsp);                                                   /* :355 */
        }
        pa.level(k, REV ? th_t_rho_rev(Ta, ra, r_ns) : th_t_rho(Ta, ra), Tre, Te, L.mdlnp);     /* :357, 361, 401-402 */
        ps.level(k, REV ? th_t_rho_rev(Ts, rsp, rs) : th_t_rho(Ts, rsp), Tre, Te, L.mdlnp);      /* :358, 362, 403-404 */
    }
    /* outflow level by linear interpolation between `out` and the level above it (:372-396) */
    double T_out_s = NAN, add_a = 0.0, add_s = 0.0;
    if (ps.out >= 0 && ps.out < nlev - 1) {
        const double p1 = lev[ps.out].p, p2 = lev[ps.out + 1].p;
        const double p_out = (p1 * ps.dT2 - p2 * ps.dT1) / (ps.dT2 - ps.dT1);
        T_out_s = (ps.Te1 * (p_out - p2) + ps.Te2 * (p1 - p_out)) / (p1 - p2);
        add_s = TH_RD * ps.dT1 * (p1 - p_out) / (p1 + p_out);
    }
    if (pa.out >= 0 && pa.out < nlev - 1) {
        const double p1 = lev[pa.out].p, p2 = lev[pa.out + 1].p;
        const double p_out = (p1 * pa.dT2 - p2 * pa.dT1) / (pa.dT2 - pa.dT1);
        add_a = TH_RD * pa.dT1 * (p1 - p_out) / (p1 + p_out);
    }
    double cape = (pa.out >= 0 ? pa.cape_at : pa.cape) + add_a;                           /* none qualifies: argmax of all-False = 0 -> the top level */
    const double capes = (ps.out >= 0 ? ps.cape_at : ps.cape) + add_s;
    cape = th_fmax(cape, 0);                                                             /* :408-409 */
    if (cape != cape) cape = 0;
    const double cape_diff = capes - cape;
    double pi = sqrt(th_fmax(A.cecd * (sst / T_out_s) * cape_diff, 0));                  /* :411 */
    if (pi != pi) pi = 0;                                                                /* :412 */
    __stcs(A.vmax + c, pi);

    /* sat_deficit (thermo.py:92-104) clipped to [0, 10] (calc_thermo.py:68); conv_q_to_rh (thermo.py:42-47) */
    /* reversible: all three entropies carry the mid-level mixing ratio as total water (thermo.py:97-101) */
    const double sp_ = REV ? th_s_unsat_rev(Tm, A.p_mid, qm, qm) : th_s_unsat(Tm, A.p_mid, qm);
    const double sps = REV ? th_s_sat_rev(Tm, A.p_mid, qm) : th_s_sat(Tm, A.p_mid);
    const double spss = REV ? th_s_sat_rev(sst, p_surf, qm) : ss;                        /* pseudoadiabatic: s_sat(sst, psl), the same call as :300 */
    __stcs(A.chi + c, th_fmin(th_fmax((sps - sp_) / (spss - sps), 0), 10));
    double es, rsm;
    th_sat(Tm, A.p_mid, es, rsm);
    const double qs = rsm / (1 + rsm);
    __stcs(A.rh_mid + c, th_fmin(th_fmax(qm / qs, 1e-5), 1));
}
