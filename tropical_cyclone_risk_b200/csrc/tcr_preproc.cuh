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
 * Mapping.  A CTA owns 32*VEC consecutive grid points for the whole month and reads every
 * sample of theirs exactly once.  Grouped input: warp (v, k) streams variable v for the days
 * g = k, k+KSPLIT, ..., each lane VEC consecutive points (one 128*VEC-byte row segment per warp load,
 * U independent loads in flight per lane, ld.global.cs).  Ungrouped input (SINGLE): the samples go
 * global -> shared as 16-byte cp.async copies, all in flight at once.  The daily means sit in shared
 * memory [day][variable][point] (float64; float32 samples when SINGLE); after one barrier two
 * threads per point read its day series back from shared memory twice (means, then demeaned products) and
 * write the 14 statistics with coalesced stores.  HBM traffic = the algorithmic minimum: 16 B per (sample, point) read once,
 * 112 B per point written.                                                                    */
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
