/*
 * tcr_kernels.cuh -- the CUDA kernels of the hot path (sm_100a).  See DESIGN.md for the
 * kernel-by-kernel roofline and the thread mapping; reference anchors are given per kernel.
 */
#pragma once
#include "tcr_device.cuh"
#include "tcr_rhs_fast.cuh"

#define TCR_FULL 0xffffffffu

/* ======================================================================================== */
/* table builders: planes -> cell records                                                    */
/* ======================================================================================== */
/* one month: planes [19][nlat][nlon] float32 -> rec [ncy][ncx][20] float4 (corner quads)      */
/* blockIdx.y = month of a batch: planes [n][19][nlat][nlon] -> rec [n][ncy][ncx][20] */
__global__ void k_build_month(const float* __restrict__ planes, float4* __restrict__ rec, int nlat, int nlon)
{
    const int ncx = nlon - 1, ncy = nlat - 1;
    const size_t total = (size_t)ncx * ncy * TCR_REC_F4;
    planes += (size_t)blockIdx.y * TCR_N_FIELDS * nlat * nlon;
    rec += (size_t)blockIdx.y * total;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        int ch = (int)(idx % TCR_REC_F4);
        size_t cell = idx / TCR_REC_F4;
        int ix = (int)(cell % ncx), iy = (int)(cell / ncx);
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ch < TCR_N_FIELDS) {
            const float* p = planes + (size_t)ch * nlat * nlon + (size_t)iy * nlon + ix;
            r = make_float4(p[0], p[nlon], p[1], p[nlon + 1]);
        }
        rec[idx] = r;
    }
}

__global__ void k_build_bathy(const int16_t* __restrict__ src, short4* __restrict__ rec, int nlat, int nlon)
{
    const int ncx = nlon - 1, ncy = nlat - 1;
    const size_t total = (size_t)ncx * ncy;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        int ix = (int)(idx % ncx), iy = (int)(idx / ncx);
        const int16_t* p = src + (size_t)iy * nlon + ix;
        rec[idx] = make_short4(p[0], p[nlon], p[1], p[nlon + 1]);
    }
}

__global__ void k_build_land(const int8_t* __restrict__ src, char4* __restrict__ rec, int nlat, int nlon)
{
    const int ncx = nlon - 1, ncy = nlat - 1;
    const size_t total = (size_t)ncx * ncy;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        int ix = (int)(idx % ncx), iy = (int)(idx / ncx);
        const int8_t* p = src + (size_t)iy * nlon + ix;
        rec[idx] = make_char4(p[0], p[nlon], p[1], p[nlon + 1]);
    }
}

/* masks u8 [8][nlat][nlon] -> rec [ncy][ncx][4 corners] of 8 bytes (byte b = mask plane b)    */
__global__ void k_build_masks(const uint8_t* __restrict__ src, uint2* __restrict__ rec, int nlat, int nlon)
{
    const int ncx = nlon - 1, ncy = nlat - 1;
    const size_t total = (size_t)ncx * ncy * 4;
    const size_t plane = (size_t)nlat * nlon;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        int corner = (int)(idx & 3);
        size_t cell = idx >> 2;
        int ix = (int)(cell % ncx) + (corner >> 1), iy = (int)(cell / ncx) + (corner & 1);
        uint32_t lo = 0, hi = 0;
        for (int b = 0; b < 4; ++b) lo |= (uint32_t)src[b * plane + (size_t)iy * nlon + ix] << (8 * b);
        for (int b = 0; b < 4; ++b) hi |= (uint32_t)src[(4 + b) * plane + (size_t)iy * nlon + ix] << (8 * b);
        rec[idx] = make_uint2(lo, hi);
    }
}

__global__ void k_build_axis(const double* __restrict__ ax, TcrNode* __restrict__ out, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        TcrNode nd;
        nd.x = ax[i];
        nd.inv = i + 1 < n ? 1.0 / (ax[i + 1] - ax[i]) : 0.0;
        nd.x1 = i + 1 < n ? ax[i + 1] : ax[i];
        nd.pad = 0.0;
        out[i] = nd;
    }
}

/* tcr_rcp_seed of the constant divisors of the RHS (out[0] = earth_R, out[1] = pi); out[2], out[3] = sine of the genesis
 * latitude bounds -- the same tcr_sin every seeding thread used to evaluate for itself */
__global__ void k_build_consts(double earth_R, double gen_lat_min, double gen_lat_max, double* __restrict__ out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        out[0] = tcr_rcp_seed(earth_R); out[1] = tcr_rcp_seed(TCR_PI);
        out[2] = tcr_sin(TCR_DEG2RAD * gen_lat_min); out[3] = tcr_sin(TCR_DEG2RAD * gen_lat_max);
    }
}

/* storm-independent harmonics of the output time grid: sc[j][k] = {sin, cos}(2 pi (k+1) t_j / T_Fs) */
__global__ void k_build_sincos(const __grid_constant__ TcrCtx cx, double2* __restrict__ sc, double2* __restrict__ sct)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cx.p.n_steps) return;
    double2 h[TCR_N_HARM];
    tcr_harmonics(cx, tcr_node_time(cx, j), h);
#pragma unroll
    for (int k = 0; k < TCR_N_HARM; ++k) {
        sc[(size_t)j * TCR_N_HARM + k] = h[k];
        sct[(size_t)k * cx.p.n_steps + j] = h[k];
    }
}

/* ======================================================================================== */
/* stand-alone bilinear sampler (the HBM-roofline kernel)                                     */
/* replaces RectBivariateSpline(kx=1,ky=1).ev on every field: util/mat.py:142-153,            */
/* bam_track.py:93-108, coupled_fast.py:35-58,125-126.  out [n][21] float64.                   */
/* ======================================================================================== */
#define EI_TILE 256
#define EI_ITEMS (EI_TILE * TCR_N_INTERP_OUT)

struct EiLoc { const float4* rec; double w00, w01, w10, w11, bathy, land; };

/* variant 0: per-lane LDG.128 of the record quads, flat (query, channel) item space so that
 * both the record reads and the float64 output writes are fully coalesced                    */
template <int MINB>
__global__ void __launch_bounds__(EI_TILE, MINB) k_env_interp(const __grid_constant__ TcrCtx cx, int64_t n,
                                                        const int32_t* __restrict__ ym, const double* __restrict__ lon,
                                                        const double* __restrict__ lat, double* __restrict__ out)
{
    __shared__ EiLoc loc[EI_TILE];
    const int tid = threadIdx.x;
    const int64_t q0 = (int64_t)blockIdx.x * EI_TILE;
    const int nq = (int)min((int64_t)EI_TILE, n - q0);
    if (tid < nq) {
        const int64_t q = q0 + tid;
        const double x = lon[q], y = lat[q];
        TcrCell c;
        tcr_cell_at(cx.tab.lon, cx.tab.lat, x, y, c);
        EiLoc l;
        l.rec = tcr_record(cx.tab, ym[q], c);
        l.w00 = c.w00; l.w01 = c.w01; l.w10 = c.w10; l.w11 = c.w11;
        l.bathy = tcr_bathy_at(cx.st, x, y);
        l.land = tcr_land_at(cx.st, x, y);
        loc[tid] = l;
    }
    __syncthreads();
    double* o = out + q0 * TCR_N_INTERP_OUT;
    const int n_items = nq * TCR_N_INTERP_OUT;
    /* 21 items per thread, in 3 batches of 7 independent loads */
#pragma unroll 1
    for (int b = 0; b < 3; ++b) {
        float4 r[7];
        int qi[7], ch[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            int i = tid + (b * 7 + j) * EI_TILE;
            qi[j] = i / TCR_N_INTERP_OUT;
            ch[j] = i - qi[j] * TCR_N_INTERP_OUT;
            r[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < n_items && ch[j] < TCR_N_FIELDS) r[j] = __ldg(loc[qi[j]].rec + ch[j]);
        }
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            int i = tid + (b * 7 + j) * EI_TILE;
            if (i < n_items) {
                const EiLoc& l = loc[qi[j]];
                double v;
                if (ch[j] < TCR_N_FIELDS)
                    v = fma((double)r[j].w, l.w11, fma((double)r[j].z, l.w10, fma((double)r[j].y, l.w01, (double)r[j].x * l.w00)));
                else
                    v = (ch[j] == TCR_N_FIELDS) ? l.bathy : l.land;
                __stcs(o + i, v);
            }
        }
    }
}

/* (Variants 1 and 4 of round 1 -- one 320-byte cp.async.bulk per query behind an mbarrier, and a warp-specialised
 * cp.async + mbarrier ring -- measured 3.6 x and 2.8 x slower than variant 0 and were removed in round 2: a per-record
 * bulk copy is too small to amortise its issue cost, and a tile design needs cell-sorted queries whose 168-byte
 * result rows then scatter.  DESIGN.md section 4.) */
__device__ __forceinline__ void tcr_cp_async16(void* dst, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

/* variant 5: one tile of EA_TILE queries per CTA, all of the tile's record quads fetched by
 * cp.async (LDGSTS, no register staging) in one burst -- EA_TILE x 320 B in flight per CTA, the
 * other resident CTAs of the SM computing meanwhile -- then the outputs are formed from shared
 * memory with a fixed channel per thread (252 = 12 queries x 21 channels of the 256 threads are
 * active per pass, item = pass * 252 + t, so loads, stores and the smem reads are all contiguous
 * and no per-item index arithmetic is left).                                                  */
template <int EA_TILE>
__global__ void __launch_bounds__(256) k_env_interp_async(const __grid_constant__ TcrCtx cx, int64_t n,
                                                          const int32_t* __restrict__ ym, const double* __restrict__ lon,
                                                          const double* __restrict__ lat, double* __restrict__ out)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* rec = reinterpret_cast<float4*>(smem_raw);                                        /* [EA_TILE][20] */
    EiLoc* loc = reinterpret_cast<EiLoc*>(smem_raw + (size_t)EA_TILE * TCR_REC_F4 * 16);       /* [EA_TILE]     */
    const int tid = threadIdx.x;
    const int64_t q0 = (int64_t)blockIdx.x * EA_TILE;
    const int nq = (int)min((int64_t)EA_TILE, n - q0);
    const int lane = tid & 31;
    for (int base = 0; base < nq; base += 256) {
        const int qq = base + tid;
        const bool valid = qq < nq;
        const float4* my_rec = nullptr;
        TcrCellLoc ll, lb;
        TcrCell c;
        double x = 0.0, y = 0.0;
        if (valid) {
            const int64_t q = q0 + qq;
            x = lon[q]; y = lat[q];
            TcrCellLoc lt;
            tcr_cell_begin(cx.tab.lon, cx.tab.lat, x, y, lt);
            tcr_cell_begin(cx.st.lon_l, cx.st.lat_l, x, y, ll);
            tcr_cell_begin(cx.st.lon_b, cx.st.lat_b, x, y, lb);
            tcr_cell_end(cx.tab.lon, cx.tab.lat, lt, c);
            my_rec = tcr_record(cx.tab, ym[q], c);
        }
        /* the warp's 32 records, one coalesced 320-byte asynchronous copy each (lanes 0..19),
         * issued before the static grids are touched */
        const unsigned long long rp = (unsigned long long)my_rec;
        const int wq0 = base + (tid & ~31);
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const unsigned long long src = __shfl_sync(TCR_FULL, rp, k);
            if (src != 0ull && lane < TCR_REC_F4)
                tcr_cp_async16(rec + (wq0 + k) * TCR_REC_F4 + lane, reinterpret_cast<const float4*>(src) + lane);
        }
        if (valid) {
            TcrCell cl, cb;
            tcr_cell_end(cx.st.lon_l, cx.st.lat_l, ll, cl);
            tcr_cell_end(cx.st.lon_b, cx.st.lat_b, lb, cb);
            EiLoc l;
            l.rec = my_rec;
            l.w00 = c.w00; l.w01 = c.w01; l.w10 = c.w10; l.w11 = c.w11;
            l.bathy = tcr_bathy_cell(cx.st, cb);
            l.land = tcr_land_cell(cx.st, cl);
            loc[qq] = l;
        }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    if (tid < 252) {
        const int ch = tid % TCR_N_INTERP_OUT, qs = tid / TCR_N_INTERP_OUT;        /* fixed channel, query sub-index 0..11 */
        double* o = out + q0 * TCR_N_INTERP_OUT + tid;
#pragma unroll 4
        for (int qq = qs; qq < nq; qq += 12, o += 252) {
            const EiLoc& l = loc[qq];
            double v;
            if (ch < TCR_N_FIELDS) {
                const float4 r = rec[qq * TCR_REC_F4 + ch];
                v = fma((double)r.w, l.w11, fma((double)r.z, l.w10, fma((double)r.y, l.w01, (double)r.x * l.w00)));
            } else {
                v = (ch == TCR_N_FIELDS) ? l.bathy : l.land;
            }
            __stcs(o, v);
        }
    }
}

/* ======================================================================================== */
/* Fourier coefficients of a wave's storms                                                    */
/* gen_f phases (bam_track.py:27) -> {A, B} = amp_k {cos, sin}(2 pi x): the angle-addition   */
/* form of bam_track.py:28-31.  coef [n][60] double2.                                         */
/* ======================================================================================== */
__global__ void k_coef_from_phases(const __grid_constant__ TcrCtx cx, int64_t n, const double* __restrict__ phases,
                                   double2* __restrict__ coef)
{
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * TCR_N_PHASES) return;
    int pair = (int)(idx % TCR_N_PHASES);
    double s, c;
    tcr_sincos2pi(phases[idx], &s, &c);
    double amp = cx.p.fourier_amp[pair % TCR_N_HARM];
    coef[idx] = make_double2(amp * c, amp * s);
}

__global__ void k_coef_from_philox(const __grid_constant__ TcrCtx cx, const unsigned int* __restrict__ n_slots,
                                   const int64_t* __restrict__ slot_att, const int32_t* __restrict__ slot_key,
                                   uint32_t run_seed, double2* __restrict__ coef)
{
    /* thread = one Philox block = the two phases (2 j, 2 j + 1) of a slot: 30 threads per slot, 32 bytes written each */
    constexpr int HALF = TCR_N_PHASES / 2;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t slot = idx / HALF;
    if (slot >= (int64_t)*n_slots) return;
    const int j = (int)(idx - slot * HALF);
    double u[2];
    tcr_draw2(run_seed, slot_key[slot], slot_att[slot], (uint32_t)j, 1u, u);
    double2 out[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        double s, c;
        tcr_sincos2pi(u[e], &s, &c);
        const double amp = cx.p.fourier_amp[(2 * j + e) % TCR_N_HARM];
        out[e] = make_double2(amp * c, amp * s);
    }
    double2* dst = coef + slot * TCR_N_PHASES + 2 * j;
    dst[0] = out[0]; dst[1] = out[1];
}

/* ======================================================================================== */
/* Fourier tables of a wave's storms: gen_f (bam_track.py:23-31) on the output time grid       */
/* ftab [n][n_steps][4].  A small dense contraction per storm (n_steps x 30 by 30 x 4) done on */
/* the fp64 pipe in the exact operation order of the specification (series-by-series fma       */
/* chains over the harmonics), so it is bit-identical to evaluating the series in the RHS.     */
/* Thread = one time node (its 15 {sin, cos} pairs stay in registers); the CTA walks a tile of */
/* storms whose coefficients sit in shared memory (broadcast reads); stores are coalesced:     */
/* adjacent threads write adjacent 32-byte nodes.                                              */
/* ======================================================================================== */
#define FT_THREADS 96
#define FT_NODES 2                     /* output nodes per thread: every coefficient fetched from shared memory feeds 2 x 2 DFMA */
#define FT_STORMS 16
__global__ void __launch_bounds__(FT_THREADS) k_fourier_table(const __grid_constant__ TcrCtx cx, int64_t n,
                                                              const unsigned int* __restrict__ n_dev,
                                                              const double2* __restrict__ coef, double* __restrict__ ftab)
{
    __shared__ __align__(16) double2 cfs[2][FT_STORMS][TCR_N_PHASES];      /* double-buffered coefficient tiles */
    const int64_t count = n_dev ? (int64_t)*n_dev : n;
    const int ns = cx.p.n_steps;
    int jn[FT_NODES];
    double2 sc[FT_NODES][TCR_N_HARM];
#pragma unroll
    for (int u = 0; u < FT_NODES; ++u) {
        jn[u] = blockIdx.x * (FT_THREADS * FT_NODES) + u * FT_THREADS + threadIdx.x;
        const int jj = min(jn[u], ns - 1);                                 /* out-of-range nodes compute a duplicate, never store */
#pragma unroll
        for (int k = 0; k < TCR_N_HARM; ++k) sc[u][k] = __ldg(cx.sc + (size_t)jj * TCR_N_HARM + k);
    }
    const int64_t stride = (int64_t)gridDim.y * FT_STORMS;
    /* asynchronous copy (LDGSTS) of one tile's coefficients: the next tile lands while this one is used */
    auto prefetch = [&](int64_t s0, int buf) {
        if (s0 < count) {
            const int nst = (int)min((int64_t)FT_STORMS, count - s0);
            for (int i = threadIdx.x; i < nst * TCR_N_PHASES; i += FT_THREADS)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(&cfs[buf][0][0] + i)),
                             "l"(coef + (size_t)s0 * TCR_N_PHASES + i) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int64_t s0 = (int64_t)blockIdx.y * FT_STORMS;
    int buf = 0;
    prefetch(s0, 0);
    for (; s0 < count; s0 += stride, buf ^= 1) {
        prefetch(s0 + stride, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        const int nst = (int)min((int64_t)FT_STORMS, count - s0);
        if (jn[0] < ns) {
#pragma unroll 1
            for (int st = 0; st < nst; ++st) {
                double F[FT_NODES][4];
#pragma unroll
                for (int u = 0; u < FT_NODES; ++u) { F[u][0] = F[u][1] = F[u][2] = F[u][3] = 0.0; }
#pragma unroll
                for (int k = 0; k < TCR_N_HARM; ++k) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const double2 ab = cfs[buf][st][i * TCR_N_HARM + k];
#pragma unroll
                        for (int u = 0; u < FT_NODES; ++u) {
                            F[u][i] = fma(ab.x, sc[u][k].x, F[u][i]);
                            F[u][i] = fma(ab.y, sc[u][k].y, F[u][i]);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < FT_NODES; ++u) {
                    if (jn[u] < ns) {
                        double2* dst = reinterpret_cast<double2*>(ftab + ((size_t)(s0 + st) * ns + jn[u]) * 4);
                        __stcs(dst, make_double2(F[u][0], F[u][1]));
                        __stcs(dst + 1, make_double2(F[u][2], F[u][3]));
                    }
                }
            }
        }
        __syncthreads();                       /* tile `buf` is free for the prefetch of the iteration after next */
    }
}

/* ---- the same tables on the FP64 tensor cores -------------------------------------------------
 * Per storm the tabulation is the dense product  F[n_steps x 4] = SC[n_steps x 30] . CF[30 x 4]  with
 * SC[j][2k], SC[j][2k+1] = {sin, cos} of harmonic k at node j and CF[2k][i], CF[2k+1][i] the storm's
 * coefficient pair of series i -- and the specification's summation order (one fma chain per (node,
 * series) over K = 0..29) is exactly what DMMA.8x8x4 computes: measured on B200, mma.sync.m8n8k4.f64
 * equals d = fma(a3,b3, fma(a2,b2, fma(a1,b1, fma(a0,b0,c)))) bit for bit (scripts/probes/dmma_order.cu,
 * 1.28 M elements), and chaining eight of them through the accumulator continues the chain (K padded
 * from 30 to 32 with zero terms, which leave the sum unchanged).  One mma covers 8 nodes x (2 storms x
 * 4 series) x 4 K.  Warp w of the CTA owns node groups of 32 (four 8-node tiles whose A fragments --
 * 32 doubles per lane -- stay in registers; for n_steps <= 384 once for the whole kernel), the CTA walks
 * double-buffered tiles of 16 storms whose coefficients arrive by cp.async; per storm pair a lane reads
 * its 8 B-fragment doubles from shared memory, issues 32 DMMA and writes four 16-byte pieces of
 * ftab[storm][node][series pair] (a warp store covers two 256-byte runs).                          */
#define FTM_WARPS 12
#define FTM_THREADS (FTM_WARPS * 32)
__global__ void __launch_bounds__(FTM_THREADS, 1) k_fourier_table_mma(const __grid_constant__ TcrCtx cx, int64_t n,
                                                                      const unsigned int* __restrict__ n_dev,
                                                                      const double2* __restrict__ coef, double* __restrict__ ftab)
{
    __shared__ __align__(16) double2 cfs[2][FT_STORMS][TCR_N_PHASES];
    const int64_t count = n_dev ? (int64_t)*n_dev : n;
    const int ns = cx.p.n_steps;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = lane >> 2, kq = lane & 3;                   /* A: node row / K offset;  B: column n = row, K offset kq */
    const int n_jg = (ns + 31) >> 5;                            /* node groups of 32 */
    const bool hoist = n_jg <= FTM_WARPS;                       /* each warp has at most one group: its A fragments are loaded once */
    const double* scd = reinterpret_cast<const double*>(cx.sc); /* [n_steps][15][2] = SC[j][K], K = 0..29 */

    double a[4][8];
    auto load_a = [&](int jg) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int j = min(jg * 32 + t * 8 + row, ns - 1);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int K = c * 4 + kq;
                a[t][c] = K < 2 * TCR_N_HARM ? __ldg(scd + (size_t)j * (2 * TCR_N_HARM) + K) : 0.0;
            }
        }
    };
    if (hoist && warp < n_jg) load_a(warp);

    auto prefetch = [&](int64_t s0, int buf) {
        if (s0 < count) {
            const int nst = (int)min((int64_t)FT_STORMS, count - s0);
            for (int i = threadIdx.x; i < nst * TCR_N_PHASES; i += FTM_THREADS)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(&cfs[buf][0][0] + i)),
                             "l"(coef + (size_t)s0 * TCR_N_PHASES + i) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int64_t stride = (int64_t)gridDim.x * FT_STORMS;
    int64_t s0 = (int64_t)blockIdx.x * FT_STORMS;
    int buf = 0;
    prefetch(s0, 0);
    for (; s0 < count; s0 += stride, buf ^= 1) {
        prefetch(s0 + stride, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        const int nst = (int)min((int64_t)FT_STORMS, count - s0);
        for (int jg = warp; jg < n_jg; jg += FTM_WARPS) {
            if (!hoist) load_a(jg);
            for (int sp = 0; sp * 2 < nst; ++sp) {
                /* B fragment: column n = row -> storm 2 sp + (n >> 2), series n & 3; element K of that column is
                 * double number K of the series' 15 coefficient pairs */
                const int st_b = min(2 * sp + (row >> 2), nst - 1);
                const double* cfd = reinterpret_cast<const double*>(&cfs[buf][st_b][(row & 3) * TCR_N_HARM]);
                double b[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int K = c * 4 + kq;
                    b[c] = K < 2 * TCR_N_HARM ? cfd[K] : 0.0;
                }
                double d[4][2];
#pragma unroll
                for (int t = 0; t < 4; ++t) { d[t][0] = 0.0; d[t][1] = 0.0; }
#pragma unroll
                for (int c = 0; c < 8; ++c) {
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                     : "+d"(d[t][0]), "+d"(d[t][1]) : "d"(a[t][c]), "d"(b[c]));
                }
                /* D: row = node, columns 2 kq, 2 kq + 1 -> storm 2 sp + (kq >> 1), series 2 (kq & 1) + {0, 1} */
                const int st_d = 2 * sp + (kq >> 1);
                if (st_d < nst) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int j = jg * 32 + t * 8 + row;
                        if (j < ns)
                            __stcs(reinterpret_cast<double2*>(ftab + ((size_t)(s0 + st_d) * ns + j) * 4 + 2 * (kq & 1)),
                                   make_double2(d[t][0], d[t][1]));
                    }
                }
            }
        }
        __syncthreads();                       /* tile `buf` is free for the prefetch of the iteration after next */
    }
}

/* ======================================================================================== */
/* the integrator: Coupled_FAST.gen_track (coupled_fast.py:229-267) incl. scipy's RK45        */
/* driver loop, t_eval dense output and the terminal event                                    */
/* ======================================================================================== */
/* One segment of a storm's Fourier ring: nodes [j0, j0 + seg) of its four series from its 60 coefficient pairs, by the 32
 * lanes of the calling warp (lane L takes nodes j0 + L, j0 + 32 + L of every 64-node pass) -- the fma chains of
 * k_fourier_table, so the bits of the full tables.  Called warp-uniformly from the integrator's macro-step boundary.
 * stage != 0: shared-space address of six warp-private rows of 32 doubles (pitch bytes apart) that hold the coefficient pairs
 * during the fill (pair e at row e / 16, column pair e % 16): every lane then reads them as broadcast LDS instead of keeping
 * 60 global loads in flight.  Not inlined: the integrator's register allocation (168, no spills) stays what it was.          */
__device__ __noinline__ void tcr_ring_fill(const TcrCtx& cx, const double2* __restrict__ coef, double* __restrict__ ring_row,
                                           int j0, int seg, int mask, uint32_t stage, uint32_t pitch)
{
    const int lane = threadIdx.x & 31;
    const int ns = cx.p.n_steps;
    if (stage) {
#pragma unroll
        for (int e = lane; e < TCR_N_PHASES; e += 32) {
            const double2 cf = __ldg(coef + e);
            asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(stage + (uint32_t)(e >> 4) * pitch + (uint32_t)(e & 15) * 16u), "d"(cf.x), "d"(cf.y) : "memory");
        }
        __syncwarp();
    }
    for (int jb = j0; jb < j0 + seg; jb += 64) {
        const int ja = jb + lane, jc = jb + 32 + lane;
        /* harmonic-major copy of the {sin, cos} table: the 32 lanes' nodes are 512 contiguous bytes per harmonic (the
         * node-major rows are 240 B apart: 32 L1 wavefronts per load instead of 4, and the L1 data pipe is the integrator's
         * busiest unit) */
        const double2* sa = cx.sct + min(ja, ns - 1);
        const double2* sb = cx.sct + min(jc, ns - 1);
        double Fa[4] = {0.0, 0.0, 0.0, 0.0}, Fb[4] = {0.0, 0.0, 0.0, 0.0};
        if (stage) {
#pragma unroll
            for (int k = 0; k < TCR_N_HARM; ++k) {
                const double2 a = __ldg(sa + (size_t)k * ns), b = __ldg(sb + (size_t)k * ns);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int e = i * TCR_N_HARM + k;
                    double2 ab;
                    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(ab.x), "=d"(ab.y) : "r"(stage + (uint32_t)(e >> 4) * pitch + (uint32_t)(e & 15) * 16u));
                    Fa[i] = fma(ab.x, a.x, Fa[i]); Fa[i] = fma(ab.y, a.y, Fa[i]);
                    Fb[i] = fma(ab.x, b.x, Fb[i]); Fb[i] = fma(ab.y, b.y, Fb[i]);
                }
            }
        } else {
#pragma unroll 1
            for (int k = 0; k < TCR_N_HARM; ++k) {
                const double2 a = __ldg(sa + (size_t)k * ns), b = __ldg(sb + (size_t)k * ns);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const double2 ab = __ldg(coef + i * TCR_N_HARM + k);
                    Fa[i] = fma(ab.x, a.x, Fa[i]); Fa[i] = fma(ab.y, a.y, Fa[i]);
                    Fb[i] = fma(ab.x, b.x, Fb[i]); Fb[i] = fma(ab.y, b.y, Fb[i]);
                }
            }
        }
        if (ja < ns && ja < j0 + seg) {
            double2* d = reinterpret_cast<double2*>(ring_row + (size_t)(ja & mask) * 4);
            d[0] = make_double2(Fa[0], Fa[1]); d[1] = make_double2(Fa[2], Fa[3]);
        }
        if (jc < ns && jc < j0 + seg) {
            double2* d = reinterpret_cast<double2*>(ring_row + (size_t)(jc & mask) * 4);
            d[0] = make_double2(Fb[0], Fb[1]); d[1] = make_double2(Fb[2], Fb[3]);
        }
    }
    __syncwarp();                              /* the staging rows and the ring nodes are the warp's to read / overwrite again */
}

struct IntegArgs {
    int64_t n;                         /* number of storms (slots) if n_dev == NULL           */
    const unsigned int* n_dev;         /* device-side count (run_years)                       */
    const int32_t* ym;                 /* [n] table index                                     */
    const double* lon0; const double* lat0; const double* v0; const double* m0; const double* h_bl;
    const double* ftab;                /* [n][n_steps][4] Fourier tables (k_fourier_table); NULL with a ring */
    /* Fourier RING (tcr_run_years): instead of a full table per integrated seed, every track-pool row owns a ring of
     * ring_nodes = 2 segments of nodes that the integrator fills on demand from the storm's coefficients, one segment ahead
     * of the storm's clock -- storms use a quarter to a third of their table, and the table was 90 % of the per-slot workspace.
     * A segment is longer than the nodes one RK attempt can span (max_step / t_step + 4: the host checks it), so an attempt
     * never needs more than the two resident segments.                                                                      */
    const double2* coef; double* ring; int ring_nodes;
    int ring_cta;                      /* lock-stepped variants: requests served by the whole CTA (1) or by the lane's own warp (0) */
    double* track;                     /* [n][n_steps][4] lon,lat,v,m -- or, with track_row, a POOL of rows  */
    /* track pool (tcr_run_years): a storm writes its samples into the row its lane holds; a TC candidate
     * keeps the row (track_row[slot] = row) and the lane draws a fresh one, every other storm's row is
     * reused by the lane's next storm -- so rows exist for the lanes in flight plus the candidates (2-3 %
     * of the storms) instead of for every integrated seed.  pool_ctl[0] = fresh rows handed out so far (row = pool_first + that: rows
     * below pool_first = the launch's lane count belong to the lanes), pool_ctl[1] = overflow flag (the host re-runs the wave with fewer slots).  */
    int32_t* track_row; unsigned int* pool_ctl; unsigned int pool_rows; unsigned int pool_first;
    int32_t* n_time; int32_t* status; int32_t* nfev; uint32_t* flags;
    unsigned long long* queue;         /* work counter, zeroed by the host                    */
    int32_t* cand_list; unsigned int* cand_count;   /* TC candidates (NULL: not collected)    */
    int lane_cap;                      /* lanes per warp that take storms (small batches)     */
    int pack;                          /* drain-phase packing of the surviving storms on/off  */
};

enum { M_IDLE = 0, M_INIT0 = 1, M_INIT1 = 2, M_WAIT = 3, M_RK = 4 };

/* Thread mapping: ONE STORM PER LANE, warps advance in lock-step MACRO STEPS of six RHS
 * evaluations -- exactly one RK45 attempt (5 stages + the FSAL evaluation), accepted or
 * rejected -- so the expensive RHS is always executed convergently and only the cheap
 * per-stage bookkeeping is predicated.  A lane whose storm ended pops the next storm from
 * the global queue and spends its macro step on the two initial-step evaluations
 * (select_initial_step); the persistent loop ends when the queue is drained.  Storm state
 * (y, K[7][4], step control) lives in registers for the storm's lifetime; the storm's Fourier
 * table is read from HBM/L2 (64 B per evaluation); only emitted samples go to HBM.
 * THREADS x MINB fixes the register budget (launch bounds): <256,1> 255 registers, <128,3> 168,
 * <128,4> 128 -- chosen at run time by tcr_set_tuning, default by measurement.   */
template <int THREADS, int MINB, int KSMEM, int LOCKSTEP_FAST>
__global__ void __launch_bounds__(THREADS, MINB) k_integrate(const __grid_constant__ TcrCtx cx, const IntegArgs A)
{
    /* LOCKSTEP_FAST = slot mask of the CTA-wide re-alignment barriers (bits 0-5) + 256 for the straight-line RHS */
    constexpr int CTA_LOCKSTEP = LOCKSTEP_FAST & 255;
    constexpr bool FAST_RHS = (LOCKSTEP_FAST & 256) != 0;
    /* Stage storage, "stage" j = 0..7: K0 (FSAL derivative), K1..K5, K6, and the step's end state y_new.
     * KSMEM = 1: K1..K5 (dead during an RHS evaluation) live in shared memory, [stage][component][thread],
     * which frees 40 registers; KSMEM = 2: all eight (64 registers); KSMEM = 0: registers only. */
    extern __shared__ __align__(16) double k_smem[];
    double Kr[8][4] = {};
    double* const ks = k_smem + (KSMEM ? threadIdx.x : 0);
    bool cta_drained = false;          /* some thread of the CTA has seen the queue empty (refreshed at the slot-1 barrier) */
    auto in_smem = [](int j) { return KSMEM == 2 || (KSMEM == 1 && j >= 1 && j <= 5); };
    auto k_off = [](int j, int i) { return ((KSMEM == 2 ? j : j - 1) * 4 + i) * THREADS; };
    auto Kg = [&](int j, int i) -> double { return in_smem(j) ? ks[k_off(j, i)] : Kr[j][i]; };
    auto Ks = [&](int j, int i, double v) { if (in_smem(j)) ks[k_off(j, i)] = v; else Kr[j][i] = v; };
    const tcr_params& p = cx.p;
    const int lane = threadIdx.x & 31;
    const double* ftab = nullptr;       /* this storm's full table, or its row's ring */
    const int fmask = A.ring ? A.ring_nodes - 1 : TCR_FTAB_FULL;
    int have = 0;                       /* ring: nodes [0, have) of this storm have been tabulated */
    const int64_t n = A.n_dev ? (int64_t)*A.n_dev : A.n;
    const int ns = p.n_steps;
    const double t_bound = p.total_time, rtol = p.rtol, atol = p.atol, max_step = p.max_step;

    int mode = M_IDLE;
    bool drained = false;
    int64_t sid = -1;
    int ym = 0, nfev = 0, n_out = 0, status = 0, n_attempts = 0;
    bool rejected = false, new_step = false, any_v = false;
    /* Step-control state.  With KSMEM = 2 it lives in shared memory (the rows of the drain-phase staging area are each
     * thread's "home" slots): it is touched a few times per RK attempt but would otherwise hold 24 registers across every
     * RHS evaluation, where the register file is the scarce resource (168 per thread at 12 warps per SM).
     * home rows: 0-3 y, 4 t, 5 h_abs, 6 g, 7 min_step, 8 h, 9 t_new, 10 h0, 11 d1; rows 12-17 are the packing mailbox */
    constexpr bool HOMES = (KSMEM == 2);
    double* const hs = k_smem + (HOMES ? 32 * THREADS + (int)threadIdx.x : 0);
    double hbl = 0.0;
    double r_y[4] = {0, 0, 0, 0}, r_t = 0.0, r_h = 0.0, r_h_abs = 0.0, r_t_new = 0.0, r_g = 0.0, r_min_step = 0.0, r_h0 = 0.0, r_d1 = 0.0;
    auto Y = [&](int i) -> double& { return HOMES ? hs[i * THREADS] : r_y[i]; };
    double& t = HOMES ? hs[4 * THREADS] : r_t;
    double& h_abs = HOMES ? hs[5 * THREADS] : r_h_abs;
    double& g = HOMES ? hs[6 * THREADS] : r_g;
    double& min_step = HOMES ? hs[7 * THREADS] : r_min_step;
    double& h = HOMES ? hs[8 * THREADS] : r_h;
    double& t_new = HOMES ? hs[9 * THREADS] : r_t_new;
    double& h0 = HOMES ? hs[10 * THREADS] : r_h0;
    double& d1 = HOMES ? hs[11 * THREADS] : r_d1;
    if (HOMES) {
#pragma unroll
        for (int i = 0; i < 12; ++i) hs[i * THREADS] = 0.0;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) Ks(j, i, 0.0);
    double* trk = nullptr;
    /* track pool: the row this lane writes (its storm's while active, its spare while idle) */
    unsigned int row = (unsigned int)(blockIdx.x * THREADS + threadIdx.x);
    const size_t row_doubles = (size_t)ns * 4;

    /* storm end: n_time / status / nfev / TC criteria (util/compute.py:185-189) */
    auto finalize = [&](int st) {
        A.n_time[sid] = n_out;
        A.status[sid] = st;
        A.nfev[sid] = (st == TCR_STATUS_VENT) ? 0 : nfev;
        uint32_t fl = 0;
        if (st != TCR_STATUS_VENT && n_out > 0) {
            const double t2 = 2.0 * 24 * 60 * 60;
            double v2d;
            if (t2 >= tcr_node_time(cx, n_out - 1)) {
                v2d = trk[(size_t)(n_out - 1) * 4 + 2];
            } else {
                int j = tcr_nodes_le(cx, t2, 1) - 1;
                if (j > n_out - 2) j = n_out - 2;
                double x0 = tcr_node_time(cx, j), x1 = tcr_node_time(cx, j + 1);
                double vj = trk[(size_t)j * 4 + 2], vj1 = trk[(size_t)(j + 1) * 4 + 2];
                double slope = (vj1 - vj) / (x1 - x0);
                v2d = slope * (t2 - x0) + vj;
            }
            if (any_v && v2d >= p.seed_v_2d_thresh) fl = TCR_FLAG_IS_TC;
        }
        A.flags[sid] = fl;
        if (fl && A.cand_list) A.cand_list[atomicAdd(A.cand_count, 1u)] = (int32_t)sid;
        if (A.track_row) {
            A.track_row[sid] = fl ? (int32_t)row : -1;
            if (fl) {                                            /* the candidate keeps its row: take a fresh one */
                const unsigned int fresh = A.pool_first + atomicAdd(A.pool_ctl, 1u);
                if (fresh < A.pool_rows) row = fresh; else A.pool_ctl[1] = 1u;
            }
        }
        mode = M_IDLE;
    };

    /* tabulate ring segments until node need_idx of the lane's storm is resident (-1: nothing needed).  The warp serves its
     * requests one at a time, all 32 lanes on one storm's segment.  Residency: after a fill, nodes [have - ring_nodes, have)
     * are in the ring and have - seg <= need_idx; the lowest node a later evaluation of the same storm can ask for is the
     * lower node of its clock, >= need_idx - 2 - (nodes one attempt spans) -- inside the ring because seg exceeds that span
     * by 4 or more.  Callers pass need_idx one node high (ring_need: a stage time t + c h may round one ulp past t_new,
     * onto a node). */
    uint32_t ring_stage = 0;
    if (HOMES) ring_stage = (uint32_t)__cvta_generic_to_shared(k_smem + (32 + 12) * THREADS + ((int)threadIdx.x & ~31));
    /* CTA-wide service (lock-stepped variants): the lanes post their requests to a list in shared memory and the CTA's warps
     * take them round-robin -- a warp that had to serve its own lanes one after the other kept its five partner warps
     * waiting at the next slot barrier for as long as the unluckiest warp's requests took */
    constexpr bool CTA_SERVE = HOMES && CTA_LOCKSTEP != 0;
    __shared__ int rq_n[2], rq_par;
    __shared__ int4 rq[CTA_SERVE ? THREADS : 1];
    if (CTA_SERVE && A.ring) {
        if (threadIdx.x < 2) rq_n[threadIdx.x] = 0;
        if (threadIdx.x == 2) rq_par = 0;
        __syncthreads();
    }
    auto ring_need = [&](int idx) { return idx < 0 ? idx : min(idx + 1, ns - 1); };
    auto ring_serve = [&](int need_idx) {
        for (;;) {
            const unsigned req = __ballot_sync(TCR_FULL, need_idx >= have);
            if (!req) break;
            const int src = __ffs((int)req) - 1;
            const long long sid_s = __shfl_sync(TCR_FULL, (long long)sid, src);
            const int have_s = __shfl_sync(TCR_FULL, have, src);
            const unsigned int row_s = __shfl_sync(TCR_FULL, row, src);
            const int seg = A.ring_nodes >> 1;
            tcr_ring_fill(cx, A.coef + (size_t)sid_s * TCR_N_PHASES, A.ring + (size_t)row_s * A.ring_nodes * 4, have_s, seg,
                          A.ring_nodes - 1, ring_stage, (uint32_t)(THREADS * sizeof(double)));
            if (lane == src) have += seg;
        }
    };

    for (;;) {
        /* ---- macro-step boundary: open the next RK attempt (RungeKutta._step_impl) ---- */
        if (mode == M_WAIT) { mode = M_RK; new_step = true; }
        if constexpr (CTA_LOCKSTEP != 0 && KSMEM == 2) {
            /* ---- drain phase: pack the surviving storms into as few warps as possible ----
             * Once the queue is empty every warp keeps issuing the full RHS instruction stream for
             * however few lanes it has left, and twelve thin warps per SM run each other's tail at a
             * third of the speed one warp would have alone.  Whenever packing frees at least one warp,
             * the live storms (17 words each: y, K0, t, h, g, ...) move through shared memory to the
             * lowest threads of the CTA; emptied warps then only meet the slot barriers. */
            const unsigned idle_b = __ballot_sync(TCR_FULL, mode == M_IDLE);
            /* nothing to pack before the queue runs dry: the CTA-wide vote (a barrier per macro step) only starts then */
            const bool may_pack = A.pack && (((CTA_LOCKSTEP & 2) == 0) || cta_drained);
            if (may_pack && __syncthreads_and(drained || idle_b == 0u)) {
                __shared__ int w_act[THREADS / 32];
                const int wid = threadIdx.x >> 5;
                const unsigned act_b = ~idle_b;
                if (lane == 0) w_act[wid] = __popc(act_b);
                __syncthreads();
                int total = 0, warps_used = 0, before = 0;
#pragma unroll
                for (int w = 0; w < THREADS / 32; ++w) {
                    const int c = w_act[w];
                    total += c; warps_used += (c > 0); if (w < wid) before += c;
                }
                if (total > 0 && total <= 32 * (warps_used - 1)) {
                    /* the live storm of rank r moves to thread r: its home rows (y, t, h_abs, g, min_step) and K0 are
                     * read first, then written into thread r's rows; what it keeps in registers goes through the mailbox */
                    double* const stg = k_smem + 32 * THREADS;
                    const bool live = mode != M_IDLE;
                    double py[4] = {0, 0, 0, 0}, pk[4] = {0, 0, 0, 0}, pt = 0.0, ph = 0.0, pg = 0.0, pm = 0.0;
                    if (live) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) { py[i] = Y(i); pk[i] = Kg(0, i); }
                        pt = t; ph = h_abs; pg = g; pm = min_step;
                    }
                    __syncthreads();
                    if (live) {
                        const int r = before + __popc(act_b & ((1u << lane) - 1u));
                        double* d = stg + r;
#pragma unroll
                        for (int i = 0; i < 4; ++i) { d[i * THREADS] = py[i]; k_smem[k_off(0, i) + r] = pk[i]; }
                        d[4 * THREADS] = pt; d[5 * THREADS] = ph; d[6 * THREADS] = pg; d[7 * THREADS] = pm;
                        d[12 * THREADS] = hbl;
                        d[13 * THREADS] = __longlong_as_double((long long)sid);
                        d[14 * THREADS] = __hiloint2double(ym, nfev);
                        d[15 * THREADS] = __hiloint2double(n_out, n_attempts);
                        d[16 * THREADS] = __hiloint2double(mode, (rejected ? 1 : 0) | (new_step ? 2 : 0) | (any_v ? 4 : 0));
                        d[17 * THREADS] = __hiloint2double(have, (int)row);
                    }
                    __syncthreads();
                    if ((int)threadIdx.x < total) {
                        const double* d = stg + threadIdx.x;
                        hbl = d[12 * THREADS];
                        sid = (int64_t)__double_as_longlong(d[13 * THREADS]);
                        ym = __double2hiint(d[14 * THREADS]); nfev = __double2loint(d[14 * THREADS]);
                        n_out = __double2hiint(d[15 * THREADS]); n_attempts = __double2loint(d[15 * THREADS]);
                        mode = __double2hiint(d[16 * THREADS]);
                        const int fl = __double2loint(d[16 * THREADS]);
                        rejected = fl & 1; new_step = (fl & 2) != 0; any_v = (fl & 4) != 0;
                        status = 100;
                        row = (unsigned int)__double2loint(d[17 * THREADS]);
                        have = __double2hiint(d[17 * THREADS]);
                        ftab = A.ring ? A.ring + (size_t)row * A.ring_nodes * 4 : A.ftab + (size_t)sid * ns * 4;
                        trk = A.track + (A.track_row ? (size_t)row : (size_t)sid) * row_doubles;
                    } else {
                        mode = M_IDLE;             /* the queue is drained: this lane never needs a row again */
                    }
                    __syncthreads();                                     /* mailbox reusable by the next packing */
                }
            }
        }
        /* ---- idle lanes take the next storm from the queue.  Once per macro step: a storm ends where an RK attempt ends
         * (only a storm that fails the ventilation pre-check frees its lane mid-step: 0.1-0.5 % of them), and a new storm's
         * first ring segment has to be tabulated before its first evaluation ---- */
        if (!drained) {
            const unsigned need = __ballot_sync(TCR_FULL, mode == M_IDLE && lane < A.lane_cap);
            if (need) {
                const int cnt = __popc(need), leader = __ffs(need) - 1;
                unsigned long long base = 0;
                if (lane == leader) base = atomicAdd(A.queue, (unsigned long long)cnt);
                base = __shfl_sync(TCR_FULL, base, leader);
                if ((need >> lane) & 1u) {
                    const int64_t my = (int64_t)base + __popc(need & ((1u << lane) - 1u));
                    if (my < n) {
                        sid = my;
                        ym = A.ym[sid];
                        Y(0) = A.lon0[sid]; Y(1) = A.lat0[sid]; Y(2) = A.v0[sid]; Y(3) = A.m0[sid];
                        hbl = 0.5 * p.Ck / A.h_bl[sid];          /* storm-constant prefactor of dv/dt, dm/dt */
                        ftab = A.ring ? A.ring + (size_t)row * A.ring_nodes * 4 : A.ftab + (size_t)sid * ns * 4;
                        have = 0;
                        trk = A.track + (A.track_row ? (size_t)row : (size_t)sid) * row_doubles;
                        nfev = 0; n_out = 0; n_attempts = 0; any_v = false; t = 0.0; status = 100;
                        mode = M_INIT0;
                    }
                }
                if ((int64_t)base + cnt >= n) drained = true;
            }
        }
        if (mode == M_RK) {
            if (new_step) {
                min_step = 10.0 * (tcr_bits2d(tcr_d2bits(t) + 1) - t);
                if (h_abs > max_step) h_abs = max_step; else if (h_abs < min_step) h_abs = min_step;
                rejected = false;
                new_step = false;
            }
            if (h_abs < min_step || tcr_isnan(h_abs) || ++n_attempts > TCR_MAX_RK_ATTEMPTS) {
                finalize(TCR_STATUS_FAILED);
            } else {
                h = h_abs;
                t_new = t + h;
                if (t_new - t_bound > 0.0) t_new = t_bound;
                h = t_new - t;
                h_abs = fabs(h);
            }
        }
        /* ---- Fourier ring: every stage time of this attempt lies in (t, t_new], so the highest node it can bracket is the
         * upper node of t_new; a new storm's first evaluation (t = 0) brackets nodes 0 and 1 ---- */
        if (A.ring) {
            const int need = ring_need(mode == M_RK ? tcr_fs_index(cx, t_new) : (mode == M_INIT0 ? 1 : -1));
            if (CTA_SERVE && A.ring_cta) {
                const int par = rq_par;
                if (need >= have) {              /* one segment is always enough here (a segment exceeds an attempt's span) */
                    rq[atomicAdd(&rq_n[par], 1)] = make_int4((int)sid, (int)row, have, 0);
                    have += A.ring_nodes >> 1;
                }
                __syncthreads();
                const int nreq = rq_n[par];
                if (threadIdx.x == 0) { rq_n[par ^ 1] = 0; rq_par = par ^ 1; }      /* the next macro step's list */
                for (int i = (int)(threadIdx.x >> 5); i < nreq; i += THREADS / 32) {
                    const int4 r = rq[i];
                    tcr_ring_fill(cx, A.coef + (size_t)r.x * TCR_N_PHASES, A.ring + (size_t)(unsigned)r.y * A.ring_nodes * 4, r.z,
                                  A.ring_nodes >> 1, A.ring_nodes - 1, ring_stage, (uint32_t)(THREADS * sizeof(double)));
                }
                /* the slot-0 barrier below orders these fills before the evaluations that read them */
            }
            ring_serve(need);                    /* free-running variants; with the CTA service nothing is left to do */
        }

#pragma unroll 1
        for (int slot = 0; slot < 6; ++slot) {
            if (slot == 1 && A.ring) {
                /* select_initial_step's second evaluation sits at t = h0, which is known only now: inside segment 0 in
                 * practice (h0 is seconds to minutes), but nothing bounds it */
                ring_serve(ring_need(mode == M_INIT1 ? tcr_fs_index(cx, 0.0 + h0) : -1));
            }
            if constexpr (CTA_LOCKSTEP != 0) {
                /* all warps of the CTA enter every RHS evaluation together: the ~45 KB of RHS code
                 * (larger than the 32 KB L1.5 instruction cache) is then fetched once per CTA instead
                 * of once per warp; the CTA retires when its last warp runs dry */
                /* CTA_LOCKSTEP = bit mask of the slots that re-align the warps (bit 0 always set) */
                if (slot == 0) { if (__syncthreads_or(mode != M_IDLE) == 0) return; }
                else if (slot == 1 && (CTA_LOCKSTEP & 2)) cta_drained = __syncthreads_or(drained) != 0;
                else if ((CTA_LOCKSTEP >> slot) & 1) __syncthreads();
            } else {
                if (slot == 0 && __ballot_sync(TCR_FULL, mode != M_IDLE) == 0u) return;   /* warp-uniform */
            }

            /* ---- evaluation point of this slot ---- */
            double te = 0.0, ye[4] = {0, 0, 0, 0};
            bool ev = false;
            if (mode == M_RK) {
                ev = true;
                switch (slot) {
                case 0:
                    te = t + RK_C1 * h;
#pragma unroll
                    for (int i = 0; i < 4; ++i) ye[i] = fma(Kg(0, i) * RK_A10, h, Y(i));
                    break;
                case 1:
                    te = t + RK_C2 * h;
#pragma unroll
                    for (int i = 0; i < 4; ++i) ye[i] = fma(fma(Kg(1, i), RK_A21, Kg(0, i) * RK_A20), h, Y(i));
                    break;
                case 2:
                    te = t + RK_C3 * h;
#pragma unroll
                    for (int i = 0; i < 4; ++i) ye[i] = fma(fma(Kg(2, i), RK_A32, fma(Kg(1, i), RK_A31, Kg(0, i) * RK_A30)), h, Y(i));
                    break;
                case 3:
                    te = t + RK_C4 * h;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        ye[i] = fma(fma(Kg(3, i), RK_A43, fma(Kg(2, i), RK_A42, fma(Kg(1, i), RK_A41, Kg(0, i) * RK_A40))), h, Y(i));
                    break;
                case 4:
                    te = t + 1.0 * h;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        ye[i] = fma(fma(Kg(4, i), RK_A54, fma(Kg(3, i), RK_A53, fma(Kg(2, i), RK_A52, fma(Kg(1, i), RK_A51, Kg(0, i) * RK_A50)))), h, Y(i));
                    break;
                default:
                    te = t + h;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        ye[i] = fma(h, fma(Kg(5, i), RK_B5, fma(Kg(4, i), RK_B4, fma(Kg(3, i), RK_B3, fma(Kg(2, i), RK_B2, Kg(0, i) * RK_B0)))), Y(i));
                        Ks(7, i, ye[i]);
                    }
                    break;
                }
            } else if (mode == M_INIT0) {
                ev = true; te = 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i) ye[i] = Y(i);
            } else if (mode == M_INIT1) {
                ev = true; te = 0.0 + h0;
#pragma unroll
                for (int i = 0; i < 4; ++i) ye[i] = Kg(7, i);
            }

            double dy[4] = {0, 0, 0, 0};
            TcrRhsAux aux = {0, 0, 0, {0, 0, 0, 0}};
            if (ev) {
                if constexpr (FAST_RHS) {
                    /* straight-line evaluation (tcr_rhs_fast.cuh); an evaluation that left the common case of any
                     * of its operations is repeated by the specification form */
                    const TcrFtab ft = {ftab, fmask};
                    if (tcr_rhs_fast(cx, ym, ft, hbl, te, ye, dy, aux)) {
                        double yy[4] = {ye[0], ye[1], ye[2], ye[3]}, dd[4];
                        TcrRhsAux ax2;
                        tcr_rhs_slow(cx, ym, ft, hbl, te, yy, dd, &ax2);
                        dy[0] = dd[0]; dy[1] = dd[1]; dy[2] = dd[2]; dy[3] = dd[3];
                        aux = ax2;
                    }
                } else {
                    tcr_rhs(cx, ym, TcrFtab{ftab, fmask}, hbl, te, ye, dy, aux);
                }
                ++nfev;
            }

            /* ---- consume ---- */
            if (mode == M_RK) {
                switch (slot) {
                case 0: Ks(1, 0, dy[0]); Ks(1, 1, dy[1]); Ks(1, 2, dy[2]); Ks(1, 3, dy[3]); break;
                case 1: Ks(2, 0, dy[0]); Ks(2, 1, dy[1]); Ks(2, 2, dy[2]); Ks(2, 3, dy[3]); break;
                case 2: Ks(3, 0, dy[0]); Ks(3, 1, dy[1]); Ks(3, 2, dy[2]); Ks(3, 3, dy[3]); break;
                case 3: Ks(4, 0, dy[0]); Ks(4, 1, dy[1]); Ks(4, 2, dy[2]); Ks(4, 3, dy[3]); break;
                case 4: Ks(5, 0, dy[0]); Ks(5, 1, dy[1]); Ks(5, 2, dy[2]); Ks(5, 3, dy[3]); break;
                default: Ks(6, 0, dy[0]); Ks(6, 1, dy[1]); Ks(6, 2, dy[2]); Ks(6, 3, dy[3]); break;
                }
            } else if (mode == M_INIT0) {
                /* ventilation pre-check (coupled_fast.py:238-244); the evaluation at (0, y0) is
                 * also f0 of the integration, so it is only counted when the storm is integrated */
                bool vent = false;
                if (aux.vpot > 0.0) {
                    double vent_index = aux.S_free * aux.chi / aux.vpot;
                    if (vent_index >= 1.0) vent = true;
                }
                if (vent) {
                    finalize(TCR_STATUS_VENT);
                } else {
                    /* select_initial_step, first half (scipy/integrate/_ivp/common.py) */
                    double sa[4], sb[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        Ks(0, i, dy[i]);
                        double scale = atol + fabs(Y(i)) * rtol;
                        sa[i] = Y(i) / scale; sb[i] = dy[i] / scale;
                    }
                    double d0 = tcr_rms4(sa);
                    d1 = tcr_rms4(sb);
                    if (d0 < 1e-5 || d1 < 1e-5) h0 = 1e-6; else h0 = 0.01 * d0 / d1;
                    if (t_bound < h0) h0 = t_bound;
#pragma unroll
                    for (int i = 0; i < 4; ++i) Ks(7, i, fma(h0, dy[i], Y(i)));
                    mode = M_INIT1;
                }
            } else if (mode == M_INIT1) {
                double sc[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    double scale = atol + fabs(Y(i)) * rtol;
                    sc[i] = (dy[i] - Kg(0, i)) / scale;
                }
                double d2 = tcr_rms4(sc) / h0, h1;
                if (d1 <= 1e-15 && d2 <= 1e-15) { h1 = h0 * 1e-3; if (h1 < 1e-6) h1 = 1e-6; }
                else h1 = tcr_pow(0.01 / (d2 > d1 ? d2 : d1), 0.2);
                double hh = 100.0 * h0;
                if (h1 < hh) hh = h1;
                if (t_bound < hh) hh = t_bound;
                if (max_step < hh) hh = max_step;
                h_abs = hh;
                { const double yl[4] = {Y(0), Y(1), Y(2), Y(3)}; g = tcr_event(p, yl); }
                /* an initial-step probe beyond the ring's length has overwritten the nodes the integration starts from */
                if (A.ring && have > A.ring_nodes) have = 0;
                mode = M_WAIT;
            }
        }

        /* ---- end of the RK attempt: error control, events, dense output ---- */
        if (mode == M_RK) {
            double en[4];
            const double ynl[4] = {Kg(7, 0), Kg(7, 1), Kg(7, 2), Kg(7, 3)};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double ay = fabs(Y(i)), an = fabs(ynl[i]);
                double mx = (tcr_isnan(ay) || tcr_isnan(an)) ? NAN : (ay > an ? ay : an);
                double scale = atol + mx * rtol;
                double e = fma(Kg(6, i), RK_E6, fma(Kg(5, i), RK_E5, fma(Kg(4, i), RK_E4, fma(Kg(3, i), RK_E3, fma(Kg(2, i), RK_E2, Kg(0, i) * RK_E0)))));
                en[i] = (e * h) / scale;
            }
            const double err = tcr_rms4(en);
            if (err < 1.0) {
                double factor;
                if (err == 0.0) factor = 10.0;
                else { factor = 0.9 * tcr_pow(err, -0.2); if (10.0 < factor) factor = 10.0; }
                if (rejected && 1.0 < factor) factor = 1.0;
                h_abs *= factor;
                /* accepted */
                const double t_old = t;
                if (t_new - t_bound >= 0.0) status = TCR_STATUS_FINISHED;
                const double g_new = tcr_event(p, ynl);
                const bool active = ((g <= 0.0) && (g_new >= 0.0)) || ((g >= 0.0) && (g_new <= 0.0));
                double t_emit = t_new;
                if (active) { status = TCR_STATUS_EVENT; if (g == 0.0) t_emit = t_old; }
                g = g_new;
                /* t_eval sampling through the dense output */
                const int i_new = tcr_nodes_le(cx, t_emit, n_out);
                if (i_new > n_out) {
                    double Q1[4], Q2[4], Q3[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        Q1[i] = fma(Kg(6, i), RK_P61, fma(Kg(5, i), RK_P51, fma(Kg(4, i), RK_P41, fma(Kg(3, i), RK_P31, fma(Kg(2, i), RK_P21, Kg(0, i) * RK_P01)))));
                        Q2[i] = fma(Kg(6, i), RK_P62, fma(Kg(5, i), RK_P52, fma(Kg(4, i), RK_P42, fma(Kg(3, i), RK_P32, fma(Kg(2, i), RK_P22, Kg(0, i) * RK_P02)))));
                        Q3[i] = fma(Kg(6, i), RK_P63, fma(Kg(5, i), RK_P53, fma(Kg(4, i), RK_P43, fma(Kg(3, i), RK_P33, fma(Kg(2, i), RK_P23, Kg(0, i) * RK_P03)))));
                    }
                    const double hd = t_new - t_old, yhd = tcr_rcp_seed(hd);
                    const double k0l[4] = {Kg(0, 0), Kg(0, 1), Kg(0, 2), Kg(0, 3)};
                    const double yl[4] = {Y(0), Y(1), Y(2), Y(3)};
                    /* four samples per trip, computed unconditionally (the last trip repeats its final
                     * sample) so that the four dependent chains interleave; only the stores are guarded */
                    for (int k0 = n_out; k0 < i_new; k0 += 4) {
                        double o[4][4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int k = min(k0 + u, i_new - 1);
                            double x = tcr_div_y(tcr_node_time(cx, k) - t_old, hd, yhd);
                            double p2 = x * x, p3 = p2 * x, p4 = p3 * x;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                double acc = (k0l[i] * 1.0) * x;
                                acc = fma(Q1[i], p2, acc);
                                acc = fma(Q2[i], p3, acc);
                                acc = fma(Q3[i], p4, acc);
                                o[u][i] = fma(hd, acc, yl[i]);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (k0 + u < i_new) {
                                double2* dst = reinterpret_cast<double2*>(trk + (size_t)(k0 + u) * 4);
                                dst[0] = make_double2(o[u][0], o[u][1]);
                                dst[1] = make_double2(o[u][2], o[u][3]);
                                if (o[u][2] >= p.seed_v_thresh) any_v = true;
                            }
                        }
                    }
                    n_out = i_new;
                }
                t = t_new;
#pragma unroll
                for (int i = 0; i < 4; ++i) { Y(i) = ynl[i]; Ks(0, i, Kg(6, i)); }
                new_step = true;
                if (status != 100) finalize(status);
            } else {
                double fac = 0.9 * tcr_pow(err, -0.2);
                if (!(fac > 0.2)) fac = 0.2;
                h_abs *= fac;
                rejected = true;
            }
        }
    }
}

/* ======================================================================================== */
/* single evaluations of Coupled_FAST.dydt (coupled_fast.py:196-207) and BetaAdvectionTrack._env_winds              */
/* (bam_track.py:116-128) at given (t, y): the inner tier of the reference's seam and the test hook of rows a5-a11.  */
/* One thread per evaluation, the same tcr_rhs the integrator runs.                                                   */
/* ======================================================================================== */
__global__ void __launch_bounds__(128) k_rhs_eval(const __grid_constant__ TcrCtx cx, int64_t n, const int32_t* __restrict__ ym,
                                                  const double* __restrict__ t, const double* __restrict__ y,
                                                  const double* __restrict__ h_bl, const double* __restrict__ ftab,
                                                  double* __restrict__ dydt, double* __restrict__ env)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double yi[4] = {y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]};
    double dy[4];
    TcrRhsAux aux = {0, 0, 0, {0, 0, 0, 0}};
    tcr_rhs(cx, ym[i], TcrFtab{ftab + (size_t)i * cx.p.n_steps * 4, TCR_FTAB_FULL}, 0.5 * cx.p.Ck / h_bl[i], t[i], yi, dy, aux);
#pragma unroll
    for (int k = 0; k < 4; ++k) { dydt[4 * i + k] = dy[k]; env[4 * i + k] = aux.wf[k]; }
}

/* ======================================================================================== */
/* post-processing of a candidate: env-wind recompute (util/compute.py:201-202), translation  */
/* speed (util/sphere.py:58-83), axi_to_max_wind (wind/tc_wind.py:6-21), nanmax(vmax) >= 18   */
/* (util/compute.py:205).  One CTA per storm; lanes stride over the output samples.           */
/* ======================================================================================== */
/* one output sample of a storm: env winds at the track point (util/compute.py:201-202), translation speed from the
 * neighbouring samples (util/sphere.py:58-83) and axi_to_max_wind (wind/tc_wind.py:6-21)                       */
/* ftab: the storm's full Fourier table, or NULL and coef: its 60 coefficient pairs (the nodes are then formed on the spot) */
__device__ __forceinline__ double tcr_post_sample(const TcrCtx& cx, int ym, const double* __restrict__ ftab, const double2* __restrict__ coef,
                                                  const double* __restrict__ trk, int nt, int k, double w[4])
{
    const tcr_params& p = cx.p;
    const double2 a = *reinterpret_cast<const double2*>(trk + (size_t)k * 4);
    const double lon = a.x, lat = a.y;
    const double v = trk[(size_t)k * 4 + 2];
    const double tk = tcr_node_time(cx, k);
    w[0] = w[1] = w[2] = w[3] = 0.0;
    if (!(tcr_isnan(lon) || tcr_isnan(tk))) {
        TcrFsNodes fsn;
        if (ftab) tcr_fs_begin(cx, TcrFtab{ftab, TCR_FTAB_FULL}, tk, fsn);
        else tcr_fs_begin_coef(cx, coef, tk, fsn);
        TcrCell c;
        tcr_cell_at(cx.tab.lon, cx.tab.lat, lon, lat, c);
        tcr_env_winds_cell(cx, tcr_record(cx.tab, ym, c), c, fsn, tk, w);
    }
    double ut, vt;
    if (nt <= 1) {
        ut = vt = NAN;
    } else {
        double lon_m, lat_m, lon_p, lat_p;
        if (k == 0) {
            const double2 b = *reinterpret_cast<const double2*>(trk + 4);
            lon_m = 2 * lon - b.x; lat_m = 2 * lat - b.y;
        } else {
            const double2 b = *reinterpret_cast<const double2*>(trk + (size_t)(k - 1) * 4);
            lon_m = b.x; lat_m = b.y;
        }
        if (k == nt - 1) {
            const double2 b = *reinterpret_cast<const double2*>(trk + (size_t)(nt - 2) * 4);
            lon_p = 2 * lon - b.x; lat_p = 2 * lat - b.y;
        } else {
            const double2 b = *reinterpret_cast<const double2*>(trk + (size_t)(k + 1) * 4);
            lon_p = b.x; lat_p = b.y;
        }
        double dlon = 0.5 * (tcr_sign(lon_p - lon_m) * tcr_haversine_km(p, lon_p, lat, lon_m, lat));
        double dlat = 0.5 * (tcr_sign(lat_p - lat_m) * tcr_haversine_km(p, lon, lat_p, lon, lat_m));
        ut = dlon * 1000.0 / p.dt_track;
        vt = dlat * 1000.0 / p.dt_track;
    }
    double G = 0.8 + 0.35 * (1.0 + tcr_tanh((lat - 35.0) / 10.0));
    if (1.0 < G) G = 1.0;
    double u_shr = w[0] - w[2], v_shr = w[1] - w[3];
    double U = G * ut + 0.1 * u_shr * v / 15.0;
    double V = G * vt + 0.1 * v_shr * v / 15.0;
    double mag_inc = sqrt(U * U + V * V);
    double vm;
    if (mag_inc == 0.0) {
        vm = fabs(v);
    } else {
        double mag_fac = (v * 0.50) / mag_inc;
        if (!tcr_isnan(mag_fac) && 1.0 < mag_fac) mag_fac = 1.0;
        double ug = v * (U / mag_inc) + U * mag_fac;
        double vg = v * (V / mag_inc) + V * mag_fac;
        vm = sqrt(ug * ug + vg * vg);
    }
    return vm;
}

struct PostArgs {
    int64_t n;                          /* storms if list == NULL */
    const int32_t* list; const unsigned int* list_count;
    const int32_t* ym; const double* ftab; const double* track;
    const double2* coef;                /* [n][60] coefficient pairs, used when ftab == NULL (the integrator kept only rings) */
    const int32_t* track_row;           /* pool row of a candidate's track (NULL: track is indexed by storm) */
    const unsigned int* pool_ctl;       /* [1] != 0: the track pool overflowed, the wave is void             */
    const int32_t* n_time; const int32_t* status;
    double* env;                        /* [n][n_steps][4]; NULL: only the kept flag is formed (tcr_run_years) */
    double* vmax;                       /* [n][n_steps]                                                        */
    uint32_t* flags;
};

__global__ void __launch_bounds__(128) k_postprocess(const __grid_constant__ TcrCtx cx, const PostArgs A)
{
    __shared__ unsigned long long best_bits;
    __shared__ int have;
    const tcr_params& p = cx.p;
    const int ns = p.n_steps;
    if (A.pool_ctl && A.pool_ctl[1]) return;
    const int64_t count = A.list ? (int64_t)*A.list_count : A.n;
    for (int64_t item = blockIdx.x; item < count; item += gridDim.x) {
        const int64_t sid = A.list ? (int64_t)A.list[item] : item;
        const int nt = A.n_time[sid];
        if (nt <= 0 || A.status[sid] == TCR_STATUS_VENT) continue;
        __syncthreads();
        if (threadIdx.x == 0) { best_bits = 0ull; have = 0; }
        __syncthreads();
        const int ym = A.ym[sid];
        const double* trk = A.track + (size_t)(A.track_row ? (int64_t)A.track_row[sid] : sid) * ns * 4;
        const double* ftab = A.ftab ? A.ftab + (size_t)sid * ns * 4 : nullptr;
        const double2* coef = A.coef ? A.coef + (size_t)sid * TCR_N_PHASES : nullptr;
        for (int k = threadIdx.x; k < nt; k += blockDim.x) {
            double w[4];
            const double vm = tcr_post_sample(cx, ym, ftab, coef, trk, nt, k, w);
            if (A.env) {
                double2* ed = reinterpret_cast<double2*>(A.env + ((size_t)sid * ns + k) * 4);
                ed[0] = make_double2(w[0], w[1]);
                ed[1] = make_double2(w[2], w[3]);
                A.vmax[(size_t)sid * ns + k] = vm;
            }
            if (!tcr_isnan(vm)) {
                /* vm >= 0: the bit pattern orders like the value */
                atomicMax(&best_bits, (unsigned long long)tcr_d2bits(fabs(vm)));
                have = 1;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t fl = A.flags[sid];
            if ((fl & TCR_FLAG_IS_TC) && have && tcr_bits2d((int64_t)best_bits) >= p.seed_vmax_thresh) fl |= TCR_FLAG_KEPT;
            A.flags[sid] = fl;
        }
    }
}

/* ======================================================================================== */
/* seeding: one thread per attempt = one pass of `while not seed_passed` (compute.py:136-175) */
/* Draw slots (stream 0): block 0 = (lon, lat), 1 = (month, low-latitude test), 2 = Box-      */
/* Muller pair for v_init, 3.. = ocean-point redraws.                                         */
/* ======================================================================================== */
struct SeedArgs {
    int n_years;
    const int64_t* wave_off;            /* [n_years+1] prefix offsets of the years' attempt ranges */
    const int64_t* k0;                  /* [n_years] first attempt index of each year's range      */
    const int32_t* ym_base; const int32_t* year_key;
    uint32_t run_seed;
    /* per attempt (flat) */
    int32_t* code; int32_t* basin; int32_t* month;
    double* lon; double* lat; double* v0; double* m0; double* pi_gen;      /* pi_gen may be NULL */
    unsigned int* blk_count;            /* [gridDim.x] attempts of the block that go on to gen_track (may be NULL) */
    int rank, world;                    /* within-year sharding: this rank integrates the attempts k with k % world == rank */
};

__device__ __forceinline__ double tcr_mask_at(const uint2 r[4], int b, const TcrCell& c)
{
    /* (double)byte without the conversion pipe: 2^52 + byte is exact */
    auto byte = [&](const uint2& q) {
        return __hiloint2double(0x43300000, (int)(((b < 4 ? q.x : q.y) >> (8 * (b & 3))) & 0xffu)) - 4503599627370496.0;
    };
    return tcr_bilin_fitpack(byte(r[0]), byte(r[1]), byte(r[2]), byte(r[3]), c);
}

__device__ __forceinline__ void tcr_mask_cell(const TcrMasks& mk, double lon, double lat, TcrCell& c, uint2 r[4])
{
    tcr_cell_at(mk.lon, mk.lat, lon, lat, c);
    const uint4* q = reinterpret_cast<const uint4*>(mk.rec + ((size_t)c.iy * mk.ncx + c.ix) * 4);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    r[0] = make_uint2(a.x, a.y); r[1] = make_uint2(a.z, a.w);
    r[2] = make_uint2(b.x, b.y); r[3] = make_uint2(b.z, b.w);
}

__global__ void __launch_bounds__(256) k_seed(const __grid_constant__ TcrCtx cx, const SeedArgs A)
{
    const tcr_params& p = cx.p;
    const int64_t total = A.wave_off[A.n_years];
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int code = -1;
    int bi = 0, mon = 1, yr = 0;
    int64_t k = 0;
    int32_t key = 0;
    double gen_lon = 0, gen_lat = 0, v0 = 0, m0 = 0, pi = 0;
    const bool active = idx < total;
    const double* b = p.basin_bounds;
    TcrCell c;
    uint2 r[4];
    bool pending = false, exhausted = false;
    if (active) {
        while (yr + 1 < A.n_years && idx >= A.wave_off[yr + 1]) ++yr;
        k = A.k0[yr] + (idx - A.wave_off[yr]);
        key = A.year_key[yr];
        double u[2];
        tcr_draw2(A.run_seed, key, k, 0, 0, u);
        const double y_min = cx.gen_y_min, y_max = cx.gen_y_max;
        gen_lon = b[0] + (b[2] - b[0]) * u[0];
        gen_lat = tcr_asin(y_min + (y_max - y_min) * u[1]) * 180.0 / TCR_PI;
        tcr_mask_cell(cx.mk, gen_lon, gen_lat, c, r);
        pending = tcr_mask_at(r, 7, c) < 1e-2;
    }
    /* The ocean-point redraw (compute.py:144-148: `while f_b.ev(lon, lat) < 1e-2: redraw uniformly in the box`).  Redraw r
     * of attempt k is a pure function of (k, r), so it need not be evaluated by the attempt's own lane: the lanes of the
     * warp whose first draw hit the ocean would otherwise idle through the tail of the slowest lane's loop (on the global
     * box 12 of 32 lanes were active on average).  Every round the 32 lanes are dealt out to the still-pending attempts
     * and test their next g = 32 / pending redraws at once; the first hit in redraw order wins, and the winner is
     * re-evaluated by its owner afterwards (one full-width pass).  Redraws are capped at max_redraws (code 3). */
    {
        __shared__ unsigned char s_owner[8][32];              /* per warp: lane of the j-th pending attempt */
        unsigned char* own = s_owner[threadIdx.x >> 5];
        const int lane = threadIdx.x & 31;
        int next_r = 0;
        bool moved = false;                                    /* the genesis point was replaced by a redraw */
        unsigned pend = __ballot_sync(TCR_FULL, pending);
        while (pend) {
            const int m = __popc(pend);
            const int g = 32 / m;
            const int pos = __popc(pend & ((1u << lane) - 1u));
            if (pending) own[pos] = (unsigned char)lane;
            __syncwarp();
            const int j = lane / g;
            const bool has = j < m;
            const int owner = has ? (int)own[j] : 0;
            const long long k_o = __shfl_sync(TCR_FULL, (long long)k, owner);
            const int32_t key_o = __shfl_sync(TCR_FULL, key, owner);
            const int r_o = __shfl_sync(TCR_FULL, next_r, owner) + (lane - j * g);
            bool okr = false;
            double lo = 0.0, la = 0.0;
            if (has && r_o < p.max_redraws) {
                double uu[2];
                tcr_draw2(A.run_seed, key_o, (int64_t)k_o, 3u + (uint32_t)r_o, 0, uu);
                lo = b[0] + (b[2] - b[0]) * uu[0]; la = b[1] + (b[3] - b[1]) * uu[1];
                TcrCell cc;
                uint2 rr[4];
                tcr_mask_cell(cx.mk, lo, la, cc, rr);
                okr = !(tcr_mask_at(rr, 7, cc) < 1e-2);
            }
            const unsigned hit = __ballot_sync(TCR_FULL, okr);
            /* the owner takes the point of its first hit -- or, when the redraws are exhausted, of the last one tested --
             * from the helper lane that evaluated it */
            const unsigned grp = pending ? (hit >> (pos * g)) & (g == 32 ? 0xffffffffu : ((1u << g) - 1u)) : 0u;
            const bool last_round = pending && !grp && next_r + g >= p.max_redraws;
            int src = lane;
            if (grp) src = pos * g + __ffs((int)grp) - 1;
            else if (last_round && p.max_redraws > next_r) src = pos * g + (p.max_redraws - 1 - next_r);
            const double lo_w = __shfl_sync(TCR_FULL, lo, src), la_w = __shfl_sync(TCR_FULL, la, src);
            if (pending) {
                if (grp) { gen_lon = lo_w; gen_lat = la_w; moved = true; pending = false; }
                else {
                    next_r += g;
                    if (next_r >= p.max_redraws) {
                        exhausted = true; pending = false;
                        if (src != lane) { gen_lon = lo_w; gen_lat = la_w; moved = true; }
                    }
                }
            }
            __syncwarp();
            pend = __ballot_sync(TCR_FULL, pending);
        }
        if (moved) tcr_mask_cell(cx.mk, gen_lon, gen_lat, c, r);
    }
    if (active) {
        double u[2];
        tcr_draw2(A.run_seed, key, k, 1, 0, u);
        mon = 1 + (int)floor(u[0] * 12.0);
        const double r_lowlat = u[1];
        /* arg max over the seven basin masks (first maximum wins, compute.py:150-153).  A basin whose four corner bytes
         * are zero interpolates to +0 exactly and can only win as basin 0 of an all-zero cell, so only the basins present
         * at the cell are evaluated: one or two instead of seven */
        const unsigned ox = r[0].x | r[1].x | r[2].x | r[3].x, oy = r[0].y | r[1].y | r[2].y | r[3].y;
        double best = (ox & 0xffu) ? tcr_mask_at(r, 0, c) : 0.0;
        bi = 0;
        unsigned present = 0u;
#pragma unroll
        for (int i = 1; i < TCR_N_BASINS; ++i)
            if (((i < 4 ? ox : oy) >> (8 * (i & 3))) & 0xffu) present |= 1u << i;
        while (present) {
            const int i = __ffs((int)present) - 1;
            present &= present - 1u;
            const double val = tcr_mask_at(r, i, c);
            if (val > best) { best = val; bi = i; }
        }
        double q = (fabs(gen_lat) - p.lat_vort_fac) / 12.0;
        if (q < 0.0) q = 0.0;
        if (q > 1.0) q = 1.0;
        const double prob = tcr_pow(q, p.lat_vort_power[bi]);
        const bool counted = !exhausted && best > 1e-3 && r_lowlat < prob;
        /* the whole-year path (pi_gen == NULL) only needs what the sequential loop would have looked at: the potential
         * intensity of a COUNTED attempt (compute.py:162-169) and the initial state of a PASSED one (:172-175); the
         * test hook (tcr_seed_attempts) evaluates everything for every attempt */
        const bool full = A.pi_gen != nullptr;
        TcrCell ce;
        const float4* rec = nullptr;
        if (full || counted) {
            tcr_cell_at(cx.tab.lon, cx.tab.lat, gen_lon, gen_lat, ce);
            rec = tcr_record(cx.tab, A.ym_base[yr] + mon - 1, ce);
            pi = tcr_bilin(__ldg(rec + CH_VPOT), ce);
        }
        if (exhausted) code = 3;
        else if (counted) code = pi > p.pi_gen_min ? 2 : 1;
        else code = 0;
        if (full || code == 2) {
            tcr_draw2(A.run_seed, key, k, 2, 0, u);
            double bs, bc;
            tcr_sincos2pi(u[1], &bs, &bc);
            const double randn = sqrt(-2.0 * tcr_log(1.0 - u[0])) * bc;
            v0 = p.seed_v_init + randn;
            const double rh = tcr_bilin(__ldg(rec + CH_RH), ce);
            const double mi = p.minit_amp / (1.0 + tcr_exp(-(rh - p.minit_center) * p.minit_slope)) + p.minit_offset;
            m0 = mi > 0.0 ? mi : 0.0;
        }
        A.code[idx] = code; A.basin[idx] = bi; A.month[idx] = mon;
        if (full || code == 2) { A.lon[idx] = gen_lon; A.lat[idx] = gen_lat; A.v0[idx] = v0; A.m0[idx] = m0; }
        if (A.pi_gen) A.pi_gen[idx] = pi;
    }
    if (A.blk_count) {
        __shared__ unsigned int s_cnt;
        if (threadIdx.x == 0) s_cnt = 0u;
        __syncthreads();
        const bool own = A.world <= 1 || (int)(k % A.world) == A.rank;
        const unsigned pass = __ballot_sync(TCR_FULL, code == 2 && own);
        if ((threadIdx.x & 31) == 0 && pass) atomicAdd(&s_cnt, (unsigned int)__popc(pass));
        __syncthreads();
        if (threadIdx.x == 0) A.blk_count[blockIdx.x] = s_cnt;
    }
}

/* exclusive scan of the seed blocks' pass counts (one CTA) -> blk_off; n_slots = min(total,
 * slot_cap); consumed[y] = W_y (k_assign_slots lowers it where the slot capacity runs out)   */
__global__ void __launch_bounds__(1024) k_scan_counts(const unsigned int* __restrict__ blk_count, unsigned int* __restrict__ blk_off,
                                                      int n_blocks, unsigned int slot_cap, unsigned int* __restrict__ n_slots,
                                                      unsigned int* __restrict__ n_pass, const int64_t* __restrict__ wave_off,
                                                      int n_years, int64_t* __restrict__ consumed)
{
    __shared__ unsigned int warp_tot[32];
    __shared__ unsigned int s_run;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_run = 0u;
    for (int y = tid; y < n_years; y += blockDim.x) consumed[y] = wave_off[y + 1] - wave_off[y];
    __syncthreads();
    for (int base = 0; base < n_blocks; base += 1024 * 4) {
        unsigned int v[4], sum = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int i = base + tid * 4 + j;
            v[j] = i < n_blocks ? blk_count[i] : 0u;
            sum += v[j];
        }
        unsigned int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned int u = __shfl_up_sync(TCR_FULL, incl, d); if (lane >= d) incl += u; }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            unsigned int t = warp_tot[lane], sc = t;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { unsigned int u = __shfl_up_sync(TCR_FULL, sc, d); if (lane >= d) sc += u; }
            warp_tot[lane] = sc - t;
        }
        __syncthreads();
        unsigned int off = s_run + warp_tot[wid] + incl - sum;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int i = base + tid * 4 + j;
            if (i < n_blocks) blk_off[i] = off;
            off += v[j];
        }
        __syncthreads();
        if (tid == 1023) s_run = off;
        __syncthreads();
    }
    if (tid == 0) { *n_pass = s_run; *n_slots = s_run < slot_cap ? s_run : slot_cap; }
}

/* slots in ATTEMPT ORDER (year-major): slot = rank of the attempt among the wave's integrated
 * attempts.  Attempts whose slot would exceed the capacity are not integrated in this wave
 * (att_slot = -2) and the year's consumed range is cut at the first of them.  Same launch
 * geometry as k_seed (blk_off is per k_seed block).                                          */
struct AssignArgs {
    int n_years;
    const int64_t* wave_off; const int64_t* k0;
    const int32_t* ym_base; const int32_t* year_key;
    const int32_t* code; const int32_t* basin; const int32_t* month;
    const double* lon; const double* lat; const double* v0; const double* m0;
    const unsigned int* blk_off; unsigned int slot_cap;
    int32_t* att_slot; int64_t* consumed;
    int32_t* s_ym; double* s_lon; double* s_lat; double* s_v0; double* s_m0; double* s_hbl;
    int64_t* s_att; int32_t* s_key;
    int rank, world;
};

__global__ void __launch_bounds__(256) k_assign_slots(const __grid_constant__ TcrCtx cx, const AssignArgs A)
{
    __shared__ unsigned int warp_cnt[8];
    const int64_t total = A.wave_off[A.n_years];
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int code = idx < total ? A.code[idx] : -1;
    int yr = 0;
    if (idx < total) while (yr + 1 < A.n_years && idx >= A.wave_off[yr + 1]) ++yr;
    /* within-year sharding: a passing attempt another rank owns is not integrated here */
    if (code == 2 && A.world > 1 && (int)((A.k0[yr] + (idx - A.wave_off[yr])) % A.world) != A.rank) code = 0;
    const unsigned pass = __ballot_sync(TCR_FULL, code == 2);
    if (lane == 0) warp_cnt[wid] = (unsigned int)__popc(pass);
    __syncthreads();
    if (idx >= total) return;
    if (code != 2) { A.att_slot[idx] = -1; return; }
    unsigned int rank = __popc(pass & ((1u << lane) - 1u));
    for (int w = 0; w < wid; ++w) rank += warp_cnt[w];
    const unsigned int s = A.blk_off[blockIdx.x] + rank;
    if (s >= A.slot_cap) {
        A.att_slot[idx] = -2;
        atomicMin(reinterpret_cast<unsigned long long*>(A.consumed + yr), (unsigned long long)(idx - A.wave_off[yr]));
        return;
    }
    const int bi = A.basin[idx], mon = A.month[idx];
    A.att_slot[idx] = (int32_t)s;
    A.s_ym[s] = A.ym_base[yr] + mon - 1;
    A.s_lon[s] = A.lon[idx]; A.s_lat[s] = A.lat[idx]; A.s_v0[s] = A.v0[idx]; A.s_m0[s] = A.m0[idx];
    A.s_hbl[s] = cx.p.atm_bl_depth[bi];
    A.s_att[s] = A.k0[yr] + (idx - A.wave_off[yr]);
    A.s_key[s] = A.year_key[yr];
}

/* ======================================================================================== */
/* ordered selection: the sequential `while nt < n_tracks` semantics of compute.py:134-209    */
/* over the indexed attempt stream.  One CTA per year.                  */
/* ======================================================================================== */
/* totals of one wave, fully parallel (thread per attempt): per year the counters of ALL consumed
 * attempts (k_select subtracts the over-shoot beyond i*), and one byte per attempt saying whether
 * it produced a kept storm (so that the ordered scan of k_select reads a dense byte stream).   */
struct WaveStatsArgs {
    int n_years;
    const int64_t* wave_off; const int64_t* consumed;
    const int32_t* code; const int32_t* basin; const int32_t* month; const int32_t* att_slot;
    const int32_t* n_time; const int32_t* nfev; const uint32_t* flags;
    uint8_t* att_kept;                   /* [total attempts]                                     */
    /* wave_glob [n_years][86] u32: counted attempts per (basin, month), their sum, exhausted redraw chains -- what within-year sharding
     * all-reduces (each rank counts the attempts it owns); wave_loc [n_years][3] u64: integrated storms, their
     * samples and RHS evaluations -- this rank's share, never reduced                                       */
    unsigned int* wave_glob; unsigned long long* wave_loc;
    const int64_t* k0; int rank, world;
};

__global__ void __launch_bounds__(256) k_wave_stats(const WaveStatsArgs A)
{
    __shared__ unsigned int s_hist[TCR_N_BASINS * 12 + 2];
    __shared__ unsigned long long s_tot[3];
    const int64_t total = A.wave_off[A.n_years];
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int tid = threadIdx.x;
    const int NG = TCR_N_BASINS * 12 + 2;                   /* 84 (basin, month) bins, counted attempts, exhausted redraw chains */
    /* year of the block's first attempt: the shared accumulators belong to it */
    const int64_t first = (int64_t)blockIdx.x * blockDim.x;
    int yr0 = 0;
    while (yr0 + 1 < A.n_years && first >= A.wave_off[yr0 + 1]) ++yr0;
    for (int i = tid; i < NG; i += blockDim.x) s_hist[i] = 0u;
    if (tid < 3) s_tot[tid] = 0ull;
    __syncthreads();
    if (idx < total) {
        int yr = yr0;
        while (yr + 1 < A.n_years && idx >= A.wave_off[yr + 1]) ++yr;
        const int64_t li = idx - A.wave_off[yr];
        uint8_t kept = 0;
        if (li < A.consumed[yr]) {
            const int code = A.code[idx], slot = A.att_slot[idx];
            const bool own = A.world <= 1 || (int)((A.k0[yr] + li) % A.world) == A.rank;
            const bool counted = (code == 1 || code == 2) && own;
            if (code == 3 && own) atomicAdd(yr == yr0 ? &s_hist[NG - 1] : &A.wave_glob[yr * NG + NG - 1], 1u);
            unsigned long long steps = 0, rhs = 0;
            if (slot >= 0) {
                steps = (unsigned long long)A.n_time[slot]; rhs = (unsigned long long)A.nfev[slot];
                if (A.flags[slot] & TCR_FLAG_KEPT) kept = 1;
            }
            const int bin = counted ? A.basin[idx] * 12 + A.month[idx] - 1 : 0;
            if (yr == yr0) {
                if (counted) { atomicAdd(&s_hist[bin], 1u); atomicAdd(&s_hist[NG - 2], 1u); }
                if (slot >= 0) { atomicAdd(&s_tot[0], 1ull); atomicAdd(&s_tot[1], steps); atomicAdd(&s_tot[2], rhs); }
            } else {                                        /* block straddles a year boundary: rare */
                if (counted) { atomicAdd(&A.wave_glob[yr * NG + bin], 1u); atomicAdd(&A.wave_glob[yr * NG + NG - 2], 1u); }
                if (slot >= 0) {
                    atomicAdd(&A.wave_loc[yr * 3 + 0], 1ull); atomicAdd(&A.wave_loc[yr * 3 + 1], steps);
                    atomicAdd(&A.wave_loc[yr * 3 + 2], rhs);
                }
            }
        }
        A.att_kept[idx] = kept;
    }
    __syncthreads();
    for (int i = tid; i < NG; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&A.wave_glob[yr0 * NG + i], s_hist[i]);
    if (tid < 3 && s_tot[tid]) atomicAdd(&A.wave_loc[yr0 * 3 + tid], s_tot[tid]);
}

struct SelectArgs {
    int n_years, n_tracks;
    const int64_t* wave_off; const int64_t* k0;
    const int64_t* consumed;    /* [n_years] attempts of the year's range that were fully processed   */
    const int32_t* code; const int32_t* basin; const int32_t* month; const int32_t* att_slot;
    const int32_t* n_time; const int32_t* nfev;
    const uint8_t* att_kept; const unsigned int* wave_glob; const unsigned long long* wave_loc;   /* k_wave_stats */
    int32_t* nt;                /* [n_years] kept so far (in/out)                              */
    int64_t* used;              /* [n_years] attempts consumed by this wave: i*+1, or consumed[y] */
    int32_t* row_slot;          /* [n_years][n_tracks] slot of a row assigned in THIS wave, else -1 */
    double* tc_month; int32_t* tc_basin; double* n_seeds;    /* outputs (device)               */
    tcr_year_stats* stats;      /* [n_years] device accumulators                               */
    const unsigned int* pool_ctl;   /* [1] != 0: the track pool overflowed -- the wave is void, nothing is committed */
    /* scratch of the three selection kernels (a year's attempt range starts on a 256-attempt block boundary) */
    unsigned int* blk_kept;     /* [blocks] kept storms of the block (k_count_kept)                      */
    unsigned int* blk_pref;     /* [blocks] kept storms of the year before the block (k_select_scan)     */
    int64_t* istar;             /* [n_years] attempt of the want-th kept storm, -1: not reached, -2: no work */
    int32_t* total_kept;        /* [n_years] kept storms in the year's processed range                   */
    unsigned long long* sel_acc;    /* [n_years][6] over-shoot: counted, integrated, steps, rhs; kept steps; exhausted */
    unsigned int* sel_hist;     /* [n_years][84] over-shoot counted attempts per (basin, month)          */
};

/* The sequential acceptance `while nt < n_tracks` (util/compute.py:134, 205-207) as an ordered selection in three
 * fully parallel steps (round 1 scanned a year's kept flags with ONE CTA: 5 % of the step at 20 000 tracks per year,
 * where a year is 40 M attempts): kept storms per 256-attempt block -> per-year exclusive scan of the block counts,
 * which also locates i* = the attempt of the want-th kept storm -> every attempt finds its own rank (block prefix +
 * ballot rank), emits its row if rank <= want, or adds itself to the over-shoot statistics if it lies beyond i*.   */
__global__ void __launch_bounds__(256) k_count_kept(const uint8_t* __restrict__ att_kept, int64_t total, unsigned int* __restrict__ blk_kept)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = __syncthreads_count(idx < total && att_kept[idx] != 0);
    if (threadIdx.x == 0) blk_kept[blockIdx.x] = (unsigned int)c;
}

__global__ void __launch_bounds__(1024) k_select_scan(const SelectArgs A)
{
    __shared__ unsigned int warp_tot[32];
    __shared__ unsigned int s_running;
    __shared__ long long s_bstar;
    const int yr = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int r = tid; r < A.n_tracks; r += blockDim.x) A.row_slot[(size_t)yr * A.n_tracks + r] = -1;
    if (A.pool_ctl[1]) return;
    const int64_t off = A.wave_off[yr];
    const int64_t W = A.consumed[yr];
    if (W == 0) { if (tid == 0) { A.istar[yr] = -2; A.total_kept[yr] = 0; } return; }
    const unsigned int want = (unsigned int)(A.n_tracks - A.nt[yr]);
    const int64_t b0 = off >> 8, nb = (W + 255) >> 8;          /* blocks holding processed attempts */
    if (tid == 0) { s_running = 0u; s_bstar = -1; }
    __syncthreads();
    for (int64_t base = 0; base < nb; base += 1024 * 4) {
        unsigned int v[4], sum = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t i = base + (int64_t)tid * 4 + j;
            v[j] = i < nb ? A.blk_kept[b0 + i] : 0u;
            sum += v[j];
        }
        unsigned int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned int u = __shfl_up_sync(TCR_FULL, incl, d); if (lane >= d) incl += u; }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            unsigned int t = warp_tot[lane], sc = t;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { unsigned int u = __shfl_up_sync(TCR_FULL, sc, d); if (lane >= d) sc += u; }
            warp_tot[lane] = sc - t;
        }
        __syncthreads();
        unsigned int pref = s_running + warp_tot[wid] + incl - sum;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t i = base + (int64_t)tid * 4 + j;
            if (i < nb) {
                A.blk_pref[b0 + i] = pref;
                if (want > 0u && pref < want && want <= pref + v[j]) s_bstar = (long long)i;      /* exactly one block qualifies */
                pref += v[j];
            }
        }
        __syncthreads();
        if (tid == blockDim.x - 1) s_running = pref;
        __syncthreads();
    }
    if (wid == 0) {
        /* i* inside block b*: the (want - prefix)-th kept flag of its 256 bytes */
        long long istar = -1;
        if (s_bstar >= 0) {
            const unsigned int need = want - A.blk_pref[b0 + s_bstar];
            const uint8_t* kb = A.att_kept + off + s_bstar * 256;
            unsigned int bits = 0u;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int64_t i = s_bstar * 256 + lane * 8 + j;
                if (i < W && kb[lane * 8 + j]) bits |= 1u << j;
            }
            const unsigned int cnt = __popc(bits);
            unsigned int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { unsigned int u = __shfl_up_sync(TCR_FULL, incl, d); if (lane >= d) incl += u; }
            const unsigned int before = incl - cnt;
            const bool mine = before < need && need <= incl;
            int pos = -1;
            if (mine) pos = lane * 8 + (int)__fns(bits, 0, (int)(need - before));
            const unsigned int who = __ballot_sync(TCR_FULL, mine);
            if (who) istar = s_bstar * 256 + __shfl_sync(TCR_FULL, pos, __ffs((int)who) - 1);
        }
        if (lane == 0) { A.istar[yr] = istar; A.total_kept[yr] = (int32_t)s_running; }
    }
}

__global__ void __launch_bounds__(256) k_select_rows(const SelectArgs A)
{
    __shared__ unsigned int s_warp[8];
    __shared__ unsigned long long s_acc[6];
    __shared__ unsigned int s_hist[TCR_N_BASINS * 12];
    if (A.pool_ctl[1]) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x;
    if (first >= A.wave_off[A.n_years]) return;
    int yr = 0;
    while (yr + 1 < A.n_years && first >= A.wave_off[yr + 1]) ++yr;          /* the whole block belongs to this year */
    const int64_t off = A.wave_off[yr], W = A.consumed[yr];
    const int64_t i_star = A.istar[yr];
    if (i_star == -2) return;
    const int64_t li = first - off + tid;
    const int64_t last = i_star >= 0 ? i_star : W - 1;
    if (first - off >= W) return;                                              /* beyond the processed range */
    const bool over = first - off + 255 > last;                                /* block reaches into the over-shoot */
    if (tid < 6) s_acc[tid] = 0ull;
    if (over) for (int i = tid; i < TCR_N_BASINS * 12; i += blockDim.x) s_hist[i] = 0u;
    const bool in = li < W;
    const bool kept = in && li <= last && A.att_kept[first + tid] != 0;
    const unsigned int kb = __ballot_sync(TCR_FULL, kept);
    if (lane == 0) s_warp[wid] = __popc(kb);
    __syncthreads();
    if (kept) {
        unsigned int rank = A.blk_pref[blockIdx.x] + __popc(kb & ((1u << lane) - 1u)) + 1u;     /* 1-based, attempt order */
        for (int w = 0; w < wid; ++w) rank += s_warp[w];
        const int nt0 = A.nt[yr];
        const int row = nt0 + (int)rank - 1;                   /* rank <= want because li <= i* (or i* was not reached) */
        const int64_t idx = first + tid;
        const int slot = A.att_slot[idx];
        A.row_slot[(size_t)yr * A.n_tracks + row] = slot;
        A.tc_month[(size_t)yr * A.n_tracks + row] = (double)A.month[idx];
        A.tc_basin[(size_t)yr * A.n_tracks + row] = A.basin[idx];
        if (slot >= 0) atomicAdd(&s_acc[4], (unsigned long long)A.n_time[slot]);       /* < 0: another rank's storm */
    }
    if (in && li > last) {                                     /* the over-shoot (last, W): what the sequential loop never ran */
        const int64_t idx = first + tid;
        const int code = A.code[idx], slot = A.att_slot[idx];
        if (code == 1 || code == 2) { atomicAdd(&s_acc[0], 1ull); atomicAdd(&s_hist[A.basin[idx] * 12 + A.month[idx] - 1], 1u); }
        if (code == 3) atomicAdd(&s_acc[5], 1ull);
        if (slot >= 0) {
            atomicAdd(&s_acc[1], 1ull); atomicAdd(&s_acc[2], (unsigned long long)A.n_time[slot]);
            atomicAdd(&s_acc[3], (unsigned long long)A.nfev[slot]);
        }
    }
    __syncthreads();
    if (tid < 6 && s_acc[tid]) atomicAdd(&A.sel_acc[yr * 6 + tid], s_acc[tid]);
    if (over) for (int i = tid; i < TCR_N_BASINS * 12; i += blockDim.x) if (s_hist[i]) atomicAdd(&A.sel_hist[yr * TCR_N_BASINS * 12 + i], s_hist[i]);
}

__global__ void __launch_bounds__(128) k_select_finish(const SelectArgs A)
{
    if (A.pool_ctl[1]) return;
    const int yr = blockIdx.x, tid = threadIdx.x;
    const int64_t i_star = A.istar[yr];
    if (i_star == -2) { if (tid == 0) A.used[yr] = 0; return; }
    const int NG = TCR_N_BASINS * 12 + 2;
    for (int i = tid; i < TCR_N_BASINS * 12; i += blockDim.x)
        A.n_seeds[(size_t)yr * TCR_N_BASINS * 12 + i] += (double)(A.wave_glob[yr * NG + i] - A.sel_hist[yr * TCR_N_BASINS * 12 + i]);
    if (tid == 0) {
        const int64_t W = A.consumed[yr];
        const int64_t last = i_star >= 0 ? i_star : W - 1;
        const unsigned long long* acc = A.sel_acc + yr * 6;
        tcr_year_stats& s = A.stats[yr];
        const unsigned long long* tot = A.wave_loc + yr * 3;
        const int nt0 = A.nt[yr];
        const int want = A.n_tracks - nt0;
        const int got = min(want, A.total_kept[yr]);
        s.attempts = A.k0[yr] + last + 1;
        s.counted_seeds += (int64_t)((unsigned long long)A.wave_glob[yr * NG + NG - 2] - acc[0]);
        s.redraw_exhausted += (int64_t)((unsigned long long)A.wave_glob[yr * NG + NG - 1] - acc[5]);
        s.integrated += (int64_t)(tot[0] - acc[1]);
        s.storm_steps += (int64_t)(tot[1] - acc[2]);
        s.rhs_evals += (int64_t)(tot[2] - acc[3]);
        s.kept_steps += (int64_t)acc[4];
        s.wasted_integrated += (int64_t)acc[1];
        s.wasted_steps += (int64_t)acc[2];
        s.wasted_rhs_evals += (int64_t)acc[3];
        s.n_kept = nt0 + got;
        s.n_waves += 1;
        A.nt[yr] = nt0 + got;
        A.used[yr] = last + 1;
    }
}

/* write the rows assigned in this wave into the caller's 9-tuple layout (compute.py:126-133, 193-207), NaN-padded
 * past n_time.  One warp per (year, row).  env winds and vmax of a kept storm are formed HERE, from its track and
 * Fourier table (the arithmetic of k_postprocess, so the same bits): the wave keeps no env / vmax rows at all.   */
struct GatherArgs {
    int n_years, n_tracks;
    const int32_t* row_slot; const int32_t* n_time;
    const int32_t* ym; const double* ftab; const double2* coef; const double* track; const int32_t* track_row;
    const unsigned int* pool_ctl;
    double* o_lon; double* o_lat; double* o_v; double* o_m; double* o_vmax; double* o_env;
};

__global__ void __launch_bounds__(256) k_gather(const __grid_constant__ TcrCtx cx, const GatherArgs A)
{
    const int ns = cx.p.n_steps;
    if (A.pool_ctl[1]) return;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= (int64_t)A.n_years * A.n_tracks) return;
    const int slot = A.row_slot[row];
    if (slot < 0) return;
    const int lane = threadIdx.x & 31;
    const int nt = A.n_time[slot];
    const int ym = A.ym[slot];
    const double* trk = A.track + (size_t)A.track_row[slot] * ns * 4;
    const double* ftab = A.ftab ? A.ftab + (size_t)slot * ns * 4 : nullptr;
    const double2* coef = A.coef ? A.coef + (size_t)slot * TCR_N_PHASES : nullptr;
    const size_t o = (size_t)row * ns;
    const bool env_al16 = (reinterpret_cast<uintptr_t>(A.o_env) & 15) == 0;
    for (int k = lane; k < ns; k += 32) {
        double2 a = make_double2(NAN, NAN), b = a, e0 = a, e1 = a;
        double vm = NAN;
        if (k < nt) {
            a = *reinterpret_cast<const double2*>(trk + (size_t)k * 4);
            b = *reinterpret_cast<const double2*>(trk + (size_t)k * 4 + 2);
            double w[4];
            vm = tcr_post_sample(cx, ym, ftab, coef, trk, nt, k, w);
            e0 = make_double2(w[0], w[1]);
            e1 = make_double2(w[2], w[3]);
        }
        A.o_lon[o + k] = a.x; A.o_lat[o + k] = a.y; A.o_v[o + k] = b.x; A.o_m[o + k] = b.y;
        A.o_vmax[o + k] = vm;
        double* ep = A.o_env + (o + k) * 4;
        if (env_al16) {
            double2* eo = reinterpret_cast<double2*>(ep);
            eo[0] = e0; eo[1] = e1;
        } else {                        /* the caller's env section starts on an odd double (odd rows x odd n_steps before it) */
            ep[0] = e0.x; ep[1] = e0.y; ep[2] = e1.x; ep[3] = e1.y;
        }
    }
}

/* ======================================================================================== */
/* return-period reduction over the finished-track tensor:                     */
/* notebooks/sample_analysis.ipynb cells 13-17.  One warp per track; a streaming, HBM-bound    */
/* pass (24 B per sample: lon, lat, vmax): the fp64 haversine is only evaluated for samples    */
/* whose latitude is within the radius of the point (a great-circle distance is never shorter  */
/* than R |dlat|, so the 1 % slack below cannot change any result).                            */
/* ======================================================================================== */
#define POI_CHUNK 384
__global__ void __launch_bounds__(256) k_poi_vmax(int64_t n_rows, int n_steps, const double* __restrict__ lon,
                                                  const double* __restrict__ lat, const double* __restrict__ vmax,
                                                  double poi_lon, double poi_lat, double radius_km, double r_km,
                                                  double* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double band_deg = radius_km / r_km * (180.0 / TCR_PI) * 1.01 + 1e-9;
    const float cos_poi = cosf((float)(poi_lat * TCR_DEG2RAD));
    const float s_half = sinf((float)fmin(0.5 * radius_km / r_km, 1.5));
    const float a_max = s_half * s_half * 1.02f + 1e-12f;          /* sin^2(radius / 2R) with slack */
    /* Two phases per chunk of 384 samples, so that the expensive part runs with full warps: (1) twelve independent
     * 256-byte latitude loads per warp in flight, the samples inside the latitude band of the point (6 % in the
     * benchmark) are compacted into a per-warp list in shared memory; (2) the list is worked off 32 at a time:
     * longitude, a float32 exclusion test, the fp64 haversine, vmax.  (Round 1 evaluated the haversine in place, a
     * handful of lanes at a time: 26 % of the HBM roofline.) */
    __shared__ unsigned short s_list[8][POI_CHUNK];
    unsigned short* list = s_list[threadIdx.x >> 5];
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n_rows; row += warps) {
        const size_t base = (size_t)row * n_steps;
        double best = -INFINITY;
        bool have = false;
        for (int k0 = 0; k0 < n_steps; k0 += POI_CHUNK) {
            double la[POI_CHUNK / 32];
#pragma unroll
            for (int u = 0; u < POI_CHUNK / 32; ++u) {
                const int k = k0 + 32 * u + lane;
                la[u] = k < n_steps ? __ldcs(lat + base + k) : INFINITY;      /* +inf is outside every band */
            }
            int cnt = 0;
#pragma unroll
            for (int u = 0; u < POI_CHUNK / 32; ++u) {
                const bool inb = !(fabs(la[u] - poi_lat) > band_deg);         /* NaN stays in, as in the dense formulation */
                const unsigned m = __ballot_sync(TCR_FULL, inb);
                if (inb) list[cnt + __popc(m & ((1u << lane) - 1u))] = (unsigned short)(32 * u + lane);
                cnt += __popc(m);
            }
            __syncwarp();
            for (int i = lane; i < cnt; i += 32) {
                const int k = k0 + list[i];
                const double lat_k = __ldg(lat + base + k);                    /* just read: a cache hit */
                const double lo = __ldcs(lon + base + k);
                /* cheap exclusion in float32: the haversine argument is at least its longitude term
                 * cos(lat1) cos(lat2) sin^2(dlon/2); 2 % slack covers the float32 error many times over */
                const float c2 = __cosf((float)(lat_k * TCR_DEG2RAD));
                const float sh = __sinf((float)((lo - poi_lon) * (0.5 * TCR_DEG2RAD)));
                if (cos_poi * c2 * sh * sh > a_max) continue;
                const double d = tcr_haversine_r(r_km, poi_lon, poi_lat, lo, lat_k);
                if (d <= radius_km) {
                    const double v = __ldcs(vmax + base + k);
                    if (!tcr_isnan(v)) { have = true; if (v > best) best = v; }
                }
            }
            __syncwarp();
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const double o = __shfl_xor_sync(TCR_FULL, best, d);
            const int h2 = __shfl_xor_sync(TCR_FULL, (int)have, d);
            if (o > best) best = o;
            have = have || (h2 != 0);
        }
        if (lane == 0) out[row] = have ? best : NAN;
    }
}

/* exceedance_count[b] = #{ i : v[i] >= bins[b] }  (cell 17); counts accumulate (zeroed by the host) */
__global__ void __launch_bounds__(256) k_exceedance(int64_t n, const double* __restrict__ v, int n_bins,
                                                    const double* __restrict__ bins, unsigned long long* __restrict__ counts)
{
    __shared__ unsigned int s_cnt[64];
    for (int b = threadIdx.x; b < n_bins; b += blockDim.x) s_cnt[b] = 0u;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = v[i];
        for (int b = 0; b < n_bins; ++b)
            if (x >= bins[b]) atomicAdd(&s_cnt[b], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < n_bins; b += blockDim.x)
        if (s_cnt[b]) atomicAdd(&counts[b], (unsigned long long)s_cnt[b]);
}

/* ======================================================================================== */
/* per-month field preparation on the device: what the top of run_tracks does   */
/* with xarray / NumPy per month (util/compute.py:107-121) and BetaAdvectionTrack does per    */
/* field (track/bam_track.py:72-74): NaN policies, PI scaling, chi transform, bilinear regrid  */
/* of the ocean climatologies to the thermo grid (util/mat.py:159-164), basin crop with        */
/* longitude re-wrapping (util/basins.py:57-75, as index maps).  raw [17][nlat_g][nlon_g]:     */
/* 14 wind statistics, vmax, chi, rh_mid; ocean [2][nlat_o][nlon_o]: mld, strat.               */
/* out planes [19][nlat_b][nlon_b] float32 (the input of k_build_month).                       */
/* ======================================================================================== */
struct PrepArgs {
    int nlat_g, nlon_g, nlat_o, nlon_o, nlat_b, nlon_b;
    double pi_reduc, sqrt_ck_cd, log_chi_fac, chi_fac;
    const float* raw; const float* ocean;
    const double* lon_g; const double* lat_g; const double* lon_o; const double* lat_o;
    const int32_t* src_col; const int32_t* src_row;
    float* out;
};

__device__ __forceinline__ void tcr_locate_plain(const double* ax, int n, double x, int& i0, double& w0, double& w1)
{
    double a = x;
    if (a < ax[0]) a = ax[0];
    if (a > ax[n - 1]) a = ax[n - 1];
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (ax[mid] <= a) lo = mid; else hi = mid; }
    if (lo > n - 2) lo = n - 2;
    const double f = 1.0 / (ax[lo + 1] - ax[lo]);
    i0 = lo; w0 = f * (ax[lo + 1] - a); w1 = f * (a - ax[lo]);
}

__device__ __forceinline__ double tcr_nan_to_num(double x)
{
    if (tcr_isnan(x)) return 0.0;
    if (x == INFINITY) return 1.7976931348623157e308;
    if (x == -INFINITY) return -1.7976931348623157e308;
    return x;
}

__global__ void __launch_bounds__(256) k_prepare_month(const PrepArgs A)
{
    const size_t plane_b = (size_t)A.nlat_b * A.nlon_b, plane_g = (size_t)A.nlat_g * A.nlon_g;
    const size_t total = plane_b * TCR_N_FIELDS;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx / plane_b);
        const size_t rem = idx - (size_t)c * plane_b;
        const int i = (int)(rem / A.nlon_b), j = (int)(rem - (size_t)i * A.nlon_b);
        const int r = A.src_row[i], q = A.src_col[j];
        double v;
        if (c < CH_CHI) {                                   /* 14 wind statistics: nan_to_num (bam_track.py:74) */
            v = tcr_nan_to_num((double)A.raw[(size_t)c * plane_g + (size_t)r * A.nlon_g + q]);
        } else if (c == CH_VPOT) {                          /* vmax * PI_reduc * sqrt(Ck/Cd), NaN -> 0 (compute.py:76,110) */
            v = tcr_nan_to_num((double)A.raw[(size_t)14 * plane_g + (size_t)r * A.nlon_g + q] * A.pi_reduc * A.sqrt_ck_cd);
        } else if (c == CH_CHI) {                           /* compute.py:113,115 */
            double x = (double)A.raw[(size_t)15 * plane_g + (size_t)r * A.nlon_g + q];
            if (tcr_isnan(x)) x = 5.0;
            double t = tcr_exp(tcr_log(x + 1e-3) + A.log_chi_fac) + A.chi_fac;
            t = tcr_isnan(t) ? t : (t < 5.0 ? t : 5.0);                     /* np.minimum / np.maximum propagate NaN */
            v = tcr_isnan(t) ? t : (t > 1e-5 ? t : 1e-5);
        } else if (c == CH_RH) {                            /* rh_mid is sampled raw (compute.py:114) */
            v = (double)A.raw[(size_t)16 * plane_g + (size_t)r * A.nlon_g + q];
        } else {                                            /* mld / strat: nan_to_num, then bilinear regrid (compute.py:117-118) */
            const float* f = A.ocean + (size_t)(c - CH_MLD) * A.nlat_o * A.nlon_o;
            int ix, iy; double wx0, wx1, wy0, wy1;
            tcr_locate_plain(A.lon_o, A.nlon_o, A.lon_g[q], ix, wx0, wx1);
            tcr_locate_plain(A.lat_o, A.nlat_o, A.lat_g[r], iy, wy0, wy1);
            const double f00 = tcr_nan_to_num((double)f[(size_t)iy * A.nlon_o + ix]);
            const double f01 = tcr_nan_to_num((double)f[(size_t)(iy + 1) * A.nlon_o + ix]);
            const double f10 = tcr_nan_to_num((double)f[(size_t)iy * A.nlon_o + ix + 1]);
            const double f11 = tcr_nan_to_num((double)f[(size_t)(iy + 1) * A.nlon_o + ix + 1]);
            double sp = f00 * wx0 * wy0;                     /* fields.bilinear / FITPACK product order */
            sp = sp + f01 * wx0 * wy1;
            sp = sp + f10 * wx1 * wy0;
            sp = sp + f11 * wx1 * wy1;
            v = sp;
        }
        A.out[idx] = (float)v;
    }
}
