/*
 * tcrisk.cu -- C ABI (include/tcrisk.h) of the B200-native tropical-cyclone ensemble integrator:
 * device-memory management, table upload and the wave scheduler around the kernels of
 * tcr_kernels.cuh.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false (the
 * arithmetic contract forbids implicit contraction; see include/tcr_libm.h).
 */
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "tcr_kernels.cuh"
#include "tcr_preproc.cuh"

#define TCR_VERSION 100

static thread_local std::string g_err;

static int set_err(const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return -1;
}

#define CK(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return set_err("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
    } while (0)

#define CKK(h)                                                                                      \
    do {                                                                                            \
        cudaError_t e_ = cudaGetLastError();                                                        \
        if (e_ != cudaSuccess)                                                                      \
            return set_err("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
        (h)->launches++;                                                                            \
    } while (0)

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }          /* early error returns (CK) must not leak call-local scratch */
    int ensure(size_t need)
    {
        if (need <= bytes) return 0;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        cudaError_t e = cudaMalloc(&p, need);
        if (e != cudaSuccess) { p = nullptr; return set_err("cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e)); }
        bytes = need;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct AxisBuf {
    DevBuf buf;
    int n = 0;
};

struct TimedSpan { int which; cudaEvent_t a, b; };

/* scratch of tcr_run_years / tcr_integrate, grown on demand and kept in the handle.  Two
 * capacities: seed ATTEMPTS per wave (56 B each) and integrated storm SLOTS per wave
 * (slot_bytes(n_steps) each: Fourier table, track, env winds, vmax).                          */
struct Workspace {
    int64_t att_cap = 0, slot_cap = 0;
    int64_t full_cap = 0;          /* slots that also own track / env / vmax rows (tcr_integrate)           */
    int64_t pool_rows = 0;         /* rows of the track pool (tcr_run_years): lanes in flight + candidates   */
    bool mem_limited = false;      /* the capacities were set by the memory budget, not by the job            */
    int ns = 0;
    int ring_nodes = 0;            /* nodes per Fourier ring (0: tcr_run_years tabulates full tables)         */
    DevBuf code, basin, month, att_slot, a_lon, a_lat, a_v0, a_m0;   /* per attempt */
    DevBuf blk_count, blk_off;                                       /* per 256-attempt seed block */
    DevBuf att_kept, wave_tot;                                       /* k_wave_stats: byte per attempt; per-year totals + histogram */
    DevBuf s_ym, s_lon, s_lat, s_v0, s_m0, s_hbl, s_att, s_key;       /* per slot */
    DevBuf n_time, status, nfev, flags, cand, track_row;
    DevBuf coef, ftab, ring, track, env, vmax;
    DevBuf counters;       /* [0] queue (u64), u32 @+8 n_slots, @+12 cand_count, @+16 n_pass, @+20 pool next, @+24 pool overflow */
    DevBuf year_i64;       /* wave_off [ny+1], k0 [ny], consumed [ny], used [ny] */
    DevBuf year_i32;       /* ym_base [ny], year_key [ny], nt [ny] */
    DevBuf row_slot, stats;
    DevBuf out;            /* device-side result block when the caller passes host pointers */
    void release_all()
    {
        DevBuf* all[] = {&code, &basin, &month, &att_slot, &a_lon, &a_lat, &a_v0, &a_m0, &blk_count, &blk_off, &att_kept, &wave_tot,
                         &s_ym, &s_lon, &s_lat, &s_v0, &s_m0, &s_hbl, &s_att, &s_key,
                         &n_time, &status, &nfev, &flags, &cand, &track_row, &coef, &ftab, &ring, &track, &env, &vmax, &counters, &year_i64,
                         &year_i32, &row_slot, &stats, &out};
        for (DevBuf* b : all) b->release();
        att_cap = slot_cap = full_cap = pool_rows = 0; mem_limited = false; ring_nodes = 0;
    }
};

}  // namespace

struct tcr_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    TcrCtx ctx;
    int num_sms = 148;
    size_t smem_optin = 0;
    int64_t launches = 0;
    /* tables */
    DevBuf rec, stage, prep;
    AxisBuf ax_lon, ax_lat;
    int nlat = 0, nlon = 0, n_ym = 0;
    size_t month_f4 = 0;
    DevBuf sincos;               /* [n_steps][15] double2, storm-independent harmonics */
    /* static */
    DevBuf bathy, land, masks;
    DevBuf etab;                 /* entropy look-up table of the thermo kernel: p_look | s_look | T_lookup */
    int etab_np = 0, etab_ns = 0;
    DevBuf etab3;                /* reversible table (select_thermo = 2): p_look | s_look | rt_look | T_lookup [np][ns][nr] */
    int etab3_np = 0, etab3_ns = 0, etab3_nr = 0;
    AxisBuf ax_lon_b, ax_lat_b, ax_lon_l, ax_lat_l, ax_lon_m, ax_lat_m;
    bool have_static = false, have_masks = false;
    /* tuning */
    int integ_variant = 21, oversub_permille = 1020, interp_variant = 0;
    double ws_frac = 0.6; size_t ws_cap = (size_t)96 << 30;   /* tcr_set_workspace_budget */
    int ftab_ring = -1;          /* tcr_run_years: Fourier rings filled by the integrator (1), full tables tabulated ahead of it (0), or
                                  * by the length of the output grid (-1, default: ring_nodes_for) */
    /* within-year sharding (tcr_set_shard): rank r of `world` integrates the attempts k with k % world == r */
    int shard_rank = 0, shard_world = 1;
    tcr_allreduce_fn allreduce = nullptr;
    void* allreduce_user = nullptr;
    bool use_recp = false;       /* a REC = 2 variant is selected: point records are derived from the cell records before each run */
    bool recp_stale = true;       /* a REC = 1 variant is selected: tables also exist as integrator records */
    /* survival statistics of earlier tcr_run_years calls on this handle: size the first wave */
    double hint_kept_rate = 0.0, hint_pass_rate = 0.0;
    std::vector<double> hint_year_rate;     /* per year slot of the previous call (same n_years): survival differs by year */
    int64_t max_wave = 0, max_slots = 0;
    Workspace ws;
    void* pinned = nullptr;      /* small pinned read-back area */
    /* device-time accounting (tcr_set_timing): event pairs around every launch of a kernel class */
    bool timing = false;
    std::vector<TimedSpan> spans;
    std::vector<cudaEvent_t> free_events;
    double class_ms[TCR_N_KERNEL_CLASSES] = {0};
    int64_t class_launches[TCR_N_KERNEL_CLASSES] = {0};
};

static cudaEvent_t take_event(tcr_handle* h)
{
    cudaEvent_t e = nullptr;
    if (!h->free_events.empty()) { e = h->free_events.back(); h->free_events.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}

/* brackets one kernel launch with events on the handle's stream when timing is on */
struct LaunchTimer {
    tcr_handle* h; int which; cudaEvent_t a = nullptr;
    LaunchTimer(tcr_handle* h_, int which_) : h(h_), which(which_)
    {
        if (h->timing) { a = take_event(h); cudaEventRecord(a, h->stream); }
    }
    ~LaunchTimer()
    {
        if (a) { cudaEvent_t b = take_event(h); cudaEventRecord(b, h->stream); h->spans.push_back({which, a, b}); }
    }
};

static int resolve_spans(tcr_handle* h)
{
    if (h->spans.empty()) return 0;
    CK(cudaStreamSynchronize(h->stream));
    for (const TimedSpan& sp : h->spans) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, sp.a, sp.b));
        h->class_ms[sp.which] += ms;
        h->class_launches[sp.which] += 1;
        h->free_events.push_back(sp.a);
        h->free_events.push_back(sp.b);
    }
    h->spans.clear();
    return 0;
}

static int make_axis(tcr_handle* h, const double* host, int n, AxisBuf& ab, TcrAxis& ax)
{
    if (n < 2) return set_err("axis needs at least 2 points (got %d)", n);
    for (int i = 1; i < n; ++i)
        if (!(host[i] > host[i - 1])) return set_err("axis must be strictly ascending (index %d)", i);
    DevBuf tmp;
    if (tmp.ensure(sizeof(double) * n)) return -1;
    if (ab.buf.ensure(sizeof(TcrNode) * n)) { tmp.release(); return -1; }
    cudaError_t e = cudaMemcpyAsync(tmp.p, host, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        k_build_axis<<<(n + 255) / 256, 256, 0, h->stream>>>(tmp.as<double>(), ab.buf.as<TcrNode>(), n);
        e = cudaGetLastError();
        h->launches++;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    tmp.release();
    if (e != cudaSuccess) return set_err("axis upload failed: %s", cudaGetErrorString(e));
    ab.n = n;
    ax.a = ab.buf.as<TcrNode>();
    ax.n = n;
    ax.lo = host[0];
    ax.hi = host[n - 1];
    ax.inv_d = (double)(n - 1) / (host[n - 1] - host[0]);
    /* arithmetic nodes only if they reproduce the tabulated ones bit for bit */
    const volatile double dx = host[1] - host[0];
    const volatile double inv_dx = 1.0 / dx;
    bool uniform = true;
    for (int i = 0; i < n && uniform; ++i) {
        const volatile double prod = (double)i * dx;          /* two roundings, as on the device (-fmad=false) */
        const volatile double xi = host[0] + prod;
        if (xi != host[i]) uniform = false;
        if (i + 1 < n) {
            const volatile double d = host[i + 1] - host[i];
            const volatile double inv = 1.0 / d;
            if (inv != inv_dx) uniform = false;
        }
    }
    ax.uniform = uniform ? 1 : 0;
    ax.dx = dx;
    ax.inv_dx = inv_dx;
    return 0;
}

static int grid_for(size_t total, int block, int num_sms)
{
    size_t g = (total + block - 1) / block;
    size_t cap = (size_t)num_sms * 32;
    return (int)std::max<size_t>(1, std::min(g, cap));
}

extern "C" {

const char* tcr_last_error(void) { return g_err.c_str(); }
int tcr_version(void) { return TCR_VERSION; }

int tcr_create(int device, const tcr_params* p, tcr_handle** out)
{
    if (!p || !out) return set_err("tcr_create: null argument");
    if (p->n_steps < 2) return set_err("tcr_create: n_steps must be >= 2");
    int n_dev = 0;
    CK(cudaGetDeviceCount(&n_dev));
    if (device < 0 || device >= n_dev) return set_err("tcr_create: device %d not present (%d devices)", device, n_dev);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return set_err("tcr_create: this library is built for sm_100a (Blackwell B200); device %d is sm_%d%d", device, prop.major, prop.minor);
    tcr_handle* h = new tcr_handle();
    h->device = device;
    if (const char* iv = getenv("TCR_INTEG_VARIANT")) {                 /* A/B runs and tests of a non-default integrate variant */
        const int v = atoi(iv);
        if (v >= 1 && v <= 32) h->integ_variant = v - 1;
    }
    if (const char* fr = getenv("TCR_FTAB_RING")) h->ftab_ring = atoi(fr) < 0 ? -1 : atoi(fr) != 0;   /* A/B runs; a test runs both in one process */
    h->num_sms = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    memset(&h->ctx, 0, sizeof h->ctx);
    h->ctx.p = *p;
    h->ctx.t_step = p->total_time / (double)(p->n_steps - 1);
    if (cudaMallocHost(&h->pinned, 1 << 16) != cudaSuccess) { delete h; return set_err("cudaMallocHost failed"); }
    if (h->sincos.ensure((size_t)2 * p->n_steps * TCR_N_HARM * sizeof(double2))) { cudaFreeHost(h->pinned); delete h; return -1; }
    h->ctx.sc = h->sincos.as<double2>();
    h->ctx.sct = h->ctx.sc + (size_t)p->n_steps * TCR_N_HARM;
    h->ctx.inv_t_step = 1.0 / h->ctx.t_step;
    k_build_sincos<<<(p->n_steps + 127) / 128, 128, 0, h->stream>>>(h->ctx, h->sincos.as<double2>(),
                                                                    h->sincos.as<double2>() + (size_t)p->n_steps * TCR_N_HARM);
    double* d_consts = reinterpret_cast<double*>(h->pinned);          /* pinned memory is device-accessible (UVA) */
    k_build_consts<<<1, 32, 0, h->stream>>>(p->earth_R, p->gen_lat_min, p->gen_lat_max, d_consts);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    h->ctx.y_earth_R = d_consts[0];
    h->ctx.y_pi = d_consts[1];
    h->ctx.gen_y_min = d_consts[2];
    h->ctx.gen_y_max = d_consts[3];
    if (e != cudaSuccess) {
        set_err("tcr_create: harmonic table build failed: %s", cudaGetErrorString(e));
        h->sincos.release(); cudaFreeHost(h->pinned); delete h;
        return -1;
    }
    h->launches++;
    *out = h;
    return 0;
}

int tcr_destroy(tcr_handle* h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    h->ws.release_all();
    DevBuf* bufs[] = {&h->rec, &h->stage, &h->prep, &h->bathy, &h->land, &h->masks, &h->sincos};
    for (DevBuf* b : bufs) b->release();
    AxisBuf* axs[] = {&h->ax_lon, &h->ax_lat, &h->ax_lon_b, &h->ax_lat_b, &h->ax_lon_l, &h->ax_lat_l, &h->ax_lon_m, &h->ax_lat_m};
    for (AxisBuf* a : axs) a->buf.release();
    if (h->pinned) cudaFreeHost(h->pinned);
    for (const TimedSpan& sp : h->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (cudaEvent_t e : h->free_events) cudaEventDestroy(e);
    delete h;
    return 0;
}

int tcr_set_stream(tcr_handle* h, void* cuda_stream)
{
    if (!h) return set_err("null handle");
    h->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return 0;
}

int tcr_synchronize(tcr_handle* h)
{
    if (!h) return set_err("null handle");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

int tcr_set_tuning(tcr_handle* h, int integ_variant, int64_t max_wave_cands, int64_t max_wave_slots, int oversub_permille)
{
    if (!h) return set_err("null handle");
    if (integ_variant > 0) {
        if (integ_variant > 32) return set_err("integrate variant must be 1 (256 thr x 1 CTA/SM), 2 (128 x 3), 3 (128 x 4) or 4 (160 x 2)");
        h->integ_variant = integ_variant - 1;
    }
    if (max_wave_cands > 0) h->max_wave = max_wave_cands;
    if (max_wave_slots > 0) h->max_slots = max_wave_slots;
    if (oversub_permille > 0) h->oversub_permille = oversub_permille;
    return 0;
}

static int ring_nodes_for(const tcr_handle* h);
int tcr_fourier_ring_nodes(tcr_handle* h)
{
    if (!h) return set_err("null handle");
    return ring_nodes_for(h);
}

int tcr_set_workspace_budget(tcr_handle* h, double fraction_of_free, int64_t cap_bytes)
{
    if (!h) return set_err("null handle");
    if (!(fraction_of_free > 0.0 && fraction_of_free <= 0.95) || cap_bytes <= 0)
        return set_err("tcr_set_workspace_budget: fraction in (0, 0.95], cap_bytes > 0");
    h->ws_frac = fraction_of_free;
    h->ws_cap = (size_t)cap_bytes;
    h->ws.mem_limited = false;                       /* the next tcr_run_years sizes its workspace again */
    return 0;
}

int tcr_set_interp_variant(tcr_handle* h, int variant)
{
    if (!h) return set_err("null handle");
    if (variant < 0 || variant > 6 || variant == 1 || variant == 4)
        return set_err("interp variant must be 0 (LDG), 2 / 3 (LDG, 4 / 5 CTAs per SM) or 5 / 6 (cp.async tile of 256 / 128 queries); 1 and 4 were removed");
    h->interp_variant = variant;
    return 0;
}

int tcr_set_shard(tcr_handle* h, int rank, int world, tcr_allreduce_fn fn, void* user)
{
    if (!h) return set_err("null handle");
    if (world < 1 || rank < 0 || rank >= world) return set_err("tcr_set_shard: rank %d of %d", rank, world);
    if (world > 1 && !fn) return set_err("tcr_set_shard: world > 1 needs an all-reduce callback");
    h->shard_rank = rank; h->shard_world = world; h->allreduce = fn; h->allreduce_user = user;
    return 0;
}

int64_t tcr_launch_count(tcr_handle* h) { return h ? h->launches : 0; }

int tcr_set_timing(tcr_handle* h, int enable)
{
    if (!h) return set_err("null handle");
    CK(cudaSetDevice(h->device));
    if (resolve_spans(h)) return -1;
    h->timing = enable != 0;
    for (int i = 0; i < TCR_N_KERNEL_CLASSES; ++i) { h->class_ms[i] = 0.0; h->class_launches[i] = 0; }
    return 0;
}

int tcr_kernel_time(tcr_handle* h, int kernel_class, double* ms, int64_t* launches)
{
    if (!h) return set_err("null handle");
    if (kernel_class < 0 || kernel_class >= TCR_N_KERNEL_CLASSES) return set_err("tcr_kernel_time: unknown kernel class %d", kernel_class);
    CK(cudaSetDevice(h->device));
    if (resolve_spans(h)) return -1;
    if (ms) *ms = h->class_ms[kernel_class];
    if (launches) *launches = h->class_launches[kernel_class];
    return 0;
}

#ifdef TCR_DEBUG_BAD
/* instrumented builds only (scripts/probes/bad_sites.py): how often each operation of the straight-line RHS left its common case */
int tcr_debug_bad(unsigned long long* out, int reset)
{
    if (cudaMemcpyFromSymbol(out, tcr_dbg_bad, sizeof(unsigned long long) * 33) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[33] = {0}; cudaMemcpyToSymbol(tcr_dbg_bad, z, sizeof z); }
    return 0;
}
#endif

int tcr_host_alloc(size_t bytes, void** out)
{
    if (!out) return set_err("null argument");
    CK(cudaMallocHost(out, bytes));
    return 0;
}

int tcr_host_free(void* p)
{
    if (p) CK(cudaFreeHost(p));
    return 0;
}

/* ---- static fields ------------------------------------------------------------------------- */
int tcr_upload_static(tcr_handle* h, int nlat_b, int nlon_b, const double* lat_b, const double* lon_b, const int16_t* bathy,
                      int nlat_l, int nlon_l, const double* lat_l, const double* lon_l, const int8_t* land)
{
    if (!h || !lat_b || !lon_b || !bathy || !lat_l || !lon_l || !land) return set_err("tcr_upload_static: null argument");
    CK(cudaSetDevice(h->device));
    TcrStatic& st = h->ctx.st;
    if (make_axis(h, lon_b, nlon_b, h->ax_lon_b, st.lon_b) || make_axis(h, lat_b, nlat_b, h->ax_lat_b, st.lat_b) ||
        make_axis(h, lon_l, nlon_l, h->ax_lon_l, st.lon_l) || make_axis(h, lat_l, nlat_l, h->ax_lat_l, st.lat_l))
        return -1;
    DevBuf tmp;
    {
        size_t n = (size_t)nlat_b * nlon_b, nc = (size_t)(nlat_b - 1) * (nlon_b - 1);
        if (tmp.ensure(n * sizeof(int16_t)) || h->bathy.ensure(nc * sizeof(short4))) { tmp.release(); return -1; }
        CK(cudaMemcpyAsync(tmp.p, bathy, n * sizeof(int16_t), cudaMemcpyHostToDevice, h->stream));
        k_build_bathy<<<grid_for(nc, 256, h->num_sms), 256, 0, h->stream>>>(tmp.as<int16_t>(), h->bathy.as<short4>(), nlat_b, nlon_b);
        CKK(h);
        CK(cudaStreamSynchronize(h->stream));
        st.bathy = h->bathy.as<short4>();
        st.ncx_b = nlon_b - 1;
    }
    {
        size_t n = (size_t)nlat_l * nlon_l, nc = (size_t)(nlat_l - 1) * (nlon_l - 1);
        if (tmp.ensure(n) || h->land.ensure(nc * sizeof(char4))) { tmp.release(); return -1; }
        CK(cudaMemcpyAsync(tmp.p, land, n, cudaMemcpyHostToDevice, h->stream));
        k_build_land<<<grid_for(nc, 256, h->num_sms), 256, 0, h->stream>>>(tmp.as<int8_t>(), h->land.as<char4>(), nlat_l, nlon_l);
        CKK(h);
        CK(cudaStreamSynchronize(h->stream));
        st.land = h->land.as<char4>();
        st.ncx_l = nlon_l - 1;
    }
    tmp.release();
    h->have_static = true;
    return 0;
}

int tcr_upload_masks(tcr_handle* h, int nlat_m, int nlon_m, const double* lat_m, const double* lon_m, const uint8_t* masks)
{
    if (!h || !lat_m || !lon_m || !masks) return set_err("tcr_upload_masks: null argument");
    CK(cudaSetDevice(h->device));
    TcrMasks& mk = h->ctx.mk;
    if (make_axis(h, lon_m, nlon_m, h->ax_lon_m, mk.lon) || make_axis(h, lat_m, nlat_m, h->ax_lat_m, mk.lat)) return -1;
    DevBuf tmp;
    size_t n = (size_t)TCR_N_MASKS * nlat_m * nlon_m, nc = (size_t)(nlat_m - 1) * (nlon_m - 1) * 4;
    if (tmp.ensure(n) || h->masks.ensure(nc * sizeof(uint2))) { tmp.release(); return -1; }
    CK(cudaMemcpyAsync(tmp.p, masks, n, cudaMemcpyHostToDevice, h->stream));
    k_build_masks<<<grid_for(nc, 256, h->num_sms), 256, 0, h->stream>>>(tmp.as<uint8_t>(), h->masks.as<uint2>(), nlat_m, nlon_m);
    CKK(h);
    CK(cudaStreamSynchronize(h->stream));
    tmp.release();
    mk.rec = h->masks.as<uint2>();
    mk.ncx = nlon_m - 1;
    h->have_masks = true;
    return 0;
}

/* ---- monthly tables ------------------------------------------------------------------------ */
int tcr_alloc_tables(tcr_handle* h, int n_ym, int nlat, int nlon, const double* lat, const double* lon)
{
    if (!h || !lat || !lon) return set_err("tcr_alloc_tables: null argument");
    if (n_ym < 1 || nlat < 2 || nlon < 2) return set_err("tcr_alloc_tables: bad shape");
    CK(cudaSetDevice(h->device));
    TcrTables& tb = h->ctx.tab;
    if (make_axis(h, lon, nlon, h->ax_lon, tb.lon) || make_axis(h, lat, nlat, h->ax_lat, tb.lat)) return -1;
    h->month_f4 = (size_t)(nlat - 1) * (nlon - 1) * TCR_REC_F4;
    if (h->rec.ensure(h->month_f4 * sizeof(float4) * (size_t)n_ym)) return -1;
    if (h->stage.ensure((size_t)TCR_N_FIELDS * nlat * nlon * sizeof(float))) return -1;
    h->nlat = nlat; h->nlon = nlon; h->n_ym = n_ym;
    tb.rec = h->rec.as<float4>();
    tb.ncx = nlon - 1; tb.ncy = nlat - 1; tb.n_ym = n_ym;
    return 0;
}

int tcr_upload_month_dev(tcr_handle* h, int ym, const float* d_planes)
{
    if (!h || !d_planes) return set_err("tcr_upload_month_dev: null argument");
    if (!h->rec.p) return set_err("tcr_upload_month: call tcr_alloc_tables first");
    if (ym < 0 || ym >= h->n_ym) return set_err("tcr_upload_month: ym %d out of range [0, %d)", ym, h->n_ym);
    CK(cudaSetDevice(h->device));
    {
        LaunchTimer lt_(h, TCR_K_BUILD);
        k_build_month<<<grid_for(h->month_f4, 256, h->num_sms), 256, 0, h->stream>>>(
            d_planes, h->rec.as<float4>() + h->month_f4 * (size_t)ym, h->nlat, h->nlon);
    }
    CKK(h);
    return 0;
}

int tcr_upload_months(tcr_handle* h, int ym0, int n_months, const float* planes)
{
    if (!h || !planes) return set_err("tcr_upload_months: null argument");
    if (!h->rec.p) return set_err("tcr_upload_months: call tcr_alloc_tables first");
    if (n_months <= 0) return 0;
    if (ym0 < 0 || ym0 + n_months > h->n_ym) return set_err("tcr_upload_months: months [%d, %d) out of range [0, %d)", ym0, ym0 + n_months, h->n_ym);
    CK(cudaSetDevice(h->device));
    const size_t plane = (size_t)h->nlat * h->nlon, bytes = plane * TCR_N_FIELDS * sizeof(float) * (size_t)n_months;
    if (h->stage.ensure(bytes)) return -1;
    CK(cudaMemcpyAsync(h->stage.p, planes, bytes, cudaMemcpyHostToDevice, h->stream));
    {
        LaunchTimer lt_(h, TCR_K_BUILD);
        dim3 grid((unsigned)grid_for(h->month_f4, 256, h->num_sms), (unsigned)n_months);
        k_build_month<<<grid, 256, 0, h->stream>>>(h->stage.as<float>(), h->rec.as<float4>() + h->month_f4 * (size_t)ym0, h->nlat, h->nlon);
    }
    CKK(h);
    return 0;
}

int tcr_upload_month(tcr_handle* h, int ym, const float* const* fields)
{
    if (!h || !fields) return set_err("tcr_upload_month: null argument");
    if (!h->rec.p) return set_err("tcr_upload_month: call tcr_alloc_tables first");
    CK(cudaSetDevice(h->device));
    const size_t plane = (size_t)h->nlat * h->nlon;
    bool contiguous = true;
    for (int i = 0; i < TCR_N_FIELDS; ++i) {
        if (!fields[i]) return set_err("tcr_upload_month: field %d is null", i);
        if (i && fields[i] != fields[i - 1] + plane) contiguous = false;
    }
    float* st = h->stage.as<float>();
    if (contiguous) {
        CK(cudaMemcpyAsync(st, fields[0], plane * TCR_N_FIELDS * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    } else {
        for (int i = 0; i < TCR_N_FIELDS; ++i)
            CK(cudaMemcpyAsync(st + plane * i, fields[i], plane * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    }
    return tcr_upload_month_dev(h, ym, st);
}

/* ---- stand-alone bilinear sampler ------------------------------------------------------------ */
static int launch_env_interp(tcr_handle* h, int64_t n, const int32_t* ym, const double* lon, const double* lat, double* out)
{
    LaunchTimer lt_(h, TCR_K_ENV_INTERP);
    if (h->interp_variant == 5 || h->interp_variant == 6) {
        const int tile = h->interp_variant == 5 ? 256 : 128;
        const int smem = tile * (TCR_REC_F4 * 16 + (int)sizeof(EiLoc));
        int64_t tiles = (n + tile - 1) / tile;
        if (tile == 256) {
            CK(cudaFuncSetAttribute(k_env_interp_async<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            k_env_interp_async<256><<<(unsigned)tiles, 256, smem, h->stream>>>(h->ctx, n, ym, lon, lat, out);
        } else {
            CK(cudaFuncSetAttribute(k_env_interp_async<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            k_env_interp_async<128><<<(unsigned)tiles, 256, smem, h->stream>>>(h->ctx, n, ym, lon, lat, out);
        }
    } else {
        int64_t tiles = (n + EI_TILE - 1) / EI_TILE;
        if (h->interp_variant == 2) k_env_interp<4><<<(unsigned)tiles, EI_TILE, 0, h->stream>>>(h->ctx, n, ym, lon, lat, out);
        else if (h->interp_variant == 3) k_env_interp<5><<<(unsigned)tiles, EI_TILE, 0, h->stream>>>(h->ctx, n, ym, lon, lat, out);
        else k_env_interp<3><<<(unsigned)tiles, EI_TILE, 0, h->stream>>>(h->ctx, n, ym, lon, lat, out);
    }
    CKK(h);
    return 0;
}

int tcr_env_interp(tcr_handle* h, int64_t n, const int32_t* ym, const double* lon, const double* lat, double* out, int on_device)
{
    if (!h) return set_err("null handle");
    if (n < 0) return set_err("tcr_env_interp: negative n");
    if (n == 0) return 0;
    if (!ym || !lon || !lat || !out) return set_err("tcr_env_interp: null argument");
    if (!h->rec.p || !h->have_static) return set_err("tcr_env_interp: tables / static fields not uploaded");
    if (n > (int64_t)2147483647 * EI_TILE) return set_err("tcr_env_interp: n too large");
    CK(cudaSetDevice(h->device));
    if (on_device) return launch_env_interp(h, n, ym, lon, lat, out);
    DevBuf in, o;
    if (in.ensure((size_t)n * 20) || o.ensure((size_t)n * TCR_N_INTERP_OUT * sizeof(double))) { in.release(); o.release(); return -1; }
    double* d_lon = in.as<double>();
    double* d_lat = d_lon + n;
    int32_t* d_ym = reinterpret_cast<int32_t*>(d_lat + n);
    /* ym is range-checked on the host path; device callers own their indices */
    for (int64_t i = 0; i < n; ++i)
        if (ym[i] < 0 || ym[i] >= h->n_ym) { in.release(); o.release(); return set_err("tcr_env_interp: ym[%lld]=%d out of range", (long long)i, ym[i]); }
    int rc = 0;
    cudaError_t e = cudaMemcpyAsync(d_lon, lon, n * 8, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_lat, lat, n * 8, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_ym, ym, n * 4, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) rc = launch_env_interp(h, n, d_ym, d_lon, d_lat, o.as<double>());
    if (e == cudaSuccess && rc == 0) e = cudaMemcpyAsync(out, o.p, (size_t)n * TCR_N_INTERP_OUT * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && rc == 0) e = cudaStreamSynchronize(h->stream);
    in.release(); o.release();
    if (rc) return rc;
    if (e != cudaSuccess) return set_err("tcr_env_interp: %s", cudaGetErrorString(e));
    return 0;
}

/* ---- workspace ----------------------------------------------------------------------------- */
/* per integrated slot: 60 coefficient pairs, the Fourier table (32 B per node), ~100 B of scalars, and its share
 * of the track pool (one row per kPoolDiv slots: candidates are 2-3 % of the storms)                       */
static const int kPoolDiv = 8;
static size_t slot_bytes(int ns, int ring_nodes)
{
    return 960 + 104 + (ring_nodes ? (size_t)ring_nodes * 32 / kPoolDiv : (size_t)ns * 32) + (size_t)ns * 32 / kPoolDiv;
}

/* Fourier ring of tcr_run_years (IntegArgs::ring): nodes per ring, or 0 for full tables.  A segment (half a ring) must exceed
 * the nodes one RK attempt can span by 4 (k_integrate, ring_serve).
 * Default: the ring where the output grid is at least four rings long.  Measured on one box (profiles/r02_ring_ab.txt): a
 * node tabulated inside the integrator costs about 3.6 x a node of the stand-alone DMMA kernel (two dependent L2 round trips and
 * ~400 instructions per 64-node segment, issued by a warp whose CTA partners wait at the next slot barrier), and storms use
 * 30-40 % of their table -- so at 361 hourly nodes (rings of 128) the ring loses (configs[1] 12.3 against 11.4 ms per step,
 * configs[2] 737 against 726 ms although it runs in 3 waves instead of 9), at 1441 nodes (900-s output, rings of 256), where the
 * tabulation kernel is 39 % of the step, it wins (300 against 345 ms).  TCR_FTAB_RING=1 / 0 at tcr_create forces either.    */
static int ring_nodes_for(const tcr_handle* h)
{
    if (!h->ftab_ring) return 0;
    const tcr_params& p = h->ctx.p;
    if (!(p.max_step > 0.0) || !(h->ctx.t_step > 0.0)) return 0;
    const double span = std::ceil(p.max_step / h->ctx.t_step) + 1.0;
    int seg = 64;
    while ((double)seg < span + 4.0 && seg < (1 << 20)) seg *= 2;
    if (2 * seg >= p.n_steps) return 0;                     /* a ring as long as the table saves nothing */
    if (h->ftab_ring < 0 && 8 * seg > p.n_steps) return 0;
    return 2 * seg;
}
static const size_t kAttemptBytes = 56;
static int64_t max_lanes(const tcr_handle* h) { return (int64_t)h->num_sms * 576; }   /* most threads per SM of any variant */

/* full: every slot owns a track / env / vmax row (tcr_integrate returns them for every storm); otherwise tracks
 * live in the pool and env winds / vmax are never stored per slot (k_gather forms them for the rows it emits) */
enum WsMode { WS_SLOTS = 0, WS_FULL = 1, WS_POOL = 2 };     /* slots only (coefficients, Fourier tables) / + per-slot rows / + track pool */
static int ws_ensure(tcr_handle* h, int64_t att_cap, int64_t slot_cap, int n_years, WsMode mode, int ring_nodes = 0)
{
    const bool full = mode == WS_FULL;
    Workspace& w = h->ws;
    const int ns = h->ctx.p.n_steps;
    if (w.ns != ns) { w.release_all(); w.ns = ns; }
    /* full Fourier tables for the slots THIS call asked for (a workspace grown by ring-mode years has none) */
    if (!ring_nodes && w.ftab.ensure((size_t)std::max<int64_t>(slot_cap, 1) * ns * 32)) return -1;
    if (att_cap > w.att_cap) {
        const size_t c = (size_t)att_cap, nb = (c + 255) / 256;
        if (w.code.ensure(c * 4) || w.basin.ensure(c * 4) || w.month.ensure(c * 4) || w.att_slot.ensure(c * 4) ||
            w.a_lon.ensure(c * 8) || w.a_lat.ensure(c * 8) || w.a_v0.ensure(c * 8) || w.a_m0.ensure(c * 8) ||
            w.blk_count.ensure(nb * 4) || w.blk_off.ensure(nb * 4) || w.att_kept.ensure(c + 16))
            return -1;
        w.att_cap = att_cap;
    }
    if (slot_cap > w.slot_cap) {
        const size_t c = (size_t)slot_cap;
        if (w.s_ym.ensure(c * 4) || w.s_lon.ensure(c * 8) || w.s_lat.ensure(c * 8) || w.s_v0.ensure(c * 8) ||
            w.s_m0.ensure(c * 8) || w.s_hbl.ensure(c * 8) || w.s_att.ensure(c * 8) || w.s_key.ensure(c * 4) ||
            w.n_time.ensure(c * 4) || w.status.ensure(c * 4) || w.nfev.ensure(c * 4) || w.flags.ensure(c * 4) ||
            w.cand.ensure(c * 4) || w.track_row.ensure(c * 4) || w.coef.ensure(c * TCR_N_PHASES * sizeof(double2)))
            return -1;
        w.slot_cap = slot_cap;
    }
    if (full) {
        if (slot_cap > w.full_cap) {
            const size_t c = (size_t)slot_cap;
            if (w.track.ensure(c * ns * 32) || w.env.ensure(c * ns * 32) || w.vmax.ensure(c * ns * 8)) return -1;
            w.full_cap = slot_cap;
        }
    } else if (mode == WS_POOL) {
        const int64_t rows = max_lanes(h) + std::max<int64_t>(4096, w.slot_cap / kPoolDiv);
        if (rows > w.pool_rows || w.track.bytes < (size_t)rows * ns * 32) {
            if (w.track.ensure((size_t)rows * ns * 32)) return -1;
            w.pool_rows = rows;
            w.full_cap = std::min<int64_t>(w.full_cap, (int64_t)(w.track.bytes / ((size_t)ns * 32)));
        }
        w.pool_rows = std::min<int64_t>((int64_t)(w.track.bytes / ((size_t)ns * 32)), 0x7fffffff);
        if (ring_nodes) {
            if (w.ring.ensure((size_t)w.pool_rows * ring_nodes * 32)) return -1;
            w.pool_rows = std::min<int64_t>(w.pool_rows, (int64_t)(w.ring.bytes / ((size_t)ring_nodes * 32)));
        }
        w.ring_nodes = ring_nodes;
    }
    if (w.counters.ensure(64)) return -1;
    const size_t ny = (size_t)std::max(n_years, 1);
    /* k_wave_stats totals + histogram, then the selection scratch: i*, kept totals, over-shoot accumulators + histogram */
    if (w.wave_tot.ensure(ny * (3 * 8 + (TCR_N_BASINS * 12 + 2) * 4) + 16 + ny * (8 + 8 + 6 * 8 + TCR_N_BASINS * 12 * 4) + 64)) return -1;
    if (w.year_i64.ensure((4 * ny + 1) * 8) || w.year_i32.ensure(3 * ny * 4) || w.stats.ensure(ny * sizeof(tcr_year_stats))) return -1;
    return 0;
}

/* Fourier tables of slots [0, n) (n on the host) or [0, *n_dev) */
static int launch_fourier_table(tcr_handle* h, int64_t n_upper, const unsigned int* n_dev)
{
    Workspace& w = h->ws;
    const int ns = h->ctx.p.n_steps;
    dim3 grid((unsigned)((ns + FT_THREADS * FT_NODES - 1) / (FT_THREADS * FT_NODES)),
              (unsigned)std::max<int64_t>(1, std::min<int64_t>((n_upper + FT_STORMS - 1) / FT_STORMS, (int64_t)h->num_sms * 8)));
    const char* ft_env = getenv("TCR_FTAB_MMA");                         /* 0: scalar fp64-pipe variant (A/B runs, tests) */
    const int use_mma = ft_env ? atoi(ft_env) : 1;
    {
        LaunchTimer lt_(h, TCR_K_FTABLE);
        if (use_mma) {
            const int64_t tiles = std::max<int64_t>(1, (n_upper + FT_STORMS - 1) / FT_STORMS);
            const int g = (int)std::min<int64_t>(tiles, (int64_t)h->num_sms * use_mma);
            k_fourier_table_mma<<<g, FTM_THREADS, 0, h->stream>>>(h->ctx, n_upper, n_dev, w.coef.as<double2>(), w.ftab.as<double>());
        } else {
            k_fourier_table<<<grid, FT_THREADS, 0, h->stream>>>(h->ctx, n_upper, n_dev, w.coef.as<double2>(), w.ftab.as<double>());
        }
    }
    CKK(h);
    return 0;
}

}  // extern "C"

template <int THREADS, int MINB, int KSMEM, int LOCKSTEP = 0>
static void launch_integrate_variant(tcr_handle* h, IntegArgs& a, int64_t n_upper)
{
    const int warps_per_cta = THREADS / 32;
    int64_t max_ctas = (int64_t)h->num_sms * MINB;
    int64_t want_ctas = (n_upper + warps_per_cta - 1) / warps_per_cta;       /* one storm per warp at least */
    int grid = (int)std::max<int64_t>(1, std::min(max_ctas, want_ctas));
    int64_t lanes = (n_upper + (int64_t)grid * warps_per_cta - 1) / ((int64_t)grid * warps_per_cta);
    a.lane_cap = (int)std::max<int64_t>(1, std::min<int64_t>(32, lanes));
    { static const int pack_on = getenv("TCR_NO_PACK") ? 0 : 1; a.pack = pack_on; }
    a.pool_first = (unsigned int)grid * THREADS;                      /* lanes own rows [0, grid x THREADS) */
    /* KSMEM 2: eight stage vectors + the 18-word staging area of the drain-phase packing */
    const size_t smem = KSMEM == 2 ? (size_t)(32 + 18) * THREADS * sizeof(double) : KSMEM == 1 ? (size_t)20 * THREADS * sizeof(double) : 0;
    LaunchTimer lt_(h, TCR_K_INTEGRATE);
    cudaFuncSetAttribute(k_integrate<THREADS, MINB, KSMEM, LOCKSTEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_integrate<THREADS, MINB, KSMEM, LOCKSTEP><<<grid, THREADS, smem, h->stream>>>(h->ctx, a);
}

static int launch_integrate(tcr_handle* h, IntegArgs& a, int64_t n_upper)
{
    switch (h->integ_variant) {
    case 0: launch_integrate_variant<256, 1, 0>(h, a, n_upper); break;
    case 1: launch_integrate_variant<128, 3, 0>(h, a, n_upper); break;
    case 2: launch_integrate_variant<128, 4, 1>(h, a, n_upper); break;
    case 3: launch_integrate_variant<160, 2, 0>(h, a, n_upper); break;
    case 4: launch_integrate_variant<128, 3, 1>(h, a, n_upper); break;
    case 5: launch_integrate_variant<192, 2, 1>(h, a, n_upper); break;
    case 6: launch_integrate_variant<192, 2, 1, 63>(h, a, n_upper); break;
    case 7: launch_integrate_variant<256, 1, 0, 63>(h, a, n_upper); break;
    case 8: launch_integrate_variant<288, 1, 1, 63>(h, a, n_upper); break;
    case 9: launch_integrate_variant<384, 1, 1, 63>(h, a, n_upper); break;
    case 10: launch_integrate_variant<512, 1, 1, 63>(h, a, n_upper); break;
    case 11: launch_integrate_variant<128, 3, 1, 63>(h, a, n_upper); break;
    case 12: launch_integrate_variant<224, 1, 0, 63>(h, a, n_upper); break;
    case 13: launch_integrate_variant<192, 2, 1, 21>(h, a, n_upper); break;      /* re-align at slots 0, 2, 4 */
    case 14: launch_integrate_variant<192, 2, 1, 9>(h, a, n_upper); break;       /* slots 0, 3 */
    case 15: launch_integrate_variant<192, 2, 1, 1>(h, a, n_upper); break;       /* slot 0 only */
    case 16: launch_integrate_variant<256, 2, 2, 63>(h, a, n_upper); break;      /* all stage storage in smem: 128 registers */
    case 17: launch_integrate_variant<192, 2, 2, 63>(h, a, n_upper); break;
    case 18: launch_integrate_variant<192, 3, 2, 63>(h, a, n_upper); break;      /* 96 registers, 18 warps/SM */
    case 19: launch_integrate_variant<224, 2, 2, 63>(h, a, n_upper); break;      /* 144 registers, 14 warps/SM */
    case 20: launch_integrate_variant<384, 1, 2, 63>(h, a, n_upper); break;      /* one 12-warp CTA per SM */
    case 21: launch_integrate_variant<192, 2, 2, 63 + 256>(h, a, n_upper); break; /* variant 17 with the straight-line RHS (tcr_rhs_fast.cuh) */
    case 22: launch_integrate_variant<192, 2, 2, 3 + 256>(h, a, n_upper); break;  /* ... re-aligned at slots 0, 1 only */
    case 23: launch_integrate_variant<192, 2, 2, 0 + 256>(h, a, n_upper); break;  /* ... free-running warps (no drain packing) */
    case 24: launch_integrate_variant<192, 2, 2, 21 + 256>(h, a, n_upper); break; /* ... slots 0, 2, 4 */
    case 25: launch_integrate_variant<256, 2, 2, 63 + 256>(h, a, n_upper); break; /* 128 registers, 16 warps/SM */
    case 26: launch_integrate_variant<256, 2, 2, 0 + 256>(h, a, n_upper); break;
    case 27: launch_integrate_variant<192, 3, 2, 0 + 256>(h, a, n_upper); break;  /* 96 registers, 18 warps/SM */
    case 28: launch_integrate_variant<128, 3, 2, 63 + 256>(h, a, n_upper); break; /* three lock-step groups of 4 warps per SM */
    case 29: launch_integrate_variant<96, 4, 2, 63 + 256>(h, a, n_upper); break;  /* four groups of 3 warps */
    case 30: launch_integrate_variant<64, 6, 2, 63 + 256>(h, a, n_upper); break;  /* six groups of 2 warps */
    case 31: launch_integrate_variant<224, 2, 2, 63 + 256>(h, a, n_upper); break; /* 14 warps/SM */
    default: return set_err("unknown integrate variant %d", h->integ_variant);
    }
    CKK(h);
    return 0;
}

extern "C" {

/* ---- integrate given seeds ------------------------------------------------------------------ */
int tcr_integrate(tcr_handle* h, int64_t n, const int32_t* ym, const double* lon0, const double* lat0,
                  const double* v0, const double* m0, const double* h_bl, const double* phases,
                  double* track, double* env, double* vmax, int32_t* n_time, int32_t* status, int32_t* nfev,
                  uint32_t* flags, int on_device)
{
    if (!h) return set_err("null handle");
    if (n < 0) return set_err("tcr_integrate: negative n");
    if (n == 0) return 0;
    if (!ym || !lon0 || !lat0 || !v0 || !m0 || !h_bl || !phases) return set_err("tcr_integrate: null input");
    if (!h->rec.p || !h->have_static) return set_err("tcr_integrate: tables / static fields not uploaded");
    if (n > 0x7fffffff) return set_err("tcr_integrate: n too large");
    CK(cudaSetDevice(h->device));
    if (ws_ensure(h, 0, n, 1, WS_FULL)) return -1;
    Workspace& w = h->ws;
    const int ns = h->ctx.p.n_steps;
    cudaStream_t s = h->stream;
    const cudaMemcpyKind in_kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (!on_device)
        for (int64_t i = 0; i < n; ++i)
            if (ym[i] < 0 || ym[i] >= h->n_ym) return set_err("tcr_integrate: ym[%lld]=%d out of range", (long long)i, ym[i]);
    DevBuf ph;
    if (ph.ensure((size_t)n * TCR_N_PHASES * 8)) return -1;
    CK(cudaMemcpyAsync(w.s_ym.p, ym, n * 4, in_kind, s));
    CK(cudaMemcpyAsync(w.s_lon.p, lon0, n * 8, in_kind, s));
    CK(cudaMemcpyAsync(w.s_lat.p, lat0, n * 8, in_kind, s));
    CK(cudaMemcpyAsync(w.s_v0.p, v0, n * 8, in_kind, s));
    CK(cudaMemcpyAsync(w.s_m0.p, m0, n * 8, in_kind, s));
    CK(cudaMemcpyAsync(w.s_hbl.p, h_bl, n * 8, in_kind, s));
    CK(cudaMemcpyAsync(ph.p, phases, (size_t)n * TCR_N_PHASES * 8, in_kind, s));
    CK(cudaMemsetAsync(w.counters.p, 0, 64, s));
    CK(cudaMemsetAsync(w.track.p, 0xff, (size_t)n * ns * 32, s));
    CK(cudaMemsetAsync(w.env.p, 0xff, (size_t)n * ns * 32, s));
    CK(cudaMemsetAsync(w.vmax.p, 0xff, (size_t)n * ns * 8, s));
    {
        int64_t total = n * TCR_N_PHASES;
        LaunchTimer lt_(h, TCR_K_COEF);
        k_coef_from_phases<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(h->ctx, n, ph.as<double>(), w.coef.as<double2>());
        CKK(h);
    }
    IntegArgs a;
    memset(&a, 0, sizeof a);
    a.n = n; a.n_dev = nullptr;
    a.ym = w.s_ym.as<int32_t>(); a.lon0 = w.s_lon.as<double>(); a.lat0 = w.s_lat.as<double>();
    a.v0 = w.s_v0.as<double>(); a.m0 = w.s_m0.as<double>(); a.h_bl = w.s_hbl.as<double>();
    a.ftab = w.ftab.as<double>(); a.track = w.track.as<double>();
    a.n_time = w.n_time.as<int32_t>(); a.status = w.status.as<int32_t>(); a.nfev = w.nfev.as<int32_t>();
    a.flags = w.flags.as<uint32_t>();
    a.queue = w.counters.as<unsigned long long>();
    a.cand_list = nullptr; a.cand_count = nullptr;
    if (launch_fourier_table(h, n, nullptr) || launch_integrate(h, a, n)) { ph.release(); return -1; }
    PostArgs pa;
    memset(&pa, 0, sizeof pa);
    pa.n = n; pa.list = nullptr; pa.list_count = nullptr;
    pa.ym = a.ym; pa.ftab = a.ftab; pa.track = a.track; pa.n_time = a.n_time; pa.status = a.status;
    pa.track_row = nullptr; pa.pool_ctl = nullptr;
    pa.env = w.env.as<double>(); pa.vmax = w.vmax.as<double>(); pa.flags = a.flags;
    {
        LaunchTimer lt_(h, TCR_K_POSTPROCESS);
        k_postprocess<<<(unsigned)std::min<int64_t>(n, (int64_t)h->num_sms * 16), 128, 0, s>>>(h->ctx, pa);
    }
    CKK(h);
    const cudaMemcpyKind out_kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (track) CK(cudaMemcpyAsync(track, w.track.p, (size_t)n * ns * 32, out_kind, s));
    if (env) CK(cudaMemcpyAsync(env, w.env.p, (size_t)n * ns * 32, out_kind, s));
    if (vmax) CK(cudaMemcpyAsync(vmax, w.vmax.p, (size_t)n * ns * 8, out_kind, s));
    if (n_time) CK(cudaMemcpyAsync(n_time, w.n_time.p, n * 4, out_kind, s));
    if (status) CK(cudaMemcpyAsync(status, w.status.p, n * 4, out_kind, s));
    if (nfev) CK(cudaMemcpyAsync(nfev, w.nfev.p, n * 4, out_kind, s));
    if (flags) CK(cudaMemcpyAsync(flags, w.flags.p, n * 4, out_kind, s));
    CK(cudaStreamSynchronize(s));
    ph.release();
    return 0;
}

/* ---- single RHS evaluations (inner tier of the seam, test hook) --------------------------------- */
int tcr_rhs_eval(tcr_handle* h, int64_t n, const int32_t* ym, const double* t, const double* y, const double* h_bl,
                 const double* phases, double* dydt, double* env_winds)
{
    if (!h) return set_err("null handle");
    if (n < 0) return set_err("tcr_rhs_eval: negative n");
    if (n == 0) return 0;
    if (!ym || !t || !y || !h_bl || !phases || !dydt || !env_winds) return set_err("tcr_rhs_eval: null argument");
    if (!h->rec.p || !h->have_static) return set_err("tcr_rhs_eval: tables / static fields not uploaded");
    if (n > 0x7fffffff) return set_err("tcr_rhs_eval: n too large");
    for (int64_t i = 0; i < n; ++i)
        if (ym[i] < 0 || ym[i] >= h->n_ym) return set_err("tcr_rhs_eval: ym[%lld]=%d out of range", (long long)i, ym[i]);
    CK(cudaSetDevice(h->device));
    if (ws_ensure(h, 0, n, 1, WS_SLOTS)) return -1;
    Workspace& w = h->ws;
    cudaStream_t s = h->stream;
    DevBuf in, out;
    if (in.ensure((size_t)n * (TCR_N_PHASES + 6) * 8) || out.ensure((size_t)n * 64)) return -1;
    double* d_ph = in.as<double>();
    double* d_t = d_ph + n * TCR_N_PHASES;
    double* d_y = d_t + n;
    double* d_hbl = d_y + 4 * n;
    CK(cudaMemcpyAsync(d_ph, phases, (size_t)n * TCR_N_PHASES * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_t, t, n * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_y, y, n * 32, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_hbl, h_bl, n * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(w.s_ym.p, ym, n * 4, cudaMemcpyHostToDevice, s));
    {
        LaunchTimer lt_(h, TCR_K_COEF);
        k_coef_from_phases<<<(unsigned)((n * TCR_N_PHASES + 255) / 256), 256, 0, s>>>(h->ctx, n, d_ph, w.coef.as<double2>());
        CKK(h);
    }
    if (launch_fourier_table(h, n, nullptr)) return -1;
    k_rhs_eval<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(h->ctx, n, w.s_ym.as<int32_t>(), d_t, d_y, d_hbl, w.ftab.as<double>(),
                                                           out.as<double>(), out.as<double>() + 4 * n);
    CKK(h);
    CK(cudaMemcpyAsync(dydt, out.p, n * 32, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(env_winds, out.as<double>() + 4 * n, n * 32, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}

/* ---- seeding only (test hook) ---------------------------------------------------------------- */
int tcr_seed_attempts(tcr_handle* h, int ym_base, int32_t year_key, uint32_t run_seed, int64_t k0, int64_t n,
                      int32_t* code, int32_t* basin, int32_t* month, double* lon, double* lat, double* v0, double* m0,
                      double* pi_gen)
{
    if (!h) return set_err("null handle");
    if (n <= 0) return n < 0 ? set_err("negative n") : 0;
    if (!code || !basin || !month || !lon || !lat || !v0 || !m0 || !pi_gen) return set_err("tcr_seed_attempts: null output");
    if (!h->rec.p || !h->have_masks) return set_err("tcr_seed_attempts: tables / masks not uploaded");
    if (ym_base < 0 || ym_base + 12 > h->n_ym) return set_err("tcr_seed_attempts: ym_base out of range");
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    DevBuf ints, dbls, yr;
    if (ints.ensure((size_t)n * 12) || dbls.ensure((size_t)n * 40) || yr.ensure(64)) { ints.release(); dbls.release(); yr.release(); return -1; }
    int64_t hy[3] = {0, n, k0};
    int32_t hk[2] = {ym_base, year_key};
    CK(cudaMemcpyAsync(yr.p, hy, sizeof hy, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(yr.as<char>() + 32, hk, sizeof hk, cudaMemcpyHostToDevice, s));
    SeedArgs a;
    memset(&a, 0, sizeof a);
    a.n_years = 1;
    a.wave_off = yr.as<int64_t>(); a.k0 = yr.as<int64_t>() + 2;
    a.ym_base = reinterpret_cast<int32_t*>(yr.as<char>() + 32); a.year_key = a.ym_base + 1;
    a.run_seed = run_seed;
    a.code = ints.as<int32_t>(); a.basin = a.code + n; a.month = a.basin + n;
    a.lon = dbls.as<double>(); a.lat = a.lon + n; a.v0 = a.lat + n; a.m0 = a.v0 + n; a.pi_gen = a.m0 + n;
    a.blk_count = nullptr;
    k_seed<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(h->ctx, a);
    CKK(h);
    CK(cudaMemcpyAsync(code, a.code, n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(basin, a.basin, n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(month, a.month, n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(lon, a.lon, n * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(lat, a.lat, n * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(v0, a.v0, n * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(m0, a.m0, n * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(pi_gen, a.pi_gen, n * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    ints.release(); dbls.release(); yr.release();
    return 0;
}

/* ---- whole years ----------------------------------------------------------------------------- */
int tcr_run_years(tcr_handle* h, int n_years, const int32_t* ym_base, const int32_t* year_key, uint32_t run_seed,
                  int n_tracks, double* lon, double* lat, double* v, double* m, double* vmax, double* env,
                  double* tc_month, int32_t* tc_basin, double* n_seeds, tcr_year_stats* stats, int on_device)
{
    if (!h) return set_err("null handle");
    if (n_years <= 0 || n_tracks <= 0) return set_err("tcr_run_years: n_years and n_tracks must be positive");
    if (!ym_base || !year_key || !lon || !lat || !v || !m || !vmax || !env || !tc_month || !tc_basin || !n_seeds)
        return set_err("tcr_run_years: null argument");
    if (!h->rec.p || !h->have_static || !h->have_masks) return set_err("tcr_run_years: tables / static fields / masks not uploaded");
    for (int y = 0; y < n_years; ++y)
        if (ym_base[y] < 0 || ym_base[y] + 12 > h->n_ym) return set_err("tcr_run_years: ym_base[%d] out of range", y);
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const int ns = h->ctx.p.n_steps;
    const size_t rows = (size_t)n_years * n_tracks;

    /* wave capacity: attempts and slots -- what the job is expected to consume in its first wave,
     * bounded by a memory budget.  A workspace that already fits is reused without touching the
     * allocator (cudaMemGetInfo / cudaMalloc are millisecond-class host calls). */
    const size_t out_bytes = on_device ? 0 : rows * ns * 72 + rows * 12 + (size_t)n_years * 84 * 8;
    const int ring_nodes = ring_nodes_for(h);
    {
        Workspace& w0 = h->ws;
        const bool hinted = h->hint_kept_rate > 0.0;
        const double pass_est = hinted ? std::min(1.0, h->hint_pass_rate * 1.25 + 0.01) : 0.30;
        double need = (double)rows * 640.0;
        if (hinted) {
            need = 0.0;
            for (int y = 0; y < n_years; ++y) {
                const double rate = ((int)h->hint_year_rate.size() == n_years && h->hint_year_rate[y] > 0.0) ? h->hint_year_rate[y]
                                                                                                             : h->hint_kept_rate;
                need += ((double)n_tracks + 4.0 * std::sqrt((double)n_tracks) + 8.0) / rate * (h->oversub_permille / 1000.0);
            }
        }
        int64_t want_att = std::max<int64_t>(65536, (int64_t)(need * 1.05));
        if (h->max_wave > 0) want_att = std::min(want_att, h->max_wave);
        want_att = std::max<int64_t>(want_att, std::max<int64_t>(4096, 512 * (int64_t)n_years));   /* at least two 256-attempt blocks per year */
        int64_t want_slot = std::min(want_att, std::max<int64_t>(4096, (int64_t)((double)want_att * pass_est) + 1024));
        /* a workspace that was sized by the memory budget is as large as it gets: asking the allocator again would only
         * re-derive the same capacity from a slightly different free-memory reading, and a one-percent growth means
         * freeing and re-allocating tens of gigabytes (measured: 150-350 ms per call, every call, at configs[3]'s shape) */
        const bool fits = w0.ns == ns && w0.ring_nodes == ring_nodes && (w0.att_cap >= want_att || w0.mem_limited) && (w0.slot_cap >= want_slot || w0.mem_limited) &&
                          (on_device || w0.out.bytes >= out_bytes + 256);
        if (!fits) {
            size_t free_b = 0, total_b = 0;
            CK(cudaMemGetInfo(&free_b, &total_b));
            if (ring_nodes) w0.ftab.release(); else w0.ring.release();
            size_t held = w0.ftab.bytes + w0.ring.bytes + w0.track.bytes + w0.env.bytes + w0.vmax.bytes + w0.coef.bytes + w0.out.bytes;
            /* share of the free memory a wave's workspace may take, and its cap (tcr_set_workspace_budget; TCR_WS_FRAC,
             * TCR_WS_CAP_GB override both for A/B runs) */
            static const double env_frac = getenv("TCR_WS_FRAC") ? atof(getenv("TCR_WS_FRAC")) : 0.0;
            static const size_t env_cap = (size_t)(getenv("TCR_WS_CAP_GB") ? atoi(getenv("TCR_WS_CAP_GB")) : 0) << 30;
            const double ws_frac = env_frac > 0.0 ? env_frac : h->ws_frac;
            const size_t ws_cap = env_cap ? env_cap : h->ws_cap;
            size_t budget = std::min<size_t>(ws_cap, (size_t)((free_b + held) * ws_frac));
            if (budget > out_bytes) budget -= out_bytes;
            const int64_t cap_mem = (int64_t)((double)budget / ((double)kAttemptBytes + pass_est * (double)slot_bytes(ns, ring_nodes)));
            /* 25 % headroom so that the next call's slightly different estimate still fits */
            int64_t att_cap = std::max<int64_t>(std::max<int64_t>(4096, 512 * (int64_t)n_years), std::min(cap_mem, want_att + want_att / 4));
            w0.mem_limited = hinted && cap_mem < want_att + want_att / 4;     /* sized from measured survival rates: final */
            int64_t slot_cap = std::min(att_cap, std::max<int64_t>(4096, (int64_t)((double)att_cap * pass_est) + 1024));
            w0.env.release(); w0.vmax.release();                       /* only tcr_integrate keeps per-slot env / vmax rows */
            if (w0.full_cap > 0) { w0.track.release(); w0.full_cap = 0; w0.pool_rows = 0; }
            if (ws_ensure(h, std::max(att_cap, w0.att_cap), std::max(slot_cap, w0.slot_cap), n_years, WS_POOL, ring_nodes)) return -1;
        } else if (ws_ensure(h, w0.att_cap, w0.slot_cap, n_years, WS_POOL, ring_nodes)) {
            return -1;
        }
    }
    Workspace& w = h->ws;
    int64_t cap = w.att_cap;
    if (h->max_wave > 0) cap = std::min(cap, std::max<int64_t>(h->max_wave, 512 * (int64_t)n_years));
    int64_t slot_cap = w.slot_cap;
    if (h->max_slots > 0) slot_cap = std::min(slot_cap, h->max_slots);
    if (w.row_slot.ensure(rows * 4)) return -1;
    const int world = h->shard_world, srank = h->shard_rank;
    /* collective on the handle's stream through the caller's callback (NCCL in production); dtype 0 u8, 1 i32, 2 i64; op 0 sum, 1 min */
    auto allreduce = [&](void* d_buf, int64_t count, int dtype, int op) -> int {
        if (h->allreduce(h->allreduce_user, d_buf, count, dtype, op, (void*)s) != 0) return set_err("tcr_run_years: all-reduce callback failed");
        return 0;
    };
    if (world > 1) {
        /* every rank must cut the attempt stream into the same waves: agree on the smaller capacities, and on the
         * survival hints that size the waves (rank 0's) */
        int64_t* pin = reinterpret_cast<int64_t*>(h->pinned) + 4096;
        int64_t* d_tmp = w.year_i64.as<int64_t>();
        pin[0] = cap; pin[1] = slot_cap;
        CK(cudaMemcpyAsync(d_tmp, pin, 16, cudaMemcpyHostToDevice, s));
        if (allreduce(d_tmp, 2, 2, 1)) return -1;
        CK(cudaMemcpyAsync(pin, d_tmp, 16, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        cap = pin[0]; slot_cap = pin[1];
    }

    /* result block on the device */
    double *d_lon, *d_lat, *d_v, *d_m, *d_vmax, *d_env, *d_month, *d_seeds;
    int32_t* d_basin;
    if (on_device) {
        d_lon = lon; d_lat = lat; d_v = v; d_m = m; d_vmax = vmax; d_env = env; d_month = tc_month; d_basin = tc_basin; d_seeds = n_seeds;
    } else {
        if (w.out.ensure(out_bytes + 256)) return -1;
        double* base = w.out.as<double>();
        d_lon = base; d_lat = d_lon + rows * ns; d_v = d_lat + rows * ns; d_m = d_v + rows * ns; d_vmax = d_m + rows * ns;
        d_env = d_vmax + rows * ns; d_month = d_env + rows * ns * 4; d_seeds = d_month + rows;
        d_basin = reinterpret_cast<int32_t*>(d_seeds + (size_t)n_years * 84);
    }
    if (world > 1) {
        /* a row is written by the rank that integrated its storm and stays all-zero bits elsewhere: the ranks' blocks
         * then merge by an integer sum (exact for every bit pattern, NaN padding and signed zeros included) */
        CK(cudaMemsetAsync(d_lon, 0, rows * ns * 8, s)); CK(cudaMemsetAsync(d_lat, 0, rows * ns * 8, s));
        CK(cudaMemsetAsync(d_v, 0, rows * ns * 8, s)); CK(cudaMemsetAsync(d_m, 0, rows * ns * 8, s));
        CK(cudaMemsetAsync(d_vmax, 0, rows * ns * 8, s)); CK(cudaMemsetAsync(d_env, 0, rows * ns * 32, s));
    }
    CK(cudaMemsetAsync(d_month, 0xff, rows * 8, s));
    CK(cudaMemsetAsync(d_basin, 0xff, rows * 4, s));
    CK(cudaMemsetAsync(d_seeds, 0, (size_t)n_years * 84 * 8, s));
    CK(cudaMemsetAsync(w.stats.p, 0, (size_t)n_years * sizeof(tcr_year_stats), s));

    int64_t* d_wave_off = w.year_i64.as<int64_t>();
    int64_t* d_k0 = d_wave_off + n_years + 1;
    int64_t* d_consumed = d_k0 + n_years;
    int64_t* d_used = d_consumed + n_years;
    int32_t* d_ym_base = w.year_i32.as<int32_t>();
    int32_t* d_key = d_ym_base + n_years;
    int32_t* d_nt = d_key + n_years;
    CK(cudaMemcpyAsync(d_ym_base, ym_base, n_years * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_key, year_key, n_years * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(d_nt, 0, n_years * 4, s));

    std::vector<int64_t> k0(n_years, 0), W(n_years, 0), att_total(n_years, 0), hoff(2 * n_years + 1, 0);
    std::vector<int32_t> nt(n_years, 0);
    if ((size_t)n_years * 12 + 32 > (1 << 16)) return set_err("tcr_run_years: too many years in one call (%d)", n_years);
    int32_t* pin_nt = reinterpret_cast<int32_t*>(h->pinned);
    int64_t* pin_used = reinterpret_cast<int64_t*>(h->pinned) + (n_years + 1) / 2;

    unsigned long long* d_queue = w.counters.as<unsigned long long>();
    unsigned int* d_nslots = reinterpret_cast<unsigned int*>(d_queue + 1);
    unsigned int* d_ncand = d_nslots + 1;
    unsigned int* d_npass = d_nslots + 2;
    unsigned int* d_pool = d_nslots + 3;                       /* [0] fresh rows handed out, [1] overflow */

    std::vector<tcr_year_stats> hstats(n_years);
    bool final_done = false;
    auto final_copies = [&]() -> int {
        CK(cudaMemcpyAsync(hstats.data(), w.stats.p, (size_t)n_years * sizeof(tcr_year_stats), cudaMemcpyDeviceToHost, s));
        if (!on_device) {
            CK(cudaMemcpyAsync(lon, d_lon, rows * ns * 8, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(lat, d_lat, rows * ns * 8, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(v, d_v, rows * ns * 8, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(m, d_m, rows * ns * 8, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(vmax, d_vmax, rows * ns * 8, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(env, d_env, rows * ns * 32, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(tc_month, d_month, rows * 8, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(tc_basin, d_basin, rows * 4, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(n_seeds, d_seeds, (size_t)n_years * 84 * 8, cudaMemcpyDeviceToHost, s));
        }
        return 0;
    };

    const int max_waves = 4096;
    const double oversub = h->oversub_permille / 1000.0;
    int64_t sum_pass = 0, sum_att_launched = 0;
    int wave = 0;
    for (; wave < max_waves; ++wave) {
        int n_active = 0;
        for (int y = 0; y < n_years; ++y) if (nt[y] < n_tracks) ++n_active;
        if (!n_active) break;
        /* attempts wanted per year: (remaining tracks + a Poisson margin of ~3.5 sigma) / survival rate --
         * this year's own rate once it has produced tracks, else the hint of the previous call for the
         * same year slot, else the handle's average, else a probe -- times the over-subscription.
         * A mop-up wave costs a full kernel tail (~2 ms), so falling short must be rare. */
        double want_total = 0.0;
        std::vector<double> want(n_years, 0.0);
        for (int y = 0; y < n_years; ++y) {
            if (nt[y] >= n_tracks) continue;
            const double remaining = (double)(n_tracks - nt[y]);
            const double target = remaining + 3.5 * std::sqrt(remaining) + 4.0;
            double wy;
            if (nt[y] > 0) wy = target / ((double)nt[y] / (double)att_total[y]) * oversub;
            else if (att_total[y] > 0) wy = (double)att_total[y] * 4.0;
            else if ((int)h->hint_year_rate.size() == n_years && h->hint_year_rate[y] > 0.0) wy = target / h->hint_year_rate[y] * oversub;
            else if (h->hint_kept_rate > 0.0) wy = target / h->hint_kept_rate * oversub;
            else wy = std::max(2048.0, remaining * 32.0);
            want[y] = wy;
            want_total += wy;
        }
        /* every year's range is a whole number of 256-attempt blocks (the selection kernels work per block): the rounding
         * is paid for out of the capacity */
        const double cap_eff = (double)std::max<int64_t>(256, cap - 256 * (int64_t)n_years);
        const double scale = want_total > cap_eff ? cap_eff / want_total : 1.0;
        for (int y = 0; y < n_years; ++y)
            W[y] = want[y] > 0.0 ? ((std::max<int64_t>(64, (int64_t)(want[y] * scale)) + 255) & ~(int64_t)255) : 0;
        if (world > 1) {
            /* the ranks must issue identical attempt ranges: take the smallest proposal of each year (the proposals only
             * differ when the handles were tuned differently or remember different survival rates) */
            int64_t* pin = reinterpret_cast<int64_t*>(h->pinned) + 4096;
            memcpy(pin, W.data(), (size_t)n_years * 8);
            CK(cudaMemcpyAsync(d_used, pin, (size_t)n_years * 8, cudaMemcpyHostToDevice, s));
            if (allreduce(d_used, n_years, 2, 1)) return -1;
            CK(cudaMemcpyAsync(pin, d_used, (size_t)n_years * 8, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            memcpy(W.data(), pin, (size_t)n_years * 8);
        }
        int64_t total = 0;
        for (int y = 0; y < n_years; ++y) {
            hoff[y] = total;
            total += W[y];
            hoff[n_years + 1 + y] = k0[y];
        }
        hoff[n_years] = total;
        if (total > w.att_cap) return set_err("tcr_run_years: internal error: wave of %lld attempts exceeds the workspace", (long long)total);
        CK(cudaMemcpyAsync(d_wave_off, hoff.data(), (2 * n_years + 1) * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemsetAsync(w.counters.p, 0, 64, s));
        const unsigned seed_blocks = (unsigned)((total + 255) / 256);

        SeedArgs sa;
        memset(&sa, 0, sizeof sa);
        sa.n_years = n_years; sa.wave_off = d_wave_off; sa.k0 = d_k0; sa.ym_base = d_ym_base; sa.year_key = d_key;
        sa.run_seed = run_seed;
        sa.code = w.code.as<int32_t>(); sa.basin = w.basin.as<int32_t>(); sa.month = w.month.as<int32_t>();
        sa.lon = w.a_lon.as<double>(); sa.lat = w.a_lat.as<double>(); sa.v0 = w.a_v0.as<double>(); sa.m0 = w.a_m0.as<double>();
        sa.pi_gen = nullptr;
        sa.blk_count = w.blk_count.as<unsigned int>();
        sa.rank = srank; sa.world = world;
        AssignArgs as;
        memset(&as, 0, sizeof as);
        as.n_years = n_years; as.wave_off = d_wave_off; as.k0 = d_k0; as.ym_base = d_ym_base; as.year_key = d_key;
        as.code = sa.code; as.basin = sa.basin; as.month = sa.month;
        as.lon = sa.lon; as.lat = sa.lat; as.v0 = sa.v0; as.m0 = sa.m0;
        as.blk_off = w.blk_off.as<unsigned int>(); as.slot_cap = (unsigned int)std::min<int64_t>(slot_cap, 0x7fffffff);
        as.att_slot = w.att_slot.as<int32_t>(); as.consumed = d_consumed;
        as.s_ym = w.s_ym.as<int32_t>(); as.s_lon = w.s_lon.as<double>(); as.s_lat = w.s_lat.as<double>();
        as.s_v0 = w.s_v0.as<double>(); as.s_m0 = w.s_m0.as<double>(); as.s_hbl = w.s_hbl.as<double>();
        as.s_att = w.s_att.as<int64_t>(); as.s_key = w.s_key.as<int32_t>();
        as.rank = srank; as.world = world;
        {
            LaunchTimer lt_(h, TCR_K_SEED);
            k_seed<<<seed_blocks, 256, 0, s>>>(h->ctx, sa);
            k_scan_counts<<<1, 1024, 0, s>>>(sa.blk_count, w.blk_off.as<unsigned int>(), (int)seed_blocks, as.slot_cap,
                                             d_nslots, d_npass, d_wave_off, n_years, d_consumed);
            k_assign_slots<<<seed_blocks, 256, 0, s>>>(h->ctx, as);
        }
        CKK(h);
        h->launches += 2;
        const int64_t slots_upper = std::min<int64_t>(slot_cap, total);
        {
            LaunchTimer lt_(h, TCR_K_COEF);
            k_coef_from_philox<<<(unsigned)((slots_upper * (TCR_N_PHASES / 2) + 255) / 256), 256, 0, s>>>(
                h->ctx, d_nslots, as.s_att, as.s_key, run_seed, w.coef.as<double2>());
        }
        CKK(h);

        IntegArgs a;
        memset(&a, 0, sizeof a);
        a.n = 0; a.n_dev = d_nslots;
        a.ym = as.s_ym; a.lon0 = as.s_lon; a.lat0 = as.s_lat; a.v0 = as.s_v0; a.m0 = as.s_m0; a.h_bl = as.s_hbl;
        a.track = w.track.as<double>();
        if (ring_nodes) {
            /* the integrator tabulates what its storms reach, segment by segment, into the rings of the track-pool rows;
             * post-processing and the gather form the few nodes they need from the coefficients */
            a.ftab = nullptr; a.coef = w.coef.as<double2>(); a.ring = w.ring.as<double>(); a.ring_nodes = ring_nodes;
            { static const int cta_on = getenv("TCR_RING_CTA") ? atoi(getenv("TCR_RING_CTA")) : 1; a.ring_cta = cta_on; }
        } else {
            a.ftab = w.ftab.as<double>();
        }
        a.n_time = w.n_time.as<int32_t>(); a.status = w.status.as<int32_t>(); a.nfev = w.nfev.as<int32_t>();
        a.flags = w.flags.as<uint32_t>();
        a.queue = d_queue; a.cand_list = w.cand.as<int32_t>(); a.cand_count = d_ncand;
        a.track_row = w.track_row.as<int32_t>(); a.pool_ctl = d_pool; a.pool_rows = (unsigned int)w.pool_rows;
        if (!ring_nodes && launch_fourier_table(h, slots_upper, d_nslots)) return -1;
        if (launch_integrate(h, a, slots_upper)) return -1;

        PostArgs pa;
        memset(&pa, 0, sizeof pa);
        pa.n = 0; pa.list = w.cand.as<int32_t>(); pa.list_count = d_ncand;
        pa.ym = a.ym; pa.ftab = a.ftab; pa.coef = a.coef; pa.track = a.track; pa.n_time = a.n_time; pa.status = a.status;
        pa.track_row = a.track_row; pa.pool_ctl = d_pool;
        pa.env = nullptr; pa.vmax = nullptr; pa.flags = a.flags;
        {
            LaunchTimer lt_(h, TCR_K_POSTPROCESS);
            k_postprocess<<<h->num_sms * 8, 128, 0, s>>>(h->ctx, pa);
        }
        CKK(h);

        WaveStatsArgs wa;
        memset(&wa, 0, sizeof wa);
        wa.n_years = n_years; wa.wave_off = d_wave_off; wa.consumed = d_consumed;
        wa.code = sa.code; wa.basin = sa.basin; wa.month = sa.month; wa.att_slot = as.att_slot;
        wa.n_time = a.n_time; wa.nfev = a.nfev; wa.flags = a.flags;
        wa.att_kept = w.att_kept.as<uint8_t>();
        const int NG = TCR_N_BASINS * 12 + 2;
        wa.wave_loc = w.wave_tot.as<unsigned long long>();
        wa.wave_glob = reinterpret_cast<unsigned int*>(wa.wave_loc + 3 * (size_t)n_years);       /* [n_years][86] + overflow word */
        wa.k0 = d_k0; wa.rank = srank; wa.world = world;
        const size_t wave_tot_bytes = (size_t)n_years * (3 * 8 + NG * 4) + 16;
        const size_t sel_bytes = (size_t)n_years * (8 + 8 + 6 * 8 + TCR_N_BASINS * 12 * 4);
        CK(cudaMemsetAsync(w.wave_tot.p, 0, ((wave_tot_bytes + 15) & ~(size_t)15) + sel_bytes, s));
        if (world > 1) {
            /* a year whose range one rank had to cut (slot capacity) is cut for everybody */
            if (allreduce(d_consumed, n_years, 2, 1)) return -1;
        }

        SelectArgs se;
        memset(&se, 0, sizeof se);
        se.n_years = n_years; se.n_tracks = n_tracks; se.wave_off = d_wave_off; se.k0 = d_k0; se.consumed = d_consumed;
        se.code = sa.code; se.basin = sa.basin; se.month = sa.month; se.att_slot = as.att_slot;
        se.n_time = a.n_time; se.nfev = a.nfev;
        se.att_kept = wa.att_kept; se.wave_glob = wa.wave_glob; se.wave_loc = wa.wave_loc;
        se.nt = d_nt; se.used = d_used; se.row_slot = w.row_slot.as<int32_t>();
        se.tc_month = d_month; se.tc_basin = d_basin; se.n_seeds = d_seeds;
        se.stats = w.stats.as<tcr_year_stats>();
        se.pool_ctl = d_pool;
        {
            char* base = w.wave_tot.as<char>() + ((wave_tot_bytes + 15) & ~(size_t)15);
            se.istar = reinterpret_cast<int64_t*>(base);
            se.sel_acc = reinterpret_cast<unsigned long long*>(base + (size_t)n_years * 8);
            se.total_kept = reinterpret_cast<int32_t*>(base + (size_t)n_years * (8 + 6 * 8));
            se.sel_hist = reinterpret_cast<unsigned int*>(base + (size_t)n_years * (8 + 6 * 8 + 8));
            se.blk_kept = w.blk_count.as<unsigned int>();          /* the seeding scan is done with both by now */
            se.blk_pref = w.blk_off.as<unsigned int>();
        }
        {
            LaunchTimer lt_(h, TCR_K_SELECT);
            k_wave_stats<<<seed_blocks, 256, 0, s>>>(wa);
            if (world > 1) {
                /* the ranks' kept flags, counted-seed histograms and pool-overflow flags become global: after this every
                 * rank selects the same rows (the one survivor-count exchange per wave of SURVEY 8e) */
                unsigned int* d_ovf = wa.wave_glob + (size_t)n_years * NG;
                CK(cudaMemcpyAsync(d_ovf, d_pool + 1, 4, cudaMemcpyDeviceToDevice, s));
                if (allreduce(wa.att_kept, (total + 3) / 4 * 4, 0, 0) || allreduce(wa.wave_glob, (int64_t)n_years * NG + 1, 1, 0)) return -1;
                CK(cudaMemcpyAsync(d_pool + 1, d_ovf, 4, cudaMemcpyDeviceToDevice, s));
            }
            k_count_kept<<<seed_blocks, 256, 0, s>>>(se.att_kept, total, se.blk_kept);
            k_select_scan<<<n_years, 1024, 0, s>>>(se);
            k_select_rows<<<seed_blocks, 256, 0, s>>>(se);
            k_select_finish<<<n_years, 128, 0, s>>>(se);
        }
        CKK(h);
        h->launches += 4;

        GatherArgs ga;
        memset(&ga, 0, sizeof ga);
        ga.n_years = n_years; ga.n_tracks = n_tracks; ga.row_slot = se.row_slot; ga.n_time = a.n_time;
        ga.ym = a.ym; ga.ftab = a.ftab; ga.coef = a.coef; ga.track = a.track; ga.track_row = a.track_row; ga.pool_ctl = d_pool;
        ga.o_lon = d_lon; ga.o_lat = d_lat; ga.o_v = d_v; ga.o_m = d_m; ga.o_vmax = d_vmax; ga.o_env = d_env;
        {
            LaunchTimer lt_(h, TCR_K_GATHER);
            k_gather<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(h->ctx, ga);
        }
        CKK(h);

        CK(cudaMemcpyAsync(pin_nt, d_nt, n_years * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(pin_used, d_consumed, n_years * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(pin_used + n_years, d_npass, 12, cudaMemcpyDeviceToHost, s));       /* n_pass, pool next, pool overflow */
        /* a wave sized from survival hints almost always completes every year: queue the final
         * read-back behind it so that the call costs ONE host synchronisation, not two */
        const bool speculative = h->hint_kept_rate > 0.0 && scale == 1.0;
        if (speculative) { if (final_copies()) return -1; }
        CK(cudaStreamSynchronize(s));
        if (reinterpret_cast<unsigned int*>(pin_used + n_years)[2] != 0u) {
            /* more TC candidates than the track pool has rows: nothing of this wave was committed (k_select and
             * k_gather stood down); repeat it with half the slots -- results do not depend on wave size */
            slot_cap = std::max<int64_t>(1024, std::min<int64_t>(slot_cap, total) / 2);
            final_done = false;
            continue;
        }
        final_done = speculative;
        sum_pass += (int64_t)(*reinterpret_cast<unsigned int*>(pin_used + n_years));
        sum_att_launched += total;
        for (int y = 0; y < n_years; ++y) {
            if (W[y] > 0) {
                nt[y] = pin_nt[y];
                /* a year whose range was cut by the slot capacity only advances by what was processed */
                const int64_t done = std::min<int64_t>(W[y], pin_used[y]);
                att_total[y] += done;
                k0[y] += done;
            }
        }
    }
    bool complete = true;
    for (int y = 0; y < n_years; ++y) if (nt[y] < n_tracks) complete = false;
    if (!final_done || !complete) { if (final_copies()) return -1; }
    CK(cudaStreamSynchronize(s));
    if (stats) memcpy(stats, hstats.data(), (size_t)n_years * sizeof(tcr_year_stats));
    if (complete && sum_att_launched > 0) {
        /* kept tracks per consumed attempt (i*+1 of every year), integrated storms per launched attempt */
        int64_t kept = 0, consumed = 0;
        for (int y = 0; y < n_years; ++y) { kept += hstats[y].n_kept; consumed += hstats[y].attempts; }
        if (kept > 0 && consumed > 0) h->hint_kept_rate = (double)kept / (double)consumed;
        h->hint_year_rate.assign(n_years, 0.0);
        for (int y = 0; y < n_years; ++y)
            if (hstats[y].attempts > 0) h->hint_year_rate[y] = (double)hstats[y].n_kept / (double)hstats[y].attempts;
        h->hint_pass_rate = (double)sum_pass / (double)sum_att_launched;
    }
    if (!complete) return set_err("tcr_run_years: wave limit reached before every year produced %d tracks", n_tracks);
    return 0;
}

/* ---- return-period reduction ----------------------------------------------------- */
int tcr_poi_vmax(tcr_handle* h, int64_t n_rows, int n_steps, const double* lon, const double* lat, const double* vmax,
                 double poi_lon, double poi_lat, double radius_km, double r_earth_m, double* out, int on_device)
{
    if (!h) return set_err("null handle");
    if (n_rows < 0 || n_steps <= 0) return set_err("tcr_poi_vmax: bad shape");
    if (n_rows == 0) return 0;
    if (!lon || !lat || !vmax || !out) return set_err("tcr_poi_vmax: null argument");
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const size_t n = (size_t)n_rows * n_steps;
    const double *d_lon = lon, *d_lat = lat, *d_v = vmax;
    double* d_out = out;
    DevBuf in, o;
    if (!on_device) {
        if (in.ensure(n * 24) || o.ensure((size_t)n_rows * 8)) { in.release(); o.release(); return -1; }
        double* b = in.as<double>();
        CK(cudaMemcpyAsync(b, lon, n * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(b + n, lat, n * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(b + 2 * n, vmax, n * 8, cudaMemcpyHostToDevice, s));
        d_lon = b; d_lat = b + n; d_v = b + 2 * n; d_out = o.as<double>();
    }
    {
        LaunchTimer lt_(h, TCR_K_POI);
        const int grid = (int)std::min<int64_t>((n_rows + 7) / 8, (int64_t)h->num_sms * 16);
        k_poi_vmax<<<grid, 256, 0, s>>>(n_rows, n_steps, d_lon, d_lat, d_v, poi_lon, poi_lat, radius_km, r_earth_m / 1000.0, d_out);
    }
    CKK(h);
    if (!on_device) {
        CK(cudaMemcpyAsync(out, d_out, (size_t)n_rows * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        in.release(); o.release();
    }
    return 0;
}

int tcr_exceedance(tcr_handle* h, int64_t n, const double* v, int n_bins, const double* bins, int64_t* counts, int on_device)
{
    if (!h) return set_err("null handle");
    if (n < 0 || n_bins < 0 || n_bins > 64) return set_err("tcr_exceedance: bad shape (at most 64 bins)");
    if (!counts || (n_bins && !bins) || (n && !v)) return set_err("tcr_exceedance: null argument");
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    DevBuf buf;
    if (buf.ensure((size_t)n_bins * 16 + 64 + (on_device ? 0 : (size_t)n * 8))) return -1;
    double* d_bins = buf.as<double>();
    unsigned long long* d_cnt = reinterpret_cast<unsigned long long*>(d_bins + n_bins);
    const double* d_v = v;
    CK(cudaMemcpyAsync(d_bins, bins, (size_t)n_bins * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(d_cnt, 0, (size_t)n_bins * 8, s));
    if (!on_device && n) {
        double* dv = reinterpret_cast<double*>(d_cnt + n_bins);
        CK(cudaMemcpyAsync(dv, v, (size_t)n * 8, cudaMemcpyHostToDevice, s));
        d_v = dv;
    }
    if (n && n_bins) {
        LaunchTimer lt_(h, TCR_K_POI);
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)h->num_sms * 8));
        k_exceedance<<<grid, 256, 0, s>>>(n, d_v, n_bins, d_bins, d_cnt);
    }
    CKK(h);
    CK(cudaMemcpyAsync(counts, d_cnt, (size_t)n_bins * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    buf.release();
    return 0;
}

/* ---- per-month field preparation on the device ---------------------------------- */
int tcr_prepare_month(tcr_handle* h, int ym, const tcr_prep_spec* sp, const float* raw, const float* ocean,
                      const double* lon_g, const double* lat_g, const double* lon_o, const double* lat_o,
                      const int32_t* src_col, const int32_t* src_row, float* planes_out)
{
    if (!h || !sp || !raw || !ocean || !lon_g || !lat_g || !lon_o || !lat_o || !src_col || !src_row)
        return set_err("tcr_prepare_month: null argument");
    if (sp->nlat_g < 2 || sp->nlon_g < 2 || sp->nlat_o < 2 || sp->nlon_o < 2 || sp->nlat_b < 2 || sp->nlon_b < 2)
        return set_err("tcr_prepare_month: bad shape");
    if (ym >= 0) {
        if (!h->rec.p) return set_err("tcr_prepare_month: call tcr_alloc_tables first");
        if (ym >= h->n_ym) return set_err("tcr_prepare_month: ym %d out of range [0, %d)", ym, h->n_ym);
        if (sp->nlat_b != h->nlat || sp->nlon_b != h->nlon)
            return set_err("tcr_prepare_month: cropped grid %dx%d differs from the table grid %dx%d", sp->nlat_b, sp->nlon_b, h->nlat, h->nlon);
    }
    for (int j = 0; j < sp->nlon_b; ++j) if (src_col[j] < 0 || src_col[j] >= sp->nlon_g) return set_err("tcr_prepare_month: src_col[%d] out of range", j);
    for (int i = 0; i < sp->nlat_b; ++i) if (src_row[i] < 0 || src_row[i] >= sp->nlat_g) return set_err("tcr_prepare_month: src_row[%d] out of range", i);
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const size_t plane_g = (size_t)sp->nlat_g * sp->nlon_g, plane_o = (size_t)sp->nlat_o * sp->nlon_o;
    const size_t plane_b = (size_t)sp->nlat_b * sp->nlon_b;
    /* one scratch block: raw | ocean | axes | index maps ; the prepared planes go to the table stage buffer */
    const size_t b_raw = plane_g * 17 * 4, b_oc = plane_o * 2 * 4;
    const size_t b_ax = ((size_t)sp->nlon_g + sp->nlat_g + sp->nlon_o + sp->nlat_o) * 8;
    const size_t b_ix = ((size_t)sp->nlon_b + sp->nlat_b) * 4;
    if (h->prep.ensure(b_raw + b_oc + b_ax + b_ix + 64)) return -1;
    if (h->stage.ensure(plane_b * TCR_N_FIELDS * sizeof(float))) return -1;
    char* base = h->prep.as<char>();
    float* d_raw = reinterpret_cast<float*>(base);
    float* d_oc = reinterpret_cast<float*>(base + b_raw);
    double* d_ax = reinterpret_cast<double*>(base + b_raw + b_oc + ((8 - (b_raw + b_oc) % 8) % 8));
    int32_t* d_ix = reinterpret_cast<int32_t*>(d_ax + sp->nlon_g + sp->nlat_g + sp->nlon_o + sp->nlat_o);
    CK(cudaMemcpyAsync(d_raw, raw, b_raw, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_oc, ocean, b_oc, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_ax, lon_g, (size_t)sp->nlon_g * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_ax + sp->nlon_g, lat_g, (size_t)sp->nlat_g * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_ax + sp->nlon_g + sp->nlat_g, lon_o, (size_t)sp->nlon_o * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_ax + sp->nlon_g + sp->nlat_g + sp->nlon_o, lat_o, (size_t)sp->nlat_o * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_ix, src_col, (size_t)sp->nlon_b * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_ix + sp->nlon_b, src_row, (size_t)sp->nlat_b * 4, cudaMemcpyHostToDevice, s));
    PrepArgs a;
    memset(&a, 0, sizeof a);
    a.nlat_g = sp->nlat_g; a.nlon_g = sp->nlon_g; a.nlat_o = sp->nlat_o; a.nlon_o = sp->nlon_o; a.nlat_b = sp->nlat_b; a.nlon_b = sp->nlon_b;
    a.pi_reduc = sp->pi_reduc; a.sqrt_ck_cd = sp->sqrt_ck_cd; a.log_chi_fac = sp->log_chi_fac; a.chi_fac = sp->chi_fac;
    a.raw = d_raw; a.ocean = d_oc;
    a.lon_g = d_ax; a.lat_g = d_ax + sp->nlon_g; a.lon_o = a.lat_g + sp->nlat_g; a.lat_o = a.lon_o + sp->nlon_o;
    a.src_col = d_ix; a.src_row = d_ix + sp->nlon_b;
    a.out = h->stage.as<float>();
    {
        LaunchTimer lt_(h, TCR_K_BUILD);
        k_prepare_month<<<grid_for(plane_b * TCR_N_FIELDS, 256, h->num_sms), 256, 0, s>>>(a);
    }
    CKK(h);
    if (planes_out) CK(cudaMemcpyAsync(planes_out, h->stage.p, plane_b * TCR_N_FIELDS * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (ym >= 0) { if (tcr_upload_month_dev(h, ym, h->stage.as<float>())) return -1; }
    if (planes_out) CK(cudaStreamSynchronize(s));
    return 0;
}

/* ---- monthly wind mean / covariance reduction (track/env_wind.py:169-228) --------- */
}  // extern "C"

template <int VEC, int KSPLIT, int U>
static int launch_wind_stats(tcr_handle* h, const WindStatArgs& a)
{
    constexpr int P = 32 * VEC;
    const size_t smem = (size_t)(a.n_groups + 1) * 4 * P * sizeof(double);
    if (smem > h->smem_optin) return set_err("tcr_wind_stats: %d day groups need %zu B of shared memory (limit %zu)", a.n_groups, smem, h->smem_optin);
    CK(cudaFuncSetAttribute(k_wind_stats<VEC, KSPLIT, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = (a.n_pts + P - 1) / P;
    if (grid > 0x7fffffffLL) return set_err("tcr_wind_stats: too many grid points");
    LaunchTimer lt_(h, TCR_K_WINDSTAT);
    k_wind_stats<VEC, KSPLIT, U><<<(unsigned)grid, 128 * KSPLIT, smem, h->stream>>>(a);
    return 0;
}

template <int P, int NT>
static int launch_wind_stats_single(tcr_handle* h, const WindStatArgs& a)
{
    const size_t smem = (size_t)a.n_groups * 4 * P * sizeof(float) + (size_t)4 * P * sizeof(double);
    if (smem > h->smem_optin) return set_err("tcr_wind_stats: %d ungrouped samples need %zu B of shared memory (limit %zu): pass daily groups (group_start) for records this long", a.n_groups, smem, h->smem_optin);
    CK(cudaFuncSetAttribute(k_wind_stats_single<P, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t grid = (a.n_pts + P - 1) / P;
    if (grid > 0x7fffffffLL) return set_err("tcr_wind_stats: too many grid points");
    LaunchTimer lt_(h, TCR_K_WINDSTAT);
    k_wind_stats_single<P, NT><<<(unsigned)grid, NT, smem, h->stream>>>(a);
    return 0;
}

extern "C" {

int tcr_wind_stats(tcr_handle* h, int n_time, int64_t n_pts, int64_t t_stride,
                   const float* ua_upper, const float* va_upper, const float* ua_lower, const float* va_lower,
                   int n_groups, const int32_t* group_start, double* out, int on_device)
{
    if (!h) return set_err("null handle");
    if (n_time <= 0 || n_pts <= 0 || n_groups <= 0 || n_groups > n_time || t_stride < n_pts) return set_err("tcr_wind_stats: bad shape");
    if (!ua_upper || !va_upper || !ua_lower || !va_lower || !group_start || !out) return set_err("tcr_wind_stats: null argument");
    if (group_start[0] != 0 || group_start[n_groups] != n_time) return set_err("tcr_wind_stats: group_start must run from 0 to n_time");
    for (int g = 0; g < n_groups; ++g)
        if (group_start[g + 1] <= group_start[g]) return set_err("tcr_wind_stats: day group %d is empty or out of order", g);
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const float* hsrc[4] = {ua_upper, va_upper, ua_lower, va_lower};
    DevBuf in, o, gs;
    if (gs.ensure((size_t)(n_groups + 1) * 4)) return -1;
    CK(cudaMemcpyAsync(gs.p, group_start, (size_t)(n_groups + 1) * 4, cudaMemcpyHostToDevice, s));
    WindStatArgs a;
    memset(&a, 0, sizeof a);
    a.n_pts = n_pts; a.n_time = n_time; a.n_groups = n_groups; a.gstart = gs.as<int32_t>();
    if (on_device) {
        for (int v = 0; v < 4; ++v) a.src[v] = hsrc[v];
        a.t_stride = t_stride; a.out = out;
    } else {
        const size_t plane = (size_t)n_time * n_pts;
        if (in.ensure(plane * 4 * sizeof(float)) || o.ensure((size_t)n_pts * 14 * sizeof(double))) { in.release(); o.release(); gs.release(); return -1; }
        for (int v = 0; v < 4; ++v) {
            CK(cudaMemcpy2DAsync(in.as<float>() + plane * v, (size_t)n_pts * 4, hsrc[v], (size_t)t_stride * 4, (size_t)n_pts * 4, n_time, cudaMemcpyHostToDevice, s));
            a.src[v] = in.as<float>() + plane * v;
        }
        a.t_stride = n_pts; a.out = o.as<double>();
    }
    const char* ws_env = getenv("TCR_WS_VARIANT");                       /* tuning sweeps only */
    const int variant = ws_env ? atoi(ws_env) : -1;
    const bool single = n_groups == n_time;
    int rc;
    /* default kernels: 64 points per CTA when the month tile fits into shared memory, else 32 (long ungrouped records) */
    int pick = single ? 11 : 9;
    if (variant < 0 && single && ((size_t)n_groups * 4 * 64 + 8 * 64) * sizeof(float) > h->smem_optin) pick = 10;
    /* an ungrouped record whose month tile does not fit even at 32 points per CTA (an hourly month) streams from global memory */
    if (variant < 0 && single && ((size_t)n_groups * 4 * 32 + 8 * 32) * sizeof(float) > h->smem_optin) pick = 15;
    if (variant < 0 && !single && (size_t)(n_groups + 1) * 4 * 64 * sizeof(double) > h->smem_optin) pick = 8;
    switch (variant >= 0 ? variant : pick) {
    case 10: rc = launch_wind_stats_single<32, 64>(h, a); break;
    case 11: rc = launch_wind_stats_single<64, 128>(h, a); break;
    case 12: rc = launch_wind_stats_single<32, 128>(h, a); break;
    case 13: rc = launch_wind_stats_single<64, 256>(h, a); break;
    case 14: rc = launch_wind_stats_single<128, 256>(h, a); break;
    case 15: {
        if (!single) { rc = set_err("tcr_wind_stats: the streaming kernel is for ungrouped records"); break; }
        const int64_t grid = (a.n_pts + 255) / 256;
        if (grid > 0x7fffffffLL) { rc = set_err("tcr_wind_stats: too many grid points"); break; }
        LaunchTimer lt_(h, TCR_K_WINDSTAT);
        k_wind_stats_stream<<<(unsigned)grid, 256, 0, h->stream>>>(a);
        rc = 0;
        break;
    }
    case 0: rc = launch_wind_stats<1, 1, 8>(h, a); break;
    case 1: rc = launch_wind_stats<2, 1, 8>(h, a); break;
    default:
    case 2: rc = launch_wind_stats<2, 2, 8>(h, a); break;
    case 3: rc = launch_wind_stats<4, 2, 4>(h, a); break;
    case 4: rc = launch_wind_stats<1, 2, 8>(h, a); break;
    case 5: rc = launch_wind_stats<2, 2, 4>(h, a); break;
    case 6: rc = launch_wind_stats<4, 4, 4>(h, a); break;
    case 7: rc = launch_wind_stats<1, 1, 16>(h, a); break;
    case 8: rc = launch_wind_stats<1, 2, 4>(h, a); break;
    case 9: rc = launch_wind_stats<2, 4, 4>(h, a); break;
    }
    if (rc) { in.release(); o.release(); gs.release(); return rc; }
    CKK(h);
    if (!on_device) CK(cudaMemcpyAsync(out, a.out, (size_t)n_pts * 14 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    in.release(); o.release(); gs.release();
    return 0;
}

/* ---- potential intensity / saturation deficit / mid-level humidity (thermo/) ------ */
int tcr_set_entropy_table(tcr_handle* h, int np, int ns, const double* p_look, const double* s_look, const double* T_lookup)
{
    if (!h) return set_err("null handle");
    if (np < 2 || ns < 2 || !p_look || !s_look || !T_lookup) return set_err("tcr_set_entropy_table: bad argument");
    for (int i = 1; i < np; ++i) if (!(p_look[i] > p_look[i - 1])) return set_err("tcr_set_entropy_table: pressure axis must ascend");
    for (int i = 1; i < ns; ++i) if (!(s_look[i] > s_look[i - 1])) return set_err("tcr_set_entropy_table: entropy axis must ascend");
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)np + ns + (size_t)np * ns;
    if (h->etab.ensure(n * sizeof(double))) return -1;
    double* d = h->etab.as<double>();
    CK(cudaMemcpyAsync(d, p_look, (size_t)np * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d + np, s_look, (size_t)ns * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d + np + ns, T_lookup, (size_t)np * ns * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->etab_np = np; h->etab_ns = ns;
    return 0;
}

int tcr_set_entropy_table_reversible(tcr_handle* h, int np, int ns, int nrt, const double* p_look, const double* s_look,
                                     const double* rt_look, const double* T_lookup)
{
    if (!h) return set_err("null handle");
    if (np < 2 || ns < 2 || nrt < 2 || !p_look || !s_look || !rt_look || !T_lookup) return set_err("tcr_set_entropy_table_reversible: bad argument");
    for (int i = 1; i < np; ++i) if (!(p_look[i] > p_look[i - 1])) return set_err("tcr_set_entropy_table_reversible: pressure axis must ascend");
    for (int i = 1; i < ns; ++i) if (!(s_look[i] > s_look[i - 1])) return set_err("tcr_set_entropy_table_reversible: entropy axis must ascend");
    for (int i = 1; i < nrt; ++i) if (!(rt_look[i] > rt_look[i - 1])) return set_err("tcr_set_entropy_table_reversible: total-water axis must ascend");
    CK(cudaSetDevice(h->device));
    const size_t nT = (size_t)np * ns * nrt, n = (size_t)np + ns + nrt + nT;
    if (h->etab3.ensure(n * sizeof(double))) return -1;
    double* d = h->etab3.as<double>();
    CK(cudaMemcpyAsync(d, p_look, (size_t)np * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d + np, s_look, (size_t)ns * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d + np + ns, rt_look, (size_t)nrt * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(d + np + ns + nrt, T_lookup, nT * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->etab3_np = np; h->etab3_ns = ns; h->etab3_nr = nrt;
    return 0;
}

static int thermo_month_impl(tcr_handle* h, bool rev, int64_t n_pts, int nlev, const double* p_env, const float* ta, const float* hus,
                             const double* sst, const double* psl, double ck_over_cd, int k_mid,
                             double* vmax, double* chi, double* rh_mid, int on_device)
{
    if (!h) return set_err("null handle");
    if (!rev && !h->etab.p) return set_err("tcr_thermo_month: call tcr_set_entropy_table first");
    if (rev && !h->etab3.p) return set_err("tcr_thermo_month_reversible: call tcr_set_entropy_table_reversible first");
    if (n_pts <= 0 || nlev < 2 || nlev > TH_MAX_LEVELS || k_mid < 0 || k_mid >= nlev) return set_err("tcr_thermo_month: bad shape (2..%d levels)", TH_MAX_LEVELS);
    if (!p_env || !ta || !hus || !sst || !psl || !vmax || !chi || !rh_mid) return set_err("tcr_thermo_month: null argument");
    for (int k = 1; k < nlev; ++k)
        if (!(p_env[k] < p_env[k - 1])) return set_err("tcr_thermo_month: levels must run from the lowest model level (highest pressure) upwards");
    CK(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    DevBuf lv, in, o;
    const size_t b_lv = (size_t)nlev * (sizeof(ThLevel) + 8);
    if (lv.ensure(b_lv)) return -1;
    double* d_p = reinterpret_cast<double*>(lv.as<char>() + (size_t)nlev * sizeof(ThLevel));
    CK(cudaMemcpyAsync(d_p, p_env, (size_t)nlev * 8, cudaMemcpyHostToDevice, s));
    ThermoArgs a;
    memset(&a, 0, sizeof a);
    a.n_pts = n_pts; a.nlev = nlev; a.k_mid = k_mid; a.p_env = d_p; a.lev = lv.as<ThLevel>();
    if (rev) {
        a.np = h->etab3_np; a.ns = h->etab3_ns; a.nr = h->etab3_nr;
        a.p_look = h->etab3.as<double>(); a.s_look = a.p_look + a.np; a.r_look = a.s_look + a.ns; a.T_look = a.r_look + a.nr;
    } else {
        a.np = h->etab_np; a.ns = h->etab_ns;
        a.p_look = h->etab.as<double>(); a.s_look = a.p_look + a.np; a.T_look = a.s_look + a.ns;
    }
    a.cecd = ck_over_cd; a.p_mid = p_env[k_mid];
    if (on_device) {
        a.ta = ta; a.hus = hus; a.sst = sst; a.psl = psl; a.vmax = vmax; a.chi = chi; a.rh_mid = rh_mid;
    } else {
        const size_t col = (size_t)nlev * n_pts * sizeof(float);
        if (in.ensure(2 * col + (size_t)n_pts * 16) || o.ensure((size_t)n_pts * 24)) { lv.release(); in.release(); o.release(); return -1; }
        char* b = in.as<char>();
        CK(cudaMemcpyAsync(b + 2 * col, sst, (size_t)n_pts * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(b + 2 * col + (size_t)n_pts * 8, psl, (size_t)n_pts * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(b, ta, col, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(b + col, hus, col, cudaMemcpyHostToDevice, s));
        a.ta = reinterpret_cast<float*>(b); a.hus = reinterpret_cast<float*>(b + col);
        a.sst = reinterpret_cast<double*>(b + 2 * col); a.psl = a.sst + n_pts;
        a.vmax = o.as<double>(); a.chi = a.vmax + n_pts; a.rh_mid = a.chi + n_pts;
    }
    {
        LaunchTimer lt_(h, TCR_K_THERMO);
        k_thermo_levels<<<1, TH_MAX_LEVELS, 0, s>>>(a);
        if (rev) k_thermo<true><<<(unsigned)((n_pts + 127) / 128), 128, 0, s>>>(a);
        else k_thermo<false><<<(unsigned)((n_pts + 127) / 128), 128, 0, s>>>(a);
    }
    CKK(h);
    h->launches++;
    if (!on_device) {
        CK(cudaMemcpyAsync(vmax, a.vmax, (size_t)n_pts * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(chi, a.chi, (size_t)n_pts * 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(rh_mid, a.rh_mid, (size_t)n_pts * 8, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    lv.release(); in.release(); o.release();
    return 0;
}

int tcr_thermo_month(tcr_handle* h, int64_t n_pts, int nlev, const double* p_env, const float* ta, const float* hus,
                     const double* sst, const double* psl, double ck_over_cd, int k_mid,
                     double* vmax, double* chi, double* rh_mid, int on_device)
{
    return thermo_month_impl(h, false, n_pts, nlev, p_env, ta, hus, sst, psl, ck_over_cd, k_mid, vmax, chi, rh_mid, on_device);
}

int tcr_thermo_month_reversible(tcr_handle* h, int64_t n_pts, int nlev, const double* p_env, const float* ta, const float* hus,
                                const double* sst, const double* psl, double ck_over_cd, int k_mid,
                                double* vmax, double* chi, double* rh_mid, int on_device)
{
    return thermo_month_impl(h, true, n_pts, nlev, p_env, ta, hus, sst, psl, ck_over_cd, k_mid, vmax, chi, rh_mid, on_device);
}

}  // extern "C"


