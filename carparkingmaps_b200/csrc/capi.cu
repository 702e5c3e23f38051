// capi.cu -- the C ABI of libcpm.so (declared in include/cpm.h): contexts, device-resident city /
// fleet handles, and the host-pointer entry points that replace the reference's Julia functions.
// No CPU fallback: every computing entry point launches the sm_100a kernels or fails.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <chrono>
#include <functional>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/cpm.h"
#include "kernels.h"

using namespace cpm;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int32_t fail(int32_t code, const char *fmt, const char *a = "", const char *b = "")
{
    char buf[512];
    snprintf(buf, sizeof buf, fmt, a, b);
    g_err = buf;
    return code;
}
#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return fail(e__ == cudaErrorMemoryAllocation ? CPM_ERR_OOM : CPM_ERR_CUDA, "%s: %s", #call, \
                        cudaGetErrorString(e__));                                                    \
    } while (0)
#define TRY(call)                                \
    do {                                         \
        int32_t r__ = (call);                    \
        if (r__ != CPM_OK) return r__;           \
    } while (0)
#define ARG(cond)                                                             \
    do {                                                                      \
        if (!(cond)) return fail(CPM_ERR_ARG, "bad argument: %s", #cond);     \
    } while (0)

// ------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------
// Device blocks are recycled: a released block goes to a per-device free list and the next ensure() takes the smallest
// cached block that fits (and is not more than 2x too large) before it asks the driver.  A batch of cities of different
// sizes (BASELINE configs[2]) otherwise spends most of its time in cudaMalloc / cudaFree of GB-sized tables.  All work
// on a block is stream-ordered on the context's streams and every destroy synchronises first, so handing a block to its
// next owner needs no further synchronisation.  The cache is emptied when an allocation fails and at cpm_ctx_destroy.
struct BlockCache {
    std::map<int, std::multimap<size_t, void *>> free_blocks;   // per device ordinal
    std::mutex mu;   // contexts of different host threads may share a device
    void *take(int dev, size_t bytes, size_t *cap)
    {
        std::lock_guard<std::mutex> lock(mu);
        auto &m = free_blocks[dev];
        auto it = m.lower_bound(bytes);
        if (it == m.end() || it->first > 2 * bytes + (1u << 20)) return nullptr;
        void *p = it->second;
        *cap = it->first;
        m.erase(it);
        return p;
    }
    void give(int dev, void *p, size_t cap)
    {
        std::lock_guard<std::mutex> lock(mu);
        free_blocks[dev].emplace(cap, p);
    }
    void flush(int dev)
    {
        std::lock_guard<std::mutex> lock(mu);
        auto &m = free_blocks[dev];
        for (auto &kv : m) cudaFree(kv.second);
        m.clear();
    }
};
static BlockCache g_blocks;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int dev = 0;
    int32_t ensure(size_t bytes)
    {
        if (bytes <= cap) return CPM_OK;
        if (p) { cudaDeviceSynchronize(); release(); }   // growing: work in flight may still use the old block (cudaFree used to wait)
        cudaGetDevice(&dev);
        const size_t want = bytes ? bytes : 1;
        if ((p = g_blocks.take(dev, want, &cap)) != nullptr) return CPM_OK;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {   // give the cached blocks back to the driver and try once more
            cudaGetLastError();
            g_blocks.flush(dev);
            e = cudaMalloc(&p, want);
        }
        if (e != cudaSuccess) {
            p = nullptr; cap = 0;
            return fail(CPM_ERR_OOM, "cudaMalloc(%s bytes): %s", std::to_string(bytes).c_str(), cudaGetErrorString(e));
        }
        cap = want;
        return CPM_OK;
    }
    void release()
    {
        if (p) g_blocks.give(dev, p, cap);
        p = nullptr; cap = 0;
    }
    template <typename T> T *as() const { return static_cast<T *>(p); }
};

struct cpm_ctx {
    int device = 0;
    cudaStream_t st = nullptr;
    bool own_stream = false;
    int64_t launches = 0;
    // scratch objects of the host API (re-used across calls)
    cpm_city *h_city = nullptr;
    cpm_fleet *h_fleet = nullptr;
    // Host-array cache of the host API: which caller array (pointer, size, sampled checksum) each resident device copy
    // mirrors.  main.jl:82-98 passes the same datamatrix to createpdrive, createpdestin and resampling, the same p_dest to
    // solveinitialvalueproblem and resampling, and the trajectory arrays from one call to the next: a match skips the upload.
    struct HostKey { const void *p = nullptr; size_t n = 0; uint64_t sig = 0; bool valid = false; };
    HostKey hk_mean, hk_sd, hk_dist, hk_pdest, hk_state, hk_trans[4];
    bool cache_inputs = true;
    int64_t uploads_skipped = 0;
    int64_t graph_replays = 0;               // passes replayed as an executable graph (small cities)
    bool graphs = true;                      // cpm_ctx_set_graphs
    struct SweepState *sweep = nullptr;      // fleets and staging kept between cpm_sweep calls
    DevBuf traj_state, traj_trans, u[4], tmp, partials, scalars;
    // optional timing of the sampling kernel: one event pair per day (consecutive launches of the step kernel)
    bool profiling = false;
    uint64_t prof_tick = 0;
    cudaStream_t copy_st = nullptr;          // host-input pipeline: uploads overlap the tables and the IVP
    cudaStream_t aux_st = nullptr;           // dataset pass: createpdrive runs beside the CDF build
    cudaEvent_t ev_aux[2] = {nullptr, nullptr};   // fork (main -> aux), join (aux -> main)
    cudaEvent_t ev_copy[4] = {nullptr, nullptr, nullptr, nullptr};   // [0] first inputs, [1] all inputs, [2] compute -> copy, [3] distances
    std::vector<cudaEvent_t> ev_hour;        // host-input pipeline: the std slab of hour t has landed
    const cudaEvent_t *step_wait = nullptr;  // set around a resampling day: step s waits for step_wait[s / steps_per_hour]
    int64_t n_step_wait = 0;
    std::vector<cudaEvent_t> ev;      // start0, stop0, start1, stop1, ...
    std::vector<int64_t> ev_launches; // step-kernel launches bracketed by each pair
};

struct cpm_city {
    cpm_ctx *ctx = nullptr;
    int64_t N = 0, T = 0;
    DevBuf dm, dist, rdist, msd, mean_sum, p_drive, p_dest, norm, cdf, rowzero;
    bool have_msd = false;   // (mean, std) pair table of the current inputs
    // CDF row block: N doubles, padded to goff doubles, then the row's guide table (G+1 uint16, padded): ldc doubles
    // in total, a multiple of 16 bytes so that a row block is one TMA bulk copy for any N
    int64_t ldc = 0;
    int32_t goff = 0, G = 0;
    bool have_inputs = false, have_meansum = false, have_pdrive = false, have_cdf = false, have_pdest = false;
};

struct cpm_fleet {
    cpm_ctx *ctx = nullptr;
    int64_t C = 0;
    DevBuf zone, counts, scalars, norm_out;
    int64_t counts_N = 0, counts_steps = 0;   // shape of the last run's counts
    int64_t car_offset = 0;
    // two equivalent representations of the car state: zone[car] (gather sampler, host API) and cars
    // grouped by zone bin (bucket sampler); each is rebuilt lazily from the other
    bool zone_valid = true, bucket_valid = false;
    int64_t bucket_N = 0;
    BucketGeom geom = {};     // slab partition of the current grouping (first_zone depends on it)
    int cur = 0;
    // perm: car ids grouped by bin; dest: (destination, rank) per slot; pos: per-slab {count, first position} rows;
    // tot: two padded per-destination arrival counter arrays (ping-pong over the steps); first_zone: per slab
    DevBuf perm[2], dest, pos, tot, first_zone, bcnt, bdrv, cnt_cur, start_tmp, cursor_tmp, lost_acc;
    int32_t last_sampler = 0;   // CPM_SAMPLER_* the last day ran on
    DevBuf cars_cnt;   // car-resident day kernel: occupancy before the first step [N+1] | net arrivals [S][N+1] | drivers [S][N+1]
    // small cities: the launch sequence of a pass over this fleet, kept as an executable graph (graph_run)
    struct PassGraph { cudaGraphExec_t exec = nullptr; uint64_t key = 0; int seen = 0; bool off = false; } pg;
};

struct Scalars {   // device-side scalar block of a fleet
    double time_sum;
    unsigned long long lost;
    unsigned long long t_hi, t_lo;   // fixed-point driving-time sum: seconds, 2^-32 s
};

static int32_t use(cpm_ctx *ctx)
{
    ARG(ctx != nullptr);
    CU(cudaSetDevice(ctx->device));
    return CPM_OK;
}

// ------------------------------------------------------------------------------------------------
// plumbing
// ------------------------------------------------------------------------------------------------
extern "C" int32_t cpm_version(void) { return CPM_VERSION; }
extern "C" const char *cpm_last_error(void) { return g_err.c_str(); }
extern "C" void cpm_internal_set_error(const char *msg) { g_err = msg ? msg : ""; }   // for multi.cu; not part of cpm.h

extern "C" int32_t cpm_device_count(int32_t *count)
{
    ARG(count != nullptr);
    int n = 0;
    CU(cudaGetDeviceCount(&n));
    *count = n;
    return CPM_OK;
}

extern "C" int32_t cpm_ctx_create(int32_t device, void *stream, cpm_ctx **out)
{
    ARG(out != nullptr);
    *out = nullptr;
    int n = 0;
    CU(cudaGetDeviceCount(&n));
    if (n <= 0) return fail(CPM_ERR_CUDA, "no CUDA device visible (libcpm has no CPU fallback)");
    ARG(device >= 0 && device < n);
    CU(cudaSetDevice(device));
    cpm_ctx *c = new (std::nothrow) cpm_ctx();
    if (!c) return fail(CPM_ERR_OOM, "host allocation failed");
    c->device = device;
    if (stream) { c->st = (cudaStream_t)stream; c->own_stream = false; }
    else {
        cudaError_t e = cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete c; return fail(CPM_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
        c->own_stream = true;
    }
    *out = c;
    return CPM_OK;
}

struct SweepState;
static void sweep_state_free(SweepState *&w);

extern "C" int32_t cpm_ctx_destroy(cpm_ctx *ctx)
{
    if (!ctx) return CPM_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->st);
    sweep_state_free(ctx->sweep);
    if (ctx->h_city) cpm_city_destroy(ctx->h_city);
    if (ctx->h_fleet) cpm_fleet_destroy(ctx->h_fleet);
    for (cudaEvent_t e : ctx->ev) cudaEventDestroy(e);
    ctx->traj_state.release(); ctx->traj_trans.release(); ctx->tmp.release(); ctx->partials.release(); ctx->scalars.release();
    for (auto &b : ctx->u) b.release();
    if (ctx->aux_st) { cudaStreamSynchronize(ctx->aux_st); cudaStreamDestroy(ctx->aux_st); for (auto e : ctx->ev_aux) if (e) cudaEventDestroy(e); }
    if (ctx->copy_st) { cudaStreamSynchronize(ctx->copy_st); cudaStreamDestroy(ctx->copy_st); for (auto e : ctx->ev_copy) if (e) cudaEventDestroy(e); for (auto e : ctx->ev_hour) cudaEventDestroy(e); }
    if (ctx->own_stream) cudaStreamDestroy(ctx->st);
    g_blocks.flush(ctx->device);   // cached (unused) blocks go back to the driver
    delete ctx;
    return CPM_OK;
}

extern "C" int32_t cpm_ctx_sync(cpm_ctx *ctx)
{
    TRY(use(ctx));
    if (ctx->aux_st) CU(cudaStreamSynchronize(ctx->aux_st));
    CU(cudaStreamSynchronize(ctx->st));
    return CPM_OK;
}
extern "C" int32_t cpm_ctx_stream(cpm_ctx *ctx, void **stream_out)
{
    ARG(ctx && stream_out);
    *stream_out = (void *)ctx->st;
    return CPM_OK;
}
extern "C" int32_t cpm_ctx_launch_count(cpm_ctx *ctx, int64_t *count)
{
    ARG(ctx && count);
    *count = ctx->launches;
    return CPM_OK;
}
extern "C" int32_t cpm_host_alloc(int64_t bytes, void **out)
{
    ARG(out && bytes >= 0);
    CU(cudaHostAlloc(out, (size_t)(bytes ? bytes : 1), cudaHostAllocDefault));
    return CPM_OK;
}
extern "C" int32_t cpm_host_free(void *p)
{
    if (p) CU(cudaFreeHost(p));
    return CPM_OK;
}

// Page-locks an EXISTING host array (e.g. a Julia Array) for the duration of a batch of calls, so that the uploads of
// the host-pointer entry points run at full PCIe rate and overlap with the kernels; undo with cpm_host_unregister.
extern "C" int32_t cpm_host_register(void *p, int64_t bytes)
{
    ARG(p != nullptr && bytes > 0);
    CU(cudaHostRegister(p, (size_t)bytes, cudaHostRegisterDefault));
    return CPM_OK;
}
extern "C" int32_t cpm_host_unregister(void *p)
{
    ARG(p != nullptr);
    CU(cudaHostUnregister(p));
    return CPM_OK;
}

extern "C" int32_t cpm_ctx_profile_begin(cpm_ctx *ctx)
{
    TRY(use(ctx));
    for (cudaEvent_t e : ctx->ev) cudaEventDestroy(e);
    ctx->ev.clear(); ctx->ev_launches.clear();
    ctx->profiling = true;
    return CPM_OK;
}
extern "C" int32_t cpm_ctx_profile_end(cpm_ctx *ctx, double *step_kernel_ms, int64_t *step_launches)
{
    TRY(use(ctx));
    ctx->profiling = false;
    CU(cudaStreamSynchronize(ctx->st));
    double ms = 0.0; int64_t n = 0;
    for (size_t k = 0; k + 1 < ctx->ev.size(); k += 2) {
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, ctx->ev[k], ctx->ev[k + 1]));
        ms += t; n += ctx->ev_launches[k / 2];
    }
    for (cudaEvent_t e : ctx->ev) cudaEventDestroy(e);
    ctx->ev.clear(); ctx->ev_launches.clear();
    if (step_kernel_ms) *step_kernel_ms = ms;
    if (step_launches) *step_launches = n;
    return CPM_OK;
}

static int32_t check_launch(cpm_ctx *ctx, int n)
{
    if (n < 0) {   // a launcher's own CUDA call (attribute, memset, launch) failed: report that call's error
        const cudaError_t e = cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? CPM_ERR_OOM : CPM_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    }
    ctx->launches += n;
    CU(cudaGetLastError());
    return CPM_OK;
}
#define LAUNCH(ctx, expr) TRY(check_launch((ctx), (expr)))

// ------------------------------------------------------------------------------------------------
// city
// ------------------------------------------------------------------------------------------------
extern "C" int32_t cpm_city_create(cpm_ctx *ctx, int64_t N, int64_t T, cpm_city **out)
{
    TRY(use(ctx));
    ARG(out && N >= 1 && T >= 1 && N <= 0x7fffffff);
    cpm_city *c = new (std::nothrow) cpm_city();
    if (!c) return fail(CPM_ERR_OOM, "host allocation failed");
    c->ctx = ctx; c->N = N; c->T = T;
    *out = c;
    return CPM_OK;
}
extern "C" int32_t cpm_city_destroy(cpm_city *c)
{
    if (!c) return CPM_OK;
    cudaSetDevice(c->ctx->device);
    if (c->ctx->aux_st) cudaStreamSynchronize(c->ctx->aux_st);   // a failed pass may have left the createpdrive chain unjoined
    if (c->ctx->copy_st) cudaStreamSynchronize(c->ctx->copy_st); // an upload may still be writing into this city's blocks
    cudaStreamSynchronize(c->ctx->st);
    c->dm.release(); c->dist.release(); c->rdist.release(); c->msd.release(); c->mean_sum.release(); c->p_drive.release(); c->p_dest.release();
    c->norm.release(); c->cdf.release(); c->rowzero.release();
    delete c;
    return CPM_OK;
}

// Drops the tables (p_drive, p_dest, CDF, pair table, scratch) and keeps the inputs: the blocks go back to the device's
// free list, where the next city of a batch finds them (BASELINE configs[2]: 34 cities resident, one set of tables).
extern "C" int32_t cpm_city_release_tables(cpm_city *c)
{
    ARG(c != nullptr);
    TRY(use(c->ctx));
    if (c->ctx->aux_st) CU(cudaStreamSynchronize(c->ctx->aux_st));
    CU(cudaStreamSynchronize(c->ctx->st));
    c->rdist.release(); c->msd.release(); c->mean_sum.release(); c->p_drive.release(); c->p_dest.release();
    c->norm.release(); c->cdf.release(); c->rowzero.release();
    c->have_meansum = c->have_pdrive = c->have_cdf = c->have_pdest = c->have_msd = false;
    return CPM_OK;
}

static int32_t city_alloc_inputs(cpm_city *c)
{
    const size_t NN = (size_t)c->N * c->N;
    TRY(c->dm.ensure(NN * c->T * 2 * sizeof(double)));
    TRY(c->dist.ensure(NN * sizeof(double)));
    return CPM_OK;
}
static void city_invalidate(cpm_city *c) { c->have_meansum = c->have_pdrive = c->have_cdf = c->have_pdest = c->have_msd = false; }

extern "C" int32_t cpm_city_upload(cpm_city *c, const double *datamatrix, const double *distance_matrix_km)
{
    ARG(c && datamatrix && distance_matrix_km);
    TRY(use(c->ctx));
    TRY(city_alloc_inputs(c));
    const size_t NN = (size_t)c->N * c->N;
    CU(cudaMemcpyAsync(c->dm.p, datamatrix, NN * c->T * 2 * sizeof(double), cudaMemcpyHostToDevice, c->ctx->st));
    CU(cudaMemcpyAsync(c->dist.p, distance_matrix_km, NN * sizeof(double), cudaMemcpyHostToDevice, c->ctx->st));
    CU(cudaStreamSynchronize(c->ctx->st));   // the caller may reuse its host buffers
    c->have_inputs = true;
    city_invalidate(c);
    return CPM_OK;
}
extern "C" int32_t cpm_city_upload_rows(cpm_city *c, const int32_t *sourceid, const int32_t *dstid, const int32_t *hod,
                                        const double *mean, const double *stddev, int64_t n_rows,
                                        const double *distance_matrix_km, int64_t *n_skipped)
{
    ARG(c && distance_matrix_km && n_rows >= 0 && (n_rows == 0 || (sourceid && dstid && hod && mean && stddev)));
    cpm_ctx *ctx = c->ctx;
    TRY(use(ctx));
    TRY(city_alloc_inputs(c));
    const size_t nr = (size_t)n_rows;
    // staging: three int32 columns, two double columns, one counter
    TRY(ctx->tmp.ensure(nr * (3 * sizeof(int32_t) + 2 * sizeof(double)) + 64));
    double *d_mean = ctx->tmp.as<double>();
    double *d_sd = d_mean + nr;
    int32_t *d_src = reinterpret_cast<int32_t *>(d_sd + nr), *d_dst = d_src + nr, *d_hod = d_dst + nr;
    unsigned long long *d_bad = reinterpret_cast<unsigned long long *>(ctx->tmp.as<char>() + (nr * (3 * sizeof(int32_t) + 2 * sizeof(double)) + 7) / 8 * 8);
    if (nr) {
        CU(cudaMemcpyAsync(d_mean, mean, nr * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
        CU(cudaMemcpyAsync(d_sd, stddev, nr * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
        CU(cudaMemcpyAsync(d_src, sourceid, nr * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->st));
        CU(cudaMemcpyAsync(d_dst, dstid, nr * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->st));
        CU(cudaMemcpyAsync(d_hod, hod, nr * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->st));
    }
    CU(cudaMemcpyAsync(c->dist.p, distance_matrix_km, (size_t)c->N * c->N * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
    LAUNCH(ctx, launch_rows_scatter(d_src, d_dst, d_hod, d_mean, d_sd, n_rows, c->N, c->T, c->dm.as<double>(), d_bad, ctx->st));
    unsigned long long bad = 0;
    CU(cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    if (n_skipped) *n_skipped = (int64_t)bad;
    c->have_inputs = true;
    city_invalidate(c);
    return CPM_OK;
}

extern "C" int32_t cpm_city_synth(cpm_city *c, uint64_t seed, double density)
{
    ARG(c != nullptr);
    TRY(use(c->ctx));
    TRY(city_alloc_inputs(c));
    LAUNCH(c->ctx, launch_synth_city(c->dm.as<double>(), c->dist.as<double>(), c->N, c->T, seed, density, c->ctx->st));
    c->have_inputs = true;
    city_invalidate(c);
    return CPM_OK;
}
extern "C" int32_t cpm_city_download_inputs(cpm_city *c, double *datamatrix, double *distance_matrix_km)
{
    ARG(c != nullptr);
    TRY(use(c->ctx));
    if (!c->have_inputs) return fail(CPM_ERR_STATE, "city has no inputs");
    const size_t NN = (size_t)c->N * c->N;
    if (datamatrix) CU(cudaMemcpyAsync(datamatrix, c->dm.p, NN * c->T * 2 * sizeof(double), cudaMemcpyDeviceToHost, c->ctx->st));
    if (distance_matrix_km) CU(cudaMemcpyAsync(distance_matrix_km, c->dist.p, NN * sizeof(double), cudaMemcpyDeviceToHost, c->ctx->st));
    CU(cudaStreamSynchronize(c->ctx->st));
    return CPM_OK;
}

static int32_t city_build_pdrive_async(cpm_city *c, double p_min, double p_max, double e_drive, cudaStream_t st = nullptr)
{
    if (!st) st = c->ctx->st;
    if (!c->have_inputs) return fail(CPM_ERR_STATE, "createpdrive needs the city inputs (upload or synth first)");
    const size_t NT = (size_t)c->N * c->T;
    TRY(c->mean_sum.ensure(NT * sizeof(double)));
    TRY(c->rdist.ensure((size_t)c->N * c->N * 2 * sizeof(double)));   // (dist, 1/dist) pairs
    TRY(c->p_drive.ensure(NT * sizeof(double)));
    if (!c->have_meansum) {   // parameter-independent: cached across sweeps
        LAUNCH(c->ctx, launch_pdrive_meansum(c->dm.as<double>(), c->dist.as<double>(), c->rdist.as<double>(), c->N, c->T,
                                             c->mean_sum.as<double>(), st));
        c->have_meansum = true;
    }
    LAUNCH(c->ctx, launch_pdrive_finish(c->mean_sum.as<double>(), c->N, c->T, p_min, p_max, e_drive, c->p_drive.as<double>(), st));
    c->have_pdrive = true;
    return CPM_OK;
}
// layout of a CDF row block for N zones: N running sums, padding to goff doubles, then the guide table (G+1 uint16)
static void cdf_layout(int64_t N, int32_t *goff, int32_t *G, int64_t *ldc)
{
    *goff = (int32_t)((N + 1) / 2 * 2);                           // 16-byte aligned guide
    *G = 0;
    if (N + 1 <= 65535) {                                          // guide entries are uint16
        int64_t g = 64;
        // 4-8 CDF entries per guide cell.  Measured at cfg2 (profiles/README.md): G = 512 beats 1,024 (fewer 2-byte guide
        // stores in the CDF build, 1 KB less per staged row) and 256 / 128 / 64 (more probes per search)
        while (g < (N + 7) / 8 && g < 8192) g *= 2;
        *G = (int32_t)g;
    }
    const int64_t guide_doubles = *G ? ((int64_t)*G + 1 + 7) / 8 * 2 : 0;   // (G+1) uint16 rounded up to 16 bytes
    *ldc = *goff + guide_doubles;
}

// Diagnostics (host arithmetic only, no device needed): the launch geometry the bucket sampler would use for N zones and C
// cars.  out8 = {supported, n_slabs, slab, threads, ring, smem_sample, smem_scatter, ldc}.  Used by the CPU tests.
extern "C" int32_t cpm_debug_bucket_geometry(int64_t N, int64_t C, int32_t with_time, int64_t *out8)
{
    ARG(out8 && N >= 1 && C >= 1);
    int32_t goff, G;
    int64_t ldc;
    cdf_layout(N, &goff, &G, &ldc);
    BucketGeom g;
    memset(&g, 0, sizeof g);
    const bool ok = G > 0 && bucket_supported(N, ldc) && bucket_geometry(N, ldc, C, with_time, &g);
    const int64_t v[8] = { ok ? 1 : 0, g.n_slabs, g.slab, g.threads, g.ring, (int64_t)g.smem_sample, (int64_t)g.smem_scatter, ldc };
    for (int k = 0; k < 8; ++k) out8[k] = v[k];
    return CPM_OK;
}

// The same for the car-resident day kernel (csrc/daycars.cu): out5 = {supported, CTAs, threads, car slots per CTA,
// bytes of shared memory}.
extern "C" int32_t cpm_debug_cars_geometry(int64_t N, int64_t C, int32_t with_time, int64_t *out5)
{
    ARG(out5 && N >= 1 && C >= 1);
    CarsGeom g;
    memset(&g, 0, sizeof g);
    const bool ok = cars_geometry(N, C, with_time, 1, &g);
    const int64_t v[5] = { ok ? 1 : 0, g.n_ctas, g.threads, g.cars_per_cta, (int64_t)g.smem };
    for (int k = 0; k < 5; ++k) out5[k] = v[k];
    return CPM_OK;
}

static int32_t city_alloc_cdf(cpm_city *c)
{
    cdf_layout(c->N, &c->goff, &c->G, &c->ldc);
    TRY(c->cdf.ensure((size_t)c->N * c->ldc * c->T * sizeof(double)));
    TRY(c->rowzero.ensure((size_t)c->N * c->T));
    return CPM_OK;
}
static int32_t city_build_pdestin_async(cpm_city *c, double e_dest, bool write_p = true, cudaEvent_t before_cdf = nullptr)
{
    if (!c->have_inputs) return fail(CPM_ERR_STATE, "createpdestin needs the city inputs (upload or synth first)");
    const size_t NN = (size_t)c->N * c->N;
    TRY(c->p_dest.ensure(NN * c->T * sizeof(double)));
    TRY(c->norm.ensure((size_t)c->N * c->T * sizeof(double)));
    TRY(city_alloc_cdf(c));
    LAUNCH(c->ctx, launch_pdestin(c->dm.as<double>(), c->N, c->T, e_dest, c->p_dest.as<double>(), c->norm.as<double>(), c->ctx->st));
    if (before_cdf) CU(cudaEventRecord(before_cdf, c->ctx->st));   // the two streaming passes are done: HBM has headroom from here
    LAUNCH(c->ctx, launch_cdf_normalise(c->p_dest.as<double>(), c->norm.as<double>(), c->N, c->T, c->cdf.as<double>(),
                                        c->ldc, c->goff, c->G, c->rowzero.as<uint8_t>(), write_p ? 1 : 0, c->ctx->st));
    c->have_cdf = true;
    c->have_pdest = write_p;   // without write_p the buffer holds the unnormalised stage only
    return CPM_OK;
}

extern "C" int32_t cpm_city_prepare_multi_day(cpm_city *c)
{
    ARG(c != nullptr);
    TRY(use(c->ctx));
    if (!c->have_inputs) return fail(CPM_ERR_STATE, "the city has no inputs");
    if (c->have_msd) return CPM_OK;
    TRY(c->msd.ensure((size_t)c->N * c->N * c->T * 2 * sizeof(double)));
    LAUNCH(c->ctx, launch_pair_table(c->dm.as<double>(), c->N, c->T, c->msd.as<double>(), c->ctx->st));
    c->have_msd = true;
    return CPM_OK;
}

extern "C" int32_t cpm_city_build_pdrive(cpm_city *c, double p_min, double p_max, double e_drive)
{
    ARG(c != nullptr);
    TRY(use(c->ctx));
    return city_build_pdrive_async(c, p_min, p_max, e_drive);
}
extern "C" int32_t cpm_city_build_pdestin(cpm_city *c, double e_dest, int32_t keep_p_dest)
{
    ARG(c != nullptr);
    TRY(use(c->ctx));
    TRY(city_build_pdestin_async(c, e_dest, keep_p_dest != 0));
    if (!keep_p_dest) {   // p_dest is only a stage of the CDF build; drop it when the caller does not need it
        CU(cudaStreamSynchronize(c->ctx->st));
        c->p_dest.release();
        c->have_pdest = false;
    }
    return CPM_OK;
}
extern "C" int32_t cpm_city_set_pdrive(cpm_city *c, const double *p_drive)
{
    ARG(c && p_drive);
    TRY(use(c->ctx));
    const size_t NT = (size_t)c->N * c->T;
    TRY(c->p_drive.ensure(NT * sizeof(double)));
    CU(cudaMemcpyAsync(c->p_drive.p, p_drive, NT * sizeof(double), cudaMemcpyHostToDevice, c->ctx->st));
    CU(cudaStreamSynchronize(c->ctx->st));
    c->have_pdrive = true;
    return CPM_OK;
}
extern "C" int32_t cpm_city_set_pdest(cpm_city *c, const double *p_dest)
{
    ARG(c && p_dest);
    TRY(use(c->ctx));
    const size_t NN = (size_t)c->N * c->N;
    TRY(c->p_dest.ensure(NN * c->T * sizeof(double)));
    TRY(city_alloc_cdf(c));
    CU(cudaMemcpyAsync(c->p_dest.p, p_dest, NN * c->T * sizeof(double), cudaMemcpyHostToDevice, c->ctx->st));
    LAUNCH(c->ctx, launch_cdf_from_pdest(c->p_dest.as<double>(), c->N, c->T, c->cdf.as<double>(), c->ldc, c->goff, c->G, c->rowzero.as<uint8_t>(), c->ctx->st));
    CU(cudaStreamSynchronize(c->ctx->st));
    c->have_cdf = true;
    c->have_pdest = true;
    return CPM_OK;
}
extern "C" int32_t cpm_city_get_pdrive(cpm_city *c, double *out)
{
    ARG(c && out);
    TRY(use(c->ctx));
    if (!c->have_pdrive) return fail(CPM_ERR_STATE, "p_drive has not been built");
    CU(cudaMemcpyAsync(out, c->p_drive.p, (size_t)c->N * c->T * sizeof(double), cudaMemcpyDeviceToHost, c->ctx->st));
    CU(cudaStreamSynchronize(c->ctx->st));
    return CPM_OK;
}
extern "C" int32_t cpm_city_get_pdest(cpm_city *c, double *out)
{
    ARG(c && out);
    TRY(use(c->ctx));
    if (!c->have_pdest) return fail(CPM_ERR_STATE, "p_dest is not resident (build with keep_p_dest=1)");
    CU(cudaMemcpyAsync(out, c->p_dest.p, (size_t)c->N * c->N * c->T * sizeof(double), cudaMemcpyDeviceToHost, c->ctx->st));
    CU(cudaStreamSynchronize(c->ctx->st));
    return CPM_OK;
}
extern "C" int32_t cpm_city_get_cdf(cpm_city *c, double *out)
{
    ARG(c && out);
    TRY(use(c->ctx));
    if (!c->have_cdf) return fail(CPM_ERR_STATE, "the CDF has not been built");
    // device rows are padded to ldc doubles; the caller gets dense [T][N][N]
    CU(cudaMemcpy2DAsync(out, (size_t)c->N * sizeof(double), c->cdf.p, (size_t)c->ldc * sizeof(double),
                         (size_t)c->N * sizeof(double), (size_t)c->N * c->T, cudaMemcpyDeviceToHost, c->ctx->st));
    CU(cudaStreamSynchronize(c->ctx->st));
    return CPM_OK;
}

extern "C" int32_t cpm_city_get_guide(cpm_city *c, int64_t hour, uint16_t *guide_out, int32_t *G_out)
{
    ARG(c && G_out && hour >= 0 && hour < c->T);
    TRY(use(c->ctx));
    if (!c->have_cdf) return fail(CPM_ERR_STATE, "the CDF has not been built");
    *G_out = c->G;
    if (c->G == 0 || !guide_out) return CPM_OK;
    // the guide of a row sits behind its N running sums (offset goff doubles inside the row block of ldc doubles)
    const char *src = c->cdf.as<char>() + ((size_t)hour * c->N * c->ldc + (size_t)c->goff) * sizeof(double);
    CU(cudaMemcpy2DAsync(guide_out, (size_t)(c->G + 1) * sizeof(uint16_t), src, (size_t)c->ldc * sizeof(double),
                         (size_t)(c->G + 1) * sizeof(uint16_t), (size_t)c->N, cudaMemcpyDeviceToHost, c->ctx->st));
    CU(cudaStreamSynchronize(c->ctx->st));
    return CPM_OK;
}

// ------------------------------------------------------------------------------------------------
// fleet
// ------------------------------------------------------------------------------------------------
extern "C" int32_t cpm_fleet_create(cpm_ctx *ctx, int64_t C, cpm_fleet **out)
{
    TRY(use(ctx));
    ARG(out && C >= 1 && C <= 0x7fffffff);
    cpm_fleet *f = new (std::nothrow) cpm_fleet();
    if (!f) return fail(CPM_ERR_OOM, "host allocation failed");
    f->ctx = ctx; f->C = C;
    int32_t r = f->zone.ensure((size_t)C * sizeof(int32_t));
    if (r == CPM_OK) r = f->scalars.ensure(sizeof(Scalars));
    if (r != CPM_OK) { f->zone.release(); f->scalars.release(); delete f; return r; }
    cudaMemsetAsync(f->zone.p, 0, (size_t)C * sizeof(int32_t), ctx->st);
    *out = f;
    return CPM_OK;
}
extern "C" int32_t cpm_fleet_destroy(cpm_fleet *f)
{
    if (!f) return CPM_OK;
    cudaSetDevice(f->ctx->device);
    if (f->ctx->aux_st) cudaStreamSynchronize(f->ctx->aux_st);
    if (f->ctx->copy_st) cudaStreamSynchronize(f->ctx->copy_st);
    cudaStreamSynchronize(f->ctx->st);
    DevBuf *bufs[] = { &f->zone, &f->counts, &f->scalars, &f->norm_out, &f->perm[0], &f->perm[1], &f->dest, &f->pos, &f->bcnt,
                       &f->bdrv, &f->cnt_cur, &f->start_tmp, &f->cursor_tmp, &f->tot, &f->first_zone, &f->lost_acc, &f->cars_cnt };
    for (DevBuf *b : bufs) b->release();
    if (f->pg.exec) cudaGraphExecDestroy(f->pg.exec);
    delete f;
    return CPM_OK;
}
extern "C" int32_t cpm_fleet_init(cpm_fleet *f, int64_t cars_per_zone, int64_t car_offset)
{
    ARG(f && cars_per_zone >= 1 && car_offset >= 0);
    TRY(use(f->ctx));
    f->car_offset = car_offset;
    LAUNCH(f->ctx, launch_fleet_init(f->zone.as<int32_t>(), f->C, cars_per_zone, car_offset, f->ctx->st));
    f->zone_valid = true; f->bucket_valid = false;
    return CPM_OK;
}
extern "C" int32_t cpm_fleet_set_zones(cpm_fleet *f, const int64_t *zones)
{
    ARG(f && zones);
    TRY(use(f->ctx));
    cpm_ctx *ctx = f->ctx;
    TRY(ctx->tmp.ensure((size_t)f->C * sizeof(int64_t)));
    CU(cudaMemcpyAsync(ctx->tmp.p, zones, (size_t)f->C * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->st));
    LAUNCH(ctx, launch_i64_to_i32(ctx->tmp.as<int64_t>(), f->zone.as<int32_t>(), f->C, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    f->zone_valid = true; f->bucket_valid = false;
    return CPM_OK;
}
extern "C" int32_t cpm_fleet_get_zones(cpm_fleet *f, int64_t *zones_out)
{
    ARG(f && zones_out);
    TRY(use(f->ctx));
    cpm_ctx *ctx = f->ctx;
    if (!f->zone_valid) return fail(CPM_ERR_STATE, "fleet zones are not materialised");
    TRY(ctx->tmp.ensure((size_t)f->C * sizeof(int64_t)));
    LAUNCH(ctx, launch_i32_to_i64(f->zone.as<int32_t>(), ctx->tmp.as<int64_t>(), f->C, ctx->st));
    CU(cudaMemcpyAsync(zones_out, ctx->tmp.p, (size_t)f->C * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return CPM_OK;
}
extern "C" int32_t cpm_fleet_copy_zones(cpm_fleet *dst, const cpm_fleet *src)
{
    ARG(dst && src && dst->C == src->C && dst->ctx == src->ctx);
    TRY(use(dst->ctx));
    if (!src->zone_valid) return fail(CPM_ERR_STATE, "source fleet zones are not materialised");
    CU(cudaMemcpyAsync(dst->zone.p, src->zone.p, (size_t)dst->C * sizeof(int32_t), cudaMemcpyDeviceToDevice, dst->ctx->st));
    dst->zone_valid = true; dst->bucket_valid = false;
    if (src->bucket_valid) {   // the source is grouped by zone already: take its grouping instead of redoing it per copy
        const size_t NB = (size_t)src->bucket_N + 1;
        TRY(dst->perm[0].ensure((size_t)dst->C * 4));
        TRY(dst->cnt_cur.ensure(NB * 4)); TRY(dst->start_tmp.ensure((NB + 1) * 4));
        TRY(dst->first_zone.ensure((size_t)src->geom.n_slabs * 4));
        CU(cudaMemcpyAsync(dst->perm[0].p, src->perm[src->cur].p, (size_t)dst->C * 4, cudaMemcpyDeviceToDevice, dst->ctx->st));
        CU(cudaMemcpyAsync(dst->cnt_cur.p, src->cnt_cur.p, NB * 4, cudaMemcpyDeviceToDevice, dst->ctx->st));
        CU(cudaMemcpyAsync(dst->start_tmp.p, src->start_tmp.p, (NB + 1) * 4, cudaMemcpyDeviceToDevice, dst->ctx->st));
        CU(cudaMemcpyAsync(dst->first_zone.p, src->first_zone.p, (size_t)src->geom.n_slabs * 4, cudaMemcpyDeviceToDevice, dst->ctx->st));
        dst->cur = 0; dst->bucket_valid = true; dst->bucket_N = src->bucket_N; dst->geom = src->geom;
    }
    return CPM_OK;
}

// ------------------------------------------------------------------------------------------------
// the day loop
// ------------------------------------------------------------------------------------------------
struct Traj {
    int64_t *state = nullptr;   // Julia layout [C,T]
    double *trans = nullptr;    // Julia layout [C,T,4]
    int64_t T = 0;
};

static RngDev make_rng(const cpm_rng &r, int64_t C)
{
    RngDev d;
    d.mode = r.mode;
    d.stream = r.stream;
    d.k0 = (uint32_t)(r.seed & 0xffffffffu);
    d.k1 = (uint32_t)(r.seed >> 32);
    d.car_offset = r.car_offset;
    d.ld = C;
    d.u_drive = r.u_drive; d.u_dest = r.u_dest; d.u_time = r.u_time; d.u_dist = r.u_dist;
    for (uint32_t i = 0; i < 10; ++i) { d.rk0[i] = d.k0 + i * 0x9E3779B9u; d.rk1[i] = d.k1 + i * 0xBB67AE85u; }
    return d;
}

static int32_t profile_open(cpm_ctx *ctx, int64_t launches)
{
    if (!ctx->profiling) return CPM_OK;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    ctx->ev.push_back(e0); ctx->ev.push_back(e1); ctx->ev_launches.push_back(launches);
    CU(cudaEventRecord(e0, ctx->st));
    return CPM_OK;
}
static int32_t profile_close(cpm_ctx *ctx)
{
    if (ctx->profiling) CU(cudaEventRecord(ctx->ev.back(), ctx->st));
    return CPM_OK;
}

// gather sampler: one thread per car and step, car-id order (last fallback; records trajectories too)
static int32_t day_gather(cpm_city *city, cpm_fleet *f, const cpm_day_opts *o, const Traj *traj)
{
    cpm_ctx *ctx = f->ctx;
    const int64_t N = city->N;
    const size_t n_counts = (size_t)N * o->n_steps;
    if (!f->zone_valid) return fail(CPM_ERR_STATE, "fleet zones are not materialised");
    CU(cudaMemsetAsync(f->counts.p, 0, 2 * n_counts * sizeof(uint32_t), ctx->st));
    StepArgs a;
    memset(&a, 0, sizeof a);
    a.p_drive = city->p_drive.as<double>();
    a.cdf = city->cdf.as<double>();
    a.ldc = city->ldc; a.goff = city->goff; a.G = city->G;
    a.rowzero = city->rowzero.as<uint8_t>();
    a.dm = city->dm.as<double>();
    a.msd = city->have_msd ? city->msd.as<double2>() : nullptr;
    a.dist = city->dist.as<double>();
    a.N = N; a.T_tab = city->T;
    a.zone = f->zone.as<int32_t>();
    a.C = f->C;
    a.with_time = o->with_time;
    a.rng = make_rng(o->rng, f->C);
    Scalars *sc = f->scalars.as<Scalars>();
    a.t_hi = &sc->t_hi; a.t_lo = &sc->t_lo; a.lost = &sc->lost;
    if (traj) { a.traj_state = traj->state; a.traj_trans = traj->trans; a.T_traj = traj->T; }
    TRY(profile_open(ctx, o->n_steps));
    for (int64_t s = 0; s < o->n_steps; ++s) {
        a.step = s;
        a.hour = s / o->steps_per_hour;
        a.parking = f->counts.as<uint32_t>() + (size_t)N * s;
        a.driving = f->counts.as<uint32_t>() + n_counts + (size_t)N * s;
        if (traj) { a.t_traj = s; a.write_next_state = (s + 1 < traj->T) ? 1 : 0; }
        if (ctx->step_wait && o->with_time && s % o->steps_per_hour == 0 && a.hour < ctx->n_step_wait)
            CU(cudaStreamWaitEvent(ctx->st, ctx->step_wait[a.hour], 0));   // this hour's std slab is still on its way up
        LAUNCH(ctx, launch_step_gather(a, ctx->st));
    }
    TRY(profile_close(ctx));
    f->bucket_valid = false;
    return CPM_OK;
}

// bucket sampler: cars grouped by zone, CDF rows staged in shared memory (bucket.cu)
// Group the cars of a fleet by zone from its car-id indexed zone array: perm[0] (car ids by bin), the bucket starts, the
// bucket sizes (cnt_cur) and the first bin of every slab.  Leaves the fleet in the state a finished bucket day leaves it in.
static int32_t fleet_bucketize_async(cpm_fleet *f, int64_t N, int64_t ldc, int with_time, cudaStream_t st = nullptr)
{
    cpm_ctx *ctx = f->ctx;
    if (!st) st = ctx->st;
    const int NB = (int)N + 1;
    if (!f->zone_valid) return fail(CPM_ERR_STATE, "fleet zones are not materialised");
    BucketGeom g;
    if (!bucket_geometry(N, ldc, f->C, with_time, &g)) return fail(CPM_ERR_ARG, "the bucket sampler does not support this city size");
    TRY(f->perm[0].ensure((size_t)f->C * 4));
    TRY(f->cnt_cur.ensure((size_t)NB * 4)); TRY(f->start_tmp.ensure((size_t)(NB + 1) * 4)); TRY(f->cursor_tmp.ensure((size_t)NB * 4));
    TRY(f->first_zone.ensure((size_t)g.n_slabs * 4));
    f->cur = 0;
    LAUNCH(ctx, launch_bucketize(f->zone.as<int32_t>(), f->C, NB, f->cnt_cur.as<uint32_t>(), f->start_tmp.as<uint32_t>(),
                                 f->cursor_tmp.as<uint32_t>(), f->perm[0].as<int32_t>(), f->first_zone.as<uint32_t>(), g, st));
    f->bucket_valid = true; f->bucket_N = N; f->geom = g;
    return CPM_OK;
}

static int ilog2(int64_t v) { int l = 0; while ((1ll << (l + 1)) <= v) ++l; return l; }

static int32_t day_bucket(cpm_city *city, cpm_fleet *f, const cpm_day_opts *o)
{
    cpm_ctx *ctx = f->ctx;
    const int64_t N = city->N, C = f->C;
    const int NB = (int)N + 1;
    const int64_t S = o->n_steps;
    BucketGeom g;
    if (!bucket_geometry(N, city->ldc, C, o->with_time, &g)) return fail(CPM_ERR_ARG, "the bucket sampler does not support this city size");
    const size_t tot_words = (size_t)NB * kTotStride;
    TRY(f->perm[0].ensure((size_t)C * 4)); TRY(f->perm[1].ensure((size_t)C * 4));
    TRY(f->dest.ensure((size_t)C * 4));                                   // (destination, rank) per slot
    TRY(f->pos.ensure((size_t)g.n_slabs * NB * sizeof(uint2)));           // per-slab {count, first position} rows
    TRY(f->tot.ensure(2 * tot_words * 4));
    TRY(f->first_zone.ensure((size_t)g.n_slabs * 4));
    TRY(f->bcnt.ensure((size_t)(S + 1) * NB * 4)); TRY(f->bdrv.ensure((size_t)S * NB * 4));
    TRY(f->cnt_cur.ensure((size_t)NB * 4)); TRY(f->start_tmp.ensure((size_t)(NB + 1) * 4)); TRY(f->cursor_tmp.ensure((size_t)NB * 4));
    uint32_t *bcnt = f->bcnt.as<uint32_t>(), *bdrv = f->bdrv.as<uint32_t>(), *start = f->start_tmp.as<uint32_t>();
    uint32_t *tot = f->tot.as<uint32_t>();
    CU(cudaMemsetAsync(bdrv, 0, (size_t)S * NB * 4, ctx->st));
    CU(cudaMemsetAsync(tot, 0, 2 * tot_words * 4, ctx->st));
    const bool grouped = f->bucket_valid && f->bucket_N == N && f->geom.n_slabs == g.n_slabs && f->geom.slab == g.slab;
    if (!grouped) TRY(fleet_bucketize_async(f, N, city->ldc, o->with_time));
    // grouped (from the previous day, a copy of a grouped fleet, or just now): bucket sizes and bucket starts are current
    CU(cudaMemcpyAsync(bcnt, f->cnt_cur.p, (size_t)NB * 4, cudaMemcpyDeviceToDevice, ctx->st));
    Scalars *sc = f->scalars.as<Scalars>();
    BucketArgs a;
    memset(&a, 0, sizeof a);
    a.p_drive = city->p_drive.as<double>(); a.cdf = city->cdf.as<double>(); a.ldc = city->ldc; a.goff = city->goff; a.G = city->G;
    a.gshift = 32 - ilog2(city->G);
    a.rowzero = city->rowzero.as<uint8_t>(); a.dm = city->dm.as<double>();
    a.msd = city->have_msd ? city->msd.as<double2>() : nullptr;
    a.N = N; a.T_tab = city->T; a.C = C; a.slab = g.slab; a.ring = g.ring; a.ring_shift = ilog2(g.ring);
    { static const int gp = getenv("CPM_GATHER_EVICT_FIRST") ? atoi(getenv("CPM_GATHER_EVICT_FIRST")) : 0; a.gather_policy = gp; }
    a.start = start; a.first_zone = f->first_zone.as<uint32_t>(); a.dr = f->dest.as<uint32_t>(); a.M = f->pos.as<uint2>();
    a.with_time = o->with_time;
    a.rng = make_rng(o->rng, C);
    a.t_hi = &sc->t_hi; a.t_lo = &sc->t_lo;
    ScatterArgs sa;
    memset(&sa, 0, sizeof sa);
    sa.NB = NB; sa.dr = a.dr; sa.C = C; sa.slab = g.slab; sa.M = a.M; sa.start_next = start;
    sa.first_zone_next = f->first_zone.as<uint32_t>();
    for (int64_t s = 0; s < S; ++s) {
        a.step = s; a.hour = s / o->steps_per_hour;
        a.perm_cur = f->perm[f->cur].as<int32_t>();
        a.drv = bdrv + (size_t)s * NB;
        a.tot = tot + (size_t)(s & 1) * tot_words;
        // events bracket the sampling kernel alone (the dominant kernel); every third launch only, so that the event
        // records do not perturb the back-to-back launch pipeline of the timed region (3 is coprime to 23, 24 and 47:
        // over a few days every hour is sampled equally often)
        if (ctx->step_wait && o->with_time && s % o->steps_per_hour == 0 && a.hour < ctx->n_step_wait)
            CU(cudaStreamWaitEvent(ctx->st, ctx->step_wait[a.hour], 0));   // this hour's std slab is still on its way up
        const bool timed = ctx->profiling && (ctx->prof_tick++ % 3 == 0);
        if (timed) TRY(profile_open(ctx, 1));
        LAUNCH(ctx, launch_bucket_sample(a, g, ctx->st));
        if (timed) TRY(profile_close(ctx));
        sa.tot = a.tot; sa.tot_zero = tot + (size_t)((s + 1) & 1) * tot_words;
        sa.cnt_next = bcnt + (size_t)(s + 1) * NB;
        sa.perm_cur = a.perm_cur; sa.perm_next = f->perm[f->cur ^ 1].as<int32_t>();
        sa.zone_out = (s == S - 1) ? f->zone.as<int32_t>() : nullptr;
        LAUNCH(ctx, launch_bucket_scatter(sa, g, ctx->st));
        f->cur ^= 1;
    }
    CU(cudaMemcpyAsync(f->cnt_cur.p, bcnt + (size_t)S * NB, (size_t)NB * 4, cudaMemcpyDeviceToDevice, ctx->st));
    LAUNCH(ctx, launch_bucket_export_counts(bcnt, bdrv, (int)N, S, f->counts.as<uint32_t>(), f->counts.as<uint32_t>() + (size_t)N * S,
                                            &sc->lost, ctx->st));
    f->bucket_valid = true; f->bucket_N = N; f->geom = g; f->zone_valid = true;
    return CPM_OK;
}

// car-resident sampler: the whole day in one launch (daycars.cu); per hour when the hour's std slab is still on its way up
// sampler of the reference-named host functions (trajectory recording): the day kernel; CPM_TRAJ_SAMPLER=1 keeps the
// per-step gather kernel
static int32_t host_traj_sampler()
{
    static const int v = getenv("CPM_TRAJ_SAMPLER") ? atoi(getenv("CPM_TRAJ_SAMPLER")) : CPM_SAMPLER_AUTO;
    return v == CPM_SAMPLER_GATHER ? CPM_SAMPLER_GATHER : CPM_SAMPLER_AUTO;
}

static int32_t day_cars(cpm_city *city, cpm_fleet *f, const cpm_day_opts *o, bool want_counts, const Traj *traj)
{
    cpm_ctx *ctx = f->ctx;
    const int64_t N = city->N, C = f->C, S = o->n_steps;
    const size_t NB = (size_t)N + 1;
    CarsGeom g;
    if (!cars_geometry(N, C, o->with_time, want_counts ? 1 : 0, &g)) return fail(CPM_ERR_ARG, "the car-resident sampler does not support this city size");
    if (!f->zone_valid) return fail(CPM_ERR_STATE, "fleet zones are not materialised");
    if (want_counts) {
        TRY(f->cars_cnt.ensure((2 * (size_t)S + 1) * NB * 4));
        CU(cudaMemsetAsync(f->cars_cnt.p, 0, (2 * (size_t)S + 1) * NB * 4, ctx->st));
    }
    Scalars *sc = f->scalars.as<Scalars>();
    DayArgs a;
    memset(&a, 0, sizeof a);
    a.p_drive = city->p_drive.as<double>(); a.cdf = city->cdf.as<double>(); a.ldc = city->ldc; a.goff = city->goff; a.G = city->G;
    a.gshift = 32 - ilog2(city->G);
    a.rowzero = city->rowzero.as<uint8_t>(); a.dm = city->dm.as<double>();
    a.msd = city->have_msd ? city->msd.as<double2>() : nullptr;
    a.N = N; a.T_tab = city->T; a.C = C; a.steps_per_hour = o->steps_per_hour; a.with_time = o->with_time;
    a.cars_per_cta = g.cars_per_cta;
    a.zone = f->zone.as<int32_t>();
    a.counts = want_counts ? 1 : 0;
    if (want_counts) { a.cnt0 = f->cars_cnt.as<uint32_t>(); a.net = a.cnt0 + NB; a.drv = a.net + (size_t)S * NB; }
    else a.lost_visits = &sc->lost;   // the scalar block was zeroed at the start of the day
    a.rng = make_rng(o->rng, C);
    a.t_hi = &sc->t_hi; a.t_lo = &sc->t_lo;
    if (traj) { a.traj_state = traj->state; a.traj_trans = traj->trans; a.T_traj = traj->T; a.dist = city->dist.as<double>(); }
    uint32_t *const cnt0 = a.cnt0;
    const bool by_hour = ctx->step_wait != nullptr && o->with_time;
    for (int64_t s0 = 0; s0 < S;) {
        const int64_t hour = s0 / o->steps_per_hour;
        int64_t n = S - s0;
        if (by_hour) {
            n = std::min<int64_t>(n, (hour + 1) * o->steps_per_hour - s0);
            if (hour < ctx->n_step_wait) CU(cudaStreamWaitEvent(ctx->st, ctx->step_wait[hour], 0));
        }
        a.step0 = s0; a.n_steps = n;
        TRY(profile_open(ctx, n));
        LAUNCH(ctx, launch_day_cars(a, g, ctx->st));
        TRY(profile_close(ctx));
        a.cnt0 = nullptr;   // the occupancy before the day's first step is counted once
        s0 += n;
    }
    if (want_counts)
        LAUNCH(ctx, launch_cars_export_counts(cnt0, a.net, a.drv, (int)N, S, f->counts.as<uint32_t>(), f->counts.as<uint32_t>() + (size_t)N * S,
                                              &sc->lost, ctx->st));
    f->zone_valid = true; f->bucket_valid = false;
    return CPM_OK;
}

// which sampler a day runs on: the car-resident kernel wherever it applies (AUTO), the bucket pair when asked for or when
// the per-zone tables of the car-resident kernel do not fit shared memory, the per-step gather kernel otherwise
static int32_t pick_sampler(const cpm_city *city, const cpm_fleet *f, int32_t want, int with_time, bool traj)
{
    if (want == CPM_SAMPLER_GATHER || city->G <= 0) return CPM_SAMPLER_GATHER;
    CarsGeom g;
    const bool cars = cars_geometry(city->N, f->C, with_time, 1, &g), bucket = bucket_supported(city->N, city->ldc) && !traj;
    if (want == CPM_SAMPLER_CARS) return cars ? CPM_SAMPLER_CARS : -1;
    if (want == CPM_SAMPLER_BUCKET) return bucket ? CPM_SAMPLER_BUCKET : -1;
    if (traj) return cars ? CPM_SAMPLER_CARS : CPM_SAMPLER_GATHER;   // the bucket pair does not record trajectories
    static const int prefer_bucket = getenv("CPM_PREFER_BUCKET") ? atoi(getenv("CPM_PREFER_BUCKET")) : 0;
    if (cars && !(prefer_bucket && bucket)) return CPM_SAMPLER_CARS;
    return bucket ? CPM_SAMPLER_BUCKET : CPM_SAMPLER_GATHER;
}

// Enqueues n_steps steps (no host sync).  Counts, driving-time sum and lost counter are reset first.
// `want_counts` = false (the IVP of a dataset pass: only the end state and the lost-car visits are used) lets the
// car-resident sampler skip the per-zone rows; the other samplers count regardless.
static int32_t run_day_async(cpm_city *city, cpm_fleet *f, const cpm_day_opts *o, const Traj *traj, bool want_counts = true)
{
    cpm_ctx *ctx = f->ctx;
    ARG(city->ctx == ctx);
    ARG(o->n_steps >= 1 && o->steps_per_hour >= 1);
    ARG((o->n_steps - 1) / o->steps_per_hour < city->T);
    ARG(o->rng.mode == CPM_RNG_PHILOX || o->rng.mode == CPM_RNG_INJECTED);
    ARG(o->sampler == CPM_SAMPLER_AUTO || o->sampler == CPM_SAMPLER_GATHER || o->sampler == CPM_SAMPLER_BUCKET || o->sampler == CPM_SAMPLER_CARS);
    if (!city->have_pdrive || !city->have_cdf) return fail(CPM_ERR_STATE, "sampling needs p_drive and the p_dest CDF (build or set them first)");
    if (o->with_time && !city->have_inputs) return fail(CPM_ERR_STATE, "travel-time sampling needs datamatrix and distance_matrix_km");
    if (o->rng.mode == CPM_RNG_INJECTED) {
        ARG(o->rng.u_drive && o->rng.u_dest);
        if (o->with_time) ARG(o->rng.u_time && (traj == nullptr || o->rng.u_dist));
    }
    const int64_t N = city->N;
    const int32_t sampler = pick_sampler(city, f, o->sampler, o->with_time, traj != nullptr);
    if (sampler < 0)
        return fail(CPM_ERR_ARG, "the bucket and car-resident samplers need N + 1 <= 65535, per-zone tables that fit shared memory and "
                                 "a guide table; the bucket sampler does not record trajectories");
    TRY(f->counts.ensure(2 * (size_t)N * o->n_steps * sizeof(uint32_t)));
    f->counts_N = N; f->counts_steps = o->n_steps;
    CU(cudaMemsetAsync(f->scalars.p, 0, sizeof(Scalars), ctx->st));
    f->last_sampler = sampler;
    if (sampler == CPM_SAMPLER_CARS) TRY(day_cars(city, f, o, want_counts, traj));
    else if (sampler == CPM_SAMPLER_BUCKET) TRY(day_bucket(city, f, o));
    else TRY(day_gather(city, f, o, traj));
    Scalars *sc = f->scalars.as<Scalars>();
    if (o->with_time) LAUNCH(ctx, launch_finalize_time(&sc->t_hi, &sc->t_lo, &sc->time_sum, ctx->st));
    return CPM_OK;
}

extern "C" int32_t cpm_run_day(cpm_city *city, cpm_fleet *fleet, const cpm_day_opts *opts)
{
    ARG(city && fleet && opts);
    TRY(use(fleet->ctx));
    return run_day_async(city, fleet, opts, nullptr);
}

extern "C" int32_t cpm_fleet_get_counts(cpm_fleet *f, int64_t *parking, int64_t *driving, double *time_sum, int64_t *n_lost)
{
    ARG(f != nullptr);
    TRY(use(f->ctx));
    if (f->counts_steps == 0) return fail(CPM_ERR_STATE, "no day has been run on this fleet");
    const size_t n = (size_t)f->counts_N * f->counts_steps;
    std::vector<uint32_t> h(2 * n);
    Scalars sc;
    CU(cudaMemcpyAsync(h.data(), f->counts.p, 2 * n * sizeof(uint32_t), cudaMemcpyDeviceToHost, f->ctx->st));
    CU(cudaMemcpyAsync(&sc, f->scalars.p, sizeof sc, cudaMemcpyDeviceToHost, f->ctx->st));
    CU(cudaStreamSynchronize(f->ctx->st));
    if (parking) for (size_t k = 0; k < n; ++k) parking[k] = h[k];
    if (driving) for (size_t k = 0; k < n; ++k) driving[k] = h[n + k];
    if (time_sum) *time_sum = sc.time_sum;
    if (n_lost) *n_lost = (int64_t)sc.lost;
    return CPM_OK;
}
extern "C" int32_t cpm_fleet_last_sampler(cpm_fleet *f, int32_t *sampler)
{
    ARG(f && sampler);
    *sampler = f->last_sampler;
    return CPM_OK;
}
extern "C" int32_t cpm_fleet_get_time_fixed(cpm_fleet *f, int64_t *seconds, int64_t *frac32)
{
    ARG(f && seconds && frac32);
    TRY(use(f->ctx));
    if (f->counts_steps == 0) return fail(CPM_ERR_STATE, "no day has been run on this fleet");
    Scalars sc;
    CU(cudaMemcpyAsync(&sc, f->scalars.p, sizeof sc, cudaMemcpyDeviceToHost, f->ctx->st));
    CU(cudaStreamSynchronize(f->ctx->st));
    *seconds = (int64_t)sc.t_hi;
    *frac32 = (int64_t)sc.t_lo;
    return CPM_OK;
}
extern "C" int32_t cpm_fleet_counts_device(cpm_fleet *f, void **dev_ptr, int64_t *n_elems)
{
    ARG(f && dev_ptr && n_elems);
    if (f->counts_steps == 0) return fail(CPM_ERR_STATE, "no day has been run on this fleet");
    *dev_ptr = f->counts.p;
    *n_elems = 2 * f->counts_N * f->counts_steps;
    return CPM_OK;
}

// ------------------------------------------------------------------------------------------------
// Small cities: a pass is a short sequence of kernels that take a few microseconds each.  graph_run captures the launch
// sequence of `body` (everything it enqueues on the context's stream and the streams forked from it) and replays it as
// one executable graph.  The first pass of a shape runs directly (it allocates), the second instantiates the graph, every
// later one re-captures (cheap: nothing is launched) and patches the executable graph's kernel arguments in place --
// seeds, parameters and pointers may all differ from pass to pass, only the shape of the sequence must not.  Same kernels,
// same arguments, same order: results are identical to the direct path.  Host cost per replayed pass: ~45 us capture,
// ~10 us update, ~25 us launch (profiles/README.md); with the one-launch day kernel a cfg1 pass is 13 launches and
// bound by the device's kernel-to-kernel latency, so the replay buys ~3 % there (0.37 -> 0.35 ms).
// CPM_GRAPH_MAX_CARS (default 300,000; 0 = never) bounds the fleets it is used for.
// ------------------------------------------------------------------------------------------------
static int64_t graph_max_cars()
{
    static const int64_t v = getenv("CPM_GRAPH_MAX_CARS") ? atoll(getenv("CPM_GRAPH_MAX_CARS")) : 300000;
    return v;
}

static int32_t graph_run(cpm_fleet *f, uint64_t key, const std::function<int32_t()> &body)
{
    cpm_ctx *ctx = f->ctx;
    cpm_fleet::PassGraph &pg = f->pg;
    const bool eligible = ctx->graphs && !pg.off && f->C <= graph_max_cars() && !ctx->profiling && ctx->st != nullptr &&
                          ctx->st != cudaStreamLegacy && ctx->st != cudaStreamPerThread;
    if (!eligible) return body();
    if (pg.key != key) {
        if (pg.exec) { cudaGraphExecDestroy(pg.exec); pg.exec = nullptr; }
        pg.key = key; pg.seen = 0;
    }
    if (pg.seen++ == 0) return body();   // first pass of this shape: direct (allocations, one-time attribute calls)
    static const bool dbg = getenv("CPM_GRAPH_DEBUG") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = dbg ? now() : 0.0;
    if (cudaStreamBeginCapture(ctx->st, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); pg.off = true; return body(); }
    const int64_t launches0 = ctx->launches;
    const int32_t rc = body();
    cudaGraph_t graph = nullptr;
    const cudaError_t e_end = cudaStreamEndCapture(ctx->st, &graph);
    const double t1 = dbg ? now() : 0.0;
    if (rc != CPM_OK) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc; }
    bool ok = e_end == cudaSuccess && graph != nullptr;
    if (ok && pg.exec) {
        cudaGraphExecUpdateResultInfo info;
        if (cudaGraphExecUpdate(pg.exec, graph, &info) != cudaSuccess) {   // the sequence changed shape: build it anew
            cudaGetLastError();
            cudaGraphExecDestroy(pg.exec); pg.exec = nullptr;
        }
    }
    if (ok && !pg.exec) ok = cudaGraphInstantiate(&pg.exec, graph, 0) == cudaSuccess;
    const double t2 = dbg ? now() : 0.0;
    if (graph) cudaGraphDestroy(graph);
    if (ok) ok = cudaGraphLaunch(pg.exec, ctx->st) == cudaSuccess;
    if (dbg) fprintf(stderr, "[graph] capture %.0f us, update/instantiate %.0f us, destroy+launch %.0f us\n", t1 - t0, t2 - t1, now() - t2);
    if (!ok) {   // nothing ran: capture is unavailable for this sequence here -- run it directly, now and from now on
        cudaGetLastError();
        if (pg.exec) { cudaGraphExecDestroy(pg.exec); pg.exec = nullptr; }
        pg.off = true;
        ctx->launches = launches0;
        return body();
    }
    ++ctx->graph_replays;
    return CPM_OK;
}

// ------------------------------------------------------------------------------------------------
// main.jl:82-98 on the device
// ------------------------------------------------------------------------------------------------
struct NormOut {   // device layout of the normalised outputs of one day: parking[N*T], traffic[T], counts[T]
    static size_t bytes(int64_t N, int64_t T) { return ((size_t)N * T + T) * sizeof(double) + (size_t)T * sizeof(unsigned long long); }
};

static int32_t normalise_async(cpm_fleet *f, int64_t N, int64_t T, int64_t C_total)
{
    TRY(f->norm_out.ensure(NormOut::bytes(N, T)));
    double *parking = f->norm_out.as<double>();
    double *traffic = parking + (size_t)N * T;
    unsigned long long *tc = reinterpret_cast<unsigned long long *>(traffic + T);
    const uint32_t *cp = f->counts.as<uint32_t>();
    LAUNCH(f->ctx, launch_normalise_counts(cp, cp + (size_t)N * T, N, T, C_total, parking, nullptr, traffic, tc, f->ctx->st));
    return CPM_OK;
}

// Enqueues main.jl:82-98 for one dataset on the context's stream (no host synchronisation): tables, fleet init, IVP,
// resampling, normalisation.  `std_hour_events` (nullable): event t = the std slab of hour t is resident; the std slab of
// datamatrix is only read there, so a host upload can still be in flight while the tables and the IVP run.
static int32_t sample_dataset_enqueue(cpm_city *city, cpm_fleet *fleet, const cpm_params *pr, int64_t cars_per_zone,
                                      uint64_t seed, int32_t sampler, const cudaEvent_t *std_hour_events,
                                      const std::function<int32_t()> &before_resampling = nullptr,
                                      cudaEvent_t dist_ready = nullptr)
{
    cpm_ctx *ctx = fleet->ctx;
    const int64_t N = city->N, T = city->T;
    ARG(fleet->C == N * cars_per_zone);
    city->have_meansum = false;   // one dataset pass recomputes everything, like the reference (no cached stages)
    TRY(fleet->lost_acc.ensure(sizeof(unsigned long long)));
    CU(cudaMemsetAsync(fleet->lost_acc.p, 0, sizeof(unsigned long long), ctx->st));
    // createpdrive (main.jl:82) and createpdestin (main.jl:85) are independent.  k_pdest_unnorm / k_pdest_norm run at
    // the HBM roofline, k_cdf_build at about half of it (bound by its transposing writes): the createpdrive chain runs
    // on a second stream from the moment the CDF build starts and uses the bandwidth that kernel leaves idle.
    if (!ctx->aux_st) {
        CU(cudaStreamCreateWithFlags(&ctx->aux_st, cudaStreamNonBlocking));
        for (auto &e : ctx->ev_aux) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    TRY(city_build_pdestin_async(city, pr->e_dest, false, ctx->ev_aux[0]));                 // main.jl:85
    CU(cudaStreamWaitEvent(ctx->aux_st, ctx->ev_aux[0], 0));
    if (dist_ready) CU(cudaStreamWaitEvent(ctx->aux_st, dist_ready, 0));   // row pipeline: the distances go up after the means
    TRY(city_build_pdrive_async(city, pr->p_min, pr->p_max, pr->e_drive, ctx->aux_st));     // main.jl:82
    // initializestates (main.jl:88) and the grouping of the cars by zone do not depend on the tables: second stream as well
    LAUNCH(ctx, launch_fleet_init(fleet->zone.as<int32_t>(), fleet->C, cars_per_zone, 0, ctx->aux_st));
    fleet->car_offset = 0; fleet->zone_valid = true; fleet->bucket_valid = false;
    if (pick_sampler(city, fleet, sampler, T > 1 ? 0 : 1, false) == CPM_SAMPLER_BUCKET)
        TRY(fleet_bucketize_async(fleet, N, city->ldc, T > 1 ? 0 : 1, ctx->aux_st));
    CU(cudaEventRecord(ctx->ev_aux[1], ctx->aux_st));
    CU(cudaStreamWaitEvent(ctx->st, ctx->ev_aux[1], 0));
    cpm_day_opts o;
    memset(&o, 0, sizeof o);
    o.steps_per_hour = 1; o.sampler = sampler;
    o.rng.mode = CPM_RNG_PHILOX; o.rng.seed = seed;
    if (T > 1) {
        o.n_steps = T - 1; o.with_time = 0; o.rng.stream = 0;                               // main.jl:91-92
        TRY(run_day_async(city, fleet, &o, nullptr, false));
        // lost-car visits of the IVP survive the next day's reset of the scalar block (device-to-device, no host sync)
        CU(cudaMemcpyAsync(fleet->lost_acc.p, &fleet->scalars.as<Scalars>()->lost, sizeof(unsigned long long),
                           cudaMemcpyDeviceToDevice, ctx->st));
    }
    if (before_resampling) TRY(before_resampling());   // row pipeline: the std column is scattered here
    // host-input pipeline: the resampling step of hour t waits for the std slab of hour t only (the upload runs on)
    ctx->step_wait = std_hour_events; ctx->n_step_wait = std_hour_events ? T : 0;
    o.n_steps = T; o.with_time = 1; o.rng.stream = 1;                                       // main.jl:95
    const int32_t rc_day = run_day_async(city, fleet, &o, nullptr);
    ctx->step_wait = nullptr; ctx->n_step_wait = 0;
    TRY(rc_day);
    TRY(normalise_async(fleet, N, T, fleet->C));                                            // saveresults.jl:6-28
    return CPM_OK;
}

// what fixes the shape of a pass's launch sequence (kind: 1 dataset pass, 2 sweep set)
static uint64_t pass_key(const cpm_city *city, const cpm_fleet *fleet, int32_t sampler, int kind)
{
    uint64_t k = 0xcbf29ce484222325ull;
    const int64_t v[] = {city->N, city->T, fleet->C, sampler, kind, city->G, city->ldc, city->have_msd ? 1 : 0};
    for (int64_t x : v) { k ^= (uint64_t)x; k *= 0x100000001b3ull; }
    return k | 1ull;
}

static int32_t sample_dataset_collect(cpm_city *city, cpm_fleet *fleet, double *parking, double *traffic, double *A_drive,
                                      int64_t *n_lost)
{
    cpm_ctx *ctx = fleet->ctx;
    const int64_t N = city->N, T = city->T;
    const double *d_parking = fleet->norm_out.as<double>();
    Scalars sc;
    unsigned long long lost_ivp = 0;
    if (parking) CU(cudaMemcpyAsync(parking, d_parking, (size_t)N * T * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    if (traffic) CU(cudaMemcpyAsync(traffic, d_parking + (size_t)N * T, (size_t)T * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaMemcpyAsync(&sc, fleet->scalars.p, sizeof sc, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaMemcpyAsync(&lost_ivp, fleet->lost_acc.p, sizeof lost_ivp, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    if (A_drive) *A_drive = sc.time_sum / ((double)fleet->C * (double)T * 60 * 60);         // averagedrivingtime.jl:10
    if (n_lost) *n_lost = (int64_t)lost_ivp + (int64_t)sc.lost;
    return CPM_OK;
}

extern "C" int32_t cpm_sample_dataset(cpm_city *city, cpm_fleet *fleet, const cpm_params *pr, int64_t cars_per_zone,
                                      uint64_t seed, int32_t sampler, double *parking, double *traffic, double *A_drive,
                                      int64_t *n_lost)
{
    ARG(city && fleet && pr && cars_per_zone >= 1);
    TRY(use(fleet->ctx));
    TRY(graph_run(fleet, pass_key(city, fleet, sampler, 1),
                  [&]() { return sample_dataset_enqueue(city, fleet, pr, cars_per_zone, seed, sampler, nullptr); }));
    return sample_dataset_collect(city, fleet, parking, traffic, A_drive, n_lost);
}

// The same from HOST inputs in one call: the upload runs on a second stream, the mean slab and the distances first; the
// tables and the IVP start as soon as those have landed, while the std slab (read only by the resampling day's travel
// times) is still crossing PCIe.  Pinned host memory (cpm_host_alloc) is needed for the copies to overlap.
extern "C" int32_t cpm_sample_dataset_host(cpm_city *city, cpm_fleet *fleet, const double *datamatrix,
                                           const double *distance_matrix_km, const cpm_params *pr, int64_t cars_per_zone,
                                           uint64_t seed, int32_t sampler, double *parking, double *traffic,
                                           double *A_drive, int64_t *n_lost)
{
    ARG(city && fleet && datamatrix && distance_matrix_km && pr && cars_per_zone >= 1);
    cpm_ctx *ctx = fleet->ctx;
    ARG(city->ctx == ctx);
    TRY(use(ctx));
    TRY(city_alloc_inputs(city));
    if (!ctx->copy_st) {
        CU(cudaStreamCreateWithFlags(&ctx->copy_st, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&ctx->ev_copy[0], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_copy[1], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_copy[2], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_copy[3], cudaEventDisableTiming));
    }
    const size_t slab = (size_t)city->N * city->N * city->T * sizeof(double);
    // the copy stream must not overwrite inputs that earlier work on the compute stream still reads
    CU(cudaEventRecord(ctx->ev_copy[2], ctx->st));
    CU(cudaStreamWaitEvent(ctx->copy_st, ctx->ev_copy[2], 0));
    CU(cudaMemcpyAsync(city->dist.p, distance_matrix_km, (size_t)city->N * city->N * sizeof(double), cudaMemcpyHostToDevice, ctx->copy_st));
    CU(cudaMemcpyAsync(city->dm.p, datamatrix, slab, cudaMemcpyHostToDevice, ctx->copy_st));                       // means
    CU(cudaEventRecord(ctx->ev_copy[0], ctx->copy_st));
    // stds, hour by hour, an event after each: the resampling step of hour t starts as soon as its slab is resident
    while ((int64_t)ctx->ev_hour.size() < city->T) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->ev_hour.push_back(e);
    }
    const size_t hour_bytes = (size_t)city->N * city->N * sizeof(double);
    for (int64_t t = 0; t < city->T; ++t) {
        CU(cudaMemcpyAsync(city->dm.as<char>() + slab + t * hour_bytes, reinterpret_cast<const char *>(datamatrix) + slab + t * hour_bytes,
                           hour_bytes, cudaMemcpyHostToDevice, ctx->copy_st));
        CU(cudaEventRecord(ctx->ev_hour[(size_t)t], ctx->copy_st));
    }
    CU(cudaEventRecord(ctx->ev_copy[1], ctx->copy_st));
    city->have_inputs = true;
    city_invalidate(city);
    int32_t rc;
    if (ctx->graphs && fleet->C <= graph_max_cars()) {
        // small city: the whole upload is a few hundred microseconds -- wait for all of it and replay the pass as a graph
        // (graph_run) instead of pipelining hour slabs
        CU(cudaStreamWaitEvent(ctx->st, ctx->ev_copy[1], 0));
        rc = graph_run(fleet, pass_key(city, fleet, sampler, 1),
                       [&]() { return sample_dataset_enqueue(city, fleet, pr, cars_per_zone, seed, sampler, nullptr); });
    } else {
        CU(cudaStreamWaitEvent(ctx->st, ctx->ev_copy[0], 0));
        rc = sample_dataset_enqueue(city, fleet, pr, cars_per_zone, seed, sampler, ctx->ev_hour.data());
    }
    if (rc != CPM_OK) { cudaStreamSynchronize(ctx->copy_st); cudaStreamSynchronize(ctx->st); return rc; }
    return sample_dataset_collect(city, fleet, parking, traffic, A_drive, n_lost);   // the caller may reuse its buffers after this
}

// main.jl:74-98 for one dataset straight from the CSV columns, pipelined: the id columns, the means and the distances
// go up first and are scattered (claims parked in the std slab); the tables and the IVP run while the std column is
// still on its way; it is scattered just before the resampling day, the only consumer of standard deviations.
extern "C" int32_t cpm_sample_dataset_rows(cpm_city *city, cpm_fleet *fleet, const int32_t *sourceid, const int32_t *dstid,
                                           const int32_t *hod, const double *mean, const double *stddev, int64_t n_rows,
                                           const double *distance_matrix_km, const cpm_params *pr, int64_t cars_per_zone,
                                           uint64_t seed, int32_t sampler, double *parking, double *traffic, double *A_drive,
                                           int64_t *n_lost, int64_t *n_skipped)
{
    ARG(city && fleet && distance_matrix_km && pr && cars_per_zone >= 1 && n_rows >= 0);
    ARG(n_rows == 0 || (sourceid && dstid && hod && mean && stddev));
    cpm_ctx *ctx = fleet->ctx;
    ARG(city->ctx == ctx);
    TRY(use(ctx));
    TRY(city_alloc_inputs(city));
    if (!ctx->copy_st) {
        CU(cudaStreamCreateWithFlags(&ctx->copy_st, cudaStreamNonBlocking));
        for (auto &e : ctx->ev_copy) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const size_t nr = (size_t)n_rows;
    const int64_t N = city->N, T = city->T;
    TRY(ctx->tmp.ensure(nr * (3 * sizeof(int32_t) + 2 * sizeof(double)) + 64));
    double *d_mean = ctx->tmp.as<double>();
    double *d_sd = d_mean + nr;
    int32_t *d_src = reinterpret_cast<int32_t *>(d_sd + nr), *d_dst = d_src + nr, *d_hod = d_dst + nr;
    unsigned long long *d_bad = reinterpret_cast<unsigned long long *>(ctx->tmp.as<char>() + (nr * (3 * sizeof(int32_t) + 2 * sizeof(double)) + 7) / 8 * 8);
    // the copy stream must not overwrite staging or inputs that earlier work on the compute stream still reads
    CU(cudaEventRecord(ctx->ev_copy[2], ctx->st));
    CU(cudaStreamWaitEvent(ctx->copy_st, ctx->ev_copy[2], 0));
    if (nr) {
        CU(cudaMemcpyAsync(d_src, sourceid, nr * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->copy_st));
        CU(cudaMemcpyAsync(d_dst, dstid, nr * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->copy_st));
        CU(cudaMemcpyAsync(d_hod, hod, nr * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->copy_st));
        CU(cudaMemcpyAsync(d_mean, mean, nr * sizeof(double), cudaMemcpyHostToDevice, ctx->copy_st));
    }
    CU(cudaEventRecord(ctx->ev_copy[0], ctx->copy_st));
    // the distances are read by createpdrive only, which runs on the second stream after the two streaming passes of createpdestin
    CU(cudaMemcpyAsync(city->dist.p, distance_matrix_km, (size_t)N * N * sizeof(double), cudaMemcpyHostToDevice, ctx->copy_st));
    CU(cudaEventRecord(ctx->ev_copy[3], ctx->copy_st));
    if (nr) CU(cudaMemcpyAsync(d_sd, stddev, nr * sizeof(double), cudaMemcpyHostToDevice, ctx->copy_st));
    CU(cudaEventRecord(ctx->ev_copy[1], ctx->copy_st));
    // compute stream: zeros(N,N,T,2) (src/createdatamatrix.jl:7) while the columns travel, then the means
    CU(cudaMemsetAsync(city->dm.p, 0, (size_t)N * N * T * 2 * sizeof(double), ctx->st));
    CU(cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), ctx->st));
    CU(cudaStreamWaitEvent(ctx->st, ctx->ev_copy[0], 0));
    LAUNCH(ctx, launch_rows_scatter_means(d_src, d_dst, d_hod, d_mean, n_rows, N, T, city->dm.as<double>(), d_bad, ctx->st));
    city->have_inputs = true;
    city_invalidate(city);
    auto scatter_stds = [&]() -> int32_t {
        CU(cudaStreamWaitEvent(ctx->st, ctx->ev_copy[1], 0));
        LAUNCH(ctx, launch_rows_scatter_stds(d_src, d_dst, d_hod, d_sd, n_rows, N, T, city->dm.as<double>(), ctx->st));
        return CPM_OK;
    };
    int32_t rc = sample_dataset_enqueue(city, fleet, pr, cars_per_zone, seed, sampler, nullptr, scatter_stds, ctx->ev_copy[3]);
    if (rc != CPM_OK) { cudaStreamSynchronize(ctx->copy_st); cudaStreamSynchronize(ctx->st); return rc; }
    unsigned long long bad = 0;
    CU(cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, ctx->st));
    rc = sample_dataset_collect(city, fleet, parking, traffic, A_drive, n_lost);   // synchronises: host columns are free again
    if (n_skipped) *n_skipped = (int64_t)bad;
    return rc;
}

static bool hk_match(cpm_ctx *ctx, const cpm_ctx::HostKey &k, const void *p, size_t n);
static void hk_set(cpm_ctx::HostKey &k, const void *p, size_t n);

// Sweep state that survives between cpm_sweep calls on a context: the grouped initial fleet, two work fleets and a
// pinned staging block.  The notebook's searches issue one trial (or a small batch) per call
// (README.md:1133-1316, :1461-1681): re-creating fleets and re-uploading the 19 MB initial state cost more than a trial.
struct SweepState {
    cpm_fleet *base = nullptr, *work[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    struct Stage { unsigned long long tc[64]; Scalars sc; } *stage = nullptr;   // pinned, [2]
    cpm_ctx::HostKey init_key;
    int64_t C = 0, N = 0;
};
static void sweep_state_free(SweepState *&w)
{
    if (!w) return;
    if (w->base) cpm_fleet_destroy(w->base);
    for (auto &f : w->work) if (f) cpm_fleet_destroy(f);
    for (auto &e : w->done) if (e) cudaEventDestroy(e);
    if (w->stage) cudaFreeHost(w->stage);
    delete w;
    w = nullptr;
}

extern "C" int32_t cpm_sweep(cpm_city *city, const int64_t *initial_zones, int64_t C, const cpm_params *params, int64_t P,
                             uint64_t seed, const uint32_t *streams, int64_t *traffic_counts, double *A_drive,
                             double *parking)
{
    ARG(city && initial_zones && params && P >= 1 && C >= 1);
    cpm_ctx *ctx = city->ctx;
    TRY(use(ctx));
    const int64_t N = city->N, T = city->T;
    ARG(T <= 64);
    SweepState *&w = ctx->sweep;
    if (w && (w->C != C || w->N != N)) sweep_state_free(w);
    if (!w) {
        w = new (std::nothrow) SweepState();
        if (!w) return fail(CPM_ERR_OOM, "host allocation failed");
        w->C = C; w->N = N;
        int32_t rc = cpm_fleet_create(ctx, C, &w->base);
        for (int k = 0; k < 2 && rc == CPM_OK; ++k) rc = cpm_fleet_create(ctx, C, &w->work[k]);
        cudaError_t e = cudaSuccess;
        for (int k = 0; k < 2 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&w->done[k], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void **>(&w->stage), 2 * sizeof(SweepState::Stage), cudaHostAllocDefault);
        if (rc != CPM_OK || e != cudaSuccess) {
            sweep_state_free(w);
            return rc != CPM_OK ? rc : fail(CPM_ERR_CUDA, "sweep state: %s", cudaGetErrorString(e));
        }
    }
    TRY(city_alloc_cdf(city));
    const bool bucket = pick_sampler(city, w->base, CPM_SAMPLER_AUTO, 1, false) == CPM_SAMPLER_BUCKET;
    // the initial state: uploaded and grouped by zone once, as long as the caller passes the same unchanged array
    const bool same_init = hk_match(ctx, w->init_key, initial_zones, (size_t)C * sizeof(int64_t)) && w->base->zone_valid &&
                           (!bucket || (w->base->bucket_valid && w->base->bucket_N == N));
    if (!same_init) {
        TRY(cpm_fleet_set_zones(w->base, initial_zones));
        if (bucket) TRY(fleet_bucketize_async(w->base, N, city->ldc, 1));   // every set takes a copy of the grouping along
        hk_set(w->init_key, initial_zones, (size_t)C * sizeof(int64_t));
    }
    if (P >= 2) TRY(cpm_city_prepare_multi_day(city));   // many days on the same inputs: (mean, std) pair table
    // sets are processed grouped by e_dest (stable), so every distinct e_dest builds its CDF once; results land at the
    // caller's index p and the Philox stream is tied to p, so the order of evaluation is invisible to the caller
    std::vector<int64_t> order((size_t)P);
    for (int64_t k = 0; k < P; ++k) order[(size_t)k] = k;
    std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return params[x].e_dest < params[y].e_dest; });
    bool have_e_dest = false;
    double cur_e_dest = 0.0;
    int64_t in_flight[2] = {-1, -1};   // the set whose results sit in work[k] / stage[k]
    auto collect = [&](int k) -> int32_t {   // results of the set that ran in slot k (its copies were enqueued behind it)
        const int64_t p = in_flight[k];
        if (p < 0) return CPM_OK;
        CU(cudaEventSynchronize(w->done[k]));
        if (traffic_counts) for (int64_t t = 0; t < T; ++t) traffic_counts[t + T * p] = (int64_t)w->stage[k].tc[t];
        if (A_drive) A_drive[p] = w->stage[k].sc.time_sum / ((double)C * (double)T * 60 * 60);
        in_flight[k] = -1;
        return CPM_OK;
    };
    int32_t rc = CPM_OK;
    for (int64_t q = 0; rc == CPM_OK && q < P; ++q) {
        const int64_t p = order[(size_t)q];
        const cpm_params &pr = params[p];
        const int k = (int)(q & 1);
        rc = collect(k);   // slot k is free again; the other slot's set is still running: the host stays one set ahead
        if (rc != CPM_OK) break;
        cpm_fleet *work = w->work[k];
        if (!have_e_dest || pr.e_dest != cur_e_dest || !city->have_cdf) {   // CDFs are shared by sets with equal e_dest
            rc = city_build_pdestin_async(city, pr.e_dest, false);
            have_e_dest = true; cur_e_dest = pr.e_dest;
        }
        cpm_day_opts o;
        memset(&o, 0, sizeof o);
        o.n_steps = T; o.steps_per_hour = 1; o.with_time = 1; o.sampler = CPM_SAMPLER_AUTO;
        o.rng.mode = CPM_RNG_PHILOX; o.rng.seed = seed;
        o.rng.stream = streams ? streams[p] : (uint32_t)(2 + p);
        // one set = p_drive from the cached row sums, a copy of the initial state, one day, the normalisation: the same
        // launch sequence for every set, replayed as a graph for small fleets (graph_run)
        if (rc == CPM_OK) rc = graph_run(work, pass_key(city, work, o.sampler, 2), [&]() -> int32_t {
            TRY(city_build_pdrive_async(city, pr.p_min, pr.p_max, pr.e_drive));   // mean_sum cached
            TRY(cpm_fleet_copy_zones(work, w->base));                             // README.md:1014,1549
            TRY(run_day_async(city, work, &o, nullptr));
            return normalise_async(work, N, T, C);
        });
        if (rc != CPM_OK) break;
        const double *d_parking = work->norm_out.as<double>();
        const unsigned long long *d_tc = reinterpret_cast<const unsigned long long *>(d_parking + (size_t)N * T + T);
        cudaError_t e = cudaMemcpyAsync(w->stage[k].tc, d_tc, (size_t)T * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&w->stage[k].sc, work->scalars.p, sizeof(Scalars), cudaMemcpyDeviceToHost, ctx->st);
        if (e == cudaSuccess && parking)
            e = cudaMemcpyAsync(parking + (size_t)N * T * p, d_parking, (size_t)N * T * sizeof(double), cudaMemcpyDeviceToHost, ctx->st);
        if (e == cudaSuccess) e = cudaEventRecord(w->done[k], ctx->st);
        if (e != cudaSuccess) { rc = fail(CPM_ERR_CUDA, "sweep copy-out: %s", cudaGetErrorString(e)); break; }
        in_flight[k] = p;
    }
    for (int k = 0; k < 2; ++k) { const int32_t r2 = collect(k); if (rc == CPM_OK) rc = r2; }
    if (rc != CPM_OK) cudaStreamSynchronize(ctx->st);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// Host API: the reference's Julia functions with host pointers
// ------------------------------------------------------------------------------------------------
// Sampled 64-bit checksum of a host array: the first and last 128 words and 65,536 evenly spaced ones.  It decides
// whether the device copy made from this array in an earlier call is still current.  Contract: the
// reference never writes into datamatrix, distance_matrix_km, p_drive or p_dest after building them, and the trajectory
// arrays change only through these functions and `state_matrix[:,1] = initial_state` (main.jl:92), which is always
// re-read; a caller that edits arrays in place between calls switches the cache off (cpm_ctx_set_input_cache).
static uint64_t host_sig(const void *p, size_t bytes)
{
    const uint64_t *w = static_cast<const uint64_t *>(p);
    const size_t n = bytes / 8;
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t)bytes;
    auto mix = [&](uint64_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 29; };
    if (n <= 65536 + 256) { for (size_t k = 0; k < n; ++k) mix(w[k]); return h; }
    for (size_t k = 0; k < 128; ++k) { mix(w[k]); mix(w[n - 1 - k]); }
    const size_t step = n / 65536;
    for (size_t k = 0; k < 65536; ++k) mix(w[k * step + (k % step)]);
    return h;
}
static bool hk_match(cpm_ctx *ctx, const cpm_ctx::HostKey &k, const void *p, size_t n)
{
    const bool hit = ctx->cache_inputs && k.valid && k.p == p && k.n == n && k.sig == host_sig(p, n);
    if (hit) ++ctx->uploads_skipped;
    return hit;
}
static void hk_set(cpm_ctx::HostKey &k, const void *p, size_t n) { k.p = p; k.n = n; k.sig = host_sig(p, n); k.valid = true; }
static void hk_clear_city(cpm_ctx *ctx) { ctx->hk_mean.valid = ctx->hk_sd.valid = ctx->hk_dist.valid = ctx->hk_pdest.valid = false; }
static void hk_clear_traj(cpm_ctx *ctx) { ctx->hk_state.valid = false; for (auto &k : ctx->hk_trans) k.valid = false; }

extern "C" int32_t cpm_ctx_set_input_cache(cpm_ctx *ctx, int32_t enabled)
{
    ARG(ctx != nullptr);
    ctx->cache_inputs = enabled != 0;
    if (!enabled) { hk_clear_city(ctx); hk_clear_traj(ctx); }
    return CPM_OK;
}
extern "C" int32_t cpm_ctx_uploads_skipped(cpm_ctx *ctx, int64_t *count)
{
    ARG(ctx && count);
    *count = ctx->uploads_skipped;
    return CPM_OK;
}

extern "C" int32_t cpm_ctx_host_last_sampler(cpm_ctx *ctx, int32_t *sampler)
{
    ARG(ctx && sampler);
    *sampler = ctx->h_fleet ? ctx->h_fleet->last_sampler : 0;
    return CPM_OK;
}
extern "C" int32_t cpm_ctx_set_graphs(cpm_ctx *ctx, int32_t enabled)
{
    ARG(ctx != nullptr);
    ctx->graphs = enabled != 0;
    return CPM_OK;
}
extern "C" int32_t cpm_ctx_graph_replays(cpm_ctx *ctx, int64_t *count)
{
    ARG(ctx && count);
    *count = ctx->graph_replays;
    return CPM_OK;
}

static int32_t host_city(cpm_ctx *ctx, int64_t N, int64_t T, cpm_city **out)
{
    if (ctx->h_city && (ctx->h_city->N != N || ctx->h_city->T != T)) { cpm_city_destroy(ctx->h_city); ctx->h_city = nullptr; hk_clear_city(ctx); }
    if (!ctx->h_city) { TRY(cpm_city_create(ctx, N, T, &ctx->h_city)); hk_clear_city(ctx); }
    *out = ctx->h_city;
    return CPM_OK;
}
static int32_t host_fleet(cpm_ctx *ctx, int64_t C, cpm_fleet **out)
{
    if (ctx->h_fleet && ctx->h_fleet->C != C) { cpm_fleet_destroy(ctx->h_fleet); ctx->h_fleet = nullptr; }
    if (!ctx->h_fleet) TRY(cpm_fleet_create(ctx, C, &ctx->h_fleet));
    *out = ctx->h_fleet;
    return CPM_OK;
}

// Brings the parts of the caller's inputs a call needs into the host-API city (asynchronously on the context's stream;
// the caller synchronises before it returns).  datamatrix = [mean slab | std slab].
static int32_t host_inputs(cpm_ctx *ctx, cpm_city *c, const double *datamatrix, const double *distance_matrix_km, bool need_mean,
                           bool need_sd, bool need_dist)
{
    TRY(city_alloc_inputs(c));
    const size_t NN = (size_t)c->N * c->N, slab = NN * c->T * sizeof(double);
    const char *dmb = reinterpret_cast<const char *>(datamatrix);
    if (need_mean && !hk_match(ctx, ctx->hk_mean, dmb, slab)) {
        CU(cudaMemcpyAsync(c->dm.p, dmb, slab, cudaMemcpyHostToDevice, ctx->st));
        hk_set(ctx->hk_mean, dmb, slab);
        ctx->hk_pdest.valid = false;   // tables derived from the old means are gone with them
        city_invalidate(c);
    }
    if (need_sd && !hk_match(ctx, ctx->hk_sd, dmb + slab, slab)) {
        CU(cudaMemcpyAsync(c->dm.as<char>() + slab, dmb + slab, slab, cudaMemcpyHostToDevice, ctx->st));
        hk_set(ctx->hk_sd, dmb + slab, slab);
        c->have_msd = false;
    }
    if (need_dist && !hk_match(ctx, ctx->hk_dist, distance_matrix_km, NN * sizeof(double))) {
        CU(cudaMemcpyAsync(c->dist.p, distance_matrix_km, NN * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
        hk_set(ctx->hk_dist, distance_matrix_km, NN * sizeof(double));
        c->have_meansum = false;
    }
    c->have_inputs = ctx->hk_mean.valid && ctx->hk_sd.valid && ctx->hk_dist.valid;
    return CPM_OK;
}

extern "C" int32_t cpm_create_pdrive(cpm_ctx *ctx, const double *datamatrix, const double *distance_matrix_km, int64_t N,
                                     int64_t T, double p_min, double p_max, double e_drive, double *p_drive_out)
{
    TRY(use(ctx));
    ARG(datamatrix && distance_matrix_km && p_drive_out && N >= 1 && T >= 1);
    cpm_city *c;
    TRY(host_city(ctx, N, T, &c));
    // only the mean slab [N,N,T,1] and the distances are read (src/createpdrive.jl:16-18)
    TRY(host_inputs(ctx, c, datamatrix, distance_matrix_km, true, false, true));
    const bool full = c->have_inputs;
    c->have_inputs = true;
    int32_t rc = city_build_pdrive_async(c, p_min, p_max, e_drive);
    if (rc == CPM_OK) rc = cpm_city_get_pdrive(c, p_drive_out);
    c->have_inputs = full;
    return rc;
}

extern "C" int32_t cpm_create_pdestin(cpm_ctx *ctx, const double *datamatrix, int64_t N, int64_t T, double e_dest,
                                      double *p_dest_out)
{
    TRY(use(ctx));
    ARG(datamatrix && p_dest_out && N >= 1 && T >= 1);
    cpm_city *c;
    TRY(host_city(ctx, N, T, &c));
    // only the mean slab [N,N,T,1] is read (src/createpdestin.jl:13,23)
    TRY(host_inputs(ctx, c, datamatrix, nullptr, true, false, false));
    const bool full = c->have_inputs;
    c->have_inputs = true;
    int32_t rc = city_build_pdestin_async(c, e_dest);
    if (rc == CPM_OK) rc = cpm_city_get_pdest(c, p_dest_out);
    c->have_inputs = full;
    // the device keeps p_dest and its sampling CDF: a later call that passes this very array back finds them resident
    if (rc == CPM_OK) hk_set(ctx->hk_pdest, p_dest_out, (size_t)N * N * T * sizeof(double));
    return rc;
}

extern "C" int32_t cpm_initialize_states(cpm_ctx *ctx, int64_t C, int64_t T, int64_t cars_per_zone, int64_t *state, double *trans)
{
    TRY(use(ctx));
    ARG(state && trans && C >= 1 && T >= 1 && cars_per_zone >= 1);
    const size_t CT = (size_t)C * T;
    TRY(ctx->traj_state.ensure(CT * sizeof(int64_t)));
    TRY(ctx->traj_trans.ensure(CT * 4 * sizeof(double)));
    LAUNCH(ctx, launch_init_states(ctx->traj_state.as<int64_t>(), ctx->traj_trans.as<double>(), C, T, cars_per_zone, ctx->st));
    CU(cudaMemcpyAsync(state, ctx->traj_state.p, CT * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaMemcpyAsync(trans, ctx->traj_trans.p, CT * 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    hk_set(ctx->hk_state, state, CT * sizeof(int64_t));
    for (int k = 0; k < 4; ++k) hk_set(ctx->hk_trans[k], trans + (size_t)k * CT, CT * sizeof(double));
    return CPM_OK;
}

// uploads injected variates (host -> device) and rewrites the rng to device pointers
static int32_t stage_rng(cpm_ctx *ctx, const cpm_rng *in, int64_t C, int64_t n_steps, bool with_time, cpm_rng *out)
{
    *out = *in;
    if (in->mode != CPM_RNG_INJECTED) return CPM_OK;
    const double *src[4] = { in->u_drive, in->u_dest, in->u_time, in->u_dist };
    const double **dst[4] = { &out->u_drive, &out->u_dest, &out->u_time, &out->u_dist };
    const int n = with_time ? 4 : 2;
    for (int k = 0; k < n; ++k) {
        ARG(src[k] != nullptr);
        const size_t bytes = (size_t)C * n_steps * sizeof(double);
        TRY(ctx->u[k].ensure(bytes));
        CU(cudaMemcpyAsync(ctx->u[k].p, src[k], bytes, cudaMemcpyHostToDevice, ctx->st));
        *dst[k] = ctx->u[k].as<double>();
    }
    return CPM_OK;
}

static int32_t host_sample(cpm_ctx *ctx, int64_t *state, double *trans, const double *p_drive, const double *p_dest,
                           const double *datamatrix, const double *distance_matrix_km, int64_t C, int64_t N, int64_t T,
                           const cpm_rng *rng, bool ivp, int64_t *initial_state_out, int64_t *n_lost_out)
{
    ARG(state && trans && p_drive && p_dest && rng && C >= 1 && N >= 1 && T >= 1);
    cpm_city *city; cpm_fleet *fleet;
    TRY(host_city(ctx, N, T, &city));
    TRY(host_fleet(ctx, C, &fleet));
    if (!ivp) {
        ARG(datamatrix && distance_matrix_km);
        TRY(host_inputs(ctx, city, datamatrix, distance_matrix_km, true, true, true));
    }
    TRY(cpm_city_set_pdrive(city, p_drive));
    const size_t pd_bytes = (size_t)N * N * T * sizeof(double);
    if (!(city->have_cdf && hk_match(ctx, ctx->hk_pdest, p_dest, pd_bytes))) {   // else: p_dest and its CDF are resident
        TRY(cpm_city_set_pdest(city, p_dest));
        hk_set(ctx->hk_pdest, p_dest, pd_bytes);
    }
    const size_t CT = (size_t)C * T;
    const bool fresh = ctx->traj_state.cap < CT * sizeof(int64_t) || ctx->traj_trans.cap < CT * 4 * sizeof(double);
    if (fresh) hk_clear_traj(ctx);
    bool pre[2] = { false, false };   // did the device copies of slots 1-2 mirror the caller's array when this call began?
    if (ivp) for (int k = 0; k < 2; ++k) { pre[k] = hk_match(ctx, ctx->hk_trans[k], trans + (size_t)k * CT, CT * sizeof(double)); if (pre[k]) --ctx->uploads_skipped; }
    TRY(ctx->traj_state.ensure(CT * sizeof(int64_t)));
    TRY(ctx->traj_trans.ensure(CT * 4 * sizeof(double)));
    int64_t *d_state = ctx->traj_state.as<int64_t>();
    double *d_trans = ctx->traj_trans.as<double>();
    // column 1 of state_matrix is the only state input (src/resampling.jl:12 at t = 1)
    CU(cudaMemcpyAsync(d_state, state, (size_t)C * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->st));
    LAUNCH(ctx, launch_i64_to_i32(d_state, fleet->zone.as<int32_t>(), C, ctx->st));
    fleet->zone_valid = true; fleet->bucket_valid = false;
    const int64_t n_steps = ivp ? T - 1 : T;
    if (!ivp)   // slots 3-4 of non-driving cars keep their previous contents (src/resampling.jl:55-77)
        for (int k = 2; k < 4; ++k)
            if (!hk_match(ctx, ctx->hk_trans[k], trans + (size_t)k * CT, CT * sizeof(double)))
                CU(cudaMemcpyAsync(d_trans + (size_t)k * CT, trans + (size_t)k * CT, CT * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
    if (n_steps >= 1) {
        cpm_day_opts o;
        memset(&o, 0, sizeof o);
        o.n_steps = n_steps; o.steps_per_hour = 1; o.with_time = ivp ? 0 : 1; o.sampler = host_traj_sampler();
        TRY(stage_rng(ctx, rng, C, n_steps, !ivp, &o.rng));
        Traj tj; tj.state = d_state; tj.trans = d_trans; tj.T = T;
        TRY(run_day_async(city, fleet, &o, &tj));
        // copy back exactly the blocks the reference writes
        if (T > 1) CU(cudaMemcpyAsync(state + C, d_state + C, (size_t)C * (T - 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->st));
        const size_t ncol = (size_t)C * n_steps;
        CU(cudaMemcpyAsync(trans, d_trans, ncol * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaMemcpyAsync(trans + CT, d_trans + CT, ncol * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
        if (!ivp) CU(cudaMemcpyAsync(trans + 2 * CT, d_trans + 2 * CT, 2 * CT * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    }
    if (initial_state_out)   // state_matrix[:,24] (src/solveinitialvalueproblem.jl:57), generalised to column T
        CU(cudaMemcpyAsync(initial_state_out, d_state + (size_t)C * (T - 1), (size_t)C * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->st));
    Scalars sc;
    CU(cudaMemcpyAsync(&sc, fleet->scalars.p, sizeof sc, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    if (n_lost_out) *n_lost_out = n_steps >= 1 ? (int64_t)sc.lost : 0;
    // Host and device trajectories agree from here on wherever this call wrote or read them.  An IVP call leaves column T of
    // slots 1-2 and all of slots 3-4 alone on both sides: those keys stay what they were (slots 1-2 only if the host
    // array still matched before the call).
    if (ctx->cache_inputs) {
        hk_set(ctx->hk_state, state, CT * sizeof(int64_t));
        for (int k = 0; k < 4; ++k) {
            if (!ivp) hk_set(ctx->hk_trans[k], trans + (size_t)k * CT, CT * sizeof(double));
            else if (k < 2) { if (pre[k]) hk_set(ctx->hk_trans[k], trans + (size_t)k * CT, CT * sizeof(double)); else ctx->hk_trans[k].valid = false; }
        }
    } else hk_clear_traj(ctx);
    return CPM_OK;
}

extern "C" int32_t cpm_solve_ivp(cpm_ctx *ctx, int64_t *state, double *trans, const double *p_drive, const double *p_dest,
                                 int64_t C, int64_t N, int64_t T, const cpm_rng *rng, int64_t *initial_state_out,
                                 int64_t *n_lost_out)
{
    TRY(use(ctx));
    ARG(initial_state_out != nullptr);
    return host_sample(ctx, state, trans, p_drive, p_dest, nullptr, nullptr, C, N, T, rng, true, initial_state_out, n_lost_out);
}

extern "C" int32_t cpm_resampling(cpm_ctx *ctx, int64_t *state, double *trans, int64_t C, int64_t N, int64_t T,
                                  const double *p_drive, const double *p_dest, const double *datamatrix,
                                  const double *distance_matrix_km, const cpm_rng *rng, int64_t *n_lost_out)
{
    TRY(use(ctx));
    return host_sample(ctx, state, trans, p_drive, p_dest, datamatrix, distance_matrix_km, C, N, T, rng, false, nullptr, n_lost_out);
}

extern "C" int32_t cpm_average_driving_time(cpm_ctx *ctx, int64_t C, int64_t T, double A_drive_in, const double *trans,
                                            double *A_drive_out)
{
    TRY(use(ctx));
    ARG(trans && A_drive_out && C >= 1 && T >= 1);
    const size_t CT = (size_t)C * T;
    const int np = sampler_grid((int64_t)CT);
    TRY(ctx->partials.ensure((size_t)np * sizeof(double)));
    TRY(ctx->scalars.ensure(sizeof(double)));
    const double *d_slot3;
    if (ctx->traj_trans.cap >= 4 * CT * sizeof(double) && hk_match(ctx, ctx->hk_trans[2], trans + 2 * CT, CT * sizeof(double)))
        d_slot3 = ctx->traj_trans.as<double>() + 2 * CT;   // the trajectory resampling() just returned is still resident
    else {
        TRY(ctx->tmp.ensure(CT * sizeof(double)));
        CU(cudaMemcpyAsync(ctx->tmp.p, trans + 2 * CT, CT * sizeof(double), cudaMemcpyHostToDevice, ctx->st));   // slot 3
        d_slot3 = ctx->tmp.as<double>();
    }
    LAUNCH(ctx, launch_sum_f64(d_slot3, (int64_t)CT, ctx->partials.as<double>(), np, ctx->scalars.as<double>(), ctx->st));
    double s = 0.0;
    CU(cudaMemcpyAsync(&s, ctx->scalars.p, sizeof s, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    *A_drive_out = A_drive_in + s / ((double)C * (double)T * 60 * 60);   // :10
    return CPM_OK;
}

extern "C" int32_t cpm_correct_parameters(double p_min_next, double p_max_next, double p_min, double p_max, double *out2)
{
    ARG(out2 != nullptr);
    // four scalars of host arithmetic: src/correctparameters.jl:5-19
    if (p_min_next < 0) p_min_next = 0;
    else if (p_min_next > p_max) p_min_next = p_max;
    if (p_max_next > 1) p_max_next = 1;
    else if (p_max_next < p_min) p_max_next = p_min;
    out2[0] = p_min_next;
    out2[1] = p_max_next;
    return CPM_OK;
}

extern "C" int32_t cpm_distance_matrix(cpm_ctx *ctx, const double *centroid_lat, const double *centroid_long, int64_t N,
                                       double *distance_matrix_km)
{
    TRY(use(ctx));
    ARG(centroid_lat && centroid_long && distance_matrix_km && N >= 1);
    const size_t n = (size_t)N;
    TRY(ctx->tmp.ensure((2 * n + n * n) * sizeof(double)));
    double *d_lat = ctx->tmp.as<double>(), *d_lon = d_lat + n, *d_out = d_lon + n;
    CU(cudaMemcpyAsync(d_lat, centroid_lat, n * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
    CU(cudaMemcpyAsync(d_lon, centroid_long, n * sizeof(double), cudaMemcpyHostToDevice, ctx->st));
    LAUNCH(ctx, launch_distance_matrix(d_lat, d_lon, N, d_out, ctx->st));
    CU(cudaMemcpyAsync(distance_matrix_km, d_out, n * n * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return CPM_OK;
}

extern "C" int32_t cpm_polygon_centroids(cpm_ctx *ctx, const double *longitude_matrix, const double *latitude_matrix, int64_t N,
                                         int64_t V, double *centroid_lat, double *centroid_long, double *area_polygon)
{
    TRY(use(ctx));
    ARG(longitude_matrix && latitude_matrix && centroid_lat && centroid_long && N >= 1 && V >= 2);
    // the reference walks at most 10,000 vertices per zone (src/processgeodata.jl:107,132): only those columns travel
    const int64_t Vd = V < 10001 ? V : 10001;
    const size_t n = (size_t)N, nv = n * (size_t)Vd;
    TRY(ctx->tmp.ensure((2 * nv + 3 * n) * sizeof(double)));
    double *d_lon = ctx->tmp.as<double>(), *d_lat = d_lon + nv, *d_cla = d_lat + nv, *d_clo = d_cla + n, *d_ar = d_clo + n;
    CU(cudaMemcpyAsync(d_lon, longitude_matrix, nv * sizeof(double), cudaMemcpyHostToDevice, ctx->st));   // [N,V] column-major:
    CU(cudaMemcpyAsync(d_lat, latitude_matrix, nv * sizeof(double), cudaMemcpyHostToDevice, ctx->st));    // the first Vd columns are contiguous
    LAUNCH(ctx, launch_polygon_centroids(d_lon, d_lat, N, Vd, d_cla, d_clo, d_ar, ctx->st));
    CU(cudaMemcpyAsync(centroid_lat, d_cla, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaMemcpyAsync(centroid_long, d_clo, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    if (area_polygon) CU(cudaMemcpyAsync(area_polygon, d_ar, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return CPM_OK;
}

extern "C" int32_t cpm_zone_histogram(cpm_ctx *ctx, const int64_t *state, const double *trans, int64_t C, int64_t N, int64_t T,
                                      double *parking, double *driving, double *traffic)
{
    TRY(use(ctx));
    ARG(state && trans && C >= 1 && N >= 1 && T >= 1);
    const size_t CT = (size_t)C * T, NT = (size_t)N * T;
    TRY(ctx->partials.ensure(2 * NT * sizeof(uint32_t) + (2 * NT + 2 * T) * sizeof(double)));
    uint32_t *cnt = ctx->partials.as<uint32_t>();
    double *outp = reinterpret_cast<double *>(cnt + 2 * NT);   // NT even multiples keep 8-byte alignment
    const int64_t *d_state;
    const double *d_flag;
    const bool resident = ctx->traj_state.cap >= CT * sizeof(int64_t) && ctx->traj_trans.cap >= 4 * CT * sizeof(double) &&
                          hk_match(ctx, ctx->hk_state, state, CT * sizeof(int64_t)) && hk_match(ctx, ctx->hk_trans[0], trans, CT * sizeof(double));
    if (resident) { d_state = ctx->traj_state.as<int64_t>(); d_flag = ctx->traj_trans.as<double>(); }
    else {
        hk_clear_traj(ctx);   // the trajectory buffers are about to be overwritten with this call's arrays
        TRY(ctx->traj_state.ensure(CT * sizeof(int64_t)));
        TRY(ctx->tmp.ensure(CT * sizeof(double)));
        CU(cudaMemcpyAsync(ctx->traj_state.p, state, CT * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->st));
        CU(cudaMemcpyAsync(ctx->tmp.p, trans, CT * sizeof(double), cudaMemcpyHostToDevice, ctx->st));   // slot 1: drive flag
        d_state = ctx->traj_state.as<int64_t>(); d_flag = ctx->tmp.as<double>();
    }
    CU(cudaMemsetAsync(cnt, 0, 2 * NT * sizeof(uint32_t), ctx->st));
    LAUNCH(ctx, launch_traj_histogram(d_state, d_flag, C, N, T, cnt, cnt + NT, ctx->st));
    LAUNCH(ctx, launch_normalise_counts(cnt, cnt + NT, N, T, C, outp, outp + NT, outp + 2 * NT,
                                        reinterpret_cast<unsigned long long *>(outp + 2 * NT + T), ctx->st));
    if (parking) CU(cudaMemcpyAsync(parking, outp, NT * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    if (driving) CU(cudaMemcpyAsync(driving, outp + NT, NT * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    if (traffic) CU(cudaMemcpyAsync(traffic, outp + 2 * NT, (size_t)T * sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return CPM_OK;
}
