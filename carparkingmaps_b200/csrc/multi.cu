// multi.cu -- several GPUs from ONE host process, through the C ABI (include/cpm.h).
//
// The reference's city loop and dataset loop are serial (main.jl:46, :71) and a Julia host is a single process, so
// the natural drop-in for "this box has 8 GPUs" is one call that takes the whole list of datasets: one host thread and
// one context per device, datasets assigned by longest-processing-time, no collective on the data path (datasets are
// independent).  The second entry splits ONE city's cars over the devices (BASELINE configs[4]): contiguous car ranges
// with global Philox counters, tables replicated, then one NCCL all-reduce (SUM) of the uint32 count blocks -- NCCL is
// loaded at run time (dlopen), so libcpm.so has no link-time dependency on it.
// Built from the public entry points of capi.cu only.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <dlfcn.h>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/cpm.h"

namespace {


double dataset_cost(int64_t N, int64_t cars_per_zone, int64_t T)
{
    return (double)N * (double)cars_per_zone * (double)(2 * T - 1) + 3.0 * (double)N * (double)N * (double)T;
}

// longest-processing-time-first: deterministic, ties broken by index (the same rule as multigpu.lpt_assign)
std::vector<std::vector<int64_t>> lpt(const std::vector<double> &cost, int n_dev)
{
    std::vector<int64_t> order(cost.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return cost[(size_t)a] > cost[(size_t)b]; });
    std::vector<double> load((size_t)n_dev, 0.0);
    std::vector<std::vector<int64_t>> out((size_t)n_dev);
    for (int64_t k : order) {
        const size_t r = (size_t)(std::min_element(load.begin(), load.end()) - load.begin());
        out[r].push_back(k);
        load[r] += cost[(size_t)k];
    }
    return out;
}

// ---- NCCL, loaded at run time --------------------------------------------------------------------
typedef struct ncclComm *ncclComm_t;
struct Nccl {
    void *h = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool load(std::string *why)
    {
        if (h) return true;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) { *why = std::string("NCCL not found (dlopen libnccl.so.2): ") + dlerror(); return false; }
        CommInitAll = (decltype(CommInitAll))dlsym(h, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
        GroupStart = (decltype(GroupStart))dlsym(h, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(h, "ncclGroupEnd");
        AllReduce = (decltype(AllReduce))dlsym(h, "ncclAllReduce");
        GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
        if (!CommInitAll || !CommDestroy || !GroupStart || !GroupEnd || !AllReduce) { *why = "libnccl lacks a required symbol"; return false; }
        return true;
    }
};
Nccl g_nccl;
constexpr int kNcclUint32 = 3, kNcclSum = 0;   // ncclUint32, ncclSum (nccl.h)

}  // namespace

extern "C" void cpm_internal_set_error(const char *msg);   // capi.cu: sets the calling thread's cpm_last_error()
static void set_err(const std::string &m) { cpm_internal_set_error(m.c_str()); }

// main.jl:46-103 / :71-103 over several GPUs: n_datasets independent datasets (datamatrix[k]: [N[k],N[k],T,2],
// distance_matrix_km[k]: [N[k],N[k]], HOST arrays), one dataset pass of main.jl:82-98 each, results per dataset.
extern "C" int32_t cpm_sample_datasets_multi(const int32_t *devices, int32_t n_devices, int64_t n_datasets, const int64_t *N,
                                             int64_t T, const double *const *datamatrix, const double *const *distance_matrix_km,
                                             const cpm_params *params, int64_t cars_per_zone, uint64_t seed,
                                             double *const *parking_out, double *const *traffic_out, double *A_drive_out,
                                             int64_t *n_lost_out)
{
    if (!devices || n_devices < 1 || n_datasets < 0 || !N || !datamatrix || !distance_matrix_km || !params || cars_per_zone < 1 || T < 1) {
        set_err("bad argument");
        return CPM_ERR_ARG;
    }
    std::vector<double> cost((size_t)n_datasets);
    for (int64_t k = 0; k < n_datasets; ++k) cost[(size_t)k] = dataset_cost(N[k], cars_per_zone, T);
    const auto assign = lpt(cost, n_devices);
    std::vector<int32_t> rc((size_t)n_devices, CPM_OK);
    std::vector<std::string> msg((size_t)n_devices);
    std::vector<std::thread> th;
    for (int d = 0; d < n_devices; ++d)
        th.emplace_back([&, d]() {
            cpm_ctx *ctx = nullptr;
            int32_t r = cpm_ctx_create(devices[d], nullptr, &ctx);
            for (int64_t k : assign[(size_t)d]) {
                if (r != CPM_OK) break;
                cpm_city *city = nullptr;
                cpm_fleet *fleet = nullptr;
                r = cpm_city_create(ctx, N[k], T, &city);
                if (r == CPM_OK) r = cpm_fleet_create(ctx, N[k] * cars_per_zone, &fleet);
                double A = 0.0;
                int64_t lost = 0;
                if (r == CPM_OK)
                    r = cpm_sample_dataset_host(city, fleet, datamatrix[k], distance_matrix_km[k], params, cars_per_zone, seed + (uint64_t)k,
                                                CPM_SAMPLER_AUTO, parking_out ? parking_out[k] : nullptr, traffic_out ? traffic_out[k] : nullptr,
                                                &A, &lost);
                if (r == CPM_OK) { if (A_drive_out) A_drive_out[k] = A; if (n_lost_out) n_lost_out[k] = lost; }
                else msg[(size_t)d] = cpm_last_error();
                cpm_fleet_destroy(fleet);   // their blocks go to the device's free list: the next dataset reuses them
                cpm_city_destroy(city);
            }
            if (r != CPM_OK && msg[(size_t)d].empty()) msg[(size_t)d] = cpm_last_error();
            cpm_ctx_destroy(ctx);
            rc[(size_t)d] = r;
        });
    for (auto &t : th) t.join();
    for (int d = 0; d < n_devices; ++d)
        if (rc[(size_t)d] != CPM_OK) { set_err("device " + std::to_string(devices[d]) + ": " + msg[(size_t)d]); return rc[(size_t)d]; }
    return CPM_OK;
}

// One city, its cars split over the devices (contiguous ranges, car_offset = first global car of the range), n_steps
// steps, then ONE NCCL all-reduce of the count blocks.  Outputs (host): parking / driving Int64 [N, n_steps] of the
// WHOLE fleet, the driving-time sum (seconds), lost-car visits, and the device time of the all-reduce alone.
extern "C" int32_t cpm_split_city_multi(const int32_t *devices, int32_t n_devices, int64_t N, int64_t T, const double *datamatrix,
                                        const double *distance_matrix_km, const cpm_params *params, int64_t cars_per_zone,
                                        int64_t n_steps, int64_t steps_per_hour, int32_t with_time, uint64_t seed, uint32_t stream,
                                        int64_t *parking_out, int64_t *driving_out, double *time_sum_out, int64_t *n_lost_out,
                                        double *allreduce_ms_out)
{
    if (!devices || n_devices < 1 || N < 1 || T < 1 || !datamatrix || !distance_matrix_km || !params || cars_per_zone < 1 || n_steps < 1 ||
        N * cars_per_zone < n_devices) {
        set_err("bad argument");
        return CPM_ERR_ARG;
    }
    std::string why;
    if (n_devices > 1 && !g_nccl.load(&why)) { set_err(why); return CPM_ERR_STATE; }
    const int64_t C = N * cars_per_zone;
    struct Rank { cpm_ctx *ctx = nullptr; cpm_city *city = nullptr; cpm_fleet *fleet = nullptr; void *counts = nullptr; int64_t n = 0;
                  void *stream = nullptr; int64_t sec = 0, frac = 0, lost = 0; int32_t rc = CPM_OK; std::string msg; };
    std::vector<Rank> rk((size_t)n_devices);
    std::vector<std::thread> th;
    for (int d = 0; d < n_devices; ++d)
        th.emplace_back([&, d]() {
            Rank &r = rk[(size_t)d];
            const int64_t base = C / n_devices, rem = C % n_devices;
            const int64_t begin = d * base + std::min<int64_t>(d, rem), count = base + (d < rem ? 1 : 0);
            r.rc = cpm_ctx_create(devices[d], nullptr, &r.ctx);
            if (r.rc == CPM_OK) r.rc = cpm_city_create(r.ctx, N, T, &r.city);
            if (r.rc == CPM_OK) r.rc = cpm_city_upload(r.city, datamatrix, distance_matrix_km);
            if (r.rc == CPM_OK) r.rc = cpm_city_build_pdrive(r.city, params->p_min, params->p_max, params->e_drive);
            if (r.rc == CPM_OK) r.rc = cpm_city_build_pdestin(r.city, params->e_dest, 0);
            if (r.rc == CPM_OK) r.rc = cpm_fleet_create(r.ctx, std::max<int64_t>(count, 1), &r.fleet);
            if (r.rc == CPM_OK) r.rc = cpm_fleet_init(r.fleet, cars_per_zone, begin);
            cpm_day_opts o;
            memset(&o, 0, sizeof o);
            o.n_steps = n_steps; o.steps_per_hour = steps_per_hour; o.with_time = with_time; o.sampler = CPM_SAMPLER_AUTO;
            o.rng.mode = CPM_RNG_PHILOX; o.rng.seed = seed; o.rng.stream = stream; o.rng.car_offset = begin;
            if (r.rc == CPM_OK && count > 0) r.rc = cpm_run_day(r.city, r.fleet, &o);
            if (r.rc == CPM_OK) r.rc = cpm_fleet_counts_device(r.fleet, &r.counts, &r.n);
            if (r.rc == CPM_OK) r.rc = cpm_ctx_stream(r.ctx, &r.stream);
            if (r.rc == CPM_OK) r.rc = cpm_ctx_sync(r.ctx);
            if (r.rc == CPM_OK) r.rc = cpm_fleet_get_time_fixed(r.fleet, &r.sec, &r.frac);
            if (r.rc == CPM_OK) r.rc = cpm_fleet_get_counts(r.fleet, nullptr, nullptr, nullptr, &r.lost);
            if (r.rc != CPM_OK) r.msg = cpm_last_error();
        });
    for (auto &t : th) t.join();
    int32_t rc = CPM_OK;
    for (int d = 0; d < n_devices && rc == CPM_OK; ++d)
        if (rk[(size_t)d].rc != CPM_OK) { rc = rk[(size_t)d].rc; set_err("device " + std::to_string(devices[d]) + ": " + rk[(size_t)d].msg); }
    double ar_ms = 0.0;
    if (rc == CPM_OK && n_devices > 1) {
        std::vector<ncclComm_t> comm((size_t)n_devices);
        int e = g_nccl.CommInitAll(comm.data(), n_devices, devices);
        if (e != 0) { set_err(std::string("ncclCommInitAll: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(e) : "error")); rc = CPM_ERR_CUDA; }
        if (rc == CPM_OK) {
            cudaEvent_t e0, e1;
            cudaSetDevice(devices[0]);
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0, (cudaStream_t)rk[0].stream);
            g_nccl.GroupStart();
            for (int d = 0; d < n_devices; ++d)
                e |= g_nccl.AllReduce(rk[(size_t)d].counts, rk[(size_t)d].counts, (size_t)rk[(size_t)d].n, kNcclUint32, kNcclSum, comm[(size_t)d],
                                      (cudaStream_t)rk[(size_t)d].stream);
            e |= g_nccl.GroupEnd();
            cudaSetDevice(devices[0]);
            cudaEventRecord(e1, (cudaStream_t)rk[0].stream);
            for (int d = 0; d < n_devices; ++d) { cudaSetDevice(devices[d]); cudaStreamSynchronize((cudaStream_t)rk[(size_t)d].stream); }
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            ar_ms = ms;
            cudaEventDestroy(e0); cudaEventDestroy(e1);
            if (e != 0) { set_err("ncclAllReduce failed"); rc = CPM_ERR_CUDA; }
            for (auto c : comm) g_nccl.CommDestroy(c);
        }
    }
    if (rc == CPM_OK) {
        // every rank now holds the whole fleet's counts; rank 0's copy goes to the caller.  The driving-time sums are exact
        // integers (seconds, 2^-32 s): added on the host, canonicalised, converted once -- the unsplit run's value bit for bit.
        rc = cpm_fleet_get_counts(rk[0].fleet, parking_out, driving_out, nullptr, nullptr);
        if (rc != CPM_OK) set_err(cpm_last_error());
        unsigned long long sec = 0, frac = 0;
        int64_t lost = 0;
        for (auto &r : rk) { sec += (unsigned long long)r.sec; frac += (unsigned long long)r.frac; lost += r.lost; }
        sec += frac >> 32; frac &= 0xffffffffull;
        if (time_sum_out) *time_sum_out = (double)sec + (double)frac * (1.0 / 4294967296.0);
        if (n_lost_out) *n_lost_out = lost;
        if (allreduce_ms_out) *allreduce_ms_out = ar_ms;
    }
    for (auto &r : rk) { cpm_fleet_destroy(r.fleet); cpm_city_destroy(r.city); cpm_ctx_destroy(r.ctx); }
    return rc;
}
