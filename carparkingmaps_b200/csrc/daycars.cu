// daycars.cu -- the car-resident day kernel: ONE launch runs every step of a day (src/resampling.jl:11-83,
// src/solveinitialvalueproblem.jl:8-55, src/saveresults.jl:8-17).
//
// Cars never interact: a car's chain z(0) -> z(1) -> ... depends on the tables and on its own draws only (Philox keyed by
// global car id and step), and the per-zone counts are sums.  So a CTA takes a contiguous block of cars, keeps their
// zones in shared memory (16 bits per car) for the whole day and walks the steps on its own -- no sort, no exchange, no
// kernel boundary between steps.  Per step a warp
//   1. decides for each of its cars whether it drives (one Philox block, the hour's p_drive column from shared memory),
//   2. queues the drivers and serves them in dense batches of 32: guide cell -> a window of a few CDF entries of the
//      origin's row (global memory; the hour's rows are L2-resident for cities up to Melbourne size) -> destination,
//   3. queues the cars that changed zone and draws their travel times in dense batches (resampling day only).
// Counts: per-CTA shared-memory histograms (net arrivals and drivers per zone), flushed once per step with one reduction
// per non-empty bin into that step's global row; occupancy at step s+1 = occupancy at step s + net arrivals (export).
// Same draws as every other sampler => identical results (tests compare with the oracle and with the bucket sampler).
#include <math_constants.h>
#include <cstdlib>
#include <map>
#include <mutex>
#include "kernels.h"

namespace cpm {

constexpr int kCarQueue = 64;          // entries per warp queue (a batch is served as soon as 32 have queued up)
constexpr uint16_t kNoCar = 0xFFFFu;   // slot beyond the CTA's last car

template <int THREADS, bool INJ, bool TIME>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) k_day_cars(const DayArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long red[2];
    constexpr int WARPS = THREADS / 32;
    constexpr int U = 4;                 // cars per lane and trip: four independent Philox chains
    const int64_t N = a.N, C = a.C, NN = a.N * a.N;
    const int Ni = (int)N, NB = Ni + 1;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    // shared memory: drive queues | far queues | p_drive column | net arrivals | drivers | rowzero column | zones
    unsigned long long *const dq = reinterpret_cast<unsigned long long *>(smem_raw) + wid * kCarQueue;
    unsigned long long *const fq = reinterpret_cast<unsigned long long *>(smem_raw) + (WARPS + wid) * kCarQueue;
    double *const pd_s = reinterpret_cast<double *>(smem_raw) + (TIME ? 2 : 1) * WARPS * kCarQueue;
    uint32_t *const hnet = reinterpret_cast<uint32_t *>(pd_s + (a.pd_smem ? N : 0));
    uint32_t *const hdrv = hnet + NB;
    uint8_t *const rz_s = reinterpret_cast<uint8_t *>(hdrv + NB);
    uint16_t *const zs = reinterpret_cast<uint16_t *>(rz_s + ((N + 3) & ~(int64_t)3));
    const int cpc = a.cars_per_cta, spw = cpc / WARPS;   // slots per CTA / per warp (spw: a multiple of 32 * U)
    const int wbase = wid * spw;
    unsigned long long t_hi = 0ull, t_lo = 0ull;
    if (tid < 2) red[tid] = 0ull;

    for (int64_t base = (int64_t)blockIdx.x * cpc; base < C; base += (int64_t)gridDim.x * cpc) {
        const int n_cars = (int)((C - base < cpc) ? (C - base) : cpc);
        const int64_t car0 = base + wbase;   // the warp's first car
        __syncthreads();
        for (int x = tid; x < 2 * NB; x += THREADS) hnet[x] = 0u;   // hnet and hdrv are adjacent
        __syncthreads();
        // the CTA's cars: zones into shared memory, and their share of the occupancy at the first step
        for (int x = tid; x < cpc; x += THREADS) {
            uint16_t z16 = kNoCar;
            if (x < n_cars) {
                const int32_t z = a.zone[base + x];
                z16 = (z >= 1 && z <= Ni) ? (uint16_t)z : (uint16_t)0;   // anything else is a lost car (documented extension)
                if (a.cnt0) atomicAdd(&hnet[z16 ? z16 - 1 : Ni], 1u);
            }
            zs[x] = z16;
        }
        __syncthreads();
        if (a.cnt0) {
            for (int x = tid; x < NB; x += THREADS) {
                const uint32_t v = hnet[x];
                if (v) { atomicAdd(a.cnt0 + x, v); hnet[x] = 0u; }
            }
        }

        for (int64_t s = a.step0; s < a.step0 + a.n_steps; ++s) {
            const int64_t hour = s / a.steps_per_hour;
            const double *__restrict__ pd_t = a.p_drive + N * hour;
            const double *__restrict__ cdf_t = a.cdf + hour * N * a.ldc;
            const int64_t uoff = a.rng.ld * s;
            if (s == a.step0 || s % a.steps_per_hour == 0) {   // a new hour: its p_drive and rowzero columns
                __syncthreads();
                if (a.pd_smem) for (int x = tid; x < Ni; x += THREADS) pd_s[x] = pd_t[x];
                for (int x = tid; x < Ni; x += THREADS) rz_s[x] = a.rowzero[hour * N + x];
            }
            __syncthreads();   // columns loaded; the previous step's flush is done with the histograms

            uint32_t dq_head = 0, fq_head = 0;
            int dq_cnt = 0, fq_cnt = 0;
            auto far_batch = [&](int n) {   // travel times of up to 32 queued cars (src/resampling.jl:61-70), one per lane
                if (lane < n) {
                    const unsigned long long e = fq[(fq_head + lane) & (kCarQueue - 1)];
                    const int64_t car = car0 + (int64_t)(e & 0xFFFull);
                    const uint32_t odp = (uint32_t)(e >> 12);
                    const int64_t od = (int64_t)(odp & 0xffffu) + N * (int64_t)(odp >> 16);
                    double mean, sd;
                    if (a.msd) { const double2 ms = __ldg(a.msd + od + NN * hour); mean = ms.x; sd = ms.y; }
                    else { mean = __ldg(a.dm + od + NN * hour); sd = __ldg(a.dm + od + NN * (hour + a.T_tab)); }
                    double u3;
                    if (INJ) u3 = a.rng.u_time[car + uoff];
                    else { uint32_t o[4]; philox_words(a.rng, car, (uint32_t)s, 1u, o); u3 = u52(o[0], o[1]); }
                    if (sd == 0.0) sd = 0.1 * mean;                                   // :65-67
                    fx_add(t_hi, t_lo, truncnormal(mean, sd, u3));                    // :68
                }
            };
            auto drive_batch = [&](int n) {   // categorical destinations of up to 32 queued drivers (src/resampling.jl:26-49)
                bool far = false;
                uint32_t slot = 0, odp = 0;
                if (lane < n) {
                    const unsigned long long e = dq[(dq_head + lane) & (kCarQueue - 1)];
                    slot = (uint32_t)(e & 0xFFFull);
                    const unsigned long long uq = e >> 12;   // the 52-bit integer of u_dest (Philox)
                    const int z = (int)zs[wbase + slot];     // 1-based origin
                    const double *__restrict__ row = cdf_t + (int64_t)(z - 1) * a.ldc;
                    const uint16_t *__restrict__ gd = reinterpret_cast<const uint16_t *>(row + a.goff);
                    const double u2 = INJ ? a.rng.u_dest[car0 + slot + uoff]
                                          : __hiloint2double((int)((uint32_t)(uq >> 32) | 0x3FF00000u), (int)(uint32_t)uq) - 1.0;
                    int32_t dest = 0;
                    if (u2 > 0.0) {
                        int32_t lo, hi;
                        if (INJ && !(u2 < 1.0)) { lo = __ldg(gd + a.G); hi = Ni; }
                        else {
                            const uint32_t c = INJ ? (uint32_t)(u2 * (double)a.G) : (uint32_t)(uq >> (52 - (32 - a.gshift)));
                            lo = __ldg(gd + c); hi = __ldg(gd + c + 1);
                        }
                        // first index in [lo, hi] whose CDF value is not below u2 (NaN counts as not below, like the binary
                        // search of the other samplers: the row is non-decreasing with a NaN tail, so the predicate is
                        // monotone and the answer is the same index).  The guide leaves a window of a few entries: four
                        // independent loads per trip instead of a chain of dependent probes.
                        const int32_t last = hi < Ni ? hi : Ni - 1;
                        for (int32_t idx = lo; idx <= last; idx += 4) {
                            double r[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) r[q] = (idx + q <= last) ? __ldg(row + idx + q) : CUDART_INF;
                            int q = 4;
                            double v = 0.0;
#pragma unroll
                            for (int k = 3; k >= 0; --k) if (!(r[k] < u2)) { q = k; v = r[k]; }
                            if (q < 4) { if (idx + q <= last && u2 <= v) dest = idx + q + 1; break; }   // else 0: lost (:32,40)
                        }
                    }
                    if (dest != z) {
                        zs[wbase + slot] = (uint16_t)dest;                            // :81-83
                        atomicSub(&hnet[z - 1], 1u);
                        atomicAdd(&hnet[dest >= 1 ? dest - 1 : Ni], 1u);
                        if (TIME && dest >= 1) { far = true; odp = (uint32_t)(z - 1) | ((uint32_t)(dest - 1) << 16); }
                    } else if (TIME) t_hi += 300ull;                                  // :58-59
                }
                if (TIME) {
                    const uint32_t m_far = __ballot_sync(0xffffffffu, far);
                    if (far) fq[(fq_head + fq_cnt + __popc(m_far & lt_mask)) & (kCarQueue - 1)] = ((unsigned long long)odp << 12) | slot;
                    fq_cnt += __popc(m_far);
                    __syncwarp();
                    if (fq_cnt >= 32) { far_batch(32); fq_head += 32; fq_cnt -= 32; __syncwarp(); }
                }
            };

            for (int k0 = 0; k0 < spw; k0 += 32 * U) {
                uint16_t z[U];
                double pdz[U];
#pragma unroll
                for (int j = 0; j < U; ++j) z[j] = zs[wbase + k0 + j * 32 + lane];
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const bool parked = z[j] != 0 && z[j] != kNoCar;
                    pdz[j] = !parked ? -1.0 : (a.pd_smem ? pd_s[z[j] - 1] : __ldg(pd_t + (z[j] - 1)));   // -1: never drives
                }
                bool srch[U];
                unsigned long long uq[U];
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const int64_t car = car0 + k0 + j * 32 + lane;
                    bool drive;
                    uq[j] = 0ull;
                    if (INJ) drive = pdz[j] >= 0.0 && a.rng.u_drive[car + uoff] <= pdz[j];   // :15 (<=; NaN p_drive parks)
                    else {
                        uint32_t o[4];
                        philox_words(a.rng, car, (uint32_t)s, 0u, o);
                        drive = u52(o[0], o[1]) <= pdz[j];
                        uq[j] = ((unsigned long long)o[3] << 20) | (unsigned long long)(o[2] >> 12);
                    }
                    srch[j] = false;
                    if (drive) {
                        atomicAdd(&hdrv[z[j] - 1], 1u);                               // src/saveresults.jl:13-15
                        if (rz_s[z[j] - 1]) { if (TIME) t_hi += 300ull; }             // all-zero row: stays, 5 min (:35-36, :58-59)
                        else {
                            srch[j] = true;
                            if (!INJ) {   // the guide cell is needed once 32 drivers have queued up: start it towards L2 now
                                const uint16_t *gd = reinterpret_cast<const uint16_t *>(cdf_t + (int64_t)(z[j] - 1) * a.ldc + a.goff);
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(gd + (uint32_t)(uq[j] >> (52 - (32 - a.gshift)))));
                            }
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const uint32_t m = __ballot_sync(0xffffffffu, srch[j]);
                    if (srch[j]) dq[(dq_head + dq_cnt + __popc(m & lt_mask)) & (kCarQueue - 1)] = (uq[j] << 12) | (uint32_t)(k0 + j * 32 + lane);
                    dq_cnt += __popc(m);
                    __syncwarp();
                    if (dq_cnt >= 32) { drive_batch(32); dq_head += 32; dq_cnt -= 32; __syncwarp(); }
                }
            }
            if (dq_cnt > 0) { drive_batch(dq_cnt); __syncwarp(); }
            if (TIME && fq_cnt > 0) far_batch(fq_cnt);

            __syncthreads();
            // this step's rows: one reduction per non-empty bin
            uint32_t *const net_s = a.net + (size_t)s * NB, *const drv_s = a.drv + (size_t)s * NB;
            for (int x = tid; x < NB; x += THREADS) {
                const uint32_t v = hnet[x], d = hdrv[x];
                if (v) { atomicAdd(net_s + x, v); hnet[x] = 0u; }
                if (d) { atomicAdd(drv_s + x, d); hdrv[x] = 0u; }
            }
        }
        __syncthreads();
        for (int x = tid; x < n_cars; x += THREADS) a.zone[base + x] = (int32_t)zs[x];
    }

    // totals: integer block reduction, a few global atomics per CTA
    __syncthreads();
    t_hi = warp_sum_u64(t_hi); t_lo = warp_sum_u64(t_lo);
    if (lane == 0) {
        if (t_hi) atomicAdd(&red[0], t_hi);
        if (t_lo) atomicAdd(&red[1], t_lo);
    }
    __syncthreads();
    if (TIME && tid == 0) {
        const unsigned long long hi = red[0] + (red[1] >> 32), lo32 = red[1] & 0xffffffffull;
        if (hi) atomicAdd(a.t_hi, hi);
        if (lo32) atomicAdd(a.t_lo, lo32);
    }
}

// occupancy per step from the first step's occupancy and the net arrivals, in the public layout: parking / driving
// uint32 [N, n_steps]; the lost-car visits of the day are the occupancy of bin N summed over the steps
__global__ void __launch_bounds__(256) k_cars_export_counts(const uint32_t *__restrict__ cnt0, const uint32_t *__restrict__ net,
                                                            const uint32_t *__restrict__ drv, int N, int64_t n_steps,
                                                            uint32_t *__restrict__ parking, uint32_t *__restrict__ driving,
                                                            unsigned long long *lost)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int NB = N + 1;
    if (x >= NB) return;
    uint32_t run = cnt0[x];
    unsigned long long lost_visits = 0ull;
    for (int64_t s = 0; s < n_steps; ++s) {
        if (x < N) { parking[s * N + x] = run; driving[s * N + x] = drv[s * NB + x]; }
        else lost_visits += run;
        run += net[s * NB + x];   // two's complement: departures were subtracted
    }
    if (x == N) *lost = lost_visits;
}

// ------------------------------------------------------------------------------------------------
// geometry and launchers
// ------------------------------------------------------------------------------------------------
static int cars_env(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

bool cars_geometry(int64_t N, int64_t C, int with_time, CarsGeom *g)
{
    if (N < 1 || N + 1 > 65535 || C < 1) return false;
    static const int threads_env = cars_env("CPM_CARS_THREADS", 1024);
    const int threads = threads_env == 512 ? 512 : 1024;
    const int per_sm = 1024 / threads, warps = threads / 32;
    const size_t cap = (size_t)(232448 - 1024 * per_sm) / per_sm - 64;   // opt-in maximum per SM minus the per-CTA reserve
    const size_t queues = (size_t)warps * kCarQueue * 8 * (with_time ? 2 : 1);
    const size_t hist = 2 * (size_t)(N + 1) * 4 + (size_t)((N + 3) & ~(int64_t)3);
    const int64_t n_ctas = (int64_t)kSMs * per_sm;
    const int64_t want = ((C + n_ctas - 1) / n_ctas + threads * 4 - 1) / ((int64_t)threads * 4) * ((int64_t)threads * 4);
    for (int pd_smem = 1; pd_smem >= 0; --pd_smem) {
        const size_t fixed = queues + hist + (pd_smem ? (size_t)N * 8 : 0) + 32;
        if (fixed + (size_t)threads * 4 * 2 > cap) continue;
        int64_t cpc = (int64_t)((cap - fixed) / 2);
        cpc = cpc / (threads * 4) * (threads * 4);            // every warp owns a multiple of 32 * 4 slots
        if (cpc > want) cpc = want;
        if (cpc > (int64_t)4096 * warps) cpc = (int64_t)4096 * warps;   // 12-bit slot numbers inside a warp
        if (cpc < threads * 4) continue;
        g->threads = threads; g->pd_smem = pd_smem; g->cars_per_cta = (int)cpc;
        const int64_t need = (C + cpc - 1) / cpc;
        g->n_ctas = (int)(need < n_ctas ? need : n_ctas);
        g->smem = fixed + (size_t)cpc * 2;
        return true;
    }
    return false;
}

template <int THREADS, bool INJ, bool TIME>
static int day_cars_launch(const DayArgs &a, const CarsGeom &g, cudaStream_t st)
{
    // the opt-in applies to the current device's copy of the kernel: once per (device, instantiation)
    static std::mutex mu;
    static std::map<int, size_t> granted;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    {
        std::lock_guard<std::mutex> lk(mu);
        size_t &have = granted[dev];
        if (g.smem > have) {
            if (cudaFuncSetAttribute(k_day_cars<THREADS, INJ, TIME>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem) != cudaSuccess) return -1;
            have = g.smem;
        }
    }
    k_day_cars<THREADS, INJ, TIME><<<(unsigned)g.n_ctas, THREADS, g.smem, st>>>(a);
    return 1;
}

int launch_day_cars(const DayArgs &a, const CarsGeom &g, cudaStream_t st)
{
    const bool inj = a.rng.mode == 1, time = a.with_time != 0;
    if (g.threads == 512) {
        if (inj) return time ? day_cars_launch<512, true, true>(a, g, st) : day_cars_launch<512, true, false>(a, g, st);
        return time ? day_cars_launch<512, false, true>(a, g, st) : day_cars_launch<512, false, false>(a, g, st);
    }
    if (inj) return time ? day_cars_launch<1024, true, true>(a, g, st) : day_cars_launch<1024, true, false>(a, g, st);
    return time ? day_cars_launch<1024, false, true>(a, g, st) : day_cars_launch<1024, false, false>(a, g, st);
}

int launch_cars_export_counts(const uint32_t *cnt0, const uint32_t *net, const uint32_t *drv, int N, int64_t n_steps,
                              uint32_t *parking, uint32_t *driving, unsigned long long *lost, cudaStream_t st)
{
    k_cars_export_counts<<<(unsigned)((N + 1 + 255) / 256), 256, 0, st>>>(cnt0, net, drv, N, n_steps, parking, driving, lost);
    return 1;
}

}  // namespace cpm
