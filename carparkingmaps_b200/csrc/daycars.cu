// daycars.cu -- the car-resident day kernel: ONE launch runs every step of a day (src/resampling.jl:11-83,
// src/solveinitialvalueproblem.jl:8-55, src/saveresults.jl:8-17).
//
// Cars never interact: a car's chain z(0) -> z(1) -> ... depends on the tables and on its own draws only (Philox keyed by
// global car id and step), and the per-zone counts are sums.  So a CTA takes a contiguous block of cars, keeps their
// zones in shared memory (16 bits per car) for the whole day and walks the steps on its own -- no sort, no exchange, no
// kernel boundary between steps.  Per step a warp
//   1. decides for each of its cars whether it drives (one Philox block against the hour's p_drive column),
//   2. queues the drivers and serves them in dense batches of 32: guide cell -> a window of a few CDF entries of the
//      origin's row (global memory) -> destination,
//   3. queues the cars that changed zone and draws their travel times in dense batches (resampling day only).
// Counts: per-CTA shared-memory histograms (net arrivals and drivers per zone), flushed once per step with one reduction
// per non-empty bin into that step's global row; occupancy at step s+1 = occupancy at step s + net arrivals (export).
// Same draws as every other sampler => identical results (tests compare with the oracle and with the bucket sampler).
#include <math_constants.h>
#include <map>
#include <mutex>
#include "kernels.h"

namespace cpm {

constexpr int kCarQueue = 64;          // entries per warp queue (a batch is served as soon as 32 have queued up)
constexpr uint16_t kNoCar = 0xFFFFu;   // slot beyond the CTA's last car
constexpr int kCarThreads = 1024;      // one CTA per SM
constexpr int kWindow = 8;             // CDF entries per window trip (4: +2 % per step, one more trip for half of the batches)

template <int THREADS, bool INJ, bool TIME, int U, bool TRAJ>
__global__ void __launch_bounds__(THREADS, THREADS == 1024 ? 1 : 2) k_day_cars(const DayArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long red[3];
    constexpr int WARPS = THREADS / 32;
    constexpr int W = kWindow;
    constexpr int E = 1;                 // queue entries per lane and batch (2 measured slower: larger partial batches, 64 registers)
    const int64_t N = a.N, C = a.C, NN = a.N * a.N;
    const int Ni = (int)N, NB = Ni + 1;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    // shared memory: drive queues | far queues | net arrivals | drivers | rowzero column | zones
    unsigned long long *const dq = reinterpret_cast<unsigned long long *>(smem_raw) + wid * kCarQueue;
    unsigned long long *const fq = reinterpret_cast<unsigned long long *>(smem_raw) + (WARPS + wid) * kCarQueue;
    // (the per-zone tables only exist when the day keeps per-step rows: a.counts)
    uint32_t *const hnet = reinterpret_cast<uint32_t *>(smem_raw) + 2 * (TIME ? 2 : 1) * WARPS * kCarQueue;
    uint32_t *const hdrv = hnet + (a.counts ? NB : 0);   // two 16-bit counters per word: a CTA holds fewer than 65,536 cars
    const int NW = (NB + 1) >> 1;                        // words of hdrv
    uint8_t *const rz_s = reinterpret_cast<uint8_t *>(hdrv + (a.counts ? NW : 0));
    uint16_t *const zs = reinterpret_cast<uint16_t *>(rz_s + (a.counts ? ((N + 3) & ~(int64_t)3) : 0));
    const int cpc = a.cars_per_cta, spw = cpc / WARPS;   // slots per CTA / per warp (spw: a multiple of 32)
    const int wbase = wid * spw;
    // With per-step rows the warps of a CTA meet at every step anyway, so the hour's rowzero column is staged in shared
    // memory; without (the IVP of a dataset pass) the warps run through the day without a single block barrier.  The
    // p_drive column is read through L1 either way (staged: +1 % per step at cfg2, and at 10^4 zones its 80 KB halve the
    // block of cars a CTA can hold: +35 % per step).
    const bool staged = a.counts != 0;
    unsigned long long t_hi = 0ull, t_lo = 0ull, lost = 0ull;
    if (tid < 3) red[tid] = 0ull;

    for (int64_t base = (int64_t)blockIdx.x * cpc; base < C; base += (int64_t)gridDim.x * cpc) {
        const int n_cars = (int)((C - base < cpc) ? (C - base) : cpc);
        const int64_t car0 = base + wbase;   // the warp's first car
        __syncthreads();
        if (a.counts) for (int x = tid; x < NB + NW; x += THREADS) hnet[x] = 0u;   // hnet and hdrv are adjacent
        __syncthreads();
        // the CTA's cars: zones into shared memory, and their share of the occupancy at the first step
        for (int x = tid; x < cpc; x += THREADS) {
            uint16_t z16 = kNoCar;
            if (x < n_cars) {
                const int32_t z = a.zone[base + x];
                z16 = (z >= 1 && z <= Ni) ? (uint16_t)z : (uint16_t)0;   // anything else is a lost car (documented extension)
                if (a.cnt0) atomicAdd(&hnet[z16 ? z16 - 1 : Ni], 1u);   // (null without counts)
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
            const uint8_t *__restrict__ rz_t = a.rowzero + hour * N;
            if (staged) {
                if (s == a.step0 || s % a.steps_per_hour == 0) {   // a new hour: its rowzero column
                    __syncthreads();
                    for (int x = tid; x < Ni; x += THREADS) rz_s[x] = rz_t[x];
                }
                __syncthreads();   // column loaded; the previous step's flush is done with the histograms
            }

            uint32_t dq_head = 0, fq_head = 0;
            int dq_cnt = 0, fq_cnt = 0;
            // travel times of up to 32 * E queued cars (src/resampling.jl:61-70), E per lane: the E table gathers of a lane
            // are in flight together
            auto far_batch = [&](int n) {
                double mean[E], sd[E], dkm[E];
                int64_t car[E];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    mean[e] = sd[e] = dkm[e] = 0.0; car[e] = 0;
                    if (lane + 32 * e < n) {
                        const unsigned long long q = fq[(fq_head + lane + 32 * e) & (kCarQueue - 1)];
                        car[e] = car0 + (int64_t)(q & 0xFFFull);
                        const uint32_t odp = (uint32_t)(q >> 12);
                        const int64_t od = (int64_t)(odp & 0xffffu) + N * (int64_t)(odp >> 16);
                        if (a.msd) { const double2 ms = __ldg(a.msd + od + NN * hour); mean[e] = ms.x; sd[e] = ms.y; }
                        else { mean[e] = __ldg(a.dm + od + NN * hour); sd[e] = __ldg(a.dm + od + NN * (hour + a.T_tab)); }
                        if (TRAJ) dkm[e] = __ldg(a.dist + od);                        // :72
                    }
                }
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    if (lane + 32 * e < n) {
                        double u3, u4 = 0.0;
                        if (INJ) { u3 = a.rng.u_time[car[e] + uoff]; if (TRAJ) u4 = a.rng.u_dist[car[e] + uoff]; }
                        else { uint32_t o[4]; philox_words(a.rng, car[e], (uint32_t)s, 1u, o); u3 = u52(o[0], o[1]); if (TRAJ) u4 = u52(o[2], o[3]); }
                        if (sd[e] == 0.0) sd[e] = 0.1 * mean[e];                      // :65-67
                        const double tt = truncnormal(mean[e], sd[e], u3);            // :68
                        fx_add(t_hi, t_lo, tt);
                        if (TRAJ) {                                                   // :69, :73-75
                            double *tr = a.traj_trans + car[e] + C * s;
                            tr[2 * C * a.T_traj] = tt;
                            tr[3 * C * a.T_traj] = truncnormal(dkm[e], 0.1 * dkm[e], u4);
                        }
                    }
                }
            };
            // categorical destinations of up to 32 * E queued drivers (src/resampling.jl:26-49), E per lane, in phases so
            // that the E guide lookups, then the E row windows, of a lane are in flight together
            auto drive_batch = [&](int n) {
                uint32_t slot[E];
                int z[E];
                int32_t lo[E], last[E], dest[E];
                double u2[E];
                const double *row[E];
                bool act[E];
#pragma unroll
                for (int e = 0; e < E; ++e) {   // queue entry, origin, guide cell
                    act[e] = lane + 32 * e < n;
                    slot[e] = 0; z[e] = 1; lo[e] = 1; last[e] = 0; dest[e] = 0; u2[e] = 0.0; row[e] = cdf_t;
                    if (act[e]) {
                        const unsigned long long q = dq[(dq_head + lane + 32 * e) & (kCarQueue - 1)];
                        slot[e] = (uint32_t)(q & 0xFFFull);
                        const unsigned long long uq = q >> 12;   // the 52-bit integer of u_dest (Philox)
                        z[e] = (int)zs[wbase + slot[e]];         // 1-based origin
                        row[e] = cdf_t + (int64_t)(z[e] - 1) * a.ldc;
                        const uint16_t *__restrict__ gd = reinterpret_cast<const uint16_t *>(row[e] + a.goff);
                        u2[e] = INJ ? a.rng.u_dest[car0 + slot[e] + uoff]
                                    : __hiloint2double((int)((uint32_t)(uq >> 32) | 0x3FF00000u), (int)(uint32_t)uq) - 1.0;
                        if (u2[e] > 0.0) {
                            int32_t hi;
                            if (INJ && !(u2[e] < 1.0)) { lo[e] = __ldg(gd + a.G); hi = Ni; }
                            else {
                                const uint32_t c = INJ ? (uint32_t)(u2[e] * (double)a.G) : (uint32_t)(uq >> (52 - (32 - a.gshift)));
                                lo[e] = __ldg(gd + c); hi = __ldg(gd + c + 1);
                            }
                            last[e] = hi < Ni ? hi : Ni - 1;
                        }
                    }
                }
                // first index in [lo, last] whose CDF value is not below u2 (NaN counts as not below, like the binary search
                // of the other samplers: the row is non-decreasing with a NaN tail, so the predicate is monotone and the
                // answer is the same index).  The guide leaves a window of a few entries: eight independent loads per trip
                // and entry instead of a chain of dependent probes.
                bool more;
                do {
                    double r[E][W];
#pragma unroll
                    for (int e = 0; e < E; ++e)
#pragma unroll
                        for (int q = 0; q < W; ++q) r[e][q] = (lo[e] + q <= last[e]) ? __ldg(row[e] + lo[e] + q) : CUDART_INF;
                    more = false;
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        if (lo[e] <= last[e]) {
                            int q = W;
                            double v = 0.0;
#pragma unroll
                            for (int k = W - 1; k >= 0; --k) if (!(r[e][k] < u2[e])) { q = k; v = r[e][k]; }
                            if (q < W) {
                                if (lo[e] + q <= last[e] && u2[e] <= v) dest[e] = lo[e] + q + 1;   // else 0: lost (:32,40)
                                lo[e] = last[e] + 1;   // done
                            } else {
                                lo[e] += W;
                                more |= lo[e] <= last[e];
                            }
                        }
                    }
                } while (__any_sync(0xffffffffu, more));
                bool far[E];
                uint32_t odp[E];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    far[e] = false; odp[e] = 0;
                    if (TRAJ && act[e]) {   // Julia-layout trajectory of a car that drew a destination (src/resampling.jl:19-21,47,58-60,82)
                        const int64_t car = car0 + slot[e];
                        double *tr = a.traj_trans + car + C * s;
                        tr[0] = 1.0;
                        tr[C * a.T_traj] = (double)dest[e];
                        if (TIME && dest[e] == z[e]) { tr[2 * C * a.T_traj] = 300.0; tr[3 * C * a.T_traj] = 1.0; }
                        if (s + 1 < a.T_traj) a.traj_state[car + C * (s + 1)] = dest[e];
                    }
                    if (act[e]) {
                        if (dest[e] != z[e]) {
                            zs[wbase + slot[e]] = (uint16_t)dest[e];                      // :81-83
                            if (a.counts) {
                                atomicSub(&hnet[z[e] - 1], 1u);
                                atomicAdd(&hnet[dest[e] >= 1 ? dest[e] - 1 : Ni], 1u);
                            }
                            if (TIME && dest[e] >= 1) { far[e] = true; odp[e] = (uint32_t)(z[e] - 1) | ((uint32_t)(dest[e] - 1) << 16); }
                        } else if (TIME) t_hi += 300ull;                                  // :58-59
                    }
                }
                if (TIME) {
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const uint32_t m_far = __ballot_sync(0xffffffffu, far[e]);
                        if (far[e]) fq[(fq_head + fq_cnt + __popc(m_far & lt_mask)) & (kCarQueue - 1)] = ((unsigned long long)odp[e] << 12) | slot[e];
                        fq_cnt += __popc(m_far);
                    }
                    __syncwarp();
                    if (fq_cnt >= 32 * E) { far_batch(32 * E); fq_head += 32 * E; fq_cnt -= 32 * E; __syncwarp(); }
                }
            };

            for (int k0 = 0; k0 < spw; k0 += 32 * U) {
                uint16_t z[U];
                double pdz[U];
#pragma unroll
                for (int j = 0; j < U; ++j) z[j] = (k0 + j * 32 < spw) ? zs[wbase + k0 + j * 32 + lane] : kNoCar;
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const bool parked = z[j] != 0 && z[j] != kNoCar;
                    if (!a.counts) lost += z[j] == 0;   // without the per-step rows the lost-car visits are counted here
                    pdz[j] = !parked ? -1.0 : __ldg(pd_t + (z[j] - 1));   // -1: never drives
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
                        if (a.counts) atomicAdd(&hdrv[(z[j] - 1) >> 1], 1u << (((z[j] - 1) & 1) * 16));   // src/saveresults.jl:13-15
                        if (staged ? rz_s[z[j] - 1] : __ldg(rz_t + (z[j] - 1))) { if (TIME) t_hi += 300ull; }             // all-zero row: stays, 5 min (:35-36, :58-59)
                        else srch[j] = true;
                    }
                    if (TRAJ && z[j] != kNoCar && !srch[j]) {   // parked, lost, or driving inside an all-zero row: the car stays
                        double *tr = a.traj_trans + car + C * s;
                        tr[0] = drive ? 1.0 : 0.0;
                        tr[C * a.T_traj] = (double)z[j];
                        if (TIME && drive) { tr[2 * C * a.T_traj] = 300.0; tr[3 * C * a.T_traj] = 1.0; }
                        if (s + 1 < a.T_traj) a.traj_state[car + C * (s + 1)] = z[j];
                    }
                }
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const uint32_t m = __ballot_sync(0xffffffffu, srch[j]);
                    if (srch[j]) dq[(dq_head + dq_cnt + __popc(m & lt_mask)) & (kCarQueue - 1)] = (uq[j] << 12) | (uint32_t)(k0 + j * 32 + lane);
                    dq_cnt += __popc(m);
                    __syncwarp();
                    if (dq_cnt >= 32 * E) { drive_batch(32 * E); dq_head += 32 * E; dq_cnt -= 32 * E; __syncwarp(); }
                }
            }
            if (dq_cnt > 0) { drive_batch(dq_cnt); __syncwarp(); }
            if (TIME) while (fq_cnt > 0) { const int n = fq_cnt < 32 * E ? fq_cnt : 32 * E; far_batch(n); fq_head += n; fq_cnt -= n; }

            if (a.counts) {
                __syncthreads();
                // this step's rows: one reduction per non-empty bin
                uint32_t *const net_s = a.net + (size_t)s * NB, *const drv_s = a.drv + (size_t)s * NB;
                for (int x = tid; x < NB; x += THREADS) {
                    const uint32_t v = hnet[x];
                    if (v) { atomicAdd(net_s + x, v); hnet[x] = 0u; }
                }
                for (int x = tid; x < NW; x += THREADS) {
                    const uint32_t d = hdrv[x];
                    if (d & 0xffffu) atomicAdd(drv_s + 2 * x, d & 0xffffu);
                    if (d >> 16) atomicAdd(drv_s + 2 * x + 1, d >> 16);
                    if (d) hdrv[x] = 0u;
                }
            }
        }
        __syncthreads();
        for (int x = tid; x < n_cars; x += THREADS) a.zone[base + x] = (int32_t)zs[x];
    }

    // totals: integer block reduction, a few global atomics per CTA
    __syncthreads();
    t_hi = warp_sum_u64(t_hi); t_lo = warp_sum_u64(t_lo);
    if (!a.counts) lost = warp_sum_u64(lost);
    if (lane == 0) {
        if (t_hi) atomicAdd(&red[0], t_hi);
        if (t_lo) atomicAdd(&red[1], t_lo);
        if (lost) atomicAdd(&red[2], lost);
    }
    __syncthreads();
    if (tid == 0) {
        if (TIME) {
            const unsigned long long hi = red[0] + (red[1] >> 32), lo32 = red[1] & 0xffffffffull;
            if (hi) atomicAdd(a.t_hi, hi);
            if (lo32) atomicAdd(a.t_lo, lo32);
        }
        if (!a.counts && red[2]) atomicAdd(a.lost_visits, red[2]);
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
bool cars_geometry(int64_t N, int64_t C, int with_time, int counts, CarsGeom *g)
{
    if (N < 1 || N + 1 > 65535 || C < 1) return false;
    const int threads = kCarThreads, per_sm = 1, warps = threads / 32;
    const size_t cap = (size_t)(232448 - 1024 * per_sm) / per_sm - 64;   // opt-in maximum per SM minus the per-CTA reserve
    const size_t queues = (size_t)warps * kCarQueue * 8 * (with_time ? 2 : 1);
    const size_t hist = counts ? (size_t)(N + 1) * 4 + (size_t)((N + 2) / 2) * 4 + (size_t)((N + 3) & ~(int64_t)3) : 0;   // hnet, hdrv (16-bit), rowzero
    const int64_t n_ctas = (int64_t)kSMs * per_sm;
    const int64_t want = ((C + n_ctas - 1) / n_ctas + threads - 1) / threads * threads;   // every warp owns a multiple of 32 slots
    const size_t fixed = queues + hist + 32;
    if (fixed + (size_t)threads * 2 > cap) return false;
    int64_t cpc = (int64_t)((cap - fixed) / 2);
    cpc = cpc / threads * threads;
    if (cpc > want) cpc = want;
    if (cpc > (int64_t)4096 * warps) cpc = (int64_t)4096 * warps;   // 12-bit slot numbers inside a warp
    if (cpc > 65535 / threads * threads) cpc = 65535 / threads * threads;   // 16-bit driver counters per CTA and bin
    if (cpc < want) {   // several blocks of cars per CTA: size them so that every CTA walks the same number
        const int64_t waves = (C + cpc * n_ctas - 1) / (cpc * n_ctas);
        const int64_t even = ((C + waves * n_ctas - 1) / (waves * n_ctas) + threads - 1) / threads * threads;
        if (even < cpc) cpc = even;
    }
    g->threads = threads; g->cars_per_cta = (int)cpc;
    const int64_t need = (C + cpc - 1) / cpc;
    g->n_ctas = (int)(need < n_ctas ? need : n_ctas);
    g->smem = fixed + (size_t)cpc * 2;
    return true;
}

template <int THREADS, bool INJ, bool TIME, int U, bool TRAJ>
static int day_cars_launch_u(const DayArgs &a, const CarsGeom &g, cudaStream_t st)
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
            if (cudaFuncSetAttribute(k_day_cars<THREADS, INJ, TIME, U, TRAJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem) != cudaSuccess) return -1;
            have = g.smem;
        }
    }
    k_day_cars<THREADS, INJ, TIME, U, TRAJ><<<(unsigned)g.n_ctas, THREADS, g.smem, st>>>(a);
    return 1;
}

int launch_day_cars(const DayArgs &a, const CarsGeom &g, cudaStream_t st)
{
    // one CTA of 1,024 threads per SM, two cars per lane and trip.  Measured at cfg2 (profiles/README.md): 2 x 512 threads
    // per SM 45 % slower per step, 2 x 768 for the IVP day (40 registers, spills) 17 % slower, four cars per lane 11 %
    // slower, one car per lane 2 % slower.
    constexpr int T = kCarThreads, U = 2;
    const bool inj = a.rng.mode == 1, time = a.with_time != 0;
    if (a.traj_trans) {   // Julia-layout trajectories (the reference-named host functions)
        if (inj) return time ? day_cars_launch_u<T, true, true, U, true>(a, g, st) : day_cars_launch_u<T, true, false, U, true>(a, g, st);
        return time ? day_cars_launch_u<T, false, true, U, true>(a, g, st) : day_cars_launch_u<T, false, false, U, true>(a, g, st);
    }
    if (inj) return time ? day_cars_launch_u<T, true, true, U, false>(a, g, st) : day_cars_launch_u<T, true, false, U, false>(a, g, st);
    return time ? day_cars_launch_u<T, false, true, U, false>(a, g, st) : day_cars_launch_u<T, false, false, U, false>(a, g, st);
}

int launch_cars_export_counts(const uint32_t *cnt0, const uint32_t *net, const uint32_t *drv, int N, int64_t n_steps,
                              uint32_t *parking, uint32_t *driving, unsigned long long *lost, cudaStream_t st)
{
    k_cars_export_counts<<<(unsigned)((N + 1 + 255) / 256), 256, 0, st>>>(cnt0, net, drv, N, n_steps, parking, driving, lost);
    return 1;
}

}  // namespace cpm
