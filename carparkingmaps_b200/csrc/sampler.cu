// sampler.cu -- per-car, per-step HMM sampling (phases 1-3 + state update + zone histogram).
// Reference: src/resampling.jl:7-85, src/solveinitialvalueproblem.jl:8-55, src/saveresults.jl:6-28.
//
// "gather" sampler: one thread per car in car-id order, car state structure-of-arrays in HBM
// (int32 zone per car), counter-based Philox draws keyed by the GLOBAL car index so results do not
// depend on grid shape or on how cars are split across GPUs.  The categorical draw is a binary
// search of the origin's fp64 CDF row (sequential-sum CDF => same index as the reference's linear scan).
#include "kernels.h"

namespace cpm {

constexpr int kStepThreads = 256;
constexpr int kBlocksPerSM = 8;

int sampler_grid(int64_t C)
{
    int64_t b = (C + kStepThreads - 1) / kStepThreads;
    const int64_t cap = (int64_t)kSMs * kBlocksPerSM;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

__global__ void __launch_bounds__(kStepThreads) k_step_gather(const StepArgs a)
{
    const int64_t N = a.N, C = a.C;
    const double *__restrict__ pd = a.p_drive + N * a.hour;
    const double *__restrict__ cdf_t = a.cdf + a.hour * N * a.ldc;
    const uint8_t *__restrict__ rz = a.rowzero + a.hour * N;
    const int64_t NN = N * N;
    const bool inj = a.rng.mode == 1;
    const int64_t uoff = a.rng.ld * a.step;
    unsigned long long t_hi = 0ull, t_lo = 0ull;
    unsigned lost = 0;

    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < C; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t z = a.zone[i];
        int32_t dest = z;
        double drive = 0.0;
        double tt = 0.0, dd = 0.0;
        bool wrote_time = false;
        if (z < 1 || z > N) {
            ++lost;   // documented extension: the reference raises BoundsError here
        } else {
            double u1, u2;
            if (inj) u1 = a.rng.u_drive[i + uoff]; else philox_pair(a.rng, i, (uint32_t)a.step, 0u, u1, u2);
            if (a.parking) atomicAdd(a.parking + (z - 1), 1u);                    // saveresults.jl:12
            if (u1 <= __ldg(pd + (z - 1))) {                                      // resampling.jl:15  (<=; NaN parks)
                drive = 1.0;
                if (a.driving) atomicAdd(a.driving + (z - 1), 1u);                // saveresults.jl:13-15
                if (inj) u2 = a.rng.u_dest[i + uoff];
                if (!rz[z - 1]) {                                                 // resampling.jl:35-36
                    const double *__restrict__ row = cdf_t + (int64_t)(z - 1) * a.ldc;
                    if (a.G > 0) {
                        const uint16_t *__restrict__ gd = reinterpret_cast<const uint16_t *>(row + a.goff);
                        dest = cdf_search_guided([&](int32_t k) { return __ldg(row + k); },
                                                 [&](int32_t b) { return (int32_t)__ldg(gd + b); }, (int32_t)N, a.G, u2);   // :38-45
                    } else dest = cdf_search([&](int32_t k) { return __ldg(row + k); }, (int32_t)N, u2);
                }
                if (a.with_time) {                                                // resampling.jl:53-78
                    if (dest == z) { tt = 300.0; dd = 1.0; wrote_time = true; }   // :58-60
                    else if (dest >= 1) {
                        double u3, u4;
                        if (inj) { u3 = a.rng.u_time[i + uoff]; u4 = a.rng.u_dist[i + uoff]; }
                        else philox_pair(a.rng, i, (uint32_t)a.step, 1u, u3, u4);
                        const int64_t od = (int64_t)(z - 1) + N * (int64_t)(dest - 1);
                        double mean, sd;                                          // :63-64
                        if (a.msd) { const double2 ms = __ldg(a.msd + od + NN * a.hour); mean = ms.x; sd = ms.y; }
                        else { mean = __ldg(a.dm + od + NN * a.hour); sd = __ldg(a.dm + od + NN * (a.hour + a.T_tab)); }
                        if (sd == 0.0) sd = 0.1 * mean;                           // :65-67
                        tt = truncnormal(mean, sd, u3);                           // :68
                        mean = __ldg(a.dist + od);                                // :72
                        dd = truncnormal(mean, 0.1 * mean, u4);                   // :73-74
                        wrote_time = true;
                    }
                    fx_add(t_hi, t_lo, tt);
                }
            }
            a.zone[i] = dest;                                                     // :81-83
        }
        if (a.traj_trans) {
            const int64_t CT = C * a.T_traj;
            double *tr = a.traj_trans + i + C * a.t_traj;
            tr[0] = drive;                                                        // :21
            tr[CT] = (double)dest;                                                // :19 / :47
            if (wrote_time) { tr[2 * CT] = tt; tr[3 * CT] = dd; }                 // :59-60, :69, :75
            if (a.write_next_state) a.traj_state[i + C * (a.t_traj + 1)] = dest;  // :82
        }
    }

    // fixed-point driving-time total: integer block reduction, then two global atomics per CTA
    if (a.with_time) {
        __shared__ unsigned long long red[2];
        if (threadIdx.x < 2) red[threadIdx.x] = 0ull;
        __syncthreads();
        t_hi = warp_sum_u64(t_hi); t_lo = warp_sum_u64(t_lo);
        if ((threadIdx.x & 31) == 0) { if (t_hi) atomicAdd(&red[0], t_hi); if (t_lo) atomicAdd(&red[1], t_lo); }
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned long long hi = red[0] + (red[1] >> 32), lo32 = red[1] & 0xffffffffull;
            if (hi) atomicAdd(a.t_hi, hi);
            if (lo32) atomicAdd(a.t_lo, lo32);
        }
    }
    if (lost) atomicAdd(a.lost, (unsigned long long)lost);
}

int launch_step_gather(const StepArgs &a, cudaStream_t st)
{
    k_step_gather<<<sampler_grid(a.C), kStepThreads, 0, st>>>(a);
    return 1;
}

__global__ void k_reduce_partials(const double *__restrict__ partials, int n, double *__restrict__ out)
{
    // single thread, fixed order: n <= 148*8, negligible
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < n; ++k) s += partials[k];
        *out = s;
    }
}
int launch_reduce_partials(const double *partials, int n, double *out, cudaStream_t st)
{
    k_reduce_partials<<<1, 32, 0, st>>>(partials, n, out);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
__global__ void k_fleet_init(int32_t *__restrict__ zone, int64_t C, int64_t cpz, int64_t car_offset)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < C; i += (int64_t)gridDim.x * blockDim.x)
        zone[i] = (int32_t)((car_offset + i) / cpz + 1);   // ceil((i+1)/cpz), src/initializestates.jl:11-16
}
int launch_fleet_init(int32_t *zone, int64_t C, int64_t cpz, int64_t car_offset, cudaStream_t st)
{
    k_fleet_init<<<sampler_grid(C), kStepThreads, 0, st>>>(zone, C, cpz, car_offset);
    return 1;
}

__global__ void k_i64_to_i32(const int64_t *__restrict__ s, int32_t *__restrict__ d, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = s[i];
        d[i] = (v < 0 || v > 0x7fffffff) ? 0 : (int32_t)v;
    }
}
__global__ void k_i32_to_i64(const int32_t *__restrict__ s, int64_t *__restrict__ d, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        d[i] = s[i];
}
int launch_i64_to_i32(const int64_t *s, int32_t *d, int64_t n, cudaStream_t st)
{
    k_i64_to_i32<<<sampler_grid(n), kStepThreads, 0, st>>>(s, d, n);
    return 1;
}
int launch_i32_to_i64(const int32_t *s, int64_t *d, int64_t n, cudaStream_t st)
{
    k_i32_to_i64<<<sampler_grid(n), kStepThreads, 0, st>>>(s, d, n);
    return 1;
}

// initializestates (src/initializestates.jl:4-22) in Julia layout, on the device
__global__ void k_init_states(int64_t *__restrict__ state, double *__restrict__ trans, int64_t C, int64_t T, int64_t cpz)
{
    const int64_t n_state = C * T, n_trans = C * T * 4;
    const int64_t full = (C / cpz) * cpz;   // the reference's zone loop covers floor(C/cpz) whole zones (:11)
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_trans; k += (int64_t)gridDim.x * blockDim.x) {
        trans[k] = 0.0;
        if (k < n_state) state[k] = (k < full) ? (k / cpz + 1) : 0;   // k < C <=> column 1
        if (k < n_state && k >= C) state[k] = 0;
    }
}
int launch_init_states(int64_t *state, double *trans, int64_t C, int64_t T, int64_t cpz, cudaStream_t st)
{
    k_init_states<<<sampler_grid(C * T * 4), kStepThreads, 0, st>>>(state, trans, C, T, cpz);
    return 1;
}

// zone histogram of a Julia-layout trajectory (src/saveresults.jl:8-17)
__global__ void k_traj_histogram(const int64_t *__restrict__ state, const double *__restrict__ trans, int64_t C, int64_t N,
                                 int64_t T, uint32_t *__restrict__ parking, uint32_t *__restrict__ driving)
{
    const int64_t n = C * T;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = k / C;
        const int64_t z = state[k];
        if (z < 1 || z > N) continue;
        atomicAdd(parking + (z - 1) + N * t, 1u);
        if (trans[k] == 1.0) atomicAdd(driving + (z - 1) + N * t, 1u);
    }
}
int launch_traj_histogram(const int64_t *state, const double *trans, int64_t C, int64_t N, int64_t T, uint32_t *parking,
                          uint32_t *driving, cudaStream_t st)
{
    k_traj_histogram<<<sampler_grid(C * T), kStepThreads, 0, st>>>(state, trans, C, N, T, parking, driving);
    return 1;
}

// deterministic two-stage fp64 sum
__global__ void __launch_bounds__(256) k_sum_stage1(const double *__restrict__ x, int64_t n, double *__restrict__ partials)
{
    double s = 0.0;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) s += x[k];
    __shared__ double sp[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sp[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = 0.0;
        for (int w = 0; w < 8; ++w) r += sp[w];
        partials[blockIdx.x] = r;
    }
}
int launch_sum_f64(const double *x, int64_t n, double *partials, int n_partials, double *out, cudaStream_t st)
{
    int g = sampler_grid(n);
    if (g > n_partials) g = n_partials;
    k_sum_stage1<<<g, 256, 0, st>>>(x, n, partials);
    k_reduce_partials<<<1, 32, 0, st>>>(partials, g, out);
    return 2;
}

// parking = counts / C (saveresults.jl:20); traffic = minmax over hours of column sums (saveresults.jl:23-28).
// One block per hour: coalesced pass over the hour's column; then one tiny block for the min-max over hours.
__global__ void __launch_bounds__(256) k_normalise_cols(const uint32_t *__restrict__ parking, const uint32_t *__restrict__ driving,
                                                        int64_t N, int64_t C, double *__restrict__ parking_out,
                                                        double *__restrict__ driving_out, unsigned long long *__restrict__ col_sum)
{
    const int64_t t = blockIdx.x;
    unsigned long long s = 0ull;
    for (int64_t z = threadIdx.x; z < N; z += blockDim.x) {
        const int64_t k = z + N * t;
        if (parking_out) parking_out[k] = (double)parking[k] / (double)C;
        const uint32_t d = driving[k];
        if (driving_out) driving_out[k] = (double)d;
        s += d;
    }
    __shared__ unsigned long long sp[8];
    s = warp_sum_u64(s);
    if ((threadIdx.x & 31) == 0) sp[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long r = 0ull;
        for (int w = 0; w < 8; ++w) r += sp[w];
        col_sum[t] = r;
    }
}
__global__ void k_normalise_traffic(const unsigned long long *__restrict__ col_sum, int64_t T, double *__restrict__ traffic_out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long mn = col_sum[0], mx = col_sum[0];
        for (int64_t t = 0; t < T; ++t) { if (col_sum[t] < mn) mn = col_sum[t]; if (col_sum[t] > mx) mx = col_sum[t]; }
        for (int64_t t = 0; t < T; ++t) traffic_out[t] = ((double)col_sum[t] - (double)mn) / ((double)mx - (double)mn);
    }
}
// traffic_counts_out [T] is required (it doubles as the column-sum scratch)
int launch_normalise_counts(const uint32_t *parking, const uint32_t *driving, int64_t N, int64_t T, int64_t C,
                            double *parking_out, double *driving_out, double *traffic_out,
                            unsigned long long *traffic_counts_out, cudaStream_t st)
{
    k_normalise_cols<<<(unsigned)T, 256, 0, st>>>(parking, driving, N, C, parking_out, driving_out, traffic_counts_out);
    if (!traffic_out) return 1;
    k_normalise_traffic<<<1, 32, 0, st>>>(traffic_counts_out, T, traffic_out);
    return 2;
}

}  // namespace cpm
