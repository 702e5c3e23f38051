// synth.cu -- synthetic city generator on the device: the Uber Movement / Melbourne data
// the reference consumes are not available offline, so benchmarks run on synthetic travel-time
// matrices of the same shape, layout and value ranges (README.md:259, :310 of the reference).
// Counter-hash based and built only from IEEE +,-,*,/,sqrt so that carparkingmaps_b200/synth.py
// (numpy) produces bit-identical arrays on the host.
#include "kernels.h"

namespace cpm {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ double hashu(uint64_t seed, uint64_t tag, uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t h = mix64(seed * 0x100000001B3ull + tag);
    h = mix64(h ^ a);
    h = mix64(h ^ b);
    h = mix64(h ^ c);
    return (double)(h >> 11) * 0x1.0p-53;
}

__device__ __forceinline__ double bump(double x)
{
    const double q = x / 2.5;
    const double r = 1.0 - q * q;
    return r > 0.0 ? r * r : 0.0;
}

// one thread per (i,j), i fastest; loops over t
__global__ void __launch_bounds__(256) k_synth_city(double *__restrict__ dm, double *__restrict__ dist, int64_t N, int64_t T,
                                                    uint64_t seed, double density)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t NN = N * N;
    if (idx >= NN) return;
    const uint64_t i = (uint64_t)(idx % N), j = (uint64_t)(idx / N);
    double d;
    if (i == j) d = 1.0;   // src/processgeodata.jl:164-166: diagonal forced to 1
    else {
        const double dx = 40.0 * hashu(seed, 1, i, 0, 0) - 40.0 * hashu(seed, 1, j, 0, 0);
        const double dy = 40.0 * hashu(seed, 2, i, 0, 0) - 40.0 * hashu(seed, 2, j, 0, 0);
        d = sqrt(dx * dx + dy * dy);
        if (d < 0.05) d = 0.05;
    }
    dist[idx] = d;
    const double v0 = 30.0 + 30.0 * hashu(seed, 3, i, j, 0);
    const double amp = 0.3 + 0.9 * hashu(seed, 4, i, j, 0);
    const double phase = -1.0 + 2.0 * hashu(seed, 5, i, j, 0);
    // free-flow time, clipped so that mean = base * c stays inside the 60..3600 s range of the real data
    // (c < 2.4) without flattening any pair to a constant (a constant pair makes NaN table entries)
    double base = (d * 3600.0) / v0;
    if (base < 60.0) base = 60.0;
    if (base > 1500.0) base = 1500.0;
    for (int64_t t = 0; t < T; ++t) {
        double mean = 0.0, sd = 0.0;
        const bool present = (i != j) && (density >= 1.0 || hashu(seed, 8, i, j, (uint64_t)t) < density);
        if (present) {
            const double h = (double)((t % 24) + 1);
            const double c = 1.0 + amp * (bump(h - 8.0 - phase) + bump(h - 17.5 - phase)) +
                             0.15 * hashu(seed, 9, i, j, (uint64_t)t);
            mean = base * c;
            sd = mean * (0.05 + 0.55 * hashu(seed, 6, i, j, (uint64_t)t));
            if (hashu(seed, 7, i, j, (uint64_t)t) < 0.02) sd = 0.0;   // exercises src/resampling.jl:65-67
        }
        dm[idx + NN * t] = mean;
        dm[idx + NN * (t + T)] = sd;
    }
}

// (mean, std) of every (origin, destination, hour) side by side: one 16-byte gather per sampled trip instead of two
// 8-byte gathers from slabs 1.1 GB apart.  Worth its 4.4 GB of build traffic only when many days are sampled on the
// same inputs (the model-selection sweep).
__global__ void __launch_bounds__(256) k_pair_table(const double *__restrict__ dm, int64_t n, double2 *__restrict__ msd)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) msd[k] = make_double2(__ldg(dm + k), __ldg(dm + k + n));
}
int launch_pair_table(const double *dm, int64_t N, int64_t T, double *msd, cudaStream_t st)
{
    const int64_t n = N * N * T;
    k_pair_table<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dm, n, reinterpret_cast<double2 *>(msd));
    return 1;
}

int launch_synth_city(double *dm, double *dist, int64_t N, int64_t T, uint64_t seed, double density, cudaStream_t st)
{
    const int64_t NN = N * N;
    k_synth_city<<<(unsigned)((NN + 255) / 256), 256, 0, st>>>(dm, dist, N, T, seed, density);
    return 1;
}

}  // namespace cpm
