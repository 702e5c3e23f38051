// tables.cu -- probability tables of the HMM: createpdrive / createpdestin and the sampling CDF.
// Reference: src/createpdrive.jl:3-38, src/createpdestin.jl:3-50 (Julia, column-major, origin index fastest).
// All kernels are HBM-bound elementwise / scan work in fp64; no tensor cores (nothing is a contraction).
#include "kernels.h"

namespace cpm {

// ------------------------------------------------------------------------------------------------
// createpdrive, part 1: mean_sum[i,t] = (sum_{j: m!=0} m[i,j,t]/dist[i,j]) / count   (:10-21)
// One thread per (i,t), i fastest -> every load of the j loop is coalesced across the warp.  The
// running sum is accumulated in ascending j exactly like the reference, so mean_sum is bit-identical
// to a sequential fp64 evaluation (fp64 add/div are IEEE on the device).
// ------------------------------------------------------------------------------------------------
constexpr int kMsUnroll = 8;

__global__ void __launch_bounds__(64) k_pdrive_meansum(const double *__restrict__ dm, const double *__restrict__ dist,
                                                       int64_t N, int64_t T, double *__restrict__ mean_sum)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * T) return;
    const int64_t i = idx % N, t = idx / N;
    const double *m = dm + i + N * N * t;
    const double *d = dist + i;
    double s = 0.0;
    int64_t cnt = 0;
    int64_t j = 0;
    for (; j + kMsUnroll <= N; j += kMsUnroll) {
        double mv[kMsUnroll], dv[kMsUnroll];
#pragma unroll
        for (int u = 0; u < kMsUnroll; ++u) { mv[u] = __ldg(m + N * (j + u)); dv[u] = __ldg(d + N * (j + u)); }
#pragma unroll
        for (int u = 0; u < kMsUnroll; ++u)
            if (mv[u] != 0.0) { s = s + mv[u] / dv[u]; ++cnt; }
    }
    for (; j < N; ++j) {
        const double mv = __ldg(m + N * j), dv = __ldg(d + N * j);
        if (mv != 0.0) { s = s + mv / dv; ++cnt; }
    }
    mean_sum[idx] = s / (double)cnt;   // 0/0 -> NaN, as in Julia (:20)
}

// createpdrive, part 2 (:22-34): NaN-propagating max/min over t, then the affine map.  One thread per zone.
__global__ void k_pdrive_finish(const double *__restrict__ mean_sum, int64_t N, int64_t T, double p_min, double p_max,
                                double e_drive, double *__restrict__ p_drive)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double mx = mean_sum[i], mn = mx;
    bool nan = false;
    for (int64_t t = 0; t < T; ++t) {
        const double v = mean_sum[i + N * t];
        nan |= isnan(v);
        if (v > mx) mx = v;
        if (v < mn) mn = v;
    }
    if (nan) mx = mn = __longlong_as_double(0x7ff8000000000000LL);
    const bool ok = mx > 0.0;   // NaN -> false: the row stays zero (:28)
    for (int64_t t = 0; t < T; ++t) {
        double p = 0.0;
        if (ok) p = p_min + (p_max - p_min) * jl_pow((mean_sum[i + N * t] - mn) / (mx - mn), e_drive);   // :31
        p_drive[i + N * t] = p;
    }
}

int launch_pdrive_meansum(const double *dm, const double *dist, int64_t N, int64_t T, double *mean_sum, cudaStream_t st)
{
    const int64_t n = N * T;
    k_pdrive_meansum<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(dm, dist, N, T, mean_sum);
    return 1;
}

int launch_pdrive_finish(const double *mean_sum, int64_t N, int64_t T, double p_min, double p_max, double e_drive,
                         double *p_drive, cudaStream_t st)
{
    k_pdrive_finish<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(mean_sum, N, T, p_min, p_max, e_drive, p_drive);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// createpdestin, pass 1 (:10-28): per (i,j) min/max over t (zeros included), then the unnormalised
// ((m-min)/(max-min))^e_dest.  One thread per (i,j), i fastest: coalesced.  TMAX > 0 keeps the T
// values in registers (T = 24 in the reference); TMAX == 0 re-reads them.
// ------------------------------------------------------------------------------------------------
template <int TMAX>
__global__ void __launch_bounds__(256) k_pdest_unnorm(const double *__restrict__ dm, int64_t N, int64_t T, double e_dest,
                                                      double *__restrict__ p_dest)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t NN = N * N;
    if (idx >= NN) return;
    double v[TMAX > 0 ? TMAX : 1];
    double mx, mn;
    bool has_nan = false;
    if constexpr (TMAX > 0) {
#pragma unroll
        for (int t = 0; t < TMAX; ++t) v[t] = (t < T) ? __ldg(dm + idx + NN * t) : 0.0;
        mx = mn = v[0];
#pragma unroll
        for (int t = 0; t < TMAX; ++t)
            if (t < T) { has_nan |= isnan(v[t]); if (v[t] > mx) mx = v[t]; if (v[t] < mn) mn = v[t]; }
    } else {
        mx = mn = __ldg(dm + idx);
        for (int64_t t = 0; t < T; ++t) {
            const double x = __ldg(dm + idx + NN * t);
            has_nan |= isnan(x); if (x > mx) mx = x; if (x < mn) mn = x;
        }
    }
    if (has_nan) mx = mn = __longlong_as_double(0x7ff8000000000000LL);
    const bool ok = mx > 0.0;   // :22
    const double den = mx - mn;
    if constexpr (TMAX > 0) {
#pragma unroll
        for (int t = 0; t < TMAX; ++t)
            if (t < T) p_dest[idx + NN * t] = ok ? jl_pow((v[t] - mn) / den, e_dest) : 0.0;   // :24
    } else {
        for (int64_t t = 0; t < T; ++t)
            p_dest[idx + NN * t] = ok ? jl_pow((__ldg(dm + idx + NN * t) - mn) / den, e_dest) : 0.0;
    }
}

// Julia's sum(::Vector{Float64}): pairwise, 1024-element sequential base case.  `stride` = N (elements of
// one origin's row are N apart in the column-major table).
__device__ double jl_sum_strided(const double *__restrict__ x, int64_t n, int64_t stride)
{
    if (n <= 1024) {
        double s = 0.0;
        int64_t k = 0;
        for (; k + 8 <= n; k += 8) {
            double a[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) a[u] = x[(k + u) * stride];
#pragma unroll
            for (int u = 0; u < 8; ++u) s += a[u];
        }
        for (; k < n; ++k) s += x[k * stride];
        return s;
    }
    const int64_t h = n / 2;
    const double a = jl_sum_strided(x, h, stride);
    return a + jl_sum_strided(x + h * stride, n - h, stride);
}

// pass 2 (:31-35): normalization_factor[i,t] = sum(p_dest[i,:,t]).  One thread per (i,t).
__global__ void __launch_bounds__(64) k_pdest_norm(const double *__restrict__ p_dest, int64_t N, int64_t T,
                                                   double *__restrict__ norm)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * T) return;
    const int64_t i = idx % N, t = idx / N;
    norm[idx] = jl_sum_strided(p_dest + i + N * N * t, N, N);
}

// ------------------------------------------------------------------------------------------------
// pass 3 (:38-46) fused with the sampling-CDF build.  Each warp owns 32 consecutive origins of one
// hour; lane = origin, so the loads over j are coalesced and the running sum is accumulated strictly
// left to right in fp64 -- the association order of the reference's linear scan
// (src/resampling.jl:38-45), which is what makes destination indices bit-exact.  The CDF is written
// row-contiguous (destination fastest) through a padded shared-memory transpose.
//   DIVIDE: p /= norm where norm > 0 and write p back (createpdestin); otherwise p is read-only
//   (p_dest supplied by the caller, as in the Julia API).
// ------------------------------------------------------------------------------------------------
template <bool DIVIDE>
__global__ void __launch_bounds__(128) k_cdf_build(double *__restrict__ p_dest, const double *__restrict__ norm, int64_t N,
                                                   int64_t T, double *__restrict__ cdf, int64_t ldc, uint8_t *__restrict__ rowzero)
{
    __shared__ double tile[4][32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t i0 = ((int64_t)blockIdx.x * 4 + w) * 32;
    const int64_t t = blockIdx.y;
    if (i0 >= N) return;
    const int64_t i = i0 + lane;
    const bool live = i < N;
    const int64_t ic = live ? i : N - 1;
    double *p = p_dest + ic + N * N * t;
    double nf = 1.0;
    bool div = false;
    if (DIVIDE) { nf = norm[ic + N * t]; div = nf > 0.0; }   // NaN -> skip (:40)
    double run = 0.0;
    double *out = cdf + (t * N + i0) * ldc;
    const int64_t rows = (N - i0 < 32) ? (N - i0) : 32;
    for (int64_t j0 = 0; j0 < N; j0 += 32) {
        double v[32];
        if (j0 + 32 <= N) {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) v[jj] = p[N * (j0 + jj)];
        } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) v[jj] = (j0 + jj < N) ? p[N * (j0 + jj)] : 0.0;
        }
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
            double x = v[jj];
            if (DIVIDE && div) {
                x = x / nf;                                            // :42
                if (live && j0 + jj < N) p[N * (j0 + jj)] = x;
            }
            run = run + x;                                             // src/resampling.jl:39
            tile[w][jj][lane] = run;
        }
        __syncwarp();
        if (j0 + lane < N)
            for (int64_t r = 0; r < rows; ++r) out[r * ldc + j0 + lane] = tile[w][lane][r];
        __syncwarp();
    }
    if (live) rowzero[t * N + i] = (run == 0.0) ? 1 : 0;               // sum(distribution) == 0 (:35)
}

// ------------------------------------------------------------------------------------------------
// Guide ("cut-point") table of every CDF row: guide[b] = #{j : cdf[j] < b/G}, b = 0..G (common.cuh:
// cdf_search_guided).  One warp per row; entry j owns the thresholds b/G in (cdf[j-1], cdf[j]] (exact: G is a power
// of two), thresholds above the last finite sum see every finite entry.  A NaN sum (and everything after it) owns none.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_guide_build(double *__restrict__ cdf, int64_t ldc, int goff, int G, int64_t N, int64_t rows)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const double *rowp = cdf + row * ldc;
    uint16_t *g = reinterpret_cast<uint16_t *>(cdf + row * ldc + goff);
    const double Gd = (double)G;
    double carry = 0.0;
    int64_t n_valid = N;
    for (int64_t jb = 0; jb < N; jb += 32) {
        const int64_t j = jb + lane;
        const bool valid = j < N;
        const double cur = valid ? rowp[j] : carry;
        double prev = __shfl_up_sync(0xffffffffu, cur, 1);
        if (lane == 0) prev = carry;
        carry = __shfl_sync(0xffffffffu, cur, 31);
        if (valid && cur > prev) {
            const double lo_d = prev * Gd, hi_d = cur * Gd;
            const int b_lo = (lo_d >= Gd) ? G + 1 : (int)floor(lo_d) + 1;
            const int b_hi = (hi_d >= Gd) ? G : (int)floor(hi_d);
            for (int b = b_lo; b <= b_hi; ++b) g[b] = (uint16_t)j;
        }
        const uint32_t nanmask = __ballot_sync(0xffffffffu, valid && cur != cur);
        if (nanmask && n_valid == N) n_valid = jb + __ffs(nanmask) - 1;
    }
    const double last = n_valid > 0 ? rowp[n_valid - 1] : 0.0;
    const double lo_d = last * Gd;
    const int b_lo = (lo_d >= Gd) ? G + 1 : (int)floor(lo_d) + 1;
    for (int b = b_lo + lane; b <= G; b += 32) g[b] = (uint16_t)n_valid;
    if (lane == 0) g[0] = 0;
}

static void launch_guide(double *cdf, int64_t ldc, int goff, int G, int64_t N, int64_t T, cudaStream_t st)
{
    const int64_t rows = N * T;
    k_guide_build<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(cdf, ldc, goff, G, N, rows);
}

int launch_pdestin(const double *dm, int64_t N, int64_t T, double e_dest, double *p_dest, double *norm, cudaStream_t st)
{
    const int64_t NN = N * N;
    const unsigned g = (unsigned)((NN + 255) / 256);
    if (T <= 24) k_pdest_unnorm<24><<<g, 256, 0, st>>>(dm, N, T, e_dest, p_dest);
    else         k_pdest_unnorm<0><<<g, 256, 0, st>>>(dm, N, T, e_dest, p_dest);
    k_pdest_norm<<<(unsigned)((N * T + 63) / 64), 64, 0, st>>>(p_dest, N, T, norm);
    return 2;
}

int launch_cdf_from_pdest(const double *p_dest, int64_t N, int64_t T, double *cdf, int64_t ldc, int goff, int G,
                          uint8_t *rowzero, cudaStream_t st)
{
    dim3 g((unsigned)((N + 127) / 128), (unsigned)T);
    k_cdf_build<false><<<g, 128, 0, st>>>(const_cast<double *>(p_dest), nullptr, N, T, cdf, ldc, rowzero);
    if (G > 0) launch_guide(cdf, ldc, goff, G, N, T, st);
    return G > 0 ? 2 : 1;
}

int launch_cdf_normalise(double *p_dest, const double *norm, int64_t N, int64_t T, double *cdf, int64_t ldc, int goff,
                         int G, uint8_t *rowzero, cudaStream_t st)
{
    dim3 g((unsigned)((N + 127) / 128), (unsigned)T);
    k_cdf_build<true><<<g, 128, 0, st>>>(p_dest, norm, N, T, cdf, ldc, rowzero);
    if (G > 0) launch_guide(cdf, ldc, goff, G, N, T, st);
    return G > 0 ? 2 : 1;
}

}  // namespace cpm
