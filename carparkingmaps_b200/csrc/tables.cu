// tables.cu -- probability tables of the HMM: createpdrive / createpdestin and the sampling CDF.
// Reference: src/createpdrive.jl:3-38, src/createpdestin.jl:3-50 (Julia, column-major, origin index fastest).
// All kernels are elementwise / scan work in fp64 over the N x N x T grid; no tensor cores (nothing is a
// contraction).  fp64 division is the expensive instruction on this part (a ~25-instruction sequence on a
// half-rate pipe), so every quotient whose divisor is shared by many elements is computed as
// "reciprocal once + two FMAs" (div_r below), which returns the same correctly rounded quotient.
#include "kernels.h"

namespace cpm {

// a / d given r = 1/d correctly rounded (Markstein): q = a*r is within an ulp or two of the quotient; the residual
// e = a - q*d is exact in an FMA; q + e*r rounds to the correctly rounded quotient.  Valid when r is finite and
// non-zero and nothing overflows; callers check that once per divisor (or once per row) and otherwise take the
// division instruction, so special values behave exactly like Julia's `/`.
__device__ __forceinline__ double div_r(double a, double d, double r)
{
    const double q = a * r;
    return __fma_rn(__fma_rn(-q, d, a), r, q);
}
__device__ __forceinline__ bool recip_ok(double r) { return fabs(r) < 1.0e300 && fabs(r) > 1.0e-300; }   // NaN -> false

// ------------------------------------------------------------------------------------------------
// createpdrive, part 1: mean_sum[i,t] = (sum_{j: m!=0} m[i,j,t]/dist[i,j]) / count   (:10-21)
// One thread per (i,t), i fastest -> every load of the j loop is coalesced across the warp.  The running sum is
// accumulated in ascending j exactly like the reference, so mean_sum is bit-identical to a sequential fp64
// evaluation.  rdist = 1/dist is computed once per (i,j) by k_rdist instead of 24 times.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rdist(const double *__restrict__ dist, int64_t n, double2 *__restrict__ rdist)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) { const double d = dist[k]; rdist[k] = make_double2(d, 1.0 / d); }
}

constexpr int kMsUnroll = 16;   // loads in flight per thread: only N*T threads exist, so memory parallelism must come from ILP

__global__ void __launch_bounds__(64) k_pdrive_meansum(const double *__restrict__ dm, const double2 *__restrict__ rdist,
                                                       int64_t N, int64_t T, double *__restrict__ mean_sum)
{
    // blocks of the same origin chunk and consecutive hours are adjacent in launch order, so the (dist, 1/dist) pairs
    // they share are fetched from HBM once and then hit L2
    const int64_t t = blockIdx.x % T;
    const int64_t i = (int64_t)(blockIdx.x / T) * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int64_t idx = i + N * t;
    const double *m = dm + i + N * N * t;
    const double2 *dr = rdist + i;
    double s = 0.0;
    int cnt = 0;
    int64_t j = 0;
    for (; j + kMsUnroll <= N; j += kMsUnroll) {
        double mv[kMsUnroll];
        double2 dv[kMsUnroll];
#pragma unroll
        for (int u = 0; u < kMsUnroll; ++u) { mv[u] = __ldg(m + N * (j + u)); dv[u] = __ldg(dr + N * (j + u)); }
#pragma unroll
        for (int u = 0; u < kMsUnroll; ++u)
            if (mv[u] != 0.0) { s = s + div_r(mv[u], dv[u].x, dv[u].y); ++cnt; }   // :15-17
    }
    for (; j < N; ++j) {
        const double mv = __ldg(m + N * j);
        const double2 dv = __ldg(dr + N * j);
        if (mv != 0.0) { s = s + div_r(mv, dv.x, dv.y); ++cnt; }
    }
    if (!(fabs(s) < 1.0e300)) {   // a zero / non-finite distance or an overflow: redo the row with the division instruction
        s = 0.0;
        for (j = 0; j < N; ++j) {
            const double mv = __ldg(m + N * j);
            if (mv != 0.0) s = s + mv / __ldg(dr + N * j).x;
        }
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
    bool has_nan = false;
    for (int64_t t = 0; t < T; ++t) {
        const double v = mean_sum[i + N * t];
        has_nan |= isnan(v);
        if (v > mx) mx = v;
        if (v < mn) mn = v;
    }
    if (has_nan) mx = mn = __longlong_as_double(0x7ff8000000000000LL);
    const bool ok = mx > 0.0;   // NaN -> false: the row stays zero (:28)
    for (int64_t t = 0; t < T; ++t) {
        double p = 0.0;
        if (ok) p = p_min + (p_max - p_min) * jl_pow((mean_sum[i + N * t] - mn) / (mx - mn), e_drive);   // :31
        p_drive[i + N * t] = p;
    }
}

int launch_pdrive_meansum(const double *dm, const double *dist, double *rdist, int64_t N, int64_t T, double *mean_sum,
                          cudaStream_t st)
{
    const int64_t NN = N * N, n = N * T;
    double2 *dr = reinterpret_cast<double2 *>(rdist);   // [N,N] pairs (dist, 1/dist)
    k_rdist<<<(unsigned)((NN + 255) / 256), 256, 0, st>>>(dist, NN, dr);
    (void)n;
    k_pdrive_meansum<<<(unsigned)(((N + 63) / 64) * T), 64, 0, st>>>(dm, dr, N, T, mean_sum);
    return 2;
}

int launch_pdrive_finish(const double *mean_sum, int64_t N, int64_t T, double p_min, double p_max, double e_drive,
                         double *p_drive, cudaStream_t st)
{
    k_pdrive_finish<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(mean_sum, N, T, p_min, p_max, e_drive, p_drive);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// createpdestin, pass 1 (:10-28): per (i,j) min/max over t (zeros included), then the unnormalised
// ((m-min)/(max-min))^e_dest.  One thread per (i,j), i fastest: coalesced.  TMAX > 0 keeps the T values in
// registers (T = 24 in the reference); TMAX == 0 re-reads them.  EK selects Julia's Float64^Int fast paths at
// compile time (2: x*x -- the reference's e_dest = 2, main.jl:38; 1: x; 3: x*x*x) or the general pow (0).
// ------------------------------------------------------------------------------------------------
template <int EK>
__device__ __forceinline__ double pow_k(double x, double e)
{
    if (EK == 2) return x * x;
    if (EK == 1) return x;
    if (EK == 3) return x * x * x;
    return jl_pow(x, e);
}

template <int TMAX, int EK>
__global__ void __launch_bounds__(256) k_pdest_unnorm(const double *__restrict__ dm, int64_t N, int64_t T, double e_dest,
                                                      double *__restrict__ p_dest)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t NN = N * N;
    if (idx >= NN) return;
    double v[TMAX > 0 ? TMAX : 1];
    double mx, mn, nan_probe = 0.0;   // fmax/fmin drop NaN; the running sum keeps it (Julia's maximum/minimum propagate NaN)
    if constexpr (TMAX > 0) {
#pragma unroll
        for (int t = 0; t < TMAX; ++t) v[t] = (t < T) ? __ldg(dm + idx + NN * t) : 0.0;
        mx = mn = v[0];
#pragma unroll
        for (int t = 0; t < TMAX; ++t)
            if (t < T) { mx = fmax(mx, v[t]); mn = fmin(mn, v[t]); nan_probe += v[t] * 0.0; }
    } else {
        mx = mn = __ldg(dm + idx);
        for (int64_t t = 0; t < T; ++t) {
            const double x = __ldg(dm + idx + NN * t);
            mx = fmax(mx, x); mn = fmin(mn, x); nan_probe += x * 0.0;
        }
    }
    if (nan_probe != nan_probe) mx = mn = __longlong_as_double(0x7ff8000000000000LL);   // NaN or Inf among the inputs
    const bool ok = mx > 0.0;   // :22
    const double den = mx - mn;
    const double rden = 1.0 / den;
    const bool fast = recip_ok(rden);   // else: den is 0 (constant pair -> 0/0 = NaN, :24), Inf or NaN
    if constexpr (TMAX > 0) {
        if (fast || !ok) {
#pragma unroll
            for (int t = 0; t < TMAX; ++t)
                if (t < T) p_dest[idx + NN * t] = ok ? pow_k<EK>(div_r(v[t] - mn, den, rden), e_dest) : 0.0;   // :24
        } else {
#pragma unroll 1
            for (int64_t t = 0; t < T; ++t)   // re-reads instead of indexing v[] dynamically (keeps v[] in registers)
                p_dest[idx + NN * t] = pow_k<EK>((__ldg(dm + idx + NN * t) - mn) / den, e_dest);
        }
    } else {
#pragma unroll 1
        for (int64_t t = 0; t < T; ++t)
            p_dest[idx + NN * t] = ok ? pow_k<EK>((__ldg(dm + idx + NN * t) - mn) / den, e_dest) : 0.0;
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
// pass 3 (:38-46) fused with the sampling-CDF build and its guide table.  Each warp owns 32 consecutive origins of
// one hour.  Scan phase: lane = origin, so the loads over j are coalesced and the running sum is accumulated
// strictly left to right in fp64 -- the association order of the reference's linear scan
// (src/resampling.jl:38-45), which is what makes destination indices bit-exact.  Write phase: through a padded
// shared-memory transpose, lane = destination, rows are written contiguously (256 B per row and tile) and each
// entry fills the guide cells it owns: guide[b] = #{j : cdf[j] < b/G}; entry j owns the thresholds b/G in
// (cdf[j-1], cdf[j]] (exact: G is a power of two), thresholds above the last finite sum see every finite entry,
// a NaN sum (and everything after it) owns none (common.cuh: cdf_search_guided).
//   DIVIDE : p /= norm where norm > 0 (createpdestin :40-42); otherwise p is the caller's p_dest (Julia API)
//   WRITE_P: also write the normalised p back (only when the caller wants p_dest itself)
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ double div_slow(double a, double d) { return a / d; }

template <bool DIVIDE, bool WRITE_P>
__global__ void __launch_bounds__(128) k_cdf_build(double *__restrict__ p_dest, const double *__restrict__ norm, int64_t N,
                                                   int64_t T, double *__restrict__ cdf, int64_t ldc, int goff, int G,
                                                   uint8_t *__restrict__ rowzero)
{
    __shared__ double tile[4][32][33];
    __shared__ uint16_t gidx[4][32][34];   // guide cell reached by each running sum: min(G, floor(run * G)); 0 for NaN
    __shared__ uint16_t gcarry[4][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t i0 = ((int64_t)blockIdx.x * 4 + w) * 32;
    const int64_t t = blockIdx.y;
    if (i0 >= N) return;
    const int64_t i = i0 + lane;
    const bool live = i < N;
    const int64_t ic = live ? i : N - 1;
    double *p = p_dest + ic + N * N * t;
    double nf = 1.0, rnf = 1.0;
    bool div = false;
    if (DIVIDE) { nf = norm[ic + N * t]; div = nf > 0.0; rnf = 1.0 / nf; }   // NaN -> skip (:40)
    const bool fast = recip_ok(rnf);
    double run = 0.0;
    double *out = cdf + (t * N + i0) * ldc;
    const int rows = (int)((N - i0 < 32) ? (N - i0) : 32);
    const double Gd = (double)G;
    gcarry[w][lane] = 0;
    __syncwarp();
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
                x = fast ? div_r(x, nf, rnf) : div_slow(x, nf);        // :42
                if (WRITE_P && live && j0 + jj < N) p[N * (j0 + jj)] = x;
            }
            run = run + x;                                             // src/resampling.jl:39
            tile[w][jj][lane] = run;
            // thresholds b/G <= run  <=>  b <= floor(run * G) (exact: G is a power of two); NaN converts to 0
            gidx[w][jj][lane] = (uint16_t)__double2int_rd(fmin(run * Gd, Gd));
        }
        __syncwarp();
        if (j0 + lane < N) {
            double *orow = out + j0 + lane;
            const uint16_t jme = (uint16_t)(j0 + lane);
            const double *tl = &tile[w][lane][0];
            const uint16_t *gl = &gidx[w][lane][0];
            const uint16_t *gp = (lane > 0) ? &gidx[w][lane - 1][0] : &gcarry[w][0];
            for (int r = 0; r < rows; ++r, orow += ldc) {
                *orow = tl[r];
                // entry j owns the guide cells (floor(prev*G), floor(cur*G)]: none when the sum did not grow or is NaN
                const int b_hi = gl[r], b_lo = gp[r] + 1;
                if (G > 0) {   // most entries own 0, 1 or 2 cells: two predicated stores, a loop only for the rest
                    uint16_t *g = reinterpret_cast<uint16_t *>(orow - (j0 + lane) + goff);
                    if (b_hi >= b_lo) g[b_lo] = jme;
                    if (b_hi > b_lo) g[b_lo + 1] = jme;
                    if (b_hi > b_lo + 1)
                        for (int b = b_lo + 2; b <= b_hi; ++b) g[b] = jme;
                }
            }
        }
        __syncwarp();
        gcarry[w][lane] = gidx[w][31][lane];
        __syncwarp();
    }
    if (!live) return;
    const bool zero = (run == 0.0);
    rowzero[t * N + i] = zero ? 1 : 0;                                  // sum(distribution) == 0 (:35)
    if (G > 0 && !zero) {
        // thresholds above the last finite running sum see every finite entry (rare slow path: a NaN sum)
        const double *myrow = cdf + (t * N + i) * ldc;
        int64_t n_valid = N;
        double last = run;
        if (run != run) {
            n_valid = 0;
            while (n_valid < N && myrow[n_valid] == myrow[n_valid]) ++n_valid;
            last = n_valid > 0 ? myrow[n_valid - 1] : 0.0;
        }
        uint16_t *g = reinterpret_cast<uint16_t *>(cdf + (t * N + i) * ldc + goff);
        for (int b = __double2int_rd(fmin(last * Gd, Gd)) + 1; b <= G; ++b) g[b] = (uint16_t)n_valid;
        g[0] = 0;
    }
}

template <int EK>
static void launch_unnorm(const double *dm, int64_t N, int64_t T, double e_dest, double *p_dest, cudaStream_t st)
{
    const unsigned g = (unsigned)((N * N + 255) / 256);
    if (T <= 24) k_pdest_unnorm<24, EK><<<g, 256, 0, st>>>(dm, N, T, e_dest, p_dest);
    else         k_pdest_unnorm<0, EK><<<g, 256, 0, st>>>(dm, N, T, e_dest, p_dest);
}

int launch_pdestin(const double *dm, int64_t N, int64_t T, double e_dest, double *p_dest, double *norm, cudaStream_t st)
{
    if (e_dest == 2.0) launch_unnorm<2>(dm, N, T, e_dest, p_dest, st);
    else if (e_dest == 1.0) launch_unnorm<1>(dm, N, T, e_dest, p_dest, st);
    else if (e_dest == 3.0) launch_unnorm<3>(dm, N, T, e_dest, p_dest, st);
    else launch_unnorm<0>(dm, N, T, e_dest, p_dest, st);
    k_pdest_norm<<<(unsigned)((N * T + 63) / 64), 64, 0, st>>>(p_dest, N, T, norm);
    return 2;
}

int launch_cdf_from_pdest(const double *p_dest, int64_t N, int64_t T, double *cdf, int64_t ldc, int goff, int G,
                          uint8_t *rowzero, cudaStream_t st)
{
    dim3 g((unsigned)((N + 127) / 128), (unsigned)T);
    k_cdf_build<false, false><<<g, 128, 0, st>>>(const_cast<double *>(p_dest), nullptr, N, T, cdf, ldc, goff, G, rowzero);
    return 1;
}

int launch_cdf_normalise(double *p_dest, const double *norm, int64_t N, int64_t T, double *cdf, int64_t ldc, int goff,
                         int G, uint8_t *rowzero, int write_p, cudaStream_t st)
{
    dim3 g((unsigned)((N + 127) / 128), (unsigned)T);
    if (write_p) k_cdf_build<true, true><<<g, 128, 0, st>>>(p_dest, norm, N, T, cdf, ldc, goff, G, rowzero);
    else         k_cdf_build<true, false><<<g, 128, 0, st>>>(p_dest, norm, N, T, cdf, ldc, goff, G, rowzero);
    return 1;
}

}  // namespace cpm
