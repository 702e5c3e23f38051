// tables.cu -- probability tables of the HMM: createpdrive / createpdestin and the sampling CDF.
// Reference: src/createpdrive.jl:3-38, src/createpdestin.jl:3-50 (Julia, column-major, origin index fastest).
// All kernels are elementwise / scan work in fp64 over the N x N x T grid; no tensor cores (nothing is a
// contraction).  fp64 division is the expensive instruction on this part (a ~25-instruction sequence on a
// half-rate pipe), so every quotient whose divisor is shared by many elements is computed as
// "reciprocal once + two FMAs" (div_r below), which returns the same correctly rounded quotient.
#include "kernels.h"
#include "async.cuh"

#include <cuda.h>   // CUtensorMap and the encoder's enums (types only: the entry point is looked up at run time)

#include <cstdlib>
#include <type_traits>

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

__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, uint64_t *bar, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = [] {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            ptr = nullptr;
        return reinterpret_cast<EncodeTiledFn>(ptr);
    }();
    return fn;
}

// ------------------------------------------------------------------------------------------------
// k_pdrive_meansum_tma: the same sums as k_pdrive_meansum with the datamatrix stream staged by TMA.
// k_pdrive_meansum has only N*T threads (one sequential fp64 chain each), so its memory parallelism comes from 16
// loads in flight per thread and it stops at about half of the HBM roofline.  Here every warp is its own CTA (all
// N*T/32 resident at once) and owns a 4-stage shared-memory ring of (8 destinations x 32 origins) tiles filled by
// one cp.async.bulk.tensor each, three tiles ahead of the running sum (zero fill outside N, and a zero mean is
// "no data", :15).  The (dist, 1/dist) pairs are shared by the 24 hours and hit L2: plain 16-byte loads, fetched one
// tile ahead into registers.  Needs even N (tensor-map strides are multiples of 16 bytes).
// ------------------------------------------------------------------------------------------------
constexpr int kMtTileJ = 8, kMtStages = 4;

// HT = hours per warp (1 or 2).  With HT = 2 a lane carries two independent chains (hours t, t+1 of its origin) that
// share every (dist, 1/dist) pair: the pair traffic from L2 -- 2/3 of what this kernel moves -- halves.
template <int HT>
__global__ void __launch_bounds__(32) k_pdrive_meansum_tma(const __grid_constant__ CUtensorMap tm_dm, const double *__restrict__ dm,
                                                           const double2 *__restrict__ rdist, int64_t N, int64_t T,
                                                           double *__restrict__ mean_sum)
{
    __shared__ __align__(128) double in[kMtStages][HT][kMtTileJ][32];
    __shared__ uint64_t bar[kMtStages];
    const int lane = threadIdx.x;
    const int ngrp = (int)(T / HT);
    const int t0 = (blockIdx.x % ngrp) * HT, i0 = (blockIdx.x / ngrp) * 32;   // consecutive CTAs: same origins, consecutive hours
    const int64_t i = (int64_t)i0 + lane;                                     // (the pairs they share are fetched from HBM once)
    const bool live = i < N;
    const int64_t ic = live ? i : N - 1;
    const double2 *dr = rdist + ic;
    const int ntiles = (int)((N + kMtTileJ - 1) / kMtTileJ);
    constexpr uint32_t kTileBytes = HT * kMtTileJ * 32 * 8;
    const uint64_t pol = l2_evict_first_policy();

    if (elect_one()) {
#pragma unroll
        for (int s = 0; s < kMtStages; ++s) mbar_init(&bar[s], 1);
        fence_mbar_init();
        for (int k = 0; k < kMtStages - 1 && k < ntiles; ++k) {
            mbar_expect_tx(&bar[k], kTileBytes);
            tma_load_3d(&in[k][0][0][0], &tm_dm, i0, k * kMtTileJ, t0, &bar[k], pol);
        }
    }
    __syncwarp();

    double2 dv[kMtTileJ], dn[kMtTileJ];
    auto fetch_pairs = [&](int k, double2 *dst) {
        const int64_t j0 = (int64_t)k * kMtTileJ;
#pragma unroll
        for (int u = 0; u < kMtTileJ; ++u) dst[u] = (j0 + u < N) ? __ldg(dr + N * (j0 + u)) : make_double2(1.0, 1.0);
    };
    fetch_pairs(0, dv);
    double s[HT];
    int cnt[HT];
#pragma unroll
    for (int h = 0; h < HT; ++h) { s[h] = 0.0; cnt[h] = 0; }
    for (int k = 0; k < ntiles; ++k) {
        const int st = k & (kMtStages - 1);
        if (k + 1 < ntiles) fetch_pairs(k + 1, dn);
        mbar_wait(&bar[st], (uint32_t)(k / kMtStages) & 1u);
#pragma unroll
        for (int u = 0; u < kMtTileJ; ++u) {
#pragma unroll
            for (int h = 0; h < HT; ++h) {
                const double mv = in[st][h][u][lane];
                if (mv != 0.0) { s[h] = s[h] + div_r(mv, dv[u].x, dv[u].y); ++cnt[h]; }   // :15-17
            }
        }
        __syncwarp();   // every lane has consumed in[st]
        const int kn = k + kMtStages - 1;
        if (kn < ntiles && elect_one()) {   // refill the stage consumed in the previous iteration
            const int sn = kn & (kMtStages - 1);
            mbar_expect_tx(&bar[sn], kTileBytes);
            tma_load_3d(&in[sn][0][0][0], &tm_dm, i0, kn * kMtTileJ, t0, &bar[sn], pol);
        }
#pragma unroll
        for (int u = 0; u < kMtTileJ; ++u) dv[u] = dn[u];
    }
    if (!live) return;
#pragma unroll
    for (int h = 0; h < HT; ++h) {
        const int64_t t = t0 + h;
        double sh = s[h];
        if (!(fabs(sh) < 1.0e300)) {   // a zero / non-finite distance or an overflow: redo the row with the division instruction
            const double *m = dm + i + N * N * t;
            sh = 0.0;
            for (int64_t j = 0; j < N; ++j) {
                const double mv = __ldg(m + N * j);
                if (mv != 0.0) sh = sh + mv / __ldg(dr + N * j).x;
            }
        }
        mean_sum[i + N * t] = sh / (double)cnt[h];   // 0/0 -> NaN, as in Julia (:20)
    }
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
    EncodeTiledFn enc = encode_tiled();
    if (enc && N % 2 == 0 && N >= 2) {   // dm's mean slab as [t][j][o] fp64, box 32 (o) x 8 (j) x HT hours
        CUtensorMap tm;
        const cuuint64_t dim[3] = {(cuuint64_t)N, (cuuint64_t)N, (cuuint64_t)T};
        const cuuint64_t str[2] = {(cuuint64_t)N * 8, (cuuint64_t)N * N * 8};
        const cuuint32_t ht = (T % 2 == 0) ? 2 : 1;
        const cuuint32_t box[3] = {32, (cuuint32_t)kMtTileJ, ht}, ones[3] = {1, 1, 1};
        if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(dm), dim, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
            const unsigned grid = (unsigned)(((N + 31) / 32) * (T / ht));
            if (ht == 2) k_pdrive_meansum_tma<2><<<grid, 32, 0, st>>>(tm, dm, dr, N, T, mean_sum);
            else         k_pdrive_meansum_tma<1><<<grid, 32, 0, st>>>(tm, dm, dr, N, T, mean_sum);
            return 2;
        }
    }
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
    double mx, mn;
    bool has_nan = false;   // fmax/fmin drop NaN, Julia's maximum/minimum propagate it (and keep Inf, as fmax/fmin do)
    if constexpr (TMAX > 0) {
#pragma unroll
        for (int t = 0; t < TMAX; ++t) v[t] = (t < T) ? __ldg(dm + idx + NN * t) : 0.0;
        mx = mn = v[0];
#pragma unroll
        for (int t = 0; t < TMAX; ++t)
            if (t < T) { mx = fmax(mx, v[t]); mn = fmin(mn, v[t]); has_nan |= v[t] != v[t]; }
    } else {
        mx = mn = __ldg(dm + idx);
        for (int64_t t = 0; t < T; ++t) {
            const double x = __ldg(dm + idx + NN * t);
            mx = fmax(mx, x); mn = fmin(mn, x); has_nan |= x != x;
        }
    }
    if (has_nan) mx = mn = __longlong_as_double(0x7ff8000000000000LL);   // NaN among the inputs
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
    const int64_t h = (n + 1) / 2;   // imid = (ifirst + ilast) >> 1: the first half gets ceil(n/2) elements (base/reduce.jl)
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

constexpr int kCdfWarps = 1;   // warps per CTA: one, so that the N*T/32 row groups spread evenly over the SMs

template <bool DIVIDE, bool WRITE_P>
__global__ void __launch_bounds__(32 * kCdfWarps) k_cdf_build(double *__restrict__ p_dest, const double *__restrict__ norm, int64_t N,
                                                   int64_t T, double *__restrict__ cdf, int64_t ldc, int goff, int G,
                                                   uint8_t *__restrict__ rowzero)
{
    __shared__ double tile[kCdfWarps][32][33];
    __shared__ uint16_t gidx[kCdfWarps][32][34];   // guide cell reached by each running sum: min(G, floor(run * G)); 0 for NaN
    __shared__ uint16_t gcarry[kCdfWarps][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t i0 = ((int64_t)blockIdx.x * kCdfWarps + w) * 32;
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
    const bool warp_straight = __all_sync(0xffffffffu, fast || !div);
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
        // One step of the running sum.  STRAIGHT: every lane of the warp takes the reciprocal path (or does not
        // divide), so the step is branch-free; otherwise lanes whose reciprocal is unusable call the division.
        auto step = [&](int jj, auto straight) {
            double x = v[jj];
            if (DIVIDE) {
                if constexpr (decltype(straight)::value) {
                    const double q = div_r(x, nf, rnf);
                    x = div ? q : x;                                   // :40-42
                } else if (div) x = fast ? div_r(x, nf, rnf) : div_slow(x, nf);
                if (WRITE_P && div && live && j0 + jj < N) p[N * (j0 + jj)] = x;
            }
            run = run + x;                                             // src/resampling.jl:39
            tile[w][jj][lane] = run;
            // thresholds b/G <= run  <=>  b <= floor(run * G) (exact: G is a power of two).  The conversion saturates
            // (Inf -> INT_MAX) and maps NaN to 0: a NaN sum owns no cell, the tail fix-up below fills them.
            gidx[w][jj][lane] = (uint16_t)min(max(__double2int_rd(run * Gd), 0), G);
        };
        if (warp_straight) {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) step(jj, std::true_type{});
        } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) step(jj, std::false_type{});   // full unroll: v[] stays in registers
        }
        __syncwarp();
        if (j0 + lane < N) {
            // This loop is 2/3 of the kernel's instructions (profiles/README.md), so it is written for instruction
            // count: one byte pointer per row (bumped by the row pitch), the guide base at a loop-invariant distance
            // from it, and a full unroll for complete tiles.
            char *rowp = reinterpret_cast<char *>(out + j0 + lane);                    // &cdf[row 0][j]
            const int64_t gdelta = ((int64_t)goff - (j0 + lane)) * 8;                  // &cdf[r][j] -> guide cells of row r
            const int64_t pitch = ldc * 8;
            const uint16_t jme = (uint16_t)(j0 + lane);
            const double *tl = &tile[w][lane][0];
            const uint16_t *gl = &gidx[w][lane][0];
            const uint16_t *gp = (lane > 0) ? &gidx[w][lane - 1][0] : &gcarry[w][0];
            // entry j owns the guide cells (floor(prev*G), floor(cur*G)]: none when the sum did not grow or is NaN.
            // Nearly all entries own 0-3 cells: three predicated stores in the straight-line body; the few that own
            // more are finished by a second pass that runs only when the tile has one.
            int maxown = 0;
            auto put = [&](int r, auto guide) {
                *reinterpret_cast<double *>(rowp) = tl[r];
                if constexpr (decltype(guide)::value) {
                    const int b_prev = gp[r], own = (int)gl[r] - b_prev;
                    uint16_t *g = reinterpret_cast<uint16_t *>(rowp + gdelta) + b_prev;
                    if (own >= 1) g[1] = jme;
                    if (own >= 2) g[2] = jme;
                    if (own >= 3) g[3] = jme;
                    maxown = max(maxown, own);
                }
                rowp += pitch;
            };
            if (G > 0) {
                if (rows == 32) {
#pragma unroll
                    for (int r = 0; r < 32; ++r) put(r, std::true_type{});
                } else {
                    for (int r = 0; r < rows; ++r) put(r, std::true_type{});
                }
                if (maxown > 3) {
                    char *gp8 = reinterpret_cast<char *>(out + j0 + lane) + gdelta;
                    for (int r = 0; r < rows; ++r, gp8 += pitch) {
                        const int b_prev = gp[r], own = (int)gl[r] - b_prev;
                        uint16_t *g = reinterpret_cast<uint16_t *>(gp8) + b_prev;
                        for (int c = 4; c <= own; ++c) g[c] = jme;
                    }
                }
            } else {
                if (rows == 32) {
#pragma unroll
                    for (int r = 0; r < 32; ++r) put(r, std::false_type{});
                } else {
                    for (int r = 0; r < rows; ++r) put(r, std::false_type{});
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
    dim3 g((unsigned)((N + 32 * kCdfWarps - 1) / (32 * kCdfWarps)), (unsigned)T);
    k_cdf_build<false, false><<<g, 32 * kCdfWarps, 0, st>>>(const_cast<double *>(p_dest), nullptr, N, T, cdf, ldc, goff, G, rowzero);
    return 1;
}

int launch_cdf_normalise(double *p_dest, const double *norm, int64_t N, int64_t T, double *cdf, int64_t ldc, int goff,
                         int G, uint8_t *rowzero, int write_p, cudaStream_t st)
{
    dim3 g((unsigned)((N + 32 * kCdfWarps - 1) / (32 * kCdfWarps)), (unsigned)T);
    if (write_p) k_cdf_build<true, true><<<g, 32 * kCdfWarps, 0, st>>>(p_dest, norm, N, T, cdf, ldc, goff, G, rowzero);
    else         k_cdf_build<true, false><<<g, 32 * kCdfWarps, 0, st>>>(p_dest, norm, N, T, cdf, ldc, goff, G, rowzero);
    return 1;
}

}  // namespace cpm
