// common.cuh -- shared device helpers for libcpm (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cpm {

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  counter = (car_lo, car_hi, step, 2*stream+phase),
// key = (seed_lo, seed_hi).  Standard Random123 round constants; pinned by known-answer vectors in tests/.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 52-bit uniform on [0,1): the resolution of Julia <= 1.6 rand() (reference era).
__device__ __forceinline__ double u52(uint32_t lo, uint32_t hi)
{
    // m * 2^-52 with m = the top 52 bits: placed in the mantissa of a double in [1, 2), minus 1 -- exact, and three
    // integer operations + one add instead of a 64-bit integer-to-double conversion and a multiply
    const uint32_t mh = (hi >> 12) | 0x3FF00000u, ml = (hi << 20) | (lo >> 12);
    return __hiloint2double((int)mh, (int)ml) - 1.0;
}

struct RngDev {
    int32_t  mode;        // 0 philox, 1 injected
    uint32_t stream;
    uint32_t k0, k1;
    int64_t  car_offset;
    int64_t  ld;          // leading dimension (= C) of the injected arrays
    const double *u_drive, *u_dest, *u_time, *u_dist;
    uint32_t rk0[10], rk1[10];   // the ten round keys (k + r * Weyl constant), precomputed on the host: the kernels read them
                                 // straight from the constant bank instead of bumping two registers per round
};

// Philox4x32-10 with the round keys of `r`: words of (car, step, phase); u52(o0,o1) and u52(o2,o3) are the two uniforms
__device__ __forceinline__ void philox_words(const RngDev &r, int64_t car, uint32_t step, uint32_t phase, uint32_t (&o)[4])
{
    const uint64_t g = (uint64_t)(r.car_offset + car);
    uint32_t c0 = (uint32_t)g, c1 = (uint32_t)(g >> 32), c2 = step, c3 = r.stream * 2u + phase;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ r.rk0[i];
        const uint32_t n2 = hi0 ^ c3 ^ r.rk1[i];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

// phase 0 -> (u_drive, u_dest); phase 1 -> (u_time, u_dist)
__device__ __forceinline__ void philox_pair(const RngDev &r, int64_t car, uint32_t step, uint32_t phase,
                                            double &ua, double &ub)
{
    uint32_t o[4];
    philox_words(r, car, step, phase, o);
    ua = u52(o[0], o[1]);
    ub = u52(o[2], o[3]);
}

// ---------------------------------------------------------------------------------------------
// Julia semantics helpers
// ---------------------------------------------------------------------------------------------
// Float64 ^ e with Julia's Float64^Int fast paths (base/math.jl) for integral e in {-1,0,1,2,3}.
__device__ __forceinline__ double jl_pow(double x, double e)
{
    if (e == 2.0) return x * x;
    if (e == 1.0) return x;
    if (e == 3.0) return x * x * x;
    if (e == 0.0) return 1.0;
    if (e == -1.0) return 1.0 / x;
    return pow(x, e);
}

// erf(x) for |x| <= 1.5 by its Maclaurin series x * sum_n (2/sqrt(pi)) (-1)^n x^(2n) / (n! (2n+1)), 24 terms in Horner form
// (relative error < 5e-16 on that range, checked against scipy.special.erf).  One straight-line chain: the library erf()
// switches between a polynomial and an exp-based branch at |x| = 1, and a warp whose lanes straddle 1 pays for both.
__device__ __forceinline__ double erf_small(double x)
{
    const double t = x * x;
    double p = -9.28672901131906e-25;
    p = fma(p, t, 2.2308786802746453e-23);
    p = fma(p, t, -5.136209054585811e-22);
    p = fma(p, t, 1.131218725924631e-20);
    p = fma(p, t, -2.3784598852774293e-19);
    p = fma(p, t, 4.763348040515068e-18);
    p = fma(p, t, -9.063970842808673e-17);
    p = fma(p, t, 1.6342614095367152e-15);
    p = fma(p, t, -2.7835162072109212e-14);
    p = fma(p, t, 4.4632242632864775e-13);
    p = fma(p, t, -6.7113668551641105e-12);
    p = fma(p, t, 9.422759064650411e-11);
    p = fma(p, t, -1.2290555301717926e-09);
    p = fma(p, t, 1.4807192815879218e-08);
    p = fma(p, t, -1.6365844691234924e-07);
    p = fma(p, t, 1.6462114365889246e-06);
    p = fma(p, t, -1.492565035840625e-05);
    p = fma(p, t, 0.00012055332981789664);
    p = fma(p, t, -0.0008548327023450852);
    p = fma(p, t, 0.005223977625442188);
    p = fma(p, t, -0.026866170645131252);
    p = fma(p, t, 0.11283791670955126);
    p = fma(p, t, -0.37612638903183754);
    p = fma(p, t, 1.1283791670955126);
    return x * p;
}

// Truncated(Normal(mu, sigma), 0.9mu, 1.1mu) by inverse CDF from one uniform (src/resampling.jl:68,74).
// Symmetric interval: z = sqrt(2) * erfinv((2u-1) * erf(b/sqrt(2))), b = 0.1*mu/sigma.
__device__ __forceinline__ double truncnormal(double mu, double sigma, double u)
{
    if (!(sigma > 0.0) || isinf(sigma) || isinf(mu) || isnan(mu)) return mu;
    double lo = 0.9 * mu, hi = 1.1 * mu;
    const double b = (0.1 * mu) / sigma;
    const double xe = b * 0.70710678118654752440;   // b / sqrt(2) to within an ulp (a division costs ~20 instructions)
    const double E = (fabs(xe) <= 1.5) ? erf_small(xe) : erf(xe);
    // z = sqrt(2) * erfinv((2u-1) * E) = Phi^-1(1/2 + q), q = (u - 1/2) * E.  For |q| <= 0.425 (every draw when
    // E <= 0.85, i.e. sigma > 0.07 mu; most draws otherwise) Wichura's AS 241 central rational (PPND16, |error| ~ 1e-16)
    // gives it in two degree-7 Horner chains and one division instead of erfinv's logarithm + branch cascade.
    const double q = (u - 0.5) * E;
    double z;
    if (fabs(q) <= 0.425) {
        const double r = 0.180625 - q * q;
        double A = 2509.0809287301226727, B = 5226.4952788528545610;
        A = fma(A, r, 33430.575583588128105); B = fma(B, r, 28729.085735721942674);
        A = fma(A, r, 67265.770927008700853); B = fma(B, r, 39307.895800092710610);
        A = fma(A, r, 45921.953931549871457); B = fma(B, r, 21213.794301586595867);
        A = fma(A, r, 13731.693765509461125); B = fma(B, r, 5394.1960214247511077);
        A = fma(A, r, 1971.5909503065514427); B = fma(B, r, 687.18700749205790830);
        A = fma(A, r, 133.14166789178437745); B = fma(B, r, 42.313330701600911252);
        A = fma(A, r, 3.3871328727963666080); B = fma(B, r, 1.0);
        z = q * A / B;
    } else {
        z = 1.41421356237309504880 * erfinv(2.0 * q);
    }
    double x = mu + sigma * z;
    if (lo > hi) { const double t = lo; lo = hi; hi = t; }
    if (x < lo) x = lo;
    if (x > hi) x = hi;
    return x;
}

// Categorical draw on one CDF row (src/resampling.jl:38-45).  The reference scans j = 1..N with a running
// fp64 sum and takes the first j with range_low < u <= range_up.  On the same left-to-right running sum that
// is lower_bound(u) followed by the acceptance test u <= cdf[j]: the test fails (sentinel 0, :32) when u lies
// above the last entry or the running sum has turned NaN (`cdf < u` is false for NaN, so the search goes left
// and stays correct on the finite prefix); u == 0.0 never satisfies `range_low < u` and also yields 0.
// Returns the 1-based destination or 0.
template <typename LoadF>
__device__ __forceinline__ int32_t cdf_search(LoadF load, int32_t n, double u)
{
    if (!(u > 0.0)) return 0;
    int32_t lo = 0, hi = n;
    while (lo < hi) {
        const int32_t mid = (lo + hi) >> 1;
        if (load(mid) < u) lo = mid + 1; else hi = mid;   // NaN compares false -> goes left
    }
    if (lo >= n) return 0;
    return (u <= load(lo)) ? lo + 1 : 0;
}

// Fixed-point driving-time accumulator: whole seconds in `hi`, 2^-32 s units in `lo`.  Integer adds commute,
// so the sum is bit-reproducible for any processing order, grid shape, sampler or GPU split.
__device__ __forceinline__ void fx_add(unsigned long long &hi, unsigned long long &lo, double x)
{
    if (!(x >= 0.0) || !(x < 9.0e15)) return;   // NaN / negative / absurd: not accumulated (cannot come from valid tables)
    const double f = floor(x);
    hi += (unsigned long long)f;
    lo += (unsigned long long)rint((x - f) * 4294967296.0);
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// The same draw with a guide ("cut-point") table: guide[b] = #{j : cdf[j] < b/G}, b = 0..G, G a power of two.
// For b = floor(u*G):  guide[b] <= lower_bound(u) <= guide[b+1], so the binary search runs over a handful of
// entries instead of N.  All comparisons are the exact fp64 ones of cdf_search, hence the same index.
template <typename LoadF, typename LoadG>
__device__ __forceinline__ int32_t cdf_search_guided(LoadF load, LoadG guide, int32_t n, int32_t G, double u)
{
    if (!(u > 0.0)) return 0;
    int32_t lo, hi;
    if (u < 1.0) { const int32_t b = (int32_t)(u * (double)G); lo = guide(b); hi = guide(b + 1); }
    else { lo = guide(G); hi = n; }
    while (lo < hi) {
        const int32_t mid = (lo + hi) >> 1;
        if (load(mid) < u) lo = mid + 1; else hi = mid;
    }
    if (lo >= n) return 0;
    return (u <= load(lo)) ? lo + 1 : 0;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace cpm
