// ingest.cu -- createdatamatrix on the device: the Uber Movement rows
// (sourceid, dstid, hod, mean_travel_time, standard_deviation_travel_time) are scattered into the dense
// datamatrix[N,N,T,2] the tables are built from.  Reference: src/createdatamatrix.jl:8-23 -- zone ID 0 becomes N,
// hour 0 becomes 24 (= T), rows are applied in file order, so for duplicate (src,dst,hod) keys the LAST row wins.
// Uploading the rows instead of the dense array moves 28 B per row instead of 16 B per (i,j,t) cell: ~9x less PCIe
// traffic for Melbourne-like data (91 % empty, README.md:310).
//
// "Last row wins" without a sort: pass 1 claims every touched cell with atomicMax(row index + 1) on the cell's mean
// slot (row indices are tiny integers = denormal bit patterns, which no travel time has), pass 2 lets the winning
// row write the std slot, pass 3 the mean slot.  Out-of-range rows are counted and skipped (the reference raises
// BoundsError).
#include "kernels.h"

namespace cpm {

__device__ __forceinline__ int64_t row_cell(const int32_t *src, const int32_t *dst, const int32_t *hod, int64_t r, int64_t N, int64_t T)
{
    int64_t i = src[r], j = dst[r], h = hod[r];
    if (i == 0) i = N;        // src/createdatamatrix.jl:9-11
    if (j == 0) j = N;        // :12-14
    if (h == 0) h = 24;       // :15-17 (hard-coded 24 in the reference)
    if (i < 1 || i > N || j < 1 || j > N || h < 1 || h > T) return -1;
    return (i - 1) + N * ((j - 1) + N * (h - 1));
}

// claim_off / dst_off: element offsets of the slab a kernel claims in / writes to (0 = mean slab, N*N*T = std slab)
__global__ void k_rows_claim(const int32_t *__restrict__ src, const int32_t *__restrict__ dst, const int32_t *__restrict__ hod,
                             int64_t n_rows, int64_t N, int64_t T, double *__restrict__ dm, int64_t claim_off,
                             unsigned long long *__restrict__ n_bad)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = row_cell(src, dst, hod, r, N, T);
        if (c < 0) { atomicAdd(n_bad, 1ull); continue; }
        atomicMax(reinterpret_cast<unsigned long long *>(dm) + claim_off + c, (unsigned long long)(r + 1));
    }
}

// the row that holds the claim of its cell writes val[r] to the cell's slot in the slab at dst_off (the slab that holds
// the claims must be written last: the value overwrites the claim)
__global__ void k_rows_write(const int32_t *__restrict__ src, const int32_t *__restrict__ dst, const int32_t *__restrict__ hod,
                             const double *__restrict__ val, int64_t n_rows, int64_t N, int64_t T, double *__restrict__ dm,
                             int64_t claim_off, int64_t dst_off)
{
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = row_cell(src, dst, hod, r, N, T);
        if (c < 0) continue;
        if (reinterpret_cast<const unsigned long long *>(dm)[claim_off + c] == (unsigned long long)(r + 1)) dm[dst_off + c] = val[r];
    }
}

int launch_rows_scatter(const int32_t *src, const int32_t *dst, const int32_t *hod, const double *mean, const double *sd,
                        int64_t n_rows, int64_t N, int64_t T, double *dm, unsigned long long *n_bad, cudaStream_t st)
{
    const int64_t slab = N * N * T;
    cudaMemsetAsync(dm, 0, (size_t)slab * 2 * sizeof(double), st);   // zeros(number_zones, number_zones, T, 2), :7
    cudaMemsetAsync(n_bad, 0, sizeof(unsigned long long), st);
    if (n_rows <= 0) return 0;
    const int g = sampler_grid(n_rows);
    k_rows_claim<<<g, 256, 0, st>>>(src, dst, hod, n_rows, N, T, dm, 0, n_bad);          // claims live in the mean slab
    k_rows_write<<<g, 256, 0, st>>>(src, dst, hod, sd, n_rows, N, T, dm, 0, slab);
    k_rows_write<<<g, 256, 0, st>>>(src, dst, hod, mean, n_rows, N, T, dm, 0, 0);
    return 3;
}

// The same scatter in two halves for the pipelined dataset pass: the means first (claims parked in the std slab, which
// nothing reads before the resampling day), the standard deviations later, when their column has arrived.
int launch_rows_scatter_means(const int32_t *src, const int32_t *dst, const int32_t *hod, const double *mean, int64_t n_rows,
                              int64_t N, int64_t T, double *dm, unsigned long long *n_bad, cudaStream_t st)
{
    const int64_t slab = N * N * T;
    if (n_rows <= 0) return 0;
    const int g = sampler_grid(n_rows);
    k_rows_claim<<<g, 256, 0, st>>>(src, dst, hod, n_rows, N, T, dm, slab, n_bad);
    k_rows_write<<<g, 256, 0, st>>>(src, dst, hod, mean, n_rows, N, T, dm, slab, 0);
    return 2;
}
int launch_rows_scatter_stds(const int32_t *src, const int32_t *dst, const int32_t *hod, const double *sd, int64_t n_rows,
                             int64_t N, int64_t T, double *dm, cudaStream_t st)
{
    const int64_t slab = N * N * T;
    if (n_rows <= 0) return 0;
    k_rows_write<<<sampler_grid(n_rows), 256, 0, st>>>(src, dst, hod, sd, n_rows, N, T, dm, slab, slab);
    return 1;
}

// Centroid distance matrix -- src/processgeodata.jl:150-166: equirectangular formula with the reference's
// constants (111.3 km per degree, 0.01745 rad per degree), symmetric, diagonal forced to 1 km.
__global__ void __launch_bounds__(256) k_distance_matrix(const double *__restrict__ lat, const double *__restrict__ lon, int64_t N,
                                                         double *__restrict__ out)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * N) return;
    const int64_t i = idx % N, j = idx / N;
    double d = 1.0;                                                         // :164-166
    if (i != j) {
        const int64_t a = i < j ? i : j, b = i < j ? j : i;                 // the reference evaluates the pair with i < j (:153-161)
        const double c = cos((lat[a] + lat[b]) / 2 * 0.01745);
        const double dl = lon[a] - lon[b], dp = lat[a] - lat[b];
        d = 111.3 * sqrt(c * c * (dl * dl) + dp * dp);                      // :158
    }
    out[idx] = d;
}
int launch_distance_matrix(const double *lat, const double *lon, int64_t N, double *out, cudaStream_t st)
{
    k_distance_matrix<<<(unsigned)((N * N + 255) / 256), 256, 0, st>>>(lat, lon, N, out);
    return 1;
}

// Polygon area and centroid -- src/processgeodata.jl:99-148.  One thread per zone walks its
// zero-terminated vertex list ([N,V] column-major: the loads of a warp are coalesced across zones) in the reference's
// order of operations: the area pass stops at the terminator and closes the ring in place, the centroid pass runs
// over all J columns without stopping at zeros.  lon/lat are the device copies (modified like the reference's).
__global__ void __launch_bounds__(128) k_polygon_centroids(double *__restrict__ lon, double *__restrict__ lat, int64_t N, int64_t V,
                                                           double *__restrict__ c_lat, double *__restrict__ c_lon,
                                                           double *__restrict__ area)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int64_t J = (V - 1 < 10000) ? V - 1 : 10000;
    double *lo = lon + i, *la = lat + i;   // vertex j (1-based) at [N * (j - 1)]
    double a = 0.0, cl = 0.0, ca = 0.0;
    const double lo1 = lo[0], la1 = la[0];
    if (lo1 != 0.0) {                                                           // :105
        double s = 0.0;
        double pl = 0.0, pa = 0.0;                                              // vertex j-1
        double xl = lo1, xa = la1;                                              // vertex j
        for (int64_t j = 1; j <= J; ++j) {                                      // :107
            if (xl == 0.0) {                                                    // :108
                s = s - (pa * xl - xa * pl);                                    // :110
                xl = lo1; xa = la1;
                lo[N * (j - 1)] = xl; la[N * (j - 1)] = xa;                     // :113-114
                s = s + (pa * xl - xa * pl);                                    // :117
                break;
            }
            const double nl = lo[N * j], na = la[N * j];                        // vertex j+1
            s = s + (xa * nl - na * xl);                                        // :120
            pl = xl; pa = xa; xl = nl; xa = na;
        }
        a = s / 2;                                                              // :123
        double sl = 0.0, sa = 0.0;
        pl = 0.0; pa = 0.0; xl = lo[0]; xa = la[0];
        for (int64_t j = 1; j <= J; ++j) {                                      // :132 (no break)
            const double nl = lo[N * j], na = la[N * j];
            if (xl == 0.0) {
                const double cr = pl * xa - xl * pa;
                sl = sl - (pl + xl) * cr;                                       // :135
                sa = sa - (pa + xa) * cr;                                       // :136
            } else {
                const double cr = xl * na - nl * xa;
                sl = sl + (xl + nl) * cr;                                       // :138
                sa = sa + (xa + na) * cr;                                       // :139
            }
            pl = xl; pa = xa; xl = nl; xa = na;
        }
        cl = -sl / (6 * a);                                                     // :142
        ca = -sa / (6 * a);                                                     // :143
    }
    area[i] = a; c_lon[i] = cl; c_lat[i] = ca;
}
int launch_polygon_centroids(double *lon, double *lat, int64_t N, int64_t V, double *c_lat, double *c_lon, double *area, cudaStream_t st)
{
    k_polygon_centroids<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(lon, lat, N, V, c_lat, c_lon, area);
    return 1;
}

}  // namespace cpm
