// atomics_bench.cu -- round-2 microbenchmarks behind the redesign of the sampling step (profiles/README.md r2).
// Each kernel processes 2.4 M "cars" (cfg2) with a cheap integer hash as the random source, so that the measured time
// is the cost of the memory operation under test, full chip, CUDA events, 20 launches after 3 warm-ups.
//   A  shared-memory atomicAdd with return, random bin in [0, 2401)          (rank inside a CTA)
//   B  the same without using the return value                                (histogram only)
//   C  global atomicAdd with return on 2,401 counters, unpadded / one per 256 B, uniform / skewed destinations
//   D  global 4-byte scatter store to a random position of a 9.6 MB array     (what k_bucket_scatter does)
//   E  the same positions written in runs of ~7 consecutive slots per warp    (CTA-local sort first)
//   F  4 dependent random 8-byte shared-memory loads per car                  (the guided CDF search)
//   G  shared-memory scatter store (4-byte, random slot of a 64 KB array)      (the CTA-local sort itself)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomics_bench.bin atomics_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
// destination bin: uniform over NB, or skewed (a quarter of the cars go to 16 hot bins)
__device__ __forceinline__ uint32_t dest_of(uint32_t i, uint32_t salt, int NB, int skew)
{
    const uint32_t h = hash32(i * 2654435761u + salt);
    if (skew && (h & 3u) == 0u) return (h >> 8) & 15u;
    return (uint32_t)(((uint64_t)h * (uint64_t)NB) >> 32);
}

constexpr int NB = 2401;

template <bool USE_RET>
__global__ void k_smem_atomic(int C, uint32_t salt, int skew, uint32_t *out, uint32_t *Mrow)
{
    __shared__ uint32_t hist[NB];
    for (int x = threadIdx.x; x < NB; x += blockDim.x) hist[x] = 0;
    __syncthreads();
    const int per = (C + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per, hi = min(C, lo + per);
    uint32_t acc = 0;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const uint32_t d = dest_of(i, salt, NB, skew);
        if (USE_RET) { const uint32_t r = atomicAdd(&hist[d], 1u); out[i] = d | (r << 16); }
        else { atomicAdd(&hist[d], 1u); acc += d; }
    }
    __syncthreads();
    for (int x = threadIdx.x; x < NB; x += blockDim.x) Mrow[(size_t)blockIdx.x * NB + x] = hist[x];
    if (!USE_RET && acc == 0xffffffffu) out[0] = acc;
}

__global__ void k_gmem_atomic(int C, uint32_t salt, int skew, int stride, uint32_t *cnt, uint32_t *out)
{
    const int per = (C + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per, hi = min(C, lo + per);
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const uint32_t d = dest_of(i, salt, NB, skew);
        const uint32_t r = atomicAdd(&cnt[(size_t)d * stride], 1u);
        out[i] = r;
    }
}

__global__ void k_scatter_random(int C, uint32_t salt, const int32_t *ids, int32_t *perm_next)
{
    const int per = (C + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per, hi = min(C, lo + per);
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const uint32_t p = (uint32_t)(((uint64_t)hash32(i * 2654435761u + salt) * (uint64_t)C) >> 32);
        perm_next[p] = ids[i];
    }
}
// runs: slot i of a CTA goes to base(run) + offset, run = i / RUN (consecutive threads -> consecutive addresses inside a run)
__global__ void k_scatter_runs(int C, uint32_t salt, int RUN, const int32_t *ids, int32_t *perm_next)
{
    const int per = (C + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per, hi = min(C, lo + per);
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const int run = i / RUN, o = i - run * RUN;
        uint32_t p = (uint32_t)(((uint64_t)hash32(run * 2654435761u + salt) * (uint64_t)(C - RUN)) >> 32);
        perm_next[p + o] = ids[i];
    }
}

__global__ void k_smem_probe(int C, uint32_t salt, const double *row_g, int n_row, double *out)
{
    extern __shared__ double row[];
    for (int x = threadIdx.x; x < n_row; x += blockDim.x) row[x] = row_g[x];
    __syncthreads();
    const int per = (C + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per, hi = min(C, lo + per);
    double acc = 0.0;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        uint32_t j = hash32(i + salt) % (uint32_t)n_row;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double v = row[j];
            acc += v;
            j = (j + 1u + (uint32_t)(v > 0.5)) % (uint32_t)n_row;   // dependent on the loaded value
        }
    }
    if (acc == 123.456) out[0] = acc;
}

__global__ void k_smem_scatter(int C, uint32_t salt, const int32_t *ids, int32_t *out)
{
    extern __shared__ int32_t buf[];   // 16,384 slots
    const int per = (C + gridDim.x - 1) / gridDim.x;
    const int lo = blockIdx.x * per, hi = min(C, lo + per);
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) buf[hash32(i + salt) & 16383u] = ids[i];
    __syncthreads();
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) out[i] = buf[(i - lo) & 16383];
}

template <typename F>
static float time_it(const char *name, F f, int reps = 20)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) f(w);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) f(100 + r);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-64s %8.2f us per launch\n", name, 1000.f * ms / reps);
    fflush(stdout);
    return ms / reps;
}

int main()
{
    const int C = 2400000;
    uint32_t *out, *M, *cnt;
    int32_t *ids, *perm;
    double *row_g, *outd;
    CK(cudaMalloc(&out, (size_t)C * 4)); CK(cudaMalloc(&M, (size_t)1184 * NB * 4)); CK(cudaMalloc(&cnt, (size_t)NB * 64 * 4));
    CK(cudaMalloc(&ids, (size_t)C * 4)); CK(cudaMalloc(&perm, (size_t)C * 4 + 4096));
    CK(cudaMalloc(&row_g, 2560 * 8)); CK(cudaMalloc(&outd, 64));
    CK(cudaMemset(ids, 1, (size_t)C * 4)); CK(cudaMemset(row_g, 0, 2560 * 8)); CK(cudaMemset(cnt, 0, (size_t)NB * 64 * 4));
    char name[128];
    const int shapes[3][2] = { {148, 1024}, {296, 512}, {592, 256} };
    for (int skew = 0; skew < 2; ++skew)
        for (auto &s : shapes) {
            snprintf(name, sizeof name, "A smem atomicAdd+ret   grid %4d x %4d  skew %d", s[0], s[1], skew);
            time_it(name, [&](int it) { k_smem_atomic<true><<<s[0], s[1]>>>(C, it, skew, out, M); });
            snprintf(name, sizeof name, "B smem atomicAdd noret grid %4d x %4d  skew %d", s[0], s[1], skew);
            time_it(name, [&](int it) { k_smem_atomic<false><<<s[0], s[1]>>>(C, it, skew, out, M); });
        }
    for (int skew = 0; skew < 2; ++skew)
        for (int stride : {1, 8, 64}) {
            snprintf(name, sizeof name, "C gmem atomicAdd+ret   grid 592 x 256  stride %3d B skew %d", stride * 4, skew);
            time_it(name, [&](int it) { k_gmem_atomic<<<592, 256>>>(C, it, skew, stride, cnt, out); });
        }
    // one atomic per (CTA, bin): 148 / 592 CTAs x 2,401 bins
    for (auto &s : shapes) {
        snprintf(name, sizeof name, "D random 4-byte global scatter grid %4d x %4d", s[0], s[1]);
        time_it(name, [&](int it) { k_scatter_random<<<s[0], s[1]>>>(C, it, ids, perm); });
    }
    for (int run : {2, 4, 7, 16, 32}) {
        snprintf(name, sizeof name, "E global scatter in runs of %2d, grid 148 x 1024", run);
        time_it(name, [&](int it) { k_scatter_runs<<<148, 1024>>>(C, it, run, ids, perm); });
    }
    CK(cudaFuncSetAttribute(k_smem_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 2560 * 8));
    for (auto &s : shapes) {
        snprintf(name, sizeof name, "F 4 dependent smem 8-byte probes per car grid %4d x %4d", s[0], s[1]);
        time_it(name, [&](int it) { k_smem_probe<<<s[0], s[1], 2560 * 8>>>(C, it, row_g, 2400, outd); });
    }
    CK(cudaFuncSetAttribute(k_smem_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 4));
    time_it("G smem 4-byte scatter + coalesced write-out grid 148 x 1024", [&](int it) { k_smem_scatter<<<148, 1024, 16384 * 4>>>(C, it, ids, (int32_t *)out); });
    printf("done\n");
    return 0;
}
