// bucket.cu -- the "bucket" sampler: cars are kept grouped by their current zone, one CTA work item = one
// zone (or a chunk of a large zone), and the origin's fp64 CDF row is staged in shared memory with a TMA
// bulk copy (cp.async.bulk + mbarrier), double-buffered across work items.
//
// Why (profiles/README.md r1a): with one thread per car in car-id order the 12-probe binary search is a chain
// of dependent L2/HBM loads under divergence; grouping by origin turns it into shared-memory probes, reads
// every CDF row once per step (the algorithmic table traffic), and makes the per-zone parked-car
// histogram free (it IS the bucket size).  Results are identical to the gather sampler and to the reference
// algorithm because every draw is keyed by (global car id, step), not by processing order.
//
// Per step:  k_bucket_sample  (phases 1-3 per car; per-item shared-memory histogram of destinations with
//                              warp-aggregated atomics; one global atomic per (item, destination) reserves a
//                              range in the destination bucket)
//            k_bucket_scatter (prefix-sum of the new bucket sizes, cars move to their new bucket; block 0 also
//                              writes the next step's work-item list)
// Reference semantics: src/resampling.jl:11-83, src/solveinitialvalueproblem.jl:11-53, src/saveresults.jl:8-17.
#include "kernels.h"

namespace cpm {

// ------------------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy helpers (PTX; sm_90+ instructions, compiled for sm_100a)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait suspends in hardware for a bounded time per call; the spin bound turns a lost TMA completion
// (which would otherwise hang the GPU) into a trap the host sees as a launch failure.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 24)) __trap();
}

// ------------------------------------------------------------------------------------------------
// shared-memory layout of k_bucket_sample
// ------------------------------------------------------------------------------------------------
struct SampleSmem {
    double *row;          // [ldc]   the origin's CDF row (TMA destination)
    double *srch_u;       // [CH]    u_dest of the compacted drivers
    uint32_t *hist;       // [N+1]   per-item destination counts, then bucket bases
    uint16_t *dest_s;     // [CH]    destination bin per car of the item
    uint16_t *rank_s;     // [CH]    rank among the item's cars with the same destination
    uint16_t *srch_idx;   // [CH]    slot of each compacted driver
    uint16_t *touched;    // [CH+2]  destination bins with a non-zero count (so per-item work is O(cars), not O(N))
    uint64_t *bar;
    uint32_t *misc;       // [0] n_search, [1] n_drive, [2] n_touched
};

__host__ __device__ inline size_t hist_bytes(int64_t N) { return ((size_t)(N + 1) * 4 + 15) / 16 * 16; }
__host__ __device__ inline size_t sample_smem_bytes(int64_t N, int64_t ldc, int CH)
{
    return (size_t)ldc * 8 + (size_t)CH * 8 + hist_bytes(N) + (size_t)CH * 2 * 3 + (size_t)(CH + 8) * 2 + 64;
}

__device__ __forceinline__ SampleSmem carve(unsigned char *p, int64_t N, int64_t ldc, int CH)
{
    SampleSmem s;
    s.row = reinterpret_cast<double *>(p); p += (size_t)ldc * 8;
    s.srch_u = reinterpret_cast<double *>(p); p += (size_t)CH * 8;
    s.hist = reinterpret_cast<uint32_t *>(p); p += hist_bytes(N);
    s.dest_s = reinterpret_cast<uint16_t *>(p); p += (size_t)CH * 2;
    s.rank_s = reinterpret_cast<uint16_t *>(p); p += (size_t)CH * 2;
    s.srch_idx = reinterpret_cast<uint16_t *>(p); p += (size_t)CH * 2;
    s.touched = reinterpret_cast<uint16_t *>(p); p += (size_t)(CH + 8) * 2;
    s.bar = reinterpret_cast<uint64_t *>(p); p += 16;
    s.misc = reinterpret_cast<uint32_t *>(p);
    return s;
}

// ------------------------------------------------------------------------------------------------
// k_bucket_sample.  One work item = up to CH cars of one zone.  Several small CTAs per SM run items in
// different phases, which is what hides the three latencies on an item's critical path (car-id load, CDF-row
// TMA, the global atomics that reserve ranges in the destination buckets).
// ------------------------------------------------------------------------------------------------
template <int THREADS, int CH>
__global__ void __launch_bounds__(THREADS) k_bucket_sample(const BucketArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int R = CH / THREADS;
    const int64_t N = a.N, NN = a.N * a.N;
    const int Ni = (int)N;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    SampleSmem s = carve(smem_raw, N, a.ldc, CH);
    const uint32_t row_bytes = (uint32_t)(a.ldc * 8);   // CDF row + its guide table in one bulk copy
    const double *__restrict__ cdf_t = a.cdf + a.hour * N * a.ldc;
    const double *__restrict__ pd_t = a.p_drive + N * a.hour;
    const uint8_t *__restrict__ rz_t = a.rowzero + a.hour * N;
    const bool inj = a.rng.mode == 1;
    const int64_t uoff = a.rng.ld * a.step;
    const int n_items = *a.n_items;

    if (tid == 0) { mbar_init(&s.bar[0], 1); fence_mbar_init(); }
    for (int x = tid; x < (int)(hist_bytes(N) / 16); x += THREADS) reinterpret_cast<uint4 *>(s.hist)[x] = make_uint4(0u, 0u, 0u, 0u);
    if (tid < 4) s.misc[tid] = 0u;
    __syncthreads();

    uint32_t par = 0u;
    unsigned long long t_hi = 0ull, t_lo = 0ull, lost = 0ull;

    uint32_t *const hist = s.hist;
    uint16_t *const touched = s.touched;
    uint32_t *const n_touched = &s.misc[2];
    int prev_nt = 0;   // #bins the previous item touched (every thread keeps its own copy: no extra barrier)

    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        const int4 it = __ldg(a.items + w);   // x = bin (0..N-1 zone, N = lost), y = first slot, z = #cars
        const int bin = it.x, lo = it.y, n = it.z;
        const bool is_zone = bin < Ni;
        if (is_zone && tid == 0) {   // the row buffer was last read before the previous item's phase-B barrier
            fence_proxy_async();
            mbar_expect_tx(&s.bar[0], row_bytes);
            tma_bulk_g2s(s.row, cdf_t + (int64_t)bin * a.ldc, row_bytes, &s.bar[0]);
        }
        const int32_t *__restrict__ perm = a.perm_cur + lo;

        // car ids of all rounds first, so their load latencies overlap each other and the clean-up below
        int32_t ids[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { const int k = r * THREADS + tid; ids[r] = (is_zone && k < n) ? __ldg(perm + k) : 0; }
        // clear the histogram bins the previous item touched (O(cars), not O(N))
        for (int i = tid; i < prev_nt; i += THREADS) hist[touched[i]] = 0u;
        __syncthreads();

        if (!is_zone) {
            // lost cars (zone 0) are carried along: never drive, never counted in a zone
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int k = r * THREADS + tid;
                if (k < n) { s.dest_s[k] = (uint16_t)bin; s.rank_s[k] = (uint16_t)k; }
            }
            if (tid == 0) { hist[bin] = (uint32_t)n; touched[0] = (uint16_t)bin; *n_touched = 1u; lost += (unsigned long long)n; }
            __syncthreads();
        } else {
            const double pd = __ldg(pd_t + bin);
            const bool rowzero = rz_t[bin] != 0;
            unsigned nd_warp = 0;
            // ---- phase A: Bernoulli drive decision for every car of the item (src/resampling.jl:11-22) ----
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r * THREADS >= n) break;          // block-uniform: the ballots below stay convergent
                const int k = r * THREADS + tid;
                const bool active = k < n;
                bool drive = false;
                double u2 = 0.0;
                if (active) {
                    const int64_t id = ids[r];
                    double u1;
                    if (inj) u1 = a.rng.u_drive[id + uoff];
                    else philox_pair(a.rng, id, (uint32_t)a.step, 0u, u1, u2);
                    drive = u1 <= pd;                                   // :15 (<=; NaN p_drive parks)
                    if (drive && inj && !rowzero) u2 = a.rng.u_dest[id + uoff];
                }
                const bool search = drive && !rowzero;
                const bool stay = active && !search;                    // parked, or all-zero row (:35-36)
                const uint32_t m_drive = __ballot_sync(0xffffffffu, drive);
                const uint32_t m_stay = __ballot_sync(0xffffffffu, stay);
                const uint32_t m_s = m_drive & ~m_stay;                 // == ballot(search)
                nd_warp += __popc(m_drive);
                // stayers: rank inside the own-zone bin via one warp-aggregated shared atomic;
                // searchers: compact (slot, u_dest) so that phase B runs without idle lanes
                uint32_t base = 0, off = 0;
                if (lane == 0) {
                    if (m_stay) {
                        base = atomicAdd(&hist[bin], (uint32_t)__popc(m_stay));
                        if (base == 0u) touched[atomicAdd(n_touched, 1u)] = (uint16_t)bin;
                    }
                    if (m_s) off = atomicAdd(&s.misc[0], (uint32_t)__popc(m_s));
                }
                base = __shfl_sync(0xffffffffu, base, 0);
                off = __shfl_sync(0xffffffffu, off, 0);
                if (stay) { s.dest_s[k] = (uint16_t)bin; s.rank_s[k] = (uint16_t)(base + __popc(m_stay & lt_mask)); }
                if (search) {
                    const uint32_t q = off + __popc(m_s & lt_mask);
                    s.srch_idx[q] = (uint16_t)k;
                    s.srch_u[q] = u2;
                }
            }
            if (rowzero && a.with_time && lane == 0) t_hi += 300ull * nd_warp;   // origin == destination: 5 min (:58-59)
            if (lane == 0 && nd_warp) atomicAdd(&s.misc[1], nd_warp);
            __syncthreads();

            // ---- phase B: categorical destination from the staged CDF row (src/resampling.jl:26-49) ----
            const int n_search = (int)s.misc[0];
            mbar_wait(&s.bar[0], par & 1u);   // every thread observes the TMA completion (phase parity stays in step)
            par ^= 1u;
            const double *row = s.row;
            const uint16_t *gd = reinterpret_cast<const uint16_t *>(s.row + a.goff);
            for (int q0 = 0; q0 < n_search; q0 += THREADS) {   // block-uniform trip count (ballot below)
                const int q = q0 + tid;
                const bool act = q < n_search;
                const int k = act ? s.srch_idx[q] : 0;
                const double u2 = act ? s.srch_u[q] : 0.0;
                // guided lower_bound on the staged row; acceptance test as in src/resampling.jl:40 (u = 0, u above the
                // last entry or a NaN sum -> sentinel 0 = lost)
                const int32_t d = cdf_search_guided([&](int32_t j) { return row[j]; }, [&](int32_t bb) { return (int32_t)gd[bb]; },
                                                    Ni, a.G, u2);
                const bool hit = d > 0;
                const int pos = d - 1;
                const int dbin = hit ? pos : Ni;
                const uint32_t old = act ? atomicAdd(&hist[dbin], 1u) : 1u;
                // first toucher of a bin appends it to the touched list: one shared atomic per warp, not per lane
                const uint32_t m_new = __ballot_sync(0xffffffffu, old == 0u);
                if (m_new) {
                    uint32_t tbase = 0;
                    if (lane == 0) tbase = atomicAdd(n_touched, (uint32_t)__popc(m_new));
                    tbase = __shfl_sync(0xffffffffu, tbase, 0);
                    if (old == 0u) touched[tbase + __popc(m_new & lt_mask)] = (uint16_t)dbin;
                }
                if (!act) continue;
                s.dest_s[k] = (uint16_t)dbin;
                s.rank_s[k] = (uint16_t)old;
                if (a.with_time) {                                   // src/resampling.jl:53-70 (time only: counts mode)
                    if (dbin == bin) t_hi += 300ull;
                    else if (hit) {
                        const int64_t id = __ldg(perm + k);
                        double u3, u4;
                        if (inj) u3 = a.rng.u_time[id + uoff];
                        else philox_pair(a.rng, id, (uint32_t)a.step, 1u, u3, u4);
                        const int64_t od = (int64_t)bin + N * (int64_t)dbin;
                        const double mean = __ldg(a.dm + od + NN * a.hour);
                        double sd = __ldg(a.dm + od + NN * (a.hour + a.T_tab));
                        if (sd == 0.0) sd = 0.1 * mean;
                        fx_add(t_hi, t_lo, truncnormal(mean, sd, u3));
                    }
                }
            }
            __syncthreads();
        }

        // ---- phase C: reserve a range in every touched destination bucket (one global atomic per item x destination) ----
        {
            const int nt = (int)*n_touched;
            prev_nt = nt;
            // batches of CB atomics per thread are issued back to back so that their round trips overlap
            constexpr int CB = 4;
            for (int i0 = 0; i0 < nt; i0 += THREADS * CB) {
                int b[CB];
                uint32_t c[CB];
#pragma unroll
                for (int u = 0; u < CB; ++u) {
                    const int i = i0 + u * THREADS + tid;
                    b[u] = (i < nt) ? (int)touched[i] : -1;
                    c[u] = (b[u] >= 0) ? hist[b[u]] : 0u;
                }
#pragma unroll
                for (int u = 0; u < CB; ++u) if (b[u] >= 0) c[u] = atomicAdd(a.cnt_next + b[u], c[u]);
#pragma unroll
                for (int u = 0; u < CB; ++u) if (b[u] >= 0) hist[b[u]] = c[u];
            }
            if (tid == 0 && is_zone && s.misc[1]) atomicAdd(a.drv + bin, s.misc[1]);   // src/saveresults.jl:13-15
        }
        __syncthreads();

        // ---- phase D: destination bin and position inside it, per slot (coalesced) ----
        if (tid == 0) { s.misc[0] = 0u; s.misc[1] = 0u; *n_touched = 0u; }   // all were consumed before the barrier above
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int k = r * THREADS + tid;
            if (k < n) {
                const int db = s.dest_s[k];
                a.dest[lo + k] = (uint16_t)db;
                a.pos[lo + k] = hist[db] + s.rank_s[k];
            }
        }
        __syncthreads();
    }

    // driving-time and lost-car totals: integer block reduction, two global atomics per CTA
    __shared__ unsigned long long red[3];
    if (tid < 3) red[tid] = 0ull;
    __syncthreads();
    t_hi = warp_sum_u64(t_hi); t_lo = warp_sum_u64(t_lo); lost = warp_sum_u64(lost);
    if (lane == 0) {
        if (t_hi) atomicAdd(&red[0], t_hi);
        if (t_lo) atomicAdd(&red[1], t_lo);
        if (lost) atomicAdd(&red[2], lost);
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned long long hi = red[0] + (red[1] >> 32), lo32 = red[1] & 0xffffffffull;
        if (hi) atomicAdd(a.t_hi, hi);
        if (lo32) atomicAdd(a.t_lo, lo32);
        if (red[2]) atomicAdd(a.lost, red[2]);
    }
}

// ------------------------------------------------------------------------------------------------
// prefix sum of bucket sizes + work-item list (device function shared by k_bucket_items and k_bucket_scatter)
// ------------------------------------------------------------------------------------------------
// exclusive block scan of v[0..n) in place (v in shared memory); returns the total.  blockDim.x <= 1024.
__device__ uint32_t block_exclusive_scan(uint32_t *v, int n, uint32_t *warp_tot /*[32]*/)
{
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const int per = (n + nt - 1) / nt;
    const int b = tid * per, e = min(b + per, n);
    uint32_t sum = 0;
    for (int k = b; k < e; ++k) sum += v[k];
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t t = (lane < (nt + 31) / 32) ? warp_tot[lane] : 0u;
        uint32_t ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += y;
        }
        warp_tot[lane] = ti - t;   // exclusive warp offsets
        if (lane == 31) warp_tot[32] = ti;
    }
    __syncthreads();
    uint32_t run = warp_tot[wid] + inc - sum;
    for (int k = b; k < e; ++k) { const uint32_t c = v[k]; v[k] = run; run += c; }
    __syncthreads();
    return warp_tot[32];
}

// start[] (shared, NB+1 entries incl. total) must already hold the exclusive scan of cnt[]
__device__ void write_items(const uint32_t *__restrict__ cnt, const uint32_t *start, uint32_t *itoff, uint32_t *warp_tot,
                            int NB, uint32_t chunk, int4 *__restrict__ items, int32_t *__restrict__ n_items)
{
    for (int x = threadIdx.x; x < NB; x += blockDim.x) itoff[x] = (cnt[x] + chunk - 1) / chunk;
    __syncthreads();
    const uint32_t total = block_exclusive_scan(itoff, NB, warp_tot);
    for (int x = threadIdx.x; x < NB; x += blockDim.x) {
        const uint32_t c = cnt[x];
        uint32_t o = itoff[x];
        for (uint32_t j = 0; j * chunk < c; ++j)
            items[o + j] = make_int4(x, (int)(start[x] + j * chunk), (int)min(chunk, c - j * chunk), 0);
    }
    if (threadIdx.x == 0) *n_items = (int32_t)total;
}

// day start / stand-alone: start[] and the item list from cnt[]
__global__ void __launch_bounds__(1024) k_bucket_items(const uint32_t *__restrict__ cnt, int NB, uint32_t chunk,
                                                       uint32_t *__restrict__ start_out, int4 *__restrict__ items,
                                                       int32_t *__restrict__ n_items)
{
    extern __shared__ uint32_t sm[];   // start[NB] | itoff[NB] | warp_tot[33]
    uint32_t *start = sm, *itoff = sm + NB, *warp_tot = sm + 2 * NB;
    for (int x = threadIdx.x; x < NB; x += blockDim.x) start[x] = cnt[x];
    __syncthreads();
    block_exclusive_scan(start, NB, warp_tot);
    if (start_out) for (int x = threadIdx.x; x < NB; x += blockDim.x) start_out[x] = start[x];
    write_items(cnt, start, itoff, warp_tot, NB, chunk, items, n_items);
}

// ------------------------------------------------------------------------------------------------
// k_bucket_scatter: cars move to their destination bucket (state update, src/resampling.jl:81-83)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_bucket_scatter(const ScatterArgs a)
{
    extern __shared__ uint32_t sm[];   // start[NB] | itoff[NB] (block 0) | warp_tot[33]
    const int NB = a.NB;
    uint32_t *start = sm, *itoff = sm + NB, *warp_tot = sm + 2 * NB;
    for (int x = threadIdx.x; x < NB; x += blockDim.x) start[x] = a.cnt_next[x];
    __syncthreads();
    block_exclusive_scan(start, NB, warp_tot);
    for (int64_t sl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; sl < a.C; sl += (int64_t)gridDim.x * blockDim.x) {
        const int db = a.dest[sl];
        const int32_t id = a.perm_cur[sl];
        a.perm_next[start[db] + a.pos[sl]] = id;
        if (a.zone_out) a.zone_out[id] = (db < NB - 1) ? db + 1 : 0;   // last step: car-id indexed state
    }
    if (blockIdx.x == 0 && a.items) {
        __syncthreads();
        write_items(a.cnt_next, start, itoff, warp_tot, NB, (uint32_t)a.chunk, a.items, a.n_items);
    }
}

// ------------------------------------------------------------------------------------------------
// bucketize a car-id indexed zone array (day start) and its inverse
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int zone_bin(int32_t z, int NB) { return (z >= 1 && z < NB) ? z - 1 : NB - 1; }

__global__ void __launch_bounds__(256) k_bucket_hist(const int32_t *__restrict__ zone, int64_t C, int NB, uint32_t *__restrict__ cnt)
{
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < C; i0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = i0 + threadIdx.x;
        const bool act = i < C;
        const int b = act ? zone_bin(zone[i], NB) : -1;
        const uint32_t grp = __match_any_sync(0xffffffffu, b);       // lanes with the same bin
        if (act && (threadIdx.x & 31) == (__ffs(grp) - 1)) atomicAdd(cnt + b, (uint32_t)__popc(grp));
    }
}

__global__ void __launch_bounds__(256) k_bucket_place(const int32_t *__restrict__ zone, int64_t C, int NB,
                                                      const uint32_t *__restrict__ start, uint32_t *__restrict__ cursor,
                                                      int32_t *__restrict__ perm)
{
    const uint32_t lt = (1u << (threadIdx.x & 31)) - 1u;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < C; i0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = i0 + threadIdx.x;
        const bool act = i < C;
        const int b = act ? zone_bin(zone[i], NB) : -1;
        const uint32_t grp = __match_any_sync(0xffffffffu, b);
        const int leader = __ffs(grp) - 1;
        uint32_t base = 0;
        if (act && (threadIdx.x & 31) == leader) base = atomicAdd(cursor + b, (uint32_t)__popc(grp));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (act) perm[start[b] + base + __popc(grp & lt)] = (int32_t)i;
    }
}

// fleet_init directly in bucketed form: car i sits in zone i/cpz + 1 -> perm = identity, cnt = histogram of that map
__global__ void __launch_bounds__(256) k_bucket_init_sorted(int32_t *__restrict__ perm, int64_t C)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < C; i += (int64_t)gridDim.x * blockDim.x) perm[i] = (int32_t)i;
}

// counts in the public layout: parking/driving uint32 [N, n_steps] from the bucket arrays with stride NB = N + 1
__global__ void k_bucket_export_counts(const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ drv, int N, int64_t n_steps,
                                       uint32_t *__restrict__ parking, uint32_t *__restrict__ driving)
{
    const int64_t n = (int64_t)N * n_steps;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = k / N, z = k % N;
        parking[k] = cnt[s * (N + 1) + z];
        driving[k] = drv[s * (N + 1) + z];
    }
}

__global__ void k_finalize_time(const unsigned long long *hi, const unsigned long long *lo, double *out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = (double)(*hi) + (double)(*lo) * (1.0 / 4294967296.0);
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
// Two shapes: small CTAs (256 threads, items of 1,024 cars) so that four fit on an SM when the CDF row is small
// (N up to ~3,000), otherwise 512 threads / 2,048 cars with as many CTAs as shared memory allows.
static int bucket_shape(int64_t N, int64_t ldc, size_t *smem_out, int *ctas_out)
{
    const size_t small = sample_smem_bytes(N, ldc, 1024), big = sample_smem_bytes(N, ldc, 2048);
    const size_t sm_total = 227 * 1024;
    if (4 * (small + 1024) <= sm_total) { *smem_out = small; *ctas_out = (int)(sm_total / (small + 1024)); return 1; }
    if (big + 1024 <= sm_total) { *smem_out = big; *ctas_out = (int)(sm_total / (big + 1024)); return 2; }
    return 0;
}

int bucket_chunk(int64_t N, int64_t ldc)
{
    size_t smem; int ctas;
    const int shape = bucket_shape(N, ldc, &smem, &ctas);
    return shape == 1 ? 1024 : shape == 2 ? 2048 : 0;
}

bool bucket_supported(int64_t N, int64_t ldc)
{
    return N + 1 <= 65535 && bucket_chunk(N, ldc) != 0;
}

// opt in to > 48 KB of dynamic shared memory, raising the per-kernel limit only when a launch needs more
template <typename K>
static void ensure_smem(K kernel, size_t bytes, size_t &granted)
{
    if (bytes > granted) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        granted = bytes;
    }
}

int launch_bucket_sample(const BucketArgs &a, cudaStream_t st)
{
    size_t smem = 0; int ctas = 0;
    const int shape = bucket_shape(a.N, a.ldc, &smem, &ctas);
    static size_t granted[3] = {0, 0, 0};
    if (shape == 1) {
        if (ctas > 8) ctas = 8;
        ensure_smem(k_bucket_sample<256, 1024>, smem, granted[1]);
        k_bucket_sample<256, 1024><<<kSMs * ctas, 256, smem, st>>>(a);
    } else if (shape == 2) {
        if (ctas > 4) ctas = 4;
        ensure_smem(k_bucket_sample<512, 2048>, smem, granted[2]);
        k_bucket_sample<512, 2048><<<kSMs * ctas, 512, smem, st>>>(a);
    } else return -1;
    return 1;
}

static size_t scan_smem(int NB) { return ((size_t)2 * NB + 40) * sizeof(uint32_t); }

int launch_bucket_items(const uint32_t *cnt, int NB, int chunk, uint32_t *start_out, int4 *items, int32_t *n_items, cudaStream_t st)
{
    static size_t granted = 48 * 1024;
    ensure_smem(k_bucket_items, scan_smem(NB), granted);
    k_bucket_items<<<1, 1024, scan_smem(NB), st>>>(cnt, NB, (uint32_t)chunk, start_out, items, n_items);
    return 1;
}

int launch_bucket_scatter(const ScatterArgs &a, cudaStream_t st)
{
    static size_t granted = 48 * 1024;
    ensure_smem(k_bucket_scatter, scan_smem(a.NB), granted);
    int64_t blocks = (a.C + 1023) / 1024;
    if (blocks > kSMs * 2) blocks = kSMs * 2;
    k_bucket_scatter<<<(unsigned)blocks, 1024, scan_smem(a.NB), st>>>(a);
    return 1;
}

int launch_bucketize(const int32_t *zone, int64_t C, int NB, int chunk, uint32_t *cnt0, uint32_t *start_tmp, uint32_t *cursor_tmp,
                     int32_t *perm, int4 *items, int32_t *n_items, cudaStream_t st)
{
    cudaMemsetAsync(cnt0, 0, (size_t)NB * sizeof(uint32_t), st);
    cudaMemsetAsync(cursor_tmp, 0, (size_t)NB * sizeof(uint32_t), st);
    const int g = sampler_grid(C);
    k_bucket_hist<<<g, 256, 0, st>>>(zone, C, NB, cnt0);
    launch_bucket_items(cnt0, NB, chunk, start_tmp, items, n_items, st);
    k_bucket_place<<<g, 256, 0, st>>>(zone, C, NB, start_tmp, cursor_tmp, perm);
    return 3;
}

int launch_bucket_export_counts(const uint32_t *cnt, const uint32_t *drv, int N, int64_t n_steps, uint32_t *parking,
                                uint32_t *driving, cudaStream_t st)
{
    k_bucket_export_counts<<<sampler_grid((int64_t)N * n_steps), 256, 0, st>>>(cnt, drv, N, n_steps, parking, driving);
    return 1;
}

int launch_finalize_time(const unsigned long long *hi, const unsigned long long *lo, double *out, cudaStream_t st)
{
    k_finalize_time<<<1, 32, 0, st>>>(hi, lo, out);
    return 1;
}

}  // namespace cpm
