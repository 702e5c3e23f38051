// bucket.cu -- the "bucket" sampler: cars are kept grouped by their current zone, so that a CTA works on one
// zone at a time with the origin's fp64 CDF row (and its guide table) staged in shared memory by a TMA bulk copy
// (cp.async.bulk + mbarrier).
//
// Why (profiles/README.md r1a): with one thread per car in car-id order the categorical draw is a chain of
// dependent L2/HBM loads under divergence; grouping by origin turns it into shared-memory probes, reads every CDF
// row once per step (the algorithmic table traffic), and makes the per-zone parked-car histogram free
// (it IS the bucket size).  Results are identical to the gather sampler and to the reference algorithm because
// every draw is keyed by (global car id, step), not by processing order.
//
// One step is a counting sort by destination zone with per-CTA histograms (no global atomics on the data path):
//   k_bucket_sample   CTA g owns a fixed slab of slots [g*S, (g+1)*S); it walks the zone segments inside its slab,
//                     does phases 1-3 of the reference step per car, ranks every car among the cars OF THIS CTA
//                     with the same destination (shared-memory histogram), writes (destination, rank) per slot
//                     and finally its histogram row M[g][*]
//   k_bucket_colscan  exclusive prefix of M over g per destination (-> where CTA g's cars start inside each
//                     destination bucket) and the new bucket sizes (= parked cars per zone at the next step)
//   k_bucket_scatter  prefix sum of the new bucket sizes, then perm_next[start[dest] + M[g][dest] + rank] = car
// Reference semantics: src/resampling.jl:11-83, src/solveinitialvalueproblem.jl:11-53, src/saveresults.jl:8-17.
#include "kernels.h"
#include "async.cuh"

namespace cpm {

// Programmatic dependent launch: the three kernels of a step are launched with programmatic stream serialization, so
// the next kernel's CTAs are scheduled (and run their shared-memory prologue) while the previous kernel drains;
// pdl_wait() blocks until the previous kernel has completed and its writes are visible.  Every global access of these
// kernels comes after pdl_wait().  Both are no-ops for a plain launch.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// shared-memory layout of k_bucket_sample
// ------------------------------------------------------------------------------------------------
struct SampleSmem {
    double *row;          // [ldc]   the origin's CDF row + guide table (TMA destination)
    double *srch_u;       // [CH]    u_dest of the compacted drivers
    uint32_t *hist;       // [N+1]   cars of this CTA per destination bin (never cleared inside a launch)
    uint16_t *srch_idx;   // [CH]    chunk-relative slot of each compacted driver
    uint64_t *bar;
    uint32_t *misc;       // [0],[1] n_search ping-pong counters, [2] scratch of the zone search
};

__host__ __device__ inline size_t hist_bytes(int64_t N) { return ((size_t)(N + 1) * 4 + 15) / 16 * 16; }
__host__ __device__ inline size_t sample_smem_bytes(int64_t N, int64_t ldc, int CH)
{
    return (size_t)ldc * 8 + (size_t)CH * 8 + hist_bytes(N) + (size_t)CH * 2 + 64;
}

__device__ __forceinline__ SampleSmem carve(unsigned char *p, int64_t N, int64_t ldc, int CH)
{
    SampleSmem s;
    s.row = reinterpret_cast<double *>(p); p += (size_t)ldc * 8;
    s.srch_u = reinterpret_cast<double *>(p); p += (size_t)CH * 8;
    s.hist = reinterpret_cast<uint32_t *>(p); p += hist_bytes(N);
    s.srch_idx = reinterpret_cast<uint16_t *>(p); p += (size_t)CH * 2;
    s.bar = reinterpret_cast<uint64_t *>(p); p += 16;
    s.misc = reinterpret_cast<uint32_t *>(p);
    return s;
}

// largest z in [0, nb) with start[z] <= x  (start is non-decreasing, start[0] = 0, start[nb] = C > x)
__device__ __forceinline__ int zone_of_slot(const uint32_t *__restrict__ start, int nb, uint32_t x)
{
    int lo = 0, hi = nb;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(start + mid) <= x) lo = mid; else hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------------
// k_bucket_sample
// ------------------------------------------------------------------------------------------------
template <int THREADS, int CH, bool INJ>   // INJ: uniform variates injected by the caller (parity tests) instead of Philox
__global__ void __launch_bounds__(THREADS) k_bucket_sample(const BucketArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int R = CH / THREADS;
    const int64_t N = a.N, NN = a.N * a.N;
    const int Ni = (int)N, NB = Ni + 1;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    SampleSmem s = carve(smem_raw, N, a.ldc, CH);
    const uint32_t row_bytes = (uint32_t)(a.ldc * 8);   // CDF row + its guide table in one bulk copy
    const double *__restrict__ cdf_t = a.cdf + a.hour * N * a.ldc;
    const double *__restrict__ pd_t = a.p_drive + N * a.hour;
    const uint8_t *__restrict__ rz_t = a.rowzero + a.hour * N;
    constexpr bool inj = INJ;
    const int64_t uoff = a.rng.ld * a.step;
    uint32_t *const hist = s.hist;

    if (tid == 0) { mbar_init(&s.bar[0], 1); fence_mbar_init(); }
    for (int x = tid; x < (int)(hist_bytes(N) / 16); x += THREADS) reinterpret_cast<uint4 *>(hist)[x] = make_uint4(0u, 0u, 0u, 0u);
    if (tid < 3) s.misc[tid] = 0u;
    __syncthreads();
    pdl_trigger();
    pdl_wait();

    uint32_t par = 0u;
    unsigned long long t_hi = 0ull, t_lo = 0ull, lost = 0ull;
    const uint64_t l2_stream = l2_evict_first_policy();
    int pp = 0;   // which n_search counter the current chunk uses

    // this CTA's slab of slots
    const uint32_t slab_lo = (uint32_t)min((int64_t)blockIdx.x * a.slab, a.C);
    const uint32_t slab_hi = (uint32_t)min((int64_t)(blockIdx.x + 1) * a.slab, a.C);
    uint32_t pos = slab_lo;
    // pull the slab's car ids towards L2 now (one 128-byte line per thread and trip); they are read chunk by chunk below
    for (uint32_t x = slab_lo + (uint32_t)tid * 32u; x < slab_hi; x += THREADS * 32u)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.perm_cur + x));
    // the slab's first zone, z = #{b : start[b] <= slab_lo} - 1, by one coalesced pass over the bucket starts (a binary
    // search would be a chain of ~12 dependent L2 loads at the head of every CTA).  Shared memory is NOT used to stage
    // the starts: the L1 it would take is worth more to the travel-time gathers of the resampling steps.
    {
        unsigned below = 0;
        for (int x = tid; x < NB; x += THREADS) below += (__ldg(a.start + x) <= slab_lo) ? 1u : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
        if (lane == 0 && below) atomicAdd(&s.misc[2], below);
    }
    __syncthreads();
    int z = (slab_lo < slab_hi) ? (int)s.misc[2] - 1 : NB;

    while (pos < slab_hi) {
        // zone segment [pos, seg_end) of bin z inside the slab; skip empty zones (usually none)
        uint32_t z_end = __ldg(a.start + z + 1);
        while (z_end <= pos) z_end = __ldg(a.start + (++z) + 1);   // start[NB] = C > pos: terminates
        const uint32_t seg_end = min(z_end, slab_hi);
        const int bin = z;
        const bool is_zone = bin < Ni;
        const double pd = is_zone ? __ldg(pd_t + bin) : -1.0;
        const bool rowzero = is_zone ? (rz_t[bin] != 0) : true;
        const bool need_row = is_zone && !rowzero;
        unsigned nd_warp = 0;

        for (uint32_t lo = pos; lo < seg_end; lo += CH) {
            const int n = (int)min((uint32_t)CH, seg_end - lo);
            const bool first_chunk = lo == pos;
            if (need_row && first_chunk && tid == 0) {   // the row buffer's last readers passed the barrier that ended phase B
                fence_proxy_async();
                mbar_expect_tx(&s.bar[0], row_bytes);
                tma_bulk_g2s(s.row, cdf_t + (int64_t)bin * a.ldc, row_bytes, &s.bar[0], l2_stream);
            }
            const int32_t *__restrict__ perm = a.perm_cur + lo;
            uint32_t *__restrict__ out = a.dr + lo;
            uint32_t *const n_search_ctr = &s.misc[pp];

            // ---- phase A: Bernoulli drive decision for every car of the chunk (src/resampling.jl:11-22) ----
            int32_t ids[R];
#pragma unroll
            for (int r = 0; r < R; ++r) { const int k = r * THREADS + tid; ids[r] = (is_zone && k < n) ? __ldg(perm + k) : 0; }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (r * THREADS >= n) break;          // block-uniform: the ballots below stay convergent
                const int k = r * THREADS + tid;
                const bool active = k < n;
                bool drive = false;
                double u2 = 0.0;
                if (active && is_zone) {
                    const int64_t id = ids[r];
                    double u1;
                    if (inj) u1 = a.rng.u_drive[id + uoff];
                    else philox_pair(a.rng, id, (uint32_t)a.step, 0u, u1, u2);
                    drive = u1 <= pd;                                   // :15 (<=; NaN p_drive parks)
                    if (drive && inj && !rowzero) u2 = a.rng.u_dest[id + uoff];
                }
                const bool search = drive && !rowzero;
                const bool stay = active && !search;                    // parked, all-zero row (:35-36), or lost (zone 0)
                const uint32_t m_drive = __ballot_sync(0xffffffffu, drive);
                const uint32_t m_stay = __ballot_sync(0xffffffffu, stay);
                const uint32_t m_s = m_drive & ~m_stay;                 // == ballot(search)
                nd_warp += __popc(m_drive);
                // stayers: rank among this CTA's cars bound for the own zone, one warp-aggregated shared atomic;
                // searchers: compact (slot, u_dest) so that phase B runs without idle lanes
                uint32_t base = 0, off = 0;
                if (lane == 0) {
                    if (m_stay) base = atomicAdd(&hist[bin], (uint32_t)__popc(m_stay));
                    if (m_s) off = atomicAdd(n_search_ctr, (uint32_t)__popc(m_s));
                }
                base = __shfl_sync(0xffffffffu, base, 0);
                off = __shfl_sync(0xffffffffu, off, 0);
                if (stay) out[k] = (uint32_t)bin | ((base + __popc(m_stay & lt_mask)) << 16);
                if (search) {
                    const uint32_t q = off + __popc(m_s & lt_mask);
                    s.srch_idx[q] = (uint16_t)k;
                    s.srch_u[q] = u2;
                }
            }
            if (tid == 0) s.misc[pp ^ 1] = 0u;   // the other counter was consumed before the previous barrier
            __syncthreads();

            // ---- phase B: categorical destination from the staged CDF row (src/resampling.jl:26-49) ----
            const int n_search = (int)*n_search_ctr;
            if (need_row) {
                if (first_chunk) { mbar_wait(&s.bar[0], par & 1u); par ^= 1u; }   // every thread observes the TMA completion
                const double *row = s.row;
                const uint16_t *gd = reinterpret_cast<const uint16_t *>(s.row + a.goff);
                for (int q = tid; q < n_search; q += THREADS) {
                    const int k = s.srch_idx[q];
                    const double u2 = s.srch_u[q];
                    // guided lower_bound on the staged row; acceptance test as in src/resampling.jl:40 (u = 0, u above the
                    // last entry or a NaN sum -> sentinel 0 = lost)
                    const int32_t d = cdf_search_guided([&](int32_t j) { return row[j]; }, [&](int32_t bb) { return (int32_t)gd[bb]; },
                                                        Ni, a.G, u2);
                    const bool hit = d > 0;
                    const int dbin = hit ? d - 1 : Ni;
                    const uint32_t rank = atomicAdd(&hist[dbin], 1u);
                    out[k] = (uint32_t)dbin | (rank << 16);
                    if (a.with_time) {                                   // src/resampling.jl:53-70 (time only: counts mode)
                        if (dbin == bin) t_hi += 300ull;
                        else if (hit) {
                            const int64_t id = __ldg(perm + k);
                            double u3, u4;
                            if (inj) u3 = a.rng.u_time[id + uoff];
                            else philox_pair(a.rng, id, (uint32_t)a.step, 1u, u3, u4);
                            const int64_t od = (int64_t)bin + N * (int64_t)dbin;
                            double mean, sd;
                            if (a.msd) { const double2 ms = __ldg(a.msd + od + NN * a.hour); mean = ms.x; sd = ms.y; }
                            else { mean = ldg_stream(a.dm + od + NN * a.hour, l2_stream); sd = ldg_stream(a.dm + od + NN * (a.hour + a.T_tab), l2_stream); }
                            if (sd == 0.0) sd = 0.1 * mean;
                            fx_add(t_hi, t_lo, truncnormal(mean, sd, u3));
                        }
                    }
                }
            }
            __syncthreads();
            pp ^= 1;
        }
        if (is_zone) {
            if (rowzero && a.with_time && lane == 0) t_hi += 300ull * nd_warp;   // origin == destination: 5 min (:58-59)
            if (lane == 0 && nd_warp) atomicAdd(a.drv + bin, nd_warp);             // src/saveresults.jl:13-15
        } else if (tid == 0) lost += (unsigned long long)(seg_end - pos);
        pos = seg_end;
        ++z;
    }

    // this CTA's histogram row: M[g][bin] = cars of slab g bound for bin
    uint32_t *__restrict__ Mrow = a.M + (size_t)blockIdx.x * NB;
    for (int x = tid; x < NB; x += THREADS) Mrow[x] = hist[x];

    // driving-time and lost-car totals: integer block reduction, a few global atomics per CTA
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
// k_bucket_colscan: M[g][b] <- sum_{g' < g} M[g'][b] (exclusive, in place); cnt_next[b] <- sum_g M[g][b].
// One CTA per 32 destination columns: 32 x 32 threads, each sums / rescans a stripe of slabs; coalesced 128 B rows.
// ------------------------------------------------------------------------------------------------
template <int PER, int COLS, int PARTS>   // PER >= ceil(G / PARTS): the stripe of one thread lives in registers (one pass over M)
__global__ void __launch_bounds__(COLS * PARTS) k_bucket_colscan(uint32_t *__restrict__ M, int G, int NB, uint32_t *__restrict__ cnt_next)
{
    __shared__ uint32_t ps[PARTS][COLS + 1];
    pdl_trigger();
    pdl_wait();
    const int col = threadIdx.x % COLS, part = threadIdx.x / COLS;
    const int c = blockIdx.x * COLS + col;
    const int per = (G + PARTS - 1) / PARTS;
    const int g0 = min(G, part * per), g1 = min(G, g0 + per);
    const bool live = c < NB;
    uint32_t v[PER];
    uint32_t sum = 0;
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const int g = g0 + u;
        v[u] = (live && g < g1) ? M[(size_t)g * NB + c] : 0u;
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) sum += v[u];
    ps[part][col] = sum;
    __syncthreads();
    uint32_t run = 0;
    for (int p = 0; p < part; ++p) run += ps[p][col];
    if (live) {
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int g = g0 + u;
            if (g < g1) M[(size_t)g * NB + c] = run;
            run += v[u];
        }
        if (part == PARTS - 1) cnt_next[c] = run;   // the last stripe ends with the column total
    }
}

// general fallback (any G): two passes over the stripe
__global__ void __launch_bounds__(1024) k_bucket_colscan_any(uint32_t *__restrict__ M, int G, int NB, uint32_t *__restrict__ cnt_next)
{
    __shared__ uint32_t ps[32][33];
    pdl_trigger();
    pdl_wait();
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), part = threadIdx.x >> 5;
    const int per = (G + 31) / 32;
    const int g0 = min(G, part * per), g1 = min(G, g0 + per);
    const bool live = c < NB;
    uint32_t sum = 0;
    if (live) for (int g = g0; g < g1; ++g) sum += M[(size_t)g * NB + c];
    ps[part][threadIdx.x & 31] = sum;
    __syncthreads();
    uint32_t run = 0;
    for (int p = 0; p < part; ++p) run += ps[p][threadIdx.x & 31];
    if (live) {
        for (int g = g0; g < g1; ++g) {
            const size_t k = (size_t)g * NB + c;
            const uint32_t v = M[k];
            M[k] = run;
            run += v;
        }
        if (part == 31) cnt_next[c] = run;
    }
}

// ------------------------------------------------------------------------------------------------
// exclusive block scan of v[0..n) in place (v in shared memory); returns the total.  blockDim.x <= 1024.
// ------------------------------------------------------------------------------------------------
__device__ uint32_t block_exclusive_scan(uint32_t *v, int n, uint32_t *warp_tot /*[33]*/)
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

// ------------------------------------------------------------------------------------------------
// k_bucket_scatter: cars move to their destination bucket (state update, src/resampling.jl:81-83).
// CTA g handles the slab CTA g of k_bucket_sample handled: off[b] = start_next[b] + M[g][b].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bucket_scatter(const ScatterArgs a)
{
    extern __shared__ uint32_t sm[];   // off[NB] | mrow[NB] | warp_tot[33]
    const int NB = a.NB;
    uint32_t *off = sm, *mrow = sm + NB, *warp_tot = sm + 2 * NB;
    pdl_trigger();
    pdl_wait();
    const uint32_t *__restrict__ Mrow = a.M + (size_t)blockIdx.x * NB;
    const int lo = (int)min((int64_t)blockIdx.x * a.slab, a.C), hi = (int)min((int64_t)(blockIdx.x + 1) * a.slab, a.C);
    const uint32_t *__restrict__ dr = a.dr;
    const int32_t *__restrict__ pc = a.perm_cur;
    // every global load this CTA needs is issued up front (slot data of the first batch, bucket sizes, its M row),
    // so the prefix sum below runs under their latency
    constexpr int PS = 16, PRE = 4;
    uint32_t v[PS];
    int32_t id[PS];
#pragma unroll
    for (int u = 0; u < PS; ++u) {
        const int sl = lo + u * 256 + threadIdx.x;
        v[u] = (sl < hi) ? __ldg(dr + sl) : 0u;
        id[u] = (sl < hi) ? __ldg(pc + sl) : 0;
    }
    for (int x0 = 0; x0 < NB; x0 += PRE * 256) {
        uint32_t c[PRE], m[PRE];
#pragma unroll
        for (int u = 0; u < PRE; ++u) {
            const int x = x0 + u * 256 + threadIdx.x;
            c[u] = (x < NB) ? __ldg(a.cnt_next + x) : 0u;
            m[u] = (x < NB) ? __ldg(Mrow + x) : 0u;
        }
#pragma unroll
        for (int u = 0; u < PRE; ++u) { const int x = x0 + u * 256 + threadIdx.x; if (x < NB) { off[x] = c[u]; mrow[x] = m[u]; } }
    }
    __syncthreads();
    const uint32_t total = block_exclusive_scan(off, NB, warp_tot);
    if (blockIdx.x == 0) {   // bucket starts of the next step, for the next k_bucket_sample
        for (int x = threadIdx.x; x < NB; x += blockDim.x) a.start_next[x] = off[x];
        if (threadIdx.x == 0) a.start_next[NB] = total;
    }
    for (int x = threadIdx.x; x < NB; x += blockDim.x) off[x] += mrow[x];
    __syncthreads();
    for (int s0 = lo; s0 < hi; s0 += PS * 256) {
        if (s0 > lo) {   // slabs larger than one register batch
#pragma unroll
            for (int u = 0; u < PS; ++u) {
                const int sl = s0 + u * 256 + threadIdx.x;
                v[u] = (sl < hi) ? __ldg(dr + sl) : 0u;
                id[u] = (sl < hi) ? __ldg(pc + sl) : 0;
            }
        }
#pragma unroll
        for (int u = 0; u < PS; ++u) {
            const int sl = s0 + u * 256 + threadIdx.x;
            if (sl < hi) {
                const int db = (int)(v[u] & 0xffffu);
                a.perm_next[off[db] + (v[u] >> 16)] = id[u];
                if (a.zone_out) a.zone_out[id[u]] = (db < NB - 1) ? db + 1 : 0;   // last step: car-id indexed state
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// bucketize a car-id indexed zone array (day start)
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

// start[0..NB] = exclusive scan of cnt[0..NB) (+ total)
__global__ void __launch_bounds__(1024) k_bucket_start(const uint32_t *__restrict__ cnt, int NB, uint32_t *__restrict__ start)
{
    extern __shared__ uint32_t sm[];   // v[NB] | warp_tot[33]
    uint32_t *v = sm, *warp_tot = sm + NB;
    for (int x = threadIdx.x; x < NB; x += blockDim.x) v[x] = cnt[x];
    __syncthreads();
    const uint32_t total = block_exclusive_scan(v, NB, warp_tot);
    for (int x = threadIdx.x; x < NB; x += blockDim.x) start[x] = v[x];
    if (threadIdx.x == 0) start[NB] = total;
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

// canonical form (whole seconds, fraction < 1 s in 2^-32 units), then ONE conversion to fp64: the same value for any
// decomposition of the work into CTAs, samplers or GPUs
__global__ void k_finalize_time(unsigned long long *hi, unsigned long long *lo, double *out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const unsigned long long h = *hi + (*lo >> 32), l = *lo & 0xffffffffull;
        *hi = h; *lo = l;
        *out = (double)h + (double)l * (1.0 / 4294967296.0);
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
// Two shapes: small CTAs (256 threads, chunks of 1,024 cars) so that four fit on an SM when the CDF row is
// small (N up to ~3,000), otherwise 512 threads / 2,048 cars with as many CTAs as shared memory allows.
static int bucket_shape(int64_t N, int64_t ldc, size_t *smem_out, int *ctas_out)
{
    const size_t small = sample_smem_bytes(N, ldc, 1024), big = sample_smem_bytes(N, ldc, 2048);
    const size_t sm_total = 227 * 1024;
    if (4 * (small + 1024) <= sm_total) { *smem_out = small; *ctas_out = 4; return 1; }   // 4 x 256 x 64 regs fill the register file
    if (big + 1024 <= sm_total) { *smem_out = big; *ctas_out = (2 * (big + 1024) <= sm_total) ? 2 : 1; return 2; }
    return 0;
}

bool bucket_supported(int64_t N, int64_t ldc)
{
    size_t smem; int ctas;
    return N + 1 <= 65535 && bucket_shape(N, ldc, &smem, &ctas) != 0;
}

// number of slabs (= CTAs of k_bucket_sample / k_bucket_scatter) and slab size: one resident wave, more only when a
// slab would exceed the 16-bit rank range
void bucket_grid(int64_t N, int64_t ldc, int64_t C, int *G_out, int64_t *slab_out)
{
    size_t smem = 0; int ctas = 1;
    bucket_shape(N, ldc, &smem, &ctas);
    int64_t G = (int64_t)kSMs * ctas;
    const int64_t min_G = (C + 65534) / 65535;
    if (G < min_G) G = min_G;
    if (G > C) G = C > 0 ? C : 1;
    *G_out = (int)G;
    *slab_out = (C + G - 1) / G;
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

template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, args...);
}

int launch_bucket_sample(const BucketArgs &a, int G, cudaStream_t st)
{
    size_t smem = 0; int ctas = 0;
    const int shape = bucket_shape(a.N, a.ldc, &smem, &ctas);
    static size_t granted[4] = {0, 0, 0, 0};
    const bool inj = a.rng.mode == 1;
    if (shape == 1) {
        if (inj) { ensure_smem(k_bucket_sample<256, 1024, true>, smem, granted[0]); launch_pdl(k_bucket_sample<256, 1024, true>, G, 256, smem, st, a); }
        else     { ensure_smem(k_bucket_sample<256, 1024, false>, smem, granted[1]); launch_pdl(k_bucket_sample<256, 1024, false>, G, 256, smem, st, a); }
    } else if (shape == 2) {
        if (inj) { ensure_smem(k_bucket_sample<512, 2048, true>, smem, granted[2]); launch_pdl(k_bucket_sample<512, 2048, true>, G, 512, smem, st, a); }
        else     { ensure_smem(k_bucket_sample<512, 2048, false>, smem, granted[3]); launch_pdl(k_bucket_sample<512, 2048, false>, G, 512, smem, st, a); }
    } else return -1;
    return 1;
}

int launch_bucket_colscan(uint32_t *M, int G, int NB, uint32_t *cnt_next, cudaStream_t st)
{
    // 8 columns x 64 stripes per CTA: ~NB/8 CTAs (two per SM at Melbourne scale) instead of NB/32 on half of the SMs
    const int per64 = (G + 63) / 64, per32 = (G + 31) / 32;
    if (per64 <= 10) launch_pdl(k_bucket_colscan<10, 8, 64>, (NB + 7) / 8, 512, 0, st, M, G, NB, cnt_next);
    else if (per32 <= 40) launch_pdl(k_bucket_colscan<40, 32, 32>, (NB + 31) / 32, 1024, 0, st, M, G, NB, cnt_next);
    else launch_pdl(k_bucket_colscan_any, (NB + 31) / 32, 1024, 0, st, M, G, NB, cnt_next);
    return 1;
}

static size_t scan_smem(int NB) { return ((size_t)2 * NB + 40) * sizeof(uint32_t); }

int launch_bucket_scatter(const ScatterArgs &a, int G, cudaStream_t st)
{
    static size_t granted = 48 * 1024;
    ensure_smem(k_bucket_scatter, scan_smem(a.NB), granted);
    launch_pdl(k_bucket_scatter, G, 256, scan_smem(a.NB), st, a);
    return 1;
}

int launch_bucketize(const int32_t *zone, int64_t C, int NB, uint32_t *cnt0, uint32_t *start, uint32_t *cursor_tmp,
                     int32_t *perm, cudaStream_t st)
{
    cudaMemsetAsync(cnt0, 0, (size_t)NB * sizeof(uint32_t), st);
    cudaMemsetAsync(cursor_tmp, 0, (size_t)NB * sizeof(uint32_t), st);
    const int g = sampler_grid(C);
    k_bucket_hist<<<g, 256, 0, st>>>(zone, C, NB, cnt0);
    static size_t granted = 48 * 1024;
    ensure_smem(k_bucket_start, scan_smem(NB), granted);
    k_bucket_start<<<1, 1024, scan_smem(NB), st>>>(cnt0, NB, start);
    k_bucket_place<<<g, 256, 0, st>>>(zone, C, NB, start, cursor_tmp, perm);
    return 3;
}

int launch_bucket_export_counts(const uint32_t *cnt, const uint32_t *drv, int N, int64_t n_steps, uint32_t *parking,
                                uint32_t *driving, cudaStream_t st)
{
    k_bucket_export_counts<<<sampler_grid((int64_t)N * n_steps), 256, 0, st>>>(cnt, drv, N, n_steps, parking, driving);
    return 1;
}

int launch_finalize_time(unsigned long long *hi, unsigned long long *lo, double *out, cudaStream_t st)
{
    k_finalize_time<<<1, 32, 0, st>>>(hi, lo, out);
    return 1;
}

}  // namespace cpm
