// bucket.cu -- the "bucket" sampler: cars are kept grouped by their current zone, so that the categorical draw of a car
// probes its origin's fp64 CDF row (and guide table) in SHARED memory, staged there by TMA bulk copies
// (cp.async.bulk + mbarrier) through a ring of row buffers.
//
// Why grouped (profiles/README.md r1a): with one thread per car in car-id order the draw is a chain of dependent,
// fully divergent L2/HBM loads; grouping by origin reads every CDF row once per step (the algorithmic table traffic,
// SURVEY 8d), turns the probes into shared-memory reads and makes the parked-car histogram free (it IS the bucket
// size).  Results are identical to the gather sampler and to the reference algorithm because every draw is keyed by
// (global car id, step), not by processing order.
//
// One step is a counting sort by destination zone, two kernels chained by programmatic dependent launch:
//   k_bucket_sample   one fat CTA per SM owns a fixed slab of slots.  Its warps walk the zone segments of the slab
//                     independently of each other (no block barrier inside the walk): per 32*U slots a warp does
//                     phases 1-3 of the reference step, ranks every car among the cars OF THIS CTA with the same
//                     destination (shared-memory histogram) and writes (destination, rank) per slot.  The warp that
//                     finishes a zone last re-arms the zone's ring buffer with the row R zones ahead.  The CTA ends
//                     with ONE global atomicAdd per destination it saw: the return value is where its cars start
//                     inside that destination's bucket (the order of cars inside a bucket is irrelevant to the
//                     results), the padded counters end up as the new bucket sizes.
//   k_bucket_scatter  CTA g re-reads slab g, sorts it by destination in shared memory (a 4-byte shared store per car),
//                     then writes it out: consecutive threads hold consecutive cars of one destination, so the global
//                     stores are runs instead of one 32-byte sector per car (profiles/r2: 16.7 -> ~7 us).  Every CTA
//                     derives the next bucket starts from the padded counters (block 0 publishes them, with the compact
//                     histogram copy); CTA g also notes which zone its slab starts in at the next step.
// Round 1's design (per-CTA histogram rows for 592 CTAs + a column-scan kernel + a sector-per-car scatter, block
// barriers around a compaction list) is in the history of this file; see profiles/README.md for the measurements.
// Reference semantics: src/resampling.jl:11-83, src/solveinitialvalueproblem.jl:11-53, src/saveresults.jl:8-17.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

#include "kernels.h"
#include "async.cuh"

namespace cpm {

// Programmatic dependent launch: the kernels of a step are launched with programmatic stream serialization, so the
// next kernel's CTAs are scheduled (and run their shared-memory prologue) while the previous kernel drains;
// pdl_wait() blocks until the previous kernel has completed and its writes are visible.  Every global access of these
// kernels comes after pdl_wait().  Both are no-ops for a plain launch.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// -DCPM_DEBUG: device-side index checks (printf + trap) for the index arithmetic of the two kernels
#ifdef CPM_DEBUG
#define CPM_CHECK(cond, fmt, ...) do { if (!(cond)) { printf("CPM_CHECK %d blk %d tid %d: " fmt "\n", __LINE__, (int)blockIdx.x, (int)threadIdx.x, __VA_ARGS__); __trap(); } } while (0)
#else
#define CPM_CHECK(cond, fmt, ...) do { } while (0)
#endif

// Loads of data the PREVIOUS kernel of the chain wrote (bucket starts, car ids, (destination, rank) words, histogram rows,
// arrival counters).  Under programmatic dependent launch this kernel's CTAs are resident before that kernel has
// finished, so such data is not read-only for the lifetime of this kernel: no ld.global.nc (__ldg), L2 is the point
// of coherence (ld.global.cg).
template <typename T> __device__ __forceinline__ T ld_prev(const T *p) { return __ldcg(p); }

constexpr int kMaxRing = 8;
constexpr size_t kSmemCap = 227 * 1024;
constexpr size_t kSampleTail = 320;      // mbarriers, ring state, group counter, reduction / scan scratch

__host__ __device__ inline size_t hist_bytes(int64_t N) { return ((size_t)(N + 1) * 4 + 15) / 16 * 16; }
__host__ __device__ inline size_t gtab_bytes(int64_t N) { return ((size_t)(N + 2) * 4 + 15) / 16 * 16; }
__host__ __device__ inline size_t zlo_bytes(int64_t N) { return ((size_t)(N + 2) * 2 + 15) / 16 * 16; }

// ------------------------------------------------------------------------------------------------
// exclusive block scan in place (array in shared memory); blockDim.x <= 1024
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
// k_bucket_sample
// ------------------------------------------------------------------------------------------------
// Shared memory: ring of R CDF row blocks | hist[N+1] | gtab[N+2] | tail (mbarriers, ring state, scratch).
//   gtab[k]   first group of the k-th bin of the slab (bins z0, z0+1, ...; a group = 32*U consecutive slots of one bin)
//   ring slot of bin ordinal k: k % R in its (k / R)-th use.  full[slot] (mbarrier, one arrival per use) completes when
//   the row has landed (or at once for a bin nobody searches); gen[slot] is the use the slot is armed for; zdone[slot]
//   counts the finished groups of that use -- the warp that finishes the last one re-arms the slot for bin k + R.
struct SampleTail {
    uint64_t full[kMaxRing];
    uint32_t zdone[kMaxRing];
    uint32_t gen[kMaxRing];
    unsigned long long red[2];
    uint32_t next_group, pad;
    uint32_t wt[34];
};
static_assert(sizeof(SampleTail) <= kSampleTail, "tail too small");

// one entry of a warp's travel-time queue: a car that drives to another zone (src/resampling.jl:61-70)
struct FarCar { int32_t id; uint32_t od; };   // od = origin | destination << 16
constexpr int kFarQueue = 64;                  // entries per warp (a power of two, >= 63: a batch is drawn as soon as 32 cars wait)

// travel-time draws of up to 32 queued cars, one per lane: (mean, std) gathered from datamatrix, second Philox call,
// inverse-CDF truncated normal, fixed-point sum.  Dense lanes whatever the share of drivers in the groups was.
template <bool INJ>
__device__ __forceinline__ void far_batch(const BucketArgs &a, const FarCar *q, uint32_t head, int n, int lane, int64_t uoff,
                                          uint64_t l2_stream, unsigned long long &t_hi, unsigned long long &t_lo)
{
    if (lane < n) {
        const FarCar e = q[(head + lane) & (kFarQueue - 1)];
        const int64_t N = a.N, NN = a.N * a.N;
        const int64_t od = (int64_t)(e.od & 0xffffu) + N * (int64_t)(e.od >> 16);
        double mean, sd;
        if (a.msd) { const double2 ms = __ldg(a.msd + od + NN * a.hour); mean = ms.x; sd = ms.y; }
        else if (a.gather_policy) {
            mean = ldg_stream(a.dm + od + NN * a.hour, l2_stream);
            sd = ldg_stream(a.dm + od + NN * (a.hour + a.T_tab), l2_stream);
        } else {
            mean = __ldg(a.dm + od + NN * a.hour);
            sd = __ldg(a.dm + od + NN * (a.hour + a.T_tab));
        }
        double u3;
        if (INJ) u3 = a.rng.u_time[e.id + uoff];
        else { uint32_t o[4]; philox_words(a.rng, e.id, (uint32_t)a.step, 1u, o); u3 = u52(o[0], o[1]); }
        if (sd == 0.0) sd = 0.1 * mean;                                   // :65-67
        fx_add(t_hi, t_lo, truncnormal(mean, sd, u3));                    // :68
    }
}

// wait until ring slot's generation word shows `use`; bounded like mbar_wait(): a logic error must trap, not hang the GPU
__device__ __forceinline__ void spin_until_gen(const uint32_t *gen, uint32_t use)
{
    for (uint32_t spin = 0; *reinterpret_cast<const volatile uint32_t *>(gen) != use; ++spin)
        if (spin > (1u << 26)) __trap();
}

template <int THREADS, int U, bool INJ, bool TIME>
__global__ void __launch_bounds__(THREADS, THREADS == 1024 ? 1 : 2) k_bucket_sample(const BucketArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int GS = 32 * U, W = THREADS / 32;
    static_assert(kFarQueue >= 64, "travel-time queue: up to 31 waiting cars + 32 new ones");
    const int64_t N = a.N;
    const int Ni = (int)N, NB = Ni + 1;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const int R = a.ring, rmask = R - 1, rshift = a.ring_shift;   // R is a power of two
    const uint32_t row_bytes = (uint32_t)(a.ldc * 8);   // CDF row + its guide table: one bulk copy
    double *const ring = reinterpret_cast<double *>(smem_raw);
    uint32_t *const hist = reinterpret_cast<uint32_t *>(smem_raw + (size_t)R * row_bytes);   // [N+1] cars of this CTA per destination
    uint32_t *const gtab = hist + hist_bytes(N) / 4;                                         // [N+2] first group of every bin of the slab
    uint16_t *const zlo = reinterpret_cast<uint16_t *>(gtab + gtab_bytes(N) / 4);            // [N+2] first slot of every bin, relative to slab_lo
    SampleTail *const tl = reinterpret_cast<SampleTail *>(reinterpret_cast<unsigned char *>(zlo) + zlo_bytes(N));
    FarCar *const farq = reinterpret_cast<FarCar *>(reinterpret_cast<unsigned char *>(tl) + kSampleTail) + (size_t)(tid >> 5) * kFarQueue;   // TIME only

    if (tid == 0) {
        for (int s = 0; s < R; ++s) mbar_init(&tl->full[s], 1);
        fence_mbar_init();
        tl->next_group = 0u;
    }
    if (tid < kMaxRing) { tl->zdone[tid] = 0u; tl->gen[tid] = 0u; }
    if (tid < 2) tl->red[tid] = 0ull;
    for (int x = tid; x < (int)(hist_bytes(N) / 16); x += THREADS) reinterpret_cast<uint4 *>(hist)[x] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    pdl_trigger();
    pdl_wait();

    const double *__restrict__ cdf_t = a.cdf + a.hour * N * a.ldc;
    const double *__restrict__ pd_t = a.p_drive + N * a.hour;
    const uint8_t *__restrict__ rz_t = a.rowzero + a.hour * N;
    const int64_t uoff = a.rng.ld * a.step;
    const uint64_t l2_stream = l2_evict_first_policy();
    unsigned long long t_hi = 0ull, t_lo = 0ull;

    // this CTA's slab of slots, and the bins it covers
    const uint32_t slab_lo = (uint32_t)min((int64_t)blockIdx.x * a.slab, a.C);
    const uint32_t slab_hi = (uint32_t)min((int64_t)(blockIdx.x + 1) * a.slab, a.C);
    const int z0 = (slab_lo < slab_hi) ? (int)ld_prev(a.first_zone + blockIdx.x) : NB;   // the bin that holds slot slab_lo
    int nz = 0;
    for (int base = 0;; base += THREADS) {   // usually one trip: a slab covers ~N / #slabs bins
        const int t = base + tid, z = z0 + t;
        uint32_t ng = 0, lo = slab_hi;
        bool valid = false;
        if (z < NB) {
            const uint32_t zs = ld_prev(a.start + z);
            if (zs < slab_hi) {
                valid = true;
                lo = max(zs, slab_lo);
                const uint32_t hi = min(ld_prev(a.start + z + 1), slab_hi);
                ng = (hi - lo + GS - 1) / GS;
            }
        }
        if (t <= NB) { gtab[t] = ng; zlo[t] = (uint16_t)(lo - slab_lo); }   // the first bin past the slab gets slab_hi: zlo[k + 1] closes bin k
        const int c = __syncthreads_count(valid);
        nz += c;
        if (c < THREADS) break;
    }
    const uint32_t n_groups = block_exclusive_scan(gtab, nz, tl->wt);
    if (tid == 0) gtab[nz] = n_groups;

    // (re-)arm ring slot `s` for the next bin of its sequence (ordinals kk, kk + R, ...) that has cars: a TMA row copy when
    // its cars may search the row, a plain arrival otherwise; bins without cars complete their use at once
    auto arm = [&](int s, int kk) {
        for (; kk < nz; kk += R) {
            if (gtab[kk + 1] == gtab[kk]) { mbar_arrive(&tl->full[s]); continue; }
            const int z = z0 + kk;
            if (z < Ni && rz_t[z] == 0) {
                fence_proxy_async();   // the buffer's last readers finished their groups before this
                mbar_expect_tx(&tl->full[s], row_bytes);
                tma_bulk_g2s(ring + (size_t)s * a.ldc, cdf_t + (int64_t)z * a.ldc, row_bytes, &tl->full[s], l2_stream);
            } else mbar_arrive(&tl->full[s]);
            break;
        }
        __threadfence_block();
        *reinterpret_cast<volatile uint32_t *>(&tl->gen[s]) = (uint32_t)(kk >> rshift);
    };
    __syncthreads();
    if (tid < R) arm(tid, tid);
    __syncthreads();

    // Groups are handed out in order through a shared counter.  A warp is always one group ahead: while it works on a
    // group it already holds the next one's number, bin and car ids (their L2 latency hides behind the current group).
    int k = 0;                 // bin ordinal cursor of this warp (groups come in increasing order)
    uint32_t q_head = 0u;      // travel-time queue of this warp (TIME)
    int q_cnt = 0;
    struct Grp { uint32_t g, x, hi; int k; int32_t id[U]; };
    auto fetch = [&](Grp &n) {
        uint32_t g = 0;
        if (lane == 0) g = atomicAdd(&tl->next_group, 1u);
        n.g = __shfl_sync(0xffffffffu, g, 0);
        if (n.g < n_groups) {
            while (n.g >= gtab[k + 1]) { ++k; CPM_CHECK(k < nz, "g %u k %d nz %d n_groups %u", n.g, k, nz, n_groups); }
            n.k = k;
            n.hi = slab_lo + zlo[k + 1];
            n.x = slab_lo + zlo[k] + (n.g - gtab[k]) * GS + lane;
            CPM_CHECK(n.hi <= slab_hi && slab_lo + zlo[k] < n.hi, "lo %u hi %u slab %u %u k %d g %u", slab_lo + zlo[k], n.hi, slab_lo, slab_hi, k, n.g);
#pragma unroll
            for (int u = 0; u < U; ++u) n.id[u] = (n.x + 32 * u < n.hi) ? ld_prev(a.perm_cur + n.x + 32 * u) : 0;
        }
    };
    Grp cur, nxt;
    fetch(cur);
    while (cur.g < n_groups) {
        fetch(nxt);
        const int kz = cur.k, z = z0 + kz;
        const uint32_t x = cur.x, hi = cur.hi;
        const bool is_zone = z < Ni;
        const double pd = is_zone ? __ldg(pd_t + z) : -1.0;
        // u <= p_drive on the 52-bit integer m of the uniform u = m * 2^-52:  m <= floor(p_drive * 2^52)  (exact: the scaling is a
        // power of two and m is an integer); p_drive >= 1 lets every car drive, a negative or NaN p_drive none (:15)
        const long long pd_m = pd >= 1.0 ? (1ll << 52) : (pd >= 0.0 ? __double2ll_rd(pd * 4503599627370496.0) : -1ll);
        const bool rowzero = is_zone ? (rz_t[z] != 0) : true;
        const int slot = kz & rmask;
        const uint32_t use = (uint32_t)(kz >> rshift);
        // ---- phase 1: Bernoulli drive decision (src/resampling.jl:11-22) ----
        double u2[U];
        uint32_t cell[U], m_stay[U];
        bool act[U], srch[U], stay[U];
        unsigned nd = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            bool drive = false;
            act[u] = x + 32 * u < hi;
            u2[u] = 0.0; cell[u] = 0u;
            if (act[u] && is_zone) {
                if (INJ) {
                    drive = a.rng.u_drive[cur.id[u] + uoff] <= pd;              // :15 (<=; NaN p_drive parks)
                    if (drive && !rowzero) u2[u] = a.rng.u_dest[cur.id[u] + uoff];
                } else {
                    uint32_t o[4];
                    philox_words(a.rng, cur.id[u], (uint32_t)a.step, 0u, o);
                    drive = (long long)(((unsigned long long)o[1] << 20) | (unsigned long long)(o[0] >> 12)) <= pd_m;
                    u2[u] = u52(o[2], o[3]);
                    cell[u] = o[3] >> a.gshift;   // floor(u2 * G): the top log2(G) bits of the 52-bit uniform
                }
            }
            srch[u] = drive && !rowzero;
            stay[u] = act[u] && !srch[u];                                       // parked, all-zero row (:35-36), or lost (bin N)
            nd += __popc(__ballot_sync(0xffffffffu, drive));
            m_stay[u] = __ballot_sync(0xffffffffu, stay[u]);
        }
        // stayers: rank among this CTA's cars bound for the own bin, one shared atomic per group
        {
            unsigned n_stay = 0;
#pragma unroll
            for (int u = 0; u < U; ++u) n_stay += __popc(m_stay[u]);
            uint32_t base = 0;
            if (lane == 0) {
                if (n_stay) base = atomicAdd(&hist[z], n_stay);
                if (is_zone && nd) atomicAdd(a.drv + z, nd);                    // src/saveresults.jl:13-15
            }
            base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (stay[u]) a.dr[x + 32 * u] = (uint32_t)z | ((base + __popc(m_stay[u] & lt_mask)) << 16);
                base += __popc(m_stay[u]);
            }
            if (TIME && rowzero && is_zone && lane == 0) t_hi += 300ull * nd;   // origin == destination: 5 min (:58-59)
        }
        // ---- phase 2: categorical destination from the staged row (src/resampling.jl:26-49): guided lower_bound,
        //      then the acceptance test of :40 (u = 0, u above the last entry or a NaN sum -> bin N = lost) ----
        bool any_srch = false;
#pragma unroll
        for (int u = 0; u < U; ++u) any_srch |= srch[u];
        if (__any_sync(0xffffffffu, any_srch)) {
            if (lane == 0) spin_until_gen(&tl->gen[slot], use);   // the slot still serves an earlier bin (rare)
            __syncwarp();
            mbar_wait(&tl->full[slot], use & 1u);   // every lane observes the completion of the row copy
            const double *row = ring + (size_t)slot * a.ldc;
            const uint16_t *gd = reinterpret_cast<const uint16_t *>(row + a.goff);
            int slo[U], shi[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                slo[u] = shi[u] = Ni;
                if (srch[u] && u2[u] > 0.0) {
                    if (INJ && !(u2[u] < 1.0)) slo[u] = gd[a.G];
                    else {
                        const uint32_t c = INJ ? (uint32_t)(u2[u] * (double)a.G) : cell[u];
                        CPM_CHECK(c < (uint32_t)a.G, "cell %u", c);
                        slo[u] = gd[c]; shi[u] = gd[c + 1];
                    }
                }
            }
            // branch-free bisection step; a finished search re-reads row[slo] (slo <= N lies inside the row block)
            auto bisect = [&]() {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool open = slo[u] < shi[u];
                    const int mid = (slo[u] + shi[u]) >> 1;
                    const bool right = row[mid] < u2[u];                                 // NaN compares false -> goes left
                    slo[u] = (open && right) ? mid + 1 : slo[u];
                    shi[u] = (open && !right) ? mid : shi[u];
                }
            };
            // a warp needs ~4 steps anyway (its widest guide cell; a lane 2.7): four steps without loop control settle every
            // cell of up to 15 entries, the loop behind them is for the rare wider one
#pragma unroll
            for (int it = 0; it < 4; ++it) bisect();
            for (;;) {
                bool any = false;
#pragma unroll
                for (int u = 0; u < U; ++u) any |= slo[u] < shi[u];
                if (!__any_sync(0xffffffffu, any)) break;
                bisect();
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                bool far = false;   // a trip to another zone: travel time from the truncated normal
                int dbin = Ni;
                if (srch[u]) {
                    const bool hit = slo[u] < Ni && u2[u] <= row[slo[u]];
                    dbin = hit ? slo[u] : Ni;
                    CPM_CHECK(dbin >= 0 && dbin <= Ni, "dbin %d slo %d shi %d", dbin, slo[u], shi[u]);
                    const uint32_t rank = atomicAdd(&hist[dbin], 1u);
                    a.dr[x + 32 * u] = (uint32_t)dbin | (rank << 16);
                    far = hit && dbin != z;
                    if (TIME && dbin == z) t_hi += 300ull;                              // :58-59
                }
                if (TIME) {   // phase 3 (time only: counts mode), src/resampling.jl:53-70: queue the car, draw in dense batches
                    const uint32_t m_far = __ballot_sync(0xffffffffu, far);
                    if (far) {
                        FarCar e; e.id = cur.id[u]; e.od = (uint32_t)z | ((uint32_t)dbin << 16);
                        farq[(q_head + q_cnt + __popc(m_far & lt_mask)) & (kFarQueue - 1)] = e;
                    }
                    q_cnt += __popc(m_far);
                    if (q_cnt >= 32) {   // a full batch: draw now (the queue never holds more than 63 cars)
                        __syncwarp();
                        far_batch<INJ>(a, farq, q_head, 32, lane, uoff, l2_stream, t_hi, t_lo);
                        q_head += 32; q_cnt -= 32;
                        __syncwarp();
                    }
                }
            }
        }
        // this group is done with the row; the warp that finishes the bin's last group re-arms the slot
        __syncwarp();
        if (lane == 0) {
            // zdone[slot] counts one use at a time: a group that never searched may get here while the slot still serves an
            // earlier bin (rare) and has to wait for its own use before it reports in
            spin_until_gen(&tl->gen[slot], use);
            __threadfence_block();
            if (atomicAdd(&tl->zdone[slot], 1u) + 1u == gtab[kz + 1] - gtab[kz]) {
                mbar_wait(&tl->full[slot], use & 1u);   // nobody may have waited for this bin's row: it must have landed before the
                tl->zdone[slot] = 0u;                   // buffer and its barrier are armed again
                arm(slot, kz + R);
            }
        }
        cur = nxt;
    }
    if (TIME) {   // the rest of the queue
        __syncwarp();
        far_batch<INJ>(a, farq, q_head, q_cnt, lane, uoff, l2_stream, t_hi, t_lo);
    }
    __syncthreads();

#ifdef CPM_DEBUG
    {   // every slot of the slab must have been ranked exactly once
        __shared__ unsigned dbg_sum;
        if (tid == 0) dbg_sum = 0;
        __syncthreads();
        unsigned part = 0;
        for (int b = tid; b < NB; b += THREADS) part += hist[b];
        atomicAdd(&dbg_sum, part);
        __syncthreads();
        if (tid == 0) CPM_CHECK(dbg_sum == slab_hi - slab_lo, "ranked %u slots of %u; z0 %d nz %d n_groups %u ring %d", dbg_sum, slab_hi - slab_lo, z0, nz, n_groups, R);
    }
#endif
    // One global atomic per destination this CTA saw: the return value is where this CTA's cars start inside the
    // destination's bucket.  The counters are one per 256 bytes (one L2 slice each: profiles/r2 micro-benchmark C); every
    // CTA starts at a different bin, so that the CTAs (which finish together) do not queue up on the same counters.
    {
        uint2 *__restrict__ Mrow = a.M + (size_t)blockIdx.x * NB;
        constexpr int B = 4;
        const int rot = (int)(((uint64_t)blockIdx.x * (uint64_t)NB) / gridDim.x);
        for (int b0 = tid; b0 < NB; b0 += B * THREADS) {
            uint32_t h[B], gb[B];
            int bb[B];
#pragma unroll
            for (int q = 0; q < B; ++q) {
                const int b = b0 + q * THREADS;
                bb[q] = b + rot < NB ? b + rot : b + rot - NB;
                h[q] = b < NB ? hist[bb[q]] : 0u;
            }
#pragma unroll
            for (int q = 0; q < B; ++q) gb[q] = h[q] ? atomicAdd(a.tot + (size_t)bb[q] * kTotStride, h[q]) : 0u;
#pragma unroll
            for (int q = 0; q < B; ++q) if (b0 + q * THREADS < NB) Mrow[bb[q]] = make_uint2(h[q], gb[q]);
        }
    }

    // driving-time total: integer block reduction, two global atomics per CTA
    if (TIME) {
        t_hi = warp_sum_u64(t_hi); t_lo = warp_sum_u64(t_lo);
        if (lane == 0) {
            if (t_hi) atomicAdd(&tl->red[0], t_hi);
            if (t_lo) atomicAdd(&tl->red[1], t_lo);
        }
        __syncthreads();
        if (tid == 0) {
            const unsigned long long hi = tl->red[0] + (tl->red[1] >> 32), lo32 = tl->red[1] & 0xffffffffull;
            if (hi) atomicAdd(a.t_hi, hi);
            if (lo32) atomicAdd(a.t_lo, lo32);
        }
    }
    (void)W;
}

// ------------------------------------------------------------------------------------------------
// exclusive block scan of two arrays at once
// ------------------------------------------------------------------------------------------------
// two arrays at once (one pass of shuffles and barriers); returns the total of v2
__device__ uint32_t block_exclusive_scan2(uint32_t *v1, uint32_t *v2, int n, uint32_t *wt /*[66]*/)
{
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const int per = (n + nt - 1) / nt;
    const int b = tid * per, e = min(b + per, n);
    uint32_t s1 = 0, s2 = 0;
    for (int k = b; k < e; ++k) { s1 += v1[k]; s2 += v2[k]; }
    uint32_t i1 = s1, i2 = s2;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y1 = __shfl_up_sync(0xffffffffu, i1, o), y2 = __shfl_up_sync(0xffffffffu, i2, o);
        if (lane >= o) { i1 += y1; i2 += y2; }
    }
    if (lane == 31) { wt[wid] = i1; wt[33 + wid] = i2; }
    __syncthreads();
    if (wid == 0) {
        const bool in = lane < (nt + 31) / 32;
        const uint32_t t1 = in ? wt[lane] : 0u, t2 = in ? wt[33 + lane] : 0u;
        uint32_t a1 = t1, a2 = t2;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y1 = __shfl_up_sync(0xffffffffu, a1, o), y2 = __shfl_up_sync(0xffffffffu, a2, o);
            if (lane >= o) { a1 += y1; a2 += y2; }
        }
        wt[lane] = a1 - t1; wt[33 + lane] = a2 - t2;
        if (lane == 31) { wt[32] = a1; wt[65] = a2; }
    }
    __syncthreads();
    uint32_t r1 = wt[wid] + i1 - s1, r2 = wt[33 + wid] + i2 - s2;
    for (int k = b; k < e; ++k) {
        const uint32_t c1 = v1[k], c2 = v2[k];
        v1[k] = r1; v2[k] = r2;
        r1 += c1; r2 += c2;
    }
    __syncthreads();
    return wt[65];
}

// largest z in [0, nb) with start[z] <= x (start non-decreasing, start[0] = 0): the bin that holds slot x
__device__ __forceinline__ int bin_of_slot(const uint32_t *start, int nb, uint32_t x)
{
    int lo = 0, hi = nb;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (start[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------------
// k_bucket_scatter: cars move to their destination bucket (state update, src/resampling.jl:81-83).
// ------------------------------------------------------------------------------------------------
template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 1024 ? 1 : 2) k_bucket_scatter(const ScatterArgs a)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int NB = a.NB, NBp = (NB + 3) & ~3;
    const int tid = threadIdx.x;
    uint32_t *const lstart = reinterpret_cast<uint32_t *>(sm_raw);   // [NBp] first position of every destination inside the sorted slab
    uint32_t *const off = lstart + NBp;                               // [NBp] bucket starts, then global position minus sorted position
    uint32_t *const wt = off + NBp;                                   // [72]
    int32_t *const sid = reinterpret_cast<int32_t *>(wt + 72);        // [slab] car ids sorted by destination
    uint16_t *const sdst = reinterpret_cast<uint16_t *>(sid + ((a.slab + 3) & ~(int64_t)3));   // [slab] their destinations
    pdl_trigger();
    pdl_wait();
    const int lo = (int)min((int64_t)blockIdx.x * a.slab, a.C), hi = (int)min((int64_t)(blockIdx.x + 1) * a.slab, a.C);
    const uint2 *__restrict__ Mrow = a.M + (size_t)blockIdx.x * NB;
    const uint32_t *__restrict__ dr = a.dr;
    const int32_t *__restrict__ pc = a.perm_cur;
    // every global load of the first batch of slots is issued up front: the scans below run under their latency
    constexpr int PS = 16;
    uint32_t v[PS];
    int32_t id[PS];
#pragma unroll
    for (int u = 0; u < PS; ++u) {
        const int sl = lo + u * THREADS + tid;
        v[u] = (sl < hi) ? ld_prev(dr + sl) : 0u;
        id[u] = (sl < hi) ? ld_prev(pc + sl) : 0;
    }
    constexpr int MQ = 4;      // bins per thread held in registers across the scan (more bins: re-read below)
    uint32_t gbq[MQ];
#pragma unroll
    for (int q = 0; q < MQ; ++q) {
        const int x = tid + q * THREADS;
        gbq[q] = 0u;
        if (x < NB) {
            const uint2 m = ld_prev(&Mrow[x]);
            lstart[x] = m.x; gbq[q] = m.y;
            off[x] = ld_prev(a.tot + (size_t)x * kTotStride);
        }
    }
    for (int x = tid + MQ * THREADS; x < NB; x += THREADS) {
        lstart[x] = ld_prev(&Mrow[x].x);
        off[x] = ld_prev(a.tot + (size_t)x * kTotStride);
    }
    __syncthreads();
    const uint32_t total = block_exclusive_scan2(lstart, off, NB, wt);
    if (blockIdx.x == 0) {   // bucket starts and sizes of the next step (the sizes are its parked-car histogram)
        for (int x = tid; x < NB; x += THREADS) {
            a.start_next[x] = off[x];
            a.cnt_next[x] = (x + 1 < NB ? off[x + 1] : total) - off[x];
            a.tot_zero[(size_t)x * kTotStride] = 0u;   // the counters of the step after this one
        }
        if (tid == 0) a.start_next[NB] = total;
    }
    // the bin that holds this slab's first slot at the next step: the largest z with start[z] <= lo
    for (int x = tid; x < NB; x += THREADS)
        if (off[x] <= (uint32_t)lo && (x + 1 == NB || off[x + 1] > (uint32_t)lo)) a.first_zone_next[blockIdx.x] = (uint32_t)x;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < MQ; ++q) { const int x = tid + q * THREADS; if (x < NB) off[x] += gbq[q] - lstart[x]; }
    for (int x = tid + MQ * THREADS; x < NB; x += THREADS) off[x] += ld_prev(&Mrow[x].y) - lstart[x];
    __syncthreads();
    // pass 1: sort the slab by destination in shared memory
    for (int s0 = lo; s0 < hi; s0 += PS * THREADS) {
        if (s0 > lo) {   // slabs larger than one register batch
#pragma unroll
            for (int u = 0; u < PS; ++u) {
                const int sl = s0 + u * THREADS + tid;
                v[u] = (sl < hi) ? ld_prev(dr + sl) : 0u;
                id[u] = (sl < hi) ? ld_prev(pc + sl) : 0;
            }
        }
#pragma unroll
        for (int u = 0; u < PS; ++u) {
            const int sl = s0 + u * THREADS + tid;
            if (sl < hi) {
                const uint32_t d = v[u] & 0xffffu, p = lstart[d] + (v[u] >> 16);
                CPM_CHECK(d < (uint32_t)NB && p < (uint32_t)(hi - lo), "slot %d d %u rank %u p %u n %d", sl, d, v[u] >> 16, p, hi - lo);
                sid[p] = id[u];
                sdst[p] = (uint16_t)d;
            }
        }
    }
    __syncthreads();
    // pass 2: consecutive threads hold consecutive cars of one destination -> the stores go out in runs
    const int n = hi - lo;
    for (int p0 = tid; p0 < n; p0 += 4 * THREADS) {   // four independent load -> load -> store chains per trip
        uint32_t d[4], pos[4];
        int32_t car[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int p = p0 + q * THREADS; d[q] = p < n ? sdst[p] : 0u; car[q] = p < n ? sid[p] : 0; }
#pragma unroll
        for (int q = 0; q < 4; ++q) pos[q] = off[d[q]] + (uint32_t)(p0 + q * THREADS);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (p0 + q * THREADS < n) {
                CPM_CHECK(pos[q] < (uint32_t)a.C, "p %d d %u pos %u", p0 + q * THREADS, d[q], pos[q]);
                a.perm_next[pos[q]] = car[q];
                if (a.zone_out) a.zone_out[car[q]] = (d[q] < (uint32_t)(NB - 1)) ? (int32_t)d[q] + 1 : 0;   // last step: car-id indexed state
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

// start[0..NB] = exclusive scan of cnt[0..NB) (+ total); first_zone[g] = the bin that holds the first slot of slab g
__global__ void __launch_bounds__(1024) k_bucket_start(const uint32_t *__restrict__ cnt, int NB, uint32_t *__restrict__ start,
                                                       uint32_t *__restrict__ first_zone, int n_slabs, int64_t slab, int64_t C)
{
    extern __shared__ uint32_t sm[];   // v[NB] | warp_tot[33]
    uint32_t *v = sm, *warp_tot = sm + NB;
    for (int x = threadIdx.x; x < NB; x += blockDim.x) v[x] = cnt[x];
    __syncthreads();
    const uint32_t total = block_exclusive_scan(v, NB, warp_tot);
    for (int x = threadIdx.x; x < NB; x += blockDim.x) start[x] = v[x];
    if (threadIdx.x == 0) start[NB] = total;
    for (int g = threadIdx.x; g < n_slabs; g += blockDim.x)
        first_zone[g] = (uint32_t)bin_of_slot(v, NB, (uint32_t)min((int64_t)g * slab, C));
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

// counts in the public layout: parking/driving uint32 [N, n_steps] from the bucket arrays with stride NB = N + 1; the
// lost-car visits of the day are the sizes of bin N over the steps (documented extension, DESIGN.md "sentinel")
__global__ void k_bucket_export_counts(const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ drv, int N, int64_t n_steps,
                                       uint32_t *__restrict__ parking, uint32_t *__restrict__ driving, unsigned long long *lost)
{
    const int64_t n = (int64_t)N * n_steps;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = k / N, z = k % N;
        parking[k] = cnt[s * (N + 1) + z];
        driving[k] = drv[s * (N + 1) + z];
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        unsigned long long v = 0ull;
        for (int64_t s = threadIdx.x; s < n_steps; s += 32) v += cnt[s * (N + 1) + N];
        v = warp_sum_u64(v);
        if (threadIdx.x == 0) *lost = v;
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
// geometry and launchers
// ------------------------------------------------------------------------------------------------
static int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

static size_t scatter_smem(int NB, int64_t slab)
{
    const size_t NBp = (size_t)((NB + 3) & ~3);
    return 2 * NBp * 4 + 72 * 4 + (size_t)((slab + 3) & ~(int64_t)3) * 4 + (size_t)((slab + 7) & ~(int64_t)7) * 2;
}

// Slabs (= CTAs of both kernels), slab size, CTA width and ring depth.  One fat CTA per SM: a slab as large as the
// 16-bit rank and the scatter kernel's shared-memory sort allow, more slabs (whole waves) only beyond that.
bool bucket_geometry(int64_t N, int64_t ldc, int64_t C, int with_time, BucketGeom *g)
{
    if (N + 1 > 65535 || C < 1 || C > 0x7fffffff) return false;
    const int NB = (int)N + 1;
    // CTA shape: 1,024 threads, one CTA per SM.  The IVP kernel (40-46 registers) may also run as two CTAs of 768 or 512
    // threads per SM (CPM_BUCKET_IVP_THREADS); the resampling kernel needs 58 registers and stays at 1,024 x 1.
    int threads = with_time ? env_int("CPM_BUCKET_THREADS", 1024) : env_int("CPM_BUCKET_IVP_THREADS", env_int("CPM_BUCKET_THREADS", 1024));
    if (threads != 512 && threads != 768) threads = 1024;
    if (with_time && threads == 768) threads = 1024;
    const int per_sm = threads == 1024 ? 1 : 2;
    const size_t cap = kSmemCap / per_sm - (per_sm > 1 ? 1024 : 0);
    // hist | gtab | zlo | tail | one travel-time queue per warp (resampling steps only)
    const size_t fixed1 = hist_bytes(N) + gtab_bytes(N) + zlo_bytes(N) + kSampleTail + (with_time ? (size_t)(threads / 32) * kFarQueue * sizeof(FarCar) : 0),
                 row = (size_t)ldc * 8;
    if (fixed1 + row > cap) return false;
    // 4 row buffers by default: deeper rings starve L1 (measured at cfg2: 8 buffers 60 us per launch, 4 buffers 49 us)
    int ring = (int)std::min<size_t>(kMaxRing, (cap - fixed1) / row);
    const int ring_env = env_int("CPM_BUCKET_RING", 4);
    if (ring_env >= 1 && ring_env < ring) ring = ring_env;
    while (ring & (ring - 1)) ring &= ring - 1;   // a power of two: slot and use of a bin are a mask and a shift
    const size_t fixed2 = scatter_smem(NB, 0) + 64;
    if (fixed2 + 6 * 256 > cap) return false;
    int64_t slab_cap = (int64_t)((cap - fixed2) / 6) & ~(int64_t)63;
    if (slab_cap > 65472) slab_cap = 65472;
    const int64_t wave = (int64_t)kSMs * per_sm;
    int64_t n = (C + 255) / 256;
    if (n > wave) n = wave;
    if (n * slab_cap < C) n = ((C + slab_cap - 1) / slab_cap + wave - 1) / wave * wave;
    g->n_slabs = (int)n;
    g->slab = (C + n - 1) / n;
    g->threads = threads;
    g->threads_scatter = per_sm == 1 ? 1024 : 512;
    g->ring = ring;
    g->smem_sample = fixed1 + (size_t)ring * row;
    g->smem_scatter = scatter_smem(NB, g->slab);
    return true;
}

bool bucket_supported(int64_t N, int64_t ldc)
{
    BucketGeom g;
    return bucket_geometry(N, ldc, 1, 1, &g) && bucket_geometry(N, ldc, 1, 0, &g);
}

// opt in to > 48 KB of dynamic shared memory.  The attribute belongs to the CURRENT device's copy of the kernel, and
// contexts on several devices and host threads share this library: what has been granted is remembered per (device,
// kernel) under a mutex, so that the steady state costs a table lookup instead of a driver call per launch (small
// cities are bound by the host's launch rate).
static cudaError_t ensure_smem_impl(const void *kernel, size_t bytes)
{
    if (bytes <= 48 * 1024) return cudaSuccess;
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, size_t> granted;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    size_t &g = granted[std::make_pair(dev, kernel)];
    if (bytes <= g) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) g = bytes;
    return e;
}
template <typename K>
static cudaError_t ensure_smem(K kernel, size_t bytes) { return ensure_smem_impl(reinterpret_cast<const void *>(kernel), bytes); }

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <int THREADS, int U, bool TIME>
static cudaError_t launch_sample_variant(const BucketArgs &a, const BucketGeom &g, cudaStream_t st)
{
    void (*k)(const BucketArgs) = a.rng.mode == 1 ? k_bucket_sample<THREADS, U, true, TIME> : k_bucket_sample<THREADS, U, false, TIME>;
    cudaError_t e = ensure_smem(k, g.smem_sample);
    if (e != cudaSuccess) return e;
    return launch_pdl(k, (unsigned)g.n_slabs, THREADS, g.smem_sample, st, a);
}

int launch_bucket_sample(const BucketArgs &a, const BucketGeom &g, cudaStream_t st)
{
    // cars per lane and group: 4 in the IVP steps by default (four independent Philox / search chains per lane), 2 in the
    // resampling steps (the travel-time code needs the registers) and in the two-CTAs-per-SM IVP shapes
    static const int U_ivp = env_int("CPM_BUCKET_U", 4) == 2 ? 2 : 4;
    cudaError_t e;
    static const int U_time = env_int("CPM_BUCKET_U_TIME", 2) == 4 ? 4 : 2;
    if (a.with_time) e = g.threads == 512 ? launch_sample_variant<512, 2, true>(a, g, st)
                       : (U_time == 4 ? launch_sample_variant<1024, 4, true>(a, g, st) : launch_sample_variant<1024, 2, true>(a, g, st));
    else if (g.threads == 768) e = launch_sample_variant<768, 2, false>(a, g, st);
    else if (g.threads == 512) e = launch_sample_variant<512, 2, false>(a, g, st);
    else e = U_ivp == 2 ? launch_sample_variant<1024, 2, false>(a, g, st) : launch_sample_variant<1024, 4, false>(a, g, st);
    return e == cudaSuccess ? 1 : -1;
}

int launch_bucket_scatter(const ScatterArgs &a, const BucketGeom &g, cudaStream_t st)
{
    cudaError_t e;
    if (g.threads_scatter == 512) {
        e = ensure_smem(k_bucket_scatter<512>, g.smem_scatter);
        if (e == cudaSuccess) e = launch_pdl(k_bucket_scatter<512>, (unsigned)g.n_slabs, 512, g.smem_scatter, st, a);
    } else {
        e = ensure_smem(k_bucket_scatter<1024>, g.smem_scatter);
        if (e == cudaSuccess) e = launch_pdl(k_bucket_scatter<1024>, (unsigned)g.n_slabs, 1024, g.smem_scatter, st, a);
    }
    return e == cudaSuccess ? 1 : -1;
}

int launch_bucketize(const int32_t *zone, int64_t C, int NB, uint32_t *cnt0, uint32_t *start, uint32_t *cursor_tmp,
                     int32_t *perm, uint32_t *first_zone, const BucketGeom &g, cudaStream_t st)
{
    if (cudaMemsetAsync(cnt0, 0, (size_t)NB * sizeof(uint32_t), st) != cudaSuccess) return -1;
    if (cudaMemsetAsync(cursor_tmp, 0, (size_t)NB * sizeof(uint32_t), st) != cudaSuccess) return -1;
    const int grid = sampler_grid(C);
    k_bucket_hist<<<grid, 256, 0, st>>>(zone, C, NB, cnt0);
    const size_t smem = ((size_t)NB + 40) * sizeof(uint32_t);
    if (ensure_smem(k_bucket_start, smem) != cudaSuccess) return -1;
    k_bucket_start<<<1, 1024, smem, st>>>(cnt0, NB, start, first_zone, g.n_slabs, g.slab, C);
    k_bucket_place<<<grid, 256, 0, st>>>(zone, C, NB, start, cursor_tmp, perm);
    return 3;
}

int launch_bucket_export_counts(const uint32_t *cnt, const uint32_t *drv, int N, int64_t n_steps, uint32_t *parking,
                                uint32_t *driving, unsigned long long *lost, cudaStream_t st)
{
    k_bucket_export_counts<<<sampler_grid((int64_t)N * n_steps), 256, 0, st>>>(cnt, drv, N, n_steps, parking, driving, lost);
    return 1;
}

int launch_finalize_time(unsigned long long *hi, unsigned long long *lo, double *out, cudaStream_t st)
{
    k_finalize_time<<<1, 32, 0, st>>>(hi, lo, out);
    return 1;
}

}  // namespace cpm
