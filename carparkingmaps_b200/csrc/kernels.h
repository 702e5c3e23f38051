// kernels.h -- internal launcher interface between capi.cu and the kernel translation units.
// All pointers are DEVICE pointers; every launcher enqueues on `st` and returns the number of
// kernels it launched (summed into the context's launch counter).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

namespace cpm {

// ---- tables.cu --------------------------------------------------------------------------------
// mean_sum[N,T] (src/createpdrive.jl:10-21)
int launch_pdrive_meansum(const double *dm, const double *dist, double *rdist, int64_t N, int64_t T, double *mean_sum,
                          cudaStream_t st);
// p_drive[N,T] from mean_sum (src/createpdrive.jl:22-34)
int launch_pdrive_finish(const double *mean_sum, int64_t N, int64_t T, double p_min, double p_max,
                         double e_drive, double *p_drive, cudaStream_t st);
// p_dest[N,N,T] (src/createpdestin.jl:3-50); norm is [N,T] scratch
int launch_pdestin(const double *dm, int64_t N, int64_t T, double e_dest, double *p_dest, double *norm,
                   cudaStream_t st);
// cdf[T][N][N] (destination fastest) + rowzero[T][N] from p_dest[N,N,T]: left-to-right fp64 running sum
int launch_cdf_from_pdest(const double *p_dest, int64_t N, int64_t T, double *cdf, int64_t ldc, int goff, int G,
                          uint8_t *rowzero, cudaStream_t st);

// fused: p_dest /= norm (where norm > 0) and the CDF build in one pass (src/createpdestin.jl:38-46)
int launch_cdf_normalise(double *p_dest, const double *norm, int64_t N, int64_t T, double *cdf, int64_t ldc, int goff,
                         int G, uint8_t *rowzero, int write_p, cudaStream_t st);

// ---- sampler.cu -------------------------------------------------------------------------------
struct StepArgs {
    // tables
    const double *p_drive;   // [N,T]
    const double *cdf;       // [T][N] row blocks of ldc doubles: N CDF entries (destination fastest), padding to
    int64_t ldc;             //        goff doubles, then the row's guide table (G+1 uint16); ldc*8 is a multiple of 16
    int32_t goff, G;         // guide offset inside a row block (in doubles) and guide resolution (0 = no guide)
    const uint8_t *rowzero;  // [T][N]
    const double *dm;        // [N,N,T,2] (with_time only)
    const double2 *msd;      // nullable: [N,N,T] (mean, std) pairs (multi-day runs)
    const double *dist;      // [N,N]     (with_time only)
    int64_t N, T_tab;
    // fleet
    int32_t *zone;           // [C] 1-based, 0 = lost
    int64_t C;
    // step
    int64_t step;            // rng step index
    int64_t hour;            // table hour (0-based)
    int32_t with_time;
    RngDev rng;
    // outputs
    uint32_t *parking;       // [N] counts of this step (nullable)
    uint32_t *driving;       // [N]
    unsigned long long *t_hi, *t_lo;   // fixed-point driving-time sum (seconds, 2^-32 s)
    unsigned long long *lost;// scalar
    // trajectory in Julia layout (nullable): state Int64 [C,T_traj], trans Float64 [C,T_traj,4]
    int64_t *traj_state;
    double *traj_trans;
    int64_t T_traj;
    int64_t t_traj;          // column written by this step (0-based)
    int32_t write_next_state;// state[:, t_traj+1] = destination
};
int sampler_grid(int64_t C);
int launch_step_gather(const StepArgs &a, cudaStream_t st);
int launch_reduce_partials(const double *partials, int n, double *out, cudaStream_t st);

// ---- daycars.cu -------------------------------------------------------------------------------
// the car-resident day kernel: one launch = steps [step0, step0 + n_steps) of a day for all cars
struct CarsGeom {
    int n_ctas, threads;
    int cars_per_cta;          // car slots a CTA keeps in shared memory (a multiple of threads)
    size_t smem;
};
struct DayArgs {
    const double *p_drive; const double *cdf; int64_t ldc; int32_t goff, G, gshift; const uint8_t *rowzero; const double *dm;
    const double2 *msd;        // nullable: (mean, std) pairs
    int64_t N, T_tab, C;
    int64_t step0, n_steps, steps_per_hour;
    int32_t with_time, cars_per_cta;
    int32_t counts;            // 0: no per-zone counts (cnt0 / net / drv unused); the lost-car visits go to lost_visits
    int32_t *zone;             // [C] in/out: 1-based zone, 0 = lost
    uint32_t *cnt0;            // nullable: [N+1] += occupancy of every bin before step0 (bin N = lost cars)
    uint32_t *net;             // [day steps][N+1] += net arrivals per bin and step (two's complement)
    uint32_t *drv;             // [day steps][N+1] += driving cars per origin and step
    unsigned long long *lost_visits;   // counts == 0 only: += cars in zone 0, summed over the steps
    // trajectory in Julia layout (nullable; column s of the day = column s of the arrays): state Int64 [C,T_traj],
    // trans Float64 [C,T_traj,4]; dist [N,N] feeds the distance draw that only the trajectory keeps
    int64_t *traj_state; double *traj_trans; int64_t T_traj; const double *dist;
    RngDev rng;
    unsigned long long *t_hi, *t_lo;
};
bool cars_geometry(int64_t N, int64_t C, int with_time, int counts, CarsGeom *g);
int launch_day_cars(const DayArgs &a, const CarsGeom &g, cudaStream_t st);
int launch_cars_export_counts(const uint32_t *cnt0, const uint32_t *net, const uint32_t *drv, int N, int64_t n_steps,
                              uint32_t *parking, uint32_t *driving, unsigned long long *lost, cudaStream_t st);

// ---- bucket.cu --------------------------------------------------------------------------------
constexpr int kTotStride = 64;   // uint32 words between two per-destination arrival counters: 256 B, one L2 slice each
struct BucketGeom {
    int n_slabs;               // CTAs of k_bucket_sample / k_bucket_scatter (whole waves of one fat CTA per SM)
    int64_t slab;              // slots per CTA (<= 65,472: ranks are 16-bit)
    int threads;               // CTA width of k_bucket_sample
    int threads_scatter;       // CTA width of k_bucket_scatter
    int ring;                  // CDF row buffers per CTA
    size_t smem_sample, smem_scatter;
};
struct BucketArgs {
    const double *p_drive; const double *cdf; int64_t ldc; int32_t goff, G, gshift; const uint8_t *rowzero; const double *dm;
    const double2 *msd;        // nullable: (mean, std) pairs
    int64_t N, T_tab;
    const int32_t *perm_cur;   // [C] car ids grouped by zone bin
    const uint32_t *start;     // [N+2] first slot of every bin (bin N = lost cars), start[N+1] = C
    const uint32_t *first_zone;// [n_slabs] the bin that holds the first slot of every slab
    uint32_t *dr;              // [C] per slot: destination bin (low 16 bits) | rank among the slab's cars with that destination
    uint2 *M;                  // [n_slabs][N+1] {cars of slab g bound for bin b, where they start inside bucket b}
    uint32_t *tot;             // [(N+1) * kTotStride] arrivals per destination bin, zero on entry
    int64_t C, slab, step, hour;
    int32_t with_time, ring, ring_shift, gather_policy;   // gather_policy 1: L2 evict-first on the travel-time gathers
    RngDev rng;
    uint32_t *drv;             // [N+1] driving cars per origin of this step
    unsigned long long *t_hi, *t_lo;
};
struct ScatterArgs {
    int NB;
    const uint32_t *tot;       // [(N+1) * kTotStride] arrivals of this step
    uint32_t *tot_zero;        // the counters of the next step but one (zeroed here)
    uint32_t *cnt_next;        // [N+1] new bucket sizes (= the parked-car histogram of the next step)
    const uint32_t *dr; const int32_t *perm_cur; int32_t *perm_next; int64_t C, slab;
    const uint2 *M;
    uint32_t *start_next;      // [N+2] written by block 0
    uint32_t *first_zone_next; // [n_slabs]
    int32_t *zone_out;         // nullable: car-id indexed zone (written on the last step of a day)
};
bool bucket_supported(int64_t N, int64_t ldc);
bool bucket_geometry(int64_t N, int64_t ldc, int64_t C, int with_time, BucketGeom *g);
int launch_bucket_sample(const BucketArgs &a, const BucketGeom &g, cudaStream_t st);
int launch_bucket_scatter(const ScatterArgs &a, const BucketGeom &g, cudaStream_t st);
int launch_bucketize(const int32_t *zone, int64_t C, int NB, uint32_t *cnt0, uint32_t *start, uint32_t *cursor_tmp,
                     int32_t *perm, uint32_t *first_zone, const BucketGeom &g, cudaStream_t st);
int launch_bucket_export_counts(const uint32_t *cnt, const uint32_t *drv, int N, int64_t n_steps, uint32_t *parking,
                                uint32_t *driving, unsigned long long *lost, cudaStream_t st);
int launch_finalize_time(unsigned long long *hi, unsigned long long *lo, double *out, cudaStream_t st);

// helpers
int launch_fleet_init(int32_t *zone, int64_t C, int64_t cars_per_zone, int64_t car_offset, cudaStream_t st);
int launch_i64_to_i32(const int64_t *src, int32_t *dst, int64_t n, cudaStream_t st);
int launch_i32_to_i64(const int32_t *src, int64_t *dst, int64_t n, cudaStream_t st);
int launch_init_states(int64_t *state, double *trans, int64_t C, int64_t T, int64_t cars_per_zone, cudaStream_t st);
// histogram of a Julia-layout trajectory (src/saveresults.jl:8-17)
int launch_traj_histogram(const int64_t *state, const double *trans, int64_t C, int64_t N, int64_t T,
                          uint32_t *parking, uint32_t *driving, cudaStream_t st);
// sum of trans[:,:,3] (src/averagedrivingtime.jl:6-9): deterministic two-stage reduction
int launch_sum_f64(const double *x, int64_t n, double *partials, int n_partials, double *out, cudaStream_t st);
// parking = counts / C, traffic = minmax(colsum(driving)) (src/saveresults.jl:20-28)
int launch_normalise_counts(const uint32_t *parking, const uint32_t *driving, int64_t N, int64_t T, int64_t C,
                            double *parking_out, double *driving_out, double *traffic_out,
                            unsigned long long *traffic_counts_out, cudaStream_t st);

// ---- ingest.cu --------------------------------------------------------------------------------
int launch_rows_scatter(const int32_t *src, const int32_t *dst, const int32_t *hod, const double *mean, const double *sd,
                        int64_t n_rows, int64_t N, int64_t T, double *dm, unsigned long long *n_bad, cudaStream_t st);

int launch_rows_scatter_means(const int32_t *src, const int32_t *dst, const int32_t *hod, const double *mean, int64_t n_rows,
                              int64_t N, int64_t T, double *dm, unsigned long long *n_bad, cudaStream_t st);
int launch_rows_scatter_stds(const int32_t *src, const int32_t *dst, const int32_t *hod, const double *sd, int64_t n_rows,
                             int64_t N, int64_t T, double *dm, cudaStream_t st);

int launch_distance_matrix(const double *lat, const double *lon, int64_t N, double *out, cudaStream_t st);
int launch_polygon_centroids(double *lon, double *lat, int64_t N, int64_t V, double *c_lat, double *c_lon, double *area, cudaStream_t st);

// ---- synth.cu ---------------------------------------------------------------------------------
int launch_pair_table(const double *dm, int64_t N, int64_t T, double *msd, cudaStream_t st);
int launch_synth_city(double *dm, double *dist, int64_t N, int64_t T, uint64_t seed, double density,
                      cudaStream_t st);

}  // namespace cpm
