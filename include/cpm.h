/*
 * cpm.h -- C ABI of libcpm.so: the B200-native (sm_100a) implementation of the
 * CarParkingMaps HMM sampling hot path.
 *
 * The reference (ArsamAryandoust/CarParkingMaps, Julia) has no FFI or plugin interface; its
 * boundary is seven global Julia functions `include`d by main.jl:11-22 plus the globals they read
 * (main.jl:37-42).  Each entry point below names the reference function it replaces (file:line,
 * relative to the reference root).  A Julia host binds them with `ccall` (see INTEGRATION.md and
 * carparkingmaps_b200/julia/CarParkingMapsGPU.jl); the Python mirror binds them with ctypes.
 *
 * Conventions
 *   - Plain C: pointers, sizes, scalars.  No C++/torch types.  No exceptions cross the ABI.
 *   - Every function returns an int32 status (CPM_OK = 0); cpm_last_error() gives the message
 *     of the last failure on the calling thread.
 *   - Arrays use Julia's memory layout: dense, column-major (first index fastest), zone IDs are
 *     1-based VALUES, state is Int64, everything else Float64.
 *   - "Host API" functions take HOST pointers, are synchronous, and keep no host pointer after
 *     they return.  "Engine API" functions work on device-resident handles (city / fleet) so a
 *     pipeline can run without host round trips.
 *   - There is no CPU fallback: every entry point that computes needs a CUDA device and fails
 *     with CPM_ERR_CUDA otherwise.
 */
#ifndef CPM_H
#define CPM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPM_VERSION 100

enum {
    CPM_OK = 0,
    CPM_ERR_ARG = 1,     /* bad argument (null pointer, negative size, ...)              */
    CPM_ERR_CUDA = 2,    /* CUDA runtime error, or no device                             */
    CPM_ERR_OOM = 3,     /* device allocation failed                                     */
    CPM_ERR_STATE = 4    /* call order violated (e.g. sampling before tables were built) */
};

/* Variate source.  The reference draws from Julia's unseeded global RNG
 * (src/resampling.jl:13,29,68,74); here the source is explicit. */
enum {
    CPM_RNG_PHILOX = 0,  /* Philox4x32-10, counter = (car_lo, car_hi, step, 2*stream+phase),
                            key = seed; phase 0 -> (u_drive,u_dest), phase 1 -> (u_time,u_dist);
                            52-bit uniforms on [0,1) like Julia's rand()                     */
    CPM_RNG_INJECTED = 1 /* u_*[i + C*t] supplied by the caller (parity testing)             */
};

typedef struct cpm_rng {
    int32_t  mode;        /* CPM_RNG_*                                                   */
    uint32_t stream;      /* philox stream id (distinct per IVP / day / trial)            */
    uint64_t seed;        /* philox key                                                   */
    int64_t  car_offset;  /* global 0-based index of local car 0 (car sharding over GPUs) */
    const double *u_drive;/* injected: [C,n_steps] column-major                           */
    const double *u_dest; /* injected: read only for driving cars                         */
    const double *u_time; /* injected: read only for drivers with origin != destination   */
    const double *u_dist; /* injected: idem                                               */
} cpm_rng;

typedef struct cpm_ctx cpm_ctx;     /* one per GPU; single caller at a time */
typedef struct cpm_city cpm_city;   /* device-resident inputs + tables of one city/dataset */
typedef struct cpm_fleet cpm_fleet; /* device-resident car state + per-zone counts */

/* ---------------------------------------------------------------------------------------- */
/* Plumbing                                                                                 */
/* ---------------------------------------------------------------------------------------- */
int32_t     cpm_version(void);
const char *cpm_last_error(void);
int32_t     cpm_device_count(int32_t *count);
/* `stream` may be NULL (the context creates its own non-blocking stream) or an existing
 * cudaStream_t, e.g. torch's current stream, so the caller's events bracket the kernels. */
int32_t cpm_ctx_create(int32_t device, void *stream, cpm_ctx **out);
int32_t cpm_ctx_destroy(cpm_ctx *ctx);
int32_t cpm_ctx_sync(cpm_ctx *ctx);
int32_t cpm_ctx_stream(cpm_ctx *ctx, void **stream_out);
/* kernels launched by this context so far (bench.py's gpu_launches) */
int32_t cpm_ctx_launch_count(cpm_ctx *ctx, int64_t *count);
/* Times the sampling (step) kernel with CUDA events on the context's stream: between begin and end,
 * every day loop is bracketed by one event pair around its consecutive step-kernel launches.
 * end() synchronises and returns the summed device time and the number of launches it covers. */
int32_t cpm_ctx_profile_begin(cpm_ctx *ctx);
int32_t cpm_ctx_profile_end(cpm_ctx *ctx, double *step_kernel_ms, int64_t *step_launches);
/* Host-array cache of the host API (default on): the device copies of datamatrix, distance_matrix_km, p_dest and the
 * trajectory arrays are keyed by (host pointer, size, sampled 64-bit checksum); when a later call of the main.jl:82-98
 * sequence passes the same unchanged array, its upload (and, for p_dest, the CDF build) is skipped.  Switch it off for
 * callers that edit those arrays in place between calls.  cpm_ctx_uploads_skipped counts the hits. */
int32_t cpm_ctx_set_input_cache(cpm_ctx *ctx, int32_t enabled);
int32_t cpm_ctx_uploads_skipped(cpm_ctx *ctx, int64_t *count);
/* which sampler the last cpm_solve_ivp / cpm_resampling of this context ran on (CPM_SAMPLER_CARS unless the day kernel's
 * tables do not fit shared memory; 0 before the first call) */
int32_t cpm_ctx_host_last_sampler(cpm_ctx *ctx, int32_t *sampler);
/* Small cities (fleets of at most CPM_GRAPH_MAX_CARS cars, default 300,000; 0 = never): from the third pass of the same
 * shape on, cpm_sample_dataset, cpm_sample_dataset_host and every set of cpm_sweep replay their launch sequence as one
 * executable CUDA graph whose kernel arguments are patched in place (same kernels, same arguments, same results; saves
 * the per-launch host cost: ~3 % of a 100-zone pass).  cpm_ctx_graph_replays counts the passes that ran that way;
 * cpm_ctx_set_graphs(ctx, 0) keeps every pass on direct launches (default: enabled). */
int32_t cpm_ctx_set_graphs(cpm_ctx *ctx, int32_t enabled);
int32_t cpm_ctx_graph_replays(cpm_ctx *ctx, int64_t *count);
/* pinned host memory for callers that want full-rate PCIe copies */
int32_t cpm_host_alloc(int64_t bytes, void **out);
int32_t cpm_host_free(void *p);
/* page-lock / unlock a host array the caller already owns (a Julia Array, a numpy array): uploads from registered
 * memory run at full PCIe rate and overlap with the kernels, like memory from cpm_host_alloc */
int32_t cpm_host_register(void *p, int64_t bytes);
int32_t cpm_host_unregister(void *p);

/* ---------------------------------------------------------------------------------------- */
/* Host API: one entry per reference function (HOST pointers, synchronous)                  */
/* ---------------------------------------------------------------------------------------- */

/* createpdrive(datamatrix, distance_matrix_km, number_zones) -- src/createpdrive.jl:3-38.
 * The globals T, p_min, p_max, e_drive it reads (src/createpdrive.jl:4,31) are passed by value.
 * datamatrix: [N,N,T,2], distance_matrix_km: [N,N], p_drive_out: [N,T]. */
int32_t cpm_create_pdrive(cpm_ctx *ctx, const double *datamatrix, const double *distance_matrix_km,
                          int64_t N, int64_t T, double p_min, double p_max, double e_drive,
                          double *p_drive_out);

/* createpdestin(datamatrix, number_zones) -- src/createpdestin.jl:3-50 (globals T, e_dest).
 * p_dest_out: [N,N,T].  An integral e_dest in {-1,0,1,2,3} takes Julia's Float64^Int fast
 * path (x*x for the reference's e_dest = 2, main.jl:38). */
int32_t cpm_create_pdestin(cpm_ctx *ctx, const double *datamatrix, int64_t N, int64_t T,
                           double e_dest, double *p_dest_out);

/* initializestates(C) -- src/initializestates.jl:4-22 (globals T, cars_per_zone).
 * state: Int64 [C,T], trans: Float64 [C,T,4]; both fully overwritten. */
int32_t cpm_initialize_states(cpm_ctx *ctx, int64_t C, int64_t T, int64_t cars_per_zone,
                              int64_t *state, double *trans);

/* solveinitialvalueproblem(state_matrix, transition_matrix, p_drive, p_dest, C, number_zones)
 * -- src/solveinitialvalueproblem.jl:4-62.  Mutates state[:,2:T], trans[:,1:T-1,1:2] in place
 * exactly as the reference does and writes state[:,T] to initial_state_out [C].
 * n_lost_out (nullable): number of (car,step) visits with zone 0 (see DESIGN.md "sentinel"). */
int32_t cpm_solve_ivp(cpm_ctx *ctx, int64_t *state, double *trans, const double *p_drive,
                      const double *p_dest, int64_t C, int64_t N, int64_t T, const cpm_rng *rng,
                      int64_t *initial_state_out, int64_t *n_lost_out);

/* resampling(state_matrix, transition_matrix, C, number_zones, p_drive, p_dest, datamatrix,
 * distance_matrix_km) -- src/resampling.jl:3-89.  Mutates state[:,2:T] and trans in place; slots
 * 3-4 of non-driving cars keep their previous contents, as in the reference. */
int32_t cpm_resampling(cpm_ctx *ctx, int64_t *state, double *trans, int64_t C, int64_t N, int64_t T,
                       const double *p_drive, const double *p_dest, const double *datamatrix,
                       const double *distance_matrix_km, const cpm_rng *rng, int64_t *n_lost_out);

/* averagedrivingtime(C, A_drive, transition_matrix) -- src/averagedrivingtime.jl:3-12. */
int32_t cpm_average_driving_time(cpm_ctx *ctx, int64_t C, int64_t T, double A_drive_in,
                                 const double *trans, double *A_drive_out);

/* correctparameters(p_min_next, p_max_next, p_min, p_max) -- src/correctparameters.jl:3-23.
 * out2 = {p_min_next, p_max_next}.  Pure host arithmetic (four scalars). */
int32_t cpm_correct_parameters(double p_min_next, double p_max_next, double p_min, double p_max,
                               double *out2);

/* The per-zone histogram and normalisation inside saveresults -- src/saveresults.jl:6-28.
 * parking: [N,T] = counts / C; driving: [N,T] raw counts; traffic: [T] min-max normalised
 * column sums of driving.  Any output pointer may be NULL. */
int32_t cpm_zone_histogram(cpm_ctx *ctx, const int64_t *state, const double *trans, int64_t C,
                           int64_t N, int64_t T, double *parking, double *driving, double *traffic);

/* The distance matrix block of processgeodata -- src/processgeodata.jl:150-166: equirectangular distance between zone
 * centroids (degrees), symmetric, diagonal 1 km.  distance_matrix_km: [N,N]. */
int32_t cpm_distance_matrix(cpm_ctx *ctx, const double *centroid_lat, const double *centroid_long, int64_t N,
                            double *distance_matrix_km);

/* The polygon block of processgeodata -- src/processgeodata.jl:99-148: signed area and centroid of every zone from
 * its zero-terminated vertex list.  longitude_matrix / latitude_matrix: [N,V] column-major as the reference fills them
 * (:17-18, row = zone, 0 = end of list); at most the first 10,000 vertices of a zone are used, as in the reference.
 * The inputs are not modified (the reference closes each ring in place; here that happens on the device copy).
 * centroid_lat / centroid_long / area_polygon: [N]; zones without vertices give 0.  area_polygon may be NULL. */
int32_t cpm_polygon_centroids(cpm_ctx *ctx, const double *longitude_matrix, const double *latitude_matrix, int64_t N,
                              int64_t V, double *centroid_lat, double *centroid_long, double *area_polygon);

/* ---------------------------------------------------------------------------------------- */
/* Engine API: device-resident city tables and fleets                                       */
/* ---------------------------------------------------------------------------------------- */
int32_t cpm_city_create(cpm_ctx *ctx, int64_t N, int64_t T, cpm_city **out);
int32_t cpm_city_destroy(cpm_city *city);
/* inputs: either copied from host, or generated on the device (SURVEY 8d synthetic city) */
int32_t cpm_city_upload(cpm_city *city, const double *datamatrix, const double *distance_matrix_km);
/* createdatamatrix on the device -- src/createdatamatrix.jl:8-23: the CSV rows (sourceid, dstid, hod, mean, std) are
 * scattered into the dense datamatrix; ID 0 -> N, hour 0 -> 24, later rows overwrite earlier ones; rows outside
 * the matrix are skipped and counted in n_skipped (nullable; the reference raises BoundsError). */
int32_t cpm_city_upload_rows(cpm_city *city, const int32_t *sourceid, const int32_t *dstid, const int32_t *hod,
                             const double *mean, const double *stddev, int64_t n_rows,
                             const double *distance_matrix_km, int64_t *n_skipped);
int32_t cpm_city_synth(cpm_city *city, uint64_t seed, double density);
int32_t cpm_city_download_inputs(cpm_city *city, double *datamatrix, double *distance_matrix_km);
/* frees the tables and scratch of a city (inputs stay resident); their device blocks are recycled by the next build */
int32_t cpm_city_release_tables(cpm_city *city);
/* optional: build the (mean, std) pair table that halves the travel-time gather traffic; pays off when many days are
 * sampled on the same inputs (cpm_sweep calls it itself) */
int32_t cpm_city_prepare_multi_day(cpm_city *city);
/* tables: built on the device from the inputs (createpdrive / createpdestin) ... */
int32_t cpm_city_build_pdrive(cpm_city *city, double p_min, double p_max, double e_drive);
int32_t cpm_city_build_pdestin(cpm_city *city, double e_dest, int32_t keep_p_dest);
/* ... or supplied by the caller (the Julia API passes p_drive / p_dest as arguments) */
int32_t cpm_city_set_pdrive(cpm_city *city, const double *p_drive);
int32_t cpm_city_set_pdest(cpm_city *city, const double *p_dest);
int32_t cpm_city_get_pdrive(cpm_city *city, double *p_drive_out);
int32_t cpm_city_get_pdest(cpm_city *city, double *p_dest_out);       /* needs keep_p_dest */
/* sampling CDF rows: [T][N origins][N destinations], destination fastest */
int32_t cpm_city_get_cdf(cpm_city *city, double *cdf_out);
/* the guide ("cut-point") tables of one table hour: guide_out [N origins][G+1] uint16, guide[b] = #{j : cdf[j] < b/G};
 * the categorical draw of src/resampling.jl:38-45 starts its search at guide[floor(u*G)].  G_out = 0: this city has no
 * guide (more than 65,534 zones).  guide_out may be NULL to query G only. */
int32_t cpm_city_get_guide(cpm_city *city, int64_t hour, uint16_t *guide_out, int32_t *G_out);

int32_t cpm_fleet_create(cpm_ctx *ctx, int64_t C, cpm_fleet **out);
int32_t cpm_fleet_destroy(cpm_fleet *fleet);
/* car i (global index car_offset + i) starts in zone floor(i_global / cars_per_zone) + 1 */
int32_t cpm_fleet_init(cpm_fleet *fleet, int64_t cars_per_zone, int64_t car_offset);
int32_t cpm_fleet_set_zones(cpm_fleet *fleet, const int64_t *zones);
int32_t cpm_fleet_get_zones(cpm_fleet *fleet, int64_t *zones_out);
int32_t cpm_fleet_copy_zones(cpm_fleet *dst, const cpm_fleet *src);

/* AUTO: the car-resident day kernel (CARS: one launch per day, cars stay on their SM; csrc/daycars.cu) where its per-zone
 * tables fit shared memory, else the bucket pair (BUCKET: cars re-grouped by zone every step; csrc/bucket.cu; no
 * trajectories), else GATHER (one thread per car and step).  Identical results: all draws are keyed by global car index
 * and step. */
enum { CPM_SAMPLER_AUTO = 0, CPM_SAMPLER_GATHER = 1, CPM_SAMPLER_BUCKET = 2, CPM_SAMPLER_CARS = 3 };

typedef struct cpm_day_opts {
    int64_t n_steps;        /* steps to run (T-1 for the IVP, T for a resampling day)          */
    int64_t steps_per_hour; /* table hour of step s is s / steps_per_hour (1 in the reference) */
    int32_t with_time;      /* 0: phases 1-2 only (IVP); 1: also sample travel time/distance    */
    int32_t sampler;        /* CPM_SAMPLER_*                                                    */
    cpm_rng rng;            /* injected arrays are DEVICE pointers in the engine API            */
} cpm_day_opts;

/* Runs n_steps of the HMM on the fleet (counts-only: per-zone parked/driving counts per step,
 * driving-time sum, lost-car visits are accumulated on the device). */
int32_t cpm_run_day(cpm_city *city, cpm_fleet *fleet, const cpm_day_opts *opts);
/* counts of the last cpm_run_day: parking/driving Int64 [N,n_steps]; any pointer may be NULL.
 * These are raw per-rank counts: a multi-GPU host all-reduces them (SUM) before normalising. */
int32_t cpm_fleet_get_counts(cpm_fleet *fleet, int64_t *parking, int64_t *driving,
                             double *time_sum, int64_t *n_lost);
/* the driving-time sum of the last cpm_run_day in fixed point: whole seconds and 2^-32 s units (exact integers, so a
 * multi-GPU host adds them with an integer all-reduce and gets the unsplit run's value bit for bit) */
int32_t cpm_fleet_get_time_fixed(cpm_fleet *fleet, int64_t *seconds, int64_t *frac32);
/* which sampler (CPM_SAMPLER_GATHER / BUCKET / CARS) the fleet's last day ran on: what CPM_SAMPLER_AUTO resolved to */
int32_t cpm_fleet_last_sampler(cpm_fleet *fleet, int32_t *sampler);
/* device pointer to the uint32 [2][N,n_steps] count block (parking then driving), for an
 * in-place NCCL all-reduce by the host when one city's cars are split over GPUs */
int32_t cpm_fleet_counts_device(cpm_fleet *fleet, void **dev_ptr, int64_t *n_elems);

/* main.jl:82-98 for one dataset, entirely on the device: build both tables, initialise C =
 * N*cars_per_zone cars, solve the IVP, resample one day, reduce.  Outputs (host, nullable):
 * parking [N,T] densities, traffic [T], A_drive increment (averagedrivingtime with A_drive_in=0). */
typedef struct cpm_params {
    double e_drive, e_dest, p_min, p_max;
} cpm_params;
int32_t cpm_sample_dataset(cpm_city *city, cpm_fleet *fleet, const cpm_params *params,
                           int64_t cars_per_zone, uint64_t seed, int32_t sampler, double *parking,
                           double *traffic, double *A_drive, int64_t *n_lost);

/* The same from HOST inputs (datamatrix [N,N,T,2], distance_matrix_km [N,N]) in one call: the upload is pipelined with
 * the computation -- the tables and the IVP start when the mean slab has landed, while the std slab (read only by the
 * resampling day) is still crossing PCIe.  Use pinned host memory (cpm_host_alloc) for the copies to overlap. */
int32_t cpm_sample_dataset_host(cpm_city *city, cpm_fleet *fleet, const double *datamatrix,
                                const double *distance_matrix_km, const cpm_params *params, int64_t cars_per_zone,
                                uint64_t seed, int32_t sampler, double *parking, double *traffic, double *A_drive,
                                int64_t *n_lost);

/* The same pass straight from the rows of the travel-time CSV (createdatamatrix's input, src/createdatamatrix.jl:8-23:
 * sourceid, dstid, hod, mean_travel_time, standard_deviation_travel_time; zone 0 -> N, hour 0 -> 24, last row wins) in
 * one call, HOST columns: ~9x less PCIe traffic than the dense array for Melbourne-like data, and pipelined -- ids,
 * means and distances go up first, the tables and the IVP run while the std column is still on its way.
 * n_skipped_out (nullable): rows outside the matrix (the reference raises BoundsError). */
int32_t cpm_sample_dataset_rows(cpm_city *city, cpm_fleet *fleet, const int32_t *sourceid, const int32_t *dstid,
                                const int32_t *hod, const double *mean, const double *stddev, int64_t n_rows,
                                const double *distance_matrix_km, const cpm_params *params, int64_t cars_per_zone,
                                uint64_t seed, int32_t sampler, double *parking_out, double *traffic_out,
                                double *A_drive_out, int64_t *n_lost_out, int64_t *n_skipped_out);

/* Model-selection sweep (README.md:956-1316, 1435-1681, 1831-2278 as a batched grid): for each of
 * P parameter sets, rebuild the table(s), restart from `initial_zones` and resample one day.
 * streams (nullable, [P]): Philox stream id of each set (default 2 + p); equal ids give common random
 * numbers, global set indices make a sharded sweep independent of the sharding.
 * Outputs (host): traffic_counts Int64 [T,P] raw driving totals per hour, A_drive [P],
 * parking (nullable) Float64 [N,T,P] densities. */
int32_t cpm_sweep(cpm_city *city, const int64_t *initial_zones, int64_t C, const cpm_params *params,
                  int64_t P, uint64_t seed, const uint32_t *streams, int64_t *traffic_counts,
                  double *A_drive, double *parking);

/* Diagnostics, host arithmetic only (no device needed): the launch geometry of the bucket sampler for N zones and C cars.
 * out8 = {supported, slabs (CTAs), slots per slab, threads per CTA, CDF row buffers, bytes of shared memory of the
 * sampling kernel, of the scatter kernel, doubles per CDF row block}. */
int32_t cpm_debug_bucket_geometry(int64_t N, int64_t C, int32_t with_time, int64_t *out8);
/* The same for the car-resident day kernel (CPM_SAMPLER_CARS): out5 = {supported, CTAs, threads per CTA, car slots a CTA
 * keeps in shared memory, bytes of shared memory}. */
int32_t cpm_debug_cars_geometry(int64_t N, int64_t C, int32_t with_time, int64_t *out5);

/* ---------------------------------------------------------------------------------------- */
/* Several GPUs from one host process                                                       */
/* ---------------------------------------------------------------------------------------- */
/* The city loop / dataset loop of main.jl:46-103, :71-103 (serial in the reference) over n_devices GPUs in ONE call: one
 * host thread and one context per device, datasets assigned by longest-processing-time on N*cars_per_zone*(2T-1) +
 * 3*N*N*T, one dataset pass of main.jl:82-98 each (cpm_sample_dataset_host; dataset k uses seed + k), no collective on
 * the data path.  HOST arrays: datamatrix[k] [N[k],N[k],T,2], distance_matrix_km[k] [N[k],N[k]]; outputs per dataset
 * (any of the output arrays may be NULL): parking_out[k] [N[k],T], traffic_out[k] [T], A_drive_out[k], n_lost_out[k]. */
int32_t cpm_sample_datasets_multi(const int32_t *devices, int32_t n_devices, int64_t n_datasets, const int64_t *N,
                                  int64_t T, const double *const *datamatrix, const double *const *distance_matrix_km,
                                  const cpm_params *params, int64_t cars_per_zone, uint64_t seed,
                                  double *const *parking_out, double *const *traffic_out, double *A_drive_out,
                                  int64_t *n_lost_out);
/* ONE city whose cars are split over the devices (contiguous ranges, global Philox counters: the result does not depend
 * on the number of devices), tables replicated, n_steps steps (table hour = step / steps_per_hour), then one NCCL
 * all-reduce (SUM) of the uint32 count blocks; NCCL is loaded with dlopen("libnccl.so.2") on first use.
 * Outputs (host): parking / driving Int64 [N,n_steps] of the whole fleet, driving-time sum (s), lost-car visits,
 * device time of the all-reduce alone (ms; 0 with one device).  Any output pointer may be NULL. */
int32_t cpm_split_city_multi(const int32_t *devices, int32_t n_devices, int64_t N, int64_t T, const double *datamatrix,
                             const double *distance_matrix_km, const cpm_params *params, int64_t cars_per_zone,
                             int64_t n_steps, int64_t steps_per_hour, int32_t with_time, uint64_t seed, uint32_t stream,
                             int64_t *parking_out, int64_t *driving_out, double *time_sum_out, int64_t *n_lost_out,
                             double *allreduce_ms_out);

#ifdef __cplusplus
}
#endif
#endif /* CPM_H */
