/*
 * cpm_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C, single-threaded, fp64, column-major, loop-for-loop restatement of the
 * hot path of ArsamAryandoust/CarParkingMaps (Julia).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the product
 * (carparkingmaps_b200/libcpm.so) never links, loads or calls it.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   - The reference's Julia cannot run in this environment (no julia binary) and the
 *     reference ships no tests or fixtures.  The oracle is pinned against the reproducible
 *     known answers printed in the reference notebook (tests/test_oracle_kat.py) and against
 *     a second, independent pure-Python restatement (oracle/pyref.py) on small cases.
 *   - PARITY UNPINNED for: Julia's global MersenneTwister stream (never seeded by the
 *     reference) and Distributions.jl's Truncated(Normal) sampler (un-vendored, version
 *     unpinned: main.jl:3).  The oracle consumes explicit uniform variates instead and
 *     samples the truncated normal by inverse CDF (same distribution, different stream).
 *
 * Conventions: every array is column-major (first index fastest) exactly as Julia lays it
 * out; zone IDs are 1-based VALUES; indices i,j,t in the comments are Julia's 1-based ones.
 *
 * Documented extension beyond the reference: a destination draw that finds no bin leaves
 * destination = 0 (src/resampling.jl:32,40-45).  The reference then crashes with a
 * BoundsError on the next read of p_drive[0,t].  The oracle (and the CUDA path) instead
 * treat zone 0 as "lost": the car never drives again, is not counted in any zone, and the
 * number of such cars is returned.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_SQRT2
#define M_SQRT2 1.41421356237309504880
#endif

typedef int64_t i64;

/* ------------------------------------------------------------------------------------ */
/* Variate sources                                                                      */
/* ------------------------------------------------------------------------------------ */

enum { ORC_RNG_STREAM = 0, ORC_RNG_INJECTED = 1, ORC_RNG_PHILOX = 2 };

typedef struct {
    int32_t  mode;        /* ORC_RNG_*                                                       */
    uint32_t stream;      /* philox: stream id (IVP / resampling / trial)                     */
    uint64_t seed;        /* stream + philox                                                  */
    i64      car_offset;  /* philox: global index of local car 0 (0-based)                    */
    double  *u_drive;     /* [C,T] injected: read; stream/philox: written when non-NULL       */
    double  *u_dest;      /* [C,T] only entries of driving cars are read / written            */
    double  *u_time;      /* [C,T] only drivers with origin != destination                    */
    double  *u_dist;      /* [C,T] idem                                                       */
    uint64_t s;           /* stream mode: splitmix64 state (set from seed by orc_rng_init)    */
} orc_rng;

void orc_rng_init(orc_rng *r) { r->s = r->seed; }

/* stream mode: 52-bit uniforms on [0,1), the resolution of Julia <= 1.6 rand() (dSFMT). */
static double stream_next(orc_rng *r)
{
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (double)(z >> 12) * 0x1.0p-52;
}

/* Philox4x32-10 (Salmon et al. 2011), the same counter-based generator the CUDA path uses.
 * counter = (car_lo, car_hi, step, 2*stream+phase), key = (seed_lo, seed_hi). */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox_pair(uint64_t seed, uint32_t stream, i64 car, uint32_t step, uint32_t phase,
                     double *ua, double *ub)
{
    uint32_t ctr[4] = { (uint32_t)((uint64_t)car & 0xffffffffu), (uint32_t)((uint64_t)car >> 32),
                        step, stream * 2u + phase };
    uint32_t key[2] = { (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32) };
    uint32_t o[4];
    philox4x32_10(ctr, key, o);
    uint64_t a = ((uint64_t)o[1] << 32) | o[0];
    uint64_t b = ((uint64_t)o[3] << 32) | o[2];
    *ua = (double)(a >> 12) * 0x1.0p-52;
    *ub = (double)(b >> 12) * 0x1.0p-52;
}

/* which: 0 drive, 1 dest, 2 time, 3 dist.  i,t are 0-based here. */
static double draw(orc_rng *r, int which, i64 i, i64 t, i64 C)
{
    double *arr = which == 0 ? r->u_drive : which == 1 ? r->u_dest : which == 2 ? r->u_time : r->u_dist;
    double u;
    if (r->mode == ORC_RNG_INJECTED) {
        return arr[i + C * t];
    } else if (r->mode == ORC_RNG_PHILOX) {
        double a, b;
        orc_philox_pair(r->seed, r->stream, r->car_offset + i, (uint32_t)t, (uint32_t)(which >> 1), &a, &b);
        u = (which & 1) ? b : a;
    } else {
        u = stream_next(r);
    }
    if (arr) arr[i + C * t] = u;
    return u;
}

/* ------------------------------------------------------------------------------------ */
/* Small Julia-semantics helpers                                                        */
/* ------------------------------------------------------------------------------------ */

/* Julia Base `^`: Float64^Int64 special-cases -1,0,1,2,3 (base/math.jl), otherwise libm pow.
 * The C ABI carries the exponent as a double; integral values in that set take the
 * multiplication form, which is what the reference's `e_dest = 2` (an Int, main.jl:38) hits. */
static double jl_pow(double x, double e)
{
    if (e == 2.0) return x * x;
    if (e == 1.0) return x;
    if (e == 3.0) return x * x * x;
    if (e == 0.0) return 1.0;
    if (e == -1.0) return 1.0 / x;
    return pow(x, e);
}

/* Julia's sum(::Vector{Float64}) = Base.mapreduce_impl(identity, +, A, 1, n, 1024) (base/reduce.jl): ranges of at  */
/* most 1024 elements are added left to right, longer ones split at imid = (ifirst + ilast) >> 1, i.e. the FIRST   */
/* half gets ceil(n/2) elements.  (The sequential part is a @simd loop in Julia, whose association order is up to   */
/* LLVM: "parity unpinned at the ulp level")                                                           */
static double jl_sum(const double *x, i64 n)
{
    if (n <= 1024) {
        double s = 0.0;
        for (i64 k = 0; k < n; ++k) s += x[k];
        return s;
    }
    i64 h = (n + 1) / 2;
    return jl_sum(x, h) + jl_sum(x + h, n - h);
}
double orc_jl_sum(const double *x, i64 n) { return jl_sum(x, n); }   /* exported for tests */

/* Julia maximum/minimum propagate NaN. */
static double jl_max(const double *x, i64 n)
{
    double m = x[0];
    for (i64 k = 0; k < n; ++k) { if (isnan(x[k])) return NAN; if (x[k] > m) m = x[k]; }
    return m;
}
static double jl_min(const double *x, i64 n)
{
    double m = x[0];
    for (i64 k = 0; k < n; ++k) { if (isnan(x[k])) return NAN; if (x[k] < m) m = x[k]; }
    return m;
}

#define DM(i, j, t, k) dm[(i) + N * ((j) + N * ((t) + T * (k)))]

/* ------------------------------------------------------------------------------------ */
/* distance matrix from zone centroids -- src/processgeodata.jl:150-166                 */
/* ------------------------------------------------------------------------------------ */
void orc_distancematrix(const double *lat, const double *lon, i64 N, double *out)
{
    const double c_lat_long = 111.3, conv_deg_rad = 0.01745;                 /* :151-152 */
    memset(out, 0, (size_t)(N * N) * sizeof(double));
    for (i64 i = 0; i < N - 1; ++i)                                          /* :154 */
        for (i64 j = i; j < N; ++j) {                                        /* :157 */
            double c = cos((lat[i] + lat[j]) / 2 * conv_deg_rad);
            double d = c_lat_long * sqrt(c * c * ((lon[i] - lon[j]) * (lon[i] - lon[j])) + (lat[i] - lat[j]) * (lat[i] - lat[j]));
            out[i + N * j] = d; out[j + N * i] = d;                          /* :159-160 */
        }
    for (i64 i = 0; i < N; ++i) out[i + N * i] = 1;                          /* :164-166 */
}

/* ------------------------------------------------------------------------------------ */
/* polygon area and centroid from zero-terminated vertex lists -- src/processgeodata.jl:99-148 */
/* lon/lat: [N,V] column-major, row i = the vertices of zone i+1 in file order, 0 = end.       */
/* Like the reference, the area pass closes the ring in place (writes the first vertex over     */
/* the terminator, :113-114) and the centroid pass does NOT stop at a zero (:130-141): it runs  */
/* over all J = min(10000, V-1) columns.  A row without a terminator inside J columns (the       */
/* reference then indexes out of bounds) leaves the ring open: its sums simply end at J.         */
/* ------------------------------------------------------------------------------------ */
void orc_polygoncentroids(double *lon, double *lat, i64 N, i64 V, double *c_lat, double *c_lon, double *area)
{
#define LO(i, j) lon[(i) + N * ((j) - 1)]
#define LA(i, j) lat[(i) + N * ((j) - 1)]
    const i64 J = (V - 1 < 10000) ? V - 1 : 10000;                            /* j+1 must exist */
    for (i64 i = 0; i < N; ++i) { c_lat[i] = 0; c_lon[i] = 0; area[i] = 0; }  /* :99-101 */
    for (i64 i = 0; i < N; ++i) {                                             /* :104 */
        if (LO(i, 1) == 0) continue;                                          /* :105 */
        double s = 0;
        for (i64 j = 1; j <= J; ++j) {                                        /* :107 */
            if (LO(i, j) == 0) {                                              /* :108 */
                s = s - (LA(i, j - 1) * LO(i, j) - LA(i, j) * LO(i, j - 1));  /* :110 */
                LO(i, j) = LO(i, 1); LA(i, j) = LA(i, 1);                     /* :113-114 */
                s = s + (LA(i, j - 1) * LO(i, j) - LA(i, j) * LO(i, j - 1));  /* :117 */
                break;
            }
            s = s + (LA(i, j) * LO(i, j + 1) - LA(i, j + 1) * LO(i, j));      /* :120 */
        }
        area[i] = s / 2;                                                      /* :123 */
    }
    for (i64 i = 0; i < N; ++i) {                                             /* :128 */
        if (LO(i, 1) == 0) continue;
        double sl = 0, sa = 0;
        for (i64 j = 1; j <= J; ++j) {                                        /* :132, no break */
            if (LO(i, j) == 0) {
                const double cr = LO(i, j - 1) * LA(i, j) - LO(i, j) * LA(i, j - 1);
                sl = sl - (LO(i, j - 1) + LO(i, j)) * cr;                     /* :135 */
                sa = sa - (LA(i, j - 1) + LA(i, j)) * cr;                     /* :136 */
            } else {
                const double cr = LO(i, j) * LA(i, j + 1) - LO(i, j + 1) * LA(i, j);
                sl = sl + (LO(i, j) + LO(i, j + 1)) * cr;                     /* :138 */
                sa = sa + (LA(i, j) + LA(i, j + 1)) * cr;                     /* :139 */
            }
        }
        c_lon[i] = -sl / (6 * area[i]);                                       /* :142 */
        c_lat[i] = -sa / (6 * area[i]);                                       /* :143 */
    }
#undef LO
#undef LA
}

/* ------------------------------------------------------------------------------------ */
/* createdatamatrix (the scatter, not the CSV parsing) -- src/createdatamatrix.jl:7-23   */
/* rows applied in order: later rows overwrite earlier ones.  Returns #rows outside the   */
/* matrix (the reference raises BoundsError there; they are skipped here).               */
/* ------------------------------------------------------------------------------------ */
i64 orc_createdatamatrix(const int32_t *src, const int32_t *dst, const int32_t *hod, const double *mean,
                         const double *sd, i64 n_rows, i64 N, i64 T, double *dm)
{
    i64 bad = 0;
    memset(dm, 0, (size_t)(N * N * T * 2) * sizeof(double));               /* :7 */
    for (i64 r = 0; r < n_rows; ++r) {                                     /* :8 */
        i64 i = src[r], j = dst[r], h = hod[r];
        if (i == 0) i = N;                                                 /* :9-11 */
        if (j == 0) j = N;                                                 /* :12-14 */
        if (h == 0) h = 24;                                                /* :15-17 */
        if (i < 1 || i > N || j < 1 || j > N || h < 1 || h > T) { ++bad; continue; }
        DM(i - 1, j - 1, h - 1, 0) = mean[r];                              /* :21 */
        DM(i - 1, j - 1, h - 1, 1) = sd[r];                                /* :22 */
    }
    return bad;
}

/* ------------------------------------------------------------------------------------ */
/* createpdrive -- src/createpdrive.jl:3-38                                             */
/* ------------------------------------------------------------------------------------ */
void orc_createpdrive(const double *dm, const double *dist, i64 N, i64 T,
                      double p_min, double p_max, double e_drive, double *p_drive)
{
    double *mean_sum = (double *)calloc((size_t)(N * T), sizeof(double));
    double *row = (double *)malloc((size_t)T * sizeof(double));
    memset(p_drive, 0, (size_t)(N * T) * sizeof(double));                  /* :4  */
    for (i64 i = 0; i < N; ++i) {
        for (i64 t = 0; t < T; ++t) {
            double s = 0.0; i64 counter = 0;                               /* :12-13 */
            for (i64 j = 0; j < N; ++j) {                                  /* :14 ascending j */
                double m = DM(i, j, t, 0);
                if (m != 0) { s = s + m / dist[i + N * j]; counter += 1; } /* :15-17 */
            }
            mean_sum[i + N * t] = s / (double)counter;                     /* :20  0/0 = NaN */
        }
        for (i64 t = 0; t < T; ++t) row[t] = mean_sum[i + N * t];
        double max_t = jl_max(row, T), min_t = jl_min(row, T);             /* :22-23 */
        if (max_t > 0) {                                                   /* :28 (NaN -> false) */
            for (i64 t = 0; t < T; ++t)
                p_drive[i + N * t] = p_min + (p_max - p_min) *
                    jl_pow((mean_sum[i + N * t] - min_t) / (max_t - min_t), e_drive);  /* :31 */
        }
    }
    free(row); free(mean_sum);
}

/* ------------------------------------------------------------------------------------ */
/* createpdestin -- src/createpdestin.jl:3-50                                           */
/* ------------------------------------------------------------------------------------ */
void orc_createpdestin(const double *dm, i64 N, i64 T, double e_dest, double *p_dest)
{
    double *col = (double *)malloc((size_t)(T > N ? T : N) * sizeof(double));
    memset(p_dest, 0, (size_t)(N * N * T) * sizeof(double));               /* :4 */
    for (i64 i = 0; i < N; ++i)
        for (i64 j = 0; j < N; ++j) {
            for (i64 t = 0; t < T; ++t) col[t] = DM(i, j, t, 0);
            double max_x = jl_max(col, T), min_x = jl_min(col, T);         /* :13-14 (zeros count) */
            if (max_x > 0)                                                 /* :22 */
                for (i64 t = 0; t < T; ++t)
                    p_dest[i + N * (j + N * t)] =
                        jl_pow((col[t] - min_x) / (max_x - min_x), e_dest);/* :24 */
        }
    for (i64 i = 0; i < N; ++i)
        for (i64 t = 0; t < T; ++t) {
            for (i64 j = 0; j < N; ++j) col[j] = p_dest[i + N * (j + N * t)];
            double nf = jl_sum(col, N);                                    /* :33 */
            if (nf > 0)                                                    /* :40 (NaN -> skip) */
                for (i64 j = 0; j < N; ++j) p_dest[i + N * (j + N * t)] /= nf;  /* :42 */
        }
    free(col);
}

/* ------------------------------------------------------------------------------------ */
/* initializestates -- src/initializestates.jl:4-22                                     */
/* ------------------------------------------------------------------------------------ */
void orc_initializestates(i64 C, i64 T, i64 cars_per_zone, i64 *state, double *trans)
{
    memset(state, 0, (size_t)(C * T) * sizeof(i64));                       /* :6  */
    memset(trans, 0, (size_t)(C * T * 4) * sizeof(double));                /* :7  */
    i64 zone = 0;
    for (i64 i = 0; i <= C - cars_per_zone; i += cars_per_zone) {          /* :11 */
        zone += 1;
        for (i64 j = 1; j <= cars_per_zone; ++j) state[(i + j - 1)] = zone;/* :14 column 1 */
    }
}

/* ------------------------------------------------------------------------------------ */
/* phases 1+2 of one step -- src/resampling.jl:11-49 == src/solveinitialvalueproblem.jl:11-49 */
/* ------------------------------------------------------------------------------------ */
#define TR(i, t, k) trans[(i) + C * ((t) + T * (k))]

static i64 phases12(i64 *state, double *trans, const double *p_drive, const double *p_dest,
                    i64 C, i64 N, i64 T, i64 t, orc_rng *rng, double *distribution)
{
    i64 lost = 0;
    for (i64 i = 0; i < C; ++i) {                                          /* :11 */
        i64 origin = state[i + C * t];
        if (origin < 1 || origin > N) {            /* extension: lost car (reference: BoundsError) */
            TR(i, t, 0) = 0; TR(i, t, 1) = (double)origin; ++lost; continue;
        }
        double u = draw(rng, 0, i, t, C);                                  /* :13 */
        double pd = p_drive[(origin - 1) + N * t];                         /* :14 */
        double drive;
        if (u <= pd) drive = 1;                                            /* :15  <=, NaN -> park */
        else { drive = 0; TR(i, t, 1) = (double)origin; }                  /* :18-19 */
        TR(i, t, 0) = drive;                                               /* :21 */
    }
    for (i64 i = 0; i < C; ++i) {                                          /* :26 */
        if (TR(i, t, 0) == 1) {                                            /* :28 */
            double u = draw(rng, 1, i, t, C);                              /* :29 */
            double range_up = 0, range_low = 0; i64 destination = 0;       /* :30-32 */
            i64 origin = state[i + C * t];
            for (i64 j = 0; j < N; ++j)                                    /* :34 strided row COPY */
                distribution[j] = p_dest[(origin - 1) + N * (j + N * t)];
            if (jl_sum(distribution, N) == 0) destination = origin;        /* :35-36 */
            else
                for (i64 j = 0; j < N; ++j) {                              /* :38 */
                    range_up = range_up + distribution[j];                 /* :39 */
                    if (range_low < u && u <= range_up) { destination = j + 1; break; }  /* :40-42 */
                    range_low = range_up;                                  /* :44 */
                }
            TR(i, t, 1) = (double)destination;                             /* :47 */
        }
    }
    return lost;
}

/* solveinitialvalueproblem -- src/solveinitialvalueproblem.jl:4-62.  Returns #lost-car visits. */
i64 orc_solveinitialvalueproblem(i64 *state, double *trans, const double *p_drive, const double *p_dest,
                                 i64 C, i64 N, i64 T, orc_rng *rng, i64 *initial_state)
{
    double *distribution = (double *)malloc((size_t)N * sizeof(double));
    i64 lost = 0;
    for (i64 t = 0; t < T - 1; ++t) {                                      /* :8 */
        lost += phases12(state, trans, p_drive, p_dest, C, N, T, t, rng, distribution);
        for (i64 i = 0; i < C; ++i) state[i + C * (t + 1)] = (i64)TR(i, t, 1);  /* :53 */
    }
    /* :57 hard-codes column 24; column T is the generalisation (T = 24 in the reference). */
    for (i64 i = 0; i < C; ++i) initial_state[i] = state[i + C * (T - 1)];
    free(distribution);
    return lost;
}

/* ------------------------------------------------------------------------------------ */
/* Truncated(Normal(mu, sigma), 0.9mu, 1.1mu) by inverse CDF.                           */
/* The interval is symmetric about mu, so with b = 0.1*mu/sigma and w = 2u-1:           */
/*     z = sqrt(2) * erfinv( w * erf(b/sqrt(2)) ),   x = mu + sigma*z, clamped.          */
/* Distributions.jl draws the same law by rejection (PARITY UNPINNED, see header).      */
/* ------------------------------------------------------------------------------------ */
static double orc_erfinv(double y)
{
    if (!(y > -1.0 && y < 1.0)) return y == 1.0 ? INFINITY : y == -1.0 ? -INFINITY : NAN;
    if (y == 0.0) return 0.0;
    double a = fabs(y);
    /* Winitzki start, then Newton on erf (central) or erfc (tail) to convergence */
    double ln = log1p(-a * a), c = 2.0 / (M_PI * 0.147) + 0.5 * ln;
    double x = sqrt(sqrt(c * c - ln / 0.147) - c);
    const double k = 1.1283791670955125739; /* 2/sqrt(pi) */
    for (int it = 0; it < 6; ++it) {
        double f = (a < 0.8) ? (erf(x) - a) : ((1.0 - a) - erfc(x));
        double d = k * exp(-x * x);
        double dx = f / d;
        /* Halley correction */
        dx = dx / (1.0 + x * dx);
        x -= dx;
        if (fabs(dx) <= 1e-17 * fabs(x)) break;
    }
    return y < 0 ? -x : x;
}
double orc_erfinv_export(double y) { return orc_erfinv(y); }

double orc_truncnormal(double mu, double sigma, double u)
{
    if (!(sigma > 0) || !isfinite(sigma) || !isfinite(mu)) return mu;   /* degenerate: extension */
    double lo = 0.9 * mu, hi = 1.1 * mu;
    double b = (0.1 * mu) / sigma;
    double E = erf(b / M_SQRT2);
    double w = 2.0 * u - 1.0;
    double z = M_SQRT2 * orc_erfinv(w * E);
    double x = mu + sigma * z;
    if (lo > hi) { double tmp = lo; lo = hi; hi = tmp; }
    if (x < lo) x = lo;
    if (x > hi) x = hi;
    return x;
}

/* resampling -- src/resampling.jl:3-89.  Returns #lost-car visits. */
i64 orc_resampling(i64 *state, double *trans, i64 C, i64 N, i64 T, const double *p_drive,
                   const double *p_dest, const double *dm, const double *dist, orc_rng *rng)
{
    double *distribution = (double *)malloc((size_t)N * sizeof(double));
    i64 lost = 0;
    for (i64 t = 0; t < T; ++t) {                                          /* :7 */
        lost += phases12(state, trans, p_drive, p_dest, C, N, T, t, rng, distribution);
        for (i64 i = 0; i < C; ++i) {                                      /* :53 */
            if (TR(i, t, 0) == 1) {                                        /* :55 */
                i64 origin = state[i + C * t];
                i64 destination = (i64)TR(i, t, 1);                        /* :57 */
                if (origin == destination) {                               /* :58 */
                    TR(i, t, 2) = 5 * 60; TR(i, t, 3) = 1;                 /* :59-60 */
                } else if (destination >= 1) {
                    double mean = DM(origin - 1, destination - 1, t, 0);   /* :63 */
                    double sd = DM(origin - 1, destination - 1, t, 1);     /* :64 */
                    if (sd == 0) sd = 0.1 * mean;                          /* :65-67 */
                    TR(i, t, 2) = orc_truncnormal(mean, sd, draw(rng, 2, i, t, C));  /* :68-69 */
                    mean = dist[(origin - 1) + N * (destination - 1)];     /* :72 */
                    sd = 0.1 * mean;                                       /* :73 */
                    TR(i, t, 3) = orc_truncnormal(mean, sd, draw(rng, 3, i, t, C));  /* :74-75 */
                }
                /* destination == 0 (sentinel): reference raises BoundsError at :63; extension: slots untouched */
            }
        }
        if (t < T - 1)                                                     /* :81 */
            for (i64 i = 0; i < C; ++i) state[i + C * (t + 1)] = (i64)TR(i, t, 1);  /* :82 */
    }
    free(distribution);
    return lost;
}

/* averagedrivingtime -- src/averagedrivingtime.jl:3-12 */
double orc_averagedrivingtime(i64 C, i64 T, double A_drive, const double *trans)
{
    double sum_driving_time = 0;
    for (i64 t = 0; t < T; ++t)
        sum_driving_time = sum_driving_time + jl_sum(&TR(0, t, 2), C);     /* :7-8 */
    return A_drive + sum_driving_time / ((double)C * (double)T * 60 * 60); /* :10 */
}

/* correctparameters -- src/correctparameters.jl:3-23 */
void orc_correctparameters(double p_min_next, double p_max_next, double p_min, double p_max, double *out)
{
    if (p_min_next < 0) p_min_next = 0;
    else if (p_min_next > p_max) p_min_next = p_max;
    if (p_max_next > 1) p_max_next = 1;
    else if (p_max_next < p_min) p_max_next = p_min;
    out[0] = p_min_next; out[1] = p_max_next;
}

/* zone histogram + normalisation -- src/saveresults.jl:6-28.
 * parking, driving: [N,T]; traffic: [T].  parking is returned divided by C (:20); driving is
 * returned as raw counts (the reference only uses its column sums).  Lost cars (zone 0) are skipped. */
void orc_zonehistogram(i64 N, i64 T, i64 C, const i64 *state, const double *trans,
                       double *parking, double *driving, double *traffic)
{
    memset(parking, 0, (size_t)(N * T) * sizeof(double));
    memset(driving, 0, (size_t)(N * T) * sizeof(double));
    for (i64 i = 0; i < C; ++i)
        for (i64 t = 0; t < T; ++t) {
            i64 index = state[i + C * t];                                  /* :10 */
            if (index < 1 || index > N) continue;
            parking[(index - 1) + N * t] += 1;                             /* :12 */
            if (TR(i, t, 0) == 1) driving[(index - 1) + N * t] += 1;       /* :13-15 */
        }
    for (i64 k = 0; k < N * T; ++k) parking[k] = parking[k] / (double)C;   /* :20 */
    for (i64 t = 0; t < T; ++t) {                                          /* :23 sum(dims=1) */
        double s = 0; for (i64 z = 0; z < N; ++z) s += driving[z + N * t];
        traffic[t] = s;
    }
    double mn = jl_min(traffic, T), mx = jl_max(traffic, T);               /* :24-25 */
    for (i64 t = 0; t < T; ++t) traffic[t] = (traffic[t] - mn) / (mx - mn);/* :26-28 */
}

/* Counts-only variant of a whole day (what saveresults consumes), used to check the CUDA
 * counts-only path at sizes where the [C,T,4] trajectory would not fit: identical draws,
 * identical algorithm, but only per-zone integer counts and the driving-time sum are kept.
 * n_steps may exceed T_tab (per-minute stepping): table hour = step / steps_per_hour.
 * with_time: 0 = IVP-style (phases 1-2 only), 1 = resampling-style (phase 3 too). */
i64 orc_day_counts(i64 *zone /*[C] in/out*/, i64 C, i64 N, i64 T_tab, i64 n_steps, i64 steps_per_hour,
                   const double *p_drive, const double *p_dest, const double *dm, const double *dist,
                   orc_rng *rng, int with_time, i64 *parking /*[N,n_steps]*/, i64 *driving /*[N,n_steps]*/,
                   double *time_sum /* scalar out */)
{
    i64 T = T_tab; /* for DM macro */
    (void)dist;   /* distances are not part of the counts-only outputs */
    double *distribution = (double *)malloc((size_t)N * sizeof(double));
    i64 lost = 0; double tsum = 0;
    memset(parking, 0, (size_t)(N * n_steps) * sizeof(i64));
    memset(driving, 0, (size_t)(N * n_steps) * sizeof(i64));
    i64 *next = (i64 *)malloc((size_t)C * sizeof(i64));
    for (i64 s = 0; s < n_steps; ++s) {
        i64 t = s / steps_per_hour;
        for (i64 i = 0; i < C; ++i) {
            i64 origin = zone[i];
            next[i] = origin;
            if (origin < 1 || origin > N) { ++lost; continue; }
            parking[(origin - 1) + N * s] += 1;
            double u = draw(rng, 0, i, s, C);
            if (!(u <= p_drive[(origin - 1) + N * t])) continue;
            driving[(origin - 1) + N * s] += 1;
            double u2 = draw(rng, 1, i, s, C);
            double range_up = 0, range_low = 0; i64 destination = 0;
            for (i64 j = 0; j < N; ++j) distribution[j] = p_dest[(origin - 1) + N * (j + N * t)];
            if (jl_sum(distribution, N) == 0) destination = origin;
            else
                for (i64 j = 0; j < N; ++j) {
                    range_up = range_up + distribution[j];
                    if (range_low < u2 && u2 <= range_up) { destination = j + 1; break; }
                    range_low = range_up;
                }
            next[i] = destination;
            if (with_time) {
                if (origin == destination) tsum += 300.0;
                else if (destination >= 1) {
                    double mean = DM(origin - 1, destination - 1, t, 0);
                    double sd = DM(origin - 1, destination - 1, t, 1);
                    if (sd == 0) sd = 0.1 * mean;
                    tsum += orc_truncnormal(mean, sd, draw(rng, 2, i, s, C));
                }
            }
        }
        memcpy(zone, next, (size_t)C * sizeof(i64));
    }
    if (time_sum) *time_sum = tsum;
    free(next); free(distribution);
    return lost;
}
