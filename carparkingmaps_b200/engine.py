"""Device-resident engine: thin object wrappers over the C ABI's context / city / fleet handles.

A pipeline built from these objects runs main.jl:82-98 of the reference without host round trips;
only the reduced results (per-zone densities, hourly traffic, A_drive) cross PCIe.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _f64F(a, shape=None):
    a = np.asarray(a)
    if a.dtype != np.float64 or not a.flags.f_contiguous:
        a = np.asfortranarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


@dataclass
class Rng:
    """Variate source of a sampling call (cpm_rng).  mode 'philox' | 'injected'."""
    mode: str = "philox"
    seed: int = 0
    stream: int = 0
    car_offset: int = 0
    u_drive: object = None   # numpy (host API) or int device pointers (engine API)
    u_dest: object = None
    u_time: object = None
    u_dist: object = None

    def to_c(self) -> L.CRng:
        r = L.CRng()
        r.mode = {"philox": L.RNG_PHILOX, "injected": L.RNG_INJECTED}[self.mode]
        r.seed, r.stream, r.car_offset = self.seed, self.stream, self.car_offset
        for k in ("u_drive", "u_dest", "u_time", "u_dist"):
            v = getattr(self, k)
            if v is None:
                setattr(r, k, None)
            elif isinstance(v, int):
                setattr(r, k, v)
            else:
                if v.dtype != np.float64 or not v.flags.f_contiguous:
                    raise ValueError(f"{k} must be a Fortran-ordered float64 array")
                setattr(r, k, v.ctypes.data)
        return r


class Context:
    """One per GPU (cpm_ctx).  `stream`: optional raw cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._h = C.c_void_p()
        L.check(L.lib().cpm_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(self._h)))
        self.device = device

    def sync(self):
        L.check(L.lib().cpm_ctx_sync(self._h))

    @property
    def stream(self) -> int:
        s = C.c_void_p()
        L.check(L.lib().cpm_ctx_stream(self._h, C.byref(s)))
        return s.value or 0

    @property
    def launches(self) -> int:
        n = C.c_int64()
        L.check(L.lib().cpm_ctx_launch_count(self._h, C.byref(n)))
        return n.value

    def set_input_cache(self, enabled: bool):
        """Host API only: reuse device copies of unchanged host arrays across calls (default on)."""
        L.check(L.lib().cpm_ctx_set_input_cache(self._h, int(bool(enabled))))

    @property
    def uploads_skipped(self) -> int:
        n = C.c_int64()
        L.check(L.lib().cpm_ctx_uploads_skipped(self._h, C.byref(n)))
        return n.value

    @property
    def host_last_sampler(self) -> int:
        """Sampler of the last solveinitialvalueproblem / resampling call through the host API (SAMPLER_CARS normally)."""
        v = C.c_int32()
        L.check(L.lib().cpm_ctx_host_last_sampler(self._h, C.byref(v)))
        return v.value

    def set_graphs(self, enabled: bool):
        """Small cities: replay repeated passes as CUDA graphs (default on)."""
        L.check(L.lib().cpm_ctx_set_graphs(self._h, int(bool(enabled))))

    @property
    def graph_replays(self) -> int:
        """Passes over small cities that ran as a replayed CUDA graph (cpm.h: cpm_ctx_graph_replays)."""
        n = C.c_int64()
        L.check(L.lib().cpm_ctx_graph_replays(self._h, C.byref(n)))
        return n.value

    def profile_begin(self):
        L.check(L.lib().cpm_ctx_profile_begin(self._h))

    def profile_end(self):
        """(summed device ms of the step-kernel launches since profile_begin, number of launches)"""
        ms, n = C.c_double(), C.c_int64()
        L.check(L.lib().cpm_ctx_profile_end(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def close(self):
        if self._h:
            L.lib().cpm_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class City:
    """Device-resident inputs (datamatrix, distance_matrix_km) and tables (p_drive, p_dest CDF) of one dataset."""

    def __init__(self, ctx: Context, N: int, T: int = 24):
        self.ctx, self.N, self.T = ctx, int(N), int(T)
        self._h = C.c_void_p()
        L.check(L.lib().cpm_city_create(ctx._h, self.N, self.T, C.byref(self._h)))

    def upload(self, datamatrix, distance_matrix_km):
        dm = _f64F(datamatrix, (self.N, self.N, self.T, 2))
        dist = _f64F(distance_matrix_km, (self.N, self.N))
        L.check(L.lib().cpm_city_upload(self._h, _ptr(dm), _ptr(dist)))
        return self

    def upload_rows(self, sourceid, dstid, hod, mean, stddev, distance_matrix_km):
        """createdatamatrix on the device (src/createdatamatrix.jl:8-23) from the CSV columns; returns #skipped rows."""
        cols = [np.ascontiguousarray(a, dtype=np.int32) for a in (sourceid, dstid, hod)]
        vals = [np.ascontiguousarray(a, dtype=np.float64) for a in (mean, stddev)]
        n = cols[0].shape[0]
        assert all(a.shape == (n,) for a in cols + vals)
        dist = _f64F(distance_matrix_km, (self.N, self.N))
        skipped = C.c_int64()
        L.check(L.lib().cpm_city_upload_rows(self._h, _ptr(cols[0]), _ptr(cols[1]), _ptr(cols[2]), _ptr(vals[0]), _ptr(vals[1]),
                                             n, _ptr(dist), C.byref(skipped)))
        return skipped.value

    def synth(self, seed: int, density: float = 1.0):
        L.check(L.lib().cpm_city_synth(self._h, seed, float(density)))
        return self

    def download_inputs(self):
        dm = np.empty((self.N, self.N, self.T, 2), order="F")
        dist = np.empty((self.N, self.N), order="F")
        L.check(L.lib().cpm_city_download_inputs(self._h, _ptr(dm), _ptr(dist)))
        return dm, dist

    def release_tables(self):
        """Free the tables (their device blocks are recycled); the inputs stay resident."""
        L.check(L.lib().cpm_city_release_tables(self._h))
        return self

    def prepare_multi_day(self):
        """Build the (mean, std) pair table: halves the travel-time gather traffic of every later resampling day."""
        L.check(L.lib().cpm_city_prepare_multi_day(self._h))
        return self

    def build_pdrive(self, p_min, p_max, e_drive):
        L.check(L.lib().cpm_city_build_pdrive(self._h, p_min, p_max, float(e_drive)))
        return self

    def build_pdestin(self, e_dest, keep_p_dest: bool = True):
        L.check(L.lib().cpm_city_build_pdestin(self._h, float(e_dest), int(keep_p_dest)))
        return self

    def set_pdrive(self, p_drive):
        L.check(L.lib().cpm_city_set_pdrive(self._h, _ptr(_f64F(p_drive, (self.N, self.T)))))
        return self

    def set_pdest(self, p_dest):
        L.check(L.lib().cpm_city_set_pdest(self._h, _ptr(_f64F(p_dest, (self.N, self.N, self.T)))))
        return self

    def pdrive(self):
        out = np.empty((self.N, self.T), order="F")
        L.check(L.lib().cpm_city_get_pdrive(self._h, _ptr(out)))
        return out

    def pdest(self):
        out = np.empty((self.N, self.N, self.T), order="F")
        L.check(L.lib().cpm_city_get_pdest(self._h, _ptr(out)))
        return out

    def cdf(self):
        """[T, N origins, N destinations] (C order: destination fastest)."""
        out = np.empty((self.T, self.N, self.N), order="C")
        L.check(L.lib().cpm_city_get_cdf(self._h, _ptr(out)))
        return out

    def guide(self, hour: int):
        """Guide tables of one table hour: uint16 [N, G+1], guide[o, b] = #{j : cdf[o, j] < b/G} (None without a guide)."""
        G = C.c_int32()
        L.check(L.lib().cpm_city_get_guide(self._h, int(hour), None, C.byref(G)))
        if G.value == 0:
            return None
        out = np.empty((self.N, G.value + 1), dtype=np.uint16)
        L.check(L.lib().cpm_city_get_guide(self._h, int(hour), _ptr(out), C.byref(G)))
        return out

    def close(self):
        if self._h:
            L.lib().cpm_city_destroy(self._h)
            self._h = C.c_void_p()


class Fleet:
    """Device-resident car state (one int32 zone per car) and the per-zone counts of the last day."""

    def __init__(self, ctx: Context, C_: int):
        self.ctx, self.C = ctx, int(C_)
        self._h = C.c_void_p()
        L.check(L.lib().cpm_fleet_create(ctx._h, self.C, C.byref(self._h)))
        self._shape = None

    def init(self, cars_per_zone: int, car_offset: int = 0):
        L.check(L.lib().cpm_fleet_init(self._h, int(cars_per_zone), int(car_offset)))
        return self

    def set_zones(self, zones):
        z = np.ascontiguousarray(zones, dtype=np.int64)
        assert z.shape == (self.C,)
        L.check(L.lib().cpm_fleet_set_zones(self._h, _ptr(z)))
        return self

    def zones(self):
        out = np.empty(self.C, dtype=np.int64)
        L.check(L.lib().cpm_fleet_get_zones(self._h, _ptr(out)))
        return out

    def run_day(self, city: City, n_steps: int, *, with_time: bool, rng: Rng, steps_per_hour: int = 1,
                sampler: int = L.SAMPLER_AUTO):
        o = L.CDayOpts()
        o.n_steps, o.steps_per_hour, o.with_time, o.sampler = int(n_steps), int(steps_per_hour), int(with_time), sampler
        o.rng = rng.to_c()
        L.check(L.lib().cpm_run_day(city._h, self._h, C.byref(o)))
        self._shape = (city.N, int(n_steps))
        return self

    def counts(self):
        """(parking int64[N,n_steps], driving int64[N,n_steps], time_sum, n_lost) of the last day."""
        N, S = self._shape
        parking = np.empty((N, S), dtype=np.int64, order="F")
        driving = np.empty((N, S), dtype=np.int64, order="F")
        ts, lost = C.c_double(), C.c_int64()
        L.check(L.lib().cpm_fleet_get_counts(self._h, _ptr(parking), _ptr(driving), C.byref(ts), C.byref(lost)))
        return parking, driving, ts.value, lost.value

    @property
    def last_sampler(self) -> int:
        """SAMPLER_GATHER / SAMPLER_BUCKET / SAMPLER_CARS: what the last day ran on (what SAMPLER_AUTO resolved to)."""
        v = C.c_int32()
        L.check(L.lib().cpm_fleet_last_sampler(self._h, C.byref(v)))
        return v.value

    def time_fixed(self):
        """(whole seconds, 2^-32 s units) of the last day's driving-time sum: exact integers."""
        s, f = C.c_int64(), C.c_int64()
        L.check(L.lib().cpm_fleet_get_time_fixed(self._h, C.byref(s), C.byref(f)))
        return s.value, f.value

    def counts_device(self):
        """(device pointer, n_elems) of the uint32 [2][N,n_steps] count block, for an NCCL all-reduce."""
        p, n = C.c_void_p(), C.c_int64()
        L.check(L.lib().cpm_fleet_counts_device(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def close(self):
        if self._h:
            L.lib().cpm_fleet_destroy(self._h)
            self._h = C.c_void_p()


def sample_dataset(city: City, fleet: Fleet, *, e_drive=0.5, e_dest=2, p_min=0.1, p_max=0.9, cars_per_zone=1000,
                   seed=0, sampler=L.SAMPLER_AUTO, want_parking=True):
    """main.jl:82-98 for one dataset on the device.  Returns (parking[N,T], traffic[T], A_drive, n_lost)."""
    pr = L.CParams(float(e_drive), float(e_dest), float(p_min), float(p_max))
    parking = np.empty((city.N, city.T), order="F") if want_parking else None
    traffic = np.empty(city.T)
    A, lost = C.c_double(), C.c_int64()
    L.check(L.lib().cpm_sample_dataset(city._h, fleet._h, C.byref(pr), int(cars_per_zone), int(seed), sampler,
                                       _ptr(parking), _ptr(traffic), C.byref(A), C.byref(lost)))
    return parking, traffic, A.value, lost.value


def sample_dataset_host(city: City, fleet: Fleet, datamatrix, distance_matrix_km, *, e_drive=0.5, e_dest=2, p_min=0.1, p_max=0.9,
                        cars_per_zone=1000, seed=0, sampler=L.SAMPLER_AUTO, want_parking=True):
    """main.jl:82-98 for one dataset from HOST arrays, the upload pipelined with the computation.
    Returns (parking[N,T], traffic[T], A_drive, n_lost)."""
    dm = _f64F(datamatrix, (city.N, city.N, city.T, 2))
    dist = _f64F(distance_matrix_km, (city.N, city.N))
    pr = L.CParams(float(e_drive), float(e_dest), float(p_min), float(p_max))
    parking = np.empty((city.N, city.T), order="F") if want_parking else None
    traffic = np.empty(city.T)
    A, lost = C.c_double(), C.c_int64()
    L.check(L.lib().cpm_sample_dataset_host(city._h, fleet._h, _ptr(dm), _ptr(dist), C.byref(pr), int(cars_per_zone), int(seed),
                                            sampler, _ptr(parking), _ptr(traffic), C.byref(A), C.byref(lost)))
    return parking, traffic, A.value, lost.value


def sample_dataset_rows(city: City, fleet: Fleet, sourceid, dstid, hod, mean, stddev, distance_matrix_km, *, e_drive=0.5, e_dest=2,
                        p_min=0.1, p_max=0.9, cars_per_zone=1000, seed=0, sampler=L.SAMPLER_AUTO, want_parking=True):
    """main.jl:74-98 for one dataset from the CSV columns (createdatamatrix on the device, src/createdatamatrix.jl:8-23),
    the upload pipelined with the computation.  Returns (parking[N,T], traffic[T], A_drive, n_lost, n_skipped_rows)."""
    cols = [np.ascontiguousarray(a, dtype=np.int32) for a in (sourceid, dstid, hod)]
    vals = [np.ascontiguousarray(a, dtype=np.float64) for a in (mean, stddev)]
    n = cols[0].shape[0]
    assert all(a.shape == (n,) for a in cols + vals)
    dist = _f64F(distance_matrix_km, (city.N, city.N))
    pr = L.CParams(float(e_drive), float(e_dest), float(p_min), float(p_max))
    parking = np.empty((city.N, city.T), order="F") if want_parking else None
    traffic = np.empty(city.T)
    A, lost, skipped = C.c_double(), C.c_int64(), C.c_int64()
    L.check(L.lib().cpm_sample_dataset_rows(city._h, fleet._h, _ptr(cols[0]), _ptr(cols[1]), _ptr(cols[2]), _ptr(vals[0]), _ptr(vals[1]), n,
                                            _ptr(dist), C.byref(pr), int(cars_per_zone), int(seed), sampler, _ptr(parking),
                                            _ptr(traffic), C.byref(A), C.byref(lost), C.byref(skipped)))
    return parking, traffic, A.value, lost.value, skipped.value


def sweep(city: City, initial_zones, params, *, seed=0, common_random_numbers=False, want_parking=False, streams=None):
    """Model-selection sweep over P parameter sets (rows of `params`: e_drive, e_dest, p_min, p_max).
    `streams`: Philox stream id per set (default 2 + p; all equal with common_random_numbers).
    Returns (traffic_counts int64[T,P], A_drive[P], parking[N,T,P] or None)."""
    params = np.asarray(params, dtype=np.float64).reshape(-1, 4)
    P = params.shape[0]
    arr = (L.CParams * P)(*[L.CParams(*map(float, row)) for row in params])
    z = np.ascontiguousarray(initial_zones, dtype=np.int64)
    tc = np.empty((city.T, P), dtype=np.int64, order="F")
    A = np.empty(P)
    parking = np.empty((city.N, city.T, P), order="F") if want_parking else None
    if streams is None and common_random_numbers:
        streams = np.ones(P, dtype=np.uint32)
    st = None if streams is None else np.ascontiguousarray(streams, dtype=np.uint32)
    assert st is None or st.shape == (P,)
    L.check(L.lib().cpm_sweep(city._h, _ptr(z), z.shape[0], arr, P, int(seed), _ptr(st), _ptr(tc), _ptr(A), _ptr(parking)))
    return tc, A, parking
