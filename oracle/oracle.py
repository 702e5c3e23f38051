"""ctypes binding of the CPU ORACLE (oracle/cpm_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package never imports this module.

Arrays follow Julia's layout: numpy arrays in Fortran order (column-major), zone IDs are
1-based values, state is int64, everything else float64.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

RNG_STREAM, RNG_INJECTED, RNG_PHILOX = 0, 1, 2


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cpm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


class _Rng(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("stream", C.c_uint32), ("seed", C.c_uint64), ("car_offset", C.c_int64),
        ("u_drive", C.c_void_p), ("u_dest", C.c_void_p), ("u_time", C.c_void_p), ("u_dist", C.c_void_p),
        ("s", C.c_uint64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp, ip, i64, d = C.c_void_p, C.c_void_p, C.c_int64, C.c_double
        L.orc_createpdrive.argtypes = [dp, dp, i64, i64, d, d, d, dp]
        L.orc_createpdestin.argtypes = [dp, i64, i64, d, dp]
        L.orc_initializestates.argtypes = [i64, i64, i64, ip, dp]
        L.orc_solveinitialvalueproblem.argtypes = [ip, dp, dp, dp, i64, i64, i64, C.POINTER(_Rng), ip]
        L.orc_solveinitialvalueproblem.restype = i64
        L.orc_resampling.argtypes = [ip, dp, i64, i64, i64, dp, dp, dp, dp, C.POINTER(_Rng)]
        L.orc_resampling.restype = i64
        L.orc_averagedrivingtime.argtypes = [i64, i64, d, dp]
        L.orc_averagedrivingtime.restype = d
        L.orc_correctparameters.argtypes = [d, d, d, d, dp]
        L.orc_zonehistogram.argtypes = [i64, i64, i64, ip, dp, dp, dp, dp]
        L.orc_truncnormal.argtypes = [d, d, d]
        L.orc_truncnormal.restype = d
        L.orc_erfinv_export.argtypes = [d]
        L.orc_erfinv_export.restype = d
        L.orc_philox_pair.argtypes = [C.c_uint64, C.c_uint32, i64, C.c_uint32, C.c_uint32, dp, dp]
        L.orc_day_counts.argtypes = [ip, i64, i64, i64, i64, i64, dp, dp, dp, dp, C.POINTER(_Rng),
                                     C.c_int, ip, ip, dp]
        L.orc_day_counts.restype = i64
        L.orc_rng_init.argtypes = [C.POINTER(_Rng)]
        L.orc_createdatamatrix.argtypes = [ip, ip, ip, dp, dp, i64, i64, i64, dp]
        L.orc_createdatamatrix.restype = i64
        L.orc_distancematrix.argtypes = [dp, dp, i64, dp]
        L.orc_polygoncentroids.argtypes = [dp, dp, i64, i64, dp, dp, dp]
        L.orc_polygoncentroids.restype = None
        L.orc_jl_sum.argtypes = [dp, i64]
        L.orc_jl_sum.restype = C.c_double
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    a = np.asarray(a)
    assert a.dtype == np.float64 and a.flags.f_contiguous, "need Fortran-ordered float64"
    return a


@dataclass
class Rng:
    """Variate source.  mode 'stream': sequential draws in the reference's order (recorded into the
    u_* arrays when given); 'injected': u_*[i,t] are inputs; 'philox': counter-based, keyed by
    (seed, stream, global car index, step)."""
    mode: str = "stream"
    seed: int = 0
    stream: int = 0
    car_offset: int = 0
    u_drive: np.ndarray | None = None
    u_dest: np.ndarray | None = None
    u_time: np.ndarray | None = None
    u_dist: np.ndarray | None = None

    def _c(self):
        r = _Rng()
        r.mode = {"stream": RNG_STREAM, "injected": RNG_INJECTED, "philox": RNG_PHILOX}[self.mode]
        r.seed, r.stream, r.car_offset = self.seed, self.stream, self.car_offset
        for k in ("u_drive", "u_dest", "u_time", "u_dist"):
            a = getattr(self, k)
            if a is not None:
                _f64(a)
            setattr(r, k, None if a is None else a.ctypes.data)
        lib().orc_rng_init(C.byref(r))
        return r

    @staticmethod
    def recording(C_, T, seed=0, mode="stream", stream=0):
        """A stream/philox source that records every variate it consumes (NaN where none was drawn)."""
        mk = lambda: np.full((C_, T), np.nan, dtype=np.float64, order="F")
        return Rng(mode=mode, seed=seed, stream=stream, u_drive=mk(), u_dest=mk(), u_time=mk(), u_dist=mk())

    def as_injected(self):
        return Rng(mode="injected", u_drive=self.u_drive, u_dest=self.u_dest, u_time=self.u_time, u_dist=self.u_dist)


def distancematrix(centroid_lat, centroid_long):
    """src/processgeodata.jl:150-166 -> distance_matrix_km[N,N]"""
    lat = np.ascontiguousarray(centroid_lat, dtype=np.float64).ravel()
    lon = np.ascontiguousarray(centroid_long, dtype=np.float64).ravel()
    out = np.zeros((lat.shape[0], lat.shape[0]), order="F")
    lib().orc_distancematrix(_p(lat), _p(lon), lat.shape[0], _p(out))
    return out


def jl_sum(x):
    """Julia's sum(::Vector{Float64}): pairwise above 1024 elements (base/reduce.jl mapreduce_impl)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    return float(lib().orc_jl_sum(_p(x), x.shape[0]))


def polygoncentroids(longitude_matrix, latitude_matrix):
    """src/processgeodata.jl:99-148 -> (centroid_lat[N], centroid_long[N], area_polygon[N]).
    The matrices are [N,V] zero-terminated vertex lists; copies are taken (the reference closes the rings in place)."""
    lon = np.array(longitude_matrix, dtype=np.float64, order="F", copy=True)
    lat = np.array(latitude_matrix, dtype=np.float64, order="F", copy=True)
    assert lon.shape == lat.shape and lon.ndim == 2
    N, V = lon.shape
    c_lat, c_lon, area = np.zeros(N), np.zeros(N), np.zeros(N)
    lib().orc_polygoncentroids(_p(lon), _p(lat), N, V, _p(c_lat), _p(c_lon), _p(area))
    return c_lat, c_lon, area


def createdatamatrix(sourceid, dstid, hod, mean, stddev, number_zones, *, T):
    """The scatter of src/createdatamatrix.jl:7-23 from the CSV columns -> (datamatrix[N,N,T,2], #skipped rows)."""
    N = int(number_zones)
    cols = [np.ascontiguousarray(a, dtype=np.int32) for a in (sourceid, dstid, hod)]
    vals = [np.ascontiguousarray(a, dtype=np.float64) for a in (mean, stddev)]
    dm = np.zeros((N, N, T, 2), order="F")
    bad = lib().orc_createdatamatrix(_p(cols[0]), _p(cols[1]), _p(cols[2]), _p(vals[0]), _p(vals[1]), cols[0].shape[0], N, T, _p(dm))
    return dm, bad


def createpdrive(datamatrix, distance_matrix_km, number_zones, *, T, p_min, p_max, e_drive):
    dm, dist = _f64(datamatrix), _f64(distance_matrix_km)
    N = int(number_zones)
    out = np.zeros((N, T), dtype=np.float64, order="F")
    lib().orc_createpdrive(_p(dm), _p(dist), N, T, p_min, p_max, float(e_drive), _p(out))
    return out


def createpdestin(datamatrix, number_zones, *, T, e_dest):
    dm = _f64(datamatrix)
    N = int(number_zones)
    out = np.zeros((N, N, T), dtype=np.float64, order="F")
    lib().orc_createpdestin(_p(dm), N, T, float(e_dest), _p(out))
    return out


def initializestates(C_, *, T, cars_per_zone):
    state = np.zeros((C_, T), dtype=np.int64, order="F")
    trans = np.zeros((C_, T, 4), dtype=np.float64, order="F")
    lib().orc_initializestates(C_, T, cars_per_zone, _p(state), _p(trans))
    return state, trans


def solveinitialvalueproblem(state, trans, p_drive, p_dest, C_, number_zones, *, T, rng: Rng):
    init = np.zeros(C_, dtype=np.int64)
    r = rng._c()
    lost = lib().orc_solveinitialvalueproblem(_p(state), _p(trans), _p(_f64(p_drive)), _p(_f64(p_dest)),
                                              C_, int(number_zones), T, C.byref(r), _p(init))
    solveinitialvalueproblem.last_lost = lost
    return init


def resampling(state, trans, C_, number_zones, p_drive, p_dest, datamatrix, distance_matrix_km, *, T, rng: Rng):
    r = rng._c()
    lost = lib().orc_resampling(_p(state), _p(trans), C_, int(number_zones), T, _p(_f64(p_drive)),
                                _p(_f64(p_dest)), _p(_f64(datamatrix)), _p(_f64(distance_matrix_km)), C.byref(r))
    resampling.last_lost = lost
    return state, trans


def averagedrivingtime(C_, A_drive, trans, *, T):
    return lib().orc_averagedrivingtime(C_, T, float(A_drive), _p(_f64(trans)))


def correctparameters(p_min_next, p_max_next, p_min, p_max):
    out = np.zeros(2)
    lib().orc_correctparameters(p_min_next, p_max_next, p_min, p_max, _p(out))
    return float(out[0]), float(out[1])


def zonehistogram(number_zones, state, trans, C_, *, T):
    N = int(number_zones)
    parking = np.zeros((N, T), order="F")
    driving = np.zeros((N, T), order="F")
    traffic = np.zeros(T)
    with np.errstate(all="ignore"):
        lib().orc_zonehistogram(N, T, C_, _p(state), _p(_f64(trans)), _p(parking), _p(driving), _p(traffic))
    return parking, driving, traffic


def truncnormal(mu, sigma, u):
    return lib().orc_truncnormal(mu, sigma, u)


def erfinv(y):
    return lib().orc_erfinv_export(y)


def philox_pair(seed, stream, car, step, phase):
    a, b = C.c_double(), C.c_double()
    lib().orc_philox_pair(seed, stream, car, step, phase, C.byref(a), C.byref(b))
    return a.value, b.value


def day_counts(zone, number_zones, T_tab, n_steps, steps_per_hour, p_drive, p_dest, datamatrix,
               distance_matrix_km, rng: Rng, with_time: bool):
    """Counts-only day: returns (parking[N,n_steps] int64, driving[N,n_steps] int64, time_sum, lost);
    `zone` (int64[C]) is advanced in place."""
    N = int(number_zones)
    C_ = zone.shape[0]
    parking = np.zeros((N, n_steps), dtype=np.int64, order="F")
    driving = np.zeros((N, n_steps), dtype=np.int64, order="F")
    ts = C.c_double(0.0)
    r = rng._c()
    lost = lib().orc_day_counts(_p(zone), C_, N, T_tab, n_steps, steps_per_hour, _p(_f64(p_drive)), _p(_f64(p_dest)),
                                _p(None if datamatrix is None else _f64(datamatrix)),
                                _p(None if distance_matrix_km is None else _f64(distance_matrix_km)),
                                C.byref(r), int(with_time), _p(parking), _p(driving), C.byref(ts))
    return parking, driving, ts.value, lost
