"""Host-side mirror of the reference's Julia interface for the hot path.

Same function names, argument order and mutation semantics as the seven global functions that
main.jl:11-22 of the reference `include`s; arrays are numpy arrays in Julia's memory layout
(Fortran order, 1-based zone values, Int64 state, Float64 elsewhere).  Like the Julia functions, they
read the model globals at call time (src/createpdrive.jl:4,31; src/createpdestin.jl:4,24;
src/initializestates.jl:6,11): here they are attributes of this module, with main.jl:37-42's defaults.

Every function is a ctypes call into libcpm.so (include/cpm.h); there is no CPU path.  The extra
keyword `rng` (default: native Philox) exists because the reference draws from Julia's unseeded
global RNG, which cannot be reproduced; tests inject the oracle's recorded variates through it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .engine import Context, Rng, _ptr

# --- the reference's globals (main.jl:37-42) ---
e_drive = 0.5
e_dest = 2
p_min = 0.1
p_max = 0.9
cars_per_zone = 1000
T = 24

_ctx: Context | None = None
_stream_counter = 0


def context() -> Context:
    """The module-level context (GPU 0) the Julia-named functions run on."""
    global _ctx
    if _ctx is None:
        _ctx = Context(0)
    return _ctx


def set_context(ctx: Context | None):
    global _ctx
    _ctx = ctx


def _next_rng(rng):
    global _stream_counter
    if rng is not None:
        return rng
    _stream_counter += 1
    return Rng(mode="philox", seed=0, stream=_stream_counter)


def _chkF(a, dtype, name):
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags.f_contiguous:
        raise TypeError(f"{name} must be a Fortran-ordered numpy array of {np.dtype(dtype).name} (Julia layout)")
    return a


def createpdrive(datamatrix, distance_matrix_km, number_zones):
    """src/createpdrive.jl:3-38 -> p_drive[N,T]"""
    N = int(number_zones)
    dm = _chkF(datamatrix, np.float64, "datamatrix")
    dist = _chkF(distance_matrix_km, np.float64, "distance_matrix_km")
    assert dm.shape == (N, N, T, 2) and dist.shape == (N, N)
    out = np.empty((N, T), order="F")
    L.check(L.lib().cpm_create_pdrive(context()._h, _ptr(dm), _ptr(dist), N, T, float(p_min), float(p_max),
                                      float(e_drive), _ptr(out)))
    return out


def createpdestin(datamatrix, number_zones):
    """src/createpdestin.jl:3-50 -> p_dest[N,N,T]"""
    N = int(number_zones)
    dm = _chkF(datamatrix, np.float64, "datamatrix")
    assert dm.shape == (N, N, T, 2)
    out = np.empty((N, N, T), order="F")
    L.check(L.lib().cpm_create_pdestin(context()._h, _ptr(dm), N, T, float(e_dest), _ptr(out)))
    return out


def initializestates(C_):
    """src/initializestates.jl:4-22 -> (state_matrix Int64[C,T], transition_matrix Float64[C,T,4])"""
    C_ = int(C_)
    state = np.empty((C_, T), dtype=np.int64, order="F")
    trans = np.empty((C_, T, 4), dtype=np.float64, order="F")
    L.check(L.lib().cpm_initialize_states(context()._h, C_, T, int(cars_per_zone), _ptr(state), _ptr(trans)))
    return state, trans


def solveinitialvalueproblem(state_matrix, transition_matrix, p_drive, p_dest, C_, number_zones, *, rng: Rng | None = None):
    """src/solveinitialvalueproblem.jl:4-62.  Mutates state_matrix / transition_matrix in place; returns initial_state."""
    C_, N = int(C_), int(number_zones)
    state = _chkF(state_matrix, np.int64, "state_matrix")
    trans = _chkF(transition_matrix, np.float64, "transition_matrix")
    pd, px = _chkF(p_drive, np.float64, "p_drive"), _chkF(p_dest, np.float64, "p_dest")
    assert state.shape == (C_, T) and trans.shape == (C_, T, 4) and pd.shape == (N, T) and px.shape == (N, N, T)
    init = np.empty(C_, dtype=np.int64)
    lost = C.c_int64()
    r = _next_rng(rng).to_c()
    L.check(L.lib().cpm_solve_ivp(context()._h, _ptr(state), _ptr(trans), _ptr(pd), _ptr(px), C_, N, T, C.byref(r),
                                  _ptr(init), C.byref(lost)))
    solveinitialvalueproblem.last_lost = lost.value
    for t in range(1, T):   # the reference prints progress from inside the loop (:54)
        if verbose:
            print(f"{int(t / 0.23)}% of the Initial Value Problem is solved")
    return init


def resampling(state_matrix, transition_matrix, C_, number_zones, p_drive, p_dest, datamatrix, distance_matrix_km, *,
               rng: Rng | None = None):
    """src/resampling.jl:3-89.  Mutates and returns (state_matrix, transition_matrix)."""
    C_, N = int(C_), int(number_zones)
    state = _chkF(state_matrix, np.int64, "state_matrix")
    trans = _chkF(transition_matrix, np.float64, "transition_matrix")
    pd, px = _chkF(p_drive, np.float64, "p_drive"), _chkF(p_dest, np.float64, "p_dest")
    dm, dist = _chkF(datamatrix, np.float64, "datamatrix"), _chkF(distance_matrix_km, np.float64, "distance_matrix_km")
    assert state.shape == (C_, T) and trans.shape == (C_, T, 4) and pd.shape == (N, T) and px.shape == (N, N, T)
    assert dm.shape == (N, N, T, 2) and dist.shape == (N, N)
    lost = C.c_int64()
    r = _next_rng(rng).to_c()
    L.check(L.lib().cpm_resampling(context()._h, _ptr(state), _ptr(trans), C_, N, T, _ptr(pd), _ptr(px), _ptr(dm),
                                   _ptr(dist), C.byref(r), C.byref(lost)))
    resampling.last_lost = lost.value
    for t in range(1, T + 1):   # :84
        if verbose:
            print(f"{int(t / 0.23)}% of the dataset is simulated")
    return state, trans


def averagedrivingtime(C_, A_drive, transition_matrix):
    """src/averagedrivingtime.jl:3-12"""
    trans = _chkF(transition_matrix, np.float64, "transition_matrix")
    assert trans.shape == (int(C_), T, 4)
    out = C.c_double()
    L.check(L.lib().cpm_average_driving_time(context()._h, int(C_), T, float(A_drive), _ptr(trans), C.byref(out)))
    return out.value


def correctparameters(p_min_next, p_max_next, p_min_, p_max_):
    """src/correctparameters.jl:3-23"""
    out = np.empty(2)
    L.check(L.lib().cpm_correct_parameters(float(p_min_next), float(p_max_next), float(p_min_), float(p_max_), _ptr(out)))
    return float(out[0]), float(out[1])


def zonehistogram(number_zones, state_matrix, transition_matrix, C_):
    """The histogram + normalisation block of saveresults (src/saveresults.jl:6-28), argument order of
    saveresults(number_zones, state_matrix, transition_matrix, ..., C).
    Returns (parking_cars[N,T] densities, driving_cars[N,T] counts, traffic_resultsmatrix[T])."""
    N, C_ = int(number_zones), int(C_)
    state = _chkF(state_matrix, np.int64, "state_matrix")
    trans = _chkF(transition_matrix, np.float64, "transition_matrix")
    parking = np.empty((N, T), order="F")
    driving = np.empty((N, T), order="F")
    traffic = np.empty(T)
    L.check(L.lib().cpm_zone_histogram(context()._h, _ptr(state), _ptr(trans), C_, N, T, _ptr(parking), _ptr(driving),
                                       _ptr(traffic)))
    return parking, driving, traffic


def distancematrix(centroid_lat, centroid_long):
    """The distance-matrix block of processgeodata (src/processgeodata.jl:150-166) -> distance_matrix_km[N,N]."""
    lat = np.ascontiguousarray(centroid_lat, dtype=np.float64).ravel()
    lon = np.ascontiguousarray(centroid_long, dtype=np.float64).ravel()
    assert lat.shape == lon.shape
    N = lat.shape[0]
    out = np.empty((N, N), order="F")
    L.check(L.lib().cpm_distance_matrix(context()._h, _ptr(lat), _ptr(lon), N, _ptr(out)))
    return out


def polygoncentroids(longitude_matrix, latitude_matrix):
    """The polygon block of processgeodata (src/processgeodata.jl:99-148): zero-terminated vertex lists [N,V] ->
    (centroid_lat[N], centroid_long[N], area_polygon[N])."""
    lon = _chkF(longitude_matrix, np.float64, "longitude_matrix")
    lat = _chkF(latitude_matrix, np.float64, "latitude_matrix")
    assert lon.ndim == 2 and lon.shape == lat.shape
    N, V = lon.shape
    c_lat, c_lon, area = np.empty(N), np.empty(N), np.empty(N)
    L.check(L.lib().cpm_polygon_centroids(context()._h, _ptr(lon), _ptr(lat), N, V, _ptr(c_lat), _ptr(c_lon), _ptr(area)))
    return c_lat, c_lon, area


verbose = False   # print the reference's progress lines (src/resampling.jl:84, src/solveinitialvalueproblem.jl:54)
