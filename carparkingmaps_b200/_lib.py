"""ctypes binding of libcpm.so (include/cpm.h).  Fails loudly: there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("CPM_LIBCPM") or os.path.join(_HERE, "libcpm.so")   # CPM_LIBCPM: a debug build of the same library
CSRC = os.path.join(_HERE, "csrc")

CPM_OK = 0
RNG_PHILOX, RNG_INJECTED = 0, 1
SAMPLER_AUTO, SAMPLER_GATHER, SAMPLER_BUCKET, SAMPLER_CARS = 0, 1, 2, 3


class CpmError(RuntimeError):
    pass


class CRng(C.Structure):
    _fields_ = [("mode", C.c_int32), ("stream", C.c_uint32), ("seed", C.c_uint64), ("car_offset", C.c_int64),
                ("u_drive", C.c_void_p), ("u_dest", C.c_void_p), ("u_time", C.c_void_p), ("u_dist", C.c_void_p)]


class CDayOpts(C.Structure):
    _fields_ = [("n_steps", C.c_int64), ("steps_per_hour", C.c_int64), ("with_time", C.c_int32),
                ("sampler", C.c_int32), ("rng", CRng)]


class CParams(C.Structure):
    _fields_ = [("e_drive", C.c_double), ("e_dest", C.c_double), ("p_min", C.c_double), ("p_max", C.c_double)]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libcpm.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j8"] + (["-B"] if force else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise CpmError("building libcpm.so failed")
    return SO_PATH


# name -> (restype, argtypes); every symbol include/cpm.h declares
_vp, _i64, _i32, _d, _u64 = C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_uint64
_pp = C.POINTER(C.c_void_p)
SIGNATURES = {
    "cpm_version": (_i32, []),
    "cpm_last_error": (C.c_char_p, []),
    "cpm_device_count": (_i32, [C.POINTER(_i32)]),
    "cpm_ctx_create": (_i32, [_i32, _vp, _pp]),
    "cpm_ctx_destroy": (_i32, [_vp]),
    "cpm_ctx_sync": (_i32, [_vp]),
    "cpm_ctx_stream": (_i32, [_vp, _pp]),
    "cpm_ctx_launch_count": (_i32, [_vp, C.POINTER(_i64)]),
    "cpm_ctx_profile_begin": (_i32, [_vp]),
    "cpm_ctx_profile_end": (_i32, [_vp, C.POINTER(_d), C.POINTER(_i64)]),
    "cpm_ctx_set_input_cache": (_i32, [_vp, _i32]),
    "cpm_ctx_uploads_skipped": (_i32, [_vp, C.POINTER(_i64)]),
    "cpm_ctx_graph_replays": (_i32, [_vp, C.POINTER(_i64)]),
    "cpm_ctx_set_graphs": (_i32, [_vp, _i32]),
    "cpm_ctx_host_last_sampler": (_i32, [_vp, C.POINTER(_i32)]),
    "cpm_host_alloc": (_i32, [_i64, _pp]),
    "cpm_host_free": (_i32, [_vp]),
    "cpm_host_register": (_i32, [_vp, _i64]),
    "cpm_host_unregister": (_i32, [_vp]),
    "cpm_create_pdrive": (_i32, [_vp, _vp, _vp, _i64, _i64, _d, _d, _d, _vp]),
    "cpm_create_pdestin": (_i32, [_vp, _vp, _i64, _i64, _d, _vp]),
    "cpm_initialize_states": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "cpm_solve_ivp": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, C.POINTER(CRng), _vp, C.POINTER(_i64)]),
    "cpm_resampling": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, C.POINTER(CRng), C.POINTER(_i64)]),
    "cpm_average_driving_time": (_i32, [_vp, _i64, _i64, _d, _vp, C.POINTER(_d)]),
    "cpm_correct_parameters": (_i32, [_d, _d, _d, _d, _vp]),
    "cpm_zone_histogram": (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "cpm_distance_matrix": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "cpm_polygon_centroids": (_i32, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "cpm_city_create": (_i32, [_vp, _i64, _i64, _pp]),
    "cpm_city_destroy": (_i32, [_vp]),
    "cpm_city_upload": (_i32, [_vp, _vp, _vp]),
    "cpm_city_upload_rows": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, C.POINTER(_i64)]),
    "cpm_city_synth": (_i32, [_vp, _u64, _d]),
    "cpm_city_download_inputs": (_i32, [_vp, _vp, _vp]),
    "cpm_city_release_tables": (_i32, [_vp]),
    "cpm_city_prepare_multi_day": (_i32, [_vp]),
    "cpm_city_build_pdrive": (_i32, [_vp, _d, _d, _d]),
    "cpm_city_build_pdestin": (_i32, [_vp, _d, _i32]),
    "cpm_city_set_pdrive": (_i32, [_vp, _vp]),
    "cpm_city_set_pdest": (_i32, [_vp, _vp]),
    "cpm_city_get_pdrive": (_i32, [_vp, _vp]),
    "cpm_city_get_pdest": (_i32, [_vp, _vp]),
    "cpm_city_get_cdf": (_i32, [_vp, _vp]),
    "cpm_city_get_guide": (_i32, [_vp, _i64, _vp, C.POINTER(_i32)]),
    "cpm_fleet_create": (_i32, [_vp, _i64, _pp]),
    "cpm_fleet_destroy": (_i32, [_vp]),
    "cpm_fleet_init": (_i32, [_vp, _i64, _i64]),
    "cpm_fleet_set_zones": (_i32, [_vp, _vp]),
    "cpm_fleet_get_zones": (_i32, [_vp, _vp]),
    "cpm_fleet_copy_zones": (_i32, [_vp, _vp]),
    "cpm_run_day": (_i32, [_vp, _vp, C.POINTER(CDayOpts)]),
    "cpm_fleet_get_counts": (_i32, [_vp, _vp, _vp, C.POINTER(_d), C.POINTER(_i64)]),
    "cpm_fleet_last_sampler": (_i32, [_vp, C.POINTER(_i32)]),
    "cpm_fleet_get_time_fixed": (_i32, [_vp, C.POINTER(_i64), C.POINTER(_i64)]),
    "cpm_fleet_counts_device": (_i32, [_vp, _pp, C.POINTER(_i64)]),
    "cpm_sample_dataset": (_i32, [_vp, _vp, C.POINTER(CParams), _i64, _u64, _i32, _vp, _vp, C.POINTER(_d), C.POINTER(_i64)]),
    "cpm_sample_dataset_host": (_i32, [_vp, _vp, _vp, _vp, C.POINTER(CParams), _i64, _u64, _i32, _vp, _vp, C.POINTER(_d), C.POINTER(_i64)]),
    "cpm_sample_dataset_rows": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, C.POINTER(CParams), _i64, _u64, _i32, _vp, _vp,
                                C.POINTER(_d), C.POINTER(_i64), C.POINTER(_i64)]),
    "cpm_sweep": (_i32, [_vp, _vp, _i64, C.POINTER(CParams), _i64, _u64, _vp, _vp, _vp, _vp]),
    "cpm_debug_bucket_geometry": (_i32, [_i64, _i64, _i32, _vp]),
    "cpm_debug_cars_geometry": (_i32, [_i64, _i64, _i32, C.POINTER(_i64)]),
    "cpm_sample_datasets_multi": (_i32, [_vp, _i32, _i64, _vp, _i64, _vp, _vp, C.POINTER(CParams), _i64, _u64, _vp, _vp, _vp, _vp]),
    "cpm_split_city_multi": (_i32, [_vp, _i32, _i64, _i64, _vp, _vp, C.POINTER(CParams), _i64, _i64, _i64, _i32, _u64, C.c_uint32,
                             _vp, _vp, C.POINTER(_d), C.POINTER(_i64), C.POINTER(_d)]),
}

_lib = None


def lib():
    """The loaded library.  Raises if it has not been built: the product has no other path."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise CpmError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(or make -C carparkingmaps_b200/csrc). There is no CPU fallback.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int):
    if rc != CPM_OK:
        raise CpmError(f"libcpm error {rc}: {lib().cpm_last_error().decode(errors='replace')}")
