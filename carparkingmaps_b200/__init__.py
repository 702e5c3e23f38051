"""carparkingmaps_b200 -- B200-native (sm_100a) HMM sampling hot path of CarParkingMaps.

Layout: csrc/ (CUDA kernels + the C ABI of include/cpm.h, built in-tree to libcpm.so),
api.py (mirror of the reference's Julia function names), engine.py (device-resident handles),
synth.py (synthetic inputs), modelselection.py (the notebook's update rules on top of the sweep),
julia/CarParkingMapsGPU.jl (the ccall shim a Julia host includes instead of src/*.jl).
"""
from . import _lib
from ._lib import CpmError, SAMPLER_AUTO, SAMPLER_BUCKET, SAMPLER_CARS, SAMPLER_GATHER, build
from .engine import City, Context, Fleet, Rng, sample_dataset, sample_dataset_host, sample_dataset_rows, sweep

__all__ = ["City", "Context", "Fleet", "Rng", "sample_dataset", "sample_dataset_host", "sample_dataset_rows", "sweep", "build", "CpmError",
           "SAMPLER_AUTO", "SAMPLER_GATHER", "SAMPLER_BUCKET", "SAMPLER_CARS"]
