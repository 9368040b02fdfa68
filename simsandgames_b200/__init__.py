"""simsandgames_b200 -- B200-native grid-fluid step for simsAndGames' flip_fluid / eulerian_fluid.

The product is libflipgpu.so (hand-written sm_100a CUDA behind the C ABI of include/flipgpu.h).
This package is the thin Python mirror of the reference's per-sim interface used by tests and
bench.py; the C++ mirror is simsandgames_b200/host/flip_fluid_gpu.hpp.  There is no CPU fallback:
importing works anywhere, creating a simulation without the built library or without a CUDA
device raises.
"""
from .binding import (FlipFluidGPU, FluidGPU, Fluid3DGPU, FlipGpuError, lib_path, load_library, build_library,
                      ShardedProjection, FlipFluidSharded, nccl_unique_id, slab_rows, FLIP_FIELDS, EULER_FIELDS, RED_BLACK, LEX_WAVEFRONT, LEX_DIAGONAL, LEX_SYSTOLIC)
from . import scenes

__all__ = ["FlipFluidGPU", "FluidGPU", "Fluid3DGPU", "FlipGpuError", "lib_path", "load_library", "build_library", "ShardedProjection", "FlipFluidSharded", "nccl_unique_id", "slab_rows",
           "FLIP_FIELDS", "EULER_FIELDS", "RED_BLACK", "LEX_WAVEFRONT", "LEX_DIAGONAL", "LEX_SYSTOLIC", "scenes"]
