"""raytracinggpu_b200 — B200-native (sm_100a) implementation of the per-pixel render hot path of
souhhcong/RaytracingGPU behind a C ABI (include/rt_b200.h, librtb200.so).

Python here is only the test/bench harness binding; the product is the shared library built from csrc/.
"""
from ._abi import (RT_RENDER_COUNT_WORK, RT_RENDER_NO_SYNC, rt_params, rt_sphere, rt_stats)  # noqa: F401
from . import sharding  # noqa: F401
from .api import (LIB_PATH, Comm, comm_available, Mesh, RtError, Scene, camera_basis, camera_z, camera_z_device, default_walls, device_count, lib, move_light,  # noqa: F401
                  params_profile, shard_rows, shard_row_count, selftest_division, selftest_division3, fma_peak_tflops, selftest_libm, selftest_xorwow, write_png, PngWriter)
