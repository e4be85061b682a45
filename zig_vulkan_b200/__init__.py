"""zig_vulkan_b200 — Python harness over the B200 voxel ray-tracing libraries.

The product is two native libraries built in-tree by ``make`` (see ``__graft_entry__.build``):

* ``libvrt.so``       sm_100a CUDA kernels + the C ABI of ``include/vrt.h`` (drop-in for the reference's
                      ``ComputePipeline`` / ``Pipeline.transfer*``, voxel_rt/ComputePipeline.zig, Pipeline.zig:560-652)
* ``libvrt_host.so``  C++ host side of ``include/vrt_host.h`` (BrickGrid, Camera, Sun, VoxelRT facade, scene producers)

This package only binds them with ctypes so tests and ``bench.py`` can drive the same entry points a Zig host
would.  There is no Python or CPU implementation of the trace path here: if the libraries are
missing, importing :mod:`zig_vulkan_b200.ffi` raises, and ``vrt_init`` fails when no CUDA device is present.
"""
from .ffi import (  # noqa: F401
    VrtError,
    Aov,
    CameraDevice,
    Config,
    Context,
    Counters,
    Grid,
    GridState,
    HostCamera,
    HostSun,
    Material,
    Renderer,
    SunDevice,
    Vox,
    bench_path_pose,
    lib,
    host_lib,
    terrain_materials,
)
from . import scenes  # noqa: F401
