"""The BASELINE.json workloads as concrete inputs, shared by tests, smoke() and bench.py.

All configs: spp 1, Camera.Config.max_bounce 0 (device value 1, Camera.zig:74), fov 75 (VoxelRT.zig:42), cubic grid of
world extent 64 with min corner (-32,-32,-32), the 8 terrain materials padded to 256, sun at (0,-1000,0) colour
(1,1.1,1) radius 0, seeded integer synthetic terrain (seed 420, the seed main.zig:120 passes).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import ffi

WORLD_EXTENT = 64.0
SEED = 420


@dataclass(frozen=True)
class Workload:
    name: str
    n_voxels: int      # voxels per axis
    brick_dim: int
    width: int
    height: int
    sun: bool          # cast one sun (shadow) ray per primary hit
    description: str

    @property
    def bricks_per_axis(self) -> int:
        return self.n_voxels // self.brick_dim

    @property
    def scale(self) -> float:
        return WORLD_EXTENT / self.bricks_per_axis


WORKLOADS = {
    "C1": Workload("C1", 64, 4, 256, 256, False, "64^3 voxel grid, 256x256 frame, primary rays only"),
    "C2": Workload("C2", 256, 4, 1920, 1080, False, "256^3 voxel grid, 1920x1080, primary rays only"),
    "C3": Workload("C3", 512, 4, 1920, 1080, True, "512^3 voxel grid, 1920x1080, primary + 1 shadow ray"),
    "C4": Workload("C4", 1024, 16, 1920, 1080, False, "1024^3 brickmap (64^3 bricks of 16^3), 1920x1080, primary rays"),
    "C5": Workload("C5", 512, 4, 3840, 2160, True, "512^3 voxel grid, 3840x2160, primary + 1 shadow ray, rows tiled across GPUs"),
}

# The reference application's own default configuration (src/main.zig:23,71,77-81,122-135, Sun.zig:4-11): 128 x 64 x 128 bricks of 4^3 at
# scale 0.5 with min corner (-32,-16,-32), internal resolution 1024x576, 2 samples per pixel, max_bounce 2 (device value 3), sun
# enabled with a disc of radius 5.  Every pixel takes the general shading path (jittered samples, scatter functions, sin-hash RNG).
REF_DEFAULT = dict(dim=(128, 64, 128), brick_dim=4, min_point=(-32.0, -16.0, -32.0), scale=0.5, width=1024, height=576, spp=2, max_bounce=2, sun_radius=5.0)


def build_ref_default_grid(seed: int = SEED) -> ffi.Grid:
    g = ffi.Grid(REF_DEFAULT["dim"], brick_dim=4, min_point=REF_DEFAULT["min_point"], scale=REF_DEFAULT["scale"])
    rc = g.fill_synthetic(seed)
    if rc != 0:
        raise ffi.VrtError(rc, "vrt_scene_synthetic_fill failed for the reference default grid")
    return g


def build_grid(n_voxels: int, brick_dim: int = 4, seed: int = SEED, brick_alloc: int = 0) -> ffi.Grid:
    """Cubic BrickGrid of world extent 64 filled with the synthetic terrain."""
    per_axis = n_voxels // brick_dim
    g = ffi.Grid((per_axis, per_axis, per_axis), brick_dim=brick_dim, brick_alloc=brick_alloc,
                 min_point=(-WORLD_EXTENT / 2,) * 3, scale=WORLD_EXTENT / per_axis)
    rc = g.fill_synthetic(seed)
    if rc != 0:
        raise ffi.VrtError(rc, f"vrt_scene_synthetic_fill failed for {n_voxels}^3 (brick_alloc={brick_alloc})")
    return g


def count_bricks(n_voxels: int, brick_dim: int, seed: int = SEED) -> int:
    """Bricks the synthetic scene occupies (lets large grids size brick_alloc tightly, BrickGrid.Config.brick_alloc)."""
    import ctypes as C

    seen = set()
    d = brick_dim
    per_axis = n_voxels // d

    @C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint8)
    def emit(_user, x, y, z, _m):
        seen.add((x // d) + per_axis * ((z // d) + per_axis * (y // d)))
        return 0

    ffi.host_lib().vrt_scene_synthetic(n_voxels, seed, C.cast(emit, C.c_void_p), None)
    return len(seen)


def camera(width: int, height: int, origin=(0.0, -8.0, 0.0), euler_deg=(0.0, 0.0, 0.0), spp: int = 1, max_bounce: int = 0) -> ffi.CameraDevice:
    cam = ffi.HostCamera(75.0, width, height, origin=origin, samples_per_pixel=spp, max_bounce=max_bounce)
    if tuple(euler_deg) != (0.0, 0.0, 0.0):
        cam.set_euler_deg(*euler_deg)
    return cam.device


def sun(enabled: bool, radius: float = 0.0) -> ffi.SunDevice:
    return ffi.HostSun(enabled=enabled, radius=radius, animate=False).device


def sweep_poses(n: int = 11, extent_scale: float = 1.0):
    """n poses sampled uniformly along the reference's benchmark fly-through (Benchmark.zig:141-173)."""
    return [ffi.bench_path_pose(i / max(n - 1, 1), extent_scale) for i in range(n)]


def camera_from_pose(width, height, origin, yaw_wxyz, spp=1, max_bounce=0) -> ffi.CameraDevice:
    cam = ffi.HostCamera(75.0, width, height, origin=origin, samples_per_pixel=spp, max_bounce=max_bounce)
    cam.set_orientation(yaw_wxyz)
    return cam.device
