"""Seeded random configurations of the path, shared by the oracle-vs-shader-text sweep (tests/test_ref_shader.py, CPU) and the
CUDA-vs-oracle sweep (tests/test_gpu_parity.py, GPU): non-cubic grids, non-power-of-two scales, random material tables, cameras
anywhere, 1-3 samples, 0-4 bounces, point / disc / no sun.  The pose is re-drawn until at least a fifth of the camera rays hit a voxel."""
import numpy as np

import zig_vulkan_b200 as zv
from zig_vulkan_b200 import ffi, scenes
from oracle import orc


def random_configuration(seed: int) -> dict:
    rng = np.random.default_rng(1000 + seed)
    bd = int(rng.choice([4, 4, 8]))
    dims = tuple(int(rng.integers(5, 20)) for _ in range(3))
    scale = float(rng.choice([0.5, 1.0, 0.7, 1.3, 2.0]))
    min_point = tuple(float(-0.5 * d * scale + rng.uniform(-1.0, 1.0)) for d in dims)
    grid = ffi.Grid(dims, brick_dim=bd, min_point=min_point, scale=scale)
    assert grid.fill_synthetic(int(rng.integers(1, 1 << 30))) == 0
    mats = zv.terrain_materials().copy()
    for i in range(1, 8):
        mats[i]["type"] = int(rng.integers(0, 5))  # lambertian, metal, dielectric, none, unknown
        # (a dielectric with refraction index 0 refracts into a NaN direction, on which the shader's DDA never advances — the
        # reference's text loops forever there, DESIGN.md "Deviations" 1 — so the sweep draws indices a material can have)
        mats[i]["type_data"] = float(rng.choice([1.0, 1.33, 1.5] if mats[i]["type"] == 2 else [0.0, 0.3, 1.0, 1.33, 1.5]))
        mats[i]["albedo_r"], mats[i]["albedo_g"], mats[i]["albedo_b"] = (float(x) for x in rng.uniform(0.05, 0.95, 3))
    extent = max(d * scale for d in dims)
    w, h, spp, bounce = int(rng.integers(24, 72)), int(rng.integers(16, 48)), int(rng.integers(1, 4)), int(rng.integers(0, 5))
    sun = scenes.sun(bool(rng.integers(0, 2)), float(rng.choice([0.0, 0.0, 2.0, 6.0])))
    sc = orc.OracleScene.from_grid(grid, mats)
    for _ in range(200):
        origin = tuple(float(x) for x in rng.uniform(-0.8 * extent, 0.8 * extent, 3))
        cam = scenes.camera(w, h, spp=spp, max_bounce=bounce, origin=origin, euler_deg=(float(rng.uniform(-80, 80)), float(rng.uniform(-180, 180)), 0.0))
        image, aov, counters = sc.render(cam, sun, aov=True)
        if counters["primary_hits"] * 5 >= w * h * spp:
            return dict(grid=grid, materials=mats, scene=sc, cam=cam, sun=sun, image=image, aov=aov, counters=counters)
    raise AssertionError(f"seed {seed}: no pose of the sweep sees the scene")
