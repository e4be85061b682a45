"""Trace one C3 frame and run the present pass a few times (target of `ncu -k regex:denoise`); prints the device time."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zig_vulkan_b200 as zv  # noqa: E402
from zig_vulkan_b200 import ffi, scenes  # noqa: E402

wl = scenes.WORKLOADS["C3"]
grid = scenes.build_grid(wl.n_voxels, wl.brick_dim)
cam = scenes.camera(wl.width, wl.height, origin=(0.0, -10.0, 28.0), euler_deg=(25.0, 0.0, 0.0))
ctx = ffi.Context(wl.width, wl.height, len(grid.brick_indices))
ctx.upload_grid(grid, zv.terrain_materials())
ctx.trace(cam, scenes.sun(True))
for i in range(6):
    ctx.denoise()
    print("denoise ms", ctx.last_denoise_ms(), flush=True)
