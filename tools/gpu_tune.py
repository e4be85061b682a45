"""Time the tuned kernel on a workload under different VRT_TUNE_* env settings (experiments only)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import zig_vulkan_b200 as zv
from zig_vulkan_b200 import ffi, scenes
POSE0 = dict(origin=(0.0, -10.0, 28.0), euler_deg=(25.0, 0.0, 0.0))
wl = scenes.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C3"]
grid = scenes.build_grid(wl.n_voxels, wl.brick_dim)
mats = zv.terrain_materials()
cam = scenes.camera(wl.width, wl.height, **POSE0)
sun = scenes.sun(wl.sun)
ctx = ffi.Context(wl.width, wl.height, len(grid.brick_indices), brick_dim=wl.brick_dim)
ctx.upload_grid(grid, mats)
ref = None
for setting in sys.argv[2:]:
    for kv in setting.split(","):
        if kv:
            k, v = kv.split("=")
            os.environ[k] = v
    ms = []
    for _ in range(15):
        ctx.trace(cam, sun)
        ms.append(ctx.last_trace_ms())
    img = ctx.read_framebuffer()
    if ref is None:
        ref = img
    print(setting, "min %.4f median %.4f" % (min(ms[2:]), sorted(ms[2:])[len(ms[2:]) // 2]), "same_image" if np.array_equal(img, ref) else "IMAGE DIFFERS", flush=True)
