"""Single-GPU timing of partitioned launches (what each rank of an N-GPU run executes, without the exchange)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zig_vulkan_b200 as zv
from zig_vulkan_b200 import ffi, scenes
POSE0 = dict(origin=(0.0, -10.0, 28.0), euler_deg=(25.0, 0.0, 0.0))
wl = scenes.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C3"]
grid = scenes.build_grid(wl.n_voxels, wl.brick_dim)
mats = zv.terrain_materials()
cam = scenes.camera(wl.width, wl.height, **POSE0)
sun = scenes.sun(wl.sun)
for world in (1, 2, 4, 8):
    for mode in ("interleave", "slab"):
        res = []
        for r in range(world):
            if mode == "interleave":
                ctx = ffi.Context(wl.width, wl.height, len(grid.brick_indices), part=(r, world))
            else:
                h = wl.height // world
                ctx = ffi.Context(wl.width, wl.height, len(grid.brick_indices), rows=(r * h, (r + 1) * h) if world > 1 else (0, 0))
            ctx.upload_grid(grid, mats)
            ms = []
            for _ in range(12):
                ctx.trace(cam, sun)
                ms.append(ctx.last_trace_ms())
            res.append(min(ms[2:]))
            ctx.close()
        print(f"world {world} {mode:10s} per-rank kernel ms: max {max(res):.4f} min {min(res):.4f}  all {[round(x,3) for x in res]}", flush=True)
