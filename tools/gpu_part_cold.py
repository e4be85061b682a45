"""Partition simulation with a cold L2 (144 MiB fill before every frame, as bench.py does): per-rank trace kernel time of the first
ranks of an N-GPU run, cost-sorted order.  python tools/gpu_part_cold.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import zig_vulkan_b200 as zv
from zig_vulkan_b200 import ffi, scenes
wl = scenes.WORKLOADS["C3"]
grid = scenes.build_grid(wl.n_voxels, wl.brick_dim)
mats = zv.terrain_materials()
cam = scenes.camera(wl.width, wl.height, origin=(0.0, -10.0, 28.0), euler_deg=(25.0, 0.0, 0.0))
sun = scenes.sun(wl.sun)
flush = torch.empty(144 << 20, dtype=torch.uint8, device="cuda")
for world in (8, 4, 2):
    res = []
    for r in range(min(world, 3)):
        ctx = ffi.Context(wl.width, wl.height, len(grid.brick_indices), part=(r, world))
        ctx.upload_grid(grid, mats)
        ctx.set_schedule(ffi.VRT_SCHED_LPT, 2)
        ms = []
        for i in range(16):
            flush.fill_(i)
            torch.cuda.synchronize()
            ctx.trace(cam, sun)
            ms.append(ctx.last_trace_kernel_ms())
        ms = sorted(ms[4:])
        res.append((ms[0], ms[len(ms) // 2]))
        ctx.close()
    print(f"world {world} cold: per-rank kernel ms (min, median) " + "  ".join(f"{a:.4f}/{b:.4f}" for a, b in res), flush=True)
