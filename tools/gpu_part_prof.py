"""What rank 0 of an 8-GPU run launches for a C3 frame (1/8 of the strips, cost-sorted order), a few times: target of
`ncu -k regex:trace_warp` for the scaling-limit analysis.  python tools/gpu_part_prof.py [world]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zig_vulkan_b200 as zv
from zig_vulkan_b200 import ffi, scenes
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
wl = scenes.WORKLOADS["C3"]
grid = scenes.build_grid(wl.n_voxels, wl.brick_dim)
cam = scenes.camera(wl.width, wl.height, origin=(0.0, -10.0, 28.0), euler_deg=(25.0, 0.0, 0.0))
ctx = ffi.Context(wl.width, wl.height, len(grid.brick_indices), part=(0, world))
ctx.upload_grid(grid, zv.terrain_materials())
ctx.set_schedule(ffi.VRT_SCHED_LPT, 2)
for i in range(10):
    ctx.trace(cam, scenes.sun(wl.sun))
    print("kernel ms", ctx.last_trace_kernel_ms(), flush=True)
