"""Where a tile's instructions go (analysis build build/ab/libvrt_stats.so, -DVRT_TILE_STATS=1): rounds, step iterations, brick
phases and voxel-loop iterations per tile, against its cost.  python tools/gpu_tilestats.py [C3]"""
import os, sys
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
W, H = wl.width, wl.height
ctx = ffi.Context(W, H, len(grid.brick_indices))
ctx.upload_grid(grid, mats)
ctx.set_schedule(ffi.VRT_SCHED_LPT, 2)
n = ctx.n_tiles
st = np.zeros((n, 8), dtype=np.uint32)
ctx._check(ctx._l.vrt_debug_tile_stats(ctx.handle, st.ctypes.data, n))  # arm
for _ in range(6):
    ctx.trace(cam, sun)
ctx._check(ctx._l.vrt_debug_tile_stats(ctx.handle, st.ctypes.data, n))
ctx.close()
names = ["rounds", "step_iters", "brick_phases", "voxel_iters", "lanes_marching_sum", "lanes_testing_sum", "lanes_parked_sum", "ticks/32"]
s = st.astype(np.float64)
cost = s[:, 7] * 32
tot = s.sum(axis=0)
print("frame totals:", {k: int(v) for k, v in zip(names[:7], tot[:7])})
print("per round: %.2f step iterations, %.1f lanes marching, %.1f lanes parked (waiting for the brick phase); per brick phase: %.1f lanes testing, %.2f voxel iterations" % (tot[1] / tot[0], tot[4] / tot[0], tot[6] / tot[0], tot[5] / tot[2], tot[3] / tot[2]))
# instruction model: rounds * a + step_iters * 10 + brick_phases * b + voxel_iters * 24
X = np.stack([s[:, 0], s[:, 1], s[:, 2], s[:, 3], np.ones(n)], axis=1)
coef, *_ = np.linalg.lstsq(X, cost, rcond=None)
print("least squares ticks ~ %.0f * rounds + %.0f * step_iters + %.0f * brick_phases + %.0f * voxel_iters + %.0f" % tuple(coef))
share = coef[:4] * tot[:4]
print("share of the modelled ticks: rounds %.0f %%, steps %.0f %%, brick phases %.0f %%, voxel iterations %.0f %%" % tuple(100 * share / share.sum()))
order = np.argsort(-cost)
print("heaviest tiles: ticks | rounds step_iters brick_phases voxel_iters | lanes/round lanes/phase parked/round")
for t in order[:12]:
    r = s[t]
    print("  %7.0f | %5.0f %6.0f %5.0f %6.0f | %5.1f %5.1f %5.1f" % (cost[t], r[0], r[1], r[2], r[3], r[4] / max(r[0], 1), r[5] / max(r[2], 1), r[6] / max(r[0], 1)))
q = np.argsort(cost)
for name, sel in (("middle 10%", q[n * 45 // 100: n * 55 // 100]), ("top 10%", q[-n // 10:]), ("top 1%", q[-n // 100:])):
    m = s[sel].mean(axis=0)
    print("%-10s ticks %7.0f | rounds %5.0f step_iters %6.0f brick_phases %5.1f voxel_iters %6.1f | lanes/round %.1f lanes/phase %.1f parked/round %.1f" % (name, cost[sel].mean(), m[0], m[1], m[2], m[3], m[4] / max(m[0], 1), m[5] / max(m[2], 1), m[6] / max(m[0], 1)))
