"""What makes a tile expensive?  Per-tile cost (clock ticks, from the LPT schedule) against per-tile traversal statistics of the
counting kernel (cells marched, voxel steps, rays).  python tools/gpu_tilecost.py [C3]"""
import json, os, sys
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
for _ in range(8):
    ctx.trace(cam, sun)
cost = ctx.sched_costs().astype(np.float64) * 32
ctx.close()
a = ffi.Context(W, H, len(grid.brick_indices), flags=ffi.VRT_FLAG_AOV)
a.upload_grid(grid, mats)
a.trace(cam, sun)
aov = a.read_aov()
a.close()
tx, ty = W // 8, H // 4
def tiles(x, f):
    return f(f(x.reshape(ty, 4, tx, 8), axis=3), axis=1).reshape(-1)
gs, vs = aov["grid_steps"].astype(np.float64), aov["voxel_steps"].astype(np.float64)
hit = (aov["flags"] & 1).astype(np.float64)
feat = {"max_grid_steps": tiles(gs, np.max), "sum_grid_steps": tiles(gs, np.sum), "max_voxel_steps": tiles(vs, np.max), "sum_voxel_steps": tiles(vs, np.sum), "hits": tiles(hit, np.sum)}
print("tiles", cost.shape[0], "cost ticks: median %.0f p90 %.0f p99 %.0f max %.0f; sum %.3g" % (np.median(cost), np.percentile(cost, 90), np.percentile(cost, 99), cost.max(), cost.sum()))
for k, v in feat.items():
    print("corr(cost, %s) = %.3f" % (k, np.corrcoef(cost, v)[0, 1]))
X = np.stack([feat["max_grid_steps"], feat["max_voxel_steps"], feat["sum_voxel_steps"], np.ones_like(cost)], axis=1)
coef, *_ = np.linalg.lstsq(X, cost, rcond=None)
print("least squares: cost ~ %.1f * max_grid_steps + %.1f * max_voxel_steps + %.2f * sum_voxel_steps + %.0f" % tuple(coef))
order = np.argsort(-cost)
print("heaviest tiles: cost, max_grid_steps, sum_grid_steps, max_voxel_steps, sum_voxel_steps, hits, tile(x,y)")
for t in order[:15]:
    print("  %7.0f %5.0f %7.0f %5.0f %7.0f %3.0f  (%d,%d)" % (cost[t], feat["max_grid_steps"][t], feat["sum_grid_steps"][t], feat["max_voxel_steps"][t], feat["sum_voxel_steps"][t], feat["hits"][t], t % tx, t // tx))
q = np.argsort(cost)
for name, sel in (("cheapest 10%", q[: len(q) // 10]), ("middle 10%", q[len(q) * 45 // 100: len(q) * 55 // 100]), ("top 1%", q[-len(q) // 100:])):
    print(name, {k: round(float(v[sel].mean()), 1) for k, v in feat.items()}, "cost %.0f" % cost[sel].mean())
top = q[-len(q) // 100:]
print("top 1%% of tiles hold %.1f %% of the cost; top 10%% hold %.1f %%" % (100 * cost[top].sum() / cost.sum(), 100 * cost[q[-len(q) // 10:]].sum() / cost.sum()))
