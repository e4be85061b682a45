"""Single-GPU timing of partitioned launches: what each rank of an N-GPU run executes, without the exchange, under the three tile
schedules (static bottom-up, cost-sorted, cost-sorted and dealt).  python tools/gpu_part.py [C3|C5]"""
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
out = {}

def best(ctx, n=14, skip=4):
    ms = []
    for _ in range(n):
        ctx.trace(cam, sun)
        ms.append(ctx.last_trace_kernel_ms())
    return min(ms[skip:])

full = ffi.Context(wl.width, wl.height, len(grid.brick_indices))
full.upload_grid(grid, mats)
t_static = best(full)
full.set_schedule(ffi.VRT_SCHED_LPT, 2)
t_lpt = best(full)
costs = full.sched_costs()
print(f"world 1: static {t_static:.4f} ms, lpt {t_lpt:.4f} ms; tile cost (ticks/32): median {np.median(costs):.0f} p99 {np.percentile(costs, 99):.0f} max {costs.max()}", flush=True)
out["1"] = {"static": t_static, "lpt": t_lpt, "cost_median": float(np.median(costs)), "cost_p99": float(np.percentile(costs, 99)), "cost_max": int(costs.max())}
full.close()
for world in (2, 4, 8):
    res = {"static": [], "lpt": [], "deal": []}
    for r in range(world):
        ctx = ffi.Context(wl.width, wl.height, len(grid.brick_indices), part=(r, world))
        ctx.upload_grid(grid, mats)
        res["static"].append(best(ctx))
        ctx.set_schedule(ffi.VRT_SCHED_LPT, 2)
        res["lpt"].append(best(ctx))
        ctx.set_schedule(ffi.VRT_SCHED_DEAL, 1000)  # keep the order sorted from the full-frame costs (a real run exchanges costs every frame)
        ctx.sched_set_costs(costs)
        res["deal"].append(best(ctx))
        ctx.close()
    out[str(world)] = {k: {"max": max(v), "min": min(v)} for k, v in res.items()}
    print(f"world {world}: per-rank kernel ms  " + "  ".join(f"{k}: max {max(v):.4f} min {min(v):.4f}" for k, v in res.items()) + f"   (ideal {t_lpt / world:.4f})", flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"part_{wl.name}.json"), "w"), indent=1)
