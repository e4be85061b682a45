"""How much a compacted sun-ray pass could save (analysis build build/ab/libvrt_stats.so, -DVRT_TILE_STATS=1): the same C3 frame
with and without the sun ray; per tile the difference is the sun trip.  python tools/gpu_sunsplit.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import zig_vulkan_b200 as zv
from zig_vulkan_b200 import ffi, scenes
POSE0 = dict(origin=(0.0, -10.0, 28.0), euler_deg=(25.0, 0.0, 0.0))
wl = scenes.WORKLOADS["C3"]
grid = scenes.build_grid(wl.n_voxels, wl.brick_dim)
mats = zv.terrain_materials()
cam = scenes.camera(wl.width, wl.height, **POSE0)
W, H = wl.width, wl.height


def stats(sun_on):
    ctx = ffi.Context(W, H, len(grid.brick_indices))
    ctx.upload_grid(grid, mats)
    ctx.set_schedule(ffi.VRT_SCHED_LPT, 2)
    n = ctx.n_tiles
    st = np.zeros((n, 8), dtype=np.uint32)
    ctx._check(ctx._l.vrt_debug_tile_stats(ctx.handle, st.ctypes.data, n))  # arm
    for _ in range(6):
        ctx.trace(cam, scenes.sun(sun_on))
    ctx._check(ctx._l.vrt_debug_tile_stats(ctx.handle, st.ctypes.data, n))
    ctx.close()
    return st.astype(np.float64)


def primary_hits_per_tile():
    ctx = ffi.Context(W, H, len(grid.brick_indices), flags=ffi.VRT_FLAG_AOV)
    ctx.upload_grid(grid, mats)
    ctx.trace(cam, scenes.sun(False))
    hit = (ctx.read_aov()["flags"] & 1).astype(np.int64)
    ctx.close()
    return hit.reshape(H // 4, 4, W // 8, 8).sum(axis=(1, 3)).reshape(-1).astype(np.float64)


off, on = stats(False), stats(True)
n = len(on)
n_sun = primary_hits_per_tile()          # primary rays of the tile that hit a voxel = lanes of its sun trip
assert len(n_sun) == n
ticks_on, ticks_off = on[:, 7] * 32, off[:, 7] * 32
sun_ticks = np.maximum(ticks_on - ticks_off, 0.0)
sun_rounds, sun_lanes = on[:, 0] - off[:, 0], on[:, 4] - off[:, 4]
print("tiles %d; frame ticks with sun %.3e, without %.3e -> sun trip %.1f %% of the frame" % (n, ticks_on.sum(), ticks_off.sum(), 100 * sun_ticks.sum() / ticks_on.sum()))
print("primary trip: %.0f rounds, %.1f lanes marching per round; sun trip: %.0f rounds, %.1f lanes marching per round" %
      (off[:, 0].sum(), off[:, 4].sum() / off[:, 0].sum(), sun_rounds.sum(), sun_lanes.sum() / max(sun_rounds.sum(), 1)))
none, full, part = n_sun == 0, n_sun == 32, (n_sun > 0) & (n_sun < 32)
print("tiles by sun lanes: none %d (%.1f %%), all 32: %d (%.1f %%), partial: %d (%.1f %%, mean %.1f lanes)" %
      (none.sum(), 100 * none.mean(), full.sum(), 100 * full.mean(), part.sum(), 100 * part.mean(), n_sun[part].mean() if part.any() else 0))
print("sun-trip ticks: full tiles %.1f %% of the frame, partial tiles %.1f %% of the frame" % (100 * sun_ticks[full].sum() / ticks_on.sum(), 100 * sun_ticks[part].sum() / ticks_on.sum()))
# bound for a perfect re-packing of the partial tiles' sun lanes into full warps: a packed warp costs what a full tile's sun trip
# costs per lane-slot; the saving cannot exceed the idle-lane share of the partial tiles' sun trips
saving = (sun_ticks[part] * (1.0 - n_sun[part] / 32.0)).sum()
print("upper bound of what packing the partial tiles' sun rays into full warps can save: %.2f %% of the frame (before the cost of the re-pack)" % (100 * saving / ticks_on.sum()))
# lanes that end early inside FULL sun trips (rays of different length) cannot be re-packed by a pass boundary at all
print("idle lane-rounds inside full tiles' sun trips: %.1f %% of their lane-rounds" % (100 * (1 - sun_lanes[full].sum() / max(32 * sun_rounds[full].sum(), 1))))
