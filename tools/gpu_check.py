"""Quick GPU sanity run: parity of both kernels against the oracle on small scenes + first timings.
Usage (on the GPU box): python tools/gpu_check.py [--full]   -> writes gpurun_out/gpu_check.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import zig_vulkan_b200 as zv  # noqa: E402
from zig_vulkan_b200 import ffi, scenes  # noqa: E402
from oracle import orc  # noqa: E402

POSE0 = dict(origin=(0.0, -10.0, 28.0), euler_deg=(25.0, 0.0, 0.0))


def compare(name, ctx_flags, grid, mats, cam, sun, ref_img, ref_aov, out):
    w, h = cam.image_width, cam.image_height
    ctx = ffi.Context(w, h, len(grid.brick_indices), brick_dim=grid.brick_dim, n_brick_alloc=grid.brick_alloc, flags=ctx_flags)
    ctx.upload_grid(grid, mats)
    ctx.trace(cam, sun)
    img = ctx.read_framebuffer()
    res = {"rgba_max_abs_diff": int(np.abs(img.astype(int) - ref_img.astype(int)).max()), "rgba_mismatch_px": int((img != ref_img).any(axis=2).sum())}
    if ctx_flags & ffi.VRT_FLAG_AOV:
        aov = ctx.read_aov()
        for f in ("flags", "grid_index", "voxel_index", "material", "shadow_grid_index", "shadow_voxel_index"):
            res["aov_" + f + "_mismatch"] = int((aov[f] != ref_aov[f]).sum())
        for f in ("t", "point", "normal"):
            res["aov_" + f + "_bitdiff"] = int((aov[f].view(np.uint32) != ref_aov[f].view(np.uint32)).sum())
        res["counters"] = ctx.counters()
    res["trace_ms"] = ctx.last_trace_ms()
    ctx.close()
    out[name] = res
    print(name, res, flush=True)


def main():
    full = "--full" in sys.argv
    out = {}
    mats = zv.terrain_materials()
    cases = [("C1", 64, 256, 256, False, 4), ("C3small", 128, 480, 270, True, 4), ("bd8", 128, 320, 200, True, 8), ("bd16", 256, 320, 200, True, 16)]
    for name, n, w, h, sun_on, bd in cases:
        grid = scenes.build_grid(n, brick_dim=bd)
        sc = orc.OracleScene.from_grid(grid, mats)
        cam = scenes.camera(w, h, **POSE0)
        sun = scenes.sun(sun_on)
        ref_img, ref_aov, cnt = sc.render(cam, sun, aov=True)
        out[name + "_oracle_counters"] = cnt
        compare(name + "_baseline_aov", ffi.VRT_FLAG_AOV | ffi.VRT_FLAG_BASELINE, grid, mats, cam, sun, ref_img, ref_aov, out)
        compare(name + "_tuned_aov", ffi.VRT_FLAG_AOV, grid, mats, cam, sun, ref_img, ref_aov, out)
        compare(name + "_tuned", 0, grid, mats, cam, sun, ref_img, ref_aov, out)
        compare(name + "_baseline", ffi.VRT_FLAG_BASELINE, grid, mats, cam, sun, ref_img, ref_aov, out)
    # bounce / spp / sun-disc mode (RNG reaches the image; both sides use the same deterministic sine)
    grid = scenes.build_grid(128)
    sc = orc.OracleScene.from_grid(grid, mats)
    cam = scenes.camera(480, 270, spp=2, max_bounce=2, **POSE0)
    sun = scenes.sun(True, radius=5.0)
    ref_img, ref_aov, cnt = sc.render(cam, sun, aov=True)
    compare("look_baseline_aov", ffi.VRT_FLAG_AOV | ffi.VRT_FLAG_BASELINE, grid, mats, cam, sun, ref_img, ref_aov, out)
    compare("look_tuned", 0, grid, mats, cam, sun, ref_img, ref_aov, out)

    if full:
        for wl in ("C2", "C3"):
            W = scenes.WORKLOADS[wl]
            t0 = time.time()
            grid = scenes.build_grid(W.n_voxels, W.brick_dim)
            print(wl, "grid built in", time.time() - t0, "s", flush=True)
            cam = scenes.camera(W.width, W.height, **POSE0)
            sun = scenes.sun(W.sun)
            for label, flags in (("tuned", 0), ("baseline", ffi.VRT_FLAG_BASELINE)):
                ctx = ffi.Context(W.width, W.height, len(grid.brick_indices), brick_dim=W.brick_dim, flags=flags)
                ctx.upload_grid(grid, mats)
                ms = []
                for _ in range(12):
                    ctx.trace(cam, sun)
                    ms.append(ctx.last_trace_ms())
                out[f"{wl}_{label}_ms"] = ms
                print(wl, label, ms, flush=True)
                ctx.close()
            ctx = ffi.Context(W.width, W.height, len(grid.brick_indices), brick_dim=W.brick_dim, flags=ffi.VRT_FLAG_AOV | ffi.VRT_FLAG_BASELINE)
            ctx.upload_grid(grid, mats)
            ctx.trace(cam, sun)
            out[f"{wl}_counters"] = ctx.counters()
            print(wl, "counters", out[f"{wl}_counters"], flush=True)
            ctx.close()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
