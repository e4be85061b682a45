"""Committed golden vectors (tests/golden/*.npz, produced by tests/golden/make_golden.py from the oracle).
CPU: the oracle still reproduces them.  GPU: both CUDA kernels reproduce them through the C ABI."""
import importlib.util
import os

import numpy as np
import pytest

import zig_vulkan_b200 as zv
from zig_vulkan_b200 import ffi
from oracle import orc

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(spec)
spec.loader.exec_module(make_golden)

EXACT_FIELDS = ("flags", "grid_index", "voxel_index", "material", "shadow_grid_index", "shadow_voxel_index")
FLOAT_FIELDS = ("t", "point", "normal")


def load(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    return z["rgba"], z["aov"], dict(zip(orc.COUNTER_NAMES, [int(v) for v in z["counters"]]))


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_oracle_reproduces_golden(name):
    rgba, aov, cnt = load(name)
    _, _, _, img, got_aov, got_cnt = make_golden.render_case(make_golden.CASES[name])
    assert np.array_equal(img, rgba)
    assert np.array_equal(got_aov, aov)
    assert got_cnt == cnt


@pytest.mark.gpu
@pytest.mark.parametrize("kernel_flags", [0, ffi.VRT_FLAG_BASELINE], ids=["tuned", "baseline"])
@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_cuda_reproduces_golden(name, kernel_flags):
    rgba, aov, cnt = load(name)
    case = make_golden.CASES[name]
    n, bd, w, h, sun_on, radius, spp, bounce, pose = case
    from zig_vulkan_b200 import scenes

    grid = scenes.build_grid(n, brick_dim=bd)
    cam = scenes.camera(w, h, spp=spp, max_bounce=bounce, **pose)
    sun = scenes.sun(sun_on, radius)
    mats = zv.terrain_materials()
    # plain context: the framebuffer
    ctx = ffi.Context(w, h, len(grid.brick_indices), brick_dim=bd, flags=kernel_flags)
    ctx.upload_grid(grid, mats)
    img = ctx.trace_to_host(cam, sun)
    assert ctx.last_trace_launches() >= 1
    ctx.close()
    assert np.array_equal(img, rgba), f"{(img != rgba).any(axis=2).sum()} pixels differ"
    # AOV context: hit records bit-exact (integers AND floats)
    ctx = ffi.Context(w, h, len(grid.brick_indices), brick_dim=bd, flags=kernel_flags | ffi.VRT_FLAG_AOV)
    ctx.upload_grid(grid, mats)
    ctx.trace(cam, sun)
    got = ctx.read_aov()
    assert np.array_equal(ctx.read_framebuffer(), rgba)
    for f in EXACT_FIELDS:
        assert np.array_equal(got[f], aov[f]), f
    for f in FLOAT_FIELDS:
        assert np.array_equal(got[f].view(np.uint32), aov[f].view(np.uint32)), f
    c = ctx.counters()
    for k in ("rays", "primary_hits", "shadow_rays", "grid_steps", "hits"):
        assert c[k] == cnt[k], k
    if kernel_flags & ffi.VRT_FLAG_BASELINE or bd != 4:
        assert c == cnt  # the reference-shape kernel also reproduces the request counters of the byte model
    ctx.close()
