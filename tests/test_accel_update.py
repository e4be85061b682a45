"""Scene edits between frames (VoxelRT.updateGridDelta, VoxelRT.zig:107-172): the derived distance planes follow the uploads by
a device-side merge of the status words + an in-place patch for new bricks.  The patched planes must be byte-identical to a
rebuild from scratch, and the frame identical to the oracle's on the edited grid."""
import numpy as np
import pytest

import zig_vulkan_b200 as zv
from zig_vulkan_b200 import ffi, scenes
from oracle import orc

pytestmark = pytest.mark.gpu
POSE0 = dict(origin=(0.0, -10.0, 28.0), euler_deg=(25.0, 0.0, 0.0))


def planes_after_full_rebuild(ctx):
    ctx.debug_force_accel_rebuild()
    return ctx.debug_dist_planes()


@pytest.mark.parametrize("bd,n", [(4, 64), (8, 128)])
def test_patched_distance_planes_equal_a_rebuild(materials, bd, n):
    grid = scenes.build_grid(n, brick_dim=bd)
    W, H = 160, 90
    cam, sun = scenes.camera(W, H, **POSE0), scenes.sun(True)
    ctx = ffi.Context(W, H, len(grid.brick_indices), brick_dim=bd)
    ctx.upload_grid(grid, materials)
    for which in range(5):
        grid.delta_reset(which)
    ctx.trace(cam, sun)
    rng = np.random.default_rng(3)
    base = ctx.debug_dist_planes().copy()
    assert np.array_equal(base, planes_after_full_rebuild(ctx))
    per_axis = n // bd
    for round_ in range(6):
        # 1 .. 5 voxels per round: floating in empty bricks (new bricks), on existing terrain (no new brick), in a grid corner
        k = int(rng.integers(1, 6))
        for _ in range(k):
            kind = int(rng.integers(0, 3))
            if kind == 0:
                x, y, z = (int(v) for v in rng.integers(0, n, 3))
            elif kind == 1:
                x, z = int(rng.integers(0, n)), int(rng.integers(0, n))
                y = int(rng.integers(0, n // 8))  # low: inside the terrain / ocean layer
            else:
                x, y, z = (int(v) for v in rng.choice([0, n - 1], 3))
            assert grid.insert(x, y, z, int(rng.integers(1, 8))) == 0
        assert ctx.upload_grid_delta(grid) > 0
        img = ctx.trace_to_host(cam, sun)
        patched = ctx.debug_dist_planes().copy()
        ref_img, _, _ = orc.OracleScene.from_grid(grid, materials).render(cam, sun)
        assert np.array_equal(img, ref_img), f"round {round_}: {(img != ref_img).any(axis=2).sum()} pixels differ"
        rebuilt = planes_after_full_rebuild(ctx)
        assert np.array_equal(patched, rebuilt), f"round {round_}: {(patched != rebuilt).sum()} distance bytes differ"
    assert not np.array_equal(base, patched)  # the edits did change the planes
    # uploading the same status words again changes nothing
    ctx.upload_brick_statuses(0, grid.statuses)
    assert np.array_equal(ctx.debug_dist_planes(), rebuilt)
    # a cleared bit (the reference never does this; a caller resetting its grid may) forces the full rebuild
    st = grid.statuses.copy()
    word = int(np.flatnonzero(st)[0])
    st[word] &= st[word] - 1
    ctx.upload_brick_statuses(0, st)
    after_clear = ctx.debug_dist_planes().copy()
    assert np.array_equal(after_clear, planes_after_full_rebuild(ctx)) and not np.array_equal(after_clear, rebuilt)
    # many new bricks at once (more than the patch list holds): full rebuild path
    ctx.upload_brick_statuses(0, grid.statuses)
    for _ in range(80):
        x, y, z = (int(v) for v in rng.integers(0, n, 3))
        grid.insert(x, y, z, 3)
    ctx.upload_grid_delta(grid)
    img = ctx.trace_to_host(cam, sun)
    ref_img, _, _ = orc.OracleScene.from_grid(grid, materials).render(cam, sun)
    assert np.array_equal(img, ref_img)
    assert np.array_equal(ctx.debug_dist_planes(), planes_after_full_rebuild(ctx))
    ctx.close()


def test_uploads_from_pinned_memory_may_be_overwritten_at_once(materials):
    """include/vrt.h promises every upload copies before returning.  With a pinned source cudaMemcpyAsync is truly asynchronous, so the
    uploads are staged through the context's own pinned ring: scribbling over the caller's buffer right after the call must not matter."""
    import torch

    grid = scenes.build_grid(64)
    cam, sun = scenes.camera(160, 90, **POSE0), scenes.sun(True)
    ref_img, _, _ = orc.OracleScene.from_grid(grid, materials).render(cam, sun)
    ctx = ffi.Context(160, 90, len(grid.brick_indices))
    ctx.upload_grid_state(grid.state)
    ctx.upload_materials(0, materials)
    for fn, arr in [(ctx._l.vrt_upload_brick_statuses, grid.statuses), (ctx._l.vrt_upload_brick_indices, grid.brick_indices), (ctx._l.vrt_upload_brick_occupancy, grid.occupancy),
                    (ctx._l.vrt_upload_brick_start_indices, grid.start_indices), (ctx._l.vrt_upload_material_indices, grid.material_indices)]:
        t = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).copy()).pin_memory()
        ctx._check(fn(ctx.handle, 0, t.data_ptr(), arr.shape[0]))
        t.fill_(0xA5)  # the caller reuses its buffer immediately
    assert np.array_equal(ctx.trace_to_host(cam, sun), ref_img)
    ctx.close()
