"""vrt_insert_voxels: BrickGrid.insert (brick/Grid.zig:129-194) for a batch of voxels on the device.  The five grid buffers must
be byte-identical to the oracle's sequential restatement of Grid.insert (oracle/vrt_oracle.cpp orc_grid_insert) over the same
voxel list — brick numbering in order of first occurrence, material blocks, last-writer-wins materials — and a frame traced
from them identical to the oracle's frame."""
import ctypes as C

import numpy as np
import pytest

import zig_vulkan_b200 as zv
from zig_vulkan_b200 import ffi, scenes
from oracle import orc

pytestmark = pytest.mark.gpu

BUFFERS = {3: "statuses", 4: "brick_indices", 5: "occupancy", 6: "start_indices", 7: "material_indices"}


def make_ctx(dim, brick_dim, brick_alloc=0, w=64, h=36):
    og = orc.OracleGrid(dim, brick_dim=brick_dim, brick_alloc=brick_alloc, min_point=(-8.0, -4.0, -6.0), scale=1.0)
    ctx = ffi.Context(w, h, dim[0] * dim[1] * dim[2], brick_dim=brick_dim, n_brick_alloc=brick_alloc)
    state = ffi.GridState.from_buffer_copy(bytes(og.state))
    ctx.upload_grid_state(state)
    return og, ctx


def assert_buffers_equal(ctx, og):
    for which, name in BUFFERS.items():
        got, ref = ctx.download_buffer(which), getattr(og, name)
        assert got.shape == ref.shape, name
        assert np.array_equal(got, ref), f"{name}: {(got != ref).sum()} of {len(ref)} elements differ"


def random_voxels(rng, n, voxel_dims, clustered=True):
    if clustered:  # a few dense blobs: many voxels per brick, many duplicates
        centres = rng.integers(0, voxel_dims, (8, 3))
        xyz = (centres[rng.integers(0, 8, n)] + rng.integers(-6, 7, (n, 3))) % voxel_dims
    else:
        xyz = rng.integers(0, voxel_dims, (n, 3))
    return np.concatenate([xyz, rng.integers(0, 256, (n, 1))], axis=1).astype(np.uint32)


@pytest.mark.parametrize("brick_dim", [4, 8])
def test_batches_match_sequential_inserts(brick_dim):
    dim = (16, 8, 12)
    og, ctx = make_ctx(dim, brick_dim)
    vdims = np.array(dim) * brick_dim
    rng = np.random.default_rng(brick_dim)
    active = 0
    for n, clustered in [(5000, True), (1, True), (3000, False), (20000, True)]:
        xyzm = random_voxels(rng, n, vdims, clustered)
        xyzm[n // 2:n // 2 + n // 8] = xyzm[:n // 8]  # the same voxels again ...
        xyzm[n // 2:n // 2 + n // 8, 3] ^= 0x55       # ... with other materials: the later insert wins
        assert og.insert_many(xyzm) == 0
        active = ctx.insert_voxels(xyzm, active)
        assert active == og.active_bricks
        assert_buffers_equal(ctx, og)
    ctx.close()


def test_large_batch_uses_every_scan_level():
    dim = (64, 64, 64)
    og, ctx = make_ctx(dim, 4)
    rng = np.random.default_rng(3)
    xyzm = random_voxels(rng, (1 << 20) + 12345, np.array(dim) * 4, clustered=False)  # > 1024^2 voxels: three scan levels
    assert og.insert_many(xyzm) == 0
    assert ctx.insert_voxels(xyzm, 0) == og.active_bricks
    assert_buffers_equal(ctx, og)
    ctx.close()


def test_rejected_batches_change_nothing():
    dim = (4, 4, 4)
    og, ctx = make_ctx(dim, 4, brick_alloc=10)
    rng = np.random.default_rng(9)
    first = random_voxels(rng, 200, np.array([8, 8, 8]), clustered=False)  # at most 8 bricks
    assert og.insert_many(first) == 0
    active = ctx.insert_voxels(first, 0)
    assert active == og.active_bricks <= 8
    bad = first.copy()
    bad[100, 1] = 16  # y outside the 16^3 voxels
    with pytest.raises(ffi.VrtError) as e:
        ctx.insert_voxels(bad, active)
    assert e.value.code == ffi.VRT_E_RANGE
    assert_buffers_equal(ctx, og)
    everywhere = random_voxels(rng, 2000, np.array([16, 16, 16]), clustered=False)  # needs far more than 10 bricks
    with pytest.raises(ffi.VrtError) as e:
        ctx.insert_voxels(everywhere, active)
    assert e.value.code == ffi.VRT_E_RANGE
    assert_buffers_equal(ctx, og)
    with pytest.raises(ffi.VrtError):
        ctx.insert_voxels(first, 11)  # more active bricks than were ever allocated
    assert ctx.insert_voxels(np.zeros((0, 4), dtype=np.uint32), active) == active
    ctx.close()


def test_scene_built_on_the_device_traces_like_the_oracle(materials):
    """The 64^3 synthetic scene inserted on the device (instead of on the host + uploads) gives the oracle's frame."""
    voxels = []

    @C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint8)
    def emit(_user, x, y, z, m):
        voxels.append((x, y, z, m))
        return 0

    assert ffi.host_lib().vrt_scene_synthetic(64, scenes.SEED, C.cast(emit, C.c_void_p), None) == 0
    xyzm = np.array(voxels, dtype=np.uint32)
    host_grid = scenes.build_grid(64)  # the same scene through the host BrickGrid
    ctx = ffi.Context(160, 90, len(host_grid.brick_indices))
    ctx.upload_grid_state(host_grid.state)
    ctx.upload_materials(0, materials)
    assert ctx.insert_voxels(xyzm, 0) == host_grid.active_bricks
    for which, name in BUFFERS.items():
        assert np.array_equal(ctx.download_buffer(which), getattr(host_grid, name)), name
    cam = scenes.camera(160, 90, origin=(0.0, -10.0, 28.0), euler_deg=(25.0, 0.0, 0.0))
    sun = scenes.sun(True)
    ref, _, _ = orc.OracleScene.from_grid(host_grid, materials).render(cam, sun)
    assert np.array_equal(ctx.trace_to_host(cam, sun), ref)
    # an edit on the device: a pillar in front of the camera, same edit on the host grid for the oracle
    pillar = np.array([(x, y, z, 7) for y in range(20, 60) for x in range(30, 34) for z in range(50, 54)], dtype=np.uint32)
    active = ctx.insert_voxels(pillar, host_grid.active_bricks)
    for x, y, z, m in pillar:
        assert host_grid.insert(int(x), int(y), int(z), int(m)) == 0
    assert active == host_grid.active_bricks
    ref2, _, _ = orc.OracleScene.from_grid(host_grid, materials).render(cam, sun)
    img2 = ctx.trace_to_host(cam, sun)
    assert np.array_equal(img2, ref2) and not np.array_equal(img2, ref)
    ctx.close()
