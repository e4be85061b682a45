"""Host side (libvrt_host.so): BrickGrid / Camera / Sun / scene producers against the reference's Zig sources."""
import ctypes as C
import math

import numpy as np
import pytest

import zig_vulkan_b200 as zv
from zig_vulkan_b200 import ffi, scenes
from oracle import orc


def test_grid_init_matches_grid_zig():
    # Grid.zig:36-114 with dims (3,2,5), scale 0.5, min (-1,-2,-3)
    g = ffi.Grid((3, 2, 5), min_point=(-1.0, -2.0, -3.0), scale=0.5)
    s = g.state
    assert (s.voxel_dim_x, s.voxel_dim_y, s.voxel_dim_z) == (12, 8, 20) and (s.dim_x, s.dim_y, s.dim_z) == (3, 2, 5)
    assert list(s.min_point_base_t) == [-1.0, -2.0, -3.0, np.float32(0.01)]
    assert list(s.max_point_scale) == [0.5, -1.0, -0.5, 0.5]
    assert len(g.statuses) == 1 and len(g.brick_indices) == 30                      # ceil(30/32), brick_count
    assert len(g.occupancy) == 30 * 8 and len(g.start_indices) == 30 and len(g.material_indices) == 30 * 64
    assert (g.start_indices == 0xFFFFFFFF).all() and not g.occupancy.any() and not g.statuses.any()
    assert g.brick_alloc == 30 and g.active_bricks == 0
    g2 = ffi.Grid((3, 2, 5), brick_alloc=7)
    assert len(g2.occupancy) == 56 and len(g2.start_indices) == 7 and len(g2.material_indices) == 7 * 64
    with pytest.raises(ffi.VrtError):
        ffi.Grid((0, 1, 1))
    with pytest.raises(ffi.VrtError):
        ffi.Grid((1, 1, 1), brick_dim=5)


def test_insert_index_math_and_y_flip():
    g = ffi.Grid((4, 4, 4))
    # voxel (5, 2, 9): flipped_y = 15-2 = 13 -> brick (1,3,2) -> grid index 1 + 4*(2 + 4*3) = 57; local (1,1,1) -> bit 21
    assert g.insert(5, 2, 9, 6) == 0
    assert g.active_bricks == 1
    assert g.statuses[57 // 32] == 1 << (57 % 32)
    assert g.brick_indices[57] == 0 and g.start_indices[0] == 0
    assert g.occupancy[21 // 8] == 1 << (21 % 8) and g.material_indices[21] == 6
    # second voxel in another brick gets brick 1 and the next 64-entry material slab (MaterialAllocator.zig:34-43)
    assert g.insert(0, 15, 0, 2) == 0  # flipped_y 0 -> brick (0,0,0) index 0, local bit 0
    assert g.active_bricks == 2 and g.brick_indices[0] == 1 and g.start_indices[1] == 64
    assert g.occupancy[8] == 1 and g.material_indices[64] == 2
    # re-inserting into a loaded brick reuses it and overwrites the material
    assert g.insert(5, 2, 9, 3) == 0 and g.active_bricks == 2 and g.material_indices[21] == 3
    # the reference asserts on out-of-range coordinates (Grid.zig:130-132)
    assert g.insert(16, 0, 0, 1) == -1 and g.insert(0, 16, 0, 1) == -1 and g.insert(0, 0, 16, 1) == -1


def test_insert_capacity():
    g = ffi.Grid((2, 2, 2), brick_alloc=2)
    assert g.insert(0, 0, 0, 1) == 0 and g.insert(4, 0, 0, 1) == 0
    assert g.insert(0, 4, 0, 1) == -2  # third brick does not fit brick_alloc = 2
    assert g.active_bricks == 2


def test_delta_tracking():
    """DeviceDataDelta (State.zig:14-57): `.empty` starts at from = to = 0, so the first range starts at element 0;
    after resetDelta the range is exactly [min, max+1)."""
    g = ffi.Grid((4, 4, 4))
    for which in range(5):
        assert g.delta(which)[0] == 0
    g.insert(5, 2, 9, 6)  # grid index 57, brick 0, bit 21
    assert g.delta(ffi_const("VRT_DELTA_STATUSES")) == (1, 0, 2)          # word 1 touched; from stays 0
    assert g.delta(ffi_const("VRT_DELTA_BRICK_INDICES")) == (1, 0, 58)
    assert g.delta(ffi_const("VRT_DELTA_OCCUPANCY")) == (1, 0, 3)
    assert g.delta(ffi_const("VRT_DELTA_START_INDICES")) == (1, 0, 1)
    assert g.delta(ffi_const("VRT_DELTA_MATERIAL_INDICES")) == (1, 0, 22)
    for which in range(5):
        g.delta_reset(which)
        assert g.delta(which)[0] == 0
    g.insert(5, 2, 9, 1)
    g.insert(6, 2, 9, 1)  # bit 22 of the same brick
    assert g.delta(ffi_const("VRT_DELTA_STATUSES")) == (1, 1, 2)
    assert g.delta(ffi_const("VRT_DELTA_BRICK_INDICES")) == (1, 57, 58)
    assert g.delta(ffi_const("VRT_DELTA_OCCUPANCY")) == (1, 2, 3)
    assert g.delta(ffi_const("VRT_DELTA_START_INDICES"))[0] == 0  # start index was already set: no delta
    assert g.delta(ffi_const("VRT_DELTA_MATERIAL_INDICES")) == (1, 21, 23)
    assert g.delta(7)[0] == -1


def ffi_const(name):
    return {"VRT_DELTA_STATUSES": 0, "VRT_DELTA_BRICK_INDICES": 1, "VRT_DELTA_OCCUPANCY": 2, "VRT_DELTA_START_INDICES": 3, "VRT_DELTA_MATERIAL_INDICES": 4}[name]


@pytest.mark.parametrize("brick_dim", [4, 8, 16])
def test_grid_builder_agrees_with_oracle_builder(brick_dim):
    """Two independent restatements of Grid.insert (product: csrc/host/vrt_grid.cpp, oracle: vrt_oracle.cpp) on random voxels."""
    rng = np.random.default_rng(brick_dim)
    dim = (5, 3, 4)
    n = 4000
    xyzm = np.stack([rng.integers(0, dim[0] * brick_dim, n), rng.integers(0, dim[1] * brick_dim, n), rng.integers(0, dim[2] * brick_dim, n),
                     rng.integers(0, 8, n)], axis=1).astype(np.uint32)
    a = ffi.Grid(dim, brick_dim=brick_dim, min_point=(-3.0, 1.0, 2.0), scale=0.75)
    b = orc.OracleGrid(dim, brick_dim=brick_dim, min_point=(-3.0, 1.0, 2.0), scale=0.75)
    assert a.insert_many(xyzm) == 0 and b.insert_many(xyzm) == 0
    assert a.active_bricks == b.active_bricks
    assert bytes(a.state) == bytes(b.state)
    for name in ("statuses", "brick_indices", "occupancy", "start_indices", "material_indices"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name


def test_synthetic_scene_is_deterministic_and_sane():
    g1 = scenes.build_grid(64)
    g2 = scenes.build_grid(64)
    for name in ("statuses", "brick_indices", "occupancy", "start_indices", "material_indices"):
        assert np.array_equal(getattr(g1, name), getattr(g2, name))
    assert 0 < g1.active_bricks < 16 ** 3
    used = g1.material_indices[: g1.active_bricks * 64]
    occ = np.unpackbits(g1.occupancy[: g1.active_bricks * 8], bitorder="little").astype(bool)
    mats = np.unique(used[occ])
    assert set(mats) <= set(range(8)) and 7 in mats and len(mats) >= 4
    assert scenes.build_grid(64, seed=1).active_bricks != 0
    assert not np.array_equal(scenes.build_grid(64, seed=1).occupancy, g1.occupancy)
    assert scenes.count_bricks(64, 4) == g1.active_bricks


def test_terrain_materials_table():
    m = zv.terrain_materials()
    assert len(m) == 256
    assert m[0]["type"] == 2 and m[0]["type_data"] == np.float32(1.333)      # water (terrain.zig:131-138)
    assert m[7]["type"] == 1 and m[7]["type_data"] == np.float32(0.45)       # iron (terrain.zig:188-195)
    assert all(m[i]["type"] == 0 for i in range(1, 7))
    assert m[2]["albedo_g"] == np.float32(0.5019)


def test_camera_init_matches_camera_zig():
    # Camera.init (Camera.zig:36-77): fov 75, 1920x1080, viewport_height config 2
    cam = ffi.HostCamera(75.0, 1920, 1080, origin=(1.0, 2.0, 3.0), samples_per_pixel=3, max_bounce=2)
    d = cam.device
    vh = 2.0 * math.tan(math.radians(75.0) * 0.5)
    vw = vh * 1920 / 1080
    assert (d.image_width, d.image_height, d.samples_per_pixel) == (1920, 1080, 3)
    assert d.max_bounce == 3  # stored +1 (Camera.zig:74)
    assert list(d.horizontal) == pytest.approx([vw, 0, 0], abs=1e-5)   # right = up x forward = (1,0,0)
    assert list(d.vertical) == pytest.approx([0, vh, 0], abs=1e-5)     # up = forward x right = (0,1,0)
    assert list(d.lower_left_corner) == pytest.approx([1 - vw / 2, 2 - vh / 2, 3 - 1], abs=1e-5)  # origin - h/2 - v/2 - forward
    assert list(d.origin) == [1.0, 2.0, 3.0]


def test_camera_motion():
    cam = ffi.HostCamera(75.0, 64, 64)
    cam.translate(0.5, (0, 0, 2))  # norm(by) * dt * speed(1) along z
    assert list(cam.device.origin) == pytest.approx([0, 0, 0.5], abs=1e-6)
    cam.turn_yaw(math.pi / 2 / 0.1 / 2)  # h_angle = angle*turn_rate = pi/4 -> quaternion (cos, 0, sin, 0) = 90 degrees about y
    d = cam.device
    # forward (0,0,1) rotated 90 deg about +y -> (1,0,0); llc = origin - h/2 - v/2 - forward
    vh = 2.0 * math.tan(math.radians(75.0) * 0.5)
    assert list(d.horizontal) == pytest.approx([0, 0, -vh], abs=1e-5)
    llc_expected = np.array([0, 0, 0.5]) - np.array(list(d.horizontal)) / 2 - np.array(list(d.vertical)) / 2 - np.array([1, 0, 0])
    assert list(d.lower_left_corner) == pytest.approx(list(llc_expected), abs=1e-5)
    ffi.host_lib().vrt_hcam_disable_input(cam.handle)
    cam.translate(1.0, (1, 0, 0))
    assert list(cam.device.origin) == pytest.approx([0, 0, 0.5], abs=1e-6)  # input disabled -> no-op (Camera.zig:114)
    cam.reset()
    assert list(cam.device.horizontal) == pytest.approx([vh, 0, 0], abs=1e-5)
    # pitch is clamped so the camera never flips (Camera.zig:136-139)
    for _ in range(100):
        cam.turn_pitch(1.0)
    assert np.isfinite(list(cam.device.lower_left_corner)).all()


def test_sun():
    s = ffi.HostSun()
    d = s.device
    assert list(d.position) == [0.0, -1000.0, 0.0] and d.enabled == 1 and d.radius == 5.0  # Sun.zig:4-11,41
    assert list(d.color) == [1.0, np.float32(1.1), 1.0]
    s.update(1.0)  # slerp_pos 0 -> position = static vector rotated by orientation[0] = identity; colour = lerp_color[0]
    d = s.device
    assert list(d.position) == pytest.approx([0, -1000, 0], abs=1e-3)
    assert list(d.color) == pytest.approx([1, 0.99, 0.823], abs=1e-6)
    for _ in range(50):
        s.update(1.0)
    p = np.array(list(s.device.position))
    assert np.linalg.norm(p) == pytest.approx(1000.0, rel=1e-3)  # stays on the sphere of radius sun_distance
    still = ffi.HostSun(animate=False)
    still.update(10.0)
    assert list(still.device.position) == [0.0, -1000.0, 0.0]


def test_sun_cycle_takes_the_short_arc():
    """Sun.update slerps through three orientations 120 degrees apart about z (Sun.zig:36-40,72).  The third segment (240 degrees
    back to the identity) has a negative quaternion dot product: zalgebra's slerp negates one operand and keeps going the short
    way (another 120 degrees), so the sun circles the grid at a steady pace — 12 degrees per update at speed 0.1 and dt 1 —
    instead of swinging 240 degrees backwards."""
    s = ffi.HostSun()
    pos = []
    for _ in range(61):  # two full cycles of 3 segments x 10 updates
        s.update(1.0)
        pos.append(np.array(list(s.device.position), dtype=np.float64))
    steps = [np.degrees(np.arccos(np.clip(np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b)), -1, 1))) for a, b in zip(pos, pos[1:])]
    assert max(steps) < 16.0 and min(steps) > 8.0, (min(steps), max(steps))
    # the winding about z is monotonic: one full turn per 30 updates
    ang = np.unwrap([np.arctan2(p[0], -p[1]) for p in pos])
    assert np.all(np.diff(ang) > 0) or np.all(np.diff(ang) < 0)
    assert abs(abs(ang[30] - ang[0]) - 2 * np.pi) < 0.15
    # mid third segment: rotation by 300 degrees about z of (0,-1000,0)
    s2 = ffi.HostSun()
    for _ in range(26):
        s2.update(1.0)
    p = np.array(list(s2.device.position))
    assert p[0] < -700 and p[1] < -300 and abs(p[2]) < 120, p


def test_bench_path():
    o, q = zv.bench_path_pose(0.0)
    assert o == [0, 0, 0] and q == [1, 0, 0, 0]                      # Benchmark.zig:146,160
    o, q = zv.bench_path_pose(1.0)
    assert o == [0, 13, 0]                                           # last way point (:156)
    o, q = zv.bench_path_pose(1.5 / 11)                              # halfway between points 1 and 2
    assert o == pytest.approx([2.5, 5, 2.5], abs=1e-4)
    o2, _ = zv.bench_path_pose(1.5 / 11, 2.0)
    assert o2 == pytest.approx([5, 10, 5], abs=1e-4)
    assert len(scenes.sweep_poses(11)) == 11


def test_benchmark_follows_the_reference_update_rule():
    """Benchmark.init / update / Report (Benchmark.zig:22-101): fixed dt frames over the 60 s path."""
    cam = ffi.HostCamera(75.0, 64, 36, origin=(9.0, 9.0, 9.0), samples_per_pixel=1, max_bounce=0)
    grid = ffi.Grid((4, 2, 3))
    b = ffi.Benchmark(cam, grid, sun_enabled=True)
    assert list(cam.device.origin) == [0.0, 0.0, 0.0]              # init: origin = path_points[0] (:28)
    assert not b.update(60.0 / 11 * 1.5)                            # halfway between way points 1 and 2 (:50-56)
    assert list(cam.device.origin) == pytest.approx([2.5, 5.0, 2.5], abs=1e-4)
    o_ref, q_ref = zv.bench_path_pose(1.5 / 11)
    assert list(cam.device.origin) == pytest.approx(o_ref, abs=1e-5)
    llc = list(cam.device.lower_left_corner)
    cam.translate(1.0, (1.0, 0.0, 0.0))                             # input is disabled while benchmarking (:27)
    assert list(cam.device.lower_left_corner) == llc
    frames = 1
    while not b.update(0.5):
        frames += 1
        assert frames < 1000
    frames += 1
    rep = b.report
    assert rep.frames == frames
    assert rep.min_frame_ms == pytest.approx(500.0) and rep.max_frame_ms == pytest.approx(60.0 / 11 * 1.5 * 1000.0, rel=1e-5)
    assert rep.avg_frame_ms == pytest.approx((60.0 / 11 * 1.5 + 0.5 * (frames - 1)) / frames * 1000.0, rel=1e-4)
    assert list(rep.voxel_dim) == [16, 8, 12] and rep.sun_enabled == 1
    assert (rep.image_width, rep.image_height, rep.max_bounce, rep.samples_per_pixel) == (64, 36, 1, 1)
    # the last eleventh of the path keeps the pose the previous segment ended with (:51,:59): close to the last way point
    assert list(cam.device.origin) == pytest.approx([0.0, 13.0, 0.0], abs=2.2)  # within one 0.5 s frame of the end of the segment
    b.close()
    short = ffi.Benchmark(cam, None, sun_enabled=False, duration_s=1.0, extent_scale=2.0)
    assert short.update(0.25) is False and short.update(0.8) is True
    assert list(short.report.voxel_dim) == [0, 0, 0] and short.report.sun_enabled == 0
    with pytest.raises(ValueError):
        ffi.Benchmark(cam, None, duration_s=0.0)
