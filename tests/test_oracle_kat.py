"""Known-answer tests of the CPU oracle on hand-computable micro-scenes.

The reference has no test for this path, so these are authored from the shader text
(assets/shaders/brick_raytracer.comp); each expected value is derived in the comment next to it.
Micro-scene: 4x4x4 bricks of 4^3 voxels, min corner (0,0,0), brick scale 1 -> world [0,4]^3, voxel edge 0.25.
"""
import numpy as np
import pytest

import zig_vulkan_b200 as zv
from zig_vulkan_b200 import ffi, scenes
from oracle import orc

F = np.float32


def micro_grid(dim=4, voxels=()):
    g = ffi.Grid((dim, dim, dim), brick_dim=4, min_point=(0.0, 0.0, 0.0), scale=1.0)
    vy = dim * 4
    for (x, fy, z, m) in voxels:  # given in DEVICE coordinates; insert() flips y (Grid.zig:135)
        assert g.insert(x, vy - 1 - fy, z, m) == 0
    return g


@pytest.fixture(scope="module")
def one_voxel(materials):
    g = micro_grid(4, [(8, 8, 8, 5)])  # device voxel (8,8,8): world box [2,2.25]^3, brick (2,2,2), local voxel (0,0,0)
    return g, orc.OracleScene.from_grid(g, materials)


def test_axis_ray_plus_z_hand_computed(one_voxel):
    """Ray (2.125, 2.125, -1) -> +z.
    slab (:522-536): t_mins = (-2.125e12, -2.125e12, 1) -> axis z, normal.z = sign(inv.z) = +1, grid_t_min = 1, grid_t_max = 5.
    start t = 1 + 1e-4 (:287); fposition = (2.125, 2.125, 1e-4) -> cell (2,2,0); side_dist.z = (1 - 1e-4)*1, x/y = 0.5e12.
    cells visited: (2,2,0)=34, (2,2,1)=38, (2,2,2)=42 (grid_index = x + 4*(z + 4*y), :318) -> 3 grid steps, all in status word 1
    -> exactly 1 status fetch (:321-326).  At cell 42: t_value = 1.9999, hit.t = (1.9999 + 1) + 0.01 = 3.0099 (:332).
    BrickHit: fposition = ((2.125,2.125,2.0099) - (2,2,2)) / 0.25 = (0.5, 0.5, 0.0396) -> voxel (0,0,0), index 0, solid at once:
    t_offset = 0.25*0.05 = 0.0125; hit.t = 3.0099 - 0.0125 = 2.9974 (:431-432); normal = last grid step = (0,0,-1) (:370);
    point = origin + t*dir + normal*0.0125 = (2.125, 2.125, 1.9974 - 0.0125) (:433)."""
    _, sc = one_voxel
    hit, a = sc.grid_hit((2.125, 2.125, -1.0), (0.0, 0.0, 1.0))
    assert hit
    assert a["flags"] & 1
    assert a["grid_index"] == 42 and a["voxel_index"] == 0 and a["material"] == 5
    assert a["grid_steps"] == 3 and a["voxel_steps"] == 1 and a["status_fetches"] == 1
    assert tuple(a["normal"]) == (0.0, 0.0, -1.0)
    assert a["t"] == pytest.approx(2.9974, abs=2e-6)
    assert a["point"][0] == F(2.125) and a["point"][1] == F(2.125)
    assert a["point"][2] == pytest.approx(1.9849, abs=2e-6)
    # the same number re-derived in float32 with the shader's operation order
    t_value = F(F(1.0) - F(F(1.0) + F(1e-4)) + F(1.0)) * F(1.0)  # side_dist.z at start: fma(1, floor(f)-f, 1) with f = fma(1.0001,1,-1)
    t_value = F(t_value + F(1.0))                                # read before the 2nd increment (:367-368)
    t_expected = F(F(F(t_value + F(1.0)) + F(F(0.01) * F(1.0))) + F(F(0.0) - F(F(0.25) * F(0.05))))
    assert a["t"] == t_expected


@pytest.mark.parametrize(
    "origin,direction,face_t",
    [
        ((2.125, 2.125, -1.0), (0, 0, 1), 3.0),   # hits face z = 2
        ((2.125, 2.125, 5.0), (0, 0, -1), 2.75),  # hits face z = 2.25
        ((-1.0, 2.125, 2.125), (1, 0, 0), 3.0),
        ((5.0, 2.125, 2.125), (-1, 0, 0), 2.75),
        ((2.125, -1.0, 2.125), (0, 1, 0), 3.0),
        ((2.125, 5.0, 2.125), (0, -1, 0), 2.75),
    ],
)
def test_single_voxel_from_six_sides(one_voxel, origin, direction, face_t):
    """Axis-aligned rays with two zero direction components (safeInverse -> 1e12, :267; those axes never win the
    strict-< ladder, :345-372).  hit.t sits within the shader's fudge offsets of the geometric face distance:
    -1e-4*scale (:287) + 0.01*scale (:332) - 0.05*voxel_scale (:431) plus, when the brick is entered from its far side,
    the voxel DDA's own accumulated steps — all < 0.02.  Normal = -direction (:304-308)."""
    _, sc = one_voxel
    hit, a = sc.grid_hit(origin, direction)
    assert hit
    assert a["grid_index"] == 42 and a["voxel_index"] == 0 and a["material"] == 5
    assert tuple(a["normal"]) == tuple(-float(d) for d in direction)
    assert abs(float(a["t"]) - face_t) < 0.02
    p = np.array(origin, dtype=np.float64) + face_t * np.array(direction, dtype=np.float64)
    assert np.abs(a["point"].astype(np.float64) - p).max() < 0.03


def test_miss_outside_and_through_empty(one_voxel):
    _, sc = one_voxel
    hit, a = sc.grid_hit((10.0, 10.0, -1.0), (0, 0, 1))  # passes beside the grid: slab test fails (:282)
    assert not hit and a["grid_steps"] == 0 and a["status_fetches"] == 0
    hit, a = sc.grid_hit((2.125, 2.125, -1.0), (0, 0, -1))  # pointing away: t_max < t_min
    assert not hit and a["grid_steps"] == 0
    hit, a = sc.grid_hit((0.5, 0.5, -1.0), (0, 0, 1))  # through 4 empty cells and out
    assert not hit and a["grid_steps"] == 4 and a["voxel_steps"] == 0
    assert a["grid_index"] == 0xFFFFFFFF and a["voxel_index"] == 0xFFFFFFFF


def test_origin_inside_grid(one_voxel):
    """Origin inside the box: every t_lower/t_upper pair straddles 0 so t_mins < 0 and grid_t_min = max(1e-5, .) = 1e-5
    (:280,533).  From (2.125,2.125,0.5): cells z=0,1,2 -> 3 steps; face at distance 1.5."""
    _, sc = one_voxel
    hit, a = sc.grid_hit((2.125, 2.125, 0.5), (0, 0, 1))
    assert hit and a["grid_index"] == 42 and a["grid_steps"] == 3
    assert abs(float(a["t"]) - 1.5) < 0.02


def test_status_word_cache_by_direction(materials):
    """8x8x8 bricks, empty.  grid_index = x + 8*(z + 8*y); one status word = 32 consecutive indices (:321).
    +x at (y=2,z=2): indices 144..151 -> word 4 only                 -> 1 fetch
    +z at (x=2,y=2): indices 130,138,...,186 -> words 4,4,4,4,5,5,5,5 -> 2 fetches
    +y at (x=2,z=2): indices 18,82,...,466  -> 8 different words      -> 8 fetches"""
    g = micro_grid(8)
    sc = orc.OracleScene.from_grid(g, materials)
    for origin, d, fetches in [((-1, 2.5, 2.5), (1, 0, 0), 1), ((2.5, 2.5, -1), (0, 0, 1), 2), ((2.5, -1, 2.5), (0, 1, 0), 8)]:
        hit, a = sc.grid_hit(origin, d)
        assert not hit and a["grid_steps"] == 8 and a["status_fetches"] == fetches


def test_diagonal_ray_tie_break_and_brick_crossing(materials):
    """Direction (1,1,1)/sqrt3 from (-1,-1,-1): all three side distances stay equal, so the ladder's tie order applies
    (:345-372: x<y false -> y<z false -> z first; then x<y false, y<z true -> y; then x).  The ray enters cell (0,0,0),
    steps z, y, x to reach (1,1,1), and so on: 3 steps per diagonal cell -> cells visited = 1 + 3*3 = 10 before leaving
    the 4^3 grid.  A voxel on the diagonal at device (5,5,5) [brick (1,1,1), local (1,1,1) -> index 1+4*(1+4*1) = 21] is hit
    after the brick is entered through its corner."""
    g = micro_grid(4, [(5, 5, 5, 2)])
    sc = orc.OracleScene.from_grid(g, materials)
    hit, a = sc.grid_hit((-1.0, -1.0, -1.0), (1, 1, 1))
    assert hit and a["material"] == 2
    assert a["grid_index"] == 1 + 4 * (1 + 4 * 1) and a["voxel_index"] == 21
    assert a["grid_steps"] == 4  # (0,0,0) -> z -> y -> x -> (1,1,1)
    # geometric distance to the voxel's corner (1.25,1.25,1.25) from (-1,-1,-1) is 2.25*sqrt(3)
    assert abs(float(a["t"]) - 2.25 * np.sqrt(3.0)) < 0.03
    empty = orc.OracleScene.from_grid(micro_grid(4), materials)
    hit, a = empty.grid_hit((-1.0, -1.0, -1.0), (1, 1, 1))
    assert not hit and a["grid_steps"] == 10


def test_full_brick_and_empty_loaded_brick(materials):
    """A completely solid brick is hit in its first voxel; the voxel index is the entry voxel of the ray."""
    vox = [(x, y, z, 3) for x in range(4, 8) for y in range(4, 8) for z in range(4, 8)]  # brick (1,1,1)
    g = micro_grid(4, vox)
    sc = orc.OracleScene.from_grid(g, materials)
    assert g.active_bricks == 1 and int(g.occupancy[:8].view(np.uint64)[0]) == 0xFFFFFFFFFFFFFFFF
    hit, a = sc.grid_hit((1.6, 1.4, -1.0), (0, 0, 1))  # local voxel (2,1,0): 2 + 4*(0 + 4*1) = 18
    assert hit and a["voxel_index"] == 18 and a["voxel_steps"] == 1 and a["material"] == 3


def test_y_flip_on_insert(materials):
    """insert(x, y=0, z) lands in the LAST device row (flipped_y = voxel_dim_y-1-y, Grid.zig:135): world y near max."""
    g = ffi.Grid((4, 4, 4), min_point=(0.0, 0.0, 0.0), scale=1.0)
    assert g.insert(0, 0, 0, 1) == 0
    sc = orc.OracleScene.from_grid(g, materials)
    hit, a = sc.grid_hit((0.125, 3.875, -1.0), (0, 0, 1))
    assert hit and a["grid_index"] == 0 + 4 * (0 + 4 * 3) and a["voxel_index"] == 0 + 4 * (0 + 4 * 3)
    hit, _ = sc.grid_hit((0.125, 0.125, -1.0), (0, 0, 1))
    assert not hit


def test_sin_and_hash():
    l = orc.lib()
    xs = np.concatenate([np.linspace(-50, 50, 4001), np.linspace(-3000, 3000, 4001)]).astype(np.float32)
    got = np.array([l.orc_sinf(float(x)) for x in xs], dtype=np.float64)
    # GLSL requires 2^-11 absolute error for sin; the deterministic sine is far inside that
    assert np.abs(got - np.sin(xs.astype(np.float64))).max() < 2e-6
    # hash12 (rand.comp:22-26) re-derived in float32 with the documented dot order
    def hash12(px, py):
        fr = lambda v: F(v - np.floor(v))
        p3 = [fr(F(px) * F(0.1031)), fr(F(py) * F(0.1031)), fr(F(px) * F(0.1031))]
        q = [F(p3[1] + F(33.33)), F(p3[2] + F(33.33)), F(p3[0] + F(33.33))]
        d = F(F(F(p3[0] * q[0]) + F(p3[1] * q[1])) + F(p3[2] * q[2]))
        p3 = [F(v + d) for v in p3]
        return fr(F(F(p3[0] + p3[1]) * p3[2]))
    for px, py in [(0.0, 0.0), (0.2, 0.4), (12.4, 7.8), (383.8, 215.8)]:
        assert l.orc_hash12(px, py) == hash12(px, py)
    assert l.orc_hash12(0.0, 0.0) == 0.0  # sample 0 has zero jitter (:167-170)


def _unorm8(c):
    c = np.clip(c, F(0), F(1))
    return (c * F(255) + F(0.5)).astype(np.uint8)


def test_sky_pixel_colour(materials):
    """A frame over an empty grid is pure background (:197-201,260-264): t = 0.5*(d.y+1); c = fma(1-t, 1, t*(0.5,0.7,1));
    c *= sun_color when the sun is enabled; c/(c+1); sqrt; unorm8 — re-derived here in float32."""
    g = micro_grid(4)
    sc = orc.OracleScene.from_grid(g, materials)
    cam = scenes.camera(16, 8, origin=(2.0, 2.0, 10.0))
    for sun_on in (False, True):
        sun = scenes.sun(sun_on)
        img, _, cnt = sc.render(cam, sun)
        assert cnt["rays"] == 128 and cnt["hits"] == 0
        H, V, L, O = (np.array(list(v), dtype=F) for v in (cam.horizontal, cam.vertical, cam.lower_left_corner, cam.origin))
        for (px, py) in [(0, 0), (7, 3), (15, 7)]:
            u, v = F(F(px) / F(15)), F(F(py) / F(7))
            d = np.array([F(np.float64(H[i]) * u + L[i]) + F(np.float64(v) * V[i] - O[i]) for i in range(3)], dtype=F)  # two fmas, one add (:475)
            inv = F(1) / np.sqrt(F(F(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]))
            d = d * inv
            t = F(0.5) * F(d[1] + F(1))
            c = np.array([F(np.float64(F(1) - t) * 1.0 + np.float64(F(t * k))) for k in (F(0.5), F(0.7), F(1.0))], dtype=F)
            if sun_on:
                c = c * np.array([1.0, 1.1, 1.0], dtype=F)
            c = c / (c + F(1))
            c = np.sqrt(c / F(1))
            assert tuple(img[py, px]) == (*_unorm8(c), 255)


def test_lit_and_shadowed_pixels(materials):
    """Ground slab (material 1, grass: albedo (0,0.6,0)) with a floating voxel above it, camera looking straight down
    (-y is up, the sun sits at (0,-1000,0), Sun.zig:41).  A lit ground pixel is albedo*sun_color -> c/(c+1) -> sqrt;
    a pixel in the voxel's shadow is (0,0,0); primary hits == shadow rays."""
    vox = [(x, 12, z, 1) for x in range(16) for z in range(16)]  # device y = 12: world y in [3, 3.25] (down)
    vox += [(x, 4, z, 7) for x in range(6, 10) for z in range(6, 10)]  # 1x1 world-unit plate above the ground at y in [1,1.25]
    g = micro_grid(4, vox)
    sc = orc.OracleScene.from_grid(g, materials)
    cam = ffi.HostCamera(75.0, 33, 33, origin=(2.0, -3.0, 2.0), samples_per_pixel=1, max_bounce=0)
    cam.set_euler_deg(89.0, 0.0, 0.0)  # pitch about x: the view direction (-forward) now points (almost) along +y = down
    sun = ffi.HostSun(enabled=True, radius=0.0, animate=False)
    sundev = sun.device
    sundev.position[0], sundev.position[1], sundev.position[2] = -498.0, -1000.0, 2.0  # oblique: dx/dy = 0.5, so the shadow shifts to +x
    img, aov, cnt = sc.render(cam.device, sundev, aov=True)
    assert cnt["primary_hits"] > 100 and cnt["shadow_rays"] == cnt["primary_hits"]
    centre = aov[16, 16]
    assert centre["material"] == 7 and (centre["flags"] & 4) == 0  # the plate itself is lit
    lit = np.array([0.0, 0.6 * 1.1, 0.0], dtype=F)
    lit = np.sqrt(lit / (lit + F(1)))
    expect_lit = (*_unorm8(lit), 255)
    ground = aov["material"] == 1
    blocked = (aov["flags"] & 4) != 0
    assert (ground & blocked).sum() > 0 and (ground & ~blocked).sum() > 0
    assert all(tuple(p) == expect_lit for p in img[ground & ~blocked])
    assert all(tuple(p) == (0, 0, 0, 255) for p in img[ground & blocked])
    # the plate (x,z in [1.5,2.5], y in [1,1.25]) shadows the ground (y = 3) shifted by 0.5*(3-1.25) .. 0.5*(3-1) in +x
    pts = aov["point"][ground & blocked]
    assert pts[:, 0].min() > 1.5 + 0.875 - 0.06 and pts[:, 0].max() < 2.5 + 1.0 + 0.06 and pts[:, 2].min() > 1.45 and pts[:, 2].max() < 2.55


def test_max_bounce_one_means_no_rng_in_image(materials):
    """Device max_bounce = 1 and sun radius 0: no Rand() output reaches the image, so spp = 1 frames are a pure
    function of traversal.  Two different 'seeds' do not exist; instead check idempotence and that radius 0 gives a
    jitter-free sun ray (blocked flags identical to a second render)."""
    g = scenes.build_grid(64)
    sc = orc.OracleScene.from_grid(g, materials)
    cam = scenes.camera(64, 36, origin=(0.0, -10.0, 28.0), euler_deg=(25.0, 0.0, 0.0))
    a, aov_a, _ = sc.render(cam, scenes.sun(True), aov=True)
    b, aov_b, _ = sc.render(cam, scenes.sun(True), aov=True, threads=1)
    assert np.array_equal(a, b) and np.array_equal(aov_a, aov_b)


def test_degenerate_direction_is_a_miss(one_voxel):
    """Deviation from the shader (DESIGN.md "Deviations"): normalize((0,0,0)) is NaN, ray_step is 0 on all axes and the
    shader's DDA would never advance; oracle and CUDA kernels define such rays as misses."""
    _, sc = one_voxel
    hit, a = sc.grid_hit((2.125, 2.125, 2.125), (0.0, 0.0, 0.0))
    assert not hit and a["grid_steps"] == 0
    img, _, cnt = sc.render(scenes.camera(1, 1), scenes.sun(True))  # u = 0/0 (:168)
    assert cnt["hits"] == 0 and tuple(img[0, 0]) == (0, 0, 0, 255)


def test_insert_sequence_semantics_hand_computed():
    """Grid.insert (brick/Grid.zig:129-194) by hand: bricks are numbered in order of first insertion (fetchAdd :147), each new
    brick takes the next 64-entry material block (MaterialAllocator.zig:34-43), a voxel inserted twice keeps the last material
    (:173-175), occupancy bits accumulate (:180-182).  These are the semantics vrt_insert_voxels must reproduce on the device."""
    g = orc.OracleGrid((2, 2, 2), brick_dim=4)  # 8^3 voxels
    # y is flipped (:135): voxel y = 7 is brick row 0.  Insert into brick cell (1,0,0) first, then (0,0,0), then (1,0,0) again.
    assert g.insert(4, 7, 0, 11) == 0   # cell 1 -> brick 0, voxel 0
    assert g.insert(0, 7, 0, 22) == 0   # cell 0 -> brick 1, voxel 0
    assert g.insert(5, 7, 0, 33) == 0   # cell 1 (brick 0), voxel 1
    assert g.insert(4, 7, 0, 44) == 0   # same voxel as the first insert: overwrites 11
    assert g.insert(0, 0, 4, 55) == 0   # y = 0 -> flipped 7 -> brick row 1, z = 4 -> brick z 1: cell 0 + 2*(1 + 2*1) = 6 -> brick 2
    assert g.active_bricks == 3
    assert list(g.brick_indices) == [1, 0, 0, 0, 0, 0, 2, 0]
    assert g.statuses[0] == 0b01000011
    assert list(g.start_indices[:3]) == [0, 64, 128] and (g.start_indices[3:] == 0xFFFFFFFF).all()
    assert g.occupancy[0] == 0b00000011 and g.occupancy[8] == 0b00000001   # brick 0: voxels 0 and 1; brick 1: voxel 0
    # brick 2: x = 0, z = 4 % 4 = 0, flipped y = 7 % 4 = 3 -> voxel 0 + 4*(0 + 4*3) = 48 -> byte 6, bit 0
    assert g.occupancy[16 + 6] == 0b00000001
    mi = g.material_indices
    assert mi[0] == 44 and mi[1] == 33 and mi[64] == 22 and mi[128 + 48] == 55
    assert g.insert(8, 0, 0, 1) == -1  # outside the grid (:130-132)
