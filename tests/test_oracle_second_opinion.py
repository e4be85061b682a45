"""A second, independent restatement of GridHit / BrickHit / AdvNormIntersect (assets/shaders/brick_raytracer.comp:267-471,
:493-536), written in Python straight from the GLSL — numpy float32 scalars for every operation, an exact rational fma — and
compared ray by ray with the C++ oracle (oracle/vrt_oracle.cpp).  The reference has no executable form of this path here
(the shader text itself, compiled under oracle/ref_shim/, is what pins the oracle: tests/test_ref_shader.py); this is a second reading of the shader in another
language with another arithmetic implementation, which is what catches slips of the C++ restatement."""
from fractions import Fraction

import numpy as np
import pytest

import zig_vulkan_b200 as zv
from zig_vulkan_b200 import scenes
from oracle import orc

F = np.float32
INF = F(np.inf)


def round_f32(x: Fraction) -> np.float32:
    """Correctly rounded (nearest, ties to even) float32 of an exact rational."""
    c = F(float(x))  # float(Fraction) is correctly rounded to double; the second rounding can be off by one float32 ulp at most
    best = None
    for cand in (np.nextafter(c, -INF), c, np.nextafter(c, INF)):
        if not np.isfinite(cand):
            continue
        err = abs(Fraction(float(cand)) - x)
        even = (int(np.float32(cand).view(np.uint32)) & 1) == 0
        if best is None or err < best[0] or (err == best[0] and even and not best[2]):
            best = (err, cand, even)
    return F(best[1])


def fma(a, b, c):  # GLSL fma(): one rounding
    if not (np.isfinite(a) and np.isfinite(b) and np.isfinite(c)):
        return F(a * b + c)
    return round_f32(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))


def gmin(a, b):
    return b if b < a else a


def gmax(a, b):
    return b if a < b else a


def sign(x):
    return F(1.0) if x > 0 else (F(-1.0) if x < 0 else F(0.0))


def safe_inverse(x):  # :267
    return F(1e12) if x == 0 else F(1.0) / x


class Scene:
    def __init__(self, grid, materials):
        s = grid.state
        self.g_min = [F(v) for v in list(s.min_point_base_t)[:3]]
        self.g_max = [F(v) for v in list(s.max_point_scale)[:3]]
        self.scale = F(s.max_point_scale[3])
        self.dim = (s.dim_x, s.dim_y, s.dim_z)
        self.bd = grid.brick_dim
        self.voxel_scale = self.scale * F(F(1.0) / F(self.bd))  # Pipeline.zig:313 brick_voxel_scale = 1 / brick_dim
        self.statuses, self.brick_indices = grid.statuses, grid.brick_indices
        self.occupancy, self.start_indices, self.material_indices = grid.occupancy, grid.start_indices, grid.material_indices
        self.materials = materials


def ray_at(o, d, t):  # :192-195
    return [fma(t, d[i], o[i]) for i in range(3)]


def side_dist_init(fpos, step, delta):  # :296-298 / :394-395
    out = []
    for i in range(3):
        fstep = F(step[i])
        inter = F(np.floor(fpos[i])) - fpos[i]
        out.append(fma(fstep, inter, fstep * F(0.5) + F(0.5)) * delta[i])
    return out


def dda_step(side, delta, pos, step, scale):  # :345-372 / :440-467; returns (t_value, axis)
    if side[0] < side[1]:
        axis = 0 if side[0] < side[2] else 2
    else:
        axis = 1 if side[1] < side[2] else 2
    t_value = side[axis] * scale
    side[axis] = side[axis] + delta[axis]
    pos[axis] += step[axis]
    return t_value, axis


def brick_hit(sc, o, d, grid_t_max, delta, step, brick_index, brick_min, hit_t, normal):  # :378-471 with t_max = grid_t_max (:339)
    p = ray_at(o, d, hit_t)
    fpos = [(p[i] - brick_min[i]) / sc.voxel_scale for i in range(3)]
    side = side_dist_init(fpos, step, delta)
    pos = [int(np.floor(fpos[i])) for i in range(3)]
    local_t_max = grid_t_max - hit_t
    t_value = F(0.0)
    bd = sc.bd
    brick_bytes = bd ** 3 // 8
    while all(0 <= pos[i] < bd for i in range(3)) and t_value <= local_t_max:
        voxel_index = pos[0] + bd * (pos[2] + bd * pos[1])
        mask_index = (voxel_index // 8) & 0xFF if bd <= 8 else voxel_index // 8  # uint8_t truncation (:413); widened for 16
        entry = int(sc.occupancy[brick_index * brick_bytes + mask_index])
        if (entry >> (voxel_index % 8)) & 1:
            start = int(sc.start_indices[brick_index]) & 0x7FFFFFFF
            index = int(sc.material_indices[start + voxel_index])
            m = sc.materials[index]
            ignore = int(m["type"]) == 3 and F(1.0) == F(m["type_data"])  # CreateRay: ignore MAT_NONE, ir 1.0 (:180-184)
            if not ignore:
                t_offset = sc.voxel_scale * F(0.05)
                return True, hit_t + (t_value - t_offset), voxel_index, index, normal
        t_value, axis = dda_step(side, delta, pos, step, sc.voxel_scale)
        normal = [F(0.0)] * 3
        normal[axis] = F(1.0) if step[axis] < 0 else F(-1.0)
    return False, hit_t, -1, -1, normal


def grid_hit(sc, origin, direction):  # CreateRay + GridHit(r, 0.00001, infinity)
    o = [F(v) for v in origin]
    dv = [F(v) for v in direction]
    with np.errstate(divide="ignore", invalid="ignore"):  # normalize(0) = 0 * inf = NaN, as in GLSL
        inv_len = F(1.0) / F(np.sqrt((dv[0] * dv[0] + dv[1] * dv[1]) + dv[2] * dv[2]))
        d = [dv[i] * inv_len for i in range(3)]
    if not np.isfinite((d[0] + d[1]) + d[2]):
        return False, None  # documented deviation: non-finite direction = miss
    inv = [safe_inverse(d[i]) for i in range(3)]
    # AdvNormIntersect (:522-536)
    t_lower = [(sc.g_min[i] - o[i]) * inv[i] for i in range(3)]
    t_upper = [(sc.g_max[i] - o[i]) * inv[i] for i in range(3)]
    t_mins = [gmin(t_lower[i], t_upper[i]) for i in range(3)]
    t_maxes = [gmax(t_lower[i], t_upper[i]) for i in range(3)]
    k = int(t_mins[1] > t_mins[0] and t_mins[1] > t_mins[2]) + int(t_mins[2] > t_mins[0] and t_mins[2] > t_mins[1]) * 2
    normal = [F(0.0)] * 3
    normal[k] = sign(inv[k])
    grid_t_min = gmax(F(0.00001), t_mins[k])
    grid_t_max = gmin(INF, gmin(gmin(t_maxes[0], t_maxes[1]), t_maxes[2]))
    if not grid_t_min <= grid_t_max:
        return False, None
    global_t = grid_t_min + F(0.0001) * sc.scale
    delta = [F(abs(inv[i])) for i in range(3)]
    step = [int(sign(d[i])) for i in range(3)]
    p = ray_at(o, d, global_t)
    fpos = [(p[i] - sc.g_min[i]) / sc.scale for i in range(3)]
    side = side_dist_init(fpos, step, delta)
    pos = [int(np.floor(fpos[i])) for i in range(3)]
    t_value = F(0.0)
    while all(0 <= pos[i] < sc.dim[i] for i in range(3)):  # global_t_value <= t_max with t_max = +inf is always true
        grid_index = pos[0] + sc.dim[0] * (pos[2] + sc.dim[2] * pos[1])
        if (int(sc.statuses[grid_index // 32]) >> (grid_index % 32)) & 1:
            brick_min = [fma(F(pos[i]), sc.scale, sc.g_min[i]) for i in range(3)]
            hit_t = (t_value + grid_t_min) + F(0.01) * sc.scale
            ok, t, voxel_index, index, n = brick_hit(sc, o, d, grid_t_max, delta, step, int(sc.brick_indices[grid_index]), brick_min, hit_t, normal)
            if ok:
                return True, dict(grid_index=grid_index, voxel_index=voxel_index, material=index, t=t, normal=n)
        t_value, axis = dda_step(side, delta, pos, step, sc.scale)
        normal = [F(0.0)] * 3
        normal[axis] = F(1.0) if step[axis] < 0 else F(-1.0)
    return False, None


@pytest.mark.parametrize("brick_dim,n_voxels", [(4, 64), (8, 64), (16, 128)])
def test_second_reading_of_the_shader_agrees_with_the_oracle(materials, brick_dim, n_voxels):
    grid = scenes.build_grid(n_voxels, brick_dim=brick_dim)
    mine = Scene(grid, materials)
    theirs = orc.OracleScene.from_grid(grid, materials)
    rng = np.random.default_rng(100 + brick_dim)
    n = 160
    origins = rng.uniform(-45, 45, (n, 3)).astype(np.float32)
    origins[: n // 3] = rng.uniform(-30, 30, (n // 3, 3)).astype(np.float32)  # inside the grid
    directions = rng.normal(size=(n, 3)).astype(np.float32)
    directions[n // 2:] = (-origins[n // 2:] + rng.uniform(-12, 12, (n - n // 2, 3))).astype(np.float32)  # aimed at the terrain
    directions[0] = (0, 1, 0)    # straight down (+y is down in device space): two zero components (safeInverse)
    origins[0] = (0.3, -40.0, 0.7)
    directions[1] = (0, 0, 0)    # zero direction: defined miss
    hits = 0
    for i in range(n):
        got, a = theirs.grid_hit(origins[i], directions[i])
        ok, m = grid_hit(mine, origins[i], directions[i])
        assert ok == got, f"ray {i}: hit flag {ok} vs oracle {got}"
        if ok:
            hits += 1
            assert (m["grid_index"], m["voxel_index"], m["material"]) == (int(a["grid_index"]), int(a["voxel_index"]), int(a["material"])), f"ray {i}"
            assert F(m["t"]).view(np.uint32) == a["t"].view(np.uint32), f"ray {i}: t {m['t']} vs {a['t']}"
            assert [float(v) for v in m["normal"]] == [float(v) for v in a["normal"]], f"ray {i}"
    assert 40 < hits < n - 20
