"""ctypes wrapper of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg / --impl reference) may import this module.
Parity is pinned to the reference's shader text through oracle/_ref (oracle/ref.py, tests/test_ref_shader.py); see oracle/vrt_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

AOV_DTYPE = np.dtype(
    [
        ("flags", "<u4"), ("grid_index", "<u4"), ("voxel_index", "<u4"), ("material", "<u4"),
        ("t", "<f4"), ("point", "<f4", 3), ("normal", "<f4", 3),
        ("shadow_grid_index", "<u4"), ("shadow_voxel_index", "<u4"),
        ("grid_steps", "<u4"), ("voxel_steps", "<u4"), ("status_fetches", "<u4"),
    ]
)
MATERIAL_DTYPE = np.dtype([("type", "<u4"), ("albedo_r", "<f4"), ("albedo_g", "<f4"), ("albedo_b", "<f4"), ("type_data", "<f4")])
COUNTER_NAMES = ("rays", "primary_hits", "shadow_rays", "grid_steps", "voxel_steps", "status_fetches", "bricks_entered", "hits")


class GridState(C.Structure):
    _fields_ = [
        ("voxel_dim_x", C.c_uint32), ("voxel_dim_y", C.c_uint32), ("voxel_dim_z", C.c_uint32),
        ("dim_x", C.c_uint32), ("dim_y", C.c_uint32), ("dim_z", C.c_uint32),
        ("padding1", C.c_uint32), ("padding2", C.c_uint32),
        ("min_point_base_t", C.c_float * 4), ("max_point_scale", C.c_float * 4),
    ]


class Scene(C.Structure):
    _fields_ = [
        ("state", GridState), ("brick_dim", C.c_uint32), ("n_materials", C.c_uint32), ("materials", C.c_void_p),
        ("statuses", C.c_void_p), ("n_statuses", C.c_uint64),
        ("brick_indices", C.c_void_p), ("n_brick_indices", C.c_uint64),
        ("occupancy", C.c_void_p), ("n_occupancy", C.c_uint64),
        ("start_indices", C.c_void_p), ("n_start_indices", C.c_uint64),
        ("material_indices", C.c_void_p), ("n_material_indices", C.c_uint64),
    ]


def build():
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        l = C.CDLL(path)
        P = C.c_void_p
        l.orc_grid_create.restype = P
        l.orc_grid_create.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(C.c_float * 3), C.c_float, C.c_float]
        l.orc_grid_destroy.argtypes = [P]
        l.orc_grid_insert.restype = C.c_int
        l.orc_grid_insert.argtypes = [P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint8]
        l.orc_grid_insert_many.restype = C.c_int
        l.orc_grid_insert_many.argtypes = [P, P, C.c_size_t]
        l.orc_grid_active_bricks.restype = C.c_uint32
        l.orc_grid_active_bricks.argtypes = [P]
        l.orc_grid_get_state.argtypes = [P, C.POINTER(GridState)]
        for n in ("statuses", "brick_indices", "occupancy", "start_indices", "material_indices"):
            f = getattr(l, "orc_grid_" + n)
            f.restype = P
            f.argtypes = [P, C.POINTER(C.c_uint64)]
        l.orc_scene_from_grid.argtypes = [P, P, C.c_uint32, C.POINTER(Scene)]
        l.orc_render.restype = C.c_int
        l.orc_render.argtypes = [C.POINTER(Scene), P, P, C.c_uint32, C.c_uint32, P, P, P, C.c_int]
        l.orc_grid_hit.restype = C.c_int
        l.orc_grid_hit.argtypes = [C.POINTER(Scene), C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3), P]
        l.orc_sinf.restype = C.c_float
        l.orc_sinf.argtypes = [C.c_float]
        l.orc_hash12.restype = C.c_float
        l.orc_hash12.argtypes = [C.c_float, C.c_float]
        l.orc_denoise.restype = C.c_int
        l.orc_denoise.argtypes = [P, C.c_uint32, C.c_uint32, P, C.c_uint32, C.c_uint32, C.c_uint32, P, C.c_int]
        l.orc_pow.restype = C.c_float
        l.orc_pow.argtypes = [C.c_float, C.c_float]
        _lib = l
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class OracleGrid:
    """Restatement of BrickGrid (brick/Grid.zig) inside the oracle; used to cross-check the product's grid builder."""

    def __init__(self, dim, brick_dim=4, brick_alloc=0, min_point=(0.0, 0.0, 0.0), scale=1.0, base_t=0.01):
        l = lib()
        mp = (C.c_float * 3)(*[float(v) for v in min_point])
        self.handle = l.orc_grid_create(dim[0], dim[1], dim[2], brick_dim, brick_alloc, C.byref(mp), scale, base_t)
        if not self.handle:
            raise RuntimeError("orc_grid_create failed")
        self.brick_dim = brick_dim

    def __del__(self):
        try:
            if self.handle:
                lib().orc_grid_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def insert(self, x, y, z, m) -> int:
        return lib().orc_grid_insert(self.handle, x, y, z, m)

    def insert_many(self, xyzm) -> int:
        a = np.ascontiguousarray(xyzm, dtype=np.uint32).reshape(-1, 4)
        return lib().orc_grid_insert_many(self.handle, _ptr(a), a.shape[0])

    @property
    def active_bricks(self) -> int:
        return lib().orc_grid_active_bricks(self.handle)

    @property
    def state(self) -> GridState:
        s = GridState()
        lib().orc_grid_get_state(self.handle, C.byref(s))
        return s

    def _array(self, name, dtype):
        n = C.c_uint64()
        p = getattr(lib(), "orc_grid_" + name)(self.handle, C.byref(n))
        buf = (C.c_uint8 * (n.value * np.dtype(dtype).itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=dtype)

    statuses = property(lambda s: s._array("statuses", np.uint32))
    brick_indices = property(lambda s: s._array("brick_indices", np.uint32))
    occupancy = property(lambda s: s._array("occupancy", np.uint8))
    start_indices = property(lambda s: s._array("start_indices", np.uint32))
    material_indices = property(lambda s: s._array("material_indices", np.uint8))


class OracleScene:
    """The UBO + six SSBOs the shader reads (brick_raytracer.comp:79-134), as numpy arrays kept alive here."""

    def __init__(self, state, brick_dim, materials, statuses, brick_indices, occupancy, start_indices, material_indices):
        self.arrays = dict(
            materials=np.ascontiguousarray(materials, dtype=MATERIAL_DTYPE),
            statuses=np.ascontiguousarray(statuses, dtype=np.uint32),
            brick_indices=np.ascontiguousarray(brick_indices, dtype=np.uint32),
            occupancy=np.ascontiguousarray(occupancy, dtype=np.uint8),
            start_indices=np.ascontiguousarray(start_indices, dtype=np.uint32),
            material_indices=np.ascontiguousarray(material_indices, dtype=np.uint8),
        )
        s = Scene()
        C.memmove(C.byref(s.state), C.byref(state), C.sizeof(GridState))
        s.brick_dim = brick_dim
        a = self.arrays
        s.materials, s.n_materials = a["materials"].ctypes.data, a["materials"].shape[0]
        s.statuses, s.n_statuses = a["statuses"].ctypes.data, a["statuses"].shape[0]
        s.brick_indices, s.n_brick_indices = a["brick_indices"].ctypes.data, a["brick_indices"].shape[0]
        s.occupancy, s.n_occupancy = a["occupancy"].ctypes.data, a["occupancy"].shape[0]
        s.start_indices, s.n_start_indices = a["start_indices"].ctypes.data, a["start_indices"].shape[0]
        s.material_indices, s.n_material_indices = a["material_indices"].ctypes.data, a["material_indices"].shape[0]
        self.c = s

    @classmethod
    def from_grid(cls, grid, materials):
        """`grid` is anything with .state/.brick_dim and the five arrays (OracleGrid or the product's Grid)."""
        return cls(grid.state, grid.brick_dim, materials, grid.statuses, grid.brick_indices, grid.occupancy, grid.start_indices, grid.material_indices)

    def render(self, camera, sun, rows=None, aov=False, threads=0):
        """brick_raytracer.comp main() over rows [begin,end).  Returns (rgba8[H,W,4], aov[H,W] or None, counters dict).
        `camera`/`sun` are any ctypes structs with the vrt_camera / vrt_sun layout."""
        w, h = camera.image_width, camera.image_height
        r0, r1 = rows if rows else (0, h)
        img = np.zeros((h, w, 4), dtype=np.uint8)
        aov_arr = np.zeros((h, w), dtype=AOV_DTYPE) if aov else None
        counters = (C.c_uint64 * 8)()
        rc = lib().orc_render(C.byref(self.c), C.byref(camera), C.byref(sun), r0, r1, _ptr(img), _ptr(aov_arr) if aov else None, counters, threads)
        if rc != 0:
            raise RuntimeError("orc_render failed")
        return img, aov_arr, dict(zip(COUNTER_NAMES, [int(v) for v in counters]))

    def grid_hit(self, origin, direction):
        """One GridHit of CreateRay(origin, direction) (brick_raytracer.comp:271-376). Returns (hit: bool, aov record)."""
        out = np.zeros(1, dtype=AOV_DTYPE)
        o = (C.c_float * 3)(*[float(v) for v in origin])
        d = (C.c_float * 3)(*[float(v) for v in direction])
        rc = lib().orc_grid_hit(C.byref(self.c), C.byref(o), C.byref(d), _ptr(out))
        return bool(rc), out[0]


def denoise(image: np.ndarray, params=(20, 0.6, 1.5, 20.0), out_width=None, out_height=None, flags=0, threads=0) -> np.ndarray:
    """image.frag:31-79 over an (H, W, 4) uint8 image; params = (samples, distribution_bias, pixel_multiplier,
    inverse_hue_tolerance) or any ctypes struct of that layout.  Returns the (out_height, out_width, 4) uint8 result."""
    image = np.ascontiguousarray(image, dtype=np.uint8)
    h, w = image.shape[:2]
    ow, oh = out_width or w, out_height or h
    if isinstance(params, C.Structure):
        pbuf = params
    else:
        class _P(C.Structure):
            _fields_ = [("samples", C.c_int32), ("distribution_bias", C.c_float), ("pixel_multiplier", C.c_float), ("inverse_hue_tolerance", C.c_float)]
        pbuf = _P(int(params[0]), float(params[1]), float(params[2]), float(params[3]))
    out = np.empty((oh, ow, 4), dtype=np.uint8)
    rc = lib().orc_denoise(_ptr(image), w, h, C.cast(C.pointer(pbuf), C.c_void_p), ow, oh, flags, _ptr(out), threads)
    if rc != 0:
        raise ValueError("orc_denoise rejected its arguments")
    return out


def algorithmic_bytes(counters: dict, n_pixels: int, brick_bytes: int) -> int:
    """Request-byte model of the reference algorithm (DESIGN.md "Algorithmic bytes"):
    4*pixels + sum_rays[4*S + (4 + brick_bytes)*B + 25*H]."""
    return 4 * n_pixels + 4 * counters["status_fetches"] + (4 + brick_bytes) * counters["bricks_entered"] + 25 * counters["hits"]
