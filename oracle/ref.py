"""ctypes wrapper of oracle/_ref/libref_shader.so — the reference's OWN shaders (brick_raytracer.comp + rand.comp, image.frag)
compiled by g++ under oracle/ref_shim/glsl_compat.h.  TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__ and bench.py's CPU legs may import this module.  It is what pins the hand-written oracle
(oracle/vrt_oracle*.cpp) to the reference's text: tests/test_ref_shader.py renders every golden case through it and requires
the oracle's output bit for bit.  Built by `make -C oracle ref` where /root/reference exists; the GPU box gets the prebuilt .so.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import orc

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libref_shader.so")
REF_SHADERS = "/root/reference/assets/shaders"
_PATH_WIDE = os.path.join(_HERE, "_ref", "libref_shader_wide.so")
_lib = None
_lib_wide = None

HIT_DTYPE = np.dtype([("hit", "<u4"), ("index", "<u4"), ("t", "<f4"), ("point", "<f4", 3), ("normal", "<f4", 3)])


def available() -> bool:
    """True when the library exists or can be built here (the reference tree is present)."""
    return os.path.exists(_PATH) or os.path.exists(os.path.join(REF_SHADERS, "brick_raytracer.comp"))


def build():
    r = subprocess.run(["make", "-C", _HERE, "ref"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("make -C oracle ref failed:\n" + r.stdout[-2000:] + r.stderr[-4000:])


def lib(wide: bool = False) -> C.CDLL:
    """wide: the variant for 16^3 bricks (mask byte index widened, oracle/ref_shim/translate.py)."""
    global _lib, _lib_wide
    if (_lib_wide if wide else _lib) is None:
        if os.path.exists(os.path.join(REF_SHADERS, "brick_raytracer.comp")):
            build()  # make decides whether anything is stale
        path = _PATH_WIDE if wide else _PATH
        if not os.path.exists(path):
            raise FileNotFoundError(path + " is missing and /root/reference is not here to build it from")
        l = C.CDLL(path)
        P = C.c_void_p
        l.ref_trace_render.argtypes = [P, P, P, C.c_uint32, C.c_uint32, P, P, C.c_int]
        l.ref_trace_render.restype = C.c_int
        l.ref_trace_grid_hit.argtypes = [P, P, P, P]
        l.ref_trace_grid_hit.restype = C.c_int
        l.ref_trace_hash12.argtypes = [C.c_float, C.c_float]
        l.ref_trace_hash12.restype = C.c_float
        l.ref_trace_rand2.argtypes = [C.c_float, C.c_float]
        l.ref_trace_rand2.restype = C.c_float
        l.ref_present_render.argtypes = [P, C.c_uint32, C.c_uint32, P, C.c_uint32, C.c_uint32, C.c_uint32, P, C.c_int]
        l.ref_present_render.restype = C.c_int
        if wide:
            _lib_wide = l
        else:
            _lib = l
    return _lib_wide if wide else _lib


def render(scene: orc.OracleScene, camera, sun, rows=None, hits=False, threads=0):
    """brick_raytracer.comp main() per pixel of rows [begin, end).  Returns (rgba8[H,W,4], hits[H,W] or None)."""
    w, h = camera.image_width, camera.image_height
    r0, r1 = rows if rows else (0, h)
    img = np.zeros((h, w, 4), dtype=np.uint8)
    hit_arr = np.zeros((h, w), dtype=HIT_DTYPE) if hits else None
    rc = lib(scene.c.brick_dim > 8).ref_trace_render(C.byref(scene.c), C.byref(camera), C.byref(sun), r0, r1, img.ctypes.data, hit_arr.ctypes.data if hits else None, threads)
    if rc != 0:
        raise RuntimeError(f"ref_trace_render failed ({rc})")
    return img, hit_arr


def grid_hit(scene: orc.OracleScene, origin, direction):
    out = np.zeros(1, dtype=HIT_DTYPE)
    o = (C.c_float * 3)(*[float(v) for v in origin])
    d = (C.c_float * 3)(*[float(v) for v in direction])
    rc = lib(scene.c.brick_dim > 8).ref_trace_grid_hit(C.byref(scene.c), C.byref(o), C.byref(d), out.ctypes.data)
    return bool(rc), out[0]


def present(image: np.ndarray, params=(20, 0.6, 1.5, 20.0), out_width=None, out_height=None, flags=0, threads=0) -> np.ndarray:
    """image.frag main() per output texel (same arguments as orc.denoise)."""
    image = np.ascontiguousarray(image, dtype=np.uint8)
    h, w = image.shape[:2]
    ow, oh = out_width or w, out_height or h

    class _P(C.Structure):
        _fields_ = [("samples", C.c_int32), ("distribution_bias", C.c_float), ("pixel_multiplier", C.c_float), ("inverse_hue_tolerance", C.c_float)]

    pbuf = _P(int(params[0]), float(params[1]), float(params[2]), float(params[3]))
    out = np.empty((oh, ow, 4), dtype=np.uint8)
    rc = lib().ref_present_render(image.ctypes.data, w, h, C.byref(pbuf), ow, oh, flags, out.ctypes.data, threads)
    if rc != 0:
        raise ValueError("ref_present_render rejected its arguments")
    return out
