"""The C ABI: struct layouts equal the reference's extern structs, and the built libraries export every symbol the
headers declare.  No compute calls here (no GPU needed)."""
import ctypes as C
import os
import re

import pytest

from zig_vulkan_b200 import ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_camera_device_layout():
    # Camera.zig:183-193 — Zig @Vector(3,f32) is 16 bytes wide and 16-byte aligned; GLSL mirror brick_raytracer.comp:58-68
    assert C.sizeof(ffi.CameraDevice) == 96
    off = {n: getattr(ffi.CameraDevice, n).offset for n, *_ in ffi.CameraDevice._fields_}
    assert off["image_width"] == 0 and off["image_height"] == 4
    assert off["horizontal"] == 16 and off["vertical"] == 32 and off["lower_left_corner"] == 48 and off["origin"] == 64
    assert off["samples_per_pixel"] == 80 and off["max_bounce"] == 84


def test_sun_device_layout():
    # Sun.zig:13-18, pushed at offset 96 (ComputePipeline.zig:259-263, 497-505)
    assert C.sizeof(ffi.SunDevice) == 32
    assert ffi.SunDevice.position.offset == 0 and ffi.SunDevice.enabled.offset == 12
    assert ffi.SunDevice.color.offset == 16 and ffi.SunDevice.radius.offset == 28
    assert C.sizeof(ffi.CameraDevice) + C.sizeof(ffi.SunDevice) == 128


def test_grid_state_layout():
    # State.zig:60-79 == BrickGridState UBO brick_raytracer.comp:79-95
    assert C.sizeof(ffi.GridState) == 64
    assert ffi.GridState.dim_x.offset == 12
    assert ffi.GridState.min_point_base_t.offset == 32 and ffi.GridState.max_point_scale.offset == 48


def test_material_layout():
    # gpu_types.zig:16-32: 20 bytes, std430 stride 20
    assert C.sizeof(ffi.Material) == 20
    assert ffi.MATERIAL_DTYPE.itemsize == 20
    assert ffi.Material.type_data.offset == 16


def test_denoise_push_constant_layout():
    # GraphicsPipeline.PushConstant (GraphicsPipeline.zig:27-32): i32 + 3 x f32 = 16 bytes, pushed as one block
    assert C.sizeof(ffi.DenoiseParams) == 16
    assert [ffi.DenoiseParams.samples.offset, ffi.DenoiseParams.distribution_bias.offset, ffi.DenoiseParams.pixel_multiplier.offset,
            ffi.DenoiseParams.inverse_hue_tolerance.offset] == [0, 4, 8, 12]
    d = ffi.DenoiseParams.default()  # GraphicsPipeline.Config defaults (:34-39)
    assert (d.samples, round(d.distribution_bias, 3), d.pixel_multiplier, d.inverse_hue_tolerance) == (20, 0.6, 1.5, 20.0)
    assert C.sizeof(ffi.BenchmarkReport) == 48
    assert ffi.RAY_DTYPE.itemsize == 32 and ffi.RAY_HIT_DTYPE.itemsize == 32  # explicit-ray mode: two 128-bit words each


def test_aov_and_config_layout():
    assert C.sizeof(ffi.Aov) == 64 and ffi.AOV_DTYPE.itemsize == 64
    assert C.sizeof(ffi.Counters) == 64
    assert C.sizeof(ffi.Config) == 64


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vrt_[a-z0-9_]+)\s*\(", text)) - {"vrt_emit_fn"})


@pytest.mark.parametrize("header,libname,table", [("vrt.h", "libvrt.so", ffi.VRT_SYMBOLS), ("vrt_host.h", "libvrt_host.so", ffi.VRT_HOST_SYMBOLS)])
def test_library_exports_every_declared_symbol(header, libname, table):
    names = _declared(header)
    assert len(names) >= 20
    dll = C.CDLL(os.path.join(ROOT, "zig_vulkan_b200", libname), mode=C.RTLD_GLOBAL) if libname == "libvrt.so" else ffi.host_lib()
    for n in names:
        assert hasattr(dll, n), f"{libname} does not export {n} declared in include/{header}"
    # the ctypes table binds exactly the declared set, so a header change cannot silently go unbound
    assert sorted(table) == names


def test_init_rejects_bad_config_without_touching_the_gpu():
    l = ffi.lib()
    h = C.c_void_p()
    cfg = ffi.Config(C.sizeof(ffi.Config), ffi.VRT_ABI_VERSION, 0, 0, 4, 256, 64, 64, 0, 0, 0, 0, 0, 0)
    assert l.vrt_init(C.byref(h), C.byref(cfg)) == -1  # VRT_E_INVALID: zero-sized image
    assert b"image size" in l.vrt_last_error(None)
    cfg = ffi.Config(C.sizeof(ffi.Config) - 4, ffi.VRT_ABI_VERSION, 16, 16, 4, 256, 64, 64, 0, 0, 0, 0, 0, 0)
    assert l.vrt_init(C.byref(h), C.byref(cfg)) == -1
    assert b"ABI mismatch" in l.vrt_last_error(None)
    cfg = ffi.Config(C.sizeof(ffi.Config), ffi.VRT_ABI_VERSION, 16, 16, 5, 256, 64, 64, 0, 0, 0, 0, 0, 0)
    assert l.vrt_init(C.byref(h), C.byref(cfg)) == -1
    assert b"brick_dim" in l.vrt_last_error(None)
    assert l.vrt_init(None, C.byref(cfg)) == -1
    # NULL handles never crash
    assert l.vrt_trace(None, None, None) == -1
    assert l.vrt_sync(None) == -1
    l.vrt_deinit(None)


def test_no_silent_cpu_fallback(has_cuda):
    """Without a CUDA device vrt_init must fail loudly (the product has no CPU path)."""
    if has_cuda:
        pytest.skip("a GPU is present")
    with pytest.raises(ffi.VrtError) as e:
        ffi.Context(16, 16, 64)
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)


def test_headers_are_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: both public headers must compile as C99 (no C++ constructs, no torch types)."""
    import shutil
    import subprocess

    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "hdr.c"
    src.write_text('#include "include/vrt.h"\n#include "include/vrt_host.h"\n'
                   'int main(void) { vrt_config c; vrt_denoise_params p = {20, 0.6f, 1.5f, 20.0f}; (void)c; (void)p; return 0; }\n')
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", ROOT, str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_c_example_builds_and_fails_loudly_without_a_gpu(tmp_path, has_cuda):
    """examples/render_frame.c (plain C99 against both libraries) compiles and links; without a CUDA device it must stop with
    the library's own error — there is no CPU path to fall back to."""
    import shutil
    import subprocess

    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    pkg = os.path.join(ROOT, "zig_vulkan_b200")
    exe = str(tmp_path / "render_frame")
    r = subprocess.run([cc, "-std=c99", "-O1", "-Wall", "-Wextra", "-Werror", "-I", ROOT, os.path.join(ROOT, "examples", "render_frame.c"), "-L", pkg,
                        "-lvrt_host", "-lvrt", "-Wl,-rpath," + pkg, "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if not has_cuda:
        run = subprocess.run([exe, "64", str(tmp_path / "f")], capture_output=True, text=True, timeout=120)
        assert run.returncode != 0 and "no CPU fallback" in run.stderr
