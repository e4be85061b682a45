"""The oracle is PINNED to the reference's own text: oracle/_ref/libref_shader.so is assets/shaders/brick_raytracer.comp +
rand.comp and image.frag themselves, compiled by g++ under oracle/ref_shim/glsl_compat.h (translate.py's lexical pass only).
Everything here is bit-exact: the hand-written oracle (oracle/vrt_oracle*.cpp), the committed golden vectors and — on the GPU box —
the CUDA kernels must reproduce what the reference's shader text computes."""
import importlib.util
import os

import numpy as np
import pytest

import zig_vulkan_b200 as zv
from zig_vulkan_b200 import ffi, scenes
from oracle import orc, ref

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(spec)
spec.loader.exec_module(make_golden)

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref is not built and /root/reference is not here to build it from")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def assert_hits_equal(aov, hits):
    hit = (aov["flags"] & 1) != 0
    assert np.array_equal(hit, hits["hit"] != 0)
    assert np.array_equal(aov["material"][hit], hits["index"][hit])
    for f in ("t", "point", "normal"):
        assert np.array_equal(bits(aov[f])[hit], bits(hits[f])[hit]), f


def test_translation_is_lexical_only():
    """Every line of the translated shaders is the reference's line up to translate.py's listed substitutions (f suffixes, swizzle
    calls, references for out / inout, main -> shader_main); the only lines that come or go are resource declarations."""
    import re
    from collections import Counter

    src_dir = "/root/reference/assets/shaders"
    if not os.path.exists(os.path.join(src_dir, "brick_raytracer.comp")):
        pytest.skip("reference tree not present")
    ref.lib()

    def norm(s):  # undo the expression-level substitutions
        s = re.sub(r"(?<![\w.])(\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)f(?![\w.])", r"\1", s)
        s = re.sub(r"\.([xyzw]{2,4}|[rgba]{2,4})\(\)", r".\1", s)
        s = re.sub(r"\b(?:in)?out\s+(\w+)\s+(\w+)", r"\1& \2", s)
        return s.replace("shader_main", "main").strip()

    gone_ok = re.compile(r"^(#version|#extension|#include|layout\b|readonly\b|//|\} (push_constant|pushConstant|brick_grid);|\};)")
    new_ok = re.compile(r"^(#include \"rand\.comp\.inc\"|Buffer<\w+> \w+;|struct \w+ \{|image2D \w+;|sampler2D \w+;|thread_local vec[24] \w+;|"
                        r"(uint|int|float) brick_\w+ = [\w.]+;|\} (push_constant|pushConstant|brick_grid);|\};)$")
    for name in ("brick_raytracer.comp", "rand.comp", "image.frag"):
        a = Counter(norm(l) for l in open(os.path.join(src_dir, name)).read().replace("\r\n", "\n").split("\n") if l.strip())
        b = Counter(norm(l) for l in open(os.path.join(os.path.dirname(HERE), "oracle", "_ref", name + ".inc")).read().split("\n")[1:] if l.strip())
        gone, new = a - b, b - a
        assert all(gone_ok.match(l) for l in gone), [l for l in gone if not gone_ok.match(l)]
        assert all(new_ok.match(l) for l in new), [l for l in new if not new_ok.match(l)]
        assert sum(gone.values()) < 40 and sum(new.values()) < 20


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_reference_shader_reproduces_golden(name):
    """The reference's shader text renders every committed golden frame bit for bit (RGBA8 and the primary hit records)."""
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    n, bd, w, h, sun_on, radius, spp, bounce, pose = make_golden.CASES[name]
    grid = scenes.build_grid(n, brick_dim=bd)
    sc = orc.OracleScene.from_grid(grid, zv.terrain_materials())
    cam = scenes.camera(w, h, spp=spp, max_bounce=bounce, **pose)
    sun = scenes.sun(sun_on, radius)
    img, hits = ref.render(sc, cam, sun, hits=True)
    assert np.array_equal(img, z["rgba"]), f"{(img != z['rgba']).any(axis=2).sum()} pixels differ"
    assert_hits_equal(z["aov"], hits)


def test_uint8_mask_index_wraps_at_16_cubed():
    """Why brick_dim 16 needs the documented extension: the UNMODIFIED shader's 8-bit mask byte index (:413) wraps above 8^3 voxels
    and renders a different frame; the oracle (and the kernels) follow the widened text."""
    n, bd, w, h, sun_on, radius, spp, bounce, pose = make_golden.CASES["bd16_128_160x90_sun"]
    grid = scenes.build_grid(n, brick_dim=bd)
    sc = orc.OracleScene.from_grid(grid, zv.terrain_materials())
    cam, sun = scenes.camera(w, h, **pose), scenes.sun(sun_on, radius)
    img = np.zeros((h, w, 4), dtype=np.uint8)
    import ctypes as C

    rc = ref.lib(False).ref_trace_render(C.byref(sc.c), C.byref(cam), C.byref(sun), 0, h, img.ctypes.data, None, 0)
    assert rc == 0
    want = np.load(os.path.join(HERE, "golden", "bd16_128_160x90_sun.npz"))["rgba"]
    assert (img != want).any()


@pytest.mark.parametrize("bd,n", [(4, 64), (8, 64), (16, 128)])
def test_grid_hit_random_rays(bd, n):
    """GridHit / BrickHit of the reference text vs the oracle on random rays from inside, outside and on the faces of the grid,
    including zero direction components (safeInverse, :267)."""
    grid = scenes.build_grid(n, brick_dim=bd)
    sc = orc.OracleScene.from_grid(grid, zv.terrain_materials())
    rng = np.random.default_rng(1234 + bd)
    mismatches = 0
    for i in range(1500):
        o = rng.uniform(-45.0, 45.0, 3).astype(np.float32)
        d = rng.normal(size=3).astype(np.float32)
        if i % 7 == 0:
            d[rng.integers(0, 3)] = 0.0
        if i % 11 == 0:
            o = np.round(o)  # origins on cell faces
        if not np.any(d):
            continue
        got_o, rec_o = sc.grid_hit(o, d)
        got_r, rec_r = ref.grid_hit(sc, o, d)
        assert got_o == got_r, (i, o, d)
        if got_o:
            assert rec_o["material"] == rec_r["index"]
            for f in ("t", "point", "normal"):
                assert np.array_equal(bits(rec_o[f]), bits(rec_r[f])), (i, f, o, d)
        mismatches += 0
    assert mismatches == 0


@pytest.mark.parametrize("case", [
    dict(n=64, bd=4, w=97, h=61, sun=True, radius=5.0, spp=3, bounce=3, pose=dict(origin=(5.0, -6.0, 20.0), euler_deg=(15.0, 30.0, 0.0))),
    dict(n=128, bd=4, w=120, h=68, sun=False, radius=0.0, spp=2, bounce=4, pose=dict(origin=(0.0, -10.0, 28.0), euler_deg=(25.0, 0.0, 0.0))),
    dict(n=64, bd=8, w=64, h=48, sun=True, radius=2.0, spp=1, bounce=2, pose=dict(origin=(-12.0, -3.0, -9.0), euler_deg=(5.0, -120.0, 0.0))),
    dict(n=128, bd=16, w=64, h=48, sun=True, radius=0.0, spp=2, bounce=1, pose=dict(origin=(0.0, -10.0, 28.0), euler_deg=(25.0, 0.0, 0.0))),
    dict(n=64, bd=4, w=80, h=45, sun=True, radius=0.0, spp=1, bounce=0, pose=dict(origin=(0.0, -8.0, 0.0), euler_deg=(0.0, 0.0, 0.0))),
], ids=["look_spp3_bounce3", "nosun_bounce4", "bd8_sun_disc", "bd16_spp2", "survey_pose0"])
def test_full_path_trace_modes_match_reference_text(case):
    """main() + RayColor + the scatter functions + the sin-hash RNG (spp > 1, bounces, sun disc): oracle == reference text."""
    grid = scenes.build_grid(case["n"], brick_dim=case["bd"])
    sc = orc.OracleScene.from_grid(grid, zv.terrain_materials())
    cam = scenes.camera(case["w"], case["h"], spp=case["spp"], max_bounce=case["bounce"], **case["pose"])
    sun = scenes.sun(case["sun"], case["radius"])
    want, aov, _ = sc.render(cam, sun, aov=True)
    img, hits = ref.render(sc, cam, sun, hits=True)
    assert np.array_equal(img, want), f"{(img != want).any(axis=2).sum()} pixels differ"
    assert_hits_equal(aov, hits)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("seed", range(40))
def test_seeded_random_configurations_match_reference_text(seed):
    """A seeded sweep nobody tuned by hand (tests/random_configs.py): non-cubic grids, non-power-of-two scales, random material tables
    (every type, random fuzz / refraction index), cameras inside and outside the grid, 1-3 samples, 0-4 bounces, point / disc / no sun.
    Oracle == the reference's shader text, frame and hit records, bit for bit."""
    from random_configs import random_configuration

    c = random_configuration(seed)
    img, hits = ref.render(c["scene"], c["cam"], c["sun"], hits=True)
    assert np.array_equal(img, c["image"]), f"{(img != c['image']).any(axis=2).sum()} pixels differ"
    assert_hits_equal(c["aov"], hits)


def test_custom_materials_and_ignore_rule():
    """Materials of every type incl. MAT_NONE (3) and an unknown type (the `default:` arm, :234-237), dielectric index = 1.0
    so that the ignore rule of :427 can fire for camera rays."""
    mats = zv.terrain_materials().copy()
    mats[1]["type"], mats[1]["type_data"] = 3, 1.0   # MAT_NONE with type_data == a fresh ray's internal_reflection: ignored voxels
    mats[2]["type"] = 7                               # unknown type: default arm
    mats[3]["type"], mats[3]["type_data"] = 2, 1.0    # dielectric with ir 1.0
    mats[4]["type"], mats[4]["type_data"] = 1, 0.3    # fuzzy metal
    grid = scenes.build_grid(64)
    sc = orc.OracleScene.from_grid(grid, mats)
    cam = scenes.camera(96, 54, spp=2, max_bounce=3, origin=(0.0, -10.0, 28.0), euler_deg=(25.0, 0.0, 0.0))
    for sun in (scenes.sun(True, 3.0), scenes.sun(False)):
        want, aov, _ = sc.render(cam, sun, aov=True)
        img, hits = ref.render(sc, cam, sun, hits=True)
        assert np.array_equal(img, want)
        assert_hits_equal(aov, hits)


def test_rng_primitives():
    l, o = ref.lib(), orc.lib()
    rng = np.random.default_rng(7)
    for x, y in rng.uniform(-300.0, 300.0, (200, 2)).astype(np.float32):
        assert bits(np.float32(l.ref_trace_hash12(x, y))) == bits(np.float32(o.orc_hash12(x, y)))


def test_present_pass_matches_reference_text():
    """image.frag: oracle == reference text on the golden frame and on random images / parameters / target sizes."""
    z = np.load(os.path.join(HERE, "golden", "denoise_c1_64_256x256.npz"))
    assert np.array_equal(ref.present(z["traced"]), z["denoised"])
    assert np.array_equal(ref.present(z["traced"], out_width=384, out_height=216, flags=ffi.VRT_DENOISE_BGRA), z["denoised_384x216_bgra"])
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)
    img[5:9, 5:9, :3] = 0  # black texels: normalize(0) = NaN poisons the weights, as upstream
    for params in [(20, 0.6, 1.5, 20.0), (5, 0.3, 3.0, 4.0), (0, 0.6, 1.5, 20.0), (64, 1.2, 0.7, 50.0)]:
        for ow, oh in [(53, 37), (80, 45), (16, 9)]:
            assert np.array_equal(ref.present(img, params, out_width=ow, out_height=oh), orc.denoise(img, params, out_width=ow, out_height=oh)), (params, ow, oh)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("seed", range(24))
def test_present_pass_seeded_random_parameters(seed):
    """image.frag on random images (flat regions, noise, black and saturated texels), random push constants (whole and fractional
    hue tolerance, small and large sample offsets, 0-80 samples) and random target sizes: oracle == the reference's text, byte for byte."""
    rng = np.random.default_rng(500 + seed)
    w, h = int(rng.integers(1, 70)), int(rng.integers(1, 50))
    img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    if seed % 3 == 0:  # large flat areas: pow(1, n) and equal weights
        img[: h // 2] = rng.integers(0, 256, 4, dtype=np.uint8)
    if seed % 4 == 1:
        img[rng.integers(0, h), rng.integers(0, w), :3] = 0  # a black texel: normalize(0) = NaN, as upstream
        img[rng.integers(0, h), rng.integers(0, w), :3] = 255
    params = (int(rng.integers(0, 81)), float(rng.choice([0.3, 0.6, 1.0, 1.7])), float(rng.choice([0.5, 1.5, 4.0, 9.0])),
              float(rng.choice([1.0, 2.0, 3.0, 20.0, 64.0, 65.0, 2.5, 0.5, 30.25])))
    ow, oh = (None, None) if seed % 2 else (int(rng.integers(1, 90)), int(rng.integers(1, 60)))
    flags = ffi.VRT_DENOISE_BGRA if seed % 5 == 0 else 0
    assert np.array_equal(ref.present(img, params, out_width=ow, out_height=oh, flags=flags), orc.denoise(img, params, out_width=ow, out_height=oh, flags=flags)), (params, w, h, ow, oh)


@pytest.mark.gpu
@pytest.mark.parametrize("kernel_flags", [0, ffi.VRT_FLAG_BASELINE], ids=["tuned", "baseline"])
def test_cuda_matches_reference_text(kernel_flags):
    """The CUDA kernels against the reference's shader text directly (no hand-written oracle in between)."""
    for n, bd, w, h, spp, bounce, radius in [(128, 4, 320, 180, 1, 0, 0.0), (64, 4, 160, 90, 2, 2, 5.0), (64, 8, 160, 90, 1, 0, 0.0), (128, 16, 160, 90, 1, 0, 0.0)]:
        grid = scenes.build_grid(n, brick_dim=bd)
        mats = zv.terrain_materials()
        sc = orc.OracleScene.from_grid(grid, mats)
        cam = scenes.camera(w, h, spp=spp, max_bounce=bounce, origin=(0.0, -10.0, 28.0), euler_deg=(25.0, 0.0, 0.0))
        sun = scenes.sun(True, radius)
        want, hits = ref.render(sc, cam, sun, hits=True)
        ctx = ffi.Context(w, h, len(grid.brick_indices), brick_dim=bd, flags=kernel_flags | ffi.VRT_FLAG_AOV)
        ctx.upload_grid(grid, mats)
        ctx.trace(cam, sun)
        assert np.array_equal(ctx.read_framebuffer(), want)
        assert_hits_equal(ctx.read_aov(), hits)
        ctx.close()
        ctx = ffi.Context(w, h, len(grid.brick_indices), brick_dim=bd, flags=kernel_flags)
        ctx.upload_grid(grid, mats)
        assert np.array_equal(ctx.trace_to_host(cam, sun), want)
        shown = ctx.denoise(None, w, h, 0)
        assert np.array_equal(shown, ref.present(want))
        ctx.close()
