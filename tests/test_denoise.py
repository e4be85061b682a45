"""The present pass (assets/shaders/image.frag:31-79): oracle known-answer tests on the CPU, CUDA-vs-oracle parity on the GPU.

The reference has no test, golden image or CPU form of this shader; image.frag itself, compiled under oracle/ref_shim/, pins the
oracle bit for bit (tests/test_ref_shader.py).  Independently of that the oracle is checked here by (i) an independent float64 numpy transcription of the shader written here, compared at +-1 LSB (the two
differ in pow / rounding at the 1e-6 level, which can move a value across a rounding boundary), (ii) closed-form cases, and
(iii) a committed golden frame.  The GPU gate is bit-exact: same FP32 operations in the same order on both sides."""
import os

import numpy as np
import pytest

import zig_vulkan_b200 as zv
from zig_vulkan_b200 import ffi, scenes
from oracle import orc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "denoise_c1_64_256x256.npz")
DEFAULT = (20, 0.6, 1.5, 20.0)  # GraphicsPipeline.zig:34-39


def shader_f64(img, params=DEFAULT, out_w=None, out_h=None):
    """image.frag transcribed line by line in float64 numpy (vectorised over fragments)."""
    samples, bias, mult, tol = params
    h, w = img.shape[:2]
    ow, oh = out_w or w, out_h or h
    tex = img[..., :3].astype(np.float64) / 255.0

    def texture(u, v):  # linear, repeat (Pipeline.zig:193-212)
        x, y = u * w - 0.5, v * h - 0.5
        fx, fy = np.floor(x), np.floor(y)
        a, b = (x - fx)[..., None], (y - fy)[..., None]
        x0, y0 = fx.astype(int) % w, fy.astype(int) % h
        x1, y1 = (x0 + 1) % w, (y0 + 1) % h
        return (1 - a) * (1 - b) * tex[y0, x0] + a * (1 - b) * tex[y0, x1] + (1 - a) * b * tex[y1, x0] + a * b * tex[y1, x1]

    def gpow(a, b):  # :29
        return np.power(np.maximum(a, 0.0), b)

    uvx, uvy = np.meshgrid((np.arange(ow) + 0.5) / ow, (np.arange(oh) + 0.5) / oh)
    golden = 2.3999632
    c, s = np.cos(golden), np.sin(golden)
    radius = np.sqrt(float(samples))
    true_radius = 0.5 / (radius * radius)
    center = texture(uvx, uvy)
    with np.errstate(invalid="ignore", divide="ignore"):
        center_norm = center / np.linalg.norm(center, axis=-1, keepdims=True)
        center_sat = np.linalg.norm(center, axis=-1)
        denoised = np.zeros_like(center)
        influence_sum = np.zeros(center.shape[:2])
        rot = np.array([0.0, 1.0])
        for k in range(samples + 1):
            rot = np.array([rot[0] * c + rot[1] * s, rot[0] * -s + rot[1] * c])  # v * mat2(c, s, -s, c)
            off = mult * rot * np.sqrt(float(k)) * 0.5
            influence = 1.0 - true_radius * gpow(off @ off, bias)
            this = texture(uvx + off[0] / w, uvy + off[1] / h)
            influence = influence * influence * influence
            this_len = np.linalg.norm(this, axis=-1)
            hue = gpow(0.5 + 0.5 * np.sum(center_norm * (this / this_len[..., None]), axis=-1), tol)
            sat = gpow(1.0 - np.abs(this_len - np.abs(center_sat)), 8.0)
            weight = influence * hue * sat
            influence_sum += weight
            denoised += this * weight[..., None]
        col = denoised / influence_sum[..., None]
    col = np.where(np.isnan(col), 0.0, np.clip(col, 0.0, 1.0))
    out = np.full((oh, ow, 4), 255, dtype=np.uint8)
    out[..., :3] = np.floor(col * 255.0 + 0.5).astype(np.uint8)
    return out


def blocky_image(w, h, seed, lo=40, hi=220):
    """Piecewise-flat colour blocks plus a little noise: edges for the hue / saturation filter to act on, no black texels."""
    rng = np.random.default_rng(seed)
    img = np.full((h, w, 4), 255, dtype=np.uint8)
    blocks = rng.integers(lo, hi, ((h + 9) // 10, (w + 9) // 10, 3))
    img[..., :3] = (np.repeat(np.repeat(blocks, 10, 0), 10, 1)[:h, :w] + rng.integers(0, 20, (h, w, 3))).astype(np.uint8)
    return img


# ------------------------------------------------------------------------------------------------- oracle KATs (CPU)

def test_pow_matches_libm_and_edge_cases():
    l = orc.lib()
    rng = np.random.default_rng(0)
    for a, b in zip(rng.uniform(0.0, 12.0, 4000), rng.choice([0.6, 8.0, 20.0, 1.0, 2.5, 0.25], 4000)):
        ref = float(np.float32(a)) ** float(np.float32(b))
        if ref > 1e-30:
            assert abs(l.orc_pow(a, b) - ref) <= 1e-5 * ref
    assert l.orc_pow(0.0, 0.6) == 0.0          # pow(0, b) = 0
    assert l.orc_pow(-3.0, 2.0) == 0.0         # the shader's max(a, 0.) macro (:29)
    assert l.orc_pow(1.0, 20.0) == 1.0
    assert l.orc_pow(4.0, 0.5) == pytest.approx(2.0, rel=1e-6)
    assert np.isnan(l.orc_pow(float("nan"), 2.0))


def test_flat_image_is_a_fixed_point():
    # every sample equals the centre: hue and saturation weights are pow(1, .) = 1, so the weighted mean is the colour itself
    for value in (1, 37, 128, 255):
        img = np.full((12, 20, 4), value, dtype=np.uint8)
        img[..., 3] = 255
        assert np.array_equal(orc.denoise(img), img)
        assert np.array_equal(orc.denoise(img, out_width=33, out_height=7), np.full((7, 33, 4), (value, value, value, 255), dtype=np.uint8))


def test_black_centre_is_nan_and_stores_zero():
    # normalize((0,0,0)) = 0 * inf = NaN poisons every weight of that fragment (as written upstream); NaN stores 0
    img = np.full((9, 9, 4), 128, dtype=np.uint8)
    img[..., 3] = 255
    img[4, 4, :3] = 0
    out = orc.denoise(img)
    assert tuple(out[4, 4]) == (0, 0, 0, 255)
    assert out[0, 0, 0] == 128


def test_zero_samples_is_all_nan():
    # samples = 0: sampleTrueRadius = 0.5 / 0 = inf and pow(0, bias) = 0 -> inf * 0 = NaN (image.frag:36,52)
    img = blocky_image(16, 8, 5)
    assert not orc.denoise(img, params=(0, 0.6, 1.5, 20.0))[..., :3].any()


def test_bgra_swaps_red_and_blue():
    img = blocky_image(24, 16, 2)
    rgba, bgra = orc.denoise(img), orc.denoise(img, flags=ffi.VRT_DENOISE_BGRA)
    assert np.array_equal(rgba[..., [2, 1, 0, 3]], bgra)


@pytest.mark.parametrize("params", [DEFAULT, (5, 0.3, 3.0, 2.0), (40, 1.0, 1.0, 30.0)])
@pytest.mark.parametrize("size", [(40, 30, None, None), (40, 30, 60, 45), (33, 17, 20, 11)])
def test_oracle_matches_float64_transcription(params, size):
    w, h, ow, oh = size
    img = blocky_image(w, h, seed=w * 131 + h)
    got = orc.denoise(img, params=params, out_width=ow, out_height=oh)
    ref = shader_f64(img, params, ow, oh)
    assert got.shape == ref.shape
    diff = np.abs(got.astype(int) - ref.astype(int))
    assert diff.max() <= 1, f"max diff {diff.max()}"
    assert (diff > 0).mean() < 0.02  # rounding-boundary cases only


def test_oracle_reproduces_golden_denoise():
    z = np.load(GOLDEN)
    assert np.array_equal(orc.denoise(z["traced"]), z["denoised"])
    assert np.array_equal(orc.denoise(z["traced"], out_width=384, out_height=216, flags=ffi.VRT_DENOISE_BGRA), z["denoised_384x216_bgra"])


# ------------------------------------------------------------------------------------------------- CUDA parity (GPU)

def _ctx_with_frame(img):
    """A context whose framebuffer holds `img` (attached caller-owned device image)."""
    import torch

    h, w = img.shape[:2]
    ctx = ffi.Context(w, h, 8)
    dev = torch.from_numpy(img.copy()).cuda()
    ctx._check(ctx._l.vrt_attach_framebuffer(ctx.handle, dev.data_ptr(), dev.numel()))
    return ctx, dev


@pytest.mark.gpu
# the first six take the padded fast path (whole hue tolerance 2..64, sample offsets within the 16-texel border: default 20, even,
# odd, 255 samples, none); the last three the general kernel (fractional tolerance, tolerance 1, offsets beyond the border)
@pytest.mark.parametrize("params", [DEFAULT, (1, 0.6, 1.5, 20.0), (64, 0.9, 2.5, 4.0), (255, 0.5, 1.0, 10.0), (0, 0.6, 1.5, 20.0), (20, 0.6, 1.5, 3.0),
                                    (20, 0.6, 1.5, 2.5), (12, 0.6, 1.5, 1.0), (30, 0.6, 8.0, 20.0)])
@pytest.mark.parametrize("size", [(160, 90, None, None), (160, 90, 240, 135), (67, 41, 50, 29), (1, 1, 3, 2)])
def test_cuda_denoise_is_bit_exact(params, size):
    w, h, ow, oh = size
    img = blocky_image(w, h, seed=7 * w + h, lo=0, hi=236)  # includes black-ish texels -> NaN paths
    ctx, dev = _ctx_with_frame(img)
    p = ffi.DenoiseParams(*params)
    for flags in (0, ffi.VRT_DENOISE_BGRA):
        got = ctx.denoise(p, ow, oh, flags)
        ref = orc.denoise(img, params=p, out_width=ow, out_height=oh, flags=flags)
        assert np.array_equal(got, ref), f"{(got != ref).any(axis=2).sum()} texels differ"
    assert ctx.last_denoise_ms() > 0.0
    ctx.close()
    del dev


@pytest.mark.gpu
def test_cuda_trace_then_denoise_matches_golden(materials):
    z = np.load(GOLDEN)
    grid = scenes.build_grid(64)
    cam = scenes.camera(256, 256, origin=(0.0, -10.0, 28.0), euler_deg=(25.0, 0.0, 0.0))
    ctx = ffi.Context(256, 256, len(grid.brick_indices))
    ctx.upload_grid(grid, materials)
    traced = ctx.trace_to_host(cam, scenes.sun(False))
    assert np.array_equal(traced, z["traced"])
    assert np.array_equal(ctx.denoise(), z["denoised"])
    assert np.array_equal(ctx.denoise(None, 384, 216, ffi.VRT_DENOISE_BGRA), z["denoised_384x216_bgra"])
    ctx.close()


@pytest.mark.gpu
def test_cuda_denoise_full_frame_properties():
    # 1920x1080: a flat frame is a fixed point; a random frame agrees with the oracle on a band of rows
    import torch

    flat = np.full((1080, 1920, 4), 255, dtype=np.uint8)
    flat[..., :3] = (90, 140, 200)
    ctx, dev = _ctx_with_frame(flat)
    assert np.array_equal(ctx.denoise(), flat)
    img = blocky_image(1920, 1080, seed=11)
    dev.copy_(torch.from_numpy(img).cuda())
    got = ctx.denoise()
    ref = orc.denoise(img)
    assert np.array_equal(got, ref)
    ctx.close()


@pytest.mark.gpu
def test_cuda_denoise_rejects_bad_arguments():
    ctx = ffi.Context(16, 16, 8)
    l, hnd = ctx._l, ctx.handle
    ok = ffi.DenoiseParams.default()
    import ctypes as C

    assert l.vrt_denoise(hnd, None, 16, 16, 0) == ffi.VRT_E_INVALID
    assert l.vrt_denoise(hnd, C.byref(ffi.DenoiseParams(256, 0.6, 1.5, 20.0)), 16, 16, 0) == ffi.VRT_E_INVALID
    assert l.vrt_denoise(hnd, C.byref(ffi.DenoiseParams(-1, 0.6, 1.5, 20.0)), 16, 16, 0) == ffi.VRT_E_INVALID
    assert l.vrt_denoise(hnd, C.byref(ok), 0, 16, 0) == ffi.VRT_E_INVALID
    assert l.vrt_denoise(hnd, C.byref(ok), 16, 16, 2) == ffi.VRT_E_INVALID
    buf = np.empty(16 * 16 * 4, dtype=np.uint8)
    assert l.vrt_read_denoised(hnd, buf.ctypes.data, buf.nbytes) == ffi.VRT_E_STATE  # nothing denoised yet
    assert l.vrt_denoise(hnd, C.byref(ok), 16, 16, 0) == ffi.VRT_OK
    assert l.vrt_read_denoised(hnd, buf.ctypes.data, 5) == ffi.VRT_E_INVALID
    ctx.close()
