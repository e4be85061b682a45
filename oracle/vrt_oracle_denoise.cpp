/*
 * vrt_oracle_denoise.cpp — CPU oracle of the reference's post-process pass.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Restates assets/shaders/image.frag:31-71 (sirBirdDenoise + main), the fragment shader the reference draws over a
 * full-screen quad (GraphicsPipeline.zig:17-22 vertices, uv 0..1) to put the compute image on the swapchain:
 *   - input  = the R8G8B8A8_UNORM compute image (Pipeline.zig:103-126) read through a sampler with linear min/mag filter and
 *     repeat addressing (Pipeline.zig:193-212),
 *   - params = the 16-byte push constant {samples, distributionBias, pixelMultiplier, inversHueTolerance}
 *     (GraphicsPipeline.zig:27-39, defaults 20 / 0.6 / 1.5 / 20),
 *   - output = one B8G8R8A8_UNORM swapchain texel per fragment (swapchain.zig:235), alpha 1.
 *
 * PARITY PINNED (see vrt_oracle.h): no CPU path, golden image or test exists upstream for this pass either, but image.frag itself
 * compiles under oracle/ref_shim/ and tests/test_ref_shader.py::test_present_pass_matches_reference_text holds this file to it bit for bit.
 *
 * Choices where GLSL / Vulkan leave the arithmetic to the implementation (the CUDA kernel is held to the same ones):
 *   - inUV of the fragment at output pixel (x, y) = ((x + 0.5) / out_width, (y + 0.5) / out_height).
 *   - texture(): unnormalised coordinate u * W - 0.5, i = floor, a = fraction; texels (i, i+1) wrapped (repeat); UNORM
 *     decode byte / 255.0f; weights in full FP32 (hardware samplers use ~8 fractional bits):
 *     ((w00*t00 + w10*t10) + w01*t01) + w11*t11 with w00 = (1-a)(1-b), w10 = a(1-b), w01 = (1-a)b, w11 = ab.
 *   - pow(a, b) (after the shader's own `max(a, 0.)` macro, image.frag:29) = exp2(b * log2(a)) with the explicit
 *     det_log2f / det_exp2f below (atanh series / degree-6 polynomial, only + - * / fmaf and bit operations), so that x86
 *     and the GPU agree bit for bit; pow(0, b) = 0; arguments below FLT_MIN count as 0.  GLSL's pow precision is
 *     "inherited from exp2(x * log2(y))", which this is — except for whole exponents 1..64 (the shader's literal 8 and the default
 *     inverse hue tolerance 20), which are a fixed chain of FP32 multiplications (binary exponentiation), the way a shader
 *     compiler folds pow(x, 8.): more accurate, and a third of the instructions of the pass.
 *   - cos / sin(GOLDEN_ANGLE) are the correctly rounded FP32 constants; sqrt and / are IEEE; normalize, dot, length, max,
 *     abs as in vrt_oracle.cpp (normalize(0) = 0 * inf = NaN: a black texel poisons the weights, exactly as written upstream).
 *   - UNORM store as in the trace path: clamp to [0,1], (uint8_t)(c * 255 + 0.5), NaN stores 0.
 *   - no contraction except where written as fmaf (build flags as for vrt_oracle.cpp).
 */
#include "vrt_oracle.h"

#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

inline float from_bits(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
inline uint32_t to_bits(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
inline float gmax(float a, float b) { return a < b ? b : a; }

// log2 of a finite a >= FLT_MIN: a = 2^e * m with m in [sqrt(1/2), sqrt(2)); ln m = 2 atanh(s), s = (m-1)/(m+1)
inline float det_log2f(float a) {
    const int32_t ia = (int32_t)to_bits(a);
    const int32_t e = (ia - 0x3f3504f3) >> 23;
    const float m = from_bits((uint32_t)(ia - e * (1 << 23)));
    const float f = m - 1.0f;
    const float s = f / (2.0f + f);
    const float z = s * s;
    float p = fmaf(z, 0.22222222f, 0.2857143f);
    p = fmaf(p, z, 0.4f);
    p = fmaf(p, z, 0.6666667f);
    const float ln_m = fmaf(s * z, p, s + s);
    return fmaf(ln_m, 1.442695f, (float)e);
}

// 2^y, y clamped to [-126, 126] (results are never denormal / infinite); NaN stays NaN
inline float det_exp2f(float y) {
    if (y != y) return y;
    y = y < -126.0f ? -126.0f : (y > 126.0f ? 126.0f : y);
    const float n = floorf(y + 0.5f);
    const float r = y - n;  // [-0.5, 0.5]
    float p = fmaf(0.0001540353f, r, 0.0013333558f);
    p = fmaf(p, r, 0.009618129f);
    p = fmaf(p, r, 0.05550411f);
    p = fmaf(p, r, 0.2402265f);
    p = fmaf(p, r, 0.6931472f);
    p = fmaf(p, r, 1.0f);
    return p * from_bits((uint32_t)(((int32_t)n + 127) << 23));
}

// image.frag:29  #define pow(a,b) pow(max(a,0.),b)
// x^n for a whole n >= 1: binary exponentiation, lowest bit first (exact FP32 products in a fixed order)
inline float det_powif(float a, unsigned n) {
    float r = 1.0f, p = a;
    while (n) {
        if (n & 1u) r = r * p;
        n >>= 1;
        if (n) p = p * p;
    }
    return r;
}
inline float gpow(float a, float b) {
    a = gmax(a, 0.0f);
    if (a != a) return a;
    if (a < 1.17549435e-38f) return 0.0f;
    if (a > 3.4028234e38f) return a;  // +inf
    if (b >= 1.0f && b <= 64.0f && b == floorf(b)) return det_powif(a, (unsigned)b);  // whole exponents: multiplication chain
    return det_exp2f(b * det_log2f(a));
}

struct V3 {
    float x, y, z;
};
inline float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float length3(V3 v) { return sqrtf(dot3(v, v)); }
inline V3 normalize3(V3 v) {
    const float inv = 1.0f / sqrtf((v.x * v.x + v.y * v.y) + v.z * v.z);
    return V3{v.x * inv, v.y * inv, v.z * inv};
}

struct Image {
    const uint8_t* rgba;
    int w, h;
};
inline int wrap(int i, int n) {
    const int m = i % n;
    return m < 0 ? m + n : m;
}
inline V3 texel(const Image& im, int x, int y) {
    const uint8_t* p = im.rgba + ((size_t)y * im.w + x) * 4;
    return V3{(float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f};
}
// texture(imageSampler, uv).rgb with the sampler of Pipeline.zig:193-212
inline V3 sample_linear_repeat(const Image& im, float u, float v) {
    const float x = u * (float)im.w - 0.5f, y = v * (float)im.h - 0.5f;
    const float fx = floorf(x), fy = floorf(y);
    const float a = x - fx, b = y - fy;
    const int x0 = wrap((int)fx, im.w), y0 = wrap((int)fy, im.h);
    const int x1 = x0 + 1 == im.w ? 0 : x0 + 1, y1 = y0 + 1 == im.h ? 0 : y0 + 1;
    const V3 t00 = texel(im, x0, y0), t10 = texel(im, x1, y0), t01 = texel(im, x0, y1), t11 = texel(im, x1, y1);
    const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
    return V3{((w00 * t00.x + w10 * t10.x) + w01 * t01.x) + w11 * t11.x, ((w00 * t00.y + w10 * t10.y) + w01 * t01.y) + w11 * t11.y,
              ((w00 * t00.z + w10 * t10.z) + w01 * t01.z) + w11 * t11.z};
}

inline uint8_t unorm8(float c) {
    if (!(c == c)) return 0;
    c = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);
    return (uint8_t)(c * 255.0f + 0.5f);
}

constexpr float kCosGolden = -0.7373688f, kSinGolden = 0.6754904f;  // cos / sin(2.3999632), image.frag:25,29

// image.frag:31-71 for the fragment at (ox, oy)
inline V3 sir_bird_denoise(const Image& im, const vrt_denoise_params& pc, float uvx, float uvy) {
    V3 denoised = V3{0.0f, 0.0f, 0.0f};                                     // :33
    const float sample_radius = sqrtf((float)pc.samples);                   // :35
    const float sample_true_radius = 0.5f / (sample_radius * sample_radius);  // :36
    const float sample_pixel_x = 1.0f / (float)im.w, sample_pixel_y = 1.0f / (float)im.h;  // :37
    const V3 center = sample_linear_repeat(im, uvx, uvy);                   // :38
    const V3 center_norm = normalize3(center);                              // :39
    const float center_sat = length3(center);                               // :40
    float influence_sum = 0.0f;                                             // :42
    float rot_x = 0.0f, rot_y = 1.0f;                                       // :45
    for (float x = 0.0f; x <= (float)pc.samples; x += 1.0f) {               // :47
        // :49  pixelRotated *= sample2D, sample2D = mat2(cos, sin, -sin, cos) (columns): v * M = (dot(v, col0), dot(v, col1))
        const float nx = rot_x * kCosGolden + rot_y * kSinGolden;
        const float ny = rot_x * (-kSinGolden) + rot_y * kCosGolden;
        rot_x = nx, rot_y = ny;
        const float sq = sqrtf(x);
        float off_x = ((pc.pixel_multiplier * rot_x) * sq) * 0.5f, off_y = ((pc.pixel_multiplier * rot_y) * sq) * 0.5f;  // :51
        float influence = 1.0f - sample_true_radius * gpow(off_x * off_x + off_y * off_y, pc.distribution_bias);      // :52
        off_x *= sample_pixel_x, off_y *= sample_pixel_y;                                                                // :53
        const V3 c = sample_linear_repeat(im, uvx + off_x, uvy + off_y);                                                 // :55
        influence *= influence * influence;                                                                              // :57
        // :61-64 hue + saturation filter; length(sampleCenterSat) of a float = abs
        influence *= gpow(0.5f + 0.5f * dot3(center_norm, normalize3(c)), pc.inverse_hue_tolerance) *
                     gpow(1.0f - fabsf(length3(c) - fabsf(center_sat)), 8.0f);
        influence_sum += influence;                                                                                      // :66
        denoised = V3{denoised.x + c.x * influence, denoised.y + c.y * influence, denoised.z + c.z * influence};         // :67
    }
    return V3{denoised.x / influence_sum, denoised.y / influence_sum, denoised.z / influence_sum};  // :70
}

}  // namespace

extern "C" int orc_denoise(const uint8_t* rgba8_in, uint32_t in_width, uint32_t in_height, const vrt_denoise_params* params, uint32_t out_width,
                           uint32_t out_height, uint32_t flags, uint8_t* out, int threads) {
    if (!rgba8_in || !params || !out || !in_width || !in_height || !out_width || !out_height || params->samples < 0) return -1;
    const Image im{rgba8_in, (int)in_width, (int)in_height};
    const bool bgra = (flags & VRT_DENOISE_BGRA) != 0u;
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if ((uint32_t)nt > out_height) nt = (int)out_height;
    auto rows = [&](uint32_t y0, uint32_t y1) {
        for (uint32_t y = y0; y < y1; y++)
            for (uint32_t x = 0; x < out_width; x++) {
                const float uvx = ((float)x + 0.5f) / (float)out_width, uvy = ((float)y + 0.5f) / (float)out_height;
                const V3 col = sir_bird_denoise(im, *params, uvx, uvy);  // image.frag:74-79
                uint8_t* p = out + ((size_t)y * out_width + x) * 4;
                p[bgra ? 2 : 0] = unorm8(col.x), p[1] = unorm8(col.y), p[bgra ? 0 : 2] = unorm8(col.z), p[3] = 255;
            }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; t++) pool.emplace_back(rows, (uint32_t)((uint64_t)out_height * t / nt), (uint32_t)((uint64_t)out_height * (t + 1) / nt));
    for (auto& th : pool) th.join();
    return 0;
}

extern "C" float orc_pow(float a, float b) { return gpow(a, b); }
