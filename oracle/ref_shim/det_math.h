/*
 * det_math.h — the transcendental built-ins the Vulkan driver would supply (sin, cos, pow), in the repo's FP discipline.
 * TEST INFRASTRUCTURE, NOT PRODUCT.  GLSL 4.50 §4.7.1 only bounds their error (sin / cos: 2^-11 absolute; pow: "inherited from
 * exp2(y * log2(x))"), so any faithful implementation is a conforming driver; these are the ones oracle/vrt_oracle*.cpp and the
 * CUDA kernels use (explicit fmaf chains, bit-identical on x86 and sm_100a), repeated here so that oracle/_ref depends on
 * nothing but the reference's shader text and this directory.
 */
#pragma once
#include <cstdint>
#include <cstring>

namespace detmath {

inline float sincos_core(float x, int quadrant_offset) {
    const float kf = __builtin_rintf(x * 0.636619747f);  // x * 2/pi, round half to even
    const int k = (int)kf + quadrant_offset;
    float r = __builtin_fmaf(kf, -1.57079601e+00f, x);  // three-term Cody-Waite pi/2
    r = __builtin_fmaf(kf, -3.13916473e-07f, r);
    r = __builtin_fmaf(kf, -5.39030253e-15f, r);
    const float s = r * r;
    float res;
    if (k & 1) {  // cosine polynomial on [-pi/4, pi/4]
        float p = 2.44677067e-5f;
        p = __builtin_fmaf(p, s, -1.38877297e-3f);
        p = __builtin_fmaf(p, s, 4.16666567e-2f);
        p = __builtin_fmaf(p, s, -5.00000000e-1f);
        res = __builtin_fmaf(p, s, 1.0f);
    } else {  // sine polynomial on [-pi/4, pi/4]
        float p = 2.86567956e-6f;
        p = __builtin_fmaf(p, s, -1.98559923e-4f);
        p = __builtin_fmaf(p, s, 8.33338592e-3f);
        p = __builtin_fmaf(p, s, -1.66666672e-1f);
        const float t = r * s;
        res = __builtin_fmaf(p, t, r);
    }
    return (k & 2) ? -res : res;
}
inline float det_sinf(float x) { return sincos_core(x, 0); }
inline float det_cosf(float x) { return sincos_core(x, 1); }  // cos(x) = sin(x + pi/2)

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
// log2 of a finite a >= FLT_MIN: a = 2^e * m with m in [sqrt(1/2), sqrt(2)); ln m = 2 atanh(s), s = (m-1)/(m+1)
inline float det_log2f(float a) {
    const int32_t ia = (int32_t)to_bits(a);
    const int32_t e = (ia - 0x3f3504f3) >> 23;
    const float m = from_bits((uint32_t)(ia - e * (1 << 23)));
    const float f = m - 1.0f;
    const float s = f / (2.0f + f);
    const float z = s * s;
    float p = __builtin_fmaf(z, 0.22222222f, 0.2857143f);
    p = __builtin_fmaf(p, z, 0.4f);
    p = __builtin_fmaf(p, z, 0.6666667f);
    const float ln_m = __builtin_fmaf(s * z, p, s + s);
    return __builtin_fmaf(ln_m, 1.442695f, (float)e);
}
// 2^y, y clamped to [-126, 126]; NaN stays NaN
inline float det_exp2f(float y) {
    if (y != y) return y;
    y = y < -126.0f ? -126.0f : (y > 126.0f ? 126.0f : y);
    const float n = __builtin_floorf(y + 0.5f);
    const float r = y - n;
    float p = __builtin_fmaf(0.0001540353f, r, 0.0013333558f);
    p = __builtin_fmaf(p, r, 0.009618129f);
    p = __builtin_fmaf(p, r, 0.05550411f);
    p = __builtin_fmaf(p, r, 0.2402265f);
    p = __builtin_fmaf(p, r, 0.6931472f);
    p = __builtin_fmaf(p, r, 1.0f);
    return p * from_bits((uint32_t)(((int32_t)n + 127) << 23));
}
// x^n for a whole n >= 1 by binary exponentiation, lowest bit first: exact FP32 products in a fixed order
inline float det_powif(float a, unsigned n) {
    float r = 1.0f, p = a;
    while (n) {
        if (n & 1u) r = r * p;
        n >>= 1;
        if (n) p = p * p;
    }
    return r;
}
// pow(x, y) for x >= 0 (image.frag:29 clamps the base itself): pow(0, y) = 0, bases below FLT_MIN count as 0; a whole exponent
// 1 <= y <= 64 (image.frag's literal 8, the default hue tolerance 20) is a chain of multiplications, as a shader compiler folds
// pow(x, 8.) — more accurate than exp2(y * log2(x)) and well inside what GLSL lets a driver do
inline float det_powf(float a, float b) {
    if (a != a) return a;
    if (a < 1.17549435e-38f) return 0.0f;
    if (a > 3.4028234e38f) return a;
    if (b >= 1.0f && b <= 64.0f && b == __builtin_floorf(b)) return det_powif(a, (unsigned)b);
    return det_exp2f(b * det_log2f(a));
}

}  // namespace detmath
