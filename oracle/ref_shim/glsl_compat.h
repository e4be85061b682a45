/*
 * glsl_compat.h — just enough of GLSL 4.50 for g++ to compile the reference's shaders AS THEY ARE.
 * TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * oracle/ref_shim/translate.py copies /root/reference/assets/shaders/{brick_raytracer,rand}.comp and image.frag into
 * oracle/_ref/ with a handful of purely lexical substitutions (listed in that script) and ref_harness.cpp includes the
 * result inside `namespace refshader`, after this header.  Every line of shader arithmetic that runs is the reference's own
 * text; this header only supplies what a Vulkan driver supplies: the vector types, the built-in functions and the resource
 * bindings.  Where the GLSL / Vulkan specifications leave a built-in's precision to the implementation, the choice made
 * here is the repo's FP discipline, the same the hand-written oracle and the CUDA kernels follow:
 *   - all arithmetic FP32, `fma()` is a fused multiply-add, nothing else is contracted (-ffp-contract=off);
 *   - `/`, sqrt: IEEE round-to-nearest;  inversesqrt is not used by the shaders;
 *   - dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z;   length(v) = sqrt(dot(v,v));
 *   - normalize(v) = v * (1.0f / sqrt(dot(v,v)));   reflect(I,N) = I - (2*dot(N,I))*N  (GLSL 8.5);
 *   - fract(x) = x - floor(x);  sign(±0) = 0;  min(x,y) = y<x ? y : x;  max(x,y) = x<y ? y : x  (GLSL 8.3);
 *   - sin / cos: det_sinf / det_cosf of det_math.h (GLSL promises 2^-11 absolute error only);
 *   - pow(x,y) = exp2(y*log2(x)) with det_log2f / det_exp2f of det_math.h (GLSL: "inherited from exp2(y*log2(x))");
 *   - float -> int conversion saturates, NaN -> 0 (GLSL: undefined out of range);
 *   - imageStore to Rgba8 / colour output to UNORM: clamp to [0,1], (uint8_t)(c*255 + 0.5f), NaN -> 0 (Vulkan 1.3 §3.11.1
 *     round-to-nearest);
 *   - storage-buffer reads outside the bound range return 0 (robustBufferAccess);
 *   - texture() on the linear / repeat sampler of Pipeline.zig:193-212: see sampler2D below.
 * This file must be included at global scope; it opens namespace refshader and leaves it OPEN (ref_harness.cpp closes it).
 */
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <type_traits>

#include "det_math.h"

#ifndef REFSHADER_NS
#define REFSHADER_NS refshader
#endif
namespace REFSHADER_NS {

typedef unsigned int uint;

template <class T>
using if_scalar = typename std::enable_if<std::is_arithmetic<T>::value, int>::type;

inline int f2i(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return -2147483647 - 1;
    return (int)f;
}

struct ivec2 {
    int x, y;
    ivec2() : x(0), y(0) {}
    template <class A, class B, if_scalar<A> = 0, if_scalar<B> = 0>
    ivec2(A a, B b) : x((int)a), y((int)b) {}
};
struct uvec3 {
    uint x, y, z;
};
struct bvec3 {
    bool x, y, z;
};
struct vec3;
struct ivec3 {
    int x, y, z;
    ivec3() : x(0), y(0), z(0) {}
    template <class A, if_scalar<A> = 0>
    explicit ivec3(A s) : x((int)s), y((int)s), z((int)s) {}
    template <class A, class B, class C, if_scalar<A> = 0, if_scalar<B> = 0, if_scalar<C> = 0>
    ivec3(A a, B b, C c) : x((int)a), y((int)b), z((int)c) {}
    explicit ivec3(const vec3& v);
};

struct vec2 {
    float x, y;
    vec3 xyx() const;
    vec2() : x(0.0f), y(0.0f) {}
    template <class A, if_scalar<A> = 0>
    explicit vec2(A s) : x((float)s), y((float)s) {}
    template <class A, class B, if_scalar<A> = 0, if_scalar<B> = 0>
    vec2(A a, B b) : x((float)a), y((float)b) {}
    vec2 xy() const { return *this; }
};

struct vec3 {
    float x, y, z;
    vec3() : x(0.0f), y(0.0f), z(0.0f) {}
    template <class A, if_scalar<A> = 0>
    explicit vec3(A s) : x((float)s), y((float)s), z((float)s) {}
    template <class A, class B, class C, if_scalar<A> = 0, if_scalar<B> = 0, if_scalar<C> = 0>
    vec3(A a, B b, C c) : x((float)a), y((float)b), z((float)c) {}
    vec3(const ivec3& i) : x((float)i.x), y((float)i.y), z((float)i.z) {}  // GLSL 4.1.10 implicit conversion ivec3 -> vec3
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    // swizzles (translate.py turns `.xyz` into `.xyz()`); r-values only, which is all the shaders use
    vec2 xy() const { return vec2(x, y); }
    vec2 xx() const { return vec2(x, x); }
    vec2 yz() const { return vec2(y, z); }
    vec2 zy() const { return vec2(z, y); }
    vec3 xyz() const { return *this; }
    vec3 rgb() const { return *this; }
    vec3 yzx() const { return vec3(y, z, x); }
    vec3 zyx() const { return vec3(z, y, x); }
    vec3 yxz() const { return vec3(y, x, z); }
    vec3 xxy() const { return vec3(x, x, y); }
    vec3 yzz() const { return vec3(y, z, z); }
};
inline vec3 vec2::xyx() const { return vec3(x, y, x); }

struct vec4 {
    float x, y, z, w;
    vec4() : x(0.0f), y(0.0f), z(0.0f), w(0.0f) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    template <class A, if_scalar<A> = 0>
    vec4(const vec3& v, A a) : x(v.x), y(v.y), z(v.z), w((float)a) {}
    vec3 xyz() const { return vec3(x, y, z); }
    vec3 rgb() const { return vec3(x, y, z); }
};

inline ivec3::ivec3(const vec3& v) : x(f2i(v.x)), y(f2i(v.y)), z(f2i(v.z)) {}

// column-major 2x2 matrix; mat2(a, b, c, d) has columns (a, b) and (c, d) (GLSL 5.4.2)
struct mat2 {
    float c0x, c0y, c1x, c1y;
    mat2(float a, float b, float c, float d) : c0x(a), c0y(b), c1x(c), c1y(d) {}
};

// ---- operators (component-wise, GLSL 5.9); a scalar operand of integer type converts to float first ----
inline vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator+(const vec3& a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(const vec3& a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3& operator+=(vec3& a, const vec3& b) { return a = a + b; }
inline vec3& operator+=(vec3& a, float s) { return a = a + s; }
inline vec2 operator+(const vec2& a, const vec2& b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator*(const vec2& a, const vec2& b) { return vec2(a.x * b.x, a.y * b.y); }
inline vec2 operator+(const vec2& a, float s) { return vec2(a.x + s, a.y + s); }
inline vec2 operator*(const vec2& a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator*(float s, const vec2& a) { return vec2(s * a.x, s * a.y); }
inline vec2& operator*=(vec2& a, const vec2& b) { return a = a * b; }
// v * m with v a row vector (GLSL 5.10): result.x = dot(v, column 0), result.y = dot(v, column 1)
inline vec2& operator*=(vec2& v, const mat2& m) { return v = vec2(v.x * m.c0x + v.y * m.c0y, v.x * m.c1x + v.y * m.c1y); }

// ---- built-in functions (GLSL 8.x); precision choices in the header comment ----
inline float fma(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
inline vec3 fma(const vec3& a, const vec3& b, const vec3& c) { return vec3(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y), fma(a.z, b.z, c.z)); }
inline float floor(float x) { return __builtin_floorf(x); }
inline vec3 floor(const vec3& v) { return vec3(floor(v.x), floor(v.y), floor(v.z)); }
inline float fract(float x) { return x - floor(x); }
inline vec2 fract(const vec2& v) { return vec2(fract(v.x), fract(v.y)); }
inline vec3 fract(const vec3& v) { return vec3(fract(v.x), fract(v.y), fract(v.z)); }
inline float abs(float x) { return __builtin_fabsf(x); }
inline vec3 abs(const vec3& v) { return vec3(abs(v.x), abs(v.y), abs(v.z)); }
inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline vec3 sign(const vec3& v) { return vec3(sign(v.x), sign(v.y), sign(v.z)); }
inline float min(float a, float b) { return b < a ? b : a; }
inline float max(float a, float b) { return a < b ? b : a; }
inline vec3 min(const vec3& a, const vec3& b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline vec3 max(const vec3& a, const vec3& b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline float sqrt(float x) { return __builtin_sqrtf(x); }
inline vec3 sqrt(const vec3& v) { return vec3(sqrt(v.x), sqrt(v.y), sqrt(v.z)); }
inline float sin(float x) { return detmath::det_sinf(x); }
inline float cos(float x) { return detmath::det_cosf(x); }
inline float pow(float a, float b) { return detmath::det_powf(a, b); }
inline float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
inline float dot(const vec3& a, const vec3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float length(float x) { return abs(x); }
inline float length(const vec3& v) { return sqrt(dot(v, v)); }
inline vec3 normalize(const vec3& v) { return v * (1.0f / sqrt(dot(v, v))); }
inline vec3 reflect(const vec3& i, const vec3& n) { return i - (2.0f * dot(n, i)) * n; }
inline bvec3 greaterThanEqual(const ivec3& a, const ivec3& b) { return bvec3{a.x >= b.x, a.y >= b.y, a.z >= b.z}; }
inline bvec3 lessThan(const ivec3& a, const ivec3& b) { return bvec3{a.x < b.x, a.y < b.y, a.z < b.z}; }
inline bool all(const bvec3& b) { return b.x && b.y && b.z; }
template <class T, class O>
inline T bitfieldExtract(T value, O offset, int bits) {  // unsigned genType (GLSL 8.8)
    if (bits == 0) return (T)0;
    const uint32_t v = (uint32_t)value >> (uint32_t)offset;
    return (T)(bits >= 32 ? v : (v & ((1u << bits) - 1u)));
}

// ---- resources ----
template <class T>
struct Buffer {  // a storage buffer block with one unsized array member
    const T* data = nullptr;
    uint64_t count = 0;
    T operator[](uint64_t i) const {
        if (i < count) return data[i];
        T zero;
        std::memset(&zero, 0, sizeof(T));
        return zero;
    }
};

inline uint8_t unorm8(float c) {
    if (!(c == c)) return 0;
    c = c < 0.0f ? 0.0f : (c > 1.0f ? 1.0f : c);
    return (uint8_t)(c * 255.0f + 0.5f);
}

struct image2D {  // layout(Rgba8) writeonly
    uint8_t* rgba8 = nullptr;
    int width = 0, height = 0;
};
inline ivec2 imageSize(const image2D& img) { return ivec2(img.width, img.height); }
inline void imageStore(const image2D& img, const ivec2& p, const vec4& c) {
    uint8_t* o = img.rgba8 + ((size_t)p.y * (size_t)img.width + (size_t)p.x) * 4;
    o[0] = unorm8(c.x), o[1] = unorm8(c.y), o[2] = unorm8(c.z), o[3] = unorm8(c.w);
}

// The sampler Pipeline.zig:193-212 creates for the compute image: linear min / mag filter, repeat addressing, R8G8B8A8_UNORM.
// texture(): unnormalised coordinate u*W - 0.5, i = floor, a = fraction, texels i and i+1 wrapped; UNORM decode byte/255.0f;
// FP32 weights, ((w00*t00 + w10*t10) + w01*t01) + w11*t11  (Vulkan 1.3 §16.8 leaves the weight precision open).
struct sampler2D {
    const uint8_t* rgba8 = nullptr;
    int width = 0, height = 0;
};
inline ivec2 textureSize(const sampler2D& s, int) { return ivec2(s.width, s.height); }
inline int wrap_repeat(int i, int n) {
    i %= n;
    return i < 0 ? i + n : i;
}
inline vec4 texture(const sampler2D& s, const vec2& uv) {
    const float fx = uv.x * (float)s.width - 0.5f, fy = uv.y * (float)s.height - 0.5f;
    const float flx = floor(fx), fly = floor(fy);
    const float a = fx - flx, b = fy - fly;
    const int i0 = wrap_repeat(f2i(flx), s.width), j0 = wrap_repeat(f2i(fly), s.height);
    const int i1 = wrap_repeat(i0 + 1, s.width), j1 = wrap_repeat(j0 + 1, s.height);
    const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
    float out[4];
    for (int c = 0; c < 4; c++) {
        const float t00 = (float)s.rgba8[((size_t)j0 * s.width + i0) * 4 + c] / 255.0f;
        const float t10 = (float)s.rgba8[((size_t)j0 * s.width + i1) * 4 + c] / 255.0f;
        const float t01 = (float)s.rgba8[((size_t)j1 * s.width + i0) * 4 + c] / 255.0f;
        const float t11 = (float)s.rgba8[((size_t)j1 * s.width + i1) * 4 + c] / 255.0f;
        out[c] = ((w00 * t00 + w10 * t10) + w01 * t01) + w11 * t11;
    }
    return vec4(out[0], out[1], out[2], out[3]);
}

// namespace refshader stays open: the translated shader text follows.
