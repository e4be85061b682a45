/*
 * vrt_oracle.cpp — CPU oracle for the voxel ray-tracing hot path.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A function-by-function FP32 restatement of the reference's compute shader
 *   assets/shaders/brick_raytracer.comp   (main, RayColor, GridHit, BrickHit, slab test, scatter fns)
 *   assets/shaders/rand.comp              (Rand, RandVec3, hash12)
 * and of the host code that builds the shader's input buffers
 *   src/modules/voxel_rt/brick/Grid.zig, brick/MaterialAllocator.zig.
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 *
 * PARITY PINNED — see vrt_oracle.h: the reference has no CPU path, no tests and no golden output for this code, but its
 * shader text compiles under oracle/ref_shim/ (oracle/_ref/libref_shader.so) and tests/test_ref_shader.py holds this file to
 * it bit for bit, on top of the hand-computed known-answer tests (tests/test_oracle_kat.py).
 *
 * FP discipline (GLSL leaves these implementation-defined; the CUDA kernels are held to the same choices):
 *   - FP32 everywhere; literals are float (GLSL literals are FP32).
 *   - only calls the shader spells `fma(...)` are fused (fmaf); NO other contraction: build with
 *     -ffp-contract=off (nvcc side: --fmad=false).
 *   - `/` and sqrt are IEEE round-to-nearest.
 *   - normalize(v) = v * (1.0f / sqrtf((x*x + y*y) + z*z));  dot(a,b) = (a.x*b.x + a.y*b.y) + a.z*b.z
 *   - fract(x) = x - floorf(x);  sign(0) = 0;  min(a,b) = b<a ? b : a;  max(a,b) = a<b ? b : a
 *   - sin(x) = orc_sinf(x) below: Cody-Waite reduction + minimax polynomials written with explicit fmaf, so
 *     it is bit-identical on x86 and on the GPU (libm sinf and CUDA sinf are not).  GLSL only promises
 *     2^-11 absolute error for sin, so any faithful-ish sine is a conforming implementation of the shader.
 *   - Rgba8 imageStore: clamp to [0,1] then (uint8_t)(c*255.0f + 0.5f); NaN stores 0.
 *   - float -> int conversion of floor(fposition) saturates and maps NaN to 0 (the CUDA cvt.rzi behaviour), so the
 *     two sides agree even where the C++ cast would be undefined.
 */
#include "vrt_oracle.h"

#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <thread>
#include <vector>

namespace {

// --------------------------------------------------------------------------------------------------
// GLSL-flavoured FP32 helpers
// --------------------------------------------------------------------------------------------------
struct V3 {
    float x, y, z;
};
struct V2 {
    float x, y;
};
struct I3 {
    int x, y, z;
};

inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 v3s(float s) { return V3{s, s, s}; }
inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator/(V3 a, V3 b) { return V3{a.x / b.x, a.y / b.y, a.z / b.z}; }
inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
inline V3 neg(V3 a) { return V3{-a.x, -a.y, -a.z}; }
inline V3 fma3(V3 a, V3 b, V3 c) { return V3{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z)}; }
inline V3 tofloat(I3 i) { return V3{(float)i.x, (float)i.y, (float)i.z}; }
// ivec3(floor(f)) with the saturating / NaN -> 0 semantics of the GPU's float->int conversion
inline int f2i(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return -2147483647 - 1;
    return (int)f;
}
inline float gmin(float a, float b) { return b < a ? b : a; }
inline float gmax(float a, float b) { return a < b ? b : a; }
inline float gsign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline float fract(float x) { return x - floorf(x); }
inline float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot2(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
inline V3 normalize3(V3 v) {
    const float inv = 1.0f / sqrtf((v.x * v.x + v.y * v.y) + v.z * v.z);
    return v * inv;
}
// GLSL reflect(I, N) = I - 2.0 * dot(N, I) * N
inline V3 reflect3(V3 i, V3 n) { return i - (2.0f * dot3(n, i)) * n; }
inline float comp(V3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

// --------------------------------------------------------------------------------------------------
// Deterministic FP32 sine (see header comment).  |x| up to ~1e5 keeps full accuracy of the reduction.
// --------------------------------------------------------------------------------------------------
inline float det_sinf(float x) {
    const float kf = rintf(x * 0.636619747f);  // x * 2/pi, round half to even
    const int k = (int)kf;
    float r = fmaf(kf, -1.57079601e+00f, x);  // three-term Cody-Waite pi/2
    r = fmaf(kf, -3.13916473e-07f, r);
    r = fmaf(kf, -5.39030253e-15f, r);
    const float s = r * r;
    float res;
    if (k & 1) {  // cosine polynomial on [-pi/4, pi/4]
        float p = 2.44677067e-5f;
        p = fmaf(p, s, -1.38877297e-3f);
        p = fmaf(p, s, 4.16666567e-2f);
        p = fmaf(p, s, -5.00000000e-1f);
        res = fmaf(p, s, 1.0f);
    } else {  // sine polynomial on [-pi/4, pi/4]
        float p = 2.86567956e-6f;
        p = fmaf(p, s, -1.98559923e-4f);
        p = fmaf(p, s, 8.33338592e-3f);
        p = fmaf(p, s, -1.66666672e-1f);
        const float t = r * s;
        res = fmaf(p, t, r);
    }
    return (k & 2) ? -res : res;
}

// --------------------------------------------------------------------------------------------------
// rand.comp
// --------------------------------------------------------------------------------------------------
// rand.comp:3   float Rand(float co) { return fract(sin(co*(91.3458)) * 47453.5453); }
inline float Rand1(float co) { return fract(det_sinf(co * 91.3458f) * 47453.5453f); }
// rand.comp:4   float Rand(vec2 co){ return fract(sin(dot(co.xy, vec2(12.9898,78.233))) * 43758.5453); }
inline float Rand2(V2 co) { return fract(det_sinf(dot2(co, V2{12.9898f, 78.233f})) * 43758.5453f); }
// rand.comp:5   float Rand(vec3 co){ return Rand(co.xy+Rand(co.z)); }
inline float Rand3(V3 co) {
    const float r = Rand1(co.z);
    return Rand2(V2{co.x + r, co.y + r});
}
// rand.comp:6-8
inline float Rand2mm(V2 co, float mn, float mx) { return mn + (mx - mn) * Rand2(co); }
// rand.comp:15-20
inline V3 RandVec3mm(V2 co, float mn, float mx) {
    const float x = Rand2mm(co, mn, mx);
    const float y = Rand2mm(V2{co.x + x, co.y + x}, mn, mx);
    const float z = Rand2mm(V2{co.x + y, co.y + y}, mn, mx);
    return v3(x, y, z);
}
// rand.comp:22-26
inline float hash12(V2 p) {
    V3 p3 = v3(fract(p.x * .1031f), fract(p.y * .1031f), fract(p.x * .1031f));
    const float d = dot3(p3, v3(p3.y + 33.33f, p3.z + 33.33f, p3.x + 33.33f));
    p3 = v3(p3.x + d, p3.y + d, p3.z + d);
    return fract((p3.x + p3.y) * p3.z);
}

// --------------------------------------------------------------------------------------------------
// shader types
// --------------------------------------------------------------------------------------------------
constexpr uint32_t MAT_LAMBERTIAN = 0, MAT_METAL = 1, MAT_DIELECTRIC = 2, MAT_NONE = 3;

struct Ray {  // brick_raytracer.comp:36-41
    V3 origin;
    V3 direction;
    float internal_reflection;
    uint32_t ignore_type_material;
};
struct HitRecord {  // brick_raytracer.comp:46-51
    V3 point;
    V3 normal;
    float t;
    uint32_t index;
};

// what the oracle records about one GridHit call, beyond the shader's outputs
struct Trace {
    uint32_t grid_index = ~0u;
    uint32_t voxel_index = ~0u;
    uint32_t grid_steps = 0;
    uint32_t voxel_steps = 0;
    uint32_t status_fetches = 0;
    uint32_t bricks_entered = 0;
};

struct Ctx {
    const orc_scene* sc;
    const vrt_camera* cam;
    const vrt_sun* sun;
    V3 g_min, g_max;
    float g_scale;
    I3 brick_dim;
    int bd;                   // spec const `brick_dimensions`
    uint32_t brick_bytes;     // spec const `brick_bytes` = bd^3/8
    float brick_voxel_scale;  // spec const: 1.0f / bd, computed on the host in f32 (Pipeline.zig:313)
};

// brick_raytracer.comp:180-184
inline Ray CreateRay(V3 origin, V3 direction) { return Ray{origin, normalize3(direction), 1.0f, MAT_NONE}; }
// brick_raytracer.comp:186-190
inline Ray CreateShadowRay(const Ctx& c, V3 origin, V3 direction) {
    const uint32_t ignore = (c.sun->enabled > 0) ? MAT_NONE : MAT_DIELECTRIC;
    return Ray{origin, normalize3(direction), 1.0f, ignore};
}
// brick_raytracer.comp:192-195
inline V3 RayAt(const Ray& r, float t) { return fma3(v3s(t), r.direction, r.origin); }
// brick_raytracer.comp:197-201
inline V3 BackgroundColor(const Ray& r) {
    const float t = 0.5f * (r.direction.y + 1.0f);
    return fma3(v3s(1.0f - t), v3s(1.0f), t * v3(0.5f, 0.7f, 1.0f));
}
// brick_raytracer.comp:267-268
inline float safeInverse(float x) { return (x == 0.0f) ? 1e12f : (1.0f / x); }
// brick_raytracer.comp:497-503
inline float minComponent(V3 v) { return gmin(gmin(v.x, v.y), v.z); }
inline int indexOfMaxComponent(V3 v) { return int(v.y > v.x && v.y > v.z) + int(v.z > v.x && v.z > v.y) * 2; }

// brick_raytracer.comp:522-536
inline bool AdvNormIntersect(V3 box_min, V3 box_max, const Ray& r, V3 inv, V3& normal, float& t_min, float& t_max) {
    const V3 t_lower = (box_min - r.origin) * inv;
    const V3 t_upper = (box_max - r.origin) * inv;
    const V3 t_mins = v3(gmin(t_lower.x, t_upper.x), gmin(t_lower.y, t_upper.y), gmin(t_lower.z, t_upper.z));
    const V3 t_maxes = v3(gmax(t_lower.x, t_upper.x), gmax(t_lower.y, t_upper.y), gmax(t_lower.z, t_upper.z));
    const int idx = indexOfMaxComponent(t_mins);
    normal = v3s(0.0f);
    const float s = gsign(comp(inv, idx));
    if (idx == 0) normal.x = s;
    else if (idx == 1) normal.y = s;
    else normal.z = s;
    t_min = gmax(t_min, comp(t_mins, idx));
    t_max = gmin(t_max, minComponent(t_maxes));
    return t_min <= t_max;
}

// The DDA advance shared verbatim by GridHit (:345-372) and BrickHit (:440-467): strict `<`, tie order
// x -> z / y -> z, t_value read BEFORE the increment.
inline void dda_step(V3& side_dist, V3 ray_delta, I3& pos, I3 ray_step, float scale, V3 normal_axis, float& t_value,
                     V3& normal) {
    if (side_dist.x < side_dist.y) {
        if (side_dist.x < side_dist.z) {
            t_value = side_dist.x * scale;
            side_dist.x += ray_delta.x;
            pos.x += ray_step.x;
            normal = v3(normal_axis.x, 0, 0);
        } else {
            t_value = side_dist.z * scale;
            side_dist.z += ray_delta.z;
            pos.z += ray_step.z;
            normal = v3(0, 0, normal_axis.z);
        }
    } else {
        if (side_dist.y < side_dist.z) {
            t_value = side_dist.y * scale;
            side_dist.y += ray_delta.y;
            pos.y += ray_step.y;
            normal = v3(0, normal_axis.y, 0);
        } else {
            t_value = side_dist.z * scale;
            side_dist.z += ray_delta.z;
            pos.z += ray_step.z;
            normal = v3(0, 0, normal_axis.z);
        }
    }
}

// brick_raytracer.comp:378-471
bool BrickHit(const Ctx& c, const Ray& r, float /*t_min*/, float t_max, V3 ray_delta, I3 ray_step, float g_scale,
              uint32_t brick_index, V3& brick_position, HitRecord& hit, Trace& tr) {
    const orc_scene& sc = *c.sc;
    const float voxel_scale = g_scale * c.brick_voxel_scale;                      // :389
    const uint64_t solid_mask_base_index = (uint64_t)brick_index * c.brick_bytes;  // :390

    const V3 fposition = (RayAt(r, hit.t) - brick_position) / v3s(voxel_scale);  // :393
    const V3 intersection_delta = v3(floorf(fposition.x), floorf(fposition.y), floorf(fposition.z)) - fposition;
    const V3 fstep = tofloat(ray_step);
    V3 side_dist = fma3(fstep, intersection_delta, fstep * 0.5f + v3s(0.5f)) * ray_delta;  // :395

    const V3 normal_axis = v3(ray_step.x < 0 ? 1.f : -1.f, ray_step.y < 0 ? 1.f : -1.f, ray_step.z < 0 ? 1.f : -1.f);

    I3 pos = I3{f2i(floorf(fposition.x)), f2i(floorf(fposition.y)), f2i(floorf(fposition.z))};  // :403
    const float local_t_max = t_max - hit.t;                                                   // :405
    float t_value = 0;
    const int bd = c.bd;
    while (pos.x >= 0 && pos.y >= 0 && pos.z >= 0 && pos.x < bd && pos.y < bd && pos.z < bd && t_value <= local_t_max) {
        tr.voxel_steps++;
        const int voxel_index = pos.x + bd * (pos.z + bd * pos.y);  // :412
        // :413 the shader truncates the byte index to uint8_t; exact for bd <= 8 (bd^3/8 <= 64).  For the
        // bd = 16 extension (512 mask bytes per brick) the index is kept in 32 bits — documented deviation.
        const uint32_t mask_index = (bd <= 8) ? (uint32_t)(uint8_t)(voxel_index / 8) : (uint32_t)(voxel_index / 8);
        const uint32_t mask_offset = (uint32_t)(voxel_index % 8);
        const uint64_t mask_at = solid_mask_base_index + mask_index;
        const uint8_t entry = mask_at < sc.n_occupancy ? sc.occupancy[mask_at] : (uint8_t)0;  // :415
        const bool hit_voxel = ((entry >> mask_offset) & 1u) != 0;                               // :417
        if (hit_voxel) {
            bool ignore_brick = false;
            {
                const uint32_t sw = brick_index < sc.n_start_indices ? sc.start_indices[brick_index] : 0u;
                const uint32_t brick_material_index = sw & 0x7fffffffu;  // :422
                const uint64_t mi = (uint64_t)brick_material_index + (uint32_t)voxel_index;
                hit.index = mi < sc.n_material_indices ? sc.material_indices[mi] : 0u;  // :425
                const vrt_material& m = sc.materials[hit.index < sc.n_materials ? hit.index : 0u];
                ignore_brick = (m.type == r.ignore_type_material) && (r.internal_reflection == m.type_data);  // :427
            }
            if (!ignore_brick) {
                const float t_offset = voxel_scale * 0.05f;          // :431
                hit.t += t_value - t_offset;                         // :432
                hit.point = RayAt(r, hit.t) + hit.normal * t_offset;  // :433
                brick_position = tofloat(pos) * voxel_scale + brick_position;  // :434
                tr.voxel_index = (uint32_t)voxel_index;
                return true;
            }
        }
        dda_step(side_dist, ray_delta, pos, ray_step, voxel_scale, normal_axis, t_value, hit.normal);  // :440-467
    }
    return false;
}

// brick_raytracer.comp:271-376
bool GridHit(const Ctx& c, const Ray& r, float t_min, float t_max, V3& hit_min, HitRecord& hit, Trace& tr) {
    const orc_scene& sc = *c.sc;
    const V3 g_min = c.g_min;
    const V3 g_max = c.g_max;
    const float g_scale = c.g_scale;
    const I3 brick_dim = c.brick_dim;

    // Deviation (DESIGN.md "Deviations"): a ray whose direction is not finite (normalize of a zero or NaN vector, e.g. the
    // 0/0 of a 1-pixel-wide image at :168) has ray_step = 0 on every axis; the shader's loop would then never advance
    // (undefined behaviour in GLSL, a hang in practice).  Such a ray is a miss here and in the CUDA kernels.
    if (std::isnan((r.direction.x + r.direction.y) + r.direction.z)) return false;

    const V3 inv_ray_dir = v3(safeInverse(r.direction.x), safeInverse(r.direction.y), safeInverse(r.direction.z));  // :278

    float grid_t_min = t_min;
    float grid_t_max = t_max;
    if (!AdvNormIntersect(g_min, g_max, r, inv_ray_dir, hit.normal, grid_t_min, grid_t_max)) return false;  // :282

    float global_t_value = grid_t_min + 0.0001f * g_scale;  // :287

    const V3 ray_delta = v3(fabsf(inv_ray_dir.x), fabsf(inv_ray_dir.y), fabsf(inv_ray_dir.z));              // :290
    const I3 ray_step = I3{(int)gsign(r.direction.x), (int)gsign(r.direction.y), (int)gsign(r.direction.z)};  // :291

    const V3 hit_point = RayAt(r, global_t_value);  // :293

    const V3 fposition = (hit_point - g_min) / v3s(g_scale);  // :296
    const V3 intersection_delta = v3(floorf(fposition.x), floorf(fposition.y), floorf(fposition.z)) - fposition;
    const V3 fstep = tofloat(ray_step);
    V3 side_dist = fma3(fstep, intersection_delta, fstep * 0.5f + v3s(0.5f)) * ray_delta;  // :298

    uint32_t brick_type_index = ~0u;  // :301
    uint32_t brick_bits = 0;

    const V3 normal_axis = v3(ray_step.x < 0 ? 1.f : -1.f, ray_step.y < 0 ? 1.f : -1.f, ray_step.z < 0 ? 1.f : -1.f);

    float t_value = 0;
    I3 pos = I3{f2i(floorf(fposition.x)), f2i(floorf(fposition.y)), f2i(floorf(fposition.z))};  // :311

    while (pos.x >= 0 && pos.y >= 0 && pos.z >= 0 && pos.x < brick_dim.x && pos.y < brick_dim.y && pos.z < brick_dim.z &&
           global_t_value <= t_max) {  // :313-317 (the PARAMETER t_max = +inf, not grid_t_max)
        tr.grid_steps++;
        const uint32_t grid_index = (uint32_t)(pos.x + brick_dim.x * (pos.z + brick_dim.z * pos.y));  // :318

        const uint32_t new_brick_type_index = grid_index / 32;  // :321
        const int brick_type_offset = (int)(grid_index % 32);
        if (brick_type_index != new_brick_type_index) {
            brick_bits = new_brick_type_index < sc.n_statuses ? sc.statuses[new_brick_type_index] : 0u;  // :324
            brick_type_index = new_brick_type_index;
            tr.status_fetches++;
        }

        const uint32_t entry_type = brick_bits & (1u << brick_type_offset);  // :328
        if (entry_type != 0) {
            V3 brick_min = fma3(tofloat(pos), v3s(g_scale), g_min);     // :331
            global_t_value = (t_value + grid_t_min) + 0.01f * g_scale;  // :332
            hit.t = global_t_value;                                     // :334

            const uint32_t brick_index = grid_index < sc.n_brick_indices ? sc.brick_indices[grid_index] : 0u;  // :337
            tr.bricks_entered++;
            if (BrickHit(c, r, t_min, grid_t_max, ray_delta, ray_step, g_scale, brick_index, brick_min, hit, tr)) {
                hit_min = brick_min;
                tr.grid_index = grid_index;
                return true;
            }
        }
        dda_step(side_dist, ray_delta, pos, ray_step, g_scale, normal_axis, t_value, hit.normal);  // :345-372
    }
    return false;
}

// brick_raytracer.comp:539-544
inline bool ScatterLambertian(const HitRecord& hit, Ray& scattered) {
    const V2 co = V2{hit.point.x + hit.point.z, hit.point.y + hit.point.z};
    const V3 scatter_dir = normalize3(hit.normal + RandVec3mm(co, -0.4f, 0.4f));
    scattered = CreateRay(hit.point, scatter_dir);
    return true;
}
// brick_raytracer.comp:546-551
inline bool ScatterMetal(const vrt_material& m, const Ray& r_in, const HitRecord& hit, Ray& scattered) {
    const V3 reflected = reflect3(r_in.direction, hit.normal);
    const float fuzz = m.type_data;
    const V2 co = V2{hit.point.x + hit.point.z, hit.point.y + hit.point.z};
    scattered = CreateRay(hit.point, reflected + RandVec3mm(co, -fuzz, fuzz));
    return dot3(scattered.direction, hit.normal) > 0;
}
// brick_raytracer.comp:564-574
inline bool transmissionDirection(float n1, float n2, V3 ray_dir, V3 normal, V3& refrac_dir) {
    const float eta = n1 / n2;
    const float c1 = -dot3(ray_dir, normal);
    const float w = eta * c1;
    const float c2m = (w - eta) * (w + eta);
    if (c2m < -1.0f) return false;
    refrac_dir = fma3(v3s(eta), ray_dir, (w - sqrtf(1.0f + c2m)) * normal);
    return true;
}
// brick_raytracer.comp:576-596
inline bool ScatterDielectric(const vrt_material& m, const Ray& r_in, const HitRecord& hit, Ray& scattered) {
    const float ir = m.type_data;
    const V2 co = V2{hit.point.x + hit.point.z, hit.point.y + hit.point.z};
    const V3 normal = normalize3(hit.normal + RandVec3mm(co, -0.05f, 0.05f));
    V3 direction = v3s(0.0f);
    const bool should_refract = transmissionDirection(ir, r_in.internal_reflection, r_in.direction, normal, direction);
    if (should_refract && Rand3(hit.point) > 0.5f) {
        scattered = CreateRay(hit.point, direction);
        scattered.ignore_type_material = MAT_DIELECTRIC;
        scattered.internal_reflection = ir;
    } else {
        direction = reflect3(r_in.direction, normal);
        scattered = CreateRay(hit.point, direction);
    }
    return true;
}

struct PixelAcc {  // per-pixel bookkeeping for AOV + counters
    vrt_aov* aov;  // nullable
    vrt_counters cnt;
};

inline void account(PixelAcc& acc, const Trace& tr, bool hit, bool shadow) {
    acc.cnt.rays++;
    acc.cnt.grid_steps += tr.grid_steps;
    acc.cnt.voxel_steps += tr.voxel_steps;
    acc.cnt.status_fetches += tr.status_fetches;
    acc.cnt.bricks_entered += tr.bricks_entered;
    if (hit) acc.cnt.hits++;
    if (shadow) acc.cnt.shadow_rays++;
}

// brick_raytracer.comp:203-265
V3 RayColor(const Ctx& c, const Ray& r, PixelAcc& acc, bool record_aov) {
    const orc_scene& sc = *c.sc;
    const bool sun_enabled = c.sun->enabled > 0;
    const V3 sun_position = v3(c.sun->position[0], c.sun->position[1], c.sun->position[2]);
    const V3 sun_color = v3(c.sun->color[0], c.sun->color[1], c.sun->color[2]);
    const float sun_radius = c.sun->radius;
    const float infinity = std::numeric_limits<float>::infinity();  // :28

    HitRecord hit{};
    Ray current_ray = r;
    int loop_count = 0;
    HitRecord shadow_hit{};
    V3 color = v3s(0.0f);
    V3 hit_v_min = v3s(0.0f);

    for (int iter = 0;; iter++) {  // iter counts loop trips; loop_count can be decremented (:236)
        if (!(loop_count < c.cam->max_bounce)) break;  // :218 (left operand of &&)
        Trace tr;
        const bool got = GridHit(c, current_ray, 0.00001f, infinity, hit_v_min, hit, tr);
        account(acc, tr, got, false);
        const bool first = record_aov && acc.aov && iter == 0;  // AOV = sample 0, primary ray
        if (first) {
            vrt_aov& a = *acc.aov;
            a.grid_steps += tr.grid_steps;
            a.voxel_steps += tr.voxel_steps;
            a.status_fetches += tr.status_fetches;
            if (got) {
                a.flags |= VRT_AOV_HIT;
                a.grid_index = tr.grid_index;
                a.voxel_index = tr.voxel_index;
                a.material = hit.index;
                a.t = hit.t;
                a.point[0] = hit.point.x, a.point[1] = hit.point.y, a.point[2] = hit.point.z;
                a.normal[0] = hit.normal.x, a.normal[1] = hit.normal.y, a.normal[2] = hit.normal.z;
            }
        }
        if (!got) break;
        if (iter == 0) acc.cnt.primary_hits++;

        loop_count += 1;  // :219
        Ray scattered = current_ray;
        bool result = false;

        const vrt_material material = sc.materials[hit.index < sc.n_materials ? hit.index : 0u];  // :223
        const V3 attenuation = v3(material.albedo_r, material.albedo_g, material.albedo_b);
        switch (material.type) {  // :225-239
            case MAT_LAMBERTIAN: result = ScatterLambertian(hit, scattered); break;
            case MAT_METAL: result = ScatterMetal(material, current_ray, hit, scattered); break;
            case MAT_DIELECTRIC: result = ScatterDielectric(material, current_ray, hit, scattered); break;
            default:
                loop_count -= 1;
                result = false;
                break;
        }
        if (sun_enabled) {  // :240-249
            const V2 co = V2{current_ray.direction.x + current_ray.direction.z, current_ray.direction.y + current_ray.direction.z};
            const V3 sun_sample_position = sun_position + RandVec3mm(co, -sun_radius, sun_radius);
            const V3 shadow_ray_dir = sun_sample_position - hit.point;
            const Ray shadow_ray = CreateShadowRay(c, hit.point, shadow_ray_dir);
            Trace str;
            V3 shadow_min = v3s(0.0f);
            const bool blocked = GridHit(c, shadow_ray, 0.00001f, infinity, shadow_min, shadow_hit, str);
            account(acc, str, blocked, true);
            if (first) {
                vrt_aov& a = *acc.aov;
                a.flags |= VRT_AOV_SHADOW_CAST;
                a.grid_steps += str.grid_steps;
                a.voxel_steps += str.voxel_steps;
                a.status_fetches += str.status_fetches;
                if (blocked) {
                    a.flags |= VRT_AOV_SHADOW_BLOCKED;
                    a.shadow_grid_index = str.grid_index;
                    a.shadow_voxel_index = str.voxel_index;
                }
            }
            if (!blocked) color = color + attenuation * sun_color;  // :248
        } else {
            color = color + attenuation;  // :251
        }
        if (!result) break;  // :255
        current_ray = scattered;
    }

    if (loop_count == 0) {  // :260-262
        color = color + BackgroundColor(current_ray) * (sun_enabled ? sun_color : v3s(1.0f));
    }
    return color / (color + v3s(1.0f));  // :264
}

// brick_raytracer.comp:474-477
inline Ray CameraGetRay(const Ctx& c, float u, float v) {
    const vrt_camera& cam = *c.cam;
    const V3 horizontal = v3(cam.horizontal[0], cam.horizontal[1], cam.horizontal[2]);
    const V3 vertical = v3(cam.vertical[0], cam.vertical[1], cam.vertical[2]);
    const V3 llc = v3(cam.lower_left_corner[0], cam.lower_left_corner[1], cam.lower_left_corner[2]);
    const V3 origin = v3(cam.origin[0], cam.origin[1], cam.origin[2]);
    const V3 ray_dir = fma3(horizontal, v3s(u), llc) + fma3(v3s(v), vertical, neg(origin));
    return CreateRay(origin, ray_dir);
}

inline uint8_t unorm8(float c) {
    if (!(c > 0.0f)) return 0;  // also NaN
    if (c > 1.0f) c = 1.0f;
    return (uint8_t)(c * 255.0f + 0.5f);
}

// brick_raytracer.comp:153-178
void shade_pixel(const Ctx& c, uint32_t px, uint32_t py, uint8_t* rgba8, vrt_aov* aov, vrt_counters& totals) {
    const vrt_camera& cam = *c.cam;
    PixelAcc acc{};
    acc.aov = aov;
    if (aov) {
        std::memset(aov, 0, sizeof(*aov));
        aov->grid_index = aov->voxel_index = aov->material = ~0u;
        aov->shadow_grid_index = aov->shadow_voxel_index = ~0u;
    }
    V3 color = v3s(0.0f);
    for (int sample_i = 0; sample_i < cam.samples_per_pixel; sample_i++) {
        const float x = (float)px;
        const float y = (float)py;
        const float flag = (float)(sample_i > 0);
        const float noise_x = hash12(V2{((x + (float)sample_i) * 0.2f) * flag, (y * 0.2f) * flag});  // :167
        const float u = (x + noise_x) / (float)(cam.image_width - 1u);                                // :168
        const float noise_y = hash12(V2{(x * 0.2f) * flag, ((y + (float)sample_i) * 0.2f) * flag});  // :169
        const float v = (y + noise_y) / (float)(cam.image_height - 1u);                               // :170
        const Ray ray = CameraGetRay(c, u, v);
        color = color + RayColor(c, ray, acc, sample_i == 0);
    }
    const float spp = (float)cam.samples_per_pixel;
    color = v3(sqrtf(color.x / spp), sqrtf(color.y / spp), sqrtf(color.z / spp));  // :176
    uint8_t* o = rgba8 + 4ull * ((uint64_t)py * cam.image_width + px);
    o[0] = unorm8(color.x), o[1] = unorm8(color.y), o[2] = unorm8(color.z), o[3] = 255;  // :177

    totals.rays += acc.cnt.rays;
    totals.primary_hits += acc.cnt.primary_hits;
    totals.shadow_rays += acc.cnt.shadow_rays;
    totals.grid_steps += acc.cnt.grid_steps;
    totals.voxel_steps += acc.cnt.voxel_steps;
    totals.status_fetches += acc.cnt.status_fetches;
    totals.bricks_entered += acc.cnt.bricks_entered;
    totals.hits += acc.cnt.hits;
}

Ctx make_ctx(const orc_scene* sc, const vrt_camera* cam, const vrt_sun* sun) {
    Ctx c{};
    c.sc = sc, c.cam = cam, c.sun = sun;
    const vrt_grid_state& s = sc->state;
    c.g_min = v3(s.min_point_base_t[0], s.min_point_base_t[1], s.min_point_base_t[2]);
    c.g_max = v3(s.max_point_scale[0], s.max_point_scale[1], s.max_point_scale[2]);
    c.g_scale = s.max_point_scale[3];
    c.brick_dim = I3{(int)s.dim_x, (int)s.dim_y, (int)s.dim_z};
    c.bd = (int)sc->brick_dim;
    c.brick_bytes = sc->brick_dim * sc->brick_dim * sc->brick_dim / 8;
    c.brick_voxel_scale = 1.0f / (float)sc->brick_dim;
    return c;
}

}  // namespace

// ====================================================================================================
// grid builder: brick/Grid.zig + brick/MaterialAllocator.zig
// ====================================================================================================
struct orc_grid {
    uint32_t brick_dim, brick_bits, brick_bytes;
    vrt_grid_state state;
    std::vector<uint32_t> statuses;
    std::vector<uint32_t> brick_indices;
    std::vector<uint8_t> occupancy;
    std::vector<uint32_t> start_indices;
    std::vector<uint8_t> material_indices;
    uint64_t brick_alloc;
    uint32_t active_bricks;  // State.active_bricks
    uint64_t next_material;  // MaterialAllocator.next_index
};

extern "C" {

// Grid.zig:36-114
orc_grid* orc_grid_create(uint32_t dim_x, uint32_t dim_y, uint32_t dim_z, uint32_t brick_dim, uint64_t brick_alloc,
                          const float min_point[3], float scale, float base_t) {
    if (!dim_x || !dim_y || !dim_z || !min_point) return nullptr;
    if (brick_dim != 4 && brick_dim != 8 && brick_dim != 16) return nullptr;
    orc_grid* g = new (std::nothrow) orc_grid();
    if (!g) return nullptr;
    g->brick_dim = brick_dim;
    g->brick_bits = brick_dim * brick_dim * brick_dim;  // State.zig:6
    g->brick_bytes = g->brick_bits / 8;                 // State.zig:7
    const uint64_t brick_count = (uint64_t)dim_x * dim_y * dim_z;
    g->brick_alloc = brick_alloc ? brick_alloc : brick_count;  // Grid.zig:51
    try {
        g->statuses.assign((brick_count + 31) / 32, 0u);                                // Grid.zig:43-45
        g->brick_indices.assign(brick_count, 0u);                                       // Grid.zig:47-49
        g->occupancy.assign(g->brick_alloc * g->brick_bytes, (uint8_t)0);               // Grid.zig:53-55
        g->start_indices.assign(g->brick_alloc, 0xffffffffu);                           // Grid.zig:57-59
        g->material_indices.assign(g->brick_alloc * (uint64_t)g->brick_bits, (uint8_t)0);  // Grid.zig:61-64
    } catch (...) {
        delete g;
        return nullptr;
    }
    vrt_grid_state& s = g->state;
    std::memset(&s, 0, sizeof(s));
    s.voxel_dim_x = dim_x * brick_dim, s.voxel_dim_y = dim_y * brick_dim, s.voxel_dim_z = dim_z * brick_dim;
    s.dim_x = dim_x, s.dim_y = dim_y, s.dim_z = dim_z;
    s.min_point_base_t[0] = min_point[0], s.min_point_base_t[1] = min_point[1], s.min_point_base_t[2] = min_point[2];
    s.min_point_base_t[3] = base_t;
    s.max_point_scale[0] = min_point[0] + (float)dim_x * scale;  // Grid.zig:74-79 (f32 mul then add)
    s.max_point_scale[1] = min_point[1] + (float)dim_y * scale;
    s.max_point_scale[2] = min_point[2] + (float)dim_z * scale;
    s.max_point_scale[3] = scale;
    g->active_bricks = 0;
    g->next_material = 0;
    return g;
}

void orc_grid_destroy(orc_grid* g) { delete g; }

// Grid.zig:129-194
int orc_grid_insert(orc_grid* g, uint32_t x, uint32_t y, uint32_t z, uint8_t material) {
    const vrt_grid_state& s = g->state;
    if (x >= s.voxel_dim_x || y >= s.voxel_dim_y || z >= s.voxel_dim_z) return -1;  // asserts :130-132
    const uint32_t bd = g->brick_dim;
    const uint32_t flipped_y = s.voxel_dim_y - 1 - y;  // :135

    // gridAt :206-211
    const uint64_t grid_index = (uint64_t)(x / bd) + (uint64_t)s.dim_x * ((z / bd) + (uint64_t)s.dim_z * (flipped_y / bd));
    const uint64_t status_index = grid_index / 32;
    const uint32_t status_offset = (uint32_t)(grid_index % 32);
    const bool loaded = (g->statuses[status_index] >> status_offset) & 1u;  // BrickStatusMask.read, State.zig:101-106
    uint32_t brick_index;
    if (loaded) {
        brick_index = g->brick_indices[grid_index];  // :143
    } else {
        if (g->active_bricks >= g->brick_alloc) return -2;
        brick_index = g->active_bricks++;  // fetchAdd :147
    }

    // voxelAt :198-203
    const uint32_t nth_bit = (x % bd) + bd * ((z % bd) + bd * (flipped_y % bd));

    uint32_t& start = g->start_indices[brick_index];
    if (start == 0xffffffffu) {  // :161
        if (g->next_material >= g->material_indices.size()) return -2;  // MaterialAllocator assert :40
        start = (uint32_t)g->next_material & 0x7fffffffu;               // value:u31, type bit = voxel_start_index = 0
        g->next_material += g->brick_bits;                              // MaterialAllocator.nextEntry :39
    }
    g->material_indices[(uint64_t)(start & 0x7fffffffu) + nth_bit] = material;  // :173-174

    g->occupancy[(uint64_t)brick_index * g->brick_bytes + nth_bit / 8] |= (uint8_t)(1u << (nth_bit % 8));  // :180-182
    g->statuses[status_index] |= (1u << status_offset);                                                   // :188
    g->brick_indices[grid_index] = brick_index;                                                           // :192
    return 0;
}

int orc_grid_insert_many(orc_grid* g, const uint32_t* xyzm, size_t n) {
    for (size_t i = 0; i < n; i++) {
        const int rc = orc_grid_insert(g, xyzm[4 * i], xyzm[4 * i + 1], xyzm[4 * i + 2], (uint8_t)xyzm[4 * i + 3]);
        if (rc) return rc;
    }
    return 0;
}

uint32_t orc_grid_active_bricks(const orc_grid* g) { return g->active_bricks; }
void orc_grid_get_state(const orc_grid* g, vrt_grid_state* out) { *out = g->state; }
const uint32_t* orc_grid_statuses(const orc_grid* g, uint64_t* n) { if (n) *n = g->statuses.size(); return g->statuses.data(); }
const uint32_t* orc_grid_brick_indices(const orc_grid* g, uint64_t* n) { if (n) *n = g->brick_indices.size(); return g->brick_indices.data(); }
const uint8_t* orc_grid_occupancy(const orc_grid* g, uint64_t* n) { if (n) *n = g->occupancy.size(); return g->occupancy.data(); }
const uint32_t* orc_grid_start_indices(const orc_grid* g, uint64_t* n) { if (n) *n = g->start_indices.size(); return g->start_indices.data(); }
const uint8_t* orc_grid_material_indices(const orc_grid* g, uint64_t* n) { if (n) *n = g->material_indices.size(); return g->material_indices.data(); }

void orc_scene_from_grid(const orc_grid* g, const vrt_material* materials, uint32_t n_materials, orc_scene* out) {
    std::memset(out, 0, sizeof(*out));
    out->state = g->state;
    out->brick_dim = g->brick_dim;
    out->materials = materials, out->n_materials = n_materials;
    out->statuses = g->statuses.data(), out->n_statuses = g->statuses.size();
    out->brick_indices = g->brick_indices.data(), out->n_brick_indices = g->brick_indices.size();
    out->occupancy = g->occupancy.data(), out->n_occupancy = g->occupancy.size();
    out->start_indices = g->start_indices.data(), out->n_start_indices = g->start_indices.size();
    out->material_indices = g->material_indices.data(), out->n_material_indices = g->material_indices.size();
}

int orc_render(const orc_scene* scene, const vrt_camera* camera, const vrt_sun* sun, uint32_t row_begin, uint32_t row_end,
               uint8_t* rgba8, vrt_aov* aov, vrt_counters* counters, int threads) {
    if (!scene || !camera || !sun || !rgba8 || !scene->materials || scene->n_materials == 0) return -1;
    if (row_end > camera->image_height || row_begin > row_end) return -1;
    const Ctx c = make_ctx(scene, camera, sun);
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    std::atomic<uint32_t> next_row{row_begin};
    std::vector<vrt_counters> per_thread((size_t)nt);
    auto worker = [&](int tid) {
        vrt_counters local{};
        for (;;) {
            const uint32_t y = next_row.fetch_add(1, std::memory_order_relaxed);
            if (y >= row_end) break;
            for (uint32_t x = 0; x < camera->image_width; x++) {
                vrt_aov* a = aov ? aov + ((uint64_t)y * camera->image_width + x) : nullptr;
                shade_pixel(c, x, y, rgba8, a, local);
            }
        }
        per_thread[(size_t)tid] = local;
    };
    if (nt == 1) {
        worker(0);
    } else {
        std::vector<std::thread> pool;
        for (int i = 0; i < nt; i++) pool.emplace_back(worker, i);
        for (auto& t : pool) t.join();
    }
    if (counters) {
        vrt_counters t{};
        for (const auto& p : per_thread) {
            t.rays += p.rays, t.primary_hits += p.primary_hits, t.shadow_rays += p.shadow_rays;
            t.grid_steps += p.grid_steps, t.voxel_steps += p.voxel_steps, t.status_fetches += p.status_fetches;
            t.bricks_entered += p.bricks_entered, t.hits += p.hits;
        }
        *counters = t;
    }
    return 0;
}

int orc_grid_hit(const orc_scene* scene, const float origin[3], const float direction[3], vrt_aov* out) {
    vrt_camera cam{};
    vrt_sun sun{};
    sun.enabled = 1;
    const Ctx c = make_ctx(scene, &cam, &sun);
    const Ray r = CreateRay(v3(origin[0], origin[1], origin[2]), v3(direction[0], direction[1], direction[2]));
    HitRecord hit{};
    V3 hit_min = v3s(0.0f);
    Trace tr;
    const bool got = GridHit(c, r, 0.00001f, std::numeric_limits<float>::infinity(), hit_min, hit, tr);
    if (out) {
        std::memset(out, 0, sizeof(*out));
        out->grid_index = out->voxel_index = out->material = ~0u;
        out->shadow_grid_index = out->shadow_voxel_index = ~0u;
        out->grid_steps = tr.grid_steps, out->voxel_steps = tr.voxel_steps, out->status_fetches = tr.status_fetches;
        // the slab-entry / last-step normal is reported even on a miss (hit.normal is written by the slab test)
        out->normal[0] = hit.normal.x, out->normal[1] = hit.normal.y, out->normal[2] = hit.normal.z;
        if (got) {
            out->flags = VRT_AOV_HIT;
            out->grid_index = tr.grid_index, out->voxel_index = tr.voxel_index, out->material = hit.index;
            out->t = hit.t;
            out->point[0] = hit.point.x, out->point[1] = hit.point.y, out->point[2] = hit.point.z;
        }
    }
    return got ? 1 : 0;
}

float orc_sinf(float x) { return det_sinf(x); }
float orc_hash12(float px, float py) { return hash12(V2{px, py}); }

}  // extern "C"
