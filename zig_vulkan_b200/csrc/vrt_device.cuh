// vrt_device.cuh — device-side FP32 building blocks shared by every trace kernel.
//
// FP discipline (DESIGN.md "FP discipline"): this translation unit is compiled with --fmad=false, so the only
// fused multiply-adds are the ones spelled fmaf() here — exactly the calls the reference shader spells fma()
// (assets/shaders/brick_raytracer.comp:194,199,298,331,395,475,572).  Division and sqrt are IEEE (nvcc
// defaults -prec-div=true -prec-sqrt=true -ftz=false).  Every helper mirrors one GLSL built-in under the
// documented choice (normalize, dot order, fract, sign, min/max, sin).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/vrt.h"

namespace vrt {

struct V3 {
    float x, y, z;
};
struct V2 {
    float x, y;
};
struct I3 {
    int x, y, z;
};

#define VRT_DI __device__ __forceinline__

VRT_DI V3 v3(float x, float y, float z) { return V3{x, y, z}; }
VRT_DI V3 v3s(float s) { return V3{s, s, s}; }
VRT_DI V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
VRT_DI V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
VRT_DI V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
VRT_DI V3 operator/(V3 a, V3 b) { return V3{a.x / b.x, a.y / b.y, a.z / b.z}; }
VRT_DI V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
VRT_DI V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
VRT_DI V3 neg(V3 a) { return V3{-a.x, -a.y, -a.z}; }
VRT_DI V3 fma3(V3 a, V3 b, V3 c) { return V3{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z)}; }
VRT_DI V3 tofloat(I3 i) { return V3{(float)i.x, (float)i.y, (float)i.z}; }
VRT_DI V3 ld3(const float* p) { return V3{p[0], p[1], p[2]}; }
VRT_DI float gmin(float a, float b) { return b < a ? b : a; }
VRT_DI float gmax(float a, float b) { return a < b ? b : a; }
VRT_DI float gsign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
VRT_DI float fract(float x) { return x - floorf(x); }
VRT_DI float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
VRT_DI float dot2(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
VRT_DI V3 normalize3(V3 v) {
    const float inv = 1.0f / sqrtf((v.x * v.x + v.y * v.y) + v.z * v.z);
    return v * inv;
}
VRT_DI V3 reflect3(V3 i, V3 n) { return i - (2.0f * dot3(n, i)) * n; }
VRT_DI float comp(V3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

// sin() of the shader's sin-hash (rand.comp:3-4).  GLSL promises only 2^-11 absolute error, and libm/CUDA
// sinf differ in the last ulp, which the hash amplifies; this explicit Cody-Waite + polynomial evaluation
// rounds identically on every IEEE machine.
VRT_DI float det_sinf(float x) {
    const float kf = rintf(x * 0.636619747f);
    const int k = (int)kf;
    float r = fmaf(kf, -1.57079601e+00f, x);
    r = fmaf(kf, -3.13916473e-07f, r);
    r = fmaf(kf, -5.39030253e-15f, r);
    const float s = r * r;
    float res;
    if (k & 1) {
        float p = 2.44677067e-5f;
        p = fmaf(p, s, -1.38877297e-3f);
        p = fmaf(p, s, 4.16666567e-2f);
        p = fmaf(p, s, -5.00000000e-1f);
        res = fmaf(p, s, 1.0f);
    } else {
        float p = 2.86567956e-6f;
        p = fmaf(p, s, -1.98559923e-4f);
        p = fmaf(p, s, 8.33338592e-3f);
        p = fmaf(p, s, -1.66666672e-1f);
        const float t = r * s;
        res = fmaf(p, t, r);
    }
    return (k & 2) ? -res : res;
}

// rand.comp:3-26
VRT_DI float Rand1(float co) { return fract(det_sinf(co * 91.3458f) * 47453.5453f); }
VRT_DI float Rand2(V2 co) { return fract(det_sinf(dot2(co, V2{12.9898f, 78.233f})) * 43758.5453f); }
VRT_DI float Rand3(V3 co) {
    const float r = Rand1(co.z);
    return Rand2(V2{co.x + r, co.y + r});
}
VRT_DI float Rand2mm(V2 co, float mn, float mx) { return mn + (mx - mn) * Rand2(co); }
VRT_DI V3 RandVec3mm(V2 co, float mn, float mx) {
    const float x = Rand2mm(co, mn, mx);
    const float y = Rand2mm(V2{co.x + x, co.y + x}, mn, mx);
    const float z = Rand2mm(V2{co.x + y, co.y + y}, mn, mx);
    return v3(x, y, z);
}
VRT_DI float hash12(V2 p) {
    V3 p3 = v3(fract(p.x * .1031f), fract(p.y * .1031f), fract(p.x * .1031f));
    const float d = dot3(p3, v3(p3.y + 33.33f, p3.z + 33.33f, p3.x + 33.33f));
    p3 = v3(p3.x + d, p3.y + d, p3.z + d);
    return fract((p3.x + p3.y) * p3.z);
}

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

// brick_raytracer.comp:180-201
VRT_DI Ray CreateRay(V3 origin, V3 direction) { return Ray{origin, normalize3(direction), 1.0f, VRT_MAT_NONE}; }
VRT_DI V3 RayAt(const Ray& r, float t) { return fma3(v3s(t), r.direction, r.origin); }
VRT_DI V3 BackgroundColor(const Ray& r) {
    const float t = 0.5f * (r.direction.y + 1.0f);
    return fma3(v3s(1.0f - t), v3s(1.0f), t * v3(0.5f, 0.7f, 1.0f));
}
// brick_raytracer.comp:267
VRT_DI float safeInverse(float x) { return (x == 0.0f) ? 1e12f : (1.0f / x); }

// brick_raytracer.comp:522-536 (+ :497-503).  Returns t_min <= t_max; ties in the arg-max go to axis 0.
VRT_DI bool AdvNormIntersect(V3 box_min, V3 box_max, const Ray& r, V3 inv, V3& normal, float& t_min, float& t_max) {
    const V3 t_lower = (box_min - r.origin) * inv;
    const V3 t_upper = (box_max - r.origin) * inv;
    const V3 t_mins = v3(gmin(t_lower.x, t_upper.x), gmin(t_lower.y, t_upper.y), gmin(t_lower.z, t_upper.z));
    const V3 t_maxes = v3(gmax(t_lower.x, t_upper.x), gmax(t_lower.y, t_upper.y), gmax(t_lower.z, t_upper.z));
    const int idx = int(t_mins.y > t_mins.x && t_mins.y > t_mins.z) + int(t_mins.z > t_mins.x && t_mins.z > t_mins.y) * 2;
    const float s = gsign(comp(inv, idx));
    normal = v3(idx == 0 ? s : 0.0f, idx == 1 ? s : 0.0f, idx == 2 ? s : 0.0f);
    t_min = gmax(t_min, comp(t_mins, idx));
    t_max = gmin(t_max, gmin(gmin(t_maxes.x, t_maxes.y), t_maxes.z));
    return t_min <= t_max;
}

// The DDA advance of GridHit (:345-372) and BrickHit (:440-467).  Returns the axis stepped (0 x, 1 y, 2 z);
// the caller derives hit.normal = -step[axis] on that axis, which is what normal_axis encodes (:304-308).
VRT_DI int dda_step(V3& side_dist, V3 ray_delta, I3& pos, I3 ray_step, float scale, float& t_value) {
    if (side_dist.x < side_dist.y) {
        if (side_dist.x < side_dist.z) {
            t_value = side_dist.x * scale;
            side_dist.x += ray_delta.x;
            pos.x += ray_step.x;
            return 0;
        }
        t_value = side_dist.z * scale;
        side_dist.z += ray_delta.z;
        pos.z += ray_step.z;
        return 2;
    }
    if (side_dist.y < side_dist.z) {
        t_value = side_dist.y * scale;
        side_dist.y += ray_delta.y;
        pos.y += ray_step.y;
        return 1;
    }
    t_value = side_dist.z * scale;
    side_dist.z += ray_delta.z;
    pos.z += ray_step.z;
    return 2;
}

// hit.normal for "stepped along `axis`": vec3 with normal_axis[axis] on that axis (:350,356,364,370).
VRT_DI V3 axis_normal(int axis, I3 ray_step) {
    const float nx = ray_step.x < 0 ? 1.0f : -1.0f;
    const float ny = ray_step.y < 0 ? 1.0f : -1.0f;
    const float nz = ray_step.z < 0 ? 1.0f : -1.0f;
    return v3(axis == 0 ? nx : 0.0f, axis == 1 ? ny : 0.0f, axis == 2 ? nz : 0.0f);
}

// side_dist initialiser shared by :293-298 and :393-395
VRT_DI V3 init_side_dist(V3 fposition, I3 ray_step, V3 ray_delta) {
    const V3 intersection_delta = v3(floorf(fposition.x), floorf(fposition.y), floorf(fposition.z)) - fposition;
    const V3 fstep = tofloat(ray_step);
    return fma3(fstep, intersection_delta, fstep * 0.5f + v3s(0.5f)) * ray_delta;
}

// brick_raytracer.comp:564-574
VRT_DI bool transmissionDirection(float n1, float n2, V3 ray_dir, V3 normal, V3& refrac_dir) {
    const float eta = n1 / n2;
    const float c1 = -dot3(ray_dir, normal);
    const float w = eta * c1;
    const float c2m = (w - eta) * (w + eta);
    if (c2m < -1.0f) return false;
    refrac_dir = fma3(v3s(eta), ray_dir, (w - sqrtf(1.0f + c2m)) * normal);
    return true;
}

// brick_raytracer.comp:539-596; `mtype` selects the switch arm of :225-239.  Returns `result`.
VRT_DI bool scatter(uint32_t mtype, float type_data, const Ray& r_in, const HitRecord& hit, Ray& scattered, int& loop_count) {
    const V2 co = V2{hit.point.x + hit.point.z, hit.point.y + hit.point.z};
    if (mtype == VRT_MAT_LAMBERTIAN) {
        const V3 scatter_dir = normalize3(hit.normal + RandVec3mm(co, -0.4f, 0.4f));
        scattered = CreateRay(hit.point, scatter_dir);
        return true;
    }
    if (mtype == VRT_MAT_METAL) {
        const V3 reflected = reflect3(r_in.direction, hit.normal);
        const float fuzz = type_data;
        scattered = CreateRay(hit.point, reflected + RandVec3mm(co, -fuzz, fuzz));
        return dot3(scattered.direction, hit.normal) > 0;
    }
    if (mtype == VRT_MAT_DIELECTRIC) {
        const float ir = type_data;
        const V3 normal = normalize3(hit.normal + RandVec3mm(co, -0.05f, 0.05f));
        V3 direction = v3s(0.0f);
        const bool should_refract = transmissionDirection(ir, r_in.internal_reflection, r_in.direction, normal, direction);
        if (should_refract && Rand3(hit.point) > 0.5f) {
            scattered = CreateRay(hit.point, direction);
            scattered.ignore_type_material = VRT_MAT_DIELECTRIC;
            scattered.internal_reflection = ir;
        } else {
            direction = reflect3(r_in.direction, normal);
            scattered = CreateRay(hit.point, direction);
        }
        return true;
    }
    loop_count -= 1;  // default: arm (:235-238)
    return false;
}

// Rgba8 imageStore conversion (:177)
VRT_DI uint32_t unorm8(float c) {
    if (!(c > 0.0f)) return 0u;
    if (c > 1.0f) c = 1.0f;
    return (uint32_t)(c * 255.0f + 0.5f);
}
VRT_DI uint32_t pack_rgba8(V3 c) { return unorm8(c.x) | (unorm8(c.y) << 8) | (unorm8(c.z) << 16) | 0xff000000u; }

// Everything a trace kernel needs, passed by value as a __grid_constant__ (lands in the constant bank, the
// CUDA analogue of the reference's push constants + UBO + descriptor set).
struct TraceParams {
    vrt_camera cam;         // push constants [0,96)
    vrt_sun sun;            // push constants [96,128)
    vrt_grid_state grid;    // binding 1
    const vrt_material* materials;      // binding 2
    const uint32_t* statuses;           // binding 3
    const uint32_t* brick_indices;      // binding 4
    const uint8_t* occupancy;           // binding 5
    const uint32_t* start_indices;      // binding 6
    const uint8_t* material_indices;    // binding 7
    unsigned long long n_statuses, n_brick_indices, n_occupancy, n_start_indices, n_material_indices;
    uint32_t n_materials;
    uint32_t materials_basic;      // every uploaded material is lambertian / metal / dielectric (precondition of the simple shading path)
    uint32_t materials_have_none;  // any uploaded material with type == MAT_NONE(3)? (enables the :427 test for type-3 rays)
    int brick_dim;            // spec const 4
    uint32_t brick_bytes;     // spec const 3
    float brick_voxel_scale;  // spec const 5 (host-computed 1.0f/brick_dim, Pipeline.zig:313)
    uint32_t row_begin, row_end;
    // interleaved partition (il_world > 0): this launch traces the 4-row strips t with t % il_world == il_rank.
    // il_gather: store strip k of this rank at rows (il_rank * il_strips_max + k) * 4 of `fb` (rank-major all-gather layout)
    uint32_t il_world, il_rank, il_gather, il_strips_max;
    uint32_t* fb;                 // RGBA8 image, binding 0 (row-major, width*height words)
    vrt_aov* aov;                 // nullable
    unsigned long long* counters; // nullable, 8 x u64 in vrt_counters order
    // derived acceleration structures (built on the device from the buffers above, vrt_kernels.cu "build_*")
    const uint4* cell_rec;                // [n_bricks] per GRID cell: 4^3 voxel mask (x, y), material start (z), brick index (w); brick_dim == 4 only
    const uint8_t* dist;                  // 8 padded directional Chebyshev distance grids (one per octant), see vrt_trav_warp.cuh
    unsigned long long dist_plane;        // bytes per octant
    uint32_t dist_log_px, dist_log_pz;    // row / plane strides of `dist` are powers of two: x + (z << log_px) + (y << (log_px+log_pz))
    uint32_t scale_pow2, voxel_scale_pow2;  // brick / voxel scale is a power of two -> divide by multiplying with the exact inverse
    float inv_scale, inv_voxel_scale;
    // persistent-kernel work queue: counter[0] = next ticket, counter[1] = warps that have left the queue; the last warp to
    // leave resets both, so every launch starts from 0 with identical parameters (frames can be replayed from a CUDA graph)
    unsigned long long* tile_counter;
    uint32_t queue_world;  // GPUs whose warps pull from this queue (1: private; > 1: it lives in rank 0's memory, reached over NVLink)
    // Tile schedule (vrt_sched.cu).  order: a permutation of this launch's tile space, most expensive tile first (longest-
    // processing-time-first keeps the tail of the launch short); ticket i traces tile order[order_offset + i * order_stride].
    // NULL: bottom-up (ground rows first, sky tiles fill the tail).  cost: clock ticks / 32 each tile took, written by whoever
    // traces it (and into every peer's copy when the tiles of one frame are dealt across GPUs), the sort key of the next order.
    const uint32_t* tile_order;
    uint32_t order_offset, order_stride;
    uint16_t* tile_cost;
    uint16_t* peer_cost[8];
    uint32_t n_cost_peers;  // entries of peer_cost in use (dealt schedules across GPUs)
    uint32_t vec_store_ok;                // framebuffer rows are 16-B aligned -> 128-bit stores
    // fused peer-store exchange (multi-GPU): framebuffers of every rank, this rank included
    uint32_t* peer_fb[8];
    uint32_t n_peers;
    // tile-major exchange (VRT_EXCHANGE_PEER_TILES): instead of four 32-byte row segments into every peer's row-major image, a tile
    // goes as ONE contiguous 128-byte record (32 texels, lane order) into the tile-major staging buffer of every rank, this one
    // included; each rank un-tiles its own staging buffer into its framebuffer after the frame barrier (untile_kernel)
    uint32_t* stage[8];
    uint32_t n_stage;
    uint32_t one;  // = 1, opaque to the compiler: multiplier of the march's index IMADs (vrt_trav_warp.cuh march_step)
};

}  // namespace vrt
