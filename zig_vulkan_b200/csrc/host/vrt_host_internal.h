// vrt_host_internal.h — private definitions shared by the host-side sources (not installed).
#pragma once

#include <cmath>
#include <cstdint>
#include <vector>

#include "../../../include/vrt_host.h"

// DeviceDataDelta (State.zig:14-57) without the mutex: like the reference's render thread, a vrt_grid is driven
// from one thread.
struct vrt_delta {
    bool active;
    uint64_t from, to;
    void touch(uint64_t index);
    void reset();
};

struct vrt_grid {
    uint32_t brick_dim = 4, brick_bits = 64, brick_bytes = 8;
    vrt_grid_state state{};
    std::vector<uint32_t> statuses;
    std::vector<uint32_t> brick_indices;
    std::vector<uint8_t> occupancy;
    std::vector<uint32_t> start_indices;
    std::vector<uint8_t> material_indices;
    uint64_t brick_alloc = 0;
    uint32_t active_bricks = 0;   // State.active_bricks
    uint64_t next_material = 0;   // MaterialAllocator.next_index
    vrt_delta delta[5];
};

// ---------------------------------------------------------------------------------------------------------------
// The slice of kooparse/zalgebra (build.zig.zon:20-23, un-vendored) that Camera.zig / Sun.zig / Benchmark.zig call.
// Restated from the library's published formulas; f32 throughout.  Not pinned by any reference test.
// ---------------------------------------------------------------------------------------------------------------
namespace vrt_host {

struct Vec3 {
    float x, y, z;
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 scale(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline Vec3 norm(Vec3 a) {
    const float l = std::sqrt(dot(a, a));
    return {a.x / l, a.y / l, a.z / l};
}
inline float lerp(float from, float to, float t) { return (1.0f - t) * from + t * to; }
inline Vec3 lerp(Vec3 a, Vec3 b, float t) { return {lerp(a.x, b.x, t), lerp(a.y, b.y, t), lerp(a.z, b.z, t)}; }
constexpr Vec3 kUp = {0.0f, 1.0f, 0.0f};
constexpr Vec3 kRight = {1.0f, 0.0f, 0.0f};
constexpr Vec3 kForward = {0.0f, 0.0f, 1.0f};

struct Quat {
    float w, x, y, z;
};
constexpr Quat kIdentity = {1.0f, 0.0f, 0.0f, 0.0f};
inline Quat mul(Quat l, Quat r) {
    return {(-l.x * r.x) - (l.y * r.y) - (l.z * r.z) + (l.w * r.w), (l.x * r.w) + (l.y * r.z) - (l.z * r.y) + (l.w * r.x),
            (-l.x * r.z) + (l.y * r.w) + (l.z * r.x) + (l.w * r.y), (l.x * r.y) - (l.y * r.x) + (l.z * r.w) + (l.w * r.z)};
}
inline float qdot(Quat a, Quat b) { return a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z; }
inline Quat qscale(Quat a, float s) { return {a.w * s, a.x * s, a.y * s, a.z * s}; }
inline Quat qadd(Quat a, Quat b) { return {a.w + b.w, a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Quat qsub(Quat a, Quat b) { return {a.w - b.w, a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Quat qnorm(Quat q) {
    const float l = std::sqrt(qdot(q, q));
    if (l == 0.0f) return kIdentity;
    return {q.w / l, q.x / l, q.y / l, q.z / l};
}
inline Quat qlerp(Quat l, Quat r, float t) { return {lerp(l.w, r.w, t), lerp(l.x, r.x, t), lerp(l.y, r.y, t), lerp(l.z, r.z, t)}; }
// zalgebra (kooparse/zalgebra@7cf3b90, build.zig.zon:20-23) is an un-vendored dependency: its source is not in the reference
// tree, so this is a restatement of its published Quaternion.slerp from the algorithm, not from the text: the dot product's
// absolute value is used (q and -q are the same rotation, so the right operand is negated when the dot is negative and the
// interpolation takes the SHORTEST arc), nearly parallel operands fall back to a normalised lerp, otherwise
// sin((1-t)a)/sin(a) * l + sin(t a)/sin(a) * r.  Call sites: Sun.zig:72 (three orientations 120 degrees apart about z: the third
// segment, 240 degrees -> 0, has dot -0.5 and would sweep the long way round without the negation), Benchmark.zig:62.
// Host-side convenience only — the device path takes finished vrt_camera / vrt_sun structs.
inline Quat qslerp(Quat l, Quat r, float t) {
    const float threshold = 0.9995f;
    float cos_theta = qdot(l, r);
    if (cos_theta < 0.0f) {
        cos_theta = -cos_theta;
        r = qscale(r, -1.0f);
    }
    if (cos_theta > threshold) return qnorm(qlerp(l, r, t));
    const float angle = std::acos(cos_theta > 1.0f ? 1.0f : cos_theta);
    const float s = std::sin(angle);
    return qadd(qscale(l, std::sin((1.0f - t) * angle) / s), qscale(r, std::sin(t * angle) / s));
}
inline Vec3 rotate(Quat q_in, Vec3 v) {
    const Quat q = qnorm(q_in);
    const Vec3 b = {q.x, q.y, q.z};
    const float b2 = dot(b, b);
    return scale(v, q.w * q.w - b2) + scale(b, dot(v, b) * 2.0f) + scale(cross(b, v), q.w * 2.0f);
}
inline float to_radians(float deg) { return deg * (3.14159265358979323846f / 180.0f); }
inline float to_degrees(float rad) { return rad * (180.0f / 3.14159265358979323846f); }
inline Quat from_axis(float degrees, Vec3 axis) {
    const float radians = to_radians(degrees);
    const float s = std::sin(radians / 2.0f);
    const Vec3 a = scale(norm(axis), s);
    return {std::cos(radians / 2.0f), a.x, a.y, a.z};
}
inline Quat from_euler(Vec3 deg) {
    const Quat x = from_axis(deg.x, kRight), y = from_axis(deg.y, kUp), z = from_axis(deg.z, kForward);
    return mul(z, mul(y, x));
}
inline Vec3 extract_euler(Quat q) {
    const float yaw = std::atan2(2.0f * (q.y * q.z + q.w * q.x), q.w * q.w - q.x * q.x - q.y * q.y + q.z * q.z);
    const float pitch = std::asin(-2.0f * (q.x * q.z - q.w * q.y));
    const float roll = std::atan2(2.0f * (q.x * q.y + q.w * q.z), q.w * q.w + q.x * q.x - q.y * q.y - q.z * q.z);
    return {to_degrees(yaw), to_degrees(pitch), to_degrees(roll)};
}

}  // namespace vrt_host

struct vrt_hcam {
    float turn_rate, normal_speed, sprint_speed, movement_speed;
    bool user_input_disabled;
    float viewport_width, viewport_height, vertical_fov;
    vrt_host::Quat pitch, yaw;
    vrt_camera d_camera;
};

struct vrt_benchmark {  // Benchmark.zig:11-20 + Report (:81-101)
    vrt_hcam* camera;
    bool sun_enabled;
    float timer, duration, fraction, extent_scale;
    float min_dt, max_dt, dt_sum;
    uint32_t samples;
    uint32_t voxel_dim[3];
};

struct vrt_hsun {
    vrt_sun device_data;
    bool animate;
    float animate_speed;
    size_t slerp_index;
    float slerp_pos;
    vrt_host::Quat slerp_orientations[3];
    vrt_host::Vec3 lerp_color[3];
    vrt_host::Vec3 static_pos_vec;
};
