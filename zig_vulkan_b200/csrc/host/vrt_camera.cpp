// vrt_camera.cpp — host camera and sun: C++ counterparts of src/modules/voxel_rt/Camera.zig and Sun.zig.
// Their only product is the 96-byte vrt_camera / 32-byte vrt_sun blocks handed to vrt_trace, exactly what the
// reference pushes as push constants (ComputePipeline.zig:488-505).
#include <cmath>
#include <cstring>
#include <new>

#include "vrt_host_internal.h"

using namespace vrt_host;

namespace {

void store3(float dst[3], Vec3 v) { dst[0] = v.x, dst[1] = v.y, dst[2] = v.z; }
Vec3 load3(const float src[3]) { return {src[0], src[1], src[2]}; }

Quat orientation(const vrt_hcam* c) { return qnorm(mul(c->yaw, c->pitch)); }         // Camera.zig:153-155
Vec3 forward_dir(const vrt_hcam* c) { return rotate(orientation(c), {0, 0, 1}); }  // :162-164

// Camera.propogatePitchChange + lowerLeftCorner (:167-180)
void propagate(vrt_hcam* c) {
    const Vec3 forward = forward_dir(c);
    const Vec3 right = norm(cross(kUp, forward));
    const Vec3 up = norm(cross(forward, right));
    const Vec3 horizontal = scale(right, c->viewport_width);
    const Vec3 vertical = scale(up, c->viewport_height);
    store3(c->d_camera.horizontal, horizontal);
    store3(c->d_camera.vertical, vertical);
    const Vec3 llc = load3(c->d_camera.origin) - scale(horizontal, 0.5f) - scale(vertical, 0.5f) - forward;
    store3(c->d_camera.lower_left_corner, llc);
}

}  // namespace

extern "C" {

void vrt_hcam_default_config(vrt_hcam_config* out) {  // Camera.Config (Camera.zig:5-14)
    if (!out) return;
    std::memset(out, 0, sizeof(*out));
    out->viewport_height = 2.0f;
    out->samples_per_pixel = 2;
    out->max_bounce = 2;
    out->turn_rate = 0.1f;
    out->normal_speed = 1.0f;
    out->sprint_speed = 2.0f;
    out->user_input_disabled = 0;
}

// Camera.init (Camera.zig:36-77)
vrt_hcam* vrt_hcam_create(float vertical_fov_deg, uint32_t image_width, uint32_t image_height, const vrt_hcam_config* cfg_in) {
    if (image_width == 0 || image_height == 0) return nullptr;
    vrt_hcam_config cfg;
    if (cfg_in) cfg = *cfg_in;
    else vrt_hcam_default_config(&cfg);
    vrt_hcam* c = new (std::nothrow) vrt_hcam();
    if (!c) return nullptr;
    const float aspect_ratio = (float)image_width / (float)image_height;
    const float a = (float)(3.14159265358979323846 * (1.0 / 180.0));  // comptime_float folded, then used in f32 math
    const float theta = vertical_fov_deg * a;
    const float viewport_height = cfg.viewport_height * std::tan(theta * 0.5f);
    const float viewport_width = aspect_ratio * viewport_height;

    c->turn_rate = cfg.turn_rate;
    c->normal_speed = cfg.normal_speed;
    c->sprint_speed = cfg.sprint_speed;
    c->movement_speed = cfg.normal_speed;
    c->user_input_disabled = cfg.user_input_disabled != 0;
    c->viewport_width = viewport_width;
    c->viewport_height = viewport_height;
    c->vertical_fov = vertical_fov_deg;
    c->pitch = kIdentity;
    c->yaw = kIdentity;
    std::memset(&c->d_camera, 0, sizeof(c->d_camera));
    c->d_camera.image_width = image_width;
    c->d_camera.image_height = image_height;
    for (int i = 0; i < 3; i++) c->d_camera.origin[i] = cfg.origin[i];
    c->d_camera.samples_per_pixel = cfg.samples_per_pixel;
    c->d_camera.max_bounce = cfg.max_bounce + 1;  // :74
    // :47-53 computes the same basis as propogatePitchChange does for the identity orientation
    propagate(c);
    return c;
}

void vrt_hcam_destroy(vrt_hcam* c) { delete c; }
void vrt_hcam_device(const vrt_hcam* c, vrt_camera* out) {
    if (c && out) *out = c->d_camera;
}
void vrt_hcam_set_origin(vrt_hcam* c, const float origin[3]) {  // :90-93
    if (!c || !origin) return;
    for (int i = 0; i < 3; i++) c->d_camera.origin[i] = origin[i];
    propagate(c);
}
void vrt_hcam_translate(vrt_hcam* c, float delta_time, const float by[3]) {  // :113-123
    if (!c || !by || c->user_input_disabled) return;
    const Vec3 n = norm(load3(by));
    const Vec3 delta = rotate(orientation(c), scale(n, delta_time * c->movement_speed));
    if (std::isnan(delta.x)) return;
    c->d_camera.origin[0] += delta.x, c->d_camera.origin[1] += delta.y, c->d_camera.origin[2] += delta.z;
    propagate(c);
}
void vrt_hcam_turn_pitch(vrt_hcam* c, float angle) {  // :125-142
    if (!c || c->user_input_disabled) return;
    const float h_angle = angle * c->turn_rate;
    const Quat prev = c->pitch;
    c->pitch = mul(c->pitch, Quat{std::cos(h_angle), std::sin(h_angle), 0.0f, 0.0f});
    if (std::fabs(extract_euler(c->pitch).x) >= 90.0f) c->pitch = prev;
    propagate(c);
}
void vrt_hcam_turn_yaw(vrt_hcam* c, float angle) {  // :144-152
    if (!c || c->user_input_disabled) return;
    const float h_angle = angle * c->turn_rate;
    c->yaw = mul(c->yaw, Quat{std::cos(h_angle), 0.0f, std::sin(h_angle), 0.0f});
    propagate(c);
}
void vrt_hcam_reset(vrt_hcam* c) {  // :105-110
    if (!c) return;
    c->user_input_disabled = false;
    c->yaw = kIdentity;
    c->pitch = kIdentity;
    propagate(c);
}
void vrt_hcam_activate_sprint(vrt_hcam* c) {
    if (c) c->movement_speed = c->normal_speed * c->sprint_speed;
}
void vrt_hcam_disable_sprint(vrt_hcam* c) {
    if (c) c->movement_speed = c->normal_speed;
}
void vrt_hcam_disable_input(vrt_hcam* c) {
    if (c) c->user_input_disabled = true;
}
void vrt_hcam_enable_input(vrt_hcam* c) {
    if (c) c->user_input_disabled = false;
}
void vrt_hcam_set_orientation(vrt_hcam* c, const float yaw_wxyz[4], const float pitch_wxyz[4]) {  // Benchmark.zig:27-31,57-66
    if (!c || !yaw_wxyz) return;
    c->yaw = Quat{yaw_wxyz[0], yaw_wxyz[1], yaw_wxyz[2], yaw_wxyz[3]};
    c->pitch = pitch_wxyz ? Quat{pitch_wxyz[0], pitch_wxyz[1], pitch_wxyz[2], pitch_wxyz[3]} : kIdentity;
    propagate(c);
}
void vrt_hcam_set_euler_deg(vrt_hcam* c, float x_deg, float y_deg, float z_deg) {
    if (!c) return;
    c->yaw = from_euler({x_deg, y_deg, z_deg});
    c->pitch = kIdentity;
    propagate(c);
}

// ------------------------------------------------------------------------------------------------ Sun.zig

void vrt_hsun_default_config(vrt_hsun_config* out) {  // Sun.Config (Sun.zig:4-11)
    if (!out) return;
    out->animate = 1;
    out->animate_speed = 0.1f;
    out->enabled = 1;
    out->color[0] = 1.0f, out->color[1] = 1.1f, out->color[2] = 1.0f;
    out->radius = 5.0f;
    out->sun_distance = 1000.0f;
}

vrt_hsun* vrt_hsun_create(const vrt_hsun_config* cfg_in) {  // Sun.init (Sun.zig:35-63)
    vrt_hsun_config cfg;
    if (cfg_in) cfg = *cfg_in;
    else vrt_hsun_default_config(&cfg);
    vrt_hsun* s = new (std::nothrow) vrt_hsun();
    if (!s) return nullptr;
    s->slerp_orientations[0] = from_euler({0, 0, 0});
    s->slerp_orientations[1] = from_euler({0, 10, 120});
    s->slerp_orientations[2] = from_euler({0, 0, 240});
    s->static_pos_vec = {0.0f, -cfg.sun_distance, 0.0f};
    s->lerp_color[0] = {1.0f, 0.99f, 0.823f};
    s->lerp_color[1] = {0.9f, 0.45f, 0.45f};
    s->lerp_color[2] = {1.0f, 0.7569f, 0.5412f};
    s->device_data.enabled = cfg.enabled ? 1u : 0u;
    s->device_data.position[0] = s->static_pos_vec.x, s->device_data.position[1] = s->static_pos_vec.y, s->device_data.position[2] = s->static_pos_vec.z;
    for (int i = 0; i < 3; i++) s->device_data.color[i] = cfg.color[i];
    s->device_data.radius = cfg.radius;
    s->animate = cfg.animate != 0;
    s->animate_speed = cfg.animate_speed;
    s->slerp_index = 0;
    s->slerp_pos = 0.0f;
    return s;
}

void vrt_hsun_destroy(vrt_hsun* s) { delete s; }
void vrt_hsun_device(const vrt_hsun* s, vrt_sun* out) {
    if (s && out) *out = s->device_data;
}

void vrt_hsun_update(vrt_hsun* s, float delta_time) {  // Sun.update (Sun.zig:65-86)
    if (!s || !s->animate || s->device_data.enabled == 0) return;
    const size_t next_index = (s->slerp_index + 1) % 3;
    const Vec3 p = rotate(qslerp(s->slerp_orientations[s->slerp_index], s->slerp_orientations[next_index], s->slerp_pos), s->static_pos_vec);
    s->device_data.position[0] = p.x, s->device_data.position[1] = p.y, s->device_data.position[2] = p.z;
    const Vec3 col = lerp(s->lerp_color[s->slerp_index], s->lerp_color[next_index], s->slerp_pos);
    s->device_data.color[0] = col.x, s->device_data.color[1] = col.y, s->device_data.color[2] = col.z;
    s->slerp_pos += s->animate_speed * delta_time;
    if (s->slerp_pos > 1.0f) {
        float ipart;
        s->slerp_pos = std::modf(s->slerp_pos, &ipart);
        s->slerp_index = next_index;
    }
}

}  // extern "C"
