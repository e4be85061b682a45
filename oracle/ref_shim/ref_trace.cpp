/*
 * ref_trace.cpp — the reference's OWN compute shader (assets/shaders/brick_raytracer.comp + rand.comp), compiled by g++ and run
 * on the host, one invocation per pixel.  TEST INFRASTRUCTURE, NOT PRODUCT (only tests/, smoke() and bench.py's CPU legs load it).
 *
 * oracle/_ref/brick_raytracer.comp.inc is the shader text after translate.py's lexical pass; this file is the "Vulkan driver":
 * it binds the UBO / SSBOs / push constants / specialization constants (ComputePipeline.zig:105-303, Pipeline.zig:273-315),
 * sets gl_GlobalInvocationID and calls main() for every pixel of the dispatch (ComputePipeline.zig:547-550).
 * Built by oracle/Makefile (`make -C oracle ref`) into oracle/_ref/libref_shader.so when /root/reference is present.
 * With -DREF_WIDE_MASK_INDEX (libref_shader_wide.so) the text is brick_raytracer_wide.comp.inc: the one-line 16^3-brick extension.
 */
#define REFSHADER_NS refshader_trace
#include "glsl_compat.h"
// (namespace refshader_trace is open)

thread_local uvec3 gl_GlobalInvocationID;

#ifdef REF_WIDE_MASK_INDEX
#include "../_ref/brick_raytracer_wide.comp.inc"
#else
#include "../_ref/brick_raytracer.comp.inc"
#endif

}  // namespace refshader_trace

#include <atomic>
#include <mutex>
#include <thread>
#include <vector>

#include "../vrt_oracle.h"  // orc_scene, vrt_camera, vrt_sun: plain C structs of the public ABI
#include "ref_shader.h"

namespace rs = refshader_trace;

static_assert(sizeof(rs::Material) == sizeof(vrt_material), "Material stride (std430: 20 bytes)");
static_assert(sizeof(rs::BrickGridState) == sizeof(vrt_grid_state), "BrickGridState (std140: 64 bytes)");

namespace {
std::mutex g_bind_mutex;  // the shader's resources are globals, as in GLSL: one dispatch at a time

void bind(const orc_scene* sc, const vrt_camera* cam, const vrt_sun* sun, uint8_t* rgba8) {
    // specialization constants (Pipeline.zig:293-315; brick_dimension is a Zig constant 4 upstream, State.zig:5)
    const uint32_t bd = sc->brick_dim;
    rs::brick_bits = bd * bd * bd;
    rs::brick_bytes = bd * bd * bd / 8;
    rs::brick_dimensions = (int)bd;
    rs::brick_voxel_scale = 1.0f / (float)bd;
    // push constants: Camera.Device @0 + Sun.Device @96 (ComputePipeline.zig:488-505), member by member
    auto& pc = rs::push_constant;
    pc.image_width = cam->image_width, pc.image_height = cam->image_height;
    pc.horizontal = rs::vec3(cam->horizontal[0], cam->horizontal[1], cam->horizontal[2]);
    pc.vertical = rs::vec3(cam->vertical[0], cam->vertical[1], cam->vertical[2]);
    pc.lower_left_corner = rs::vec3(cam->lower_left_corner[0], cam->lower_left_corner[1], cam->lower_left_corner[2]);
    pc.origin = rs::vec3(cam->origin[0], cam->origin[1], cam->origin[2]);
    pc.paddin = 0.0f;
    pc.samples_per_pixel = cam->samples_per_pixel, pc.max_bounce = cam->max_bounce;
    pc.sun_position = rs::vec3(sun->position[0], sun->position[1], sun->position[2]);
    pc.sun_enabled = sun->enabled;
    pc.sun_color = rs::vec3(sun->color[0], sun->color[1], sun->color[2]);
    pc.sun_radius = sun->radius;
    // UBO (binding 1) and the six SSBOs (bindings 2-7)
    std::memcpy(&rs::brick_grid, &sc->state, sizeof(vrt_grid_state));
    rs::materials.data = reinterpret_cast<const rs::Material*>(sc->materials), rs::materials.count = sc->n_materials;
    rs::brick_type_bits.data = sc->statuses, rs::brick_type_bits.count = sc->n_statuses;
    rs::brick_indices.data = sc->brick_indices, rs::brick_indices.count = sc->n_brick_indices;
    rs::brick_solid_mask.data = sc->occupancy, rs::brick_solid_mask.count = sc->n_occupancy;
    rs::brick_type_and_index.data = sc->start_indices, rs::brick_type_and_index.count = sc->n_start_indices;
    rs::material_indices.data = sc->material_indices, rs::material_indices.count = sc->n_material_indices;
    // storage image (binding 0)
    rs::img_output.rgba8 = rgba8, rs::img_output.width = (int)cam->image_width, rs::img_output.height = (int)cam->image_height;
}

void record_hit(bool got, const rs::HitRecord& hit, ref_hit* out) {
    std::memset(out, 0, sizeof(*out));
    if (!got) return;
    out->hit = 1u, out->index = hit.index, out->t = hit.t;
    out->point[0] = hit.point.x, out->point[1] = hit.point.y, out->point[2] = hit.point.z;
    out->normal[0] = hit.normal.x, out->normal[1] = hit.normal.y, out->normal[2] = hit.normal.z;
}
}  // namespace

extern "C" int ref_trace_render(const orc_scene* scene, const vrt_camera* camera, const vrt_sun* sun, uint32_t row_begin, uint32_t row_end,
                                uint8_t* rgba8, ref_hit* hits, int threads) {
    if (!scene || !camera || !sun || !rgba8 || row_begin > row_end || row_end > camera->image_height) return -1;
    // a 1-pixel-wide / -high image makes u or v = 0/0 (:168,:170): the shader's DDA would never terminate (DESIGN.md "Deviations")
    if (camera->image_width < 2 || camera->image_height < 2) return -2;
    std::lock_guard<std::mutex> lock(g_bind_mutex);
    bind(scene, camera, sun, rgba8);
    int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    std::atomic<uint32_t> next_row{row_begin};
    const uint32_t width = camera->image_width, height = camera->image_height;
    auto worker = [&]() {
        for (;;) {
            const uint32_t y = next_row.fetch_add(1);
            if (y >= row_end) break;
            for (uint32_t x = 0; x < width; x++) {
                rs::gl_GlobalInvocationID = rs::uvec3{x, y, 0u};
                rs::shader_main();  // brick_raytracer.comp:153
                if (hits) {  // the primary ray of sample 0 once more, through the shader's own CameraGetRay + GridHit (:167-171, :218)
                    const float u = ((float)x + 0.0f) / (float)(width - 1u), v = ((float)y + 0.0f) / (float)(height - 1u);
                    const rs::Ray ray = rs::CameraGetRay(u, v);
                    rs::vec3 hit_min;
                    rs::HitRecord hit;
                    hit.t = 0.0f, hit.index = 0u;
                    const bool got = rs::GridHit(ray, 0.00001f, rs::infinity, hit_min, hit);
                    record_hit(got, hit, hits + (size_t)y * width + x);
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; t++) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    return 0;
}

extern "C" int ref_trace_grid_hit(const orc_scene* scene, const float origin[3], const float direction[3], ref_hit* out) {
    if (!scene || !origin || !direction || !out) return -1;
    std::lock_guard<std::mutex> lock(g_bind_mutex);
    vrt_camera cam;
    vrt_sun sun;
    std::memset(&cam, 0, sizeof(cam));
    std::memset(&sun, 0, sizeof(sun));
    bind(scene, &cam, &sun, nullptr);
    const rs::Ray ray = rs::CreateRay(rs::vec3(origin[0], origin[1], origin[2]), rs::vec3(direction[0], direction[1], direction[2]));
    rs::vec3 hit_min;
    rs::HitRecord hit;
    hit.t = 0.0f, hit.index = 0u;
    const bool got = rs::GridHit(ray, 0.00001f, rs::infinity, hit_min, hit);
    record_hit(got, hit, out);
    return got ? 1 : 0;
}

extern "C" float ref_trace_hash12(float px, float py) { return rs::hash12(rs::vec2(px, py)); }
extern "C" float ref_trace_rand2(float x, float y) { return rs::Rand(rs::vec2(x, y)); }
