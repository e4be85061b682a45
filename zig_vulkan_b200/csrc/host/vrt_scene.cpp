// vrt_scene.cpp — scene producers for the benchmark/parity configs and the reference's benchmark fly-through.
//
// The reference fills its grid from value-noise terrain (terrain/terrain.zig:43-127) seeded with 420
// (main.zig:120), but that generator is not reproducible (it reads a dangling std.Random and races on it across
// threads).  The synthetic scene below mirrors its SHAPE — a height field filled for
// y in [h/2, h) with grass/dirt/rock bands by height, water up to an ocean level, plus a few metal spheres so
// every material type is on screen — using 32-bit integer arithmetic only, so the same voxels come out on every
// compiler and CPU.
#include <cmath>
#include <cstring>
#include <limits>
#include <new>

#include "vrt_host_internal.h"

using namespace vrt_host;

namespace {

// lowbias32-style integer finaliser
inline uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
inline uint32_t hash3(uint32_t a, uint32_t b, uint32_t c, uint32_t seed) {
    return mix32(a * 0x9e3779b1u ^ mix32(b * 0x85ebca77u ^ mix32(c * 0xc2b2ae3du ^ seed)));
}

// value noise on an integer lattice of `period` voxels, bilinear in 16.16 fixed point; returns [0, 65535]
uint32_t lattice_noise(uint32_t x, uint32_t z, uint32_t period, uint32_t seed) {
    const uint32_t ix = x / period, iz = z / period;
    const uint32_t fx = ((x % period) << 16) / period, fz = ((z % period) << 16) / period;
    const uint32_t v00 = hash3(ix, iz, 0, seed) & 0xffffu, v10 = hash3(ix + 1, iz, 0, seed) & 0xffffu;
    const uint32_t v01 = hash3(ix, iz + 1, 0, seed) & 0xffffu, v11 = hash3(ix + 1, iz + 1, 0, seed) & 0xffffu;
    const uint64_t top = (uint64_t)v00 * (65536u - fx) + (uint64_t)v10 * fx;  // 32.16
    const uint64_t bot = (uint64_t)v01 * (65536u - fx) + (uint64_t)v11 * fx;
    return (uint32_t)((top * (65536u - fz) + bot * fz) >> 32);
}

struct Sphere {
    int64_t cx, cy, cz, r;
};

int emit_to_grid(void* user, uint32_t x, uint32_t y, uint32_t z, uint8_t material) {
    return vrt_grid_insert(static_cast<vrt_grid*>(user), x, y, z, material);
}

// Benchmark.Configuration (Benchmark.zig:141-173)
constexpr float kBenchDuration = 60.0f;
const Vec3 kBenchPoints[VRT_BENCH_PATH_POINTS] = {{0, 0, 0},      {2, 5, 0},      {3, 5, 5},      {5, 2, 1},    {10, 0, 10}, {20, -20, 20},
                                                  {10, -25, 15}, {10, -22, 20}, {10, -30, 25}, {5, -10, 10}, {0, 13, 0}};
const Vec3 kBenchEulers[VRT_BENCH_PATH_POINTS] = {{0, 0, 0},    {0, 45, 0},   {10, -20, 0}, {20, 180, 0}, {50, 90, 0}, {60, 0, 0},
                                                  {80, -10, 0}, {75, -40, 0}, {80, -10, 0}, {80, -90, 0}, {0, -145, 0}};

}  // namespace

extern "C" {

// terrain/terrain.zig:130-196
uint32_t vrt_scene_terrain_materials(vrt_material* out, uint32_t capacity) {
    static const vrt_material table[8] = {
        {VRT_MAT_DIELECTRIC, 0.117f, 0.45f, 0.85f, 1.333f},  // water
        {VRT_MAT_LAMBERTIAN, 0.0f, 0.6f, 0.0f, 0.0f},        // grass 1
        {VRT_MAT_LAMBERTIAN, 0.0f, 0.5019f, 0.0f, 0.0f},     // grass 2
        {VRT_MAT_LAMBERTIAN, 0.301f, 0.149f, 0.0f, 0.0f},    // dirt 1
        {VRT_MAT_LAMBERTIAN, 0.4f, 0.2f, 0.0f, 0.0f},        // dirt 2
        {VRT_MAT_LAMBERTIAN, 0.275f, 0.275f, 0.275f, 0.0f},  // rock 1
        {VRT_MAT_LAMBERTIAN, 0.225f, 0.225f, 0.225f, 0.0f},  // rock 2
        {VRT_MAT_METAL, 0.6f, 0.337f, 0.282f, 0.45f},        // iron
    };
    const uint32_t n = capacity < 8 ? capacity : 8;
    if (out) std::memcpy(out, table, n * sizeof(vrt_material));
    return 8;
}

int vrt_scene_synthetic(uint32_t n, uint32_t seed, vrt_emit_fn emit, void* user) { return vrt_scene_synthetic_box(n, n, n, seed, emit, user); }

// The same scene in an nx x ny x nz voxel box (the reference's default grid is 512 x 256 x 512, main.zig:77-81): heights scale with
// ny (terrain.zig:81 takes them from voxel_dim_y), lattice periods with nx.  nx = ny = nz gives vrt_scene_synthetic's scene exactly.
int vrt_scene_synthetic_box(uint32_t nx, uint32_t ny, uint32_t nz, uint32_t seed, vrt_emit_fn emit, void* user) {
    if (!emit || nx < 16 || nx > 4096 || ny < 16 || ny > 4096 || nz < 16 || nz > 4096) return -1;
    const uint32_t terrain_max = ny / 2;  // terrain.zig:81: voxel_dim_y * 0.5
    const uint32_t floor_h = ny / 16;
    const uint32_t ocean = ny / 8;        // main.zig:120 passes ocean_level 20 of 256 (~1/12); 1/8 here so lakes are visible
    const uint32_t p1 = nx / 4 > 4 ? nx / 4 : 4, p2 = nx / 16 > 2 ? nx / 16 : 2;
    for (uint32_t x = 0; x < nx; x++) {
        for (uint32_t z = 0; z < nz; z++) {
            const uint32_t v = (3u * lattice_noise(x, z, p1, seed) + lattice_noise(x, z, p2, seed ^ 0x5bd1e995u)) >> 2;  // [0, 65535]
            const uint32_t height = floor_h + (uint32_t)(((uint64_t)v * (terrain_max - floor_h)) >> 16);
            uint32_t y = height / 2;  // terrain.zig:97
            for (; y < height; y++) {
                // terrain.zig:99-102: lerp(1, 3.4, y / terrain_max) + rng * 0.5 -> class 1..3, then one of 2 variants
                const uint32_t h = hash3(x, y, z, seed ^ 0x1b873593u);
                const uint32_t val = 256u + (y * 614u) / terrain_max + (h & 127u);  // 8.8 fixed point
                const uint32_t cls = val >> 8;                                       // 1 grass, 2 dirt, 3 rock
                const uint8_t material = (uint8_t)(1u + (cls - 1u) * 2u + ((h >> 8) & 1u));
                const int rc = emit(user, x, y, z, material);
                if (rc) return rc;
            }
            for (; y < ocean; y++) {  // terrain.zig:105-107
                const int rc = emit(user, x, y, z, 0);
                if (rc) return rc;
            }
        }
    }
    // K solid iron spheres floating above the terrain (below the pose-0 camera height of 0.625 n)
    const uint32_t k_spheres = nx / 32 > 1 ? nx / 32 : 1;
    for (uint32_t k = 0; k < k_spheres; k++) {
        Sphere s;
        s.cx = hash3(k, 1, 0, seed) % nx;
        s.cz = hash3(k, 2, 0, seed) % nz;
        s.cy = ny / 4 + hash3(k, 3, 0, seed) % (ny / 4);
        s.r = ny / 32 + hash3(k, 4, 0, seed) % (ny / 32 + 1);
        for (int64_t x = s.cx - s.r; x <= s.cx + s.r; x++) {
            if (x < 0 || x >= (int64_t)nx) continue;
            for (int64_t z = s.cz - s.r; z <= s.cz + s.r; z++) {
                if (z < 0 || z >= (int64_t)nz) continue;
                for (int64_t y = s.cy - s.r; y <= s.cy + s.r; y++) {
                    if (y < 0 || y >= (int64_t)ny) continue;
                    const int64_t dx = x - s.cx, dy = y - s.cy, dz = z - s.cz;
                    if (dx * dx + dy * dy + dz * dz > s.r * s.r) continue;
                    const int rc = emit(user, (uint32_t)x, (uint32_t)y, (uint32_t)z, 7);
                    if (rc) return rc;
                }
            }
        }
    }
    return 0;
}

int vrt_scene_synthetic_fill(vrt_grid* g, uint32_t seed) {
    if (!g) return -1;
    const vrt_grid_state& s = g->state;
    return vrt_scene_synthetic_box(s.voxel_dim_x, s.voxel_dim_y, s.voxel_dim_z, seed, emit_to_grid, g);
}

// Benchmark.Configuration (Benchmark.zig:141-173) evaluated the way Benchmark.update does (:50-66), with
// timer = t * benchmark_duration: way points lerp, orientations lerp (the reference lerps the quaternions without
// renormalising; Camera.orientation normalises afterwards).
void vrt_bench_path_pose(float t, float extent_scale, float origin_out[3], float yaw_wxyz_out[4]) {
    if (t < 0.0f) t = 0.0f;
    if (t > 1.0f) t = 1.0f;
    const float duration = kBenchDuration;
    const float timer = t * duration;
    const float fraction = duration / (float)VRT_BENCH_PATH_POINTS;  // path_point_fraction == path_orientation_fraction
    size_t index = (size_t)std::floor(timer / fraction);
    Vec3 origin = kBenchPoints[VRT_BENCH_PATH_POINTS - 1];
    Quat yaw = from_euler(kBenchEulers[VRT_BENCH_PATH_POINTS - 1]);
    if (index < VRT_BENCH_PATH_POINTS - 1) {
        const float pos = std::fmod(timer, fraction) / fraction;
        origin = lerp(kBenchPoints[index], kBenchPoints[index + 1], pos);
        const Quat l = index == 0 ? kIdentity : from_euler(kBenchEulers[index]);
        yaw = qlerp(l, from_euler(kBenchEulers[index + 1]), pos);
    }
    if (origin_out) origin_out[0] = origin.x * extent_scale, origin_out[1] = origin.y * extent_scale, origin_out[2] = origin.z * extent_scale;
    if (yaw_wxyz_out) yaw_wxyz_out[0] = yaw.w, yaw_wxyz_out[1] = yaw.x, yaw_wxyz_out[2] = yaw.y, yaw_wxyz_out[3] = yaw.z;
}

// Benchmark.init (Benchmark.zig:22-44)
vrt_benchmark* vrt_benchmark_create(vrt_hcam* camera, const vrt_grid* grid, int sun_enabled, float duration_s, float extent_scale) {
    if (!camera || !(duration_s > 0.0f)) return nullptr;
    vrt_benchmark* b = new (std::nothrow) vrt_benchmark();
    if (!b) return nullptr;
    b->camera = camera, b->sun_enabled = sun_enabled != 0, b->timer = 0.0f, b->duration = duration_s, b->extent_scale = extent_scale;
    b->fraction = duration_s / (float)VRT_BENCH_PATH_POINTS;  // path_point_fraction == path_orientation_fraction (:23-24)
    b->min_dt = std::numeric_limits<float>::max(), b->max_dt = 0.0f, b->dt_sum = 0.0f, b->samples = 0;  // Report.init (:88-101)
    b->voxel_dim[0] = b->voxel_dim[1] = b->voxel_dim[2] = 0;
    if (grid) b->voxel_dim[0] = grid->state.voxel_dim_x, b->voxel_dim[1] = grid->state.voxel_dim_y, b->voxel_dim[2] = grid->state.voxel_dim_z;
    vrt_hcam_disable_input(camera);  // :27
    const float origin[3] = {kBenchPoints[0].x * extent_scale, kBenchPoints[0].y * extent_scale, kBenchPoints[0].z * extent_scale};
    const float yaw[4] = {1.0f, 0.0f, 0.0f, 0.0f};
    vrt_hcam_set_origin(camera, origin);             // :28
    vrt_hcam_set_orientation(camera, yaw, nullptr);  // :30-32 yaw = orientations[0] = identity, pitch = identity
    return b;
}

void vrt_benchmark_destroy(vrt_benchmark* b) { delete b; }

// Benchmark.update (Benchmark.zig:47-75): returns 1 when the fly-through has completed
int vrt_benchmark_update(vrt_benchmark* b, float dt) {
    if (!b) return 1;
    b->timer += dt;
    const size_t index = (size_t)std::floor(b->timer / b->fraction);  // same index for points and orientations
    if (index < VRT_BENCH_PATH_POINTS - 1) {                           // the last segment keeps the pose it ended with (:51,:59)
        const float pos = std::fmod(b->timer, b->fraction) / b->fraction;
        const Vec3 o = lerp(kBenchPoints[index], kBenchPoints[index + 1], pos);
        const float origin[3] = {o.x * b->extent_scale, o.y * b->extent_scale, o.z * b->extent_scale};
        const Quat l = index == 0 ? kIdentity : from_euler(kBenchEulers[index]);
        const Quat q = qlerp(l, from_euler(kBenchEulers[index + 1]), pos);  // lerp without renormalising, as upstream (:63)
        const float yaw[4] = {q.w, q.x, q.y, q.z};
        vrt_hcam_set_origin(b->camera, origin);
        vrt_hcam_set_orientation(b->camera, yaw, nullptr);
    }
    if (dt < b->min_dt) b->min_dt = dt;  // :69-72
    if (dt > b->max_dt) b->max_dt = dt;
    b->dt_sum += dt;
    b->samples += 1;
    return b->timer >= b->duration ? 1 : 0;
}

// Benchmark.Report (Benchmark.zig:81-139) as numbers instead of a log line
void vrt_benchmark_get_report(const vrt_benchmark* b, vrt_benchmark_report* out) {
    if (!b || !out) return;
    out->min_frame_ms = b->samples ? b->min_dt * 1000.0f : 0.0f;
    out->max_frame_ms = b->max_dt * 1000.0f;
    out->avg_frame_ms = b->samples ? b->dt_sum / (float)b->samples * 1000.0f : 0.0f;
    out->frames = b->samples;
    for (int i = 0; i < 3; i++) out->voxel_dim[i] = b->voxel_dim[i];
    out->sun_enabled = b->sun_enabled ? 1u : 0u;
    out->image_width = b->camera->d_camera.image_width, out->image_height = b->camera->d_camera.image_height;
    out->max_bounce = b->camera->d_camera.max_bounce, out->samples_per_pixel = b->camera->d_camera.samples_per_pixel;
}

}  // extern "C"
