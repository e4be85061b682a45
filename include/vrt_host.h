/*
 * vrt_host.h — C API of libvrt_host.so: the host side ABOVE the C ABI of vrt.h, i.e. the C++ restatement of
 * the reference's Zig host objects that feed the ray-tracing path.  (The reference's host language is Zig; no
 * Zig toolchain exists in this environment, so the host side is C++ — see INTEGRATION.md for the .zig binding.)
 * It is a plain client of libvrt.so: everything it does to the device goes through vrt_* calls.
 *
 *   vrt_grid_*      <- BrickGrid            src/modules/voxel_rt/brick/Grid.zig, State.zig, MaterialAllocator.zig
 *   vrt_hcam_*      <- Camera               src/modules/voxel_rt/Camera.zig
 *   vrt_hsun_*      <- Sun                  src/modules/voxel_rt/Sun.zig
 *   vrt_renderer_*  <- VoxelRT facade       src/modules/VoxelRT.zig (init/draw/pushMaterials/updateGridDelta)
 *   vrt_scene_*     <- scene producers      terrain materials (terrain/terrain.zig:130-196) + the seeded synthetic
 *                                           scene of the benchmark configs (integer-only; DESIGN.md "Synthetic scene")
 *   vrt_bench_path_* <- benchmark fly-through  src/modules/voxel_rt/Benchmark.zig:141-173
 */
#ifndef VRT_HOST_H
#define VRT_HOST_H

#include "vrt.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ BrickGrid */
typedef struct vrt_grid vrt_grid;

/* BrickGrid.init (Grid.zig:36-114). brick_alloc 0 = all bricks (Grid.zig:51). NULL on bad arguments / OOM. */
vrt_grid* vrt_grid_create(uint32_t dim_x, uint32_t dim_y, uint32_t dim_z, uint32_t brick_dim, uint64_t brick_alloc,
                          const float min_point[3], float scale, float base_t);
void vrt_grid_destroy(vrt_grid* g);
/* BrickGrid.insert (Grid.zig:129-194): y is flipped. 0 ok, -1 out of range (the reference asserts), -2 capacity. */
int vrt_grid_insert(vrt_grid* g, uint32_t x, uint32_t y, uint32_t z, uint8_t material);
/* n voxels as packed {x,y,z,material} uint32 quadruples; stops at the first failure and returns its code. */
int vrt_grid_insert_many(vrt_grid* g, const uint32_t* xyzm, size_t n);
uint32_t vrt_grid_active_bricks(const vrt_grid* g);
uint32_t vrt_grid_brick_dim(const vrt_grid* g);
uint64_t vrt_grid_brick_alloc(const vrt_grid* g);
void vrt_grid_get_state(const vrt_grid* g, vrt_grid_state* out);
const uint32_t* vrt_grid_statuses(const vrt_grid* g, uint64_t* count);
const uint32_t* vrt_grid_brick_indices(const vrt_grid* g, uint64_t* count);
const uint8_t* vrt_grid_occupancy(const vrt_grid* g, uint64_t* count);
const uint32_t* vrt_grid_start_indices(const vrt_grid* g, uint64_t* count);
const uint8_t* vrt_grid_material_indices(const vrt_grid* g, uint64_t* count);

/* DeviceDataDelta (State.zig:14-57).  which: */
#define VRT_DELTA_STATUSES 0
#define VRT_DELTA_BRICK_INDICES 1
#define VRT_DELTA_OCCUPANCY 2
#define VRT_DELTA_START_INDICES 3
#define VRT_DELTA_MATERIAL_INDICES 4
/* 1 if the delta is active (and [*from,*to) is its element range), 0 if inactive, -1 bad `which`. */
int vrt_grid_delta_peek(vrt_grid* g, int which, uint64_t* from, uint64_t* to);
void vrt_grid_delta_reset(vrt_grid* g, int which); /* resetDelta (State.zig:33-37) */

/* ------------------------------------------------------------------ Camera */
typedef struct vrt_hcam vrt_hcam;

typedef struct vrt_hcam_config { /* Camera.Config (Camera.zig:5-14) */
    float viewport_height;       /* 2 */
    float origin[3];             /* 0,0,0 */
    int32_t samples_per_pixel;   /* 2 */
    int32_t max_bounce;          /* 2 (stored +1 on the device, Camera.zig:74) */
    float turn_rate;             /* 0.1 */
    float normal_speed;          /* 1 */
    float sprint_speed;          /* 2 */
    uint32_t user_input_disabled;
} vrt_hcam_config;

void vrt_hcam_default_config(vrt_hcam_config* out);
vrt_hcam* vrt_hcam_create(float vertical_fov_deg, uint32_t image_width, uint32_t image_height, const vrt_hcam_config* cfg);
void vrt_hcam_destroy(vrt_hcam* c);
void vrt_hcam_device(const vrt_hcam* c, vrt_camera* out);   /* d_camera */
void vrt_hcam_set_origin(vrt_hcam* c, const float origin[3]);
void vrt_hcam_translate(vrt_hcam* c, float delta_time, const float by[3]);
void vrt_hcam_turn_pitch(vrt_hcam* c, float angle);
void vrt_hcam_turn_yaw(vrt_hcam* c, float angle);
void vrt_hcam_reset(vrt_hcam* c);
void vrt_hcam_activate_sprint(vrt_hcam* c);
void vrt_hcam_disable_sprint(vrt_hcam* c);
void vrt_hcam_disable_input(vrt_hcam* c);
void vrt_hcam_enable_input(vrt_hcam* c);
/* What Benchmark.update does (Benchmark.zig:50-66): set yaw/pitch quaternions {w,x,y,z} and re-derive the basis. */
void vrt_hcam_set_orientation(vrt_hcam* c, const float yaw_wxyz[4], const float pitch_wxyz[4]);
/* Convenience: orientation from Euler angles in degrees (zalgebra Quat.fromEulerAngles) as yaw, identity pitch. */
void vrt_hcam_set_euler_deg(vrt_hcam* c, float x_deg, float y_deg, float z_deg);

/* ------------------------------------------------------------------ Sun */
typedef struct vrt_hsun vrt_hsun;

typedef struct vrt_hsun_config { /* Sun.Config (Sun.zig:4-11) */
    uint32_t animate;       /* 1 */
    float animate_speed;    /* 0.1 */
    uint32_t enabled;       /* 1 */
    float color[3];         /* 1, 1.1, 1 */
    float radius;           /* 5 */
    float sun_distance;     /* 1000 */
} vrt_hsun_config;

void vrt_hsun_default_config(vrt_hsun_config* out);
vrt_hsun* vrt_hsun_create(const vrt_hsun_config* cfg);
void vrt_hsun_destroy(vrt_hsun* s);
void vrt_hsun_device(const vrt_hsun* s, vrt_sun* out);   /* device_data */
void vrt_hsun_update(vrt_hsun* s, float delta_time);     /* Sun.update (Sun.zig:65-86) */

/* ------------------------------------------------------------------ VoxelRT facade */
typedef struct vrt_renderer vrt_renderer;

typedef struct vrt_renderer_config { /* VoxelRT.Config (VoxelRT.zig:22-28) */
    uint32_t internal_resolution_width;   /* 1280 */
    uint32_t internal_resolution_height;  /* 720 */
    uint32_t material_buffer;             /* Pipeline.Config.material_buffer = 256 */
    vrt_hcam_config camera;
    vrt_hsun_config sun;
    int32_t device;
    uint32_t flags;                       /* VRT_FLAG_* forwarded to vrt_init */
    uint32_t row_begin, row_end;          /* forwarded to vrt_init */
} vrt_renderer_config;

void vrt_renderer_default_config(vrt_renderer_config* out);
/* VoxelRT.init (VoxelRT.zig:39-70): camera(fov 75) + sun + pipeline, then transferGridState.
 * The renderer borrows `grid` (the reference "takes ownership" only nominally: deinit does not free it). */
int vrt_renderer_create(vrt_renderer** out, vrt_grid* grid, const vrt_renderer_config* cfg);
void vrt_renderer_destroy(vrt_renderer* r);
const char* vrt_renderer_last_error(const vrt_renderer* r);
vrt_hcam* vrt_renderer_camera(vrt_renderer* r);
vrt_hsun* vrt_renderer_sun(vrt_renderer* r);
vrt_ctx* vrt_renderer_ctx(vrt_renderer* r);
/* VoxelRT.pushMaterials (VoxelRT.zig:85-87) */
int vrt_renderer_push_materials(vrt_renderer* r, const vrt_material* materials, size_t count);
/* VoxelRT.updateGridDelta (VoxelRT.zig:107-172): upload the five dirty ranges, reset them. */
int vrt_renderer_update_grid_delta(vrt_renderer* r);
/* VoxelRT.updateSun (VoxelRT.zig:80-82) */
void vrt_renderer_update_sun(vrt_renderer* r, float delta_time);
/* VoxelRT.draw -> Pipeline.draw -> ComputePipeline.dispatch (VoxelRT.zig:76-78): enqueue one frame. */
int vrt_renderer_draw(vrt_renderer* r);
/* draw + copy the frame to host memory. */
int vrt_renderer_draw_to_host(vrt_renderer* r, uint8_t* rgba8_host, size_t bytes);
/* The graphics half of Pipeline.draw (Pipeline.zig:441-540): draw, then the present pass (image.frag, GraphicsPipeline.zig) into
 * an out_width x out_height image copied to host memory.  params NULL = GraphicsPipeline.Config defaults (:34-39);
 * flags = VRT_DENOISE_*. */
int vrt_renderer_present_to_host(vrt_renderer* r, const vrt_denoise_params* params, uint32_t out_width, uint32_t out_height, uint32_t flags,
                                 uint8_t* host, size_t bytes);

/* ------------------------------------------------------------------ scene producers */
/* The 8 terrain materials (terrain/terrain.zig:130-196): water, grass x2, dirt x2, rock x2, iron. */
uint32_t vrt_scene_terrain_materials(vrt_material* out, uint32_t capacity);

/* Seeded, integer-only synthetic terrain + metal spheres for an n_voxels^3 cube (DESIGN.md "Synthetic scene").
 * Every voxel is reported through `emit(user, x, y, z, material)` in x -> z -> y order; a non-zero return from
 * emit aborts and is returned.  vrt_scene_synthetic_fill is the common case emit = vrt_grid_insert. */
typedef int (*vrt_emit_fn)(void* user, uint32_t x, uint32_t y, uint32_t z, uint8_t material);
int vrt_scene_synthetic(uint32_t n_voxels, uint32_t seed, vrt_emit_fn emit, void* user);
/* The same scene in an nx x ny x nz voxel box (heights scale with ny as in terrain.zig:81); nx = ny = nz is vrt_scene_synthetic. */
int vrt_scene_synthetic_box(uint32_t nx, uint32_t ny, uint32_t nz, uint32_t seed, vrt_emit_fn emit, void* user);
int vrt_scene_synthetic_fill(vrt_grid* g, uint32_t seed);

/* ------------------------------------------------------------------ MagicaVoxel .vox reader
 * vox/loader.zig + vox/types.zig (VOX v150: MAIN / PACK / SIZE / XYZI / RGBA, unknown chunks skipped) and the two loops
 * of main.zig:96-118 that feed a model into the material table and the grid. */
typedef struct vrt_vox vrt_vox;
#define VRT_VOX_E_INVALID_ID (-10)            /* ParseError.InvalidId            (loader.zig:33)  */
#define VRT_VOX_E_EXPECTED_SIZE_HEADER (-11)  /* ParseError.ExpectedSizeHeader                    */
#define VRT_VOX_E_EXPECTED_XYZI_HEADER (-12)  /* ParseError.ExpectedXyziHeader                    */
#define VRT_VOX_E_EXPECTED_RGBA_HEADER (-13)  /* ParseError.ExpectedRgbaHeader                    */
#define VRT_VOX_E_UNEXPECTED_VERSION (-14)    /* ParseError.UnexpectedVersion                     */
#define VRT_VOX_E_INVALID_FILE_CONTENT (-15)  /* ParseError.InvalidFileContent (also: truncated)  */
#define VRT_VOX_E_IO (-16)                    /* file could not be opened / read                  */
int vrt_vox_validate_header(const uint8_t* buffer, size_t len);                       /* validateHeader, loader.zig:215-228 */
int vrt_vox_parse(vrt_vox** out, const uint8_t* buffer, size_t len, int strict);      /* parseBuffer,    loader.zig:41-197  */
int vrt_vox_load(vrt_vox** out, const char* path, int strict);                        /* load,           loader.zig:9-30    */
void vrt_vox_destroy(vrt_vox* v);
int32_t vrt_vox_num_models(const vrt_vox* v);
int vrt_vox_model_size(const vrt_vox* v, int32_t model, int32_t size_xyz[3]);
const uint8_t* vrt_vox_model_xyzi(const vrt_vox* v, int32_t model, uint64_t* count); /* {x,y,z,color_index} bytes */
const uint8_t* vrt_vox_palette(const vrt_vox* v);                                     /* 256 x {r,g,b,a} */
uint32_t vrt_vox_materials(const vrt_vox* v, vrt_material* materials, uint32_t capacity, uint32_t material_base);  /* main.zig:96-108 */
int vrt_vox_insert_into_grid(const vrt_vox* v, int32_t model, vrt_grid* grid, uint32_t off_x, uint32_t off_y, uint32_t off_z,
                             uint32_t material_base);                                 /* main.zig:110-118 */

/* Benchmark fly-through (Benchmark.zig:141-173): 11 way points x 11 orientations over 60 s.  t in [0,1] is the
 * normalised position along the path; offsets are scaled by `extent_scale` (1 = the reference's own units). */
#define VRT_BENCH_PATH_POINTS 11
void vrt_bench_path_pose(float t, float extent_scale, float origin_out[3], float yaw_wxyz_out[4]);

/* Benchmark (Benchmark.zig:22-139): drives a camera along that path by accumulated frame time and keeps the frame-time report.
 * duration_s = 60 is the reference's benchmark_duration; extent_scale as above. */
typedef struct vrt_benchmark vrt_benchmark;
typedef struct vrt_benchmark_report { /* Benchmark.Report + what Report.print logs (:103-139) */
    float min_frame_ms, max_frame_ms, avg_frame_ms;
    uint32_t frames;
    uint32_t voxel_dim[3];
    uint32_t sun_enabled;
    uint32_t image_width, image_height;
    int32_t max_bounce, samples_per_pixel;
} vrt_benchmark_report;
vrt_benchmark* vrt_benchmark_create(vrt_hcam* camera, const vrt_grid* grid /* nullable */, int sun_enabled, float duration_s, float extent_scale);
void vrt_benchmark_destroy(vrt_benchmark* b);
/* Benchmark.update: advance by one frame of dt seconds (moves the camera); returns 1 once the path has completed. */
int vrt_benchmark_update(vrt_benchmark* b, float dt);
void vrt_benchmark_get_report(const vrt_benchmark* b, vrt_benchmark_report* out);
/* The reference's benchmark mode (main.zig: VoxelRT.createBenchmark + update per frame): draw frames along the path, each
 * frame's measured wall time (enqueue + device + sync) is the next dt, until the path completes. */
int vrt_renderer_run_benchmark(vrt_renderer* r, float duration_s, float extent_scale, vrt_benchmark_report* out);

#ifdef __cplusplus
}
#endif
#endif /* VRT_HOST_H */
