/*
 * vrt_oracle.h — C API of the CPU oracle.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline leg / --impl reference) may load this
 * library.  Nothing under zig_vulkan_b200/ links, imports or executes it.
 *
 * PARITY PINNED TO THE REFERENCE'S TEXT: the reference (Avokadoen/zig_vulkan @ 176598f) implements this path only as a
 * GLSL compute shader, ships no CPU path, no golden images and no tests for it, and its own build (zig + glslang + Vulkan)
 * cannot run here.  But the shader source compiles: oracle/_ref/libref_shader.so is assets/shaders/brick_raytracer.comp +
 * rand.comp (and image.frag) THEMSELVES, built by g++ under oracle/ref_shim/glsl_compat.h after a purely lexical pass
 * (oracle/ref_shim/translate.py).  tests/test_ref_shader.py requires this restatement to equal that library bit for bit —
 * RGBA8 and hit records, every golden case, random rays, the RNG-driven modes, a full 1080p C3 frame in bench.py.  What
 * stays a choice of this repo is only what GLSL leaves to the driver (the precision of normalize / sin / pow, FMA
 * contraction, UNORM rounding): the FP discipline listed in vrt_oracle.cpp, which glsl_compat.h implements identically.
 *
 * Struct layouts come from include/vrt.h (the public ABI, itself checked against the reference's
 * `extern struct`s by tests/test_abi.py).
 */
#ifndef VRT_ORACLE_H
#define VRT_ORACLE_H

#include "../include/vrt.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- grid builder: restates brick/Grid.zig:36-211 + brick/MaterialAllocator.zig:34-43 ---- */
typedef struct orc_grid orc_grid;

orc_grid* orc_grid_create(uint32_t dim_x, uint32_t dim_y, uint32_t dim_z, uint32_t brick_dim,
                          uint64_t brick_alloc /* 0 = all bricks (Grid.zig:51) */,
                          const float min_point[3], float scale, float base_t);
void orc_grid_destroy(orc_grid* g);
/* Grid.insert (Grid.zig:129-194).  0 ok; -1 coordinate out of range; -2 out of brick / material capacity. */
int orc_grid_insert(orc_grid* g, uint32_t x, uint32_t y, uint32_t z, uint8_t material);
/* n packed {x,y,z,material} uint32 quadruples; stops at the first failure. */
int orc_grid_insert_many(orc_grid* g, const uint32_t* xyzm, size_t n);
uint32_t orc_grid_active_bricks(const orc_grid* g);
void orc_grid_get_state(const orc_grid* g, vrt_grid_state* out);
const uint32_t* orc_grid_statuses(const orc_grid* g, uint64_t* count);
const uint32_t* orc_grid_brick_indices(const orc_grid* g, uint64_t* count);
const uint8_t* orc_grid_occupancy(const orc_grid* g, uint64_t* count);
const uint32_t* orc_grid_start_indices(const orc_grid* g, uint64_t* count);
const uint8_t* orc_grid_material_indices(const orc_grid* g, uint64_t* count);

/* ---- the scene the kernel sees: the UBO + six SSBOs of brick_raytracer.comp:79-134 ---- */
typedef struct orc_scene {
    vrt_grid_state state;
    uint32_t brick_dim;
    uint32_t n_materials;
    const vrt_material* materials;
    const uint32_t* statuses;        uint64_t n_statuses;
    const uint32_t* brick_indices;   uint64_t n_brick_indices;
    const uint8_t* occupancy;        uint64_t n_occupancy;
    const uint32_t* start_indices;   uint64_t n_start_indices;
    const uint8_t* material_indices; uint64_t n_material_indices;
} orc_scene;

void orc_scene_from_grid(const orc_grid* g, const vrt_material* materials, uint32_t n_materials, orc_scene* out);

/* Restates brick_raytracer.comp main() for image rows [row_begin,row_end).  rgba8 is the FULL image buffer
 * (width*height*4); aov (nullable) is full-image sized too; counters (nullable) are totals over the rows.
 * threads <= 0 -> hardware_concurrency. */
int orc_render(const orc_scene* scene, const vrt_camera* camera, const vrt_sun* sun,
               uint32_t row_begin, uint32_t row_end,
               uint8_t* rgba8, vrt_aov* aov, vrt_counters* counters, int threads);

/* One GridHit (brick_raytracer.comp:271-376) of CreateRay(origin, direction) with t_min=1e-5, t_max=+inf.
 * Returns 1 on hit, 0 on miss; `out` gets the primary fields of vrt_aov. */
int orc_grid_hit(const orc_scene* scene, const float origin[3], const float direction[3], vrt_aov* out);

/* The FP32 sine both sides use for the shader's sin-hash (rand.comp:3-4); exposed for tests. */
float orc_sinf(float x);
/* rand.comp:22-26 */
float orc_hash12(float px, float py);

/* ---- post-process pass: restates assets/shaders/image.frag:31-79 (vrt_oracle_denoise.cpp) ---- */
int orc_denoise(const uint8_t* rgba8_in, uint32_t in_width, uint32_t in_height, const vrt_denoise_params* params, uint32_t out_width,
                uint32_t out_height, uint32_t flags /* VRT_DENOISE_* */, uint8_t* out, int threads);
/* The pow() both sides use for that pass (after the shader's max(a, 0) macro); exposed for tests. */
float orc_pow(float a, float b);

#ifdef __cplusplus
}
#endif
#endif
