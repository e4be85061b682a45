/*
 * ref_shader.h — C API of oracle/_ref/libref_shader.so: the reference's own shaders compiled for the host (ref_trace.cpp,
 * ref_present.cpp).  TEST INFRASTRUCTURE, NOT PRODUCT.
 */
#ifndef REF_SHADER_H
#define REF_SHADER_H

#include "../vrt_oracle.h"

#ifdef __cplusplus
extern "C" {
#endif

/* HitRecord of the primary ray (brick_raytracer.comp:48-53) + GridHit's return value */
typedef struct ref_hit {
    uint32_t hit;
    uint32_t index; /* material index */
    float t;
    float point[3];
    float normal[3];
} ref_hit;

/* brick_raytracer.comp main() for every pixel of rows [row_begin, row_end) of the dispatch; rgba8 = the whole storage image
 * (width*height*4); hits (nullable, width*height) = the shader's GridHit on each pixel's sample-0 camera ray.
 * threads <= 0: all host threads.  -2: image narrower than 2 pixels (the shader divides by width-1). */
int ref_trace_render(const orc_scene* scene, const vrt_camera* camera, const vrt_sun* sun, uint32_t row_begin, uint32_t row_end,
                     uint8_t* rgba8, ref_hit* hits, int threads);
/* GridHit(CreateRay(origin, direction), 0.00001, infinity) (:271); returns 1 / 0 */
int ref_trace_grid_hit(const orc_scene* scene, const float origin[3], const float direction[3], ref_hit* out);
float ref_trace_hash12(float px, float py); /* rand.comp:22 */
float ref_trace_rand2(float x, float y);    /* rand.comp:4 */

/* image.frag main() for every fragment of an out_width x out_height target (uv = pixel centre / size, full-screen quad
 * GraphicsPipeline.zig:17-22); input = the R8G8B8A8_UNORM compute image through the linear / repeat sampler (Pipeline.zig:193-212);
 * out = RGBA8 (flags 0) or BGRA8 (VRT_DENOISE_BGRA, the swapchain format, swapchain.zig:235). */
int ref_present_render(const uint8_t* rgba8_in, uint32_t in_width, uint32_t in_height, const vrt_denoise_params* params, uint32_t out_width,
                       uint32_t out_height, uint32_t flags, uint8_t* out, int threads);

#ifdef __cplusplus
}
#endif
#endif
