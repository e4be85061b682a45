/* render_frame.c — the drop-in boundary used from plain C, the way the reference's main.zig drives its renderer
 * (main.zig:77-138,156-195): build a BrickGrid, hand it to the VoxelRT facade, push materials, ship the dirty ranges,
 * draw one frame, run the present pass, write both images as PPM.
 *
 *   cc -O2 -I. examples/render_frame.c -Lzig_vulkan_b200 -lvrt_host -lvrt -Wl,-rpath,$PWD/zig_vulkan_b200 -o render_frame
 *   ./render_frame 128 frame        (needs a CUDA device: libvrt has no CPU fallback)
 */
#include <stdio.h>
#include <stdlib.h>

#include "include/vrt.h"
#include "include/vrt_host.h"

static int write_ppm(const char* path, const uint8_t* rgba, uint32_t w, uint32_t h) {
    FILE* f = fopen(path, "wb");
    if (!f) return -1;
    fprintf(f, "P6\n%u %u\n255\n", w, h);
    for (size_t i = 0; i < (size_t)w * h; i++) fwrite(rgba + 4 * i, 1, 3, f);
    return fclose(f);
}

int main(int argc, char** argv) {
    const uint32_t n_voxels = argc > 1 ? (uint32_t)atoi(argv[1]) : 128u; /* cubic scene, world extent 64 */
    const char* stem = argc > 2 ? argv[2] : "frame";
    const uint32_t per_axis = n_voxels / 4u;
    const float min_point[3] = {-32.0f, -32.0f, -32.0f};

    vrt_grid* grid = vrt_grid_create(per_axis, per_axis, per_axis, 4, 0, min_point, 64.0f / (float)per_axis, 0.01f); /* BrickGrid.init */
    if (!grid || vrt_scene_synthetic_fill(grid, 420) != 0) return fprintf(stderr, "grid build failed\n"), 1;

    vrt_renderer_config cfg;
    vrt_renderer_default_config(&cfg); /* VoxelRT.Config: 1280x720, spp 2, max_bounce 2, sun on */
    vrt_renderer* r = NULL;
    if (vrt_renderer_create(&r, grid, &cfg) != VRT_OK) return fprintf(stderr, "vrt_renderer_create: %s\n", vrt_renderer_last_error(NULL)), 1;

    vrt_material materials[256] = {{0}};
    vrt_scene_terrain_materials(materials, 256);
    const float origin[3] = {0.0f, -10.0f, 28.0f};
    vrt_hcam_set_origin(vrt_renderer_camera(r), origin);
    vrt_hcam_set_euler_deg(vrt_renderer_camera(r), 25.0f, 0.0f, 0.0f);

    const uint32_t w = cfg.internal_resolution_width, h = cfg.internal_resolution_height;
    uint8_t* traced = malloc((size_t)w * h * 4);
    uint8_t* shown = malloc((size_t)w * h * 4);
    int rc = vrt_renderer_push_materials(r, materials, 256);                 /* VoxelRT.pushMaterials */
    if (rc == VRT_OK) rc = vrt_renderer_update_grid_delta(r);                /* VoxelRT.updateGridDelta */
    if (rc == VRT_OK) rc = vrt_renderer_draw_to_host(r, traced, (size_t)w * h * 4); /* VoxelRT.draw */
    if (rc == VRT_OK) rc = vrt_renderer_present_to_host(r, NULL, w, h, 0, shown, (size_t)w * h * 4); /* image.frag */
    if (rc != VRT_OK) return fprintf(stderr, "render failed: %s\n", vrt_renderer_last_error(r)), 1;

    char path[512];
    snprintf(path, sizeof path, "%s_traced.ppm", stem);
    write_ppm(path, traced, w, h);
    snprintf(path, sizeof path, "%s_presented.ppm", stem);
    write_ppm(path, shown, w, h);
    float ms = 0.0f;
    vrt_last_trace_ms(vrt_renderer_ctx(r), &ms);
    printf("%ux%u frame of a %u^3 scene: trace %.3f ms on the device\n", w, h, n_voxels, ms);

    free(traced), free(shown);
    vrt_renderer_destroy(r);
    vrt_grid_destroy(grid);
    return 0;
}
