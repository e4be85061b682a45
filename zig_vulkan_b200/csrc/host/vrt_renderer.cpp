// vrt_renderer.cpp — the VoxelRT facade (src/modules/VoxelRT.zig) on top of the C ABI of vrt.h.
// Everything that touches the device goes through vrt_* calls; this file has no CUDA in it.
#include <cstdio>
#include <cstring>
#include <new>

#include <chrono>

#include "vrt_host_internal.h"

struct vrt_renderer {
    vrt_hcam* camera = nullptr;
    vrt_hsun* sun = nullptr;
    vrt_grid* grid = nullptr;  // borrowed
    vrt_ctx* ctx = nullptr;    // Pipeline + ComputePipeline
    char err[512] = "no error";
};

namespace {
char g_create_error[512] = "no error";

int pass(vrt_renderer* r, int rc) {
    if (rc != VRT_OK) std::snprintf(r->err, sizeof(r->err), "%s", vrt_last_error(r->ctx));
    return rc;
}
}  // namespace

extern "C" {

void vrt_renderer_default_config(vrt_renderer_config* out) {  // VoxelRT.Config (VoxelRT.zig:22-28)
    if (!out) return;
    std::memset(out, 0, sizeof(*out));
    out->internal_resolution_width = 1280;
    out->internal_resolution_height = 720;
    out->material_buffer = 256;  // Pipeline.Config.material_buffer (Pipeline.zig:30)
    vrt_hcam_default_config(&out->camera);
    vrt_hsun_default_config(&out->sun);
}

// VoxelRT.init (VoxelRT.zig:39-70)
int vrt_renderer_create(vrt_renderer** out, vrt_grid* grid, const vrt_renderer_config* cfg_in) {
    if (!out) return VRT_E_INVALID;
    *out = nullptr;
    if (!grid) {
        std::snprintf(g_create_error, sizeof(g_create_error), "vrt_renderer_create: grid is NULL");
        return VRT_E_INVALID;
    }
    vrt_renderer_config cfg;
    if (cfg_in) cfg = *cfg_in;
    else vrt_renderer_default_config(&cfg);
    vrt_renderer* r = new (std::nothrow) vrt_renderer();
    if (!r) return VRT_E_OOM;
    r->grid = grid;
    r->camera = vrt_hcam_create(75.0f, cfg.internal_resolution_width, cfg.internal_resolution_height, &cfg.camera);  // :42
    r->sun = vrt_hsun_create(&cfg.sun);                                                                              // :46
    if (!r->camera || !r->sun) {
        std::snprintf(g_create_error, sizeof(g_create_error), "vrt_renderer_create: bad resolution or out of memory");
        vrt_renderer_destroy(r);
        return VRT_E_INVALID;
    }
    // Pipeline.init (Pipeline.zig:272-316): buffer capacities come from the host grid's array lengths
    vrt_config vc;
    std::memset(&vc, 0, sizeof(vc));
    vc.struct_size = sizeof(vc);
    vc.abi_version = VRT_ABI_VERSION;
    vc.width = cfg.internal_resolution_width;
    vc.height = cfg.internal_resolution_height;
    vc.brick_dim = grid->brick_dim;
    vc.material_capacity = cfg.material_buffer;
    vc.n_bricks = grid->brick_indices.size();
    vc.n_brick_alloc = grid->brick_alloc;
    vc.device = cfg.device;
    vc.flags = cfg.flags;
    vc.row_begin = cfg.row_begin, vc.row_end = cfg.row_end;
    int rc = vrt_init(&r->ctx, &vc);
    if (rc != VRT_OK) {
        std::snprintf(g_create_error, sizeof(g_create_error), "%s", vrt_last_error(nullptr));
        vrt_renderer_destroy(r);
        return rc;
    }
    rc = vrt_upload_grid_state(r->ctx, &grid->state);  // :62
    if (rc != VRT_OK) {
        std::snprintf(g_create_error, sizeof(g_create_error), "%s", vrt_last_error(r->ctx));
        vrt_renderer_destroy(r);
        return rc;
    }
    *out = r;
    return VRT_OK;
}

// VoxelRT.deinit (VoxelRT.zig:174-178); the grid stays with the caller
void vrt_renderer_destroy(vrt_renderer* r) {
    if (!r) return;
    vrt_deinit(r->ctx);
    vrt_hcam_destroy(r->camera);
    vrt_hsun_destroy(r->sun);
    delete r;
}

const char* vrt_renderer_last_error(const vrt_renderer* r) { return r ? r->err : g_create_error; }
vrt_hcam* vrt_renderer_camera(vrt_renderer* r) { return r ? r->camera : nullptr; }
vrt_hsun* vrt_renderer_sun(vrt_renderer* r) { return r ? r->sun : nullptr; }
vrt_ctx* vrt_renderer_ctx(vrt_renderer* r) { return r ? r->ctx : nullptr; }

int vrt_renderer_push_materials(vrt_renderer* r, const vrt_material* materials, size_t count) {  // VoxelRT.zig:85-87
    if (!r) return VRT_E_INVALID;
    return pass(r, vrt_upload_materials(r->ctx, 0, materials, count));
}

// VoxelRT.updateGridDelta (VoxelRT.zig:107-172)
int vrt_renderer_update_grid_delta(vrt_renderer* r) {
    if (!r) return VRT_E_INVALID;
    vrt_grid* g = r->grid;
    int rc;
    vrt_delta* d = &g->delta[VRT_DELTA_STATUSES];
    if (d->active) {
        if ((rc = pass(r, vrt_upload_brick_statuses(r->ctx, d->from, g->statuses.data() + d->from, d->to - d->from))) != VRT_OK) return rc;
        d->reset();
    }
    d = &g->delta[VRT_DELTA_BRICK_INDICES];
    if (d->active) {
        if ((rc = pass(r, vrt_upload_brick_indices(r->ctx, d->from, g->brick_indices.data() + d->from, d->to - d->from))) != VRT_OK) return rc;
        d->reset();
    }
    d = &g->delta[VRT_DELTA_OCCUPANCY];
    if (d->active) {
        if ((rc = pass(r, vrt_upload_brick_occupancy(r->ctx, d->from, g->occupancy.data() + d->from, d->to - d->from))) != VRT_OK) return rc;
        d->reset();
    }
    d = &g->delta[VRT_DELTA_START_INDICES];
    if (d->active) {
        if ((rc = pass(r, vrt_upload_brick_start_indices(r->ctx, d->from, g->start_indices.data() + d->from, d->to - d->from))) != VRT_OK) return rc;
        d->reset();
    }
    d = &g->delta[VRT_DELTA_MATERIAL_INDICES];
    if (d->active) {
        if ((rc = pass(r, vrt_upload_material_indices(r->ctx, d->from, g->material_indices.data() + d->from, d->to - d->from))) != VRT_OK) return rc;
        d->reset();
    }
    return VRT_OK;
}

void vrt_renderer_update_sun(vrt_renderer* r, float delta_time) {  // VoxelRT.zig:80-82
    if (r) vrt_hsun_update(r->sun, delta_time);
}

// VoxelRT.draw -> Pipeline.draw -> ComputePipeline.dispatch(camera.*, sun.*) (VoxelRT.zig:76-78, Pipeline.zig:441-446)
int vrt_renderer_draw(vrt_renderer* r) {
    if (!r) return VRT_E_INVALID;
    return pass(r, vrt_trace(r->ctx, &r->camera->d_camera, &r->sun->device_data));
}

int vrt_renderer_draw_to_host(vrt_renderer* r, uint8_t* rgba8_host, size_t bytes) {
    if (!r) return VRT_E_INVALID;
    return pass(r, vrt_trace_to_host(r->ctx, &r->camera->d_camera, &r->sun->device_data, rgba8_host, bytes));
}

int vrt_renderer_present_to_host(vrt_renderer* r, const vrt_denoise_params* params, uint32_t out_width, uint32_t out_height, uint32_t flags,
                                 uint8_t* host, size_t bytes) {
    if (!r) return VRT_E_INVALID;
    const vrt_denoise_params defaults = {20, 0.6f, 1.5f, 20.0f};  // GraphicsPipeline.Config (GraphicsPipeline.zig:34-39)
    int rc = pass(r, vrt_trace(r->ctx, &r->camera->d_camera, &r->sun->device_data));
    if (rc != VRT_OK) return rc;
    rc = pass(r, vrt_denoise(r->ctx, params ? params : &defaults, out_width, out_height, flags));
    if (rc != VRT_OK) return rc;
    return pass(r, vrt_read_denoised(r->ctx, host, bytes));
}

// main.zig's benchmark mode: `while (!benchmark.update(dt)) draw`, dt = the frame's wall time
int vrt_renderer_run_benchmark(vrt_renderer* r, float duration_s, float extent_scale, vrt_benchmark_report* out) {
    if (!r || !out) return VRT_E_INVALID;
    vrt_benchmark* b = vrt_benchmark_create(r->camera, r->grid, r->sun->device_data.enabled != 0u, duration_s, extent_scale);
    if (!b) return VRT_E_INVALID;
    int rc = VRT_OK;
    for (;;) {
        const auto t0 = std::chrono::steady_clock::now();
        rc = pass(r, vrt_trace(r->ctx, &r->camera->d_camera, &r->sun->device_data));
        if (rc == VRT_OK) rc = pass(r, vrt_sync(r->ctx));
        if (rc != VRT_OK) break;
        const float dt = std::chrono::duration<float>(std::chrono::steady_clock::now() - t0).count();
        if (vrt_benchmark_update(b, dt)) break;
    }
    vrt_benchmark_get_report(b, out);
    vrt_benchmark_destroy(b);
    vrt_hcam_enable_input(r->camera);
    return rc;
}

}  // extern "C"
