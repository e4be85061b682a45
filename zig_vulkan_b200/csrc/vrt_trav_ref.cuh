// vrt_trav_ref.cuh — GridHit/BrickHit transliterated from the shader, memory accesses included: the one-word
// status cache (:321-326), one byte load per voxel step (:415), brick_indices -> occupancy -> start index ->
// material index dependent chain.  This is "the reference algorithm recompiled for sm_100a": the baseline the
// tuned traversal (vrt_trav_tuned.cuh) is measured against, and the kernel the AOV/counter path runs, since
// the request-byte model of DESIGN.md is defined on exactly these accesses.
#pragma once

#include "vrt_shade.cuh"

namespace vrt {

struct RefTraversal {
    // brick_raytracer.comp:378-471
    template <bool COUNT>
    static VRT_DI bool brick_hit(const TraceParams& P, const Ray& r, float t_max, V3 ray_delta, I3 ray_step, float g_scale,
                                 uint32_t brick_index, V3 brick_position, HitRecord& hit, TraceInfo& ti) {
        const float voxel_scale = g_scale * P.brick_voxel_scale;                                   // :389
        const unsigned long long solid_mask_base_index = (unsigned long long)brick_index * P.brick_bytes;  // :390
        const V3 fposition = (RayAt(r, hit.t) - brick_position) / v3s(voxel_scale);  // :393
        V3 side_dist = init_side_dist(fposition, ray_step, ray_delta);               // :394-395
        I3 pos = I3{(int)floorf(fposition.x), (int)floorf(fposition.y), (int)floorf(fposition.z)};  // :403
        const float local_t_max = t_max - hit.t;                                                   // :405
        float t_value = 0.0f;
        const int bd = P.brick_dim;
        while (pos.x >= 0 && pos.y >= 0 && pos.z >= 0 && pos.x < bd && pos.y < bd && pos.z < bd && t_value <= local_t_max) {
            if (COUNT) ti.voxel_steps++;
            const int voxel_index = pos.x + bd * (pos.z + bd * pos.y);  // :412
            const uint32_t mask_index = (bd <= 8) ? (uint32_t)(uint8_t)(voxel_index / 8) : (uint32_t)(voxel_index / 8);  // :413
            const uint32_t mask_offset = (uint32_t)(voxel_index % 8);
            const unsigned long long mask_at = solid_mask_base_index + mask_index;
            const uint32_t entry = mask_at < P.n_occupancy ? (uint32_t)__ldg(P.occupancy + mask_at) : 0u;  // :415
            if ((entry >> mask_offset) & 1u) {                                                           // :417
                const uint32_t sw = brick_index < P.n_start_indices ? __ldg(P.start_indices + brick_index) : 0u;
                const unsigned long long mi = (unsigned long long)(sw & 0x7fffffffu) + (uint32_t)voxel_index;  // :422
                hit.index = mi < P.n_material_indices ? (uint32_t)__ldg(P.material_indices + mi) : 0u;        // :425
                const vrt_material m = load_material(P, hit.index);
                const bool ignore_brick = (m.type == r.ignore_type_material) && (r.internal_reflection == m.type_data);  // :427
                if (!ignore_brick) {
                    const float t_offset = voxel_scale * 0.05f;           // :431
                    hit.t += t_value - t_offset;                          // :432
                    hit.point = RayAt(r, hit.t) + hit.normal * t_offset;  // :433
                    if (COUNT) ti.voxel_index = (uint32_t)voxel_index;
                    return true;
                }
            }
            const int axis = dda_step(side_dist, ray_delta, pos, ray_step, voxel_scale, t_value);  // :440-467
            hit.normal = axis_normal(axis, ray_step);
        }
        return false;
    }

    // brick_raytracer.comp:271-376 with t_min = 0.00001, t_max = infinity (the only way it is called, :218,247)
    template <bool COUNT>
    static VRT_DI bool grid_hit(const TraceParams& P, const Ray& r, HitRecord& hit, TraceInfo& ti) {
        const V3 g_min = v3(P.grid.min_point_base_t[0], P.grid.min_point_base_t[1], P.grid.min_point_base_t[2]);
        const V3 g_max = v3(P.grid.max_point_scale[0], P.grid.max_point_scale[1], P.grid.max_point_scale[2]);
        const float g_scale = P.grid.max_point_scale[3];
        const I3 brick_dim = I3{(int)P.grid.dim_x, (int)P.grid.dim_y, (int)P.grid.dim_z};

        // A non-finite direction has ray_step 0 on every axis and would never leave the loop: treated as a miss
        // (DESIGN.md "Deviations"; the oracle does the same).
        if (isnan((r.direction.x + r.direction.y) + r.direction.z)) return false;
        const V3 inv_ray_dir = v3(safeInverse(r.direction.x), safeInverse(r.direction.y), safeInverse(r.direction.z));  // :278
        float grid_t_min = 0.00001f;
        float grid_t_max = __int_as_float(0x7f800000);
        if (!AdvNormIntersect(g_min, g_max, r, inv_ray_dir, hit.normal, grid_t_min, grid_t_max)) return false;  // :282

        const float global_t_value = grid_t_min + 0.0001f * g_scale;  // :287
        const V3 ray_delta = v3(fabsf(inv_ray_dir.x), fabsf(inv_ray_dir.y), fabsf(inv_ray_dir.z));  // :290
        const I3 ray_step = I3{(int)gsign(r.direction.x), (int)gsign(r.direction.y), (int)gsign(r.direction.z)};  // :291
        const V3 fposition = (RayAt(r, global_t_value) - g_min) / v3s(g_scale);  // :293-296
        V3 side_dist = init_side_dist(fposition, ray_step, ray_delta);           // :297-298

        uint32_t brick_type_index = ~0u;  // :301
        uint32_t brick_bits = 0u;
        float t_value = 0.0f;
        I3 pos = I3{(int)floorf(fposition.x), (int)floorf(fposition.y), (int)floorf(fposition.z)};  // :311

        // :313-317 — `global_t_value <= t_max` with t_max = +inf is always true for finite t and dropped.
        while (pos.x >= 0 && pos.y >= 0 && pos.z >= 0 && pos.x < brick_dim.x && pos.y < brick_dim.y && pos.z < brick_dim.z) {
            if (COUNT) ti.grid_steps++;
            const uint32_t grid_index = (uint32_t)(pos.x + brick_dim.x * (pos.z + brick_dim.z * pos.y));  // :318
            const uint32_t new_brick_type_index = grid_index / 32u;                                      // :321
            if (brick_type_index != new_brick_type_index) {
                brick_bits = new_brick_type_index < P.n_statuses ? __ldg(P.statuses + new_brick_type_index) : 0u;  // :324
                brick_type_index = new_brick_type_index;
                if (COUNT) ti.status_fetches++;
            }
            if (brick_bits & (1u << (grid_index % 32u))) {  // :328
                const V3 brick_min = fma3(tofloat(pos), v3s(g_scale), g_min);  // :331
                hit.t = (t_value + grid_t_min) + 0.01f * g_scale;              // :332-334
                const uint32_t brick_index = grid_index < P.n_brick_indices ? __ldg(P.brick_indices + grid_index) : 0u;  // :337
                if (COUNT) ti.bricks_entered++;
                if (brick_hit<COUNT>(P, r, grid_t_max, ray_delta, ray_step, g_scale, brick_index, brick_min, hit, ti)) {
                    if (COUNT) ti.grid_index = grid_index;
                    return true;
                }
            }
            const int axis = dda_step(side_dist, ray_delta, pos, ray_step, g_scale, t_value);  // :345-372
            hit.normal = axis_normal(axis, ray_step);
        }
        return false;
    }

    template <bool COUNT>
    static VRT_DI bool shadow_hit(const TraceParams& P, const Ray& r, HitRecord& hit, TraceInfo& ti) {
        return grid_hit<COUNT>(P, r, hit, ti);
    }
};

}  // namespace vrt
