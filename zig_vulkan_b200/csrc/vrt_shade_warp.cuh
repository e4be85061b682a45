// vrt_shade_warp.cuh — main() + RayColor of the reference shader (brick_raytracer.comp:153-265) in warp-uniform form:
// all 32 lanes of a warp walk the sample / bounce loops together (a lane whose pixel is outside the image or whose
// path has ended simply carries active = false), so that the traversal calls inside are warp-cooperative
// (vrt_trav_warp.cuh).  Per lane, the sequence of operations is exactly the one of vrt_shade.cuh::RayColor.
#pragma once

#include "vrt_trav_warp.cuh"

namespace vrt {

template <int BD, bool AOV>
VRT_DI V3 ray_color_warp(const TraceParams& P, const Ray& r, bool lane_on, vrt_aov* aov, PixelCounters& pc, bool record_aov) {
    const bool sun_enabled = P.sun.enabled > 0;
    const V3 sun_color = ld3(P.sun.color);
    const int max_bounce = P.cam.max_bounce;

    HitRecord hit;
    hit.point = v3s(0.0f), hit.normal = v3s(0.0f), hit.t = 0.0f, hit.index = 0u;
    Ray current_ray = r;
    int loop_count = 0;
    V3 color = v3s(0.0f);
    bool alive = lane_on;  // still inside the while loop of :218-258

    for (int iter = 0;; iter++) {
        const bool want = alive && loop_count < max_bounce;  // :218 (left operand of &&)
        if (!__any_sync(kFullMask, want)) break;
        alive = want;
        TraceInfo ti;
        reset(ti);
        // :427 can only pass for rays that carry a real ignore type (after a refraction) or when a type-3 material exists
        const bool ignore_test = current_ray.ignore_type_material != VRT_MAT_NONE || P.materials_have_none != 0u;
        const bool got = grid_hit_warp<BD, AOV ? 2 : 0>(P, current_ray, want, true, ignore_test, hit, ti);
        if (want) account<AOV>(pc, ti, got, false);
        const bool first = AOV && record_aov && iter == 0 && want;
        if (first) {
            aov->grid_steps += ti.grid_steps;
            aov->voxel_steps += ti.voxel_steps;
            aov->status_fetches += ti.status_fetches;
            if (got) {
                aov->flags |= VRT_AOV_HIT;
                aov->grid_index = ti.grid_index;
                aov->voxel_index = ti.voxel_index;
                aov->material = hit.index;
                aov->t = hit.t;
                aov->point[0] = hit.point.x, aov->point[1] = hit.point.y, aov->point[2] = hit.point.z;
                aov->normal[0] = hit.normal.x, aov->normal[1] = hit.normal.y, aov->normal[2] = hit.normal.z;
            }
        }
        if (!got) alive = false;  // the while condition failed
        if (AOV && alive && iter == 0) pc.primary_hits++;

        Ray scattered = current_ray;
        V3 attenuation = v3s(0.0f);
        bool result = false;
        Ray shadow_ray = current_ray;
        if (alive) {
            loop_count += 1;  // :219
            const vrt_material material = load_material(P, hit.index);  // :223
            attenuation = v3(material.albedo_r, material.albedo_g, material.albedo_b);
            if (loop_count < max_bounce || material.type > VRT_MAT_DIELECTRIC) {
                result = scatter(material.type, material.type_data, current_ray, hit, scattered, loop_count);  // :225-239
            }  // else: last allowed bounce — the guard at :218 fails next trip whatever `result`/`scattered` are
            if (sun_enabled) {  // :240-244
                const V2 co = V2{current_ray.direction.x + current_ray.direction.z, current_ray.direction.y + current_ray.direction.z};
                // radius 0: every component of RandVec3 is (-0) + 0 * Rand = +0 exactly (Rand is finite), so the three sines are skipped
                const V3 jitter = P.sun.radius == 0.0f ? v3s(0.0f) : RandVec3mm(co, -P.sun.radius, P.sun.radius);
                const V3 sun_sample_position = ld3(P.sun.position) + jitter;
                shadow_ray = CreateRay(hit.point, sun_sample_position - hit.point);  // CreateShadowRay: ignore type MAT_NONE when enabled (:188)
            }
        }
        if (sun_enabled) {  // warp-uniform (push constant)
            HitRecord shadow_hit;
            shadow_hit.point = v3s(0.0f), shadow_hit.normal = v3s(0.0f), shadow_hit.t = 0.0f, shadow_hit.index = 0u;
            TraceInfo sti;
            reset(sti);
            const bool blocked = grid_hit_warp<BD, AOV ? 2 : 0>(P, shadow_ray, alive, false, P.materials_have_none != 0u, shadow_hit, sti);  // :247
            if (alive) {
                account<AOV>(pc, sti, blocked, true);
                if (first) {
                    aov->flags |= VRT_AOV_SHADOW_CAST;
                    aov->grid_steps += sti.grid_steps;
                    aov->voxel_steps += sti.voxel_steps;
                    aov->status_fetches += sti.status_fetches;
                    if (blocked) {
                        aov->flags |= VRT_AOV_SHADOW_BLOCKED;
                        aov->shadow_grid_index = sti.grid_index;
                        aov->shadow_voxel_index = sti.voxel_index;
                    }
                }
                if (!blocked) color = color + attenuation * sun_color;  // :248
            }
        } else if (alive) {
            color = color + attenuation;  // :251
        }
        if (alive) {
            if (!result) alive = false;  // :255
            else current_ray = scattered;
        }
    }

    if (loop_count == 0) {  // :260-262
        color = color + BackgroundColor(current_ray) * (sun_enabled ? sun_color : v3s(1.0f));
    }
    return color / (color + v3s(1.0f));  // :264
}

// brick_raytracer.comp:153-178 for this lane's pixel (px, py); every lane of the warp must call it.
template <int BD, bool AOV>
VRT_DI uint32_t shade_pixel_warp(const TraceParams& P, uint32_t px, uint32_t py, bool inside, PixelCounters& pc) {
    vrt_aov local_aov;
    if (AOV) {
        local_aov.flags = 0u;
        local_aov.grid_index = local_aov.voxel_index = local_aov.material = ~0u;
        local_aov.t = 0.0f;
        local_aov.point[0] = local_aov.point[1] = local_aov.point[2] = 0.0f;
        local_aov.normal[0] = local_aov.normal[1] = local_aov.normal[2] = 0.0f;
        local_aov.shadow_grid_index = local_aov.shadow_voxel_index = ~0u;
        local_aov.grid_steps = local_aov.voxel_steps = local_aov.status_fetches = 0u;
    }
    V3 color = v3s(0.0f);
    const int spp = P.cam.samples_per_pixel;
    for (int sample_i = 0; sample_i < spp; sample_i++) {
        const float x = (float)px;
        const float y = (float)py;
        const float flag = (float)(sample_i > 0);
        const float noise_x = hash12(V2{((x + (float)sample_i) * 0.2f) * flag, (y * 0.2f) * flag});  // :167
        const float u = (x + noise_x) / (float)(P.cam.image_width - 1u);                              // :168
        const float noise_y = hash12(V2{(x * 0.2f) * flag, ((y + (float)sample_i) * 0.2f) * flag});  // :169
        const float v = (y + noise_y) / (float)(P.cam.image_height - 1u);                             // :170
        const Ray ray = CameraGetRay(P, u, v);
        color = color + ray_color_warp<BD, AOV>(P, ray, inside, &local_aov, pc, sample_i == 0);
    }
    const float fspp = (float)spp;
    if (spp != 1) color = color / v3s(fspp);  // x / 1.0f == x
    color = v3(sqrtf(color.x), sqrtf(color.y), sqrtf(color.z));  // :176
    if (AOV && inside && P.aov) P.aov[(size_t)py * P.cam.image_width + px] = local_aov;
    return pack_rgba8(color);
}

// The common configuration — device max_bounce == 1, one sample per pixel, sun radius 0 (or sun off), every material of
// type lambertian / metal / dielectric — needs none of the scatter / RNG code: sample 0 has zero pixel noise (hash12(0,0) = 0,
// x + 0 = x, :167-170), the scatter result is dead at the last allowed bounce (:218) and the sun jitter is exactly 0.  Per
// lane the operations below are those of shade_pixel_warp; the camera ray and the sun ray go through ONE copy of the
// traversal (a two-trip loop), which keeps the kernel small enough for the instruction cache (17 KB instead of 58 KB).
template <int BD>
VRT_DI uint32_t shade_pixel_warp_simple(const TraceParams& P, uint32_t px, uint32_t py, bool inside) {
    const bool sun_enabled = P.sun.enabled > 0;
    const float u = (float)px / (float)(P.cam.image_width - 1u);   // :168 with noise 0
    const float v = (float)py / (float)(P.cam.image_height - 1u);  // :170
    Ray ray = CameraGetRay(P, u, v);
    const float primary_dir_y = ray.direction.y;
    V3 attenuation = v3s(0.0f);
    V3 color = v3s(0.0f);
    bool on = inside, primary_hit = false;
    const int trips = sun_enabled ? 2 : 1;
#pragma unroll 1
    for (int trip = 0; trip < trips; trip++) {
        HitRecord hit;
        hit.point = v3s(0.0f), hit.normal = v3s(0.0f), hit.t = 0.0f, hit.index = 0u;
        TraceInfo ti;
        reset(ti);
        const bool got = grid_hit_warp<BD, 0>(P, ray, on, trip == 0, false, hit, ti);  // :218 / :247
        if (trip == 0) {
            primary_hit = on && got;
            on = primary_hit;
            if (primary_hit) {
                const vrt_material material = load_material(P, hit.index);  // :223
                attenuation = v3(material.albedo_r, material.albedo_g, material.albedo_b);
                if (sun_enabled) ray = CreateRay(hit.point, (ld3(P.sun.position) + v3s(0.0f)) - hit.point);  // :241-244, jitter = (+0,+0,+0) at radius 0
                else color = color + attenuation;                                              // :251
            }
        } else if (on && !got) {
            color = color + attenuation * ld3(P.sun.color);  // :248
        }
    }
    if (!primary_hit) {  // :260-262
        Ray pr = ray;
        pr.direction.y = primary_dir_y;
        color = color + BackgroundColor(pr) * (sun_enabled ? ld3(P.sun.color) : v3s(1.0f));
    }
    color = color / (color + v3s(1.0f));                          // :264
    color = v3(sqrtf(color.x), sqrtf(color.y), sqrtf(color.z));  // :176, spp == 1
    return pack_rgba8(color);
}

}  // namespace vrt
