// vrt_shade.cuh — RayColor + main() of the reference shader, generic over the traversal implementation.
//
// `Trav` provides   static bool grid_hit<COUNT>(const TraceParams&, const Ray&, HitRecord&, TraceInfo&)
// with the semantics of GridHit(r, 0.00001, infinity, hit_min, hit) (brick_raytracer.comp:271-376), and
// shadow_hit<COUNT>(...) — the same call for the sun ray (:247), where only the boolean is consumed.  Both the
// transliteration kernel and the tuned kernel instantiate this same code, so shading can never diverge.
#pragma once

#include "vrt_device.cuh"

namespace vrt {

struct TraceInfo {
    uint32_t grid_index, voxel_index;
    uint32_t grid_steps, voxel_steps, status_fetches, bricks_entered;
};
VRT_DI void reset(TraceInfo& t) {
    t.grid_index = t.voxel_index = ~0u;
    t.grid_steps = t.voxel_steps = t.status_fetches = t.bricks_entered = 0u;
}

struct PixelCounters {
    uint32_t rays, primary_hits, shadow_rays, grid_steps, voxel_steps, status_fetches, bricks_entered, hits;
};

template <bool COUNT>
VRT_DI void account(PixelCounters& pc, const TraceInfo& ti, bool hit, bool shadow) {
    if (COUNT) {
        pc.rays++;
        pc.grid_steps += ti.grid_steps;
        pc.voxel_steps += ti.voxel_steps;
        pc.status_fetches += ti.status_fetches;
        pc.bricks_entered += ti.bricks_entered;
        pc.hits += hit ? 1u : 0u;
        pc.shadow_rays += shadow ? 1u : 0u;
    }
}

VRT_DI vrt_material load_material(const TraceParams& P, uint32_t index) {
    const vrt_material* m = P.materials + (index < P.n_materials ? index : 0u);
    vrt_material out;
    out.type = __ldg(&m->type);
    out.albedo_r = __ldg(&m->albedo_r);
    out.albedo_g = __ldg(&m->albedo_g);
    out.albedo_b = __ldg(&m->albedo_b);
    out.type_data = __ldg(&m->type_data);
    return out;
}

// brick_raytracer.comp:203-265
template <class Trav, bool AOV>
VRT_DI V3 RayColor(const TraceParams& P, const Ray& r, vrt_aov* aov, PixelCounters& pc, bool record_aov) {
    const bool sun_enabled = P.sun.enabled > 0;
    const V3 sun_color = ld3(P.sun.color);
    const int max_bounce = P.cam.max_bounce;

    HitRecord hit;
    hit.point = v3s(0.0f), hit.normal = v3s(0.0f), hit.t = 0.0f, hit.index = 0u;
    Ray current_ray = r;
    int loop_count = 0;
    V3 color = v3s(0.0f);

    for (int iter = 0;; iter++) {
        if (!(loop_count < max_bounce)) break;  // :218
        TraceInfo ti;
        reset(ti);
        const bool got = Trav::template grid_hit<AOV>(P, current_ray, hit, ti);
        account<AOV>(pc, ti, got, false);
        const bool first = AOV && record_aov && iter == 0;
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
        if (!got) break;
        if (AOV && iter == 0) pc.primary_hits++;

        loop_count += 1;  // :219
        Ray scattered = current_ray;
        const vrt_material material = load_material(P, hit.index);  // :223
        const V3 attenuation = v3(material.albedo_r, material.albedo_g, material.albedo_b);
        bool result;
        if (loop_count < max_bounce || material.type > VRT_MAT_DIELECTRIC) {
            result = scatter(material.type, material.type_data, current_ray, hit, scattered, loop_count);  // :225-239
        } else {
            // Last allowed bounce: the loop guard (:218) fails next trip whatever `result` and `scattered`
            // are, and the scatter functions have no side effects, so their evaluation is skipped.
            result = false;
        }
        if (sun_enabled) {  // :240-249
            const V2 co = V2{current_ray.direction.x + current_ray.direction.z, current_ray.direction.y + current_ray.direction.z};
            const V3 sun_sample_position = ld3(P.sun.position) + RandVec3mm(co, -P.sun.radius, P.sun.radius);
            const V3 shadow_ray_dir = sun_sample_position - hit.point;
            // CreateShadowRay (:186-190): ignore type is MAT_NONE whenever the sun is enabled, and it is only
            // called when enabled.
            const Ray shadow_ray = CreateRay(hit.point, shadow_ray_dir);
            HitRecord shadow_hit;
            shadow_hit.point = v3s(0.0f), shadow_hit.normal = v3s(0.0f), shadow_hit.t = 0.0f, shadow_hit.index = 0u;
            TraceInfo sti;
            reset(sti);
            const bool blocked = Trav::template shadow_hit<AOV>(P, shadow_ray, shadow_hit, sti);
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
        } else {
            color = color + attenuation;  // :251
        }
        if (!result) break;  // :255
        current_ray = scattered;
    }

    if (loop_count == 0) {  // :260-262
        color = color + BackgroundColor(current_ray) * (sun_enabled ? sun_color : v3s(1.0f));
    }
    return color / (color + v3s(1.0f));  // :264
}

// brick_raytracer.comp:474-477
VRT_DI Ray CameraGetRay(const TraceParams& P, float u, float v) {
    const V3 origin = ld3(P.cam.origin);
    const V3 ray_dir = fma3(ld3(P.cam.horizontal), v3s(u), ld3(P.cam.lower_left_corner)) + fma3(v3s(v), ld3(P.cam.vertical), neg(origin));
    return CreateRay(origin, ray_dir);
}

// brick_raytracer.comp:153-178 for pixel (px, py); returns the packed Rgba8 texel.
template <class Trav, bool AOV>
VRT_DI uint32_t shade_pixel(const TraceParams& P, uint32_t px, uint32_t py, PixelCounters& pc) {
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
        color = color + RayColor<Trav, AOV>(P, ray, &local_aov, pc, sample_i == 0);
    }
    const float fspp = (float)spp;
    color = v3(sqrtf(color.x / fspp), sqrtf(color.y / fspp), sqrtf(color.z / fspp));  // :176
    if (AOV && P.aov) P.aov[(size_t)py * P.cam.image_width + px] = local_aov;
    return pack_rgba8(color);
}

// Fold one thread's counters into the frame totals: warp shuffle reduction, one atomic per warp per field.
VRT_DI void flush_counters(const TraceParams& P, const PixelCounters& pc) {
    if (!P.counters) return;
    const uint32_t vals[8] = {pc.rays, pc.primary_hits, pc.shadow_rays, pc.grid_steps, pc.voxel_steps, pc.status_fetches, pc.bricks_entered, pc.hits};
#pragma unroll
    for (int i = 0; i < 8; i++) {
        unsigned long long v = vals[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(P.counters + i, v);
    }
}

}  // namespace vrt
