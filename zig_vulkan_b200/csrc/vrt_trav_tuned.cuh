// vrt_trav_tuned.cuh — the B200 traversal: same per-cell FP32 arithmetic as the shader (so hit records are
// bit-identical to the oracle), different memory system.
//
// What changes relative to the transliteration (vrt_trav_ref.cuh):
//   * brick level: the 1-bit-per-brick status buffer (x-major u32 words, reloaded on almost every y/z step,
//     brick_raytracer.comp:321-326) is re-tiled on the device into a 64-ary mask pyramid:
//         coarse   1 bit  per 4x4x4 bricks   -> staged ONCE per CTA into shared memory by a TMA bulk copy
//         status64 1 u64  per 4x4x4 bricks   -> shared memory when it fits (<= 64^3 bricks), else L1/L2
//     so a ray marching through empty space issues no global loads at all, and a status word is reused for
//     ~4 cells in every direction instead of only along x.
//   * voxel level: the 8-byte brick mask is fetched ONCE per brick entry as one u64 from a grid-indexed
//     copy (`occ_dense`), replacing the dependent brick_indices -> byte-per-voxel-step loads (:337,:415); the
//     voxel DDA then runs entirely in registers.
//   * the material chain (brick_indices -> start index -> material index) is only walked for rays whose
//     shading needs it (primary/bounce rays), not for sun rays.
//   * scheduling: persistent CTAs (one wave), each warp pulls 8x4-pixel tiles from a global counter, so
//     sky tiles that die after a slab test do not idle an SM while terrain tiles march hundreds of cells.
//   * framebuffer: a warp's 8x4 tile is written as eight 128-bit stores (4 texels each, gathered by shfl).
//
// The per-cell arithmetic sequence (side_dist accumulation by repeated FP32 adds, t_value read before the
// increment, tie order) is untouched: skipping k cells with one multiply would round differently and can
// flip which voxel is hit (SURVEY.md §7 "Hard parts").
#pragma once

#include "vrt_kernels.cuh"
#include "vrt_shade.cuh"
#include "vrt_trav_ref.cuh"

namespace vrt {

extern __shared__ __align__(128) unsigned char vrt_smem[];

struct SmemHeader {
    unsigned long long mbar;   // mbarrier the TMA bulk copies complete on
    uint32_t status_in_smem;   // 1 if status64 was staged next to the coarse bits
    uint32_t coarse_bytes;     // padded to 16 B (cp.async.bulk granularity)
    uint32_t pad[12];          // header = 64 B, keeps the payload 16-B aligned
};
static_assert(sizeof(SmemHeader) == 64, "SmemHeader must stay 64 bytes");

VRT_DI const uint32_t* smem_coarse() { return reinterpret_cast<const uint32_t*>(vrt_smem + sizeof(SmemHeader)); }
VRT_DI const unsigned long long* smem_status64() {
    const SmemHeader* h = reinterpret_cast<const SmemHeader*>(vrt_smem);
    return h->status_in_smem ? reinterpret_cast<const unsigned long long*>(vrt_smem + sizeof(SmemHeader) + h->coarse_bytes) : nullptr;
}

// hit.normal as (axis, sign): every normal this path produces has one non-zero component (:350-370, :530-531)
struct AxisNormal {
    int axis;
    float sign;
};
VRT_DI V3 to_v3(AxisNormal n) { return v3(n.axis == 0 ? n.sign : 0.0f, n.axis == 1 ? n.sign : 0.0f, n.axis == 2 ? n.sign : 0.0f); }
VRT_DI AxisNormal step_normal(int axis, I3 ray_step) {
    const int s = axis == 0 ? ray_step.x : (axis == 1 ? ray_step.y : ray_step.z);
    return AxisNormal{axis, s < 0 ? 1.0f : -1.0f};  // normal_axis (:304-308)
}

// brick_indices -> start index -> material_indices (:337, :422-425)
VRT_DI uint32_t material_index_at(const TraceParams& P, uint32_t grid_index, int voxel_index) {
    const uint32_t brick_index = __ldg(P.brick_indices + grid_index);
    const uint32_t sw = brick_index < P.n_start_indices ? __ldg(P.start_indices + brick_index) : 0u;
    const unsigned long long mi = (unsigned long long)(sw & 0x7fffffffu) + (uint32_t)voxel_index;
    return mi < P.n_material_indices ? (uint32_t)__ldg(P.material_indices + mi) : 0u;
}

// Voxel-level DDA over one 4^3 brick whose 64-bit mask is in registers (brick_raytracer.comp:378-471).
// Returns the voxel index hit, or -1.  On a hit, hit.t / hit.point / hit.normal are final.
// IGNORE_TEST: evaluate :427 (needs the material of every solid voxel met); sets hit.index as the shader does.
template <bool IGNORE_TEST>
VRT_DI int brick_hit_u64(const TraceParams& P, const Ray& r, float grid_t_max, V3 ray_delta, I3 ray_step, float g_scale, V3 brick_position,
                         unsigned long long occ, uint32_t grid_index, HitRecord& hit, AxisNormal& n) {
    const float voxel_scale = g_scale * P.brick_voxel_scale;                     // :389
    const V3 fposition = (RayAt(r, hit.t) - brick_position) / v3s(voxel_scale);  // :393
    V3 side_dist = init_side_dist(fposition, ray_step, ray_delta);               // :394-395
    I3 pos = I3{(int)floorf(fposition.x), (int)floorf(fposition.y), (int)floorf(fposition.z)};  // :403
    const float local_t_max = grid_t_max - hit.t;                                              // :405
    float t_value = 0.0f;
    // 0 <= pos < 4 on all three axes  <=>  ((x|y|z) & ~3) == 0 in two's complement
    while ((((uint32_t)pos.x | (uint32_t)pos.y | (uint32_t)pos.z) & ~3u) == 0u && t_value <= local_t_max) {
        const int voxel_index = pos.x + 4 * (pos.z + 4 * pos.y);  // :412
        if ((occ >> voxel_index) & 1ull) {                        // :415-417
            bool ignore_brick = false;
            if (IGNORE_TEST) {
                hit.index = material_index_at(P, grid_index, voxel_index);  // :425
                const vrt_material m = load_material(P, hit.index);
                ignore_brick = (m.type == r.ignore_type_material) && (r.internal_reflection == m.type_data);  // :427
            }
            if (!ignore_brick) {
                const float t_offset = voxel_scale * 0.05f;           // :431
                hit.t += t_value - t_offset;                          // :432
                hit.normal = to_v3(n);
                hit.point = RayAt(r, hit.t) + hit.normal * t_offset;  // :433
                return voxel_index;
            }
        }
        n = step_normal(dda_step(side_dist, ray_delta, pos, ray_step, voxel_scale, t_value), ray_step);  // :440-467
    }
    return -1;
}

// brick_raytracer.comp:271-376 with t_min = 0.00001, t_max = +inf, for brick_dim == 4.
// NEED_MATERIAL: produce hit.index (primary / bounce rays); sun rays only need the boolean.
// IGNORE_TEST:   the ray can actually ignore voxels (:427) — see TunedTrav.
// COUNT:         record the hit cell / voxel and the number of cells marched in `ti` (AOV parity path).
template <bool NEED_MATERIAL, bool IGNORE_TEST, bool COUNT>
VRT_DI bool grid_hit_tuned(const TraceParams& P, const uint32_t* __restrict__ s_coarse, const unsigned long long* __restrict__ s_status64,
                           const Ray& r, HitRecord& hit, TraceInfo& ti) {
    const V3 g_min = v3(P.grid.min_point_base_t[0], P.grid.min_point_base_t[1], P.grid.min_point_base_t[2]);
    const V3 g_max = v3(P.grid.max_point_scale[0], P.grid.max_point_scale[1], P.grid.max_point_scale[2]);
    const float g_scale = P.grid.max_point_scale[3];
    const int dim_x = (int)P.grid.dim_x, dim_y = (int)P.grid.dim_y, dim_z = (int)P.grid.dim_z;

    if (isnan((r.direction.x + r.direction.y) + r.direction.z)) return false;  // see RefTraversal::grid_hit
    const V3 inv_ray_dir = v3(safeInverse(r.direction.x), safeInverse(r.direction.y), safeInverse(r.direction.z));  // :278
    float grid_t_min = 0.00001f;
    float grid_t_max = __int_as_float(0x7f800000);
    V3 slab_normal;
    if (!AdvNormIntersect(g_min, g_max, r, inv_ray_dir, slab_normal, grid_t_min, grid_t_max)) return false;  // :282
    AxisNormal n;
    n.axis = slab_normal.x != 0.0f ? 0 : (slab_normal.y != 0.0f ? 1 : 2);
    n.sign = n.axis == 0 ? slab_normal.x : (n.axis == 1 ? slab_normal.y : slab_normal.z);

    const float global_t_value = grid_t_min + 0.0001f * g_scale;  // :287
    const V3 ray_delta = v3(fabsf(inv_ray_dir.x), fabsf(inv_ray_dir.y), fabsf(inv_ray_dir.z));  // :290
    const I3 ray_step = I3{(int)gsign(r.direction.x), (int)gsign(r.direction.y), (int)gsign(r.direction.z)};  // :291
    const V3 fposition = (RayAt(r, global_t_value) - g_min) / v3s(g_scale);  // :293-296
    V3 side_dist = init_side_dist(fposition, ray_step, ray_delta);           // :297-298
    float t_value = 0.0f;
    I3 pos = I3{(int)floorf(fposition.x), (int)floorf(fposition.y), (int)floorf(fposition.z)};  // :311

    uint32_t cached_super = ~0u;
    unsigned long long sbits = 0ull;

    // :313-317 (the `global_t_value <= t_max` term is `finite <= +inf`)
    while ((uint32_t)pos.x < (uint32_t)dim_x && (uint32_t)pos.y < (uint32_t)dim_y && (uint32_t)pos.z < (uint32_t)dim_z) {
        if (COUNT) ti.grid_steps++;
        const uint32_t super = (uint32_t)(pos.x >> 2) + P.sdim_x * ((uint32_t)(pos.z >> 2) + P.sdim_z * (uint32_t)(pos.y >> 2));
        if (super != cached_super) {
            cached_super = super;
            sbits = 0ull;
            if ((s_coarse[super >> 5] >> (super & 31u)) & 1u) sbits = s_status64 ? s_status64[super] : __ldg(P.status64 + super);
        }
        const uint32_t local = (uint32_t)(pos.x & 3) + 4u * ((uint32_t)(pos.z & 3) + 4u * (uint32_t)(pos.y & 3));
        if ((sbits >> local) & 1ull) {  // :328
            const uint32_t grid_index = (uint32_t)(pos.x + dim_x * (pos.z + dim_z * pos.y));  // :318
            const unsigned long long occ = __ldg(P.occ_dense + grid_index);
            if (occ != 0ull) {  // an all-empty loaded brick: BrickHit would step through it and miss
                const V3 brick_min = fma3(tofloat(pos), v3s(g_scale), g_min);  // :331
                hit.t = (t_value + grid_t_min) + 0.01f * g_scale;              // :332-334
                const int voxel_index =
                    brick_hit_u64<IGNORE_TEST>(P, r, grid_t_max, ray_delta, ray_step, g_scale, brick_min, occ, grid_index, hit, n);
                if (voxel_index >= 0) {
                    if (NEED_MATERIAL && !IGNORE_TEST) hit.index = material_index_at(P, grid_index, voxel_index);
                    if (COUNT) ti.grid_index = grid_index, ti.voxel_index = (uint32_t)voxel_index;
                    return true;
                }
            }
        }
        n = step_normal(dda_step(side_dist, ray_delta, pos, ray_step, g_scale, t_value), ray_step);  // :345-372
    }
    return false;
}

// Trav interface of vrt_shade.cuh.
// :427 — a hit is ignored only when materials[idx].type == ray.ignore_type AND ir == type_data.  Camera rays,
// lambert/metal bounces and sun rays carry ignore_type == MAT_NONE(3); unless a type-3 material was uploaded
// (P.materials_have_none) that test can never pass, so it — and for sun rays the whole material chain — is
// skipped.  Rays that carry ignore_type == DIELECTRIC (after a refraction) evaluate it.
struct TunedTrav {
    template <bool COUNT>
    static VRT_DI bool grid_hit(const TraceParams& P, const Ray& r, HitRecord& hit, TraceInfo& ti) {
        if (r.ignore_type_material != VRT_MAT_NONE || P.materials_have_none)
            return grid_hit_tuned<true, true, COUNT>(P, smem_coarse(), smem_status64(), r, hit, ti);
        return grid_hit_tuned<true, false, COUNT>(P, smem_coarse(), smem_status64(), r, hit, ti);
    }
    template <bool COUNT>
    static VRT_DI bool shadow_hit(const TraceParams& P, const Ray& r, HitRecord& hit, TraceInfo& ti) {
        if (P.materials_have_none) return grid_hit_tuned<false, true, COUNT>(P, smem_coarse(), smem_status64(), r, hit, ti);
        return grid_hit_tuned<false, false, COUNT>(P, smem_coarse(), smem_status64(), r, hit, ti);
    }
};

cudaError_t launch_trace_tuned(const TraceParams& P, bool aov, cudaStream_t stream, LaunchInfo* info);

}  // namespace vrt
