// vrt_trav_warp.cuh — the B200 traversal: a warp-cooperative GridHit.
//
// Results are bit-identical to the shader (assets/shaders/brick_raytracer.comp:271-471): every ray still visits the
// same cells in the same order and advances side_dist by the same sequence of FP32 additions — skipping k cells with
// one multiply would round differently and can flip which voxel is hit.  What changes is everything
// AROUND that arithmetic:
//
//   * Safe-step counts instead of per-cell occupancy tests.  A derived byte grid `dist` holds, per brick cell and per
//     direction octant, the L1 distance D (in cells) to the nearest loaded brick or grid face lying in that octant
//     (0 = loaded brick, 255 = the one-cell border around the grid, bit 7 = no loaded brick in the octant at all).
//     A DDA only moves along its own octant, one cell along one axis per step — after j steps it is at L1 distance
//     exactly j — so from a cell with distance D the next D-1 steps cannot land on a loaded brick or leave the grid
//     and run with no memory access and no bounds test:
//     three compares, one predicate op, predicated FADDs and IMADs (march_step).  Only the D-th step is followed by a lookup.
//     The border makes "left the grid" a byte value, so the march carries ONE linear index, not three coordinates.
//   * Warp rounds.  The 32 rays of a tile look their cells up TOGETHER, reduce the distances found to the warp minimum
//     k (redux.sync) and all take k steps: the step loop is warp-uniform (no per-lane counters, no divergence) and the
//     lookup latency is paid once per round instead of once per ray.
//   * Warp supersteps.  Phase A: the rays march in rounds until one of them parks on a loaded brick (or all have left
//     the grid).  Phase B, right after that round: the parked rays run the voxel-level DDA together — the rays of a tile
//     mostly reach a surface in the same round, so the expensive brick entry (three divides, a 64-bit mask fetch) is still
//     executed coherently instead of once per ray at 32 different times, and no ray waits for a slower one to park.
//   * The 4^3 voxel mask of a brick and the start of its materials are one 128-bit load from a grid-indexed record (`cell_rec`) and the voxel DDA runs in
//     registers on ONE packed integer (bounds guard + voxel index, brick_hit_warp4); 8^3 / 16^3 bricks keep the current 32-bit
//     mask word in a register (brick_hit_warp_n).  The shader does a dependent brick_indices load plus one byte load per
//     voxel step (:337,:415).
//   * Divisions by the brick/voxel scale become exact multiplications when the scale is a power of two (the result of
//     x / 2^k and x * 2^-k is the same correctly-rounded number).
//   * The material chain (brick_indices -> start index -> material index -> material) is walked once per hit, and not at
//     all for sun rays, unless the ray can ignore voxels (:427), in which case it is evaluated where the shader does.
#pragma once

#include "vrt_kernels.cuh"
#include "vrt_shade.cuh"

namespace vrt {

constexpr unsigned kFullMask = 0xffffffffu;

// -DVRT_TILE_STATS=1 (analysis builds only, tools/gpu_tilestats.py): per-warp counters of where a tile's instructions go —
// rounds, step-loop iterations, brick phases, voxel-loop iterations — kept in shared memory, flushed per tile by the trace kernel.
#ifndef VRT_PHASE_WAIT_ALL
#define VRT_PHASE_WAIT_ALL 0  // A/B: 1 = round 1's policy, the brick phase starts only when no ray of the warp is marching any more
#endif
#ifndef VRT_TILE_STATS
#define VRT_TILE_STATS 0
#endif
#if VRT_TILE_STATS
__device__ uint32_t* g_tile_stats;  // [tiles][8]
VRT_DI uint32_t* warp_stats() {
    __shared__ uint32_t s[32][8];
    return s[threadIdx.x >> 5];
}
#define VRT_STAT(i, v)                                          \
    do {                                                        \
        const uint32_t v_ = (v); /* evaluated by every lane */  \
        if ((threadIdx.x & 31u) == 0u) warp_stats()[i] += v_;   \
    } while (0)
#else
#define VRT_STAT(i, v) \
    do {               \
    } while (0)
#endif

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
    const uint32_t brick_index = grid_index < P.n_brick_indices ? __ldg(P.brick_indices + grid_index) : 0u;
    const uint32_t sw = brick_index < P.n_start_indices ? __ldg(P.start_indices + brick_index) : 0u;
    const unsigned long long mi = (unsigned long long)(sw & 0x7fffffffu) + (uint32_t)voxel_index;
    return mi < P.n_material_indices ? (uint32_t)__ldg(P.material_indices + mi) : 0u;
}

// x / s, as a multiplication when s is a power of two (bit-identical, see header)
VRT_DI V3 div_scale(V3 v, float s, float inv_s, bool exact) { return exact ? v * inv_s : v / v3s(s); }

// One brick-level DDA step (:345-372) on the linear cell index: strict '<', tie order x -> z / y -> z, the picked
// axis' side value gets its delta added (the shader's own sequence of FP32 additions per axis) and the cell index moves
// by that axis' stride.  take_x = (sx<sy)&(sx<sz); take_y = !(sx<sy)&(sy<sz); take_z = neither.
// Written in PTX for the instruction mix: on sm_100a the ALU pipe (FSETP, PLOP3, IADD3) issues a warp instruction every
// 2 cycles per scheduler and the FMA pipe (FADD, IMAD) every cycle, and the march is bound by the ALU pipe — so the step
// is 3 setp (the second and third take the first as their AND input) + 1 predicate op on the ALU pipe, and 3 predicated
// add.f32 + 3 predicated mad.lo (stride * 1 + idx, `one` is a register so that it stays an IMAD) on the FMA pipe.
VRT_DI void march_step(float& sx, float& sy, float& sz, float dx, float dy, float dz, int stx, int sty, int stz, int& idx, int one) {
    asm("{\n\t"
        ".reg .pred p1, px, py, pxy;\n\t"
        "setp.lt.f32 p1, %0, %1;\n\t"
        "setp.lt.and.f32 px, %0, %2, p1;\n\t"
        "setp.lt.and.f32 py, %1, %2, !p1;\n\t"
        "or.pred pxy, px, py;\n\t"
        "@px add.rn.f32 %0, %0, %4;\n\t"
        "@py add.rn.f32 %1, %1, %5;\n\t"
        "@!pxy add.rn.f32 %2, %2, %6;\n\t"
        "@px mad.lo.s32 %3, %7, %10, %3;\n\t"
        "@py mad.lo.s32 %3, %8, %10, %3;\n\t"
        "@!pxy mad.lo.s32 %3, %9, %10, %3;\n\t"
        "}"
        : "+f"(sx), "+f"(sy), "+f"(sz), "+r"(idx)
        : "f"(dx), "f"(dy), "f"(dz), "r"(stx), "r"(sty), "r"(stz), "r"(one));
}

// Voxel-level DDA inside one 4^3 brick (brick_raytracer.comp:378-471) with the whole state in registers: the brick's 64-bit
// mask, and ONE integer that carries both the voxel index and the bounds test —
//   bits 0-11: (x+4) | (z+4) << 4 | (y+4) << 8   (a coordinate is inside [0,4) iff bit 2 of its field is set: 3 = -1 and 8 = 4
//              are the only values a single step can leave the brick with, and neither carries into the next field),
//   bits 12- : voxel_index = x + 4 * (z + 4 * y) (:412), moved by the same predicated add as the fields.
// The step is march_step (the shader's ladder and additions, :440-467); the axis of the last step — the hit normal — is
// recovered after the loop from the difference of the last two indices.  Returns the voxel index hit or -1.
template <int INFO>
VRT_DI int brick_hit_warp4(const TraceParams& P, const Ray& r, bool ignore_test, float grid_t_max, V3 ray_delta, I3 ray_step, float g_scale,
                           V3 brick_position, unsigned long long occ, uint32_t grid_index, unsigned lanes, HitRecord& hit, AxisNormal& n, TraceInfo& ti,
                           int one) {
    const float voxel_scale = g_scale * P.brick_voxel_scale;                                                  // :389
    const V3 fposition = div_scale(RayAt(r, hit.t) - brick_position, voxel_scale, P.inv_voxel_scale, P.voxel_scale_pow2 != 0u);  // :393
    const V3 side_dist = init_side_dist(fposition, ray_step, ray_delta);                                      // :394-395
    float sx = side_dist.x, sy = side_dist.y, sz = side_dist.z;
    const int px = (int)floorf(fposition.x), py = (int)floorf(fposition.y), pz = (int)floorf(fposition.z);    // :403
    const float local_t_max = grid_t_max - hit.t;                                                            // :405
    float t_value = 0.0f;
    constexpr int kInside = 0x444;
    // a start position outside the brick (FP error at the brick face) never enters the loop (:407-410): clear the guard bits
    const bool start_inside = (uint32_t)px < 4u && (uint32_t)py < 4u && (uint32_t)pz < 4u;
    int state = start_inside ? ((px + 4) | ((pz + 4) << 4) | ((py + 4) << 8) | ((px + 4 * (pz + 4 * py)) << 12)) : 0;
    const int stx = ray_step.x * (1 + (1 << 12)), stz = ray_step.z * ((1 << 4) + (4 << 12)), sty = ray_step.y * ((1 << 8) + (16 << 12));
    int prev = state;
    int found = -1;
#if VRT_TILE_STATS
    uint32_t my_iters = 0u;
#endif
    while ((state & kInside) == kInside && t_value <= local_t_max) {  // :407-411
        if (INFO == 2) ti.voxel_steps++;
#if VRT_TILE_STATS
        my_iters++;
#endif
        const int voxel_index = state >> 12;  // :412
        if ((occ >> voxel_index) & 1ull) {    // :415-417
            bool ignore_brick = false;
            if (ignore_test) {
                hit.index = material_index_at(P, grid_index, voxel_index);  // :425
                const vrt_material m = load_material(P, hit.index);
                ignore_brick = (m.type == r.ignore_type_material) && (r.internal_reflection == m.type_data);  // :427
            }
            if (!ignore_brick) {
                found = voxel_index;
                break;
            }
        }
        t_value = fminf(fminf(sx, sy), sz) * voxel_scale;  // the side value of the axis about to step is the minimum (:443,449,457,463)
        prev = state;
        march_step(sx, sy, sz, ray_delta.x, ray_delta.y, ray_delta.z, stx, sty, stz, state, one);  // :440-467
    }
    __syncwarp(lanes);  // the rays leave the loop at different trips: finish hits (and, in the caller, misses) together
#if VRT_TILE_STATS
    {
        const uint32_t mx = __reduce_max_sync(lanes, my_iters);
        if ((threadIdx.x & 31u) == (uint32_t)(__ffs((int)lanes) - 1)) warp_stats()[3] += mx;
    }
#endif
    if (found >= 0) {
        const int moved = (state >> 12) - (prev >> 12);  // 0: hit in the entry voxel, the normal stays the brick-level one
        if (moved != 0) {
            const int a = moved < 0 ? -moved : moved;
            n = step_normal(a == 1 ? 0 : (a == 4 ? 2 : 1), ray_step);
        }
        const float t_offset = voxel_scale * 0.05f;           // :431
        hit.t += t_value - t_offset;                          // :432
        hit.normal = to_v3(n);
        hit.point = RayAt(r, hit.t) + hit.normal * t_offset;  // :433
    }
    return found;
}

#ifndef VRT_TMA_MASKS
#define VRT_TMA_MASKS 0
#endif
#if VRT_TMA_MASKS
// ---- TMA staging of 16^3 brick masks (brick_dim 16, BASELINE config 4) — MEASURED AND LEFT OFF (build with -DVRT_TMA_MASKS=1) -----
// The north star names "brick mask staged through TMA into shared memory".  This is that, where it has its best case; it is
// bit-exact and 14 % slower than the register-cached mask word (profiles/README.md "TMA staging", ncu captures of both builds
// under profiles/): the bulk copy + mbarrier wait sit in front of every brick test, while the word cache already removes most
// loads.  For 4^3 bricks the whole mask is one 8-byte load — there is nothing to stage.
// A 16^3 brick's occupancy mask is 512 bytes = four cache lines, and a voxel DDA touches them one word at a time along a
// dependent chain.  The rays a warp parks in one phase sit on few distinct bricks (a 16^3 brick covers many pixels), so phase B
// first copies the masks of up to kStageSlots distinct bricks into shared memory with one bulk-async copy each
// (cp.async.bulk, SASS UBLKCP; completion on a per-warp mbarrier) and the DDA then reads words with LDS.  Lanes whose brick
// did not get a slot read global memory as before.
constexpr int kStageSlots = 2;
constexpr int kStageWords = 128;  // 512 bytes
struct BrickStage {
    alignas(128) uint32_t mask[8][kStageSlots][kStageWords];  // [warp of the CTA][slot][word]
    alignas(8) unsigned long long bar[8];
    uint32_t phase[8];
};
VRT_DI BrickStage& brick_stage() {
    __shared__ BrickStage s;
    return s;
}
VRT_DI uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// every thread of the CTA, once, before the first traversal (kernels with brick_dim == 16 only)
VRT_DI void brick_stage_init() {
    BrickStage& s = brick_stage();
    if ((threadIdx.x & 31u) == 0u) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s.bar[threadIdx.x >> 5])) : "memory");
        s.phase[threadIdx.x >> 5] = 0u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
}
// Called by the parked lanes of a warp (mask `lanes`) with their brick's mask offset; returns this lane's slot or -1.
VRT_DI int brick_stage_masks(const TraceParams& P, unsigned lanes, uint32_t brick_index, unsigned long long mask_base) {
    BrickStage& s = brick_stage();
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const bool in_bounds = mask_base + 4ull * kStageWords <= P.n_occupancy;
    const unsigned peers = __match_any_sync(lanes, brick_index);
    const int leader = __ffs(peers) - 1;
    const bool is_leader = (int)lane == leader && in_bounds;
    const unsigned leaders = __ballot_sync(lanes, is_leader);
    const int my_rank = __popc(leaders & ((1u << lane) - 1u));
    const int leader_rank = __shfl_sync(lanes, my_rank, leader);
    int slot = (((leaders >> leader) & 1u) && leader_rank < kStageSlots) ? leader_rank : -1;
    const int n_staged = min(__popc(leaders), kStageSlots);
    if (n_staged == 0) return -1;  // uniform over `lanes`
    const uint32_t bar = smem_u32(&s.bar[warp]);
    const uint32_t phase = s.phase[warp];
    __syncwarp(lanes);
    if ((int)lane == __ffs(lanes) - 1) {
        // the slots were last read through the generic proxy (previous phase B): order those reads before the async writes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(n_staged * 4 * kStageWords)) : "memory");
        s.phase[warp] = phase ^ 1u;
    }
    __syncwarp(lanes);
    if (is_leader && my_rank < kStageSlots) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(&s.mask[warp][my_rank][0])),
                     "l"(P.occupancy + mask_base), "r"((uint32_t)(4 * kStageWords)), "r"(bar)
                     : "memory");
    }
    uint32_t done = 0u;
    for (int tries = 0; tries < (1 << 22) && !done; tries++) {  // bounded: a copy that never lands must not hang the GPU
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(phase)
            : "memory");
    }
    return done ? slot : -1;
}
#endif  // VRT_TMA_MASKS

// The same for 8^3 and 16^3 bricks (brick_dim is a specialization constant upstream, State.zig:5; 16 is this repo's documented
// extension).  State layout: 6-bit fields (x+bd) | (z+bd) << 6 | (y+bd) << 12 — inside iff bit log2(bd) of every field is set — and
// voxel_index in bits 18+.  The brick's mask lives in the occupancy buffer (:415); the 32-bit word holding the current voxel's
// bit is kept in a register and re-read only when the voxel index leaves it (x steps and most z steps stay inside a word),
// instead of one byte load per voxel step.  (The shader's uint8_t truncation of voxel_index / 8, :413, cannot trigger for an
// in-range index at brick_dim 8 and is widened at 16, as in the oracle.)
template <int INFO>
VRT_DI int brick_hit_warp_n(const TraceParams& P, const Ray& r, bool ignore_test, float grid_t_max, V3 ray_delta, I3 ray_step, float g_scale,
                            V3 brick_position, uint32_t grid_index, unsigned lanes, HitRecord& hit, AxisNormal& n, TraceInfo& ti, int one) {
    const int bd = P.brick_dim;
    const float voxel_scale = g_scale * P.brick_voxel_scale;                                                  // :389
    const V3 fposition = div_scale(RayAt(r, hit.t) - brick_position, voxel_scale, P.inv_voxel_scale, P.voxel_scale_pow2 != 0u);  // :393
    const V3 side_dist = init_side_dist(fposition, ray_step, ray_delta);                                      // :394-395
    float sx = side_dist.x, sy = side_dist.y, sz = side_dist.z;
    const int px = (int)floorf(fposition.x), py = (int)floorf(fposition.y), pz = (int)floorf(fposition.z);    // :403
    const float local_t_max = grid_t_max - hit.t;                                                            // :405
    float t_value = 0.0f;
    const int inside = bd | (bd << 6) | (bd << 12);
    const bool start_inside = (uint32_t)px < (uint32_t)bd && (uint32_t)py < (uint32_t)bd && (uint32_t)pz < (uint32_t)bd;  // :407-410
    int state = start_inside ? ((px + bd) | ((pz + bd) << 6) | ((py + bd) << 12) | ((px + bd * (pz + bd * py)) << 18)) : 0;
    const int stx = ray_step.x * (1 + (1 << 18)), stz = ray_step.z * ((1 << 6) + (bd << 18)), sty = ray_step.y * ((1 << 12) + ((bd * bd) << 18));
    const uint32_t brick_index = grid_index < P.n_brick_indices ? __ldg(P.brick_indices + grid_index) : 0u;  // :337
    const unsigned long long mask_base = (unsigned long long)brick_index * P.brick_bytes;                    // :390
#if VRT_TMA_MASKS
    const int slot = bd == 16 ? brick_stage_masks(P, lanes, brick_index, mask_base) : -1;                     // 16^3: mask staged in shared memory
    const uint32_t* staged = slot >= 0 ? brick_stage().mask[threadIdx.x >> 5][slot] : nullptr;
#endif
    int word_at = -1;
    uint32_t word = 0u;
    int prev = state;
    int found = -1;
    while ((state & inside) == inside && t_value <= local_t_max) {  // :407-411
        if (INFO == 2) ti.voxel_steps++;
        const int voxel_index = state >> 18;  // :412
        if ((voxel_index >> 5) != word_at) {
            word_at = voxel_index >> 5;
            const unsigned long long at = mask_base + 4ull * (unsigned long long)word_at;
#if VRT_TMA_MASKS
            if (staged) word = staged[word_at];
            else
#endif
            word = at < P.n_occupancy ? __ldg(reinterpret_cast<const uint32_t*>(P.occupancy + at)) : 0u;  // bytes :415 reads one at a time
        }
        if ((word >> (voxel_index & 31)) & 1u) {  // :415-417
            bool ignore_brick = false;
            if (ignore_test) {
                hit.index = material_index_at(P, grid_index, voxel_index);  // :425
                const vrt_material m = load_material(P, hit.index);
                ignore_brick = (m.type == r.ignore_type_material) && (r.internal_reflection == m.type_data);  // :427
            }
            if (!ignore_brick) {
                found = voxel_index;
                break;
            }
        }
        t_value = fminf(fminf(sx, sy), sz) * voxel_scale;
        prev = state;
        march_step(sx, sy, sz, ray_delta.x, ray_delta.y, ray_delta.z, stx, sty, stz, state, one);  // :440-467
    }
    __syncwarp(lanes);
    if (found >= 0) {
        const int moved = (state >> 18) - (prev >> 18);
        if (moved != 0) {
            const int a = moved < 0 ? -moved : moved;
            n = step_normal(a == 1 ? 0 : (a == bd ? 2 : 1), ray_step);
        }
        const float t_offset = voxel_scale * 0.05f;           // :431
        hit.t += t_value - t_offset;                          // :432
        hit.normal = to_v3(n);
        hit.point = RayAt(r, hit.t) + hit.normal * t_offset;  // :433
    }
    return found;
}

// Reference grid index (:318) of the padded linear cell index used by the march.
VRT_DI uint32_t cell_grid_index(const TraceParams& P, int idx, int log_px, int log_pzx) {
    const int x = (idx & ((1 << log_px) - 1)) - 1, z = ((idx >> log_px) & ((1 << (log_pzx - log_px)) - 1)) - 1, y = (idx >> log_pzx) - 1;
    return (uint32_t)(x + (int)P.grid.dim_x * (z + (int)P.grid.dim_z * y));
}

// GridHit(r, 0.00001, infinity, ...) (:271-376) for the 32 rays of a warp.  Every lane of the warp must call this;
// `active` says whether the lane has a ray.  need_material: produce hit.index (camera / bounce rays; sun rays only
// need the boolean).  ignore_test: the ray can ignore voxels (:427).
// INFO: 0 = hit record only; 1 = also the hit cell / voxel indices in `ti`; 2 = also the shader's step / fetch counters
// (disables the free-octant shortcut so that the counters equal the shader's).
template <int BD, int INFO>
VRT_DI bool grid_hit_warp(const TraceParams& P, const Ray& r, bool active, bool need_material, bool ignore_test, HitRecord& hit, TraceInfo& ti) {
    const V3 g_min = v3(P.grid.min_point_base_t[0], P.grid.min_point_base_t[1], P.grid.min_point_base_t[2]);
    const float g_scale = P.grid.max_point_scale[3];
    const int log_px = (int)P.dist_log_px, log_pzx = (int)(P.dist_log_px + P.dist_log_pz);

    float grid_t_min = 0.00001f, grid_t_max = __int_as_float(0x7f800000);
    V3 ray_delta = v3s(0.0f);
    I3 ray_step = I3{0, 0, 0};
    AxisNormal n = AxisNormal{0, 0.0f};
    float sx = 0.0f, sy = 0.0f, sz = 0.0f, t_side = 0.0f;
    int idx = 0, stx = 0, sty = 0, stz = 0;
    bool marching = false;

    if (active && !isnan((r.direction.x + r.direction.y) + r.direction.z)) {  // non-finite direction: a miss (DESIGN.md "Deviations")
        const V3 g_max = v3(P.grid.max_point_scale[0], P.grid.max_point_scale[1], P.grid.max_point_scale[2]);
        const V3 inv_ray_dir = v3(safeInverse(r.direction.x), safeInverse(r.direction.y), safeInverse(r.direction.z));  // :278
        V3 slab_normal;
        if (AdvNormIntersect(g_min, g_max, r, inv_ray_dir, slab_normal, grid_t_min, grid_t_max)) {  // :282
            n.axis = slab_normal.x != 0.0f ? 0 : (slab_normal.y != 0.0f ? 1 : 2);
            n.sign = n.axis == 0 ? slab_normal.x : (n.axis == 1 ? slab_normal.y : slab_normal.z);
            const float global_t_value = grid_t_min + 0.0001f * g_scale;  // :287
            ray_delta = v3(fabsf(inv_ray_dir.x), fabsf(inv_ray_dir.y), fabsf(inv_ray_dir.z));  // :290
            ray_step = I3{(int)gsign(r.direction.x), (int)gsign(r.direction.y), (int)gsign(r.direction.z)};  // :291
            const V3 fposition = div_scale(RayAt(r, global_t_value) - g_min, g_scale, P.inv_scale, P.scale_pow2 != 0u);  // :293-296
            const V3 side_dist = init_side_dist(fposition, ray_step, ray_delta);  // :297-298
            sx = side_dist.x, sy = side_dist.y, sz = side_dist.z;
            const int px = (int)floorf(fposition.x), py = (int)floorf(fposition.y), pz = (int)floorf(fposition.z);  // :311
            // :313-315 for the start cell; afterwards the border bytes of `dist` stand in for the bounds test
            if ((uint32_t)px < P.grid.dim_x && (uint32_t)py < P.grid.dim_y && (uint32_t)pz < P.grid.dim_z) {
                idx = (px + 1) + ((pz + 1) << log_px) + ((py + 1) << log_pzx);
                stx = ray_step.x, stz = ray_step.z << log_px, sty = ray_step.y << log_pzx;
                marching = true;
            }
        }
    }
    const float dx = ray_delta.x, dy = ray_delta.y, dz = ray_delta.z;
    // `idx` addresses the distance grid of this ray's direction octant (a zero step never moves, either sign's grid is
    // valid for it): the octant's plane offset is folded into the index and removed again when a cell is decoded
    const int obase = (int)(((ray_step.x < 0 ? 1u : 0u) | (ray_step.y < 0 ? 2u : 0u) | (ray_step.z < 0 ? 4u : 0u)) * (uint32_t)P.dist_plane);
    idx += obase;
    const uint8_t* __restrict__ dist = P.dist;
    const int one = (int)P.one;
    int last_stride = 0;  // stride of the most recent step (0: none yet -> slab normal)
    bool result = false;
    constexpr bool COUNT = INFO == 2;
    uint32_t cnt_word = ~0u;  // COUNT: the reference's one-word status cache (:301,:321-326)
    constexpr uint32_t kIdle = 0xffffu;
    enum : int { kMarching = 0, kParked = 1, kDone = 2 };
    int mode = marching ? kMarching : kDone;
    // deltas / strides actually applied in the step loop: zero while this lane is parked or finished
    float fdx = marching ? dx : 0.0f, fdy = marching ? dy : 0.0f, fdz = marching ? dz : 0.0f;
    int fsx = stx, fsy = sty, fsz = stz;  // (stx.. are 0 when not marching)

    while (__any_sync(kFullMask, mode != kDone)) {
        // ---- phase A (:313-373 without the per-cell tests): rounds of { every marching ray looks its cell up; all of
        // them take k = min over the warp of the distances found steps }.  k is warp-uniform, so the step loop has no
        // per-lane counter and no divergence; rays that are parked / finished ride along with zero deltas and strides
        // (x + 0.0f == x), which keeps the parked rays' DDA state intact for the step after a brick miss.
        for (;;) {
            uint32_t d = kIdle;
            if (mode == kMarching) {
                d = __ldg(dist + idx);
                // bit 7: left the grid (border byte 255, :313-315), or — exact shortcut — no loaded brick exists anywhere in
                // the octant this DDA can reach, so the shader's loop would only step through empty cells until it
                // leaves the grid.  (COUNT keeps marching through free octants so that its step counters equal the shader's.)
                const bool out = COUNT ? d == kDistBorder : (d & kDistFree) != 0u;
                if (!out) {
                    d &= 0x7fu;
                    if (COUNT) {  // an in-grid cell = one iteration of the shader's loop; emulate its one-word status cache (:321-326)
                        ti.grid_steps++;
                        const uint32_t gi = cell_grid_index(P, idx - obase, log_px, log_pzx);
                        if ((gi >> 5) != cnt_word) cnt_word = gi >> 5, ti.status_fetches++;
                    }
                }
                if (out || d == 0u) {  // d == 0: status bit set (:328) -> park for phase B
                    mode = out ? kDone : kParked;
                    d = kIdle;
                    fdx = fdy = fdz = 0.0f, fsx = fsy = fsz = 0;
                }
            }
            const uint32_t k = __reduce_min_sync(kFullMask, d);
            if (k == kIdle) break;
            VRT_STAT(0, 1u);                                           // rounds
            VRT_STAT(1, k);                                            // step-loop iterations
            VRT_STAT(4, (uint32_t)__popc(__ballot_sync(kFullMask, d != kIdle)));  // lanes marching in this round
            VRT_STAT(6, (uint32_t)__popc(__ballot_sync(kFullMask, mode == kParked)));  // lanes waiting in it for the brick phase
            // A round in which a ray parked is the last one before the brick phase: parked rays do not wait for the others to park
            // too.  (Waiting for all of them — fewer, fuller brick phases — cost the expensive tiles 2-3x their rounds: each
            // superstep lasted as long as its slowest ray, profiles/r02_tile_statistics_C3.txt.)  A scheduling choice only: every
            // ray visits the same cells either way.
            const bool to_bricks = !VRT_PHASE_WAIT_ALL && __any_sync(kFullMask, mode == kParked);
            const bool on = d != kIdle;
            if (COUNT) {
                for (uint32_t i = 1; i < k; i++) {  // k-1 steps onto cells known to be empty and inside
                    march_step(sx, sy, sz, fdx, fdy, fdz, fsx, fsy, fsz, idx, one);
                    if (on) {
                        ti.grid_steps++;
                        const uint32_t gi = cell_grid_index(P, idx - obase, log_px, log_pzx);
                        if ((gi >> 5) != cnt_word) cnt_word = gi >> 5, ti.status_fetches++;
                    }
                }
            } else {
                // k-1 steps onto cells known to be empty and inside.  Most rounds are short (near a surface k is 1 - 3): blocks of four,
                // then the two low bits of the count by themselves — a handful of instructions of loop control per round, not a
                // general unrolled loop's prologue + remainder loop.
                uint32_t n = k - 1u;
#pragma unroll 1
                for (; n >= 4u; n -= 4u) {
                    march_step(sx, sy, sz, fdx, fdy, fdz, fsx, fsy, fsz, idx, one);
                    march_step(sx, sy, sz, fdx, fdy, fdz, fsx, fsy, fsz, idx, one);
                    march_step(sx, sy, sz, fdx, fdy, fdz, fsx, fsy, fsz, idx, one);
                    march_step(sx, sy, sz, fdx, fdy, fdz, fsx, fsy, fsz, idx, one);
                }
                if (n & 2u) {
                    march_step(sx, sy, sz, fdx, fdy, fdz, fsx, fsy, fsz, idx, one);
                    march_step(sx, sy, sz, fdx, fdy, fdz, fsx, fsy, fsz, idx, one);
                }
                if (n & 1u) march_step(sx, sy, sz, fdx, fdy, fdz, fsx, fsy, fsz, idx, one);
            }
            // the k-th step lands on a cell that is looked up next round; remember its side value and axis
            if (on) t_side = fminf(fminf(sx, sy), sz);  // = side_dist.<axis> before the increment (the picked side is the minimum)
            const int before = idx;
            march_step(sx, sy, sz, fdx, fdy, fdz, fsx, fsy, fsz, idx, one);
            if (on) last_stride = idx - before;
            if (to_bricks) break;
        }
        // ---- phase B: the parked rays test their bricks together (:329-342)
        const unsigned parked_lanes = __ballot_sync(kFullMask, mode == kParked);
        if (parked_lanes) {
            VRT_STAT(2, 1u);                                // brick phases
            VRT_STAT(5, (uint32_t)__popc(parked_lanes));    // lanes testing a brick in them
        }
        if (mode == kParked) {
            if (last_stride != 0) {
                const int a = last_stride < 0 ? -last_stride : last_stride;
                n = step_normal(a == 1 ? 0 : (a == (1 << log_px) ? 2 : 1), ray_step);
            }
            const int cell = idx - obase;
            const int x = (cell & ((1 << log_px) - 1)) - 1, z = ((cell >> log_px) & ((1 << (log_pzx - log_px)) - 1)) - 1, y = (cell >> log_pzx) - 1;
            const uint32_t grid_index = (uint32_t)(x + (int)P.grid.dim_x * (z + (int)P.grid.dim_z * y));  // :318
            unsigned long long occ = 0ull;
            uint32_t mat_base = 0u;
            if (BD == 4) {
                const uint4 rec = __ldg(P.cell_rec + grid_index);
                occ = (unsigned long long)rec.x | ((unsigned long long)rec.y << 32), mat_base = rec.z;
            }
            const V3 brick_min = fma3(v3((float)x, (float)y, (float)z), v3s(g_scale), g_min);  // :331
            const float t_value = t_side * g_scale;                                            // :347,353,361,367
            hit.t = (t_value + grid_t_min) + 0.01f * g_scale;                                  // :332-334
            if (COUNT) ti.bricks_entered++;
            const int voxel_index = BD == 4 ? brick_hit_warp4<INFO>(P, r, ignore_test, grid_t_max, ray_delta, ray_step, g_scale, brick_min, occ, grid_index, parked_lanes, hit, n, ti, one)
                                             : brick_hit_warp_n<INFO>(P, r, ignore_test, grid_t_max, ray_delta, ray_step, g_scale, brick_min, grid_index, parked_lanes, hit, n, ti, one);
            if (voxel_index >= 0) {
                if (need_material && !ignore_test) {
                    if (BD == 4) {  // material_indices[start + voxel_index] (:425) with the start index already in hand
                        const unsigned long long mi = (unsigned long long)mat_base + (uint32_t)voxel_index;
                        hit.index = mi < P.n_material_indices ? (uint32_t)__ldg(P.material_indices + mi) : 0u;
                    } else {
                        hit.index = material_index_at(P, grid_index, voxel_index);
                    }
                }
                if (INFO >= 1) ti.grid_index = grid_index, ti.voxel_index = (uint32_t)voxel_index;
                result = true;
                mode = kDone;
            } else {  // :345-372, then the next cell is looked up
                t_side = fminf(fminf(sx, sy), sz);
                const int before = idx;
                march_step(sx, sy, sz, dx, dy, dz, stx, sty, stz, idx, one);
                last_stride = idx - before;
                fdx = dx, fdy = dy, fdz = dz, fsx = stx, fsy = sty, fsz = stz;  // marching again
                mode = kMarching;
            }
        }
    }
    return result;
}

}  // namespace vrt
