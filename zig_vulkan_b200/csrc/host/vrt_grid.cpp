// vrt_grid.cpp — host-side brick grid: the C++ counterpart of the reference's BrickGrid
// (src/modules/voxel_rt/brick/Grid.zig, State.zig, MaterialAllocator.zig).
//
// It produces, byte for byte, the five arrays + 64-byte state block that the trace kernels consume, and it tracks
// the dirty element range of every array (State.zig:14-57) so that edits are shipped as partial uploads.
// brick_dim is a runtime parameter here (the reference fixes it to 4 at State.zig:5).
#include <cstring>
#include <limits>
#include <new>
#include <vector>

#include "../../../include/vrt_host.h"
#include "vrt_host_internal.h"

namespace {
constexpr uint32_t kUnsetStart = 0xffffffffu;  // Brick.unset_index (State.zig:122-123)
}

// DeviceDataDelta.registerDelta (State.zig:39-46)
void vrt_delta::touch(uint64_t index) {
    active = true;
    if (index < from) from = index;
    if (index + 1 > to) to = index + 1;
}
// DeviceDataDelta.resetDelta (State.zig:33-37)
void vrt_delta::reset() {
    active = false;
    from = std::numeric_limits<uint64_t>::max();
    to = 0;
}

extern "C" {

// BrickGrid.init (Grid.zig:36-114)
vrt_grid* vrt_grid_create(uint32_t dim_x, uint32_t dim_y, uint32_t dim_z, uint32_t brick_dim, uint64_t brick_alloc, const float min_point[3],
                          float scale, float base_t) {
    if (!min_point || dim_x == 0 || dim_y == 0 || dim_z == 0) return nullptr;  // assert Grid.zig:38
    if (brick_dim != 4 && brick_dim != 8 && brick_dim != 16) return nullptr;
    const uint64_t brick_count = (uint64_t)dim_x * dim_y * dim_z;
    if (brick_count > 0xffffffffull) return nullptr;  // brick_count is a u32 in the reference
    vrt_grid* g = new (std::nothrow) vrt_grid();
    if (!g) return nullptr;
    g->brick_dim = brick_dim;
    g->brick_bits = brick_dim * brick_dim * brick_dim;
    g->brick_bytes = g->brick_bits / 8;
    g->brick_alloc = brick_alloc ? brick_alloc : brick_count;  // Grid.zig:51
    try {
        g->statuses.assign((size_t)((brick_count + 31) / 32), 0u);                    // :43-45
        g->brick_indices.assign((size_t)brick_count, 0u);                             // :47-49
        g->occupancy.assign((size_t)(g->brick_alloc * g->brick_bytes), (uint8_t)0);   // :53-55
        g->start_indices.assign((size_t)g->brick_alloc, kUnsetStart);                 // :57-59
        g->material_indices.assign((size_t)(g->brick_alloc * g->brick_bits), (uint8_t)0);  // :61-64
    } catch (const std::bad_alloc&) {
        delete g;
        return nullptr;
    }
    vrt_grid_state& s = g->state;
    std::memset(&s, 0, sizeof(s));
    s.voxel_dim_x = dim_x * brick_dim;
    s.voxel_dim_y = dim_y * brick_dim;
    s.voxel_dim_z = dim_z * brick_dim;
    s.dim_x = dim_x, s.dim_y = dim_y, s.dim_z = dim_z;
    for (int i = 0; i < 3; i++) s.min_point_base_t[i] = min_point[i];
    s.min_point_base_t[3] = base_t;
    const uint32_t dims[3] = {dim_x, dim_y, dim_z};
    for (int i = 0; i < 3; i++) s.max_point_scale[i] = min_point[i] + (float)dims[i] * scale;  // :74-79
    s.max_point_scale[3] = scale;
    // DeviceDataDelta.empty (State.zig:15-20): inactive, from = to = 0.  `from` therefore stays 0 until the first
    // resetDelta, so the very first upload of every array starts at element 0 — kept as is.
    for (vrt_delta& d : g->delta) d = vrt_delta{false, 0, 0};
    return g;
}

void vrt_grid_destroy(vrt_grid* g) { delete g; }

// BrickGrid.insert (Grid.zig:129-194)
int vrt_grid_insert(vrt_grid* g, uint32_t x, uint32_t y, uint32_t z, uint8_t material) {
    if (!g) return -1;
    const vrt_grid_state& s = g->state;
    if (x >= s.voxel_dim_x || y >= s.voxel_dim_y || z >= s.voxel_dim_z) return -1;  // :130-132
    const uint32_t d = g->brick_dim;
    const uint32_t fy = s.voxel_dim_y - 1u - y;  // :135 "Flip Y for more intuitive coordinates"

    const uint64_t cell = (uint64_t)(x / d) + (uint64_t)s.dim_x * ((uint64_t)(z / d) + (uint64_t)s.dim_z * (fy / d));  // gridAt :206-211
    const uint64_t word = cell / 32;
    const uint32_t bit = (uint32_t)(cell % 32);
    uint32_t brick;
    if ((g->statuses[word] >> bit) & 1u) {  // BrickStatusMask.read == .loaded
        brick = g->brick_indices[cell];     // :143
    } else {
        if (g->active_bricks >= g->brick_alloc) return -2;  // the reference would index out of bounds here
        brick = g->active_bricks++;                         // fetchAdd :147
    }

    const uint32_t nth = (x % d) + d * ((z % d) + d * (fy % d));  // voxelAt :198-203

    uint32_t& start = g->start_indices[brick];
    if (start == kUnsetStart) {  // :161
        if (g->next_material >= g->material_indices.size()) return -2;  // MaterialAllocator.nextEntry assert :40
        start = (uint32_t)g->next_material & 0x7fffffffu;                // value:u31, type = voxel_start_index (0)
        g->next_material += g->brick_bits;                               // :39
        g->delta[VRT_DELTA_START_INDICES].touch(brick);                  // :167
    }
    const uint64_t mat_at = (uint64_t)(start & 0x7fffffffu) + nth;  // :173
    g->material_indices[mat_at] = material;
    g->delta[VRT_DELTA_MATERIAL_INDICES].touch(mat_at);  // :176

    const uint64_t occ_at = (uint64_t)brick * g->brick_bytes + nth / 8;  // :179-182
    g->occupancy[occ_at] |= (uint8_t)(1u << (nth % 8));
    g->delta[VRT_DELTA_OCCUPANCY].touch(occ_at);  // :185

    g->statuses[word] |= 1u << bit;  // :188
    g->delta[VRT_DELTA_STATUSES].touch(word);
    g->brick_indices[cell] = brick;  // :192
    g->delta[VRT_DELTA_BRICK_INDICES].touch(cell);
    return 0;
}

int vrt_grid_insert_many(vrt_grid* g, const uint32_t* xyzm, size_t n) {
    if (!g || (!xyzm && n)) return -1;
    for (size_t i = 0; i < n; i++) {
        const int rc = vrt_grid_insert(g, xyzm[4 * i], xyzm[4 * i + 1], xyzm[4 * i + 2], (uint8_t)xyzm[4 * i + 3]);
        if (rc) return rc;
    }
    return 0;
}

uint32_t vrt_grid_active_bricks(const vrt_grid* g) { return g ? g->active_bricks : 0; }
uint32_t vrt_grid_brick_dim(const vrt_grid* g) { return g ? g->brick_dim : 0; }
uint64_t vrt_grid_brick_alloc(const vrt_grid* g) { return g ? g->brick_alloc : 0; }
void vrt_grid_get_state(const vrt_grid* g, vrt_grid_state* out) {
    if (g && out) *out = g->state;
}

#define VRT_GRID_ARRAY(name, type, member)                        \
    const type* name(const vrt_grid* g, uint64_t* count) {        \
        if (!g) return nullptr;                                   \
        if (count) *count = g->member.size();                     \
        return g->member.data();                                  \
    }
VRT_GRID_ARRAY(vrt_grid_statuses, uint32_t, statuses)
VRT_GRID_ARRAY(vrt_grid_brick_indices, uint32_t, brick_indices)
VRT_GRID_ARRAY(vrt_grid_occupancy, uint8_t, occupancy)
VRT_GRID_ARRAY(vrt_grid_start_indices, uint32_t, start_indices)
VRT_GRID_ARRAY(vrt_grid_material_indices, uint8_t, material_indices)
#undef VRT_GRID_ARRAY

int vrt_grid_delta_peek(vrt_grid* g, int which, uint64_t* from, uint64_t* to) {
    if (!g || which < 0 || which >= 5) return -1;
    const vrt_delta& d = g->delta[which];
    if (from) *from = d.from;
    if (to) *to = d.to;
    return d.active ? 1 : 0;
}

void vrt_grid_delta_reset(vrt_grid* g, int which) {
    if (g && which >= 0 && which < 5) g->delta[which].reset();
}

}  // extern "C"
