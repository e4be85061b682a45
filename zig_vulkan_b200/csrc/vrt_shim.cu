// vrt_shim.cu — implementation of the C ABI in include/vrt.h on the CUDA runtime.
//
// One vrt_ctx == one ComputePipeline + its device buffer + its target image (reference:
// src/modules/voxel_rt/ComputePipeline.zig, Pipeline.zig:103-126,272-316).  The Vulkan pieces map as:
//   device-local buffer with UBO+6 SSBOs at offsets (ComputePipeline.zig:86-100)  -> seven cudaMalloc'd arrays
//   StagingRamp.transferToBuffer + flush (render/StagingRamp.zig:318-495)          -> cudaMemcpyAsync on the ctx stream
//   push constants + descriptor set (ComputePipeline.zig:488-545)                  -> one __grid_constant__ TraceParams
//   vkQueueSubmit + complete_fence (ComputePipeline.zig:423-459)                   -> stream order + cudaStreamSynchronize
// NCCL is loaded lazily with dlopen so that the library has no link-time dependency on it.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <new>

#include <cuda_runtime.h>
#include <nccl.h>

#include "../../include/vrt.h"
#include "vrt_kernels.cuh"

using namespace vrt;

namespace {

char g_init_error[512] = "no error";

// finite, normal, exactly a power of two: x / s == x * (1/s) bit for bit
bool is_pow2(float s) {
    uint32_t u;
    std::memcpy(&u, &s, 4);
    const uint32_t e = (u >> 23) & 0xffu;
    return (u & 0x807fffffu) == 0u && e > 1u && e < 253u;
}

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

bool load_nccl(char* err, size_t err_len) {
    if (g_nccl.handle) return true;
    // RTLD_NOLOAD first: reuse the copy the host process already mapped (e.g. PyTorch's bundled libnccl.so.2)
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        snprintf(err, err_len, "dlopen(libnccl.so.2) failed: %s", dlerror());
        return false;
    }
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(dlsym(h, "ncclAllGather"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(h, "ncclAllReduce"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.AllReduce || !g_nccl.CommDestroy || !g_nccl.GetErrorString) {
        snprintf(err, err_len, "libnccl is missing a required symbol");
        dlclose(h);
        return false;
    }
    g_nccl.handle = h;
    return true;
}

}  // namespace

struct vrt_ctx {
    vrt_config cfg;
    uint32_t row_begin, row_end;
    bool interleave = false;          // VRT_FLAG_INTERLEAVE: strips t % part_world == part_rank
    uint32_t part_rank = 0, part_world = 1, strips_max = 0;
    uint32_t* d_gather = nullptr;     // rank-major all-gather buffer of an interleaved partition
    int* d_barrier = nullptr;         // 4 bytes all-reduced after a peer-store frame
    int* h_barrier_error = nullptr;   // pinned, device-visible: set by the flag barrier when a peer never arrives
    uint32_t barrier_frame = 0;       // frames signalled through the flag barrier
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    bool timing_valid = false;

    // the seven reference buffers (Pipeline.zig:273-283)
    vrt_grid_state grid{};
    bool have_grid = false;
    vrt_material* d_materials = nullptr;
    uint32_t* d_statuses = nullptr;
    uint32_t* d_brick_indices = nullptr;
    uint8_t* d_occupancy = nullptr;
    uint32_t* d_start_indices = nullptr;
    uint8_t* d_material_indices = nullptr;
    size_t n_materials = 0, n_statuses = 0, n_brick_indices = 0, n_occupancy = 0, n_start_indices = 0, n_material_indices = 0;
    vrt_material* h_materials = nullptr;  // host mirror, to answer "is there a type-3 material" (:427 shortcut)

    // target image
    uint32_t* d_fb_own = nullptr;
    uint32_t* d_fb = nullptr;
    size_t fb_bytes = 0;
    // vrt_trace_to_host_async: second framebuffer + copy stream + per-slot events
    uint32_t* d_fb_ring1 = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_traced[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    bool slot_used[2] = {false, false};
    uint64_t stage_seq = 0; // frames exchanged tile-major (parity picks the staging buffer)
    uint64_t ring_seq = 0;  // frames that went through the two-slot ring (async frames; every frame in the peer-store modes)

    // post-process output (vrt_denoise)
    uint32_t* d_denoised = nullptr;
    float4* d_dn_decoded = nullptr;  // the traced image UNORM-decoded to float4 (scratch of the pass)
    uint32_t dn_width = 0, dn_height = 0;
    cudaEvent_t ev_dn_begin = nullptr, ev_dn_end = nullptr;
    bool dn_timing_valid = false;

    // debug
    vrt_aov* d_aov = nullptr;
    unsigned long long* d_counters = nullptr;

    // derived acceleration structures (vrt_trav_warp.cuh)
    uint4* d_cell_rec = nullptr;  // per grid cell: voxel mask + material start + brick index (brick_dim 4)
    uint8_t* d_dist = nullptr;
    uint8_t* d_dist_tmp = nullptr;  // 6 x n_bricks bytes of scratch for the separable distance scans
    size_t dist_plane = 0;          // bytes per octant
    uint32_t dist_log_px = 0, dist_log_pz = 0;
    uint32_t accel_dim[3] = {0, 0, 0};
    bool accel_dirty = true;   // status words were uploaded: the distance planes MAY be stale (d_accel_delta knows)
    bool occ_dirty = true;     // brick indices / occupancy bytes changed: the per-cell records are stale, the distance planes are not
    bool accel_force = true;   // the distance planes must be rebuilt from scratch (new grid, device-side insert)
    AccelDelta* d_accel_delta = nullptr;  // which status bits the uploads since the last build changed
    uint32_t* d_status_stage = nullptr;   // uploaded status words land here; status_merge_kernel compares them with the live ones

    // persistent-kernel work queue {next ticket, warps that left}; the kernel resets it itself
    unsigned long long* d_tile_counter = nullptr;

    // tile schedule (vrt_sched.cu): costs live behind the frame ring + flags in the IPC allocation so that peers can write them
    uint32_t sched_mode = VRT_SCHED_STATIC, sched_interval = 8;
    uint32_t* d_order = nullptr;
    uint32_t* d_sched_scratch = nullptr;
    uint16_t* d_cost[2] = {nullptr, nullptr};
    uint32_t tiles_global = 0;      // 8x4-pixel tiles of the whole image
    uint32_t sched_tiles = 0;       // tile space the current order covers (0: order not initialised)
    uint64_t sched_frames = 0;      // frames traced under the current schedule
    cudaEvent_t ev_kernel_end = nullptr;  // after the trace kernel, before the exchange

    // pinned staging ring of the uploads (render/StagingRamp.zig:98-175: N host-visible buffers filled round-robin)
    uint8_t* h_stage = nullptr;
    cudaEvent_t ev_stage[4] = {nullptr, nullptr, nullptr, nullptr};
    bool stage_used[4] = {false, false, false, false};
    int stage_next = 0;
    size_t stage_off = 0;  // fill level of buffer stage_next

    // multi-GPU
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    uint32_t exchange_mode = VRT_EXCHANGE_ALLGATHER;
    void* peer_fb[8] = {nullptr};
    bool peers_open = false;

    uint32_t last_launches = 0;
    char err[512] = "no error";
};

namespace {

int fail(vrt_ctx* ctx, int code, const char* fmt, ...) {
    char* dst = ctx ? ctx->err : g_init_error;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define VRT_CUDA(ctx, call)                                                                                   \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess) {                                                                             \
            const int code__ = (e__ == cudaErrorMemoryAllocation) ? VRT_E_OOM : VRT_E_CUDA;                   \
            return fail(ctx, code__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
        }                                                                                                     \
    } while (0)

// Uploads go through a ring of pinned staging buffers, like the reference's StagingRamp (render/StagingRamp.zig:98-175,318-360):
// the caller's bytes are copied into the next free buffer and DMA'd from there on the ctx stream, so the call returns as soon
// as the bytes are staged — whatever memory `src` is (pageable, pinned, mapped) the caller may overwrite it immediately and is
// never blocked by the transfer itself; a buffer is reused only after the copy that read it has finished.
constexpr size_t kStageChunk = 16u << 20;
constexpr int kStageBuffers = 4;
int stage_upload(vrt_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx->h_stage) {
        if (cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_stage), kStageChunk * kStageBuffers, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            ctx->h_stage = nullptr;
            return fail(ctx, VRT_E_OOM, "cannot allocate the %zu-byte pinned staging ring", kStageChunk * kStageBuffers);
        }
        for (int i = 0; i < kStageBuffers; i++) VRT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_stage[i], cudaEventDisableTiming));
    }
    const uint8_t* s = static_cast<const uint8_t*>(src);
    uint8_t* d = static_cast<uint8_t*>(dst);
    while (bytes) {
        if (ctx->stage_off >= kStageChunk) {  // this buffer is full: on to the next one, once the copies that read it are done
            ctx->stage_next = (ctx->stage_next + 1) % kStageBuffers;
            ctx->stage_off = 0;
            if (ctx->stage_used[ctx->stage_next]) VRT_CUDA(ctx, cudaEventSynchronize(ctx->ev_stage[ctx->stage_next]));
            ctx->stage_used[ctx->stage_next] = false;
        }
        const int b = ctx->stage_next;
        const size_t room = kStageChunk - ctx->stage_off;
        const size_t n = bytes < room ? bytes : room;
        uint8_t* stage = ctx->h_stage + (size_t)b * kStageChunk + ctx->stage_off;
        std::memcpy(stage, s, n);
        VRT_CUDA(ctx, cudaMemcpyAsync(d, stage, n, cudaMemcpyHostToDevice, ctx->stream));
        VRT_CUDA(ctx, cudaEventRecord(ctx->ev_stage[b], ctx->stream));  // = the latest copy out of buffer b
        ctx->stage_used[b] = true;
        ctx->stage_off += (n + 255u) & ~(size_t)255u;
        s += n, d += n, bytes -= n;
    }
    return VRT_OK;
}

template <class T>
int upload_range(vrt_ctx* ctx, T* dst, size_t capacity, size_t offset, const T* data, size_t count, const char* what) {
    if (!ctx) return VRT_E_INVALID;
    if (count == 0) return VRT_OK;
    if (!data) return fail(ctx, VRT_E_INVALID, "%s: data is NULL", what);
    if (offset > capacity || count > capacity - offset)
        return fail(ctx, VRT_E_RANGE, "%s: [%zu, %zu) outside capacity %zu", what, offset, offset + count, capacity);
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    return stage_upload(ctx, dst + offset, data, count * sizeof(T));
}

// Bring the per-cell records / the distance planes up to date with the uploads (stream-ordered, nothing waits).
int rebuild_accel(vrt_ctx* ctx, const TraceParams& P, LaunchInfo* info) {
    const size_t n_cells = (size_t)ctx->grid.dim_x * ctx->grid.dim_y * ctx->grid.dim_z;
    if (ctx->accel_force) {
        VRT_CUDA(ctx, cudaMemsetAsync(&ctx->d_accel_delta->force_full, 1, 1, ctx->stream));  // little-endian: the word becomes >= 1
        ctx->accel_dirty = true;
    }
    VRT_CUDA(ctx, launch_build_accel(P, ctx->d_cell_rec, ctx->d_dist, ctx->d_dist_tmp, n_cells, !ctx->accel_dirty, ctx->d_accel_delta, ctx->stream, info));
    ctx->accel_dirty = ctx->occ_dirty = ctx->accel_force = false;
    return VRT_OK;
}

// the tile-major exchange is for the tuned kernel on partitions aligned to the image's 4-row strips
bool tiles_exchange_ok(const vrt_ctx* c) {
    return !(c->cfg.flags & (VRT_FLAG_AOV | VRT_FLAG_BASELINE)) &&
           (c->interleave || (c->row_begin % kStripRows == 0 && (c->row_end % kStripRows == 0 || c->row_end == c->cfg.height)));
}
size_t cost_bytes(const vrt_ctx* c) { return ((size_t)c->tiles_global * sizeof(uint16_t) + 255u) & ~(size_t)255u; }
size_t stage_bytes(const vrt_ctx* c) { return (size_t)c->tiles_global * 128u; }  // tile-major staging of one frame (VRT_EXCHANGE_PEER_TILES)
size_t stage_offset(const vrt_ctx* c, int parity) { return 2 * c->fb_bytes + kPeerFlagBytes + 2 * cost_bytes(c) + (size_t)parity * stage_bytes(c); }

void fill_params(const vrt_ctx* c, const vrt_camera* cam, const vrt_sun* sun, TraceParams& P) {
    std::memset(&P, 0, sizeof(P));
    P.cam = *cam;
    P.sun = *sun;
    P.grid = c->grid;
    P.materials = c->d_materials;
    P.statuses = c->d_statuses;
    P.brick_indices = c->d_brick_indices;
    P.occupancy = c->d_occupancy;
    P.start_indices = c->d_start_indices;
    P.material_indices = c->d_material_indices;
    P.n_statuses = c->n_statuses, P.n_brick_indices = c->n_brick_indices, P.n_occupancy = c->n_occupancy;
    P.n_start_indices = c->n_start_indices, P.n_material_indices = c->n_material_indices;
    P.n_materials = (uint32_t)c->n_materials;
    P.materials_have_none = 0, P.materials_basic = 1;
    for (size_t i = 0; i < c->n_materials; i++) {
        if (c->h_materials[i].type == VRT_MAT_NONE) P.materials_have_none = 1;
        if (c->h_materials[i].type > VRT_MAT_DIELECTRIC) P.materials_basic = 0;
    }
    P.brick_dim = (int)c->cfg.brick_dim;
    P.brick_bytes = c->cfg.brick_dim * c->cfg.brick_dim * c->cfg.brick_dim / 8;
    P.brick_voxel_scale = 1.0f / (float)c->cfg.brick_dim;  // Pipeline.zig:313
    P.row_begin = c->row_begin, P.row_end = c->row_end;
    P.il_world = c->interleave ? c->part_world : 0u;
    P.il_rank = c->part_rank;
    P.il_gather = 0u, P.il_strips_max = c->strips_max;
    P.fb = c->d_fb;
    P.aov = c->d_aov;
    P.counters = c->d_counters;
    P.cell_rec = c->d_cell_rec;
    P.dist = c->d_dist;
    P.dist_plane = c->dist_plane;
    P.dist_log_px = c->dist_log_px, P.dist_log_pz = c->dist_log_pz;
    const float scale = c->grid.max_point_scale[3];
    const float voxel_scale = scale * P.brick_voxel_scale;  // :389, same f32 product the kernels form
    P.scale_pow2 = is_pow2(scale) ? 1u : 0u;
    P.voxel_scale_pow2 = is_pow2(voxel_scale) ? 1u : 0u;
    P.inv_scale = P.scale_pow2 ? 1.0f / scale : 0.0f;
    P.inv_voxel_scale = P.voxel_scale_pow2 ? 1.0f / voxel_scale : 0.0f;
    P.tile_counter = c->d_tile_counter;
    P.vec_store_ok = (cam->image_width % 4 == 0) && ((reinterpret_cast<uintptr_t>(c->d_fb) & 15u) == 0);
    P.n_peers = 0;
    P.one = 1u;
    if (c->world > 1 && (c->exchange_mode == VRT_EXCHANGE_PEER_STORE || c->exchange_mode == VRT_EXCHANGE_PEER_FLAGS || c->exchange_mode == VRT_EXCHANGE_PEER_PUSH ||
                         c->exchange_mode == VRT_EXCHANGE_PEER_TILES) && c->peers_open) {
        const size_t slot_words = (c->d_fb == c->d_fb_ring1) ? c->fb_bytes / 4 : 0;  // same ring slot on every rank
        for (int r = 0; r < c->world; r++)
            if (r != c->rank) P.peer_fb[P.n_peers++] = static_cast<uint32_t*>(c->peer_fb[r]) + slot_words;
        if (c->exchange_mode == VRT_EXCHANGE_PEER_TILES && tiles_exchange_ok(c)) {
            // tile-major exchange: every rank's staging buffer of this frame's parity, this rank's own included (peer_fb[rank] is d_fb_own)
            for (int r = 0; r < c->world; r++)
                P.stage[P.n_stage++] = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(c->peer_fb[r]) + stage_offset(c, (int)(c->stage_seq & 1u)));
        }
    }
    // tile schedule (tuned kernel, not the counting / AOV variant)
    if (c->sched_mode != VRT_SCHED_STATIC && !(c->cfg.flags & (VRT_FLAG_AOV | VRT_FLAG_BASELINE))) {
        const int parity = (int)(c->sched_frames & 1u);
        P.tile_order = c->d_order;
        P.tile_cost = c->d_cost[parity];
        P.order_offset = 0u, P.order_stride = 1u;
        if (c->sched_mode == VRT_SCHED_DEAL || c->sched_mode == VRT_SCHED_SHARED) {
            // the whole image is one tile space: this rank takes every part_world-th entry of the order (DEAL), or whatever ticket its
            // warps draw from the ONE queue all ranks share — rank 0's, reached through the peer mapping (SHARED)
            if (c->sched_mode == VRT_SCHED_DEAL) {
                P.order_offset = c->part_rank, P.order_stride = c->part_world;
            } else if (c->world > 1 && P.n_peers) {
                P.tile_counter = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(c->peer_fb[0]) + 2 * c->fb_bytes + kSharedQueueOffset);
                P.queue_world = (uint32_t)c->world;
            }
            P.il_world = 0u, P.il_gather = 0u;
            P.row_begin = 0u, P.row_end = c->cfg.height;
            const size_t cost_off = 2 * c->fb_bytes + kPeerFlagBytes + (size_t)parity * cost_bytes(c);
            for (int r = 0; r < c->world && P.n_peers; r++)
                if (r != c->rank) P.peer_cost[P.n_cost_peers++] = reinterpret_cast<uint16_t*>(static_cast<uint8_t*>(c->peer_fb[r]) + cost_off);
        }
    }
}

// (Re)allocate the distance grid for the grid dimensions in ctx->grid.
int ensure_accel(vrt_ctx* ctx) {
    const uint32_t dx = ctx->grid.dim_x, dy = ctx->grid.dim_y, dz = ctx->grid.dim_z;
    if (ctx->d_dist && dx == ctx->accel_dim[0] && dy == ctx->accel_dim[1] && dz == ctx->accel_dim[2]) return VRT_OK;
    if (ctx->d_dist) cudaFree(ctx->d_dist);
    if (ctx->d_dist_tmp) cudaFree(ctx->d_dist_tmp);
    ctx->d_dist = nullptr, ctx->d_dist_tmp = nullptr;
    uint32_t lx = 1, lz = 1;
    while ((1u << lx) < dx + 2) lx++;  // one border cell on each side; power-of-two strides make the index decodable with shifts
    while ((1u << lz) < dz + 2) lz++;
    ctx->dist_log_px = lx, ctx->dist_log_pz = lz;
    ctx->dist_plane = ((size_t)(dy + 2)) << (lx + lz);
    if (8 * ctx->dist_plane > 0x7fffffffull)
        return fail(ctx, VRT_E_INVALID, "grid %ux%ux%u is too large for the 31-bit cell index of the march (8 padded distance planes)", dx, dy, dz);
    VRT_CUDA(ctx, cudaMalloc(&ctx->d_dist, 8 * ctx->dist_plane));
    VRT_CUDA(ctx, cudaMalloc(&ctx->d_dist_tmp, 6 * (size_t)dx * dy * dz));
    VRT_CUDA(ctx, cudaMemsetAsync(ctx->d_dist, 255, 8 * ctx->dist_plane, ctx->stream));  // the border; interiors are rewritten by every build
    ctx->accel_dim[0] = dx, ctx->accel_dim[1] = dy, ctx->accel_dim[2] = dz;
    ctx->accel_dirty = ctx->accel_force = true;
    return VRT_OK;
}

}  // namespace

extern "C" {

const char* vrt_last_error(const vrt_ctx* ctx) { return ctx ? ctx->err : g_init_error; }

int vrt_init(vrt_ctx** out_ctx, const vrt_config* cfg) {
    if (!out_ctx) return fail(nullptr, VRT_E_INVALID, "vrt_init: out_ctx is NULL");
    *out_ctx = nullptr;
    if (!cfg) return fail(nullptr, VRT_E_INVALID, "vrt_init: cfg is NULL");
    if (cfg->struct_size != sizeof(vrt_config) || cfg->abi_version != VRT_ABI_VERSION)
        return fail(nullptr, VRT_E_INVALID, "vrt_init: ABI mismatch (struct_size %u vs %zu, version %u vs %u)", cfg->struct_size,
                    sizeof(vrt_config), cfg->abi_version, VRT_ABI_VERSION);
    if (cfg->width == 0 || cfg->height == 0 || cfg->width > 32768 || cfg->height > 32768)
        return fail(nullptr, VRT_E_INVALID, "vrt_init: bad image size %ux%u", cfg->width, cfg->height);
    if (cfg->brick_dim != 4 && cfg->brick_dim != 8 && cfg->brick_dim != 16)
        return fail(nullptr, VRT_E_INVALID, "vrt_init: brick_dim must be 4, 8 or 16 (got %u)", cfg->brick_dim);
    if (cfg->n_bricks == 0 || cfg->n_bricks > 0xffffffffull)
        return fail(nullptr, VRT_E_INVALID, "vrt_init: n_bricks out of range");
    if (cfg->row_begin > cfg->row_end || cfg->row_end > cfg->height)
        return fail(nullptr, VRT_E_INVALID, "vrt_init: bad row slab [%u, %u) for height %u", cfg->row_begin, cfg->row_end, cfg->height);
    if (cfg->flags & VRT_FLAG_INTERLEAVE) {
        if (cfg->row_begin != 0 || cfg->row_end != 0) return fail(nullptr, VRT_E_INVALID, "vrt_init: VRT_FLAG_INTERLEAVE excludes a row slab");
        if (cfg->part_world < 1 || cfg->part_world > 8 || cfg->part_rank >= cfg->part_world)
            return fail(nullptr, VRT_E_INVALID, "vrt_init: bad interleaved partition %u of %u", cfg->part_rank, cfg->part_world);
        if (cfg->flags & VRT_FLAG_BASELINE) return fail(nullptr, VRT_E_INVALID, "vrt_init: the baseline kernel only supports row slabs");
    }

    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(nullptr, VRT_E_CUDA, "vrt_init: no CUDA device (%s) — this library has no CPU fallback", cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= n_dev) return fail(nullptr, VRT_E_INVALID, "vrt_init: device %d of %d", cfg->device, n_dev);

    vrt_ctx* ctx = new (std::nothrow) vrt_ctx();
    if (!ctx) return fail(nullptr, VRT_E_OOM, "vrt_init: host allocation failed");
    ctx->cfg = *cfg;
    if (ctx->cfg.material_capacity == 0) ctx->cfg.material_capacity = 256;  // Pipeline.Config.material_buffer
    if (ctx->cfg.n_brick_alloc == 0) ctx->cfg.n_brick_alloc = cfg->n_bricks;  // Grid.zig:51
    ctx->row_begin = cfg->row_begin;
    ctx->row_end = (cfg->row_begin == 0 && cfg->row_end == 0) ? cfg->height : cfg->row_end;
    if (cfg->flags & VRT_FLAG_INTERLEAVE) {
        ctx->interleave = true;
        ctx->part_rank = cfg->part_rank, ctx->part_world = cfg->part_world;
        const uint32_t strips = (cfg->height + kStripRows - 1) / kStripRows;
        ctx->strips_max = (strips + cfg->part_world - 1) / cfg->part_world;
    }

    const uint32_t bits = cfg->brick_dim * cfg->brick_dim * cfg->brick_dim;
    ctx->n_materials = ctx->cfg.material_capacity;
    ctx->n_statuses = (cfg->n_bricks + 31) / 32;                      // Grid.zig:43
    ctx->n_brick_indices = cfg->n_bricks;                             // Grid.zig:47
    ctx->n_occupancy = ctx->cfg.n_brick_alloc * (bits / 8);           // Grid.zig:53
    ctx->n_start_indices = ctx->cfg.n_brick_alloc;                    // Grid.zig:57
    ctx->n_material_indices = ctx->cfg.n_brick_alloc * (size_t)bits;  // Grid.zig:61-62
    ctx->fb_bytes = (size_t)cfg->width * cfg->height * 4;
    ctx->tiles_global = ((cfg->width + 7u) / 8u) * ((cfg->height + kStripRows - 1u) / kStripRows);

#define INIT_CUDA(call)                                                                                          \
    do {                                                                                                         \
        cudaError_t e__ = (call);                                                                                \
        if (e__ != cudaSuccess) {                                                                                \
            const int code__ = (e__ == cudaErrorMemoryAllocation) ? VRT_E_OOM : VRT_E_CUDA;                      \
            fail(nullptr, code__, "vrt_init: %s failed: %s", #call, cudaGetErrorString(e__));                    \
            vrt_deinit(ctx);                                                                                     \
            return code__;                                                                                       \
        }                                                                                                        \
    } while (0)

    INIT_CUDA(cudaSetDevice(cfg->device));
    INIT_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    INIT_CUDA(cudaEventCreate(&ctx->ev_begin));
    INIT_CUDA(cudaEventCreate(&ctx->ev_end));
    INIT_CUDA(cudaMalloc(&ctx->d_materials, ctx->n_materials * sizeof(vrt_material)));
    INIT_CUDA(cudaMalloc(&ctx->d_statuses, ctx->n_statuses * 4));
    INIT_CUDA(cudaMalloc(&ctx->d_status_stage, ctx->n_statuses * 4));
    INIT_CUDA(cudaMalloc(&ctx->d_accel_delta, sizeof(AccelDelta)));
    INIT_CUDA(cudaMemsetAsync(ctx->d_accel_delta, 0, sizeof(AccelDelta), ctx->stream));
    INIT_CUDA(cudaMalloc(&ctx->d_brick_indices, ctx->n_brick_indices * 4));
    INIT_CUDA(cudaMalloc(&ctx->d_occupancy, ctx->n_occupancy));
    INIT_CUDA(cudaMalloc(&ctx->d_start_indices, ctx->n_start_indices * 4));
    INIT_CUDA(cudaMalloc(&ctx->d_material_indices, ctx->n_material_indices));
    // slot 0 + slot 1 of the frame ring + barrier flags + the two tile-cost arrays + two tile-major staging frames: one allocation = one IPC handle
    const size_t shared_bytes = 2 * ctx->fb_bytes + kPeerFlagBytes + 2 * cost_bytes(ctx) + 2 * stage_bytes(ctx);
    INIT_CUDA(cudaMalloc(&ctx->d_fb_own, shared_bytes));
    ctx->d_fb_ring1 = ctx->d_fb_own + ctx->fb_bytes / 4;
    ctx->d_cost[0] = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(ctx->d_fb_own) + 2 * ctx->fb_bytes + kPeerFlagBytes);
    ctx->d_cost[1] = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(ctx->d_cost[0]) + cost_bytes(ctx));
    INIT_CUDA(cudaMalloc(&ctx->d_order, (size_t)ctx->tiles_global * 4));
    INIT_CUDA(cudaMalloc(&ctx->d_sched_scratch, sched_scratch_words(ctx->tiles_global) * 4));
    INIT_CUDA(cudaEventCreate(&ctx->ev_kernel_end));
    INIT_CUDA(cudaMalloc(&ctx->d_tile_counter, 16));
    ctx->d_fb = ctx->d_fb_own;
    ctx->h_materials = new (std::nothrow) vrt_material[ctx->n_materials]();
    if (!ctx->h_materials) {
        vrt_deinit(ctx);
        return fail(nullptr, VRT_E_OOM, "vrt_init: host allocation failed");
    }
    // Same initial contents as Grid.init gives the host arrays (Grid.zig:45-64); the reference leaves
    // the device buffer uninitialised until the first transfer.
    INIT_CUDA(cudaMemsetAsync(ctx->d_materials, 0, ctx->n_materials * sizeof(vrt_material), ctx->stream));
    INIT_CUDA(cudaMemsetAsync(ctx->d_statuses, 0, ctx->n_statuses * 4, ctx->stream));
    INIT_CUDA(cudaMemsetAsync(ctx->d_brick_indices, 0, ctx->n_brick_indices * 4, ctx->stream));
    INIT_CUDA(cudaMemsetAsync(ctx->d_occupancy, 0, ctx->n_occupancy, ctx->stream));
    INIT_CUDA(cudaMemsetAsync(ctx->d_start_indices, 0xff, ctx->n_start_indices * 4, ctx->stream));
    INIT_CUDA(cudaMemsetAsync(ctx->d_material_indices, 0, ctx->n_material_indices, ctx->stream));
    INIT_CUDA(cudaMemsetAsync(ctx->d_fb_own, 0, shared_bytes, ctx->stream));
    INIT_CUDA(cudaMemsetAsync(ctx->d_tile_counter, 0, 16, ctx->stream));
    if (cfg->brick_dim == 4) INIT_CUDA(cudaMalloc(&ctx->d_cell_rec, cfg->n_bricks * sizeof(uint4)));
    if (cfg->flags & VRT_FLAG_AOV) {
        INIT_CUDA(cudaMalloc(&ctx->d_aov, (size_t)cfg->width * cfg->height * sizeof(vrt_aov)));
        INIT_CUDA(cudaMalloc(&ctx->d_counters, 8 * sizeof(unsigned long long)));
        INIT_CUDA(cudaMemsetAsync(ctx->d_aov, 0, (size_t)cfg->width * cfg->height * sizeof(vrt_aov), ctx->stream));
        INIT_CUDA(cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
    }
    INIT_CUDA(cudaStreamSynchronize(ctx->stream));
#undef INIT_CUDA
    *out_ctx = ctx;
    return VRT_OK;
}

void vrt_deinit(vrt_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);  // ComputePipeline.deinit waits for the fence first (:386-393)
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    if (ctx->peers_open) {
        for (int r = 0; r < ctx->world; r++)
            if (r != ctx->rank && ctx->peer_fb[r]) cudaIpcCloseMemHandle(ctx->peer_fb[r]);
    }
    cudaFree(ctx->d_materials), cudaFree(ctx->d_statuses), cudaFree(ctx->d_brick_indices), cudaFree(ctx->d_occupancy);
    cudaFree(ctx->d_start_indices), cudaFree(ctx->d_material_indices), cudaFree(ctx->d_fb_own), cudaFree(ctx->d_aov);
    cudaFree(ctx->d_counters), cudaFree(ctx->d_cell_rec), cudaFree(ctx->d_dist), cudaFree(ctx->d_dist_tmp);
    cudaFree(ctx->d_order), cudaFree(ctx->d_sched_scratch), cudaFree(ctx->d_status_stage), cudaFree(ctx->d_accel_delta);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    for (int i = 0; i < 4; i++)
        if (ctx->ev_stage[i]) cudaEventDestroy(ctx->ev_stage[i]);
    if (ctx->ev_kernel_end) cudaEventDestroy(ctx->ev_kernel_end);
    cudaFree(ctx->d_tile_counter), cudaFree(ctx->d_gather), cudaFree(ctx->d_barrier), cudaFree(ctx->d_denoised), cudaFree(ctx->d_dn_decoded);
    if (ctx->h_barrier_error) cudaFreeHost(ctx->h_barrier_error);
    if (ctx->ev_dn_begin) cudaEventDestroy(ctx->ev_dn_begin);
    if (ctx->ev_dn_end) cudaEventDestroy(ctx->ev_dn_end);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    for (int i = 0; i < 2; i++) {
        if (ctx->ev_traced[i]) cudaEventDestroy(ctx->ev_traced[i]);
        if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
    if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete[] ctx->h_materials;
    delete ctx;
}

int vrt_upload_grid_state(vrt_ctx* ctx, const vrt_grid_state* state) {
    if (!ctx) return VRT_E_INVALID;
    if (!state) return fail(ctx, VRT_E_INVALID, "vrt_upload_grid_state: state is NULL");
    const uint64_t bricks = (uint64_t)state->dim_x * state->dim_y * state->dim_z;
    if (bricks == 0 || bricks > ctx->cfg.n_bricks)
        return fail(ctx, VRT_E_RANGE, "vrt_upload_grid_state: %ux%ux%u bricks exceed the %llu sized at init", state->dim_x, state->dim_y,
                    state->dim_z, (unsigned long long)ctx->cfg.n_bricks);
    if (state->dim_x > 4096 || state->dim_y > 4096 || state->dim_z > 4096)
        return fail(ctx, VRT_E_INVALID, "vrt_upload_grid_state: grid dimension above 4096 bricks");
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    if (ctx->have_grid && (state->dim_x != ctx->grid.dim_x || state->dim_y != ctx->grid.dim_y || state->dim_z != ctx->grid.dim_z)) ctx->accel_force = true;
    ctx->grid = *state;
    ctx->have_grid = true;
    return ensure_accel(ctx);
}

int vrt_upload_materials(vrt_ctx* ctx, size_t offset, const vrt_material* data, size_t count) {
    const int rc = upload_range(ctx, ctx ? ctx->d_materials : nullptr, ctx ? ctx->n_materials : 0, offset, data, count, "vrt_upload_materials");
    if (rc == VRT_OK && count) std::memcpy(ctx->h_materials + offset, data, count * sizeof(vrt_material));
    return rc;
}
int vrt_upload_brick_statuses(vrt_ctx* ctx, size_t offset, const uint32_t* data, size_t count) {
    if (!ctx) return VRT_E_INVALID;
    if (!ctx->have_grid) {  // no grid dimensions yet to place the changed bits in: plain copy, full rebuild later
        const int rc = upload_range(ctx, ctx->d_statuses, ctx->n_statuses, offset, data, count, "vrt_upload_brick_statuses");
        if (rc == VRT_OK && count) ctx->accel_dirty = ctx->accel_force = true;
        return rc;
    }
    // Grid.insert re-registers the status word of every voxel it touches (brick/Grid.zig:188-189), so the per-frame delta upload
    // (VoxelRT.zig:112-123) mostly carries words that are already there.  The words go to a staging copy and a kernel merges
    // them, recording which bits really changed: the next trace then leaves the distance planes alone, patches them for a few new
    // bricks, or rebuilds them — decided on the device, the upload never waits for the answer.
    const int rc = upload_range(ctx, ctx->d_status_stage, ctx->n_statuses, offset, data, count, "vrt_upload_brick_statuses");
    if (rc != VRT_OK || count == 0) return rc;
    VRT_CUDA(ctx, launch_status_merge(ctx->d_statuses + offset, ctx->d_status_stage + offset, count, offset, ctx->grid, ctx->d_accel_delta, ctx->stream, nullptr));
    ctx->accel_dirty = true;
    return VRT_OK;
}
int vrt_upload_brick_indices(vrt_ctx* ctx, size_t offset, const uint32_t* data, size_t count) {
    const int rc = upload_range(ctx, ctx ? ctx->d_brick_indices : nullptr, ctx ? ctx->n_brick_indices : 0, offset, data, count, "vrt_upload_brick_indices");
    if (rc == VRT_OK && count) ctx->occ_dirty = true;  // the per-cell records are indexed through brick_indices; the distance planes only read the status bits
    return rc;
}
int vrt_upload_brick_occupancy(vrt_ctx* ctx, size_t offset, const uint8_t* data, size_t count) {
    const int rc = upload_range(ctx, ctx ? ctx->d_occupancy : nullptr, ctx ? ctx->n_occupancy : 0, offset, data, count, "vrt_upload_brick_occupancy");
    if (rc == VRT_OK && count) ctx->occ_dirty = true;  // voxels inside bricks: the distance planes (brick level) are unaffected
    return rc;
}
int vrt_upload_brick_start_indices(vrt_ctx* ctx, size_t offset, const uint32_t* data, size_t count) {
    const int rc = upload_range(ctx, ctx ? ctx->d_start_indices : nullptr, ctx ? ctx->n_start_indices : 0, offset, data, count, "vrt_upload_brick_start_indices");
    if (rc == VRT_OK && count) ctx->occ_dirty = true;  // the per-cell records carry the material start index
    return rc;
}
int vrt_upload_material_indices(vrt_ctx* ctx, size_t offset, const uint8_t* data, size_t count) {
    return upload_range(ctx, ctx ? ctx->d_material_indices : nullptr, ctx ? ctx->n_material_indices : 0, offset, data, count, "vrt_upload_material_indices");
}

namespace {

bool peer_mode(const vrt_ctx* c) {
    return c->world > 1 && c->peers_open &&
           (c->exchange_mode == VRT_EXCHANGE_PEER_STORE || c->exchange_mode == VRT_EXCHANGE_PEER_FLAGS || c->exchange_mode == VRT_EXCHANGE_PEER_PUSH ||
            c->exchange_mode == VRT_EXCHANGE_PEER_TILES);
}
bool scattered(const vrt_ctx* c) { return c->sched_mode == VRT_SCHED_DEAL || c->sched_mode == VRT_SCHED_SHARED; }
bool owns_fb(const vrt_ctx* c) { return c->d_fb == c->d_fb_own || c->d_fb == c->d_fb_ring1; }

int ensure_ring(vrt_ctx* ctx) {  // copy stream + per-slot events of the two-slot frame ring
    if (ctx->copy_stream) return VRT_OK;
    VRT_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        VRT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_traced[i], cudaEventDisableTiming));
        VRT_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming));
    }
    return VRT_OK;
}

// The frame about to be traced goes into slot `slot` of the ring (or stays in the attached image): make the stream wait until the
// slot's previous contents have reached the host (an earlier vrt_trace_to_host_async may still be copying them).
int claim_slot(vrt_ctx* ctx, int slot) {
    if (ctx->slot_used[slot]) {
        VRT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[slot], 0));
        ctx->slot_used[slot] = false;
    }
    ctx->d_fb = slot ? ctx->d_fb_ring1 : ctx->d_fb_own;
    return VRT_OK;
}

// One frame: (derived structures) -> trace kernel -> exchange -> (schedule).  `ring`: this frame uses the ring slot ring_seq & 1.
int trace_frame(vrt_ctx* ctx, const vrt_camera* camera, const vrt_sun* sun, bool ring) {
    if (!camera || !sun) return fail(ctx, VRT_E_INVALID, "vrt_trace: camera or sun is NULL");
    if (!ctx->have_grid) return fail(ctx, VRT_E_STATE, "vrt_trace: vrt_upload_grid_state has not been called");
    if (camera->image_width != ctx->cfg.width || camera->image_height != ctx->cfg.height)
        return fail(ctx, VRT_E_INVALID, "vrt_trace: camera image %ux%u != target image %ux%u", camera->image_width, camera->image_height,
                    ctx->cfg.width, ctx->cfg.height);
    if (camera->samples_per_pixel < 1) return fail(ctx, VRT_E_INVALID, "vrt_trace: samples_per_pixel < 1");
    const bool pipelined = ring;  // vrt_trace_to_host_async: nobody can ask for this frame's device time, so its three timing events are not recorded
    const bool gather = ctx->world > 1 && ctx->exchange_mode == VRT_EXCHANGE_ALLGATHER;
    if (ctx->world > 1 && !ctx->comm && ctx->exchange_mode != VRT_EXCHANGE_PEER_FLAGS && ctx->exchange_mode != VRT_EXCHANGE_PEER_PUSH &&
        ctx->exchange_mode != VRT_EXCHANGE_PEER_TILES && ctx->exchange_mode != VRT_EXCHANGE_HOST)  // none of these needs a communicator
        return fail(ctx, VRT_E_STATE, "vrt_trace: world > 1 but vrt_comm_init has not been called");
    if (scattered(ctx) && ctx->world > 1 && !peer_mode(ctx))
        return fail(ctx, VRT_E_STATE, "vrt_trace: VRT_SCHED_DEAL / SHARED scatter a rank's tiles over the image and need a peer-store exchange mode");
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));

    // Frame ring.  Async frames always use it.  In the peer-store modes EVERY frame does: a peer's frame k+1 stores straight into this
    // rank's framebuffer, and the frame barrier only orders them after this rank's TRACE of frame k — with one buffer they could land
    // while the host is still reading frame k.  With two slots, frame k+1 goes to the other slot and frame k+2 cannot start on any rank
    // before this rank has traced k+1, which its host does after it has finished with frame k.
    if (ring || (peer_mode(ctx) && owns_fb(ctx))) {
        const int rc = ensure_ring(ctx);
        if (rc != VRT_OK) return rc;
        const int slot = (int)(ctx->ring_seq & 1u);
        const int rc2 = claim_slot(ctx, slot);
        if (rc2 != VRT_OK) return rc2;
        ring = true;
    } else if (owns_fb(ctx) && ctx->copy_stream) {  // a blocking frame after async ones: same hazard on whatever slot d_fb points at
        const int rc = claim_slot(ctx, ctx->d_fb == ctx->d_fb_ring1 ? 1 : 0);
        if (rc != VRT_OK) return rc;
    }

    TraceParams P;
    fill_params(ctx, camera, sun, P);
    LaunchInfo info = {0u};
    const bool aov = (ctx->cfg.flags & VRT_FLAG_AOV) != 0;
    const TraceKernel which = (ctx->cfg.flags & VRT_FLAG_BASELINE) ? KERNEL_REF : KERNEL_TUNED;

    if (!pipelined) VRT_CUDA(ctx, cudaEventRecord(ctx->ev_begin, ctx->stream));
    if (which == KERNEL_TUNED && (ctx->accel_dirty || ctx->occ_dirty)) {  // (vrt_trace_rays does the same)
        // Uploads changed statuses / indices / occupancy: rebuild the derived structures before tracing.  Stream order gives
        // upload -> build -> trace, where the reference has no barrier at all between its staging copy and the next dispatch
        // (edits land one frame late, Pipeline.zig:540).  Occupancy-only edits (voxels inside loaded bricks — most of
        // Grid.insert's traffic) leave the brick-level distance planes alone.
        const int rcb = rebuild_accel(ctx, P, &info);
        if (rcb != VRT_OK) return rcb;
    }
    if (aov) VRT_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
    if (gather && ctx->interleave) {  // trace straight into this rank's slice of the rank-major gather buffer
        P.fb = ctx->d_gather;
        P.il_gather = 1u;
        P.vec_store_ok = (camera->image_width % 4 == 0) ? 1u : 0u;
    }
    if (P.tile_order) {  // a schedule is attached: (re)initialise it when the tile space it permutes has changed
        TraceParams Q = P;
        Q.tile_order = nullptr;
        const uint32_t space = trace_tile_space(Q);
        if (space != ctx->sched_tiles) {
            VRT_CUDA(ctx, launch_sched_init(ctx->d_order, space, ctx->stream, &info));
            ctx->sched_tiles = space, ctx->sched_frames = 0;
            P.tile_cost = ctx->d_cost[0];
            if (scattered(ctx) && P.n_peers) {  // parity 0 again
                const size_t cost_off = 2 * ctx->fb_bytes + kPeerFlagBytes;
                uint32_t p = 0;
                for (int r = 0; r < ctx->world; r++)
                    if (r != ctx->rank) P.peer_cost[p++] = reinterpret_cast<uint16_t*>(static_cast<uint8_t*>(ctx->peer_fb[r]) + cost_off);
            }
        }
    }
    // (with a shared queue a rank does not know in advance which tiles it will trace: the stores stay in the trace kernel)
    const bool push = peer_mode(ctx) && ctx->exchange_mode == VRT_EXCHANGE_PEER_PUSH && P.vec_store_ok && which == KERNEL_TUNED && !aov && ctx->sched_mode != VRT_SCHED_SHARED;
    if (push) {  // the trace kernel keeps its pixels local (the costs of a dealt schedule still go to the peers); a copy kernel ships them afterwards
        TraceParams Q = P;
        Q.n_peers = 0u;
        VRT_CUDA(ctx, launch_trace(Q, which, aov, ctx->stream, &info));
    } else {
        VRT_CUDA(ctx, launch_trace(P, which, aov, ctx->stream, &info));
    }
    if (!pipelined) VRT_CUDA(ctx, cudaEventRecord(ctx->ev_kernel_end, ctx->stream));
    if (push) VRT_CUDA(ctx, launch_push_tiles(P, ctx->stream, &info));

    if (gather) {
        // in place: the kernel already wrote this rank's pixels at their offset in the gathered buffer
        if (ctx->interleave) {
            const size_t slab_bytes = (size_t)ctx->strips_max * kStripRows * ctx->cfg.width * 4;
            const uint8_t* send = reinterpret_cast<const uint8_t*>(ctx->d_gather) + (size_t)ctx->rank * slab_bytes;
            const ncclResult_t r = g_nccl.AllGather(send, ctx->d_gather, slab_bytes, ncclUint8, ctx->comm, ctx->stream);
            if (r != ncclSuccess) return fail(ctx, VRT_E_NCCL, "ncclAllGather failed: %s", g_nccl.GetErrorString(r));
            VRT_CUDA(ctx, launch_deinterleave(ctx->d_gather, ctx->d_fb, ctx->cfg.width, ctx->cfg.height, (uint32_t)ctx->world, ctx->strips_max, ctx->stream, &info));
        } else {
            const size_t slab_bytes = (size_t)(ctx->row_end - ctx->row_begin) * ctx->cfg.width * 4;
            const uint8_t* send = reinterpret_cast<const uint8_t*>(ctx->d_fb) + (size_t)ctx->row_begin * ctx->cfg.width * 4;
            const ncclResult_t r = g_nccl.AllGather(send, ctx->d_fb, slab_bytes, ncclUint8, ctx->comm, ctx->stream);
            if (r != ncclSuccess) return fail(ctx, VRT_E_NCCL, "ncclAllGather failed: %s", g_nccl.GetErrorString(r));
        }
    } else if (ctx->world > 1 && ctx->exchange_mode != VRT_EXCHANGE_HOST) {
        // peer-store: the pixels are already in every rank's framebuffer; the frame barrier makes "every rank has finished
        // frame k" visible in stream order.  The NEXT frame's remote stores land in the other ring slot of every peer, so the
        // barrier also waits until this rank has copied that slot's previous frame to its host (pipelined frames).
        const int other = (int)((ctx->ring_seq + 1u) & 1u);
        if (ring && ctx->slot_used[other]) VRT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copied[other], 0));
        if (ctx->exchange_mode == VRT_EXCHANGE_PEER_FLAGS || ctx->exchange_mode == VRT_EXCHANGE_PEER_PUSH || ctx->exchange_mode == VRT_EXCHANGE_PEER_TILES) {
            uint32_t* flags[8] = {nullptr};
            for (int r = 0; r < ctx->world; r++) flags[r] = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(ctx->peer_fb[r]) + 2 * ctx->fb_bytes);
            VRT_CUDA(ctx, launch_peer_barrier(flags, (uint32_t)ctx->rank, (uint32_t)ctx->world, ++ctx->barrier_frame, ctx->h_barrier_error, ctx->stream, &info));
        } else {
            const ncclResult_t r = g_nccl.AllReduce(ctx->d_barrier, ctx->d_barrier, 1, ncclInt32, ncclSum, ctx->comm, ctx->stream);
            if (r != ncclSuccess) return fail(ctx, VRT_E_NCCL, "ncclAllReduce failed: %s", g_nccl.GetErrorString(r));
        }
    }
    if (P.n_stage) {  // every rank's tiles of this frame are in this rank's staging buffer now: into the row-major framebuffer
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(ctx->d_fb_own) + stage_offset(ctx, (int)(ctx->stage_seq & 1u)));
        VRT_CUDA(ctx, launch_untile(mine, ctx->d_fb, ctx->cfg.width, ctx->cfg.height, P.vec_store_ok != 0u, ctx->stream, &info));
        ctx->stage_seq++;
    }
    if (P.tile_order) {
        // every `interval` frames (and after the first): sort the costs this frame reported into the next frames' order.  Dealt
        // schedules sort AFTER the frame barrier, when every rank's costs of this frame have arrived: all ranks sort the same array.
        if (ctx->sched_frames % ctx->sched_interval == 0)
            VRT_CUDA(ctx, launch_sched_sort(P.tile_cost, ctx->sched_tiles, ctx->d_order, ctx->d_sched_scratch, ctx->stream, &info));
        ctx->sched_frames++;
    }
    if (!pipelined) {
        VRT_CUDA(ctx, cudaEventRecord(ctx->ev_end, ctx->stream));
        ctx->timing_valid = true;
    }
    ctx->last_launches = info.launches;
    if (ring) ctx->ring_seq++;
    return VRT_OK;
}

// device -> host copy of what this context holds of the frame in d_fb, into the same place of a full-image host buffer: the whole
// image after a device-side exchange; only this rank's own rows / strips when there is none (one row slab, or — VRT_EXCHANGE_HOST with
// an interleaved partition — one strided copy of the 4-row strips k * world + rank).
int copy_frame_to_host(vrt_ctx* ctx, uint8_t* host, cudaStream_t stream) {
    const uint8_t* fb = reinterpret_cast<const uint8_t*>(ctx->d_fb);
    const size_t row_bytes = (size_t)ctx->cfg.width * 4;
    const bool host_mode = ctx->exchange_mode == VRT_EXCHANGE_HOST;
    if (ctx->interleave && host_mode && !scattered(ctx)) {
        const size_t chunk = kStripRows * row_bytes, pitch = (size_t)ctx->part_world * chunk, first = (size_t)ctx->part_rank * chunk;
        const uint32_t strips = (ctx->cfg.height + kStripRows - 1) / kStripRows;
        const uint32_t mine = strips > ctx->part_rank ? (strips - ctx->part_rank + ctx->part_world - 1) / ctx->part_world : 0u;
        if (mine == 0) return VRT_OK;
        const uint32_t last_strip = (mine - 1) * ctx->part_world + ctx->part_rank;
        const uint32_t last_rows = (last_strip + 1) * kStripRows <= ctx->cfg.height ? kStripRows : ctx->cfg.height - last_strip * kStripRows;
        const uint32_t full = last_rows == kStripRows ? mine : mine - 1;
        if (full) VRT_CUDA(ctx, cudaMemcpy2DAsync(host + first, pitch, fb + first, pitch, chunk, full, cudaMemcpyDeviceToHost, stream));
        if (full != mine) {
            const size_t at = (size_t)last_strip * chunk;
            VRT_CUDA(ctx, cudaMemcpyAsync(host + at, fb + at, last_rows * row_bytes, cudaMemcpyDeviceToHost, stream));
        }
        return VRT_OK;
    }
    const bool whole = !host_mode && (ctx->world > 1 || ctx->interleave || scattered(ctx));
    const size_t from = whole ? 0 : (size_t)ctx->row_begin * row_bytes;
    const size_t n = whole ? ctx->fb_bytes : (size_t)(ctx->row_end - ctx->row_begin) * row_bytes;
    if (ctx->interleave && !whole && scattered(ctx))
        return fail(ctx, VRT_E_STATE, "VRT_EXCHANGE_HOST cannot be combined with VRT_SCHED_DEAL / SHARED (a rank's tiles are scattered over the image)");
    VRT_CUDA(ctx, cudaMemcpyAsync(host + from, fb + from, n, cudaMemcpyDeviceToHost, stream));
    return VRT_OK;
}

}  // namespace

int vrt_trace(vrt_ctx* ctx, const vrt_camera* camera, const vrt_sun* sun) {
    if (!ctx) return VRT_E_INVALID;
    return trace_frame(ctx, camera, sun, false);
}

int vrt_sync(vrt_ctx* ctx) {
    if (!ctx) return VRT_E_INVALID;
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    VRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->copy_stream) VRT_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    if (ctx->h_barrier_error && *reinterpret_cast<volatile int*>(ctx->h_barrier_error) != 0)
        return fail(ctx, VRT_E_STATE, "vrt_sync: the peer-flag frame barrier timed out (a rank did not trace the same frame)");
    return VRT_OK;
}

int vrt_trace_to_host_async(vrt_ctx* ctx, const vrt_camera* camera, const vrt_sun* sun, uint8_t* rgba8_host, size_t bytes) {
    if (!ctx) return VRT_E_INVALID;
    if (rgba8_host && bytes != ctx->fb_bytes) return fail(ctx, VRT_E_INVALID, "vrt_trace_to_host_async: need a %zu-byte buffer", ctx->fb_bytes);
    if (!owns_fb(ctx)) return fail(ctx, VRT_E_STATE, "vrt_trace_to_host_async: not available while a caller-owned framebuffer is attached");
    const int slot = (int)(ctx->ring_seq & 1u);
    const int rc = trace_frame(ctx, camera, sun, true);
    if (rc != VRT_OK) return rc;
    VRT_CUDA(ctx, cudaEventRecord(ctx->ev_traced[slot], ctx->stream));
    VRT_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_traced[slot], 0));
    if (rgba8_host) {  // NULL: take part in the frame ring (multi-GPU ranks that do not need the pixels on their host) without a copy
        const int rcc = copy_frame_to_host(ctx, rgba8_host, ctx->copy_stream);
        if (rcc != VRT_OK) return rcc;
    }
    VRT_CUDA(ctx, cudaEventRecord(ctx->ev_copied[slot], ctx->copy_stream));
    ctx->slot_used[slot] = true;
    return VRT_OK;
}

int vrt_read_framebuffer(vrt_ctx* ctx, uint8_t* rgba8_host, size_t bytes) {
    if (!ctx) return VRT_E_INVALID;
    if (!rgba8_host || bytes != ctx->fb_bytes) return fail(ctx, VRT_E_INVALID, "vrt_read_framebuffer: need a %zu-byte buffer", ctx->fb_bytes);
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    VRT_CUDA(ctx, cudaMemcpyAsync(rgba8_host, ctx->d_fb, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    VRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VRT_OK;
}

int vrt_trace_to_host(vrt_ctx* ctx, const vrt_camera* camera, const vrt_sun* sun, uint8_t* rgba8_host, size_t bytes) {
    if (!ctx) return VRT_E_INVALID;
    if (!rgba8_host || bytes != ctx->fb_bytes) return fail(ctx, VRT_E_INVALID, "vrt_trace_to_host: need a %zu-byte buffer", ctx->fb_bytes);
    const int rc = trace_frame(ctx, camera, sun, false);
    if (rc != VRT_OK) return rc;
    const int rcc = copy_frame_to_host(ctx, rgba8_host, ctx->stream);
    if (rcc != VRT_OK) return rcc;
    VRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VRT_OK;
}

int vrt_insert_voxels(vrt_ctx* ctx, const uint32_t* xyzm_host, size_t count, uint32_t* active_bricks) {
    if (!ctx) return VRT_E_INVALID;
    if (!active_bricks) return fail(ctx, VRT_E_INVALID, "vrt_insert_voxels: active_bricks is NULL");
    if (count == 0) return VRT_OK;
    if (!xyzm_host) return fail(ctx, VRT_E_INVALID, "vrt_insert_voxels: xyzm is NULL");
    if (!ctx->have_grid) return fail(ctx, VRT_E_STATE, "vrt_insert_voxels: vrt_upload_grid_state has not been called");
    if (count >= 0xffffffffull) return fail(ctx, VRT_E_RANGE, "vrt_insert_voxels: at most 2^32 - 2 voxels per call");
    if (*active_bricks > ctx->n_start_indices) return fail(ctx, VRT_E_RANGE, "vrt_insert_voxels: active_bricks %u exceeds n_brick_alloc %zu", *active_bricks, ctx->n_start_indices);
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    const uint32_t bits = ctx->cfg.brick_dim * ctx->cfg.brick_dim * ctx->cfg.brick_dim;
    InsertBuffers B;
    B.state = ctx->grid, B.brick_dim = ctx->cfg.brick_dim;
    B.n_cells = (size_t)ctx->grid.dim_x * ctx->grid.dim_y * ctx->grid.dim_z;
    if (B.n_cells > ctx->n_brick_indices) return fail(ctx, VRT_E_RANGE, "vrt_insert_voxels: grid state has more cells than n_bricks");
    B.statuses = ctx->d_statuses, B.brick_indices = ctx->d_brick_indices, B.occupancy = ctx->d_occupancy;
    B.start_indices = ctx->d_start_indices, B.material_indices = ctx->d_material_indices;
    // scratch: one allocation for the voxel list + flags + ranks + scan levels + totals + per-cell first positions
    const size_t scan_entries = insert_scan_scratch_entries(count);
    const size_t bytes = count * 16 + count * 8 * 2 + scan_entries * 8 + 4 * 8 + B.n_cells * 4;
    uint8_t* scratch = nullptr;
    if (cudaMalloc(&scratch, bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, VRT_E_OOM, "vrt_insert_voxels: cannot allocate %zu bytes of scratch", bytes);
    }
    uint32_t* d_xyzm = reinterpret_cast<uint32_t*>(scratch);
    B.flags = reinterpret_cast<unsigned long long*>(scratch + count * 16);
    B.ranks = B.flags + count;
    B.scan_scratch = B.ranks + count;
    B.totals = B.scan_scratch + scan_entries;
    B.first_pos = reinterpret_cast<uint32_t*>(B.totals + 4);
    uint32_t* last_writer = nullptr;
    int rc = VRT_OK;
    unsigned long long totals[4] = {0, 0, 0, 0};
    LaunchInfo info = {0u};
    auto cuda_ok = [&](cudaError_t e) {
        if (e == cudaSuccess) return true;
        rc = fail(ctx, VRT_E_CUDA, "vrt_insert_voxels: %s", cudaGetErrorString(e));
        return false;
    };
    do {
        if (!cuda_ok(cudaMemcpyAsync(d_xyzm, xyzm_host, count * 16, cudaMemcpyHostToDevice, ctx->stream))) break;
        if (!cuda_ok(launch_insert_prepare(B, d_xyzm, count, ctx->stream, &info))) break;
        if (!cuda_ok(cudaMemcpyAsync(totals, B.totals, sizeof(totals), cudaMemcpyDeviceToHost, ctx->stream))) break;
        if (!cuda_ok(cudaStreamSynchronize(ctx->stream))) break;
        if (totals[2] != 0ull) {
            rc = fail(ctx, VRT_E_RANGE, "vrt_insert_voxels: a voxel lies outside the %ux%ux%u grid (nothing inserted)", ctx->grid.voxel_dim_x, ctx->grid.voxel_dim_y, ctx->grid.voxel_dim_z);
            break;
        }
        const uint32_t new_bricks = (uint32_t)(totals[0] >> 32), touched = (uint32_t)totals[0];
        if ((size_t)*active_bricks + new_bricks > ctx->n_start_indices) {
            rc = fail(ctx, VRT_E_RANGE, "vrt_insert_voxels: %u new bricks on top of %u exceed n_brick_alloc %zu (nothing inserted)", new_bricks, *active_bricks, ctx->n_start_indices);
            break;
        }
        const size_t lw_bytes = (size_t)touched * bits * 4;
        if (cudaMalloc(&last_writer, lw_bytes) != cudaSuccess) {
            cudaGetLastError();
            rc = fail(ctx, VRT_E_OOM, "vrt_insert_voxels: cannot allocate %zu bytes of scratch", lw_bytes);
            break;
        }
        if (!cuda_ok(cudaMemsetAsync(last_writer, 0, lw_bytes, ctx->stream))) break;
        if (!cuda_ok(launch_insert_commit(B, d_xyzm, count, *active_bricks, last_writer, ctx->stream, &info))) break;
        if (!cuda_ok(cudaStreamSynchronize(ctx->stream))) break;
        *active_bricks += new_bricks;
        ctx->accel_dirty = ctx->accel_force = true;  // the status words changed on the device, behind the merge kernel's back
    } while (false);
    cudaFree(last_writer);
    cudaFree(scratch);
    return rc;
}

int vrt_download_buffer(vrt_ctx* ctx, uint32_t which, size_t offset, void* host, size_t count) {
    if (!ctx) return VRT_E_INVALID;
    const uint8_t* src = nullptr;
    size_t capacity = 0, elem = 4;
    switch (which) {
        case VRT_BUFFER_STATUSES: src = reinterpret_cast<const uint8_t*>(ctx->d_statuses), capacity = ctx->n_statuses; break;
        case VRT_BUFFER_BRICK_INDICES: src = reinterpret_cast<const uint8_t*>(ctx->d_brick_indices), capacity = ctx->n_brick_indices; break;
        case VRT_BUFFER_OCCUPANCY: src = ctx->d_occupancy, capacity = ctx->n_occupancy, elem = 1; break;
        case VRT_BUFFER_START_INDICES: src = reinterpret_cast<const uint8_t*>(ctx->d_start_indices), capacity = ctx->n_start_indices; break;
        case VRT_BUFFER_MATERIAL_INDICES: src = ctx->d_material_indices, capacity = ctx->n_material_indices, elem = 1; break;
        case VRT_BUFFER_DEBUG_DIST: {  // the derived distance planes, brought up to date first (what the next trace would use)
            if (!ctx->have_grid || !ctx->d_dist) return fail(ctx, VRT_E_STATE, "vrt_download_buffer: no grid state yet");
            VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
            if (ctx->accel_dirty || ctx->occ_dirty || ctx->accel_force) {
                vrt_camera cam;
                vrt_sun sun;
                std::memset(&cam, 0, sizeof(cam));
                std::memset(&sun, 0, sizeof(sun));
                TraceParams P;
                fill_params(ctx, &cam, &sun, P);
                const int rcb = rebuild_accel(ctx, P, nullptr);
                if (rcb != VRT_OK) return rcb;
            }
            src = ctx->d_dist, capacity = 8 * ctx->dist_plane, elem = 1;
            break;
        }
        default: return fail(ctx, VRT_E_INVALID, "vrt_download_buffer: unknown buffer %u", which);
    }
    if (count == 0) return VRT_OK;
    if (!host) return fail(ctx, VRT_E_INVALID, "vrt_download_buffer: host is NULL");
    if (offset > capacity || count > capacity - offset) return fail(ctx, VRT_E_RANGE, "vrt_download_buffer: [%zu, %zu) outside capacity %zu", offset, offset + count, capacity);
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    VRT_CUDA(ctx, cudaMemcpyAsync(host, src + offset * elem, count * elem, cudaMemcpyDeviceToHost, ctx->stream));
    VRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VRT_OK;
}

int vrt_debug_tile_stats(vrt_ctx* ctx, uint32_t* host, size_t tiles) {
    if (!ctx || !host || tiles > ctx->tiles_global) return VRT_E_INVALID;
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    static uint32_t* d_stats = nullptr;  // analysis only: one buffer per process
    static size_t d_tiles = 0;
    if (d_tiles < ctx->tiles_global) {
        cudaFree(d_stats);
        VRT_CUDA(ctx, cudaMalloc(&d_stats, (size_t)ctx->tiles_global * 32));
        d_tiles = ctx->tiles_global;
        VRT_CUDA(ctx, cudaMemset(d_stats, 0, (size_t)ctx->tiles_global * 32));
        if (debug_set_tile_stats(d_stats) != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, VRT_E_STATE, "vrt_debug_tile_stats: this build has no tile statistics (-DVRT_TILE_STATS=1)");
        }
        std::memset(host, 0, tiles * 32);
        return VRT_OK;  // armed: the next traces fill it
    }
    VRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    VRT_CUDA(ctx, cudaMemcpy(host, d_stats, tiles * 32, cudaMemcpyDeviceToHost));
    return VRT_OK;
}

int vrt_debug_force_accel_rebuild(vrt_ctx* ctx) {
    if (!ctx) return VRT_E_INVALID;
    ctx->accel_dirty = ctx->accel_force = true;
    return VRT_OK;
}

int vrt_denoise(vrt_ctx* ctx, const vrt_denoise_params* params, uint32_t out_width, uint32_t out_height, uint32_t flags) {
    if (!ctx) return VRT_E_INVALID;
    if (!params) return fail(ctx, VRT_E_INVALID, "vrt_denoise: params is NULL");
    if (params->samples < 0 || params->samples > 255) return fail(ctx, VRT_E_INVALID, "vrt_denoise: samples %d outside [0, 255]", params->samples);
    if (out_width == 0 || out_height == 0) return fail(ctx, VRT_E_INVALID, "vrt_denoise: empty output image");
    if (flags & ~VRT_DENOISE_BGRA) return fail(ctx, VRT_E_INVALID, "vrt_denoise: unknown flags 0x%x", flags);
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    if (out_width != ctx->dn_width || out_height != ctx->dn_height) {
        if (ctx->d_denoised) {
            VRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            VRT_CUDA(ctx, cudaFree(ctx->d_denoised));
            ctx->d_denoised = nullptr, ctx->dn_width = ctx->dn_height = 0;
        }
        if (cudaMalloc(&ctx->d_denoised, (size_t)out_width * out_height * 4) != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, VRT_E_OOM, "vrt_denoise: cannot allocate a %ux%u output image", out_width, out_height);
        }
        ctx->dn_width = out_width, ctx->dn_height = out_height;
    }
    if (!ctx->d_dn_decoded && cudaMalloc(&ctx->d_dn_decoded, denoise_scratch_float4(ctx->cfg.width, ctx->cfg.height) * sizeof(float4)) != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, VRT_E_OOM, "vrt_denoise: cannot allocate the %ux%u float4 scratch image", ctx->cfg.width, ctx->cfg.height);
    }
    if (!ctx->ev_dn_begin) {
        VRT_CUDA(ctx, cudaEventCreate(&ctx->ev_dn_begin));
        VRT_CUDA(ctx, cudaEventCreate(&ctx->ev_dn_end));
    }
    LaunchInfo info = {0u};
    VRT_CUDA(ctx, cudaEventRecord(ctx->ev_dn_begin, ctx->stream));
    VRT_CUDA(ctx, launch_denoise(ctx->d_fb, ctx->d_dn_decoded, ctx->cfg.width, ctx->cfg.height, *params, ctx->d_denoised, out_width, out_height, (flags & VRT_DENOISE_BGRA) != 0u,
                                 ctx->stream, &info));
    VRT_CUDA(ctx, cudaEventRecord(ctx->ev_dn_end, ctx->stream));
    ctx->dn_timing_valid = true;
    return VRT_OK;
}

int vrt_read_denoised(vrt_ctx* ctx, uint8_t* host, size_t bytes) {
    if (!ctx) return VRT_E_INVALID;
    if (!ctx->d_denoised) return fail(ctx, VRT_E_STATE, "vrt_read_denoised: vrt_denoise has not been called");
    const size_t need = (size_t)ctx->dn_width * ctx->dn_height * 4;
    if (!host || bytes != need) return fail(ctx, VRT_E_INVALID, "vrt_read_denoised: need a %zu-byte buffer", need);
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    VRT_CUDA(ctx, cudaMemcpyAsync(host, ctx->d_denoised, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    VRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VRT_OK;
}

int vrt_denoised_device_ptr(vrt_ctx* ctx, void** out_device_ptr) {
    if (!ctx || !out_device_ptr) return VRT_E_INVALID;
    if (!ctx->d_denoised) return fail(ctx, VRT_E_STATE, "vrt_denoised_device_ptr: vrt_denoise has not been called");
    *out_device_ptr = ctx->d_denoised;
    return VRT_OK;
}

int vrt_last_denoise_ms(vrt_ctx* ctx, float* out_ms) {
    if (!ctx || !out_ms) return VRT_E_INVALID;
    if (!ctx->dn_timing_valid) return fail(ctx, VRT_E_STATE, "vrt_last_denoise_ms: no vrt_denoise yet");
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    VRT_CUDA(ctx, cudaEventSynchronize(ctx->ev_dn_end));
    VRT_CUDA(ctx, cudaEventElapsedTime(out_ms, ctx->ev_dn_begin, ctx->ev_dn_end));
    return VRT_OK;
}

int vrt_trace_rays(vrt_ctx* ctx, const vrt_ray* rays_device, vrt_ray_hit* hits_device, size_t count) {
    if (!ctx) return VRT_E_INVALID;
    if (count == 0) return VRT_OK;
    if (!rays_device || !hits_device) return fail(ctx, VRT_E_INVALID, "vrt_trace_rays: NULL buffer");
    if ((reinterpret_cast<uintptr_t>(rays_device) | reinterpret_cast<uintptr_t>(hits_device)) & 15u)
        return fail(ctx, VRT_E_INVALID, "vrt_trace_rays: buffers must be 16-byte aligned");
    if (!ctx->have_grid) return fail(ctx, VRT_E_STATE, "vrt_trace_rays: vrt_upload_grid_state has not been called");
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    vrt_camera cam;
    vrt_sun sun;
    std::memset(&cam, 0, sizeof(cam));
    std::memset(&sun, 0, sizeof(sun));
    TraceParams P;
    fill_params(ctx, &cam, &sun, P);
    // the ray list is this context's own: its private queue, no tile schedule, no peers
    P.tile_counter = ctx->d_tile_counter, P.queue_world = 0u;
    P.tile_order = nullptr, P.tile_cost = nullptr, P.n_cost_peers = 0u, P.n_peers = 0u, P.n_stage = 0u;
    LaunchInfo info = {0u};
    if (ctx->accel_dirty || ctx->occ_dirty) {
        const int rcb = rebuild_accel(ctx, P, &info);
        if (rcb != VRT_OK) return rcb;
    }
    VRT_CUDA(ctx, launch_trace_rays(P, rays_device, hits_device, count, ctx->stream, &info));
    ctx->last_launches = info.launches;
    return VRT_OK;
}

int vrt_trace_rays_host(vrt_ctx* ctx, const vrt_ray* rays_host, vrt_ray_hit* hits_host, size_t count) {
    if (!ctx) return VRT_E_INVALID;
    if (count == 0) return VRT_OK;
    if (!rays_host || !hits_host) return fail(ctx, VRT_E_INVALID, "vrt_trace_rays_host: NULL buffer");
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    void* d = nullptr;
    VRT_CUDA(ctx, cudaMalloc(&d, count * 64));
    vrt_ray* d_rays = static_cast<vrt_ray*>(d);
    vrt_ray_hit* d_hits = reinterpret_cast<vrt_ray_hit*>(static_cast<uint8_t*>(d) + count * 32);
    int rc = VRT_OK;
    cudaError_t e = cudaMemcpyAsync(d_rays, rays_host, count * 32, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) rc = vrt_trace_rays(ctx, d_rays, d_hits, count);
    if (e == cudaSuccess && rc == VRT_OK) e = cudaMemcpyAsync(hits_host, d_hits, count * 32, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && rc == VRT_OK) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (rc != VRT_OK) return rc;
    if (e != cudaSuccess) return fail(ctx, VRT_E_CUDA, "vrt_trace_rays_host: %s", cudaGetErrorString(e));
    return VRT_OK;
}

int vrt_read_aov(vrt_ctx* ctx, vrt_aov* aov_host, size_t count) {
    if (!ctx) return VRT_E_INVALID;
    if (!ctx->d_aov) return fail(ctx, VRT_E_STATE, "vrt_read_aov: context was not created with VRT_FLAG_AOV");
    if (!aov_host || count != (size_t)ctx->cfg.width * ctx->cfg.height) return fail(ctx, VRT_E_INVALID, "vrt_read_aov: need width*height records");
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    VRT_CUDA(ctx, cudaMemcpyAsync(aov_host, ctx->d_aov, count * sizeof(vrt_aov), cudaMemcpyDeviceToHost, ctx->stream));
    VRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VRT_OK;
}

int vrt_get_counters(vrt_ctx* ctx, vrt_counters* out) {
    if (!ctx) return VRT_E_INVALID;
    if (!ctx->d_counters) return fail(ctx, VRT_E_STATE, "vrt_get_counters: context was not created with VRT_FLAG_AOV");
    if (!out) return fail(ctx, VRT_E_INVALID, "vrt_get_counters: out is NULL");
    unsigned long long h[8];
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    VRT_CUDA(ctx, cudaMemcpyAsync(h, ctx->d_counters, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    VRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    out->rays = h[0], out->primary_hits = h[1], out->shadow_rays = h[2], out->grid_steps = h[3];
    out->voxel_steps = h[4], out->status_fetches = h[5], out->bricks_entered = h[6], out->hits = h[7];
    return VRT_OK;
}

int vrt_last_trace_ms(vrt_ctx* ctx, float* out_ms) {
    if (!ctx) return VRT_E_INVALID;
    if (!out_ms) return fail(ctx, VRT_E_INVALID, "vrt_last_trace_ms: out is NULL");
    if (!ctx->timing_valid) return fail(ctx, VRT_E_STATE, "vrt_last_trace_ms: no trace yet");
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    VRT_CUDA(ctx, cudaEventSynchronize(ctx->ev_end));
    VRT_CUDA(ctx, cudaEventElapsedTime(out_ms, ctx->ev_begin, ctx->ev_end));
    return VRT_OK;
}

int vrt_last_trace_launches(vrt_ctx* ctx, uint32_t* out) {
    if (!ctx) return VRT_E_INVALID;
    if (!out) return fail(ctx, VRT_E_INVALID, "vrt_last_trace_launches: out is NULL");
    *out = ctx->last_launches;
    return VRT_OK;
}

int vrt_set_stream(vrt_ctx* ctx, void* cuda_stream) {
    if (!ctx) return VRT_E_INVALID;
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    VRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return VRT_OK;
}

int vrt_attach_framebuffer(vrt_ctx* ctx, void* device_ptr, size_t bytes) {
    if (!ctx) return VRT_E_INVALID;
    if (device_ptr && bytes != ctx->fb_bytes) return fail(ctx, VRT_E_INVALID, "vrt_attach_framebuffer: need %zu bytes", ctx->fb_bytes);
    if (device_ptr && (reinterpret_cast<uintptr_t>(device_ptr) & 3u)) return fail(ctx, VRT_E_INVALID, "vrt_attach_framebuffer: pointer not 4-byte aligned");
    ctx->d_fb = device_ptr ? static_cast<uint32_t*>(device_ptr) : ctx->d_fb_own;
    return VRT_OK;
}

int vrt_framebuffer_device_ptr(vrt_ctx* ctx, void** out_device_ptr) {
    if (!ctx) return VRT_E_INVALID;
    if (!out_device_ptr) return fail(ctx, VRT_E_INVALID, "vrt_framebuffer_device_ptr: out is NULL");
    *out_device_ptr = ctx->d_fb;
    return VRT_OK;
}

int vrt_last_trace_kernel_ms(vrt_ctx* ctx, float* out_ms) {
    if (!ctx) return VRT_E_INVALID;
    if (!out_ms) return fail(ctx, VRT_E_INVALID, "vrt_last_trace_kernel_ms: out is NULL");
    if (!ctx->timing_valid) return fail(ctx, VRT_E_STATE, "vrt_last_trace_kernel_ms: no trace yet");
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    VRT_CUDA(ctx, cudaEventSynchronize(ctx->ev_kernel_end));
    VRT_CUDA(ctx, cudaEventElapsedTime(out_ms, ctx->ev_begin, ctx->ev_kernel_end));
    return VRT_OK;
}

int vrt_set_schedule(vrt_ctx* ctx, uint32_t mode, uint32_t interval) {
    if (!ctx) return VRT_E_INVALID;
    if (mode > VRT_SCHED_SHARED) return fail(ctx, VRT_E_INVALID, "vrt_set_schedule: unknown mode %u", mode);
    if (mode == VRT_SCHED_DEAL || mode == VRT_SCHED_SHARED) {
        if (!ctx->interleave) return fail(ctx, VRT_E_STATE, "vrt_set_schedule: VRT_SCHED_DEAL / SHARED need a context created with VRT_FLAG_INTERLEAVE (part_rank / part_world)");
        if (!ctx->d_fb_own || !owns_fb(ctx)) return fail(ctx, VRT_E_STATE, "vrt_set_schedule: VRT_SCHED_DEAL / SHARED need the context-owned framebuffer");
    }
    if (mode != ctx->sched_mode) ctx->sched_tiles = 0;  // another tile space / another set of tiles per rank: start from the default order
    ctx->sched_mode = mode;
    ctx->sched_interval = interval ? interval : 8u;
    return VRT_OK;
}

int vrt_sched_get_costs(vrt_ctx* ctx, uint16_t* costs_host, size_t count) {
    if (!ctx) return VRT_E_INVALID;
    if (!costs_host || count > ctx->tiles_global) return fail(ctx, VRT_E_INVALID, "vrt_sched_get_costs: at most %u tiles", ctx->tiles_global);
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    // the array the most recent frame wrote (frames alternate between two, so that a fast peer's next frame cannot disturb a sort)
    const int parity = ctx->sched_frames ? (int)((ctx->sched_frames - 1u) & 1u) : 0;
    VRT_CUDA(ctx, cudaMemcpyAsync(costs_host, ctx->d_cost[parity], count * sizeof(uint16_t), cudaMemcpyDeviceToHost, ctx->stream));
    VRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return VRT_OK;
}

int vrt_sched_set_costs(vrt_ctx* ctx, const uint16_t* costs_host, size_t count) {
    if (!ctx) return VRT_E_INVALID;
    if (!costs_host || count == 0 || count > ctx->tiles_global) return fail(ctx, VRT_E_INVALID, "vrt_sched_set_costs: 1 .. %u tiles", ctx->tiles_global);
    if (ctx->sched_mode == VRT_SCHED_STATIC) return fail(ctx, VRT_E_STATE, "vrt_sched_set_costs: no schedule selected (vrt_set_schedule)");
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    for (int i = 0; i < 2; i++) {
        const int rc = stage_upload(ctx, ctx->d_cost[i], costs_host, count * sizeof(uint16_t));
        if (rc != VRT_OK) return rc;
    }
    LaunchInfo info = {0u};
    VRT_CUDA(ctx, launch_sched_sort(ctx->d_cost[0], (uint32_t)count, ctx->d_order, ctx->d_sched_scratch, ctx->stream, &info));
    ctx->sched_tiles = (uint32_t)count;
    ctx->sched_frames = 1;  // the order is in place: no re-initialisation, next sort at the next multiple of the interval
    return VRT_OK;
}

int vrt_comm_get_unique_id(uint8_t id_out[VRT_NCCL_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) == VRT_NCCL_ID_BYTES, "ncclUniqueId size");
    if (!id_out) return fail(nullptr, VRT_E_INVALID, "vrt_comm_get_unique_id: out is NULL");
    if (!load_nccl(g_init_error, sizeof(g_init_error))) return VRT_E_NCCL;
    ncclUniqueId id;
    const ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, VRT_E_NCCL, "ncclGetUniqueId failed: %s", g_nccl.GetErrorString(r));
    std::memcpy(id_out, &id, sizeof(id));
    return VRT_OK;
}

int vrt_comm_init(vrt_ctx* ctx, int rank, int world, const uint8_t id[VRT_NCCL_ID_BYTES]) {
    if (!ctx) return VRT_E_INVALID;
    if (world < 1 || world > 8 || rank < 0 || rank >= world || !id) return fail(ctx, VRT_E_INVALID, "vrt_comm_init: bad rank/world %d/%d", rank, world);
    const uint32_t rows = ctx->row_end - ctx->row_begin;
    if (ctx->interleave) {
        if ((uint32_t)world != ctx->part_world || (uint32_t)rank != ctx->part_rank)
            return fail(ctx, VRT_E_INVALID, "vrt_comm_init: rank %d of %d does not match the interleaved partition %u of %u given to vrt_init", rank, world,
                        ctx->part_rank, ctx->part_world);
        if (world > 1 && ctx->cfg.width % 4 != 0) return fail(ctx, VRT_E_INVALID, "vrt_comm_init: the interleaved exchange needs width %% 4 == 0");
    } else if (world > 1 && (ctx->cfg.height % (uint32_t)world != 0 || rows != ctx->cfg.height / (uint32_t)world || ctx->row_begin != rows * (uint32_t)rank)) {
        return fail(ctx, VRT_E_INVALID, "vrt_comm_init: rank %d of %d must own rows [%u, %u)", rank, world, ctx->cfg.height / world * rank,
                    ctx->cfg.height / world * (rank + 1));
    }
    if (!load_nccl(ctx->err, sizeof(ctx->err))) return VRT_E_NCCL;
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    ncclUniqueId nid;
    std::memcpy(&nid, id, sizeof(nid));
    const ncclResult_t r = g_nccl.CommInitRank(&ctx->comm, world, nid, rank);
    if (r != ncclSuccess) return fail(ctx, VRT_E_NCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
    ctx->rank = rank, ctx->world = world;
    if (world > 1 && !ctx->d_barrier) {
        VRT_CUDA(ctx, cudaMalloc(&ctx->d_barrier, 4));
        VRT_CUDA(ctx, cudaMemsetAsync(ctx->d_barrier, 0, 4, ctx->stream));
    }
    if (world > 1 && ctx->interleave && !ctx->d_gather)
        VRT_CUDA(ctx, cudaMalloc(&ctx->d_gather, (size_t)world * ctx->strips_max * kStripRows * ctx->cfg.width * 4));
    if (world > 1) {
        // NCCL connects its channels lazily, inside the first collective of each kind: pay for that here (every rank calls
        // vrt_comm_init), not inside the first frames a caller times.
        ncclResult_t r = g_nccl.AllReduce(ctx->d_barrier, ctx->d_barrier, 1, ncclInt32, ncclSum, ctx->comm, ctx->stream);
        if (r == ncclSuccess) {
            if (ctx->interleave) {
                const size_t slab_bytes = (size_t)ctx->strips_max * kStripRows * ctx->cfg.width * 4;
                r = g_nccl.AllGather(reinterpret_cast<const uint8_t*>(ctx->d_gather) + (size_t)rank * slab_bytes, ctx->d_gather, slab_bytes, ncclUint8, ctx->comm, ctx->stream);
            } else {
                const size_t slab_bytes = (size_t)rows * ctx->cfg.width * 4;
                r = g_nccl.AllGather(reinterpret_cast<const uint8_t*>(ctx->d_fb_own) + (size_t)ctx->row_begin * ctx->cfg.width * 4, ctx->d_fb_own, slab_bytes, ncclUint8, ctx->comm, ctx->stream);
            }
        }
        if (r != ncclSuccess) return fail(ctx, VRT_E_NCCL, "vrt_comm_init: warm-up collective failed: %s", g_nccl.GetErrorString(r));
        VRT_CUDA(ctx, cudaMemsetAsync(ctx->d_barrier, 0, 4, ctx->stream));
        VRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return VRT_OK;
}

int vrt_comm_get_ipc_handle(vrt_ctx* ctx, uint8_t handle_out[VRT_IPC_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == VRT_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    if (!ctx) return VRT_E_INVALID;
    if (!handle_out) return fail(ctx, VRT_E_INVALID, "vrt_comm_get_ipc_handle: out is NULL");
    if (!owns_fb(ctx)) return fail(ctx, VRT_E_STATE, "vrt_comm_get_ipc_handle: only the context-owned framebuffer can be shared");
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    cudaIpcMemHandle_t h;
    VRT_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->d_fb_own));
    std::memcpy(handle_out, &h, sizeof(h));
    return VRT_OK;
}

int vrt_comm_open_peers(vrt_ctx* ctx, int rank, int world, const uint8_t* handles) {
    if (!ctx) return VRT_E_INVALID;
    if (world < 1 || world > 8 || rank < 0 || rank >= world || !handles) return fail(ctx, VRT_E_INVALID, "vrt_comm_open_peers: bad arguments");
    VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
    for (int r = 0; r < world; r++) {
        if (r == rank) {
            ctx->peer_fb[r] = ctx->d_fb_own;
            continue;
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)r * VRT_IPC_HANDLE_BYTES, sizeof(h));
        VRT_CUDA(ctx, cudaIpcOpenMemHandle(&ctx->peer_fb[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    ctx->rank = rank, ctx->world = world;
    ctx->peers_open = true;
    return VRT_OK;
}

int vrt_comm_set_exchange(vrt_ctx* ctx, uint32_t mode) {
    if (!ctx) return VRT_E_INVALID;
    if (mode > VRT_EXCHANGE_PEER_TILES) return fail(ctx, VRT_E_INVALID, "vrt_comm_set_exchange: unknown mode %u", mode);
    if ((mode == VRT_EXCHANGE_PEER_STORE || mode == VRT_EXCHANGE_PEER_FLAGS || mode == VRT_EXCHANGE_PEER_PUSH || mode == VRT_EXCHANGE_PEER_TILES) && !ctx->peers_open)
        return fail(ctx, VRT_E_STATE, "vrt_comm_set_exchange: call vrt_comm_open_peers first");
    if ((mode == VRT_EXCHANGE_PEER_FLAGS || mode == VRT_EXCHANGE_PEER_PUSH || mode == VRT_EXCHANGE_PEER_TILES) && !ctx->h_barrier_error) {
        VRT_CUDA(ctx, cudaSetDevice(ctx->cfg.device));
        VRT_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_barrier_error), sizeof(int), cudaHostAllocMapped));
        *ctx->h_barrier_error = 0;
    }
    ctx->exchange_mode = mode;
    if ((mode == VRT_EXCHANGE_PEER_FLAGS || mode == VRT_EXCHANGE_PEER_PUSH || mode == VRT_EXCHANGE_PEER_TILES) && ctx->world > 1) {
        // one round of the flag barrier now (every rank makes this call): the first touch of each peer mapping happens here, and
        // all ranks leave set-up together
        uint32_t* flags[8] = {nullptr};
        for (int r = 0; r < ctx->world; r++) flags[r] = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(ctx->peer_fb[r]) + 2 * ctx->fb_bytes);
        VRT_CUDA(ctx, launch_peer_barrier(flags, (uint32_t)ctx->rank, (uint32_t)ctx->world, ++ctx->barrier_frame, ctx->h_barrier_error, ctx->stream, nullptr));
        VRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (*reinterpret_cast<volatile int*>(ctx->h_barrier_error) != 0) return fail(ctx, VRT_E_STATE, "vrt_comm_set_exchange: a rank did not join the flag barrier");
    }
    return VRT_OK;
}

}  // extern "C"
