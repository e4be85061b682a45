// vrt_kernels.cuh — host-callable launchers of the device code (implemented in vrt_kernels.cu).
#pragma once

#include "vrt_device.cuh"

namespace vrt {

enum TraceKernel : int {
    KERNEL_REF = 0,    // transliteration, one thread per pixel (VRT_FLAG_BASELINE and the AOV path)
    KERNEL_TUNED = 1,  // warp-cooperative traversal over the derived distance grid (vrt_trav_warp.cuh)
};

constexpr uint32_t kDistBorder = 255u;  // dist byte of the one-cell border around the grid
constexpr uint32_t kDistFree = 0x80u;   // dist bit: no loaded brick anywhere in this cell's octant (low 7 bits: distance, <= 126)

struct LaunchInfo {
    uint32_t launches;  // kernels enqueued
};

// Enqueue the kernels that trace rows [P.row_begin, P.row_end) into P.fb.
cudaError_t launch_trace(const TraceParams& P, TraceKernel which, bool aov, cudaStream_t stream, LaunchInfo* info);

// What the status uploads since the last rebuild changed, kept on the device so that no upload has to wait for an answer:
// the cells whose status bit went 0 -> 1 (the only thing Grid.insert ever does to a status word, brick/Grid.zig:188), or "too many /
// a bit was cleared / never built" = rebuild everything.
constexpr uint32_t kAccelMaxNew = 32;
struct AccelDelta {
    uint32_t n_new;       // bits that went 0 -> 1 (may exceed kAccelMaxNew: then only the count is meaningful)
    uint32_t force_full;  // a bit went 1 -> 0, or the distance planes were never built for this grid
    uint32_t lo[3], hi[3];  // bounding box (cell coordinates) of the new cells
    uint32_t cell[kAccelMaxNew];
};
// Write `count` freshly uploaded status words (staged at `staged`) over dst[0..count) and record what changed in *delta.
cudaError_t launch_status_merge(uint32_t* dst, const uint32_t* staged, size_t count, size_t first_word, const vrt_grid_state& grid, AccelDelta* delta,
                                cudaStream_t stream, LaunchInfo* info);
// Rebuild the derived structures from the reference-format buffers: cell_rec always (brick_dim 4); the distance planes according
// to *delta — untouched if no status bit changed, patched in place for a few new bricks, rebuilt by the three line scans otherwise;
// *delta is cleared at the end.  occ_only: the distance planes are known to be current (host-side knowledge).  tmp: 6 * n_bricks bytes.
cudaError_t launch_build_accel(const TraceParams& P, uint4* cell_rec, uint8_t* dist, uint8_t* tmp, size_t n_bricks, bool occ_only,
                               AccelDelta* delta, cudaStream_t stream, LaunchInfo* info);
// Tiles (8x4 pixels) of the launch launch_trace_tuned would make for P with no schedule attached: the tile space an order for it permutes.
uint32_t trace_tile_space(const TraceParams& P);
// Tile schedule (vrt_sched.cu): order[i] = n - 1 - i; stable sort of the tiles by cost, most expensive first.
size_t sched_scratch_words(uint32_t n_tiles);
cudaError_t launch_sched_init(uint32_t* order, uint32_t n_tiles, cudaStream_t stream, LaunchInfo* info);
cudaError_t launch_sched_sort(const uint16_t* cost, uint32_t n_tiles, uint32_t* order, uint32_t* scratch, cudaStream_t stream, LaunchInfo* info);
// Explicit-ray mode: GridHit on caller-supplied rays (device pointers).
cudaError_t launch_trace_rays(const TraceParams& P, const vrt_ray* rays, vrt_ray_hit* hits, size_t count, cudaStream_t stream, LaunchInfo* info);
// After the all-gather of an interleaved partition: rank-major strips -> row-major frame (width % 4 == 0).
cudaError_t launch_deinterleave(const uint32_t* gathered, uint32_t* frame, uint32_t width, uint32_t height, uint32_t world, uint32_t strips_max,
                                cudaStream_t stream, LaunchInfo* info);
// image.frag:31-79 over `image` (width x height RGBA8, linear / repeat sampling) -> out (out_width x out_height RGBA8 or BGRA8); vrt_denoise.cu.
// decoded: denoise_scratch_float4(width, height) float4 of scratch (the UNORM-decoded image with its repeat border of kDenoisePad
// texels, then the per-launch sample tables).
constexpr int kDenoisePad = 16;
constexpr size_t kDenoiseScratchTail = 192 + 256;  // 3 x 256 floats, then 256 packed taps
inline size_t denoise_scratch_float4(uint32_t width, uint32_t height) { return ((size_t)width + 2 * kDenoisePad) * ((size_t)height + 2 * kDenoisePad) + kDenoiseScratchTail; }
cudaError_t launch_denoise(const uint32_t* image, float4* decoded, uint32_t width, uint32_t height, const vrt_denoise_params& params, uint32_t* out,
                           uint32_t out_width, uint32_t out_height, bool bgra, cudaStream_t stream, LaunchInfo* info);
// VRT_EXCHANGE_PEER_PUSH: copy the tiles P describes (as the trace launch would enumerate them) from P.fb into every P.peer_fb.
cudaError_t launch_push_tiles(const TraceParams& P, cudaStream_t stream, LaunchInfo* info);
// VRT_EXCHANGE_PEER_TILES: un-tile this rank's tile-major staging buffer (128 bytes per 8x4 tile of the whole image) into fb.
cudaError_t launch_untile(const uint32_t* stage, uint32_t* fb, uint32_t width, uint32_t height, bool vec_ok, cudaStream_t stream, LaunchInfo* info);
// Frame barrier over peer-mapped flag words (fused peer-store exchange, VRT_EXCHANGE_PEER_FLAGS); see vrt_kernels.cu.
cudaError_t launch_peer_barrier(uint32_t* const flags[8], uint32_t rank, uint32_t world, uint32_t frame, int* error, cudaStream_t stream, LaunchInfo* info);
// BrickGrid.insert for a batch of voxels on the device (vrt_build.cu): `prepare` only reads the grid buffers (first occurrences,
// flags, scan; totals[0] = new bricks << 32 | touched cells, totals[2] != 0: a voxel outside the grid), `commit` applies the batch.
struct InsertBuffers {
    vrt_grid_state state;
    uint32_t brick_dim;
    size_t n_cells;
    uint32_t* statuses;
    uint32_t* brick_indices;
    uint8_t* occupancy;
    uint32_t* start_indices;
    uint8_t* material_indices;
    uint32_t* first_pos;               // scratch, n_cells
    unsigned long long* flags;         // scratch, n voxels
    unsigned long long* ranks;         // scratch, n voxels
    unsigned long long* scan_scratch;  // scratch, insert_scan_scratch_entries(n)
    unsigned long long* totals;        // scratch, 4
};
size_t insert_scan_scratch_entries(size_t n);
cudaError_t launch_insert_prepare(const InsertBuffers& B, const uint32_t* xyzm, size_t n, cudaStream_t stream, LaunchInfo* info);
cudaError_t launch_insert_commit(const InsertBuffers& B, const uint32_t* xyzm, size_t n, uint32_t active_before, uint32_t* last_writer /* touched * brick_bits, zeroed */,
                                 cudaStream_t stream, LaunchInfo* info);
// analysis builds (-DVRT_TILE_STATS=1): device buffer of 8 words per tile the trace kernel fills; cudaErrorNotSupported otherwise
cudaError_t debug_set_tile_stats(uint32_t* device_buffer);
constexpr size_t kSharedQueueOffset = 128;  // inside the flag words' 256 bytes: the tile queue all ranks pull from under VRT_SCHED_SHARED (rank 0's copy)
constexpr size_t kPeerFlagBytes = 256;  // flag words appended to the IPC-shared framebuffer allocation
constexpr uint32_t kStripRows = 4;  // rows per strip of the interleaved partition (= the tile height of the trace kernel)
cudaError_t launch_trace_tuned(const TraceParams& P, bool aov, cudaStream_t stream, LaunchInfo* info);

}  // namespace vrt
