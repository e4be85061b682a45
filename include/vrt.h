/*
 * vrt.h — C ABI of the B200 voxel ray-tracing hot path (libvrt.so).
 *
 * This is the drop-in boundary for the ONE data-parallel path of Avokadoen/zig_vulkan that this
 * repo accelerates: the per-pixel two-level 3D-DDA traversal + hit shading that the reference runs as
 * the Vulkan compute shader assets/shaders/brick_raytracer.comp.  The entry points replace, one for one,
 * the host-side objects that feed and launch that shader (all file:line below are relative to the
 * reference tree):
 *
 *   vrt_init / vrt_deinit            <- ComputePipeline.init / deinit        (voxel_rt/ComputePipeline.zig:67,385)
 *                                       + the compute image + buffer sizing  (voxel_rt/Pipeline.zig:103-126,272-316)
 *   vrt_upload_grid_state            <- Pipeline.transferGridState           (voxel_rt/Pipeline.zig:560-571)
 *   vrt_upload_materials             <- Pipeline.transferMaterials           (voxel_rt/Pipeline.zig:573-583)
 *   vrt_upload_brick_statuses        <- Pipeline.transferBrickStatuses       (voxel_rt/Pipeline.zig:585-595)
 *   vrt_upload_brick_indices         <- Pipeline.transferBrickIndices        (voxel_rt/Pipeline.zig:597-607)
 *   vrt_upload_brick_occupancy       <- Pipeline.transferBrickOccupancy      (voxel_rt/Pipeline.zig:609-624)
 *   vrt_upload_brick_start_indices   <- Pipeline.transferBrickStartIndex     (voxel_rt/Pipeline.zig:626-641)
 *   vrt_upload_material_indices      <- Pipeline.transferMaterialIndices     (voxel_rt/Pipeline.zig:643-652)
 *   vrt_trace                        <- ComputePipeline.dispatch             (voxel_rt/ComputePipeline.zig:417-463)
 *   vrt_sync                         <- the complete_fence wait              (voxel_rt/ComputePipeline.zig:423-434)
 *   vrt_read_framebuffer             <- (new) the reference samples the compute image in image.frag and never
 *                                       reads it back; closest analogue Texture.copyToHost (render/Texture.zig:185-237)
 *
 * Plain pointers and sizes only; no C++/torch/CUDA types cross this boundary (streams and device pointers
 * travel as void*).  Every function returns 0 on success or a negative vrt_status; nothing aborts or throws.
 * A context is NOT thread-safe (the reference drives its pipeline from one thread, src/main.zig:156-195).
 * All uploads copy from the caller's buffer before returning, so the caller may mutate it immediately.
 */
#ifndef VRT_H
#define VRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VRT_ABI_VERSION 1u

typedef enum vrt_status {
    VRT_OK = 0,
    VRT_E_INVALID = -1, /* bad argument / bad config                                 */
    VRT_E_OOM = -2,     /* device or pinned-host allocation failed                    */
    VRT_E_RANGE = -3,   /* (offset, count) outside the buffer sized at vrt_init       */
    VRT_E_CUDA = -4,    /* CUDA runtime error; text in vrt_last_error                 */
    VRT_E_NCCL = -5,    /* NCCL error; text in vrt_last_error                         */
    VRT_E_STATE = -6    /* call order violated (e.g. trace before grid state upload)  */
} vrt_status;

/* ---------------------------------------------------------------------------------------------------
 * Byte-exact device structs.  These are what the reference uploads verbatim.
 * ------------------------------------------------------------------------------------------------- */

/* brick/State.zig:60-79 `Device` == GLSL `BrickGridState` UBO (brick_raytracer.comp:79-95).  64 bytes. */
typedef struct vrt_grid_state {
    uint32_t voxel_dim_x, voxel_dim_y, voxel_dim_z; /* voxels per axis                         */
    uint32_t dim_x, dim_y, dim_z;                   /* bricks per axis                         */
    uint32_t padding1, padding2;
    float min_point_base_t[4];                      /* xyz = grid min corner, w = base_t       */
    float max_point_scale[4];                       /* xyz = grid max corner, w = brick scale  */
} vrt_grid_state;

/* Camera.zig:183-193 `Device` == GLSL push constants bytes [0,96) (brick_raytracer.comp:58-68).
 * Zig @Vector(3,f32) is 16-byte aligned and 16 bytes wide, hence the explicit pads.  96 bytes. */
typedef struct vrt_camera {
    uint32_t image_width, image_height;
    uint32_t _pad0[2];
    float horizontal[3];        float _pad1;
    float vertical[3];          float _pad2;
    float lower_left_corner[3]; float _pad3;
    float origin[3];            float _pad4;   /* GLSL `paddin` */
    int32_t samples_per_pixel;
    int32_t max_bounce;                        /* device value = user max_bounce + 1 (Camera.zig:74) */
    uint32_t _pad5[2];
} vrt_camera;

/* Sun.zig:13-18 `Device` == GLSL push constants bytes [96,128) (brick_raytracer.comp:70-74).  32 bytes. */
typedef struct vrt_sun {
    float position[3];
    uint32_t enabled;
    float color[3];
    float radius;
} vrt_sun;

/* gpu_types.zig:16-32 `Material` == GLSL `Material` (brick_raytracer.comp:97-104).  20 bytes, stride 20. */
typedef struct vrt_material {
    uint32_t type;      /* 0 lambertian, 1 metal, 2 dielectric */
    float albedo_r, albedo_g, albedo_b;
    float type_data;    /* metal: fuzz; dielectric: index of refraction */
} vrt_material;

#define VRT_MAT_LAMBERTIAN 0u
#define VRT_MAT_METAL 1u
#define VRT_MAT_DIELECTRIC 2u
#define VRT_MAT_NONE 3u

/* Per-pixel debug record ("AOV") written by vrt_trace when VRT_FLAG_AOV is set: the traversal result of
 * sample 0 / bounce 0.  This is what the parity tests compare bit-for-bit with the oracle.  64 bytes. */
typedef struct vrt_aov {
    uint32_t flags;              /* VRT_AOV_* bits                                                    */
    uint32_t grid_index;         /* brick-grid cell of the primary hit (GridHit `grid_index`), ~0 miss */
    uint32_t voxel_index;        /* voxel inside that brick (BrickHit `voxel_index`), ~0 on miss       */
    uint32_t material;           /* material_indices[] value at the hit (hit.index), ~0 on miss        */
    float t;                     /* hit.t                                                             */
    float point[3];              /* hit.point                                                         */
    float normal[3];             /* hit.normal                                                        */
    uint32_t shadow_grid_index;  /* cell that blocked the sun ray, ~0 if lit / not cast               */
    uint32_t shadow_voxel_index;
    uint32_t grid_steps;         /* brick-level DDA cells visited, primary + shadow                   */
    uint32_t voxel_steps;        /* voxel-level DDA cells visited, primary + shadow                   */
    uint32_t status_fetches;     /* status-word loads under the reference's 1-word cache (:321-326)    */
} vrt_aov;

#define VRT_AOV_HIT 1u             /* primary ray hit a voxel                */
#define VRT_AOV_SHADOW_CAST 2u     /* a sun ray was traced from the hit      */
#define VRT_AOV_SHADOW_BLOCKED 4u  /* ... and it hit something               */

/* Frame totals of the request-byte model that defines roofline "algorithmic bytes" (DESIGN.md):
 * BYTES = 4*pixels + sum_rays[4*S + (4 + brick_bytes)*B + 25*H]. */
typedef struct vrt_counters {
    uint64_t rays;            /* primary + shadow (+ bounce) rays cast                 */
    uint64_t primary_hits;
    uint64_t shadow_rays;
    uint64_t grid_steps;
    uint64_t voxel_steps;
    uint64_t status_fetches;  /* S: sum over rays                                      */
    uint64_t bricks_entered;  /* B: occupied bricks entered (BrickHit calls)           */
    uint64_t hits;            /* H: rays that ended in a voxel hit (primary + shadow)  */
} vrt_counters;

/* ---------------------------------------------------------------------------------------------------
 * Context
 * ------------------------------------------------------------------------------------------------- */

typedef struct vrt_ctx vrt_ctx;

#define VRT_FLAG_AOV 1u        /* allocate + fill the AOV buffer and counters on every trace (debug)   */
#define VRT_FLAG_BASELINE 2u   /* use the one-thread-per-pixel transliteration kernel, not the tuned one */
#define VRT_FLAG_INTERLEAVE 4u /* partition by interleaved strips of 4 image rows (part_rank / part_world) instead of a row slab:
                                  strip t belongs to rank t % part_world.  Balances sky against terrain across GPUs. */

/* What ComputePipeline.init receives as ImageInfo + StateConfigs + specialization constants
 * (ComputePipeline.zig:67-73, Pipeline.zig:272-316), flattened. */
typedef struct vrt_config {
    uint32_t struct_size;        /* = sizeof(vrt_config); ABI guard                                      */
    uint32_t abi_version;        /* = VRT_ABI_VERSION                                                    */
    uint32_t width, height;      /* full image size (ImageInfo.width/height)                             */
    uint32_t brick_dim;          /* spec const id 4 `brick_dimensions`; reference: 4. 4, 8 or 16.        */
    uint32_t material_capacity;  /* Pipeline.Config.material_buffer; reference: 256                      */
    uint64_t n_bricks;           /* dim_x*dim_y*dim_z  -> statuses ceil(n/32) words, indices n words     */
    uint64_t n_brick_alloc;      /* BrickGrid.Config.brick_alloc -> occupancy/start/material capacities  */
    int32_t device;              /* CUDA device ordinal this context owns                                */
    uint32_t flags;              /* VRT_FLAG_*                                                           */
    uint32_t row_begin, row_end; /* image rows [begin,end) this context traces; 0,0 = all rows.
                                    Multi-GPU runs give each rank one slab of the same full image.       */
    uint32_t part_rank, part_world; /* VRT_FLAG_INTERLEAVE: this context traces the 4-row strips t with
                                    t % part_world == part_rank (row_begin/row_end must be 0).            */
} vrt_config;

int vrt_init(vrt_ctx** out_ctx, const vrt_config* cfg);
void vrt_deinit(vrt_ctx* ctx);

/* Never NULL; valid until the next call on the same ctx (or on NULL: the last vrt_init failure). */
const char* vrt_last_error(const vrt_ctx* ctx);

/* ---------------------------------------------------------------------------------------------------
 * Uploads: (offset in ELEMENTS, pointer, element count) exactly like Pipeline.transfer*; partial ranges
 * are how the reference ships per-frame edits (VoxelRT.updateGridDelta, VoxelRT.zig:107-172).
 * ------------------------------------------------------------------------------------------------- */
int vrt_upload_grid_state(vrt_ctx* ctx, const vrt_grid_state* state);
int vrt_upload_materials(vrt_ctx* ctx, size_t offset, const vrt_material* data, size_t count);
int vrt_upload_brick_statuses(vrt_ctx* ctx, size_t offset, const uint32_t* data, size_t count);
int vrt_upload_brick_indices(vrt_ctx* ctx, size_t offset, const uint32_t* data, size_t count);
int vrt_upload_brick_occupancy(vrt_ctx* ctx, size_t offset, const uint8_t* data, size_t count);
int vrt_upload_brick_start_indices(vrt_ctx* ctx, size_t offset, const uint32_t* data, size_t count);
int vrt_upload_material_indices(vrt_ctx* ctx, size_t offset, const uint8_t* data, size_t count);

/* ---------------------------------------------------------------------------------------------------
 * Trace
 * ------------------------------------------------------------------------------------------------- */

/* Enqueue one frame on the context's stream and return (dispatch returns a semaphore the same way).
 * Multi-GPU contexts also enqueue the framebuffer exchange configured with vrt_comm_*. */
int vrt_trace(vrt_ctx* ctx, const vrt_camera* camera, const vrt_sun* sun);

/* Block until everything enqueued so far has finished. */
int vrt_sync(vrt_ctx* ctx);

/* Copy the RGBA8 framebuffer (width*height*4 bytes, row 0 first) to host memory; implies vrt_sync. */
int vrt_read_framebuffer(vrt_ctx* ctx, uint8_t* rgba8_host, size_t bytes);

/* vrt_trace + readback of this context's rows in one call, the frame a host-side caller sees: camera/sun go
 * host->device, rows [row_begin,row_end) come back into `rgba8_host` (full-image layout, only those rows are
 * written).  Blocks until the pixels are in host memory. */
int vrt_trace_to_host(vrt_ctx* ctx, const vrt_camera* camera, const vrt_sun* sun, uint8_t* rgba8_host, size_t bytes);

/* The same frame for a STREAM of frames: enqueue the trace and the device->host copy of this context's rows and return
 * without waiting.  Two frames may be in flight (the context double-buffers its framebuffer and copies on a second
 * stream), so the copy of frame k overlaps the trace of frame k+1 — the CUDA counterpart of the reference keeping one
 * compute frame in flight while the previous one is presented (ComputePipeline.zig:423-434, Pipeline.zig:494-517).
 * `rgba8_host` must stay valid until vrt_sync() (or until two further async calls have been made) and should be pinned
 * host memory, otherwise the copy is staged and does not overlap.  `rgba8_host` may be NULL: the frame is traced (and
 * exchanged) in the same ring slot but not copied — in a multi-GPU run every rank must issue the same sequence of frame
 * calls, and only the ranks that need the pixels on their host pass a buffer.  Not available while a caller-owned
 * framebuffer is attached (VRT_E_STATE). */
int vrt_trace_to_host_async(vrt_ctx* ctx, const vrt_camera* camera, const vrt_sun* sun, uint8_t* rgba8_host, size_t bytes);

/* ---------------------------------------------------------------------------------------------------
 * Explicit rays (extension).  The reference only ever traces the rays its shader generates from the camera block
 * (brick_raytracer.comp:474-477); this entry point runs the same GridHit/BrickHit (:271-471) on caller-supplied rays —
 * picking, collision, light probes.  32 bytes in, 32 bytes out per ray, both moved with 128-bit accesses.
 * ------------------------------------------------------------------------------------------------- */
typedef struct vrt_ray {      /* 32 bytes, 16-byte aligned */
    float origin[3];
    float _pad0;
    float direction[3];       /* normalised by the kernel exactly like CreateRay (:180-184) */
    float _pad1;
} vrt_ray;

typedef struct vrt_ray_hit {  /* 32 bytes, 16-byte aligned */
    uint32_t hit;             /* 1 = a voxel was hit                              */
    uint32_t grid_index;      /* GridHit `grid_index` (:318), ~0 on miss          */
    uint32_t voxel_index;     /* BrickHit `voxel_index` (:412), ~0 on miss        */
    uint32_t material;        /* material_indices[] value at the hit, ~0 on miss  */
    float t;                  /* hit.t                                            */
    float normal[3];          /* hit.normal                                       */
} vrt_ray_hit;

/* rays / hits are DEVICE pointers (count elements each); enqueued on the ctx stream like vrt_trace. */
int vrt_trace_rays(vrt_ctx* ctx, const vrt_ray* rays_device, vrt_ray_hit* hits_device, size_t count);
/* Same with HOST pointers: copies in, traces, copies out, blocks. */
int vrt_trace_rays_host(vrt_ctx* ctx, const vrt_ray* rays_host, vrt_ray_hit* hits_host, size_t count);

/* Debug / parity (requires VRT_FLAG_AOV). */
int vrt_read_aov(vrt_ctx* ctx, vrt_aov* aov_host, size_t count);
int vrt_get_counters(vrt_ctx* ctx, vrt_counters* out);

/* Milliseconds the device spent in the last vrt_trace / vrt_trace_to_host (kernels + exchange), CUDA events on ctx's stream.
 * (Pipelined frames, vrt_trace_to_host_async, are not timed.) */
int vrt_last_trace_ms(vrt_ctx* ctx, float* out_ms);
/* The part of it before the exchange: (derived-structure rebuild +) the trace kernel alone.  exchange = total - this. */
int vrt_last_trace_kernel_ms(vrt_ctx* ctx, float* out_ms);
/* Number of kernels the last vrt_trace launched. */
int vrt_last_trace_launches(vrt_ctx* ctx, uint32_t* out);

/* ---------------------------------------------------------------------------------------------------
 * Tile schedule (extension; the reference leaves workgroup order to the Vulkan driver, ComputePipeline.zig:547-550).
 * The trace kernel's persistent warps pull 8x4-pixel tiles from a queue.  STATIC: bottom-up image order.  LPT: every tile
 * reports what it cost and every `interval` frames the costs are sorted into the next frames' order, most expensive first
 * (longest-processing-time-first list scheduling: the launch no longer ends on a 50 us tile that happened to start last).
 * DEAL (contexts created with VRT_FLAG_INTERLEAVE): the sorted list of ALL tiles of the image is dealt round-robin to the
 * part_world ranks instead of giving each rank fixed strips — every GPU gets the same mix of cheap and expensive tiles; needs
 * a peer-store exchange mode when part_world > 1 GPUs really share the frame (a rank's tiles are scattered over the image),
 * and every rank must call it with the same arguments.  The image is the same whatever the schedule.
 * ------------------------------------------------------------------------------------------------- */
#define VRT_SCHED_STATIC 0u
#define VRT_SCHED_LPT 1u
#define VRT_SCHED_DEAL 2u
#define VRT_SCHED_SHARED 3u /* like DEAL, but nothing is dealt in advance: the warps of ALL ranks draw their tiles from one queue (in rank 0's
                               memory, atomics over NVLink) in cost order — whichever GPU has a free warp takes the next most expensive tile, so
                               the ranks finish together whatever the frame looks like.  Same requirements as DEAL; every rank must launch the
                               same frame. */
int vrt_set_schedule(vrt_ctx* ctx, uint32_t mode, uint32_t interval /* frames between sorts; 0 = 8 */);
/* Debug / tests: read the per-tile costs of the last frame (clock ticks / 32, 0 = never traced), or install costs and sort them
 * into the order right away.  count <= tiles of the image = ceil(width/8) * ceil(height/4). */
int vrt_sched_get_costs(vrt_ctx* ctx, uint16_t* costs_host, size_t count);
int vrt_sched_set_costs(vrt_ctx* ctx, const uint16_t* costs_host, size_t count);

/* ---------------------------------------------------------------------------------------------------
 * Scene edits on the device: BrickGrid.insert (brick/Grid.zig:129-194) for a batch of voxels, the step in front of the
 * uploads.  The five grid buffers of the ctx end up byte-identical to what the host loop
 *     for (i < count) grid.insert(x[i], y[i], z[i], material[i])
 * followed by the transfer* uploads would have produced: new bricks are numbered in order of first occurrence, their
 * material block starts at brick * brick_dim^3, a voxel inserted twice keeps the last material.  Differences, both stricter:
 * a voxel outside the grid or a batch that needs more bricks than n_brick_alloc fails with VRT_E_RANGE and changes nothing
 * (Grid.insert asserts / indexes out of bounds).  The host BrickGrid is not updated; read the buffers back with
 * vrt_download_buffer if it must follow.
 * ------------------------------------------------------------------------------------------------- */
/* xyzm_host: count packed {x, y, z, material} uint32 quadruples (the layout of vrt_grid_insert_many).  *active_bricks: in = bricks
 * allocated so far (BrickGrid.State.active_bricks; 0 for a fresh ctx), out = after the batch.  Blocks until done. */
int vrt_insert_voxels(vrt_ctx* ctx, const uint32_t* xyzm_host, size_t count, uint32_t* active_bricks);

#define VRT_BUFFER_STATUSES 3u         /* binding numbers of brick_raytracer.comp:109-134; element = uint32 */
#define VRT_BUFFER_BRICK_INDICES 4u    /* uint32 */
#define VRT_BUFFER_OCCUPANCY 5u        /* uint8  */
#define VRT_BUFFER_START_INDICES 6u    /* uint32 */
#define VRT_BUFFER_MATERIAL_INDICES 7u /* uint8  */
#define VRT_BUFFER_DEBUG_DIST 100u     /* debug / tests: the derived distance planes (uint8, 8 octants x padded grid), after
                                          bringing them up to date with the uploads made so far */
/* Debug / tests: the next trace rebuilds the distance planes from scratch instead of patching them for the bricks added since. */
int vrt_debug_force_accel_rebuild(vrt_ctx* ctx);
/* Analysis builds only (-DVRT_TILE_STATS=1, tools/gpu_tilestats.py; VRT_E_STATE otherwise): the first call arms per-tile counters, later
 * calls read 8 words per tile: rounds, step-loop iterations, brick phases, voxel-loop iterations, lanes marching (summed over rounds),
 * lanes testing (summed over brick phases), voxel hits, clock ticks / 32. */
int vrt_debug_tile_stats(vrt_ctx* ctx, uint32_t* host, size_t tiles);
/* Copy `count` elements starting at element `offset` of a grid buffer to the host (the inverse of vrt_upload_*).  Blocks. */
int vrt_download_buffer(vrt_ctx* ctx, uint32_t which, size_t offset, void* host, size_t count);

/* ---------------------------------------------------------------------------------------------------
 * Post-process: the reference's present pass (assets/shaders/image.frag:31-79, "sirBirdDenoise"), the step right after
 * the compute dispatch in Pipeline.draw (Pipeline.zig:441-540).  Replaces GraphicsPipeline's full-screen draw: reads the
 * traced RGBA8 image through a linear / repeat sampler (Pipeline.zig:193-212) and writes one denoised texel per pixel of
 * an out_width x out_height target (the swapchain extent; it may differ from the trace resolution).
 * ------------------------------------------------------------------------------------------------- */
typedef struct vrt_denoise_params { /* GraphicsPipeline.PushConstant (GraphicsPipeline.zig:27-32), 16 bytes */
    int32_t samples;                /* default 20  (Config, GraphicsPipeline.zig:34-39); 0 <= samples <= 255 */
    float distribution_bias;        /* default 0.6 */
    float pixel_multiplier;         /* default 1.5 */
    float inverse_hue_tolerance;    /* default 20  */
} vrt_denoise_params;

#define VRT_DENOISE_BGRA 1u /* store B,G,R,A bytes (the reference's B8G8R8A8_UNORM swapchain, swapchain.zig:235) instead of R,G,B,A */

/* Enqueue the pass on the ctx stream, after whatever vrt_trace put there: input = the framebuffer currently traced into,
 * output = a ctx-owned out_width*out_height*4-byte device image (re-allocated when the size changes). */
int vrt_denoise(vrt_ctx* ctx, const vrt_denoise_params* params, uint32_t out_width, uint32_t out_height, uint32_t flags);
/* Blocking read-back of the last vrt_denoise output (bytes = out_width*out_height*4). */
int vrt_read_denoised(vrt_ctx* ctx, uint8_t* host, size_t bytes);
/* Device address of the last vrt_denoise output. */
int vrt_denoised_device_ptr(vrt_ctx* ctx, void** out_device_ptr);
/* Milliseconds the device spent in the last vrt_denoise. */
int vrt_last_denoise_ms(vrt_ctx* ctx, float* out_ms);

/* ---------------------------------------------------------------------------------------------------
 * Interop (all optional)
 * ------------------------------------------------------------------------------------------------- */

/* Run on a caller-owned CUDA stream (cudaStream_t passed as void*); NULL restores the ctx's own stream. */
int vrt_set_stream(vrt_ctx* ctx, void* cuda_stream);
/* Trace into a caller-owned device image (width*height*4 bytes) — the reference's pipeline likewise does
 * not own its target image (ComputePipeline.zig:65-66).  NULL restores the ctx's own framebuffer. */
int vrt_attach_framebuffer(vrt_ctx* ctx, void* device_ptr, size_t bytes);
/* Device address of the framebuffer currently traced into. */
int vrt_framebuffer_device_ptr(vrt_ctx* ctx, void** out_device_ptr);

/* ---------------------------------------------------------------------------------------------------
 * Multi-GPU: one process (and one ctx) per GPU, each tracing a row slab of the same image; the slabs are
 * exchanged right after the trace kernel on the same stream so that every rank ends with the full frame.
 *   exchange mode 0: one in-place ncclAllGather over NVLink.  Row slabs must be equal; interleaved strips are traced
 *                    into a rank-major gather buffer and re-ordered into the frame by a copy kernel after the gather.
 *   exchange mode 1: fused — the trace kernel itself stores its pixels into every peer's framebuffer
 *                    through NVLink peer mappings (vrt_comm_open_peers); the only collective left is a 4-byte
 *                    all-reduce that orders frame k+1's stores after every rank has consumed frame k.
 *   exchange mode 2: mode 1 without any collective call: the frame barrier is a 32-thread kernel that publishes / polls
 *                    per-rank flag words through the same NVLink peer mappings (release / acquire at system scope).
 * ------------------------------------------------------------------------------------------------- */
#define VRT_NCCL_ID_BYTES 128
#define VRT_IPC_HANDLE_BYTES 64
#define VRT_EXCHANGE_ALLGATHER 0u
#define VRT_EXCHANGE_PEER_STORE 1u
#define VRT_EXCHANGE_PEER_FLAGS 2u
#define VRT_EXCHANGE_PEER_PUSH 4u /* mode 2 with the peer stores moved out of the trace kernel: the kernel writes its own framebuffer only and a
                                     copy kernel then ships this rank's tiles to every peer (128-bit loads / stores), followed by the flag barrier */
#define VRT_EXCHANGE_PEER_TILES 5u /* mode 2 with a tile-major wire format: the trace kernel stores each finished 8x4 tile as ONE contiguous 128-byte
                                      record into a staging buffer of every rank (its own included) instead of four 32-byte row segments into
                                      every peer's image; after the flag barrier each rank un-tiles its staging buffer into its framebuffer.
                                      A quarter of the NVLink packets, each four times the size.  Tuned kernel, strip-aligned partitions. */
#define VRT_EXCHANGE_HOST 3u /* no device-side exchange at all: the consumer is the host.  vrt_trace_to_host(_async) copies only this
                                rank's own rows / 4-row strips into their place of the full-image host buffer, so N ranks given the same
                                (shared, pinned) buffer assemble the frame over N PCIe links at once instead of funnelling it through one
                                GPU.  Needs no communicator and no peer mappings; the device framebuffers stay partial; the ranks'
                                host code decides when a frame is complete (all ranks' vrt_sync / a host barrier).  Not with VRT_SCHED_DEAL. */

int vrt_comm_get_unique_id(uint8_t id_out[VRT_NCCL_ID_BYTES]);
int vrt_comm_init(vrt_ctx* ctx, int rank, int world, const uint8_t id[VRT_NCCL_ID_BYTES]);
int vrt_comm_get_ipc_handle(vrt_ctx* ctx, uint8_t handle_out[VRT_IPC_HANDLE_BYTES]);
int vrt_comm_open_peers(vrt_ctx* ctx, int rank, int world, const uint8_t* handles /* world*VRT_IPC_HANDLE_BYTES */);
int vrt_comm_set_exchange(vrt_ctx* ctx, uint32_t mode);

#ifdef __cplusplus
}
#endif
#endif /* VRT_H */
