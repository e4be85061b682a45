// vrt_kernels.cu — the sm_100a kernels of the voxel ray-tracing hot path and their launchers.
//
// Compile with:  nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false -lineinfo  (see Makefile).
// --fmad=false is part of the contract: parity with the oracle is bit-exact only if no mul+add pair is
// contracted beyond the explicit fmaf() calls (DESIGN.md "FP discipline").
#include "vrt_kernels.cuh"
#include "vrt_shade.cuh"
#include "vrt_shade_warp.cuh"
#include "vrt_trav_ref.cuh"

namespace vrt {

// ----------------------------------------------------------------------------------------------------
// Baseline: the reference dispatch shape — one thread per pixel, a warp covers 32 consecutive pixels of a
// row exactly like a 32x32 GLSL workgroup's subgroups do (ComputePipeline.zig:547-550,588-597), and the
// shader's own memory accesses (vrt_trav_ref.cuh).
// ----------------------------------------------------------------------------------------------------
template <bool AOV>
__global__ void __launch_bounds__(256) trace_ref_kernel(const __grid_constant__ TraceParams P) {
    const uint32_t px = blockIdx.x * 32u + threadIdx.x;
    const uint32_t py = P.row_begin + blockIdx.y * 8u + threadIdx.y;
    PixelCounters pc = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    if (px < P.cam.image_width && py < P.row_end) {  // :156-159
        const uint32_t texel = shade_pixel<RefTraversal, AOV>(P, px, py, pc);
        P.fb[(size_t)py * P.cam.image_width + px] = texel;
    }
    if (AOV) flush_counters(P, pc);
}

// ----------------------------------------------------------------------------------------------------
// Tuned: persistent CTAs; each warp pulls 8x4-pixel tiles from a global counter (sky tiles that die after a slab
// test do not idle an SM while terrain tiles march hundreds of cells) and traces the 32 rays cooperatively
// (vrt_trav_warp.cuh).  The tile is written as eight 128-bit stores (4 texels each, gathered by shfl); in the
// fused multi-GPU exchange the same stores also go to every peer's framebuffer over NVLink.
// ----------------------------------------------------------------------------------------------------
// 8 CTAs of 128 threads per SM = 32 warps at 64 registers (40 bytes of spills outside the march): measured 4.5 % faster than
// 24 warps at 80 registers and no spills, 1 % faster than 36 warps at 56 (profiles/r02_ab_occupancy_prefetch_C3.txt).  The warps of a
// CTA share nothing, so the CTA size only sets the granularity of the register file split.
#ifndef VRT_TUNED_THREADS
#define VRT_TUNED_THREADS 128
#endif
#ifndef VRT_TUNED_BLOCKS
#define VRT_TUNED_BLOCKS 8
#endif
#ifndef VRT_GENERAL_BLOCKS
#define VRT_GENERAL_BLOCKS 8  // the general shading path spills 470 bytes at 64 registers and is still 4 % faster than 24 warps at 80 (REF workload)
#endif
constexpr int kTunedThreads = VRT_TUNED_THREADS;
constexpr uint32_t kTileW = 8, kTileH = 4;

// 3 CTAs of 256 threads per SM (80 registers): measured 10 % faster than 2 (102 registers, no spills) and equal to 4 (64, spills).
// SIMPLE: the configuration shade_pixel_warp_simple covers (chosen per launch by launch_trace_tuned).
// A warp leaves the work queue: the last one out resets the ticket counter for the next launch (see TraceParams::tile_counter).
// With a queue shared by several GPUs (VRT_SCHED_SHARED) every rank launches the same grid, so the warps of all of them add up.
VRT_DI void leave_queue(const TraceParams& P, uint32_t lane) {
    if (lane == 0) {
        const unsigned long long warps = (unsigned long long)gridDim.x * (blockDim.x >> 5) * (P.queue_world ? P.queue_world : 1u);
        if (P.queue_world > 1u) {
            __threadfence_system();
            if (atomicAdd_system(P.tile_counter + 1, 1ull) == warps - 1ull) {
                *reinterpret_cast<volatile unsigned long long*>(P.tile_counter) = 0ull;
                *reinterpret_cast<volatile unsigned long long*>(P.tile_counter + 1) = 0ull;
                __threadfence_system();
            }
        } else {
            __threadfence();
            if (atomicAdd(P.tile_counter + 1, 1ull) == warps - 1ull) {
                P.tile_counter[0] = 0ull;
                P.tile_counter[1] = 0ull;
            }
        }
    }
}
VRT_DI unsigned long long draw_ticket(const TraceParams& P) {
    return P.queue_world > 1u ? atomicAdd_system(P.tile_counter, 1ull) : atomicAdd(P.tile_counter, 1ull);
}

template <int BD, bool AOV, bool SIMPLE>
__global__ void __launch_bounds__(kTunedThreads, SIMPLE ? VRT_TUNED_BLOCKS : VRT_GENERAL_BLOCKS) trace_warp_kernel(const __grid_constant__ TraceParams P, const uint32_t tiles_x, const uint32_t tiles_total) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lx = lane & (kTileW - 1u), ly = lane >> 3;
    const uint32_t width = P.cam.image_width;
    PixelCounters pc = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#if VRT_TMA_MASKS
    if (BD != 4 && P.brick_dim == 16) brick_stage_init();
#endif

    for (;;) {
        unsigned long long t = 0ull;
        if (lane == 0) t = draw_ticket(P);
        t = __shfl_sync(kFullMask, t, 0);
        if (t >= (unsigned long long)tiles_total) break;
        // Scheduled: the most expensive tiles of the previous frames first (vrt_sched.cu).  Otherwise bottom-up: in the reference's
        // convention image row 0 is up (sky); starting with the ground rows leaves the cheap sky tiles to fill the tail of the launch.
        const uint32_t tile = P.tile_order ? __ldg(P.tile_order + P.order_offset + (uint32_t)t * P.order_stride) : tiles_total - 1u - (uint32_t)t;
        const long long tick0 = P.tile_cost ? clock64() : 0ll;
#if VRT_TILE_STATS
        if (lane < 8u) warp_stats()[lane] = 0u;
        __syncwarp();
#endif
        const uint32_t px = (tile % tiles_x) * kTileW + lx;
        const uint32_t strip = tile / tiles_x;  // this launch's k-th strip of kTileH rows
        const uint32_t py = P.il_world ? (strip * P.il_world + P.il_rank) * kTileH + ly : P.row_begin + strip * kTileH + ly;
        const bool inside = px < width && py < P.row_end;  // :156-159
        const uint32_t texel = SIMPLE ? shade_pixel_warp_simple<BD>(P, px, py, inside) : shade_pixel_warp<BD, AOV>(P, px, py, inside, pc);
        const uint32_t out_row = P.il_gather ? (P.il_rank * P.il_strips_max + strip) * kTileH + ly : py;

        if (P.n_stage) {  // tile-major exchange: one fully coalesced 128-byte store per rank (NVLink packets 4x larger, 4x fewer)
            const size_t at = ((size_t)(py / kTileH) * tiles_x + tile % tiles_x) * 32u + lane;  // global tile id of this tile
            for (uint32_t p = 0; p < P.n_stage; p++) P.stage[p][at] = texel;
        } else {
        // 128-bit framebuffer stores: lanes with lx in {0,4} gather the 4 texels to their right
        const uint32_t t1 = __shfl_down_sync(kFullMask, texel, 1);
        const uint32_t t2 = __shfl_down_sync(kFullMask, texel, 2);
        const uint32_t t3 = __shfl_down_sync(kFullMask, texel, 3);
        const bool vec = P.vec_store_ok && ((px & ~3u) + 3u < width);  // this lane's group of 4 texels is whole
        if (vec) {
            if ((lx & 3u) == 0u && py < P.row_end) {
                const uint4 v = make_uint4(texel, t1, t2, t3);
                *reinterpret_cast<uint4*>(P.fb + (size_t)out_row * width + px) = v;
                for (uint32_t p = 0; p < P.n_peers; p++) *reinterpret_cast<uint4*>(P.peer_fb[p] + (size_t)out_row * width + px) = v;
            }
        } else if (inside) {
            P.fb[(size_t)out_row * width + px] = texel;
            for (uint32_t p = 0; p < P.n_peers; p++) P.peer_fb[p][(size_t)out_row * width + px] = texel;
        }
        }
#if VRT_TILE_STATS
        __syncwarp();
        if (lane < 8u && g_tile_stats) g_tile_stats[(size_t)tile * 8u + lane] = lane == 7u ? (uint32_t)((clock64() - tick0) >> 5) : warp_stats()[lane];
#endif
        if (P.tile_cost) {  // what this tile cost: the sort key of the next frames' order (lane p also tells peer p)
            const long long ticks = (clock64() - tick0) >> 5;
            const uint16_t c = (uint16_t)(ticks > 65535ll ? 65535ll : (ticks < 1ll ? 1ll : ticks));
            if (lane == 0) P.tile_cost[tile] = c;
            if (lane < P.n_cost_peers) P.peer_cost[lane][tile] = c;
        }
    }
    leave_queue(P, lane);
    if (AOV) flush_counters(P, pc);
}

uint32_t trace_tile_space(const TraceParams& P) {
    const uint32_t rows = P.row_end - P.row_begin;
    const uint32_t tiles_x = (P.cam.image_width + kTileW - 1) / kTileW;
    uint32_t tiles_y = (rows + kTileH - 1) / kTileH;
    if (P.il_world) tiles_y = (tiles_y + P.il_world - 1 - P.il_rank) / P.il_world;  // strips t = k * world + rank < tiles_y
    return tiles_x * tiles_y;
}

template <int BD, bool AOV, bool SIMPLE>
cudaError_t launch_warp_kernel(const TraceParams& P, int num_sms, cudaStream_t stream, LaunchInfo* info) {
    int blocks_per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, trace_warp_kernel<BD, AOV, SIMPLE>, kTunedThreads, 0);
    if (e != cudaSuccess) return e;
    if (blocks_per_sm < 1) return cudaErrorLaunchOutOfResources;
    const uint32_t tiles_x = (P.cam.image_width + kTileW - 1) / kTileW;
    uint32_t tiles_total = trace_tile_space(P);
    if (P.tile_order && P.order_stride > 1u)  // this launch's share of a schedule dealt across ranks: tickets i with offset + i * stride < all tiles
        tiles_total = P.order_offset < tiles_total ? (tiles_total - P.order_offset + P.order_stride - 1u) / P.order_stride : 0u;
    if (tiles_total == 0) return cudaSuccess;
    const uint32_t warps_per_block = kTunedThreads / 32;
    // Short partitioned launches (a rank's share of a frame split over 4-8 GPUs: fewer than ~5 tiles per resident warp) are as long as their most
    // expensive tiles, whose time is a chain of dependent L2 round trips — half the resident warps leave those tiles twice the issue
    // slots and cost nothing when the SMs would run dry anyway: -8 % at 1/4 and 1/8 of a 1080p frame, +29 % on the whole frame
    // (profiles/r02_ab_resident_ctas_C3.txt).
    // Only for a rank's share of a frame (where it was measured): a small whole frame of a lighter scene lost 4 % to it (the bench's
    // ref_default view, 9 216 tiles).
    const bool partitioned = P.il_world > 1u || (P.tile_order && P.order_stride > 1u) || P.row_end - P.row_begin < P.cam.image_height;
    if (SIMPLE && partitioned && blocks_per_sm > 4 &&
        (unsigned long long)tiles_total < 5ull * (unsigned long long)num_sms * (unsigned long long)blocks_per_sm * warps_per_block)
        blocks_per_sm = 4;
    uint32_t grid = (uint32_t)(num_sms * blocks_per_sm);  // one resident wave: a multiple of the SM count
    const uint32_t needed = (tiles_total + warps_per_block - 1) / warps_per_block;
    if (grid > needed) grid = needed;
    trace_warp_kernel<BD, AOV, SIMPLE><<<grid, kTunedThreads, 0, stream>>>(P, tiles_x, tiles_total);
    if (info) info->launches++;
    return cudaGetLastError();
}

// The exchange of VRT_EXCHANGE_PEER_PUSH: after the trace kernel, copy this rank's tiles from its own framebuffer (L2-hot) into the
// same place of every peer's framebuffer.  Same tile enumeration as the trace kernel (ticket -> tile -> pixels); a warp moves four
// tiles per trip — lane = (tile of the four, row of the tile, left / right half) -> one 128-bit load and one 128-bit store per peer.
// Against storing from inside the trace kernel this keeps the NVLink traffic out of the march's instruction stream.
__global__ void __launch_bounds__(256) push_tiles_kernel(const __grid_constant__ TraceParams P, const uint32_t tiles_x, const uint32_t tiles_total) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t sub = lane >> 3, row = (lane >> 1) & 3u, half = lane & 1u;
    const uint32_t width = P.cam.image_width;
    for (uint32_t t0 = warp * 4u; t0 < tiles_total; t0 += n_warps * 4u) {
        const uint32_t t = t0 + sub;
        if (t >= tiles_total) continue;
        const uint32_t tile = P.tile_order ? __ldg(P.tile_order + P.order_offset + t * P.order_stride) : tiles_total - 1u - t;
        const uint32_t px = (tile % tiles_x) * kTileW + half * 4u;
        const uint32_t strip = tile / tiles_x;
        const uint32_t py = (P.il_world ? (strip * P.il_world + P.il_rank) * kTileH : P.row_begin + strip * kTileH) + row;
        if (px + 3u >= width || py >= P.row_end) continue;  // (width % 4 == 0 is a precondition of this mode)
        const size_t at = (size_t)py * width + px;
        const uint4 v = *reinterpret_cast<const uint4*>(P.fb + at);
        for (uint32_t p = 0; p < P.n_peers; p++) *reinterpret_cast<uint4*>(P.peer_fb[p] + at) = v;
    }
}

cudaError_t launch_push_tiles(const TraceParams& P, cudaStream_t stream, LaunchInfo* info) {
    const uint32_t tiles_x = (P.cam.image_width + kTileW - 1) / kTileW;
    uint32_t tiles_total = trace_tile_space(P);
    if (P.tile_order && P.order_stride > 1u) tiles_total = P.order_offset < tiles_total ? (tiles_total - P.order_offset + P.order_stride - 1u) / P.order_stride : 0u;
    if (tiles_total == 0 || P.n_peers == 0) return cudaSuccess;
    const uint32_t warps = (tiles_total + 3u) / 4u;
    uint32_t blocks = (warps + 7u) / 8u;
    if (blocks > 148u * 4u) blocks = 148u * 4u;
    push_tiles_kernel<<<blocks, 256, 0, stream>>>(P, tiles_x, tiles_total);
    if (info) info->launches++;
    return cudaGetLastError();
}

// VRT_EXCHANGE_PEER_TILES, receiving side: the tile-major staging buffer (every rank's tiles of this frame, 128 bytes each) -> the
// row-major framebuffer.  One warp per tile; lanes with lx in {0, 4} write 4 texels as one 128-bit store.
__global__ void __launch_bounds__(256) untile_kernel(const uint32_t* __restrict__ stage, uint32_t* __restrict__ fb, uint32_t width, uint32_t height, uint32_t tiles_x,
                                                     uint32_t tiles_total, uint32_t vec_ok) {
    const uint32_t lane = threadIdx.x & 31u, lx = lane & (kTileW - 1u), ly = lane >> 3;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t t = warp; t < tiles_total; t += n_warps) {
        const uint32_t texel = __ldg(stage + (size_t)t * 32u + lane);
        const uint32_t px = (t % tiles_x) * kTileW + lx, py = (t / tiles_x) * kTileH + ly;
        const uint32_t t1 = __shfl_down_sync(kFullMask, texel, 1), t2 = __shfl_down_sync(kFullMask, texel, 2), t3 = __shfl_down_sync(kFullMask, texel, 3);
        if (vec_ok && (px & ~3u) + 3u < width) {
            if ((lx & 3u) == 0u && py < height) *reinterpret_cast<uint4*>(fb + (size_t)py * width + px) = make_uint4(texel, t1, t2, t3);
        } else if (px < width && py < height) {
            fb[(size_t)py * width + px] = texel;
        }
    }
}

cudaError_t launch_untile(const uint32_t* stage, uint32_t* fb, uint32_t width, uint32_t height, bool vec_ok, cudaStream_t stream, LaunchInfo* info) {
    const uint32_t tiles_x = (width + kTileW - 1) / kTileW, tiles_total = tiles_x * ((height + kTileH - 1) / kTileH);
    uint32_t blocks = (tiles_total + 7u) / 8u;
    if (blocks > 148u * 8u) blocks = 148u * 8u;
    untile_kernel<<<blocks, 256, 0, stream>>>(stage, fb, width, height, tiles_x, tiles_total, vec_ok ? 1u : 0u);
    if (info) info->launches++;
    return cudaGetLastError();
}

cudaError_t launch_trace_tuned(const TraceParams& P, bool aov, cudaStream_t stream, LaunchInfo* info) {
    cudaError_t e;
    static int sm_count[64] = {0};  // per device ordinal; contexts are one-per-GPU and single-threaded
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (sm_count[dev] == 0 && (e = cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    const int n = sm_count[dev];
    // shade_pixel_warp_simple's preconditions (vrt_shade_warp.cuh)
    const bool simple = !aov && P.cam.max_bounce == 1 && P.cam.samples_per_pixel == 1 && (P.sun.enabled == 0u || P.sun.radius == 0.0f) && P.materials_basic != 0u;
    if (P.brick_dim == 4) {
        if (simple) return launch_warp_kernel<4, false, true>(P, n, stream, info);
        return aov ? launch_warp_kernel<4, true, false>(P, n, stream, info) : launch_warp_kernel<4, false, false>(P, n, stream, info);
    }
    if (simple) return launch_warp_kernel<0, false, true>(P, n, stream, info);
    return aov ? launch_warp_kernel<0, true, false>(P, n, stream, info) : launch_warp_kernel<0, false, false>(P, n, stream, info);
}

// ----------------------------------------------------------------------------------------------------
// Explicit rays: a warp takes 32 consecutive rays (two 128-bit loads each), runs the cooperative GridHit and writes
// 32-byte hit records (two 128-bit stores each).  Persistent warps pull blocks of 32 rays from the work counter.
// ----------------------------------------------------------------------------------------------------
template <int BD>
__global__ void __launch_bounds__(kTunedThreads, VRT_TUNED_BLOCKS) trace_rays_kernel(const __grid_constant__ TraceParams P, const float4* __restrict__ rays,
                                                                     uint4* __restrict__ hits, const unsigned long long count) {
    const uint32_t lane = threadIdx.x & 31u;
    const bool ignore_test = P.materials_have_none != 0u;  // CreateRay's ignore type is MAT_NONE (:182)
    const unsigned long long blocks = (count + 31ull) / 32ull;
#if VRT_TMA_MASKS
    if (BD != 4 && P.brick_dim == 16) brick_stage_init();
#endif
    for (;;) {  // blocks of 32 rays from the same work counter as the pixel kernel: rays differ in cost by orders of magnitude
        unsigned long long t = 0ull;
        if (lane == 0) t = draw_ticket(P);
        t = __shfl_sync(kFullMask, t, 0);
        if (t >= blocks) break;
        const unsigned long long base = t * 32ull;
        const unsigned long long i = base + lane;
        const bool active = i < count;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f), d = make_float4(0.f, 0.f, 1.f, 0.f);
        if (active) o = __ldg(rays + 2 * i), d = __ldg(rays + 2 * i + 1);
        const Ray r = CreateRay(v3(o.x, o.y, o.z), v3(d.x, d.y, d.z));
        HitRecord hit;
        hit.point = v3s(0.0f), hit.normal = v3s(0.0f), hit.t = 0.0f, hit.index = 0u;
        TraceInfo ti;
        reset(ti);
        const bool got = grid_hit_warp<BD, 1>(P, r, active, true, ignore_test, hit, ti);
        if (active) {
            hits[2 * i] = make_uint4(got ? 1u : 0u, got ? ti.grid_index : ~0u, got ? ti.voxel_index : ~0u, got ? hit.index : ~0u);
            hits[2 * i + 1] = make_uint4(__float_as_uint(got ? hit.t : 0.0f), __float_as_uint(hit.normal.x), __float_as_uint(hit.normal.y), __float_as_uint(hit.normal.z));
        }
    }
    leave_queue(P, lane);
}

cudaError_t launch_trace_rays(const TraceParams& P, const vrt_ray* rays, vrt_ray_hit* hits, size_t count, cudaStream_t stream, LaunchInfo* info) {
    if (count == 0) return cudaSuccess;
    int dev = 0, sms = 0;
    cudaError_t e;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    const unsigned long long warps_needed = (count + 31) / 32;
    unsigned grid = (unsigned)(sms * VRT_TUNED_BLOCKS);
    const unsigned long long blocks_needed = (warps_needed + kTunedThreads / 32 - 1) / (kTunedThreads / 32);
    if (grid > blocks_needed) grid = (unsigned)blocks_needed;
    if (P.brick_dim == 4) trace_rays_kernel<4><<<grid, kTunedThreads, 0, stream>>>(P, reinterpret_cast<const float4*>(rays), reinterpret_cast<uint4*>(hits), count);
    else trace_rays_kernel<0><<<grid, kTunedThreads, 0, stream>>>(P, reinterpret_cast<const float4*>(rays), reinterpret_cast<uint4*>(hits), count);
    if (info) info->launches++;
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------------
// Derived acceleration structures (consumed by vrt_trav_warp.cuh).
// ----------------------------------------------------------------------------------------------------
VRT_DI bool status_bit(const TraceParams& P, size_t g) { return (__ldg(P.statuses + g / 32) >> (g % 32)) & 1u; }

// cell_rec[g] = everything a brick test at grid cell g needs, in one 128-bit load: the 64-bit voxel mask of the brick (0 if not
// loaded) and where its materials start (brick_type_and_index[brick] & 0x7fffffff, :422).  The shader reaches the same through
// brick_indices -> brick_solid_mask bytes, and brick_indices -> brick_type_and_index -> material_indices on a hit (:337,:415,:422-425).
__global__ void __launch_bounds__(256) build_cell_rec_kernel(const __grid_constant__ TraceParams P, uint4* __restrict__ cell_rec, size_t n_bricks) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_bricks) return;
    unsigned long long occ = 0ull;
    uint32_t mat_base = 0u, brick = 0u;
    if (status_bit(P, g)) {
        brick = g < P.n_brick_indices ? __ldg(P.brick_indices + g) : 0u;
        if (((unsigned long long)brick + 1ull) * 8ull <= P.n_occupancy) occ = __ldg(reinterpret_cast<const unsigned long long*>(P.occupancy) + brick);
        mat_base = (brick < P.n_start_indices ? __ldg(P.start_indices + brick) : 0u) & 0x7fffffffu;
    }
    cell_rec[g] = make_uint4((uint32_t)occ, (uint32_t)(occ >> 32), mat_base, brick);
}

// Directional distance grids.  For each of the 8 direction octants o (bit0: x decreasing, bit1: y decreasing, bit2: z
// decreasing) and every cell p:
//   dist = min over blockers q in the closed octant of p (every coordinate of q - p has the octant's sign or is 0) of the
//          L1 distance |q - p|_1, capped at 126; blockers are loaded bricks (status bit set) and every cell outside the grid.
//          A DDA moves one cell along one axis per step and only along its own octant, so after j steps it is at L1
//          distance exactly j inside the octant: from p it can take dist-1 steps blind, the dist-th lands on a cell to look up.
//   free = no loaded brick at all in the closed octant (bit 7): the DDA can only leave the grid, an exact instant miss.
// L1 is separable into three one-sided scans (x, z, y): walking a line from the end the octant looks towards,
//   d(c) = min(prev(c), 1 + d(next c)),  d(outside) = 0;     free(c) = prev_free(c) && free(next c),  free(outside) = true
// with prev = the previous pass' value (pass x: 0 / not free at loaded bricks, infinity / free elsewhere).  One thread per
// line and output variant, O(cells) whatever the scene.  byte = d | free << 7  (0 = loaded brick; never 255 = border).
//   axis 0 (x): in = status bits,             variants v = xneg                         -> out[v]
//   axis 1 (z): in = out of axis 0 [xneg],    variants v = xneg | zneg << 1             -> out[v]
//   axis 2 (y): in = out of axis 1 [x|z<<1],  variants v = xneg | yneg << 1 | zneg << 2 -> padded dist planes
constexpr uint32_t kDistCap = 126u;

__global__ void __launch_bounds__(128) dist_scan_kernel(const __grid_constant__ TraceParams P, const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                        size_t n_bricks, int axis, const AccelDelta* __restrict__ delta) {
    if (delta && !(delta->force_full || delta->n_new > kAccelMaxNew)) return;  // nothing changed, or dist_patch_kernel has done it
    const size_t dim_x = P.grid.dim_x, dim_y = P.grid.dim_y, dim_z = P.grid.dim_z;
    const size_t lines = axis == 0 ? dim_z * dim_y : (axis == 1 ? dim_x * dim_y : dim_x * dim_z);
    const size_t line = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (line >= lines) return;
    const int v = (int)blockIdx.y;
    size_t base, stride;
    int len, neg, vin;
    if (axis == 0) {
        base = line * dim_x, stride = 1, len = (int)dim_x, neg = v & 1, vin = 0;
    } else if (axis == 1) {
        base = (line % dim_x) + dim_x * dim_z * (line / dim_x), stride = dim_x, len = (int)dim_z, neg = v >> 1, vin = v & 1;
    } else {
        base = line, stride = dim_x * dim_z, len = (int)dim_y, neg = (v >> 1) & 1, vin = (v & 1) | ((v >> 2) << 1);
    }
    uint32_t carry_d = 0u, carry_free = kDistFree;  // the cell beyond the grid: a blocker, but not a brick
    for (int i = 0; i < len; i++) {
        const int c = neg ? i : len - 1 - i;  // walk from the end the octant looks towards
        const size_t g = base + (size_t)c * stride;
        uint32_t prev;
        if (axis == 0) prev = status_bit(P, g) ? 0u : (kDistCap | kDistFree);
        else prev = in[(size_t)vin * n_bricks + g];
        const uint32_t d = min(min(prev & 0x7fu, carry_d + 1u), kDistCap);
        const uint32_t fr = prev & carry_free & kDistFree;
        carry_d = d, carry_free = fr;
        const uint32_t byte = d | fr;
        if (axis == 2) {
            const uint32_t x = (uint32_t)(g % dim_x), z = (uint32_t)((g / dim_x) % dim_z);
            out[(size_t)v * P.dist_plane + (size_t)(x + 1) + ((size_t)(z + 1) << P.dist_log_px) + ((size_t)(c + 1) << (P.dist_log_px + P.dist_log_pz))] = (uint8_t)byte;
        } else {
            out[(size_t)v * n_bricks + g] = (uint8_t)byte;
        }
    }
}

// Status words that were uploaded again: copy them over the live ones and note which bits changed.  Grid.insert re-registers the
// status word of every voxel it touches (brick/Grid.zig:188-189), so most uploaded words are identical to what is there.
__global__ void __launch_bounds__(256) status_merge_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ staged, size_t count, size_t first_word,
                                                           uint32_t dim_x, uint32_t dim_z, uint64_t n_cells, AccelDelta* __restrict__ delta) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t was = dst[i], now = staged[i];
    if (was == now) return;
    dst[i] = now;
    if (was & ~now) delta->force_full = 1u;
    uint32_t added = now & ~was;
    while (added) {
        const uint32_t bit = (uint32_t)__ffs((int)added) - 1u;
        added &= added - 1u;
        const uint64_t g = (uint64_t)(first_word + i) * 32u + bit;
        if (g >= n_cells) continue;  // padding bits of the last word
        const uint32_t slot = atomicAdd(&delta->n_new, 1u);
        if (slot < kAccelMaxNew) {
            delta->cell[slot] = (uint32_t)g;
            const uint32_t x = (uint32_t)(g % dim_x), z = (uint32_t)((g / dim_x) % dim_z), y = (uint32_t)(g / ((uint64_t)dim_x * dim_z));
            atomicMin(&delta->lo[0], x), atomicMin(&delta->lo[1], y), atomicMin(&delta->lo[2], z);
            atomicMax(&delta->hi[0], x), atomicMax(&delta->hi[1], y), atomicMax(&delta->hi[2], z);
        }
    }
}

cudaError_t launch_status_merge(uint32_t* dst, const uint32_t* staged, size_t count, size_t first_word, const vrt_grid_state& grid, AccelDelta* delta,
                                cudaStream_t stream, LaunchInfo* info) {
    if (count == 0) return cudaSuccess;
    status_merge_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(dst, staged, count, first_word, grid.dim_x, grid.dim_z,
                                                                             (uint64_t)grid.dim_x * grid.dim_y * grid.dim_z, delta);
    if (info) info->launches++;
    return cudaGetLastError();
}

// A few bricks were added (status bits 0 -> 1): patch the distance planes in place instead of rebuilding them.  A new brick q can
// only LOWER values, and only for the cells p that have q in their closed octant: dist_o(p) = min(dist_o(p), |q - p|_1), and p's
// octant is not free any more.  One thread per cell and octant; O(cells) bytes touched in the worst case, against the three
// serial line scans of the full rebuild.  Same bytes as a rebuild from scratch (tests/test_accel_update.py).
__device__ __forceinline__ void dist_patch_cell(const TraceParams& P, uint8_t* __restrict__ dist, const AccelDelta* __restrict__ delta, uint32_t n, int o, size_t cell);
__global__ void __launch_bounds__(256) dist_patch_kernel(const __grid_constant__ TraceParams P, uint8_t* __restrict__ dist, const AccelDelta* __restrict__ delta) {
    const uint32_t n = delta->n_new;
    if (n == 0u || n > kAccelMaxNew || delta->force_full) return;
    const uint32_t dim_x = P.grid.dim_x, dim_y = P.grid.dim_y, dim_z = P.grid.dim_z;
    const int o = (int)blockIdx.y;  // bit0: x decreasing, bit1: y decreasing, bit2: z decreasing (dist_scan_kernel)
    // a modest fixed grid walks the cells (most launches find n == 0 above and must cost next to nothing)
    for (size_t cell = (size_t)blockIdx.x * blockDim.x + threadIdx.x; cell < (size_t)dim_x * dim_y * dim_z; cell += (size_t)gridDim.x * blockDim.x)
        dist_patch_cell(P, dist, delta, n, o, cell);
}

__device__ __forceinline__ void dist_patch_cell(const TraceParams& P, uint8_t* __restrict__ dist, const AccelDelta* __restrict__ delta, uint32_t n, int o, size_t cell) {
    const uint32_t dim_x = P.grid.dim_x, dim_z = P.grid.dim_z;
    const int px = (int)(cell % dim_x), pz = (int)((cell / dim_x) % dim_z), py = (int)(cell / ((size_t)dim_x * dim_z));
    // the octant of p must contain at least the extreme new cell
    if ((o & 1) ? px < (int)delta->lo[0] : px > (int)delta->hi[0]) return;
    if ((o & 2) ? py < (int)delta->lo[1] : py > (int)delta->hi[1]) return;
    if ((o & 4) ? pz < (int)delta->lo[2] : pz > (int)delta->hi[2]) return;
    uint32_t best = 0xffffu;
    for (uint32_t j = 0; j < n; j++) {
        const uint32_t g = delta->cell[j];
        const int qx = (int)(g % dim_x), qz = (int)((g / dim_x) % dim_z), qy = (int)(g / (dim_x * dim_z));
        const int dx = (o & 1) ? px - qx : qx - px, dy = (o & 2) ? py - qy : qy - py, dz = (o & 4) ? pz - qz : qz - pz;
        if (dx >= 0 && dy >= 0 && dz >= 0) best = min(best, (uint32_t)(dx + dy + dz));
    }
    if (best == 0xffffu) return;
    const uint32_t padded = (uint32_t)(px + 1) + ((uint32_t)(pz + 1) << P.dist_log_px) + ((uint32_t)(py + 1) << (P.dist_log_px + P.dist_log_pz));
    uint8_t* at = dist + (size_t)o * P.dist_plane + padded;
    const uint32_t cur = *at;
    const uint32_t d = min(min(cur & 0x7fu, best), kDistCap);
    if (d != cur) *at = (uint8_t)d;  // free bit cleared
}

__global__ void accel_delta_reset_kernel(AccelDelta* delta) {
    delta->n_new = 0u, delta->force_full = 0u;
    for (int a = 0; a < 3; a++) delta->lo[a] = 0xffffffffu, delta->hi[a] = 0u;
}

// tmp: 6 * n_bricks bytes of scratch.  The border bytes of `dist` (255) are written once when it is allocated.
cudaError_t launch_build_accel(const TraceParams& P, uint4* cell_rec, uint8_t* dist, uint8_t* tmp, size_t n_bricks, bool occ_only,
                               AccelDelta* delta, cudaStream_t stream, LaunchInfo* info) {
    const unsigned blocks = (unsigned)((n_bricks + 255) / 256);
    if (cell_rec) {
        build_cell_rec_kernel<<<blocks, 256, 0, stream>>>(P, cell_rec, n_bricks);
        if (info) info->launches++;
    }
    if (occ_only) return cudaGetLastError();
    uint8_t* tmp_x = tmp;                 // [2][n]
    uint8_t* tmp_z = tmp + 2 * n_bricks;  // [4][n]
    const size_t dx = P.grid.dim_x, dy = P.grid.dim_y, dz = P.grid.dim_z;
    if (delta) {
        dist_patch_kernel<<<dim3(blocks < 592u ? blocks : 592u, 8), 256, 0, stream>>>(P, dist, delta);
        if (info) info->launches++;
    }
    dist_scan_kernel<<<dim3((unsigned)((dz * dy + 127) / 128), 2), 128, 0, stream>>>(P, nullptr, tmp_x, n_bricks, 0, delta);
    dist_scan_kernel<<<dim3((unsigned)((dx * dy + 127) / 128), 4), 128, 0, stream>>>(P, tmp_x, tmp_z, n_bricks, 1, delta);
    dist_scan_kernel<<<dim3((unsigned)((dx * dz + 127) / 128), 8), 128, 0, stream>>>(P, tmp_z, dist, n_bricks, 2, delta);
    if (info) info->launches += 3;
    if (delta) {
        accel_delta_reset_kernel<<<1, 1, 0, stream>>>(delta);
        if (info) info->launches++;
    }
    return cudaGetLastError();
}

// Rank-major gather layout -> row-major frame after the all-gather of an interleaved partition: strip k of rank r holds
// image rows (k * world + r) * 4 ... + 3.  One thread per 16 bytes (4 texels).
__global__ void __launch_bounds__(256) deinterleave_kernel(const uint4* __restrict__ gathered, uint4* __restrict__ frame, uint32_t width4, uint32_t height,
                                                           uint32_t world, uint32_t strips_max) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)world * strips_max * kTileH * width4;
    if (i >= total) return;
    const uint32_t x = (uint32_t)(i % width4);
    const uint32_t grow = (uint32_t)(i / width4);  // row in the gather layout
    const uint32_t r = grow / (strips_max * kTileH), k = (grow / kTileH) % strips_max, ly = grow % kTileH;
    const uint32_t py = (k * world + r) * kTileH + ly;
    if (py < height) frame[(size_t)py * width4 + x] = gathered[i];
}

cudaError_t launch_deinterleave(const uint32_t* gathered, uint32_t* frame, uint32_t width, uint32_t height, uint32_t world, uint32_t strips_max,
                                cudaStream_t stream, LaunchInfo* info) {
    const size_t total = (size_t)world * strips_max * kTileH * (width / 4);
    deinterleave_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const uint4*>(gathered), reinterpret_cast<uint4*>(frame), width / 4,
                                                                            height, world, strips_max);
    if (info) info->launches++;
    return cudaGetLastError();
}

// Frame barrier of the fused peer-store exchange without a collective library: lane r < world publishes "this rank has finished
// frame `frame`" in rank r's flag array (a system-scope release store through the NVLink peer mapping, after the trace kernel's
// peer stores, which the kernel boundary + cumulative fence order before it), then waits until rank r's flag for the same frame
// has arrived in this rank's array (acquire).  After the kernel every peer's pixels of this frame are visible here.  A peer that
// never arrives trips a ~17 s timeout (2^35 cycles) that raises *error instead of hanging the GPU.
struct PeerFlags {
    uint32_t* flags[8];  // flags[r] = rank r's array of 8 words (this rank's own array included)
};
__global__ void __launch_bounds__(32) peer_barrier_kernel(const PeerFlags peers, const uint32_t rank, const uint32_t world, const uint32_t frame, int* error) {
    const uint32_t r = threadIdx.x;
    if (r >= world) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peers.flags[r] + rank), "r"(frame) : "memory");
    const uint32_t* mine = peers.flags[rank] + r;
    const long long t0 = clock64();
    for (;;) {
        uint32_t seen;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
        if ((int32_t)(seen - frame) >= 0) break;
        if (clock64() - t0 > (1ll << 35)) {
            *reinterpret_cast<volatile int*>(error) = 1;
            break;
        }
    }
}

cudaError_t launch_peer_barrier(uint32_t* const flags[8], uint32_t rank, uint32_t world, uint32_t frame, int* error, cudaStream_t stream, LaunchInfo* info) {
    PeerFlags pf;
    for (int i = 0; i < 8; i++) pf.flags[i] = flags[i];
    peer_barrier_kernel<<<1, 32, 0, stream>>>(pf, rank, world, frame, error);
    if (info) info->launches++;
    return cudaGetLastError();
}

cudaError_t debug_set_tile_stats(uint32_t* device_buffer) {
#if VRT_TILE_STATS
    return cudaMemcpyToSymbol(g_tile_stats, &device_buffer, sizeof(device_buffer));
#else
    (void)device_buffer;
    return cudaErrorNotSupported;
#endif
}

cudaError_t launch_trace(const TraceParams& P, TraceKernel which, bool aov, cudaStream_t stream, LaunchInfo* info) {
    const uint32_t rows = P.row_end - P.row_begin;
    if (rows == 0 || P.cam.image_width == 0) return cudaSuccess;
    if (which == KERNEL_REF) {
        const dim3 block(32, 8);
        const dim3 grid((P.cam.image_width + 31) / 32, (rows + 7) / 8);
        if (aov) trace_ref_kernel<true><<<grid, block, 0, stream>>>(P);
        else trace_ref_kernel<false><<<grid, block, 0, stream>>>(P);
        if (info) info->launches++;
        return cudaGetLastError();
    }
    return launch_trace_tuned(P, aov, stream, info);
}

}  // namespace vrt
