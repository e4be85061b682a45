// vrt_kernels.cu — the sm_100a kernels of the voxel ray-tracing hot path and their launchers.
//
// Compile with:  nvcc -gencode arch=compute_100a,code=sm_100a --fmad=false -lineinfo  (see Makefile).
// --fmad=false is part of the contract: parity with the oracle is bit-exact only if no mul+add pair is
// contracted beyond the explicit fmaf() calls (DESIGN.md "FP discipline").
#include "vrt_kernels.cuh"
#include "vrt_shade.cuh"
#include "vrt_trav_ref.cuh"
#include "vrt_trav_tuned.cuh"

namespace vrt {

// ----------------------------------------------------------------------------------------------------
// Baseline: the reference dispatch shape — one thread per pixel, a warp covers 32 consecutive pixels of a
// row exactly like a 32x32 GLSL workgroup's subgroups do (ComputePipeline.zig:547-550,588-597).
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
// Tuned: persistent CTAs, warp = one 8x4 pixel tile at a time, mask pyramid staged by TMA bulk copy.
// ----------------------------------------------------------------------------------------------------
constexpr int kTunedThreads = 256;
constexpr uint32_t kTileW = 8, kTileH = 4;

VRT_DI uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One thread: arm the mbarrier with the byte count and issue the bulk copies (SASS: UBLKCP + SYNCS).
VRT_DI void stage_pyramid(const TraceParams& P, bool stage_status) {
    SmemHeader* h = reinterpret_cast<SmemHeader*>(vrt_smem);
    const uint32_t mbar = smem_u32(&h->mbar);
    const uint32_t dst_coarse = smem_u32(vrt_smem + sizeof(SmemHeader));
    const uint32_t bytes = P.coarse_bytes + (stage_status ? P.status64_bytes : 0u);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_coarse), "l"(P.coarse),
                 "r"(P.coarse_bytes), "r"(mbar)
                 : "memory");
    if (stage_status) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_coarse + P.coarse_bytes),
                     "l"(P.status64), "r"(P.status64_bytes), "r"(mbar)
                     : "memory");
    }
}

VRT_DI void wait_pyramid() {
    const uint32_t mbar = smem_u32(&reinterpret_cast<SmemHeader*>(vrt_smem)->mbar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "VRT_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
        "@p bra VRT_DONE;\n"
        "bra VRT_WAIT;\n"
        "VRT_DONE:\n"
        "}\n" ::"r"(mbar)
        : "memory");
}

template <bool AOV>
__global__ void __launch_bounds__(kTunedThreads, 2)
    trace_tuned_kernel(const __grid_constant__ TraceParams P, const uint32_t tiles_x, const uint32_t tiles_total, const uint32_t stage_status) {
    SmemHeader* h = reinterpret_cast<SmemHeader*>(vrt_smem);
    if (threadIdx.x == 0) {
        h->status_in_smem = stage_status;
        h->coarse_bytes = P.coarse_bytes;
        stage_pyramid(P, stage_status != 0u);
    }
    __syncthreads();
    wait_pyramid();

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lx = lane & (kTileW - 1u), ly = lane >> 3;
    const uint32_t width = P.cam.image_width;
    PixelCounters pc = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};

    for (;;) {
        unsigned long long t = 0ull;
        if (lane == 0) t = atomicAdd(P.tile_counter, 1ull) - P.tile_base;
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= (unsigned long long)tiles_total) break;
        const uint32_t tile = (uint32_t)t;
        const uint32_t px = (tile % tiles_x) * kTileW + lx;
        const uint32_t py = P.row_begin + (tile / tiles_x) * kTileH + ly;
        const bool inside = px < width && py < P.row_end;  // :156-159
        uint32_t texel = 0u;
        if (inside) texel = shade_pixel<TunedTrav, AOV>(P, px, py, pc);

        // 128-bit framebuffer stores: lanes with lx in {0,4} gather the 4 texels to their right
        const uint32_t t1 = __shfl_down_sync(0xffffffffu, texel, 1);
        const uint32_t t2 = __shfl_down_sync(0xffffffffu, texel, 2);
        const uint32_t t3 = __shfl_down_sync(0xffffffffu, texel, 3);
        const bool vec = P.vec_store_ok && ((px & ~3u) + 3u < width);  // this lane's group of 4 texels is whole
        if (vec) {
            if ((lx & 3u) == 0u && py < P.row_end) {
                const uint4 v = make_uint4(texel, t1, t2, t3);
                *reinterpret_cast<uint4*>(P.fb + (size_t)py * width + px) = v;
                for (uint32_t p = 0; p < P.n_peers; p++) *reinterpret_cast<uint4*>(P.peer_fb[p] + (size_t)py * width + px) = v;
            }
        } else if (inside) {
            P.fb[(size_t)py * width + px] = texel;
            for (uint32_t p = 0; p < P.n_peers; p++) P.peer_fb[p][(size_t)py * width + px] = texel;
        }
    }
    if (AOV) flush_counters(P, pc);
}

cudaError_t launch_trace_tuned(const TraceParams& P, bool aov, cudaStream_t stream, LaunchInfo* info) {
    if (P.brick_dim != 4) {
        // 8^3 / 16^3 bricks (the C4 extension) have no u64-per-brick mask; they run the transliteration.
        const uint32_t rows = P.row_end - P.row_begin;
        const dim3 block(32, 8);
        const dim3 grid((P.cam.image_width + 31) / 32, (rows + 7) / 8);
        if (aov) trace_ref_kernel<true><<<grid, block, 0, stream>>>(P);
        else trace_ref_kernel<false><<<grid, block, 0, stream>>>(P);
        if (info) info->launches++;
        return cudaGetLastError();
    }
    // status64 is staged into shared memory when two CTAs of it fit next to L1 (<= 64 KiB each)
    const bool stage_status = P.status64_bytes <= 64u * 1024u;
    const size_t smem = sizeof(SmemHeader) + P.coarse_bytes + (stage_status ? P.status64_bytes : 0u);
    cudaError_t e;
    static int sm_count[64] = {0};  // per device ordinal; contexts are one-per-GPU and single-threaded
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (sm_count[dev] == 0) {
        if ((e = cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(trace_tuned_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(trace_tuned_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)) != cudaSuccess) return e;
    }
    const int num_sms = sm_count[dev];
    int blocks_per_sm = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, aov ? trace_tuned_kernel<true> : trace_tuned_kernel<false>, kTunedThreads, smem)) != cudaSuccess) return e;
    if (blocks_per_sm < 1) return cudaErrorLaunchOutOfResources;
    const uint32_t rows = P.row_end - P.row_begin;
    const uint32_t tiles_x = (P.cam.image_width + kTileW - 1) / kTileW;
    const uint32_t tiles_y = (rows + kTileH - 1) / kTileH;
    const uint32_t tiles_total = tiles_x * tiles_y;
    const uint32_t warps_per_block = kTunedThreads / 32;
    uint32_t grid = (uint32_t)(num_sms * blocks_per_sm);
    const uint32_t needed = (tiles_total + warps_per_block - 1) / warps_per_block;
    if (grid > needed) grid = needed;
    if (aov) trace_tuned_kernel<true><<<grid, kTunedThreads, smem, stream>>>(P, tiles_x, tiles_total, stage_status ? 1u : 0u);
    else trace_tuned_kernel<false><<<grid, kTunedThreads, smem, stream>>>(P, tiles_x, tiles_total, stage_status ? 1u : 0u);
    if (info) {
        info->launches++;
        info->counter_advance = (unsigned long long)tiles_total + (unsigned long long)grid * warps_per_block;  // every warp overshoots once
    }
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------------
// Derived mask pyramid (see vrt_trav_tuned.cuh for how it is consumed).
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) build_occ_dense_kernel(const __grid_constant__ TraceParams P, unsigned long long* __restrict__ occ_dense,
                                                              size_t n_bricks) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_bricks) return;
    unsigned long long occ = 0ull;
    const uint32_t word = __ldg(P.statuses + g / 32);
    if ((word >> (g % 32)) & 1u) {
        const unsigned long long bi = __ldg(P.brick_indices + g);
        if ((bi + 1ull) * 8ull <= P.n_occupancy) occ = __ldg(reinterpret_cast<const unsigned long long*>(P.occupancy) + bi);
    }
    occ_dense[g] = occ;
}

__global__ void __launch_bounds__(256) build_status64_kernel(const __grid_constant__ TraceParams P, unsigned long long* __restrict__ status64,
                                                             uint32_t* __restrict__ coarse, size_t n_super) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long bits = 0ull;
    if (s < n_super) {
        const uint32_t sx = (uint32_t)(s % P.sdim_x);
        const uint32_t sz = (uint32_t)((s / P.sdim_x) % P.sdim_z);
        const uint32_t sy = (uint32_t)(s / ((size_t)P.sdim_x * P.sdim_z));
        for (uint32_t ly = 0; ly < 4; ly++) {
            const uint32_t y = sy * 4 + ly;
            if (y >= P.grid.dim_y) break;
            for (uint32_t lz = 0; lz < 4; lz++) {
                const uint32_t z = sz * 4 + lz;
                if (z >= P.grid.dim_z) break;
                for (uint32_t lx = 0; lx < 4; lx++) {
                    const uint32_t x = sx * 4 + lx;
                    if (x >= P.grid.dim_x) break;
                    const size_t g = x + (size_t)P.grid.dim_x * (z + (size_t)P.grid.dim_z * y);
                    const uint32_t word = __ldg(P.statuses + g / 32);
                    if ((word >> (g % 32)) & 1u) bits |= 1ull << (lx + 4 * (lz + 4 * ly));
                }
            }
        }
        status64[s] = bits;
    }
    // 1 bit per super cell; blockDim is a multiple of 32 and s is warp-contiguous
    const uint32_t any = __ballot_sync(0xffffffffu, bits != 0ull);
    if ((threadIdx.x & 31) == 0 && s < n_super) coarse[s / 32] = any;
}

cudaError_t launch_build_accel(const TraceParams& P, unsigned long long* occ_dense, unsigned long long* status64, uint32_t* coarse,
                               size_t n_bricks, size_t n_super, cudaStream_t stream, LaunchInfo* info) {
    if (occ_dense) {
        const unsigned blocks = (unsigned)((n_bricks + 255) / 256);
        build_occ_dense_kernel<<<blocks, 256, 0, stream>>>(P, occ_dense, n_bricks);
        if (info) info->launches++;
    }
    {
        const unsigned blocks = (unsigned)((n_super + 255) / 256);
        build_status64_kernel<<<blocks, 256, 0, stream>>>(P, status64, coarse, n_super);
        if (info) info->launches++;
    }
    return cudaGetLastError();
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
