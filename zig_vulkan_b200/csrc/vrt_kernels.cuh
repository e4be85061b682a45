// vrt_kernels.cuh — host-callable launchers of the device code (implemented in vrt_kernels.cu).
#pragma once

#include "vrt_device.cuh"

namespace vrt {

enum TraceKernel : int {
    KERNEL_REF = 0,    // transliteration, one thread per pixel (VRT_FLAG_BASELINE and the AOV path)
    KERNEL_TUNED = 1,  // warp-tile traversal over the derived mask pyramid
};

struct LaunchInfo {
    uint32_t launches;                 // kernels enqueued
    unsigned long long counter_advance;  // how far the launch moves *tile_counter (tuned kernel)
};

// Enqueue the kernels that trace rows [P.row_begin, P.row_end) into P.fb.
cudaError_t launch_trace(const TraceParams& P, TraceKernel which, bool aov, cudaStream_t stream, LaunchInfo* info);

// Rebuild the derived mask pyramid (occ_dense / status64 / coarse) from the reference-format buffers.
cudaError_t launch_build_accel(const TraceParams& P, unsigned long long* occ_dense, unsigned long long* status64,
                               uint32_t* coarse, size_t n_bricks, size_t n_super, cudaStream_t stream, LaunchInfo* info);

}  // namespace vrt
