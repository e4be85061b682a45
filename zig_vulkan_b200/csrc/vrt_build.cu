// vrt_build.cu — BrickGrid.insert (src/modules/voxel_rt/brick/Grid.zig:129-194, MaterialAllocator.zig:34-43) for a whole
// batch of voxels on the device, with the result a sequential loop of inserts would leave in the five reference buffers:
//   * a brick is allocated when its first voxel arrives (active_bricks.fetchAdd, :147), so NEW bricks are numbered in the order
//     of their first occurrence in the batch: first occurrence = atomicMin of the input position per grid cell, numbering =
//     exclusive scan of the first-occurrence flags over the input order;
//   * the brick's material block is allocated by the same insert (nextEntry bumps by brick_bits, :161-167), so
//     start_indices[brick] = brick * brick_bits;
//   * occupancy and status bits are OR-ed (:180-188) — order-free;
//   * the material index of a voxel inserted more than once is the LAST one written (:173-175): the largest input position per
//     voxel wins an atomicMax in a scratch block per touched brick.
// Not the hot path (it runs when the scene changes); the scan is a plain three-level block scan.
#include <cstdint>

#include "../../include/vrt.h"
#include "vrt_kernels.cuh"

namespace vrt {

namespace {

constexpr uint32_t kNoPos = 0xffffffffu;
constexpr int kScanBlock = 1024;

struct InsertGeom {
    uint32_t voxel_dim_x, voxel_dim_y, voxel_dim_z, dim_x, dim_z, brick_dim, brick_bits;
};

__device__ __forceinline__ bool locate(const InsertGeom& g, const uint32_t* __restrict__ xyzm, size_t i, uint32_t& cell, uint32_t& nth) {
    const uint4 v = reinterpret_cast<const uint4*>(xyzm)[i];
    if (v.x >= g.voxel_dim_x || v.y >= g.voxel_dim_y || v.z >= g.voxel_dim_z) return false;  // :130-132
    const uint32_t d = g.brick_dim;
    const uint32_t fy = g.voxel_dim_y - 1u - v.y;                     // :135 Y flip
    cell = (v.x / d) + g.dim_x * ((v.z / d) + g.dim_z * (fy / d));    // gridAt :206-211
    nth = (v.x % d) + d * ((v.z % d) + d * (fy % d));                 // voxelAt :198-203
    return true;
}

// pass 1: first input position per grid cell; any out-of-range voxel raises *bad
__global__ void __launch_bounds__(256) first_pos_kernel(InsertGeom g, const uint32_t* __restrict__ xyzm, size_t n, uint32_t* __restrict__ first_pos, int* __restrict__ bad) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    uint32_t cell, nth;
    if (!locate(g, xyzm, i, cell, nth)) {
        *bad = 1;
        return;
    }
    atomicMin(first_pos + cell, (uint32_t)i);
}

// pass 2: per input voxel, (is the first voxel of a NEW brick) << 32 | (is the first voxel of its cell in this batch)
__global__ void __launch_bounds__(256) flag_kernel(InsertGeom g, const uint32_t* __restrict__ xyzm, size_t n, const uint32_t* __restrict__ first_pos,
                                                   const uint32_t* __restrict__ statuses, unsigned long long* __restrict__ flags) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    uint32_t cell, nth;
    unsigned long long f = 0ull;
    if (locate(g, xyzm, i, cell, nth) && first_pos[cell] == (uint32_t)i) {
        const bool loaded = (statuses[cell >> 5] >> (cell & 31u)) & 1u;  // BrickStatusMask.read (:141)
        f = 1ull | (loaded ? 0ull : (1ull << 32));
    }
    flags[i] = f;
}

// block-level exclusive scan of u64 (both 32-bit halves are independent counters that never carry: n < 2^32)
__global__ void __launch_bounds__(kScanBlock) scan_block_kernel(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out, size_t n,
                                                                unsigned long long* __restrict__ block_sums) {
    __shared__ unsigned long long warp_sums[kScanBlock / 32];
    const size_t i = (size_t)blockIdx.x * kScanBlock + threadIdx.x;
    const unsigned long long v = i < n ? in[i] : 0ull;
    unsigned long long incl = v;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (unsigned)o) incl += up;
    }
    if (lane == 31u) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long up = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= (unsigned)o) w += up;
        }
        warp_sums[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const unsigned long long base = warp ? warp_sums[warp - 1] : 0ull;
    if (i < n) out[i] = base + incl - v;
    if (block_sums && threadIdx.x == kScanBlock - 1) block_sums[blockIdx.x] = base + incl;
}
__global__ void __launch_bounds__(kScanBlock) scan_add_kernel(unsigned long long* __restrict__ data, size_t n, const unsigned long long* __restrict__ block_offsets) {
    const size_t i = (size_t)blockIdx.x * kScanBlock + threadIdx.x;
    if (i < n) data[i] += block_offsets[blockIdx.x];
}

// pass 3 (first voxel of a cell): allocate the brick if new, remember the cell's scratch block
__global__ void __launch_bounds__(256) allocate_kernel(InsertGeom g, const uint32_t* __restrict__ xyzm, size_t n, uint32_t* __restrict__ first_pos,
                                                       const unsigned long long* __restrict__ flags, const unsigned long long* __restrict__ ranks, uint32_t active_before,
                                                       uint32_t* __restrict__ statuses, uint32_t* __restrict__ brick_indices, uint32_t* __restrict__ start_indices) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const unsigned long long f = flags[i];
    if (!(f & 1ull)) return;
    uint32_t cell, nth;
    locate(g, xyzm, i, cell, nth);
    const unsigned long long r = ranks[i];
    if (f >> 32) {
        const uint32_t brick = active_before + (uint32_t)(r >> 32);  // fetchAdd order (:147)
        brick_indices[cell] = brick;                                  // :192
        start_indices[brick] = (brick * g.brick_bits) & 0x7fffffffu;  // nextEntry (:161-166), type bit 0 = voxel_start_index
        atomicOr(statuses + (cell >> 5), 1u << (cell & 31u));         // :188
    }
    first_pos[cell] = (uint32_t)r;  // from here on: the cell's scratch block (touched-brick rank)
}

// pass 4: occupancy bits + the last writer of every voxel
__global__ void __launch_bounds__(256) occupy_kernel(InsertGeom g, const uint32_t* __restrict__ xyzm, size_t n, const uint32_t* __restrict__ cell_rank,
                                                     const uint32_t* __restrict__ brick_indices, uint32_t* __restrict__ occupancy_words, uint32_t* __restrict__ last_writer) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    uint32_t cell, nth;
    locate(g, xyzm, i, cell, nth);
    const uint32_t brick = brick_indices[cell];
    const size_t bit = (size_t)brick * g.brick_bits + nth;  // byte bit / 8, bit bit % 8 (:179-182); little-endian words
    atomicOr(occupancy_words + (bit >> 5), 1u << (bit & 31u));
    atomicMax(last_writer + (size_t)cell_rank[cell] * g.brick_bits + nth, (uint32_t)i + 1u);
}

// pass 5: the last writer stores the material index
__global__ void __launch_bounds__(256) material_kernel(InsertGeom g, const uint32_t* __restrict__ xyzm, size_t n, const uint32_t* __restrict__ cell_rank,
                                                       const uint32_t* __restrict__ brick_indices, const uint32_t* __restrict__ start_indices,
                                                       const uint32_t* __restrict__ last_writer, uint8_t* __restrict__ material_indices) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    uint32_t cell, nth;
    locate(g, xyzm, i, cell, nth);
    if (last_writer[(size_t)cell_rank[cell] * g.brick_bits + nth] != (uint32_t)i + 1u) return;
    const uint32_t start = start_indices[brick_indices[cell]] & 0x7fffffffu;
    material_indices[(size_t)start + nth] = (uint8_t)xyzm[4 * i + 3];  // :173-175
}

cudaError_t exclusive_scan(unsigned long long* data, size_t n, unsigned long long* scratch, cudaStream_t stream, unsigned long long* total_out_device) {
    // scratch: block sums of every level, laid out one after the other (n / 1024 + n / 1024^2 + ... + 3 entries)
    const size_t blocks = (n + kScanBlock - 1) / kScanBlock;
    scan_block_kernel<<<(unsigned)blocks, kScanBlock, 0, stream>>>(data, data, n, scratch);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (blocks == 1) {
        if (total_out_device) e = cudaMemcpyAsync(total_out_device, scratch, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stream);
        return e;
    }
    e = exclusive_scan(scratch, blocks, scratch + blocks, stream, total_out_device);  // block sums -> block offsets (total = sum of all)
    if (e != cudaSuccess) return e;
    scan_add_kernel<<<(unsigned)blocks, kScanBlock, 0, stream>>>(data, n, scratch);
    return cudaGetLastError();
}

}  // namespace

size_t insert_scan_scratch_entries(size_t n) {
    size_t total = 1, level = n;
    while (level > 1) {
        level = (level + kScanBlock - 1) / kScanBlock;
        total += level;
    }
    return total + 1;
}

cudaError_t launch_insert_prepare(const InsertBuffers& B, const uint32_t* xyzm, size_t n, cudaStream_t stream, LaunchInfo* info) {
    const InsertGeom g = {B.state.voxel_dim_x, B.state.voxel_dim_y, B.state.voxel_dim_z, B.state.dim_x, B.state.dim_z, B.brick_dim, B.brick_dim * B.brick_dim * B.brick_dim};
    const unsigned blocks = (unsigned)((n + 255) / 256);
    cudaError_t e;
    if ((e = cudaMemsetAsync(B.first_pos, 0xff, B.n_cells * sizeof(uint32_t), stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(B.totals, 0, 4 * sizeof(unsigned long long), stream)) != cudaSuccess) return e;
    first_pos_kernel<<<blocks, 256, 0, stream>>>(g, xyzm, n, B.first_pos, reinterpret_cast<int*>(B.totals + 2));
    flag_kernel<<<blocks, 256, 0, stream>>>(g, xyzm, n, B.first_pos, B.statuses, B.flags);
    if ((e = cudaMemcpyAsync(B.ranks, B.flags, n * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stream)) != cudaSuccess) return e;
    if ((e = exclusive_scan(B.ranks, n, B.scan_scratch, stream, B.totals)) != cudaSuccess) return e;
    if (info) info->launches += 4;
    return cudaGetLastError();
}

cudaError_t launch_insert_commit(const InsertBuffers& B, const uint32_t* xyzm, size_t n, uint32_t active_before, uint32_t* last_writer, cudaStream_t stream,
                                 LaunchInfo* info) {
    const InsertGeom g = {B.state.voxel_dim_x, B.state.voxel_dim_y, B.state.voxel_dim_z, B.state.dim_x, B.state.dim_z, B.brick_dim, B.brick_dim * B.brick_dim * B.brick_dim};
    const unsigned blocks = (unsigned)((n + 255) / 256);
    allocate_kernel<<<blocks, 256, 0, stream>>>(g, xyzm, n, B.first_pos, B.flags, B.ranks, active_before, B.statuses, B.brick_indices, B.start_indices);
    occupy_kernel<<<blocks, 256, 0, stream>>>(g, xyzm, n, B.first_pos, B.brick_indices, reinterpret_cast<uint32_t*>(B.occupancy), last_writer);
    material_kernel<<<blocks, 256, 0, stream>>>(g, xyzm, n, B.first_pos, B.brick_indices, B.start_indices, last_writer, B.material_indices);
    if (info) info->launches += 3;
    return cudaGetLastError();
}

}  // namespace vrt
