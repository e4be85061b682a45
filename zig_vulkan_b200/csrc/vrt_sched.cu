// vrt_sched.cu — the tile schedule of the trace kernel: longest-processing-time-first.
//
// The reference dispatches ceil(W/32) x ceil(H/32) workgroups in raster order and lets the hardware scheduler sort it out
// (ComputePipeline.zig:547-550).  Here a frame is 8x4-pixel tiles pulled from a queue by persistent warps, and tile costs differ
// by two orders of magnitude (a sky tile ends at its first lookup, a tile grazing the terrain marches hundreds of cells for
// 50-60 us).  When a GPU holds only a few tiles per resident warp — a 1080p frame split over 8 GPUs is 8100 tiles on 3552 warps —
// the launch is as long as the unluckiest warp's last tile, so the order matters more than the throughput.  Every tile reports
// what it cost (clock ticks / 32, TraceParams::tile_cost); every `interval` frames the costs are sorted, most expensive first,
// into the order the next frames pull their tiles in.  With several GPUs on one frame the sorted list is dealt round-robin
// (tile i of the list goes to rank i % world): every rank gets the same mix, and because every rank sorts the same cost array
// with the same STABLE sort, all ranks agree on who traces what without talking to each other.
//
// The sort: one pass of a stable counting sort on an 8-bit monotonic key (4-bit exponent / 4-bit mantissa of the 16-bit cost,
// 6 % resolution — plenty for list scheduling).  Three small kernels: per-block histograms, one scan, a stable scatter.
#include "vrt_kernels.cuh"

namespace vrt {

namespace {

constexpr int kSortThreads = 1024;  // tiles per block: 64 blocks for a 1080p frame, 254 for 4K -> a 16 K / 65 K-entry scan
constexpr int kBins = 256;

// monotonic 8-bit key of a 16-bit cost; bin 0 = most expensive
__device__ __forceinline__ uint32_t cost_bin(uint32_t c) {
    uint32_t q;
    if (c < 16u) {
        q = c;
    } else {
        const uint32_t e = 31u - (uint32_t)__clz((int)c);  // 4 .. 15
        q = (e - 3u) * 16u + ((c >> (e - 4u)) & 15u);       // 16 .. 207
    }
    return 255u - q;
}

// The cost arrays are NOT touched here: with a dealt schedule a faster peer may already be writing this frame's costs into them
// (its first frame does not wait for this rank's set-up), and every tile's cost is written before the first sort reads it anyway.
__global__ void __launch_bounds__(256) sched_init_kernel(uint32_t* __restrict__ order, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    order[i] = n - 1u - i;  // bottom-up: ground rows first (vrt_kernels.cu)
}

// hist[bin * nblk + blk] = tiles of block blk in bin
__global__ void __launch_bounds__(kSortThreads) sched_hist_kernel(const uint16_t* __restrict__ cost, uint32_t n, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[kBins];
    if (threadIdx.x < kBins) sh[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t i = blockIdx.x * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&sh[cost_bin(cost[i])], 1u);
    __syncthreads();
    if (threadIdx.x < kBins) hist[threadIdx.x * gridDim.x + blockIdx.x] = sh[threadIdx.x];
}

// exclusive scan of cnt words of shared memory in place (+ the running carry of earlier chunks), all 1024 threads
__device__ __forceinline__ void scan_chunk(uint32_t* stage, uint32_t cnt, uint32_t* warp_sums, uint32_t& carry) {
    const uint32_t per = (cnt + 1023u) / 1024u;
    const uint32_t begin = min(threadIdx.x * per, cnt), end = min(begin + per, cnt);
    uint32_t sum = 0u;
    for (uint32_t i = begin; i < end; i++) sum += stage[i];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t incl = sum;
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)lane >= d) incl += v;
    }
    if (lane == 31u) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0u) {
        uint32_t w = warp_sums[lane];
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, w, d);
            if ((int)lane >= d) w += v;
        }
        warp_sums[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    uint32_t run = carry + incl - sum + (warp ? warp_sums[warp - 1u] : 0u);
    for (uint32_t i = begin; i < end; i++) {
        const uint32_t v = stage[i];
        stage[i] = run;
        run += v;
    }
    __syncthreads();
    if (threadIdx.x == 1023u) carry = run;  // the last thread's running total = everything so far
}

// exclusive scan of m words in place, one CTA; m <= 64 * 1024 words are staged through shared memory so that the global
// loads / stores are coalesced (a thread's serial chunk of `per` words is strided in global memory)
constexpr uint32_t kScanSmemWords = 8192;
__global__ void __launch_bounds__(1024) sched_scan_kernel(uint32_t* __restrict__ data, uint32_t m) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t stage[kScanSmemWords];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0u;
    for (uint32_t base = 0; base < m; base += kScanSmemWords) {  // 2 trips for a 1080p frame, 8 for 4K
        const uint32_t cnt = min(kScanSmemWords, m - base);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < cnt; i += 1024u) stage[i] = data[base + i];
        __syncthreads();
        scan_chunk(stage, cnt, warp_sums, carry);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < cnt; i += 1024u) data[base + i] = stage[i];
    }
}
// order[base(bin, blk) + rank of this tile among the block's tiles of the same bin, in tile order] = tile   (stable)
__global__ void __launch_bounds__(kSortThreads) sched_scatter_kernel(const uint16_t* __restrict__ cost, uint32_t n, const uint32_t* __restrict__ base,
                                                                     uint32_t* __restrict__ order) {
    __shared__ uint16_t warp_count[kSortThreads / 32][kBins];
    if (threadIdx.x < kBins)
        for (int w = 0; w < kSortThreads / 32; w++) warp_count[w][threadIdx.x] = 0;
    __syncthreads();
    const uint32_t i = blockIdx.x * kSortThreads + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const bool valid = i < n;
    const uint32_t bin = valid ? cost_bin(cost[i]) : (uint32_t)kBins;
    const uint32_t same = __match_any_sync(0xffffffffu, bin);
    const uint32_t rank_in_warp = (uint32_t)__popc(same & ((1u << lane) - 1u));
    if (valid && rank_in_warp == 0u) warp_count[warp][bin] = (uint16_t)__popc(same);
    __syncthreads();
    if (threadIdx.x < kBins) {  // per bin (one thread each): exclusive prefix over the warps
        uint32_t acc = 0u;
        for (int w = 0; w < kSortThreads / 32; w++) {
            const uint32_t c = warp_count[w][threadIdx.x];
            warp_count[w][threadIdx.x] = (uint16_t)acc;
            acc += c;
        }
    }
    __syncthreads();
    if (valid) order[base[bin * gridDim.x + blockIdx.x] + warp_count[warp][bin] + rank_in_warp] = i;
}

}  // namespace

size_t sched_scratch_words(uint32_t n_tiles) { return (size_t)kBins * ((n_tiles + kSortThreads - 1) / kSortThreads); }

cudaError_t launch_sched_init(uint32_t* order, uint32_t n_tiles, cudaStream_t stream, LaunchInfo* info) {
    if (n_tiles == 0) return cudaSuccess;
    sched_init_kernel<<<(n_tiles + 255) / 256, 256, 0, stream>>>(order, n_tiles);
    if (info) info->launches++;
    return cudaGetLastError();
}

cudaError_t launch_sched_sort(const uint16_t* cost, uint32_t n_tiles, uint32_t* order, uint32_t* scratch, cudaStream_t stream, LaunchInfo* info) {
    if (n_tiles == 0) return cudaSuccess;
    const uint32_t nblk = (n_tiles + kSortThreads - 1) / kSortThreads;
    sched_hist_kernel<<<nblk, kSortThreads, 0, stream>>>(cost, n_tiles, scratch);
    sched_scan_kernel<<<1, 1024, 0, stream>>>(scratch, (uint32_t)kBins * nblk);
    sched_scatter_kernel<<<nblk, kSortThreads, 0, stream>>>(cost, n_tiles, scratch, order);
    if (info) info->launches += 3;
    return cudaGetLastError();
}

}  // namespace vrt
