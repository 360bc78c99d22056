// Helper kernels around the LF-walk kernels of kernels_locate.cuh: hit counting, exclusive scan, row expansion,
// left-part assembly of extractUntilBoundary.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.cuh"
#include "walk_lane.h"

namespace fmgpu {

// hits per pattern: min(count, maxMatches), maxMatches <= 0 = unlimited (FmIndex.java:544)
__global__ void k_hits(const int32_t* __restrict__ counts, uint32_t n_pat, int32_t max_hits, int32_t* __restrict__ n_hits) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pat) {
        const int32_t c = counts[i];
        n_hits[i] = (max_hits > 0 && c > max_hits) ? max_hits : c;
    }
}

// exclusive prefix sum int32 -> uint64 in three passes (1024 elements per block)
constexpr int SCAN_BLOCK = 1024;
__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t* total) {
    __shared__ uint64_t warp_sums[32];
    const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint64_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t y = __shfl_up_sync(FULL, x, o);
        if (lane >= (unsigned)o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint64_t s = warp_sums[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t y = __shfl_up_sync(FULL, s, o);
            if (lane >= (unsigned)o) s += y;
        }
        warp_sums[lane] = s;
    }
    __syncthreads();
    const uint64_t before = wid ? warp_sums[wid - 1] : 0;
    if (total) *total = warp_sums[31];
    __syncthreads();
    return before + x - v;
}
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_local(const int32_t* __restrict__ in, uint32_t n, uint64_t* __restrict__ out,
                                                            uint64_t* __restrict__ block_sums) {
    const uint64_t i = (uint64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const uint64_t v = i < n ? (uint64_t)(in[i] > 0 ? in[i] : 0) : 0;
    uint64_t total;
    const uint64_t ex = block_exclusive_scan(v, &total);
    if (i < n) out[i] = ex;
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_sums(uint64_t* __restrict__ block_sums, uint32_t n_blocks, uint64_t* __restrict__ grand) {
    uint64_t carry = 0;
    for (uint32_t base = 0; base < n_blocks; base += SCAN_BLOCK) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t v = i < n_blocks ? block_sums[i] : 0;
        uint64_t total;
        const uint64_t ex = block_exclusive_scan(v, &total);
        if (i < n_blocks) block_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *grand = carry;
}
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_apply(uint64_t* __restrict__ out, uint32_t n, const uint64_t* __restrict__ block_sums,
                                                            const uint64_t* __restrict__ grand) {
    const uint64_t i = (uint64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
    if (i < n) out[i] += block_sums[blockIdx.x];
    if (i == 0) out[n] = *grand;
}

// rows[hit_off[p] + t] = start[p] + t  (one warp per pattern, coalesced)
__global__ void k_expand_rows(const uint32_t* __restrict__ ranges, const int32_t* __restrict__ n_hits, const uint64_t* __restrict__ hit_off,
                              uint32_t n_pat, uint32_t* __restrict__ rows) {
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < n_pat; p += warps) {
        const uint32_t sp = ranges[2 * (uint64_t)p];
        const int32_t k = n_hits[p];
        const uint64_t o = hit_off[p];
        for (int32_t t = (int32_t)lane; t < k; t += 32) rows[o + (uint64_t)t] = sp + (uint32_t)t;
    }
}

// arena[i][offset + q] = left[i][down-1-q]: the left part was produced right-to-left (one warp per item)
__global__ void k_eub_assemble(const uint16_t* __restrict__ left, const int32_t* __restrict__ down_len, uint32_t n, int32_t dst_len,
                               int32_t offset, uint16_t* __restrict__ arena) {
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n; w += warps) {
        int32_t d = down_len[w];
        if (d > dst_len - offset) d = dst_len - offset;  // (a left part that does not fit behind `offset` has status 9)
        const uint64_t slot = (uint64_t)w * (uint64_t)dst_len;
        for (int32_t q = (int32_t)lane; q < d; q += 32) arena[slot + (uint64_t)(offset + q)] = left[slot + (uint64_t)(d - 1 - q)];
    }
}

}  // namespace fmgpu
