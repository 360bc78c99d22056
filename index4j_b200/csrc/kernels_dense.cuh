// Device-side denser sampling of the SA rows (layout.h: DevIndex::dmarks / dsa), built once per load by the LF kernels
// themselves.  The serialized FmIndex keeps suffixes[] for the rows whose suffix starts at a multiple of sampleRate
// (fm/FmIndex.java:343-357) and locate LF-walks every hit to the nearest such row (:531-537), (sampleRate - 1) / 2 steps on
// average.  Walking the text once from every sampled row — sampleRate - 1 LF steps each, all walks in lockstep — visits every
// row together with its text position, so the rows of all multiples of dense_rate can be marked and given their position:
//   k_dense_seeds   row number of every sampled row (the RRR vector only answers access / rank: one thread per row)
//   k_dense_walk<0> walk, mark the rows whose position is a multiple of dense_rate (atomicOr into the mark records)
//   k_dense_popc / the exclusive scan of kernels_lf.cuh / k_dense_fill : marked rows before every record
//   k_dense_walk<1> the same walk again, now writing dsa[rank of the row] = position
// locate then tests the plain mark record instead of the RRR vector (k_locate<., true>) and reads dsa.  Results are the
// reference's: the position of a hit is the sampled position plus the distance walked, whichever sampled row ends the walk.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.cuh"
#include "lf_lane.h"

namespace fmgpu {

constexpr int DENSE_THREADS = 512;

__global__ void __launch_bounds__(DENSE_THREADS) k_dense_seeds(const DevIndex ix, uint32_t* __restrict__ seed_row, uint32_t n_seeds) {
    extern __shared__ uint32_t smem[];
    uint16_t* inv = reinterpret_cast<uint16_t*>(smem);
    uint16_t* cbase = inv + 32768;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(ix.rrr_inv);
        for (uint32_t i = threadIdx.x; i < 16384u; i += blockDim.x) smem[i] = __ldg(src + i);
        if (threadIdx.x < 16) cbase[threadIdx.x] = __ldg(ix.rrr_cbase + threadIdx.x);
    }
    __syncthreads();
    RrrTab R;
    R.inv = inv;
    R.cbase = cbase;
    for (uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; row < ix.length; row += (uint64_t)gridDim.x * blockDim.x) {
        const Rec32 G = ld256(sg_addr(ix, (uint32_t)row));
        uint32_t bit = 0, rank = 0;
        sampled_access_rank(ix, R, G, (uint32_t)row, &bit, &rank);
        if (bit && rank < n_seeds) seed_row[rank] = (uint32_t)row;
    }
}

// PASS 0: mark; PASS 1: write dsa.  Item k < n_seeds: the k-th sampled row (position suffixes[k]); item n_seeds: row 0, the
// sentinel's suffix (position length - 1), which covers the positions behind the last multiple of sampleRate.
template <int PASS>
__global__ void __launch_bounds__(DENSE_THREADS) k_dense_walk(const DevIndex ix, const uint32_t* __restrict__ seed_row, uint32_t n_seeds,
                                                              uint32_t dense_rate, uint32_t* __restrict__ mark_words,
                                                              uint32_t* __restrict__ dsa, uint32_t n_dense, unsigned int* fail_flag) {
    extern __shared__ uint32_t smem[];
    const SmemTables T = stage_tables(ix, smem);
    auto visit = [&](uint32_t row, uint32_t p) {
        if (p % dense_rate != 0u) return;
        const uint32_t q = row / DENSE_ROWS_PER_REC, o = row % DENSE_ROWS_PER_REC;
        if (PASS == 0) {
            atomicOr(mark_words + (uint64_t)q * 8u + 1u + (o >> 5), 1u << (o & 31u));
        } else {
            const Rec32 G = *reinterpret_cast<const Rec32*>(mark_words + (uint64_t)q * 8u);
            uint32_t bit = 0, rank = 0;
            dense_access_rank(G, row, &bit, &rank);
            if (bit && rank < n_dense) dsa[rank] = p;
            else atomicExch(fail_flag, 1u);
        }
    };
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k <= n_seeds; k += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t row = 0, p = 0;
        const int what = dense_item(ix, seed_row, n_seeds, k, &row, &p);
        if (what == 1) continue;
        if (what == 2 || !dense_walk_item(ix, T, row, p, visit)) atomicExch(fail_flag, 1u);  // not a consistent index: no dense samples
    }
}

__global__ void k_dense_popc(const Rec32* __restrict__ marks, uint32_t n_rec, int32_t* __restrict__ ones) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_rec) return;
    const Rec32 G = marks[q];
    uint32_t n = 0;
#pragma unroll
    for (int k = 1; k < 8; ++k) n += (uint32_t)__popc(G.w[k]);
    ones[q] = (int32_t)n;
}
__global__ void k_dense_fill(Rec32* __restrict__ marks, uint32_t n_rec, const uint64_t* __restrict__ before) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n_rec) marks[q].w[0] = (uint32_t)before[q];
}

}  // namespace fmgpu
