// Batched WaveletFixedBlockBoosting.rank / inverseSelect on the index's wavelet structure (the library's public wavelet API,
// wavelet/WaveletFixedBlockBoosting.java:1010-1285 and :1305-1537): one lane per query, the same record walks the FM-index
// kernels use (rank_single, dlevel_descend of lane_logic.h).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.cuh"
#include "kernels_locate.cuh"

namespace fmgpu {

// rank(position, symbol): occurrences of `symbol` in [0, position).  Order of the checks as in the reference (:1012-1026):
// position == 0 -> 0; position > size is clamped; symbol >= alphabet -> 0; negative symbol / negative position index out of
// the arrays (status 9).
__global__ void __launch_bounds__(256) k_wavelet_rank(const DevIndex ix, const int64_t* __restrict__ pos, const int32_t* __restrict__ sym,
                                                      uint32_t n, int64_t* __restrict__ out, int32_t* __restrict__ status) {
    extern __shared__ uint32_t smem[];
    const SmemTables T = stage_tables(ix, smem);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int64_t p = pos[i];
        const int32_t s = sym[i];
        int64_t v = 0;
        int32_t st = 0;
        if (p != 0) {
            if (p > (int64_t)ix.length) p = (int64_t)ix.length;
            if (s < (int32_t)ix.sigma) {
                if (s < 0 || p < 0) {
                    st = 9;
                } else {
                    uint32_t r = 0, a = 0, b = 0, c = 0;
                    st = (int32_t)rank_single(ix, T, (uint32_t)p, (uint32_t)s, &r, &a, &b, &c);
                    v = st ? 0 : (int64_t)r;
                }
            }
        }
        out[i] = v;
        status[i] = st;
    }
}

// inverseSelect(position) = (rank(position, bwt[position]) << 32) | bwt[position], the bare symbol for position 0 (:1336, :1512).
// Single-symbol blocks keep only the low byte of the symbol (:1329-1332) and add the position inside the block to the
// pre-summed boundary ranks (descriptor word 5).  Positions outside [0, size) index out of the arrays: status 9.
__global__ void __launch_bounds__(256) k_wavelet_inverse_select(const DevIndex ix, const int64_t* __restrict__ pos, uint32_t n,
                                                                int64_t* __restrict__ out, int32_t* __restrict__ status) {
    extern __shared__ uint32_t smem[];
    const SmemTables T = stage_tables(ix, smem);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int64_t p64 = pos[i];
        if (p64 < 0 || p64 >= (int64_t)ix.length) {
            out[i] = 0;
            status[i] = 9;
            continue;
        }
        const uint32_t p = (uint32_t)p64;
        const SbDesc sd = T.sb[p >> SB_LOG];
        const uint32_t blk = sd.first_block + ((p & SB_MASK) >> sd.block_log);
        uint32_t r = p & ((1u << sd.block_log) - 1u);
        const Rec32 D = ld256(ix.blocks + blk);
        uint32_t sym = 0, rk = 0;
        if (D.w[1] & 1u) {
            sym = (D.w[1] >> 8) & 0xffffu;
            rk = D.w[5] + r;
        } else {
            uint32_t sec = D.w[0], nrec = D.w[4], levels = 0;
            for (;;) {
                const Rec32 X = ld256(ix.sectors + (sec + r / SECTOR_BITS));
                const Rec32 N = ld256(ix.nodes + nrec);
                if (dlevel_descend(X, N, r % SECTOR_BITS, &r, &nrec, &sec, &sym, &rk, &levels)) break;
            }
        }
        out[i] = p == 0u ? (int64_t)sym : (((int64_t)rk << 32) | (int64_t)sym);
        status[i] = 0;
    }
}

// RrrVector.rankOnes / access per position over the 32-block group records (the (class, offset) -> block table in shared
// memory, like k_locate)
__global__ void __launch_bounds__(256) k_rrr_rank_access(const DevIndex ix, const int32_t* __restrict__ pos, uint32_t n,
                                                         int32_t* __restrict__ rank_out, int32_t* __restrict__ access_out,
                                                         int32_t* __restrict__ status) {
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
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int32_t p = pos[i];
        int32_t rk = 0, bit = 0, st = 0;
        if (p < 0) {  // :360-362 / :316-323
            st = 11;
        } else if ((uint32_t)p >= ix.length) {
            rk = (int32_t)ix.s_total_ones;
            st = 11;
        } else {
            uint32_t b = 0, r = 0;
            const Rec32 G = ld256(sg_addr(ix, (uint32_t)p));
            sampled_access_rank(ix, R, G, (uint32_t)p, &b, &r);
            rk = (int32_t)r;
            bit = (int32_t)b;
        }
        rank_out[i] = rk;
        access_out[i] = bit;
        status[i] = st;
    }
}

}  // namespace fmgpu
