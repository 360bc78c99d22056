// Index PRODUCTION on the device, stage 2: from the suffix array (already on the GPU) to the BWT and the
// sampled structures of the FmIndex constructor (fm/FmIndex.java:343-394) — the BWT gather, the sampled-row marks, the SA
// samples in row order and the inverse-SA samples — so that 2 + 0.125 + 0.25 bytes per char go back to the host instead of the
// 4-byte suffix array, and the host's three passes of random accesses over the text disappear.  Not part of the query path.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels_lf.cuh"

namespace fmgpu {

// row i: bwt[i] = text[SA[i] - 1] (the sentinel for SA[i] == 0, :374-394); row sampled iff SA[i] % sampleRate == 0 (:343-357):
// mask word (one bit per row, LSB first), its popcount, and positions[SA[i] / sampleRate] = i (:359-370)
__global__ void __launch_bounds__(256) k_build_bwt_mask(const uint16_t* __restrict__ codes, const int32_t* __restrict__ sa, uint32_t length,
                                                        uint32_t sample_rate, uint16_t* __restrict__ bwt, uint32_t* __restrict__ mask,
                                                        int32_t* __restrict__ counts, int32_t* __restrict__ positions) {
    const uint32_t n_words = (length + 31u) / 32u;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_words; w += warps) {
        const uint32_t i = w * 32u + lane;
        bool samp = false;
        if (i < length) {
            const uint32_t p = (uint32_t)sa[i];
            bwt[i] = codes[p ? p - 1u : length - 1u];
            samp = p % sample_rate == 0u;
            if (samp && positions) positions[p / sample_rate] = (int32_t)i;
        }
        const unsigned m = __ballot_sync(FULL, samp);
        if (lane == 0) {
            mask[w] = m;
            counts[w] = __popc(m);
        }
    }
}

// suffixes[k] = SA value of the k-th sampled row (:343-357): word offsets from the exclusive scan of the popcounts
__global__ void __launch_bounds__(256) k_build_compact(const int32_t* __restrict__ sa, const uint32_t* __restrict__ mask,
                                                       const uint64_t* __restrict__ word_off, uint32_t length, int32_t* __restrict__ suffixes,
                                                       int32_t* __restrict__ positions, uint32_t sample_rate) {
    const uint32_t n_words = (length + 31u) / 32u;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_words; w += warps) {
        const unsigned m = mask[w];
        if ((m >> lane) & 1u) suffixes[word_off[w] + (uint64_t)__popc(m & ((1u << lane) - 1u))] = sa[w * 32u + lane];
    }
    // positions[(length - 1) / sampleRate + 1] = positions[0] (:369: the cyclic wrap entry); positions[0] was written by
    // k_build_bwt_mask (the row of text position 0), which ran before this kernel
    if (positions && blockIdx.x == 0 && threadIdx.x == 0) positions[(length - 1u) / sample_rate + 1u] = positions[0];
}

}  // namespace fmgpu
