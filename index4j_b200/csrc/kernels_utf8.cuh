// UTF-8 byte patterns (the reference's convertBytePatternToCharPattern + count / locate, fm/FmIndex.java:239-298): the bytes
// cross PCIe (1 byte per char for log text instead of the 2 of a Java char[]) and are decoded on the device by the pre-pass,
// which also builds the pattern descriptors the backward-search kernel consumes.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.cuh"
#include "utf8_lane.h"

namespace fmgpu {

// One thread per pattern: decode bytes[off[i], off[i+1]) into chars[off[i] ..] (a pattern never has more chars than bytes,
// so the patterns' char ranges cannot overlap), descriptor {offset, length in chars, code of the last char}, length histogram.
// A pattern whose conversion throws gets length 0 here; k_utf8_merge writes its status after the search.
__global__ void __launch_bounds__(256) k_prepass_utf8(const uint8_t* __restrict__ bytes, const uint64_t* __restrict__ pat_off, uint32_t n_pat,
                                                      const uint16_t* __restrict__ char2code, uint16_t* chars,
                                                      PatDesc* __restrict__ pats, uint32_t* __restrict__ bins, int32_t* __restrict__ conv_status,
                                                      int32_t* __restrict__ conv_value, uint32_t kmer_q, uint32_t kmer_stride, uint32_t sigma) {
    __shared__ uint32_t h[LEN_BINS];
    for (uint32_t i = threadIdx.x; i < LEN_BINS; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pat; i += gridDim.x * blockDim.x) {
        const uint64_t a = pat_off[i], b = pat_off[i + 1];
        int32_t value = 0;
        uint32_t last = 0;
        const int64_t n = b > a ? utf8_convert(bytes, a, b, chars + a, &value, &last) : 0;
        PatDesc d;
        d.off = a;
        d.len = n > 0 ? (uint32_t)n : 0u;
        (void)last;
        d.last = d.len ? pattern_start(chars, a + d.len, d.len, char2code, kmer_q, kmer_stride, sigma) : 0u;  // reads this thread's own chars
        pats[i] = d;
        conv_status[i] = n < 0 ? (int32_t)(-n) : 0;
        conv_value[i] = value;
        atomicAdd(&h[d.len < LEN_BINS - 1 ? d.len : LEN_BINS - 1], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < LEN_BINS; i += blockDim.x)
        if (h[i]) atomicAdd(&bins[i], h[i]);
}

// status of the conversion overrides the search's (the Java caller never reaches count / locate); with `values` the offending
// code point of "Found a character that exceeds" goes where the count would be
__global__ void k_utf8_merge(const int32_t* __restrict__ conv_status, const int32_t* __restrict__ conv_value, uint32_t n_pat,
                             int32_t* __restrict__ counts, int32_t* __restrict__ status, int values) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pat) return;
    const int32_t st = conv_status[i];
    if (!st) return;
    if (status) status[i] = st;
    if (counts) counts[i] = (values && st == ST_CHAR_EXCEEDS_) ? conv_value[i] : 0;
}

}  // namespace fmgpu
