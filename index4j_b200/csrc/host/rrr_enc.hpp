// Host-side RRR ENCODER (index producer only; not the query path, not the oracle).
//
// Produces the four streams the reference's RrrVector serializes
// (indices/src/main/java/com/dynatrace/bitsequence/RrrVector.java:225-286 constructor,
// :430-440 write): 4-bit classes, variable-width offsets, sampled offset bit positions and
// sampled prefix sums, with samples every `sampleSize` BLOCKS of 15 bits.
//
// The (class, offset) <-> 15-bit block tables are not copied from the reference: they are
// generated from the ordering rule that reproduces the reference literal exactly (blocks grouped
// by popcount; inside a class ordered by DESCENDING value of the block read LSB-first as a
// 15-bit binary number).  tests/test_rrr_tables.py pins the generated tables by sha256 against
// the reference literal (RrrVector.java:488-8682 and :8705-16899).
#pragma once
#include <cstdint>
#include <vector>

#include "packed.hpp"

namespace fmhost {

struct RrrTables {
    uint16_t block_to_offset[32768];  // PRECOMPUTED_OFFSETS
    uint16_t inverse[32768];          // INVERSE_VALUES (index = class_base + offset)
    uint16_t class_base[16];          // CARDINALITY_OFFSETS
    int bits_needed[16];              // BITS_NEEDED_BINOMIAL_COEFFICIENTS
    RrrTables() {
        int cnt[16] = {0};
        for (int v = 0; v < 32768; ++v) cnt[__builtin_popcount(v)]++;
        int acc = 0;
        for (int k = 0; k < 16; ++k) {
            class_base[k] = (uint16_t)acc;
            bits_needed[k] = min_bits((uint64_t)cnt[k]);
            acc += cnt[k];
        }
        class_base[15] = 32767;
        int fill[16] = {0};
        for (int r = 32767; r >= 0; --r) {
            int v = 0;
            for (int b = 0; b < 15; ++b)
                if (r & (1 << b)) v |= 1 << (14 - b);
            int k = __builtin_popcount(v);
            block_to_offset[v] = (uint16_t)fill[k];
            inverse[class_base[k] + fill[k]] = (uint16_t)v;
            fill[k]++;
        }
    }
};

inline const RrrTables& rrr_tables() {
    static const RrrTables t;
    return t;
}

struct RrrEnc {
    int32_t sample_size = 0;
    int32_t length = 0;  // bits
    int32_t total_ones = 0;
    int32_t bits_per_offset_position = 1;
    IntVec classes;                  // width 4
    std::vector<uint64_t> offsets;   // VariableWidthIntVector words
    IntVec sampled_offset_pos;       // lengthOfSampledOffsets
    IntVec prefix_sums;

    size_t estimated_memory_usage() const {  // RrrVector.java:416-421
        return classes.size_in_bytes() + offsets.size() * 8 + sampled_offset_pos.size_in_bytes() +
               prefix_sums.size_in_bytes();
    }

    void encode(const BitString& bv, int sampleSize) {
        const RrrTables& T = rrr_tables();
        sample_size = sampleSize;
        length = (int32_t)bv.nbits;
        const int64_t len = (int64_t)bv.nbits;
        const int64_t nblocks = len / 15 + ((len % 15) ? 1 : 0);
        classes.init(nblocks, 4);
        uint64_t total_bits = 0;
        int64_t ones = 0;
        for (int64_t b = 0; b < nblocks; ++b) {
            int64_t lo = b * 15;
            int take = (int)((len - lo) < 15 ? (len - lo) : 15);
            uint32_t v = (uint32_t)bv.get_bits((uint64_t)lo, take);
            int k = __builtin_popcount(v);
            classes.set(b, (uint64_t)k);
            total_bits += (uint64_t)T.bits_needed[k];
            ones += k;
        }
        total_ones = (int32_t)ones;
        offsets.assign((total_bits + 63) / 64, 0);
        offsets.push_back(0);  // slack for straddling writes; dropped below
        bits_per_offset_position = min_bits(total_bits);
        sampled_offset_pos.init(nblocks / sampleSize + 1, bits_per_offset_position);
        prefix_sums.init(nblocks / sampleSize + 2, min_bits((uint64_t)ones));
        uint64_t cur_bits = 0;
        int64_t cur_sample = 0;
        uint64_t prefix = 0;
        for (int64_t b = 0; b < nblocks; ++b) {
            int64_t lo = b * 15;
            int take = (int)((len - lo) < 15 ? (len - lo) : 15);
            uint32_t v = (uint32_t)bv.get_bits((uint64_t)lo, take);
            int k = __builtin_popcount(v);
            int nb = T.bits_needed[k];
            uint64_t off = T.block_to_offset[v];
            int sh = (int)(cur_bits & 63);
            offsets[cur_bits >> 6] |= off << sh;
            if (sh + nb > 64) offsets[(cur_bits >> 6) + 1] |= off >> (64 - sh);
            if (b % sampleSize == 0) {
                sampled_offset_pos.set(cur_sample, cur_bits);
                prefix_sums.set(cur_sample, prefix);
                ++cur_sample;
            }
            cur_bits += (uint64_t)nb;
            prefix += (uint64_t)k;
        }
        prefix_sums.set(cur_sample, prefix);
        offsets.pop_back();
    }

    // getEstimatedMemoryUsage() of an all-zero vector of `len` bits, computed in closed form
    // (used by the block-size search, WaveletFixedBlockBoosting.java:960-965).
    static int64_t zero_vector_memory(int64_t len, int sampleSize) {
        int64_t nblocks = len / 15 + ((len % 15) ? 1 : 0);
        auto words = [](uint64_t bits) { return (int64_t)((bits + 63) / 64); };
        uint64_t total_bits = (uint64_t)nblocks;  // class 0 still costs one offset bit
        int64_t w = words((uint64_t)nblocks * 4) + words(total_bits) +
                    words((uint64_t)(nblocks / sampleSize + 1) * (uint64_t)min_bits(total_bits)) +
                    words((uint64_t)(nblocks / sampleSize + 2) * 1);
        return w * 8;
    }
};

}  // namespace fmhost
