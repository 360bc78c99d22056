// Host-side (index PRODUCER) bit-packed containers.
//
// These follow the storage conventions of the reference so that what we serialize is readable by
// the reference's own FmIndex.read():
//   IntVector              indices/src/main/java/com/dynatrace/intsequence/IntVector.java:45-143
//   VariableWidthIntVector indices/src/main/java/com/dynatrace/intsequence/VariableWidthIntVector.java:41-140
//   minimumNumberOfBits    indices/src/main/java/com/dynatrace/intsequence/Common.java:169-175
// Element k of a width-w vector occupies bits [k*w, k*w+w) of the LSB-first bit string made of
// 64-bit words.  This code is only used to BUILD indexes (tests, bench); it is not on the GPU
// query path and it is not the oracle.
#pragma once
#include <cstdint>
#include <vector>

namespace fmhost {

inline int min_bits(uint64_t v) { return v == 0 ? 1 : 64 - __builtin_clzll(v); }

inline uint64_t low_mask(int bits) { return bits >= 64 ? ~0ULL : ((1ULL << bits) - 1); }

// Plain LSB-first bit string with append/set/get of short fields.
struct BitString {
    std::vector<uint64_t> w;
    uint64_t nbits = 0;
    void resize(uint64_t n) {
        nbits = n;
        w.assign((n + 63) / 64 + 1, 0);  // +1 slack word so straddling reads never fault
    }
    inline void set1(uint64_t i) { w[i >> 6] |= 1ULL << (i & 63); }
    // n bits from 32-bit words (bit k of the string = bit k % 32 of word k / 32); bits past n are dropped
    void from_words32(const uint32_t* words, uint64_t n) {
        resize(n);
        const uint64_t nw32 = (n + 31) / 32;
        for (uint64_t k = 0; k < nw32; ++k) w[k >> 1] |= (uint64_t)words[k] << (32 * (k & 1));
        if (n & 63) w[n >> 6] &= (1ULL << (n & 63)) - 1;
    }
    inline bool get(uint64_t i) const { return (w[i >> 6] >> (i & 63)) & 1; }
    // read `len` (<= 57) bits starting at bit i
    inline uint64_t get_bits(uint64_t i, int len) const {
        uint64_t lo = w[i >> 6] >> (i & 63);
        int sh = (int)(i & 63);
        if (sh + len > 64) lo |= w[(i >> 6) + 1] << (64 - sh);
        return lo & low_mask(len);
    }
    inline void put_bits(uint64_t i, uint64_t v, int len) {
        int sh = (int)(i & 63);
        w[i >> 6] |= v << sh;
        if (sh + len > 64) w[(i >> 6) + 1] |= v >> (64 - sh);
    }
};

struct IntVec {
    std::vector<uint64_t> data;
    int32_t length = 0;
    int32_t width = 1;
    IntVec() {}
    IntVec(int64_t len, int w) { init(len, w); }
    void init(int64_t len, int w) {
        length = (int32_t)len;
        width = w;
        uint64_t bits = (uint64_t)len * (uint64_t)w;
        data.assign((bits + 63) / 64, 0);  // IntVector.java:45-54
    }
    inline void set(int64_t pos, uint64_t v) {
        uint64_t bp = (uint64_t)pos * width;
        v &= low_mask(width);
        int sh = (int)(bp & 63);
        size_t wi = bp >> 6;
        data[wi] = (data[wi] & ~(low_mask(width) << sh)) | (v << sh);
        if (sh + width > 64) {
            int spill = sh + width - 64;
            data[wi + 1] = (data[wi + 1] & ~low_mask(spill)) | (v >> (width - spill));
        }
    }
    inline uint64_t get(int64_t pos) const {
        uint64_t bp = (uint64_t)pos * width;
        int sh = (int)(bp & 63);
        size_t wi = bp >> 6;
        uint64_t lo = data[wi] >> sh;
        if (sh + width > 64) lo |= data[wi + 1] << (64 - sh);
        return lo & low_mask(width);
    }
    size_t size_in_bytes() const { return data.size() * 8; }
};

}  // namespace fmhost
