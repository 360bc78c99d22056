// Host-side reader of the reference's serialized FmIndex (product code; independent of oracle/).
//
// Grammar = what FmIndex.write emits through java.io.ObjectOutput (big-endian primitives):
//   fm/FmIndex.java:948-975 (read :983-1025), intsequence/IntVector.java:196-227,
//   intsequence/VariableWidthIntVector.java:175-198, bitsequence/RrrVector.java:430-469,
//   wavelet/WaveletFixedBlockBoosting.java:286-322,1544-1570 (+ :1597-1613, :1630-1667),
//   serial-version check serialization/Serialization.java:46-56.
// The ObjectOutputStream framing written by Serialization.writeToByteArray (:67-78) — stream
// header AC ED 00 05, block-data records 0x77 <u8 len> / 0x7A <i32 len> — is stripped if present.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace fmgpu_host {

struct FormatError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

class JavaIn {
  public:
    JavaIn(const uint8_t* buf, size_t len) {
        if (len >= 4 && buf[0] == 0xAC && buf[1] == 0xED && buf[2] == 0x00 && buf[3] == 0x05) {
            owned_.reserve(len);
            size_t i = 4;
            while (i < len) {
                const uint8_t tag = buf[i++];
                size_t n = 0;
                if (tag == 0x77) {
                    if (i >= len) throw FormatError("object stream: truncated short block header");
                    n = buf[i++];
                } else if (tag == 0x7A) {
                    if (i + 4 > len) throw FormatError("object stream: truncated long block header");
                    n = ((size_t)buf[i] << 24) | ((size_t)buf[i + 1] << 16) | ((size_t)buf[i + 2] << 8) | (size_t)buf[i + 3];
                    i += 4;
                } else {
                    throw FormatError("object stream: only block-data records are expected");
                }
                if (n > len - i) throw FormatError("object stream: truncated block-data record");
                owned_.insert(owned_.end(), buf + i, buf + i + n);
                i += n;
            }
            p_ = owned_.data();
            end_ = p_ + owned_.size();
        } else {
            p_ = buf;
            end_ = buf + len;
        }
    }
    uint8_t u8() {
        need(1);
        return *p_++;
    }
    int16_t i16() {
        need(2);
        uint16_t v = (uint16_t)((p_[0] << 8) | p_[1]);
        p_ += 2;
        return (int16_t)v;
    }
    int32_t i32() {
        need(4);
        uint32_t v = ((uint32_t)p_[0] << 24) | ((uint32_t)p_[1] << 16) | ((uint32_t)p_[2] << 8) | (uint32_t)p_[3];
        p_ += 4;
        return (int32_t)v;
    }
    int64_t i64() {
        need(8);
        uint64_t v = 0;
        for (int k = 0; k < 8; ++k) v = (v << 8) | p_[k];
        p_ += 8;
        return (int64_t)v;
    }
    int32_t count32(const char* what) {  // an array length
        int32_t n = i32();
        if (n < 0) throw FormatError(std::string("negative length of ") + what);
        return n;
    }
    void raw(void* dst, size_t n) {
        need(n);
        memcpy(dst, p_, n);
        p_ += n;
    }
    void i64_array(uint64_t* dst, size_t n) {
        need(n * 8);
        for (size_t i = 0; i < n; ++i) {
            uint64_t v;
            memcpy(&v, p_ + 8 * i, 8);
            dst[i] = __builtin_bswap64(v);
        }
        p_ += n * 8;
    }
    void version() {  // Serialization.checkSerialVersion
        const uint8_t v = u8();
        if (v != 0) {
            char msg[96];
            snprintf(msg, sizeof msg, "Incompatible serial versions! Expected version 0 but was %d.", (int)v);
            throw FormatError(msg);
        }
    }
    size_t remaining() const { return (size_t)(end_ - p_); }

  private:
    void need(size_t n) const {
        if ((size_t)(end_ - p_) < n) throw FormatError("unexpected end of serialized index");
    }
    std::vector<uint8_t> owned_;
    const uint8_t* p_ = nullptr;
    const uint8_t* end_ = nullptr;
};

// Bit-packed vector as serialized (element k = bits [k*width, k*width+width), LSB-first words).
struct PackedInts {
    int32_t length = 0, width = 0;
    std::vector<uint64_t> words;
    void read(JavaIn& in) {
        in.version();
        length = in.count32("IntVector");
        width = in.i32();
        if (width < 0 || width > 64) throw FormatError("IntVector width out of range");
        const uint64_t bits = (uint64_t)length * (uint64_t)width;
        words.resize((size_t)((bits + 63) / 64) + 1, 0);  // +1 slack word for straddling reads
        in.i64_array(words.data(), words.size() - 1);
    }
    inline uint64_t field(uint64_t bitpos, int len) const {
        const size_t w = (size_t)(bitpos >> 6);
        const int sh = (int)(bitpos & 63);
        uint64_t v = words[w] >> sh;
        if (sh + len > 64) v |= words[w + 1] << (64 - sh);
        return len >= 64 ? v : (v & ((1ULL << len) - 1));
    }
    inline uint64_t get(int64_t k) const { return field((uint64_t)k * (uint64_t)width, width); }
};

struct RrrStream {
    int32_t sample_size = 0, length = 0, total_ones = 0, bits_per_offset_position = 0;
    PackedInts classes;
    std::vector<uint64_t> offsets;  // +1 slack word
    PackedInts sampled_offset_pos, prefix_sums;
    void read(JavaIn& in) {
        in.version();
        sample_size = in.i32();
        length = in.i32();
        total_ones = in.i32();
        bits_per_offset_position = in.i32();
        if (sample_size < 1 || length < 0) throw FormatError("RrrVector header out of range");
        classes.read(in);
        in.version();
        const int32_t nw = in.count32("VariableWidthIntVector");
        offsets.resize((size_t)nw + 1, 0);
        in.i64_array(offsets.data(), (size_t)nw);
        sampled_offset_pos.read(in);
        prefix_sums.read(in);
        const int64_t nblocks = ((int64_t)length + 14) / 15;
        if (classes.width != 4 || classes.length < nblocks) throw FormatError("RrrVector class stream too short");
    }
};

struct BlockHdr {
    int32_t bv_rank, bv_offset, var_off;
    int16_t sigma_m1, tree_height;
};

struct SuperBlockHdr {
    int16_t sigma_m1 = 0, block_size_log = 0;
    RrrStream rank_support;
    std::vector<BlockHdr> blocks;
    std::vector<uint8_t> var;
    std::vector<int16_t> mapping;
};

struct WfbbStream {
    int64_t size = 0;
    int32_t sigma = 0, rrr_rate = 0;
    std::vector<int64_t> count, hyper_rank;
    std::vector<int32_t> sb_rank;
    std::vector<int16_t> global_mapping;
    std::vector<SuperBlockHdr> sbs;
    void read(JavaIn& in) {
        in.version();
        size = in.i64();
        sigma = in.i32();
        rrr_rate = in.i32();
        if (size < 0 || size > 0x7fffffffLL || sigma < 1 || sigma > 32768) throw FormatError("wavelet header out of range");
        count.resize((size_t)in.count32("count"));
        for (auto& v : count) v = in.i64();
        hyper_rank.resize((size_t)in.count32("hyperBlockRank"));
        for (auto& v : hyper_rank) v = in.i64();
        sb_rank.resize((size_t)in.count32("superBlockRank"));
        for (auto& v : sb_rank) v = in.i32();
        global_mapping.resize((size_t)in.count32("globalMapping"));
        for (auto& v : global_mapping) v = in.i16();
        sbs.resize((size_t)in.count32("superBlockHeaderItems"));
        for (auto& s : sbs) {
            s.sigma_m1 = in.i16();
            s.block_size_log = in.i16();
            if (s.block_size_log < 1 || s.block_size_log > 20) throw FormatError("blockSizeLog out of range");
            s.rank_support.read(in);
            s.blocks.resize((size_t)in.count32("blockHeaders"));
            for (auto& b : s.blocks) {
                b.bv_rank = in.i32();
                b.bv_offset = in.i32();
                b.var_off = in.i32();
                b.sigma_m1 = in.i16();
                b.tree_height = in.i16();
            }
            s.var.resize((size_t)in.count32("varBlockHeadersData"));
            if (!s.var.empty()) in.raw(s.var.data(), s.var.size());
            s.mapping.resize((size_t)in.count32("mapping"));
            for (auto& m : s.mapping) m = in.i16();
        }
        const int64_t nsb = (size + (1LL << 20) - 1) >> 20;
        if ((int64_t)sbs.size() != nsb) throw FormatError("superblock count does not match size");
        if (sb_rank.size() != (size_t)nsb * (size_t)sigma || global_mapping.size() != (size_t)nsb * (size_t)sigma ||
            count.size() != (size_t)sigma || hyper_rank.size() < (size_t)sigma)
            throw FormatError("wavelet directory arrays have unexpected sizes");
    }
};

struct FmStream {
    int32_t sample_rate = 0;
    bool enable_extract = false;
    int32_t bw_suffixes = 0, bw_positions = 0, length = 0;
    std::vector<std::pair<int32_t, int16_t>> map;  // monotonicMap entries in stream order
    std::vector<int32_t> C, lookup;
    PackedInts suffixes, positions;
    RrrStream sampled;
    WfbbStream wf;
    void read(JavaIn& in) {
        in.version();
        sample_rate = in.i32();
        enable_extract = in.u8() != 0;
        bw_suffixes = in.i32();
        bw_positions = in.i32();
        length = in.i32();
        if (sample_rate < 1 || length < 1) throw FormatError("FmIndex header out of range");
        map.resize((size_t)in.count32("monotonicMap"));
        for (auto& kv : map) {
            kv.first = in.i32();
            kv.second = in.i16();
        }
        C.resize((size_t)in.count32("cumulativeCounts"));
        for (auto& v : C) v = in.i32();
        lookup.resize((size_t)in.count32("monotonicLookUp"));
        for (auto& v : lookup) v = in.i32();
        suffixes.read(in);
        if (enable_extract) positions.read(in);
        sampled.read(in);
        wf.read(in);
        if (wf.size != (int64_t)length) throw FormatError("wavelet size differs from index length");
        if (sampled.length != length) throw FormatError("sampled-row bitvector length differs from index length");
    }
};

}  // namespace fmgpu_host
