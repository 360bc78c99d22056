// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// A CPU restatement of the reference's FM-index query path (dynatrace-oss/index4j, Java) used as
// the parity checker for the CUDA engine and as the CPU baseline ("port") in bench.py.  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// this library.  The product (index4j_b200/, libfmgpu) never links, imports or calls it.
//
// Parity status: PINNED against the reference's own known answers and golden strings
// (tests/test_oracle_golden.py: WaveletFixedBlockBoostingTest.java:57-69,86-132;
// RrrVectorTest.java:70-122; FmIndexTest.java:195-200,376-400,430-496,564-578) and against naive
// text scans in the style of the reference's test oracle (test/.../util/Util.java:108-279).
// The reference itself cannot be executed here (no JVM in the image), so there is no
// oracle/_ref build; the (class,offset) tables are pinned by sha256 against the Java literal.
//
// Every function names the reference lines it follows.  Path shorthand:
//   FM   = indices/src/main/java/com/dynatrace/fm/FmIndex.java
//   WF   = indices/src/main/java/com/dynatrace/wavelet/WaveletFixedBlockBoosting.java
//   RRR  = indices/src/main/java/com/dynatrace/bitsequence/RrrVector.java
//   IV   = indices/src/main/java/com/dynatrace/intsequence/IntVector.java
//   VW   = indices/src/main/java/com/dynatrace/intsequence/VariableWidthIntVector.java
//   SER  = indices/src/main/java/com/dynatrace/serialization/Serialization.java
// The data layout deliberately mirrors the Java object graph (one object per superblock, separate
// arrays per field) so that the CPU baseline has the reference's memory behaviour.
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

thread_local std::string g_err;

struct JavaThrow {  // stands for an unchecked Java exception
    int status;
    int n;
};
enum {
    ST_OK = 0,
    ST_NOT_ENABLED = 1,      // RuntimeException("Text recovery not enabled at build time")      FM:566,611
    ST_POS_NEGATIVE = 2,     // RuntimeException("Requested position less than 0")               FM:570,615
    ST_STOP_TOO_LONG = 3,    // RuntimeException("Stop position longer than index string")       FM:574
    ST_POS_TOO_LONG = 4,     // RuntimeException("Requested position longer than index string")  FM:619
    ST_DST_TOO_SMALL = 5,    // RuntimeException("Supplied destination is not large enough")     FM:591
    ST_DST_ZERO = 6,         // IllegalArgumentException("Supplied destination for extraction has size zero") FM:623
    ST_NO_BOUNDARY = 7,      // IllegalArgumentException("Boundary does not exist")              FM:659,792,849
    ST_DOES_NOT_FIT = 8,     // RuntimeException("Extraction does not fit ... Currently extracted: N") FM:733,817,894
    ST_INDEX_OOB = 9,        // ArrayIndexOutOfBoundsException (e.g. quirk Q4, WF:1022-1026)
    ST_CHAR_EXCEEDS = 10,    // RuntimeException("Found a character that exceeds (32767): it was N")  FM:262-267
    ST_NO_TERMINATION = 12,  // the reference never returns: an LF walk of locate (FM:531-537) longer than `length` steps has
                             // visited a row twice, i.e. runs in a cycle (possible only behind quirk Q1, WF:1329-1332)
};

// ---------------------------------------------------------------------------------------------
// Stream reader: java.io.DataInput semantics (big-endian).  Accepts the ObjectOutputStream framing
// produced by SER:67-78 (magic AC ED 00 05 + block-data records 0x77/0x7A) or the raw primitives.
// ---------------------------------------------------------------------------------------------
struct Reader {
    std::vector<uint8_t> data;
    size_t p = 0;
    Reader(const uint8_t* buf, size_t len) {
        if (len >= 4 && buf[0] == 0xAC && buf[1] == 0xED && buf[2] == 0x00 && buf[3] == 0x05) {
            size_t i = 4;
            data.reserve(len);
            while (i < len) {
                uint8_t tag = buf[i++];
                size_t n;
                if (tag == 0x77) {
                    if (i + 1 > len) throw std::runtime_error("truncated block-data header");
                    n = buf[i++];
                } else if (tag == 0x7A) {
                    if (i + 4 > len) throw std::runtime_error("truncated block-data header");
                    n = ((size_t)buf[i] << 24) | ((size_t)buf[i + 1] << 16) | ((size_t)buf[i + 2] << 8) | buf[i + 3];
                    i += 4;
                } else {
                    throw std::runtime_error("unexpected tag in object stream");
                }
                if (i + n > len) throw std::runtime_error("truncated block-data record");
                data.insert(data.end(), buf + i, buf + i + n);
                i += n;
            }
        } else {
            data.assign(buf, buf + len);
        }
    }
    void need(size_t n) {
        if (p + n > data.size()) throw std::runtime_error("unexpected end of stream");
    }
    uint8_t u8() {
        need(1);
        return data[p++];
    }
    int16_t i16() {
        need(2);
        int16_t v = (int16_t)((data[p] << 8) | data[p + 1]);
        p += 2;
        return v;
    }
    int32_t i32() {
        need(4);
        uint32_t v = ((uint32_t)data[p] << 24) | ((uint32_t)data[p + 1] << 16) | ((uint32_t)data[p + 2] << 8) | data[p + 3];
        p += 4;
        return (int32_t)v;
    }
    int64_t i64() {
        need(8);
        uint64_t v = 0;
        for (int k = 0; k < 8; ++k) v = (v << 8) | data[p + k];
        p += 8;
        return (int64_t)v;
    }
    void version() {  // SER:46-56
        uint8_t v = u8();
        if (v != 0) {
            char msg[96];
            snprintf(msg, sizeof msg, "Incompatible serial versions! Expected version 0 but was %d.", (int)v);
            throw std::runtime_error(msg);
        }
    }
};

inline uint64_t low_bits(int n) { return n >= 64 ? ~0ULL : ((1ULL << n) - 1); }  // Common.LOW_BITS_SET

// IV:129-143 / VW:127-140 — identical arithmetic on an arbitrary bit position.
inline uint64_t get_bits(const std::vector<uint64_t>& d, uint64_t bitpos, int len) {
    size_t wi = (size_t)(bitpos >> 6);
    int off = (int)(bitpos & 63);
    uint64_t left = d.at(wi) >> off;
    if (off + len > 64) {
        uint64_t right = (d.at(wi + 1) & low_bits((off + len) & 63)) << (64 - off);
        return left | right;
    }
    return left & low_bits(len);
}

struct IntVector {
    std::vector<uint64_t> data;
    int32_t length = 0, width = 0;
    void read(Reader& r) {  // IV:211-227
        r.version();
        length = r.i32();
        width = r.i32();
        uint64_t bits = (uint64_t)length * (uint64_t)width;
        size_t words = (size_t)(bits % 64 == 0 ? bits / 64 : bits / 64 + 1);
        data.resize(words);
        for (size_t i = 0; i < words; ++i) data[i] = (uint64_t)r.i64();
    }
    inline uint64_t get(int64_t pos, int len) const { return get_bits(data, (uint64_t)pos * (uint64_t)width, len); }
};

// (class, offset) -> 15-bit block.  Generated, not copied: blocks grouped by popcount, inside a
// class ordered by descending value of the block read LSB-first (RRR:8692-8698, :8705-16899;
// pinned by sha256 in tests/test_rrr_tables.py).
struct Tables {
    uint16_t inverse[32768];
    uint16_t card_off[16];
    int bits_needed[16];  // RRR:111-129
    Tables() {
        int cnt[16] = {0};
        for (int v = 0; v < 32768; ++v) cnt[__builtin_popcount(v)]++;
        int acc = 0;
        for (int k = 0; k < 16; ++k) {
            card_off[k] = (uint16_t)acc;
            int b = 0;
            while ((1 << b) <= cnt[k]) ++b;  // minimumNumberOfBits(C(15,k))
            bits_needed[k] = b;
            acc += cnt[k];
        }
        int fill[16] = {0};
        for (int r = 32767; r >= 0; --r) {
            int v = 0;
            for (int b = 0; b < 15; ++b)
                if (r & (1 << b)) v |= 1 << (14 - b);
            int k = __builtin_popcount(v);
            inverse[card_off[k] + fill[k]++] = (uint16_t)v;
        }
    }
};
const Tables& tables() {
    static const Tables t;
    return t;
}

struct RrrVector {
    int32_t sample_size = 0, length = 0, total_ones = 0, bits_per_offset_position = 0;
    IntVector classes;
    std::vector<uint64_t> offsets;
    IntVector sampled_offset_pos, prefix_sums;
    void read(Reader& r) {  // RRR:448-469
        r.version();
        sample_size = r.i32();
        length = r.i32();
        total_ones = r.i32();
        bits_per_offset_position = r.i32();
        classes.read(r);
        r.version();  // VW:189-198
        int32_t nw = r.i32();
        offsets.resize((size_t)nw);
        for (int32_t i = 0; i < nw; ++i) offsets[(size_t)i] = (uint64_t)r.i64();
        sampled_offset_pos.read(r);
        prefix_sums.read(r);
    }
    // RRR:358-396
    int32_t rank_ones(int32_t position) const {
        if (position < 0) return 0;
        if (position >= length) return total_ones;
        const Tables& T = tables();
        int32_t block_id = position / 15;
        int32_t sampled = block_id / sample_size;
        int32_t prefix = (int32_t)prefix_sums.get(sampled, prefix_sums.width);
        int32_t cur = (int32_t)sampled_offset_pos.get(sampled, bits_per_offset_position);
        int32_t i;
        for (i = sampled * sample_size; i < position / 15; ++i) {
            int c = (int)classes.get(i, 4);
            prefix += c;
            cur += T.bits_needed[c];
        }
        int c = (int)classes.get(i, 4);
        int nb = T.bits_needed[c];
        uint64_t off = get_bits(offsets, (uint64_t)cur, nb);
        uint32_t block = T.inverse[(uint32_t)T.card_off[c] + (uint32_t)off];
        int use = position - i * 15;
        return prefix + __builtin_popcount(block & (uint32_t)low_bits(use));
    }
    // RRR:314-349 (throws IllegalArgumentException outside [0,length))
    bool access(int32_t position) const {
        if (position < 0 || position >= length) throw JavaThrow{ST_INDEX_OOB, position};
        const Tables& T = tables();
        int32_t block_id = position / 15;
        int32_t sampled = block_id / sample_size;
        int64_t cur = (int64_t)sampled_offset_pos.get(sampled, bits_per_offset_position);
        int32_t i;
        for (i = sampled * sample_size; i < position / 15; ++i) {
            int c = (int)classes.get(i, 4);
            cur += T.bits_needed[c];
        }
        int c = (int)classes.get(i, 4);
        int nb = T.bits_needed[c];
        uint64_t off = get_bits(offsets, (uint64_t)cur, nb);
        uint32_t block = T.inverse[(uint32_t)T.card_off[c] + (uint32_t)off];
        return ((block >> (position % 15)) & 1u) == 1u;
    }
};

struct BlockHeaderItem {  // WF:1589-1605
    int32_t bv_rank, bv_offset, var_off;
    int16_t sigma, tree_height;
};
struct SuperBlockHeaderItem {  // WF:1621-1649
    int16_t sigma = 0, block_size_log = 0;
    RrrVector rank_support;
    std::vector<BlockHeaderItem> block_headers;
    std::vector<int8_t> var;  // Java byte[] (signed)
    std::vector<int16_t> mapping;
};

inline int64_t u16le(const std::vector<int8_t>& v, int64_t p) {  // ((b1 << 8) & 0xff00 | b0 & 0xff)
    return (((int32_t)v.at((size_t)p + 1) << 8) & 0x00ff00) | ((int32_t)v.at((size_t)p) & 0x0000ff);
}
inline int64_t u24le(const std::vector<int8_t>& v, int64_t p) {
    return (((int32_t)v.at((size_t)p + 2) << 16) & 0xff0000) | (((int32_t)v.at((size_t)p + 1) << 8) & 0x00ff00) |
           ((int32_t)v.at((size_t)p) & 0x0000ff);
}

// Work counters (instrumentation for the roofline accounting and the tests that cross-check the GPU's own counters; not part
// of the reference).  They must not slow the timed CPU arm down: every thread counts into its own thread-local tally, which
// is added to the process-wide totals once, when the thread ends (the batch calls join their workers) or when orc_fm_stats
// reads them — no shared cache line is touched inside rank / inverseSelect.  -DORACLE_NO_STATS (the build bench.py times as the
// CPU baseline) compiles the counters out altogether.
#ifndef ORACLE_NO_STATS
std::atomic<uint64_t> g_totals[4];  // ranks, rank_levels, lf_steps, lf_levels
struct Tally {
    uint64_t v[4] = {0, 0, 0, 0};
    void flush() {
        for (int i = 0; i < 4; ++i)
            if (v[i]) {
                g_totals[i].fetch_add(v[i], std::memory_order_relaxed);
                v[i] = 0;
            }
    }
    ~Tally() { flush(); }
};
thread_local Tally t_tally;
#define ORC_COUNT(slot, n) (t_tally.v[slot] += (uint64_t)(n))
#else
#define ORC_COUNT(slot, n) ((void)0)
#endif

struct Wfbb {
    int64_t size = 0;
    int32_t sigma = 0, rrr_rate = 0;
    std::vector<int64_t> count, hyper_rank;
    std::vector<int32_t> sb_rank;
    std::vector<int16_t> global_mapping;
    std::vector<SuperBlockHeaderItem> sbs;

    void read(Reader& r) {  // WF:286-322
        r.version();
        size = r.i64();
        sigma = r.i32();
        rrr_rate = r.i32();
        count.resize((size_t)r.i32());
        for (auto& v : count) v = r.i64();
        hyper_rank.resize((size_t)r.i32());
        for (auto& v : hyper_rank) v = r.i64();
        sb_rank.resize((size_t)r.i32());
        for (auto& v : sb_rank) v = r.i32();
        global_mapping.resize((size_t)r.i32());
        for (auto& v : global_mapping) v = r.i16();
        sbs.resize((size_t)r.i32());
        for (auto& s : sbs) {  // WF:1630-1649
            s.sigma = r.i16();
            s.block_size_log = r.i16();
            s.rank_support.read(r);
            s.block_headers.resize((size_t)r.i32());
            for (auto& b : s.block_headers) {
                b.bv_rank = r.i32();
                b.bv_offset = r.i32();
                b.var_off = r.i32();
                b.sigma = r.i16();
                b.tree_height = r.i16();
            }
            s.var.resize((size_t)r.i32());
            for (auto& b : s.var) b = (int8_t)r.u8();
            s.mapping.resize((size_t)r.i32());
            for (auto& m : s.mapping) m = r.i16();
        }
    }

    // WF:250-278
    static int64_t restore_code(int64_t block_c, const std::vector<int8_t>& hdr, int64_t pos, int64_t tree_height) {
        int32_t code = 0, code_length = 1;
        int64_t leaf_count = 0;
        while (code_length < tree_height) {
            code <<= 1;
            int64_t level_leaf = u16le(hdr, pos);
            if (leaf_count + level_leaf > block_c) {
                code += (int32_t)(block_c - leaf_count);
                break;
            } else {
                code += (int32_t)level_leaf;
                ++code_length;
                leaf_count += level_leaf;
                pos += 4;
            }
        }
        if (code_length == tree_height) {
            code <<= 1;
            code += (int32_t)(block_c - leaf_count);
        }
        return ((int64_t)code << 32) | (int64_t)code_length;
    }
    // WF:232-248
    static int64_t symbol_from_header(const std::vector<int8_t>& hdr, int64_t pos, int64_t code, int64_t code_length) {
        int64_t block_c = 0, temp = 0;
        for (int64_t i = 1; i < code_length; ++i) {
            int64_t level_leaf = u16le(hdr, pos);
            pos += 4;
            temp += level_leaf;
            block_c += level_leaf;
            temp <<= 1;
        }
        block_c += code - temp;
        return block_c;
    }

    // WF:1010-1285.  `symbol` is a Java short.
    int64_t rank(int64_t position, int16_t symbol) const {
        if (position == 0) return 0;
        if (position > size) position = size;
        if (symbol >= sigma) return 0;
        ORC_COUNT(0, 1);
        if (symbol < 0) throw JavaThrow{ST_INDEX_OOB, symbol};
        int64_t hb = position >> 32;
        int64_t sb = position >> 20;
        if ((size_t)(sb * sigma + symbol) >= global_mapping.size()) throw JavaThrow{ST_INDEX_OOB, (int)sb};  // Q4
        int16_t sb_c = global_mapping[(size_t)(sb * sigma + symbol)];
        int64_t sb_index = position & ((1LL << 20) - 1);
        const SuperBlockHeaderItem& S = sbs.at((size_t)sb);
        int64_t sb_sigma = (int64_t)S.sigma + 1;
        int64_t bl = S.block_size_log;
        int64_t bs = 1LL << bl;
        int64_t blocks_log = 20 - bl;
        int64_t block_index = position & (bs - 1);
        int64_t cur_block_size = std::min<int64_t>(bs, size - (position - block_index));
        int64_t block_id = sb_index >> bl;
        int64_t rank_sb = sb_rank[(size_t)(sb * sigma + symbol)];
        int64_t rank_hb = hyper_rank.at((size_t)(hb * sigma + symbol));
        if (sb_c >= sb_sigma) return rank_hb + rank_sb;  // WF:1040
        int16_t block_c = S.mapping.at((size_t)(((int64_t)sb_c << blocks_log) + block_id));
        if (block_c == sigma - 1) {  // WF:1048: absent from this block
            ++block_id;
            int64_t blocks_in_sb = 1LL << blocks_log;
            while (block_id < blocks_in_sb && S.mapping.at((size_t)(((int64_t)sb_c << blocks_log) + block_id)) == sigma - 1) ++block_id;
            if (block_id == blocks_in_sb) {
                if ((sb + 1) * (1LL << 20) >= size) return count.at((size_t)symbol);
                return rank_hb + sb_rank.at((size_t)((sb + 1) * sigma + symbol));
            }
            block_c = S.mapping.at((size_t)(((int64_t)sb_c << blocks_log) + block_id));
            const BlockHeaderItem& H = S.block_headers.at((size_t)block_id);
            int64_t ptr = (int64_t)H.var_off + ((int64_t)H.tree_height - 1) * 4;  // WF:1081 (no height>0 guard, no clamp repair: quirk Q3)
            return rank_hb + rank_sb + u24le(S.var, ptr + (int64_t)block_c * 5 + 2);
        }
        const BlockHeaderItem& H = S.block_headers.at((size_t)block_id);
        int64_t var_off = H.var_off;
        int64_t tree_height = H.tree_height;
        int64_t vptr = var_off;
        int64_t tmp = vptr;
        if (tree_height > 0) tmp += (tree_height - 1) * 4;
        int64_t value = u16le(S.var, tmp + 5 * (int64_t)block_c);
        if (value != symbol) ++block_c;  // WF:1128 clamp repair
        int64_t rank_blk = u24le(S.var, tmp + (int64_t)block_c * 5 + 2);
        if (tree_height == 0) return rank_hb + rank_sb + rank_blk + block_index;  // WF:1141
        int64_t code_result = restore_code(block_c, S.var, var_off, tree_height);
        int32_t code = (int32_t)((uint64_t)code_result >> 32);
        int32_t code_length = (int32_t)code_result;
        int64_t bv_rank = H.bv_rank;
        int64_t bv_offset = H.bv_offset;
        int64_t internal_nodes = 1, left_sib = 0, left_sib_bv = 0;
        int64_t node_bv_size = cur_block_size;
        int64_t depth_total = node_bv_size;
        int64_t node_rank = block_index;
        int64_t block_sigma = (int64_t)H.sigma + 1;
        int64_t second = var_off + (tree_height - 1) * 4 + block_sigma * 5;
        ORC_COUNT(1, code_length);
        for (int64_t depth = 0; depth < code_length; ++depth) {
            int64_t rank1 = S.rank_support.rank_ones((int32_t)(bv_offset + left_sib_bv + node_rank));
            int64_t left_ones = 0;
            if (left_sib > 0) left_ones = u16le(S.var, second + 2 * (left_sib - 1));
            rank1 -= bv_rank + left_ones;
            int64_t node_ones = u16le(S.var, second + 2 * left_sib) - left_ones;
            int64_t node_zeros = node_bv_size - node_ones;
            int64_t rank0 = node_rank - rank1;
            bv_rank += u16le(S.var, second + 2 * (internal_nodes - 1));
            second += 2 * internal_nodes;
            left_sib <<= 1;
            int64_t next_bit = (int64_t)code & (1LL << (code_length - depth - 1));
            if (next_bit != 0) {
                node_rank = rank1;
                node_bv_size = node_ones;
                ++left_sib;
                left_sib_bv += node_zeros;
            } else {
                node_rank = rank0;
                node_bv_size = node_zeros;
            }
            if (depth + 1 != code_length) {
                int64_t next_leaf = u16le(S.var, vptr);
                vptr += 2;
                int64_t next_total = u16le(S.var, vptr) + 1;
                vptr += 2;
                left_sib_bv -= (depth_total - next_total);
                bv_offset += depth_total;
                depth_total = next_total;
                internal_nodes <<= 1;
                internal_nodes -= next_leaf;
                left_sib -= next_leaf;
            }
        }
        return rank_hb + rank_sb + rank_blk + node_rank;
    }

    // WF:1305-1537
    int64_t inverse_select(int64_t position) const {
        if (position < 0) throw JavaThrow{ST_INDEX_OOB, (int)position};
        int64_t hb = position >> 32;
        int64_t sb = position >> 20;
        int64_t sb_index = position & ((1LL << 20) - 1);
        if ((size_t)sb >= sbs.size()) throw JavaThrow{ST_INDEX_OOB, (int)sb};
        const SuperBlockHeaderItem& S = sbs[(size_t)sb];
        int64_t bl = S.block_size_log;
        int64_t bs = 1LL << bl;
        int64_t block_index = position & (bs - 1);
        int64_t cur_block_size = std::min<int64_t>(bs, size - (position - block_index));
        int64_t block_id = sb_index >> bl;
        if ((size_t)block_id >= S.block_headers.size()) throw JavaThrow{ST_INDEX_OOB, (int)block_id};
        const BlockHeaderItem& H = S.block_headers[(size_t)block_id];
        int64_t var_off = H.var_off;
        int64_t tree_height = H.tree_height;
        int64_t ptr8 = var_off;
        const int64_t copy_ptr8 = ptr8;
        int64_t tmp8 = ptr8;
        if (tree_height > 0) tmp8 += (tree_height - 1) * 4;
        const int64_t ptr32 = tmp8;
        ORC_COUNT(2, 1);
        if (tree_height == 0) {
            int8_t b0 = S.var.at((size_t)ptr32);
            int8_t b1 = S.var.at((size_t)ptr32 + 1);
            // WF:1332 — only the low byte survives (quirk Q1)
            int32_t c = (((((int32_t)(int16_t)b1) << 8) & 0x00ff00) | ((int32_t)(int16_t)b0)) & 0x00ff;
            if (position == 0) return c;
            int64_t rank_blk = u24le(S.var, ptr32 + 2);
            int64_t rank_sb = sb_rank.at((size_t)(sb * sigma + c));
            int64_t rank_hb = hyper_rank.at((size_t)(hb * sigma + c));
            int64_t result = rank_hb + rank_sb + rank_blk + block_index;
            return (int64_t)(((uint64_t)result << 32) | (uint64_t)c);
        }
        int64_t code = 0, code_length = 0;
        int64_t bv_rank = H.bv_rank;
        int64_t bv_offset = H.bv_offset;
        int64_t internal_nodes = 1, left_sib = 0, left_sib_bv = 0;
        int64_t node_bv_size = cur_block_size;
        int64_t depth_total = node_bv_size;
        int64_t node_rank = block_index;
        int64_t block_sigma = (int64_t)H.sigma + 1;
        int64_t second = var_off + (tree_height - 1) * 4 + block_sigma * 5;
        for (int64_t depth = 0;; ++depth) {
            int32_t rank_pos = (int32_t)(bv_offset + left_sib_bv + node_rank);
            int64_t rank1 = S.rank_support.rank_ones(rank_pos);
            bool next_bit = S.rank_support.access(rank_pos);
            int64_t left_ones = 0;
            if (left_sib > 0) left_ones = u16le(S.var, second + 2 * (left_sib - 1));
            rank1 -= bv_rank + left_ones;
            int64_t node_ones = u16le(S.var, second + 2 * left_sib) - left_ones;
            int64_t node_zeros = node_bv_size - node_ones;
            int64_t rank0 = node_rank - rank1;
            bv_rank += u16le(S.var, second + 2 * (internal_nodes - 1));
            second += internal_nodes * 2;
            left_sib <<= 1;
            code <<= 1;
            ++code_length;
            if (next_bit) {
                code |= 1;
                node_rank = rank1;
                node_bv_size = node_ones;
                ++left_sib;
                left_sib_bv += node_zeros;
            } else {
                node_rank = rank0;
                node_bv_size = node_zeros;
            }
            if (depth + 1 < tree_height) {
                int64_t next_leaf = u16le(S.var, ptr8);
                ptr8 += 2;
                int64_t next_total = u16le(S.var, ptr8) + 1;
                ptr8 += 2;
                left_sib_bv -= (depth_total - next_total);
                bv_offset += depth_total;
                depth_total = next_total;
                internal_nodes <<= 1;
                internal_nodes -= next_leaf;
                if (left_sib >= next_leaf) left_sib -= next_leaf;
                else break;
            } else {
                break;
            }
        }
        ORC_COUNT(3, code_length);
        int64_t block_c = symbol_from_header(S.var, copy_ptr8, code, code_length);
        int32_t c = (int32_t)u16le(S.var, ptr32 + 5 * block_c);
        if (position == 0) return c;
        int64_t rank_blk = u24le(S.var, ptr32 + block_c * 5 + 2);
        int64_t rank_sb = sb_rank.at((size_t)(sb * sigma + c));
        int64_t rank_hb = hyper_rank.at((size_t)(hb * sigma + c));
        int64_t result = rank_hb + rank_sb + rank_blk + node_rank;
        return (int64_t)(((uint64_t)result << 32) | (uint64_t)c);
    }
};

struct FmIndex {
    int32_t sample_rate = 0;
    bool enable_extract = false;
    int32_t bw_suffixes = 0, bw_positions = 0, length = 0;
    std::unordered_map<int32_t, int16_t> map;  // monotonicMap (HashMap<Integer,Short>)
    std::vector<int32_t> C, lookup;
    IntVector suffixes, positions;
    RrrVector sampled;
    Wfbb wf;

    void read(Reader& r) {  // FM:983-1025
        r.version();
        sample_rate = r.i32();
        enable_extract = r.u8() != 0;
        bw_suffixes = r.i32();
        bw_positions = r.i32();
        length = r.i32();
        int32_t nk = r.i32();
        for (int32_t i = 0; i < nk; ++i) {
            int32_t k = r.i32();
            int16_t v = r.i16();
            map[k] = v;
        }
        C.resize((size_t)r.i32());
        for (auto& v : C) v = r.i32();
        lookup.resize((size_t)r.i32());
        for (auto& v : lookup) v = r.i32();
        suffixes.read(r);
        if (enable_extract) positions.read(r);
        sampled.read(r);
        wf.read(r);
    }
    inline int16_t code(uint16_t ch) const {  // monotonicMap.getOrDefault((int) ch, (short) 0)
        auto it = map.find((int32_t)ch);
        return it == map.end() ? (int16_t)0 : it->second;
    }

    // FM:455-474
    int32_t count(const uint16_t* pattern, int32_t offset, int32_t len) const {
        int32_t i = (offset + len) - 1;
        if (i < 0) throw JavaThrow{ST_INDEX_OOB, i};  // pattern[-1] (quirk Q7)
        int16_t c = code(pattern[i]);
        if (c == 0) return 0;
        int32_t start = C.at((size_t)c), end = C.at((size_t)c + 1);
        while (start < end && i >= offset + 1) {
            c = code(pattern[--i]);
            if (c == 0) return 0;
            start = (int32_t)(C[(size_t)c] + wf.rank(start, c));
            end = (int32_t)(C[(size_t)c] + wf.rank(end, c));
        }
        return std::max(0, end - start);
    }

    // FM:504-552.  `cap` is locations.length (ArrayIndexOutOfBounds beyond it).
    int32_t locate(const uint16_t* pattern, int32_t offset, int32_t len, int32_t* locations, int64_t cap, int32_t max_matches) const {
        int32_t i = (offset + len) - 1;
        if (i < 0) throw JavaThrow{ST_INDEX_OOB, i};
        int16_t c = code(pattern[i]);
        if (c == 0) return 0;
        int32_t start = C.at((size_t)c), end = C.at((size_t)c + 1);
        int32_t k = 0;
        while (start < end && i >= offset + 1) {
            c = code(pattern[--i]);
            if (c == 0) return 0;
            start = (int32_t)(C[(size_t)c] + wf.rank(start, c));
            end = (int32_t)(C[(size_t)c] + wf.rank(end, c));
        }
        if (start < end) {
            i = start + 1;
            while (i <= end) {
                int32_t j = i, distance = 0;
                while (!sampled.access(j - 1)) {
                    int64_t tuple = wf.inverse_select(j - 1);
                    c = (int16_t)tuple;
                    int32_t rk = (int32_t)wf.rank(j, c);
                    j = C.at((size_t)c) + rk;
                    ++distance;
                    if (distance > length) throw JavaThrow{ST_NO_TERMINATION, distance};
                }
                if (k >= cap) throw JavaThrow{ST_INDEX_OOB, k};
                locations[k] = (int32_t)(suffixes.get(sampled.rank_ones(j) - 1, bw_suffixes) + (uint64_t)distance);
                ++k;
                if (k == max_matches) break;
                ++i;
            }
        }
        return k;
    }

    inline int32_t lf(int32_t& sample_position) const {  // FM:597-599 (one LF step; returns the code)
        int16_t c = (int16_t)wf.inverse_select((int64_t)sample_position - 1);
        sample_position = (int32_t)(C.at((size_t)c) + wf.rank(sample_position, c));
        return c;
    }

    // FM:564-608
    int32_t extract(int32_t start, int32_t stop, uint16_t* dst, int64_t dst_len, int32_t offset) const {
        if (!enable_extract) throw JavaThrow{ST_NOT_ENABLED, 0};
        if (start < 0) throw JavaThrow{ST_POS_NEGATIVE, 0};
        if (stop >= length) throw JavaThrow{ST_STOP_TOO_LONG, 0};
        int32_t sample_position = (int32_t)(positions.get((stop / sample_rate) + 1, bw_positions) + 1);
        int32_t skip = sample_rate - stop % sample_rate;
        if ((stop / sample_rate) == positions.length - 2) skip = length - stop;
        int32_t range = stop - start;
        if (dst_len - offset < range) throw JavaThrow{ST_DST_TOO_SMALL, 0};
        int32_t remaining = range, distance = 0;
        while (remaining > 0) {
            int32_t c = lf(sample_position);
            if (distance >= skip) {
                int64_t at = (int64_t)remaining - 1 + offset;
                if (at < 0 || at >= dst_len) throw JavaThrow{ST_INDEX_OOB, (int)at};
                dst[at] = (uint16_t)lookup.at((size_t)c);
                remaining--;
            }
            distance++;
        }
        return range;
    }

    void check_bounds(int32_t from, int64_t dst_len) const {  // FM:610-626
        if (!enable_extract) throw JavaThrow{ST_NOT_ENABLED, 0};
        if (from < 0) throw JavaThrow{ST_POS_NEGATIVE, 0};
        if (from >= length) throw JavaThrow{ST_POS_TOO_LONG, 0};
        if (dst_len == 0) throw JavaThrow{ST_DST_ZERO, 0};
    }
    inline void sample_for(int32_t from, int32_t& sample_position, int32_t& skip) const {  // FM:645-653
        sample_position = (int32_t)(positions.get((from / sample_rate) + 1, bw_positions) + 1);
        skip = sample_rate - from % sample_rate;
        if ((from / sample_rate) == positions.length - 2) skip = length - from;
    }
    static inline void put(uint16_t* dst, int64_t dst_len, int64_t at, uint16_t v) {
        if (at < 0 || at >= dst_len) throw JavaThrow{ST_INDEX_OOB, (int)at};
        dst[at] = v;
    }
    static void arraycopy(uint16_t* dst, int64_t dst_len, int64_t src, int64_t to, int64_t n) {  // System.arraycopy
        if (n < 0 || src < 0 || to < 0 || src + n > dst_len || to + n > dst_len) throw JavaThrow{ST_INDEX_OOB, (int)to};
        memmove(dst + to, dst + src, (size_t)n * sizeof(uint16_t));
    }

    // FM:640-759
    int32_t extract_until_boundary(int32_t from, uint16_t* dst, int64_t dst_len, int32_t offset, uint16_t boundary) const {
        check_bounds(from, dst_len);
        int32_t sample_position, skip;
        sample_for(from, sample_position, skip);
        int64_t down_pos = dst_len - 1;
        int16_t mb = code(boundary);
        if (mb == 0) throw JavaThrow{ST_NO_BOUNDARY, 0};
        int64_t remaining = dst_len;
        int32_t distance = 0;
        while (remaining > 0) {
            int32_t c = lf(sample_position);
            if (distance >= skip) {
                if (c == mb) break;
                if (c == 0) break;
                put(dst, dst_len, down_pos--, (uint16_t)lookup.at((size_t)c));
                remaining--;
            }
            distance++;
        }
        int64_t down_len = dst_len - (down_pos + 1);
        arraycopy(dst, dst_len, down_pos + 1, offset, down_len);
        const int32_t step = 4;
        int64_t up_pos;
        int64_t final_pos = -1;
        int32_t times = 1;
        while (final_pos == -1) {
            int32_t prev_from = from;
            from += step;
            from = std::min(from, length - 1);
            int32_t rem = from - prev_from;
            up_pos = (int64_t)(times - 1) * step + rem - 1;
            sample_for(from, sample_position, skip);
            distance = 0;
            while (rem > 0) {
                int32_t c = lf(sample_position);
                if (distance >= skip) {
                    if (c == mb) {
                        if (up_pos == 0) return 0;
                        final_pos = up_pos;
                    }
                    if (offset + down_len + up_pos >= dst_len) throw JavaThrow{ST_DOES_NOT_FIT, (int)(offset + down_len + up_pos)};
                    put(dst, dst_len, offset + down_len + (up_pos--), (uint16_t)lookup.at((size_t)c));
                    rem--;
                }
                distance++;
            }
            if (from == length - 1) {  // FM:745-752 (quirk Q5)
                final_pos = (up_pos < 0) ? 1 : up_pos + from - prev_from;
                break;
            }
            ++times;
        }
        return (int32_t)(down_len + final_pos);
    }

    // FM:772-831
    int32_t extract_until_boundary_left(int32_t from, uint16_t* dst, int64_t dst_len, int32_t offset, uint16_t boundary) const {
        ++from;
        check_bounds(from, dst_len);
        int32_t sample_position, skip;
        sample_for(from, sample_position, skip);
        int64_t down_pos = dst_len - 1;
        int16_t mb = code(boundary);
        if (mb == 0) throw JavaThrow{ST_NO_BOUNDARY, 0};
        int32_t distance = 0;
        while (true) {
            int32_t c = lf(sample_position);
            if (distance >= skip) {
                if (c == mb) break;
                if (c == 0) break;
                put(dst, dst_len, down_pos--, (uint16_t)lookup.at((size_t)c));
                if (down_pos == offset) throw JavaThrow{ST_DOES_NOT_FIT, (int)(dst_len - offset)};
            }
            distance++;
        }
        int64_t down_len = dst_len - (down_pos + 1);
        arraycopy(dst, dst_len, down_pos + 1, offset, down_len);
        return (int32_t)down_len;
    }

    // FM:844-922
    int32_t extract_until_boundary_right(int32_t from, uint16_t* dst, int64_t dst_len, int32_t offset, uint16_t boundary) const {
        check_bounds(from, dst_len);
        int16_t mb = code(boundary);
        if (mb == 0) throw JavaThrow{ST_NO_BOUNDARY, 0};
        const int32_t step = 4;
        int64_t up_pos;
        int64_t final_pos = -1;
        int32_t times = 1;
        while (final_pos == -1) {
            int32_t prev_from = from;
            from += step;
            from = std::min(from, length - 1);
            int32_t rem = from - prev_from;
            up_pos = (int64_t)(times - 1) * step + rem - 1;
            int32_t sample_position, skip;
            sample_for(from, sample_position, skip);
            int32_t distance = 0;
            while (rem > 0) {
                int32_t c = lf(sample_position);
                if (distance >= skip) {
                    if (c == mb) {
                        if (up_pos == 0) return 0;
                        final_pos = up_pos;
                    }
                    if (offset + up_pos >= dst_len) throw JavaThrow{ST_DOES_NOT_FIT, (int)(offset + up_pos)};
                    if (up_pos > 0) {
                        put(dst, dst_len, offset + (up_pos--) - 1, (uint16_t)lookup.at((size_t)c));
                    }
                    rem--;
                }
                distance++;
            }
            if (from == length - 1) {
                final_pos = up_pos + from - prev_from;
                break;
            }
            ++times;
        }
        return (int32_t)(final_pos - 1);
    }
};

// FmIndex.convertBytePatternToCharPattern (fm/FmIndex.java:239-298), restated branch by branch.  `pattern` is a Java byte[]
// of `array_len` bytes: reading past its end throws ArrayIndexOutOfBounds (a multi-byte sequence cut off by the end of the
// array).  Returns the number of chars written to `destination`.
int64_t convert_byte_pattern_to_char_pattern(const int8_t* pattern, int64_t array_len, int64_t offset, int64_t length, uint16_t* destination) {
    auto at = [&](int64_t i) -> int32_t {
        if (i < 0 || i >= array_len) throw std::out_of_range("pattern");
        return (int32_t)pattern[i];  // Java promotes the (signed) byte to int
    };
    int64_t pos = offset;
    int64_t i = 0;
    while (pos < length + offset) {  // :245
        const int32_t firstByte = at(pos);
        uint16_t nextChar;
        if (firstByte < 0) {  // :247
            if (((firstByte & 0xF0) >> 3) == 30) {  // :249 four-byte sequence
                const int32_t secondByte = at(pos + 1), thirdByte = at(pos + 2), fourthByte = at(pos + 3);
                pos += 4;
                const int32_t beforeConversion =
                    (((firstByte & 0x07) << 18) | ((secondByte & 0x3F) << 12) | ((thirdByte & 0x3F) << 6) | (fourthByte & 0x3F)) & 0x1FFFFF;  // :255-260
                if (beforeConversion > 32767) throw JavaThrow{ST_CHAR_EXCEEDS, beforeConversion};  // :261-267
                nextChar = (uint16_t)beforeConversion;
            } else if (((firstByte & 0xE0) >> 4) == 14) {  // :270 three-byte sequence
                const int32_t secondByte = at(pos + 1), thirdByte = at(pos + 2);
                pos += 3;
                nextChar = (uint16_t)((((firstByte & 0x0F) << 12) | ((secondByte & 0x3F) << 6) | (thirdByte & 0x3F)) & 0xFFFF);  // :274-279
            } else {  // :280 two-byte sequence (every other negative first byte, continuation bytes included)
                const int32_t secondByte = at(pos + 1);
                pos += 2;
                nextChar = (uint16_t)((((firstByte & 0x1F) << 6) | (secondByte & 0x3F)) & 0x7FF);  // :284-287
            }
        } else {  // :289 single byte
            ++pos;
            nextChar = (uint16_t)firstByte;
        }
        destination[i++] = nextChar;  // :294
    }
    return i;
}

template <typename F>
int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const JavaThrow& t) {
        g_err = "java exception status " + std::to_string(t.status);
        return -100 - t.status;
    } catch (const std::out_of_range& e) {
        g_err = std::string("ArrayIndexOutOfBounds: ") + e.what();
        return -100 - ST_INDEX_OOB;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

template <typename F>
void parallel_for(int64_t n, int threads, F&& f) {
    if (threads <= 1 || n < 2) {
        for (int64_t i = 0; i < n; ++i) f(i);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        int64_t lo = n * t / threads, hi = n * (t + 1) / threads;  // contiguous slice per thread
        pool.emplace_back([lo, hi, &f]() {
            for (int64_t i = lo; i < hi; ++i) f(i);
        });
    }
    for (auto& th : pool) th.join();
}

}  // namespace

extern "C" {

const char* orc_last_error(void) { return g_err.c_str(); }

int orc_fm_load(const uint8_t* buf, uint64_t len, void** out) {
    return guarded([&]() {
        Reader r(buf, (size_t)len);
        FmIndex* f = new FmIndex();
        try {
            f->read(r);
        } catch (...) {
            delete f;
            throw;
        }
        *out = f;
    });
}
void orc_fm_free(void* h) { delete (FmIndex*)h; }
int32_t orc_fm_input_length(void* h) { return ((FmIndex*)h)->length; }            // FM:929
int32_t orc_fm_alphabet_length(void* h) { return (int32_t)((FmIndex*)h)->map.size(); }  // FM:939
int32_t orc_fm_sample_rate(void* h) { return ((FmIndex*)h)->sample_rate; }
int32_t orc_fm_extract_enabled(void* h) { return ((FmIndex*)h)->enable_extract ? 1 : 0; }
int32_t orc_fm_block_size_log(void* h, int64_t sb) { return ((FmIndex*)h)->wf.sbs.at((size_t)sb).block_size_log; }
int64_t orc_fm_num_superblocks(void* h) { return (int64_t)((FmIndex*)h)->wf.sbs.size(); }

// status: 0 or a Java-exception status (ST_*); result in *out.
int orc_fm_count(void* h, const uint16_t* pat, int32_t off, int32_t len, int32_t* out) {
    int status = 0;
    try {
        *out = ((FmIndex*)h)->count(pat, off, len);
    } catch (const JavaThrow& t) {
        status = t.status;
    } catch (const std::out_of_range&) {
        status = ST_INDEX_OOB;
    }
    return status;
}
int orc_fm_locate(void* h, const uint16_t* pat, int32_t off, int32_t len, int32_t* locations, int64_t cap, int32_t max_matches,
                  int32_t* out) {
    int status = 0;
    try {
        *out = ((FmIndex*)h)->locate(pat, off, len, locations, cap, max_matches);
    } catch (const JavaThrow& t) {
        status = t.status;
    } catch (const std::out_of_range&) {
        status = ST_INDEX_OOB;
    }
    return status;
}
int orc_fm_extract(void* h, int32_t start, int32_t stop, uint16_t* dst, int64_t dst_len, int32_t offset, int32_t* out) {
    int status = 0;
    try {
        *out = ((FmIndex*)h)->extract(start, stop, dst, dst_len, offset);
    } catch (const JavaThrow& t) {
        status = t.status;
        *out = t.n;
    } catch (const std::out_of_range&) {
        status = ST_INDEX_OOB;
    }
    return status;
}
// mode: 0 = extractUntilBoundary, 1 = ...Left, 2 = ...Right.  On ST_DOES_NOT_FIT *out = N of the message.
int orc_fm_extract_until_boundary(void* h, int32_t from, uint16_t* dst, int64_t dst_len, int32_t offset, uint16_t boundary,
                                  int32_t mode, int32_t* out) {
    int status = 0;
    try {
        FmIndex* f = (FmIndex*)h;
        *out = mode == 0   ? f->extract_until_boundary(from, dst, dst_len, offset, boundary)
               : mode == 1 ? f->extract_until_boundary_left(from, dst, dst_len, offset, boundary)
                           : f->extract_until_boundary_right(from, dst, dst_len, offset, boundary);
    } catch (const JavaThrow& t) {
        status = t.status;
        *out = t.n;
    } catch (const std::out_of_range&) {
        status = ST_INDEX_OOB;
    }
    return status;
}

// Batch forms (one shared immutable index, `threads` host threads each owning a contiguous slice,
// like sharing one @ThreadSafe FmIndex across a Java thread pool).
void orc_fm_count_batch(void* h, const uint16_t* chars, const uint64_t* pat_off, uint32_t n_pat, int32_t* counts, int32_t* status,
                        int32_t threads) {
    FmIndex* f = (FmIndex*)h;
    parallel_for((int64_t)n_pat, threads, [&](int64_t i) {
        int32_t len = (int32_t)(pat_off[i + 1] - pat_off[i]);
        int32_t c = 0;
        int st = orc_fm_count(f, chars + pat_off[i], 0, len, &c);
        counts[i] = c;
        if (status) status[i] = st;
    });
}
// convertBytePatternToCharPattern + count per pattern; every pattern is its own byte[] (bytes[pat_off[i], pat_off[i+1])).
// status 10: the converter's "Found a character that exceeds (32767): it was N", N in counts[i].
void orc_fm_count_batch_utf8(void* h, const uint8_t* bytes, const uint64_t* pat_off, uint32_t n_pat, int32_t* counts, int32_t* status,
                             int32_t threads) {
    FmIndex* f = (FmIndex*)h;
    parallel_for((int64_t)n_pat, threads, [&](int64_t i) {
        const int64_t len = (int64_t)(pat_off[i + 1] - pat_off[i]);
        std::vector<uint16_t> dst((size_t)len + 1);
        int32_t c = 0;
        int st = 0;
        try {
            const int64_t n = convert_byte_pattern_to_char_pattern((const int8_t*)bytes + pat_off[i], len, 0, len, dst.data());
            st = orc_fm_count(f, dst.data(), 0, (int32_t)n, &c);
        } catch (const JavaThrow& t) {
            st = t.status;
            c = t.n;
        } catch (const std::out_of_range&) {
            st = ST_INDEX_OOB;
        }
        counts[i] = c;
        if (status) status[i] = st;
    });
}
void orc_fm_locate_batch_utf8(void* h, const uint8_t* bytes, const uint64_t* pat_off, uint32_t n_pat, int32_t max_matches, int32_t* n_hits,
                              int32_t* positions, int64_t stride, int32_t* status, int32_t threads) {
    FmIndex* f = (FmIndex*)h;
    parallel_for((int64_t)n_pat, threads, [&](int64_t i) {
        const int64_t len = (int64_t)(pat_off[i + 1] - pat_off[i]);
        std::vector<uint16_t> dst((size_t)len + 1);
        int32_t k = 0;
        int st = 0;
        try {
            const int64_t n = convert_byte_pattern_to_char_pattern((const int8_t*)bytes + pat_off[i], len, 0, len, dst.data());
            st = orc_fm_locate(f, dst.data(), 0, (int32_t)n, positions + i * stride, stride, max_matches, &k);
        } catch (const JavaThrow& t) {
            st = t.status;
        } catch (const std::out_of_range&) {
            st = ST_INDEX_OOB;
        }
        n_hits[i] = k;
        if (status) status[i] = st;
    });
}
// the converter alone: returns the char count, or -100 - status (value of the offending code point in *value)
int64_t orc_convert_utf8(const uint8_t* bytes, int64_t array_len, int64_t offset, int64_t length, uint16_t* dst, int32_t* value) {
    try {
        return convert_byte_pattern_to_char_pattern((const int8_t*)bytes, array_len, offset, length, dst);
    } catch (const JavaThrow& t) {
        if (value) *value = t.n;
        return -100 - t.status;
    } catch (const std::out_of_range&) {
        return -100 - ST_INDEX_OOB;
    }
}
// positions of pattern i are written at positions + i*stride (stride >= max hits expected).
void orc_fm_locate_batch(void* h, const uint16_t* chars, const uint64_t* pat_off, uint32_t n_pat, int32_t max_matches, int32_t* n_hits,
                         int32_t* positions, int64_t stride, int32_t* status, int32_t threads) {
    FmIndex* f = (FmIndex*)h;
    parallel_for((int64_t)n_pat, threads, [&](int64_t i) {
        int32_t len = (int32_t)(pat_off[i + 1] - pat_off[i]);
        int32_t k = 0;
        int st = orc_fm_locate(f, chars + pat_off[i], 0, len, positions + i * stride, stride, max_matches, &k);
        n_hits[i] = k;
        if (status) status[i] = st;
    });
}
// every item: extract(start, stop, destination = new char[stride], offset)
void orc_fm_extract_batch(void* h, const int32_t* start, const int32_t* stop, uint32_t n, uint16_t* arena, int64_t stride, int32_t* len_out,
                          int32_t* status, int32_t threads, int32_t offset) {
    FmIndex* f = (FmIndex*)h;
    parallel_for((int64_t)n, threads, [&](int64_t i) {
        int32_t k = 0;
        int st = orc_fm_extract(f, start[i], stop[i], arena + i * stride, stride, offset, &k);
        len_out[i] = k;
        status[i] = st;
    });
}
// every item: extractUntilBoundary*(from, destination = new char[dst_len], offset, boundary)
void orc_fm_extract_until_boundary_batch(void* h, const int32_t* from, uint32_t n, uint16_t boundary, int32_t dst_len, int32_t mode,
                                         uint16_t* arena, int32_t* len_out, int32_t* status, int32_t threads, int32_t offset) {
    FmIndex* f = (FmIndex*)h;
    parallel_for((int64_t)n, threads, [&](int64_t i) {
        int32_t k = 0;
        int st = orc_fm_extract_until_boundary(f, from[i], arena + (int64_t)i * dst_len, dst_len, offset, boundary, mode, &k);
        len_out[i] = k;
        status[i] = st;
    });
}

// Work counters since the last reset: rank queries, tree levels walked by rank, LF steps
// (inverseSelect calls) and their levels.  Used by tests to cross-check the GPU's own counters.
void orc_fm_stats(void* h, uint64_t* out4, int32_t reset) {
    (void)h;  // process-wide totals (one index is queried at a time by the tests that read them)
#ifndef ORACLE_NO_STATS
    t_tally.flush();
    for (int i = 0; i < 4; ++i) {
        out4[i] = g_totals[i].load();
        if (reset) g_totals[i] = 0;
    }
#else
    (void)reset;
    for (int i = 0; i < 4; ++i) out4[i] = 0;
#endif
}
int32_t orc_has_stats(void) {
#ifndef ORACLE_NO_STATS
    return 1;
#else
    return 0;
#endif
}

// Direct access to the two lower layers of the loaded index (for layer tests).
int orc_fm_wfbb_rank(void* h, int64_t pos, int32_t sym, int64_t* out) {
    int status = 0;
    try {
        *out = ((FmIndex*)h)->wf.rank(pos, (int16_t)sym);
    } catch (const JavaThrow& t) {
        status = t.status;
    } catch (const std::out_of_range&) {
        status = ST_INDEX_OOB;
    }
    return status;
}
int orc_fm_wfbb_inverse_select(void* h, int64_t pos, int64_t* out) {
    int status = 0;
    try {
        *out = ((FmIndex*)h)->wf.inverse_select(pos);
    } catch (const JavaThrow& t) {
        status = t.status;
    } catch (const std::out_of_range&) {
        status = ST_INDEX_OOB;
    }
    return status;
}
int32_t orc_fm_sampled_rank(void* h, int32_t pos) { return ((FmIndex*)h)->sampled.rank_ones(pos); }
int32_t orc_fm_sampled_access(void* h, int32_t pos) {
    try {
        return ((FmIndex*)h)->sampled.access(pos) ? 1 : 0;
    } catch (const JavaThrow&) {
        return -1;
    }
}

// Stand-alone structures (the reference's own unit tests exercise them outside an FmIndex).
int orc_wfbb_load(const uint8_t* buf, uint64_t len, void** out) {
    return guarded([&]() {
        Reader r(buf, (size_t)len);
        Wfbb* w = new Wfbb();
        try {
            w->read(r);
        } catch (...) {
            delete w;
            throw;
        }
        *out = w;
    });
}
void orc_wfbb_free(void* h) { delete (Wfbb*)h; }
int orc_wfbb_rank(void* h, int64_t pos, int32_t sym, int64_t* out) {
    int status = 0;
    try {
        *out = ((Wfbb*)h)->rank(pos, (int16_t)sym);
    } catch (const JavaThrow& t) {
        status = t.status;
    } catch (const std::out_of_range&) {
        status = ST_INDEX_OOB;
    }
    return status;
}
int orc_wfbb_inverse_select(void* h, int64_t pos, int64_t* out) {
    int status = 0;
    try {
        *out = ((Wfbb*)h)->inverse_select(pos);
    } catch (const JavaThrow& t) {
        status = t.status;
    } catch (const std::out_of_range&) {
        status = ST_INDEX_OOB;
    }
    return status;
}
int orc_rrr_load(const uint8_t* buf, uint64_t len, void** out) {
    return guarded([&]() {
        Reader r(buf, (size_t)len);
        RrrVector* v = new RrrVector();
        try {
            v->read(r);
        } catch (...) {
            delete v;
            throw;
        }
        *out = v;
    });
}
void orc_rrr_free(void* h) { delete (RrrVector*)h; }
int32_t orc_rrr_rank_ones(void* h, int32_t pos) { return ((RrrVector*)h)->rank_ones(pos); }
int32_t orc_rrr_access(void* h, int32_t pos) {
    try {
        return ((RrrVector*)h)->access(pos) ? 1 : 0;
    } catch (const JavaThrow&) {
        return -1;
    }
}
// The generated (class,offset)->block table, for the sha256 pin.
void orc_rrr_inverse_table(uint16_t* out32768) { memcpy(out32768, tables().inverse, sizeof(uint16_t) * 32768); }

}  // extern "C"
