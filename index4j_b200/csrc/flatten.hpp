// Host-side re-layout: parsed Java structures (jstream.hpp) -> the sector-record device format
// (layout.h).  Runs once per fmgpu_index_load_serialized.  The information content is unchanged:
// every record is a pre-evaluation of the position-independent part of what the reference's
// WaveletFixedBlockBoosting.rank / inverseSelect and RrrVector.rankOnes / access compute per call
// (wavelet/WaveletFixedBlockBoosting.java:1010-1537, bitsequence/RrrVector.java:314-396).
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "jstream.hpp"
#include "layout.h"

namespace fmgpu_host {

using fmgpu::Rec32;

// (class, offset) -> 15-bit block.  The reference ships these tables as literals (RrrVector.java:8692-8698, :8705-16899); here
// they are computed by combinatorial unranking: class k = blocks with k ones, C(15, k) of them; inside a class the blocks are
// in descending order of their value read LSB-first, so offset `off` of class k is found bit by bit — bit j (the j-th position
// of the block) is set iff off < C(14 - j, ones still to place - 1), else off skips those C(..) blocks.  (The oracle enumerates
// bit-reversed values instead; tests pin both against the sha256 of the Java literal.)
struct RrrTables {
    uint16_t inverse[32768];
    uint16_t class_base[16];
    uint8_t bits_needed[16];  // RrrVector.java:111-129
    RrrTables() {
        uint32_t binom[16][16] = {{0}};  // binom[n][k] = C(n, k)
        for (int n = 0; n < 16; ++n) {
            binom[n][0] = 1;
            for (int k = 1; k <= n; ++k) binom[n][k] = binom[n - 1][k - 1] + (k <= n - 1 ? binom[n - 1][k] : 0u);
        }
        uint32_t acc = 0;
        for (int k = 0; k < 16; ++k) {
            const uint32_t members = binom[15][k];
            class_base[k] = (uint16_t)acc;
            int b = 0;
            while ((1u << b) <= members) ++b;
            bits_needed[k] = (uint8_t)b;
            for (uint32_t off = 0; off < members; ++off) {
                uint32_t v = 0, rest = off, ones = (uint32_t)k;
                for (int j = 0; j < 15 && ones; ++j) {
                    const uint32_t with_bit = binom[14 - j][ones - 1];  // blocks of this class that have bit j set, given the prefix
                    if (rest < with_bit) {
                        v |= 1u << j;
                        --ones;
                    } else {
                        rest -= with_bit;
                    }
                }
                inverse[acc + off] = (uint16_t)v;
            }
            acc += members;
        }
    }
};
inline const RrrTables& rrr_tables() {
    static const RrrTables t;
    return t;
}

// Whole RRR vector -> plain LSB-first bits (one slack word at the end).
inline void rrr_decode_all(const RrrStream& r, std::vector<uint64_t>& bits) {
    const RrrTables& T = rrr_tables();
    const int64_t len = r.length;
    const int64_t nblocks = (len + 14) / 15;
    bits.assign((size_t)((len + 63) / 64) + 2, 0);
    uint64_t bitpos = 0;
    const uint64_t off_bits_avail = (uint64_t)(r.offsets.size() - 1) * 64;
    for (int64_t b = 0; b < nblocks; ++b) {
        const int cls = (int)r.classes.get(b);
        const int nb = T.bits_needed[cls];
        if (bitpos + (uint64_t)nb > off_bits_avail) throw FormatError("RrrVector offset stream too short");
        const size_t w = (size_t)(bitpos >> 6);
        const int sh = (int)(bitpos & 63);
        uint64_t off = r.offsets[w] >> sh;
        if (sh + nb > 64) off |= r.offsets[w + 1] << (64 - sh);
        off &= (1ULL << nb) - 1;
        const uint32_t idx = (uint32_t)T.class_base[cls] + (uint32_t)off;
        const uint64_t v = idx < 32768 ? T.inverse[idx] : 0;
        const uint64_t p = (uint64_t)b * 15;
        bits[p >> 6] |= v << (p & 63);
        if ((p & 63) + 15 > 64) bits[(p >> 6) + 1] |= v >> (64 - (p & 63));
        bitpos += (uint64_t)nb;
    }
}

inline uint32_t bits_get(const std::vector<uint64_t>& bits, uint64_t pos, int n) {  // n <= 32
    const size_t w = (size_t)(pos >> 6);
    const int sh = (int)(pos & 63);
    uint64_t v = bits[w] >> sh;
    if (sh + n > 64) v |= bits[w + 1] << (64 - sh);
    return (uint32_t)(n >= 32 ? (v & 0xffffffffULL) : (v & ((1ULL << n) - 1)));
}

struct FlatIndex {
    fmgpu::DevIndex meta{};  // pointers unset
    std::vector<uint32_t> C;
    std::vector<uint16_t> char2code, code2char;
    std::vector<fmgpu::SbDesc> sb;
    std::vector<fmgpu::Cell8> cells;
    std::vector<Rec32> sectors, occ, blocks, nodes, sgroups, sa, isa;
    std::vector<uint32_t> soffsets;
    int32_t alphabet_length = 0;
    std::vector<uint32_t> occ_base, occ_used;  // per block: first occurrence record reserved by the sizing pass / records in use
};

// Explicit shape of one block's Huffman-shaped wavelet tree, rebuilt from the variable-size
// block header (layout written by encodeBlock, WaveletFixedBlockBoosting.java:742-809).
struct BlockTree {
    struct Node {
        uint32_t start, size;  // bit range in the superblock's level bitvector
        int32_t child[2];      // >= 0: internal node id; < 0: -(leaf local index + 1)
        uint32_t sector;       // even-depth nodes: first record (global Rec32 index, even), set by the caller
        uint32_t depth;        // root = 0
        uint32_t enode;        // even-depth nodes: index of the node record (global), set by the caller
    };
    std::vector<Node> nodes;  // BFS order, node 0 = root
    struct Leaf {
        int32_t parent;
        uint8_t bit, len;
        uint32_t code;
    };
    std::vector<Leaf> leaves;     // by block-local symbol index (canonical code order)
    std::vector<uint16_t> sym;    // header symbol per leaf
    std::vector<uint32_t> brank;  // rankAtBlockBoundary per leaf
    std::vector<uint32_t> occ;    // occurrences of the leaf's symbol in the block (= elements that reach the leaf)
    int h = 0;
    uint32_t n_sectors = 0, n_occ = 0, n_even = 0;  // level records / occurrence records / even-depth internal nodes
};

struct VarReader {
    const std::vector<uint8_t>& v;
    bool ok = true;
    explicit VarReader(const std::vector<uint8_t>& v_) : v(v_) {}
    uint32_t u16(int64_t p) {
        if (p < 0 || (size_t)p + 2 > v.size()) {
            ok = false;
            return 0;
        }
        return (uint32_t)v[(size_t)p] | ((uint32_t)v[(size_t)p + 1] << 8);
    }
    uint32_t u24(int64_t p) {
        if (p < 0 || (size_t)p + 3 > v.size()) {
            ok = false;
            return 0;
        }
        return (uint32_t)v[(size_t)p] | ((uint32_t)v[(size_t)p + 1] << 8) | ((uint32_t)v[(size_t)p + 2] << 16);
    }
};

// occurrence records a (block, symbol) pair with `occ` occurrences in a block of `block_size` positions needs (layout.h):
// one list record while its positions fit one, else (upper bound for the sizing pass) one bit per position
inline uint32_t occ_kind(uint32_t occ, uint32_t) { return occ <= fmgpu::OCC_RANGE_MAX ? fmgpu::CELL_OCC_LIST : fmgpu::CELL_OCC_BITS; }
inline uint32_t occ_records(uint32_t occ, uint32_t block_size) {
    return occ_kind(occ, block_size) == fmgpu::CELL_OCC_LIST ? 1u : block_size / fmgpu::OCC_BITS_PER_REC + 1;
}

// `sigma`: the wavelet alphabet size — header symbols must lie below it (the kernels index C[] and the rank directories with them)
inline void build_block_tree(const SuperBlockHdr& S, size_t b, uint32_t cur_block_size, BlockTree& T, int32_t sigma) {
    const BlockHdr& H = S.blocks[b];
    const int h = H.tree_height;
    const int sig = (int)H.sigma_m1 + 1;
    T.h = h;
    T.nodes.clear();
    T.leaves.clear();
    T.sym.clear();
    T.brank.clear();
    T.occ.clear();
    T.n_sectors = 0;
    T.n_occ = 0;
    T.n_even = 0;
    if (sig < 1 || h < 0) throw FormatError("block header out of range");
    VarReader R(S.var);
    const int64_t var_off = H.var_off;
    const int64_t ptr32 = var_off + (h > 0 ? (int64_t)(h - 1) * 4 : 0);
    T.sym.resize((size_t)sig);
    T.brank.resize((size_t)sig);
    for (int i = 0; i < sig; ++i) {
        T.sym[(size_t)i] = (uint16_t)R.u16(ptr32 + 5 * (int64_t)i);
        T.brank[(size_t)i] = R.u24(ptr32 + 5 * (int64_t)i + 2);
    }
    if (!R.ok) throw FormatError("variable block header truncated");
    for (int i = 0; i < sig; ++i)
        if ((int32_t)T.sym[(size_t)i] >= sigma) throw FormatError("block header symbol outside the alphabet");
    if (h == 0) return;  // run block: no tree
    if (sig < 2) throw FormatError("tree block with a single symbol");
    T.leaves.resize((size_t)sig);
    T.occ.assign((size_t)sig, 0);
    int64_t third = ptr32 + 5 * (int64_t)sig;
    std::vector<uint32_t> level;  // node ids of the current depth, left to right
    T.nodes.push_back({(uint32_t)H.bv_offset, cur_block_size, {0, 0}, 0, 0, 0});
    level.push_back(0);
    uint32_t level_start = (uint32_t)H.bv_offset;
    int64_t leaves_before = 0;
    std::vector<uint32_t> next;
    for (int d = 0; d < h; ++d) {
        const size_t n_int = level.size();
        if (n_int == 0) throw FormatError("wavelet tree level without internal nodes");
        uint32_t depth_total = 0;
        for (uint32_t id : level) depth_total += T.nodes[id].size;
        const uint32_t next_start = level_start + depth_total;
        const int64_t nleaf_next = (d + 1 < h) ? (int64_t)R.u16(var_off + 4 * (int64_t)d) : (int64_t)(2 * n_int);
        if (nleaf_next > (int64_t)(2 * n_int)) throw FormatError("wavelet tree leaf count exceeds level width");
        next.clear();
        uint32_t run = 0, prev_cum = 0;
        for (size_t j = 0; j < n_int; ++j) {
            const uint32_t id = level[j];
            const uint32_t cum = R.u16(third + 2 * (int64_t)j);
            const uint32_t ones = cum - prev_cum;
            prev_cum = cum;
            const uint32_t size = T.nodes[id].size;
            if (ones > size) throw FormatError("wavelet node has more ones than bits");
            for (int bit = 0; bit < 2; ++bit) {
                const int64_t slot = 2 * (int64_t)j + bit;
                const uint32_t csize = bit ? ones : size - ones;
                if (slot < nleaf_next) {
                    const int64_t li = leaves_before + slot;
                    if (li >= sig) throw FormatError("wavelet tree has more leaves than symbols");
                    T.nodes[id].child[bit] = (int32_t)(-(li + 1));
                    T.leaves[(size_t)li].parent = (int32_t)id;
                    T.leaves[(size_t)li].bit = (uint8_t)bit;
                    T.occ[(size_t)li] = csize;
                } else {
                    const uint32_t nid = (uint32_t)T.nodes.size();
                    T.nodes[id].child[bit] = (int32_t)nid;
                    T.nodes.push_back({next_start + run, csize, {0, 0}, 0, (uint32_t)d + 1u, 0});
                    run += csize;
                    next.push_back(nid);
                }
            }
        }
        if (!R.ok) throw FormatError("variable block header truncated");
        third += 2 * (int64_t)n_int;
        leaves_before += nleaf_next;
        level.swap(next);
        level_start = next_start;
    }
    if (leaves_before != sig) throw FormatError("wavelet tree leaf count differs from block alphabet");
    // parents of internal nodes, for path reconstruction
    std::vector<int32_t> parent(T.nodes.size(), -1);
    std::vector<uint8_t> pbit(T.nodes.size(), 0);
    for (size_t id = 0; id < T.nodes.size(); ++id)
        for (int bit = 0; bit < 2; ++bit)
            if (T.nodes[id].child[bit] >= 0 && (size_t)T.nodes[id].child[bit] != 0) {
                parent[(size_t)T.nodes[id].child[bit]] = (int32_t)id;
                pbit[(size_t)T.nodes[id].child[bit]] = (uint8_t)bit;
            }
    for (int i = 0; i < sig; ++i) {
        BlockTree::Leaf& L = T.leaves[(size_t)i];
        uint32_t code = L.bit;
        int len = 1;
        int32_t id = L.parent;
        while (parent[(size_t)id] >= 0) {
            code |= (uint32_t)pbit[(size_t)id] << len;
            ++len;
            id = parent[(size_t)id];
        }
        if (len > 32) throw FormatError("Huffman code longer than 32 bits");
        L.len = (uint8_t)len;
        L.code = code;  // bit (len-1) = root decision
        T.n_occ += occ_records(T.occ[(size_t)i], cur_block_size);
    }
    // level records of the even-depth nodes (inverseSelect walks them, lf_lane.h)
    for (size_t id = 0; id < T.nodes.size(); ++id) {
        const auto& n = T.nodes[id];
        if ((n.depth & 1u) == 0) {
            if (n.size > 65536u) throw FormatError("wavelet node larger than a block");
            T.n_sectors += n.size / fmgpu::SECTOR_BITS + 1;
            ++T.n_even;
        }
    }
}

struct SbPlan {
    uint32_t first_block = 0, rows = 0;
    uint64_t sector_base = 0, node_base = 0, occ_base = 0;
    uint64_t n_sectors = 0, n_nodes = 0, n_occ = 0;
};

inline uint32_t sb_block_size(const WfbbStream& W, size_t sb, size_t b) {
    const int64_t bs = 1LL << W.sbs[sb].block_size_log;
    const int64_t beg = ((int64_t)sb << 20) + (int64_t)b * bs;
    return (uint32_t)std::min<int64_t>(bs, W.size - beg);
}

template <typename F>
inline void parallel_sbs(size_t n, int threads, F&& f) {
    std::atomic<size_t> next(0);
    std::atomic<bool> failed(false);
    std::string err;
    auto worker = [&]() {
        try {
            for (;;) {
                const size_t i = next.fetch_add(1);
                if (i >= n || failed.load()) break;
                f(i);
            }
        } catch (const std::exception& e) {
            if (!failed.exchange(true)) err = e.what();
        }
    };
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = (int)std::max<size_t>(1, n);
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
    if (failed.load()) throw FormatError(err);
}

inline void put_cell(fmgpu::Cell8& c, uint32_t kind, uint32_t value) {
    c.value = value;
    c.info = kind << fmgpu::CELL_KIND_SHIFT;
}

// occurrence structure of one (block, symbol) pair: positions (ascending, block-relative) of the symbol's occurrences, its
// code length in the block's tree -> the cell's kind / record pointer and its records in F.occ[at ..] (layout.h)
inline uint32_t put_occ(fmgpu::Cell8& cell, const std::vector<uint16_t>& pos, uint32_t code_len, uint32_t block_size, uint32_t at,
                        std::vector<Rec32>& occ) {
    const uint32_t n = (uint32_t)pos.size();
    uint32_t kind = occ_kind(n, block_size);
    if (at > fmgpu::CELL_PTR_MASK) throw FormatError("more occurrence records than a cell can address");
    auto put16 = [](Rec32& R, uint32_t k, uint32_t v) {  // u16 slot k of a record
        uint32_t& w = R.w[k >> 1];
        w = (k & 1u) ? ((w & 0x0000ffffu) | (v << 16)) : ((w & 0xffff0000u) | v);
    };
    if (kind == fmgpu::CELL_OCC_LIST) {  // one record for the whole block
        cell.info = (kind << fmgpu::CELL_KIND_SHIFT) | at;
        Rec32& X = occ[(size_t)at];
        memset(&X, 0xff, sizeof X);
        X.w[0] = (code_len & 0xffu) << 24;
        for (uint32_t k = 0; k < n; ++k) put16(X, 2 + k, pos[k]);
        return 1;
    }
    // position lists over fixed ranges where no range holds more than 14 occurrences: the largest range that qualifies
    for (uint32_t shift : {12u, 10u}) {
        const uint32_t R = 1u << shift;
        if (R >= block_size) continue;
        bool ok = true;
        for (uint32_t i = fmgpu::OCC_RANGE_MAX; i < n && ok; ++i) ok = ((uint32_t)pos[i] >> shift) != ((uint32_t)pos[i - fmgpu::OCC_RANGE_MAX] >> shift);
        if (!ok) continue;
        kind = shift == 12u ? fmgpu::CELL_OCC_R4K : fmgpu::CELL_OCC_R1K;
        cell.info = (kind << fmgpu::CELL_KIND_SHIFT) | at;
        const uint32_t nrec = (block_size + R - 1) >> shift;
        size_t i = 0;
        for (uint32_t q = 0; q < nrec; ++q) {
            Rec32& X = occ[(size_t)at + q];
            memset(&X, 0xff, sizeof X);
            X.w[0] = (uint32_t)i | ((code_len & 0xffu) << 24);
            uint32_t k = 2;  // slots 2..15 = w1..w7
            while (i < n && ((uint32_t)pos[i] >> shift) == q) put16(X, k++, pos[i++]);
        }
        return nrec;
    }
    cell.info = (kind << fmgpu::CELL_KIND_SHIFT) | at;
    const uint32_t nrec = block_size / fmgpu::OCC_BITS_PER_REC + 1;
    size_t i = 0;
    for (uint32_t q = 0; q < nrec; ++q) {
        Rec32& R = occ[(size_t)at + q];
        memset(&R, 0, sizeof R);
        R.w[0] = (uint32_t)i | ((code_len & 0xffu) << 24);  // occurrences before the record (<= 65536)
        const uint32_t lo = q * fmgpu::OCC_BITS_PER_REC, hi = lo + fmgpu::OCC_BITS_PER_REC;
        while (i < n && pos[i] < hi) {
            const uint32_t bit = pos[i] - lo;
            R.w[1 + (bit >> 5)] |= 1u << (bit & 31u);
            ++i;
        }
    }
    return nrec;
}

inline void flatten_superblock(const WfbbStream& W, size_t sb, const SbPlan& P, FlatIndex& F) {
    const SuperBlockHdr& S = W.sbs[sb];
    const int32_t sigma = W.sigma;
    const int bl = S.block_size_log;
    const int blocks_log = 20 - bl;
    const int64_t blocks_in_sb = 1LL << blocks_log;
    const int64_t sb_sigma = (int64_t)S.sigma_m1 + 1;
    const size_t nblk = S.blocks.size();

    std::vector<uint64_t> bits;
    rrr_decode_all(S.rank_support, bits);
    const uint64_t nbits = (uint64_t)S.rank_support.length;

    // --- per block: tree, level + node records (inverseSelect), descriptors, and per leaf the positions of its occurrences
    std::vector<BlockTree> trees(nblk);
    std::vector<std::vector<std::vector<uint16_t>>> leaf_pos(nblk);  // [block][leaf] -> ascending positions in the block
    uint64_t sec = P.sector_base, node = P.node_base, occ_at = P.occ_base;
    std::vector<uint64_t> block_occ_base(nblk, 0);
    std::vector<uint16_t> cur, nxt;
    for (size_t b = 0; b < nblk; ++b) {
        BlockTree& T = trees[b];
        const uint32_t bsize = sb_block_size(W, sb, b);
        build_block_tree(S, b, bsize, T, W.sigma);
        Rec32& D = F.blocks[(size_t)P.first_block + b];
        memset(&D, 0, sizeof D);
        if (T.h == 0) {
            const uint32_t c8 = (uint32_t)T.sym[0] & 0xffu;
            D.w[1] = 1u | (c8 << 8);  // inverseSelect keeps the low byte only (:1329-1332)
            // inverseSelect's rank inside a single-symbol block (:1338-1352): boundary ranks of the TRUNCATED symbol at the
            // superblock / hyperblock level + the block boundary rank of the header entry + position in block
            if (c8 < (uint32_t)sigma)
                D.w[5] = (uint32_t)((uint64_t)W.hyper_rank[c8] + (uint64_t)(int64_t)W.sb_rank[sb * (size_t)sigma + c8] + T.brank[0]);
            continue;
        }
        block_occ_base[b] = occ_at;
        occ_at += T.n_occ;
        // record arrays and node-record indices of the even-depth nodes
        for (size_t id = 0; id < T.nodes.size(); ++id) {
            auto& n = T.nodes[id];
            if (n.depth & 1u) continue;
            n.enode = (uint32_t)node++;
            n.sector = (uint32_t)sec;
            sec += n.size / fmgpu::SECTOR_BITS + 1;
        }
        auto leaf_entry = [&](int32_t c, uint32_t flag, uint32_t* e) {
            const size_t li = (size_t)(-c - 1);
            const uint32_t sy = T.sym[li];
            const uint64_t base = (uint64_t)W.hyper_rank[sy < (uint32_t)sigma ? sy : 0] +
                                  (uint64_t)(sy < (uint32_t)sigma ? W.sb_rank[sb * (size_t)sigma + sy] : 0) + T.brank[li];
            e[0] = flag | sy;
            e[1] = (uint32_t)base;
        };
        for (size_t id = 0; id < T.nodes.size(); ++id) {
            const BlockTree::Node& n = T.nodes[id];
            if ((uint64_t)n.start + n.size > nbits) throw FormatError("wavelet node exceeds the level bitvector");
            if (n.depth & 1u) continue;
            const BlockTree::Node* ch[2] = {n.child[0] >= 0 ? &T.nodes[(size_t)n.child[0]] : nullptr,
                                            n.child[1] >= 0 ? &T.nodes[(size_t)n.child[1]] : nullptr};
            for (int t = 0; t < 2; ++t)
                if (ch[t] && (uint64_t)ch[t]->start + ch[t]->size > nbits) throw FormatError("wavelet node exceeds the level bitvector");
            // level records (layout.h): {c00 | c01 << 16, c10 | c11 << 16, 96 bits of this node, for the same 96 positions the bit
            // each element has one level further down (in the child it goes to; 0 if that child is a leaf)}
            const uint32_t nrec = n.size / fmgpu::SECTOR_BITS + 1;
            uint32_t cnt2[2][2] = {{0, 0}, {0, 0}}, cpos[2] = {0, 0};  // elements so far by (bit here, bit one level down)
            for (uint32_t q = 0; q < nrec; ++q) {
                Rec32& X = F.sectors[(size_t)n.sector + q];
                memset(&X, 0, sizeof X);
                for (int t = 0; t < 2; ++t)
                    for (int u = 0; u < 2; ++u)
                        if (cnt2[t][u] > 0xffffu) throw FormatError("wavelet node prefix count exceeds 16 bits");
                X.w[0] = cnt2[0][0] | (cnt2[0][1] << 16);
                X.w[1] = cnt2[1][0] | (cnt2[1][1] << 16);
                const uint32_t lo = q * fmgpu::SECTOR_BITS;
                const uint32_t valid = lo < n.size ? std::min<uint32_t>(fmgpu::SECTOR_BITS, n.size - lo) : 0u;
                for (uint32_t i = 0; i < valid; ++i) {
                    const uint32_t t = bits_get(bits, (uint64_t)n.start + lo + i, 1);
                    uint32_t cb = 0;
                    if (ch[t]) {
                        if (cpos[t] >= ch[t]->size) throw FormatError("wavelet child smaller than its parent's share");
                        cb = bits_get(bits, (uint64_t)ch[t]->start + cpos[t], 1);
                    }
                    ++cpos[t];
                    ++cnt2[t][cb];
                    if (t) X.w[2 + (i >> 5)] |= 1u << (i & 31u);
                    if (cb) X.w[5 + (i >> 5)] |= 1u << (i & 31u);
                }
            }
            // node record: entries [t][u] = {c, a}; child t a leaf: [t][0] = {LEAF1 | sym, boundary rank}
            Rec32& NR = F.nodes[(size_t)n.enode];
            memset(&NR, 0, sizeof NR);
            for (int t = 0; t < 2; ++t) {
                if (!ch[t]) {
                    leaf_entry(n.child[t], fmgpu::LEAF1_FLAG, &NR.w[4 * t]);
                    continue;
                }
                for (int u = 0; u < 2; ++u) {
                    const int32_t g = ch[t]->child[u];
                    uint32_t* e = &NR.w[4 * t + 2 * u];
                    if (g < 0) {
                        leaf_entry(g, fmgpu::LEAF_FLAG, e);
                    } else {
                        e[0] = T.nodes[(size_t)g].enode;
                        e[1] = T.nodes[(size_t)g].sector;
                    }
                }
            }
            if (id == 0) {
                D.w[0] = n.sector;
                D.w[4] = n.enode;
            }
        }
        // which leaf every position of the block ends in: the tree's bitvectors are stable partitions of the block, level by
        // level; the elements of a node in node order are the block positions that reach it, ascending
        auto& LP = leaf_pos[b];
        LP.assign(T.leaves.size(), {});
        for (size_t li = 0; li < T.leaves.size(); ++li) LP[li].reserve(T.occ[li]);
        std::vector<std::vector<uint16_t>> elems(T.nodes.size());
        elems[0].resize(bsize);
        for (uint32_t i = 0; i < bsize; ++i) elems[0][i] = (uint16_t)i;
        for (size_t id = 0; id < T.nodes.size(); ++id) {  // BFS order: parents before children
            const BlockTree::Node& n = T.nodes[id];
            std::vector<uint16_t>& E = elems[id];
            if (E.size() != n.size) throw FormatError("wavelet node size differs from the elements that reach it");
            for (uint32_t i = 0; i < n.size; ++i) {
                const uint32_t t = bits_get(bits, (uint64_t)n.start + i, 1);
                const int32_t c = n.child[t];
                if (c >= 0) elems[(size_t)c].push_back(E[i]);
                else LP[(size_t)(-c - 1)].push_back(E[i]);
            }
            std::vector<uint16_t>().swap(E);
        }
        for (size_t li = 0; li < T.leaves.size(); ++li)
            if (LP[li].size() != T.occ[li]) throw FormatError("wavelet leaf size differs from the elements that reach it");
    }

    // --- cells: rank(pos in block b, sym) with everything but the position-dependent part pre-evaluated
    VarReader R(S.var);
    std::vector<uint32_t> occ_next(nblk);
    for (size_t b = 0; b < nblk; ++b) occ_next[b] = (uint32_t)block_occ_base[b];
    for (int32_t sym = 0; sym < sigma; ++sym) {
        const int64_t sb_c = W.global_mapping[sb * (size_t)sigma + (size_t)sym];
        const uint64_t rank_sb = (uint64_t)(int64_t)W.sb_rank[sb * (size_t)sigma + (size_t)sym];
        const uint64_t rank_hb = (uint64_t)W.hyper_rank[(size_t)sym];
        auto cell_at = [&](size_t row) -> fmgpu::Cell8& { return F.cells[((size_t)P.first_block + row) * (size_t)sigma + (size_t)sym]; };
        if (sb_c >= sb_sigma) {  // :1040 symbol absent from the superblock
            for (size_t row = 0; row < P.rows; ++row) put_cell(cell_at(row), fmgpu::CELL_CONST, (uint32_t)(rank_hb + rank_sb));
            continue;
        }
        if (sb_c < 0) {
            for (size_t row = 0; row < P.rows; ++row) put_cell(cell_at(row), fmgpu::CELL_THROW, 0);
            continue;
        }
        // right-to-left sweep = the reference's "closest block to the right in which c occurs" scan (:1048-1069)
        int64_t next_present = -1;
        for (int64_t b = blocks_in_sb - 1; b >= 0; --b) {
            const int64_t mi = (sb_c << blocks_log) + b;
            const bool in_range = mi >= 0 && (size_t)mi < S.mapping.size();
            const int32_t block_c = in_range ? (int32_t)S.mapping[(size_t)mi] : 0;
            const bool absent = in_range && block_c == sigma - 1;
            if ((size_t)b < P.rows) {
                fmgpu::Cell8& cell = cell_at((size_t)b);
                if (!in_range) {
                    put_cell(cell, fmgpu::CELL_THROW, 0);
                } else if (absent) {
                    if (next_present < 0) {
                        uint64_t v;
                        if ((((int64_t)sb + 1) << 20) >= W.size) v = (uint64_t)W.count[(size_t)sym];
                        else v = rank_hb + (uint64_t)(int64_t)W.sb_rank[(sb + 1) * (size_t)sigma + (size_t)sym];
                        put_cell(cell, fmgpu::CELL_CONST, (uint32_t)v);
                    } else if ((size_t)next_present >= nblk) {
                        put_cell(cell, fmgpu::CELL_THROW, 0);
                    } else {
                        // :1071-1107 — rank taken from the later block's header with the CLAMPED code and without
                        // the treeHeight>0 guard (quirks Q3 and the run-block variant): restated literally.
                        const BlockHdr& H2 = S.blocks[(size_t)next_present];
                        const int64_t bc2 = S.mapping[(size_t)((sb_c << blocks_log) + next_present)];
                        const int64_t ptr = (int64_t)H2.var_off + ((int64_t)H2.tree_height - 1) * 4 + bc2 * 5 + 2;
                        R.ok = true;
                        const uint32_t r24 = R.u24(ptr);
                        if (!R.ok) put_cell(cell, fmgpu::CELL_THROW, 0);
                        else put_cell(cell, fmgpu::CELL_CONST, (uint32_t)(rank_hb + rank_sb + r24));
                    }
                } else if ((size_t)b >= nblk) {
                    put_cell(cell, fmgpu::CELL_THROW, 0);
                } else {
                    // :1113-1138 present: clamp repair, boundary rank
                    const BlockHdr& H = S.blocks[(size_t)b];
                    const BlockTree& T = trees[(size_t)b];
                    const int h = H.tree_height;
                    const int64_t tmp = (int64_t)H.var_off + (h > 0 ? (int64_t)(h - 1) * 4 : 0);
                    int64_t bc = block_c;
                    R.ok = true;
                    const uint32_t value = R.u16(tmp + 5 * bc);
                    if (value != (uint32_t)sym) ++bc;
                    const uint32_t rank_blk = R.u24(tmp + bc * 5 + 2);
                    if (!R.ok) {
                        put_cell(cell, fmgpu::CELL_THROW, 0);
                    } else if (h == 0) {
                        put_cell(cell, fmgpu::CELL_RUN, (uint32_t)(rank_hb + rank_sb + rank_blk));
                    } else if (bc < 0 || (size_t)bc >= T.leaves.size()) {
                        put_cell(cell, fmgpu::CELL_THROW, 0);
                    } else {
                        // the level walk of :1185-1279 along the leaf's code counts the elements among the first r positions
                        // of the block that reach the leaf: stored as the leaf's occurrence structure (layout.h)
                        const BlockTree::Leaf& L = T.leaves[(size_t)bc];
                        put_cell(cell, fmgpu::CELL_THROW, (uint32_t)(rank_hb + rank_sb + rank_blk));  // kind set by put_occ
                        const std::vector<uint16_t>& pos = leaf_pos[(size_t)b][(size_t)bc];
                        const uint32_t bsize = sb_block_size(W, sb, (size_t)b);
                        const uint32_t need = occ_records((uint32_t)pos.size(), bsize);
                        const uint32_t at = occ_next[(size_t)b];
                        if ((uint64_t)at + need > block_occ_base[(size_t)b] + T.n_occ) {
                            // two alphabet symbols mapped to the same leaf (a corrupt header): the block's occurrence records are
                            // sized for one structure per leaf
                            put_cell(cell, fmgpu::CELL_THROW, 0);
                        } else {
                            // (`need` is the upper bound the sizing pass reserved; range lists use less: compact_occ closes the gaps)
                            occ_next[(size_t)b] += put_occ(cell, pos, (uint32_t)L.len, bsize, at, F.occ);
                        }
                    }
                }
            }
            if (in_range && !absent) next_present = b;
        }
    }

    for (size_t b = 0; b < nblk; ++b) {
        F.occ_base[(size_t)P.first_block + b] = (uint32_t)block_occ_base[b];
        F.occ_used[(size_t)P.first_block + b] = trees[b].h != 0 ? occ_next[b] - (uint32_t)block_occ_base[b] : 0u;
    }

    // single-symbol blocks: the LF step needs rank(j, c') for the symbol c' inverseSelect decodes (low byte only) and
    // j inside the same block; that is the (block, c') cell, which never needs a record here — copy it into the descriptor
    for (size_t b = 0; b < nblk; ++b) {
        if (trees[b].h != 0) continue;
        Rec32& D = F.blocks[(size_t)P.first_block + b];
        const uint32_t c = (D.w[1] >> 8) & 0xffffu;
        if (c >= (uint32_t)sigma) {  // rank of a symbol outside the alphabet is 0 (:1018-1020)
            D.w[2] = 0;
            D.w[3] = fmgpu::CELL_CONST;
            continue;
        }
        const fmgpu::Cell8& cell = F.cells[((size_t)P.first_block + b) * (size_t)sigma + c];
        const uint32_t kind = cell.info >> fmgpu::CELL_KIND_SHIFT;
        if (kind != fmgpu::CELL_CONST && kind != fmgpu::CELL_RUN && kind != fmgpu::CELL_THROW)
            throw FormatError("single-symbol block with a tree walk");
        D.w[2] = cell.value;
        D.w[3] = kind;
    }
}

// The sizing pass reserves every (block, symbol) pair's occurrence records by an upper bound (a bit vector unless the pair has
// <= 15 occurrences); pairs that ended up as range lists use fewer.  Close the gaps: blocks keep their order, records move down,
// the cells' record pointers follow.
inline void compact_occ(FlatIndex& F, uint32_t sigma) {
    const size_t n_blocks = F.occ_base.size();
    uint32_t at = 0;
    for (size_t b = 0; b < n_blocks; ++b) {
        const uint32_t used = F.occ_used[b], from = F.occ_base[b];
        if (used == 0) continue;
        const uint32_t delta = from - at;
        if (delta) {
            memmove(&F.occ[at], &F.occ[from], (size_t)used * sizeof(Rec32));
            fmgpu::Cell8* row = &F.cells[b * (size_t)sigma];
            for (uint32_t c = 0; c < sigma; ++c)
                if ((row[c].info >> fmgpu::CELL_KIND_SHIFT) >= fmgpu::CELL_OCC_FIRST) row[c].info -= delta;
        }
        at += used;
    }
    F.occ.resize((size_t)at + 1);
    F.occ.back() = Rec32{};
    F.occ.shrink_to_fit();
    std::vector<uint32_t>().swap(F.occ_base);
    std::vector<uint32_t>().swap(F.occ_used);
}

inline void flatten_sampled(const RrrStream& r, FlatIndex& F) {
    const RrrTables& T = rrr_tables();
    const int64_t nblocks = ((int64_t)r.length + 14) / 15;
    const int64_t ngroups = (nblocks + fmgpu::SGROUP_BLOCKS - 1) / fmgpu::SGROUP_BLOCKS + 1;  // +1: rank at position == length
    F.sgroups.assign((size_t)ngroups, Rec32{});
    uint64_t bitpos = 0, ones = 0;
    for (int64_t g = 0; g < ngroups; ++g) {
        Rec32& G = F.sgroups[(size_t)g];
        G.w[0] = (uint32_t)ones;
        G.w[1] = (uint32_t)bitpos;
        uint32_t sub_bits = 0, sub_ones = 0, gb = 0, go = 0;
        for (uint32_t k = 0; k < fmgpu::SGROUP_BLOCKS; ++k) {
            const int64_t b = g * (int64_t)fmgpu::SGROUP_BLOCKS + k;
            if (k && (k % 8) == 0) {
                sub_bits |= gb << (10 * (k / 8 - 1));
                sub_ones |= go << (10 * (k / 8 - 1));
            }
            uint32_t cls = 0;
            if (b < nblocks) {
                cls = (uint32_t)r.classes.get(b);
                gb += T.bits_needed[cls];
                go += cls;
            }
            G.w[4 + k / 8] |= cls << (4 * (k % 8));
        }
        G.w[2] = sub_bits;
        G.w[3] = sub_ones;
        bitpos += gb;
        ones += go;
        if (bitpos > 0xffffffffULL) throw FormatError("sampled-row offset stream exceeds 2^32 bits");
    }
    if (bitpos > (uint64_t)(r.offsets.size() - 1) * 64) throw FormatError("RrrVector offset stream shorter than its classes need");
    if (ones != (uint64_t)(uint32_t)r.total_ones) throw FormatError("RrrVector totalOnes differs from the sum of its block classes");
    // offset stream verbatim (LSB-first 64-bit words == little-endian 32-bit word pairs), padded
    const size_t nw64 = r.offsets.size();
    F.soffsets.assign(nw64 * 2 + 16, 0);
    memcpy(F.soffsets.data(), r.offsets.data(), nw64 * 8);
}

// `limit`: every sample must lie below it (the LF kernels use inverse-SA samples as rows, SA samples are results)
inline void unpack_samples(const PackedInts& v, std::vector<Rec32>& out, uint64_t limit, const char* what) {
    const size_t n = (size_t)v.length;
    out.assign(n / 8 + 1, Rec32{});
    for (size_t i = 0; i < n; ++i) {
        const uint64_t x = (uint64_t)v.get((int64_t)i);
        if (x >= limit) throw FormatError(std::string(what) + " sample outside the index");
        out[i / 8].w[i % 8] = (uint32_t)x;
    }
}

// budget for the dense (block x symbol) cell table; beyond it the index is rejected for now
constexpr uint64_t MAX_CELL_BYTES = 48ULL << 30;

// wavelet_only: the stream held a bare WaveletFixedBlockBoosting (fm.wf filled, everything else empty): only the wavelet
// records are produced (rank / inverseSelect entry points)
inline void flatten(const FmStream& fm, int threads, FlatIndex& F, bool wavelet_only = false) {
    const WfbbStream& W = fm.wf;
    fmgpu::DevIndex& M = F.meta;
    M.length = (uint32_t)fm.length;
    M.sample_rate = (uint32_t)fm.sample_rate;
    M.n_c = (uint32_t)fm.C.size();
    M.n_lookup = (uint32_t)fm.lookup.size();
    M.sigma = (uint32_t)W.sigma;
    M.n_sb = (uint32_t)W.sbs.size();
    M.q4 = (W.size % (1LL << 20)) == 0 ? 1u : 0u;
    M.extract_enabled = fm.enable_extract ? 1u : 0u;
    M.n_isa = fm.enable_extract ? (uint32_t)fm.positions.length : 0u;
    M.n_sa = (uint32_t)fm.suffixes.length;
    M.s_total_ones = (uint32_t)fm.sampled.total_ones;
    F.alphabet_length = (int32_t)fm.map.size();

    F.C.assign(fm.C.begin(), fm.C.end());
    F.char2code.assign(65536, 0);
    for (const auto& kv : fm.map)
        if (kv.first >= 0 && kv.first < 65536) F.char2code[(size_t)kv.first] = (uint16_t)kv.second;
    F.code2char.resize(fm.lookup.size());
    for (size_t i = 0; i < fm.lookup.size(); ++i) F.code2char[i] = (uint16_t)fm.lookup[i];
    for (const auto& kv : fm.map)
        if (kv.second < 0 || (size_t)kv.second + 1 >= fm.C.size()) throw FormatError("alphabet code outside cumulativeCounts");
    if (!wavelet_only) {
        // cumulativeCounts (fm/FmIndex.java:307-327): SA ranges of the symbol codes — non-decreasing, inside [0, length], one
        // entry beyond the last wavelet symbol (the kernels form C[sym] and C[sym + 1] for every symbol the wavelet can return)
        if (fm.C.size() < (size_t)W.sigma + 1) throw FormatError("cumulativeCounts shorter than the wavelet alphabet");
        for (size_t i = 0; i < fm.C.size(); ++i)
            if (fm.C[i] < 0 || (int64_t)fm.C[i] > (int64_t)fm.length || (i && fm.C[i] < fm.C[i - 1]))
                throw FormatError("cumulativeCounts not a non-decreasing sequence inside the index");
    }

    // pass 1: sizes
    const size_t nsb = W.sbs.size();
    std::vector<SbPlan> plan(nsb);
    parallel_sbs(nsb, threads, [&](size_t sb) {
        const SuperBlockHdr& S = W.sbs[sb];
        const int64_t bs = 1LL << S.block_size_log;
        const int64_t sb_size = std::min<int64_t>(1LL << 20, W.size - ((int64_t)sb << 20));
        const size_t expect = (size_t)((sb_size + bs - 1) / bs);
        if (S.blocks.size() != expect) throw FormatError("block count does not match superblock size");
        SbPlan& P = plan[sb];
        P.rows = (uint32_t)expect;
        if (sb + 1 == nsb && (sb_size % bs) == 0 && sb_size != (1LL << 20)) P.rows += 1;  // row for position == size
        BlockTree T;
        for (size_t b = 0; b < S.blocks.size(); ++b) {
            build_block_tree(S, b, sb_block_size(W, sb, b), T, W.sigma);
            P.n_sectors += T.n_sectors;
            P.n_nodes += T.n_even;
            P.n_occ += T.n_occ;
        }
    });
    uint64_t blocks_total = 0, sectors_total = 0, nodes_total = 0, occ_total = 0;
    F.sb.assign(nsb + 1, fmgpu::SbDesc{0, 16});  // +1: a lane may form the (unused) descriptor address of position == length
    for (size_t sb = 0; sb < nsb; ++sb) {
        SbPlan& P = plan[sb];
        P.first_block = (uint32_t)blocks_total;
        P.sector_base = sectors_total;
        P.node_base = nodes_total;
        P.occ_base = occ_total;
        blocks_total += P.rows;
        sectors_total += P.n_sectors;
        nodes_total += P.n_nodes;
        occ_total += P.n_occ;
        F.sb[sb].first_block = P.first_block;
        F.sb[sb].block_log = (uint32_t)W.sbs[sb].block_size_log;
    }
    if (sectors_total >= 0xffffffffULL || nodes_total >= 0x7fffffffULL || occ_total >= 0xffffffffULL || blocks_total >= 0xffffffffULL)
        throw FormatError("index too large for 32-bit record indices");
    const uint64_t cell_bytes = blocks_total * (uint64_t)W.sigma * sizeof(fmgpu::Cell8);
    if (cell_bytes > MAX_CELL_BYTES)
        throw FormatError("alphabet x block count too large for the dense cell table (" + std::to_string(cell_bytes >> 20) + " MiB)");
    if (blocks_total * (uint64_t)W.sigma >= 0xffffffffULL) throw FormatError("more than 2^32 (block, symbol) cells");
    F.cells.assign((size_t)(blocks_total * (uint64_t)W.sigma), fmgpu::Cell8{0u, (uint32_t)fmgpu::CELL_THROW << fmgpu::CELL_KIND_SHIFT});
    F.sectors.assign((size_t)sectors_total + 1, Rec32{});
    F.nodes.assign((size_t)nodes_total + 1, Rec32{});
    F.occ.assign((size_t)occ_total + 1, Rec32{});
    F.blocks.assign((size_t)blocks_total + 1, Rec32{});
    F.occ_base.assign((size_t)blocks_total, 0);
    F.occ_used.assign((size_t)blocks_total, 0);

    // pass 2: fill
    parallel_sbs(nsb, threads, [&](size_t sb) { flatten_superblock(W, sb, plan[sb], F); });
    compact_occ(F, (uint32_t)W.sigma);

    if (wavelet_only) {
        F.sgroups.assign(1, Rec32{});
        F.soffsets.assign(16, 0);
        F.sa.assign(1, Rec32{});
        F.isa.assign(1, Rec32{});
        return;
    }
    flatten_sampled(fm.sampled, F);
    if ((uint64_t)(uint32_t)fm.sampled.total_ones > (uint64_t)fm.suffixes.length) throw FormatError("more sampled rows than suffix-array samples");
    // the sampled structures must cover the text the way the constructor lays them out (fm/FmIndex.java:343-370): one SA sample per
    // sampleRate text positions, length / sampleRate + 2 inverse-SA samples (the LF kernels index them without further checks)
    if (fm.sample_rate < 1) throw FormatError("sampleRate < 1");
    if ((int64_t)fm.suffixes.length < ((int64_t)fm.length - 1) / fm.sample_rate + 1) throw FormatError("fewer suffix-array samples than length / sampleRate");
    if (fm.enable_extract && (int64_t)fm.positions.length != (int64_t)fm.length / fm.sample_rate + 2)
        throw FormatError("inverse suffix-array samples: expected length / sampleRate + 2 entries");
    if ((int64_t)fm.sampled.length != (int64_t)fm.length) throw FormatError("sampled-row vector length differs from the index length");
    unpack_samples(fm.suffixes, F.sa, (uint64_t)fm.length, "suffix-array");
    if (fm.enable_extract) unpack_samples(fm.positions, F.isa, (uint64_t)fm.length, "inverse suffix-array");
    else F.isa.assign(1, Rec32{});
}

// a bare RrrVector stream: only the group records / offset stream of the rank / access kernels
inline void flatten_rrr(const RrrStream& r, FlatIndex& F) {
    fmgpu::DevIndex& M = F.meta;
    M = fmgpu::DevIndex{};
    M.length = (uint32_t)r.length;
    M.sample_rate = 1;
    M.s_total_ones = (uint32_t)r.total_ones;
    F.C.assign(2, 0);
    F.char2code.assign(65536, 0);
    F.code2char.assign(1, 0);
    F.sb.assign(1, fmgpu::SbDesc{0, 16});
    F.cells.assign(1, fmgpu::Cell8{0u, 0u});
    for (auto* v : {&F.sectors, &F.occ, &F.blocks, &F.nodes, &F.sa, &F.isa}) v->assign(1, Rec32{});
    flatten_sampled(r, F);
}

}  // namespace fmgpu_host
