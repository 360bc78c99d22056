// Host-side ENCODER for the fixed-block-boosting wavelet structure (index producer only; not on
// the GPU query path and not the oracle).
//
// It emits the same fields the reference's WaveletFixedBlockBoosting serializes
// (indices/src/main/java/com/dynatrace/wavelet/WaveletFixedBlockBoosting.java:1544-1570,
// SuperBlockHeaderItem :1651-1667, BlockHeaderItem :1607-1613) and follows the reference's
// structural decisions so that the produced index is one the Java reader accepts:
//   * superblocks of 2^20 symbols, one block size 2^9..2^16 per superblock chosen by the size
//     estimate of encodeSuperBlock (:853-987);
//   * per block a canonical-Huffman-shaped wavelet tree; level bitvectors concatenated in BFS
//     order into one bitvector per superblock (:602-709), RRR-coded (:534);
//   * byte-packed variable block header with the three sections of encodeBlock (:742-809);
//   * mapping rows clamped to sigma-2 with sigma-1 meaning "absent" (:457-472).
// Huffman ties: the reference's PriorityQueue comparator (:1684-1707) orders by (weight, first
// symbol of the merged list); symbol sets are disjoint so that order is total and independent of
// heap internals — we reproduce it with an explicit (weight, first) key.
// The implementation is our own (level-wise stable partition instead of per-bit lists, threads
// over superblocks).
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <queue>
#include <thread>
#include <vector>

#include "packed.hpp"
#include "rrr_enc.hpp"

namespace fmhost {

static const int SB_LOG = 20;
static const int64_t SB_SIZE = 1LL << SB_LOG;

struct BlockHeader {
    int32_t bv_rank = 0, bv_offset = 0, var_off = 0;
    int16_t sigma_m1 = 0, tree_height = 0;
};

struct SuperBlock {
    int16_t sigma_m1 = 0;
    int16_t block_size_log = 0;
    RrrEnc rank;
    std::vector<BlockHeader> blocks;
    std::vector<uint8_t> var;
    std::vector<int16_t> mapping;
};

struct WfbbEnc {
    int64_t size = 0;
    int32_t sigma = 0;
    int32_t rrr_rate = 0;
    std::vector<int64_t> count, hyper_rank;
    std::vector<int32_t> sb_rank;
    std::vector<int16_t> global_mapping;
    std::vector<SuperBlock> sbs;
};

namespace wfbb_detail {

struct HuffNode {
    uint64_t w;
    int32_t first;
    int32_t left, right;  // -1 for leaves
};
struct HuffCmp {
    const std::vector<HuffNode>* nodes;
    bool operator()(int a, int b) const {  // priority_queue is a max-heap: invert
        const HuffNode& x = (*nodes)[a];
        const HuffNode& y = (*nodes)[b];
        if (x.w != y.w) return x.w > y.w;
        return x.first > y.first;
    }
};

// syms: present symbols (ascending), freq[sym] their counts.  Writes len[sym]; returns max len.
struct Huffman {
    std::vector<HuffNode> nodes;
    std::vector<int> stack, depth;
    int run(const std::vector<int32_t>& syms, const uint32_t* freq, uint8_t* len) {
        nodes.clear();
        for (int32_t s : syms) {
            nodes.push_back({freq[s], s, -1, -1});
            len[s] = 0;
        }
        if (syms.size() < 2) return 0;
        HuffCmp cmp{&nodes};
        std::priority_queue<int, std::vector<int>, HuffCmp> pq(cmp);
        for (int i = 0; i < (int)syms.size(); ++i) pq.push(i);
        while (pq.size() > 1) {
            int x = pq.top();
            pq.pop();
            int y = pq.top();
            pq.pop();
            nodes.push_back({nodes[x].w + nodes[y].w, nodes[x].first, x, y});
            pq.push((int)nodes.size() - 1);
        }
        // depth of every leaf
        int maxlen = 0;
        stack.clear();
        std::vector<int>& st = stack;
        depth.assign(nodes.size(), 0);
        st.push_back((int)nodes.size() - 1);
        while (!st.empty()) {
            int v = st.back();
            st.pop_back();
            if (nodes[v].left < 0) {
                len[nodes[v].first] = (uint8_t)depth[v];
                if (depth[v] > maxlen) maxlen = depth[v];
            } else {
                depth[nodes[v].left] = depth[v] + 1;
                depth[nodes[v].right] = depth[v] + 1;
                st.push_back(nodes[v].left);
                st.push_back(nodes[v].right);
            }
        }
        return maxlen;
    }
};

inline void put16(std::vector<uint8_t>& v, size_t& p, uint32_t x) {
    v[p++] = (uint8_t)(x & 0xff);
    v[p++] = (uint8_t)((x >> 8) & 0xff);
}

struct SbEncoder {
    const uint16_t* text;  // whole BWT
    int64_t size;
    int32_t sigma;
    int32_t rrr_rate;
    Huffman huff;
    std::vector<uint32_t> freq_small;  // [block][sigma] at the smallest block size, merged upward
    std::vector<uint8_t> len;
    std::vector<int32_t> present;

    void encode(int64_t sb, const int64_t* count_at_start, WfbbEnc& out) {
        SuperBlock& S = out.sbs[sb];
        const int64_t sb_beg = sb * SB_SIZE;
        const int64_t sb_end = std::min(sb_beg + SB_SIZE, size);
        const int64_t sb_size = sb_end - sb_beg;
        const uint16_t* t = text + sb_beg;
        len.assign(sigma, 0);

        // ranks at the superblock boundary, superblock alphabet (:827-851)
        std::vector<uint32_t> sbfreq(sigma, 0);
        for (int64_t i = 0; i < sb_size; ++i) sbfreq[t[i]]++;
        int32_t sb_sigma = 0;
        for (int32_t c = 0; c < sigma; ++c) {
            out.sb_rank[sb * sigma + c] = (int32_t)count_at_start[c];  // single hyperblock: rank 0
            if (sbfreq[c]) out.global_mapping[sb * sigma + c] = (int16_t)sb_sigma++;
        }
        S.sigma_m1 = (int16_t)(sb_sigma - 1);

        // block-size search (:853-987)
        const int smallest_log = 9;
        const int64_t max_blocks = SB_SIZE >> smallest_log;
        freq_small.assign((size_t)max_blocks * sigma, 0);
        int best_log = 0;
        int64_t best_size = 0;
        int64_t compressed_bv = 0, prev_uncompressed = 0;
        for (int bl = smallest_log; bl <= 16; ++bl) {
            const int64_t bs = 1LL << bl;
            const int64_t nblk = (sb_size + bs - 1) / bs;
            int64_t enc = 14 * nblk + (int64_t)sb_sigma * (SB_SIZE / bs);
            if (bl == smallest_log) {
                for (int64_t i = 0; i < sb_size; ++i) freq_small[(size_t)(i >> bl) * sigma + t[i]]++;
            } else {
                const int64_t prev_nblk = (sb_size + (bs / 2) - 1) / (bs / 2);
                for (int64_t b = 0; b < prev_nblk; b += 2) {
                    uint32_t* dst = &freq_small[(size_t)(b >> 1) * sigma];
                    const uint32_t* a = &freq_small[(size_t)b * sigma];
                    const uint32_t* c2 = (b + 1 < prev_nblk) ? &freq_small[(size_t)(b + 1) * sigma] : nullptr;
                    for (int32_t c = 0; c < sigma; ++c) dst[c] = a[c] + (c2 ? c2[c] : 0);
                }
            }
            int64_t uncompressed = 0;
            for (int64_t b = 0; b < nblk; ++b) {
                const uint32_t* f = &freq_small[(size_t)b * sigma];
                present.clear();
                for (int32_t c = 0; c < sigma; ++c)
                    if (f[c]) present.push_back(c);
                const int64_t bsig = (int64_t)present.size();
                enc += bsig * 4 + (bsig - 1) * 2;
                int maxlen = huff.run(present, f, len.data());
                if (maxlen > 1) enc += (int64_t)(maxlen - 1) * 3;
                for (int32_t c : present) uncompressed += (int64_t)f[c] * len[c];
            }
            if (uncompressed > 0) {
                if (bl == smallest_log) {
                    compressed_bv = RrrEnc::zero_vector_memory(uncompressed, rrr_rate);
                } else if (prev_uncompressed > 0) {
                    double scale = (double)uncompressed / (double)prev_uncompressed;
                    compressed_bv = (int64_t)((double)compressed_bv * scale);
                } else {
                    compressed_bv = 0;  // Java: 0 * (x/0.0) = NaN -> (long) NaN == 0
                }
                enc += compressed_bv;
            }
            prev_uncompressed = uncompressed;
            if (bl == smallest_log || enc < best_size) {
                best_log = bl;
                best_size = enc;
            }
        }
        encode_blocks(sb, t, sb_size, best_log, out);
    }

    void encode_blocks(int64_t sb, const uint16_t* t, int64_t sb_size, int bl, WfbbEnc& out) {
        SuperBlock& S = out.sbs[sb];
        const int64_t bs = 1LL << bl;
        const int64_t blocks_per_sb = SB_SIZE / bs;
        const int64_t nblk = (sb_size + bs - 1) / bs;
        const int32_t sb_sigma = (int32_t)S.sigma_m1 + 1;
        S.block_size_log = (int16_t)bl;
        S.mapping.assign((size_t)sb_sigma * blocks_per_sb, (int16_t)(sigma - 1));
        S.blocks.assign((size_t)nblk, BlockHeader());

        struct BlockPlan {
            std::vector<int32_t> order;  // symbols sorted by (len, sym)
            std::vector<uint8_t> lens;   // parallel to order
            int maxlen;
        };
        std::vector<BlockPlan> plans((size_t)nblk);
        std::vector<uint32_t> f(sigma, 0);
        int64_t bv_size_total = 0, var_total = 0;
        for (int64_t b = 0; b < nblk; ++b) {
            const int64_t beg = b * bs, end = std::min(beg + bs, sb_size);
            present.clear();
            for (int64_t i = beg; i < end; ++i)
                if (f[t[i]]++ == 0) present.push_back(t[i]);
            std::sort(present.begin(), present.end());
            BlockPlan& P = plans[b];
            P.maxlen = huff.run(present, f.data(), len.data());
            P.order = present;
            std::sort(P.order.begin(), P.order.end(), [&](int32_t a, int32_t c) {
                if (len[a] != len[c]) return len[a] < len[c];
                return a < c;
            });
            P.lens.resize(P.order.size());
            int64_t bvsz = 0;
            const int64_t bsig = (int64_t)P.order.size();
            for (int64_t i = 0; i < bsig; ++i) {
                int32_t s = P.order[i];
                P.lens[i] = len[s];
                if (bsig > 1) bvsz += (int64_t)f[s] * len[s];
                int16_t sbc = out.global_mapping[sb * sigma + s];
                int64_t clamped = std::min<int64_t>(sigma - 2, i);
                S.mapping[(size_t)sbc * blocks_per_sb + b] = (int16_t)clamped;
            }
            BlockHeader& H = S.blocks[b];
            H.bv_offset = (int32_t)bv_size_total;
            H.var_off = (int32_t)var_total;
            H.tree_height = (int16_t)P.maxlen;
            H.sigma_m1 = (int16_t)(bsig - 1);
            bv_size_total += bvsz;
            if (P.maxlen > 1) var_total += (int64_t)(P.maxlen - 1) * 4;
            var_total += bsig * 5 + (bsig - 1) * 2;
            for (int32_t s : present) f[s] = 0;
        }
        S.var.assign((size_t)var_total, 0);

        BitString bv;
        bv.resize((uint64_t)bv_size_total);
        std::vector<int64_t> block_rank(sigma, 0);
        std::vector<uint32_t> local_of(sigma, 0), code_of_local, cur, nxt;
        std::vector<uint32_t> ones_per_node;
        int64_t bv_rank = 0;
        for (int64_t b = 0; b < nblk; ++b) {
            const int64_t beg = b * bs, end = std::min(beg + bs, sb_size);
            BlockPlan& P = plans[b];
            BlockHeader& H = S.blocks[b];
            const int bsig = (int)P.order.size();
            const int maxlen = P.maxlen;
            // canonical codes (:537-555)
            code_of_local.assign(bsig, 0);
            {
                uint32_t c = 0;
                for (int i = 0; i < bsig; ++i) {
                    if (i != 0) c = (c + 1) << (P.lens[i] - P.lens[i - 1]);
                    code_of_local[i] = c;
                    local_of[P.order[i]] = (uint32_t)i;
                }
            }
            ones_per_node.clear();
            int64_t ones_total = 0;
            uint64_t wpos = (uint64_t)H.bv_offset;
            if (bsig > 1) {
                cur.resize((size_t)(end - beg));
                for (int64_t i = beg; i < end; ++i) cur[(size_t)(i - beg)] = local_of[t[i]];
                for (int d = 0; d < maxlen && !cur.empty(); ++d) {
                    nxt.clear();
                    size_t i = 0;
                    while (i < cur.size()) {
                        // one node = maximal run of elements sharing the d-bit code prefix
                        const uint32_t li = cur[i];
                        const uint32_t prefix = code_of_local[li] >> (P.lens[li] - d);
                        size_t j = i;
                        uint32_t ones = 0;
                        while (j < cur.size()) {
                            const uint32_t lj = cur[j];
                            if ((code_of_local[lj] >> (P.lens[lj] - d)) != prefix) break;
                            const uint32_t bit = (code_of_local[lj] >> (P.lens[lj] - d - 1)) & 1u;
                            if (bit) {
                                bv.set1(wpos);
                                ++ones;
                            } else if (P.lens[lj] > d + 1) {
                                nxt.push_back(lj);
                            }
                            ++wpos;
                            ++j;
                        }
                        for (size_t k = i; k < j; ++k) {
                            const uint32_t lk = cur[k];
                            if (((code_of_local[lk] >> (P.lens[lk] - d - 1)) & 1u) && P.lens[lk] > d + 1)
                                nxt.push_back(lk);
                        }
                        ones_per_node.push_back(ones);
                        ones_total += ones;
                        i = j;
                    }
                    cur.swap(nxt);
                }
            }
            // variable-size header (:742-809)
            size_t p = (size_t)H.var_off;
            std::vector<uint32_t> leaves_at(maxlen + 1, 0), level_total(maxlen + 1, 0);
            for (int i = 0; i < bsig; ++i) {
                const int L = P.lens[i];
                if (L < maxlen) leaves_at[L]++;
            }
            {
                // level_total[d] = total frequency of symbols with code length > d
                std::vector<uint32_t> fcount(bsig, 0);
                for (int64_t i = beg; i < end; ++i) fcount[local_of[t[i]]]++;
                for (int i = 0; i < bsig; ++i)
                    for (int d = 1; d < P.lens[i]; ++d) level_total[d] += fcount[i];
            }
            for (int d = 1; d < maxlen; ++d) {
                put16(S.var, p, leaves_at[d] & 0xffff);
                put16(S.var, p, (level_total[d] - 1) & 0xffff);
            }
            for (int i = 0; i < bsig; ++i) {
                const int32_t s = P.order[i];
                const int64_t r = block_rank[s];
                put16(S.var, p, (uint32_t)s & 0xffff);
                S.var[p++] = (uint8_t)(r & 0xff);
                S.var[p++] = (uint8_t)((r >> 8) & 0xff);
                S.var[p++] = (uint8_t)((r >> 16) & 0xff);
            }
            {
                int64_t nodes_this_level = 1;
                size_t ptr = 0;
                for (int d = 0; d < maxlen; ++d) {
                    uint32_t cum = 0;
                    for (int64_t j = 0; j < nodes_this_level; ++j) {
                        cum += ones_per_node[ptr++];
                        put16(S.var, p, cum & 0xffff);
                    }
                    if (d + 1 != maxlen) {
                        nodes_this_level = (nodes_this_level << 1) - (int64_t)leaves_at[d + 1];
                    }
                }
            }
            H.bv_rank = (int32_t)bv_rank;
            bv_rank += ones_total;
            for (int64_t i = beg; i < end; ++i) block_rank[t[i]]++;
        }
        S.rank.encode(bv, rrr_rate);
    }
};

}  // namespace wfbb_detail

// text: symbols in [0, sigma); sigma = max symbol + 1 (WaveletFixedBlockBoosting.java:130-154).
inline void wfbb_encode(const uint16_t* text, int64_t size, int32_t sigma, int32_t rrr_rate,
                        int threads, WfbbEnc& out) {
    out.size = size;
    out.sigma = sigma;
    out.rrr_rate = rrr_rate;
    const int64_t nsb = (size + SB_SIZE - 1) / SB_SIZE;
    const int64_t nhb = (size + (1LL << 32) - 1) >> 32;
    out.count.assign(sigma, 0);
    out.hyper_rank.assign((size_t)nhb * sigma, 0);  // size < 2^32: one hyperblock, ranks 0
    out.sb_rank.assign((size_t)nsb * sigma, 0);
    out.global_mapping.assign((size_t)nsb * sigma, (int16_t)(sigma - 1));
    out.sbs.assign((size_t)nsb, SuperBlock());

    // symbol counts before each superblock
    std::vector<int64_t> start_counts((size_t)(nsb + 1) * sigma, 0);
    for (int64_t sb = 0; sb < nsb; ++sb) {
        int64_t* nxt = &start_counts[(size_t)(sb + 1) * sigma];
        const int64_t* cur = &start_counts[(size_t)sb * sigma];
        std::memcpy(nxt, cur, sizeof(int64_t) * sigma);
        const int64_t beg = sb * SB_SIZE, end = std::min(beg + SB_SIZE, size);
        for (int64_t i = beg; i < end; ++i) nxt[text[i]]++;
    }
    for (int32_t c = 0; c < sigma; ++c) out.count[c] = start_counts[(size_t)nsb * sigma + c];

    if (threads < 1) threads = 1;
    if (threads > nsb) threads = (int)std::max<int64_t>(1, nsb);
    std::atomic<int64_t> next(0);
    auto worker = [&]() {
        wfbb_detail::SbEncoder E;
        E.text = text;
        E.size = size;
        E.sigma = sigma;
        E.rrr_rate = rrr_rate;
        for (;;) {
            int64_t sb = next.fetch_add(1);
            if (sb >= nsb) break;
            E.encode(sb, &start_counts[(size_t)sb * sigma], out);
        }
    };
    std::vector<std::thread> pool;
    for (int i = 1; i < threads; ++i) pool.emplace_back(worker);
    worker();
    for (auto& th : pool) th.join();
}

}  // namespace fmhost
