// Per-lane code of one backward-search step (FmIndex.count, fm/FmIndex.java:464-471): the two rank queries
// rank(start, c) and rank(end, c) of WaveletFixedBlockBoosting.rank (wavelet/WaveletFixedBlockBoosting.java:1010-1285)
// run as TWO TRACKS of one fused walk.  Host/device code: k_count (kernels.cuh) runs it per lane, the host layout test
// (tests/support/flatcheck.cpp) replays it against the CPU oracle.
//
// Track B carries `end`, track A carries `start` (off when start == 0, rank(0, .) == 0, :1012).  Per step a lane issues
//   1. the (block, symbol) cell of each track (one load when both positions lie in the same block), and at the same time
//   2. SPECULATIVELY the root level record of each track: its address needs no memory access (root-record directory,
//      layout.h), so the DRAM access of the first two tree levels overlaps the cell fetch instead of following it;
//   3. per further two levels one record per track (one when both tracks read the same record).
// All loads of a stage are issued before any is used, so a lane has up to 4 records in flight and the dependent chain of a
// step is max(cell, root record) + (code length / 2 - 1) records, whatever the two tracks' blocks are.  (Before v5 the start
// track of a step whose positions straddle a block boundary was a second, serial walk: 34 % of the steps of the
// configs[1] workload.)
#pragma once
#include <cstdint>

#include "lane_logic.h"

// 1 = speculative root fetch (experiment knob, tools/gpu_v5_cycle.sh).  Measured on B200 (profiles/experiments): the fused
// two-track walk alone runs the configs[1] batch at 895 M patterns/s, with the speculative fetch on top 790-826 M/s although
// only 1.65 M of the ~45 M speculative loads per launch are wasted: k_count is bound by issued instructions at 14/32 lane
// utilisation, not by the cell -> record latency chain, and the directory lookups add instructions.  Default off.
#ifndef COUNT_SPEC_ROOT
#define COUNT_SPEC_ROOT 0
#endif

// a record variable that is only read after a (conditional) load: left uninitialised on the device, zeroed on the host
#if defined(__CUDA_ARCH__)
#define FMGPU_UNSET
#else
#define FMGPU_UNSET {}
#endif

namespace fmgpu {

using CountTables = SmemTables;  // C, superblock descriptors, root-record directory

struct CountCounters {
    uint32_t ranks, levels, loads, recs, spec_wasted;
};

// --- q-gram start table (layout.h) ---------------------------------------------------------------------------------
// PatDesc.last of a pattern whose chars end at chars[b - 1]: the code of its last char, or — when the index has a q-gram
// start table, the pattern has at least q chars and its last q chars are all in the wavelet alphabet — the table index
FMGPU_HD uint32_t pattern_start(const uint16_t* chars, uint64_t b, uint32_t len, const uint16_t* char2code, uint32_t kmer_q,
                                uint32_t kmer_stride, uint32_t sigma) {
    const uint32_t last = (uint32_t)FMGPU_LDG16(char2code + chars[b - 1]);
    if (kmer_q < 2u || len < kmer_q || last == 0u || last >= sigma) return last;
    uint32_t idx = last;
    for (uint32_t k = 1; k < kmer_q; ++k) {
        const uint32_t c = (uint32_t)FMGPU_LDG16(char2code + chars[b - 1 - k]);
        if (c == 0u || c >= sigma) return last;
        idx = idx * kmer_stride + c;
    }
    return PAT_KMER | idx;
}
// the start state of a pattern whose descriptor carries a table index: true and {*sp, *ep} = the SA range after its last q chars
// (FmIndex.count after q - 1 steps, fm/FmIndex.java:455-474), or false when the entry is not usable
FMGPU_HD bool start_table_lookup(const DevIndex& ix, uint32_t last, uint32_t* sp, uint32_t* ep) {
    const U32x2 r = ix.kmer[last & ~PAT_KMER];
    if (r.x == 0xffffffffu) return false;
    *sp = r.x;
    *ep = r.y;
    return true;
}

struct Track {
    uint32_t base, code, len, r;  // boundary rank; Huffman code left-aligned (bit 31 = next decision); levels left; position
};

FMGPU_HD void track_open(Track& t, const Rec32& cell, uint32_t r, bool on, uint32_t* err) {
    const uint32_t kind = (cell.w[2] >> 8) & 0xffu;
    const uint32_t L = cell.w[2] & 0xffu;
    t.base = cell.w[0];
    t.len = (on && kind == CELL_NORMAL) ? L : 0u;
    t.code = t.len ? cell.w[1] << (32u - t.len) : 0u;
    // CONST: the cell is the answer; RUN: boundary rank + position inside the single-symbol block (:1141-1146)
    t.r = (kind == CELL_CONST || kind == CELL_THROW) ? 0u : r;
    if (on && kind == CELL_THROW) *err = 1u;
}

// the two levels of one record for one track
FMGPU_HD void track_levels(Track& t, const Rec32& x) {
    // a code that ends at the record's first level has u == 0: the left-aligned code shifts in zeros
    t.r = dlevel_rank(x, t.r % SECTOR_BITS, t.code >> 31, (t.code >> 30) & 1u);
    t.code <<= 2;
    t.len = t.len >= 2u ? t.len - 2u : 0u;
}

// One backward-search step for the lane: on return *sp / *ep hold rank(start, c) / rank(end, c) (NOT yet offset by C[c]).
// Returns 1 where the reference throws (THROW cells).  `on` = the lane takes part in this step.
template <bool STATS>
FMGPU_HD uint32_t count_step(const DevIndex& ix, const CountTables& T, uint32_t c, uint32_t* sp, uint32_t* ep, bool on, CountCounters& cnt) {
    if (!on) return 0u;
    const uint32_t s = *sp, e = *ep;
    const bool on_a = s != 0u;
    const uint32_t sbi_b = e >> SB_LOG, sbi_a = s >> SB_LOG;
    const SbDesc db = T.sb[sbi_b];
    const SbDesc da = T.sb[sbi_a];
    const uint32_t blk_b = db.first_block + ((e & SB_MASK) >> db.block_log);
    const uint32_t blk_a = da.first_block + ((s & SB_MASK) >> da.block_log);
    const uint32_t rb = e & ((1u << db.block_log) - 1u);
    const uint32_t ra = s & ((1u << da.block_log) - 1u);
    const bool split = on_a && blk_a != blk_b;
    if (STATS) cnt.ranks += on_a ? 2u : 1u;

    // stage 0: cells and speculative root records, all issued before any is consumed
    uint32_t root_b = 0, root_a = 0;
#if COUNT_SPEC_ROOT
    const bool tree_b = root_record(T, sbi_b, blk_b, &root_b);
    const bool tree_a = on_a && root_record(T, sbi_a, blk_a, &root_a);
#else
    const bool tree_b = false, tree_a = false;
    (void)sbi_a;
#endif
    const Rec32* pb = ix.sectors + (root_b + rb / SECTOR_BITS);
    const Rec32* pa = ix.sectors + (root_a + ra / SECTOR_BITS);
    const bool ld_a = tree_a && !(tree_b && pa == pb);
    const Rec32 cell_b = FMGPU_LD256(ix.cells + ((uint64_t)blk_b * ix.sigma + c));
    Rec32 cell_a = cell_b;
    if (split) cell_a = FMGPU_LD256(ix.cells + ((uint64_t)blk_a * ix.sigma + c));
    Rec32 xb FMGPU_UNSET, xa FMGPU_UNSET;  // never interpreted unless loaded
    if (tree_b) xb = FMGPU_LD256(pb);
    if (ld_a) xa = FMGPU_LD256(pa);
    if (STATS) cnt.loads += (split ? 2u : 1u) + (tree_b ? 1u : 0u) + (ld_a ? 1u : 0u);

    uint32_t err = 0;
    Track A, B;
    track_open(B, cell_b, rb, true, &err);
    track_open(A, cell_a, ra, on_a, &err);
    const uint32_t pairs_b = (B.len + 1u) >> 1, pairs_a = (A.len + 1u) >> 1;
    if (STATS) {
        cnt.levels += A.len + B.len;
        cnt.recs += pairs_a + pairs_b;
        cnt.spec_wasted += (tree_b && !B.len ? 1u : 0u) + (ld_a && !A.len ? 1u : 0u);
    }

    // levels 0-1: the speculative root records (a NORMAL cell's first record is its block's root: cell.w[3] == root)
#if COUNT_SPEC_ROOT
    if (B.len) track_levels(B, xb);
    if (A.len) track_levels(A, ld_a ? xa : xb);
#endif

    // further levels: record index from the cell (w4..w7 = even-depth nodes at depth 2, 4, 6, 8; longer codes: overflow chunk)
    const uint32_t inl_b = pairs_b > CELL_INLINE_PAIRS ? CELL_INLINE_PAIRS - 1u : pairs_b;
    const uint32_t inl_a = pairs_a > CELL_INLINE_PAIRS ? CELL_INLINE_PAIRS - 1u : pairs_a;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t k = COUNT_SPEC_ROOT ? 1u : 0u; k < CELL_INLINE_PAIRS; ++k) {
        const bool go_b = k < inl_b, go_a = k < inl_a;
        if (go_a | go_b) {
            const Rec32* qb = ix.sectors + (cell_b.w[3 + k] + B.r / SECTOR_BITS);
            const Rec32* qa = ix.sectors + (cell_a.w[3 + k] + A.r / SECTOR_BITS);
            const bool l_a = go_a && !(go_b && qa == qb);
            Rec32 yb FMGPU_UNSET, ya FMGPU_UNSET;
            if (go_b) yb = FMGPU_LD256(qb);
            if (l_a) ya = FMGPU_LD256(qa);
            if (STATS) cnt.loads += (go_b ? 1u : 0u) + (l_a ? 1u : 0u);
            if (go_b) track_levels(B, yb);
            if (go_a) track_levels(A, l_a ? ya : yb);
        }
    }
    // codes longer than 10 bits (large alphabets): the rest of the path comes from the overflow chunks
    if (pairs_a > CELL_INLINE_PAIRS || pairs_b > CELL_INLINE_PAIRS) {
        const uint32_t* more_b = reinterpret_cast<const uint32_t*>(ix.ovf + cell_b.w[7]);
        const uint32_t* more_a = reinterpret_cast<const uint32_t*>(ix.ovf + cell_a.w[7]);
        const uint32_t top = pairs_a > pairs_b ? pairs_a : pairs_b;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (uint32_t k = CELL_INLINE_PAIRS - 1u; k < top; ++k) {
            const bool go_b = pairs_b > CELL_INLINE_PAIRS && k < pairs_b;
            const bool go_a = pairs_a > CELL_INLINE_PAIRS && k < pairs_a;
            if (go_b) {
                const Rec32 y = FMGPU_LD256(ix.sectors + (FMGPU_LDG32(more_b + (k - (CELL_INLINE_PAIRS - 1u))) + B.r / SECTOR_BITS));
                track_levels(B, y);
            }
            if (go_a) {
                const Rec32 y = FMGPU_LD256(ix.sectors + (FMGPU_LDG32(more_a + (k - (CELL_INLINE_PAIRS - 1u))) + A.r / SECTOR_BITS));
                track_levels(A, y);
            }
            if (STATS) cnt.loads += (go_b ? 1u : 0u) + (go_a ? 1u : 0u);
        }
    }
    // a rank never exceeds the number of positions; the clamp only matters for a corrupt (but loadable) index, whose
    // boundary ranks could otherwise send the next step outside the directories
    const uint32_t va = A.base + A.r, vb = B.base + B.r;
    *sp = on_a ? (va < ix.length ? va : ix.length) : 0u;
    *ep = vb < ix.length ? vb : ix.length;
    return err;
}

}  // namespace fmgpu
