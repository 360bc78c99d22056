// Per-lane code of one backward-search step (FmIndex.count, fm/FmIndex.java:464-471): the two rank queries
// rank(start, c) and rank(end, c) of WaveletFixedBlockBoosting.rank (wavelet/WaveletFixedBlockBoosting.java:1010-1285)
// run as TWO TRACKS of one fused step.  Host/device code: k_count (kernels.cuh) runs it per lane, the host layout test
// (tests/support/flatcheck.cpp) replays it against the CPU oracle.
//
// Track B carries `end`, track A carries `start` (off when start == 0, rank(0, .) == 0, :1012).  Per step a lane issues
//   1. the (block, symbol) cell of each track (one load when both positions lie in the same block), then
//   2. AT MOST ONE occurrence record per track (layout.h: a short sorted position list or a bit vector of the symbol's
//      occurrences in the block; one load when both tracks need the same record; none for CONST / RUN cells).
// All loads of a stage are issued before any is used.  A step is therefore two dependent memory round trips whatever the
// code lengths of the symbols are — rounds 1-2 walked the wavelet levels here (one record per two tree levels), and a warp
// ran to the deepest of its 64 tracks: 2.9 record trips per step for 1.15 needed per track.
#pragma once
#include <cstdint>

#include "lane_logic.h"

// a record variable that is only read after a (conditional) load: left uninitialised on the device, zeroed on the host
#if defined(__CUDA_ARCH__)
#define FMGPU_UNSET
#else
#define FMGPU_UNSET {}
#endif

namespace fmgpu {

using CountTables = SmemTables;  // C, superblock descriptors

struct CountCounters {
    uint32_t ranks, levels, loads, recs;
    uint32_t kinds[8];  // rank tracks by cell kind (CellKind)
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

// One backward-search step for the lane: on return *sp / *ep hold rank(start, c) / rank(end, c) (NOT yet offset by C[c]).
// Returns 1 where the reference throws (THROW cells).  `on` = the lane takes part in this step.
template <bool STATS>
FMGPU_HD uint32_t count_step(const DevIndex& ix, const CountTables& T, uint32_t c, uint32_t* sp, uint32_t* ep, bool on, CountCounters& cnt) {
    if (!on) return 0u;
    const uint32_t s = *sp, e = *ep;
    const bool on_a = s != 0u;
    const SbDesc db = T.sb[e >> SB_LOG];
    const SbDesc da = T.sb[s >> SB_LOG];
    const uint32_t blk_b = db.first_block + ((e & SB_MASK) >> db.block_log);
    const uint32_t blk_a = da.first_block + ((s & SB_MASK) >> da.block_log);
    const uint32_t rb = e & ((1u << db.block_log) - 1u);
    const uint32_t ra = s & ((1u << da.block_log) - 1u);
    const bool split = on_a && blk_a != blk_b;
    if (STATS) cnt.ranks += on_a ? 2u : 1u;

    // stage 0: the cells
    const Cell8 cell_b = FMGPU_LDCELL(ix.cells + ((uint64_t)blk_b * ix.sigma + c));
    Cell8 cell_a = cell_b;
    if (split) cell_a = FMGPU_LDCELL(ix.cells + ((uint64_t)blk_a * ix.sigma + c));
    if (STATS) cnt.loads += split ? 2u : 1u;
    const uint32_t kind_b = cell_kind(cell_b);
    const uint32_t kind_a = on_a ? cell_kind(cell_a) : (uint32_t)CELL_CONST;
    const bool need_b = kind_b >= CELL_OCC_FIRST;
    const bool need_a = kind_a >= CELL_OCC_FIRST;
    const uint32_t err = ((!need_b && kind_b != CELL_CONST && kind_b != CELL_RUN) || (!need_a && kind_a != CELL_CONST && kind_a != CELL_RUN)) ? 1u : 0u;

    // stage 1: one occurrence record per track that needs one
    const Rec32* rec_b = need_b ? occ_record(ix, cell_b, kind_b, rb) : nullptr;
    const Rec32* rec_a = need_a ? occ_record(ix, cell_a, kind_a, ra) : nullptr;
    uint32_t part_b = kind_b == CELL_RUN ? rb : 0u;  // boundary rank + position inside the single-symbol block (:1141-1146)
    uint32_t part_a = kind_a == CELL_RUN ? ra : 0u;
    const bool shared = need_a && need_b && rec_a == rec_b;
    Rec32 yb FMGPU_UNSET, ya FMGPU_UNSET;  // never interpreted unless loaded
    if (need_b) yb = FMGPU_LD256_OCC(rec_b);
    if (need_a && !shared) ya = FMGPU_LD256_OCC(rec_a);
    uint32_t len_b = 0, len_a = 0;
    if (need_b) part_b = occ_in_record(yb, kind_b, rb, &len_b);
    if (need_a) part_a = occ_in_record(shared ? yb : ya, kind_a, ra, &len_a);
    if (STATS) {
        cnt.loads += (need_b ? 1u : 0u) + (need_a && !shared ? 1u : 0u);
        cnt.recs += (need_b ? 1u : 0u) + (need_a ? 1u : 0u);
        cnt.levels += len_b + len_a;
        ++cnt.kinds[kind_b & 7u];
        if (on_a) ++cnt.kinds[kind_a & 7u];
    }

    // a rank never exceeds the number of positions; the clamp only matters for a corrupt (but loadable) index, whose
    // boundary ranks could otherwise send the next step outside the directories
    const uint32_t va = cell_a.value + part_a, vb = cell_b.value + part_b;
    *sp = on_a ? (va < ix.length ? va : ix.length) : 0u;
    *ep = vb < ix.length ? vb : ix.length;
    return err;
}

}  // namespace fmgpu
