// Backward search, lane-per-pattern with refill ("flat" kernel, k_count v6): per-lane code.
//
// FmIndex.count (fm/FmIndex.java:455-474) = per pattern char one step of two rank queries, WaveletFixedBlockBoosting.rank
// (wavelet/WaveletFixedBlockBoosting.java:1010-1285), i.e. per step one (block, symbol) cell fetch followed by 0..n level-record
// fetches (two tree levels per record, layout.h).  The lockstep kernel (v3-v5, count_lane.h) runs the 32 patterns of a warp step
// by step and a step lasts as long as the DEEPEST of its 64 tracks: 2.9 record trips per step where a lane needs 1.15
// (DESIGN.md section 4.1).  Here the lanes are decoupled: every trip of the warp loop a lane issues the ONE dependent fetch its own
// pattern needs next — its pattern descriptor (a lane that has finished a pattern takes the next one from the work queue), the
// cell(s) of its step, or the level record(s) of its two tracks — and all 32 lanes issue together, so a warp trip costs one
// memory round trip whatever state its lanes are in.  The three fetch kinds are all 32-byte records, which keeps the load
// section of the trip uniform; only the (short) register arithmetic after it diverges.
//
// Host/device code: k_count_flat (kernels.cuh) runs it per lane, tests/support/flatcheck.cpp replays it against the CPU oracle.
#pragma once
#include <cstdint>

#include "count_lane.h"  // pattern_start, start_table_lookup, CountCounters, track arithmetic (dlevel_rank)
#include "lane_logic.h"

namespace fmgpu {

// --- pattern descriptor of the flat kernel (one 32-byte record per pattern, written by k_prep_flat): the state of
// FmIndex.count when it enters its loop (:456-463), or — with the q-gram start table — after the q - 1 steps the table covers.
//   w0, w1 : offset of the pattern's first char in the char array (u64)
//   w2     : i     — index (inside the pattern) of the char consumed last; the loop runs while start < end && i >= 1
//   w3     : flags — FD_ERR: the reference throws (empty pattern: pattern[-1], :456-457); FD_DEAD: the last char is unknown (:458), result 0
//   w4, w5 : start, end
//   w6     : alphabet code of pattern[i - 1] (the char the first loop iteration consumes), w7: raw char pattern[i - 2]
constexpr uint32_t FD_ERR = 1u, FD_DEAD = 2u;

FMGPU_HD Rec32 flat_make_desc(const DevIndex& ix, const uint32_t* C, const uint16_t* chars, uint64_t a, uint64_t b, const uint16_t* char2code,
                              uint32_t kmer_q) {
    Rec32 d;
    for (int k = 0; k < 8; ++k) d.w[k] = 0;
    d.w[0] = (uint32_t)a;
    d.w[1] = (uint32_t)(a >> 32);
    const uint32_t len = b > a ? (uint32_t)(b - a) : 0u;
    if (len == 0u) {  // pattern[-1]: ArrayIndexOutOfBounds (:456-457)
        d.w[3] = FD_ERR;
        return d;
    }
    const uint16_t* pch = chars + a;
    uint32_t c = pattern_start(chars, b, len, char2code, kmer_q, ix.kmer_stride, ix.sigma);
    int32_t i = (int32_t)len - 1;
    uint32_t sp = 0, ep = 0;
    bool from_table = false;
    if (c & PAT_KMER) {
        if (start_table_lookup(ix, c, &sp, &ep)) {
            i -= (int32_t)ix.kmer_q - 1;
            from_table = true;
        } else {
            c = (uint32_t)FMGPU_LDG16(char2code + pch[i]);  // that q-gram throws on its way: step by step
        }
    }
    if (!from_table) {
        if (c == 0u) {  // :458
            d.w[3] = FD_DEAD;
            return d;
        }
        sp = C[c];
        ep = C[c + 1];
    }
    d.w[2] = (uint32_t)i;
    d.w[4] = sp;
    d.w[5] = ep;
    d.w[6] = i >= 1 ? (uint32_t)FMGPU_LDG16(char2code + pch[i - 1]) : 0u;
    d.w[7] = i >= 2 ? (uint32_t)pch[i - 2] : 0u;
    return d;
}

enum FlatState : uint32_t { FS_IDLE = 0, FS_DESC = 1, FS_CELL = 2, FS_LEVEL = 3 };

struct FlatTrack {
    uint32_t base, code, r, rec, next, cell;
    uint32_t len, k;  // tree levels left; index of the record pair the next record belongs to (0 = root)
};

struct FlatLane {
    uint32_t state, pat, err;
    const uint16_t* pch;
    int32_t i;
    uint32_t c, cnext, raw2, sp, ep;
    FlatTrack A, B;  // A: start (off when start == 0: rank(0, .) == 0, :1012), B: end
};

struct FlatOut {
    int32_t* counts;
    int32_t* status;   // may be null
    uint32_t* ranges;  // may be null: SA range per pattern (locate)
};

FMGPU_HD void flat_finish_pattern(FlatLane& L, const FlatOut& O) {
    const int32_t result = (!L.err && L.ep > L.sp) ? (int32_t)(L.ep - L.sp) : 0;  // :473
    O.counts[L.pat] = result;
    if (O.status) O.status[L.pat] = L.err ? 9 : 0;
    if (O.ranges) {
        O.ranges[2 * (uint64_t)L.pat] = L.sp;
        O.ranges[2 * (uint64_t)L.pat + 1] = result > 0 ? L.ep : L.sp;
    }
    L.state = FS_IDLE;
}

// loop head of FmIndex.count (:464-468): either the pattern is finished, or the lane is set up for the cell fetch of its next step
FMGPU_HD void flat_next_step(const DevIndex& ix, const SmemTables& T, FlatLane& L, const FlatOut& O, const uint16_t* char2code) {
    if (!(L.sp < L.ep && L.i >= 1)) return flat_finish_pattern(L, O);
    --L.i;
    L.c = L.cnext;
    if (L.c == 0u || L.c >= ix.sigma) {  // :466-468 unknown symbol => 0 ; rank of a symbol >= sigma is 0 => empty range
        L.sp = L.ep = 0;
        return flat_finish_pattern(L, O);
    }
    if (ix.q4 && L.ep >= ix.length) {  // rank(size, .) on a superblock boundary throws (:1022-1026)
        L.err = 1;
        return flat_finish_pattern(L, O);
    }
    // chars are mapped to codes on the fly, fetched two steps (raw char) and one step (its code) ahead
    if (L.i >= 1) L.cnext = (uint32_t)FMGPU_LDG16(char2code + L.raw2);
    if (L.i >= 2) L.raw2 = (uint32_t)FMGPU_LDG16(L.pch + (L.i - 2));
    const uint32_t s = L.sp, e = L.ep;
    const SbDesc db = T.sb[e >> SB_LOG];
    const SbDesc da = T.sb[s >> SB_LOG];
    const uint32_t blk_b = db.first_block + ((e & SB_MASK) >> db.block_log);
    const uint32_t blk_a = da.first_block + ((s & SB_MASK) >> da.block_log);
    L.B.r = e & ((1u << db.block_log) - 1u);
    L.A.r = s & ((1u << da.block_log) - 1u);
    L.B.cell = blk_b * ix.sigma + L.c;  // < 2^32: the loader rejects larger cell tables (flatten.hpp)
    L.A.cell = blk_a * ix.sigma + L.c;
    L.state = FS_CELL;
}

// a fetched cell opens a track (count_lane.h: track_open); returns the first level record it needs
FMGPU_HD void flat_open(const DevIndex& ix, FlatTrack& t, const Rec32& cell, bool on, uint32_t* err, CountCounters& cnt, bool stats) {
    const uint32_t kind = (cell.w[2] >> 8) & 0xffu;
    const uint32_t Lc = cell.w[2] & 0xffu;
    t.base = cell.w[0];
    t.len = (on && kind == CELL_NORMAL) ? Lc : 0u;
    t.code = t.len ? cell.w[1] << (32u - t.len) : 0u;
    if (kind == CELL_CONST || kind == CELL_THROW) t.r = 0u;  // CONST: the cell is the answer; RUN: boundary rank + position in block
    if (!on) {
        t.r = 0u;
        t.base = 0u;
    }
    if (on && kind == CELL_THROW) *err = 1u;
    t.k = 0u;
    if (stats) {
        cnt.levels += t.len;
        cnt.recs += (t.len + 1u) >> 1;
    }
    if (t.len > 2u * CELL_INLINE_PAIRS) {
        // codes longer than 10 bits (large alphabets): walk this track to its end right here — rare, and the path beyond the
        // inline pointers lives in the overflow chunks (layout.h)
        const uint32_t pairs = (t.len + 1u) >> 1;
        const uint32_t inl = CELL_INLINE_PAIRS - 1u;
        const uint32_t* more = reinterpret_cast<const uint32_t*>(ix.ovf + cell.w[7]);
        for (uint32_t k = 0; k < pairs; ++k) {
            const uint32_t node = k < inl ? rec_word(cell, 3u + k) : FMGPU_LDG32(more + (k - inl));
            const Rec32 x = FMGPU_LD256(ix.sectors + (node + t.r / SECTOR_BITS));
            t.r = dlevel_rank(x, t.r % SECTOR_BITS, t.code >> 31, (t.code >> 30) & 1u);
            t.code <<= 2;
            if (stats) ++cnt.loads;
        }
        t.len = 0u;
        return;
    }
    t.rec = cell.w[3] + t.r / SECTOR_BITS;
    t.next = cell.w[4];
}

// One trip of a lane that holds a pattern: fetch + process, then the step bookkeeping.  `desc_base`: the descriptors.
template <bool STATS>
FMGPU_HD void flat_trip(const DevIndex& ix, const SmemTables& T, FlatLane& L, const FlatOut& O, const Rec32* descs, const uint16_t* chars,
                        const uint16_t* char2code, CountCounters& cnt) {
    const uint32_t state = L.state;
    // ---- addresses (every kind of fetch is one 32-byte record)
    const Rec32* pb;
    const Rec32* pa;
    bool ldb, lda;
    if (state == FS_LEVEL) {
        pb = ix.sectors + L.B.rec;
        pa = ix.sectors + L.A.rec;
        ldb = L.B.len != 0u;
        lda = L.A.len != 0u && !(ldb && L.A.rec == L.B.rec);
    } else if (state == FS_CELL) {
        pb = ix.cells + L.B.cell;
        pa = ix.cells + L.A.cell;
        ldb = true;
        lda = L.sp != 0u && L.A.cell != L.B.cell;
    } else {
        pb = descs + L.pat;
        pa = pb;
        ldb = true;
        lda = false;
    }
    // pointer of the record pair after the next one (codes of 5..10 bits): a 4-byte load from the cell, issued with the records
    uint32_t nn_b = 0, nn_a = 0;
    const bool deep_b = state == FS_LEVEL && L.B.len > 4u, deep_a = state == FS_LEVEL && L.A.len > 4u;
    Rec32 xb FMGPU_UNSET, xa FMGPU_UNSET;
    if (ldb) xb = FMGPU_LD256(pb);
    if (lda) xa = FMGPU_LD256(pa);
    if (deep_b) nn_b = FMGPU_LDG32(reinterpret_cast<const uint32_t*>(ix.cells + L.B.cell) + (5u + L.B.k));
    if (deep_a) nn_a = FMGPU_LDG32(reinterpret_cast<const uint32_t*>(ix.cells + L.A.cell) + (5u + L.A.k));
    if (STATS && state != FS_DESC) cnt.loads += (ldb ? 1u : 0u) + (lda ? 1u : 0u);

    // ---- process
    if (state == FS_LEVEL) {
        if (L.A.len) {
            const Rec32& x = lda ? xa : xb;
            L.A.r = dlevel_rank(x, L.A.r % SECTOR_BITS, L.A.code >> 31, (L.A.code >> 30) & 1u);
            L.A.code <<= 2;
            L.A.len = L.A.len >= 2u ? L.A.len - 2u : 0u;
            L.A.rec = L.A.next + L.A.r / SECTOR_BITS;
            L.A.next = nn_a;
            ++L.A.k;
        }
        if (L.B.len) {
            L.B.r = dlevel_rank(xb, L.B.r % SECTOR_BITS, L.B.code >> 31, (L.B.code >> 30) & 1u);
            L.B.code <<= 2;
            L.B.len = L.B.len >= 2u ? L.B.len - 2u : 0u;
            L.B.rec = L.B.next + L.B.r / SECTOR_BITS;
            L.B.next = nn_b;
            ++L.B.k;
        }
    } else if (state == FS_CELL) {
        const bool on_a = L.sp != 0u;
        if (STATS) cnt.ranks += on_a ? 2u : 1u;
        flat_open(ix, L.B, xb, true, &L.err, cnt, STATS);
        flat_open(ix, L.A, lda ? xa : xb, on_a, &L.err, cnt, STATS);
        if (L.err) return flat_finish_pattern(L, O);
        L.state = FS_LEVEL;
    } else {  // descriptor: the pattern enters the loop of FmIndex.count
        L.pch = chars + (((uint64_t)xb.w[1] << 32) | xb.w[0]);
        L.i = (int32_t)xb.w[2];
        L.sp = xb.w[4];
        L.ep = xb.w[5];
        L.cnext = xb.w[6];
        L.raw2 = xb.w[7];
        L.err = xb.w[3] & FD_ERR;
        if (xb.w[3]) {  // the reference throws / returns 0 before its loop
            L.sp = L.ep = 0;
            return flat_finish_pattern(L, O);
        }
        return flat_next_step(ix, T, L, O, char2code);
    }
    // ---- both tracks at their leaves: close the step (:469-470) and set the next one up
    if ((L.A.len | L.B.len) == 0u) {
        const uint32_t va = L.A.base + L.A.r, vb = L.B.base + L.B.r;
        const uint32_t base = T.C[L.c];
        // a rank never exceeds the number of positions; the clamps only matter for a corrupt (but loadable) index
        L.sp = L.sp != 0u ? base + (va < ix.length ? va : ix.length) : base;
        const uint32_t e = base + (vb < ix.length ? vb : ix.length);
        L.ep = e < ix.length ? e : ix.length;
        flat_next_step(ix, T, L, O, char2code);
    }
}

}  // namespace fmgpu
