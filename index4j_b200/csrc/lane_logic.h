// Per-lane logic of the query path: what one lane does with each 32-byte record it fetched.
// Pure register arithmetic (no memory access besides the small C / superblock tables), so the same
// functions run inside the sm_100a kernels and, compiled by g++, inside the host-side layout test
// (tests/support/flatcheck.cpp) that replays them against the CPU oracle.
//
// Reference semantics restated here (paths under indices/src/main/java/com/dynatrace/):
//   dlevel_rank / rank_single   wavelet/WaveletFixedBlockBoosting.java:1010-1285 (rank)
//   dlevel_descend              wavelet/WaveletFixedBlockBoosting.java:1305-1537 (inverseSelect)
//   rrr_*                       bitsequence/RrrVector.java:111-129 + tables :8692-16899
//   eub_*    fm/FmIndex.java:692-758, 772-831, 844-922 (control flow of extractUntilBoundary*)
#pragma once
#include <cstdint>

#include "layout.h"
#include "ldrec.h"

#if defined(__CUDACC__)
#define FMGPU_HD __host__ __device__ __forceinline__
#else
#define FMGPU_HD inline
#endif

namespace fmgpu {

FMGPU_HD uint32_t popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(v);
#else
    return (uint32_t)__builtin_popcount(v);
#endif
}

// word k (0..7) of a record, k not known at compile time (select chain keeps the record in registers)
FMGPU_HD uint32_t rec_word(const Rec32& s, uint32_t k) {
    uint32_t v = s.w[0];
#pragma unroll
    for (int i = 1; i < 8; ++i)
        if (k == (uint32_t)i) v = s.w[i];
    return v;
}

// mask of the low min(max(width, 0), 32) bits
FMGPU_HD uint32_t low_mask_clamped(int width) {
#if defined(__CUDA_ARCH__)
    uint32_t m;
    const uint32_t w = (uint32_t)(width > 0 ? width : 0);
    asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(m) : "r"(0u), "r"(w));
    return m;
#else
    return width >= 32 ? 0xffffffffu : (width <= 0 ? 0u : ((1u << width) - 1u));
#endif
}

struct SmemTables {  // C array and superblock descriptors (shared memory when they fit)
    const uint32_t* C;
    const SbDesc* sb;
};

// ------------------------------------------------------------------------------------------
// occurrence structures of the (block, symbol) cells (layout.h): rank inside the block = occurrences among its first r positions
// ------------------------------------------------------------------------------------------
// how many of the n_words * 2 ascending u16 values packed in w[first ..] are < r (padding 0xffff never is: r <= 65535)
FMGPU_HD uint32_t count_u16_below(const Rec32& x, int first, int n_words, uint32_t r) {
    uint32_t n = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k >= first && k < first + n_words) n += ((x.w[k] & 0xffffu) < r ? 1u : 0u) + ((x.w[k] >> 16) < r ? 1u : 0u);
    return n;
}
FMGPU_HD uint32_t cell_kind(const Cell8& c) { return c.info >> CELL_KIND_SHIFT; }
// the ONE record an OCC_* cell needs for position r of its block
FMGPU_HD const Rec32* occ_record(const DevIndex& ix, const Cell8& cell, uint32_t kind, uint32_t r) {
    const uint32_t q = kind == CELL_OCC_BITS ? r / OCC_BITS_PER_REC : r >> ((uint32_t)(OCC_RANGE_SHIFTS >> (8u * kind)) & 31u);
    return ix.occ + ((cell.info & CELL_PTR_MASK) + q);
}

// ------------------------------------------------------------------------------------------
// level records (layout.h): two tree levels from one 64-byte record
// ------------------------------------------------------------------------------------------
// elements among the first b positions of a level record whose bits at the record's two levels are (t, u): one pass over
// (plane0 ^ ~T) & (plane1 ^ ~U) & mask.  A code that ends at the record's first level (child t is a leaf) is the case u = 0:
// plane 1 holds 0 for every element that goes to a leaf child, so the same expression counts "bit t at this level".
FMGPU_HD uint32_t dlevel_count(const Rec32& x, uint32_t b, uint32_t t, uint32_t u) {
    const uint32_t tm = t ? 0u : 0xffffffffu;
    const uint32_t um = u ? 0u : 0xffffffffu;
    uint32_t n = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) n += popc32((x.w[2 + k] ^ tm) & (x.w[5 + k] ^ um) & low_mask_clamped((int)b - 32 * k));
    return n;
}
// elements before the record with bits (t, u): w0 = c00 | c01 << 16, w1 = c10 | c11 << 16 (node-relative, < 65536)
FMGPU_HD uint32_t dlevel_base(const Rec32& x, uint32_t t, uint32_t u) { return ((t ? x.w[1] : x.w[0]) >> (u << 4)) & 0xffffu; }
// bit b (0..95) of plane 0 / plane 1
FMGPU_HD uint32_t plane_bit(const Rec32& x, uint32_t plane, uint32_t b) {
    const uint32_t wi = b >> 5;
    const uint32_t w = plane ? (wi == 0u ? x.w[5] : (wi == 1u ? x.w[6] : x.w[7])) : (wi == 0u ? x.w[2] : (wi == 1u ? x.w[3] : x.w[4]));
    return (w >> (b & 31u)) & 1u;
}

// The rank walk of WaveletFixedBlockBoosting.rank (:1185-1279) for the two levels a record holds: position b inside the
// record (b = r mod 96, r = position in the even-depth node), code bits t (this level) and u (next level; 0 when the code
// ends at this level).  Returns the position in grandchild (t, u) — in child t when the code ends here.
FMGPU_HD uint32_t dlevel_rank(const Rec32& x, uint32_t b, uint32_t t, uint32_t u) { return dlevel_base(x, t, u) + dlevel_count(x, b, t, u); }

// occurrences of the symbol among the first r positions of the block, from the record occ_record pointed at (y);
// *len (work counters) = the symbol's code length in the block's tree
FMGPU_HD uint32_t occ_in_record(const Rec32& y, uint32_t kind, uint32_t r, uint32_t* len) {
    *len = y.w[0] >> 24;
    uint32_t n = y.w[0] & 0xffffffu;
    if (kind != CELL_OCC_BITS) {  // position list: 14 u16 slots in w1..w7 (padding 0xffff is never below r <= 65535)
#pragma unroll
        for (int k = 1; k < 8; ++k) n += ((y.w[k] & 0xffffu) < r ? 1u : 0u) + ((y.w[k] >> 16) < r ? 1u : 0u);
        return n;
    }
    const uint32_t b = r % OCC_BITS_PER_REC;
#pragma unroll
    for (int k = 0; k < 7; ++k) n += popc32(y.w[1 + k] & low_mask_clamped((int)b - 32 * k));
    return n;
}

// The descent of WaveletFixedBlockBoosting.inverseSelect (:1386-1505) through the two levels of a record with the node
// record N of the even-depth node.  Returns true at a leaf (*sym, *rank = rank(pos, sym) incl. the boundary rank);
// false: continue at node record *nrec / level record *sec with position *r.
FMGPU_HD bool dlevel_descend(const Rec32& x, const Rec32& N, uint32_t b, uint32_t* r, uint32_t* nrec, uint32_t* sec, uint32_t* sym,
                             uint32_t* rank, uint32_t* levels) {
    const uint32_t t = plane_bit(x, 0, b);
    const uint32_t u = plane_bit(x, 1, b);
    const uint32_t e0 = t ? N.w[4] : N.w[0];
    const uint32_t a0 = t ? N.w[5] : N.w[1];
    const bool leaf1 = (e0 & LEAF1_FLAG) != 0u;
    const uint32_t r2 = dlevel_rank(x, b, t, u);  // leaf child: u == 0 and the count is that of bit t alone
    if (leaf1) {
        ++*levels;
        *sym = e0 & 0xffffu;
        *rank = a0 + r2;
        return true;
    }
    *levels += 2;
    const uint32_t e1 = t ? N.w[6] : N.w[2];
    const uint32_t a1 = t ? N.w[7] : N.w[3];
    const uint32_t e = u ? e1 : e0;
    const uint32_t a = u ? a1 : a0;
    if (e & LEAF_FLAG) {
        *sym = e & 0xffffu;
        *rank = a + r2;
        return true;
    }
    *nrec = e;
    *sec = a;
    *r = r2;
    return false;
}

// ------------------------------------------------------------------------------------------
// rank(pos, sym), one query (wavelet/WaveletFixedBlockBoosting.java:1010-1285).  Returns 0, or 9 where the reference
// throws (position == size on a superblock boundary, :1022-1026).  *n_rank / *n_level / *n_rec count cells fetched / tree levels walked / level records fetched.
// ------------------------------------------------------------------------------------------
FMGPU_HD uint32_t rank_single(const DevIndex& ix, const SmemTables& T, uint32_t pos, uint32_t sym, uint32_t* out, uint32_t* n_rank,
                              uint32_t* n_level, uint32_t* n_rec) {
    *out = 0;
    if (pos == 0) return 0u;            // :1012
    if (pos > ix.length) pos = ix.length;  // :1015
    if (sym >= ix.sigma) return 0u;     // :1018
    if (ix.q4 && pos == ix.length) return 9u;
    const SbDesc sd = T.sb[pos >> SB_LOG];
    const uint32_t blk = sd.first_block + ((pos & SB_MASK) >> sd.block_log);
    const uint32_t r = pos & ((1u << sd.block_log) - 1u);
    const Cell8 cell = FMGPU_LDCELL(ix.cells + ((uint64_t)blk * ix.sigma + sym));
    ++*n_rank;
    const uint32_t kind = cell_kind(cell);
    if (kind == CELL_CONST) {
        *out = cell.value;
        return 0u;
    }
    if (kind == CELL_RUN) {  // :1141-1146
        *out = cell.value + r;
        return 0u;
    }
    if (kind < CELL_OCC_FIRST) return 9u;  // THROW
    const Rec32 y = FMGPU_LD256(occ_record(ix, cell, kind, r));
    ++*n_rec;
    uint32_t len = 0;
    *out = cell.value + occ_in_record(y, kind, r, &len);
    *n_level += len;
    return 0u;
}

// BITS_NEEDED_BINOMIAL_COEFFICIENTS (RrrVector.java:111-129), one nibble per class
// classes 0..15 need 1,4,7,9,11,12,13,13,13,13,12,11,9,7,4,1 bits (nibble c of the constant = class c)
constexpr unsigned long long RRR_BITS_NEEDED = 0x1479BCDDDDCB9741ULL;
FMGPU_HD uint32_t rrr_bits(uint32_t cls) { return (uint32_t)(RRR_BITS_NEEDED >> (4u * cls)) & 15u; }

FMGPU_HD const Rec32* sg_addr(const DevIndex& ix, uint32_t pos) { return ix.sgroups + ((pos / RRR_BLOCK) >> 5); }

// binom[b*16 + k] = C(b, k) for b, k in 0..14 (0 when k > b)
FMGPU_HD void fill_binom(uint16_t* binom) {
    for (int b = 0; b < 15; ++b)
        for (int k = 0; k < 16; ++k) {
            uint32_t v = 0;
            if (k == 0) v = 1;
            else if (b > 0 && k <= b) v = (uint32_t)binom[(b - 1) * 16 + (k - 1)] + (uint32_t)binom[(b - 1) * 16 + k];
            binom[b * 16 + k] = (uint16_t)v;
        }
}

// (class, offset) -> 15-bit block by combinatorial unranking with a 15x15 binomial table
// (binom[b*16 + k] = C(b, k)); only the bits below `upto` (<= 15) are produced.
// Order of the reference tables: inside a class, descending value of the block read LSB-first.
FMGPU_HD uint32_t rrr_unrank(const uint16_t* binom, uint32_t cls, uint32_t off, uint32_t upto) {
    uint32_t v = 0, k = cls;
    for (uint32_t j = 0; j < upto && k; ++j) {
        const uint32_t c = binom[(14u - j) * 16u + (k - 1u)];
        if (off < c) {
            v |= 1u << j;
            --k;
        } else {
            off -= c;
        }
    }
    return v;
}

// ------------------------------------------------------------------------------------------
// extractUntilBoundary* control flow, given what the walks found.
//   down_len : chars collected by the left walk (fm/FmIndex.java:664-686)
//   rel      : distance from `from` to the first boundary at or right of it, or -1 when the text
//              ends first (positions from .. length-2 hold no boundary)
// Both replay the reference's 4-char chunk loop arithmetically (:692-758 and :860-921).
// ------------------------------------------------------------------------------------------
struct EubOut {
    int32_t status, value;
};

// `offset` = the reference's offset argument: it shifts where the chars land and enters the "does not fit" test and its N
// (:732-737, :894-898), not the returned length.
FMGPU_HD EubOut eub_right_chunks(int32_t from, int32_t down_len, int32_t rel, int32_t length, int32_t dst_len, bool right_only,
                                 int32_t offset = 0) {
    EubOut o;
    o.status = 0;
    o.value = 0;
    int32_t final_pos = -1, times = 1;
    {
        // Skip the chunks in which provably nothing happens, so the loop below only runs the chunk that ends the call (and at most
        // one more): chunk t is uneventful while it is a full 4-char chunk before the end of the text (t < te), lies before the
        // boundary's chunk (t < th) and — full chunks only get closer to the destination's end — does not overflow (t < to).
        // The loop body itself stays the reference's arithmetic.  (A lane used to run this loop alone, ~30 trips per record,
        // while the other 31 lanes of its warp waited: 38 % of k_extract's issued instructions.)
        const int64_t l1 = (int64_t)length - 1;
        const int64_t te = (int64_t)from >= l1 ? 1 : (l1 - from + 3) / 4;
        const int64_t th = rel >= 0 ? (int64_t)rel / 4 + 1 : te;
        const int64_t x = (int64_t)dst_len - offset - (right_only ? 0 : down_len);  // full chunk t overflows iff 4t - 1 >= x
        const int64_t to = x <= 3 ? 1 : (x + 4) / 4;
        int64_t t0 = te < th ? te : th;
        if (to < t0) t0 = to;
        if (t0 > 1) {
            times = (int32_t)t0;
            from += (int32_t)(4 * (t0 - 1));
        }
    }
    for (;;) {
        const int32_t prev = from;
        from += 4;
        if (from > length - 1) from = length - 1;
        const int32_t rem = from - prev;
        const int32_t top = (times - 1) * 4 + rem - 1;  // upStreamPos of the first char processed
        const int32_t lo = top - rem + 1;               // ... and of the last one
        // chars are visited from `top` down to `lo`; the only boundary that matters is the one at rel
        if (rem > 0) {
            // visiting order: a throw at `top` precedes everything except a boundary AT top
            const bool hit = rel >= lo && rel <= top;
            if (hit && rel == top && top == 0) {
                o.value = 0;
                return o;  // "return 0" (:719 / :877)
            }
            const int32_t limit = offset + (right_only ? top : down_len + top);
            if (limit >= dst_len) {
                o.status = 8;
                o.value = limit;
                return o;
            }
            if (hit) {
                if (rel == 0) {
                    o.value = 0;
                    return o;
                }
                final_pos = rel;
            }
        }
        if (from == length - 1) {  // :745-752 / :908-915 (quirk Q5)
            // upStreamPos after the chunk; the right-only variant does not decrement at 0 (:899-902)
            int32_t up = lo - 1;
            if (right_only && rem > 0 && lo == 0) up = 0;
            if (right_only) final_pos = up + rem;
            else final_pos = up < 0 ? 1 : up + rem;
            break;
        }
        if (final_pos != -1) break;
        ++times;
    }
    o.value = right_only ? final_pos - 1 : down_len + final_pos;
    return o;
}

}  // namespace fmgpu
