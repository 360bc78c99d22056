// Per-lane logic of the query path: what one lane does with each 32-byte record it fetched.
// Pure register arithmetic (no memory access besides the small C / superblock tables), so the same
// functions run inside the sm_100a kernels and, compiled by g++, inside the host-side layout test
// (tests/support/flatcheck.cpp) that replays them against the CPU oracle.
//
// Reference semantics restated here (paths under indices/src/main/java/com/dynatrace/):
//   rank_*   wavelet/WaveletFixedBlockBoosting.java:1010-1285 (rank)
//   lf_*     wavelet/WaveletFixedBlockBoosting.java:1305-1537 (inverseSelect)
//   sg_*     bitsequence/RrrVector.java:314-396 (access, rankOnes) + tables :8692-16899
//   eub_*    fm/FmIndex.java:692-758, 772-831, 844-922 (control flow of extractUntilBoundary*)
#pragma once
#include <cstdint>

#include "layout.h"

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

// ones among the first `nbits` (< 224) payload bits of a level sector, plus its running count
FMGPU_HD uint32_t sector_rank(const Rec32& s, uint32_t nbits) {
    uint32_t ones = s.w[0];
#pragma unroll
    for (int k = 0; k < 7; ++k) ones += popc32(s.w[1 + k] & low_mask_clamped((int)nbits - 32 * k));
    return ones;
}
FMGPU_HD uint32_t sector_bit(const Rec32& s, uint32_t b) { return (rec_word(s, 1u + (b >> 5)) >> (b & 31u)) & 1u; }

struct SmemTables {  // C array and superblock descriptors (shared memory when they fit)
    const uint32_t* C;
    const SbDesc* sb;
};

// ------------------------------------------------------------------------------------------
// rank(pos, sym)
// ------------------------------------------------------------------------------------------
struct RankSt {
    uint32_t base, code, r, L, d, inl, ovf, bix;
    uint32_t p0, p1, p2, p3, p4, p5, p6, p7;  // sector of the node at depth d, d+1, ... (shift register)
};
enum RankOut : uint32_t { RK_MORE = 0, RK_DONE = 1, RK_THROW = 2 };

// Guards of rank (:1012-1020) and the address of the (block, symbol) cell.  RK_MORE: fetch *addr;
// RK_DONE: answer in *val; RK_THROW: the reference's ArrayIndexOutOfBounds for position == size
// on a superblock boundary (:1022-1026).
FMGPU_HD uint32_t rank_begin(const DevIndex& ix, const SmemTables& T, uint32_t pos, uint32_t sym, RankSt& s, const Rec32** addr,
                             uint32_t* val) {
    if (pos == 0) {
        *val = 0;
        return RK_DONE;
    }
    if (pos > ix.length) pos = ix.length;
    if (sym >= ix.sigma) {
        *val = 0;
        return RK_DONE;
    }
    if (ix.q4 && pos == ix.length) return RK_THROW;
    const SbDesc sd = T.sb[pos >> SB_LOG];
    const uint32_t blk = sd.first_block + ((pos & SB_MASK) >> sd.block_log);
    s.bix = pos & ((1u << sd.block_log) - 1u);
    *addr = ix.cells + ((uint64_t)blk * ix.sigma + sym);
    return RK_MORE;
}

FMGPU_HD uint32_t rank_on_cell(const DevIndex& ix, const Rec32& A, RankSt& s, const Rec32** addr, uint32_t* val) {
    const uint32_t kind = (A.w[2] >> 8) & 0xffu;
    if (kind == CELL_CONST) {
        *val = A.w[0];
        return RK_DONE;
    }
    if (kind == CELL_RUN) {  // :1141-1146
        *val = A.w[0] + s.bix;
        return RK_DONE;
    }
    if (kind == CELL_THROW) return RK_THROW;
    s.base = A.w[0];
    s.code = A.w[1];
    s.L = A.w[2] & 0xffu;
    s.r = s.bix;
    s.d = 0;
    s.p0 = A.w[3];
    s.p1 = A.w[4];
    s.p2 = A.w[5];
    s.p3 = A.w[6];
    s.p4 = A.w[7];
    s.ovf = A.w[7];
    s.inl = s.L > CELL_INLINE_LEVELS ? 4u : s.L;
    *addr = ix.sectors + (s.p0 + s.r / SECTOR_BITS);
    return RK_MORE;
}

// One level of the walk (:1185-1279).  RK_DONE with *val, or RK_MORE with *addr; *want_ovf says the
// next fetch is a path chunk rather than a level sector.
FMGPU_HD uint32_t rank_on_level(const DevIndex& ix, const Rec32& A, RankSt& s, const Rec32** addr, uint32_t* val, bool* want_ovf) {
    const uint32_t ones = sector_rank(A, s.r % SECTOR_BITS);
    const uint32_t bit = (s.code >> (s.L - 1u - s.d)) & 1u;
    s.r = bit ? ones : s.r - ones;
    ++s.d;
    *want_ovf = false;
    if (s.d == s.L) {
        *val = s.base + s.r;
        return RK_DONE;
    }
    s.p0 = s.p1;
    s.p1 = s.p2;
    s.p2 = s.p3;
    s.p3 = s.p4;
    s.p4 = s.p5;
    s.p5 = s.p6;
    s.p6 = s.p7;
    --s.inl;
    if (s.inl == 0) {
        *addr = ix.ovf + s.ovf;
        *want_ovf = true;
    } else {
        *addr = ix.sectors + (s.p0 + s.r / SECTOR_BITS);
    }
    return RK_MORE;
}

FMGPU_HD void rank_on_ovf(const DevIndex& ix, const Rec32& A, RankSt& s, const Rec32** addr) {
    s.p0 = A.w[0];
    s.p1 = A.w[1];
    s.p2 = A.w[2];
    s.p3 = A.w[3];
    s.p4 = A.w[4];
    s.p5 = A.w[5];
    s.p6 = A.w[6];
    s.p7 = A.w[7];
    s.inl = 8;
    ++s.ovf;
    *addr = ix.sectors + (s.p0 + s.r / SECTOR_BITS);
}

// ------------------------------------------------------------------------------------------
// inverseSelect(pos): symbol at pos and its rank, walking DOWN the block's tree.
// ------------------------------------------------------------------------------------------
struct LfSt {
    uint32_t r, c0, c1, a0, a1, nrec, bmask;
};
enum LfOut : uint32_t { LF_MORE = 0, LF_LEAF = 1, LF_RUN = 2 };

FMGPU_HD void lf_begin(const DevIndex& ix, const SmemTables& T, uint32_t pos, LfSt& s, const Rec32** addr) {
    const SbDesc sd = T.sb[pos >> SB_LOG];
    const uint32_t blk = sd.first_block + ((pos & SB_MASK) >> sd.block_log);
    s.bmask = (1u << sd.block_log) - 1u;
    s.r = pos & s.bmask;
    *addr = ix.blocks + blk;
}
// LF_RUN: single-symbol block, *sym = the symbol as inverseSelect decodes it (low byte only, :1329-1332);
// LF_MORE: fetch the root's level sector at *addr.
FMGPU_HD uint32_t lf_on_block(const DevIndex& ix, const Rec32& D, LfSt& s, const Rec32** addr, uint32_t* sym) {
    if (D.w[1] & 1u) {
        *sym = (D.w[1] >> 8) & 0xffffu;
        return LF_RUN;
    }
    s.c0 = D.w[4];
    s.c1 = D.w[5];
    s.a0 = D.w[6];
    s.a1 = D.w[7];
    *addr = ix.sectors + (D.w[0] + s.r / SECTOR_BITS);
    return LF_MORE;
}
FMGPU_HD void lf_take_record(const Rec32& B, LfSt& s) {  // node record fetched alongside the level sector
    const uint32_t h = (s.nrec & 1u) * 4u;
    s.c0 = h ? B.w[4] : B.w[0];
    s.c1 = h ? B.w[5] : B.w[1];
    s.a0 = h ? B.w[6] : B.w[2];
    s.a1 = h ? B.w[7] : B.w[3];
}
// One level (:1386-1505).  LF_LEAF: *sym and *rk = rank(pos, sym); LF_MORE: next level sector at
// *addr_a and the child's node record at *addr_b.
FMGPU_HD uint32_t lf_on_level(const DevIndex& ix, const Rec32& A, LfSt& s, const Rec32** addr_a, const Rec32** addr_b, uint32_t* sym,
                              uint32_t* rk) {
    const uint32_t b = s.r % SECTOR_BITS;
    const uint32_t ones = sector_rank(A, b);
    const uint32_t bit = sector_bit(A, b);
    const uint32_t cb = bit ? s.c1 : s.c0;
    const uint32_t ab = bit ? s.a1 : s.a0;
    s.r = bit ? ones : s.r - ones;
    if (cb & LEAF_FLAG) {
        *sym = cb & 0xffffu;
        *rk = ab + s.r;
        return LF_LEAF;
    }
    s.nrec = cb;
    *addr_a = ix.sectors + (ab + s.r / SECTOR_BITS);
    *addr_b = ix.nodes + (cb >> 1);
    return LF_MORE;
}

// ------------------------------------------------------------------------------------------
// sampled-row bitvector (RRR): access(pos) and rankOnes(pos) from one group record (+ offset bits)
// ------------------------------------------------------------------------------------------
// BITS_NEEDED_BINOMIAL_COEFFICIENTS (RrrVector.java:111-129), one nibble per class
// classes 0..15 need 1,4,7,9,11,12,13,13,13,13,12,11,9,7,4,1 bits (nibble c of the constant = class c)
constexpr unsigned long long RRR_BITS_NEEDED = 0x1479BCDDDDCB9741ULL;
FMGPU_HD uint32_t rrr_bits(uint32_t cls) { return (uint32_t)(RRR_BITS_NEEDED >> (4u * cls)) & 15u; }

struct SgSt {
    uint32_t ones, offb, cls, use;
};
enum SgOut : uint32_t { SG_DONE = 0, SG_OFFSET = 1 };

FMGPU_HD const Rec32* sg_addr(const DevIndex& ix, uint32_t pos) { return ix.sgroups + ((pos / RRR_BLOCK) >> 5); }

// SG_DONE: *bit = access(pos), *rank = rankOnes(pos).  SG_OFFSET: the block's offset field is
// needed; fetch offsets record *addr_a (and *addr_b when *straddle).
FMGPU_HD uint32_t sg_on_group(const DevIndex& ix, const Rec32& G, uint32_t pos, SgSt& s, uint32_t* bit, uint32_t* rank,
                              const Rec32** addr_a, const Rec32** addr_b, bool* straddle) {
    const uint32_t blk = pos / RRR_BLOCK;
    const uint32_t k = blk & 31u, sub = k >> 3, kk = k & 7u;
    s.use = pos - blk * RRR_BLOCK;
    uint32_t offb = G.w[1], ones = G.w[0];
    if (sub) {
        offb += (G.w[2] >> (10u * (sub - 1u))) & 1023u;
        ones += (G.w[3] >> (10u * (sub - 1u))) & 1023u;
    }
    const uint32_t word = rec_word(G, 4u + sub);
#pragma unroll
    for (uint32_t i = 0; i < 7; ++i) {
        if (i < kk) {
            const uint32_t c = (word >> (4u * i)) & 15u;
            ones += c;
            offb += rrr_bits(c);
        }
    }
    const uint32_t cls = (word >> (4u * kk)) & 15u;
    s.ones = ones;
    s.offb = offb;
    s.cls = cls;
    if (cls == 0u) {
        *bit = 0;
        *rank = ones;
        return SG_DONE;
    }
    if (cls == 15u) {
        *bit = 1;
        *rank = ones + s.use;
        return SG_DONE;
    }
    const Rec32* base = reinterpret_cast<const Rec32*>(ix.soffsets);
    *addr_a = base + (offb >> 8);
    *addr_b = base + (offb >> 8) + 1;
    *straddle = ((offb & 255u) + rrr_bits(cls)) > 256u;
    return SG_OFFSET;
}

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

FMGPU_HD void sg_on_offset(const Rec32& A, const Rec32& B, bool straddle, const uint16_t* binom, const SgSt& s, uint32_t* bit,
                           uint32_t* rank) {
    const uint32_t nb = rrr_bits(s.cls);
    const uint32_t o = s.offb & 255u, wi = o >> 5, sh = o & 31u;
    const uint32_t lo = rec_word(A, wi);
    const uint32_t hi = wi == 7u ? (straddle ? B.w[0] : 0u) : rec_word(A, wi + 1u);
    const unsigned long long both = ((unsigned long long)hi << 32) | lo;
    const uint32_t off = (uint32_t)(both >> sh) & ((1u << nb) - 1u);
    const uint32_t block = rrr_unrank(binom, s.cls, off, s.use + 1u);
    *bit = (block >> s.use) & 1u;
    *rank = s.ones + popc32(block & ((1u << s.use) - 1u));
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

FMGPU_HD EubOut eub_right_chunks(int32_t from, int32_t down_len, int32_t rel, int32_t length, int32_t dst_len, bool right_only) {
    EubOut o;
    o.status = 0;
    o.value = 0;
    int32_t final_pos = -1, times = 1;
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
            const int32_t limit = right_only ? top : down_len + top;
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
