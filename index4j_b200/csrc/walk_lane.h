// The LF-walk lane machine: one lane = one work item of locate / extract / extractUntilBoundary*.
// Host/device code (see lane_logic.h): the kernels drive it with 256-bit loads, the host layout
// test drives it with plain memory reads and compares against the CPU oracle.
//
// Reference semantics (fm/FmIndex.java): locate :526-548, extract :564-608,
// extractUntilBoundary :640-759, ...Left :772-831, ...Right :844-922; one LF step is
// `c = (short) inverseSelect(j-1); j = C[c] + rank(j, c)` (:532-535, :597-599).
//
// LF step on the device: inverseSelect(j-1) walks DOWN the block's tree reading, per level, one
// level sector and the node record of the node it is in; its leaf gives the symbol AND
// rank(j-1, c).  Because bwt[j-1] == c, rank(j, c) == rank(j-1, c) + 1 whenever j lies in the
// same block as j-1, so the reference's second tree walk is skipped; when j starts a new block
// (or the block is a single-symbol run, where the reference decodes only the low byte of the
// symbol) the generic rank(j, c) path is taken, exactly as the reference would.
#pragma once
#include <cstdint>

#include "lane_logic.h"

namespace fmgpu {

enum WalkMode : int { WM_LOCATE = 0, WM_EXTRACT = 1, WM_EUB = 2 };

enum WalkPhase : uint32_t {
    W_IDLE = 0,
    W_ITEM,    // item inputs are being loaded
    W_ISA,     // ISA sample record
    W_ACCESS,  // sampled-row group record (+ speculative block descriptor)
    W_OFFSET,  // sampled-row offset bits
    W_BLOCK,   // block descriptor
    W_LEVEL,   // level sector (+ node record)
    W_RCELL,   // generic rank: cell
    W_RLEVEL,  // generic rank: level sector
    W_ROVF,    // generic rank: path chunk
    W_SA,      // SA sample record
    W_EXIT
};

struct WalkParams {
    uint32_t n_items;
    // locate: SA rows in, text positions out (in place)
    uint32_t* rows_pos;
    // extract
    const int32_t* start;
    const int32_t* stop;
    const uint64_t* arena_off;
    // extractUntilBoundary*
    const int32_t* from;
    uint32_t mb;        // alphabet code of the boundary char (0 = not in the alphabet)
    int32_t dst_len;
    int32_t eub_mode;   // FMGPU_MODE_*
    uint16_t* left;     // left-part scratch, item i at i*dst_len, chars in walk order (text order reversed)
    int32_t* down_len;  // chars in the left part
    // outputs
    uint16_t* arena;
    int32_t* len_out;
    int32_t* status_out;
    const uint16_t* binom;  // 15x16 binomial table (shared memory on the device)
};

struct ItemRaw {
    int32_t a, b;
    uint64_t o0, o1;
};

template <int MODE>
FMGPU_HD ItemRaw walk_load_item(const WalkParams& P, uint32_t w) {
    ItemRaw r;
    r.a = r.b = 0;
    r.o0 = r.o1 = 0;
    if (MODE == WM_LOCATE) {
        r.a = (int32_t)P.rows_pos[w];
    } else if (MODE == WM_EXTRACT) {
        r.a = P.start[w];
        r.b = P.stop[w];
        r.o0 = P.arena_off[w];
        r.o1 = P.arena_off[w + 1];
    } else {
        r.a = P.from[w];
    }
    return r;
}

struct WalkCounters {
    uint32_t lf_steps, lf_levels, ranks, rank_levels, sbits;
};

template <int MODE>
struct WalkLane {
    uint32_t phase;
    uint32_t w;          // item index
    uint32_t j;          // Java's j / samplePosition: row + 1
    uint32_t sym;        // symbol of the LF step in flight
    uint32_t dist;       // LF steps taken (locate: distance; extract: distance)
    const Rec32* addr_a;
    const Rec32* addr_b;
    bool need_b, straddle, have_rec;
    LfSt lf;
    RankSt rk;
    SgSt sg;
    // extract / eub
    int32_t from, skip, remaining, k, down, cur, end, pb, rel, stage;
    uint64_t out0;  // first arena element of the item
    uint32_t isa_idx;

    FMGPU_HD void init() {
        phase = W_IDLE;
        need_b = false;
        straddle = false;
        have_rec = false;
        addr_a = addr_b = nullptr;
        rk.p5 = rk.p6 = rk.p7 = 0;
    }
    FMGPU_HD bool needs_a() const { return phase >= W_ISA && phase <= W_SA; }

    FMGPU_HD void finish(const WalkParams& P, int32_t status, int32_t value) {
        if (MODE == WM_LOCATE) {
            if (status) P.rows_pos[w] = 0xffffffffu;
            if (status && P.status_out) P.status_out[w] = status;  // per hit; folded into per-pattern status by the caller
        } else {
            P.len_out[w] = value;
            P.status_out[w] = status;
            if (MODE == WM_EUB) P.down_len[w] = (status == 0 || status == 8) ? down : 0;
        }
        phase = W_IDLE;
    }

    // --- LF step plumbing -----------------------------------------------------------------
    FMGPU_HD void start_lf(const DevIndex& ix, const SmemTables& T) {  // LF from row j: inverseSelect(j-1)
        lf_begin(ix, T, j - 1u, lf, &addr_a);
        need_b = false;
        have_rec = true;  // the root's record arrives with the block descriptor
        phase = W_BLOCK;
    }
    FMGPU_HD void start_isa(const DevIndex& ix, uint32_t idx) {
        isa_idx = idx;
        addr_a = ix.isa + (idx >> 3);
        need_b = false;
        phase = W_ISA;
    }
    FMGPU_HD void start_access(const DevIndex& ix, const SmemTables& T) {  // sampled.access(j-1), block descriptor speculatively
        addr_a = sg_addr(ix, j - 1u);
        const Rec32* blk;
        lf_begin(ix, T, j - 1u, lf, &blk);
        addr_b = blk;
        need_b = true;
        phase = W_ACCESS;
    }
    FMGPU_HD void start_generic_rank(const DevIndex& ix, const SmemTables& T, const WalkParams& P, WalkCounters& cnt) {
        uint32_t val = 0;
        need_b = false;
        const uint32_t o = rank_begin(ix, T, j, sym, rk, &addr_a, &val);
        if (o == RK_MORE) phase = W_RCELL;
        else if (o == RK_THROW) finish(P, 9, 0);
        else lf_complete(ix, T, P, val, cnt);
    }
    FMGPU_HD void on_block(const DevIndex& ix, const SmemTables& T, const WalkParams& P, const Rec32& D, WalkCounters& cnt) {
        ++cnt.lf_steps;
        uint32_t s = 0;
        const uint32_t o = lf_on_block(ix, D, lf, &addr_a, &s);
        if (o == LF_RUN) {
            sym = s;
            start_generic_rank(ix, T, P, cnt);
        } else {
            need_b = false;
            have_rec = true;
            phase = W_LEVEL;
        }
    }

    // --- per-mode: an LF step finished; `sym` is its symbol, `rank_j` = rank(j, sym) ------------
    FMGPU_HD void lf_complete(const DevIndex& ix, const SmemTables& T, const WalkParams& P, uint32_t rank_j, WalkCounters& cnt) {
        j = T.C[sym] + rank_j;
        if (MODE == WM_LOCATE) {
            ++dist;
            start_access(ix, T);
        } else if (MODE == WM_EXTRACT) {
            if ((int32_t)dist >= skip) {  // :601-604
                P.arena[out0 + (uint64_t)(remaining - 1)] = ix.code2char[sym];
                --remaining;
            }
            ++dist;
            if (remaining > 0) start_lf(ix, T);
            else finish(P, 0, k);
        } else {
            eub_on_char(ix, T, P);
        }
        (void)cnt;
    }

    // --- extractUntilBoundary*: what to do with the char an LF step produced -------------------
    FMGPU_HD void eub_begin_right(const DevIndex& ix, const WalkParams& P) {
        // interval [cur, end): walk from the ISA sample at `end` (wrap sample in the last interval)
        const uint32_t sr = ix.sample_rate;
        const uint32_t s = (uint32_t)cur / sr;
        const uint64_t e = ((uint64_t)s + 1u) * sr;
        end = e < (uint64_t)ix.length ? (int32_t)e : (int32_t)ix.length;
        dist = 0;
        pb = -1;
        (void)P;
        start_isa(ix, s + 1u);
    }
    FMGPU_HD void eub_finish_right(const WalkParams& P, const DevIndex& ix, int32_t rel) {
        const EubOut o = eub_right_chunks(from, down, rel, (int32_t)ix.length, P.dst_len, P.eub_mode == 2);
        finish(P, o.status, o.value);
    }
    FMGPU_HD void eub_on_char(const DevIndex& ix, const SmemTables& T, const WalkParams& P) {
        const uint64_t slot = (uint64_t)w * (uint64_t)P.dst_len;
        if (stage == 0) {  // left walk (:664-686, :797-826)
            bool stop_left = false;
            if ((int32_t)dist >= skip) {
                if (sym == P.mb || sym == 0u) {
                    stop_left = true;
                } else if (P.eub_mode == 1) {
                    const int32_t idx = P.dst_len - 1 - k;  // downStreamPos
                    if (idx < 0) {
                        finish(P, 9, 0);
                        return;
                    }
                    P.left[slot + (uint64_t)k] = ix.code2char[sym];
                    ++k;
                    down = k;
                    if (idx - 1 == 0) {  // :817-821
                        finish(P, 8, P.dst_len);
                        return;
                    }
                } else {
                    P.left[slot + (uint64_t)k] = ix.code2char[sym];
                    ++k;
                    --remaining;
                    if (remaining == 0) stop_left = true;
                }
            }
            ++dist;
            if (!stop_left) {
                start_lf(ix, T);
                return;
            }
            down = k;
            if (P.eub_mode == 1) {
                finish(P, 0, k);
                return;
            }
            stage = 1;
            cur = from;
            eub_begin_right(ix, P);
            return;
        }
        // right walk over one sample interval: step `dist` produced text[end-1-dist]
        const int32_t p = end - 1 - (int32_t)dist;
        if (p >= from && p < (int32_t)ix.length - 1) {
            if (sym == P.mb) pb = p;
            if (P.eub_mode == 2) {
                const int32_t idx = p - from - 1;
                if (idx >= 0 && idx < P.dst_len) P.arena[slot + (uint64_t)idx] = ix.code2char[sym];
            } else {
                const int64_t idx = (int64_t)down + (p - from);
                if (idx < P.dst_len) P.arena[slot + (uint64_t)idx] = ix.code2char[sym];
            }
        }
        ++dist;
        if (p > cur) {
            start_lf(ix, T);
            return;
        }
        if (rel < 0 && pb >= 0) rel = pb - from;  // intervals go left to right: the first boundary seen is the nearest
        bool more = true;
        if (rel >= 0) {
            // The reference reads whole 4-char chunks; when the boundary's chunk is also the one that reaches the
            // end of the text, its end-of-text rule (quirk Q5) returns chars beyond the boundary: fetch them too.
            const int64_t chunk_end = (int64_t)from + 4 * ((int64_t)rel / 4 + 1);
            more = chunk_end >= (int64_t)ix.length - 1 && end < (int32_t)ix.length;
        } else if (end >= (int32_t)ix.length) {
            more = false;
        } else if ((int64_t)end - from > (int64_t)P.dst_len + 8) {
            rel = 0x3fffffff;  // the destination overflows before any boundary
            more = false;
        }
        if (more) {
            cur = end;
            eub_begin_right(ix, P);
        } else {
            eub_finish_right(P, ix, rel);
        }
    }

    // --- item start ---------------------------------------------------------------------------
    FMGPU_HD void begin_item(const DevIndex& ix, const SmemTables& T, const WalkParams& P, const ItemRaw& raw) {
        dist = 0;
        k = 0;
        down = 0;
        stage = 0;
        rel = -1;
        if (MODE == WM_LOCATE) {
            j = (uint32_t)raw.a + 1u;  // :527-529
            start_access(ix, T);
        } else if (MODE == WM_EXTRACT) {
            const int32_t st = raw.a, sp = raw.b;
            out0 = raw.o0;
            if (!ix.extract_enabled) return finish(P, 1, 0);
            if (st < 0) return finish(P, 2, 0);
            if (sp >= (int32_t)ix.length) return finish(P, 3, 0);
            const int32_t sr = (int32_t)ix.sample_rate;
            const int32_t idx = sp / sr + 1;
            if (idx < 0 || idx >= (int32_t)ix.n_isa) return finish(P, 9, 0);
            skip = sr - sp % sr;
            if (sp / sr == (int32_t)ix.n_isa - 2) skip = (int32_t)ix.length - sp;
            const int32_t range = sp - st;
            const int64_t room = (int64_t)(raw.o1 - raw.o0);
            if (room < (int64_t)range) return finish(P, 5, 0);
            remaining = range;
            k = range;
            if (range <= 0) return finish(P, 0, range);
            start_isa(ix, (uint32_t)idx);
        } else {
            from = raw.a;
            if (P.eub_mode == 1) ++from;  // :774
            if (!ix.extract_enabled) return finish(P, 1, 0);
            if (from < 0) return finish(P, 2, 0);
            if (from >= (int32_t)ix.length) return finish(P, 4, 0);
            if (P.dst_len == 0) return finish(P, 6, 0);
            if (P.mb == 0u) return finish(P, 7, 0);
            remaining = P.dst_len;
            if (P.eub_mode == 2) {
                stage = 1;
                cur = from;
                eub_begin_right(ix, P);
            } else {
                const int32_t sr = (int32_t)ix.sample_rate;
                skip = sr - from % sr;
                if (from / sr == (int32_t)ix.n_isa - 2) skip = (int32_t)ix.length - from;
                start_isa(ix, (uint32_t)(from / sr + 1));
            }
        }
    }

    // --- one trip: consume the records fetched for the current phase ------------------------------
    FMGPU_HD void step(const DevIndex& ix, const SmemTables& T, const WalkParams& P, const ItemRaw& raw, const Rec32& A, const Rec32& B,
                       WalkCounters& cnt) {
        switch (phase) {
            case W_ITEM:
                begin_item(ix, T, P, raw);
                break;
            case W_ISA:
                j = rec_word(A, isa_idx & 7u) + 1u;  // :579 / :645
                start_lf(ix, T);
                break;
            case W_ACCESS: {
                ++cnt.sbits;
                uint32_t bit = 0, rank = 0;
                const Rec32* oa = nullptr;
                const Rec32* ob = nullptr;
                bool st = false;
                const uint32_t o = sg_on_group(ix, A, j - 1u, sg, &bit, &rank, &oa, &ob, &st);
                if (o == SG_OFFSET) {
                    addr_a = oa;
                    addr_b = ob;
                    need_b = st;
                    straddle = st;
                    phase = W_OFFSET;
                } else if (bit) {
                    isa_idx = rank;  // suffixes[rankOnes(j) - 1], rankOnes(j) = rankOnes(j-1) + 1  (:538-542)
                    addr_a = ix.sa + (rank >> 3);
                    need_b = false;
                    phase = W_SA;
                } else {
                    on_block(ix, T, P, B, cnt);  // the speculative block descriptor is the LF step's first record
                }
                break;
            }
            case W_OFFSET: {
                uint32_t bit = 0, rank = 0;
                sg_on_offset(A, B, straddle, P.binom, sg, &bit, &rank);
                if (bit) {
                    isa_idx = rank;
                    addr_a = ix.sa + (rank >> 3);
                    need_b = false;
                    phase = W_SA;
                } else {
                    start_lf(ix, T);
                }
                break;
            }
            case W_SA:
                P.rows_pos[w] = rec_word(A, isa_idx & 7u) + dist;
                phase = W_IDLE;
                break;
            case W_BLOCK:
                on_block(ix, T, P, A, cnt);
                break;
            case W_LEVEL: {
                ++cnt.lf_levels;
                if (!have_rec) lf_take_record(B, lf);
                uint32_t s = 0, r = 0;
                const uint32_t o = lf_on_level(ix, A, lf, &addr_a, &addr_b, &s, &r);
                if (o == LF_MORE) {
                    need_b = true;
                    have_rec = false;
                } else {
                    sym = s;
                    need_b = false;
                    if ((j & lf.bmask) != 0u) lf_complete(ix, T, P, r + 1u, cnt);  // rank(j, c) = rank(j-1, c) + 1 inside a block
                    else start_generic_rank(ix, T, P, cnt);
                }
                break;
            }
            case W_RCELL: {
                ++cnt.ranks;
                uint32_t val = 0;
                const uint32_t o = rank_on_cell(ix, A, rk, &addr_a, &val);
                if (o == RK_MORE) phase = W_RLEVEL;
                else if (o == RK_THROW) finish(P, 9, 0);
                else lf_complete(ix, T, P, val, cnt);
                break;
            }
            case W_RLEVEL: {
                ++cnt.rank_levels;
                uint32_t val = 0;
                bool want_ovf = false;
                const uint32_t o = rank_on_level(ix, A, rk, &addr_a, &val, &want_ovf);
                if (o == RK_DONE) lf_complete(ix, T, P, val, cnt);
                else if (want_ovf) phase = W_ROVF;
                break;
            }
            case W_ROVF:
                rank_on_ovf(ix, A, rk, &addr_a);
                phase = W_RLEVEL;
                break;
            default:
                break;
        }
    }
};

}  // namespace fmgpu
