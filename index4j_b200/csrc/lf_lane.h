// Per-lane code of one LF step, of the sampled-row test and of the extract / extractUntilBoundary* work item (lockstep kernels
// k_locate / k_extract).
//
// Plain sequential functions: a warp runs them for its 32 work items together and the hardware reconverges the lanes after
// each data-dependent loop (the round-1 phase machine, which executed every phase's code every trip, needed ~6x more issued
// instructions per LF step).  Host/device code: the kernels fetch records with 256-bit loads, the host layout test
// (tests/support/flatcheck.cpp) with plain reads.
//
// Reference semantics (paths under indices/src/main/java/com/dynatrace/):
//   sampled_access_rank  bitsequence/RrrVector.java:314-349 (access) and :358-396 (rankOnes)
//   lf_step              fm/FmIndex.java:532-535 / :597-599  c = (short) inverseSelect(j-1); j = C[c] + rank(j, c)
//                        with inverseSelect = wavelet/WaveletFixedBlockBoosting.java:1305-1537
#pragma once
#include <cstdint>

#include "lane_logic.h"
#include "ldrec.h"
#include "walk_lane.h"  // WalkParams, ItemRaw, walk_load_item, WalkMode

namespace fmgpu {

// (class, offset) -> 15-bit block: inv[cbase[cls] + off] — the reference's INVERSE_VALUES / CARDINALITY_OFFSETS
// tables (RrrVector.java:8692-16899), regenerated from their ordering rule; shared memory on the device.
struct RrrTab {
    const uint16_t* inv;    // [32768]
    const uint16_t* cbase;  // [16]
};

struct LfCounters {
    uint32_t lf_steps, lf_levels, ranks, rank_levels, sbits, recs;  // recs: level records fetched (descents + generic ranks)
};

// access(pos) and rankOnes(pos) of the sampled-row vector from its group record G (already fetched)
FMGPU_HD void sampled_access_rank(const DevIndex& ix, const RrrTab& R, const Rec32& G, uint32_t pos, uint32_t* bit, uint32_t* rank) {
    const uint32_t blk = pos / RRR_BLOCK;
    const uint32_t k = blk & 31u, sub = k >> 3, kk = k & 7u;
    const uint32_t use = pos - blk * RRR_BLOCK;
    uint32_t offb = G.w[1], ones = G.w[0];
    if (sub) {
        offb += (G.w[2] >> (10u * (sub - 1u))) & 1023u;
        ones += (G.w[3] >> (10u * (sub - 1u))) & 1023u;
    }
    const uint32_t word = rec_word(G, 4u + sub);
    // classes of the blocks before this one inside its 8-block word: sum of nibbles / of their offset widths
    const uint32_t before = word & low_mask_clamped((int)(4u * kk));
#pragma unroll
    for (uint32_t i = 0; i < 7; ++i) {
        const uint32_t c = (before >> (4u * i)) & 15u;
        ones += c;
        offb += i < kk ? rrr_bits(c) : 0u;
    }
    const uint32_t cls = (word >> (4u * kk)) & 15u;
    if (cls == 0u) {
        *bit = 0;
        *rank = ones;
        return;
    }
    if (cls == 15u) {
        *bit = 1;
        *rank = ones + use;
        return;
    }
    const uint32_t nb = rrr_bits(cls);
    const uint32_t wi = offb >> 5, sh = offb & 31u;
    const uint32_t lo = FMGPU_LDG32(ix.soffsets + wi);
    const uint32_t hi = FMGPU_LDG32(ix.soffsets + wi + 1u);  // the stream is padded
    const unsigned long long both = ((unsigned long long)hi << 32) | lo;
    const uint32_t off = (uint32_t)(both >> sh) & ((1u << nb) - 1u);
    const uint32_t block = R.inv[((uint32_t)R.cbase[cls] + off) & 32767u];  // the mask only matters for a corrupt offset
    *bit = (block >> use) & 1u;
    *rank = ones + popc32(block & ((1u << use) - 1u));
}

// The same test on the device-side dense marks (layout.h: DevIndex::dmarks): is row `pos` marked, and how many marked rows
// precede it (= its index in dsa)?  G = dmarks[pos / 224], already fetched.
FMGPU_HD void dense_access_rank(const Rec32& G, uint32_t pos, uint32_t* bit, uint32_t* rank) {
    const uint32_t o = pos % DENSE_ROWS_PER_REC, wi = o >> 5, sh = o & 31u;
    const uint32_t word = rec_word(G, 1u + wi);
    uint32_t r = G.w[0];
#pragma unroll
    for (uint32_t k = 0; k < 6; ++k) r += k < wi ? popc32(G.w[1 + k]) : 0u;
    *bit = (word >> sh) & 1u;
    *rank = r + popc32(word & ((1u << sh) - 1u));
}

// One LF step from row j (Java's 1-based j): returns the new j.  D is the block descriptor of position j-1
// (already fetched), bmask the block-size mask of that position's superblock.
//   * tree block: inverseSelect walks DOWN the block's tree, two levels per level record (plus the node record of
//     the even-depth node, fetched alongside), and its leaf gives the symbol AND rank(j-1, c); bwt[j-1] == c, so
//     rank(j, c) = rank(j-1, c) + 1 whenever j lies in the same block;
//   * single-symbol block: the descriptor carries the symbol as inverseSelect decodes it (low byte only,
//     :1329-1332) and the pre-evaluated (block, symbol) cell, so rank(j, c) = value [+ (j mod block)] needs no
//     further record when j lies in the same block;
//   * otherwise (j starts a new block) the generic rank(j, c) is evaluated exactly as the reference would.
FMGPU_HD uint32_t lf_step(const DevIndex& ix, const SmemTables& T, const Rec32& D, uint32_t j, uint32_t bmask, uint32_t* sym_out,
                          uint32_t* err, LfCounters& cnt) {
    const uint32_t pos = j - 1u;
    const uint32_t jrel = j & bmask;
    uint32_t sym = 0, rank_j = 0;
    bool have = false;
    ++cnt.lf_steps;
    if (D.w[1] & 1u) {
        sym = (D.w[1] >> 8) & 0xffffu;
        if (jrel != 0u) {
            const uint32_t kind = D.w[3];
            if (kind == CELL_THROW) {
                *err = 1;
                *sym_out = sym;
                return j;
            }
            rank_j = D.w[2] + (kind == CELL_RUN ? jrel : 0u);
            have = true;
        }
    } else {
        uint32_t r = pos & bmask;
        uint32_t sec = D.w[0], nrec = D.w[4];
        for (;;) {
            const Rec32 X = FMGPU_LD256(ix.sectors + (sec + r / SECTOR_BITS));
            const Rec32 N = FMGPU_LD256(ix.nodes + nrec);
            uint32_t rk = 0;
            ++cnt.recs;
            if (dlevel_descend(X, N, r % SECTOR_BITS, &r, &nrec, &sec, &sym, &rk, &cnt.lf_levels)) {
                rank_j = rk + 1u;
                break;
            }
        }
        have = jrel != 0u;
    }
    *sym_out = sym;
    if (!have) {
        const uint32_t st = rank_single(ix, T, j, sym, &rank_j, &cnt.ranks, &cnt.rank_levels, &cnt.recs);
        if (st) {
            *err = 1;
            return j;
        }
    }
    // 1 <= LF(j) <= length - 1 on a consistent index; the clamps keep a corrupt (but loadable) one inside the tables
    const uint32_t jn = T.C[sym] + rank_j;
    return jn == 0u ? 1u : (jn < ix.length ? jn : ix.length);
}

// One item of the dense-sample build (kernels_dense.cuh): from `row`, whose suffix starts at text position p, LF-walk towards
// the start of the text and hand every (row, position) visited to `visit`, up to (not including) the next multiple of
// sampleRate — that row is a sampled row and the start of its own item.  Returns false where an LF step fails.
template <typename Visit>
FMGPU_HD bool dense_walk_item(const DevIndex& ix, const SmemTables& T, uint32_t row, uint32_t p, Visit&& visit) {
    LfCounters cnt;
    cnt.lf_steps = cnt.lf_levels = cnt.ranks = cnt.rank_levels = cnt.sbits = cnt.recs = 0;
    visit(row, p);
    while (p != 0u) {
        const SbDesc sd = T.sb[row >> SB_LOG];
        const uint32_t blk = sd.first_block + ((row & SB_MASK) >> sd.block_log);
        const Rec32 D = FMGPU_LD256(ix.blocks + blk);
        uint32_t sym = 0, err = 0;
        const uint32_t jn = lf_step(ix, T, D, row + 1u, (1u << sd.block_log) - 1u, &sym, &err, cnt);
        if (err) return false;
        row = jn - 1u;
        --p;
        if (p % ix.sample_rate == 0u) break;
        visit(row, p);
    }
    return true;
}
// the items: k < n_seeds = the k-th sampled row (position suffixes[k]; n_seeds = ones of the sampled-row vector — suffixes[]
// itself has length / sampleRate + 1 entries, one more than there are sampled rows when sampleRate divides length,
// fm/FmIndex.java:343-344); k == n_seeds = row 0, the sentinel's suffix (position length - 1), which covers the positions
// behind the last multiple of sampleRate.  Returns 0 = walk it, 1 = nothing to do, 2 = not a consistent index.
FMGPU_HD uint32_t dense_seed_count(const DevIndex& ix) { return ix.s_total_ones < ix.n_sa ? ix.s_total_ones : ix.n_sa; }
FMGPU_HD int dense_item(const DevIndex& ix, const uint32_t* seed_row, uint32_t n_seeds, uint64_t k, uint32_t* row, uint32_t* p) {
    if (k < n_seeds) {
        *row = seed_row[k];
        *p = reinterpret_cast<const uint32_t*>(ix.sa)[k];
        return (*row >= ix.length || *p >= ix.length || *p % ix.sample_rate != 0u) ? 2 : 0;
    }
    *row = 0;
    *p = ix.length - 1u;
    return *p % ix.sample_rate == 0u ? 1 : 0;  // 1: the sentinel's row is a sampled row itself
}

// ------------------------------------------------------------------------------------------------
// extract / extractUntilBoundary* work item of the lockstep kernel k_extract: the control flow around the LF
// steps (fm/FmIndex.java:564-608, :640-759, :772-831, :844-922), expressed as "what happens after an LF step produced a
// char" so that every trip of the warp loop is: [ISA sample if a walk starts] -> one LF step -> on_char().
// ------------------------------------------------------------------------------------------------
template <int MODE>
struct ExLane {
    uint32_t w, j, dist, isa_idx;
    bool active, need_isa;
    int32_t from, skip, remaining, k, down, cur, end, pb, rel, stage;
    uint64_t out0;

    FMGPU_HD void init() {
        active = false;
        need_isa = false;
        w = j = dist = isa_idx = 0;
        from = skip = remaining = k = down = cur = end = pb = rel = stage = 0;
        out0 = 0;
    }
    FMGPU_HD void finish(const WalkParams& P, int32_t status, int32_t value) {
        P.len_out[w] = value;
        P.status_out[w] = status;
        if (MODE == WM_EUB && P.down_len) P.down_len[w] = (status == 0 || status == 8) ? down : 0;
        active = false;
    }
    FMGPU_HD void start_isa(uint32_t idx) {
        isa_idx = idx;
        need_isa = true;
    }
    FMGPU_HD void begin_right(const DevIndex& ix) {  // interval [cur, end): walk from the ISA sample at `end`
        const uint32_t sr = ix.sample_rate;
        const uint32_t s = (uint32_t)cur / sr;
        const uint64_t e = ((uint64_t)s + 1u) * sr;
        end = e < (uint64_t)ix.length ? (int32_t)e : (int32_t)ix.length;
        dist = 0;
        pb = -1;
        start_isa(s + 1u);
    }
    // checks of extract (:566-593) / checkBoundsForExtraction (:610-626) and the first ISA sample
    FMGPU_HD void begin(const DevIndex& ix, const WalkParams& P, uint32_t item, const ItemRaw& raw) {
        w = item;
        active = true;
        need_isa = false;
        dist = 0;
        k = 0;
        down = 0;
        stage = 0;
        rel = -1;
        if (MODE == WM_EXTRACT) {
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
            const int64_t room = (int64_t)(raw.o1 - raw.o0) - (int64_t)P.offset;  // destination.length - offset (:591)
            if (room < (int64_t)range) return finish(P, 5, 0);
            out0 += (uint64_t)P.offset;
            remaining = range;
            k = range;
            if (range <= 0) return finish(P, 0, range);
            start_isa((uint32_t)idx);
        } else {
            from = raw.a;
            if (P.eub_mode == EUB_SCAN) {  // before any early exit: a hit the reference throws on owns no record
                P.win_of[w] = -1;
                P.at_bound[w] = 0;
            }
            if (P.eub_mode == 1) ++from;  // :774
            if (!ix.extract_enabled) return finish(P, 1, 0);
            if (from < 0) return finish(P, 2, 0);
            if (from >= (int32_t)ix.length) return finish(P, 4, 0);
            if (P.dst_len == 0) return finish(P, 6, 0);
            if (P.mb == 0u) return finish(P, 7, 0);
            remaining = P.dst_len;
            if (P.eub_mode == 2 || P.eub_mode == EUB_RECORD) {
                stage = 1;
                cur = from;
                begin_right(ix);
            } else {
                const int32_t sr = (int32_t)ix.sample_rate;
                skip = sr - from % sr;
                if (from / sr == (int32_t)ix.n_isa - 2) skip = (int32_t)ix.length - from;
                start_isa((uint32_t)(from / sr + 1));
            }
        }
    }
    // the record that starts at text position S belongs to the first hit that claims it; returns that hit's id
    FMGPU_HD uint32_t claim_record(const WalkParams& P, uint32_t S, uint32_t me) {
        const unsigned long long mine = ((unsigned long long)(S + 1u) << 32) | me;
        uint32_t slot = (S * 2654435761u) & P.claim_mask;
        for (;;) {
#if defined(__CUDA_ARCH__)
            const unsigned long long old = atomicCAS(P.claims + slot, 0ull, mine);
#else
            const unsigned long long old = P.claims[slot];
            if (old == 0ull) P.claims[slot] = mine;
#endif
            if (old == 0ull) return me;
            if ((uint32_t)(old >> 32) == S + 1u) return (uint32_t)old;
            slot = (slot + 1u) & P.claim_mask;
        }
    }
    // an LF step produced `sym` (the char left of the previous one)
    FMGPU_HD void on_char(const DevIndex& ix, const WalkParams& P, uint32_t sym) {
        if (MODE == WM_EXTRACT) {
            if ((int32_t)dist >= skip) {  // :601-604
                P.arena[out0 + (uint64_t)(remaining - 1)] = ix.code2char[sym];
                --remaining;
            }
            ++dist;
            if (remaining <= 0) finish(P, 0, k);
            return;
        }
        const uint64_t slot = (uint64_t)w * (uint64_t)P.dst_len;
        if (MODE == WM_EUB && P.eub_mode == EUB_SCAN) {
            // fused locate -> extractUntilBoundary: the left walk of FmIndex.extractUntilBoundary (:664-686) without storing its
            // chars — it finds the start S of the hit's record, which the first hit to arrive claims
            if ((int32_t)dist >= skip) {
                if (sym == P.mb || sym == 0u) {
                    down = k;
                    P.win_of[w] = (int32_t)claim_record(P, (uint32_t)(from - k), w);
                    return finish(P, 0, k);
                }
                ++k;
                --remaining;
                if (remaining == 0) {  // the left part alone fills the destination: no record start (win_of stays -1)
                    down = k;
                    return finish(P, 0, k);
                }
            } else if ((int32_t)dist + 1 == skip) {
                P.at_bound[w] = sym == P.mb ? 1 : 0;  // this step produced text[from]
            }
            ++dist;
            return;
        }
        if (stage == 0) {  // left walk (:664-686, :797-826)
            bool stop_left = false;
            if ((int32_t)dist >= skip) {
                if (sym == P.mb || sym == 0u) {
                    stop_left = true;
                } else if (P.eub_mode == 1) {
                    const int32_t idx = P.dst_len - 1 - k;  // downStreamPos
                    if (idx < 0) return finish(P, 9, 0);
                    P.left[slot + (uint64_t)k] = ix.code2char[sym];
                    ++k;
                    down = k;
                    if (idx - 1 == P.offset) return finish(P, 8, P.dst_len - P.offset);  // :817-821
                } else {
                    P.left[slot + (uint64_t)k] = ix.code2char[sym];
                    ++k;
                    --remaining;
                    if (remaining == 0) stop_left = true;
                }
            }
            ++dist;
            if (!stop_left) return;
            down = k;
            // System.arraycopy(destination, downStreamPos + 1, destination, offset, downStreamLength) (:688-690, :827-829)
            // throws when the left part does not fit behind `offset`
            if ((int64_t)P.offset + k > (int64_t)P.dst_len) return finish(P, 9, 0);
            if (P.eub_mode == 1) return finish(P, 0, k);
            stage = 1;
            cur = from;
            begin_right(ix);
            return;
        }
        // right walk over one sample interval: step `dist` produced text[end-1-dist]
        const int32_t p = end - 1 - (int32_t)dist;
        if (p >= from && p < (int32_t)ix.length - 1) {
            if (sym == P.mb) pb = p;
            if (P.eub_mode == EUB_RECORD) {
                const int64_t idx = (int64_t)p - from;
                if (idx < P.dst_len) P.arena[slot + (uint64_t)idx] = ix.code2char[sym];
            } else if (P.eub_mode == 2) {
                const int64_t idx = (int64_t)P.offset + (p - from - 1);
                if (p - from - 1 >= 0 && idx < P.dst_len) P.arena[slot + (uint64_t)idx] = ix.code2char[sym];
            } else {
                const int64_t idx = (int64_t)P.offset + down + (p - from);
                if (idx < P.dst_len) P.arena[slot + (uint64_t)idx] = ix.code2char[sym];
            }
        }
        ++dist;
        if (p > cur) return;
        if (rel < 0 && pb >= 0) rel = pb - from;  // intervals go left to right: the first boundary seen is the nearest
        bool more = true;
        if (P.eub_mode == EUB_RECORD) {
            // text[S, E) of a distinct record: stop at the first boundary — or go on to the end of the text when the record ends
            // within a few chars of it (the reference's end-of-text rule, quirk Q5, returns chars beyond the boundary there, and
            // which ones depends on the hit) — or when the destination is certainly too small
            int32_t value = rel;
            if (rel >= 0) {
                more = (int64_t)from + rel >= (int64_t)ix.length - 10 && end < (int32_t)ix.length;
            } else if (end >= (int32_t)ix.length) {
                more = false;  // the text ends first: rel stays -1
            } else if ((int64_t)end - from > (int64_t)P.dst_len + 8) {
                more = false;
                value = REL_NONE;
            }
            if (more) {
                cur = end;
                begin_right(ix);
            } else {
                finish(P, 0, value);
            }
            return;
        }
        if (rel >= 0) {
            // The reference reads whole 4-char chunks; when the boundary's chunk is also the one that reaches the end of
            // the text, its end-of-text rule (quirk Q5) returns chars beyond the boundary: fetch them too.
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
            begin_right(ix);
        } else {
            const EubOut o = eub_right_chunks(from, down, rel, (int32_t)ix.length, P.dst_len, P.eub_mode == 2, P.offset);
            finish(P, o.status, o.value);
        }
    }
    // one trip: [ISA sample] -> LF step -> on_char
    FMGPU_HD void trip(const DevIndex& ix, const SmemTables& T, const WalkParams& P, LfCounters& cnt) {
        if (need_isa) {
            const Rec32 I = FMGPU_LD256(ix.isa + (isa_idx >> 3));
            j = rec_word(I, isa_idx & 7u) + 1u;  // :579 / :645
            need_isa = false;
        }
        const uint32_t pos = j - 1u;
        const SbDesc sd = T.sb[pos >> SB_LOG];
        const uint32_t blk = sd.first_block + ((pos & SB_MASK) >> sd.block_log);
        const uint32_t bmask = (1u << sd.block_log) - 1u;
        const Rec32 D = FMGPU_LD256(ix.blocks + blk);
        uint32_t sym = 0, err = 0;
        const uint32_t jn = lf_step(ix, T, D, j, bmask, &sym, &err, cnt);
        if (err) return finish(P, 9, 0);
        j = jn;
        on_char(ix, P, sym);
    }
};

}  // namespace fmgpu
