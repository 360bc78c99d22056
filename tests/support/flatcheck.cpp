// TEST SUPPORT (not product code): replays the product's host re-layout (flatten.hpp) and the
// per-lane logic of the kernels (lane_logic.h, lf_lane.h) on the CPU, one lane at a time, so that
// tests can compare the device data layout and lane state machines with the oracle without a GPU.
// The kernels run exactly these functions per lane; only the warp-level scheduling differs.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../index4j_b200/csrc/flatten.hpp"
#include "../../index4j_b200/csrc/jstream.hpp"
#include "../../index4j_b200/csrc/count_lane.h"
#include "../../index4j_b200/csrc/lf_lane.h"
#include "../../index4j_b200/csrc/utf8_lane.h"
#include "../../index4j_b200/csrc/host_pack.hpp"

using namespace fmgpu;

namespace {
thread_local std::string g_err;
struct FC {
    fmgpu_host::FlatIndex F;
    DevIndex ix;
    SmemTables T;
    CountTables CT;
    uint16_t binom[15 * 16];
    std::vector<U32x2> kmer;  // q-gram start table built by fc_build_start_table (host twin of build_start_table, fmgpu.cu)
};
const Rec32 ZERO{};

// rank(pos, sym) through the cell / level records; returns status (0 or 9)
int host_rank(const FC& h, uint32_t pos, uint32_t sym, uint32_t* out, uint64_t* n_rank, uint64_t* n_level) {
    uint32_t nr = 0, nl = 0, nrec = 0;
    const uint32_t st = rank_single(h.ix, h.T, pos, sym, out, &nr, &nl, &nrec);
    if (n_rank) *n_rank += nr;
    if (n_level) *n_level += nl;
    return (int)st;
}

template <int MODE>
void run_walk(const FC& h, const WalkParams& P, uint64_t* counters) {
    LfCounters cnt{};
    for (uint32_t w = 0; w < P.n_items; ++w) {
        ExLane<MODE> lane;
        lane.init();
        lane.begin(h.ix, P, w, walk_load_item<MODE>(P, w));
        while (lane.active) lane.trip(h.ix, h.T, P, cnt);
    }
    if (counters) {
        counters[0] += cnt.ranks;
        counters[1] += cnt.rank_levels;
        counters[2] += cnt.lf_steps;
        counters[3] += cnt.lf_levels;
    }
}
}  // namespace

extern "C" {

const char* fc_last_error(void) { return g_err.c_str(); }

int fc_load(const uint8_t* buf, uint64_t len, int threads, void** out) {
    try {
        FC* h = new FC();
        fmgpu_host::JavaIn in(buf, (size_t)len);
        fmgpu_host::FmStream fm;
        fm.read(in);
        fmgpu_host::flatten(fm, threads, h->F);
        h->ix = h->F.meta;
        h->ix.C = h->F.C.data();
        h->ix.char2code = h->F.char2code.data();
        h->ix.code2char = h->F.code2char.data();
        h->ix.sb = h->F.sb.data();
        h->ix.cells = h->F.cells.data();
        h->ix.sectors = h->F.sectors.data();
        h->ix.occ = h->F.occ.data();
        h->ix.blocks = h->F.blocks.data();
        h->ix.nodes = h->F.nodes.data();
        h->ix.sgroups = h->F.sgroups.data();
        h->ix.soffsets = h->F.soffsets.data();
        h->ix.sa = h->F.sa.data();
        h->ix.isa = h->F.isa.data();
        h->T.C = h->ix.C;
        h->T.sb = h->ix.sb;
        h->CT.C = h->ix.C;
        h->CT.sb = h->ix.sb;
        fill_binom(h->binom);
        *out = h;
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -2;
    }
}
void fc_free(void* h) { delete (FC*)h; }
int32_t fc_alphabet_length(void* h) { return ((FC*)h)->F.alphabet_length; }
void fc_sizes(void* hv, uint64_t* out8) {
    FC* h = (FC*)hv;
    out8[0] = h->F.cells.size() * sizeof(Cell8);
    out8[1] = h->F.sectors.size() * 32;
    out8[2] = h->F.nodes.size() * 32;
    out8[3] = h->F.blocks.size() * 32;
    out8[4] = h->F.occ.size() * 32;
    out8[5] = h->F.sgroups.size() * 32 + h->F.soffsets.size() * 4;
    out8[6] = h->F.sa.size() * 32;
    out8[7] = h->F.isa.size() * 32;
}

// design aid: number of (block, symbol) cells by kind {NORMAL, CONST, RUN, THROW}, and blocks / run blocks
void fc_cell_kinds(void* hv, uint64_t* out6) {
    FC* h = (FC*)hv;
    for (int i = 0; i < 6; ++i) out6[i] = 0;
    for (const Cell8& c : h->F.cells) out6[cell_kind(c) & 3u]++;
    out6[4] = h->F.blocks.size();
    for (const Rec32& b : h->F.blocks) out6[5] += b.w[1] & 1u;
}

// (block, symbol) cells by kind (CellKind 0..7)
void fc_cell_kinds8(void* hv, uint64_t* out8) {
    FC* h = (FC*)hv;
    for (int i = 0; i < 8; ++i) out8[i] = 0;
    for (const Cell8& c : h->F.cells) out8[cell_kind(c) & 7u]++;
}

// up to `max` (block, symbol) cells of one kind: out[4 i ..] = {symbol, first text row of the block, block size, 0}
uint32_t fc_find_cells(void* hv, uint32_t kind, uint32_t max, uint32_t* out) {
    FC* h = (FC*)hv;
    const uint32_t sigma = h->ix.sigma;
    const size_t n_blocks = h->F.cells.size() / sigma;
    uint32_t n = 0;
    for (uint32_t sb = 0; sb < h->ix.n_sb && n < max; ++sb) {
        const SbDesc sd = h->F.sb[sb];
        const size_t first = sd.first_block, last = sb + 1 < h->ix.n_sb ? h->F.sb[sb + 1].first_block : n_blocks;
        for (size_t b = first; b < last && n < max; ++b)
            for (uint32_t c = 0; c < sigma && n < max; ++c)
                if (cell_kind(h->F.cells[b * sigma + c]) == kind) {
                    out[4 * n] = c;
                    out[4 * n + 1] = (sb << SB_LOG) + (uint32_t)((b - first) << sd.block_log);
                    out[4 * n + 2] = 1u << sd.block_log;
                    out[4 * n + 3] = 0;
                    ++n;
                }
    }
    return n;
}

int fc_rank(void* h, uint32_t pos, uint32_t sym, int64_t* out) {
    uint32_t v = 0;
    const int st = host_rank(*(FC*)h, pos, sym, &v, nullptr, nullptr);
    *out = v;
    return st;
}

// FmIndex.count of one pattern over the flat layout: the kernel's per-lane code (pattern_start / start_table_lookup / count_step,
// count_lane.h) driven sequentially.  use_table: start from the q-gram start table when the handle has one (like k_count).
static void count_one(FC& h, const uint16_t* pat, int64_t len, bool use_table, CountCounters& cnt, int32_t* result_out, int32_t* st_out,
                      uint32_t* sp_out, uint32_t* ep_out) {
    int32_t result = 0, st = 0;
    uint32_t sp = 0, ep = 0;
    if (len == 0) {
        st = 9;
    } else {
        int64_t i = len - 1;
        uint32_t c = pattern_start(pat, (uint64_t)len, (uint32_t)len, h.ix.char2code, use_table ? h.ix.kmer_q : 0u, h.ix.kmer_stride, h.ix.sigma);
        bool from_table = false;
        if (c & PAT_KMER) {
            if (start_table_lookup(h.ix, c, &sp, &ep)) {
                i -= (int64_t)h.ix.kmer_q - 1;
                from_table = true;
            } else {
                c = h.ix.char2code[pat[i]];
            }
        }
        if (from_table || c != 0) {
            if (!from_table) {
                sp = h.ix.C[c];
                ep = h.ix.C[c + 1];
            }
            bool zero = false;
            while (sp < ep && i >= 1) {
                c = h.ix.char2code[pat[--i]];
                if (c == 0 || c >= h.ix.sigma) {
                    zero = true;
                    sp = ep = 0;
                    break;
                }
                if (h.ix.q4 && ep >= h.ix.length) {
                    st = 9;
                    break;
                }
                uint32_t a = sp, b = ep;
                if (count_step<true>(h.ix, h.CT, c, &a, &b, true, cnt)) {
                    st = 9;
                    break;
                }
                sp = h.ix.C[c] + a;
                ep = h.ix.C[c] + b;
                ep = ep < h.ix.length ? ep : h.ix.length;
            }
            if (!zero && !st) result = ep > sp ? (int32_t)(ep - sp) : 0;
        }
    }
    *result_out = st ? 0 : result;
    *st_out = st;
    *sp_out = sp;
    *ep_out = (!st && result > 0) ? ep : sp;
}

static void count_batch_impl(FC& h, const uint16_t* chars, const uint64_t* pat_off, uint32_t n_pat, int32_t* counts, int32_t* status,
                             uint32_t* ranges, uint64_t* counters, bool use_table) {
    CountCounters cnt{};
    uint64_t n_rank = 0, n_level = 0, n_rec = 0, n_load = 0, kinds[8] = {0};
    for (uint32_t p = 0; p < n_pat; ++p) {
        int32_t result = 0, st = 0;
        uint32_t sp = 0, ep = 0;
        count_one(h, chars + pat_off[p], (int64_t)(pat_off[p + 1] - pat_off[p]), use_table, cnt, &result, &st, &sp, &ep);
        counts[p] = result;
        if (status) status[p] = st;
        if (ranges) {
            ranges[2 * p] = sp;
            ranges[2 * p + 1] = ep;
        }
        n_rank += cnt.ranks;
        n_level += cnt.levels;
        n_rec += cnt.recs;
        n_load += cnt.loads;
        for (int k = 0; k < 8; ++k) kinds[k] += cnt.kinds[k];
        cnt = CountCounters{};
    }
    if (counters) {
        counters[0] += n_rank;
        counters[1] += n_level;
        for (int k = 0; k < 8; ++k) counters[8 + k] += kinds[k];
        counters[6] += n_load;
        counters[7] += n_rec;
    }
}

void fc_count_batch(void* hv, const uint16_t* chars, const uint64_t* pat_off, uint32_t n_pat, int32_t* counts, int32_t* status,
                    uint32_t* ranges, uint64_t* counters) {
    count_batch_impl(*(FC*)hv, chars, pat_off, n_pat, counts, status, ranges, counters, false);
}
// the same starting every pattern from the q-gram start table (fc_build_start_table first)
void fc_count_batch_table(void* hv, const uint16_t* chars, const uint64_t* pat_off, uint32_t n_pat, int32_t* counts, int32_t* status,
                          uint32_t* ranges, uint64_t* counters) {
    count_batch_impl(*(FC*)hv, chars, pat_off, n_pat, counts, status, ranges, counters, true);
}

// Host twin of build_start_table (fmgpu.cu): for every q-gram of codes the chars can spell, the (sp, ep) the step-by-step search
// reports for it as a pattern; q-grams that end in an error are left unusable.  Returns the number of usable entries.
uint64_t fc_build_start_table(void* hv, uint32_t q) {
    FC& h = *(FC*)hv;
    const uint64_t S = h.ix.sigma;
    uint64_t n_entries = 1;
    for (uint32_t k = 0; k < q; ++k) n_entries *= S;
    h.kmer.assign((size_t)n_entries, U32x2{0xffffffffu, 0u});
    h.ix.kmer_q = 0;
    std::vector<uint16_t> pat(q);
    CountCounters cnt{};
    uint64_t usable = 0;
    for (uint64_t idx = 0; idx < n_entries; ++idx) {
        uint64_t v = idx;
        bool ok = true;
        for (uint32_t k = 0; k < q; ++k) {
            const uint32_t code = (uint32_t)(v % S);
            v /= S;
            if (code == 0 || code >= h.F.code2char.size() || h.ix.char2code[h.F.code2char[code]] != code) ok = false;
            else pat[k] = h.F.code2char[code];
        }
        if (!ok) continue;
        int32_t result = 0, st = 0;
        uint32_t sp = 0, ep = 0;
        count_one(h, pat.data(), (int64_t)q, false, cnt, &result, &st, &sp, &ep);
        if (st) continue;
        h.kmer[(size_t)idx] = U32x2{sp, ep};
        ++usable;
    }
    h.ix.kmer = h.kmer.data();
    h.ix.kmer_q = q;
    h.ix.kmer_stride = (uint32_t)S;
    return usable;
}

void fc_extract(void* hv, const int32_t* start, const int32_t* stop, uint32_t n, uint16_t* arena, const uint64_t* arena_off,
                int32_t* len_out, int32_t* status, uint64_t* counters, int32_t offset) {
    FC& h = *(FC*)hv;
    WalkParams P{};
    P.offset = offset;
    P.n_items = n;
    P.start = start;
    P.stop = stop;
    P.arena_off = arena_off;
    P.arena = arena;
    P.len_out = len_out;
    P.status_out = status;
    run_walk<WM_EXTRACT>(h, P, counters);
}

void fc_eub(void* hv, const int32_t* from, uint32_t n, uint16_t boundary, int32_t dst_len, int32_t mode, uint16_t* arena,
            int32_t* len_out, int32_t* status, uint64_t* counters, int32_t offset) {
    FC& h = *(FC*)hv;
    std::vector<uint16_t> left((size_t)n * (size_t)(dst_len > 0 ? dst_len : 1), 0);
    std::vector<int32_t> down(n, 0);
    WalkParams P{};
    P.n_items = n;
    P.from = from;
    P.mb = h.ix.char2code[boundary];
    P.dst_len = dst_len;
    P.eub_mode = mode;
    P.offset = offset;
    P.left = left.data();
    P.down_len = down.data();
    P.arena = arena;
    P.len_out = len_out;
    P.status_out = status;
    run_walk<WM_EUB>(h, P, counters);
    if (mode != 2)
        for (uint32_t w = 0; w < n; ++w)  // what k_eub_assemble does
            for (int32_t q = 0; q < down[w] && q < dst_len - offset; ++q)
                arena[(size_t)w * dst_len + offset + q] = left[(size_t)w * dst_len + (down[w] - 1 - q)];
}

// fused locate -> extractUntilBoundary (kernels_records.cuh) replayed on the host: scan + claim, record numbering, record
// extraction, per-hit arithmetic — the same lane code and the same per-hit formula as the kernels
static void rec_results_host(const FC& h, const int32_t* from, const int32_t* down, const int32_t* win, const int32_t* atb,
                             const std::vector<uint64_t>& uidx, const std::vector<int32_t>& rel_of, uint32_t n, int32_t dst_len, int32_t* rec_index,
                             int32_t* len_out, int32_t* status) {
    for (uint32_t i = 0; i < n; ++i) {
        if (status[i] != 0) {
            rec_index[i] = -1;
            len_out[i] = 0;
            continue;
        }
        const int32_t f = from[i], d = down[i], w = win[i];
        int32_t rel, u = -1;
        if (w < 0) {
            rel = atb[i] ? 0 : REL_NONE;
        } else {
            u = (int32_t)uidx[(size_t)w];
            const int32_t r = rel_of[(size_t)u];
            rel = (r >= 0 && r != REL_NONE) ? r - d : r;
        }
        const EubOut o = eub_right_chunks(f, d, rel, (int32_t)h.ix.length, dst_len, false, 0);
        rec_index[i] = u;
        len_out[i] = o.value;
        status[i] = o.status;
    }
}
uint64_t fc_records(void* hv, const int32_t* from, uint32_t n, uint16_t boundary, int32_t dst_len, int32_t* rec_index, int32_t* len_out,
                    int32_t* status, uint16_t* rec_arena, uint64_t* counters) {
    FC& h = *(FC*)hv;
    uint32_t table = 1024;
    while (table < 2u * n) table <<= 1;
    std::vector<unsigned long long> claims(table, 0ull);
    std::vector<int32_t> down(n, 0), win(n, 0), atb(n, 0);
    for (uint32_t i = 0; i < n; ++i) win[i] = (int32_t)i;  // poisoned: the lane code must initialise every entry
    WalkParams P{};
    P.n_items = n;
    P.from = from;
    P.mb = h.ix.char2code[boundary];
    P.dst_len = dst_len;
    P.eub_mode = EUB_SCAN;
    P.claims = claims.data();
    P.claim_mask = table - 1u;
    P.win_of = win.data();
    P.at_bound = atb.data();
    P.len_out = down.data();
    P.status_out = status;
    run_walk<WM_EUB>(h, P, counters);
    std::vector<uint64_t> uidx(n + 1, 0);
    uint64_t n_rec = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uidx[i] = n_rec;
        if (win[i] == (int32_t)i) ++n_rec;
    }
    std::vector<int32_t> start((size_t)n_rec + 1, 0), rel((size_t)n_rec + 1, 0), dummy((size_t)n_rec + 1, 0);
    for (uint32_t i = 0; i < n; ++i)
        if (win[i] == (int32_t)i) start[(size_t)uidx[i]] = from[i] - down[i];
    WalkParams Q{};
    Q.n_items = (uint32_t)n_rec;
    Q.from = start.data();
    Q.mb = P.mb;
    Q.dst_len = dst_len;
    Q.eub_mode = EUB_RECORD;
    Q.arena = rec_arena;
    Q.len_out = rel.data();
    Q.status_out = dummy.data();
    run_walk<WM_EUB>(h, Q, counters);
    rec_results_host(h, from, down.data(), win.data(), atb.data(), uidx, rel, n, dst_len, rec_index, len_out, status);
    return n_rec;
}

// locate (k_locate): the straight-line lane code of lf_lane.h, one hit at a time
void fc_locate_rows(void* hv, uint32_t* rows_pos, uint32_t n, uint64_t* counters) {
    FC& h = *(FC*)hv;
    const fmgpu_host::RrrTables& RT = fmgpu_host::rrr_tables();
    RrrTab R;
    R.inv = RT.inverse;
    R.cbase = RT.class_base;
    LfCounters cnt{};
    for (uint32_t w = 0; w < n; ++w) {
        uint32_t j = rows_pos[w] + 1u, dist = 0;
        for (;;) {
            const uint32_t pos = j - 1u;
            const SbDesc sd = h.T.sb[pos >> SB_LOG];
            const uint32_t blk = sd.first_block + ((pos & SB_MASK) >> sd.block_log);
            const uint32_t bmask = (1u << sd.block_log) - 1u;
            uint32_t bit = 0, rank = 0;
            ++cnt.sbits;
            sampled_access_rank(h.ix, R, *sg_addr(h.ix, pos), pos, &bit, &rank);
            if (bit) {
                rows_pos[w] = rec_word(h.ix.sa[rank >> 3], rank & 7u) + dist;
                break;
            }
            uint32_t sym = 0, err = 0;
            const uint32_t jn = lf_step(h.ix, h.T, h.ix.blocks[blk], j, bmask, &sym, &err, cnt);
            if (err || dist >= h.ix.length) {  // as k_locate: the reference throws / never returns (walk in a cycle)
                rows_pos[w] = err ? 0xffffffffu : 0xfffffffeu;
                break;
            }
            j = jn;
            ++dist;
        }
    }
    if (counters) {
        counters[0] += cnt.ranks;
        counters[1] += cnt.rank_levels;
        counters[2] += cnt.lf_steps;
        counters[3] += cnt.lf_levels;
        counters[4] += cnt.sbits;
    }
}
// Dense-sample build (kernels_dense.cuh / fmgpu.cu build_dense_samples) replayed on the host with the same lane code: seeds
// from the sampled-row vector, the marking walk, the rank counters, the walk that writes the positions.  marks: n_rec * 8
// words, dsa: (length - 1) / rate + 1 entries.  Returns the number of marked rows, or -1 where the device build gives up.
int64_t fc_dense_build(void* hv, uint32_t rate, uint32_t* marks, uint32_t* dsa) {
    FC& h = *(FC*)hv;
    const DevIndex& ix = h.ix;
    const fmgpu_host::RrrTables& RT = fmgpu_host::rrr_tables();
    RrrTab R;
    R.inv = RT.inverse;
    R.cbase = RT.class_base;
    const uint32_t n_rec = (ix.length + DENSE_ROWS_PER_REC - 1) / DENSE_ROWS_PER_REC;
    const uint32_t n_dense = (ix.length - 1u) / rate + 1u;
    const uint32_t n_seeds = dense_seed_count(ix);
    std::vector<uint32_t> seeds((size_t)n_seeds + 1, 0xffffffffu);
    for (uint32_t row = 0; row < ix.length; ++row) {  // k_dense_seeds
        uint32_t bit = 0, rank = 0;
        sampled_access_rank(ix, R, *sg_addr(ix, row), row, &bit, &rank);
        if (bit && rank < n_seeds) seeds[rank] = row;
    }
    std::memset(marks, 0, (size_t)n_rec * 32);
    bool failed = false;
    for (int pass = 0; pass < 2; ++pass) {
        auto visit = [&](uint32_t row, uint32_t p) {
            if (p % rate != 0u) return;
            const uint32_t q = row / DENSE_ROWS_PER_REC, o = row % DENSE_ROWS_PER_REC;
            if (pass == 0) {
                marks[(size_t)q * 8u + 1u + (o >> 5)] |= 1u << (o & 31u);
            } else {
                uint32_t bit = 0, rank = 0;
                dense_access_rank(*reinterpret_cast<const Rec32*>(marks + (size_t)q * 8u), row, &bit, &rank);
                if (bit && rank < n_dense) dsa[rank] = p;
                else failed = true;
            }
        };
        for (uint64_t k = 0; k <= n_seeds; ++k) {  // k_dense_walk<pass>
            uint32_t row = 0, p = 0;
            const int what = dense_item(ix, seeds.data(), n_seeds, k, &row, &p);
            if (what == 1) continue;
            if (what == 2 || !dense_walk_item(ix, h.T, row, p, visit)) failed = true;
        }
        if (pass == 0) {  // k_dense_popc + scan + k_dense_fill
            uint32_t before = 0;
            for (uint32_t q = 0; q < n_rec; ++q) {
                marks[(size_t)q * 8u] = before;
                for (int k = 1; k < 8; ++k) before += popc32(marks[(size_t)q * 8u + k]);
            }
            if (before != n_dense) return -1;
        }
    }
    return failed ? -1 : (int64_t)n_dense;
}
// k_locate<., DENSE = true>, one hit at a time
void fc_locate_rows_dense(void* hv, const uint32_t* marks, const uint32_t* dsa, uint32_t* rows_pos, uint32_t n, uint64_t* lf_steps) {
    FC& h = *(FC*)hv;
    LfCounters cnt{};
    for (uint32_t w = 0; w < n; ++w) {
        uint32_t j = rows_pos[w] + 1u, dist = 0;
        for (;;) {
            const uint32_t pos = j - 1u;
            const SbDesc sd = h.T.sb[pos >> SB_LOG];
            const uint32_t blk = sd.first_block + ((pos & SB_MASK) >> sd.block_log);
            uint32_t bit = 0, rank = 0;
            dense_access_rank(*reinterpret_cast<const Rec32*>(marks + (size_t)(pos / DENSE_ROWS_PER_REC) * 8u), pos, &bit, &rank);
            if (bit) {
                rows_pos[w] = dsa[rank] + dist;
                break;
            }
            uint32_t sym = 0, err = 0;
            const uint32_t jn = lf_step(h.ix, h.T, h.ix.blocks[blk], j, (1u << sd.block_log) - 1u, &sym, &err, cnt);
            if (err || dist >= h.ix.length) {
                rows_pos[w] = err ? 0xffffffffu : 0xfffffffeu;
                break;
            }
            j = jn;
            ++dist;
        }
    }
    if (lf_steps) *lf_steps += cnt.lf_steps;
}
void fc_sampled(void* hv, uint32_t pos, int32_t* bit, int32_t* rank) {
    FC& h = *(FC*)hv;
    const fmgpu_host::RrrTables& RT = fmgpu_host::rrr_tables();
    RrrTab R;
    R.inv = RT.inverse;
    R.cbase = RT.class_base;
    uint32_t b = 0, r = 0;
    sampled_access_rank(h.ix, R, *sg_addr(h.ix, pos), pos, &b, &r);
    *bit = (int32_t)b;
    *rank = (int32_t)r;
}

// inverseSelect(pos) through the block descriptor / level + node records, as k_wavelet_inverse_select walks them
int fc_inverse_select(void* hv, int64_t p64, int64_t* out) {
    FC& h = *(FC*)hv;
    *out = 0;
    if (p64 < 0 || p64 >= (int64_t)h.ix.length) return 9;
    const uint32_t p = (uint32_t)p64;
    const SbDesc sd = h.T.sb[p >> SB_LOG];
    const uint32_t blk = sd.first_block + ((p & SB_MASK) >> sd.block_log);
    uint32_t r = p & ((1u << sd.block_log) - 1u);
    const Rec32& D = h.ix.blocks[blk];
    uint32_t sym = 0, rk = 0;
    if (D.w[1] & 1u) {
        sym = (D.w[1] >> 8) & 0xffffu;
        rk = D.w[5] + r;
    } else {
        uint32_t sec = D.w[0], nrec = D.w[4], levels = 0;
        for (;;) {
            const Rec32& X = h.ix.sectors[sec + r / SECTOR_BITS];
            const Rec32& N = h.ix.nodes[nrec];
            if (dlevel_descend(X, N, r % SECTOR_BITS, &r, &nrec, &sec, &sym, &rk, &levels)) break;
        }
    }
    *out = p == 0u ? (int64_t)sym : (((int64_t)rk << 32) | (int64_t)sym);
    return 0;
}

// the device-side UTF-8 decoder (utf8_lane.h) on one pattern: returns the char count or -status (*value = code point)
int64_t fc_utf8_convert(const uint8_t* bytes, uint64_t len, uint16_t* out, int32_t* value) {
    uint32_t last = 0;
    return utf8_convert(bytes, 0, len, out, value, &last);
}

// all 32768 (class, offset) pairs through the device unranking, in table order
// the product's own (class, offset) -> block table (flatten.hpp: what the loader decodes with and uploads to the device)
void fc_product_rrr_table(uint16_t* out32768, uint16_t* class_base16, uint8_t* bits16) {
    const fmgpu_host::RrrTables& T = fmgpu_host::rrr_tables();
    std::memcpy(out32768, T.inverse, sizeof T.inverse);
    std::memcpy(class_base16, T.class_base, sizeof T.class_base);
    std::memcpy(bits16, T.bits_needed, sizeof T.bits_needed);
}
void fc_unrank_table(uint16_t* out32768) {
    uint16_t binom[15 * 16];
    fill_binom(binom);
    int cnt[16] = {0};
    for (int v = 0; v < 32768; ++v) cnt[__builtin_popcount(v)]++;
    size_t at = 0;
    for (uint32_t cls = 0; cls < 16; ++cls)
        for (int off = 0; off < cnt[cls]; ++off) out32768[at++] = (uint16_t)rrr_unrank(binom, cls, (uint32_t)off, 15);
}

}  // extern "C"

// ---- design aid (not a test): cost model of a warp-lockstep backward search.  Patterns are taken in groups
// of 32 in the given order; per step the warp needs 1 cell trip + max over its lanes of the level count.
extern "C" void fc_lockstep_sim(void* hv, const uint16_t* chars, const uint64_t* pat_off, uint32_t n_pat, uint64_t* out8) {
    FC& h = *(FC*)hv;
    uint64_t warp_steps = 0, sum_max = 0, lane_steps = 0, sum_levels = 0, diff_steps = 0, hist[16] = {0};
    for (uint32_t g = 0; g < n_pat; g += 32) {
        const uint32_t m = std::min<uint32_t>(32, n_pat - g);
        uint32_t sp[32], ep[32];
        int64_t i[32];
        bool alive[32];
        int64_t maxlen = 0;
        for (uint32_t l = 0; l < m; ++l) {
            const uint64_t a = pat_off[g + l], b = pat_off[g + l + 1];
            i[l] = (int64_t)(b - a) - 1;
            alive[l] = i[l] >= 0;
            if (alive[l]) {
                const uint32_t c = h.ix.char2code[chars[b - 1]];
                if (!c) alive[l] = false;
                else {
                    sp[l] = h.ix.C[c];
                    ep[l] = h.ix.C[c + 1];
                }
            }
            maxlen = std::max<int64_t>(maxlen, i[l]);
        }
        for (int64_t s = 0; s < maxlen; ++s) {
            uint32_t mx = 0;
            bool any = false;
            for (uint32_t l = 0; l < m; ++l) {
                if (!alive[l]) continue;
                if (!(sp[l] < ep[l] && i[l] >= 1)) {
                    alive[l] = false;
                    continue;
                }
                const uint32_t c = h.ix.char2code[chars[pat_off[g + l] + (uint64_t)(--i[l])]];
                if (!c) {
                    alive[l] = false;
                    continue;
                }
                any = true;
                uint64_t r1 = 0, l1 = 0, r2 = 0, l2 = 0;
                uint32_t a = 0, b = 0;
                host_rank(h, sp[l], c, &a, &r1, &l1);
                host_rank(h, ep[l], c, &b, &r2, &l2);
                const bool diff = (sp[l] >> 9) != (ep[l] >> 9) && sp[l] != 0 &&
                                  ((sp[l] >> 16) != (ep[l] >> 16));  // rough: different 64K block
                uint32_t lv = (uint32_t)std::max(l1, l2);
                if (diff) {
                    lv = (uint32_t)(l1 + l2) + 1;  // two sequential walks
                    ++diff_steps;
                }
                mx = std::max(mx, lv);
                sum_levels += l1 + l2;
                hist[std::min<uint64_t>(15, l2)]++;
                ++lane_steps;
                sp[l] = h.ix.C[c] + a;
                ep[l] = h.ix.C[c] + b;
            }
            if (!any) break;
            ++warp_steps;
            sum_max += mx;
        }
    }
    out8[0] = warp_steps;
    out8[1] = sum_max;
    out8[2] = lane_steps;
    out8[3] = sum_levels;
    out8[4] = diff_steps;
    for (int k = 0; k < 16; ++k) out8[8 + k] = hist[k];
}

// ---- design aid (not a test): structure of the lockstep backward search over a length-ordered batch.
// out[0] warp steps, [1] lane steps, [2] sum over warp steps of the deepest lane's record count, [3] single-row lane steps
// (end - start == 1, same block), [4] warp steps whose active lanes are ALL single-row, [5] split lane steps (two blocks),
// [6] lane steps with both positions in one root record, [7] sum of per-lane record counts (track B),
// [16..31] histogram of per-lane records (track B), [32..47] histogram of the per-warp-step maximum,
// [64..127] single-row lane steps by step index, [128..191] lane steps by step index, [192..255] all-single warp steps by step index,
// [256..319] warp steps by step index
extern "C" void fc_warp_sim2(void* hv, const uint16_t* chars, const uint64_t* pat_off, const uint32_t* order, uint32_t n_pat, uint64_t* out) {
    FC& h = *(FC*)hv;
    for (int k = 0; k < 320; ++k) out[k] = 0;
    for (uint32_t g = 0; g < n_pat; g += 32) {
        const uint32_t m = std::min<uint32_t>(32, n_pat - g);
        uint32_t sp[32], ep[32];
        int64_t i[32];
        bool alive[32];
        int64_t maxlen = 0;
        for (uint32_t l = 0; l < m; ++l) {
            const uint32_t p = order[g + l];
            const uint64_t a = pat_off[p], b = pat_off[p + 1];
            i[l] = (int64_t)(b - a) - 1;
            alive[l] = i[l] >= 0;
            if (alive[l]) {
                const uint32_t c = h.ix.char2code[chars[b - 1]];
                if (!c) alive[l] = false;
                else {
                    sp[l] = h.ix.C[c];
                    ep[l] = h.ix.C[c + 1];
                }
            }
            maxlen = std::max<int64_t>(maxlen, i[l]);
        }
        for (int64_t s = 0; s < maxlen; ++s) {
            uint32_t mx = 0, n_act = 0, n_single = 0;
            const int sk = (int)std::min<int64_t>(s, 63);
            for (uint32_t l = 0; l < m; ++l) {
                if (!alive[l]) continue;
                if (!(sp[l] < ep[l] && i[l] >= 1)) {
                    alive[l] = false;
                    continue;
                }
                const uint32_t p = order[g + l];
                const uint32_t c = h.ix.char2code[chars[pat_off[p] + (uint64_t)(--i[l])]];
                if (!c) {
                    alive[l] = false;
                    continue;
                }
                ++n_act;
                const SbDesc da = h.T.sb[sp[l] >> SB_LOG], db = h.T.sb[ep[l] >> SB_LOG];
                const uint32_t blk_a = da.first_block + ((sp[l] & SB_MASK) >> da.block_log);
                const uint32_t blk_b = db.first_block + ((ep[l] & SB_MASK) >> db.block_log);
                const bool same_blk = blk_a == blk_b || sp[l] == 0;
                const bool single = ep[l] - sp[l] == 1 && blk_a == blk_b;
                uint64_t r1 = 0, l1 = 0, r2 = 0, l2 = 0;
                uint32_t a = 0, b = 0;
                host_rank(h, sp[l], c, &a, &r1, &l1);
                host_rank(h, ep[l], c, &b, &r2, &l2);
                const uint32_t pb = (uint32_t)((l2 + 1) / 2), pa = (uint32_t)((l1 + 1) / 2);
                mx = std::max(mx, std::max(pa, pb));
                out[7] += pb;
                out[16 + std::min<uint32_t>(15, pb)]++;
                if (single) {
                    ++n_single;
                    out[3]++;
                    out[64 + sk]++;
                }
                if (!same_blk) out[5]++;
                if (same_blk && sp[l] != 0 &&
                    (sp[l] & ((1u << da.block_log) - 1u)) / SECTOR_BITS == (ep[l] & ((1u << db.block_log) - 1u)) / SECTOR_BITS)
                    out[6]++;
                out[1]++;
                out[128 + sk]++;
                sp[l] = h.ix.C[c] + a;
                ep[l] = h.ix.C[c] + b;
            }
            if (!n_act) break;
            out[0]++;
            out[2] += mx;
            out[32 + std::min<uint32_t>(15, mx)]++;
            out[256 + sk]++;
            if (n_single == n_act) {
                out[4]++;
                out[192 + sk]++;
            }
        }
    }
}

// ---- design aid: for every rank query of the backward search (track B of each lane step), the record count of its walk
// and how often the symbol occurs in its block: out[pairs * 8 + bucket], buckets <=10, <=26, <=42, <=100, <=1000, more;
// out[64 + blockSizeLog] = lane steps by block size
extern "C" void fc_sparse_sim(void* hv, const uint16_t* chars, const uint64_t* pat_off, uint32_t n_pat, uint64_t* out) {
    FC& h = *(FC*)hv;
    for (int k = 0; k < 96; ++k) out[k] = 0;
    for (uint32_t p = 0; p < n_pat; ++p) {
        const uint64_t a0 = pat_off[p], b0 = pat_off[p + 1];
        if (b0 <= a0) continue;
        int64_t i = (int64_t)(b0 - a0) - 1;
        uint32_t c = h.ix.char2code[chars[b0 - 1]];
        if (!c) continue;
        uint32_t sp = h.ix.C[c], ep = h.ix.C[c + 1];
        while (sp < ep && i >= 1) {
            c = h.ix.char2code[chars[a0 + (uint64_t)(--i)]];
            if (!c) break;
            uint64_t r1 = 0, l1 = 0, r2 = 0, l2 = 0;
            uint32_t a = 0, b = 0;
            host_rank(h, sp, c, &a, &r1, &l1);
            host_rank(h, ep, c, &b, &r2, &l2);
            const SbDesc db = h.T.sb[ep >> SB_LOG];
            const uint32_t lo = ep & ~((1u << db.block_log) - 1u);
            uint32_t hi = lo + (1u << db.block_log);
            if (hi > h.ix.length) hi = h.ix.length;
            uint32_t x0 = 0, x1 = 0;
            uint64_t d0 = 0, d1 = 0;
            host_rank(h, lo, c, &x0, &d0, &d1);
            host_rank(h, hi, c, &x1, &d0, &d1);
            const uint32_t occ = x1 - x0;
            const int bucket = occ <= 10 ? 0 : occ <= 26 ? 1 : occ <= 42 ? 2 : occ <= 100 ? 3 : occ <= 1000 ? 4 : 5;
            out[std::min<uint64_t>(7, (l2 + 1) / 2) * 8 + bucket]++;
            out[64 + db.block_log]++;
            sp = h.ix.C[c] + a;
            ep = h.ix.C[c] + b;
        }
    }
}

// Packed transport of the host-pointer count call (host_pack.hpp), host half: the pool narrows n_groups chunks of a char[] into
// bytes exactly as count_host_enqueue submits them (parts_per_group parts per chunk).  wide_out[g] = OR of the chunk's chars.
extern "C" int32_t fc_pack_threads(void) { return (int32_t)fmgpu_host::PackPool::get().threads(); }
extern "C" void fc_pack_narrow(const uint16_t* chars, uint64_t n, uint32_t n_groups, uint32_t parts, uint8_t* bytes_out, uint32_t* wide_out) {
    std::vector<std::atomic<uint32_t>> wide(n_groups);
    for (auto& w : wide) w.store(0);
    std::atomic<uint32_t>* wp = wide.data();
    auto job = fmgpu_host::PackPool::get().submit(n_groups, parts, [=](uint32_t g, uint32_t part) {
        const uint64_t c0 = n * g / n_groups, c1 = n * (g + 1) / n_groups;
        const uint64_t a = c0 + (c1 - c0) * part / parts, b = c0 + (c1 - c0) * (part + 1) / parts;
        const uint32_t m = fmgpu_host::narrow_u16(chars + a, bytes_out + a, (size_t)(b - a));
        wp[g].fetch_or(m);
    });
    for (uint32_t g = 0; g < n_groups; ++g) job->wait_group(g);  // in order, as the uploader does
    job->wait_all();
    for (uint32_t g = 0; g < n_groups; ++g) wide_out[g] = wide[g].load();
}
