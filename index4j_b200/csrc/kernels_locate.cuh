// k_locate — LF walks of FmIndex.locate (fm/FmIndex.java:526-548), warp-lockstep.
//
// One lane = one hit (an SA row).  Every trip of the warp loop each live lane performs ONE iteration of the
// reference's while loop: test sampledSuffixes.access(j-1); if the row is sampled fetch its SA sample and write
// `suffixes[rankOnes(j)-1] + distance`, else take one LF step.  The group record of the sampled-row test and the
// block descriptor of the LF step are fetched together (the LF step follows in 31 of 32 trips at sampleRate 32).
// The code is plain SIMT (lf_lane.h): lanes diverge inside the level loop / the offset decode and the hardware
// reconverges them, which needs ~6x fewer issued instructions per LF step than the phase machine of k_walk
// (profiles/r01_k_walk_v1_ncu_summary.txt: 6.2 of 32 lanes active, issue-bound).  Lanes whose hit is finished
// are refilled at the top of the next trip from a warp-local chunk of the global work queue.
//
// RRR decode of the sampled-row vector uses the (class, offset) -> block table in SHARED memory (64 KB).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.cuh"
#include "lf_lane.h"

namespace fmgpu {

#ifndef LOCATE_THREADS
#define LOCATE_THREADS 640
#endif
#ifndef LOCATE_MIN_CTAS
#define LOCATE_MIN_CTAS 2
#endif
// CTA shape of the dense-sample instantiations (no 64 KB table in shared memory: the register file is the only limit)
#ifndef LOCATE_DENSE_THREADS
#define LOCATE_DENSE_THREADS 640
#endif
#ifndef LOCATE_DENSE_MIN_CTAS
#define LOCATE_DENSE_MIN_CTAS 2
#endif
constexpr int locate_threads(bool dense) { return dense ? LOCATE_DENSE_THREADS : LOCATE_THREADS; }
constexpr uint32_t LOCATE_TAB_WORDS = (32768u * 2u + 16u * 2u) / 4u;  // inverse table + class bases

inline size_t locate_smem_bytes(const DevIndex& ix, bool dense = false) { return (dense ? 0 : LOCATE_TAB_WORDS * 4) + tables_smem_bytes(ix); }

// STATS: keep the work counters (fmgpu_set_stats); the production instantiation carries none.
// DENSE: the sampled-row test reads the device-side dense marks (layout.h: dmarks / dsa, kernels_dense.cuh) instead of the RRR
// vector — one plain 224-row record, no offset stream, no (class, offset) table in shared memory — and the walk ends at the
// nearest multiple of dense_rate.
template <bool STATS, bool DENSE>
__global__ void __launch_bounds__(DENSE ? LOCATE_DENSE_THREADS : LOCATE_THREADS, DENSE ? LOCATE_DENSE_MIN_CTAS : LOCATE_MIN_CTAS)
k_locate(const DevIndex ix, uint32_t* __restrict__ rows_pos, uint32_t n_items, uint32_t chunk, unsigned int* queue,
         unsigned long long* stats, const uint64_t* __restrict__ hit_off, uint32_t n_pat, int32_t* __restrict__ status, uint32_t row_base) {
    extern __shared__ uint32_t smem[];
    uint16_t* inv = reinterpret_cast<uint16_t*>(smem);
    uint16_t* cbase = inv + 32768;
    if (!DENSE) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(ix.rrr_inv);
        for (uint32_t i = threadIdx.x; i < 16384u; i += blockDim.x) smem[i] = __ldg(src + i);
        if (threadIdx.x < 16) cbase[threadIdx.x] = __ldg(ix.rrr_cbase + threadIdx.x);
    }
    const SmemTables T = stage_tables(ix, smem + (DENSE ? 0u : LOCATE_TAB_WORDS));  // ends with __syncthreads()
    RrrTab R;
    R.inv = inv;
    R.cbase = cbase;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;

    LfCounters cnt;
    cnt.lf_steps = cnt.lf_levels = cnt.ranks = cnt.rank_levels = cnt.sbits = cnt.recs = 0;

    uint32_t j = 0, dist = 0, w = 0;
    bool active = false;
    uint32_t next = 0, end = 0;  // warp-uniform: the warp's chunk of the work queue
    bool exhausted = false;

    for (;;) {
        const unsigned idle = __ballot_sync(FULL, !active);
        if (idle && !exhausted) {
            if (next == end) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(queue, chunk);
                base = __shfl_sync(FULL, base, 0);
                if (base >= n_items) {
                    exhausted = true;
                } else {
                    next = base;
                    end = base + chunk < n_items ? base + chunk : n_items;
                }
            }
            const uint32_t avail = end - next;
            const uint32_t mine = __popc(idle & lt_mask);
            if (!active && mine < avail) {
                w = next + mine;
                j = rows_pos[w] + 1u;  // :527-529
                dist = 0;
                active = true;
            }
            const uint32_t need = __popc(idle);
            next += need < avail ? need : avail;
        }
        if (!__any_sync(FULL, active)) {
            if (exhausted) break;
            continue;
        }
        if (active) {
            const uint32_t pos = j - 1u;
            const SbDesc sd = T.sb[pos >> SB_LOG];
            const uint32_t blk = sd.first_block + ((pos & SB_MASK) >> sd.block_log);
            const uint32_t bmask = (1u << sd.block_log) - 1u;
            const Rec32 G = ld256(DENSE ? ix.dmarks + pos / DENSE_ROWS_PER_REC : sg_addr(ix, pos));
            const Rec32 D = ld256(ix.blocks + blk);
            uint32_t bit = 0, rank = 0;
            ++cnt.sbits;
            if (DENSE) dense_access_rank(G, pos, &bit, &rank);
            else sampled_access_rank(ix, R, G, pos, &bit, &rank);  // :531
            Rec32 SA;
            uint32_t dense_sa = 0;
            // issued before the other lanes' LF step, consumed after it
            if (bit && DENSE) dense_sa = __ldg(ix.dsa + (rank < ix.n_dense ? rank : 0u));
            if (bit && !DENSE) SA = ld256(ix.sa + (rank >> 3));
            if (!bit) {
                uint32_t sym = 0, err = 0;
                const uint32_t jn = lf_step(ix, T, D, j, bmask, &sym, &err, cnt);  // :532-536
                // The reference throws out of locate() when an LF step indexes outside its arrays (status 9), and never returns
                // when a walk runs into a cycle — possible only where inverseSelect truncates a run-block symbol (quirk Q1): a
                // walk of more than `length` steps has visited a row twice (status 12).  The hit's pattern gets the status.
                if (!err && dist >= ix.length) err = 2;
                if (err) {
                    rows_pos[w] = 0xffffffffu;
                    active = false;
                    if (status) {
                        uint32_t lo = 0, hi = n_pat;  // last pattern with hit_off[p] <= the hit's row in the whole call
                        while (hi - lo > 1u) {
                            const uint32_t mid = (lo + hi) >> 1;
                            if (hit_off[mid] <= (uint64_t)w + row_base) lo = mid;
                            else hi = mid;
                        }
                        atomicMax(status + lo, err == 2 ? 12 : 9);
                    }
                } else {
                    j = jn;
                    ++dist;
                }
            }
            if (bit) {
                rows_pos[w] = (DENSE ? dense_sa : rec_word(SA, rank & 7u)) + dist;  // suffixes[rankOnes(j) - 1] + distance, rankOnes(j) = rankOnes(j-1) + 1 (:538-542)
                active = false;
            }
        }
    }

    if (!STATS) return;
    for (int o = 16; o; o >>= 1) {
        cnt.ranks += __shfl_xor_sync(FULL, cnt.ranks, o);
        cnt.rank_levels += __shfl_xor_sync(FULL, cnt.rank_levels, o);
        cnt.lf_steps += __shfl_xor_sync(FULL, cnt.lf_steps, o);
        cnt.lf_levels += __shfl_xor_sync(FULL, cnt.lf_levels, o);
        cnt.recs += __shfl_xor_sync(FULL, cnt.recs, o);
        cnt.sbits += __shfl_xor_sync(FULL, cnt.sbits, o);
    }
    if (lane == 0 && stats) {
        atomicAdd(stats + 0, (unsigned long long)cnt.ranks);
        atomicAdd(stats + 1, (unsigned long long)cnt.rank_levels);
        atomicAdd(stats + 2, (unsigned long long)cnt.lf_steps);
        atomicAdd(stats + 3, (unsigned long long)cnt.lf_levels);
        atomicAdd(stats + 7, (unsigned long long)cnt.recs);
        atomicAdd(stats + 4, (unsigned long long)cnt.sbits);
    }
}

// k_extract<MODE> — LF walks of FmIndex.extract (fm/FmIndex.java:564-608) and extractUntilBoundary{,Left,Right}
// (:640-922), warp-lockstep: one lane = one extraction, every trip each live lane takes one LF step (preceded by
// the ISA-sample fetch when a walk starts) and hands the char to ExLane::on_char (lf_lane.h).
// 576 threads x 2 CTAs = 36 warps/SM at 56 registers (measured: extractUntilBoundary 5.09 ms per 1 M records vs 5.35 ms at
// 256 x 4 / 63 registers, 5.33 ms at 640 x 2 / 48 registers with spills)
#ifndef EXTRACT_THREADS
#define EXTRACT_THREADS 576
#endif
#ifndef EXTRACT_MIN_CTAS
#define EXTRACT_MIN_CTAS 2
#endif
template <int MODE, bool STATS>
__global__ void __launch_bounds__(EXTRACT_THREADS, EXTRACT_MIN_CTAS)
k_extract(const DevIndex ix, WalkParams P, uint32_t chunk, unsigned int* queue, unsigned long long* stats) {
    extern __shared__ uint32_t smem[];
    const SmemTables T = stage_tables(ix, smem);  // ends with __syncthreads()
    if (MODE == WM_EUB) P.mb = (uint32_t)__ldg(ix.char2code + (P.mb & 0xffffu));
    const unsigned lane_id = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane_id) - 1u;

    LfCounters cnt;
    cnt.lf_steps = cnt.lf_levels = cnt.ranks = cnt.rank_levels = cnt.sbits = cnt.recs = 0;
    ExLane<MODE> lane;
    lane.init();
    uint32_t next = 0, end = 0;
    bool exhausted = false;

    for (;;) {
        const unsigned idle = __ballot_sync(FULL, !lane.active);
        if (idle && !exhausted) {
            if (next == end) {
                unsigned base = 0;
                if (lane_id == 0) base = atomicAdd(queue, chunk);
                base = __shfl_sync(FULL, base, 0);
                if (base >= P.n_items) {
                    exhausted = true;
                } else {
                    next = base;
                    end = base + chunk < P.n_items ? base + chunk : P.n_items;
                }
            }
            const uint32_t avail = end - next;
            const uint32_t mine = __popc(idle & lt_mask);
            if (!lane.active && mine < avail) {
                const uint32_t item = next + mine;
                const ItemRaw raw = walk_load_item<MODE>(P, item);
                lane.begin(ix, P, item, raw);  // may finish at once (error statuses, empty ranges)
            }
            const uint32_t need = __popc(idle);
            next += need < avail ? need : avail;
        }
        if (!__any_sync(FULL, lane.active)) {
            if (exhausted) break;
            continue;
        }
        if (lane.active) lane.trip(ix, T, P, cnt);
    }

    if (!STATS) return;
    for (int o = 16; o; o >>= 1) {
        cnt.ranks += __shfl_xor_sync(FULL, cnt.ranks, o);
        cnt.rank_levels += __shfl_xor_sync(FULL, cnt.rank_levels, o);
        cnt.lf_steps += __shfl_xor_sync(FULL, cnt.lf_steps, o);
        cnt.lf_levels += __shfl_xor_sync(FULL, cnt.lf_levels, o);
        cnt.recs += __shfl_xor_sync(FULL, cnt.recs, o);
    }
    if (lane_id == 0 && stats) {
        atomicAdd(stats + 0, (unsigned long long)cnt.ranks);
        atomicAdd(stats + 1, (unsigned long long)cnt.rank_levels);
        atomicAdd(stats + 2, (unsigned long long)cnt.lf_steps);
        atomicAdd(stats + 3, (unsigned long long)cnt.lf_levels);
        atomicAdd(stats + 7, (unsigned long long)cnt.recs);
    }
}

}  // namespace fmgpu
