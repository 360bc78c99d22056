// sm_100a kernels of the FM-index query path (hand-written; no library calls on the hot path).
//
// Execution model.  Every memory access of the query path is ONE 32-byte record =
// one DRAM sector, fetched with a single 256-bit load (LDG.E.256).  A lane runs a small state
// machine: each trip of the warp loop every lane issues at most one such load for whatever state
// it is in (pattern descriptor, (block,symbol) cell, wavelet level sector, path overflow chunk),
// then all lanes post-process their record with a few ALU ops.  Loads of all 32 lanes are in
// flight together regardless of how their states diverge, which is what a dependent-gather
// workload needs: memory-level parallelism, not lock-step control flow.
//
// Backward search (FmIndex.count, fm/FmIndex.java:455-474): a PAIR of adjacent lanes owns one
// pattern — the even lane carries `start`, the odd lane `end`; their two rank queries per step run
// concurrently and meet through warp shuffles.  Finished pairs are refilled from a global queue
// with one __ballot_sync + one atomicAdd per warp.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "lane_logic.h"
#include "layout.h"
#include "ldrec.h"

// experiment knobs (tools/variants.sh): L2 eviction priority of the backward-search loads (ld256 / ld256_keep / ld256_stream)
#ifndef COUNT_LD_CELL
#define COUNT_LD_CELL ld256
#endif
#ifndef COUNT_LD_SECTOR
#define COUNT_LD_SECTOR ld256
#endif

namespace fmgpu {

constexpr unsigned FULL = 0xffffffffu;
#ifndef COUNT_THREADS
#define COUNT_THREADS 512
#endif
constexpr int CTA_THREADS = COUNT_THREADS;
constexpr uint32_t SMEM_C_MAX = 4096;   // entries of C kept in shared memory
constexpr uint32_t SMEM_SB_MAX = 2048;  // superblock descriptors kept in shared memory

// Index tables small enough for shared memory (C array, superblock descriptors).
__device__ __forceinline__ SmemTables stage_tables(const DevIndex& ix, uint32_t* smem) {
    SmemTables t;
    uint32_t used = 0;
    if (ix.n_c <= SMEM_C_MAX) {
        for (uint32_t i = threadIdx.x; i < ix.n_c; i += blockDim.x) smem[i] = ix.C[i];
        t.C = smem;
        used = ix.n_c;
    } else {
        t.C = ix.C;
    }
    if (ix.n_sb <= SMEM_SB_MAX) {
        uint32_t* d = smem + used;
        const uint32_t* s = reinterpret_cast<const uint32_t*>(ix.sb);
        for (uint32_t i = threadIdx.x; i < 2 * ix.n_sb; i += blockDim.x) d[i] = s[i];
        t.sb = reinterpret_cast<const SbDesc*>(d);
    } else {
        t.sb = ix.sb;
    }
    __syncthreads();
    return t;
}
inline size_t tables_smem_bytes(const DevIndex& ix) {
    size_t n = 0;
    if (ix.n_c <= SMEM_C_MAX) n += ix.n_c;
    if (ix.n_sb <= SMEM_SB_MAX) n += 2 * (size_t)ix.n_sb;
    return n * 4 + 16;
}

// ---------------------------------------------------------------------------------------------
// Pre-pass: one descriptor per pattern (offset, length, alphabet code of the last char =
// monotonicMap.getOrDefault(ch, 0), FmIndex.java:457) and the histogram of pattern lengths.
// A warp runs 32 patterns of (nearly) equal length in lockstep, so the batch is ordered by length
// with a counting sort (lengths >= LEN_BINS-1 share the last bin).
// ---------------------------------------------------------------------------------------------
constexpr uint32_t LEN_BINS = 1024;

__global__ void __launch_bounds__(256) k_prepass(const uint16_t* __restrict__ chars, const uint64_t* __restrict__ pat_off, uint32_t n_pat,
                                                 const uint16_t* __restrict__ char2code, PatDesc* __restrict__ pats,
                                                 uint32_t* __restrict__ bins) {
    __shared__ uint32_t h[LEN_BINS];
    for (uint32_t i = threadIdx.x; i < LEN_BINS; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pat; i += gridDim.x * blockDim.x) {
        const uint64_t a = pat_off[i], b = pat_off[i + 1];
        PatDesc d;
        d.off = a;
        d.len = b > a ? (uint32_t)(b - a) : 0u;
        d.last = d.len ? (uint32_t)__ldg(char2code + chars[b - 1]) : 0u;
        pats[i] = d;
        atomicAdd(&h[d.len < LEN_BINS - 1 ? d.len : LEN_BINS - 1], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < LEN_BINS; i += blockDim.x)
        if (h[i]) atomicAdd(&bins[i], h[i]);
}
// exclusive scan of the bins, longest patterns first (they are the long poles of the launch)
__global__ void __launch_bounds__(LEN_BINS) k_len_scan(uint32_t* __restrict__ bins) {
    __shared__ uint32_t s[LEN_BINS];
    const uint32_t t = threadIdx.x;
    s[t] = bins[LEN_BINS - 1 - t];
    __syncthreads();
    for (uint32_t o = 1; o < LEN_BINS; o <<= 1) {
        const uint32_t v = t >= o ? s[t - o] : 0u;
        __syncthreads();
        s[t] += v;
        __syncthreads();
    }
    bins[LEN_BINS - 1 - t] = t ? s[t - 1] : 0u;
}
// Scatter pattern ids into length order.  A block ranks its tile inside shared memory (the value returned by the
// shared-memory atomic is the pattern's rank among the block's patterns of that length) and takes ONE global
// atomic per (block, length) for the base, instead of one per pattern on ~60 hot addresses.
constexpr uint32_t SCATTER_PER_THREAD = 8;
__global__ void __launch_bounds__(256) k_len_scatter(const PatDesc* __restrict__ pats, uint32_t n_pat, uint32_t* __restrict__ bins,
                                                     uint32_t* __restrict__ order) {
    __shared__ uint32_t h[LEN_BINS];
    const uint32_t tile = 256u * SCATTER_PER_THREAD;
    for (uint32_t base = blockIdx.x * tile; base < n_pat; base += gridDim.x * tile) {
        for (uint32_t i = threadIdx.x; i < LEN_BINS; i += blockDim.x) h[i] = 0;
        __syncthreads();
        uint32_t bin[SCATTER_PER_THREAD], rk[SCATTER_PER_THREAD];
#pragma unroll
        for (uint32_t k = 0; k < SCATTER_PER_THREAD; ++k) {
            const uint32_t i = base + k * 256u + threadIdx.x;
            bin[k] = LEN_BINS;
            if (i < n_pat) {
                const uint32_t len = pats[i].len;
                bin[k] = len < LEN_BINS - 1 ? len : LEN_BINS - 1;
                rk[k] = atomicAdd(&h[bin[k]], 1u);
            }
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < LEN_BINS; i += blockDim.x)
            if (h[i]) h[i] = atomicAdd(&bins[i], h[i]);
        __syncthreads();
#pragma unroll
        for (uint32_t k = 0; k < SCATTER_PER_THREAD; ++k) {
            const uint32_t i = base + k * 256u + threadIdx.x;
            if (bin[k] < LEN_BINS) order[h[bin[k]] + rk[k]] = i;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Backward search (FmIndex.count, fm/FmIndex.java:455-474) — warp-lockstep.
//
// A warp takes 32 patterns of equal length from the length-ordered batch; lane = pattern.  All
// lanes perform step k of their pattern together: one (block, symbol) cell fetch, then the level
// loop.  The two rank queries of a step, rank(start, c) and rank(end, c), walk the SAME tree path
// when start and end lie in the same block, so a lane carries both positions down one walk (one
// sector fetch per level when they share a 224-bit sector, two otherwise); when the blocks differ
// the lane runs a second walk.  The code is plain SIMT loops — the hardware reconverges the warp
// after each level loop — which costs ~5x fewer issued instructions per rank than the lane state
// machines of v1/v2 (profiles/r01_k_count_v1_ncu_summary.txt, ..._v2_...: issue-bound at 43-47 %
// lane utilisation); the price is that a step lasts as long as its deepest walk.
// ---------------------------------------------------------------------------------------------
struct WalkOut {
    uint32_t a, b, err;
};

// two wavelet levels (one level record) for two positions of the same even-depth node
// (WaveletFixedBlockBoosting.java:1185-1279); `two` = the code has a second level below this one
__device__ __forceinline__ void dlevel_pair(const DevIndex& ix, uint32_t node, uint32_t t, uint32_t u, bool two, uint32_t& ra, uint32_t& rb,
                                            uint32_t& n_load) {
    const uint32_t qa = ra / SECTOR_BITS, qb = rb / SECTOR_BITS;
    const Rec32 A = COUNT_LD_SECTOR(ix.sectors + (node + qa));
    Rec32 B = A;
    if (qb != qa) B = COUNT_LD_SECTOR(ix.sectors + (node + qb));
    n_load += qb != qa ? 2u : 1u;
    ra = dlevel_rank(A, ra, ra - qa * SECTOR_BITS, t, u, two);
    rb = dlevel_rank(B, rb, rb - qb * SECTOR_BITS, t, u, two);
}

// rank(., c) of two positions of ONE block: ra/rb are block-relative positions (rb == ra for a single query)
__device__ __forceinline__ WalkOut walk_pair(const DevIndex& ix, uint32_t blk, uint32_t c, uint32_t ra, uint32_t rb, bool on,
                                             uint32_t queries, uint32_t& n_level, uint32_t& n_load, uint32_t& n_rec) {
    WalkOut o;
    o.a = o.b = o.err = 0;
    if (!on) return o;
    const Rec32 cell = COUNT_LD_CELL(ix.cells + ((uint64_t)blk * ix.sigma + c));
    ++n_load;
    const uint32_t kind = (cell.w[2] >> 8) & 0xffu;
    const uint32_t base = cell.w[0];
    if (kind != CELL_NORMAL) {
        const uint32_t run = kind == CELL_RUN ? 0xffffffffu : 0u;  // :1141-1146
        o.a = base + (ra & run);
        o.b = base + (rb & run);
        o.err = kind == CELL_THROW;
        return o;
    }
    const uint32_t code = cell.w[1];
    const uint32_t L = cell.w[2] & 0xffu;
    const uint32_t pairs = (L + 1u) >> 1;
    const uint32_t inl = pairs > CELL_INLINE_PAIRS ? CELL_INLINE_PAIRS - 1u : pairs;
    const uint32_t* more = reinterpret_cast<const uint32_t*>(ix.ovf + cell.w[7]);  // codes longer than 10 bits: rest of the path
#pragma unroll
    for (uint32_t k = 0; k < CELL_INLINE_PAIRS; ++k)
        if (k < inl) {
            const uint32_t d = 2u * k;
            const bool two = d + 1u < L;
            dlevel_pair(ix, cell.w[3 + k], (code >> (L - 1u - d)) & 1u, two ? (code >> (L - 2u - d)) & 1u : 0u, two, ra, rb, n_load);
        }
#pragma unroll 1
    for (uint32_t k = inl; k < pairs; ++k) {
        const uint32_t d = 2u * k;
        const bool two = d + 1u < L;
        dlevel_pair(ix, __ldg(more + (k - inl)), (code >> (L - 1u - d)) & 1u, two ? (code >> (L - 2u - d)) & 1u : 0u, two, ra, rb, n_load);
    }
    n_level += L * queries;
    n_rec += pairs * queries;
    o.a = base + ra;
    o.b = base + rb;
    return o;
}

#ifndef COUNT_MIN_CTAS
#define COUNT_MIN_CTAS 2
#endif
__global__ void __launch_bounds__(CTA_THREADS, COUNT_MIN_CTAS)
k_count(const DevIndex ix, const uint16_t* __restrict__ chars, const PatDesc* __restrict__ pats, const uint32_t* __restrict__ order,
        uint32_t n_pat, int32_t* __restrict__ counts, int32_t* __restrict__ status, uint32_t* __restrict__ ranges, unsigned int* queue,
        unsigned long long* stats) {
    extern __shared__ uint32_t smem[];
    const SmemTables T = stage_tables(ix, smem);
    const unsigned lane = threadIdx.x & 31u;
    uint32_t n_rank = 0, n_level = 0, n_load = 0, n_rec = 0;

    for (;;) {
        unsigned batch = 0;
        if (lane == 0) batch = atomicAdd(queue, 1u);
        batch = __shfl_sync(FULL, batch, 0);
        if ((uint64_t)batch * 32u >= n_pat) break;
        const uint32_t slot = batch * 32u + lane;
        const bool have = slot < n_pat;
        const uint32_t pat = have ? order[slot] : 0u;
        PatDesc pd;
        pd.off = 0;
        pd.len = 0;
        pd.last = 0;
        if (have) pd = pats[pat];

        uint32_t c = pd.last, sp = 0, ep = 0, err = 0;
        int32_t i = (int32_t)pd.len - 1;
        bool alive = have;
        if (have && pd.len == 0) {  // pattern[-1]: ArrayIndexOutOfBounds (FmIndex.java:456-457)
            err = 1;
            alive = false;
        } else if (have && c == 0) {  // :458
            alive = false;
        } else if (have) {
            sp = T.C[c];
            ep = T.C[c + 1];
        }
        // the pattern's chars are mapped to alphabet codes on the fly, fetched two steps ahead (raw char) and one
        // step ahead (its code) so that neither load is on the step's critical path
        const uint16_t* pch = chars + pd.off;
        uint32_t cnext = (alive && i >= 1) ? (uint32_t)__ldg(ix.char2code + __ldg(pch + (i - 1))) : 0u;
        uint32_t raw2 = (alive && i >= 2) ? (uint32_t)__ldg(pch + (i - 2)) : 0u;

        for (;;) {
            // :464  while (start < end && i >= offset + 1)
            bool go = alive && sp < ep && i >= 1;
            if (go) {
                --i;
                c = cnext;
                if (c == 0 || c >= ix.sigma) {  // :466-468 unknown symbol => 0 ; rank of a symbol >= sigma is 0 => empty range
                    sp = ep = 0;
                    alive = false;
                    go = false;
                } else if (ix.q4 && ep >= ix.length) {  // rank(size, .) on a superblock boundary throws (:1022-1026)
                    err = 1;
                    alive = false;
                    go = false;
                }
            } else {
                alive = false;
            }
            if (!__any_sync(FULL, go)) break;
            if (go && i >= 1) cnext = (uint32_t)__ldg(ix.char2code + raw2);
            if (go && i >= 2) raw2 = (uint32_t)__ldg(pch + (i - 2));

            // block of each position; start == 0 needs no query (rank(0, c) == 0, :1012)
            const SbDesc se = T.sb[ep >> SB_LOG];
            const uint32_t blk_e = se.first_block + ((ep & SB_MASK) >> se.block_log);
            const uint32_t re = ep & ((1u << se.block_log) - 1u);
            const SbDesc ss = T.sb[sp >> SB_LOG];
            const uint32_t blk_s = ss.first_block + ((sp & SB_MASK) >> ss.block_log);
            const uint32_t rs = sp & ((1u << ss.block_log) - 1u);
            const bool with_s = sp != 0u && blk_s == blk_e;
            const bool second = go && sp != 0u && blk_s != blk_e;
            n_rank += go ? (sp != 0u ? 2u : 1u) : 0u;

            const WalkOut w1 = walk_pair(ix, blk_e, c, with_s ? rs : re, re, go, with_s ? 2u : 1u, n_level, n_load, n_rec);
            uint32_t val_s = with_s ? w1.a : 0u;
            const uint32_t val_e = w1.b;
            uint32_t e2 = 0;
            if (__any_sync(FULL, second)) {
                const WalkOut w2 = walk_pair(ix, blk_s, c, rs, rs, second, 1u, n_level, n_load, n_rec);
                if (second) val_s = w2.a;
                e2 = w2.err;
            }
            if (go) {
                if (w1.err | e2) {
                    err = 1;
                    alive = false;
                } else {
                    sp = T.C[c] + val_s;  // :469-470
                    ep = T.C[c] + val_e;
                }
            }
        }
        if (have) {
            const int32_t result = (!err && ep > sp) ? (int32_t)(ep - sp) : 0;  // :473
            counts[pat] = result;
            if (status) status[pat] = err ? 9 : 0;
            if (ranges) {
                ranges[2 * (uint64_t)pat] = sp;
                ranges[2 * (uint64_t)pat + 1] = result > 0 ? ep : sp;
            }
        }
    }

    for (int o = 16; o; o >>= 1) {
        n_rank += __shfl_xor_sync(FULL, n_rank, o);
        n_level += __shfl_xor_sync(FULL, n_level, o);
        n_load += __shfl_xor_sync(FULL, n_load, o);
        n_rec += __shfl_xor_sync(FULL, n_rec, o);
    }
    if (lane == 0 && stats) {
        atomicAdd(stats + 0, (unsigned long long)n_rank);
        atomicAdd(stats + 1, (unsigned long long)n_level);
        atomicAdd(stats + 6, (unsigned long long)n_load);
        atomicAdd(stats + 7, (unsigned long long)n_rec);
    }
}

}  // namespace fmgpu
