// sm_100a kernels of the FM-index query path (hand-written; no library calls on the hot path).
//
// Execution model.  Every memory access of the query path is ONE self-contained record — an 8-byte
// (block, symbol) cell or a 32-byte record = one DRAM sector, fetched with a single 256-bit load (LDG.E.256); a record
// carries everything the operation needs from that address.  Backward search (FmIndex.count, fm/FmIndex.java:455-474) runs
// warp-lockstep: the batch is ordered by pattern length (k_prepass / k_len_scan / k_len_scatter), a
// warp takes 32 patterns of equal length, lane = pattern, and every lane executes count_step
// (count_lane.h) once per pattern char.  Work is taken from a global queue, one atomic per warp batch.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "count_lane.h"
#include "lane_logic.h"
#include "layout.h"
#include "ldrec.h"

namespace fmgpu {

constexpr unsigned FULL = 0xffffffffu;
#ifndef COUNT_THREADS
#define COUNT_THREADS 576  // x 2 CTAs = 36 warps per SM at 54 registers (measured: 0.617 ms vs 0.648 ms at 512 x 2)
#endif
constexpr int CTA_THREADS = COUNT_THREADS;
constexpr int PIPE_CTA_THREADS = COUNT_THREADS >= 320 ? 320 : COUNT_THREADS;  // chunked host call (fmgpu.cu): small CTAs retire early and leave room for the next chunk's pre-pass (192..512 measured within 3 %)
constexpr uint32_t SMEM_C_MAX = 4096;   // entries of C kept in shared memory
constexpr uint32_t SMEM_SB_MAX = 2048;  // superblock descriptors kept in shared memory

// Index tables small enough for shared memory: C array and superblock descriptors.
__device__ __forceinline__ SmemTables stage_tables(const DevIndex& ix, uint32_t* smem) {
    SmemTables t;
    uint32_t used = 0;
    auto stage = [&](const uint32_t* src, uint32_t words, bool fits) -> const uint32_t* {
        if (!fits) return src;
        uint32_t* d = smem + used;
        for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) d[i] = src[i];
        used += (words + 1u) & ~1u;  // keep 8-byte alignment for the 64-bit entries
        return d;
    };
    t.C = stage(ix.C, ix.n_c, ix.n_c <= SMEM_C_MAX);
    t.sb = reinterpret_cast<const SbDesc*>(stage(reinterpret_cast<const uint32_t*>(ix.sb), 2 * ix.n_sb, ix.n_sb <= SMEM_SB_MAX));
    __syncthreads();
    return t;
}
inline size_t tables_smem_bytes(const DevIndex& ix) {
    size_t n = 0;
    if (ix.n_c <= SMEM_C_MAX) n += (ix.n_c + 1u) & ~1u;
    if (ix.n_sb <= SMEM_SB_MAX) n += 2 * (size_t)ix.n_sb;
    return n * 4 + 16;
}
constexpr size_t TABLES_SMEM_MAX_BYTES = (SMEM_C_MAX + 2 * (size_t)SMEM_SB_MAX) * 4 + 16;

// ---------------------------------------------------------------------------------------------
// Pre-pass: one descriptor per pattern (offset, length, alphabet code of the last char =
// monotonicMap.getOrDefault(ch, 0), FmIndex.java:457) and the histogram of pattern lengths.
// A warp runs 32 patterns of (nearly) equal length in lockstep, so the batch is ordered by length
// with a counting sort (lengths >= LEN_BINS-1 share the last bin).
// ---------------------------------------------------------------------------------------------
constexpr uint32_t LEN_BINS = 1024;

__global__ void __launch_bounds__(256) k_prepass(const uint16_t* __restrict__ chars, const uint64_t* __restrict__ pat_off, uint32_t n_pat,
                                                 const uint16_t* __restrict__ char2code, PatDesc* __restrict__ pats,
                                                 uint32_t* __restrict__ bins, uint32_t kmer_q, uint32_t kmer_stride, uint32_t sigma) {
    __shared__ uint32_t h[LEN_BINS];
    for (uint32_t i = threadIdx.x; i < LEN_BINS; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pat; i += gridDim.x * blockDim.x) {
        const uint64_t a = pat_off[i], b = pat_off[i + 1];
        PatDesc d;
        d.off = a;
        d.len = b > a ? (uint32_t)(b - a) : 0u;
        d.last = d.len ? pattern_start(chars, b, d.len, char2code, kmer_q, kmer_stride, sigma) : 0u;
        pats[i] = d;
        atomicAdd(&h[d.len < LEN_BINS - 1 ? d.len : LEN_BINS - 1], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < LEN_BINS; i += blockDim.x)
        if (h[i]) atomicAdd(&bins[i], h[i]);
}
// Packed transport of the host-pointer count call (host_pack.hpp): a chunk arrived as one byte per char and chunk-relative
// uint32 offsets; widen both to what the pre-pass reads (UTF-16 units at the chunk's absolute char offsets, uint64 offsets).
__global__ void __launch_bounds__(256) k_unpack_narrow(const uint8_t* __restrict__ bytes, uint64_t n_chars, uint16_t* __restrict__ chars,
                                                       const uint32_t* __restrict__ off32, uint32_t n_off, uint64_t base,
                                                       uint64_t* __restrict__ off64) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = t; i < n_chars; i += stride) chars[i] = bytes[i];
    for (uint64_t i = t; i < n_off; i += stride) off64[i] = base + off32[i];
}

// exclusive scan of the bins, longest patterns first (they are the long poles of the launch).  256 threads x 4 bins: small
// enough to share an SM with resident k_count CTAs of an earlier chunk (the chunked host call overlaps launches).
constexpr uint32_t SCAN_THREADS = 256;
__global__ void __launch_bounds__(SCAN_THREADS) k_len_scan(uint32_t* __restrict__ bins) {
    static_assert(LEN_BINS == 4 * SCAN_THREADS, "k_len_scan handles four bins per thread");
    __shared__ uint32_t s[SCAN_THREADS];
    const uint32_t t = threadIdx.x;
    uint32_t v[4], sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // reversed order: element e = bins[LEN_BINS - 1 - e]
        v[k] = bins[LEN_BINS - 1 - (4 * t + k)];
        sum += v[k];
    }
    s[t] = sum;
    __syncthreads();
    for (uint32_t o = 1; o < SCAN_THREADS; o <<= 1) {
        const uint32_t x = t >= o ? s[t - o] : 0u;
        __syncthreads();
        s[t] += x;
        __syncthreads();
    }
    uint32_t run = t ? s[t - 1] : 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        bins[LEN_BINS - 1 - (4 * t + k)] = run;
        run += v[k];
    }
}
// Scatter pattern ids into length order.  A block ranks its tile inside shared memory (the value returned by the
// shared-memory atomic is the pattern's rank among the block's patterns of that length) and takes ONE global
// atomic per (block, length) for the base, instead of one per pattern on ~60 hot addresses.
constexpr uint32_t SCATTER_PER_THREAD = 8;
__global__ void __launch_bounds__(256, 8) k_len_scatter(const PatDesc* __restrict__ pats, uint32_t n_pat, uint32_t* __restrict__ bins,
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
// A warp takes 32 patterns of equal length from the length-ordered batch; lane = pattern.  All lanes perform step k of
// their pattern together; what a lane does in a step is count_step (count_lane.h): the two rank queries rank(start, c) and
// rank(end, c) as two tracks of one fused step — the 8-byte cells first, then at most one occurrence record per track.  The
// code is plain SIMT — the hardware reconverges the warp after each step — which costs ~5x fewer issued instructions per rank
// than the lane state machines of v1/v2 (profiles/r01_k_count_v1_ncu_summary.txt, ..._v2_...).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ CountTables stage_count_tables(const DevIndex& ix, uint32_t* smem) { return stage_tables(ix, smem); }
constexpr size_t COUNT_SMEM_MAX_BYTES = TABLES_SMEM_MAX_BYTES;
inline size_t count_smem_bytes(const DevIndex& ix) { return tables_smem_bytes(ix); }

#ifndef COUNT_MIN_CTAS
#define COUNT_MIN_CTAS 2
#endif
#ifndef COUNT_BLOCK_ASSIGN
#define COUNT_BLOCK_ASSIGN 1
#endif
// STATS: count rank queries / levels / records (fmgpu_set_stats; bench.py's roofline accounting and the tests) — off in production
template <bool STATS>
__global__ void __launch_bounds__(CTA_THREADS, COUNT_MIN_CTAS)
k_count(const DevIndex ix, const uint16_t* __restrict__ chars, const PatDesc* __restrict__ pats, const uint32_t* __restrict__ order,
        uint32_t n_pat, int32_t* __restrict__ counts, int32_t* __restrict__ status, uint32_t* __restrict__ ranges, unsigned int* queue,
        unsigned long long* stats) {
    extern __shared__ __align__(8) uint32_t smem[];
    const CountTables T = stage_count_tables(ix, smem);
    const unsigned lane = threadIdx.x & 31u;
    CountCounters cnt;
    cnt.ranks = cnt.levels = cnt.loads = cnt.recs = 0;
    for (int k = 0; k < 8; ++k) cnt.kinds[k] = 0;

    // Work distribution: the first batch of every warp is static — CTA b takes the contiguous batches [b*W, (b+1)*W) of the
    // length-ordered batch, so the warps of a CTA run patterns of (nearly) the same length and the CTA retires as a whole
    // when they are done (short-pattern CTAs early), which lets the CTAs of the NEXT launch (the chunked host call runs
    // several launches on different streams) backfill the SM instead of waiting for one long-pattern warp per CTA.  Further
    // batches come from the global queue.
    const unsigned warps_per_cta = blockDim.x >> 5;
    bool first = COUNT_BLOCK_ASSIGN != 0;
    for (;;) {
        unsigned batch = 0;
        if (first) {
            batch = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
            first = false;
        } else {
            if (lane == 0) batch = atomicAdd(queue, 1u);
            batch = __shfl_sync(FULL, batch, 0) + (COUNT_BLOCK_ASSIGN ? gridDim.x * warps_per_cta : 0u);
        }
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
        const uint16_t* pch = chars + pd.off;
        bool from_table = false;
        if (have && (c & PAT_KMER)) {
            // q-gram start table: the range after the pattern's last q chars, i.e. the state after q - 1 steps (layout.h)
            if (start_table_lookup(ix, c, &sp, &ep)) {
                i -= (int32_t)ix.kmer_q - 1;
                from_table = true;
            } else {
                c = (uint32_t)__ldg(ix.char2code + __ldg(pch + i));  // that q-gram throws on its way: step by step
            }
        }
        if (have && pd.len == 0) {  // pattern[-1]: ArrayIndexOutOfBounds (FmIndex.java:456-457)
            err = 1;
            alive = false;
        } else if (have && !from_table && c == 0) {  // :458
            alive = false;
        } else if (have && !from_table) {
            sp = T.C[c];
            ep = T.C[c + 1];
        }
        // the pattern's chars are mapped to alphabet codes on the fly, fetched two steps ahead (raw char) and one
        // step ahead (its code) so that neither load is on the step's critical path
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

            uint32_t val_s = sp, val_e = ep;
            const uint32_t e1 = count_step<STATS>(ix, T, c, &val_s, &val_e, go, cnt);
            if (go) {
                if (e1) {
                    err = 1;
                    alive = false;
                } else {
                    sp = T.C[c] + val_s;  // :469-470
                    ep = T.C[c] + val_e;
                    ep = ep < ix.length ? ep : ix.length;  // no-op on a consistent index (see count_step)
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

    if (!STATS) return;
    for (int o = 16; o; o >>= 1) {
        cnt.ranks += __shfl_xor_sync(FULL, cnt.ranks, o);
        cnt.levels += __shfl_xor_sync(FULL, cnt.levels, o);
        cnt.loads += __shfl_xor_sync(FULL, cnt.loads, o);
        cnt.recs += __shfl_xor_sync(FULL, cnt.recs, o);
        for (int k = 0; k < 8; ++k) cnt.kinds[k] += __shfl_xor_sync(FULL, cnt.kinds[k], o);
    }
    if (lane == 0 && stats) {
        atomicAdd(stats + 0, (unsigned long long)cnt.ranks);
        atomicAdd(stats + 1, (unsigned long long)cnt.levels);
        atomicAdd(stats + 6, (unsigned long long)cnt.loads);
        atomicAdd(stats + 7, (unsigned long long)cnt.recs);
        for (int k = 0; k < 8; ++k) atomicAdd(stats + 8 + k, (unsigned long long)cnt.kinds[k]);
    }
}

}  // namespace fmgpu
