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

namespace fmgpu {

constexpr unsigned FULL = 0xffffffffu;
constexpr int CTA_THREADS = 256;
constexpr uint32_t SMEM_C_MAX = 4096;   // entries of C kept in shared memory
constexpr uint32_t SMEM_SB_MAX = 2048;  // superblock descriptors kept in shared memory

__device__ __forceinline__ Rec32 ld256(const Rec32* p) {
    Rec32 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p));
    return r;
}

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
// Pre-pass: UTF-16 units -> alphabet codes (monotonicMap.getOrDefault(ch, 0), FmIndex.java:457,465)
// and one descriptor per pattern.  Fully coalesced; negligible next to the search.
// ---------------------------------------------------------------------------------------------
__global__ void k_prepass(const uint16_t* __restrict__ chars, const uint64_t* __restrict__ pat_off, uint32_t n_pat, uint64_t total_chars,
                          const uint16_t* __restrict__ char2code, uint16_t* __restrict__ codes, PatDesc* __restrict__ pats) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t t0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t i = t0; i < total_chars; i += stride) codes[i] = __ldg(char2code + chars[i]);
    for (uint64_t i = t0; i < n_pat; i += stride) {
        const uint64_t a = pat_off[i], b = pat_off[i + 1];
        PatDesc d;
        d.off = a;
        d.len = b > a ? (uint32_t)(b - a) : 0u;
        d.last = d.len ? (uint32_t)__ldg(char2code + chars[b - 1]) : 0u;
        pats[i] = d;
    }
}

// ---------------------------------------------------------------------------------------------
// Backward search.
// ---------------------------------------------------------------------------------------------
enum CountPhase : uint32_t { CP_IDLE = 0, CP_FETCH, CP_CELL, CP_LEVEL, CP_OVF, CP_WAIT, CP_EXIT };

__global__ void __launch_bounds__(CTA_THREADS)
k_count(const DevIndex ix, const uint16_t* __restrict__ codes, const PatDesc* __restrict__ pats, uint32_t n_pat,
        int32_t* __restrict__ counts, int32_t* __restrict__ status, uint32_t* __restrict__ ranges, unsigned int* queue,
        unsigned long long* stats) {
    extern __shared__ uint32_t smem[];
    const SmemTables T = stage_tables(ix, smem);
    const unsigned lane = threadIdx.x & 31u;
    const bool odd = lane & 1u;
    const unsigned pair_shift = lane & ~1u;

    uint32_t phase = CP_IDLE;
    uint32_t pat = 0, sp = 0, ep = 0, c = 0, cnext = 0, val = 0, err = 0;
    int32_t i = 0;
    uint64_t off = 0;
    const Rec32* addr = nullptr;
    RankSt rs;
    rs.p5 = rs.p6 = rs.p7 = 0;
    uint32_t n_rank = 0, n_level = 0;

    for (;;) {
        // refill idle pairs from the global queue (one atomic per warp)
        const unsigned idle = __ballot_sync(FULL, phase == CP_IDLE && !odd);
        if (idle) {
            const int leader = __ffs(idle) - 1;
            unsigned base = 0;
            if ((int)lane == leader) base = atomicAdd(queue, (unsigned)__popc(idle));
            base = __shfl_sync(FULL, base, leader);
            const unsigned mine = base + __popc(idle & ((1u << lane) - 1u));
            const unsigned got = __shfl_sync(FULL, mine, pair_shift);
            if (phase == CP_IDLE) {
                pat = got;
                if (pat < n_pat) {
                    phase = CP_FETCH;
                    addr = reinterpret_cast<const Rec32*>(pats) + (pat >> 1);
                } else {
                    phase = CP_EXIT;
                }
            }
        }
        if (!__any_sync(FULL, phase != CP_EXIT)) break;

        Rec32 A;
        if (phase >= CP_FETCH && phase <= CP_OVF) A = ld256(addr);

        bool step = false;  // (sp, ep, i) hold a fresh range: decide how the pattern goes on
        if (phase == CP_FETCH) {
            const bool hi = pat & 1u;
            off = (uint64_t)(hi ? A.w[4] : A.w[0]) | ((uint64_t)(hi ? A.w[5] : A.w[1]) << 32);
            const uint32_t len = hi ? A.w[6] : A.w[2];
            c = hi ? A.w[7] : A.w[3];
            i = (int32_t)len - 1;
            err = 0;
            if (len == 0) {  // pattern[-1]: ArrayIndexOutOfBounds (FmIndex.java:456-457)
                err = 1;
                sp = ep = 0;
            } else if (c == 0) {  // :458
                sp = ep = 0;
                i = 0;
            } else {
                sp = T.C[c];
                ep = T.C[c + 1];
                cnext = i >= 1 ? (uint32_t)__ldg(codes + off + (uint64_t)(i - 1)) : 0u;
            }
            step = true;
        } else if (phase == CP_CELL) {
            ++n_rank;
            const uint32_t o = rank_on_cell(ix, A, rs, &addr, &val);
            if (o == RK_MORE) phase = CP_LEVEL;
            else {
                phase = CP_WAIT;
                err |= (o == RK_THROW);
            }
        } else if (phase == CP_LEVEL) {
            ++n_level;
            bool want_ovf = false;
            const uint32_t o = rank_on_level(ix, A, rs, &addr, &val, &want_ovf);
            if (o == RK_DONE) phase = CP_WAIT;
            else if (want_ovf) phase = CP_OVF;
        } else if (phase == CP_OVF) {
            rank_on_ovf(ix, A, rs, &addr);
            phase = CP_LEVEL;
        }

        // the two ranks of a step meet here
        const unsigned waitm = __ballot_sync(FULL, phase == CP_WAIT);
        const unsigned errm = __ballot_sync(FULL, err != 0);
        const uint32_t mine = T.C[c] + val;  // :469-470
        const uint32_t theirs = __shfl_xor_sync(FULL, mine, 1);
        if (((waitm >> pair_shift) & 3u) == 3u) {
            sp = odd ? theirs : mine;
            ep = odd ? mine : theirs;
            err = (errm >> pair_shift) & 3u;
            step = true;
        }
        if (step) {
            bool finished = true;
            int32_t result = 0;
            if (!err) {
                if (sp < ep && i >= 1) {  // :464
                    --i;
                    c = cnext;
                    if (c != 0) {  // :466-468: unknown symbol => 0
                        finished = false;
                        cnext = i >= 1 ? (uint32_t)__ldg(codes + off + (uint64_t)(i - 1)) : 0u;
                        const uint32_t o = rank_begin(ix, T, odd ? ep : sp, c, rs, &addr, &val);
                        if (o == RK_MORE) phase = CP_CELL;
                        else {
                            phase = CP_WAIT;
                            err = (o == RK_THROW);
                        }
                    }
                } else {
                    result = ep > sp ? (int32_t)(ep - sp) : 0;  // :473
                }
            }
            if (finished) {
                if (!odd) {
                    counts[pat] = err ? 0 : result;
                    if (status) status[pat] = err ? 9 : 0;
                    if (ranges) {
                        ranges[2 * (uint64_t)pat] = sp;
                        ranges[2 * (uint64_t)pat + 1] = (!err && result > 0) ? ep : sp;
                    }
                }
                phase = CP_IDLE;
            }
        }
    }

    // work counters (roofline accounting): warp-reduce, one atomic pair per warp
    for (int o = 16; o; o >>= 1) {
        n_rank += __shfl_xor_sync(FULL, n_rank, o);
        n_level += __shfl_xor_sync(FULL, n_level, o);
    }
    if (lane == 0 && stats) {
        atomicAdd(stats + 0, (unsigned long long)n_rank);
        atomicAdd(stats + 1, (unsigned long long)n_level);
    }
}

}  // namespace fmgpu
