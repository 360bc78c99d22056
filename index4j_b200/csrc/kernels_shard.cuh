// Sharded index (texts beyond the reference's 2^31-char limit, fm/FmIndex.java:131,155,335-341; BASELINE.json configs[4]): one
// FmIndex per GPU over a text shard with max_pattern_len - 1 chars of overlap.  The per-shard locate results are composed
// here: a hit is OWNED by the shard in which it starts before owned_len; at most max_hits hits per pattern survive globally,
// lowest shard first, SA order inside a shard.  Three kernels around the two NCCL exchanges (index4j_b200/sharded.py):
//   k_shard_keep   per pattern: how many of its local hits this shard owns (capped at max_hits)
//   k_shard_plan   from every rank's kept counts: what each rank contributes after the global cut, global hits per pattern
//   k_shard_pack   this rank's contribution, compacted in pattern order, as global (int64) positions   [before the exchange]
//   k_shard_merge  every rank's contribution into the final per-pattern order                            [after the exchange]
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fmgpu {

// one warp per pattern: owned hits among its local hits [hit_off[p], hit_off[p+1])
__global__ void k_shard_keep(const int32_t* __restrict__ pos, const uint64_t* __restrict__ hit_off, uint32_t n_pat, int32_t owned_len,
                             int32_t max_hits, int32_t* __restrict__ kept) {
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5, lane = threadIdx.x & 31u;
    for (uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < n_pat; p += warps) {
        const uint64_t a = hit_off[p], b = hit_off[p + 1];
        uint32_t n = 0;
        for (uint64_t i = a + lane; i < b; i += 32) n += pos[i] >= 0 && pos[i] < owned_len ? 1u : 0u;
        for (int o = 16; o; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
        if (lane == 0) kept[p] = (max_hits > 0 && n > (uint32_t)max_hits) ? max_hits : (int32_t)n;
    }
}

// all_kept[r * n_pat + p]: kept count of rank r.  take[r * n_pat + p] = what rank r contributes after the global cut (rank
// order), n_hits[p] = global hits of the pattern.
__global__ void k_shard_plan(const int32_t* __restrict__ all_kept, uint32_t n_pat, uint32_t world, int32_t max_hits, int32_t* __restrict__ take,
                             int32_t* __restrict__ n_hits) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pat) return;
    int64_t before = 0;
    for (uint32_t r = 0; r < world; ++r) {
        int64_t k = all_kept[(uint64_t)r * n_pat + p];
        if (max_hits > 0) {
            const int64_t room = (int64_t)max_hits - before;
            k = k < room ? k : (room > 0 ? room : 0);
        }
        take[(uint64_t)r * n_pat + p] = (int32_t)k;
        before += k;
    }
    n_hits[p] = before > 0x7fffffffLL ? 0x7fffffff : (int32_t)before;
}

// one warp per pattern: the first mine[p] owned hits, in SA order, + text_start -> send[send_off[p] ..]
__global__ void k_shard_pack(const int32_t* __restrict__ pos, const uint64_t* __restrict__ hit_off, uint32_t n_pat, int32_t owned_len,
                             int64_t text_start, const int32_t* __restrict__ mine, const uint64_t* __restrict__ send_off, int64_t* __restrict__ send) {
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5, lane = threadIdx.x & 31u;
    for (uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < n_pat; p += warps) {
        const uint64_t a = hit_off[p], b = hit_off[p + 1], o = send_off[p];
        const uint32_t want = (uint32_t)mine[p];
        uint32_t done = 0;
        for (uint64_t i = a; i < b && done < want; i += 32) {
            const uint64_t j = i + lane;
            const bool own = j < b && pos[j] >= 0 && pos[j] < owned_len;
            const unsigned m = __ballot_sync(0xffffffffu, own);
            const uint32_t at = done + __popc(m & ((1u << lane) - 1u));
            if (own && at < want) send[o + at] = (int64_t)pos[j] + text_start;
            done += __popc(m);
        }
    }
}

// one warp per pattern: out[hit_off[p] + (contributions of lower ranks) + t] = recv_r[roff_r[p] + t]; rank r's contribution
// starts at recv + rank_base[r], its per-pattern offsets are the exclusive scan of take[r][.] — computed here by a running sum
// kept per (rank, pattern) in `roff` (exclusive scans done by the caller with k_scan_*).
__global__ void k_shard_merge(const int64_t* __restrict__ recv, const uint64_t* __restrict__ rank_base, const uint64_t* __restrict__ roff,
                              const int32_t* __restrict__ take, const uint64_t* __restrict__ out_off, uint32_t n_pat, uint32_t world,
                              int64_t* __restrict__ out) {
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5, lane = threadIdx.x & 31u;
    for (uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < n_pat; p += warps) {
        uint64_t o = out_off[p];
        for (uint32_t r = 0; r < world; ++r) {
            const uint64_t k = (uint64_t)take[(uint64_t)r * n_pat + p];
            const int64_t* src = recv + rank_base[r] + roff[(uint64_t)r * (n_pat + 1) + p];
            for (uint64_t t = lane; t < k; t += 32) out[o + t] = src[t];
            o += k;
        }
    }
}

}  // namespace fmgpu
