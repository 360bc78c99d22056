// Fused locate -> extractUntilBoundary (the reference's "extracting whole records" flow: README.md:98-107,
// jmh/.../FmIndexThroughputBenchmark.java:231-249): for every located hit the record around it, each DISTINCT record read
// from the index once.
//
//   1. k_extract<WM_EUB> in mode EUB_SCAN (lf_lane.h): per hit the LEFT walk of FmIndex.extractUntilBoundary (fm/FmIndex.java:664-686)
//      without storing chars; it yields down = chars left of the hit and the record start S = from - down, and the first hit to
//      arrive claims S in a hash table.
//   2. k_rec_mark + exclusive scan + k_rec_starts: the claiming hits ("winners") get consecutive record numbers u, rec_start[u] = S.
//   3. k_extract<WM_EUB> in mode EUB_RECORD: text[S, E) of every distinct record into row u of the record arena, rec_rel[u] = E - S.
//   4. k_rec_results: per hit, what extractUntilBoundary(from, new char[dst_len], 0, boundary) returns or throws — pure
//      arithmetic on (from, down, E): the reference's 4-char chunk loop (eub_right_chunks, lane_logic.h) decides length, the
//      "does not fit" status and its N per HIT (they depend on where the hit sits in its record), the chars are the record's.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "lane_logic.h"
#include "walk_lane.h"

namespace fmgpu {

__global__ void k_rec_mark(const int32_t* __restrict__ win_of, uint32_t n, int32_t* __restrict__ is_winner) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) is_winner[i] = win_of[i] == (int32_t)i ? 1 : 0;
}

// winners: rec_start[u] = from - down, u = uidx[hit]
__global__ void k_rec_starts(const int32_t* __restrict__ win_of, const int32_t* __restrict__ from, const int32_t* __restrict__ down,
                             const uint64_t* __restrict__ uidx, uint32_t n, int32_t* __restrict__ rec_start) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && win_of[i] == (int32_t)i) rec_start[uidx[i]] = from[i] - down[i];
}

// per hit: record number, returned length / N, status.  scan_status: what the scan's bounds checks found (FmIndex.java:610-626, :658)
__global__ void k_rec_results(const int32_t* __restrict__ from, const int32_t* __restrict__ down, const int32_t* __restrict__ win_of,
                              const int32_t* __restrict__ at_bound, const uint64_t* __restrict__ uidx, const int32_t* __restrict__ rec_rel,
                              uint32_t n, int32_t length, int32_t dst_len, int32_t* __restrict__ rec_index, int32_t* __restrict__ len_out,
                              int32_t* __restrict__ status /* in: scan status, out: final */) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (status[i] != 0) {  // the reference throws before any walk
        rec_index[i] = -1;
        len_out[i] = 0;
        return;
    }
    const int32_t f = from[i], d = down[i], w = win_of[i];
    int32_t rel, u = -1;
    if (w < 0) {
        // the left part alone fills the destination: the chunk loop throws on its first char unless text[from] is the boundary
        rel = at_bound[i] ? 0 : REL_NONE;
    } else {
        u = (int32_t)uidx[w];
        const int32_t r = rec_rel[u];
        rel = (r >= 0 && r != REL_NONE) ? (f - d + r) - f : r;
    }
    const EubOut o = eub_right_chunks(f, d, rel, length, dst_len, false, 0);
    rec_index[i] = u;
    len_out[i] = o.value;
    status[i] = o.status;
}

}  // namespace fmgpu
