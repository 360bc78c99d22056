// Parameters and work-item inputs of the LF-walk kernels (locate / extract / extractUntilBoundary*); the lane code
// itself is in lf_lane.h.  Host/device code (the host layout test replays the lanes with plain memory reads).
//
// Reference semantics (fm/FmIndex.java): locate :526-548, extract :564-608, extractUntilBoundary :640-759,
// ...Left :772-831, ...Right :844-922.
#pragma once
#include <cstdint>

#include "lane_logic.h"

namespace fmgpu {

enum WalkMode : int { WM_LOCATE = 0, WM_EXTRACT = 1, WM_EUB = 2 };
// internal values of WalkParams::eub_mode beside FMGPU_MODE_BOTH / LEFT / RIGHT (0, 1, 2)
constexpr int32_t EUB_SCAN = 3, EUB_RECORD = 4;
constexpr int32_t REL_NONE = 0x3fffffff;  // no boundary before the destination overflows / the text ends

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
    // fused locate -> extractUntilBoundary (kernels_records.cuh): eub_mode EUB_SCAN walks LEFT only and claims the record that
    // starts at S = from - (chars walked) in `claims` (open addressing, entry = (S + 1) << 32 | id of the first hit that
    // claimed it); eub_mode EUB_RECORD extracts text[S, E) of every distinct record, `from` holding S
    unsigned long long* claims;
    uint32_t claim_mask;
    int32_t* win_of;    // EUB_SCAN: per hit the id of the hit that owns its record (itself: winner), -1: no record start found
    int32_t* at_bound;  // EUB_SCAN: 1 if text[from] is the boundary char
    int32_t offset;     // the reference's `offset` argument (extract :564, extractUntilBoundary* :640 / :772 / :844), >= 0
    uint16_t* left;     // left-part scratch, item i at i*dst_len, chars in walk order (text order reversed)
    int32_t* down_len;  // chars in the left part
    // outputs
    uint16_t* arena;
    int32_t* len_out;
    int32_t* status_out;
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

}  // namespace fmgpu
