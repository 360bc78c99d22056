// FmIndex.convertBytePatternToCharPattern (fm/FmIndex.java:239-298) as per-lane code: one UTF-8 sequence -> one UTF-16 unit,
// with the reference's branch structure (any negative first byte that is neither 1111xxxx nor 111xxxxx is read as a two-byte
// sequence, code points above 32767 of a four-byte sequence are an error, three-byte results are cut to 16 bits).
// Host/device code: k_prepass_utf8 (kernels_utf8.cuh) runs it, tests/support/flatcheck.cpp replays it against the oracle.
#pragma once
#include <cstdint>

#include "lane_logic.h"

namespace fmgpu {

constexpr int32_t ST_INDEX_OOB_ = 9;      // FMGPU_ST_INDEX_OOB: a sequence runs past the end of the pattern's byte[]
constexpr int32_t ST_CHAR_EXCEEDS_ = 10;  // FMGPU_ST_CHAR_EXCEEDS ("Found a character that exceeds (32767): it was N")

// Decodes the sequence that starts at bytes[pos] (pos < end).  Returns 0 and sets *ch / *adv, or the status of the Java
// exception (*value = the offending code point for ST_CHAR_EXCEEDS_).
FMGPU_HD int32_t utf8_next(const uint8_t* bytes, uint64_t pos, uint64_t end, uint32_t* ch, uint32_t* adv, int32_t* value) {
    const uint32_t b0 = bytes[pos];
    if (b0 < 0x80u) {  // :289 single byte
        *ch = b0;
        *adv = 1;
        return 0;
    }
    if ((b0 & 0xF0u) == 0xF0u) {  // :249 four bytes
        if (pos + 3 >= end) return ST_INDEX_OOB_;
        const uint32_t cp = (((b0 & 0x07u) << 18) | ((bytes[pos + 1] & 0x3Fu) << 12) | ((bytes[pos + 2] & 0x3Fu) << 6) | (bytes[pos + 3] & 0x3Fu)) & 0x1FFFFFu;
        if (cp > 32767u) {  // :261-267
            *value = (int32_t)cp;
            return ST_CHAR_EXCEEDS_;
        }
        *ch = cp;
        *adv = 4;
        return 0;
    }
    if ((b0 & 0xE0u) == 0xE0u) {  // :270 three bytes
        if (pos + 2 >= end) return ST_INDEX_OOB_;
        *ch = (((b0 & 0x0Fu) << 12) | ((bytes[pos + 1] & 0x3Fu) << 6) | (bytes[pos + 2] & 0x3Fu)) & 0xFFFFu;
        *adv = 3;
        return 0;
    }
    if (pos + 1 >= end) return ST_INDEX_OOB_;  // :280 two bytes
    *ch = (((b0 & 0x1Fu) << 6) | (bytes[pos + 1] & 0x3Fu)) & 0x7FFu;
    *adv = 2;
    return 0;
}

// Whole pattern bytes[a, b) -> chars written at out[0..) (*last = the final one): returns the char count or -status.
FMGPU_HD int64_t utf8_convert(const uint8_t* bytes, uint64_t a, uint64_t b, uint16_t* out, int32_t* value, uint32_t* last) {
    uint64_t pos = a, n = 0;
    while (pos < b) {
        uint32_t ch = 0, adv = 0;
        const int32_t st = utf8_next(bytes, pos, b, &ch, &adv, value);
        if (st) return -(int64_t)st;
        out[n++] = (uint16_t)ch;
        *last = ch;
        pos += adv;
    }
    return (int64_t)n;
}

}  // namespace fmgpu
