// textnorm.hpp — the character map of the reference's KyteaFullwidthFilter
// (vaporetto_rules/src/string_filters/kytea_fullwidth.rs:13-118): the pre-filter the `predict` CLI applies to
// every line before prediction unless --no-norm is given (predict/src/main.rs:98,154).  One character maps to one
// character, so boundaries computed on the filtered text apply to the original text unchanged.
//
// Restated as arithmetic instead of a match table: printable ASCII except ' ' # $ ; \ ^ ` | ~ moves to the
// full-width block (+0xFEE0), with four exceptions, plus ten non-ASCII sources (half-width CJK punctuation and
// dash look-alikes).  Checked entry by entry against tests/golden/kytea_fullwidth_map.json.
#pragma once
#include <cstdint>

#include "common.hpp"

namespace vpt {

VPT_HD uint32_t kytea_fullwidth(uint32_t c) {
    if (c < 0x80u) {
        if (c < 0x21u || c > 0x7Eu) return c;
        // unmapped: # $ ; (below 64) and \ ^ ` | ~ (64 and above)
        const uint64_t keep_lo = (1ull << 0x23) | (1ull << 0x24) | (1ull << 0x3B);
        const uint64_t keep_hi = (1ull << (0x5C - 64)) | (1ull << (0x5E - 64)) | (1ull << (0x60 - 64)) |
                                 (1ull << (0x7C - 64)) | (1ull << (0x7E - 64));
        if (((c < 64u ? keep_lo >> c : keep_hi >> (c - 64u)) & 1ull) != 0) return c;
        if (c == 0x22u) return 0x201Du;  // " -> right double quotation mark
        if (c == 0x27u) return 0x2019u;  // ' -> right single quotation mark
        if (c == 0x2Du) return 0x2212u;  // - -> minus sign
        if (c == 0x2Eu) return 0x3002u;  // . -> ideographic full stop
        return c + 0xFEE0u;
    }
    if (c < 0x2013u || (c > 0x2500u && c < 0xFF0Du) || c > 0xFF65u) return c;
    switch (c) {
        case 0x2013u: case 0x2015u: case 0x2500u: case 0xFF0Du: return 0x30FCu;  // dashes -> prolonged sound mark
        case 0xFF5Eu: return 0x301Cu;  // fullwidth tilde -> wave dash
        case 0xFF61u: return 0x3002u;  // halfwidth ideographic full stop
        case 0xFF62u: return 0x300Cu;  // halfwidth corner brackets
        case 0xFF63u: return 0x300Du;
        case 0xFF64u: return 0x3001u;  // halfwidth ideographic comma
        case 0xFF65u: return 0x30FBu;  // halfwidth katakana middle dot
        default: return c;
    }
}

}  // namespace vpt
