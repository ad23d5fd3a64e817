// textnorm.hpp — character-level helpers shared by the kernels (and the host): UTF-8 decoding, the reference's
// character types, and the character map of the reference's KyteaFullwidthFilter
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

// CharacterType::get_type (reference sentence.rs:50-67): 1 Digit, 2 Roman, 3 Hiragana, 4 Katakana, 5 Kanji, 6 Other
VPT_HD uint32_t char_type(uint32_t c) {
    if (c < 0x80) {
        if (c - 0x30u <= 9u) return 1;
        if ((c | 0x20u) - 0x61u <= 25u) return 2;
        return 6;
    }
    if (c - 0x3040u <= 0x56u) return 3;                                  // 3040..3096
    if (c - 0x30A0u <= 0x5Au || c - 0x30FCu <= 3u) return 4;             // 30A0..30FA, 30FC..30FF
    if (c - 0x4E00u <= 0x51FFu || c - 0x3400u <= 0x19BFu) return 5;      // 4E00..9FFF, 3400..4DBF
    if (c < 0xF900u) return 6;
    if (c <= 0xFAFFu) return 5;                                          // F900..FAFF
    if (c - 0xFF10u <= 9u) return 1;
    if (c - 0xFF21u <= 25u || c - 0xFF41u <= 25u) return 2;
    if (c - 0xFF66u <= 0x39u) return 4;                                  // FF66..FF9F
    if (c < 0x20000u) return 6;
    if (c <= 0x2A6DFu || c - 0x2A700u <= 0x103Fu || c - 0x2B740u <= 0xDFu || c - 0x2B820u <= 0x168Fu ||
        c - 0x2F800u <= 0x21Fu)
        return 5;
    return 6;
}

// Character types of the BMP by table, for the tile kernels' shared memory: entries [0, 256) are indexed by c >> 8 and
// hold the type of the whole 256-code-point page, or 0x80 | k for the four pages that hold more than one type (00: ASCII,
// 30: kana, 4D: CJK Extension A ends at U+4DBF, FF: full-width / half-width forms); entries [256 + 256 k, 256 + 256 (k + 1))
// are sub-table k, indexed by c & 255.  (tests/native/utf8_window_test.cpp checks every BMP code point against the
// reference's ranges: every other page is uniform.  Until the last session of round 2 page 4D was taken for uniform:
// U+4DC0..U+4DFF, the Yijing hexagram symbols, were typed Kanji instead of Other by the tile kernels.)
constexpr int kTypeTableBytes = 256 + 4 * 256;
VPT_HD uint32_t type_table_entry(uint32_t i) {
    if (i < 256u) return i == 0x00u ? 0x80u : i == 0x30u ? 0x81u : i == 0xFFu ? 0x82u : i == 0x4Du ? 0x83u : char_type(i << 8);
    const uint32_t k = (i - 256u) >> 8;
    const uint32_t page = k == 0u ? 0x00u : k == 1u ? 0x30u : k == 2u ? 0xFFu : 0x4Du;
    return char_type((page << 8) | (i & 255u));
}
// type of c < 0x10000 from the table built of type_table_entry(0 .. kTypeTableBytes)
VPT_HD uint32_t type_from_table(const uint8_t* tab, uint32_t c) {
    uint32_t ty = tab[c >> 8];
    if (ty & 0x80u) ty = tab[256u + ((ty & 3u) << 8) + (c & 255u)];
    return ty;
}

// Decodes the code point whose lead byte is the low byte of x (valid UTF-8 assumed).
VPT_HD uint32_t decode_cp(uint32_t x) {
    const uint32_t b0 = x & 0xFF;
    if (b0 < 0x80) return b0;
    const uint32_t b1 = (x >> 8) & 0x3F;
    if (b0 < 0xE0) return ((b0 & 0x1F) << 6) | b1;
    const uint32_t b2 = (x >> 16) & 0x3F;
    if (b0 < 0xF0) return ((b0 & 0x0F) << 12) | (b1 << 6) | b2;
    const uint32_t b3 = (x >> 24) & 0x3F;
    return ((b0 & 0x07) << 18) | (b1 << 12) | (b2 << 6) | b3;
}

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
