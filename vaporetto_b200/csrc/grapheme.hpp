// Extended grapheme clusters (UAX #29, rules GB3 .. GB13 / GB999) as a per-character state machine, shared by host and
// device code.  Used by the `--wsconst G` post-filter: the reference's ConcatGraphemeClustersFilter
// (vaporetto_rules/src/sentence_filters/concat_grapheme_clusters.rs:10-35) clears every boundary inside a cluster of
// `unicode_segmentation::graphemes(true)`.  Property data: grapheme_tables.hpp (generated, see its header).
#pragma once
#include <cstdint>

#include "common.hpp"
#include "grapheme_tables.hpp"

namespace vpt {

enum : uint32_t { kGcbOther = 0, kGcbCR, kGcbLF, kGcbControl, kGcbExtend, kGcbZWJ, kGcbRI, kGcbPrepend, kGcbSpacingMark,
                  kGcbL, kGcbV, kGcbT, kGcbLV, kGcbLVT };

// class word of a code point: binary search over the sorted ranges (`table` = kGraphemeTable or its device copy)
VPT_HD uint32_t grapheme_class(const GraphemeRange* table, uint32_t c) {
    int lo = 0, hi = kGraphemeRanges - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const uint32_t l = table[mid].lo, h = table[mid].hi;
        if (c < l) hi = mid - 1;
        else if (c > h) lo = mid + 1;
        else return table[mid].cls;
    }
    return 0;
}

struct GraphemeState {
    uint32_t prev = 0;      // class word of the previous character
    uint32_t ri_odd = 0;    // the run of Regional_Indicators ending at the previous character has odd length
    uint32_t emoji = 0;     // 1: Extended_Pictographic Extend*   2: ... followed by ZWJ
    uint32_t incb = 0;      // 1: InCB=Consonant [Extend|Linker]* without a Linker yet   2: with a Linker
    uint32_t started = 0;
};

// Feeds the next character; returns true when a cluster boundary lies BEFORE it (false for the first character).
VPT_HD bool grapheme_step(GraphemeState& st, uint32_t cw) {
    const uint32_t c = cw & 15u, p = st.prev & 15u;
    const bool ext_pict = (cw & 0x10u) != 0;
    const uint32_t incb = (cw >> 5) & 3u;
    bool brk;
    if (!st.started) brk = false;
    else if (p == kGcbCR && c == kGcbLF) brk = false;                                              // GB3
    else if (p == kGcbControl || p == kGcbCR || p == kGcbLF) brk = true;                           // GB4
    else if (c == kGcbControl || c == kGcbCR || c == kGcbLF) brk = true;                           // GB5
    else if (p == kGcbL && (c == kGcbL || c == kGcbV || c == kGcbLV || c == kGcbLVT)) brk = false; // GB6
    else if ((p == kGcbLV || p == kGcbV) && (c == kGcbV || c == kGcbT)) brk = false;               // GB7
    else if ((p == kGcbLVT || p == kGcbT) && c == kGcbT) brk = false;                              // GB8
    else if (c == kGcbExtend || c == kGcbZWJ) brk = false;                                         // GB9
    else if (c == kGcbSpacingMark) brk = false;                                                    // GB9a
    else if (p == kGcbPrepend) brk = false;                                                        // GB9b
    else if (st.incb == 2 && incb == 1) brk = false;                                               // GB9c
    else if (st.emoji == 2 && ext_pict) brk = false;                                               // GB11
    else if (p == kGcbRI && c == kGcbRI && st.ri_odd) brk = false;                                 // GB12, GB13
    else brk = true;                                                                               // GB999
    // state after this character
    st.ri_odd = c == kGcbRI ? ((p == kGcbRI && st.started) ? st.ri_odd ^ 1u : 1u) : 0u;
    if (ext_pict) st.emoji = 1;
    else if (st.emoji == 1 && c == kGcbExtend) st.emoji = 1;
    else if (st.emoji == 1 && c == kGcbZWJ) st.emoji = 2;
    else st.emoji = 0;
    if (incb == 1) st.incb = 1;
    else if (st.incb && incb == 3) st.incb = 2;
    else if (st.incb && incb == 2) { /* keep */ }
    else st.incb = 0;
    st.prev = cw;
    st.started = 1;
    return brk;
}

}  // namespace vpt
