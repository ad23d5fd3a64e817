// Decoders of a character's 4-byte window (the lead byte is the low byte of x: the four bytes from the lead on), used by
// the stream stage of k_fused (fused_kernel.cuh).  Plain arithmetic: tests/native/utf8_window_test.cpp runs both over
// every window on the host against a byte-by-byte restatement of str::from_utf8's rules (reference sentence.rs:160-196).
#pragma once
#include <cstdint>

#include "common.hpp"

namespace vpt {

// Code point of the character whose lead byte is the low byte of x.  `len` is the length the lead byte announces (0 for an
// empty slot); `bad` is set when the bytes after the lead are not the continuation bytes it asks for, for overlong forms,
// surrogates, values above U+10FFFF and the lead bytes F8..FF.  Together with "the lengths add up to the bytes of the
// tile" and "no sentence starts on a continuation byte" this is str::from_utf8 (reference sentence.rs:160-196).
VPT_HD uint32_t decode_any(uint32_t x, bool& bad, uint32_t& len) {
    const uint32_t b0 = x & 0xFFu;
    const uint32_t t = ((x << 4) & 0x3F000u) | ((x >> 10) & 0xFC0u) | ((x >> 24) & 0x3Fu);  // b1<<12 | b2<<6 | b3
    const uint32_t l = (b0 >= 0xC0u) + (b0 >= 0xE0u) + (b0 >= 0xF0u);                      // continuation bytes
    const uint32_t tail = t >> (18u - 6u * l);
    const uint32_t head = (b0 & (0x3Fu >> l)) << (6u * l);
    const uint32_t c = b0 < 0x80u ? b0 : (head | tail);
    const uint32_t minc = 1u << ((0x100B0700u >> (8u * l)) & 31u);  // 1, 0x80, 0x800, 0x10000
    const uint32_t cm = (0xC0C0C0C0u >> (8u * (3u - l))) & 0xFFFFFF00u;  // bits 7..6 of the bytes 1 .. l
    bad = b0 >= 0x80u && (c < minc || c - 0xD800u < 0x800u || c > 0x10FFFFu || b0 >= 0xF8u || (x & cm) != (0x80808080u & cm));
    len = x ? l + 1u : 0u;
    return c;
}

// The same for a window whose lead byte is ASCII or E0..EF (what a chunk of Japanese text holds):
// ascii = (x & 0x80) == 0, three = (x & 0xF0) == 0xE0, one of them true (the caller has them for its warp vote).
VPT_HD uint32_t decode_ascii_or_three(uint32_t x, bool ascii, bool three, bool& bad, uint32_t& len) {
    const uint32_t c3 = ((x & 0x0Fu) << 12) | ((x >> 2) & 0xFC0u) | ((x >> 16) & 0x3Fu);
    bad = three && (c3 < 0x800u || c3 - 0xD800u < 0x800u || (x & 0x00C0C000u) != 0x00808000u);
    len = three ? 3u : (x ? 1u : 0u);
    return ascii ? (x & 0xFFu) : c3;
}

}  // namespace vpt
