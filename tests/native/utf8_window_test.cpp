// TEST INFRASTRUCTURE: the 4-byte window decoders of k_fused's stream stage (vaporetto_b200/csrc/utf8_window.hpp) on the
// host, over every window, against a byte-by-byte restatement of the rules of Rust's str::from_utf8 (what the
// reference's Sentence::parse_raw accepts, sentence.rs:160-196) applied to the window's first character.
//   decode_any:             every lead byte that is not a continuation byte x every combination of the bytes behind it
//   decode_ascii_or_three:  every ASCII / E0..EF lead x every combination of the two bytes behind it; must also agree
//                           with decode_any there
//   char_type / the BMP type table of the tile kernels (textnorm.hpp): every code point against the reference's ranges
// Prints "utf8 window ok <cases>" and exits 0, or the first mismatch and exits 1.
#include <cstdint>
#include <cstdio>

#include "../../vaporetto_b200/csrc/textnorm.hpp"
#include "../../vaporetto_b200/csrc/utf8_window.hpp"

namespace {

// first character of the byte string b[0..3] under str::from_utf8's rules: returns false when it is malformed
bool ref_decode(const uint8_t* b, uint32_t& cp, uint32_t& len) {
    const uint8_t b0 = b[0];
    auto cont = [](uint8_t x) { return (x & 0xC0) == 0x80; };
    if (b0 < 0x80) { cp = b0; len = 1; return true; }
    if (b0 < 0xC2) { len = b0 < 0xC0 ? 1 : 2; cp = 0; return false; }      // continuation byte / overlong 2-byte lead
    if (b0 < 0xE0) { len = 2; cp = ((b0 & 0x1Fu) << 6) | (b[1] & 0x3Fu); return cont(b[1]); }
    if (b0 < 0xF0) {
        len = 3;
        cp = ((b0 & 0x0Fu) << 12) | ((b[1] & 0x3Fu) << 6) | (b[2] & 0x3Fu);
        if (!cont(b[1]) || !cont(b[2])) return false;
        if (b0 == 0xE0 && b[1] < 0xA0) return false;                        // overlong
        if (b0 == 0xED && b[1] > 0x9F) return false;                        // surrogates
        return true;
    }
    if (b0 < 0xF5) {
        len = 4;
        cp = ((b0 & 0x07u) << 18) | ((b[1] & 0x3Fu) << 12) | ((b[2] & 0x3Fu) << 6) | (b[3] & 0x3Fu);
        if (!cont(b[1]) || !cont(b[2]) || !cont(b[3])) return false;
        if (b0 == 0xF0 && b[1] < 0x90) return false;                        // overlong
        if (b0 == 0xF4 && b[1] > 0x8F) return false;                        // > U+10FFFF
        return true;
    }
    len = 4; cp = 0;                                                        // F5..FF: never valid
    return false;
}

// CharacterType::get_type, range by range as the reference lists them (sentence.rs:50-67)
uint32_t ref_type(uint32_t c) {
    auto in = [c](uint32_t lo, uint32_t hi) { return c >= lo && c <= hi; };
    if (in(0x30, 0x39) || in(0xFF10, 0xFF19)) return 1;
    if (in(0x41, 0x5A) || in(0x61, 0x7A) || in(0xFF21, 0xFF3A) || in(0xFF41, 0xFF5A)) return 2;
    if (in(0x3040, 0x3096)) return 3;
    if (in(0x30A0, 0x30FA) || in(0x30FC, 0x30FF) || in(0xFF66, 0xFF9F)) return 4;
    if (in(0x3400, 0x4DBF) || in(0x4E00, 0x9FFF) || in(0xF900, 0xFAFF) || in(0x20000, 0x2A6DF) || in(0x2A700, 0x2B73F) ||
        in(0x2B740, 0x2B81F) || in(0x2B820, 0x2CEAF) || in(0x2F800, 0x2FA1F))
        return 5;
    return 6;
}

// the arithmetic char_type against the reference's ranges for every code point, and the tile kernels' BMP type table
// (page table + sub-tables of the mixed pages) against char_type for every BMP code point
bool check_types() {
    for (uint32_t c = 0; c < 0x110000u; ++c)
        if (vpt::char_type(c) != ref_type(c)) { printf("char_type mismatch at U+%04X: %u, want %u\n", c, vpt::char_type(c), ref_type(c)); return false; }
    uint8_t tab[vpt::kTypeTableBytes];
    for (int i = 0; i < vpt::kTypeTableBytes; ++i) tab[i] = uint8_t(vpt::type_table_entry(uint32_t(i)));
    for (uint32_t c = 0; c < 0x10000u; ++c)
        if (vpt::type_from_table(tab, c) != ref_type(c)) {
            printf("type table mismatch at U+%04X: %u, want %u\n", c, vpt::type_from_table(tab, c), ref_type(c));
            return false;
        }
    return true;
}

}  // namespace

int main() {
    if (!check_types()) return 1;
    uint64_t cases = 0;
    const uint8_t probes3[] = {0x00, 0x41, 0x7F, 0x80, 0xBF, 0xC0, 0xE3, 0xFF};
    for (uint32_t b0 = 0; b0 < 256; ++b0) {
        if ((b0 & 0xC0) == 0x80) continue;  // (a slot's lead byte is never a continuation byte: the scatter step picks starts)
        const bool four = b0 >= 0xF0;
        for (uint32_t b1 = 0; b1 < 256; ++b1)
            for (uint32_t b2 = 0; b2 < 256; ++b2)
                for (uint32_t k3 = 0; k3 < (four ? 256u : 8u); ++k3) {
                    const uint32_t b3 = four ? k3 : probes3[k3];
                    const uint8_t b[4] = {uint8_t(b0), uint8_t(b1), uint8_t(b2), uint8_t(b3)};
                    const uint32_t x = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
                    uint32_t rcp = 0, rlen = 0;
                    const bool rok = ref_decode(b, rcp, rlen);
                    bool bad = false;
                    uint32_t len = 0;
                    const uint32_t c = vpt::decode_any(x, bad, len);
                    ++cases;
                    // lengths: the lead byte's announcement (F8..FF announce four bytes like F0..F7; the window of an
                    // all-zero slot has length 0 and is never a character of a sentence: NUL is rejected before)
                    const uint32_t want_len = x == 0 ? 0u : (b0 < 0xC0 ? 1u : b0 < 0xE0 ? 2u : b0 < 0xF0 ? 3u : 4u);
                    if (len != want_len || bad != !rok || (rok && c != rcp)) {
                        printf("decode_any mismatch at %02x %02x %02x %02x: c=%x bad=%d len=%u, want c=%x ok=%d len=%u\n", b0, b1, b2,
                               b3, c, int(bad), len, rcp, int(rok), want_len);
                        return 1;
                    }
                    const bool ascii = (x & 0x80u) == 0, three = (x & 0xF0u) == 0xE0u;
                    if (ascii || three) {
                        bool bad2 = false;
                        uint32_t len2 = 0;
                        const uint32_t c2 = vpt::decode_ascii_or_three(x, ascii, three, bad2, len2);
                        if (c2 != c || bad2 != bad || len2 != len) {
                            printf("decode_ascii_or_three mismatch at %02x %02x %02x %02x: c=%x bad=%d len=%u, decode_any c=%x bad=%d len=%u\n",
                                   b0, b1, b2, b3, c2, int(bad2), len2, c, int(bad), len);
                            return 1;
                        }
                    }
                }
    }
    printf("utf8 window ok %llu\n", static_cast<unsigned long long>(cases));
    return 0;
}
