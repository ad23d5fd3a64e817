// vaporetto_b200 — sm_100a kernels for the Predictor::predict hot path.
//
// Replaces (reference, vaporetto/src): predictor.rs:518-543 (predict), char_scorer/boundary_scorer.rs:93-113
// and char_scorer/boundary_tag_scorer.rs:121-147 (pattern walk + weight add), predictor.rs:176-213
// (PositionalWeight::add_score), type_scorer/boundary_scorer_cache.rs:59-81 (type table),
// type_scorer/boundary_scorer.rs:64-80 / boundary_tag_scorer.rs:95-116 (type automaton variants),
// sentence.rs:50-67,160-196 (get_type / parse_raw: done on device from raw UTF-8).
//
// Pipeline per batch:
//   k_count        one warp per sentence: chars per sentence, validation (empty / NUL / bad UTF-8),
//                  group-local exclusive offsets (64 sentences per CTA)
//   k_scan_groups  one CTA: exclusive scan of the per-group totals
//   k_score_fast   one warp per sentence, one lane per character: UTF-8 decode into a per-warp shared
//                  ring, longest-suffix lookup in the perfect-hash node table (one 32-byte LDG.256 record
//                  per probe), warp-shuffle gather of the 6-wide weight rows into per-boundary sums,
//                  type-table add, bias, threshold, coalesced stores.
//   k_score_general  same skeleton for models whose rows do not fit the inline window, for the type
//                  automaton variant and for tag-state output: rows scatter with global atomics.
// No tensor cores: integer indexing + scatter/gather add.
#include <cstdint>

#include <cuda_runtime.h>

#include "device_model.hpp"
#include "keys.hpp"

namespace vpt {

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kRing = 256;          // per-warp ring of decoded characters (power of two)
constexpr int kRingMask = kRing - 1;
constexpr unsigned kFull = 0xFFFFFFFFu;

struct Rec32 {
    uint32_t v[8];
};

__device__ __forceinline__ Rec32 load_record(const void* base, uint32_t slot) {
    Rec32 r;
    const char* p = static_cast<const char*>(base) + (size_t(slot) << 5);
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]),
                   "=r"(r.v[7])
                 : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t slot_of(const DevTable& t, uint64_t key) {
    const uint64_t h = mix64(key + t.salt);
    const uint32_t bucket = __umulhi(uint32_t(h >> 32), t.nbuckets);
    const uint32_t seed = __ldg(t.seeds + bucket);
    const uint64_t h2 = mix64(key ^ (uint64_t(seed) + 1) * 0x9E3779B97F4A7C15ULL);
    return __umulhi(uint32_t(h2 >> 32), t.nslots);
}

// One probe: returns true when the node with `key` exists; rec/slot are valid then.
__device__ __forceinline__ bool probe(const DevTable& t, uint64_t key, Rec32& rec, uint32_t& slot) {
    slot = slot_of(t, key);
    rec = load_record(t.records, slot);
    const uint64_t k = (uint64_t(rec.v[1]) << 32) | rec.v[0];
    return (k & ~kExtFlag) == key;
}

// CharacterType::get_type (reference sentence.rs:50-67)
__device__ __forceinline__ uint32_t char_type(uint32_t c) {
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

// Decodes the code point whose lead byte is the low byte of x (valid UTF-8 assumed).
__device__ __forceinline__ uint32_t decode_cp(uint32_t x) {
    const uint32_t b0 = x & 0xFF;
    if (b0 < 0x80) return b0;
    const uint32_t b1 = (x >> 8) & 0x3F;
    if (b0 < 0xE0) return ((b0 & 0x1F) << 6) | b1;
    const uint32_t b2 = (x >> 16) & 0x3F;
    if (b0 < 0xF0) return ((b0 & 0x0F) << 12) | (b1 << 6) | b2;
    const uint32_t b3 = (x >> 24) & 0x3F;
    return ((b0 & 0x07) << 18) | (b1 << 12) | (b2 << 6) | b3;
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(kFull, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

struct Rings {
    uint32_t cp[kRing];  // code points
    uint32_t bp[kRing];  // byte position of the character relative to the sentence start
    uint8_t ty[kRing];   // character types
};

// Decodes the 128-byte window starting at the 4-byte aligned position `wpos`, appends its characters
// (those whose lead byte lies in [b0, b1)) to the ring at index nd.., returns how many were appended.
__device__ __forceinline__ uint32_t decode_window(const uint8_t* __restrict__ text, uint64_t wpos, uint64_t b0,
                                                  uint64_t b1, uint32_t nd, Rings& r, int lane) {
    const uint64_t addr = wpos + 4u * uint32_t(lane);
    uint32_t lo = 0, hi = 0;
    if (addr < b1) {
        lo = __ldg(reinterpret_cast<const uint32_t*>(text + addr));
        if (addr + 4 < b1) hi = __ldg(reinterpret_cast<const uint32_t*>(text + addr + 4));
    }
    uint32_t smask = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint64_t p = addr + j;
        const uint32_t b = (lo >> (8 * j)) & 0xFF;
        if (p >= b0 && p < b1 && (b & 0xC0) != 0x80) smask |= 1u << j;
    }
    const uint32_t cnt = __popc(smask);
    const uint32_t incl = warp_incl_scan(cnt, lane);
    uint32_t idx = nd + incl - cnt;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (smask & (1u << j)) {
            const uint32_t x = __funnelshift_r(lo, hi, 8 * j);
            const uint32_t c = decode_cp(x);
            r.cp[idx & kRingMask] = c;
            r.bp[idx & kRingMask] = uint32_t(addr + j - b0);
            r.ty[idx & kRingMask] = uint8_t(char_type(c));
            ++idx;
        }
    }
    return __shfl_sync(kFull, incl, 31);
}

// Steps back from byte position `pos` (a character start, > b0) to the previous character; returns its
// code point and updates pos.
__device__ __forceinline__ uint32_t prev_char(const uint8_t* __restrict__ text, uint64_t b0, uint64_t& pos) {
    uint64_t q = pos - 1;
    uint32_t x = __ldg(text + q);
    uint32_t bytes = x;
    while ((x & 0xC0) == 0x80 && q > b0) {
        --q;
        x = __ldg(text + q);
        bytes = (bytes << 8) | x;
    }
    pos = q;
    return decode_cp(bytes);
}

// ------------------------------------------------------------------------------------------------
// k_count
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_count(BatchArgs a) {
    __shared__ uint32_t s_nout[kGroup], s_nch[kGroup];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t gbase = uint64_t(blockIdx.x) * kGroup;
    const uint8_t* __restrict__ text = a.text;
    for (int i = warp; i < kGroup; i += kWarpsPerBlock) {
        const uint64_t s = gbase + i;
        uint32_t nch = 0, nout = 0;
        if (s < a.n_sent) {
            const uint64_t b0 = a.offsets[s], b1 = a.offsets[s + 1];
            uint32_t starts = 0, conts = 0, expect = 0, flags = 0;  // flags: 1 NUL, 2 malformed
            for (uint64_t wpos = b0 & ~3ull; wpos < b1; wpos += 128) {
                const uint64_t addr = wpos + 4u * uint32_t(lane);
                if (addr < b1) {
                    const uint32_t lo = __ldg(reinterpret_cast<const uint32_t*>(text + addr));
                    const uint32_t hi = (addr + 4 < b1) ? __ldg(reinterpret_cast<const uint32_t*>(text + addr + 4)) : 0u;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint64_t p = addr + j;
                        if (p < b0 || p >= b1) continue;
                        const uint32_t x = __funnelshift_r(lo, hi, 8 * j);
                        const uint32_t b = x & 0xFF;
                        if ((b & 0xC0) == 0x80) { ++conts; continue; }
                        ++starts;
                        if (b < 0x80) { if (b == 0) flags |= 1; continue; }
                        const uint32_t c1 = (x >> 8) & 0xFF, c2 = (x >> 16) & 0xFF, c3 = x >> 24;
                        uint32_t len;
                        bool ok;
                        if (b < 0xC2) { len = 1; ok = false; }
                        else if (b < 0xE0) { len = 2; ok = (c1 & 0xC0) == 0x80; }
                        else if (b < 0xF0) {
                            len = 3;
                            ok = (c1 & 0xC0) == 0x80 && (c2 & 0xC0) == 0x80 && !(b == 0xE0 && c1 < 0xA0) &&
                                 !(b == 0xED && c1 > 0x9F);
                        } else if (b < 0xF5) {
                            len = 4;
                            ok = (c1 & 0xC0) == 0x80 && (c2 & 0xC0) == 0x80 && (c3 & 0xC0) == 0x80 &&
                                 !(b == 0xF0 && c1 < 0x90) && !(b == 0xF4 && c1 > 0x8F);
                        } else { len = 1; ok = false; }
                        if (p + len > b1) ok = false;
                        if (!ok) flags |= 2;
                        expect += len - 1;
                    }
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                starts += __shfl_xor_sync(kFull, starts, d);
                conts += __shfl_xor_sync(kFull, conts, d);
                expect += __shfl_xor_sync(kFull, expect, d);
                flags |= __shfl_xor_sync(kFull, flags, d);
            }
            if (conts != expect) flags |= 2;
            nch = starts;
            nout = nch > 0 ? nch - 1 : 0;
            const int st = (flags & 2) ? 3 : (flags & 1) ? 2 : (nch == 0 ? 1 : 0);
            if (lane == 0) {
                a.n_chars[s] = nch;
                a.status[s] = st;
            }
        }
        if (lane == 0) { s_nout[i] = nout; s_nch[i] = nch; }
    }
    __syncthreads();
    if (warp == 0) {
        const uint32_t v0 = s_nout[2 * lane], v1 = s_nout[2 * lane + 1];
        const uint32_t u0 = s_nch[2 * lane], u1 = s_nch[2 * lane + 1];
        const uint32_t iv = warp_incl_scan(v0 + v1, lane), iu = warp_incl_scan(u0 + u1, lane);
        const uint64_t s0 = gbase + 2 * lane;
        if (s0 < a.n_sent) { a.local_bound[s0] = iv - v0 - v1; a.local_char[s0] = iu - u0 - u1; }
        if (s0 + 1 < a.n_sent) { a.local_bound[s0 + 1] = iv - v1; a.local_char[s0 + 1] = iu - u1; }
        if (lane == 31) { a.group_bound[blockIdx.x] = iv; a.group_char[blockIdx.x] = iu; }
    }
}

// Exclusive scan of the per-group totals (in place); element [ngroups] receives the grand total.
__global__ void __launch_bounds__(1024) k_scan_groups(uint64_t* gb, uint64_t* gc, uint64_t ngroups) {
    __shared__ uint64_t s_wb[32], s_wc[32];
    __shared__ uint64_t s_carry[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_carry[0] = 0; s_carry[1] = 0; }
    __syncthreads();
    for (uint64_t base = 0; base < ngroups; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const uint64_t vb = i < ngroups ? gb[i] : 0, vc = i < ngroups ? gc[i] : 0;
        uint64_t ib = vb, ic = vc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t ob = __shfl_up_sync(kFull, ib, d), oc = __shfl_up_sync(kFull, ic, d);
            if (lane >= d) { ib += ob; ic += oc; }
        }
        if (lane == 31) { s_wb[warp] = ib; s_wc[warp] = ic; }
        __syncthreads();
        if (warp == 0) {
            uint64_t wb = s_wb[lane], wc = s_wc[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t ob = __shfl_up_sync(kFull, wb, d), oc = __shfl_up_sync(kFull, wc, d);
                if (lane >= d) { wb += ob; wc += oc; }
            }
            s_wb[lane] = wb;
            s_wc[lane] = wc;
        }
        __syncthreads();
        const uint64_t pb = (warp ? s_wb[warp - 1] : 0) + s_carry[0], pc = (warp ? s_wc[warp - 1] : 0) + s_carry[1];
        if (i < ngroups) { gb[i] = pb + ib - vb; gc[i] = pc + ic - vc; }
        __syncthreads();
        if (threadIdx.x == 0) { s_carry[0] += s_wb[31]; s_carry[1] += s_wc[31]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { gb[ngroups] = s_carry[0]; gc[ngroups] = s_carry[1]; }
}

// ------------------------------------------------------------------------------------------------
// shared per-sentence prologue
// ------------------------------------------------------------------------------------------------
struct SentInfo {
    uint64_t b0, b1, obase, cbase;
    uint32_t n, nout;
    int status;
};

__device__ __forceinline__ SentInfo sentence_info(const BatchArgs& a, uint64_t s, int lane) {
    SentInfo si;
    si.b0 = a.offsets[s];
    si.b1 = a.offsets[s + 1];
    si.n = a.n_chars[s];
    si.status = a.status[s];
    si.nout = si.n > 0 ? si.n - 1 : 0;
    const uint64_t grp = s / kGroup;
    si.obase = a.group_bound[grp] + a.local_bound[s];
    si.cbase = a.group_char[grp] + a.local_char[s];
    if (lane == 0) {
        a.bound_offsets[s] = si.obase;
        if (a.char_offsets) a.char_offsets[s] = si.cbase;
        if (s + 1 == a.n_sent) {
            a.bound_offsets[s + 1] = si.obase + si.nout;
            if (a.char_offsets) a.char_offsets[s + 1] = si.cbase + si.n;
        }
    }
    return si;
}

// Type-table index of boundary g: the 2W character types around it, zero outside the sentence
// (reference type_scorer/boundary_scorer_cache.rs:59-81).
__device__ __forceinline__ uint32_t type_index(const Rings& r, int64_t g, uint32_t n, int w) {
    uint32_t idx = 0;
    for (int k = 0; k < 2 * w; ++k) {
        const int64_t i = g - w + 1 + k;
        const uint32_t t = (i >= 0 && i < int64_t(n)) ? r.ty[i & kRingMask] : 0u;
        idx = (idx << 3) | t;
    }
    return idx;
}

// Longest-suffix lookup for the text ending at ring index g.  Symbols are code points (types=false)
// or character types (types=true).  Returns true and the record of the deepest existing node.
template <bool kTypes>
__device__ __forceinline__ bool find_node(const DevTable& t, const Rings& r, const uint8_t* __restrict__ text,
                                          uint64_t b0, uint32_t g, Rec32& rec, uint32_t& slot) {
    uint32_t c3, c2 = 0, c1 = 0;
    if (kTypes) {
        c3 = r.ty[g & kRingMask];
        if (g >= 1) c2 = r.ty[(g - 1) & kRingMask];
        if (g >= 2) c1 = r.ty[(g - 2) & kRingMask];
    } else {
        c3 = r.cp[g & kRingMask];
        if (g >= 1) c2 = r.cp[(g - 1) & kRingMask];
        if (g >= 2) c1 = r.cp[(g - 2) & kRingMask];
    }
    bool found = probe(t, shallow_key(c1, c2, c3), rec, slot);
    const bool depth3 = found && c1 != 0;
    if (!found && c1 != 0) found = probe(t, shallow_key(0, c2, c3), rec, slot);
    if (!found && c2 != 0) found = probe(t, shallow_key(0, 0, c3), rec, slot);
    if (depth3 && (rec.v[1] >> 31) && g >= 3) {
        // deeper nodes exist: keep walking backwards through the text (rare: patterns longer than 3)
        uint64_t pos = b0 + r.bp[(g - 2) & kRingMask];
        uint32_t node = __ldg(t.slot_node + slot);
        while (pos > b0) {
            uint32_t sym = prev_char(text, b0, pos);
            if (kTypes) sym = char_type(sym);
            Rec32 nrec;
            uint32_t nslot;
            if (!probe(t, deep_key(node, sym), nrec, nslot)) break;
            rec = nrec;
            slot = nslot;
            if (!(rec.v[1] >> 31)) break;
            node = __ldg(t.slot_node + slot);
        }
    }
    return found;
}

// ------------------------------------------------------------------------------------------------
// k_score_fast
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_score_fast(DevModel m, BatchArgs a) {
    __shared__ Rings s_rings[kWarpsPerBlock];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t s = uint64_t(blockIdx.x) * kWarpsPerBlock + warp;
    if (s >= a.n_sent) return;
    Rings& r = s_rings[warp];
    const SentInfo si = sentence_info(a, s, lane);
    const uint8_t* __restrict__ text = a.text;
    const uint32_t n = si.n;
    if (si.status != 0) {
        for (uint32_t i = lane; i < si.nout; i += 32) { a.scores[si.obase + i] = 0; a.boundaries[si.obase + i] = 0; }
        if (a.char_states) for (uint32_t i = lane; i < n; i += 32) a.char_states[si.cbase + i] = kNoPattern;
        if (a.type_states) for (uint32_t i = lane; i < n; i += 32) a.type_states[si.cbase + i] = kNoPattern;
        return;
    }
    const int tw = m.type_cache_window;
    const int r0 = m.ct.r0;
    uint64_t wpos = si.b0 & ~3ull;
    uint32_t nd = 0;
    int32_t prev_main = 0, carry_r = 0;
    uint32_t prev_g = 0;
    bool have_prev = false;
    for (uint32_t cb = 0; cb < n; cb += 32) {
        const uint32_t need = min(n, cb + 32u + uint32_t(tw));
        while (nd < need && wpos < si.b1) {
            nd += decode_window(text, wpos, si.b0, si.b1, nd, r, lane);
            wpos += 128;
        }
        __syncwarp();
        const uint32_t g = cb + lane;
        const bool active = g < n;
        int32_t d[kInlineWidth];
#pragma unroll
        for (int j = 0; j < kInlineWidth; ++j) d[j] = 0;
        if (active && m.ct.present) {
            Rec32 rec;
            uint32_t slot;
            if (find_node<false>(m.ct, r, text, si.b0, g, rec, slot)) {
#pragma unroll
                for (int j = 0; j < kInlineWidth; ++j) d[j] = int32_t(rec.v[2 + j]);
            }
        }
        // gather: boundary (lane) <- row entry j of the source lane (lane - r0 - j); sources that fall
        // into the neighbouring 32-character chunks are carried in registers (same lane index).
        int32_t mainv = 0, to_prev = 0, to_next = 0;
#pragma unroll
        for (int j = 0; j < kInlineWidth; ++j) {
            const int src = lane - r0 - j;
            const int32_t v = __shfl_sync(kFull, d[j], src & 31);
            if (src < 0) to_next += v;
            else if (src >= 32) to_prev += v;
            else mainv += v;
        }
        int32_t tsc = 0;
        if (tw > 0 && g + 1 < n) tsc = __ldg(m.type_cache + type_index(r, int64_t(g), n, tw));
        mainv += m.bias + tsc + carry_r;
        if (have_prev && prev_g + 1 < n) {
            const int32_t fin = prev_main + to_prev;
            a.scores[si.obase + prev_g] = fin;
            a.boundaries[si.obase + prev_g] = fin > 0 ? 1 : 0;
        }
        if (a.char_states && active) a.char_states[si.cbase + g] = kNoPattern;
        if (a.type_states && active) a.type_states[si.cbase + g] = kNoPattern;
        prev_main = mainv;
        prev_g = g;
        have_prev = true;
        carry_r = to_next;
        __syncwarp();
    }
    if (have_prev && prev_g + 1 < n) {
        a.scores[si.obase + prev_g] = prev_main;
        a.boundaries[si.obase + prev_g] = prev_main > 0 ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------------
// k_score_general
// ------------------------------------------------------------------------------------------------
template <bool kTypes>
__device__ __forceinline__ void scatter_general(const DevTable& t, const Rings& r, const uint8_t* __restrict__ text,
                                                const SentInfo& si, uint32_t g, int32_t* scores, uint32_t* states) {
    Rec32 rec;
    uint32_t slot;
    uint32_t pid = kNoPattern;
    if (find_node<kTypes>(t, r, text, si.b0, g, rec, slot)) {
        pid = rec.v[2];
        const uint32_t row = rec.v[3];
        if (row != kNoPattern) {
            const int32_t off = int32_t(rec.v[4]);
            const uint32_t len = rec.v[5];
            for (uint32_t k = 0; k < len; ++k) {
                const int64_t i = int64_t(g) + off + int64_t(k);
                if (i >= 0 && i < int64_t(si.nout)) atomicAdd(scores + si.obase + i, __ldg(t.pool + row + k));
            }
        }
    }
    if (states) states[si.cbase + g] = pid;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_score_general(DevModel m, BatchArgs a) {
    __shared__ Rings s_rings[kWarpsPerBlock];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t s = uint64_t(blockIdx.x) * kWarpsPerBlock + warp;
    if (s >= a.n_sent) return;
    Rings& r = s_rings[warp];
    const SentInfo si = sentence_info(a, s, lane);
    const uint8_t* __restrict__ text = a.text;
    const uint32_t n = si.n;
    uint32_t* cstates = a.char_states;
    uint32_t* tstates = a.type_states;
    if (si.status != 0) {
        for (uint32_t i = lane; i < si.nout; i += 32) { a.scores[si.obase + i] = 0; a.boundaries[si.obase + i] = 0; }
        if (cstates) for (uint32_t i = lane; i < n; i += 32) cstates[si.cbase + i] = kNoPattern;
        if (tstates) for (uint32_t i = lane; i < n; i += 32) tstates[si.cbase + i] = kNoPattern;
        return;
    }
    const int tw = m.type_cache_window;
    // pass 1: scores = bias + type table; states = none
    {
        uint64_t wpos = si.b0 & ~3ull;
        uint32_t nd = 0;
        for (uint32_t cb = 0; cb < n; cb += 32) {
            const uint32_t need = min(n, cb + 32u + uint32_t(tw));
            while (nd < need && wpos < si.b1) {
                nd += decode_window(text, wpos, si.b0, si.b1, nd, r, lane);
                wpos += 128;
            }
            __syncwarp();
            const uint32_t g = cb + lane;
            if (g + 1 < n) {
                int32_t v = m.bias;
                if (tw > 0) v += __ldg(m.type_cache + type_index(r, int64_t(g), n, tw));
                a.scores[si.obase + g] = v;
            }
            if (g < n) {
                if (cstates) cstates[si.cbase + g] = kNoPattern;
                if (tstates) tstates[si.cbase + g] = kNoPattern;
            }
            __syncwarp();
        }
    }
    __threadfence();
    __syncwarp();
    // pass 2: pattern rows
    if (m.ct.present || m.tt.present) {
        uint64_t wpos = si.b0 & ~3ull;
        uint32_t nd = 0;
        for (uint32_t cb = 0; cb < n; cb += 32) {
            const uint32_t need = min(n, cb + 32u);
            while (nd < need && wpos < si.b1) {
                nd += decode_window(text, wpos, si.b0, si.b1, nd, r, lane);
                wpos += 128;
            }
            __syncwarp();
            const uint32_t g = cb + lane;
            if (g < n) {
                if (m.ct.present)
                    scatter_general<false>(m.ct, r, text, si, g, a.scores, m.emit_states ? cstates : nullptr);
                if (m.tt.present)
                    scatter_general<true>(m.tt, r, text, si, g, a.scores, m.emit_states ? tstates : nullptr);
            }
            __syncwarp();
        }
    }
    __threadfence();
    __syncwarp();
    // pass 3: threshold
    for (uint32_t i = lane; i < si.nout; i += 32) {
        const int32_t v = __ldcg(a.scores + si.obase + i);
        a.boundaries[si.obase + i] = v > 0 ? 1 : 0;
    }
}

}  // namespace

cudaError_t launch_count(const BatchArgs& a, cudaStream_t stream) {
    if (a.n_sent == 0) return cudaSuccess;
    const uint64_t ngroups = (a.n_sent + kGroup - 1) / kGroup;
    k_count<<<unsigned(ngroups), kWarpsPerBlock * 32, 0, stream>>>(a);
    k_scan_groups<<<1, 1024, 0, stream>>>(a.group_bound, a.group_char, ngroups);
    return cudaGetLastError();
}

static bool use_fast(const DevModel& m) {
    return (!m.ct.present || m.ct.fast) && !m.tt.present && !m.emit_states;
}

cudaError_t launch_score(const DevModel& m, const BatchArgs& a, cudaStream_t stream) {
    if (a.n_sent == 0) return cudaSuccess;
    const uint64_t nblocks = (a.n_sent + kWarpsPerBlock - 1) / kWarpsPerBlock;
    if (use_fast(m)) k_score_fast<<<unsigned(nblocks), kWarpsPerBlock * 32, 0, stream>>>(m, a);
    else k_score_general<<<unsigned(nblocks), kWarpsPerBlock * 32, 0, stream>>>(m, a);
    return cudaGetLastError();
}

int launches_per_batch(const DevModel&) { return 3; }

}  // namespace vpt
