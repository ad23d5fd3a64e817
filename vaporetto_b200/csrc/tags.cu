// vaporetto_b200 — k_tags: Predictor::predict_tags on the device (reference predictor.rs:546-637).
//
// One warp per sentence walks its characters 32 at a time (the same window decoder as the sentence-warp scorer: it
// yields the byte position of every character).  A lane whose character ends a token (final boundary after it, or the
// last character) finds the token's first character by walking back over the boundaries, hashes the token's bytes and
// looks it up in the token table; for a known token it starts from the tag model's bias, adds the weight vectors keyed
// by (pattern id at characters last .. last + rel, token, rel) of both scorers -- the pattern's own vector plus those of
// its suffix patterns with the reference's truncation (tags.hpp) -- and takes the first strict maximum of every tag slot
// with at least two candidates (predictor.rs:286-304).  Integer adds wrap as in the reference's release build.
#include <algorithm>
#include <atomic>

#include "kernels_common.cuh"
#include "tags.hpp"

namespace vpt {

namespace {

constexpr int kTagWarps = 8;

// Calls f(byte) for every byte of the token -- of its KyteaFullwidthFilter image when norm != 0 (one character maps to
// one character, the byte length may change) -- and returns how many bytes that were.
template <typename F>
__device__ __forceinline__ uint32_t token_bytes(const uint8_t* __restrict__ bytes, uint32_t len, int norm, F f) {
    if (!norm) {
        for (uint32_t i = 0; i < len; ++i) f(uint32_t(__ldg(bytes + i)));
        return len;
    }
    uint32_t out = 0;
    for (uint32_t i = 0; i < len;) {
        const uint32_t b0 = __ldg(bytes + i);
        const uint32_t l = b0 < 0x80u ? 1u : b0 < 0xE0u ? 2u : b0 < 0xF0u ? 3u : 4u;
        uint32_t c = l == 1 ? b0 : b0 & (0x3Fu >> (l - 1));
        for (uint32_t k = 1; k < l && i + k < len; ++k) c = (c << 6) | (__ldg(bytes + i + k) & 0x3Fu);
        i += l;
        c = kytea_fullwidth(c);
        if (c < 0x80u) { f(c); out += 1; }
        else if (c < 0x800u) { f(0xC0u | (c >> 6)); f(0x80u | (c & 0x3Fu)); out += 2; }
        else if (c < 0x10000u) { f(0xE0u | (c >> 12)); f(0x80u | ((c >> 6) & 0x3Fu)); f(0x80u | (c & 0x3Fu)); out += 3; }
        else { f(0xF0u | (c >> 18)); f(0x80u | ((c >> 12) & 0x3Fu)); f(0x80u | ((c >> 6) & 0x3Fu)); f(0x80u | (c & 0x3Fu)); out += 4; }
    }
    return out;
}

__device__ __forceinline__ bool token_lookup(const DevTags& t, const uint8_t* __restrict__ bytes, uint32_t len, int norm, uint32_t& tid) {
    if (len == 0 || len > 4u * t.max_token_bytes) return false;
    uint64_t h = kTagHashInit;
    const uint32_t nlen = token_bytes(bytes, len, norm, [&](uint32_t b) { h = tag_hash_step(h, b); });
    if (nlen > t.max_token_bytes) return false;
    h = tag_hash_finish(h);
    for (uint32_t s = uint32_t(h >> 20) & t.tok_mask;; s = (s + 1) & t.tok_mask) {
        const TagTokenEntry e = t.tok_tab[s];
        if (e.hash == 0) return false;
        if (e.hash == h && e.len == nlen) {
            bool same = true;
            uint32_t idx = 0;
            token_bytes(bytes, len, norm, [&](uint32_t b) { same = same && __ldg(t.tok_bytes + e.str_off + idx) == b; ++idx; });
            if (same) { tid = e.tid; return true; }
        }
    }
}

constexpr int kTagRelRegs = 4;  // rel positions whose chains are kept in registers (window 3: rel 0..3)

// true when pattern `want` lies on the suffix chain that starts at `pid` (positions 0..4 from the chain record, the rest
// through the link table: only patterns longer than five suffix levels, i.e. dictionary words, get there)
__device__ __forceinline__ bool on_chain(uint32_t want, uint32_t pid, const uint4 ch, const uint32_t* __restrict__ link) {
    if (want == pid || want == ch.x || want == ch.y || want == ch.z || want == ch.w) return true;
    if (ch.w == kNoPattern) return false;
    for (uint32_t q = __ldg(link + ch.w); q != kNoPattern; q = __ldg(link + q))
        if (q == want) return true;
    return false;
}

// scores += the tag weights of one scorer for the token whose last character is `i` (add_tag_scores,
// char_scorer/boundary_tag_scorer.rs:154-174, type_scorer/boundary_tag_scorer.rs:123-143): for every rel position the
// pattern found at character i + rel contributes its own vector and those of its suffix patterns (merged in the
// reference at build time), element k of a shorter pattern's vector only while every longer one on the chain is longer
// than k.  The token's key list is sorted by rel, longest pattern first: one pass per rel position applies the rule.
// cnt[r]: entries with rel position r (0 .. 3); rest: entries with rel position >= 4 behind them.
__device__ __forceinline__ void add_scorer(const TagKey* __restrict__ keys, const uint8_t* cnt, uint32_t rest,
                                           const TagChain* __restrict__ chains, const uint32_t* __restrict__ link,
                                           const int32_t* __restrict__ pool, const uint32_t* __restrict__ states, uint32_t npat,
                                           uint32_t rels, uint32_t i, uint32_t n, int32_t* scores, uint32_t nscores) {
    // the patterns at the rel positions and their chains: all loads first (they are independent)
    uint32_t pid[kTagRelRegs];
    uint4 ch[kTagRelRegs];
#pragma unroll
    for (int r = 0; r < kTagRelRegs; ++r) {
        pid[r] = kNoPattern;
        if (cnt[r] != 0 && uint32_t(r) < rels && i + uint32_t(r) < n) pid[r] = states[i + r];
        if (pid[r] >= npat) pid[r] = kNoPattern;
    }
#pragma unroll
    for (int r = 0; r < kTagRelRegs; ++r) {
        ch[r] = make_uint4(kNoPattern, kNoPattern, kNoPattern, kNoPattern);
        if (pid[r] != kNoPattern) ch[r] = __ldg(reinterpret_cast<const uint4*>(chains + pid[r]));
    }
#pragma unroll
    for (int r = 0; r < kTagRelRegs; ++r) {
        const uint32_t nk = cnt[r];
        if (pid[r] != kNoPattern) {
            uint32_t limit = nscores;
            for (uint32_t j = 0; j < nk && limit; ++j) {
                const uint4 e = __ldg(reinterpret_cast<const uint4*>(keys + j));  // pid, off, len, rel
                if (on_chain(e.x, pid[r], ch[r], link)) {
                    const uint32_t upto = min(limit, e.z);
                    for (uint32_t k = 0; k < upto; ++k) scores[k] = int32_t(uint32_t(scores[k]) + uint32_t(__ldg(pool + e.y + k)));
                    limit = upto;
                }
            }
        }
        keys += nk;
    }
    // wider windows: rel positions from 4 on, the chain straight from the tables
    uint32_t cur_rel = 0xFFFFFFFFu, limit = 0;
    for (uint32_t j = 0; j < rest; ++j) {
        const uint4 e = __ldg(reinterpret_cast<const uint4*>(keys + j));
        if (e.w != cur_rel) { cur_rel = e.w; limit = nscores; }
        if (limit == 0 || e.w >= rels || i + e.w >= n) continue;
        const uint32_t q = states[i + e.w];
        if (q < npat && on_chain(e.x, q, __ldg(reinterpret_cast<const uint4*>(chains + q)), link)) {
            const uint32_t upto = min(limit, e.z);
            for (uint32_t k = 0; k < upto; ++k) scores[k] = int32_t(uint32_t(scores[k]) + uint32_t(__ldg(pool + e.y + k)));
            limit = upto;
        }
    }
}

// Tag prediction of one token (bytes [bytes, bytes + len) of the text; `i` = index of its last character inside the
// sentence's `n` characters whose pattern-id states start at cst / tst): token lookup, bias + tag weights of both scorers,
// first strict maximum per tag slot (TagPredictor::predict, predictor.rs:286-304).  Returns the token id or -1 and the
// chosen candidates in cand[].
// (the part behind the token lookup: `tid` is a token of the table)
__device__ __forceinline__ int32_t tag_score_token(const DevTags& t, uint32_t tid, const uint32_t* __restrict__ cst,
                                                   const uint32_t* __restrict__ tst, uint32_t i, uint32_t n, int32_t* cand,
                                                   uint32_t* n_unserved) {
    TagTokenInfo ti;
    {
        const uint4* q = reinterpret_cast<const uint4*>(t.tok_info + tid);
        *reinterpret_cast<uint4*>(&ti) = __ldg(q);
        *(reinterpret_cast<uint4*>(&ti) + 1) = __ldg(q + 1);
    }
    if (!ti.usable) {
        if (n_unserved) atomicAdd(n_unserved, 1u);
        return -1;
    }
    int32_t scores[kTagMaxScores];
    const uint32_t ns = ti.bias_len;
    for (uint32_t k = 0; k < ns; ++k) scores[k] = __ldg(t.pool + ti.bias_off + k);
    const uint32_t n_ckeys = uint32_t(ti.ckeys[0]) + ti.ckeys[1] + ti.ckeys[2] + ti.ckeys[3] + ti.c_rest;
    const uint32_t n_tkeys = uint32_t(ti.tkeys[0]) + ti.tkeys[1] + ti.tkeys[2] + ti.tkeys[3] + ti.t_rest;
    if (cst && n_ckeys)
        add_scorer(t.keys + ti.key_off, ti.ckeys, ti.c_rest, t.c_chain, t.c_link, t.pool, cst, t.n_char_patterns, t.char_rels, i, n,
                   scores, ns);
    if (tst && n_tkeys)
        add_scorer(t.keys + ti.key_off + n_ckeys, ti.tkeys, ti.t_rest, t.t_chain, t.t_link, t.pool, tst, t.n_type_patterns,
                   t.type_rels, i, n, scores, ns);
    uint32_t off = 0;
    const uint32_t nt = t.n_tags;
    for (uint32_t k = 0; k < ti.n_slots && k < nt; ++k) {
        const uint32_t nc = ti.cand[k];
        if (nc >= 2) {
            if (off + nc > ns) {
                if (n_unserved) atomicAdd(n_unserved, 1u);  // the host path reports the model error
                return -1;
            }
            uint32_t best = 0;
            int32_t mx = INT32_MIN;
            for (uint32_t c = 0; c < nc; ++c)
                if (scores[off + c] > mx) { best = c; mx = scores[off + c]; }
            cand[k] = int32_t(best);
            off += nc;
        } else {
            cand[k] = nc == 1 ? 0 : -1;
        }
    }
    return int32_t(tid);
}

__device__ __forceinline__ int32_t tag_token_at(const DevTags& t, const uint8_t* __restrict__ bytes, uint32_t len,
                                                const uint32_t* __restrict__ cst, const uint32_t* __restrict__ tst, uint32_t i,
                                                uint32_t n, int32_t* cand, uint32_t* n_unserved, int norm) {
    uint32_t tid = 0;
    if (!token_lookup(t, bytes, len, norm, tid)) return -1;
    return tag_score_token(t, tid, cst, tst, i, n, cand, n_unserved);
}

// kLocateOnly: phase 1 of the per-token path (descriptors for k_tok_lookup / k_tok_score); otherwise the kernel predicts the tags itself.
template <bool kLocateOnly>
__global__ void __launch_bounds__(kTagWarps * 32, kLocateOnly ? 8 : 1) k_tags(DevTags t, TagArgs a) {
    __shared__ Rings s_rings[kTagWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // (the work list of the per-token kernels starts empty: they run behind this kernel on the stream)
    if (kLocateOnly && a.tok_work && blockIdx.x == 0 && threadIdx.x == 0) a.tok_work[0] = 0;
    const uint64_t s = uint64_t(blockIdx.x) * kTagWarps + warp;
    if (s >= a.n_sent) return;
    Rings& r = s_rings[warp];
    const uint64_t b0 = a.offsets[s], b1 = a.offsets[s + 1] - (a.trims ? a.trims[s] : 0);
    const uint64_t cb = a.char_offsets[s] - a.char_base;
    const uint32_t n = uint32_t(a.char_offsets[s + 1] - a.char_offsets[s]);
    const uint64_t bo = a.bound_offsets[s] - a.bound_base;
    const uint32_t nt = t.n_tags;
    if (a.status[s] != 0) {
        if (a.tok_base) return;  // a rejected sentence has no tokens
        for (uint32_t i = lane; i < n; i += 32) {
            a.tag_token[cb + i] = -1;
            for (uint32_t k = 0; k < nt; ++k) a.tag_cand[(cb + i) * nt + k] = -1;
        }
        return;
    }
    uint64_t wpos = b0 & ~3ull;
    uint32_t nd = 0;
    uint32_t tok_rank = 0;  // tokens that end before this chunk of characters
    uint32_t chunk_start = 0;  // first character of the token that is open at the start of the chunk
    for (uint32_t c0 = 0; c0 < n; c0 += 32) {
        // byte positions of the characters up to c0 + 32 (one more: the end of the chunk's last character)
        const uint32_t need = min(n, c0 + 33u);
        while (nd < need && wpos < b1) {
            nd += decode_window(a.text, wpos, b0, b1, nd, r, lane, false);
            wpos += 128;
        }
        __syncwarp();
        const uint32_t i = c0 + lane;
        int32_t tok = -1;
        int32_t cand[kTagMaxSlots];
#pragma unroll
        for (int k = 0; k < kTagMaxSlots; ++k) cand[k] = -1;
        const bool ends = i < n && (i + 1 == n || a.boundaries[bo + i] == 1);
        const unsigned endm_all = __ballot_sync(kFull, ends);
        uint32_t desc_x = 0, desc_y = 0, desc_z = 0, desc_w = 0;
        if (i < n) {
            if (ends) {
                // the token starts behind the previous token end: in this chunk (ballot) or before it (carried)
                const unsigned before = endm_all & ((1u << lane) - 1u);
                const uint32_t start = before ? c0 + 32u - uint32_t(__clz(before)) : chunk_start;
                const bool near = i - start < uint32_t(kRing - 40);  // the ring still holds the token's first character
                const uint32_t sb = near ? r.bp[start & kRingMask] : 0u;
                const uint32_t eb = i + 1 < n ? r.bp[(i + 1) & kRingMask] : uint32_t(b1 - b0);
                if (kLocateOnly) {
                    // phase 1 of the per-token path: where the token is; k_tok_lookup and k_tok_score do the rest, one thread per token
                    desc_x = near ? uint32_t(b0 + sb - a.text_base) : 0u;
                    desc_y = uint32_t((b0 + sb - a.text_base) >> 32);
                    desc_z = uint32_t(cb + i);
                    desc_w = (near ? min(eb - sb, 0xFFFFu) : 0u) | (min(n - i, 0xFFFFu) << 16);
                    if (!near && a.n_unserved) atomicAdd(a.n_unserved, 1u);
                } else if (near) {
                    if (!kLocateOnly) tok = tag_token_at(t, a.text + b0 + sb, eb - sb, a.char_states ? a.char_states + cb : nullptr,
                                       a.type_states ? a.type_states + cb : nullptr, i, n, cand, a.n_unserved, a.norm);
                } else if (a.n_unserved) {
                    atomicAdd(a.n_unserved, 1u);  // a token longer than the ring: left to the host path
                }
            }
        }
        __syncwarp();
        if (endm_all) chunk_start = c0 + 32u - uint32_t(__clz(endm_all));
        if (a.tok_base) {
            // per-token records: the token's rank is the number of boundaries before its last character
            const unsigned endm = endm_all;
            if (ends) {
                const uint64_t rec = a.tok_base[s] + tok_rank + __popc(endm & ((1u << lane) - 1u));
                if (kLocateOnly) {
                    a.tok_desc[rec] = make_uint4(desc_x, desc_y, desc_z, desc_w);
                } else {
                    a.tok_ids[rec] = tok;
                    for (uint32_t k = 0; k < nt; ++k) a.tok_cands[rec * nt + k] = (tok >= 0 && cand[k] >= 0) ? uint8_t(cand[k]) : uint8_t(255);
                }
            }
            tok_rank += __popc(endm);
        } else if (i < n) {
            a.tag_token[cb + i] = tok;
            for (uint32_t k = 0; k < nt; ++k) a.tag_cand[(cb + i) * nt + k] = tok >= 0 ? cand[k] : -1;
        }
        __syncwarp();
    }
}

// Phase 2 of the per-token path, one thread per token record (descriptor written by k_tags): consecutive lanes read
// consecutive text.  Two kernels, because only some tokens have a tag model (41 % on the config-3 workload) and the work
// behind the lookup is 5-10 x the lookup: with both in one kernel a warp runs the long path at the lane density of the
// tokens that have a model.
//   k_tok_lookup  every token: hash + table lookup; the token id (or -1) is final here; tokens with a model are appended
//                 to a work list (one atomic per warp)
//   k_tok_score   the listed tokens, dense lanes: bias + tag weights + arg-max
// Both are grid-stride loops over a resident grid (the token count is known on the device only).
constexpr int kTokTagThreads = 128;
constexpr int kTokWorkHead = 4;  // words in front of the work list: [0] = entries
__global__ void __launch_bounds__(kTokTagThreads) k_tok_lookup(DevTags t, TagArgs a) {
    const uint64_t ntok = a.tok_base[a.n_sent];
    const uint32_t nt = t.n_tags;
    const int lane = threadIdx.x & 31;
    const uint64_t stride = uint64_t(gridDim.x) * kTokTagThreads;
    // (warp-uniform trip count: the list append is a warp operation)
    for (uint64_t rec0 = uint64_t(blockIdx.x) * kTokTagThreads + (threadIdx.x & ~31u); rec0 < ntok; rec0 += stride) {
        const uint64_t rec = rec0 + uint32_t(lane);
        bool listed = false;
        if (rec < ntok) {
            const uint4 d = a.tok_desc[rec];
            const uint32_t len = d.w & 0xFFFFu;
            int32_t tok = -1;
            uint32_t tid = 0;
            if (len && token_lookup(t, a.text + a.text_base + ((uint64_t(d.y) << 32) | d.x), len, a.norm, tid)) {
                if (__ldg(&t.tok_info[tid].usable)) {
                    tok = int32_t(tid);
                    listed = true;
                } else if (a.n_unserved) {
                    atomicAdd(a.n_unserved, 1u);
                }
            }
            a.tok_ids[rec] = tok;
            if (!listed)
                for (uint32_t k = 0; k < nt; ++k) a.tok_cands[rec * nt + k] = uint8_t(255);
        }
        const unsigned m = __ballot_sync(kFull, listed);
        if (m) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(a.tok_work, uint32_t(__popc(m)));
            base = __shfl_sync(kFull, base, 0);
            if (listed) a.tok_work[kTokWorkHead + base + __popc(m & ((1u << lane) - 1u))] = uint32_t(rec);
        }
    }
}

__global__ void __launch_bounds__(kTokTagThreads) k_tok_score(DevTags t, TagArgs a) {
    const uint32_t nw = a.tok_work[0];
    const uint32_t nt = t.n_tags;
    for (uint32_t w = blockIdx.x * kTokTagThreads + threadIdx.x; w < nw; w += gridDim.x * kTokTagThreads) {
        const uint32_t rec = a.tok_work[kTokWorkHead + w];
        const uint4 d = a.tok_desc[rec];
        int32_t cand[kTagMaxSlots];
#pragma unroll
        for (int k = 0; k < kTagMaxSlots; ++k) cand[k] = -1;
        // (characters [d.z, d.z + n_after) are the token's last character and what follows it in its sentence)
        const int32_t tok = tag_score_token(t, uint32_t(a.tok_ids[rec]), a.char_states ? a.char_states + d.z : nullptr,
                                            a.type_states ? a.type_states + d.z : nullptr, 0, d.w >> 16, cand, a.n_unserved);
        if (tok < 0) a.tok_ids[rec] = -1;
        for (uint32_t k = 0; k < nt; ++k) a.tok_cands[uint64_t(rec) * nt + k] = (tok >= 0 && cand[k] >= 0) ? uint8_t(cand[k]) : uint8_t(255);
    }
}

// ---- compact outputs -----------------------------------------------------------------------------------------------------

constexpr int kPackThreads = 256;

// boundaries (one byte each) -> one bit each; thread per output word
__global__ void __launch_bounds__(kPackThreads) k_pack_bits(CompactArgs c) {
    const uint64_t w = uint64_t(blockIdx.x) * kPackThreads + threadIdx.x;
    const uint64_t nwords = (uint64_t(c.bit_base) + c.n_bound + 31) / 32;
    if (w >= nwords) return;
    // word w holds the boundaries [32 w - bit_base, 32 w - bit_base + 32)
    const int64_t first = int64_t(32 * w) - int64_t(c.bit_base);
    uint32_t v = 0;
    if (first >= 0 && uint64_t(first) + 32 <= c.n_bound && ((reinterpret_cast<uintptr_t>(c.boundaries) + first) & 3) == 0) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(c.boundaries + first);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t x = __ldg(q + j) & 0x01010101u;  // bytes are 0 / 1
            v |= (((x * 0x01020408u) >> 24) & 15u) << (4 * j);
        }
    } else {
        for (int j = 0; j < 32; ++j) {
            const int64_t i = first + j;
            if (i >= 0 && uint64_t(i) < c.n_bound && c.boundaries[i]) v |= 1u << j;
        }
    }
    c.bits[w] = v;
}

// status as one byte, tokens per sentence (boundaries set + 1 for a scored sentence), their prefix inside the block of
// 256 sentences and the block's total
__global__ void __launch_bounds__(kPackThreads) k_sentence_info(CompactArgs c) {
    __shared__ uint32_t s_w[kPackThreads / 32];
    const uint64_t s = uint64_t(blockIdx.x) * kPackThreads + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t ntok = 0;
    if (s < c.n_sent) {
        const int32_t st = c.status[s];
        c.status8[s] = uint8_t(st);
        if (c.n_tokens) {
            const uint32_t n = c.n_chars[s];
            if (st == 0 && n > 0) {
                const uint8_t* b = c.boundaries + (c.bound_offsets[s] - c.bound_base);
                ntok = 1;
                for (uint32_t i = 0; i + 1 < n; ++i) ntok += b[i];
            }
            c.n_tokens[s] = ntok;
        }
    }
    if (!c.n_tokens) return;
    const uint32_t incl = warp_incl_scan(ntok, lane);
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kPackThreads / 32; ++w) {
        if (w < warp) base += s_w[w];
        tot += s_w[w];
    }
    if (s < c.n_sent) c.tok_local[s] = base + incl - ntok;
    if (threadIdx.x == 0) c.tok_blk[blockIdx.x] = tot;
}

// exclusive prefix of the block totals (one block; 1024 totals per round), the grand total behind them
__global__ void __launch_bounds__(1024) k_token_scan(CompactArgs c) {
    __shared__ uint64_t s_w[32];
    __shared__ uint64_t s_carry;
    const uint64_t nblk = (c.n_sent + kPackThreads - 1) / kPackThreads;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint64_t lo = 0; lo < nblk; lo += 1024) {
        const uint64_t i = lo + threadIdx.x;
        const uint64_t v = i < nblk ? c.tok_blk[i] : 0;
        uint64_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t o = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        uint64_t base = s_carry;
        for (int w = 0; w < warp; ++w) base += s_w[w];
        if (i < nblk) c.tok_blk[i] = base + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = base + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        c.tok_blk[nblk] = s_carry;
        c.tok_base[c.n_sent] = s_carry;
        if (c.tok_total_host) *c.tok_total_host = s_carry;
    }
}

__global__ void __launch_bounds__(kPackThreads) k_token_base(CompactArgs c) {
    const uint64_t s = uint64_t(blockIdx.x) * kPackThreads + threadIdx.x;
    if (s < c.n_sent) c.tok_base[s] = c.tok_blk[blockIdx.x] + c.tok_local[s];
}

}  // namespace

cudaError_t launch_compact(const CompactArgs& c, cudaStream_t stream) {
    const uint64_t nwords = (uint64_t(c.bit_base) + c.n_bound + 31) / 32;
    if (nwords) k_pack_bits<<<unsigned((nwords + kPackThreads - 1) / kPackThreads), kPackThreads, 0, stream>>>(c);
    if (c.n_sent) {
        k_sentence_info<<<unsigned((c.n_sent + kPackThreads - 1) / kPackThreads), kPackThreads, 0, stream>>>(c);
        if (c.n_tokens) {
            k_token_scan<<<1, 1024, 0, stream>>>(c);
            k_token_base<<<unsigned((c.n_sent + kPackThreads - 1) / kPackThreads), kPackThreads, 0, stream>>>(c);
        }
    }
    return cudaGetLastError();
}

cudaError_t launch_tags(const DevTags& t, const TagArgs& a, cudaStream_t stream) {
    if (a.n_sent == 0) return cudaSuccess;
    const uint64_t nblocks = (a.n_sent + kTagWarps - 1) / kTagWarps;
    if (a.tok_desc && a.tok_base) k_tags<true><<<unsigned(nblocks), kTagWarps * 32, 0, stream>>>(t, a);
    else k_tags<false><<<unsigned(nblocks), kTagWarps * 32, 0, stream>>>(t, a);
    if (a.tok_desc && a.tok_base && a.max_tokens) {
        // (the number of tokens is known on the device only: resident grids, grid-stride loops)
        if (!a.tok_work || a.max_tokens > 0xFFFFFFFFull) return cudaErrorInvalidValue;
        static std::atomic<int> sm_count[64] = {};  // (written with the same value by whoever comes first)
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
        if (sm_count[dev] == 0) {
            int v = 0;
            e = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
            if (e != cudaSuccess) return e;
            sm_count[dev] = v;
        }
        const uint64_t want = (a.max_tokens + kTokTagThreads - 1) / kTokTagThreads;
        // (resident blocks per SM: 16 x 128 threads at 32 registers, 8 x 128 at 56)
        k_tok_lookup<<<unsigned(std::min<uint64_t>(want, uint64_t(sm_count[dev]) * 16)), kTokTagThreads, 0, stream>>>(t, a);
        k_tok_score<<<unsigned(std::min<uint64_t>(want, uint64_t(sm_count[dev]) * 8)), kTokTagThreads, 0, stream>>>(t, a);
    }
    return cudaGetLastError();
}

}  // namespace vpt
