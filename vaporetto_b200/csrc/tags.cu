// vaporetto_b200 — k_tags: Predictor::predict_tags on the device (reference predictor.rs:546-637).
//
// One warp per sentence walks its characters 32 at a time (the same window decoder as the sentence-warp scorer: it
// yields the byte position of every character).  A lane whose character ends a token (final boundary after it, or the
// last character) finds the token's first character by walking back over the boundaries, hashes the token's bytes and
// looks it up in the token table; for a known token it starts from the tag model's bias, adds the weight vectors keyed
// by (pattern id at characters last .. last + rel, token, rel) of both scorers -- the pattern's own vector plus those of
// its suffix patterns with the reference's truncation (tags.hpp) -- and takes the first strict maximum of every tag slot
// with at least two candidates (predictor.rs:286-304).  Integer adds wrap as in the reference's release build.
#include "kernels_common.cuh"
#include "tags.hpp"

namespace vpt {

namespace {

constexpr int kTagWarps = 8;

__device__ __forceinline__ bool token_lookup(const DevTags& t, const uint8_t* __restrict__ bytes, uint32_t len, uint32_t& tid) {
    if (len == 0 || len > t.max_token_bytes) return false;
    uint64_t h = kTagHashInit;
    for (uint32_t i = 0; i < len; ++i) h = tag_hash_step(h, __ldg(bytes + i));
    h = tag_hash_finish(h);
    for (uint32_t s = uint32_t(h >> 20) & t.tok_mask;; s = (s + 1) & t.tok_mask) {
        const TagTokenEntry e = t.tok_tab[s];
        if (e.hash == 0) return false;
        if (e.hash == h && e.len == len) {
            bool same = true;
            for (uint32_t i = 0; i < len && same; ++i) same = __ldg(t.tok_bytes + e.str_off + i) == __ldg(bytes + i);
            if (same) { tid = e.tid; return true; }
        }
    }
}

// scores += the merged weight of pattern `pid` for (token, rel): own vector plus the suffix chain's, truncated
__device__ __forceinline__ void add_chain(const TagWeightSlot* __restrict__ tab, uint32_t mask, const uint32_t* __restrict__ link,
                                          const int32_t* __restrict__ pool, uint32_t pid, uint32_t tid, uint32_t rel,
                                          int32_t* scores, uint32_t nscores) {
    uint32_t limit = nscores;
    bool first = true;
    for (uint32_t q = pid; q != kNoPattern && limit > 0; q = __ldg(link + q)) {
        const uint64_t key = tag_weight_key(q, tid, rel);
        for (uint32_t s = tag_weight_slot(key, mask);; s = (s + 1) & mask) {
            const TagWeightSlot e = tab[s];
            if (e.key == 0) break;
            if (e.key == key) {
                // element k of a suffix's vector counts only while every longer pattern's own vector is longer than k
                // (the first vector found fixes the length of the merged vector)
                const uint32_t upto = min(limit, e.len);
                for (uint32_t k = 0; k < upto; ++k) scores[k] = int32_t(uint32_t(scores[k]) + uint32_t(__ldg(pool + e.off + k)));
                limit = upto;
                first = false;
                break;
            }
        }
    }
    (void)first;
}

__global__ void __launch_bounds__(kTagWarps * 32) k_tags(DevTags t, TagArgs a) {
    __shared__ Rings s_rings[kTagWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
        if (i < n) {
            if (ends) {
                uint32_t start = i;
                while (start > 0 && a.boundaries[bo + start - 1] == 0) --start;
                uint32_t tid = 0;
                const bool near = i - start < uint32_t(kRing - 40);  // the ring still holds the token's first character
                const uint32_t sb = near ? r.bp[start & kRingMask] : 0u;
                const uint32_t eb = i + 1 < n ? r.bp[(i + 1) & kRingMask] : uint32_t(b1 - b0);
                if (near && token_lookup(t, a.text + b0 + sb, eb - sb, tid)) {
                    const TagTokenInfo ti = t.tok_info[tid];
                    if (!ti.usable) {
                        if (a.n_unserved) atomicAdd(a.n_unserved, 1u);
                    } else {
                        int32_t scores[kTagMaxScores];
                        const uint32_t ns = ti.bias_len;
                        for (uint32_t k = 0; k < ns; ++k) scores[k] = __ldg(t.pool + ti.bias_off + k);
                        if (a.char_states)
                            for (uint32_t rel = 0; rel < t.char_rels && i + rel < n; ++rel) {
                                const uint32_t pid = a.char_states[cb + i + rel];
                                if (pid < t.n_char_patterns && t.c_any[pid])
                                    add_chain(t.cw_tab, t.cw_mask, t.c_link, t.pool, pid, tid, rel, scores, ns);
                            }
                        if (a.type_states)
                            for (uint32_t rel = 0; rel < t.type_rels && i + rel < n; ++rel) {
                                const uint32_t pid = a.type_states[cb + i + rel];
                                if (pid < t.n_type_patterns && t.t_any[pid])
                                    add_chain(t.tw_tab, t.tw_mask, t.t_link, t.pool, pid, tid, rel, scores, ns);
                            }
                        // TagPredictor::predict: first strict maximum per slot with >= 2 candidates
                        uint32_t off = 0;
                        bool ok = true;
                        for (uint32_t k = 0; k < ti.n_slots && k < nt; ++k) {
                            const uint32_t nc = ti.cand[k];
                            if (nc >= 2) {
                                if (off + nc > ns) { ok = false; break; }
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
                        if (ok) tok = int32_t(tid);
                        else if (a.n_unserved) atomicAdd(a.n_unserved, 1u);  // the host path reports the model error
                    }
                } else if (!near && a.n_unserved) {
                    atomicAdd(a.n_unserved, 1u);  // a token longer than the ring: left to the host path
                }
            }
        }
        __syncwarp();
        if (a.tok_base) {
            // per-token records: the token's rank is the number of boundaries before its last character
            const unsigned endm = __ballot_sync(kFull, ends);
            if (ends) {
                const uint64_t rec = a.tok_base[s] + tok_rank + __popc(endm & ((1u << lane) - 1u));
                a.tok_ids[rec] = tok;
                for (uint32_t k = 0; k < nt; ++k) a.tok_cands[rec * nt + k] = (tok >= 0 && cand[k] >= 0) ? uint8_t(cand[k]) : uint8_t(255);
            }
            tok_rank += __popc(endm);
        } else if (i < n) {
            a.tag_token[cb + i] = tok;
            for (uint32_t k = 0; k < nt; ++k) a.tag_cand[(cb + i) * nt + k] = tok >= 0 ? cand[k] : -1;
        }
        __syncwarp();
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

// status as one byte, tokens per sentence (boundaries set + 1 for a scored sentence)
__global__ void __launch_bounds__(kPackThreads) k_sentence_info(CompactArgs c) {
    const uint64_t s = uint64_t(blockIdx.x) * kPackThreads + threadIdx.x;
    if (s >= c.n_sent) return;
    const int32_t st = c.status[s];
    c.status8[s] = uint8_t(st);
    if (!c.n_tokens) return;
    uint32_t ntok = 0;
    const uint32_t n = c.n_chars[s];
    if (st == 0 && n > 0) {
        const uint8_t* b = c.boundaries + (c.bound_offsets[s] - c.bound_base);
        ntok = 1;
        for (uint32_t i = 0; i + 1 < n; ++i) ntok += b[i];
    }
    c.n_tokens[s] = ntok;
}

// exclusive prefix of n_tokens (one block; a chunk has at most a few hundred thousand sentences)
__global__ void __launch_bounds__(1024) k_token_scan(CompactArgs c) {
    __shared__ uint64_t s_part[1024];
    const uint64_t per = (c.n_sent + 1023) / 1024;
    const uint64_t lo = min(c.n_sent, per * threadIdx.x), hi = min(c.n_sent, lo + per);
    uint64_t sum = 0;
    for (uint64_t i = lo; i < hi; ++i) sum += c.n_tokens[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    // Hillis-Steele over the 1024 partial sums
    for (int d = 1; d < 1024; d <<= 1) {
        const uint64_t v = threadIdx.x >= unsigned(d) ? s_part[threadIdx.x - d] : 0;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    uint64_t run = s_part[threadIdx.x] - sum;
    for (uint64_t i = lo; i < hi; ++i) { c.tok_base[i] = run; run += c.n_tokens[i]; }
    if (threadIdx.x == 1023) {
        c.tok_base[c.n_sent] = s_part[1023];
        if (c.tok_total_host) *c.tok_total_host = s_part[1023];
    }
}

}  // namespace

cudaError_t launch_compact(const CompactArgs& c, cudaStream_t stream) {
    const uint64_t nwords = (uint64_t(c.bit_base) + c.n_bound + 31) / 32;
    if (nwords) k_pack_bits<<<unsigned((nwords + kPackThreads - 1) / kPackThreads), kPackThreads, 0, stream>>>(c);
    if (c.n_sent) {
        k_sentence_info<<<unsigned((c.n_sent + kPackThreads - 1) / kPackThreads), kPackThreads, 0, stream>>>(c);
        if (c.n_tokens) k_token_scan<<<1, 1024, 0, stream>>>(c);
    }
    return cudaGetLastError();
}

cudaError_t launch_tags(const DevTags& t, const TagArgs& a, cudaStream_t stream) {
    if (a.n_sent == 0) return cudaSuccess;
    const uint64_t nblocks = (a.n_sent + kTagWarps - 1) / kTagWarps;
    k_tags<<<unsigned(nblocks), kTagWarps * 32, 0, stream>>>(t, a);
    return cudaGetLastError();
}

}  // namespace vpt
