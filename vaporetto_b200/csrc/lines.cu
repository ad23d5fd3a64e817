// lines.cu — the callers either side of the scoring path, on the device:
//
//   * line splitting: raw file bytes -> sentence byte ranges, with the semantics of Rust's `BufRead::lines`
//     that the reference CLI loops over (predict/src/main.rs:126-130): a line ends at '\n', a '\r' directly
//     before it is dropped, a final line without '\n' still counts, a trailing '\n' adds no empty line;
//   * tokenised output: `Sentence::write_tokenized_text` (sentence.rs:850-886) for sentences without tags —
//     a ' ' between tokens, a '\' before each ' ', '\' and '/' of the surface — followed by the '\n' the CLI
//     writes after every line (predict/src/main.rs:140,149); rejected lines produce the bare '\n'.
//
// Both are byte-streaming work: SWAR byte tests on 32-bit words, warp prefix sums, no tables.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "device_model.hpp"

namespace vpt {

namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(kFull, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// bit 7 of every byte of x that is zero (exact per byte: no borrow between bytes)
__device__ __forceinline__ uint32_t zero_bytes(uint32_t x) {
    return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;
}
__device__ __forceinline__ uint32_t eq_bytes(uint32_t x, uint32_t c) { return zero_bytes(x ^ (c * 0x01010101u)); }

// bit 7 of the bytes of the word at `addr` that lie inside [b0, b1) (requires addr < b1, addr + 4 > b0)
__device__ __forceinline__ uint32_t inside80(uint32_t addr, uint32_t b0, uint32_t b1) {
    const uint32_t from = b0 > addr ? b0 - addr : 0u;
    const uint32_t to = b1 - addr < 4u ? b1 - addr : 4u;
    return (from >= 4u ? 0u : 0x80808080u << (8 * from)) & (0x80808080u >> (8 * (4 - to)));
}

// ------------------------------------------------------------------------------------------------
// line splitting
// ------------------------------------------------------------------------------------------------
constexpr int kSplitThreads = 256;
constexpr int kSplitBytesPerThread = kSplitBlockBytes / kSplitThreads;  // 32
static_assert(kSplitBytesPerThread == 32, "two 16-byte loads per thread");

// '\n' bytes of the thread's 32 bytes as a bit mask (bit i = byte i), bytes at or beyond n_bytes excluded
__device__ __forceinline__ uint32_t newline_mask(const uint8_t* __restrict__ text, uint64_t n_bytes, uint64_t pos) {
    uint32_t mask = 0;
    if (pos >= n_bytes) return 0;
    const uint4* p = reinterpret_cast<const uint4*>(text + pos);
    const uint4 v0 = __ldg(p);
    const uint4 v1 = pos + 16 < n_bytes ? __ldg(p + 1) : make_uint4(0, 0, 0, 0);
    const uint32_t w[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t z = eq_bytes(w[i], 0x0Au);
        mask |= (((z >> 7) | (z >> 14) | (z >> 21) | (z >> 28)) & 15u) << (4 * i);
    }
    const uint64_t left = n_bytes - pos;
    if (left < 32) mask &= (1u << left) - 1u;
    return mask;
}

// pass 1: number of '\n' in every 8 KB block
__global__ void __launch_bounds__(kSplitThreads) k_nl_count(SplitArgs s) {
    __shared__ uint32_t s_w[kSplitThreads / 32];
    const uint64_t pos = uint64_t(blockIdx.x) * kSplitBlockBytes + uint64_t(threadIdx.x) * kSplitBytesPerThread;
    const uint32_t c = __popc(newline_mask(s.text, s.n_bytes, pos));
    const uint32_t w = __reduce_add_sync(kFull, c);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int i = 0; i < kSplitThreads / 32; ++i) t += s_w[i];
        s.blk[blockIdx.x] = t;
    }
}

// pass 2 (one CTA): exclusive scan of the block counts; *n_lines = newlines (+1 for an unterminated last line)
__global__ void __launch_bounds__(1024) k_nl_scan(SplitArgs s, uint64_t nblk) {
    __shared__ uint64_t s_w[32];
    __shared__ uint64_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base < nblk; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const uint32_t v = i < nblk ? s.blk[i] : 0u;
        const uint32_t incl = warp_incl_scan_u32(v, lane);
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint64_t w = s_w[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t o = __shfl_up_sync(kFull, w, d);
                if (lane >= d) w += o;
            }
            s_w[lane] = w;
        }
        __syncthreads();
        const uint64_t excl = s_carry + (warp ? s_w[warp - 1] : 0) + incl - v;
        if (i < nblk) s.blk_base[i] = excl;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += s_w[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const bool open_tail = s.n_bytes > 0 && s.text[s.n_bytes - 1] != 0x0A;
        *s.n_lines = s_carry + (open_tail ? 1 : 0);
    }
}

// pass 3: line l ends at its '\n' (exclusive): offsets[l + 1] = position after it, trims[l] = 1 (+1 for "\r\n")
__global__ void __launch_bounds__(kSplitThreads) k_nl_write(SplitArgs s) {
    __shared__ uint32_t s_w[kSplitThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t pos = uint64_t(blockIdx.x) * kSplitBlockBytes + uint64_t(threadIdx.x) * kSplitBytesPerThread;
    uint32_t mask = newline_mask(s.text, s.n_bytes, pos);
    const uint32_t c = __popc(mask);
    const uint32_t incl = warp_incl_scan_u32(c, lane);
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t before = incl - c;
    for (int i = 0; i < warp; ++i) before += s_w[i];
    uint64_t line = s.blk_base[blockIdx.x] + before;
    while (mask) {
        const int b = __ffs(mask) - 1;
        mask &= mask - 1;
        const uint64_t p = pos + uint64_t(b);
        s.offsets[line + 1] = p + 1;
        s.trims[line] = (p > 0 && s.text[p - 1] == 0x0D) ? 2 : 1;
        ++line;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        s.offsets[0] = 0;
        if (s.n_bytes > 0 && s.text[s.n_bytes - 1] != 0x0A) {
            const uint64_t nl = *s.n_lines;
            s.offsets[nl] = s.n_bytes;
            s.trims[nl - 1] = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// tokenised output
// ------------------------------------------------------------------------------------------------
constexpr int kTokThreads = 256;

// Output bytes of every sentence of a 64-sentence group (one warp per sentence, 8 sentences per warp), their
// exclusive prefix inside the group, and the group total.
__global__ void __launch_bounds__(kTokThreads) k_tok_count(TokArgs t) {
    __shared__ uint32_t s_len[kGroup];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t gbase = uint64_t(blockIdx.x) * kGroup;
    const int ns = int(min(uint64_t(kGroup), t.n_sent - gbase));
    for (int i = warp; i < kGroup; i += kTokThreads / 32) {
        uint32_t len = 0;
        if (i < ns) {
            const uint64_t s = gbase + i;
            len = 1;  // the '\n'
            if (t.status[s] == 0) {
                const uint64_t o0 = t.offsets[s];
                const uint64_t a0 = o0 & ~3ull;
                const uint32_t b0 = uint32_t(o0 - a0), b1 = uint32_t(t.offsets[s + 1] - a0) - (t.trims ? t.trims[s] : 0);
                const uint8_t* __restrict__ base = t.text + a0;
                uint32_t cnt = 0;
                for (uint32_t addr = 4u * uint32_t(lane); addr < b1; addr += 128) {
                    const uint32_t lo = __ldg(reinterpret_cast<const uint32_t*>(base + addr));
                    const uint32_t in80 = inside80(addr, b0, b1);
                    cnt += __popc((eq_bytes(lo, 0x20u) | eq_bytes(lo, 0x2Fu) | eq_bytes(lo, 0x5Cu)) & in80);
                }
                // word boundaries of the sentence
                const uint32_t nch = t.n_chars[s];
                if (nch > 1) {
                    const uint64_t q0 = t.bound_offsets[s];
                    const uint64_t qa = q0 & ~3ull;
                    const uint32_t c0 = uint32_t(q0 - qa), c1 = c0 + nch - 1;
                    const uint8_t* __restrict__ bb = t.boundaries + qa;
                    for (uint32_t addr = 4u * uint32_t(lane); addr < c1; addr += 128) {
                        const uint32_t w = *reinterpret_cast<const uint32_t*>(bb + addr);
                        cnt += __popc(eq_bytes(w, 1u) & inside80(addr, c0, c1));
                    }
                }
                len += (b1 - b0) + __reduce_add_sync(kFull, cnt);
            }
        }
        if (lane == 0) s_len[i] = len;
    }
    __syncthreads();
    if (warp == 0) {
        const uint32_t v0 = s_len[2 * lane], v1 = s_len[2 * lane + 1];
        const uint32_t iv = warp_incl_scan_u32(v0 + v1, lane);
        const uint64_t s0 = gbase + 2 * lane;
        if (s0 < t.n_sent) t.tok_local[s0] = iv - v0 - v1;
        if (s0 + 1 < t.n_sent) t.tok_local[s0 + 1] = iv - v1;
        if (lane == 31) t.tok_group[blockIdx.x] = iv;
    }
}

// Exclusive scan of the group totals in place; element [ngroups] receives the grand total.
__global__ void __launch_bounds__(1024) k_tok_scan(uint64_t* g, uint64_t ngroups) {
    __shared__ uint64_t s_w[32];
    __shared__ uint64_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint64_t base = 0; base < ngroups; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const uint64_t v = i < ngroups ? g[i] : 0;
        uint64_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t o = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint64_t w = s_w[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t o = __shfl_up_sync(kFull, w, d);
                if (lane >= d) w += o;
            }
            s_w[lane] = w;
        }
        __syncthreads();
        if (i < ngroups) g[i] = s_carry + (warp ? s_w[warp - 1] : 0) + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += s_w[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) g[ngroups] = s_carry;
}

// One warp per sentence: every byte of the surface goes to  base + index + (escapes and spaces before it),
// preceded by its own ' ' (a word boundary before this character) and '\' (escape).
__global__ void __launch_bounds__(kTokThreads) k_tok_write(TokArgs t) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t gbase = uint64_t(blockIdx.x) * kGroup;
    const int ns = int(min(uint64_t(kGroup), t.n_sent - gbase));
    const uint64_t gout = t.tok_group[blockIdx.x];
    for (int i = warp; i < ns; i += kTokThreads / 32) {
        const uint64_t s = gbase + i;
        uint8_t* __restrict__ out = t.out + gout + t.tok_local[s];
        if (t.status[s] != 0) {
            if (lane == 0) out[0] = 0x0A;
            continue;
        }
        const uint64_t o0 = t.offsets[s];
        const uint64_t a0 = o0 & ~3ull;
        const uint32_t b0 = uint32_t(o0 - a0), b1 = uint32_t(t.offsets[s + 1] - a0) - (t.trims ? t.trims[s] : 0);
        const uint8_t* __restrict__ base = t.text + a0;
        const uint8_t* __restrict__ bnd = t.boundaries + t.bound_offsets[s];
        uint32_t chars = 0, extra = 0;  // characters / inserted bytes before this window
        for (uint32_t w0 = 0; w0 < b1; w0 += 128) {
            const uint32_t addr = w0 + 4u * uint32_t(lane);
            uint32_t lo = 0, in80 = 0;
            if (addr < b1) {
                lo = __ldg(reinterpret_cast<const uint32_t*>(base + addr));
                in80 = inside80(addr, b0, b1);
            }
            const uint32_t st80 = ~(lo & ~(lo << 1)) & in80;  // character starts (not 10xxxxxx)
            const uint32_t nst = __popc(st80);
            const uint32_t st_incl = warp_incl_scan_u32(nst, lane);
            // a ' ' goes before character k >= 1 when boundary k-1 is a word boundary
            uint32_t sp80 = 0;
            {
                uint32_t k = chars + st_incl - nst;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (st80 & (0x80u << (8 * j))) {
                        if (k >= 1 && bnd[k - 1] == 1) sp80 |= 0x80u << (8 * j);
                        ++k;
                    }
                }
            }
            const uint32_t esc80 = (eq_bytes(lo, 0x20u) | eq_bytes(lo, 0x2Fu) | eq_bytes(lo, 0x5Cu)) & in80;
            const uint32_t nex = __popc(sp80) + __popc(esc80);
            const uint32_t ex_incl = warp_incl_scan_u32(nex, lane);
            uint32_t at = (addr - b0) + extra + ex_incl - nex;  // output index of byte 0 of this word (if inside)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t bit = 0x80u << (8 * j);
                if (in80 & bit) {
                    if (sp80 & bit) out[at++] = 0x20;
                    if (esc80 & bit) out[at++] = 0x5C;
                    out[at] = uint8_t(lo >> (8 * j));
                }
                ++at;
            }
            chars += __shfl_sync(kFull, st_incl, 31);
            extra += __shfl_sync(kFull, ex_incl, 31);
        }
        if (lane == 0) out[(b1 - b0) + extra] = 0x0A;
    }
}

}  // namespace

cudaError_t launch_split_count(const SplitArgs& s, cudaStream_t stream) {
    const uint64_t nblk = (s.n_bytes + kSplitBlockBytes - 1) / kSplitBlockBytes;
    if (nblk) k_nl_count<<<unsigned(nblk), kSplitThreads, 0, stream>>>(s);
    k_nl_scan<<<1, 1024, 0, stream>>>(s, nblk);
    return cudaGetLastError();
}

cudaError_t launch_split_write(const SplitArgs& s, cudaStream_t stream) {
    const uint64_t nblk = (s.n_bytes + kSplitBlockBytes - 1) / kSplitBlockBytes;
    if (nblk == 0) return cudaSuccess;
    k_nl_write<<<unsigned(nblk), kSplitThreads, 0, stream>>>(s);
    return cudaGetLastError();
}

cudaError_t launch_tok_count(const TokArgs& t, cudaStream_t stream) {
    if (t.n_sent == 0) return cudaSuccess;
    const uint64_t ngroups = (t.n_sent + kGroup - 1) / kGroup;
    k_tok_count<<<unsigned(ngroups), kTokThreads, 0, stream>>>(t);
    k_tok_scan<<<1, 1024, 0, stream>>>(t.tok_group, ngroups);
    return cudaGetLastError();
}

cudaError_t launch_tok_write(const TokArgs& t, cudaStream_t stream) {
    if (t.n_sent == 0) return cudaSuccess;
    const uint64_t ngroups = (t.n_sent + kGroup - 1) / kGroup;
    k_tok_write<<<unsigned(ngroups), kTokThreads, 0, stream>>>(t);
    return cudaGetLastError();
}

}  // namespace vpt
