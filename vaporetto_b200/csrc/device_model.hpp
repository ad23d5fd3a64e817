// Device-side view of a predictor (pointers into the flat model blob in HBM) and the launch API
// between the C ABI (capi.cpp) and the kernels (kernels.cu).
#pragma once
#include <cstdint>

#include <cuda_runtime.h>

#include "keys.hpp"

namespace vpt {

struct DevTable {
    const void* records = nullptr;     // nslots x 32 B (FastRecord or GeneralRecord)
    const uint8_t* seeds = nullptr;    // nbuckets
    const uint32_t* slot_node = nullptr;
    const uint32_t* slot_pid = nullptr;
    const int32_t* pool = nullptr;     // general rows / overflow rows
    const uint64_t* slot_ovf = nullptr; // fast tables with overflow rows: ptr | off16 << 32 | len16 << 48
    uint64_t salt = 0;
    HashK hk{};                        // hash multipliers derived from the salt (keys.hpp)
    uint32_t nslots = 0;
    uint32_t nbuckets = 0;
    int32_t r0 = 0;
    uint32_t max_depth = 0;
    int32_t present = 0;
    int32_t fast = 0;
    int32_t seed16 = 0;                // seeds are 16-bit (dense tables; never staged in shared memory)
    int32_t has_overflow = 0;          // fast table whose deep records may carry kOvfFlag
};

struct DevModel {
    DevTable ct;                         // char + dictionary patterns
    DevTable tt;                         // type patterns (automaton variant only)
    const int32_t* type_cache = nullptr; // 8^(2W) table (cache variant only)
    int32_t type_cache_window = 0;       // 0 = no cache table
    const int32_t* type_a = nullptr;     // split tables for window 3 (T = A[t0..t3] + B[t2..t5]), or null
    const int32_t* type_b = nullptr;
    const uint32_t* type_state3 = nullptr;  // tag variant: pattern id by (t[-2] t[-1] t[0]) code, or null
    int32_t bias = 0;
    int32_t char_window = 0;
    int32_t type_window = 0;
    int32_t emit_states = 0;             // tag variant: pattern-id states are meaningful
    int32_t kytea_norm = 0;              // per call: score KyteaFullwidthFilter(text) (textnorm.hpp) instead of text
};

// Per-batch device buffers (all device pointers).
struct BatchArgs {
    const uint8_t* text = nullptr;        // concatenated UTF-8, 16-byte aligned, readable up to a multiple of 16
    const uint64_t* offsets = nullptr;    // [n_sent + 1] byte offsets into text
    const uint8_t* trims = nullptr;       // nullable [n_sent]: separator bytes at the end of [offsets[i], offsets[i+1])
                                          // that are not part of sentence i (line terminators, see lines.cu)
    uint64_t n_sent = 0;
    // scratch written by the count pass
    uint32_t* n_chars = nullptr;          // [n_sent]
    int32_t* status = nullptr;            // [n_sent] 0 ok, 1 empty, 2 NUL, 3 invalid UTF-8
    uint32_t* local_bound = nullptr;      // [n_sent] boundary offset inside its 64-sentence group
    uint32_t* local_char = nullptr;       // [n_sent] char offset inside its group
    uint64_t* group_bound = nullptr;      // [n_groups + 1]
    uint64_t* group_char = nullptr;       // [n_groups + 1]
    uint32_t* ticket = nullptr;           // work counter of the persistent tile kernel (zeroed by the scan)
    bool prezeroed = false;               // k_fused: group_bound / group_char / ticket are already zero (no memset node)
    bool self_clean = false;              // k_fused, single-CTA launches only: zero them again before the kernel ends
    uint64_t* totals_host = nullptr;      // nullable: pinned host memory [2]; the scan also stores the batch's boundary
                                          // and character totals there (no D2H copy queued behind bulk copies)
    // outputs
    int32_t* scores = nullptr;            // [sum(max(chars_i - 1, 0))] (nullable: boundaries only)
    uint8_t* boundaries = nullptr;        // same length
    uint64_t* bound_offsets = nullptr;    // [n_sent + 1]; values are chunk-local offsets + bound_base
    uint64_t bound_base = 0;              // added to the bound_offsets / char_offsets written out
    uint64_t char_base = 0;
    uint64_t* char_offsets = nullptr;     // [n_sent + 1] (nullable)
    uint32_t* char_states = nullptr;      // [sum(chars_i)] (nullable)
    uint32_t* type_states = nullptr;      // [sum(chars_i)] (nullable)
};

constexpr int kGroup = 64;  // sentences per count-pass block

// Launch helpers; all asynchronous on `stream`.  Return cudaError_t of the launch.
cudaError_t launch_count(const BatchArgs& a, cudaStream_t stream);  // count + scan
cudaError_t launch_count_only(const BatchArgs& a, cudaStream_t stream);
cudaError_t launch_scan_only(const BatchArgs& a, cudaStream_t stream);
cudaError_t launch_score(const DevModel& m, const BatchArgs& a, cudaStream_t stream);
// fused.cu: validation, counts, output offsets (decoupled look-back) and scoring of a batch in one launch, for the
// inline-row model shapes fused_ok() accepts; needs neither launch_count nor the count pass's scratch arrays
// (group_bound / group_char / ticket are used as look-back descriptors and cleared by a memset node)
bool fused_ok(const DevModel& m);
cudaError_t launch_fused(const DevModel& m, const BatchArgs& a, cudaStream_t stream);
// count (+ scan) + score, or the fused launch when the model qualifies
cudaError_t launch_batch(const DevModel& m, const BatchArgs& a, cudaStream_t stream);
// number of kernel launches issued by launch_batch for this model
int launches_per_batch(const DevModel& m);
// true when launch_score accepts BatchArgs::scores == nullptr (boundaries only) for this model
bool scores_optional(const DevModel& m);

// ---- lines.cu: device-side line splitting and tokenised output -------------------------------------
constexpr int kSplitBlockBytes = 8192;  // bytes of text per CTA of the line splitter

struct SplitArgs {
    const uint8_t* text = nullptr;   // 16-byte aligned, readable up to a multiple of 16
    uint64_t n_bytes = 0;
    uint32_t* blk = nullptr;         // [n_blocks] newline count per block (scratch)
    uint64_t* blk_base = nullptr;    // [n_blocks] lines before each block (scratch)
    uint64_t* n_lines = nullptr;     // device scalar: number of lines
    uint64_t* n_lines_host = nullptr;  // nullable: pinned host memory, receives the same number
    uint64_t* offsets = nullptr;     // [n_lines + 1] out: line starts (+ n_bytes)
    uint8_t* trims = nullptr;        // [n_lines] out: terminator bytes of each line (0, 1 or 2)
};
cudaError_t launch_split_count(const SplitArgs& s, cudaStream_t stream);  // fills *n_lines (and the scratch)
cudaError_t launch_split_write(const SplitArgs& s, cudaStream_t stream);  // fills offsets / trims

struct TokArgs {
    const uint8_t* text = nullptr;          // as BatchArgs (4-byte aligned is enough)
    const uint64_t* offsets = nullptr;
    const uint8_t* trims = nullptr;         // nullable
    uint64_t n_sent = 0;
    const int32_t* status = nullptr;        // from the count pass
    const uint32_t* n_chars = nullptr;
    const uint8_t* boundaries = nullptr;    // 4-byte aligned; from the scoring pass
    const uint64_t* bound_offsets = nullptr;  // index of a sentence's first boundary in `boundaries`
    uint64_t* tok_state = nullptr;          // [n_groups] scratch: look-back state of every 64-sentence group
    uint32_t* ticket = nullptr;             // scratch: group ticket (the 8 bytes after tok_state)
    uint64_t* total = nullptr;              // device scalar out: total output bytes
    uint64_t* total_host = nullptr;         // nullable: pinned host memory, receives the same number
    uint8_t* out = nullptr;                 // tokenised lines, each terminated by '\n'
    // tags (vpt_tokenize_lines_tags): the per-token records of the tag prediction and the model's tag strings; a token's
    // "/tag" suffixes are written behind its surface (Sentence::write_tokenized_text, sentence.rs:850-886)
    const uint64_t* tok_base = nullptr;     // [n_sent + 1] index of a sentence's first token record; nullptr = no tags
    const int32_t* tok_ids = nullptr;       // [n_tokens] token id or -1
    const uint8_t* tok_cands = nullptr;     // [n_tokens * n_tags] chosen candidate per slot, 255 = none
    uint32_t n_tags = 0;
    const uint32_t* ts_slot = nullptr;      // [n_token_ids] first slot entry of a token id
    const uint32_t* ts_cand = nullptr;      // [sum of slots] first string reference of a slot
    const uint2* ts_ref = nullptr;          // [sum of candidates] (offset, length) of the escaped tag string
    const uint8_t* ts_bytes = nullptr;
};
// zeroes tok_state[0 .. n_groups] (ticket included), then one pass: lengths, offsets (look-back), output bytes
cudaError_t launch_tokenize(const TokArgs& t, cudaStream_t stream);
// KyteaWsConstFilter for the character types in `mask` (bit t = CharacterType t): clears boundaries between two
// characters of such a type; uses text / offsets / trims / status / n_chars / bound_offsets of `t`
cudaError_t launch_wsconst(const TokArgs& t, uint8_t* boundaries, uint32_t mask, bool norm, cudaStream_t stream);
cudaError_t launch_grapheme(const TokArgs& t, uint8_t* boundaries, bool norm, cudaStream_t stream);

}  // namespace vpt
