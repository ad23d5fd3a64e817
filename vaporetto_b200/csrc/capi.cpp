// C ABI (include/vaporetto_b200.h) — host side of the predictor: model parsing, table build, upload,
// batch staging, tag prediction and the Sentence helpers.  Compiled with nvcc (needs cuda_runtime.h).
#include "../../include/vaporetto_b200.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

#include "builder.hpp"
#include "common.hpp"
#include "device_model.hpp"
#include "model.hpp"
#include "predictor_build.hpp"
#include "grapheme.hpp"
#include "tags.hpp"
#include "zstd_loader.hpp"
#include "textnorm.hpp"

using namespace vpt;

struct vpt_model {
    Model m;
};

namespace {

void cuda_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess)
        throw Error(kCudaError, std::string("CUDA error in ") + what + ": " + cudaGetErrorString(e));
}

// Per-call scratch: a stream plus grow-only device buffers.
struct Scratch {
    cudaStream_t stream = nullptr;      // copy-in + kernels
    cudaStream_t stream_out = nullptr;  // copy-out: a chunk's D2H never delays the next chunk's H2D on `stream`
    cudaEvent_t ev_kernels = nullptr;   // recorded on `stream` after a chunk's last kernel
    cudaEvent_t ev_out = nullptr;       // recorded on `stream_out` after a chunk's last D2H copy
    void* d_text = nullptr; size_t text_cap = 0;
    void* d_off = nullptr; size_t off_cap = 0;
    void* d_ws = nullptr; size_t ws_cap = 0;
    void* d_status = nullptr; size_t status_cap = 0;
    void* d_boff = nullptr; size_t boff_cap = 0;
    void* d_coff = nullptr; size_t coff_cap = 0;
    void* d_scores = nullptr; size_t scores_cap = 0;
    void* d_bounds = nullptr; size_t bounds_cap = 0;
    void* d_cst = nullptr; size_t cst_cap = 0;
    void* d_tst = nullptr; size_t tst_cap = 0;
    // line splitting / tokenised output (vpt_tokenize_lines)
    void* d_trims = nullptr; size_t trims_cap = 0;
    void* d_blk = nullptr; size_t blk_cap = 0;
    void* d_blkbase = nullptr; size_t blkbase_cap = 0;
    void* d_tokg = nullptr; size_t tokg_cap = 0;
    void* d_out = nullptr; size_t out_cap = 0;
    void* d_tok = nullptr; size_t tok_cap = 0;     // tag prediction outputs (vpt_predict_batch_tags)
    void* d_cand = nullptr; size_t cand_cap = 0;
    void* d_bits = nullptr; size_t bits_cap = 0;   // compact outputs (vpt_predict_batch_compact)
    void* d_st8 = nullptr; size_t st8_cap = 0;
    void* d_ntok = nullptr; size_t ntok_cap = 0;
    void* d_tokbase = nullptr; size_t tokbase_cap = 0;
    void* d_tokdesc = nullptr; size_t tokdesc_cap = 0;
    void* d_tokwork = nullptr; size_t tokwork_cap = 0;  // work list of the per-token tag kernels (TagArgs::tok_work)
    void* d_toklocal = nullptr; size_t toklocal_cap = 0;
    void* d_tokblk = nullptr; size_t tokblk_cap = 0;
    uint64_t* h_totals = nullptr;  // pinned, 8 x u64: boundaries, chars, lines, output bytes, tokens, first bit word
    uint32_t* h_side = nullptr; size_t side_cap = 0;  // pinned: first bit word of every chunk (vpt_predict_batch_compact)
    uint8_t* h_io = nullptr;       // pinned staging of the single-sentence call (vpt_predict), kSingleIoBytes
    void* d_io = nullptr;          // its device twin
    ~Scratch() {
        for (void* p : {d_text, d_off, d_ws, d_status, d_boff, d_coff, d_scores, d_bounds, d_cst, d_tst, d_trims, d_blk,
                        d_blkbase, d_tokg, d_out, d_tok, d_cand, d_bits, d_st8, d_ntok, d_tokbase, d_tokdesc, d_tokwork, d_toklocal, d_tokblk})
            if (p) cudaFree(p);
        if (h_totals) cudaFreeHost(h_totals);
        if (h_io) cudaFreeHost(h_io);
        if (h_side) cudaFreeHost(h_side);
        if (d_io) cudaFree(d_io);
        if (ev_kernels) cudaEventDestroy(ev_kernels);
        if (ev_out) cudaEventDestroy(ev_out);
        if (stream_out) cudaStreamDestroy(stream_out);
        if (stream) cudaStreamDestroy(stream);
    }
    static void ensure(void*& p, size_t& cap, size_t need) {
        if (need <= cap) return;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = align_up(need + need / 8 + 256, 256);
        cuda_check(cudaMalloc(&p, want), "cudaMalloc(scratch)");
        cap = want;
    }
};

}  // namespace

struct vpt_predictor : HostPredictor {
    int device = 0;
    bool from_blob = false;
    void* d_blob = nullptr;
    DevModel dm;
    // device-side tag prediction (tags.hpp): one allocation holding all tables; dt.tok_tab == nullptr when unavailable
    void* d_tags = nullptr;
    DevTags dt;
    // scratch pool
    mutable std::mutex mu;
    mutable std::vector<std::unique_ptr<Scratch>> pool;

    ~vpt_predictor() {
        if (d_blob) {
            cudaSetDevice(device);
            pool.clear();
            cudaFree(d_blob);
            if (d_tags) cudaFree(d_tags);
        }
    }
};

namespace {

int fail(const Error& e) {
    set_last_error(e.what());
    return e.code;
}
int fail(const std::exception& e) {
    set_last_error(std::string("internal error: ") + e.what());
    return kInternal;
}

#define VPT_API_BEGIN try {
#define VPT_API_END \
    } catch (const Error& e) { return fail(e); } \
      catch (const std::exception& e) { return fail(e); }

DevTable dev_table(const BlobTable& bt, const uint8_t* base) {
    DevTable d;
    d.present = bt.present;
    if (!bt.present) return d;
    d.fast = bt.fast;
    d.r0 = bt.r0;
    d.max_depth = bt.max_depth;
    d.nslots = bt.nslots;
    d.nbuckets = bt.nbuckets;
    d.salt = bt.salt;
    d.hk = hash_consts(bt.salt);
    d.seed16 = bt.seed_bits == 16;
    d.has_overflow = bt.has_overflow;
    d.slot_ovf = bt.has_overflow ? reinterpret_cast<const uint64_t*>(base + bt.ovf_off) : nullptr;
    d.records = base + bt.rec_off;
    d.seeds = base + bt.seeds_off;
    d.slot_node = reinterpret_cast<const uint32_t*>(base + bt.node_off);
    d.slot_pid = reinterpret_cast<const uint32_t*>(base + bt.pid_off);
    d.pool = reinterpret_cast<const int32_t*>(base + bt.pool_off);
    return d;
}

// Every section a blob header points at must lie inside the blob, and the geometry must be what the kernels were
// built for: a truncated or corrupted blob is an InvalidModel error, not an out-of-bounds read on the device.
void validate_blob_header(const BlobHeader& h, uint64_t len) {
    auto inside = [&](uint64_t off, uint64_t bytes) { return off >= sizeof(BlobHeader) && off <= len && bytes <= len - off; };
    auto bad = [](const char* what) { return Error(kInvalidModel, std::string("InvalidModelError: model blob: bad ") + what); };
    auto table = [&](const BlobTable& t, const char* name) {
        if (!t.present) return;
        if (t.nslots == 0 || t.nbuckets == 0 || (t.seed_bits != 8 && t.seed_bits != 16)) throw bad(name);
        if (!inside(t.rec_off, uint64_t(t.nslots) * 32) || !inside(t.seeds_off, uint64_t(t.nbuckets) * (t.seed_bits / 8)) ||
            !inside(t.node_off, uint64_t(t.nslots) * 4) || !inside(t.pid_off, uint64_t(t.nslots) * 4) || !inside(t.pool_off, 4))
            throw bad(name);
        if (t.has_overflow && !inside(t.ovf_off, uint64_t(t.nslots) * 8)) throw bad(name);
        if (t.r0 < -64 || t.r0 > 64) throw bad(name);
    };
    table(h.ct, "char table");
    table(h.tt, "type table");
    if (h.char_window < 0 || h.char_window > 255 || h.type_window < 0 || h.type_window > 255) throw bad("window");
    if (h.type_cache_window < 0 || h.type_cache_window > 3) throw bad("type table window");
    if (h.type_cache_window && !inside(h.type_cache_off, (uint64_t(4) << (6 * h.type_cache_window)))) throw bad("type table");
    if (h.type_a_off && (!inside(h.type_a_off, 4 * 4096) || !inside(h.type_b_off, 4 * 4096))) throw bad("split type tables");
    if (h.type_state3_off && !inside(h.type_state3_off, 4 * 512)) throw bad("type state table");
}

void upload(vpt_predictor& p) {
    if (p.device == -1) return;  // host-only handle (tag prediction / Sentence helpers); cannot score
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw Error(kCudaError, std::string("no usable CUDA device (vaporetto_b200 has no CPU fallback): ") +
                                    cudaGetErrorString(e));
    if (p.device < 0 || p.device >= ndev) throw Error(kInvalidArgument, "InvalidArgumentError: device: out of range");
    cuda_check(cudaSetDevice(p.device), "cudaSetDevice");
    cuda_check(cudaMalloc(&p.d_blob, align_up(p.blob.size(), 256)), "cudaMalloc(model)");
    cuda_check(cudaMemcpy(p.d_blob, p.blob.data(), p.blob.size(), cudaMemcpyHostToDevice), "cudaMemcpy(model)");
    const uint8_t* base = static_cast<const uint8_t*>(p.d_blob);
    const BlobHeader& h = p.hdr;
    p.dm = DevModel();
    p.dm.ct = dev_table(h.ct, base);
    p.dm.tt = dev_table(h.tt, base);
    p.dm.type_cache_window = h.type_cache_window;
    p.dm.type_cache = h.type_cache_window ? reinterpret_cast<const int32_t*>(base + h.type_cache_off) : nullptr;
    p.dm.type_a = h.type_a_off ? reinterpret_cast<const int32_t*>(base + h.type_a_off) : nullptr;
    p.dm.type_b = h.type_b_off ? reinterpret_cast<const int32_t*>(base + h.type_b_off) : nullptr;
    p.dm.type_state3 = h.type_state3_off ? reinterpret_cast<const uint32_t*>(base + h.type_state3_off) : nullptr;
    p.dm.bias = h.bias;
    p.dm.char_window = h.char_window;
    p.dm.type_window = h.type_window;
    p.dm.emit_states = h.emit_states;
}

// Tag tables of a tag predictor: built on the host (tags_build.cpp), one device allocation.
void upload_tags(vpt_predictor& p) {
    if (p.device == -1 || !p.predict_tags || p.n_tags == 0) return;
    const TagTablesHost t = build_tag_tables(p);
    if (!t.usable) return;
    struct Part { const void* src; size_t bytes; size_t off; };
    std::vector<Part> parts;
    size_t total = 0;
    auto add = [&](const void* src, size_t bytes) {
        parts.push_back({src, bytes, total});
        total = align_up(total + bytes, 256);
        return parts.size() - 1;
    };
    const size_t i_tok = add(t.tok_tab.data(), t.tok_tab.size() * sizeof(TagTokenEntry));
    const size_t i_bytes = add(t.tok_bytes.data(), t.tok_bytes.size());
    const size_t i_info = add(t.tok_info.data(), t.tok_info.size() * sizeof(TagTokenInfo));
    const size_t i_pool = add(t.pool.data(), t.pool.size() * 4);
    const size_t i_keys = add(t.keys.data(), t.keys.size() * sizeof(TagKey));
    const size_t i_tss = add(t.ts_slot.data(), t.ts_slot.size() * 4);
    const size_t i_tsc = add(t.ts_cand.data(), t.ts_cand.size() * 4);
    const size_t i_tsr = add(t.ts_ref.data(), t.ts_ref.size() * 4);
    const size_t i_tsb = add(t.ts_bytes.data(), t.ts_bytes.size());
    const size_t i_cc = add(t.c_chain.data(), t.c_chain.size() * sizeof(TagChain));
    const size_t i_tc = add(t.t_chain.data(), t.t_chain.size() * sizeof(TagChain));
    const size_t i_cl = add(t.c_link.data(), t.c_link.size() * 4);
    const size_t i_tl = add(t.t_link.data(), t.t_link.size() * 4);
    cuda_check(cudaSetDevice(p.device), "cudaSetDevice");
    cuda_check(cudaMalloc(&p.d_tags, total + 256), "cudaMalloc(tag tables)");
    uint8_t* base = static_cast<uint8_t*>(p.d_tags);
    for (const Part& q : parts)
        if (q.bytes) cuda_check(cudaMemcpy(base + q.off, q.src, q.bytes, cudaMemcpyHostToDevice), "cudaMemcpy(tag tables)");
    DevTags& d = p.dt;
    d.tok_tab = reinterpret_cast<const TagTokenEntry*>(base + parts[i_tok].off);
    d.tok_bytes = base + parts[i_bytes].off;
    d.tok_info = reinterpret_cast<const TagTokenInfo*>(base + parts[i_info].off);
    d.pool = reinterpret_cast<const int32_t*>(base + parts[i_pool].off);
    d.keys = reinterpret_cast<const TagKey*>(base + parts[i_keys].off);
    d.ts_slot = reinterpret_cast<const uint32_t*>(base + parts[i_tss].off);
    d.ts_cand = reinterpret_cast<const uint32_t*>(base + parts[i_tsc].off);
    d.ts_ref = reinterpret_cast<const uint2*>(base + parts[i_tsr].off);
    d.ts_bytes = base + parts[i_tsb].off;
    d.max_suffix = t.max_suffix;
    d.c_chain = reinterpret_cast<const TagChain*>(base + parts[i_cc].off);
    d.t_chain = reinterpret_cast<const TagChain*>(base + parts[i_tc].off);
    d.c_link = reinterpret_cast<const uint32_t*>(base + parts[i_cl].off);
    d.t_link = reinterpret_cast<const uint32_t*>(base + parts[i_tl].off);
    d.tok_mask = t.tok_mask;
    d.n_tags = t.n_tags;
    d.char_rels = p.char_tags ? t.char_rels : 0;
    d.type_rels = p.type_tags ? t.type_rels : 0;
    d.max_token_bytes = t.max_token_bytes;
    d.n_char_patterns = uint32_t(p.char_suffix_link.size());
    d.n_type_patterns = uint32_t(p.type_suffix_link.size());
}

struct ScratchLease {
    const vpt_predictor& p;
    std::unique_ptr<Scratch> s;
    explicit ScratchLease(const vpt_predictor& pr) : p(pr) {
        {
            std::lock_guard<std::mutex> g(p.mu);
            if (!p.pool.empty()) { s = std::move(p.pool.back()); p.pool.pop_back(); }
        }
        if (!s) {
            s.reset(new Scratch());
            cuda_check(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking), "cudaStreamCreate");
            cuda_check(cudaStreamCreateWithFlags(&s->stream_out, cudaStreamNonBlocking), "cudaStreamCreate");
            cuda_check(cudaEventCreateWithFlags(&s->ev_kernels, cudaEventDisableTiming), "cudaEventCreate");
            cuda_check(cudaEventCreateWithFlags(&s->ev_out, cudaEventDisableTiming), "cudaEventCreate");
            cuda_check(cudaMallocHost(reinterpret_cast<void**>(&s->h_totals), 64), "cudaMallocHost");
        }
    }
    ~ScratchLease() {
        // a lease can end on an exception path while copies into the caller's buffers or kernels are still queued:
        // drain both streams before the scratch becomes reusable and control returns to the caller
        if (s) {
            if (s->stream) cudaStreamSynchronize(s->stream);
            if (s->stream_out) cudaStreamSynchronize(s->stream_out);
        }
        std::lock_guard<std::mutex> g(p.mu);
        p.pool.push_back(std::move(s));
    }
};

struct WorkspaceLayout {
    size_t n_chars, local_bound, local_char, group_bound, group_char, ticket, total;
};
WorkspaceLayout workspace_layout(size_t n) {
    WorkspaceLayout l;
    // (the look-back descriptors of k_fused live in group_bound / group_char: its tiles hold at least 32 sentences)
    const size_t ng = (n + 31) / 32;
    size_t o = 0;
    l.n_chars = o; o = align_up(o + 4 * n, 256);
    l.local_bound = o; o = align_up(o + 4 * n, 256);
    l.local_char = o; o = align_up(o + 4 * n, 256);
    l.group_bound = o; o = align_up(o + 8 * (ng + 1), 256);
    l.group_char = o; o = align_up(o + 8 * (ng + 1), 256);
    l.ticket = o; o += 256;
    l.total = o + 256;
    return l;
}

void bind_workspace(BatchArgs& a, void* ws, size_t n) {
    const WorkspaceLayout l = workspace_layout(n);
    uint8_t* b = static_cast<uint8_t*>(ws);
    a.n_chars = reinterpret_cast<uint32_t*>(b + l.n_chars);
    a.local_bound = reinterpret_cast<uint32_t*>(b + l.local_bound);
    a.local_char = reinterpret_cast<uint32_t*>(b + l.local_char);
    a.group_bound = reinterpret_cast<uint64_t*>(b + l.group_bound);
    a.group_char = reinterpret_cast<uint64_t*>(b + l.group_char);
    a.ticket = reinterpret_cast<uint32_t*>(b + l.ticket);
}

// ---- host-side text helpers ------------------------------------------------------------------------

// Sentence::parse_raw checks (reference sentence.rs:160-196)
void check_raw_text(const uint8_t* s, size_t n) {
    if (!is_valid_utf8(s, n)) throw Error(kInvalidArgument, "InvalidArgumentError: text: must be valid UTF-8");
    if (memchr(s, 0, n) != nullptr) throw Error(kInvalidArgument, "InvalidArgumentError: text: must not contain NULL");
    if (n == 0) throw Error(kInvalidArgument, "InvalidArgumentError: text: must contain at least one character");
}

// CharacterType::get_type (reference sentence.rs:50-67)
uint8_t host_char_type(uint32_t c) {
    struct R { uint32_t lo, hi; uint8_t t; };
    static const R ranges[] = {
        {0x30, 0x39, 1}, {0xFF10, 0xFF19, 1},
        {0x41, 0x5A, 2}, {0x61, 0x7A, 2}, {0xFF21, 0xFF3A, 2}, {0xFF41, 0xFF5A, 2},
        {0x3040, 0x3096, 3},
        {0x30A0, 0x30FA, 4}, {0x30FC, 0x30FF, 4}, {0xFF66, 0xFF9F, 4},
        {0x3400, 0x4DBF, 5}, {0x4E00, 0x9FFF, 5}, {0xF900, 0xFAFF, 5}, {0x20000, 0x2A6DF, 5},
        {0x2A700, 0x2B73F, 5}, {0x2B740, 0x2B81F, 5}, {0x2B820, 0x2CEAF, 5}, {0x2F800, 0x2FA1F, 5},
    };
    for (const R& r : ranges) if (c >= r.lo && c <= r.hi) return r.t;
    return 6;
}

// byte offset of every character start, plus the end (char_to_str_pos, sentence.rs:100)
std::vector<uint32_t> char_starts(const uint8_t* s, size_t n) {
    std::vector<uint32_t> v;
    for (size_t i = 0; i < n; ++i) if ((s[i] & 0xC0) != 0x80) v.push_back(uint32_t(i));
    v.push_back(uint32_t(n));
    return v;
}

void add_truncated(const std::vector<int32_t>& w, std::vector<int32_t>& ys) {  // WeightVector::add_scores
    const size_t n = std::min(w.size(), ys.size());
    for (size_t i = 0; i < n; ++i) ys[i] = wrapping_add(ys[i], w[i]);
}

// Tag weight of pattern `pid` for one (token, rel) table, as the reference's build-time suffix merge would have
// produced it (PositionalWeightWithTag +=, predictor.rs:242-262, applied along the chain of suffix patterns,
// char_scorer.rs:50-78): a pattern's own vector plus the merged vector of its longest suffix pattern truncated
// to the own length; without an own vector, the suffix's merged vector unchanged.
bool merged_tag_weight(const std::unordered_map<uint32_t, std::vector<int32_t>>& table,
                       const std::vector<uint32_t>& suffix_link, uint32_t pid, std::vector<int32_t>& out) {
    // collect the chain entries (longest pattern first)
    std::vector<const std::vector<int32_t>*> chain;
    for (uint32_t q = pid; q != kNoPattern; q = suffix_link[q]) {
        auto it = table.find(q);
        if (it != table.end()) chain.push_back(&it->second);
    }
    const int n = int(chain.size());
    if (n == 0) return false;
    // evaluate from the shortest suffix outwards
    out = *chain[n - 1];
    for (int i = n - 2; i >= 0; --i) {
        std::vector<int32_t> cur = *chain[i];
        const size_t m = std::min(cur.size(), out.size());
        for (size_t k = 0; k < m; ++k) cur[k] = wrapping_add(cur[k], out[k]);
        out.swap(cur);
    }
    return true;
}

void add_tag_scores(const TagWeightMap& tw, const std::vector<uint32_t>& suffix_link, uint32_t token, size_t pos,
                    const uint32_t* states, size_t n, std::vector<int32_t>& scores) {  // boundary_tag_scorer.rs:154-174
    const auto& per_rel = tw[token];
    std::vector<int32_t> w;
    for (size_t r = 0; r < per_rel.size() && pos + r < n; ++r) {
        const uint32_t pid = states[pos + r];
        if (pid == kNoPattern || pid >= suffix_link.size() || per_rel[r].empty()) continue;
        if (merged_tag_weight(per_rel[r], suffix_link, pid, w)) add_truncated(w, scores);
    }
}

}  // namespace

extern "C" {

const char* vpt_last_error(void) { return last_error(); }
const char* vpt_version(void) { return "vaporetto_b200 0.1.0 sm_100a"; }

int vpt_model_read(const uint8_t* data, size_t len, vpt_model** out, size_t* consumed) {
    VPT_API_BEGIN
    if (!out) throw Error(kInvalidArgument, "InvalidArgumentError: out: must not be NULL");
    *out = nullptr;
    std::unique_ptr<vpt_model> m(new vpt_model());
    m->m = Model::read(data, len, consumed);
    *out = m.release();
    return kOk;
    VPT_API_END
}

int vpt_model_read_zstd(const uint8_t* data, size_t len, vpt_model** out) {
    VPT_API_BEGIN
    if (!out) throw Error(kInvalidArgument, "InvalidArgumentError: out: must not be NULL");
    *out = nullptr;
    if (len && !data) throw Error(kInvalidArgument, "InvalidArgumentError: data: must not be NULL");
    std::unique_ptr<vpt_model> m(new vpt_model());
    if (is_zstd_frame(data, len)) {
        const std::vector<uint8_t> raw = zstd_decode_all(data, len);
        m->m = Model::read(raw.data(), raw.size(), nullptr);
    } else {
        m->m = Model::read(data, len, nullptr);
    }
    *out = m.release();
    return kOk;
    VPT_API_END
}

int vpt_model_read_kytea(const uint8_t* data, size_t len, vpt_model** out) {
    VPT_API_BEGIN
    if (!out) throw Error(kInvalidArgument, "InvalidArgumentError: out: must not be NULL");
    *out = nullptr;
    std::unique_ptr<vpt_model> m(new vpt_model());
    m->m = Model::from_kytea(data, len);
    *out = m.release();
    return kOk;
    VPT_API_END
}

int vpt_model_to_vec(const vpt_model* model, uint8_t** bytes_out, uint64_t* len_out) {
    VPT_API_BEGIN
    if (!model || !bytes_out || !len_out) throw Error(kInvalidArgument, "InvalidArgumentError: model/out: must not be NULL");
    *bytes_out = nullptr;
    const std::vector<uint8_t> v = model->m.to_vec();
    uint8_t* buf = static_cast<uint8_t*>(malloc(v.size() ? v.size() : 1));
    if (!buf) throw Error(kInternal, "internal error: out of memory");
    memcpy(buf, v.data(), v.size());
    *bytes_out = buf;
    *len_out = v.size();
    return kOk;
    VPT_API_END
}

uint64_t vpt_model_dictionary_len(const vpt_model* model) { return model ? model->m.dict.size() : 0; }

int vpt_model_dictionary_get(const vpt_model* model, uint64_t index, const char** word, const int32_t** weights,
                             uint64_t* n_weights, const char** comment) {
    VPT_API_BEGIN
    if (!model) throw Error(kInvalidArgument, "InvalidArgumentError: model: must not be NULL");
    if (index >= model->m.dict.size()) throw Error(kInvalidArgument, "InvalidArgumentError: index: out of range");
    const DictEntry& e = model->m.dict[index];
    if (word) *word = e.word.c_str();
    if (weights) *weights = e.weights.data();
    if (n_weights) *n_weights = e.weights.size();
    if (comment) *comment = e.comment.c_str();
    return kOk;
    VPT_API_END
}

int vpt_model_replace_dictionary(vpt_model* model, const char* const* words, const int32_t* const* weights,
                                 const uint64_t* n_weights, const char* const* comments, uint64_t n_records) {
    VPT_API_BEGIN
    if (!model) throw Error(kInvalidArgument, "InvalidArgumentError: model: must not be NULL");
    if (n_records && (!words || !weights || !n_weights))
        throw Error(kInvalidArgument, "InvalidArgumentError: words/weights/n_weights: must not be NULL");
    std::vector<DictEntry> dict;
    dict.reserve(n_records);
    for (uint64_t i = 0; i < n_records; ++i) {
        DictEntry e;
        e.word = words[i] ? words[i] : "";
        if (!is_valid_utf8(reinterpret_cast<const uint8_t*>(e.word.data()), e.word.size()))
            throw Error(kInvalidArgument, "InvalidArgumentError: word: must be valid UTF-8");
        // WordWeightRecord::new (dict_model.rs:39-50)
        if (n_weights[i] != utf8_to_codepoints(e.word).size() + 1)
            throw Error(kInvalidArgument, "InvalidArgumentError: weights: does not match the length of the `word`");
        if (n_weights[i] && !weights[i]) throw Error(kInvalidArgument, "InvalidArgumentError: weights: must not be NULL");
        e.weights.assign(weights[i], weights[i] + n_weights[i]);
        e.comment = (comments && comments[i]) ? comments[i] : "";
        dict.push_back(std::move(e));
    }
    model->m.dict = std::move(dict);
    return kOk;
    VPT_API_END
}

void vpt_model_free(vpt_model* model) { delete model; }

int vpt_predictor_new(vpt_model* model, int predict_tags, int device, vpt_predictor** out) {
    std::unique_ptr<vpt_model> owned(model);  // consumed like Predictor::new(model, ..)
    VPT_API_BEGIN
    if (!out || !model) throw Error(kInvalidArgument, "InvalidArgumentError: model/out: must not be NULL");
    *out = nullptr;
    std::unique_ptr<vpt_predictor> p(new vpt_predictor());
    static_cast<HostPredictor&>(*p) = build_host_predictor(owned->m, predict_tags != 0);
    p->device = device;
    upload(*p);
    upload_tags(*p);
    *out = p.release();
    return kOk;
    VPT_API_END
}

void vpt_predictor_free(vpt_predictor* predictor) { delete predictor; }

int vpt_predictor_get_info(const vpt_predictor* p, vpt_predictor_info* o) {
    VPT_API_BEGIN
    if (!p || !o) throw Error(kInvalidArgument, "InvalidArgumentError: predictor/out: must not be NULL");
    memset(o, 0, sizeof *o);
    o->device = p->device;
    o->predict_tags = p->predict_tags;
    o->n_tags = int32_t(p->n_tags);
    o->char_scorer = p->hdr.char_variant;
    o->type_scorer = p->hdr.type_variant;
    o->fast_path = (!p->hdr.ct.present || p->hdr.ct.fast) && !p->hdr.tt.present;
    o->bias = p->hdr.bias;
    o->char_window = p->hdr.char_window;
    o->type_window = p->hdr.type_window;
    o->n_char_patterns = p->hdr.ct.n_patterns;
    o->n_type_patterns = p->hdr.tt.n_patterns;
    o->n_char_nodes = p->hdr.ct.n_nodes;
    o->n_type_nodes = p->hdr.tt.n_nodes;
    o->max_char_pattern_len = uint32_t(p->hdr.max_char_pattern_len);
    o->blob_bytes = p->blob.size();
    o->kernel_launches_per_batch = launches_per_batch(p->dm);
    return kOk;
    VPT_API_END
}

int vpt_blob_build(vpt_model* model, int predict_tags, uint8_t** blob_out, uint64_t* len_out) {
    std::unique_ptr<vpt_model> owned(model);
    VPT_API_BEGIN
    if (!model || !blob_out || !len_out) throw Error(kInvalidArgument, "InvalidArgumentError: model/out: must not be NULL");
    *blob_out = nullptr;
    HostPredictor hp = build_host_predictor(owned->m, predict_tags != 0);
    uint8_t* buf = static_cast<uint8_t*>(malloc(hp.blob.size()));
    if (!buf) throw Error(kInternal, "internal error: out of memory");
    memcpy(buf, hp.blob.data(), hp.blob.size());
    *blob_out = buf;
    *len_out = hp.blob.size();
    return kOk;
    VPT_API_END
}

void vpt_blob_free(uint8_t* blob) { free(blob); }

uint64_t vpt_predictor_blob_size(const vpt_predictor* p) { return p ? p->blob.size() : 0; }

int vpt_predictor_blob_export(const vpt_predictor* p, void* dst, uint64_t capacity) {
    VPT_API_BEGIN
    if (!p || !dst) throw Error(kInvalidArgument, "InvalidArgumentError: predictor/dst: must not be NULL");
    if (capacity < p->blob.size()) throw Error(kInvalidArgument, "InvalidArgumentError: capacity: too small");
    memcpy(dst, p->blob.data(), p->blob.size());
    return kOk;
    VPT_API_END
}

int vpt_predictor_from_blob(const void* blob, uint64_t len, int device, vpt_predictor** out) {
    VPT_API_BEGIN
    if (!blob || !out) throw Error(kInvalidArgument, "InvalidArgumentError: blob/out: must not be NULL");
    *out = nullptr;
    BlobHeader h;
    if (len < sizeof h) throw Error(kInvalidModel, "InvalidModelError: blob too short");
    memcpy(&h, blob, sizeof h);
    if (memcmp(h.magic, kBlobMagic, 8) != 0 || h.total_bytes != len)
        throw Error(kInvalidModel, "InvalidModelError: not a vaporetto_b200 model blob");
    validate_blob_header(h, len);
    std::unique_ptr<vpt_predictor> p(new vpt_predictor());
    p->device = device;
    p->from_blob = true;
    p->hdr = h;
    p->blob.assign(static_cast<const uint8_t*>(blob), static_cast<const uint8_t*>(blob) + len);
    upload(*p);
    *out = p.release();
    return kOk;
    VPT_API_END
}

uint64_t vpt_workspace_size(size_t n_sent) { return workspace_layout(n_sent).total; }

int vpt_device_pci_bus_id(int device, char* buf, size_t capacity) {
    VPT_API_BEGIN
    if (!buf || capacity < 16) throw Error(kInvalidArgument, "InvalidArgumentError: buf: needs at least 16 bytes");
    cuda_check(cudaDeviceGetPCIBusId(buf, int(capacity), device), "cudaDeviceGetPCIBusId");
    return kOk;
    VPT_API_END
}

static void require_device(const vpt_predictor* p) {
    if (!p) throw Error(kInvalidArgument, "InvalidArgumentError: predictor: must not be NULL");
    if (p->device < 0 || !p->d_blob)
        throw Error(kCudaError, "this predictor was created without a CUDA device (device = -1): it cannot score; "
                                "vaporetto_b200 has no CPU fallback");
}

static void run_batch_dev(const vpt_predictor* p, const uint8_t* d_utf8, const uint64_t* d_byte_offsets, size_t n_sent,
                          void* d_workspace, uint64_t workspace_bytes, int32_t* d_scores, uint8_t* d_boundaries,
                          uint64_t* d_bound_offsets, int32_t* d_status, uint32_t* d_char_states,
                          uint32_t* d_type_states, uint64_t* d_char_offsets, void* cuda_stream, float* stage_ms) {
    require_device(p);
    if (stage_ms) stage_ms[0] = stage_ms[1] = stage_ms[2] = 0.f;
    if (n_sent == 0) return;
    if (!d_utf8 || !d_byte_offsets || !d_workspace || !d_scores || !d_boundaries || !d_bound_offsets || !d_status)
        throw Error(kInvalidArgument, "InvalidArgumentError: device buffers: must not be NULL");
    if (workspace_bytes < workspace_layout(n_sent).total)
        throw Error(kInvalidArgument, "InvalidArgumentError: workspace: too small");
    if (reinterpret_cast<uintptr_t>(d_utf8) & 15)
        throw Error(kInvalidArgument, "InvalidArgumentError: d_utf8: must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    BatchArgs a;
    a.text = d_utf8;
    a.offsets = d_byte_offsets;
    a.n_sent = n_sent;
    bind_workspace(a, d_workspace, n_sent);
    a.status = d_status;
    a.scores = d_scores;
    a.boundaries = d_boundaries;
    a.bound_offsets = d_bound_offsets;
    a.char_offsets = d_char_offsets;
    a.char_states = d_char_states;
    a.type_states = d_type_states;
    if (!stage_ms) {
        cuda_check(launch_batch(p->dm, a, st), "launch(batch)");
        return;
    }
    cudaEvent_t ev[4];
    for (auto& e : ev) cuda_check(cudaEventCreate(&e), "cudaEventCreate");
    const bool fused = fused_ok(p->dm);  // one launch: the count and scan stages do not exist
    cuda_check(cudaEventRecord(ev[0], st), "record");
    if (!fused) cuda_check(launch_count_only(a, st), "launch(count)");
    cuda_check(cudaEventRecord(ev[1], st), "record");
    if (!fused) cuda_check(launch_scan_only(a, st), "launch(scan)");
    cuda_check(cudaEventRecord(ev[2], st), "record");
    cuda_check(launch_score(p->dm, a, st), "launch(score)");
    cuda_check(cudaEventRecord(ev[3], st), "record");
    cuda_check(cudaStreamSynchronize(st), "sync(profiled)");
    for (int i = 0; i < 3; ++i) cuda_check(cudaEventElapsedTime(&stage_ms[i], ev[i], ev[i + 1]), "elapsed");
    for (auto& e : ev) cudaEventDestroy(e);
}

int vpt_predict_batch_dev(const vpt_predictor* p, const uint8_t* d_utf8, const uint64_t* d_byte_offsets, size_t n_sent,
                          void* d_workspace, uint64_t workspace_bytes, int32_t* d_scores, uint8_t* d_boundaries,
                          uint64_t* d_bound_offsets, int32_t* d_status, uint32_t* d_char_states,
                          uint32_t* d_type_states, uint64_t* d_char_offsets, void* cuda_stream) {
    VPT_API_BEGIN
    run_batch_dev(p, d_utf8, d_byte_offsets, n_sent, d_workspace, workspace_bytes, d_scores, d_boundaries,
                  d_bound_offsets, d_status, d_char_states, d_type_states, d_char_offsets, cuda_stream, nullptr);
    return kOk;
    VPT_API_END
}

int vpt_predict_batch_dev_profiled(const vpt_predictor* p, const uint8_t* d_utf8, const uint64_t* d_byte_offsets,
                                   size_t n_sent, void* d_workspace, uint64_t workspace_bytes, int32_t* d_scores,
                                   uint8_t* d_boundaries, uint64_t* d_bound_offsets, int32_t* d_status,
                                   uint32_t* d_char_states, uint32_t* d_type_states, uint64_t* d_char_offsets,
                                   void* cuda_stream, float* stage_ms) {
    VPT_API_BEGIN
    if (!stage_ms) throw Error(kInvalidArgument, "InvalidArgumentError: stage_ms: must not be NULL");
    run_batch_dev(p, d_utf8, d_byte_offsets, n_sent, d_workspace, workspace_bytes, d_scores, d_boundaries,
                  d_bound_offsets, d_status, d_char_states, d_type_states, d_char_offsets, cuda_stream, stage_ms);
    return kOk;
    VPT_API_END
}

namespace {

// sentences per pipelined chunk of vpt_predict_batch (tuning knob: env VPT_CHUNK_SENTENCES)
size_t chunk_sentences() {
    static const size_t v = [] {
        const char* e = getenv("VPT_CHUNK_SENTENCES");
        const long long x = e ? atoll(e) : 0;
        return x >= 1024 ? size_t(x) : size_t(262144);
    }();
    return v;
}

// Chunk sizes of a copy/compute/copy pipeline over `total` units: small chunks first (the first copy-in and
// kernels are not overlapped with anything) growing by doubling to `big`, equal chunks of at most `big` in the
// middle, halving again to `small_down` at the end (the last copy-out is not overlapped either).
std::vector<size_t> ramp_schedule(size_t total, size_t big, size_t small_up, size_t small_down) {
    std::vector<size_t> up, down, out;
    size_t sum = 0;
    for (size_t v = std::max<size_t>(small_up, 1); v < big; v *= 2) { up.push_back(v); sum += v; }
    for (size_t v = std::max<size_t>(small_down, 1); v < big; v *= 2) { down.push_back(v); sum += v; }
    if (total <= sum + big) {
        // too small for the full ramps: equal chunks of about a quarter
        const size_t c = std::max<size_t>(std::min(big, (total + 3) / 4), std::min(small_up, big));
        for (size_t lo = 0; lo < total; lo += c) out.push_back(std::min(c, total - lo));
        return out;
    }
    const size_t middle = total - sum;
    const size_t nmid = (middle + big - 1) / big;
    out = up;
    for (size_t i = 0; i < nmid; ++i) out.push_back(middle / nmid + (i < middle % nmid ? 1 : 0));
    out.insert(out.end(), down.rbegin(), down.rend());
    return out;
}

// Pipeline trace (env VPT_TRACE=1): per chunk, CUDA-event times of copy-in end, kernels start / end and copy-out
// end are printed to stderr when the call returns.
bool pipeline_trace() {
    const char* e = getenv("VPT_TRACE");
    return e && *e == '1';
}
struct TraceEvents {
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t sub[4] = {nullptr, nullptr, nullptr, nullptr};  // optional marks between the kernels of a chunk
    void mark_sub(int i, cudaStream_t st) {
        if (!sub[i]) cuda_check(cudaEventCreate(&sub[i]), "cudaEventCreate");
        cuda_check(cudaEventRecord(sub[i], st), "cudaEventRecord");
    }
    void mark(int i, cudaStream_t st) {
        if (!ev[i]) cuda_check(cudaEventCreate(&ev[i]), "cudaEventCreate");
        cuda_check(cudaEventRecord(ev[i], st), "cudaEventRecord");
    }
    void destroy() {
        for (cudaEvent_t& e : ev) if (e) { cudaEventDestroy(e); e = nullptr; }
        for (cudaEvent_t& e : sub) if (e) { cudaEventDestroy(e); e = nullptr; }
    }
    void print(const char* what, size_t c, unsigned long long units, const TraceEvents& first) const {
        float t[4] = {0, 0, 0, 0};
        for (int i = 0; i < 4; ++i)
            if (ev[i] && first.ev[0]) cudaEventElapsedTime(&t[i], first.ev[0], ev[i]);
        fprintf(stderr, "[vpt %s] chunk %zu (%llu): copy-in end %.3f  kernels %.3f..%.3f  copy-out end %.3f ms", what, c,
                units, t[0], t[1], t[2], t[3]);
        for (int i = 0; i < 4; ++i)
            if (sub[i] && first.ev[0]) {
                float x = 0;
                cudaEventElapsedTime(&x, first.ev[0], sub[i]);
                fprintf(stderr, "%s%.3f", i ? " " : "  marks ", x);
            }
        fprintf(stderr, "\n");
    }
};

struct ChunkState {
    size_t s_lo = 0, n = 0;       // sentence range
    uint64_t byte_lo = 0, nbytes = 0;
    cudaEvent_t counted = nullptr;
    TraceEvents tr;
    BatchArgs a;
};

// stage 1 of a chunk: H2D of text + offsets, count + scan, totals to pinned host memory
void chunk_count(Scratch& s, ChunkState& ch, const uint8_t* utf8, const uint64_t* byte_offsets) {
    cudaStream_t st = s.stream;
    const WorkspaceLayout wl = workspace_layout(ch.n);
    const size_t shift = size_t(ch.byte_lo & 15);
    Scratch::ensure(s.d_text, s.text_cap, shift + ch.nbytes + 64);
    Scratch::ensure(s.d_off, s.off_cap, 8 * (ch.n + 1));
    Scratch::ensure(s.d_ws, s.ws_cap, wl.total);
    Scratch::ensure(s.d_status, s.status_cap, 4 * ch.n);
    Scratch::ensure(s.d_boff, s.boff_cap, 8 * (ch.n + 1));
    Scratch::ensure(s.d_coff, s.coff_cap, 8 * (ch.n + 1));
    if (ch.nbytes)
        cuda_check(cudaMemcpyAsync(static_cast<uint8_t*>(s.d_text) + shift, utf8 + ch.byte_lo, ch.nbytes,
                                   cudaMemcpyHostToDevice, st), "H2D(text)");
    cuda_check(cudaMemcpyAsync(s.d_off, byte_offsets + ch.s_lo, 8 * (ch.n + 1), cudaMemcpyHostToDevice, st), "H2D(offsets)");
    // the kernels overwrite buffers the previous chunk of this scratch may still be copying out
    cuda_check(cudaStreamWaitEvent(st, s.ev_out, 0), "cudaStreamWaitEvent");
    if (pipeline_trace()) ch.tr.mark(0, st);
    BatchArgs& a = ch.a;
    a = BatchArgs();
    // offsets are absolute in the caller's buffer: bias the text pointer so that text[offset] is right
    a.text = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(s.d_text) + shift - uintptr_t(ch.byte_lo));
    a.offsets = static_cast<const uint64_t*>(s.d_off);
    a.n_sent = ch.n;
    bind_workspace(a, s.d_ws, ch.n);
    a.status = static_cast<int32_t*>(s.d_status);
    a.bound_offsets = static_cast<uint64_t*>(s.d_boff);
    a.char_offsets = static_cast<uint64_t*>(s.d_coff);
    // the totals come back through pinned host memory the scan kernel writes itself: an 8-byte D2H copy would
    // queue behind the bulk copy-out of earlier chunks on the copy engine
    a.totals_host = &s.h_totals[0];
    cuda_check(launch_count(a, st), "launch(count)");
    if (!ch.counted) cuda_check(cudaEventCreateWithFlags(&ch.counted, cudaEventDisableTiming), "cudaEventCreate");
    cuda_check(cudaEventRecord(ch.counted, st), "cudaEventRecord");
}

}  // namespace

int vpt_predict_batch(const vpt_predictor* p, const uint8_t* utf8, const uint64_t* byte_offsets, size_t n_sent,
                      int32_t* scores_out, uint8_t* boundaries_out, size_t out_capacity, uint64_t* bound_offsets_out,
                      int32_t* status_out, uint32_t* char_states_out, uint32_t* type_states_out,
                      size_t states_capacity, uint64_t* char_offsets_out, uint64_t* n_boundaries_out,
                      uint64_t* n_chars_out) {
    VPT_API_BEGIN
    require_device(p);
    if (n_boundaries_out) *n_boundaries_out = 0;
    if (n_chars_out) *n_chars_out = 0;
    if (!byte_offsets || !bound_offsets_out)
        throw Error(kInvalidArgument, "InvalidArgumentError: byte_offsets/bound_offsets_out: must not be NULL");
    if (n_sent == 0) { bound_offsets_out[0] = 0; if (char_offsets_out) char_offsets_out[0] = 0; return kOk; }
    if (byte_offsets[n_sent] < byte_offsets[0])
        throw Error(kInvalidArgument, "InvalidArgumentError: byte_offsets: must be non-decreasing");
    if (byte_offsets[n_sent] > byte_offsets[0] && !utf8)
        throw Error(kInvalidArgument, "InvalidArgumentError: utf8: must not be NULL");
    cuda_check(cudaSetDevice(p->device), "cudaSetDevice");

    // The batch is cut into chunks that flow through three streams so that the H2D copy of chunk c+2, the
    // kernels of chunk c+1 and the D2H copy of chunk c overlap.  Each chunk is an independent batch on the
    // device; its output offsets are rebased with the running totals (known on the host after its count pass).
    const size_t kChunkSentences = chunk_sentences();
    const std::vector<size_t> sizes = ramp_schedule(n_sent, kChunkSentences, kChunkSentences / 8, kChunkSentences / 4);
    const size_t nchunks = sizes.size();
    constexpr int kDepth = 4;
    std::unique_ptr<ScratchLease> lease[kDepth];
    for (int i = 0; i < kDepth && size_t(i) < nchunks; ++i) lease[i].reset(new ScratchLease(*p));
    std::vector<ChunkState> chunks(nchunks);
    struct EventGuard {
        std::vector<ChunkState>& c;
        ~EventGuard() { for (auto& x : c) { if (x.counted) cudaEventDestroy(x.counted); x.tr.destroy(); } }
    } guard{chunks};
    for (size_t c = 0, lo = 0; c < nchunks; lo += sizes[c], ++c) {
        ChunkState& ch = chunks[c];
        ch.s_lo = lo;
        ch.n = sizes[c];
        ch.byte_lo = byte_offsets[ch.s_lo];
        if (byte_offsets[ch.s_lo + ch.n] < ch.byte_lo)
            throw Error(kInvalidArgument, "InvalidArgumentError: byte_offsets: must be non-decreasing");
        ch.nbytes = byte_offsets[ch.s_lo + ch.n] - ch.byte_lo;
    }
    const bool want_states = char_states_out || type_states_out;
    uint64_t nb_total = 0, nc_total = 0;
    bool overflow = false;
    constexpr size_t kAhead = kDepth - 1;  // chunks whose copy-in + count pass are issued ahead of the scoring
    for (size_t c = 0; c < std::min<size_t>(kAhead, nchunks); ++c) chunk_count(*lease[c % kDepth]->s, chunks[c], utf8, byte_offsets);
    for (size_t c = 0; c < nchunks; ++c) {
        ChunkState& ch = chunks[c];
        Scratch& s = *lease[c % kDepth]->s;
        cudaStream_t st = s.stream;
        cuda_check(cudaEventSynchronize(ch.counted), "sync(count)");
        const uint64_t nb = s.h_totals[0], nc = s.h_totals[1];
        if (nb_total + nb > out_capacity || (nb && !boundaries_out) || (want_states && nc_total + nc > states_capacity))
            overflow = true;
        if (!overflow) {
            BatchArgs& a = ch.a;
            Scratch::ensure(s.d_bounds, s.bounds_cap, nb + 4);
            a.scores = nullptr;  // boundaries only, when the caller wants no scores and the kernel can skip them
            if (scores_out || !scores_optional(p->dm)) {
                Scratch::ensure(s.d_scores, s.scores_cap, 4 * nb + 4);
                a.scores = static_cast<int32_t*>(s.d_scores);
            }
            a.boundaries = static_cast<uint8_t*>(s.d_bounds);
            if (char_states_out) { Scratch::ensure(s.d_cst, s.cst_cap, 4 * nc + 4); a.char_states = static_cast<uint32_t*>(s.d_cst); }
            if (type_states_out) { Scratch::ensure(s.d_tst, s.tst_cap, 4 * nc + 4); a.type_states = static_cast<uint32_t*>(s.d_tst); }
            a.bound_base = nb_total;
            a.char_base = nc_total;
            if (pipeline_trace()) ch.tr.mark(1, st);
            cuda_check(launch_score(p->dm, a, st), "launch(score)");
            if (pipeline_trace()) ch.tr.mark(2, st);
            cuda_check(cudaEventRecord(s.ev_kernels, st), "cudaEventRecord");
            cudaStream_t so = s.stream_out;
            cuda_check(cudaStreamWaitEvent(so, s.ev_kernels, 0), "cudaStreamWaitEvent");
            if (nb) {
                if (scores_out) cuda_check(cudaMemcpyAsync(scores_out + nb_total, s.d_scores, 4 * nb, cudaMemcpyDeviceToHost, so), "D2H(scores)");
                cuda_check(cudaMemcpyAsync(boundaries_out + nb_total, s.d_bounds, nb, cudaMemcpyDeviceToHost, so), "D2H(boundaries)");
            }
            // the last element of a chunk's offsets is the first of the next chunk's: copy n (+1 for the last chunk)
            const size_t noff = ch.n + (c + 1 == nchunks ? 1 : 0);
            cuda_check(cudaMemcpyAsync(bound_offsets_out + ch.s_lo, s.d_boff, 8 * noff, cudaMemcpyDeviceToHost, so), "D2H(offsets)");
            if (char_offsets_out)
                cuda_check(cudaMemcpyAsync(char_offsets_out + ch.s_lo, s.d_coff, 8 * noff, cudaMemcpyDeviceToHost, so), "D2H(offsets)");
            if (status_out) cuda_check(cudaMemcpyAsync(status_out + ch.s_lo, s.d_status, 4 * ch.n, cudaMemcpyDeviceToHost, so), "D2H(status)");
            if (nc && char_states_out) cuda_check(cudaMemcpyAsync(char_states_out + nc_total, s.d_cst, 4 * nc, cudaMemcpyDeviceToHost, so), "D2H(states)");
            if (nc && type_states_out) cuda_check(cudaMemcpyAsync(type_states_out + nc_total, s.d_tst, 4 * nc, cudaMemcpyDeviceToHost, so), "D2H(states)");
            cuda_check(cudaEventRecord(s.ev_out, so), "cudaEventRecord");
            if (pipeline_trace()) ch.tr.mark(3, so);
        }
        nb_total += nb;
        nc_total += nc;
        if (c + kAhead < nchunks) chunk_count(*lease[(c + kAhead) % kDepth]->s, chunks[c + kAhead], utf8, byte_offsets);
    }
    for (int i = 0; i < kDepth; ++i)
        if (lease[i]) {
            cuda_check(cudaStreamSynchronize(lease[i]->s->stream), "sync(score)");
            cuda_check(cudaStreamSynchronize(lease[i]->s->stream_out), "sync(copy-out)");
        }
    if (pipeline_trace())
        for (size_t c = 0; c < nchunks; ++c) chunks[c].tr.print("batch", c, chunks[c].n, chunks[0].tr);
    if (n_boundaries_out) *n_boundaries_out = nb_total;
    if (n_chars_out) *n_chars_out = nc_total;
    if (overflow) throw Error(kInvalidArgument, "InvalidArgumentError: out_capacity/states_capacity: too small for the batch");
    return kOk;
    VPT_API_END
}

namespace {

// bytes per pipelined chunk of vpt_tokenize_lines (tuning knob: env VPT_CHUNK_BYTES)
size_t chunk_bytes() {
    const char* e = getenv("VPT_CHUNK_BYTES");
    const long long x = e ? atoll(e) : 0;
    return x >= 64 ? size_t(x) : size_t(16) << 20;
}

constexpr uint64_t kMaxLineChunk = uint64_t(1) << 30;  // 32-bit group-local output offsets (3 bytes out per byte in)

struct LineChunk {
    uint64_t byte_lo = 0, nbytes = 0;
    uint64_t n_lines = 0;
    cudaEvent_t split = nullptr, done = nullptr;
    TraceEvents tr;
    SplitArgs sp;
};

// stage 0 of a chunk: H2D of the text, newline counts, the number of lines to pinned host memory
void lines_stage0(Scratch& s, LineChunk& ch, const uint8_t* utf8) {
    cudaStream_t st = s.stream;
    const size_t nblk = (ch.nbytes + kSplitBlockBytes - 1) / kSplitBlockBytes;
    Scratch::ensure(s.d_text, s.text_cap, ch.nbytes + 64);
    Scratch::ensure(s.d_blk, s.blk_cap, 4 * nblk + 4);
    Scratch::ensure(s.d_blkbase, s.blkbase_cap, 8 * nblk + 16);
    cuda_check(cudaMemcpyAsync(s.d_text, utf8 + ch.byte_lo, ch.nbytes, cudaMemcpyHostToDevice, st), "H2D(text)");
    if (pipeline_trace()) ch.tr.mark(0, st);
    SplitArgs& sp = ch.sp;
    sp = SplitArgs();
    sp.text = static_cast<const uint8_t*>(s.d_text);
    sp.n_bytes = ch.nbytes;
    sp.blk = static_cast<uint32_t*>(s.d_blk);
    sp.blk_base = static_cast<uint64_t*>(s.d_blkbase);
    sp.n_lines = sp.blk_base + nblk;  // the element after the per-block bases
    sp.n_lines_host = &s.h_totals[2];  // read back through pinned host memory written by the kernel (see chunk_count)
    cuda_check(launch_split_count(sp, st), "launch(split)");
    if (!ch.split) cuda_check(cudaEventCreateWithFlags(&ch.split, cudaEventDisableTiming), "cudaEventCreate");
    cuda_check(cudaEventRecord(ch.split, st), "cudaEventRecord");
}

// stage 1: line offsets, count + score, tokenised bytes; the output size to pinned host memory
void lines_stage1(const vpt_predictor& p, Scratch& s, LineChunk& ch, bool normalize, uint32_t wsconst, bool tags) {
    cudaStream_t st = s.stream;
    cuda_check(cudaEventSynchronize(ch.split), "sync(split)");
    const size_t n = size_t(s.h_totals[2]);
    ch.n_lines = n;
    if (!ch.done) cuda_check(cudaEventCreateWithFlags(&ch.done, cudaEventDisableTiming), "cudaEventCreate");
    s.h_totals[3] = 0;
    if (n == 0) { cuda_check(cudaEventRecord(ch.done, st), "cudaEventRecord"); return; }
    const WorkspaceLayout wl = workspace_layout(n);
    const size_t ng = (n + kGroup - 1) / kGroup;
    Scratch::ensure(s.d_off, s.off_cap, 8 * (n + 1));
    Scratch::ensure(s.d_trims, s.trims_cap, n + 4);
    Scratch::ensure(s.d_ws, s.ws_cap, wl.total);
    Scratch::ensure(s.d_status, s.status_cap, 4 * n);
    Scratch::ensure(s.d_boff, s.boff_cap, 8 * (n + 1));
    Scratch::ensure(s.d_bounds, s.bounds_cap, ch.nbytes + 4);
    Scratch::ensure(s.d_tokg, s.tokg_cap, 8 * (ng + 2));
    // surface bytes + at most one '\\' per byte + at most one ' ' per character + one '\n' per line
    // (with tags: every token -- at most one per byte -- may get the longest "/tag/.." suffix of the model)
    Scratch::ensure(s.d_out, s.out_cap, 3 * ch.nbytes + n + 4 + (tags ? size_t(ch.nbytes) * p.dt.max_suffix : 0));
    ch.sp.offsets = static_cast<uint64_t*>(s.d_off);
    ch.sp.trims = static_cast<uint8_t*>(s.d_trims);
    // the kernels overwrite d_out, which the previous chunk of this scratch may still be copying out
    cuda_check(cudaStreamWaitEvent(st, s.ev_out, 0), "cudaStreamWaitEvent");
    if (pipeline_trace()) ch.tr.mark(1, st);
    cuda_check(launch_split_write(ch.sp, st), "launch(split)");
    BatchArgs a;
    a.text = ch.sp.text;
    a.offsets = ch.sp.offsets;
    a.trims = ch.sp.trims;
    a.n_sent = n;
    bind_workspace(a, s.d_ws, n);
    a.status = static_cast<int32_t*>(s.d_status);
    a.bound_offsets = static_cast<uint64_t*>(s.d_boff);
    // the tokenised text needs the boundaries only (every character is at least one byte: the chunk's bytes
    // bound its boundaries)
    a.scores = nullptr;
    DevModel dm = p.dm;
    dm.kytea_norm = normalize ? 1 : 0;
    if (!scores_optional(dm)) {
        Scratch::ensure(s.d_scores, s.scores_cap, 4 * ch.nbytes + 4);
        a.scores = static_cast<int32_t*>(s.d_scores);
    }
    a.boundaries = static_cast<uint8_t*>(s.d_bounds);
    if (tags) {
        // tag prediction needs the pattern-id states and the character offsets of the sentences
        Scratch::ensure(s.d_cst, s.cst_cap, 4 * ch.nbytes + 16);
        Scratch::ensure(s.d_tst, s.tst_cap, 4 * ch.nbytes + 16);
        Scratch::ensure(s.d_coff, s.coff_cap, 8 * (n + 1));
        a.char_states = static_cast<uint32_t*>(s.d_cst);
        a.type_states = static_cast<uint32_t*>(s.d_tst);
        a.char_offsets = static_cast<uint64_t*>(s.d_coff);
    }
    if (pipeline_trace()) ch.tr.mark_sub(0, st);  // after the line offsets
    if (!fused_ok(dm)) cuda_check(launch_count(a, st), "launch(count)");
    if (pipeline_trace()) ch.tr.mark_sub(1, st);  // after count + scan (none when the scoring launch is fused)
    cuda_check(launch_score(dm, a, st), "launch(score)");
    if (pipeline_trace()) ch.tr.mark_sub(2, st);  // after the scoring kernel
    TokArgs t;
    t.text = a.text;
    t.offsets = a.offsets;
    t.trims = a.trims;
    t.n_sent = n;
    t.status = a.status;
    t.n_chars = a.n_chars;
    t.boundaries = a.boundaries;
    t.bound_offsets = a.bound_offsets;
    t.tok_state = static_cast<uint64_t*>(s.d_tokg);
    t.ticket = reinterpret_cast<uint32_t*>(t.tok_state + ng);
    t.total = t.tok_state + ng + 1;
    t.total_host = &s.h_totals[3];
    t.out = static_cast<uint8_t*>(s.d_out);
    cuda_check(launch_wsconst(t, a.boundaries, wsconst, normalize, st), "launch(wsconst)");
    if (wsconst & 0x80u) cuda_check(launch_grapheme(t, a.boundaries, normalize, st), "launch(grapheme)");
    if (tags) {
        // tokens per sentence and their prefix, then the tag prediction into per-token records (the post-filters ran:
        // fill_tags sees the final boundaries, predict/src/main.rs:157-160)
        CompactArgs k;
        k.n_sent = n;
        k.status = a.status;
        k.n_chars = a.n_chars;
        k.boundaries = a.boundaries;
        k.bound_offsets = a.bound_offsets;
        k.n_bound = 0;  // no bit stream on this path
        Scratch::ensure(s.d_st8, s.st8_cap, n + 16);
        Scratch::ensure(s.d_ntok, s.ntok_cap, 4 * n + 16);
        Scratch::ensure(s.d_tokbase, s.tokbase_cap, 8 * (n + 1) + 16);
        Scratch::ensure(s.d_toklocal, s.toklocal_cap, 4 * n + 16);
        Scratch::ensure(s.d_tokblk, s.tokblk_cap, 8 * (n / 256 + 4));
        k.status8 = static_cast<uint8_t*>(s.d_st8);
        k.n_tokens = static_cast<uint32_t*>(s.d_ntok);
        k.tok_base = static_cast<uint64_t*>(s.d_tokbase);
        k.tok_local = static_cast<uint32_t*>(s.d_toklocal);
        k.tok_blk = static_cast<uint64_t*>(s.d_tokblk);
        cuda_check(launch_compact(k, st), "launch(compact)");
        Scratch::ensure(s.d_tok, s.tok_cap, 4 * ch.nbytes + 16);
        Scratch::ensure(s.d_cand, s.cand_cap, ch.nbytes * std::max<size_t>(p.n_tags, 1) + 16);
        Scratch::ensure(s.d_tokdesc, s.tokdesc_cap, 16 * ch.nbytes + 16);
        TagArgs g;
        g.text = a.text;
        g.offsets = a.offsets;
        g.trims = a.trims;
        g.n_sent = n;
        g.status = a.status;
        g.boundaries = a.boundaries;
        g.bound_offsets = a.bound_offsets;
        g.char_offsets = a.char_offsets;
        g.char_states = p.dt.char_rels ? a.char_states : nullptr;
        g.type_states = p.dt.type_rels ? a.type_states : nullptr;
        g.tok_base = k.tok_base;
        g.tok_ids = static_cast<int32_t*>(s.d_tok);
        g.tok_cands = static_cast<uint8_t*>(s.d_cand);
        g.tok_desc = static_cast<uint4*>(s.d_tokdesc);
        g.max_tokens = ch.nbytes;
        Scratch::ensure(s.d_tokwork, s.tokwork_cap, 4 * ch.nbytes + 32);
        g.tok_work = static_cast<uint32_t*>(s.d_tokwork);
        g.norm = normalize ? 1 : 0;
        cuda_check(launch_tags(p.dt, g, st), "launch(tags)");
        t.tok_base = k.tok_base;
        t.tok_ids = g.tok_ids;
        t.tok_cands = g.tok_cands;
        t.n_tags = uint32_t(p.n_tags);
        t.ts_slot = p.dt.ts_slot;
        t.ts_cand = p.dt.ts_cand;
        t.ts_ref = p.dt.ts_ref;
        t.ts_bytes = p.dt.ts_bytes;
    }
    cuda_check(launch_tokenize(t, st), "launch(tok)");
    if (pipeline_trace()) ch.tr.mark(2, st);
    cuda_check(cudaEventRecord(ch.done, st), "cudaEventRecord");
}

}  // namespace

uint32_t vpt_kytea_fullwidth(uint32_t code_point) { return kytea_fullwidth(code_point); }

namespace {
int tokenize_lines_impl(const vpt_predictor* p, const uint8_t* utf8, size_t n_bytes, int no_norm, uint32_t wsconst_types, bool tags,
                        uint8_t* out, size_t out_capacity, uint64_t* out_len, uint64_t* n_lines_out) {
    VPT_API_BEGIN
    require_device(p);
    if (tags) {
        if (!p->predict_tags || p->from_blob)
            throw Error(kInvalidArgument, "InvalidArgumentError: this predictor is created with predict_tags = false");
        if (p->n_tags && !p->dt.tok_tab)
            throw Error(kUnsupported, "this tag model exceeds the limits of the device path (tags.hpp); use vpt_fill_tags");
        if (p->n_tags == 0) tags = false;  // predictor.rs:553-555: nothing to predict
    }
    if (out_len) *out_len = 0;
    if (n_lines_out) *n_lines_out = 0;
    if (wsconst_types & ~0xFEu)
        throw Error(kInvalidArgument, "InvalidArgumentError: wsconst_types: bits 1..6 (Digit .. Other) and 7 (grapheme clusters) only");
    if (n_bytes && !utf8) throw Error(kInvalidArgument, "InvalidArgumentError: utf8: must not be NULL");
    if (n_bytes == 0) return kOk;
    cuda_check(cudaSetDevice(p->device), "cudaSetDevice");

    // cut the buffer into chunks that end after a '\n' (memrchr from the nominal cut; a line longer than a
    // chunk extends it to the line's end)
    const size_t big = chunk_bytes();
    const std::vector<size_t> sizes = ramp_schedule(n_bytes, big, big / 8, big / 8);
    std::vector<LineChunk> chunks;
    size_t cum = 0, k = 0;
    for (size_t lo = 0; lo < n_bytes;) {
        // every chunk ends at the next nominal cut of the schedule (a line that ran past cuts skips them)
        do { cum = k < sizes.size() ? cum + sizes[k++] : n_bytes; } while (cum <= lo);
        size_t hi = std::min(n_bytes, cum);
        if (hi < n_bytes) {
            const void* q = memrchr(utf8 + lo, 0x0A, hi - lo);
            if (q) hi = size_t(static_cast<const uint8_t*>(q) - utf8) + 1;
            else {
                const void* f = memchr(utf8 + hi, 0x0A, n_bytes - hi);
                hi = f ? size_t(static_cast<const uint8_t*>(f) - utf8) + 1 : n_bytes;
            }
        }
        if (hi - lo > kMaxLineChunk) throw Error(kInvalidArgument, "InvalidArgumentError: utf8: a line is longer than 1 GiB");
        LineChunk ch;
        ch.byte_lo = lo;
        ch.nbytes = hi - lo;
        chunks.push_back(ch);
        lo = hi;
    }
    const size_t nchunks = chunks.size();
    constexpr int kDepth = 4;
    std::unique_ptr<ScratchLease> lease[kDepth];
    for (int i = 0; i < kDepth && size_t(i) < nchunks; ++i) lease[i].reset(new ScratchLease(*p));
    struct EventGuard {
        std::vector<LineChunk>& c;
        ~EventGuard() {
            for (auto& x : c) {
                if (x.split) cudaEventDestroy(x.split);
                if (x.done) cudaEventDestroy(x.done);
                x.tr.destroy();
            }
        }
    } guard{chunks};

    // chunks c+2, c+3 are copied in and split while chunk c+1 is scored and chunk c is copied out
    uint64_t total = 0, lines = 0;
    bool overflow = false;
    const bool trace = pipeline_trace();
    const auto host_t0 = std::chrono::steady_clock::now();
    auto host_ms = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
    for (size_t c = 0; c < std::min<size_t>(3, nchunks); ++c) lines_stage0(*lease[c % kDepth]->s, chunks[c], utf8);
    lines_stage1(*p, *lease[0]->s, chunks[0], no_norm == 0, wsconst_types, tags);
    for (size_t c = 0; c < nchunks; ++c) {
        const double h0 = host_ms();
        if (c + 3 < nchunks) lines_stage0(*lease[(c + 3) % kDepth]->s, chunks[c + 3], utf8);
        const double h1 = host_ms();
        if (c + 1 < nchunks) lines_stage1(*p, *lease[(c + 1) % kDepth]->s, chunks[c + 1], no_norm == 0, wsconst_types, tags);
        const double h2 = host_ms();
        Scratch& s = *lease[c % kDepth]->s;
        cuda_check(cudaEventSynchronize(chunks[c].done), "sync(tokenize)");
        if (trace)
            fprintf(stderr, "[vpt lines host] iteration %zu: begins %.3f  copy-in+split issued %.3f  kernels issued %.3f  "
                            "chunk done seen %.3f ms\n", c, h0, h1, h2, host_ms());
        const uint64_t nb = s.h_totals[3];
        if (total + nb > out_capacity || (nb && !out)) overflow = true;
        if (!overflow && nb) {
            cuda_check(cudaMemcpyAsync(out + total, s.d_out, nb, cudaMemcpyDeviceToHost, s.stream_out), "D2H(text)");
            cuda_check(cudaEventRecord(s.ev_out, s.stream_out), "cudaEventRecord");
        }
        if (pipeline_trace()) chunks[c].tr.mark(3, s.stream_out);
        total += nb;
        lines += chunks[c].n_lines;
    }
    for (int i = 0; i < kDepth; ++i)
        if (lease[i]) {
            cuda_check(cudaStreamSynchronize(lease[i]->s->stream), "sync(tokenize)");
            cuda_check(cudaStreamSynchronize(lease[i]->s->stream_out), "sync(copy-out)");
        }
    if (pipeline_trace())
        for (size_t c = 0; c < nchunks; ++c) chunks[c].tr.print("lines", c, chunks[c].nbytes, chunks[0].tr);
    if (out_len) *out_len = total;
    if (n_lines_out) *n_lines_out = lines;
    if (overflow) throw Error(kInvalidArgument, "InvalidArgumentError: out_capacity: too small for the tokenized text");
    return kOk;
    VPT_API_END
}
}  // namespace

int vpt_tokenize_lines(const vpt_predictor* p, const uint8_t* utf8, size_t n_bytes, int no_norm, uint32_t wsconst_types,
                       uint8_t* out, size_t out_capacity, uint64_t* out_len, uint64_t* n_lines_out) {
    return tokenize_lines_impl(p, utf8, n_bytes, no_norm, wsconst_types, false, out, out_capacity, out_len, n_lines_out);
}

int vpt_tokenize_lines_tags(const vpt_predictor* p, const uint8_t* utf8, size_t n_bytes, int no_norm, uint32_t wsconst_types,
                            uint8_t* out, size_t out_capacity, uint64_t* out_len, uint64_t* n_lines_out) {
    return tokenize_lines_impl(p, utf8, n_bytes, no_norm, wsconst_types, true, out, out_capacity, out_len, n_lines_out);
}

namespace {

constexpr size_t kSingleMaxBytes = 2048;   // sentences up to this size take the one-round-trip path of vpt_predict
constexpr size_t kSingleIoBytes = 65536;

// Predictor::predict for ONE sentence (reference predictor.rs:518-543) with one launch and no copy calls: the text and the
// offsets are written into pinned host memory the kernel reads directly (zero-copy over PCIe; the tile loader's bulk copy
// takes a host address like any other), the results are stored straight into the same pinned block, and the look-back
// words + ticket live in a small device block that the kernel itself leaves zeroed (BatchArgs::self_clean).  The call is
// a memcpy into the pinned block, one kernel launch, one stream synchronisation.  (The batch pipeline costs several copies,
// a memset node and three synchronisations per call.)
bool predict_single_fast(const vpt_predictor* p, const uint8_t* utf8, size_t n_bytes, int32_t* scores_out, uint8_t* boundaries_out,
                         size_t out_capacity, uint32_t* char_states_out, uint32_t* type_states_out, size_t states_capacity,
                         uint64_t* n_chars_out) {
    if (n_bytes == 0 || n_bytes > kSingleMaxBytes || !fused_ok(p->dm) || (p->dm.ct.present && p->dm.ct.has_overflow)) return false;
    cuda_check(cudaSetDevice(p->device), "cudaSetDevice");
    ScratchLease lease(*p);
    Scratch& s = *lease.s;
    if (!s.h_io) {
        cuda_check(cudaHostAlloc(reinterpret_cast<void**>(&s.h_io), kSingleIoBytes, cudaHostAllocMapped), "cudaHostAlloc");
        cuda_check(cudaMalloc(&s.d_io, 256), "cudaMalloc(single)");
        cuda_check(cudaMemset(s.d_io, 0, 256), "cudaMemset(single)");
    }
    // layout of the pinned block: input part first, then the outputs
    const size_t o_text = 0;                                   // text, then 64 readable bytes
    const size_t o_off = align_up(n_bytes + 64, 16);           // u64 offsets[2]
    const size_t in_bytes = o_off + 16;
    const size_t o_boff = align_up(in_bytes, 16);              // u64 bound_offsets[2]
    const size_t o_coff = o_boff + 16;                         // u64 char_offsets[2]
    const size_t o_stat = o_coff + 16;                         // i32 status, u32 n_chars
    const size_t o_bnd = o_stat + 16;                          // u8 boundaries
    const size_t o_sc = align_up(o_bnd + n_bytes, 16);         // i32 scores
    const size_t o_cst = o_sc + 4 * n_bytes;                   // u32 states
    const size_t o_tst = o_cst + 4 * n_bytes;
    const size_t end = o_tst + 4 * n_bytes;
    if (end > kSingleIoBytes) return false;
    memcpy(s.h_io + o_text, utf8, n_bytes);
    memset(s.h_io + o_text + n_bytes, 0, o_off - n_bytes);
    uint64_t offs[2] = {0, n_bytes};
    memcpy(s.h_io + o_off, offs, 16);
    cudaStream_t st = s.stream;
    uint8_t* hd = nullptr;  // the device's address of the pinned block (the same value under unified addressing)
    cuda_check(cudaHostGetDevicePointer(reinterpret_cast<void**>(&hd), s.h_io, 0), "cudaHostGetDevicePointer");
    uint8_t* d = static_cast<uint8_t*>(s.d_io);
    BatchArgs a;
    a.text = hd + o_text;
    a.offsets = reinterpret_cast<const uint64_t*>(hd + o_off);
    a.n_sent = 1;
    a.group_bound = reinterpret_cast<uint64_t*>(d);
    a.group_char = reinterpret_cast<uint64_t*>(d + 64);
    a.ticket = reinterpret_cast<uint32_t*>(d + 128);
    a.prezeroed = true;
    a.self_clean = true;
    a.n_chars = reinterpret_cast<uint32_t*>(hd + o_stat + 4);
    a.status = reinterpret_cast<int32_t*>(hd + o_stat);
    a.bound_offsets = reinterpret_cast<uint64_t*>(hd + o_boff);
    a.char_offsets = reinterpret_cast<uint64_t*>(hd + o_coff);
    a.boundaries = hd + o_bnd;
    a.scores = scores_out ? reinterpret_cast<int32_t*>(hd + o_sc) : nullptr;
    const bool want_states = char_states_out || type_states_out;
    if (want_states) {
        a.char_states = reinterpret_cast<uint32_t*>(hd + o_cst);
        a.type_states = reinterpret_cast<uint32_t*>(hd + o_tst);
    }
    cuda_check(launch_fused(p->dm, a, st), "launch(single)");
    cuda_check(cudaStreamSynchronize(st), "sync(single)");
    int32_t status;
    uint32_t n;
    memcpy(&status, s.h_io + o_stat, 4);
    memcpy(&n, s.h_io + o_stat + 4, 4);
    if (n_chars_out) *n_chars_out = n;
    if (status != 0) throw Error(kInternal, "internal error: device validation disagrees with host validation");
    const size_t nb = n > 0 ? n - 1 : 0;
    if (nb > out_capacity || (nb && !boundaries_out) || (want_states && n > states_capacity))
        throw Error(kInvalidArgument, "InvalidArgumentError: out_capacity/states_capacity: too small for the batch");
    if (nb) {
        memcpy(boundaries_out, s.h_io + o_bnd, nb);
        if (scores_out) memcpy(scores_out, s.h_io + o_sc, 4 * nb);
    }
    if (char_states_out) memcpy(char_states_out, s.h_io + o_cst, 4 * size_t(n));
    if (type_states_out) memcpy(type_states_out, s.h_io + o_tst, 4 * size_t(n));
    return true;
}

}  // namespace

int vpt_predict(const vpt_predictor* p, const uint8_t* utf8, size_t n_bytes, int32_t* scores_out, uint8_t* boundaries_out,
                size_t out_capacity, uint32_t* char_states_out, uint32_t* type_states_out, size_t states_capacity,
                uint64_t* n_chars_out) {
    VPT_API_BEGIN
    if (!p) throw Error(kInvalidArgument, "InvalidArgumentError: predictor: must not be NULL");
    if (n_bytes && !utf8) throw Error(kInvalidArgument, "InvalidArgumentError: utf8: must not be NULL");
    check_raw_text(utf8, n_bytes);
    require_device(p);
    if (predict_single_fast(p, utf8, n_bytes, scores_out, boundaries_out, out_capacity, char_states_out, type_states_out,
                            states_capacity, n_chars_out))
        return kOk;
    const uint64_t offs[2] = {0, n_bytes};
    uint64_t boff[2], nb = 0, nc = 0;
    int32_t status = 0;
    uint8_t dummy = 0;
    int rc = vpt_predict_batch(p, utf8, offs, 1, scores_out, boundaries_out ? boundaries_out : &dummy, out_capacity, boff,
                               &status, char_states_out, type_states_out, states_capacity, nullptr, &nb, &nc);
    if (n_chars_out) *n_chars_out = nc;
    if (rc != kOk) return rc;
    if (status != 0) throw Error(kInternal, "internal error: device validation disagrees with host validation");
    return kOk;
    VPT_API_END
}

int vpt_char_types(const uint8_t* utf8, size_t n_bytes, uint8_t* types_out, size_t capacity, uint64_t* n_chars_out) {
    VPT_API_BEGIN
    if (n_bytes && !utf8) throw Error(kInvalidArgument, "InvalidArgumentError: utf8: must not be NULL");
    check_raw_text(utf8, n_bytes);
    std::vector<uint32_t> cps = utf8_to_codepoints(std::string(reinterpret_cast<const char*>(utf8), n_bytes));
    if (n_chars_out) *n_chars_out = cps.size();
    if (cps.size() > capacity) throw Error(kInvalidArgument, "InvalidArgumentError: capacity: too small");
    for (size_t i = 0; i < cps.size(); ++i) types_out[i] = host_char_type(cps[i]);
    return kOk;
    VPT_API_END
}

int vpt_split_linebreaks(const uint8_t* utf8, size_t n_bytes, uint8_t* boundaries, size_t n_boundaries) {
    VPT_API_BEGIN
    if (n_bytes && !utf8) throw Error(kInvalidArgument, "InvalidArgumentError: utf8: must not be NULL");
    check_raw_text(utf8, n_bytes);
    const std::vector<uint32_t> cps = utf8_to_codepoints(std::string(reinterpret_cast<const char*>(utf8), n_bytes));
    if (cps.size() != n_boundaries + 1 || (n_boundaries && !boundaries))
        throw Error(kInvalidArgument, "InvalidArgumentError: boundaries: one per pair of adjacent characters");
    auto lb = [](uint32_t c) { return c == 0x0D || c == 0x0A; };
    for (size_t i = 0; i + 1 < cps.size(); ++i)
        if (lb(cps[i]) || lb(cps[i + 1])) boundaries[i] = 1;
    return kOk;
    VPT_API_END
}

int vpt_concat_grapheme_clusters(const uint8_t* utf8, size_t n_bytes, uint8_t* boundaries, size_t n_boundaries) {
    VPT_API_BEGIN
    if (n_bytes && !utf8) throw Error(kInvalidArgument, "InvalidArgumentError: utf8: must not be NULL");
    check_raw_text(utf8, n_bytes);
    const std::vector<uint32_t> cps = utf8_to_codepoints(std::string(reinterpret_cast<const char*>(utf8), n_bytes));
    if (cps.size() != n_boundaries + 1 || (n_boundaries && !boundaries))
        throw Error(kInvalidArgument, "InvalidArgumentError: boundaries: one per pair of adjacent characters");
    GraphemeState st;
    for (size_t i = 0; i < cps.size(); ++i) {
        const bool brk = grapheme_step(st, grapheme_class(kGraphemeTable, cps[i]));
        if (!brk && i > 0) boundaries[i - 1] = 0;
    }
    return kOk;
    VPT_API_END
}

uint32_t vpt_tag_n_tokens(const vpt_predictor* p) { return p ? uint32_t(p->tag_preds.size()) : 0; }

const char* vpt_tag_string(const vpt_predictor* p, uint32_t token_id, uint32_t slot, uint32_t cand) {
    if (!p || token_id >= p->tag_preds.size()) return nullptr;
    const auto& t = p->tag_preds[token_id].tags;
    if (slot >= t.size() || cand >= t[slot].size()) return nullptr;
    return t[slot][cand].c_str();
}

uint32_t vpt_tag_n_candidates(const vpt_predictor* p, uint32_t token_id, uint32_t slot) {
    if (!p || token_id >= p->tag_preds.size()) return 0;
    const auto& t = p->tag_preds[token_id].tags;
    return slot < t.size() ? uint32_t(t[slot].size()) : 0;
}

uint32_t vpt_tag_score_len(const vpt_predictor* p, uint32_t token_id) {
    if (!p || token_id >= p->tag_preds.size()) return 0;
    return uint32_t(p->tag_preds[token_id].bias.size());
}

int vpt_predict_tags_batch_dev(const vpt_predictor* p, const uint8_t* d_utf8, const uint64_t* d_byte_offsets, size_t n_sent,
                               const int32_t* d_status, const uint8_t* d_boundaries, const uint64_t* d_bound_offsets,
                               const uint64_t* d_char_offsets, const uint32_t* d_char_states, const uint32_t* d_type_states,
                               int32_t* d_tag_token, int32_t* d_tag_cand, uint32_t* d_unserved, void* cuda_stream) {
    VPT_API_BEGIN
    require_device(p);
    if (!p->predict_tags || p->from_blob)
        throw Error(kInvalidArgument, "InvalidArgumentError: this predictor is created with predict_tags = false");
    if (n_sent == 0) return kOk;
    if (!p->dt.tok_tab)
        throw Error(kUnsupported, "this tag model exceeds the limits of the device path (tags.hpp); use vpt_fill_tags");
    if (!d_utf8 || !d_byte_offsets || !d_status || !d_boundaries || !d_bound_offsets || !d_char_offsets || !d_tag_token || !d_tag_cand)
        throw Error(kInvalidArgument, "InvalidArgumentError: device buffers: must not be NULL");
    if ((p->dt.char_rels && !d_char_states) || (p->dt.type_rels && !d_type_states))
        throw Error(kInvalidArgument, "InvalidArgumentError: states: required for tag prediction");
    TagArgs t;
    t.text = d_utf8;
    t.offsets = d_byte_offsets;
    t.n_sent = n_sent;
    t.status = d_status;
    t.boundaries = d_boundaries;
    t.bound_offsets = d_bound_offsets;
    t.char_offsets = d_char_offsets;
    t.char_states = p->dt.char_rels ? d_char_states : nullptr;
    t.type_states = p->dt.type_rels ? d_type_states : nullptr;
    t.tag_token = d_tag_token;
    t.tag_cand = d_tag_cand;
    t.n_unserved = d_unserved;
    cuda_check(launch_tags(p->dt, t, static_cast<cudaStream_t>(cuda_stream)), "launch(tags)");
    return kOk;
    VPT_API_END
}

int vpt_predict_batch_tags(const vpt_predictor* p, const uint8_t* utf8, const uint64_t* byte_offsets, size_t n_sent,
                           int32_t* scores_out, uint8_t* boundaries_out, size_t out_capacity, uint64_t* bound_offsets_out,
                           int32_t* status_out, int32_t* tag_token_out, int32_t* tag_cand_out, size_t chars_capacity,
                           uint64_t* char_offsets_out, uint64_t* n_boundaries_out, uint64_t* n_chars_out,
                           uint64_t* n_unserved_out) {
    VPT_API_BEGIN
    require_device(p);
    if (n_boundaries_out) *n_boundaries_out = 0;
    if (n_chars_out) *n_chars_out = 0;
    if (n_unserved_out) *n_unserved_out = 0;
    if (!p->predict_tags || p->from_blob)
        throw Error(kInvalidArgument, "InvalidArgumentError: this predictor is created with predict_tags = false");
    if (!p->dt.tok_tab)
        throw Error(kUnsupported, "this tag model exceeds the limits of the device path (tags.hpp); use vpt_fill_tags");
    if (!byte_offsets || !bound_offsets_out || !char_offsets_out || !tag_token_out || !tag_cand_out || !boundaries_out)
        throw Error(kInvalidArgument, "InvalidArgumentError: output buffers: must not be NULL");
    if (n_sent == 0) { bound_offsets_out[0] = 0; char_offsets_out[0] = 0; return kOk; }
    if (byte_offsets[n_sent] < byte_offsets[0])
        throw Error(kInvalidArgument, "InvalidArgumentError: byte_offsets: must be non-decreasing");
    cuda_check(cudaSetDevice(p->device), "cudaSetDevice");
    // one chunk: text in, scoring with the pattern-id states kept on the device, tag prediction, results out
    ScratchLease lease(*p);
    Scratch& s = *lease.s;
    cudaStream_t st = s.stream;
    const uint64_t byte_lo = byte_offsets[0], nbytes = byte_offsets[n_sent] - byte_lo;
    const size_t shift = size_t(byte_lo & 15);
    const WorkspaceLayout wl = workspace_layout(n_sent);
    const size_t nt = p->n_tags;
    Scratch::ensure(s.d_text, s.text_cap, shift + nbytes + 64);
    Scratch::ensure(s.d_off, s.off_cap, 8 * (n_sent + 1));
    Scratch::ensure(s.d_ws, s.ws_cap, wl.total);
    Scratch::ensure(s.d_status, s.status_cap, 4 * n_sent);
    Scratch::ensure(s.d_boff, s.boff_cap, 8 * (n_sent + 1));
    Scratch::ensure(s.d_coff, s.coff_cap, 8 * (n_sent + 1));
    Scratch::ensure(s.d_scores, s.scores_cap, 4 * nbytes + 16);      // a character has at least one byte
    Scratch::ensure(s.d_bounds, s.bounds_cap, nbytes + 16);
    Scratch::ensure(s.d_cst, s.cst_cap, 4 * nbytes + 16);
    Scratch::ensure(s.d_tst, s.tst_cap, 4 * nbytes + 16);
    Scratch::ensure(s.d_tok, s.tok_cap, 4 * nbytes + 16);
    Scratch::ensure(s.d_cand, s.cand_cap, 4 * nbytes * nt + 16);
    if (nbytes)
        cuda_check(cudaMemcpyAsync(static_cast<uint8_t*>(s.d_text) + shift, utf8 + byte_lo, nbytes, cudaMemcpyHostToDevice, st), "H2D(text)");
    cuda_check(cudaMemcpyAsync(s.d_off, byte_offsets, 8 * (n_sent + 1), cudaMemcpyHostToDevice, st), "H2D(offsets)");
    BatchArgs a;
    a.text = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(s.d_text) + shift - uintptr_t(byte_lo));
    a.offsets = static_cast<const uint64_t*>(s.d_off);
    a.n_sent = n_sent;
    bind_workspace(a, s.d_ws, n_sent);
    a.status = static_cast<int32_t*>(s.d_status);
    a.bound_offsets = static_cast<uint64_t*>(s.d_boff);
    a.char_offsets = static_cast<uint64_t*>(s.d_coff);
    a.scores = static_cast<int32_t*>(s.d_scores);
    a.boundaries = static_cast<uint8_t*>(s.d_bounds);
    a.char_states = static_cast<uint32_t*>(s.d_cst);
    a.type_states = static_cast<uint32_t*>(s.d_tst);
    a.totals_host = &s.h_totals[0];
    s.h_totals[2] = 0;
    cuda_check(launch_batch(p->dm, a, st), "launch(batch)");
    uint32_t* d_unserved = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(s.d_ws) + wl.ticket + 128);
    cuda_check(cudaMemsetAsync(d_unserved, 0, 4, st), "memset");
    TagArgs t;
    t.text = a.text;
    t.offsets = a.offsets;
    t.n_sent = n_sent;
    t.status = a.status;
    t.boundaries = a.boundaries;
    t.bound_offsets = a.bound_offsets;
    t.char_offsets = a.char_offsets;
    t.char_states = p->dt.char_rels ? a.char_states : nullptr;
    t.type_states = p->dt.type_rels ? a.type_states : nullptr;
    t.tag_token = static_cast<int32_t*>(s.d_tok);
    t.tag_cand = static_cast<int32_t*>(s.d_cand);
    t.n_unserved = d_unserved;
    cuda_check(launch_tags(p->dt, t, st), "launch(tags)");
    cuda_check(cudaStreamSynchronize(st), "sync(tags)");  // the totals are in pinned host memory now
    const uint64_t nb = s.h_totals[0], nc = s.h_totals[1];
    if (n_boundaries_out) *n_boundaries_out = nb;
    if (n_chars_out) *n_chars_out = nc;
    if (nb > out_capacity || nc > chars_capacity)
        throw Error(kInvalidArgument, "InvalidArgumentError: out_capacity/chars_capacity: too small for the batch");
    uint32_t unserved = 0;
    cuda_check(cudaMemcpyAsync(&unserved, d_unserved, 4, cudaMemcpyDeviceToHost, st), "D2H");
    if (nb) {
        if (scores_out) cuda_check(cudaMemcpyAsync(scores_out, s.d_scores, 4 * nb, cudaMemcpyDeviceToHost, st), "D2H(scores)");
        cuda_check(cudaMemcpyAsync(boundaries_out, s.d_bounds, nb, cudaMemcpyDeviceToHost, st), "D2H(boundaries)");
    }
    cuda_check(cudaMemcpyAsync(bound_offsets_out, s.d_boff, 8 * (n_sent + 1), cudaMemcpyDeviceToHost, st), "D2H(offsets)");
    cuda_check(cudaMemcpyAsync(char_offsets_out, s.d_coff, 8 * (n_sent + 1), cudaMemcpyDeviceToHost, st), "D2H(offsets)");
    if (status_out) cuda_check(cudaMemcpyAsync(status_out, s.d_status, 4 * n_sent, cudaMemcpyDeviceToHost, st), "D2H(status)");
    if (nc) {
        cuda_check(cudaMemcpyAsync(tag_token_out, s.d_tok, 4 * nc, cudaMemcpyDeviceToHost, st), "D2H(tags)");
        cuda_check(cudaMemcpyAsync(tag_cand_out, s.d_cand, 4 * nc * nt, cudaMemcpyDeviceToHost, st), "D2H(tags)");
    }
    cuda_check(cudaStreamSynchronize(st), "sync(copy-out)");
    if (n_unserved_out) *n_unserved_out = unserved;
    return kOk;
    VPT_API_END
}

int vpt_predict_batch_compact(const vpt_predictor* p, const uint8_t* utf8, const uint64_t* byte_offsets, size_t n_sent,
                              uint32_t* boundary_bits_out, size_t bits_capacity_words, uint32_t* n_chars_out,
                              uint8_t* status_out, uint32_t* n_tokens_out, int32_t* token_ids_out, uint8_t* token_cands_out,
                              size_t token_capacity, uint64_t* n_boundaries_out, uint64_t* n_tokens_total_out,
                              uint64_t* n_unserved_out) {
    VPT_API_BEGIN
    require_device(p);
    if (n_boundaries_out) *n_boundaries_out = 0;
    if (n_tokens_total_out) *n_tokens_total_out = 0;
    if (n_unserved_out) *n_unserved_out = 0;
    const bool want_tags = token_ids_out != nullptr || token_cands_out != nullptr;
    const bool want_tokens = want_tags || n_tokens_out != nullptr;
    if (want_tags) {
        if (!p->predict_tags || p->from_blob)
            throw Error(kInvalidArgument, "InvalidArgumentError: this predictor is created with predict_tags = false");
        if (!p->dt.tok_tab)
            throw Error(kUnsupported, "this tag model exceeds the limits of the device path (tags.hpp); use vpt_fill_tags");
        if (!token_ids_out || (p->n_tags && !token_cands_out) || !n_tokens_out)
            throw Error(kInvalidArgument, "InvalidArgumentError: token_ids_out/token_cands_out/n_tokens_out: must not be NULL");
    }
    if (!byte_offsets || !n_chars_out || !status_out)
        throw Error(kInvalidArgument, "InvalidArgumentError: byte_offsets/n_chars_out/status_out: must not be NULL");
    if (n_sent == 0) return kOk;
    if (byte_offsets[n_sent] < byte_offsets[0])
        throw Error(kInvalidArgument, "InvalidArgumentError: byte_offsets: must be non-decreasing");
    if (byte_offsets[n_sent] > byte_offsets[0] && !utf8)
        throw Error(kInvalidArgument, "InvalidArgumentError: utf8: must not be NULL");
    cuda_check(cudaSetDevice(p->device), "cudaSetDevice");

    // Chunks flow through three host-side stages, each one chunk behind the previous one, so that the device always has
    // the next chunk's kernels queued while the host waits for a chunk's totals:
    //   A  copy-in + count pass                         -> boundaries / characters of the chunk (pinned host words)
    //   B  scoring (+ tag prediction) + compaction      -> the chunk's first boundary bit is known from A's totals
    //   C  copy-out (bit words, per-sentence words, token records at the running token total)
    const size_t kChunkSentences = chunk_sentences();
    const std::vector<size_t> sizes = ramp_schedule(n_sent, kChunkSentences, kChunkSentences / 8, kChunkSentences / 4);
    const size_t nchunks = sizes.size();
    constexpr int kDepth = 4;
    std::unique_ptr<ScratchLease> lease[kDepth];
    for (int i = 0; i < kDepth && size_t(i) < nchunks; ++i) lease[i].reset(new ScratchLease(*p));
    struct CChunk {
        ChunkState cs;
        uint64_t nb = 0, nc = 0, nb_base = 0;
        cudaEvent_t kernels = nullptr;
        bool issued = false, copied = false;
    };
    std::vector<CChunk> chunks(nchunks);
    struct EventGuard {
        std::vector<CChunk>& c;
        ~EventGuard() { for (auto& x : c) { if (x.cs.counted) cudaEventDestroy(x.cs.counted); if (x.kernels) cudaEventDestroy(x.kernels); x.cs.tr.destroy(); } }
    } guard{chunks};
    for (size_t c = 0, lo = 0; c < nchunks; lo += sizes[c], ++c) {
        ChunkState& ch = chunks[c].cs;
        ch.s_lo = lo;
        ch.n = sizes[c];
        ch.byte_lo = byte_offsets[ch.s_lo];
        if (byte_offsets[ch.s_lo + ch.n] < ch.byte_lo)
            throw Error(kInvalidArgument, "InvalidArgumentError: byte_offsets: must be non-decreasing");
        ch.nbytes = byte_offsets[ch.s_lo + ch.n] - ch.byte_lo;
    }
    const size_t nt = want_tags ? p->n_tags : 0;
    uint64_t nb_total = 0, tok_total = 0, unserved_total = 0;
    bool overflow = false;
    // pinned words that receive every chunk's first bit word (merged into the output at the end)
    Scratch& s0 = *lease[0]->s;
    if (s0.side_cap < nchunks) {
        if (s0.h_side) { cudaFreeHost(s0.h_side); s0.h_side = nullptr; s0.side_cap = 0; }
        const size_t want = std::max<size_t>(256, 2 * nchunks);
        cuda_check(cudaMallocHost(reinterpret_cast<void**>(&s0.h_side), 4 * want), "cudaMallocHost");
        s0.side_cap = want;
    }
    uint32_t* const h_side = s0.h_side;
    std::vector<uint32_t> h_unserved(nchunks, 0);

    auto stage_b = [&](size_t c) {
        CChunk& cc = chunks[c];
        ChunkState& ch = cc.cs;
        Scratch& s = *lease[c % kDepth]->s;
        cudaStream_t st = s.stream;
        cuda_check(cudaEventSynchronize(ch.counted), "sync(count)");
        cc.nb = s.h_totals[0];
        cc.nc = s.h_totals[1];
        cc.nb_base = nb_total;
        nb_total += cc.nb;
        if (!cc.kernels) cuda_check(cudaEventCreateWithFlags(&cc.kernels, cudaEventDisableTiming), "cudaEventCreate");
        if ((nb_total + 31) / 32 > bits_capacity_words || (nb_total && !boundary_bits_out)) { overflow = true; return; }
        BatchArgs& a = ch.a;
        Scratch::ensure(s.d_bounds, s.bounds_cap, cc.nb + 4);
        a.scores = nullptr;
        if (!scores_optional(p->dm)) {
            Scratch::ensure(s.d_scores, s.scores_cap, 4 * cc.nb + 4);
            a.scores = static_cast<int32_t*>(s.d_scores);
        }
        a.boundaries = static_cast<uint8_t*>(s.d_bounds);
        if (want_tags) {
            Scratch::ensure(s.d_cst, s.cst_cap, 4 * cc.nc + 4);
            Scratch::ensure(s.d_tst, s.tst_cap, 4 * cc.nc + 4);
            a.char_states = static_cast<uint32_t*>(s.d_cst);
            a.type_states = static_cast<uint32_t*>(s.d_tst);
        }
        // (chunk-local offsets: bound_base / char_base stay 0)
        if (pipeline_trace()) ch.tr.mark(1, st);
        cuda_check(launch_score(p->dm, a, st), "launch(score)");
        CompactArgs k;
        k.n_sent = ch.n;
        k.status = a.status;
        k.n_chars = a.n_chars;
        k.boundaries = a.boundaries;
        k.bound_offsets = a.bound_offsets;
        k.n_bound = cc.nb;
        k.bit_base = uint32_t(cc.nb_base & 31);
        const size_t nwords = size_t((k.bit_base + cc.nb + 31) / 32);
        Scratch::ensure(s.d_bits, s.bits_cap, 4 * nwords + 16);
        Scratch::ensure(s.d_st8, s.st8_cap, ch.n + 16);
        k.bits = static_cast<uint32_t*>(s.d_bits);
        k.status8 = static_cast<uint8_t*>(s.d_st8);
        if (want_tokens) {
            Scratch::ensure(s.d_ntok, s.ntok_cap, 4 * ch.n + 16);
            Scratch::ensure(s.d_tokbase, s.tokbase_cap, 8 * (ch.n + 1) + 16);
            k.n_tokens = static_cast<uint32_t*>(s.d_ntok);
            k.tok_base = static_cast<uint64_t*>(s.d_tokbase);
            Scratch::ensure(s.d_toklocal, s.toklocal_cap, 4 * ch.n + 16);
            Scratch::ensure(s.d_tokblk, s.tokblk_cap, 8 * (ch.n / 256 + 4));
            k.tok_local = static_cast<uint32_t*>(s.d_toklocal);
            k.tok_blk = static_cast<uint64_t*>(s.d_tokblk);
            k.tok_total_host = &s.h_totals[4];
        }
        s.h_totals[4] = 0;
        cuda_check(launch_compact(k, st), "launch(compact)");
        if (want_tags) {
            // a token has at least one character
            Scratch::ensure(s.d_tok, s.tok_cap, 4 * cc.nc + 16);
            Scratch::ensure(s.d_cand, s.cand_cap, cc.nc * std::max<size_t>(nt, 1) + 16);
            uint32_t* d_unserved = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(s.d_ws) + workspace_layout(ch.n).ticket + 128);
            cuda_check(cudaMemsetAsync(d_unserved, 0, 4, st), "memset");
            TagArgs t;
            t.text = a.text;
            t.offsets = a.offsets;
            t.n_sent = ch.n;
            t.status = a.status;
            t.boundaries = a.boundaries;
            t.bound_offsets = a.bound_offsets;
            t.char_offsets = a.char_offsets;
            t.char_states = p->dt.char_rels ? a.char_states : nullptr;
            t.type_states = p->dt.type_rels ? a.type_states : nullptr;
            t.n_unserved = d_unserved;
            t.tok_base = k.tok_base;
            t.tok_ids = static_cast<int32_t*>(s.d_tok);
            t.tok_cands = static_cast<uint8_t*>(s.d_cand);
            Scratch::ensure(s.d_tokdesc, s.tokdesc_cap, 16 * cc.nc + 16);
            t.tok_desc = static_cast<uint4*>(s.d_tokdesc);
            t.max_tokens = cc.nc;
            Scratch::ensure(s.d_tokwork, s.tokwork_cap, 4 * cc.nc + 32);
            t.tok_work = static_cast<uint32_t*>(s.d_tokwork);
            t.text_base = 0;
            cuda_check(launch_tags(p->dt, t, st), "launch(tags)");
            cuda_check(cudaMemcpyAsync(&h_unserved[c], d_unserved, 4, cudaMemcpyDeviceToHost, st), "D2H(unserved)");
        }
        if (pipeline_trace()) ch.tr.mark(2, st);
        cuda_check(cudaEventRecord(cc.kernels, st), "cudaEventRecord");
        cc.issued = true;
    };
    auto stage_c = [&](size_t c) {
        CChunk& cc = chunks[c];
        ChunkState& ch = cc.cs;
        Scratch& s = *lease[c % kDepth]->s;
        if (!cc.issued) return;
        cuda_check(cudaEventSynchronize(cc.kernels), "sync(kernels)");
        const uint64_t ntok = want_tokens ? s.h_totals[4] : 0;
        if (want_tags && tok_total + ntok > token_capacity) { overflow = true; cc.issued = false; }
        cudaStream_t so = s.stream_out;
        cuda_check(cudaStreamWaitEvent(so, cc.kernels, 0), "cudaStreamWaitEvent");
        if (!overflow) {
            const uint32_t bit_base = uint32_t(cc.nb_base & 31);
            const size_t nwords = size_t((bit_base + cc.nb + 31) / 32);
            const uint64_t w0 = cc.nb_base >> 5;
            if (nwords) {
                // the chunk's first word may share its low bits with the previous chunk: it comes back through a pinned
                // word and is merged on the host at the end; the other words go straight to their place
                if (bit_base == 0) boundary_bits_out[w0] = 0;
                cuda_check(cudaMemcpyAsync(&h_side[c], s.d_bits, 4, cudaMemcpyDeviceToHost, so), "D2H(bits)");
                if (nwords > 1)
                    cuda_check(cudaMemcpyAsync(boundary_bits_out + w0 + 1, static_cast<uint32_t*>(s.d_bits) + 1, 4 * (nwords - 1),
                                               cudaMemcpyDeviceToHost, so), "D2H(bits)");
            }
            cuda_check(cudaMemcpyAsync(n_chars_out + ch.s_lo, ch.a.n_chars, 4 * ch.n, cudaMemcpyDeviceToHost, so), "D2H(n_chars)");
            cuda_check(cudaMemcpyAsync(status_out + ch.s_lo, s.d_st8, ch.n, cudaMemcpyDeviceToHost, so), "D2H(status)");
            if (n_tokens_out) cuda_check(cudaMemcpyAsync(n_tokens_out + ch.s_lo, s.d_ntok, 4 * ch.n, cudaMemcpyDeviceToHost, so), "D2H(n_tokens)");
            if (want_tags && ntok) {
                cuda_check(cudaMemcpyAsync(token_ids_out + tok_total, s.d_tok, 4 * ntok, cudaMemcpyDeviceToHost, so), "D2H(tokens)");
                if (nt) cuda_check(cudaMemcpyAsync(token_cands_out + tok_total * nt, s.d_cand, ntok * nt, cudaMemcpyDeviceToHost, so), "D2H(tokens)");
            }
        }
        cuda_check(cudaEventRecord(s.ev_out, so), "cudaEventRecord");
        if (pipeline_trace()) ch.tr.mark(3, so);
        cc.copied = !overflow && cc.nb != 0;
        tok_total += ntok;
    };
    // A runs kDepth - 2 chunks ahead of B, B one chunk ahead of C (a scratch is free again when its chunk's C is done)
    constexpr size_t kAheadA = kDepth - 2;
    for (size_t c = 0; c < std::min<size_t>(kAheadA, nchunks); ++c) chunk_count(*lease[c % kDepth]->s, chunks[c].cs, utf8, byte_offsets);
    for (size_t step = 0; step < nchunks + 1; ++step) {
        if (step + kAheadA < nchunks) chunk_count(*lease[(step + kAheadA) % kDepth]->s, chunks[step + kAheadA].cs, utf8, byte_offsets);
        if (step < nchunks) stage_b(step);
        if (step >= 1) stage_c(step - 1);
    }
    for (int i = 0; i < kDepth; ++i)
        if (lease[i]) {
            cuda_check(cudaStreamSynchronize(lease[i]->s->stream), "sync(score)");
            cuda_check(cudaStreamSynchronize(lease[i]->s->stream_out), "sync(copy-out)");
        }
    for (uint32_t u : h_unserved) unserved_total += u;
    for (size_t c = 0; c < nchunks; ++c)
        if (chunks[c].copied) boundary_bits_out[chunks[c].nb_base >> 5] |= h_side[c];
    if (pipeline_trace())
        for (size_t c = 0; c < nchunks; ++c) chunks[c].cs.tr.print("compact", c, chunks[c].cs.n, chunks[0].cs.tr);
    if (n_boundaries_out) *n_boundaries_out = nb_total;
    if (n_tokens_total_out) *n_tokens_total_out = tok_total;
    if (n_unserved_out) *n_unserved_out = unserved_total;
    if (overflow) throw Error(kInvalidArgument, "InvalidArgumentError: bits_capacity_words/token_capacity: too small for the batch");
    return kOk;
    VPT_API_END
}

int vpt_unpack_boundaries(const uint32_t* boundary_bits, uint64_t first_bit, uint64_t n, uint8_t* boundaries_out) {
    VPT_API_BEGIN
    if (n && (!boundary_bits || !boundaries_out)) throw Error(kInvalidArgument, "InvalidArgumentError: buffers: must not be NULL");
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t b = first_bit + i;
        boundaries_out[i] = uint8_t((boundary_bits[b >> 5] >> (b & 31)) & 1u);
    }
    return kOk;
    VPT_API_END
}

int vpt_fill_tags(const vpt_predictor* p, const uint8_t* utf8, size_t n_bytes, const uint8_t* boundaries,
                  const uint32_t* char_states, const uint32_t* type_states, int32_t* tag_token_out,
                  int32_t* tag_cand_out, int32_t* tag_scores_out, size_t score_stride) {
    VPT_API_BEGIN
    if (!p) throw Error(kInvalidArgument, "InvalidArgumentError: predictor: must not be NULL");
    if (!p->predict_tags || p->from_blob)
        throw Error(kInvalidArgument, "InvalidArgumentError: this predictor is created with predict_tags = false");
    if (!utf8 || !boundaries || !tag_token_out || !tag_cand_out)
        throw Error(kInvalidArgument, "InvalidArgumentError: utf8/boundaries/tag_token_out/tag_cand_out: must not be NULL");
    check_raw_text(utf8, n_bytes);
    const std::vector<uint32_t> pos = char_starts(utf8, n_bytes);
    const size_t n = pos.size() - 1;
    const size_t nt = p->n_tags;
    for (size_t i = 0; i < n; ++i) tag_token_out[i] = -1;
    for (size_t i = 0; i < n * nt; ++i) tag_cand_out[i] = -1;
    if (nt == 0) return kOk;  // predictor.rs:553-555
    if ((p->char_tags && !char_states) || (p->type_tags && !type_states))
        throw Error(kInvalidArgument, "InvalidArgumentError: states: required for tag prediction");
    std::vector<int32_t> scores;
    auto run = [&](size_t start, size_t last) {  // token = chars [start, last]
        std::string tok(reinterpret_cast<const char*>(utf8) + pos[start], pos[last + 1] - pos[start]);
        auto it = p->token_ids.find(tok);
        if (it == p->token_ids.end()) return;
        const uint32_t tid = it->second;
        const TagPredictorHost& tp = p->tag_preds[tid];
        scores.assign(tp.bias.size(), 0);
        add_truncated(tp.bias, scores);
        if (p->char_tags) add_tag_scores(p->char_tag_weight, p->char_suffix_link, tid, last, char_states, n, scores);
        if (p->type_tags) add_tag_scores(p->type_tag_weight, p->type_suffix_link, tid, last, type_states, n, scores);
        // TagPredictor::predict (predictor.rs:286-304): first strict maximum per slot with >= 2 candidates
        size_t off = 0;
        for (size_t k = 0; k < tp.tags.size() && k < nt; ++k) {
            const size_t ncand = tp.tags[k].size();
            if (ncand >= 2) {
                if (off + ncand > scores.size())
                    throw Error(kInvalidModel, "InvalidModelError: tag bias is shorter than the number of candidates");
                size_t best = 0;
                int32_t mx = INT32_MIN;
                for (size_t c = 0; c < ncand; ++c)
                    if (scores[off + c] > mx) { best = c; mx = scores[off + c]; }
                tag_cand_out[last * nt + k] = int32_t(best);
                off += ncand;
            } else {
                tag_cand_out[last * nt + k] = ncand == 1 ? 0 : -1;
            }
        }
        tag_token_out[last] = int32_t(tid);
        if (tag_scores_out) {
            const size_t m = std::min(score_stride, scores.size());
            memcpy(tag_scores_out + last * score_stride, scores.data(), m * 4);
        }
    };
    bool have = true;
    size_t start = 0;
    for (size_t i = 0; i + 1 < n; ++i) {
        const uint8_t b = boundaries[i];
        if (b == 2) have = false;
        else if (b == 1) {
            if (have) run(start, i);
            have = true;
            start = i + 1;
        }
    }
    if (have) run(start, n - 1);
    return kOk;
    VPT_API_END
}

int vpt_write_tokenized_text(const vpt_predictor* p, const uint8_t* utf8, size_t n_bytes, const uint8_t* boundaries,
                             const int32_t* tag_token, const int32_t* tag_cand, char* buf, size_t capacity,
                             uint64_t* len_out) {
    VPT_API_BEGIN
    if (!utf8 || !boundaries) throw Error(kInvalidArgument, "InvalidArgumentError: utf8/boundaries: must not be NULL");
    check_raw_text(utf8, n_bytes);
    const std::vector<uint32_t> pos = char_starts(utf8, n_bytes);
    const size_t n = pos.size() - 1;
    const size_t nt = (p && tag_token && tag_cand) ? p->n_tags : 0;
    std::string out;
    auto put = [&](const char* s, size_t l) {
        for (size_t i = 0; i < l; ++i) {
            if (s[i] == ' ' || s[i] == '\\' || s[i] == '/') out.push_back('\\');
            out.push_back(s[i]);
        }
    };
    auto emit = [&](size_t st, size_t en) {  // chars [st, en)
        if (!out.empty()) out.push_back(' ');
        put(reinterpret_cast<const char*>(utf8) + pos[st], pos[en] - pos[st]);
        if (nt) {
            const size_t i = en - 1;
            int last = -1;
            for (size_t k = 0; k < nt; ++k) if (tag_cand[i * nt + k] >= 0) last = int(k);
            for (int k = 0; k <= last; ++k) {
                out.push_back('/');
                const int32_t c = tag_cand[i * nt + size_t(k)];
                if (c >= 0) {
                    const char* t = vpt_tag_string(p, uint32_t(tag_token[i]), uint32_t(k), uint32_t(c));
                    if (t) put(t, strlen(t));
                }
            }
        }
    };
    // TokenIterator (sentence.rs:1273-1299): tokens adjacent to Unknown boundaries are skipped
    size_t start = 0;
    bool skip = false;
    for (size_t i = 0; i + 1 < n; ++i) {
        const uint8_t b = boundaries[i];
        if (b == 1) {
            if (!skip) emit(start, i + 1);
            skip = false;
            start = i + 1;
        } else if (b == 2) skip = true;
    }
    if (!skip) emit(start, n);
    if (len_out) *len_out = out.size();
    if (buf && capacity) {
        const size_t m = std::min(capacity - 1, out.size());
        memcpy(buf, out.data(), m);
        buf[m] = 0;
    }
    return kOk;
    VPT_API_END
}

}  // extern "C"
