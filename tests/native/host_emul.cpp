// TEST INFRASTRUCTURE: interprets the flat model blob produced by the product's host builder
// (vaporetto_b200/csrc/predictor_build.cpp) on the CPU, following the same probe sequence as the kernels
// (kernels.cu: find_node / k_score_fast / k_score_general), so that table construction can be checked
// against the oracle without a GPU.  Not part of the product; never linked into libvaporetto_b200.so.
#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../vaporetto_b200/csrc/predictor_build.hpp"
#include "../../vaporetto_b200/csrc/common.hpp"
#include "../../vaporetto_b200/csrc/tags.hpp"
#include "../../vaporetto_b200/csrc/tags_token.hpp"

using namespace vpt;

namespace {

struct Tab {
    const BlobTable* bt;
    const uint8_t* base;
    TableGeom geom() const { TableGeom g; g.nslots = bt->nslots; g.nbuckets = bt->nbuckets; g.salt = bt->salt; g.seed_bits = bt->seed_bits; return g; }
    const uint8_t* seeds() const { return base + bt->seeds_off; }
    const uint32_t* rec(uint32_t slot) const { return reinterpret_cast<const uint32_t*>(base + bt->rec_off + size_t(slot) * 32); }
    const uint32_t* slot_node() const { return reinterpret_cast<const uint32_t*>(base + bt->node_off); }
    const int32_t* pool() const { return reinterpret_cast<const int32_t*>(base + bt->pool_off); }
    const uint32_t* slot_pid() const { return reinterpret_cast<const uint32_t*>(base + bt->pid_off); }
    const uint64_t* slot_ovf() const { return reinterpret_cast<const uint64_t*>(base + bt->ovf_off); }
    bool probe(uint64_t key, const uint32_t*& r, uint32_t& slot, bool deep = false) const {
        slot = table_slot(geom(), seeds(), key);
        r = rec(slot);
        uint64_t k = (uint64_t(r[1]) << 32) | r[0];
        const uint64_t flags = deep ? (kExtFlag | kOvfFlag) : ((key >> 42) ? kExtFlag : (kExtFlag | kChildMaskField));
        return (k & ~flags) == key;
    }
};

uint8_t ctype(uint32_t c) {
    if ((c >= 0x30 && c <= 0x39) || (c >= 0xFF10 && c <= 0xFF19)) return 1;
    if ((c >= 0x41 && c <= 0x5A) || (c >= 0x61 && c <= 0x7A) || (c >= 0xFF21 && c <= 0xFF3A) || (c >= 0xFF41 && c <= 0xFF5A)) return 2;
    if (c >= 0x3040 && c <= 0x3096) return 3;
    if ((c >= 0x30A0 && c <= 0x30FA) || (c >= 0x30FC && c <= 0x30FF) || (c >= 0xFF66 && c <= 0xFF9F)) return 4;
    if ((c >= 0x3400 && c <= 0x4DBF) || (c >= 0x4E00 && c <= 0x9FFF) || (c >= 0xF900 && c <= 0xFAFF) || (c >= 0x20000 && c <= 0x2A6DF) ||
        (c >= 0x2A700 && c <= 0x2B73F) || (c >= 0x2B740 && c <= 0x2B81F) || (c >= 0x2B820 && c <= 0x2CEAF) || (c >= 0x2F800 && c <= 0x2FA1F)) return 5;
    return 6;
}

bool find_node(const Tab& t, const std::vector<uint32_t>& sym, size_t g, const uint32_t*& rec, uint32_t& slot, bool& deep_hit) {
    deep_hit = false;
    uint32_t c3 = sym[g], c2 = g >= 1 ? sym[g - 1] : 0, c1 = g >= 2 ? sym[g - 2] : 0;
    bool found = t.probe(shallow_key(c1, c2, c3), rec, slot);
    bool depth3 = found && c1 != 0;
    if (!found && c1 != 0) found = t.probe(shallow_key(0, c2, c3), rec, slot);
    if (!found && c2 != 0) found = t.probe(shallow_key(0, 0, c3), rec, slot);
    if (depth3 && (rec[1] >> 31) && g >= 3) {
        size_t i = g - 2;
        uint32_t node = t.slot_node()[slot];
        while (i > 0) {
            --i;
            const uint32_t* nrec; uint32_t nslot;
            if (!t.probe(deep_key(node, sym[i]), nrec, nslot, true)) break;
            rec = nrec; slot = nslot; deep_hit = true;
            if (!(rec[1] >> 31)) break;
            node = t.slot_node()[slot];
        }
    }
    return found;
}

// The probe order of k_fused (fused_kernel.cuh, stages 1-3 of the stream stage), word for word: the node of the last TWO
// symbols first (one symbol at a sentence start), compared on its 32-bit halves with the child mask and the deep-key marker
// bits masked as the kernel masks them; then the 3-symbol node if the 2-symbol record's child mask has the bit of the
// third symbol, or the 1-symbol node if the 2-symbol node does not exist; a 3-symbol node with extensions walks on.
// Must select the same record as find_node (which asks for the longest node first): that is what the builder's child
// masks promise.
bool find_node_fused(const Tab& t, const std::vector<uint32_t>& sym, size_t g, const uint32_t*& rec, uint32_t& slot, bool& deep_hit) {
    deep_hit = false;
    const uint32_t c = sym[g], c2 = g >= 1 ? sym[g - 1] : 0, c1 = g >= 2 ? sym[g - 2] : 0;
    const uint32_t slotA = table_slot(t.geom(), t.seeds(), shallow_key(0, c2, c));
    const uint32_t* rA = t.rec(slotA);
    const uint32_t klo = c | (c2 << 21), khi = c2 >> 11;
    const bool f1 = rA[0] == klo && (rA[1] & 0x600003FFu) == khi;
    const bool w3 = f1 && c1 != 0 && ((rA[1] >> (10u + child_bit(c1))) & 1u) != 0, w1 = !f1 && c2 != 0;
    bool f2 = false;
    uint32_t slotB = 0;
    const uint32_t* rB = nullptr;
    if (w3 || w1) {
        slotB = table_slot(t.geom(), t.seeds(), w3 ? shallow_key(c1, c2, c) : shallow_key(0, 0, c));
        rB = t.rec(slotB);
        const uint32_t kloB = w3 ? klo : c, khiB = w3 ? (khi | (c1 << 10)) : 0u;
        f2 = rB[0] == kloB && (rB[1] & 0x7FFFFFFFu) == khiB;
    }
    if (f2) {
        rec = rB;
        slot = slotB;
        if (w3 && (rec[1] >> 31) && g >= 3) {
            size_t i = g - 2;
            uint32_t node = t.slot_node()[slot];
            while (i > 0) {
                --i;
                const uint32_t* nrec; uint32_t nslot;
                if (!t.probe(deep_key(node, sym[i]), nrec, nslot, true)) break;
                rec = nrec; slot = nslot; deep_hit = true;
                if (!(rec[1] >> 31)) break;
                node = t.slot_node()[slot];
            }
        }
        return true;
    }
    if (f1) { rec = rA; slot = slotA; return true; }
    return false;
}

}  // namespace

extern "C" {

// returns n_chars (>0) or -(status).  scores: n-1; states: n each (nullable).  info[0]=fast path flag.
long emul_predict(const uint8_t* model, size_t model_len, int predict_tags, const uint8_t* utf8, size_t nbytes,
                  int32_t* scores, uint32_t* cstates, uint32_t* tstates, int32_t* info) {
    try {
        size_t consumed = 0;
        Model m = Model::read(model, model_len, &consumed);
        HostPredictor hp = build_host_predictor(m, predict_tags != 0);
        const uint8_t* base = hp.blob.data();
        BlobHeader h;
        memcpy(&h, base, sizeof h);
        Tab ct{&h.ct, base}, tt{&h.tt, base};
        std::vector<uint32_t> cps = utf8_to_codepoints(std::string(reinterpret_cast<const char*>(utf8), nbytes));
        const size_t n = cps.size();
        std::vector<uint32_t> tys(n);
        for (size_t i = 0; i < n; ++i) tys[i] = ctype(cps[i]);
        const long nout = long(n) - 1;
        const bool fast = (!h.ct.present || h.ct.fast) && !h.tt.present;
        if (info) { info[0] = fast; info[1] = h.char_variant; info[2] = h.type_variant; }
        for (long i = 0; i < nout; ++i) {
            int32_t v = h.bias;
            if (h.type_cache_window) {
                const int w = h.type_cache_window;
                uint32_t idx = 0;
                for (int k = 0; k < 2 * w; ++k) {
                    long j = i - w + 1 + k;
                    idx = (idx << 3) | ((j >= 0 && j < long(n)) ? tys[size_t(j)] : 0u);
                }
                const int32_t full = reinterpret_cast<const int32_t*>(base + h.type_cache_off)[idx];
                if (h.type_a_off) {
                    const int32_t sp = wrapping_add(reinterpret_cast<const int32_t*>(base + h.type_a_off)[idx >> 6],
                                                    reinterpret_cast<const int32_t*>(base + h.type_b_off)[idx & 4095]);
                    if (sp != full) throw Error(kInternal, "split type tables disagree with the full table");
                }
                v = wrapping_add(v, full);
            }
            scores[i] = v;
        }
        for (size_t g = 0; g < n; ++g) {
            if (cstates) cstates[g] = kNoPattern;
            if (tstates) {
                tstates[g] = kNoPattern;
                if (h.emit_states && h.type_state3_off) {
                    const uint32_t t2 = g >= 1 ? tys[g - 1] : 0, t1 = (g >= 2 && t2) ? tys[g - 2] : 0;
                    tstates[g] = reinterpret_cast<const uint32_t*>(base + h.type_state3_off)[(t1 << 6) | (t2 << 3) | tys[g]];
                }
            }
        }
        for (int which = 0; which < 2; ++which) {
            const Tab& t = which ? tt : ct;
            if (!t.bt->present) continue;
            const std::vector<uint32_t>& sym = which ? tys : cps;
            for (size_t g = 0; g < n; ++g) {
                const uint32_t* rec; uint32_t slot; bool deep_hit;
                const bool found = find_node(t, sym, g, rec, slot, deep_hit);
                if (which == 0 && t.bt->fast) {
                    // (the char table in the inline format is what k_fused reads: its probe order must agree)
                    const uint32_t* rec2 = nullptr; uint32_t slot2 = 0; bool deep2 = false;
                    const bool found2 = find_node_fused(t, sym, g, rec2, slot2, deep2);
                    if (found2 != found || (found && (slot2 != slot || deep2 != deep_hit)))
                        throw Error(kInternal, "k_fused probe order selects a different node record");
                }
                if (!found) continue;
                if (t.bt->fast) {
                    if (h.emit_states && which == 0 && cstates) cstates[g] = t.slot_pid()[slot];
                    for (int j = 0; j < kInlineWidth; ++j) {
                        long i = long(g) + t.bt->r0 + j;
                        if (i >= 0 && i < nout) scores[i] = wrapping_add(scores[i], int32_t(rec[2 + j]));
                    }
                    if (deep_hit && ((uint64_t(rec[1]) << 32) & kOvfFlag)) {
                        if (!t.bt->has_overflow) throw Error(kInternal, "overflow flag on a table without overflow rows");
                        const uint64_t dsc = t.slot_ovf()[slot];
                        const uint32_t ptr = uint32_t(dsc);
                        const int off = int(int16_t(uint16_t(dsc >> 32)));
                        const int len = int(uint16_t(dsc >> 48));
                        for (int k = 0; k < len; ++k) {
                            long i = long(g) + off + k;
                            if (i >= 0 && i < nout) scores[i] = wrapping_add(scores[i], t.pool()[ptr + k]);
                        }
                    }
                } else {
                    if (h.emit_states) { if (which == 0 && cstates) cstates[g] = rec[2]; if (which == 1 && tstates) tstates[g] = rec[2]; }
                    if (rec[3] != kNoPattern) {
                        const int32_t off = int32_t(rec[4]);
                        for (uint32_t k = 0; k < rec[5]; ++k) {
                            long i = long(g) + off + long(k);
                            if (i >= 0 && i < nout) scores[i] = wrapping_add(scores[i], t.pool()[rec[3] + k]);
                        }
                    }
                }
            }
        }
        return long(n);
    } catch (const Error& e) {
        set_last_error(e.what());
        return -long(e.code);
    }
}

// Tag prediction from the flat tag tables (tags_build.cpp: token table, token info, key lists, chain tables) with the
// kernels' OWN per-token code (csrc/tags_token.hpp: token_lookup, add_scorer, tag_score_token compile for the host too):
// token by hash + byte compare (over the bytes of its KyteaFullwidthFilter image when norm != 0), bias, for every rel
// position the own vectors of the patterns on the suffix chain of the pattern found there, first strict maximum per slot.
// boundaries: n_chars - 1 bytes (1 = token boundary); cstates / tstates: pattern ids per character (of the text the
// scorer saw: the filtered text when norm != 0).  tag_token: n_chars (token id at a token's last character, else -1);
// tag_cand: n_chars x n_tags.  Returns n_tags or -(status); *unserved counts tokens beyond the device limits.
long emul_predict_tags(const uint8_t* model, size_t model_len, const uint8_t* utf8, size_t nbytes, const uint8_t* boundaries,
                       const uint32_t* cstates, const uint32_t* tstates, int norm, int32_t* tag_token, int32_t* tag_cand,
                       int32_t* unserved) {
    try {
        size_t consumed = 0;
        Model m = Model::read(model, model_len, &consumed);
        HostPredictor hp = build_host_predictor(m, true);
        const TagTablesHost t = build_tag_tables(hp);
        if (!t.usable) throw Error(kInternal, "tag tables not usable");
        DevTags d;  // the same fields capi.cpp fills, pointing at the host tables
        d.tok_tab = t.tok_tab.data();
        d.tok_bytes = t.tok_bytes.data();
        d.tok_info = t.tok_info.data();
        d.pool = t.pool.data();
        d.keys = t.keys.data();
        d.c_chain = t.c_chain.data();
        d.t_chain = t.t_chain.data();
        d.c_link = t.c_link.data();
        d.t_link = t.t_link.data();
        d.tok_mask = t.tok_mask;
        d.n_tags = t.n_tags;
        d.char_rels = hp.char_tags ? t.char_rels : 0;
        d.type_rels = hp.type_tags ? t.type_rels : 0;
        d.max_token_bytes = t.max_token_bytes;
        d.n_char_patterns = uint32_t(hp.char_suffix_link.size());
        d.n_type_patterns = uint32_t(hp.type_suffix_link.size());
        std::vector<uint32_t> start;
        for (size_t i = 0; i < nbytes; ++i) if ((utf8[i] & 0xC0) != 0x80) start.push_back(uint32_t(i));
        const size_t n = start.size();
        start.push_back(uint32_t(nbytes));
        const uint32_t nt = t.n_tags;
        uint32_t uns = 0;
        size_t tok_start = 0;
        for (size_t i = 0; i < n; ++i) {
            tag_token[i] = -1;
            for (uint32_t k = 0; k < nt; ++k) tag_cand[i * nt + k] = -1;
            const bool ends = i + 1 == n || boundaries[i] == 1;
            if (!ends) continue;
            int32_t cand[kTagMaxSlots];
            for (int k = 0; k < kTagMaxSlots; ++k) cand[k] = -1;
            const int32_t tok = tag_token_at(d, utf8 + start[tok_start], start[i + 1] - start[tok_start],
                                             d.char_rels ? cstates : nullptr, d.type_rels ? tstates : nullptr, uint32_t(i),
                                             uint32_t(n), cand, &uns, norm);
            tok_start = i + 1;
            tag_token[i] = tok;
            for (uint32_t k = 0; k < nt; ++k) tag_cand[i * nt + k] = tok >= 0 ? cand[k] : -1;
        }
        if (unserved) *unserved = int32_t(uns);
        return long(nt);
    } catch (const Error& e) {
        set_last_error(e.what());
        return -long(e.code);
    }
}

const char* emul_last_error() { return last_error(); }

}
