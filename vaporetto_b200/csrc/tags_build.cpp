// Host-side construction of the flat tag-prediction tables (tags.hpp) from a HostPredictor.
#include <algorithm>
#include <cstring>
#include <string>

#include "common.hpp"
#include "tags.hpp"

namespace vpt {

namespace {

uint32_t pow2_at_least(size_t n) {
    uint32_t c = 16;
    while (c < n) c <<= 1;
    return c;
}

void fill_weight_table(const TagWeightMap& tw, std::vector<TagWeightSlot>& tab, uint32_t& mask, std::vector<int32_t>& pool,
                       uint32_t& rels) {
    size_t n = 0;
    rels = 0;
    for (const auto& per_rel : tw) {
        rels = std::max<uint32_t>(rels, uint32_t(per_rel.size()));
        for (const auto& m : per_rel) n += m.size();
    }
    const uint32_t cap = pow2_at_least(2 * n + 16);
    mask = cap - 1;
    tab.assign(cap, TagWeightSlot{0, 0, 0});
    for (size_t tid = 0; tid < tw.size(); ++tid)
        for (size_t rel = 0; rel < tw[tid].size(); ++rel)
            for (const auto& kv : tw[tid][rel]) {
                const uint64_t key = tag_weight_key(kv.first, uint32_t(tid), uint32_t(rel));
                uint32_t s = tag_weight_slot(key, mask);
                while (tab[s].key != 0) s = (s + 1) & mask;
                tab[s].key = key;
                tab[s].off = uint32_t(pool.size());
                tab[s].len = uint32_t(kv.second.size());
                pool.insert(pool.end(), kv.second.begin(), kv.second.end());
            }
}

// 1 for every pattern that has an own tag weight or whose suffix chain reaches one
std::vector<uint8_t> chain_flags(const TagWeightMap& tw, const std::vector<uint32_t>& link) {
    std::vector<uint8_t> own(link.size(), 0), any(link.size(), 2);  // 2 = unknown
    for (const auto& per_rel : tw)
        for (const auto& m : per_rel)
            for (const auto& kv : m)
                if (kv.first < own.size()) own[kv.first] = 1;
    std::vector<uint32_t> path;
    for (uint32_t p = 0; p < link.size(); ++p) {
        if (any[p] != 2) continue;
        path.clear();
        uint32_t q = p;
        uint8_t v = 0;
        while (true) {
            if (any[q] != 2) { v = any[q]; break; }
            path.push_back(q);
            if (own[q]) { v = 1; break; }
            if (link[q] == kNoPattern) { v = 0; break; }
            q = link[q];
        }
        for (uint32_t x : path) any[x] = v;
        // (nodes after an own entry on the path keep "unknown" and are resolved by their own iteration)
    }
    return any;
}

}  // namespace

TagTablesHost build_tag_tables(const HostPredictor& hp) {
    TagTablesHost t;
    t.n_tags = uint32_t(hp.n_tags);
    t.n_tokens = uint32_t(hp.tag_preds.size());
    if (!hp.predict_tags || hp.n_tags == 0 || hp.n_tags > size_t(kTagMaxSlots) || hp.tag_preds.size() >= (1u << 24)) return t;
    t.pool.push_back(0);  // offset 0 is never a real vector (keeps "off" non-zero for debugging)
    // token info + bias
    t.tok_info.resize(hp.tag_preds.size());
    for (size_t i = 0; i < hp.tag_preds.size(); ++i) {
        const TagPredictorHost& tp = hp.tag_preds[i];
        TagTokenInfo& ti = t.tok_info[i];
        memset(&ti, 0, sizeof ti);
        ti.usable = tp.bias.size() <= size_t(kTagMaxScores) && tp.tags.size() <= size_t(kTagMaxSlots);
        ti.bias_off = uint32_t(t.pool.size());
        ti.bias_len = uint16_t(std::min<size_t>(tp.bias.size(), 65535));
        ti.n_slots = uint8_t(std::min<size_t>(tp.tags.size(), size_t(kTagMaxSlots)));
        for (size_t k = 0; k < size_t(kTagMaxSlots); ++k) {
            const size_t nc = k < tp.tags.size() ? tp.tags[k].size() : 0;
            ti.cand[k] = uint8_t(std::min<size_t>(nc, 255));
            if (nc >= 255) ti.usable = 0;
        }
        t.pool.insert(t.pool.end(), tp.bias.begin(), tp.bias.end());
    }
    // token table (the map already holds "last insert wins" for duplicate tokens)
    const uint32_t cap = pow2_at_least(2 * hp.token_ids.size() + 16);
    t.tok_mask = cap - 1;
    t.tok_tab.assign(cap, TagTokenEntry{0, 0, 0, 0, 0});
    for (const auto& kv : hp.token_ids) {
        uint64_t h = kTagHashInit;
        for (unsigned char c : kv.first) h = tag_hash_step(h, c);
        h = tag_hash_finish(h);
        uint32_t s = uint32_t(h >> 20) & t.tok_mask;
        while (t.tok_tab[s].hash != 0) s = (s + 1) & t.tok_mask;
        t.tok_tab[s].hash = h;
        t.tok_tab[s].tid = kv.second;
        t.tok_tab[s].str_off = uint32_t(t.tok_bytes.size());
        t.tok_tab[s].len = uint32_t(kv.first.size());
        t.tok_bytes.insert(t.tok_bytes.end(), kv.first.begin(), kv.first.end());
        t.max_token_bytes = std::max<uint32_t>(t.max_token_bytes, uint32_t(kv.first.size()));
    }
    t.tok_bytes.resize(t.tok_bytes.size() + 16, 0);
    if (hp.char_tags) fill_weight_table(hp.char_tag_weight, t.cw_tab, t.cw_mask, t.pool, t.char_rels);
    if (hp.type_tags) fill_weight_table(hp.type_tag_weight, t.tw_tab, t.tw_mask, t.pool, t.type_rels);
    if (t.cw_tab.empty()) { t.cw_tab.assign(16, TagWeightSlot{0, 0, 0}); t.cw_mask = 15; }
    if (t.tw_tab.empty()) { t.tw_tab.assign(16, TagWeightSlot{0, 0, 0}); t.tw_mask = 15; }
    t.c_link = hp.char_suffix_link;
    t.t_link = hp.type_suffix_link;
    if (hp.char_tags) t.c_any = chain_flags(hp.char_tag_weight, t.c_link);
    if (hp.type_tags) t.t_any = chain_flags(hp.type_tag_weight, t.t_link);
    if (t.c_link.empty()) t.c_link.push_back(kNoPattern);
    if (t.t_link.empty()) t.t_link.push_back(kNoPattern);
    if (t.c_any.empty()) t.c_any.assign(t.c_link.size(), 0);
    if (t.t_any.empty()) t.t_any.assign(t.t_link.size(), 0);
    t.usable = true;
    return t;
}

}  // namespace vpt
