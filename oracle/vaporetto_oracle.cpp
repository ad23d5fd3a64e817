// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// CPU restatement of the reference's `Predictor::predict` path
// (daac-tools/vaporetto @ 1d215b8, crate vaporetto 0.6.5).  Only `tests/`,
// `__graft_entry__.smoke()` and bench.py's cpu_baseline / `--impl reference`
// legs may load this library; the product (`vaporetto_b200/`) never does.
//
// Parity status: PINNED.  The reference itself (Rust) cannot be compiled in
// this image (no cargo/rustc; daachorse/bincode are un-vendored crates), so
// this restatement is pinned against every known-answer test the reference
// holds for the path (tests/test_oracle_golden.py) and
// against the reference's binary fixtures committed under tests/golden/.
//
// Third-party arithmetic restated from its published behaviour:
//  * daachorse 1.x (vaporetto/Cargo.toml:17) `find_overlapping_no_suffix_iter`:
//    at every text position report only the longest pattern ending there;
//    `find_overlapping_iter`: all patterns ending there.  Built here as a
//    classic goto/failure Aho-Corasick automaton and walked as a double array
//    with a code-point mapper (daachorse's charwise layout; the layout does not
//    affect results, it only keeps the timed CPU arm honest: struct AC).
//  * bincode 2.0.1 `config::standard()` (vaporetto/Cargo.toml:16): varint
//    integers, zigzag for signed, length-prefixed Vec/String.
//
// Each function cites the reference file:line it follows
// (paths relative to /root/reference/vaporetto/src).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

#include "grapheme_tables.hpp"

#if defined(__linux__)
#include <pthread.h>
#include <sched.h>
#endif

namespace ora {

using std::string;
using std::vector;

static thread_local string g_err;

struct Error : std::runtime_error {
    int code;
    Error(int c, const string& m) : std::runtime_error(m), code(c) {}
};
enum { OK = 0, INVALID_MODEL = 1, INVALID_ARGUMENT = 2, DECODE_ERROR = 3, IO_ERROR = 5 };

// ---------------------------------------------------------------------------
// bincode standard-config decoder (bincode 2.0.1 varint spec).
// ---------------------------------------------------------------------------
struct Reader {
    const uint8_t* p;
    size_t n, pos = 0;
    uint8_t u8() {
        if (pos >= n) throw Error(DECODE_ERROR, "unexpected end of model data");
        return p[pos++];
    }
    uint64_t fixed(int bytes) {
        uint64_t v = 0;
        for (int i = 0; i < bytes; ++i) v |= uint64_t(u8()) << (8 * i);
        return v;
    }
    uint64_t varint() {
        uint8_t b = u8();
        if (b < 251) return b;
        if (b == 251) return fixed(2);
        if (b == 252) return fixed(4);
        if (b == 253) return fixed(8);
        throw Error(DECODE_ERROR, "unsupported varint width");
    }
    int64_t zigzag() {
        uint64_t u = varint();
        return int64_t(u >> 1) ^ -int64_t(u & 1);
    }
    size_t len() {
        uint64_t l = varint();
        if (l > n - pos + 8 && l > (1ull << 40)) throw Error(DECODE_ERROR, "length too large");
        return size_t(l);
    }
    string bytes() {
        size_t l = len();
        if (l > n - pos) throw Error(DECODE_ERROR, "unexpected end of model data");
        string s(reinterpret_cast<const char*>(p + pos), l);
        pos += l;
        return s;
    }
    vector<int32_t> vec_i32() {
        size_t l = len();
        if (l > n - pos) throw Error(DECODE_ERROR, "unexpected end of model data");
        vector<int32_t> v(l);
        for (auto& x : v) x = int32_t(zigzag());
        return v;
    }
};

static bool valid_utf8(const string& s) {
    size_t i = 0, n = s.size();
    auto b = [&](size_t k) { return uint8_t(s[k]); };
    while (i < n) {
        uint8_t c = b(i);
        if (c < 0x80) { i += 1; continue; }
        int need; uint32_t cp;
        if (c >= 0xC2 && c <= 0xDF) { need = 1; cp = c & 0x1F; }
        else if (c >= 0xE0 && c <= 0xEF) { need = 2; cp = c & 0x0F; }
        else if (c >= 0xF0 && c <= 0xF4) { need = 3; cp = c & 0x07; }
        else return false;
        if (i + size_t(need) >= n) return false;
        for (int k = 1; k <= need; ++k) {
            if ((b(i + k) & 0xC0) != 0x80) return false;
            cp = (cp << 6) | (b(i + k) & 0x3F);
        }
        if (need == 2 && (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF))) return false;
        if (need == 3 && (cp < 0x10000 || cp > 0x10FFFF)) return false;
        i += need + 1;
    }
    return true;
}

// ---------------------------------------------------------------------------
// Model data (model.rs:41-47,61-70; ngram_model.rs:6-27; dict_model.rs:18-22)
// ---------------------------------------------------------------------------
struct NgramData { string ngram; vector<int32_t> weights; };
struct TagWeight { uint8_t rel_position; vector<int32_t> weights; };
struct TagNgramData { string ngram; vector<TagWeight> weights; };
struct WordWeightRecord { string word; vector<int32_t> weights; string comment; };
struct TagModel {
    string token;
    vector<vector<string>> tags;
    vector<TagNgramData> char_ngram_model, type_ngram_model;
    vector<int32_t> bias;
};
struct Model {
    vector<NgramData> char_ngram_model, type_ngram_model;
    vector<WordWeightRecord> dict_model;
    int32_t bias = 0;
    uint8_t char_window_size = 0, type_window_size = 0;
    vector<TagModel> tag_models;
};

static const char MODEL_MAGIC[] = "VaporettoTokenizer 0.5.0\n";  // model.rs:15

static vector<NgramData> read_ngrams(Reader& r, bool is_str) {
    size_t l = r.len();
    vector<NgramData> v;
    for (size_t i = 0; i < l; ++i) {
        NgramData d;
        d.ngram = r.bytes();
        if (is_str && !valid_utf8(d.ngram)) throw Error(DECODE_ERROR, "invalid UTF-8 in model string");
        d.weights = r.vec_i32();
        v.push_back(std::move(d));
    }
    return v;
}
static vector<TagNgramData> read_tag_ngrams(Reader& r, bool is_str) {
    size_t l = r.len();
    vector<TagNgramData> v;
    for (size_t i = 0; i < l; ++i) {
        TagNgramData d;
        d.ngram = r.bytes();
        if (is_str && !valid_utf8(d.ngram)) throw Error(DECODE_ERROR, "invalid UTF-8 in model string");
        size_t m = r.len();
        for (size_t j = 0; j < m; ++j) {
            TagWeight w;
            w.rel_position = r.u8();
            w.weights = r.vec_i32();
            d.weights.push_back(std::move(w));
        }
        v.push_back(std::move(d));
    }
    return v;
}

// Model::read_slice (model.rs:127-134)
static Model model_read(const uint8_t* data, size_t n, size_t* consumed) {
    const size_t ml = sizeof(MODEL_MAGIC) - 1;
    if (n < ml || memcmp(data, MODEL_MAGIC, ml) != 0) throw Error(INVALID_MODEL, "model version mismatch");
    Reader r{data + ml, n - ml};
    Model m;
    m.char_ngram_model = read_ngrams(r, true);
    m.type_ngram_model = read_ngrams(r, false);
    size_t nd = r.len();
    for (size_t i = 0; i < nd; ++i) {
        WordWeightRecord w;
        w.word = r.bytes();
        if (!valid_utf8(w.word)) throw Error(DECODE_ERROR, "invalid UTF-8 in model string");
        w.weights = r.vec_i32();
        w.comment = r.bytes();
        m.dict_model.push_back(std::move(w));
    }
    m.bias = int32_t(r.zigzag());
    m.char_window_size = r.u8();
    m.type_window_size = r.u8();
    size_t nt = r.len();
    for (size_t i = 0; i < nt; ++i) {
        TagModel t;
        t.token = r.bytes();
        size_t ns = r.len();
        for (size_t j = 0; j < ns; ++j) {
            size_t nc = r.len();
            vector<string> c;
            for (size_t k = 0; k < nc; ++k) c.push_back(r.bytes());
            t.tags.push_back(std::move(c));
        }
        t.char_ngram_model = read_tag_ngrams(r, true);
        t.type_ngram_model = read_tag_ngrams(r, false);
        t.bias = r.vec_i32();
        m.tag_models.push_back(std::move(t));
    }
    if (consumed) *consumed = ml + r.pos;
    return m;
}

// Model::to_vec (model.rs:99-104): MODEL_MAGIC + bincode `config::standard()` encoding of ModelData
// (varint: < 251 one byte, else marker 251/252/253 + u16/u32/u64 little endian; zigzag for signed).
namespace kw {
static void uvar(string& o, uint64_t v) {
    if (v < 251) { o.push_back(char(v)); return; }
    int w = v <= 0xFFFF ? 2 : v <= 0xFFFFFFFFull ? 4 : 8;
    o.push_back(char(w == 2 ? 251 : w == 4 ? 252 : 253));
    for (int k = 0; k < w; ++k) o.push_back(char((v >> (8 * k)) & 0xFF));
}
static void zz(string& o, int64_t v) { uvar(o, v < 0 ? (uint64_t(~v) << 1) | 1 : uint64_t(v) << 1); }
static void bytes(string& o, const string& b) { uvar(o, b.size()); o += b; }
static void vec_i32(string& o, const vector<int32_t>& v) { uvar(o, v.size()); for (int32_t x : v) zz(o, x); }
}  // namespace kw
static string model_to_vec(const Model& m) {
    string o(MODEL_MAGIC, sizeof(MODEL_MAGIC) - 1);
    for (const auto* ng : {&m.char_ngram_model, &m.type_ngram_model}) {
        kw::uvar(o, ng->size());
        for (const NgramData& d : *ng) { kw::bytes(o, d.ngram); kw::vec_i32(o, d.weights); }
    }
    kw::uvar(o, m.dict_model.size());
    for (const WordWeightRecord& w : m.dict_model) { kw::bytes(o, w.word); kw::vec_i32(o, w.weights); kw::bytes(o, w.comment); }
    kw::zz(o, m.bias);
    o.push_back(char(m.char_window_size));
    o.push_back(char(m.type_window_size));
    kw::uvar(o, m.tag_models.size());
    for (const TagModel& t : m.tag_models) {
        kw::bytes(o, t.token);
        kw::uvar(o, t.tags.size());
        for (const auto& c : t.tags) { kw::uvar(o, c.size()); for (const string& x : c) kw::bytes(o, x); }
        for (const auto* ng : {&t.char_ngram_model, &t.type_ngram_model}) {
            kw::uvar(o, ng->size());
            for (const TagNgramData& d : *ng) {
                kw::bytes(o, d.ngram);
                kw::uvar(o, d.weights.size());
                for (const auto& w : d.weights) { o.push_back(char(w.rel_position)); kw::vec_i32(o, w.weights); }
            }
        }
        kw::vec_i32(o, t.bias);
    }
    return o;
}

// ---------------------------------------------------------------------------
// KyTea model reader and converter (kytea_model.rs), restated struct by struct
// ---------------------------------------------------------------------------
namespace kytea {

struct In {
    const uint8_t* p; size_t n; size_t i = 0;
    void need(size_t k) { if (n - i < k) throw Error(IO_ERROR, "failed to fill whole buffer"); }
    uint8_t u8() { need(1); return p[i++]; }
    template <class T> T le() { need(sizeof(T)); T v; memcpy(&v, p + i, sizeof(T)); i += sizeof(T); return v; }
    // BufRead::read_until / read_line: up to and including the delimiter
    string until(uint8_t d) {
        size_t s = i;
        while (i < n && p[i] != d) ++i;
        if (i < n) ++i;
        return string(reinterpret_cast<const char*>(p + s), i - s);
    }
};

// KyteaConfig (kytea_model.rs:11-63)
struct Config { uint32_t n_tags; uint8_t char_w, type_w, dict_n; vector<uint32_t> char_map; };

static vector<uint32_t> decode_chars(const string& s) {
    vector<uint32_t> out;
    for (size_t i = 0; i < s.size();) {
        uint8_t b = uint8_t(s[i]);
        int l = b < 0x80 ? 1 : b < 0xE0 ? 2 : b < 0xF0 ? 3 : 4;
        uint32_t c = l == 1 ? b : b & (0xFF >> (l + 1));
        for (int k = 1; k < l; ++k) c = (c << 6) | (uint8_t(s[i + k]) & 0x3F);
        out.push_back(c);
        i += size_t(l);
    }
    return out;
}
static Config read_config(In& r) {
    Config c;
    r.until('\n');                       // model_tag
    r.u8(); r.u8();                       // do_ws, do_tags
    c.n_tags = r.le<uint32_t>();
    c.char_w = r.u8(); r.u8();            // char_w, char_n
    c.type_w = r.u8(); r.u8();            // type_w, type_n
    c.dict_n = r.u8();
    r.u8();                               // bias
    r.le<double>();                       // epsilon
    r.u8();                               // solver_type
    string cm = r.until(0);
    if (!valid_utf8(cm)) throw Error(DECODE_ERROR, "invalid utf-8 sequence");
    c.char_map = decode_chars(cm);
    return c;
}
// Readable for char / String (kytea_model.rs:90-130)
static uint32_t read_char(const Config& c, In& r) {
    size_t idx = r.le<uint16_t>();
    if (idx == 0 || idx > c.char_map.size()) throw Error(INVALID_MODEL, "character index out of range");  // (panics)
    return c.char_map[idx - 1];
}
static vector<uint32_t> read_string(const Config& c, In& r) {
    uint32_t size = r.le<uint32_t>();
    vector<uint32_t> s;
    for (uint32_t k = 0; k < size; ++k) s.push_back(read_char(c, r));
    return s;
}
static vector<int16_t> read_vec_i16(const Config&, In& r) {
    uint32_t size = r.le<uint32_t>();
    vector<int16_t> v;
    for (uint32_t k = 0; k < size; ++k) v.push_back(r.le<int16_t>());
    return v;
}

// State / Dictionary<T> (kytea_model.rs:132-217)
struct State { vector<std::pair<uint32_t, uint32_t>> gotos; vector<uint32_t> outputs; bool is_branch; };
template <class T> struct Dictionary {
    bool some = false;
    uint8_t n_dicts = 0;
    vector<State> states;
    vector<T> entries;
    // dump_items (:152-167): explicit stack, gotos pushed in reverse
    vector<std::pair<vector<uint32_t>, const T*>> dump_items() const {
        vector<std::pair<vector<uint32_t>, const T*>> result;
        vector<std::pair<size_t, vector<uint32_t>>> stack{{0, {}}};
        size_t guard = 0;
        while (!stack.empty()) {
            auto [idx, word] = stack.back();
            stack.pop_back();
            if (idx >= states.size() || ++guard > 4 * states.size() + 4) throw Error(INVALID_MODEL, "bad dictionary");
            const State& st = states[idx];
            if (st.is_branch) {
                if (st.outputs.empty() || st.outputs[0] >= entries.size()) throw Error(INVALID_MODEL, "bad dictionary output");
                result.push_back({word, &entries[st.outputs[0]]});
            }
            for (auto it = st.gotos.rbegin(); it != st.gotos.rend(); ++it) {
                vector<uint32_t> w = word;
                w.push_back(it->first);
                stack.push_back({it->second, w});
            }
        }
        return result;
    }
};
template <class T, class F> static Dictionary<T> read_dictionary(const Config& c, In& r, F read_entry) {
    Dictionary<T> d;
    d.n_dicts = r.u8();
    uint32_t n_states = r.le<uint32_t>();
    if (n_states == 0) return d;
    d.some = true;
    for (uint32_t s = 0; s < n_states; ++s) {
        State st;
        r.le<uint32_t>();  // failure
        uint32_t n_gotos = r.le<uint32_t>();
        for (uint32_t g = 0; g < n_gotos; ++g) {
            uint32_t k = read_char(c, r);
            uint32_t v = r.le<uint32_t>();
            st.gotos.push_back({k, v});
        }
        std::sort(st.gotos.begin(), st.gotos.end());
        uint32_t n_outputs = r.le<uint32_t>();
        for (uint32_t o = 0; o < n_outputs; ++o) st.outputs.push_back(r.le<uint32_t>());
        st.is_branch = r.u8() != 0;
        d.states.push_back(std::move(st));
    }
    uint32_t n_entries = r.le<uint32_t>();
    for (uint32_t e = 0; e < n_entries; ++e) d.entries.push_back(read_entry(c, r));
    return d;
}

// FeatureLookup<i16> (:220-262), Option<LinearModel> (:264-300)
struct FeatureLookup {
    Dictionary<vector<int16_t>> char_dict, type_dict, self_dict;
    vector<int16_t> dict_vec, biases, tag_dict_vec, tag_unk_vec;
};
struct LinearModel { bool some = false; bool has_lookup = false; FeatureLookup fl; };
static LinearModel read_linear_model(const Config& c, In& r) {
    LinearModel m;
    uint32_t n_classes = r.le<uint32_t>();
    if (n_classes == 0) return m;
    m.some = true;
    r.u8();                                                   // solver_type
    for (uint32_t k = 0; k < n_classes; ++k) r.le<int32_t>();  // labels
    r.u8();                                                   // bias
    r.le<double>();                                           // multiplier
    if (r.u8() == 0) return m;                                // FeatureLookup::read: active
    m.has_lookup = true;
    m.fl.char_dict = read_dictionary<vector<int16_t>>(c, r, read_vec_i16);
    m.fl.type_dict = read_dictionary<vector<int16_t>>(c, r, read_vec_i16);
    m.fl.self_dict = read_dictionary<vector<int16_t>>(c, r, read_vec_i16);
    m.fl.dict_vec = read_vec_i16(c, r);
    m.fl.biases = read_vec_i16(c, r);
    m.fl.tag_dict_vec = read_vec_i16(c, r);
    m.fl.tag_unk_vec = read_vec_i16(c, r);
    return m;
}

// ModelTagEntry (:302-342), ProbTagEntry (:344-377)
struct ModelTagEntry { uint8_t in_dict; };
static ModelTagEntry read_model_tag_entry(const Config& c, In& r) {
    read_string(c, r);  // word
    for (uint32_t t = 0; t < c.n_tags; ++t) {
        uint32_t size = r.le<uint32_t>();
        for (uint32_t k = 0; k < size; ++k) { read_string(c, r); r.u8(); }
    }
    ModelTagEntry e{r.u8()};
    for (uint32_t t = 0; t < c.n_tags; ++t) read_linear_model(c, r);
    return e;
}
struct ProbTagEntry {};
static ProbTagEntry read_prob_tag_entry(const Config& c, In& r) {
    read_string(c, r);
    for (uint32_t t = 0; t < c.n_tags; ++t) {
        uint32_t size = r.le<uint32_t>();
        for (uint32_t k = 0; k < size; ++k) { read_string(c, r); r.le<double>(); }
    }
    return {};
}

static string to_utf8(const vector<uint32_t>& w) {  // chars -> String
    string s;
    for (uint32_t c : w) {
        if (c < 0x80) s.push_back(char(c));
        else if (c < 0x800) { s.push_back(char(0xC0 | (c >> 6))); s.push_back(char(0x80 | (c & 0x3F))); }
        else if (c < 0x10000) {
            s.push_back(char(0xE0 | (c >> 12))); s.push_back(char(0x80 | ((c >> 6) & 0x3F))); s.push_back(char(0x80 | (c & 0x3F)));
        } else {
            s.push_back(char(0xF0 | (c >> 18))); s.push_back(char(0x80 | ((c >> 12) & 0x3F)));
            s.push_back(char(0x80 | ((c >> 6) & 0x3F))); s.push_back(char(0x80 | (c & 0x3F)));
        }
    }
    return s;
}

// KyteaModel::read (:423-450) followed by TryFrom<KyteaModel> for Model (:453-550)
static Model convert(const uint8_t* data, size_t n) {
    In r{data, n};
    Config config = read_config(r);
    LinearModel wordseg_model = read_linear_model(config, r);
    for (uint32_t t = 0; t < config.n_tags; ++t) {
        uint32_t size = r.le<uint32_t>();                       // global_tags: Vec<String>
        for (uint32_t k = 0; k < size; ++k) read_string(config, r);
        read_linear_model(config, r);                           // global_models
    }
    auto dict = read_dictionary<ModelTagEntry>(config, r, read_model_tag_entry);
    read_dictionary<ProbTagEntry>(config, r, read_prob_tag_entry);

    if (!wordseg_model.some) throw Error(INVALID_MODEL, "no word segmentation model.");
    if (!wordseg_model.has_lookup) throw Error(INVALID_MODEL, "no lookup data.");
    const FeatureLookup& fl = wordseg_model.fl;
    if (fl.biases.empty()) throw Error(INVALID_MODEL, "no bias.");  // (index panic in the reference)
    Model m;
    m.bias = fl.biases[0];
    if (!fl.char_dict.some) throw Error(INVALID_MODEL, "no character dictionary.");
    if (!fl.type_dict.some) throw Error(INVALID_MODEL, "no type dictionary.");
    auto slice = [](const vector<int16_t>& v, size_t window, size_t len) {
        if (len > 2 * window + 1 || 2 * window + 1 - len > v.size()) throw Error(INVALID_MODEL, "weight slice out of range");
        return vector<int32_t>(v.begin(), v.begin() + long(2 * window + 1 - len));
    };
    for (auto& [ngram, v] : fl.char_dict.dump_items())
        m.char_ngram_model.push_back(NgramData{to_utf8(ngram), slice(*v, config.char_w, ngram.size())});
    for (auto& [ngram, v] : fl.type_dict.dump_items()) {
        string bytes = to_utf8(ngram);
        bool skip = false;
        for (char& t : bytes) {
            switch (uint8_t(t)) {
                case 'D': t = 1; break;
                case 'R': t = 2; break;
                case 'H': t = 3; break;
                case 'T': t = 4; break;
                case 'K': t = 5; break;
                case 'O': t = 6; break;
                case 4: skip = true; break;  // vaporetto issue #110
                default: throw Error(INVALID_MODEL, "unsupported character type: " + std::to_string(int(uint8_t(t))));
            }
            if (skip) break;
        }
        if (skip) continue;
        m.type_ngram_model.push_back(NgramData{bytes, slice(*v, config.type_w, ngram.size())});
    }
    if (dict.some) {
        for (auto& [w, data] : dict.dump_items()) {
            if (w.empty() || config.dict_n == 0) throw Error(INVALID_MODEL, "empty dictionary word");
            size_t idx = std::min<size_t>(w.size(), config.dict_n) - 1;
            int32_t left = 0, inside = 0, right = 0;
            for (size_t j = 0; j < dict.n_dicts; ++j) {
                if (j < 8 && ((data->in_dict >> j) & 1) == 1) {
                    size_t offset = 3 * size_t(config.dict_n) * j + 3 * idx;
                    if (offset + 2 >= fl.dict_vec.size()) throw Error(INVALID_MODEL, "dict_vec index out of range");
                    left += fl.dict_vec[offset];
                    inside += fl.dict_vec[offset + 1];
                    right += fl.dict_vec[offset + 2];
                }
            }
            vector<int32_t> weights(w.size() + 1, inside);
            weights.front() = left;
            weights.back() = right;
            m.dict_model.push_back(WordWeightRecord{to_utf8(w), weights, ""});
        }
    }
    m.char_window_size = config.char_w;
    m.type_window_size = config.type_w;
    return m;
}

}  // namespace kytea

// ---------------------------------------------------------------------------
// PositionalWeight (predictor.rs:138-165) and the tag variant (:217-262)
// ---------------------------------------------------------------------------
static inline int32_t wadd(int32_t a, int32_t b) { return int32_t(uint32_t(a) + uint32_t(b)); }

struct PW {
    int offset = 0;  // i16 in the reference
    vector<int32_t> weight;
    // AddAssign (predictor.rs:149-165)
    void add(const PW& o) {
        int new_offset = std::min(offset, o.offset);
        size_t shift = size_t(offset - new_offset);
        size_t oshift = size_t(o.offset - new_offset);
        size_t new_size = std::max(shift + weight.size(), oshift + o.weight.size());
        weight.resize(new_size, 0);
        std::rotate(weight.begin(), weight.end() - (new_size ? shift % new_size : 0), weight.end());
        for (size_t i = 0; i < o.weight.size() && oshift + i < weight.size(); ++i)
            weight[oshift + i] = wadd(weight[oshift + i], o.weight[i]);
        offset = new_offset;
    }
};

// Test switch (ora_set_states_only): build a tag predictor whose patterns, pattern ids and boundary weights are the
// reference's, but skip copying the tag weights along the suffix chains (the literal restatement of the build-time
// merge takes minutes on 20 000 tag models).  Scores, boundaries and pattern-id states are unaffected; predict_tags of
// such a predictor is meaningless.
static bool g_states_only = false;

struct PWTag {
    std::optional<PW> weight;
    std::map<std::pair<size_t, uint8_t>, vector<int32_t>> tag_info;
    // AddAssign (predictor.rs:242-262)
    void add(const PWTag& o) {
        if (weight) {
            if (o.weight) weight->add(*o.weight);
        } else {
            weight = o.weight;
        }
        if (g_states_only) return;
        for (auto& kv : o.tag_info) {
            auto it = tag_info.find(kv.first);
            if (it == tag_info.end()) {
                tag_info[kv.first] = kv.second;
            } else {
                auto& w = it->second;
                for (size_t i = 0; i < w.size() && i < kv.second.size(); ++i) w[i] = wadd(w[i], kv.second[i]);
            }
        }
    }
};

// start offsets of the proper suffixes of `s` that begin on a symbol boundary
static vector<size_t> suffix_starts(const string& s, bool utf8) {
    vector<size_t> v;
    for (size_t j = 1; j < s.size(); ++j)
        if (!utf8 || (uint8_t(s[j]) & 0xC0) != 0x80) v.push_back(j);
    return v;
}

// CharWeightMerger / TypeWeightMerger (char_scorer.rs:29-79, type_scorer.rs:38-88)
template <class W>
struct Merger {
    struct Cell { W w; bool done = false; };
    std::map<string, Cell> map;  // BTreeMap: byte-lexicographic order
    bool utf8;
    explicit Merger(bool u) : utf8(u) {}
    void add(const string& ngram, const W& w) {
        auto it = map.find(ngram);
        if (it != map.end()) it->second.w.add(w);
        else map.emplace(ngram, Cell{w, false});
    }
    vector<std::pair<string, W>> merge() {
        vector<Cell*> stack;
        for (auto& kv : map) {
            if (kv.second.done) continue;
            stack.push_back(&kv.second);
            for (size_t j : suffix_starts(kv.first, utf8)) {
                auto it = map.find(kv.first.substr(j));
                if (it != map.end()) {
                    stack.push_back(&it->second);
                    if (it->second.done) break;
                }
            }
            Cell* from = stack.back();
            stack.pop_back();
            from->done = true;
            while (!stack.empty()) {
                Cell* to = stack.back();
                stack.pop_back();
                to->done = true;
                to->w.add(from->w);
                from = to;
            }
        }
        vector<std::pair<string, W>> out;
        for (auto& kv : map) out.emplace_back(kv.first, std::move(kv.second.w));
        return out;
    }
};

// ---------------------------------------------------------------------------
// Aho-Corasick over symbols (code points for the char scorer — the reference's default
// `charwise-pma` feature — or type bytes for the type scorer).  Semantics of daachorse; see header.
// goto function: open-addressing hash (state, symbol) -> state; classic failure links.
// ---------------------------------------------------------------------------
struct AC {
    // Double-array form of the finished automaton (the layout daachorse's CharwiseDoubleArrayAhoCorasick walks:
    // 16-byte states {base, check, fail, output}, child = base XOR code, code = the symbol's rank by frequency from a
    // mapper table, bases unique so that `check == code` identifies the edge).  Same transitions as the hash form below
    // -- the layout does not affect results -- but one state record per transition instead of two hash-table lines:
    // this is what the timed CPU arm walks, so that it is not slower than the reference for a reason the reference
    // does not have.  ORA_AC=hash in the environment keeps the hash walk (tests compare the two).
    struct DAState { uint32_t base, check, fail; int32_t out; };
    static constexpr uint32_t kNone = 0xFFFFFFFFu;
    vector<DAState> da;
    vector<uint32_t> code_of;  // symbol -> code (1-based), 0 = the symbol occurs in no pattern
    bool use_da = false;
    vector<int32_t> fail, out, own, depth;
    vector<uint64_t> hkey;   // (state << 21 | symbol) + 1, 0 = empty
    vector<int32_t> hval;
    uint64_t hmask = 0;
    static inline uint64_t hmix(uint64_t x) {
        x ^= x >> 29; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 32;
        return x;
    }
    inline int find(int s, uint32_t c) const {
        const uint64_t k = ((uint64_t(uint32_t(s)) << 21) | c) + 1;
        for (uint64_t i = hmix(k) & hmask;; i = (i + 1) & hmask) {
            if (hkey[i] == k) return hval[i];
            if (hkey[i] == 0) return -1;
        }
    }
    void insert(int s, uint32_t c, int v) {
        const uint64_t k = ((uint64_t(uint32_t(s)) << 21) | c) + 1;
        uint64_t i = hmix(k) & hmask;
        while (hkey[i] != 0) i = (i + 1) & hmask;
        hkey[i] = k;
        hval[i] = v;
    }
    static vector<uint32_t> symbols(const string& p, bool utf8) {
        vector<uint32_t> v;
        if (!utf8) { for (unsigned char c : p) v.push_back(c); return v; }
        for (size_t i = 0; i < p.size();) {
            uint8_t b = uint8_t(p[i]);
            int l = b < 0x80 ? 1 : b < 0xE0 ? 2 : b < 0xF0 ? 3 : 4;
            uint32_t cp = l == 1 ? b : l == 2 ? (b & 0x1F) : l == 3 ? (b & 0x0F) : (b & 0x07);
            for (int k = 1; k < l; ++k) cp = (cp << 6) | (uint8_t(p[i + k]) & 0x3F);
            v.push_back(cp);
            i += size_t(l);
        }
        return v;
    }
    void build(const vector<string>& pats, bool utf8) {
        size_t total = 1;
        for (auto& p : pats) total += p.size();
        size_t cap = 16;
        while (cap < total * 2) cap <<= 1;
        hkey.assign(cap, 0);
        hval.assign(cap, 0);
        hmask = cap - 1;
        fail.assign(1, 0); out.assign(1, -1); own.assign(1, -1); depth.assign(1, 0);
        vector<vector<std::pair<uint32_t, int>>> kids(1);
        for (size_t i = 0; i < pats.size(); ++i) {
            if (pats[i].empty()) throw Error(INVALID_MODEL, "failed to build the automaton");
            int s = 0;
            for (uint32_t c : symbols(pats[i], utf8)) {
                int nx = find(s, c);
                if (nx < 0) {
                    nx = int(fail.size());
                    fail.push_back(0); out.push_back(-1); own.push_back(-1); depth.push_back(depth[s] + 1);
                    kids.emplace_back();
                    kids[s].emplace_back(c, nx);
                    insert(s, c, nx);
                }
                s = nx;
            }
            if (own[s] >= 0) throw Error(INVALID_MODEL, "failed to build the automaton");  // duplicate pattern
            own[s] = int(i);
        }
        vector<int> q;
        for (auto& e : kids[0]) { fail[e.second] = 0; q.push_back(e.second); }
        for (size_t h = 0; h < q.size(); ++h) {
            const int s = q[h];
            out[s] = own[s] >= 0 ? own[s] : out[fail[s]];
            for (auto& e : kids[s]) {
                int t = fail[s], nx;
                while ((nx = find(t, e.first)) < 0 && t != 0) t = fail[t];
                fail[e.second] = nx < 0 ? 0 : nx;
                q.push_back(e.second);
            }
        }
        {
            const char* e = std::getenv("ORA_AC");
            use_da = !(e && string(e) == "hash");
            if (use_da) build_da(kids);
        }
        // rehash compactly (2 slots per transition) and pack key+value side by side for the walk
        size_t ncap = 16;
        while (ncap < fail.size() * 2) ncap <<= 1;
        vector<uint64_t> ok;
        vector<int32_t> ov;
        ok.swap(hkey);
        ov.swap(hval);
        hkey.assign(ncap, 0);
        hval.assign(ncap, 0);
        hmask = ncap - 1;
        for (size_t i = 0; i < ok.size(); ++i) {
            if (!ok[i]) continue;
            uint64_t j = hmix(ok[i]) & hmask;
            while (hkey[j] != 0) j = (j + 1) & hmask;
            hkey[j] = ok[i];
            hval[j] = ov[i];
        }
    }
    // Lays the automaton out as a double array: states in breadth-first order, every state's children at base XOR code.
    void build_da(const vector<vector<std::pair<uint32_t, int>>>& kids) {
        // codes by symbol frequency over the edges (frequent symbols get small codes: children stay close to the base)
        std::unordered_map<uint32_t, uint64_t> freq;
        uint32_t max_sym = 0;
        for (auto& ks : kids) for (auto& e : ks) { ++freq[e.first]; max_sym = std::max(max_sym, e.first); }
        vector<std::pair<uint64_t, uint32_t>> order;
        for (auto& kv : freq) order.emplace_back(kv.second, kv.first);
        std::sort(order.begin(), order.end(), [](auto& a, auto& b) { return a.first != b.first ? a.first > b.first : a.second < b.second; });
        code_of.assign(size_t(max_sym) + 1, 0);
        for (size_t i = 0; i < order.size(); ++i) code_of[order[i].second] = uint32_t(i + 1);
        uint32_t span = 1;  // XOR with a code changes only the bits below `span`
        while (span <= order.size()) span <<= 1;
        const size_t n = kids.size();
        vector<uint32_t> pos(n, kNone);  // trie state -> double-array index
        vector<uint8_t> occ, base_used;
        // free slots as a doubly linked list in ascending order (index `cap` = list end); a search that does not find
        // a base among the first kTries free slots opens a fresh block of `span` slots (keeps the build linear)
        vector<uint32_t> nxt, prv;
        uint32_t head = kNone, tail = kNone;
        auto grow = [&]() {
            const size_t old = da.size();
            da.resize(old + span, DAState{kNone, kNone, 0, -1});
            occ.resize(old + span, 0);
            base_used.resize(old + span, 0);
            nxt.resize(old + span, kNone);
            prv.resize(old + span, kNone);
            for (size_t i = old; i < old + span; ++i) {
                prv[i] = tail;
                if (tail != kNone) nxt[tail] = uint32_t(i); else head = uint32_t(i);
                tail = uint32_t(i);
            }
            return uint32_t(old);
        };
        auto take = [&](uint32_t i) {
            occ[i] = 1;
            if (prv[i] != kNone) nxt[prv[i]] = nxt[i]; else head = nxt[i];
            if (nxt[i] != kNone) prv[nxt[i]] = prv[i]; else tail = prv[i];
        };
        da.clear();
        grow();
        pos[0] = 0;
        take(0);
        constexpr int kTries = 16384;  // (fill 60-90 % on the bench models; the build stays at seconds)
        uint32_t cursor = kNone;
        vector<int> q{0};
        vector<uint32_t> codes;
        for (size_t h = 0; h < q.size(); ++h) {
            const int st = q[h];
            if (kids[st].empty()) continue;
            codes.clear();
            for (auto& e : kids[st]) codes.push_back(code_of[e.first]);
            uint32_t base = kNone;
            int tries = 0;
            bool wrapped = false;
            // next fit: the search resumes where the last one ended (slots in front of the cursor were rejected by recent
            // states; single-child states still come back to them through the wrap-around at the list end)
            // (a taken slot keeps its forward link: following the links from it reaches the next free slot behind it)
            while (cursor != kNone && occ[cursor]) cursor = nxt[cursor];
            if (cursor == kNone) cursor = head;
            for (uint32_t f = cursor;; f = nxt[f]) {
                if (f == kNone && !wrapped) { f = head; wrapped = true; }
                if (f == kNone || ++tries > kTries) f = grow();  // (a fresh block: every slot free, no base used)
                cursor = f;
                const uint32_t b = f ^ codes[0];
                if (base_used[b]) continue;
                bool ok = true;
                for (uint32_t c : codes)
                    if (occ[b ^ c]) { ok = false; break; }
                if (ok) { base = b; break; }
            }
            base_used[base] = 1;
            da[pos[st]].base = base;
            for (size_t k = 0; k < kids[st].size(); ++k) {
                const uint32_t idx = base ^ codes[k];
                take(idx);
                da[idx].check = codes[k];
                pos[kids[st][k].second] = idx;
                q.push_back(kids[st][k].second);
            }
        }
        if (std::getenv("ORA_AC_DEBUG")) fprintf(stderr, "[oracle] double array: %zu states in %zu slots\n", n, da.size());
        for (size_t st = 0; st < n; ++st) {
            da[pos[st]].fail = pos[fail[st]];
            da[pos[st]].out = out[st];
        }
    }
    inline int da_child(int s, uint32_t code) const {
        const uint32_t b = da[s].base;
        if (b == kNone) return -1;
        const uint32_t idx = b ^ code;
        return idx < da.size() && da[idx].check == code ? int(idx) : -1;
    }
    inline int step_da(int s, uint32_t c) const {
        const uint32_t code = c < code_of.size() ? code_of[c] : 0u;
        if (code == 0) return 0;  // a symbol outside every pattern: all the way back to the root
        for (;;) {
            const int nx = da_child(s, code);
            if (nx >= 0) return nx;
            if (s == 0) return 0;
            s = int(da[s].fail);
        }
    }
    inline int step_hash(int s, uint32_t c) const {
        for (;;) {
            const int nx = find(s, c);
            if (nx >= 0) return nx;
            if (s == 0) return 0;
            s = fail[s];
        }
    }
    // (state numbers are indices of whichever form is walked: callers only feed them back)
    inline int step(int s, uint32_t c) const { return use_da ? step_da(s, c) : step_hash(s, c); }
    inline int out_of(int s) const { return use_da ? da[s].out : out[s]; }
};

// CharacterType::get_type (sentence.rs:50-67)
static inline uint8_t get_type(uint32_t c) {
    if ((c >= 0x30 && c <= 0x39) || (c >= 0xFF10 && c <= 0xFF19)) return 1;
    if ((c >= 0x41 && c <= 0x5A) || (c >= 0x61 && c <= 0x7A) || (c >= 0xFF21 && c <= 0xFF3A) ||
        (c >= 0xFF41 && c <= 0xFF5A)) return 2;
    if (c >= 0x3040 && c <= 0x3096) return 3;
    if ((c >= 0x30A0 && c <= 0x30FA) || (c >= 0x30FC && c <= 0x30FF) || (c >= 0xFF66 && c <= 0xFF9F)) return 4;
    if ((c >= 0x3400 && c <= 0x4DBF) || (c >= 0x4E00 && c <= 0x9FFF) || (c >= 0xF900 && c <= 0xFAFF) ||
        (c >= 0x20000 && c <= 0x2A6DF) || (c >= 0x2A700 && c <= 0x2B73F) || (c >= 0x2B740 && c <= 0x2B81F) ||
        (c >= 0x2B820 && c <= 0x2CEAF) || (c >= 0x2F800 && c <= 0x2FA1F)) return 5;
    return 6;
}

// Sentence state touched by predict (sentence.rs:85-101)
struct Sentence {
    string text;
    vector<uint32_t> chars;  // code points
    vector<uint8_t> char_types;
    vector<uint8_t> boundaries;
    vector<int32_t> boundary_scores;  // padded strip
    size_t score_padding = 0;
    vector<uint32_t> char_pma_states, type_pma_states;
    vector<uint32_t> str_to_char_pos, char_to_str_pos;
    size_t len() const { return char_types.size(); }
    // Sentence::parse_raw (sentence.rs:160-196); text must be valid UTF-8
    void parse_raw(const char* s, size_t nbytes) {
        text.assign(s, nbytes);
        chars.clear(); char_types.clear(); boundaries.clear(); str_to_char_pos.clear(); char_to_str_pos.clear();
        boundary_scores.clear(); char_pma_states.clear(); type_pma_states.clear(); score_padding = 0;
        if (!valid_utf8(text)) throw Error(INVALID_ARGUMENT, "InvalidArgumentError: text: must be valid UTF-8");
        char_to_str_pos.push_back(0);
        size_t pos = 0;
        while (pos < nbytes) {
            uint8_t b = uint8_t(s[pos]);
            uint32_t cp; int l;
            if (b < 0x80) { cp = b; l = 1; }
            else if (b < 0xE0) { cp = b & 0x1F; l = 2; }
            else if (b < 0xF0) { cp = b & 0x0F; l = 3; }
            else { cp = b & 0x07; l = 4; }
            for (int k = 1; k < l; ++k) cp = (cp << 6) | (uint8_t(s[pos + k]) & 0x3F);
            if (cp == 0) throw Error(INVALID_ARGUMENT, "InvalidArgumentError: text: must not contain NULL");
            chars.push_back(cp);
            char_types.push_back(get_type(cp));
            pos += l;
            char_to_str_pos.push_back(uint32_t(pos));
        }
        if (char_types.empty())
            throw Error(INVALID_ARGUMENT, "InvalidArgumentError: text: must contain at least one character");
        str_to_char_pos.assign(pos + 1, 0);
        for (size_t i = 0; i < char_to_str_pos.size(); ++i) str_to_char_pos[char_to_str_pos[i]] = uint32_t(i);
        boundaries.assign(char_types.size() - 1, 2);  // Unknown
    }
};

// PositionalWeight<WeightVector>::add_score (predictor.rs:176-213).  The Fixed
// branch is the Variable branch on a zero-padded vector, so one ragged add with
// clipping at both ends of the strip covers both (the Fixed branch never clips
// for models the reference accepts; it would panic otherwise).
static inline void add_score(const PW& w, long end, vector<int32_t>& ys) {
    long pos = end + w.offset;
    long n = long(ys.size());
    for (long k = 0; k < long(w.weight.size()); ++k) {
        long p = pos + k;
        if (p < 0) continue;
        if (p >= n) break;
        ys[p] = wadd(ys[p], w.weight[k]);
    }
}

struct TagTable {  // tag_weight[token_id][rel_position] : pattern id -> weights
    vector<vector<std::unordered_map<uint32_t, vector<int32_t>>>> t;
};

// WeightVector::add_scores (predictor.rs:81-107): zip truncates; Fixed pads to 8.
static inline void add_tag_vec(const vector<int32_t>& w, vector<int32_t>& ys) {
    for (size_t i = 0; i < w.size() && i < ys.size(); ++i) ys[i] = wadd(ys[i], w[i]);
}

// One automaton-backed scorer: CharScorerBoundary / CharScorerBoundaryTag /
// TypeScorerBoundary / TypeScorerBoundaryTag
// (char_scorer/boundary_scorer.rs:56-113, char_scorer/boundary_tag_scorer.rs:62-174,
//  type_scorer/boundary_scorer.rs:45-80, type_scorer/boundary_tag_scorer.rs:51-143)
struct PmaScorer {
    AC pma;
    vector<std::optional<PW>> weights;
    // flat copy of `weights` for the hot loop: offset, begin, length (len < 0: no boundary weight)
    vector<int32_t> w_off, w_len;
    vector<uint32_t> w_begin;
    vector<int32_t> w_data;
    // one record per pattern for the walk: short rows inline (the reference's WeightVector::Fixed([i32; 8]) lives inside
    // its Option<PositionalWeight> the same way, predictor.rs:60-70): one cache line per match instead of four arrays
    struct WRec { int32_t off, len; uint32_t begin; int32_t fixed[8]; uint32_t pad; };
    vector<WRec> w_rec;
    TagTable tag_weight;
    bool tag_variant = false;
    bool is_char = true;
    vector<string> pattern_strings;

    void build(const vector<NgramData>& ngrams, const vector<WordWeightRecord>& dict, uint8_t window,
               const vector<const vector<TagNgramData>*>& tag_ngrams, bool is_char_) {
        is_char = is_char_;
        tag_variant = !tag_ngrams.empty();
        Merger<PWTag> merger(is_char);
        for (auto& d : ngrams) {
            PWTag w; w.weight = PW{-int(window), d.weights};
            merger.add(d.ngram, w);
        }
        for (auto& d : dict) {
            size_t word_len = 0;
            for (unsigned char c : d.word) if ((c & 0xC0) != 0x80) ++word_len;
            if (word_len > 32767)
                throw Error(INVALID_MODEL, "words must be shorter than or equal to 32767 characters");
            PWTag w; w.weight = PW{-int(word_len), d.weights};
            merger.add(d.word, w);
        }
        tag_weight.t.assign(tag_ngrams.size(),
                            vector<std::unordered_map<uint32_t, vector<int32_t>>>(size_t(window) + 1));
        for (size_t i = 0; i < tag_ngrams.size(); ++i)
            for (auto& d : *tag_ngrams[i])
                for (auto& w : d.weights) {
                    PWTag pw; pw.tag_info[{i, w.rel_position}] = w.weights;
                    merger.add(d.ngram, pw);
                }
        vector<string> pats;
        auto merged = merger.merge();
        for (size_t i = 0; i < merged.size(); ++i) {
            pats.push_back(merged[i].first);
            weights.push_back(merged[i].second.weight);
            for (auto& kv : merged[i].second.tag_info) {
                if (kv.first.second >= tag_weight.t[kv.first.first].size())
                    throw Error(INVALID_MODEL, "tag rel_position exceeds the window size");  // ref: index panic
                tag_weight.t[kv.first.first][kv.first.second][uint32_t(i)] = kv.second;
            }
        }
        pattern_strings = pats;
        pma.build(pats, is_char);
        for (auto& w : weights) {
            w_begin.push_back(uint32_t(w_data.size()));
            if (w) { w_off.push_back(w->offset); w_len.push_back(int32_t(w->weight.size())); w_data.insert(w_data.end(), w->weight.begin(), w->weight.end()); }
            else { w_off.push_back(0); w_len.push_back(-1); }
        }
        w_rec.assign(weights.size(), WRec{});
        for (size_t i = 0; i < weights.size(); ++i) {
            WRec& r = w_rec[i];
            r.off = w_off[i]; r.len = w_len[i]; r.begin = w_begin[i];
            for (int k = 0; k < 8; ++k) r.fixed[k] = (k < w_len[i]) ? w_data[w_begin[i] + size_t(k)] : 0;
        }
    }

    // add_scores: walk, longest match per end position (boundary_scorer.rs:93-113 etc.)
    void add_scores(Sentence& s) const {
        vector<uint32_t>* states = nullptr;
        if (tag_variant) {
            states = is_char ? &s.char_pma_states : &s.type_pma_states;
            states->assign(s.len(), 0xFFFFFFFFu);
        }
        int st = 0;
        const size_t n = s.len();
        const long ny = long(s.boundary_scores.size());
        int32_t* ys = s.boundary_scores.data();
        for (size_t i = 0; i < n; ++i) {
            st = pma.step(st, is_char ? s.chars[i] : uint32_t(s.char_types[i]));
            const int p = pma.out_of(st);
            if (p < 0) continue;
            const size_t end = i + 1;  // char index (== str_to_char_pos[m.end()] of the bytewise reference path)
            const WRec& wr = w_rec[size_t(p)];
            if (wr.len >= 0) {
                // PositionalWeight::add_score (predictor.rs:176-213), ragged add clipped at both strip ends
                const long pos = long(end + s.score_padding) - 1 + wr.off;
                const int32_t* w = wr.len <= 8 ? wr.fixed : w_data.data() + wr.begin;
                if (pos >= 0 && pos + wr.len <= ny) {
                    // the row lies inside the strip (always, for the reference's Fixed rows over its padded strip):
                    // straight adds, no per-element clipping
                    int32_t* y = ys + pos;
                    for (long k = 0; k < wr.len; ++k) y[k] = wadd(y[k], w[k]);
                } else {
                    for (long k = 0; k < wr.len; ++k) {
                        const long q = pos + k;
                        if (q < 0) continue;
                        if (q >= ny) break;
                        ys[q] = wadd(ys[q], w[k]);
                    }
                }
            }
            if (states) (*states)[end - 1] = uint32_t(p);
        }
    }

    // add_tag_scores (boundary_tag_scorer.rs:154-174 / :123-143)
    void add_tag_scores(uint32_t token_id, size_t pos, const vector<uint32_t>& states,
                        vector<int32_t>& scores) const {
        const auto& tw = tag_weight.t[token_id];
        for (size_t r = 0; r < tw.size() && pos + r < states.size(); ++r) {
            auto it = tw[r].find(states[pos + r]);
            if (it != tw[r].end()) add_tag_vec(it->second, scores);
        }
    }
};

// TypeScorerBoundaryCache (type_scorer/boundary_scorer_cache.rs:22-110)
struct TypeCache {
    vector<int32_t> scores;
    uint8_t window = 0;
    size_t mask = 0;
    void build(const vector<NgramData>& model, uint8_t w) {
        {   // DoubleArrayAhoCorasick::new rejects empty and duplicate patterns
            std::map<string, int> seen;
            for (auto& d : model) {
                if (d.ngram.empty() || seen.count(d.ngram))
                    throw Error(INVALID_MODEL, "invalid character type n-grams");
                seen[d.ngram] = 1;
            }
        }
        window = w;
        size_t seq = size_t(w) * 2;
        size_t all = size_t(1) << (3 * seq);
        mask = all - 1;
        scores.assign(all, 0);
        vector<uint8_t> s(seq);
        for (size_t id = 0; id < all; ++id) {
            size_t x = id; bool ok = true;
            for (size_t k = seq; k-- > 0;) { s[k] = uint8_t(x & 7); if (s[k] == 7) { ok = false; break; } x >>= 3; }
            if (!ok) continue;
            int32_t y = 0;
            // find_overlapping_iter: every occurrence of every n-gram
            for (auto& d : model) {
                size_t L = d.ngram.size();
                for (size_t end = L; end <= seq; ++end) {
                    if (memcmp(d.ngram.data(), s.data() + end - L, L) != 0) continue;
                    size_t idx = seq - end;
                    if (idx < d.weights.size()) y = wadd(y, d.weights[idx]);
                }
            }
            scores[id] = y;
        }
    }
    void add_scores(Sentence& s) const {
        s.type_pma_states.clear();
        size_t seqid = 0;
        for (size_t i = 0; i < window; ++i)
            seqid = ((seqid << 3) | (i < s.char_types.size() ? s.char_types[i] : 0)) & mask;
        for (size_t i = 0; i < s.boundaries.size(); ++i) {
            size_t j = i + window;
            seqid = ((seqid << 3) | (j < s.char_types.size() ? s.char_types[j] : 0)) & mask;
            int32_t& y = s.boundary_scores[s.score_padding + i];
            y = wadd(y, scores[seqid]);
        }
    }
};

struct TagPredictor {  // predictor.rs:264-304
    vector<vector<string>> tags;
    vector<int32_t> bias;  // WeightVector::from: zero-padded to >= 8 (predictor.rs:118-135)
};

struct Predictor {
    std::unique_ptr<PmaScorer> char_scorer, type_pma;
    std::unique_ptr<TypeCache> type_cache;
    int32_t bias = 0;
    bool predict_tags = false;
    std::unordered_map<string, std::pair<uint32_t, TagPredictor>> tag_predictor;
    size_t n_tags = 0;

    // Predictor::new (predictor.rs:450-508), CharScorer::new (char_scorer.rs:92-124),
    // TypeScorer::new (type_scorer.rs:104-143)
    Predictor(const Model& m, bool ptags) : bias(m.bias), predict_tags(ptags) {
        vector<const vector<TagNgramData>*> tag_char, tag_type;
        if (ptags) {
            for (size_t i = 0; i < m.tag_models.size(); ++i) {
                auto& t = m.tag_models[i];
                n_tags = std::max(n_tags, t.tags.size());
                TagPredictor tp{t.tags, t.bias};
                if (tp.bias.size() < 8) tp.bias.resize(8, 0);
                tag_predictor[t.token] = {uint32_t(i), std::move(tp)};
                tag_char.push_back(&t.char_ngram_model);
                tag_type.push_back(&t.type_ngram_model);
            }
        }
        if (!((m.char_ngram_model.empty() && m.dict_model.empty()) || m.char_window_size == 0)) {
            char_scorer.reset(new PmaScorer());
            char_scorer->build(m.char_ngram_model, m.dict_model, m.char_window_size, tag_char, true);
        }
        if (!(m.type_ngram_model.empty() || m.type_window_size == 0)) {
            if (tag_type.empty() && m.type_window_size <= 3) {
                type_cache.reset(new TypeCache());
                type_cache->build(m.type_ngram_model, m.type_window_size);
            } else {
                type_pma.reset(new PmaScorer());
                type_pma->build(m.type_ngram_model, {}, m.type_window_size, tag_type, false);
            }
        }
    }

    // Predictor::predict (predictor.rs:518-543)
    void predict(Sentence& s) const {
        s.score_padding = 7;
        s.boundary_scores.assign(s.score_padding * 2 + s.len() - 1, bias);
        if (char_scorer) char_scorer->add_scores(s);
        if (type_cache) type_cache->add_scores(s);
        if (type_pma) type_pma->add_scores(s);
        for (size_t i = 0; i < s.boundaries.size(); ++i)
            s.boundaries[i] = s.boundary_scores[s.score_padding + i] > 0 ? 1 : 0;
    }

    // TagPredictor::predict (predictor.rs:286-304): tag index per slot, -1 = None
    static void pick(const TagPredictor& tp, const vector<int32_t>& scores, int32_t* out, size_t n_out) {
        size_t offset = 0;
        for (size_t k = 0; k < tp.tags.size() && k < n_out; ++k) {
            size_t nc = tp.tags[k].size();
            if (nc >= 2) {
                size_t idx = 0; int32_t mx = INT32_MIN;
                for (size_t i = 0; i < nc; ++i) {
                    int32_t sc = scores.at(offset + i);
                    if (sc > mx) { idx = i; mx = sc; }
                }
                out[k] = int32_t(idx);
                offset += nc;
            } else {
                out[k] = nc == 1 ? 0 : -1;
            }
        }
    }

    // Predictor::predict_tags (predictor.rs:546-637).  Output per char position i,
    // slot k: tag_token[i] = token id owning the tags (or -1), tag_idx[i*n_tags+k] =
    // candidate index (or -1).
    void fill_tags(const Sentence& s, vector<int32_t>& tag_token, vector<int32_t>& tag_idx,
                   vector<vector<int32_t>>* raw_scores) const {
        if (!predict_tags) throw Error(INVALID_ARGUMENT, "this predictor is created with predict_tags = false");
        tag_token.assign(s.len(), -1);
        tag_idx.assign(s.len() * n_tags, -1);
        if (raw_scores) raw_scores->assign(s.len(), {});
        if (n_tags == 0) return;
        auto run = [&](size_t start, size_t i) {  // token = chars [start, i], last char index i
            string tok = s.text.substr(s.char_to_str_pos[start], s.char_to_str_pos[i + 1] - s.char_to_str_pos[start]);
            auto it = tag_predictor.find(tok);
            if (it == tag_predictor.end()) return;
            const TagPredictor& tp = it->second.second;
            vector<int32_t> scores(tp.bias.size(), 0);
            add_tag_vec(tp.bias, scores);
            if (char_scorer) char_scorer->add_tag_scores(it->second.first, i, s.char_pma_states, scores);
            if (type_pma) type_pma->add_tag_scores(it->second.first, i, s.type_pma_states, scores);
            pick(tp, scores, &tag_idx[i * n_tags], n_tags);
            tag_token[i] = int32_t(it->second.first);
            if (raw_scores) (*raw_scores)[i] = scores;
        };
        bool have = true; size_t start = 0;
        for (size_t i = 0; i < s.boundaries.size(); ++i) {
            uint8_t b = s.boundaries[i];
            if (b == 2) { have = false; }
            else if (b == 1) {
                if (have) run(start, i);
                have = true; start = i + 1;
            }
        }
        if (have) run(start, s.len() - 1);
    }
};

// Sentence::write_tokenized_text (sentence.rs:850-886) + TokenIterator (:1273-1299)
static string write_tokenized(const Predictor& p, const Sentence& s, const vector<int32_t>* tag_token,
                              const vector<int32_t>* tag_idx) {
    string buf;
    auto esc = [&](const string& t) {
        for (char c : t) { if (c == ' ' || c == '\\' || c == '/') buf.push_back('\\'); buf.push_back(c); }
    };
    vector<string> tok_names(p.tag_predictor.size());
    vector<const TagPredictor*> tps(p.tag_predictor.size(), nullptr);
    for (auto& kv : p.tag_predictor) tps[kv.second.first] = &kv.second.second;
    size_t start = 0, n = s.len();
    bool skip = false;
    auto emit = [&](size_t st, size_t en) {  // chars [st, en)
        if (!buf.empty()) buf.push_back(' ');
        esc(s.text.substr(s.char_to_str_pos[st], s.char_to_str_pos[en] - s.char_to_str_pos[st]));
        if (tag_token && p.n_tags) {
            size_t i = en - 1;
            int last = -1;
            for (size_t k = 0; k < p.n_tags; ++k) if ((*tag_idx)[i * p.n_tags + k] >= 0) last = int(k);
            for (int k = 0; k <= last; ++k) {
                buf.push_back('/');
                int ci = (*tag_idx)[i * p.n_tags + size_t(k)];
                if (ci >= 0) esc(tps[size_t((*tag_token)[i])]->tags[size_t(k)][size_t(ci)]);
            }
        }
    };
    for (size_t i = 0; i + 1 < n; ++i) {
        uint8_t b = s.boundaries[i];
        if (b == 1) {
            if (!skip) emit(start, i + 1);
            skip = false; start = i + 1;
        } else if (b == 2) skip = true;
    }
    if (!skip) emit(start, n);
    return buf;
}

}  // namespace ora

// ---------------------------------------------------------------------------
// C interface for tests / bench (ctypes)
// ---------------------------------------------------------------------------
using namespace ora;

#define ORA_TRY try {
#define ORA_CATCH(ret)                                   \
    } catch (const Error& e) { g_err = e.what(); return ret(e.code); } \
      catch (const std::exception& e) { g_err = e.what(); return ret(99); }
static inline int idret(int c) { return c; }

extern "C" {

const char* ora_last_error() { return g_err.c_str(); }

int ora_model_read(const uint8_t* data, size_t n, void** out, size_t* consumed) {
    ORA_TRY
    *out = new Model(model_read(data, n, consumed));
    return 0;
    ORA_CATCH(idret)
}
void ora_model_free(void* m) { delete static_cast<Model*>(m); }

// KyteaModel::read + Model::try_from (kytea_model.rs:423-550)
int ora_model_from_kytea(const uint8_t* data, size_t n, void** out) {
    ORA_TRY
    *out = new Model(kytea::convert(data, n));
    return 0;
    ORA_CATCH(idret)
}

// Model::to_vec (model.rs:99-104); returns the size, or -(needed) when cap is too small
long ora_model_to_vec(const void* m, uint8_t* buf, size_t cap) {
    const string v = model_to_vec(*static_cast<const Model*>(m));
    if (v.size() > cap) return -long(v.size());
    memcpy(buf, v.data(), v.size());
    return long(v.size());
}

int ora_predictor_new(const void* model, int predict_tags, void** out) {
    ORA_TRY
    *out = new Predictor(*static_cast<const Model*>(model), predict_tags != 0);
    return 0;
    ORA_CATCH(idret)
}
void ora_predictor_free(void* p) { delete static_cast<Predictor*>(p); }
int ora_predictor_n_tags(const void* p) { return int(static_cast<const Predictor*>(p)->n_tags); }

// 0: cache table, 1: automaton, -1: none
int ora_type_variant(const void* p) {
    auto* pr = static_cast<const Predictor*>(p);
    return pr->type_cache ? 0 : (pr->type_pma ? 1 : -1);
}

// merged pattern table of the char (which=0) or type (which=1) automaton scorer, for
// the merger known-answer tests: writes "<pattern bytes>\t<offset>\t<w0,w1,..>\n" lines.
long ora_dump_patterns(const void* p, int which, char* buf, size_t cap) {
    auto* pr = static_cast<const Predictor*>(p);
    const PmaScorer* sc = which == 0 ? pr->char_scorer.get() : pr->type_pma.get();
    if (!sc) return 0;
    string out;
    // recover pattern strings by walking the trie
    vector<string> pats(sc->weights.size());
    pats = sc->pattern_strings;
    for (size_t i = 0; i < pats.size(); ++i) {
        out += pats[i]; out += '\t';
        if (sc->weights[i]) {
            out += std::to_string(sc->weights[i]->offset); out += '\t';
            for (size_t k = 0; k < sc->weights[i]->weight.size(); ++k) {
                if (k) out += ',';
                out += std::to_string(sc->weights[i]->weight[k]);
            }
        } else out += "none\t";
        out += '\n';
    }
    if (out.size() + 1 > cap) return -long(out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return long(out.size());
}

// predict one sentence.  Returns n_chars (>=1) or -(error code).
// scores/boundaries: n_chars-1 entries; states: n_chars entries (nullable).
long ora_predict(const void* p, const char* utf8, size_t nbytes, int32_t* scores, uint8_t* boundaries,
                 uint32_t* char_states, uint32_t* type_states) {
    auto neg = [](int c) { return -long(c); };
    ORA_TRY
    auto* pr = static_cast<const Predictor*>(p);
    Sentence s;
    s.parse_raw(utf8, nbytes);
    pr->predict(s);
    size_t nb = s.boundaries.size();
    for (size_t i = 0; i < nb; ++i) {
        if (scores) scores[i] = s.boundary_scores[s.score_padding + i];
        if (boundaries) boundaries[i] = s.boundaries[i];
    }
    if (char_states) for (size_t i = 0; i < s.len(); ++i)
        char_states[i] = i < s.char_pma_states.size() ? s.char_pma_states[i] : 0xFFFFFFFFu;
    if (type_states) for (size_t i = 0; i < s.len(); ++i)
        type_states[i] = i < s.type_pma_states.size() ? s.type_pma_states[i] : 0xFFFFFFFFu;
    return long(s.len());
    ORA_CATCH(neg)
}

// char types of a sentence (Sentence::char_types). Returns n_chars or -(code).
long ora_char_types(const char* utf8, size_t nbytes, uint8_t* out) {
    auto neg = [](int c) { return -long(c); };
    ORA_TRY
    Sentence s; s.parse_raw(utf8, nbytes);
    memcpy(out, s.char_types.data(), s.len());
    return long(s.len());
    ORA_CATCH(neg)
}

// predict + (optional) fill_tags + write_tokenized_text.  `boundaries_in` (nullable)
// overrides the predicted boundaries before tag filling (post-filter hook).
long ora_tokenize(const void* p, const char* utf8, size_t nbytes, int fill_tags, char* buf, size_t cap) {
    auto neg = [](int c) { return -long(c); };
    ORA_TRY
    auto* pr = static_cast<const Predictor*>(p);
    Sentence s; s.parse_raw(utf8, nbytes);
    pr->predict(s);
    vector<int32_t> tt, ti;
    if (fill_tags) pr->fill_tags(s, tt, ti, nullptr);
    string out = write_tokenized(*pr, s, fill_tags ? &tt : nullptr, fill_tags ? &ti : nullptr);
    if (out.size() + 1 > cap) return -long(1000000 + out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return long(out.size());
    ORA_CATCH(neg)
}

// KyteaFullwidthFilter (vaporetto_rules/src/string_filters/kytea_fullwidth.rs:13-118): the match arms as a table,
// in the order of the source.  Checked against tests/golden/kytea_fullwidth_map.json.
static uint32_t kytea_fullwidth_cp(uint32_t c) {
    static const uint32_t kMap[][2] = {
        {'a', 0xFF41}, {'b', 0xFF42}, {'c', 0xFF43}, {'d', 0xFF44}, {'e', 0xFF45}, {'f', 0xFF46}, {'g', 0xFF47},
        {'h', 0xFF48}, {'i', 0xFF49}, {'j', 0xFF4A}, {'k', 0xFF4B}, {'l', 0xFF4C}, {'m', 0xFF4D}, {'n', 0xFF4E},
        {'o', 0xFF4F}, {'p', 0xFF50}, {'q', 0xFF51}, {'r', 0xFF52}, {'s', 0xFF53}, {'t', 0xFF54}, {'u', 0xFF55},
        {'v', 0xFF56}, {'w', 0xFF57}, {'x', 0xFF58}, {'y', 0xFF59}, {'z', 0xFF5A},
        {'A', 0xFF21}, {'B', 0xFF22}, {'C', 0xFF23}, {'D', 0xFF24}, {'E', 0xFF25}, {'F', 0xFF26}, {'G', 0xFF27},
        {'H', 0xFF28}, {'I', 0xFF29}, {'J', 0xFF2A}, {'K', 0xFF2B}, {'L', 0xFF2C}, {'M', 0xFF2D}, {'N', 0xFF2E},
        {'O', 0xFF2F}, {'P', 0xFF30}, {'Q', 0xFF31}, {'R', 0xFF32}, {'S', 0xFF33}, {'T', 0xFF34}, {'U', 0xFF35},
        {'V', 0xFF36}, {'W', 0xFF37}, {'X', 0xFF38}, {'Y', 0xFF39}, {'Z', 0xFF3A},
        {'0', 0xFF10}, {'1', 0xFF11}, {'2', 0xFF12}, {'3', 0xFF13}, {'4', 0xFF14}, {'5', 0xFF15}, {'6', 0xFF16},
        {'7', 0xFF17}, {'8', 0xFF18}, {'9', 0xFF19},
        {'(', 0xFF08}, {')', 0xFF09}, {'{', 0xFF5B}, {'}', 0xFF5D}, {'<', 0xFF1C}, {'>', 0xFF1E},
        {0xFF62, 0x300C}, {0xFF63, 0x300D}, {'[', 0xFF3B}, {']', 0xFF3D}, {'-', 0x2212}, {0xFF5E, 0x301C},
        {'.', 0x3002}, {0xFF0D, 0x30FC}, {'/', 0xFF0F}, {'_', 0xFF3F}, {',', 0xFF0C}, {'%', 0xFF05}, {'?', 0xFF1F},
        {0xFF64, 0x3001}, {0x2015, 0x30FC}, {'"', 0x201D}, {0x27, 0x2019}, {0xFF65, 0x30FB}, {0x2500, 0x30FC},
        {'+', 0xFF0B}, {':', 0xFF1A}, {0x2013, 0x30FC}, {'!', 0xFF01}, {0xFF61, 0x3002}, {'&', 0xFF06},
        {'*', 0xFF0A}, {'@', 0xFF20}, {'=', 0xFF1D},
    };
    for (const auto& e : kMap) if (e[0] == c) return e[1];
    return c;
}

static void append_utf8(string& out, uint32_t c) {
    if (c < 0x80) out.push_back(char(c));
    else if (c < 0x800) { out.push_back(char(0xC0 | (c >> 6))); out.push_back(char(0x80 | (c & 0x3F))); }
    else if (c < 0x10000) {
        out.push_back(char(0xE0 | (c >> 12))); out.push_back(char(0x80 | ((c >> 6) & 0x3F))); out.push_back(char(0x80 | (c & 0x3F)));
    } else {
        out.push_back(char(0xF0 | (c >> 18))); out.push_back(char(0x80 | ((c >> 12) & 0x3F)));
        out.push_back(char(0x80 | ((c >> 6) & 0x3F))); out.push_back(char(0x80 | (c & 0x3F)));
    }
}

uint32_t ora_kytea_fullwidth(uint32_t c) { return kytea_fullwidth_cp(c); }

// The loop of the reference CLI (predict/src/main.rs:126-181) over a buffer:
//   --no-norm:  for line in stdin.lines() { if s.update_raw(line).is_ok() { predict; write_tokenized_text }; out "\n" }
//   default:    line_preproc = KyteaFullwidthFilter(line); if s.update_raw(line_preproc).is_ok() { predict(s);
//               s_orig.update_raw(line); s_orig.boundaries = s.boundaries; write_tokenized_text(s_orig); } out "\n"
// `BufRead::lines` (std): a line ends at '\n'; a '\r' directly before that '\n' is dropped; the last line may be
// unterminated; a trailing '\n' adds no empty line.  Lines update_raw rejects (empty / NUL) print an empty
// line.  A line that is not valid UTF-8 makes `lines()` fail and the CLI stop; the batch interface this
// checks prints an empty line for it instead (documented difference).  Returns the output size (or
// -(1000000 + needed) when cap is too small); *n_lines receives the number of lines.
// KyteaWsConstFilter::filter (vaporetto_rules/src/sentence_filters/kytea_wsconst.rs:27-44)
static void wsconst_filter(Sentence& s, uint8_t char_type) {
    for (size_t i = 0; i + 1 < s.char_types.size(); ++i)
        if (s.char_types[i] == char_type && s.char_types[i + 1] == char_type) s.boundaries[i] = 0;
}

// ConcatGraphemeClustersFilter::filter (vaporetto_rules/src/sentence_filters/concat_grapheme_clusters.rs:10-35): every
// boundary inside an extended grapheme cluster becomes NotWordBoundary.  The clusters are those of
// unicode-segmentation 1.12 `graphemes(true)` (a Cargo dependency that is not vendored under /root/reference: UAX #29
// rules GB3-GB13, GB999, restated here rule by rule with look-back loops; property data grapheme_tables.hpp, pinned in
// tests/test_oracle_lines.py against the `regex` module's \X and the reference's own test vectors).
static uint32_t grapheme_class_of(uint32_t c) {
    int lo = 0, hi = ora_tables::kGraphemeRanges - 1;
    while (lo <= hi) {
        const int mid = (lo + hi) / 2;
        const auto& r = ora_tables::kGraphemeTable[mid];
        if (c < r.lo) hi = mid - 1;
        else if (c > r.hi) lo = mid + 1;
        else return r.cls;
    }
    return 0;
}
// true when there is NO cluster boundary between characters i-1 and i of `cw` (class words)
static bool grapheme_joined(const vector<uint32_t>& cw, size_t i) {
    enum { Other, CR, LF, Control, Extend, ZWJ, RI, Prepend, SpacingMark, L, V, T, LV, LVT };
    const uint32_t p = cw[i - 1] & 15u, c = cw[i] & 15u;
    if (p == CR && c == LF) return true;                                            // GB3
    if (p == Control || p == CR || p == LF) return false;                           // GB4
    if (c == Control || c == CR || c == LF) return false;                           // GB5
    if (p == L && (c == L || c == V || c == LV || c == LVT)) return true;           // GB6
    if ((p == LV || p == V) && (c == V || c == T)) return true;                     // GB7
    if ((p == LVT || p == T) && c == T) return true;                                // GB8
    if (c == Extend || c == ZWJ) return true;                                       // GB9
    if (c == SpacingMark) return true;                                              // GB9a
    if (p == Prepend) return true;                                                  // GB9b
    if (((cw[i] >> 5) & 3u) == 1u) {                                                // GB9c: Consonant [Extend Linker]* Linker [Extend Linker]* x Consonant
        bool linker = false;
        size_t j = i;
        while (j > 0) {
            const uint32_t b = (cw[j - 1] >> 5) & 3u;
            if (b == 3u) linker = true;
            else if (b != 2u) break;
            --j;
        }
        if (linker && j > 0 && ((cw[j - 1] >> 5) & 3u) == 1u) return true;
    }
    if ((cw[i] & 0x10u) && p == ZWJ) {                                              // GB11: ExtPict Extend* ZWJ x ExtPict
        size_t j = i - 1;
        while (j > 0 && (cw[j - 1] & 15u) == Extend) --j;
        if (j > 0 && (cw[j - 1] & 0x10u)) return true;
    }
    if (p == RI && c == RI) {                                                       // GB12, GB13: pairs of RI
        size_t n = 0;
        for (size_t j = i; j > 0 && (cw[j - 1] & 15u) == RI; --j) ++n;
        if (n % 2 == 1) return true;
    }
    return false;                                                                   // GB999
}
static void grapheme_filter(Sentence& s) {
    vector<uint32_t> cw(s.chars.size());
    for (size_t i = 0; i < cw.size(); ++i) cw[i] = grapheme_class_of(s.chars[i]);
    for (size_t i = 1; i < cw.size(); ++i)
        if (grapheme_joined(cw, i)) s.boundaries[i - 1] = 0;
}
// grapheme cluster lengths (in characters) of a string: test hook for the pin against regex \X
long ora_grapheme_lengths(const char* utf8, size_t nbytes, uint32_t* lens, size_t cap) {
    auto neg = [](int c) { return -long(c); };
    ORA_TRY
    Sentence s;
    s.parse_raw(utf8, nbytes);
    s.boundaries.assign(s.chars.size() ? s.chars.size() - 1 : 0, 1);
    grapheme_filter(s);
    size_t n = 0, run = 1;
    for (size_t i = 0; i < s.chars.size(); ++i) {
        const bool last = i + 1 == s.chars.size();
        if (last || s.boundaries[i]) { if (n < cap) lens[n] = uint32_t(run); ++n; run = 1; }
        else ++run;
    }
    return long(n);
    ORA_CATCH(neg)
}

// `wsconst_types`: bit t set = a `--wsconst` option for CharacterType t, bit 7 = `--wsconst G` (post_filters, main.rs:100-106; applied to
// the sentence that was predicted, main.rs:138,157).
// `predict_tags` != 0: the CLI's --predict-tags (main.rs:130-136,159-166: fill_tags on the sentence that was predicted,
// tags copied to the original line's sentence before write_tokenized_text).
static long tokenize_lines_impl(const void* p, const char* utf8, size_t nbytes, int no_norm, uint32_t wsconst_types,
                                int predict_tags, char* buf, size_t cap, uint64_t* n_lines) {
    auto neg = [](int c) { return -long(c); };
    ORA_TRY
    auto* pr = static_cast<const Predictor*>(p);
    string out;
    uint64_t nl = 0;
    size_t lo = 0;
    Sentence s, s_orig;
    vector<int32_t> tt, ti;
    while (lo < nbytes) {
        const void* q = memchr(utf8 + lo, '\n', nbytes - lo);
        size_t end = q ? size_t(static_cast<const char*>(q) - utf8) : nbytes;
        const size_t next = q ? end + 1 : nbytes;
        if (q && end > lo && utf8[end - 1] == '\r') --end;
        ++nl;
        bool ok = true;
        try { s_orig.parse_raw(utf8 + lo, end - lo); } catch (const Error&) { ok = false; }
        if (ok) {
            if (no_norm) {
                pr->predict(s_orig);
                for (uint8_t t = 1; t <= 6; ++t) if (wsconst_types & (1u << t)) wsconst_filter(s_orig, t);
                if (wsconst_types & 0x80u) grapheme_filter(s_orig);
                if (predict_tags) pr->fill_tags(s_orig, tt, ti, nullptr);
            } else {
                string pre;
                for (uint32_t c : s_orig.chars) append_utf8(pre, kytea_fullwidth_cp(c));
                s.parse_raw(pre.data(), pre.size());
                pr->predict(s);
                for (uint8_t t = 1; t <= 6; ++t) if (wsconst_types & (1u << t)) wsconst_filter(s, t);
                if (wsconst_types & 0x80u) grapheme_filter(s);
                if (predict_tags) pr->fill_tags(s, tt, ti, nullptr);
                s_orig.boundaries = s.boundaries;
            }
            out += write_tokenized(*pr, s_orig, predict_tags ? &tt : nullptr, predict_tags ? &ti : nullptr);
        }
        out.push_back('\n');
        lo = next;
    }
    if (n_lines) *n_lines = nl;
    if (out.size() > cap) return -long(1000000 + out.size());
    memcpy(buf, out.data(), out.size());
    return long(out.size());
    ORA_CATCH(neg)
}

long ora_tokenize_lines(const void* p, const char* utf8, size_t nbytes, int no_norm, uint32_t wsconst_types, char* buf,
                        size_t cap, uint64_t* n_lines) {
    return tokenize_lines_impl(p, utf8, nbytes, no_norm, wsconst_types, 0, buf, cap, n_lines);
}
long ora_tokenize_lines_tags(const void* p, const char* utf8, size_t nbytes, int no_norm, uint32_t wsconst_types, char* buf,
                             size_t cap, uint64_t* n_lines) {
    return tokenize_lines_impl(p, utf8, nbytes, no_norm, wsconst_types, 1, buf, cap, n_lines);
}

// predict + fill_tags: tag_token[n_chars], tag_idx[n_chars*n_tags]. Returns n_chars.
long ora_predict_tags(const void* p, const char* utf8, size_t nbytes, int32_t* tag_token, int32_t* tag_idx) {
    auto neg = [](int c) { return -long(c); };
    ORA_TRY
    auto* pr = static_cast<const Predictor*>(p);
    Sentence s; s.parse_raw(utf8, nbytes);
    pr->predict(s);
    vector<int32_t> tt, ti;
    pr->fill_tags(s, tt, ti, nullptr);
    memcpy(tag_token, tt.data(), tt.size() * 4);
    if (!ti.empty()) memcpy(tag_idx, ti.data(), ti.size() * 4);
    return long(s.len());
    ORA_CATCH(neg)
}

// raw add_tag_scores of one scorer (which: 0 char, 1 type) after predict, for the
// known-answer tests char_scorer.rs:405-525 / type_scorer.rs:367-473.
int ora_add_tag_scores(const void* p, int which, const char* utf8, size_t nbytes, uint32_t token_id, size_t pos,
                       int32_t* scores, size_t n_scores) {
    ORA_TRY
    auto* pr = static_cast<const Predictor*>(p);
    Sentence s; s.parse_raw(utf8, nbytes);
    pr->predict(s);
    vector<int32_t> sc(scores, scores + n_scores);
    const PmaScorer* scr = which == 0 ? pr->char_scorer.get() : pr->type_pma.get();
    if (!scr || !scr->tag_variant) throw Error(INVALID_ARGUMENT, "unsupported");
    scr->add_tag_scores(token_id, pos, which == 0 ? s.char_pma_states : s.type_pma_states, sc);
    memcpy(scores, sc.data(), n_scores * 4);
    return 0;
    ORA_CATCH(idret)
}

// Batch predict over a concatenated buffer, `nthreads` host threads, sentences sharded
// contiguously (CPU baseline procedure; mirrors the loop at
// predict/src/main.rs:152-181 minus I/O).  bound_offsets[n+1] must be precomputed
// (chars_i - 1 prefix sums) unless scores==NULL and boundaries==NULL.
// status[i] (nullable) = 0 ok / error code.
int ora_predict_batch(const void* p, const char* utf8, const uint64_t* byte_offsets, size_t n_sent,
                      const uint64_t* bound_offsets, int32_t* scores, uint8_t* boundaries, int32_t* status,
                      int nthreads) {
    ORA_TRY
    auto* pr = static_cast<const Predictor*>(p);
    if (nthreads < 1) nthreads = 1;
    auto work = [&](size_t lo, size_t hi) {
        Sentence s;
        for (size_t i = lo; i < hi; ++i) {
            try {
                s.parse_raw(utf8 + byte_offsets[i], size_t(byte_offsets[i + 1] - byte_offsets[i]));
                pr->predict(s);
                if (bound_offsets) {
                    size_t o = size_t(bound_offsets[i]);
                    for (size_t k = 0; k < s.boundaries.size(); ++k) {
                        if (scores) scores[o + k] = s.boundary_scores[s.score_padding + k];
                        if (boundaries) boundaries[o + k] = s.boundaries[k];
                    }
                }
                if (status) status[i] = 0;
            } catch (const Error& e) {
                if (status) status[i] = e.code;
            }
        }
    };
    if (nthreads == 1) { work(0, n_sent); return 0; }
    vector<std::thread> th;
    // shard by bytes
    uint64_t total = byte_offsets[n_sent] - byte_offsets[0];
    size_t lo = 0;
    for (int t = 0; t < nthreads; ++t) {
        uint64_t target = byte_offsets[0] + total * uint64_t(t + 1) / uint64_t(nthreads);
        size_t hi = t + 1 == nthreads ? n_sent
                                      : size_t(std::lower_bound(byte_offsets, byte_offsets + n_sent + 1, target) - byte_offsets);
        if (hi < lo) hi = lo;
        if (hi > n_sent) hi = n_sent;
        th.emplace_back(work, lo, hi);
        lo = hi;
    }
    for (auto& t : th) t.join();
    return 0;
    ORA_CATCH(idret)
}

void ora_set_states_only(int on) { g_states_only = on != 0; }

// Pattern-id states of a whole batch (tag predictors): char_states / type_states are indexed by char_offsets[i] + k.
int ora_predict_batch_states(const void* p, const char* utf8, const uint64_t* byte_offsets, size_t n_sent,
                             const uint64_t* char_offsets, uint32_t* char_states, uint32_t* type_states, int nthreads) {
    ORA_TRY
    auto* pr = static_cast<const Predictor*>(p);
    if (nthreads < 1) nthreads = 1;
    std::atomic<size_t> next{0};
    auto work = [&]() {
        Sentence s;
        for (;;) {
            const size_t lo = next.fetch_add(256, std::memory_order_relaxed);
            if (lo >= n_sent) break;
            for (size_t i = lo; i < std::min(n_sent, lo + 256); ++i) {
                try {
                    s.parse_raw(utf8 + byte_offsets[i], size_t(byte_offsets[i + 1] - byte_offsets[i]));
                    pr->predict(s);
                    const size_t o = size_t(char_offsets[i]);
                    for (size_t k = 0; k < s.len(); ++k) {
                        if (char_states) char_states[o + k] = k < s.char_pma_states.size() ? s.char_pma_states[k] : 0xFFFFFFFFu;
                        if (type_states) type_states[o + k] = k < s.type_pma_states.size() ? s.type_pma_states[k] : 0xFFFFFFFFu;
                    }
                } catch (const Error&) {
                }
            }
        }
    };
    vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return 0;
    ORA_CATCH(idret)
}

// CPU-baseline timing loop (bench.py `cpu_baseline` / `--impl reference`): the same per-sentence work as
// ora_predict_batch (parse_raw + predict, the loop of predict/src/main.rs:152-181 minus I/O), run `reps` times by a
// pool of `nthreads` threads that is created ONCE for the call: each thread is pinned to one CPU, the repetitions are
// separated by barriers, and sentences are handed out dynamically in blocks (no static shards: one slow core does
// not stall the step).  seconds[r] = wall time of repetition r between the barriers.  Returns 0 or an error code.
int ora_bench_batch(const void* p, const char* utf8, const uint64_t* byte_offsets, size_t n_sent, int nthreads, int reps,
                    double* seconds) {
    ORA_TRY
    auto* pr = static_cast<const Predictor*>(p);
    if (nthreads < 1) nthreads = 1;
    if (reps < 1) reps = 1;
    const size_t kBlock = 256;
    std::atomic<size_t> next{0};
    std::atomic<int> arrived{0};
    std::atomic<int> phase{0};
    std::atomic<uint64_t> sink{0};
    auto barrier = [&](int& local_phase) {  // sense-reversing spin barrier (threads are pinned, steps are milliseconds+)
        const int ph = local_phase;
        if (arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == nthreads) {
            arrived.store(0, std::memory_order_relaxed);
            phase.store(ph + 1, std::memory_order_release);
        } else {
            while (phase.load(std::memory_order_acquire) == ph) {
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
        }
        local_phase = ph + 1;
    };
    const int ncpu = int(std::max(1u, std::thread::hardware_concurrency()));
    auto worker = [&](int t) {
#if defined(__linux__)
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET(t % ncpu, &set);
        pthread_setaffinity_np(pthread_self(), sizeof set, &set);
#endif
        Sentence s;
        int local_phase = 0;
        uint64_t acc = 0;
        std::chrono::steady_clock::time_point t0;
        for (int r = 0; r < reps; ++r) {
            barrier(local_phase);
            if (t == 0) t0 = std::chrono::steady_clock::now();
            for (;;) {
                const size_t lo = next.fetch_add(kBlock, std::memory_order_relaxed);
                if (lo >= n_sent) break;
                const size_t hi = std::min(n_sent, lo + kBlock);
                for (size_t i = lo; i < hi; ++i) {
                    try {
                        s.parse_raw(utf8 + byte_offsets[i], size_t(byte_offsets[i + 1] - byte_offsets[i]));
                        pr->predict(s);
                        for (uint8_t b : s.boundaries) acc += b;
                    } catch (const Error&) {
                    }
                }
            }
            barrier(local_phase);
            if (t == 0) {
                seconds[r] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                next.store(0, std::memory_order_relaxed);
            }
        }
        sink.fetch_add(acc, std::memory_order_relaxed);
    };
    vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(worker, t);
    worker(0);
    for (auto& t : th) t.join();
    return sink.load() == ~0ull ? 1 : 0;
    ORA_CATCH(idret)
}

// PositionalWeight += (predictor.rs:149-165) exposed for the known-answer tests :678-747.
// y (len ny, capacity cap) += x; returns new length, *off_y updated.
long ora_pw_add(int* off_y, int32_t* y, size_t ny, size_t cap, int off_x, const int32_t* x, size_t nx) {
    PW a{*off_y, vector<int32_t>(y, y + ny)}, b{off_x, vector<int32_t>(x, x + nx)};
    a.add(b);
    if (a.weight.size() > cap) return -1;
    memcpy(y, a.weight.data(), a.weight.size() * 4);
    *off_y = a.offset;
    return long(a.weight.size());
}

// chars per sentence (so callers can size outputs). Returns 0 or error code.
int ora_count_chars(const char* utf8, const uint64_t* byte_offsets, size_t n_sent, uint64_t* n_chars) {
    for (size_t i = 0; i < n_sent; ++i) {
        uint64_t c = 0;
        for (uint64_t b = byte_offsets[i]; b < byte_offsets[i + 1]; ++b) c += (uint8_t(utf8[b]) & 0xC0) != 0x80;
        n_chars[i] = c;
    }
    return 0;
}

}  // extern "C"
