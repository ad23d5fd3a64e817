// kytea_model.cpp — KyTea binary model -> Model, the loader-side data format next to the native model file
// (reference vaporetto/src/kytea_model.rs: KyteaModel::read :423-450 and TryFrom<KyteaModel> for Model :453-550).
//
// The reference first materialises the whole KyTea model (config, linear models, dictionaries of tag entries) and
// then converts it.  Here the file is walked once: sections that the conversion ignores (tag models, global tag
// models, the sub-word dictionary) are parsed only to find their end, and the three tries that matter (character
// n-grams, type n-grams, word dictionary) are flattened in the order the reference's `dump_items` produces
// (pre-order, children by ascending character), so that `Model::to_vec` gives the same bytes as the reference's
// conversion followed by `Model::write`.
#include <algorithm>
#include <cstring>
#include <functional>

#include "common.hpp"
#include "model.hpp"

namespace vpt {

namespace {

class KyteaReader {
public:
    KyteaReader(const uint8_t* p, size_t n) : p_(p), n_(n) {}

    uint8_t u8() { need(1); return p_[i_++]; }
    uint16_t u16() { return uint16_t(le(2)); }
    uint32_t u32() { return uint32_t(le(4)); }
    int16_t i16() { return int16_t(uint16_t(le(2))); }
    void skip(size_t k) { need(k); i_ += k; }
    // bytes up to and including `delim` (std::io::BufRead::read_until keeps the delimiter, so does this)
    std::string until(uint8_t delim) {
        const void* q = memchr(p_ + i_, delim, n_ - i_);
        const size_t end = q ? size_t(static_cast<const uint8_t*>(q) - p_) + 1 : n_;
        std::string s(reinterpret_cast<const char*>(p_ + i_), end - i_);
        i_ = end;
        return s;
    }
    // element count of a vector whose items take at least `min_item_bytes` each
    size_t count(size_t min_item_bytes) {
        const uint32_t c = u32();
        if (min_item_bytes && c > (n_ - i_) / min_item_bytes) eof();
        return c;
    }

    // KyTea strings are sequences of 1-based indices into the model's character map
    std::vector<uint32_t> char_map;
    uint32_t chr() {
        const uint16_t idx = u16();
        if (idx == 0 || idx > char_map.size()) throw Error(kInvalidModel, "InvalidModelError: character index out of range");
        return char_map[idx - 1];
    }
    std::vector<uint32_t> str() {
        const size_t n = count(2);
        std::vector<uint32_t> s(n);
        for (auto& c : s) c = chr();
        return s;
    }

private:
    [[noreturn]] static void eof() { throw Error(kIoError, "IOError: failed to fill whole buffer"); }
    void need(size_t k) { if (n_ - i_ < k) eof(); }
    uint64_t le(int w) {
        need(size_t(w));
        uint64_t v = 0;
        memcpy(&v, p_ + i_, size_t(w));  // little-endian host
        i_ += size_t(w);
        return v;
    }
    const uint8_t* p_;
    size_t n_, i_ = 0;
};

void append_utf8(std::string& out, uint32_t c) {
    if (c < 0x80) out.push_back(char(c));
    else if (c < 0x800) { out.push_back(char(0xC0 | (c >> 6))); out.push_back(char(0x80 | (c & 0x3F))); }
    else if (c < 0x10000) {
        out.push_back(char(0xE0 | (c >> 12)));
        out.push_back(char(0x80 | ((c >> 6) & 0x3F)));
        out.push_back(char(0x80 | (c & 0x3F)));
    } else {
        out.push_back(char(0xF0 | (c >> 18)));
        out.push_back(char(0x80 | ((c >> 12) & 0x3F)));
        out.push_back(char(0x80 | ((c >> 6) & 0x3F)));
        out.push_back(char(0x80 | (c & 0x3F)));
    }
}
std::string to_utf8(const std::vector<uint32_t>& s) {
    std::string out;
    for (uint32_t c : s) append_utf8(out, c);
    return out;
}

// A KyTea dictionary: an Aho-Corasick automaton stored state by state, followed by its entries
// (kytea_model.rs:174-217).  `Entry` is read by `read_entry`.
template <class Entry>
struct Trie {
    struct Node {
        std::vector<std::pair<uint32_t, uint32_t>> next;  // (character, state), ascending
        uint32_t first_output = 0;
        bool has_output = false, is_branch = false;
    };
    bool present = false;
    uint8_t n_dicts = 0;
    std::vector<Node> nodes;
    std::vector<Entry> entries;

    void read(KyteaReader& r, const std::function<Entry(KyteaReader&)>& read_entry) {
        n_dicts = r.u8();
        const size_t n_states = r.count(13);
        if (n_states == 0) return;
        present = true;
        nodes.resize(n_states);
        for (Node& nd : nodes) {
            r.u32();  // failure link: not needed, the tries are only enumerated
            const size_t n_next = r.count(6);
            nd.next.resize(n_next);
            for (auto& e : nd.next) { e.first = r.chr(); e.second = r.u32(); }
            std::sort(nd.next.begin(), nd.next.end());
            const size_t n_out = r.count(4);
            for (size_t k = 0; k < n_out; ++k) {
                const uint32_t o = r.u32();
                if (k == 0) { nd.first_output = o; nd.has_output = true; }
            }
            nd.is_branch = r.u8() != 0;
        }
        const size_t n_entries = r.count(1);
        entries.reserve(n_entries);
        for (size_t k = 0; k < n_entries; ++k) entries.push_back(read_entry(r));
    }

    // every (word, entry) of the trie, pre-order with children in ascending character order — the order of
    // `Dictionary::dump_items` (kytea_model.rs:152-167)
    void for_each(const std::function<void(const std::vector<uint32_t>&, const Entry&)>& f) const {
        if (!present) return;
        std::vector<uint32_t> word;
        struct Frame { uint32_t node; size_t child; };
        std::vector<Frame> stack{{0, 0}};
        size_t visits = 1;
        auto visit = [&](uint32_t node) {
            const Node& nd = at(node);
            if (nd.is_branch) {
                if (!nd.has_output || nd.first_output >= entries.size())
                    throw Error(kInvalidModel, "InvalidModelError: dictionary output out of range");
                f(word, entries[nd.first_output]);
            }
        };
        visit(0);
        while (!stack.empty()) {
            Frame& fr = stack.back();
            const Node& nd = at(fr.node);
            if (fr.child == nd.next.size()) {
                stack.pop_back();
                if (!word.empty()) word.pop_back();
                continue;
            }
            const auto& e = nd.next[fr.child++];
            // a trie reaches every state once; shared or cyclic states would make the enumeration explode or never end
            if (++visits > nodes.size()) throw Error(kInvalidModel, "InvalidModelError: dictionary is not a tree");
            word.push_back(e.first);
            stack.push_back({e.second, 0});
            visit(e.second);
        }
    }

private:
    const Node& at(uint32_t i) const {
        if (i >= nodes.size()) throw Error(kInvalidModel, "InvalidModelError: dictionary state out of range");
        return nodes[i];
    }
};

std::vector<int16_t> read_i16s(KyteaReader& r) {
    const size_t n = r.count(2);
    std::vector<int16_t> v(n);
    for (auto& x : v) x = r.i16();
    return v;
}

// FeatureLookup<i16> (kytea_model.rs:220-262) of a linear model
struct Lookup {
    bool present = false;
    Trie<std::vector<int16_t>> chars, types, selfs;
    std::vector<int16_t> dict_vec, biases;
};

// Option<LinearModel> (kytea_model.rs:264-300); returns false for "no model"
bool read_linear_model(KyteaReader& r, Lookup* keep) {
    const size_t n_classes = r.count(4);
    if (n_classes == 0) return false;
    r.u8();                      // solver type
    r.skip(4 * n_classes);       // labels
    r.u8();                      // bias flag
    r.skip(8);                   // multiplier (f64)
    Lookup local;
    Lookup& lk = keep ? *keep : local;
    if (r.u8() == 0) return true;  // feature lookup not active
    lk.present = true;
    lk.chars.read(r, read_i16s);
    lk.types.read(r, read_i16s);
    lk.selfs.read(r, read_i16s);
    lk.dict_vec = read_i16s(r);
    lk.biases = read_i16s(r);
    read_i16s(r);  // tag dictionary vector
    read_i16s(r);  // tag unknown vector
    return true;
}

}  // namespace

Model Model::from_kytea(const uint8_t* data, size_t len) {
    if (data == nullptr && len) throw Error(kInvalidArgument, "InvalidArgumentError: data: must not be NULL");
    KyteaReader r(data, len);

    // ---- KyteaConfig (kytea_model.rs:28-63) ----
    r.until('\n');  // model tag line
    r.u8();         // do_ws
    r.u8();         // do_tags
    const uint32_t n_tags = r.u32();
    const uint8_t char_w = r.u8();
    r.u8();         // char_n
    const uint8_t type_w = r.u8();
    r.u8();         // type_n
    const uint8_t dict_n = r.u8();
    r.u8();         // bias
    r.skip(8);      // epsilon
    r.u8();         // solver type
    {
        const std::string cm = r.until(0);
        if (!is_valid_utf8(reinterpret_cast<const uint8_t*>(cm.data()), cm.size()))
            throw Error(kDecodeError, "DecodeError: invalid utf-8 sequence in the character map");
        r.char_map = utf8_to_codepoints(cm);
    }

    // ---- word segmentation model, global tag models ----
    Lookup ws;
    const bool have_ws = read_linear_model(r, &ws);
    for (uint32_t t = 0; t < n_tags; ++t) {
        const size_t n = r.count(4);
        for (size_t k = 0; k < n; ++k) r.str();  // global tags
        read_linear_model(r, nullptr);
    }

    // ---- word dictionary: only `in_dict` of an entry is used (ModelTagEntry, kytea_model.rs:302-342) ----
    Trie<uint8_t> words;
    words.read(r, [n_tags](KyteaReader& rr) {
        rr.str();  // word
        for (uint32_t t = 0; t < n_tags; ++t) {
            const size_t n = rr.count(5);
            for (size_t k = 0; k < n; ++k) { rr.str(); rr.u8(); }
        }
        const uint8_t in_dict = rr.u8();
        for (uint32_t t = 0; t < n_tags; ++t) read_linear_model(rr, nullptr);
        return in_dict;
    });
    // ---- sub-word dictionary (ProbTagEntry, kytea_model.rs:344-377): parsed for its end only ----
    Trie<uint8_t> subwords;
    subwords.read(r, [n_tags](KyteaReader& rr) {
        rr.str();
        for (uint32_t t = 0; t < n_tags; ++t) {
            const size_t n = rr.count(12);
            for (size_t k = 0; k < n; ++k) { rr.str(); rr.skip(8); }
        }
        return uint8_t(0);
    });

    // ---- conversion (kytea_model.rs:453-550) ----
    if (!have_ws) throw Error(kInvalidModel, "InvalidModelError: no word segmentation model.");
    if (!ws.present) throw Error(kInvalidModel, "InvalidModelError: no lookup data.");
    if (ws.biases.empty()) throw Error(kInvalidModel, "InvalidModelError: no bias.");
    if (!ws.chars.present) throw Error(kInvalidModel, "InvalidModelError: no character dictionary.");
    if (!ws.types.present) throw Error(kInvalidModel, "InvalidModelError: no type dictionary.");

    Model m;
    m.bias = ws.biases[0];
    m.char_window = char_w;
    m.type_window = type_w;

    auto head = [](const std::vector<int16_t>& v, size_t window, size_t n) {
        // the first 2 * window - n + 1 weights of the entry (the reference panics where this throws)
        if (n > 2 * window + 1 || 2 * window + 1 - n > v.size())
            throw Error(kInvalidModel, "InvalidModelError: n-gram longer than the window or weight vector too short");
        return std::vector<int32_t>(v.begin(), v.begin() + long(2 * window + 1 - n));
    };
    ws.chars.for_each([&](const std::vector<uint32_t>& w, const std::vector<int16_t>& v) {
        m.char_ngrams.push_back(NgramEntry{to_utf8(w), head(v, char_w, w.size())});
    });
    ws.types.for_each([&](const std::vector<uint32_t>& w, const std::vector<int16_t>& v) {
        std::string bytes = to_utf8(w), ngram;
        for (unsigned char b : bytes) {
            switch (b) {
                case 'D': ngram.push_back(1); break;
                case 'R': ngram.push_back(2); break;
                case 'H': ngram.push_back(3); break;
                case 'T': ngram.push_back(4); break;
                case 'K': ngram.push_back(5); break;
                case 'O': ngram.push_back(6); break;
                case 4: return;  // some distributed KyTea models hold this invalid type: the n-gram is dropped
                default: throw Error(kInvalidModel, "InvalidModelError: unsupported character type: " + std::to_string(int(b)));
            }
        }
        m.type_ngrams.push_back(NgramEntry{ngram, head(v, type_w, w.size())});
    });
    words.for_each([&](const std::vector<uint32_t>& w, const uint8_t& in_dict) {
        if (w.empty() || dict_n == 0) throw Error(kInvalidModel, "InvalidModelError: empty dictionary word");
        const size_t idx = std::min<size_t>(w.size(), dict_n) - 1;
        int32_t left = 0, inside = 0, right = 0;
        for (size_t j = 0; j < words.n_dicts; ++j) {
            if (j < 8 && ((in_dict >> j) & 1)) {
                const size_t off = 3 * size_t(dict_n) * j + 3 * idx;
                if (off + 2 >= ws.dict_vec.size()) throw Error(kInvalidModel, "InvalidModelError: dictionary weights out of range");
                left = wrapping_add(left, ws.dict_vec[off]);
                inside = wrapping_add(inside, ws.dict_vec[off + 1]);
                right = wrapping_add(right, ws.dict_vec[off + 2]);
            }
        }
        std::vector<int32_t> weights(w.size() + 1, inside);
        weights.front() = left;
        weights.back() = right;
        m.dict.push_back(DictEntry{to_utf8(w), std::move(weights), ""});
    });
    return m;
}

}  // namespace vpt
