#include "model.hpp"

#include <cstring>

#include "common.hpp"

namespace vpt {

static thread_local std::string g_last_error;
void set_last_error(const std::string& m) { g_last_error = m; }
const char* last_error() { return g_last_error.c_str(); }

bool is_valid_utf8(const uint8_t* s, size_t n) {
    size_t i = 0;
    while (i < n) {
        uint8_t c = s[i];
        if (c < 0x80) { ++i; continue; }
        size_t extra;
        uint32_t cp, lo;
        if (c >= 0xC2 && c <= 0xDF) { extra = 1; cp = c & 0x1Fu; lo = 0x80; }
        else if ((c & 0xF0) == 0xE0) { extra = 2; cp = c & 0x0Fu; lo = 0x800; }
        else if (c >= 0xF0 && c <= 0xF4) { extra = 3; cp = c & 0x07u; lo = 0x10000; }
        else return false;
        if (n - i <= extra) return false;
        for (size_t k = 1; k <= extra; ++k) {
            uint8_t t = s[i + k];
            if ((t & 0xC0) != 0x80) return false;
            cp = (cp << 6) | (t & 0x3Fu);
        }
        if (cp < lo || cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) return false;
        i += extra + 1;
    }
    return true;
}

std::vector<uint32_t> utf8_to_codepoints(const std::string& s) {
    std::vector<uint32_t> out;
    out.reserve(s.size());
    const uint8_t* p = reinterpret_cast<const uint8_t*>(s.data());
    size_t n = s.size(), i = 0;
    while (i < n) {
        uint8_t c = p[i];
        uint32_t cp;
        size_t l;
        if (c < 0x80) { cp = c; l = 1; }
        else if (c < 0xE0) { cp = c & 0x1Fu; l = 2; }
        else if (c < 0xF0) { cp = c & 0x0Fu; l = 3; }
        else { cp = c & 0x07u; l = 4; }
        for (size_t k = 1; k < l && i + k < n; ++k) cp = (cp << 6) | (p[i + k] & 0x3Fu);
        out.push_back(cp);
        i += l;
    }
    return out;
}

namespace {

// bincode 2.0.1 `config::standard()` decoder: little-endian variable-width integers
// (single byte below 251; marker 251/252/253 followed by u16/u32/u64), zigzag for signed
// values, Vec/String as length + payload.
class Cursor {
public:
    Cursor(const uint8_t* p, size_t n) : p_(p), n_(n) {}
    size_t pos() const { return i_; }

    uint8_t byte() {
        need(1);
        return p_[i_++];
    }
    uint64_t uvar() {
        uint8_t tag = byte();
        int width;
        switch (tag) {
            case 251: width = 2; break;
            case 252: width = 4; break;
            case 253: width = 8; break;
            case 254: case 255: throw Error(kDecodeError, "DecodeError: unsupported integer width");
            default: return tag;
        }
        need(size_t(width));
        uint64_t v = 0;
        memcpy(&v, p_ + i_, size_t(width));  // little-endian host
        i_ += size_t(width);
        return v;
    }
    int32_t i32() {
        uint64_t u = uvar();
        int64_t v = (u & 1) ? ~int64_t(u >> 1) : int64_t(u >> 1);
        if (v < INT32_MIN || v > INT32_MAX) throw Error(kDecodeError, "DecodeError: i32 out of range");
        return int32_t(v);
    }
    size_t count(size_t min_item_bytes) {
        uint64_t c = uvar();
        if (min_item_bytes && c > (n_ - i_) / min_item_bytes) throw Error(kDecodeError, "DecodeError: unexpected end");
        return size_t(c);
    }
    std::string blob(bool utf8) {
        size_t l = count(1);
        std::string s(reinterpret_cast<const char*>(p_ + i_), l);
        if (utf8 && !is_valid_utf8(p_ + i_, l)) throw Error(kDecodeError, "DecodeError: invalid UTF-8 string");
        i_ += l;
        return s;
    }
    std::vector<int32_t> ints() {
        size_t c = count(1);
        std::vector<int32_t> v;
        v.reserve(c);
        for (size_t k = 0; k < c; ++k) v.push_back(i32());
        return v;
    }

private:
    void need(size_t k) {
        if (n_ - i_ < k) throw Error(kDecodeError, "DecodeError: unexpected end");
    }
    const uint8_t* p_;
    size_t n_, i_ = 0;
};

void read_ngram_list(Cursor& c, bool utf8, std::vector<NgramEntry>& out) {
    size_t n = c.count(2);
    out.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        NgramEntry e;
        e.ngram = c.blob(utf8);
        e.weights = c.ints();
        out.push_back(std::move(e));
    }
}

void read_tag_ngram_list(Cursor& c, bool utf8, std::vector<TagNgramEntry>& out) {
    size_t n = c.count(2);
    out.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        TagNgramEntry e;
        e.ngram = c.blob(utf8);
        size_t m = c.count(2);
        for (size_t j = 0; j < m; ++j) {
            TagWeightEntry w;
            w.rel_position = c.byte();
            w.weights = c.ints();
            e.weights.push_back(std::move(w));
        }
        out.push_back(std::move(e));
    }
}

const char kMagic[] = "VaporettoTokenizer 0.5.0\n";  // reference model.rs:15

}  // namespace

namespace {

// bincode 2.0.1 `config::standard()` encoder (the inverse of Cursor)
class Sink {
public:
    std::vector<uint8_t> out;
    void byte(uint8_t b) { out.push_back(b); }
    void uvar(uint64_t v) {
        if (v < 251) { byte(uint8_t(v)); return; }
        int width;
        if (v <= 0xFFFFu) { byte(251); width = 2; }
        else if (v <= 0xFFFFFFFFull) { byte(252); width = 4; }
        else { byte(253); width = 8; }
        for (int k = 0; k < width; ++k) byte(uint8_t(v >> (8 * k)));
    }
    void i32(int32_t v) { uvar(v < 0 ? (uint64_t(~int64_t(v)) << 1) | 1u : uint64_t(v) << 1); }
    void blob(const std::string& s) {
        uvar(s.size());
        out.insert(out.end(), s.begin(), s.end());
    }
    void ints(const std::vector<int32_t>& v) {
        uvar(v.size());
        for (int32_t x : v) i32(x);
    }
};

void write_tag_ngram_list(Sink& k, const std::vector<TagNgramEntry>& list) {
    k.uvar(list.size());
    for (const TagNgramEntry& e : list) {
        k.blob(e.ngram);
        k.uvar(e.weights.size());
        for (const TagWeightEntry& w : e.weights) { k.byte(w.rel_position); k.ints(w.weights); }
    }
}

}  // namespace

std::vector<uint8_t> Model::to_vec() const {
    Sink k;
    k.out.assign(kMagic, kMagic + sizeof(kMagic) - 1);
    for (const auto* list : {&char_ngrams, &type_ngrams}) {
        k.uvar(list->size());
        for (const NgramEntry& e : *list) { k.blob(e.ngram); k.ints(e.weights); }
    }
    k.uvar(dict.size());
    for (const DictEntry& e : dict) { k.blob(e.word); k.ints(e.weights); k.blob(e.comment); }
    k.i32(bias);
    k.byte(char_window);
    k.byte(type_window);
    k.uvar(tag_models.size());
    for (const TagModelEntry& t : tag_models) {
        k.blob(t.token);
        k.uvar(t.tags.size());
        for (const auto& cands : t.tags) {
            k.uvar(cands.size());
            for (const std::string& c : cands) k.blob(c);
        }
        write_tag_ngram_list(k, t.char_ngrams);
        write_tag_ngram_list(k, t.type_ngrams);
        k.ints(t.bias);
    }
    return std::move(k.out);
}

Model Model::read(const uint8_t* data, size_t len, size_t* consumed) {
    const size_t ml = sizeof(kMagic) - 1;
    if (data == nullptr || len < ml || memcmp(data, kMagic, ml) != 0)
        throw Error(kInvalidModel, "InvalidModelError: model version mismatch");
    Cursor c(data + ml, len - ml);
    Model m;
    read_ngram_list(c, true, m.char_ngrams);
    read_ngram_list(c, false, m.type_ngrams);
    size_t nd = c.count(3);
    m.dict.reserve(nd);
    for (size_t i = 0; i < nd; ++i) {
        DictEntry e;
        e.word = c.blob(true);
        e.weights = c.ints();
        e.comment = c.blob(true);
        m.dict.push_back(std::move(e));
    }
    m.bias = c.i32();
    m.char_window = c.byte();
    m.type_window = c.byte();
    size_t nt = c.count(5);
    m.tag_models.reserve(nt);
    for (size_t i = 0; i < nt; ++i) {
        TagModelEntry t;
        t.token = c.blob(true);
        size_t ns = c.count(1);
        for (size_t j = 0; j < ns; ++j) {
            size_t nc = c.count(1);
            std::vector<std::string> cands;
            for (size_t k = 0; k < nc; ++k) cands.push_back(c.blob(true));
            t.tags.push_back(std::move(cands));
        }
        read_tag_ngram_list(c, true, t.char_ngrams);
        read_tag_ngram_list(c, false, t.type_ngrams);
        t.bias = c.ints();
        m.tag_models.push_back(std::move(t));
    }
    if (consumed) *consumed = ml + c.pos();
    return m;
}

}  // namespace vpt
