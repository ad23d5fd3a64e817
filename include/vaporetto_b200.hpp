// vaporetto_b200.hpp — header-only C++ mirror of the reference's Rust API over the C ABI (vaporetto_b200.h).
//
// Same names, argument meaning and error behaviour as the crate (vaporetto/src/lib.rs:82-91):
//   vaporetto::Model      model.rs:58   (read / read_slice / to_vec; read_kytea = KyteaModel::read + try_from,
//                                         kytea_model.rs:423-550)
//   vaporetto::Predictor  predictor.rs:434   (new(model, predict_tags), predict(&mut Sentence);
//                                             tokenize_lines = the `predict` CLI loop, predict/src/main.rs:126-181)
//   vaporetto::Sentence   sentence.rs:85   (from_raw, update_raw, as_raw_text, char_types, boundaries,
//                                           boundaries_mut, boundary_scores, fill_tags, tags, n_tags,
//                                           iter_tokens, write_tokenized_text)
//   vaporetto::CharacterBoundary / CharacterType   sentence.rs:9-29,70-82
//   vaporetto::VaporettoError   errors.rs:15-38 (thrown as a C++ exception; `Result<_, VaporettoError>`)
// Where the reference panics (fill_tags on a predictor created with predict_tags = false, predictor.rs:547-551)
// this mirror throws VaporettoError(InvalidArgument).
#pragma once
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "vaporetto_b200.h"

namespace vaporetto {

enum class CharacterBoundary : uint8_t { NotWordBoundary = 0, WordBoundary = 1, Unknown = 2 };
enum class CharacterType : uint8_t { Digit = 1, Roman = 2, Hiragana = 3, Katakana = 4, Kanji = 5, Other = 6 };

class VaporettoError : public std::runtime_error {
public:
    VaporettoError(int code, const std::string& msg) : std::runtime_error(msg), code_(code) {}
    int code() const { return code_; }  // vpt_status
private:
    int code_;
};

namespace detail {
inline void check(int rc) {
    if (rc != VPT_OK) throw VaporettoError(rc, vpt_last_error());
}
}  // namespace detail

class Predictor;

/// `vaporetto::Model` — an on-disk model image (raw, un-zstd'd bytes).
class Model {
public:
    /// `Model::read_slice(&[u8]) -> Result<(Model, &[u8])>`: returns the model and the number of bytes consumed.
    static std::pair<Model, size_t> read_slice(const uint8_t* data, size_t len) {
        vpt_model* h = nullptr;
        size_t used = 0;
        detail::check(vpt_model_read(data, len, &h, &used));
        return {Model(h), used};
    }
    /// `Model::read(R: Read)`: the whole buffer is the model.
    static Model read(const std::vector<uint8_t>& bytes) { return read_slice(bytes.data(), bytes.size()).first; }

    /// `Model::read(&mut zstd::Decoder::new(file)?)` (predict/src/main.rs:110-111): a *.model.zst image (or a raw one).
    static Model read_zstd(const std::vector<uint8_t>& bytes) {
        vpt_model* h = nullptr;
        detail::check(vpt_model_read_zstd(bytes.data(), bytes.size(), &h));
        return Model(h);
    }
    /// `KyteaModel::read` + `Model::try_from(KyteaModel)` (kytea_model.rs:423-550): converts a KyTea binary model.
    static Model read_kytea(const std::vector<uint8_t>& bytes) {
        vpt_model* h = nullptr;
        detail::check(vpt_model_read_kytea(bytes.data(), bytes.size(), &h));
        return Model(h);
    }
    /// `Model::to_vec` (model.rs:99-104): the model file image.
    std::vector<uint8_t> to_vec() const {
        uint8_t* p = nullptr;
        uint64_t n = 0;
        detail::check(vpt_model_to_vec(h_, &p, &n));
        std::vector<uint8_t> out(p, p + n);
        vpt_blob_free(p);
        return out;
    }

    Model(Model&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    Model& operator=(Model&& o) noexcept { std::swap(h_, o.h_); return *this; }
    Model(const Model&) = delete;
    ~Model() { if (h_) vpt_model_free(h_); }

private:
    friend class Predictor;
    explicit Model(vpt_model* h) : h_(h) {}
    vpt_model* release() { vpt_model* h = h_; h_ = nullptr; return h; }
    vpt_model* h_;
};

class Sentence;

/// `vaporetto::Token` (sentence.rs:1195-1258)
struct Token {
    const Sentence* sentence;
    size_t start_, end_;
    std::string surface() const;
    size_t start() const { return start_; }
    size_t end() const { return end_; }
    std::vector<std::optional<std::string>> tags() const;
};

/// `vaporetto::Predictor` resident on one CUDA device.
class Predictor {
public:
    /// `Predictor::new(model: Model, predict_tags: bool) -> Result<Predictor>` — consumes the model.
    /// `device`: CUDA ordinal; -1 = host-only handle (tags / helpers only, scoring fails: no CPU fallback).
    Predictor(Model&& model, bool predict_tags, int device = 0) {
        detail::check(vpt_predictor_new(model.release(), predict_tags ? 1 : 0, device, &h_));
        detail::check(vpt_predictor_get_info(h_, &info_));
    }
    Predictor(Predictor&& o) noexcept : h_(o.h_), info_(o.info_) { o.h_ = nullptr; }
    Predictor(const Predictor&) = delete;
    ~Predictor() { if (h_) vpt_predictor_free(h_); }

    /// `Predictor::predict(&self, &mut Sentence)` (predictor.rs:518-543).
    inline void predict(Sentence& s) const;

    /// The loop of the reference's `predict` CLI over a buffer of raw lines (predict/src/main.rs:126-181;
    /// `vpt_tokenize_lines`): line splitting, the KyteaFullwidthFilter pre-filter (unless `no_norm`), prediction and
    /// `write_tokenized_text` + '\n' all run on the device.  Returns the output text.
    /// `predict_tags`: the CLI's --predict-tags (`vpt_tokenize_lines_tags`: fill_tags + tags in the output text).
    std::string tokenize_lines(const std::string& text, bool no_norm = false, uint32_t wsconst_types = 0,
                               bool predict_tags = false) const {
        size_t n_lines = 0;
        for (char c : text) n_lines += c == '\n';
        std::string out((predict_tags ? 19 : 3) * text.size() + n_lines + 1, '\0');
        uint64_t n_out = 0, nl = 0;
        for (int attempt = 0; attempt < 2; ++attempt) {
            const int rc = predict_tags
                ? vpt_tokenize_lines_tags(h_, reinterpret_cast<const uint8_t*>(text.data()), text.size(), no_norm ? 1 : 0,
                                          wsconst_types, reinterpret_cast<uint8_t*>(&out[0]), out.size(), &n_out, &nl)
                : vpt_tokenize_lines(h_, reinterpret_cast<const uint8_t*>(text.data()), text.size(), no_norm ? 1 : 0,
                                     wsconst_types, reinterpret_cast<uint8_t*>(&out[0]), out.size(), &n_out, &nl);
            if (rc != 0 && attempt == 0 && n_out > out.size()) { out.assign(size_t(n_out) + 1, '\0'); continue; }  // long tag strings
            detail::check(rc);
            break;
        }
        out.resize(size_t(n_out));
        return out;
    }

    /// Result of `predict_batch_compact`: see `vpt_predict_batch_compact` (include/vaporetto_b200.h).
    struct CompactResult {
        std::vector<uint32_t> boundary_bits;   // the batch's boundaries, one bit each
        std::vector<uint32_t> n_chars;         // per sentence
        std::vector<uint8_t> status;           // per sentence
        std::vector<uint32_t> n_tokens;        // per sentence
        std::vector<int32_t> token_ids;        // per token (tags requested)
        std::vector<uint8_t> token_cands;      // per token x n_tags, 255 = none
        uint64_t n_boundaries = 0, n_unserved = 0;
        /// boundaries [first_bit, first_bit + n) as bytes (0 / 1)
        std::vector<uint8_t> boundaries(uint64_t first_bit, uint64_t n) const {
            std::vector<uint8_t> out(n);
            detail::check(vpt_unpack_boundaries(boundary_bits.data(), first_bit, n, out.data()));
            return out;
        }
    };
    /// predict (+ predict_tags when `tags`) for a batch of sentences given as concatenated UTF-8 + byte offsets, with
    /// compact results: one bit per boundary, one record per token.
    CompactResult predict_batch_compact(const std::string& text, const std::vector<uint64_t>& byte_offsets, bool tags = false) const {
        CompactResult r;
        const size_t n = byte_offsets.empty() ? 0 : byte_offsets.size() - 1;
        const size_t cap = text.size() + 1;
        r.boundary_bits.assign(cap / 32 + 2, 0);
        r.n_chars.assign(n, 0);
        r.status.assign(n, 0);
        r.n_tokens.assign(n, 0);
        const size_t nt = tags ? size_t(info_.n_tags) : 0;
        if (tags) { r.token_ids.assign(cap, -1); r.token_cands.assign(cap * (nt ? nt : 1), 255); }
        uint64_t ntok = 0;
        if (n)
            detail::check(vpt_predict_batch_compact(h_, reinterpret_cast<const uint8_t*>(text.data()), byte_offsets.data(), n,
                                                    r.boundary_bits.data(), r.boundary_bits.size(), r.n_chars.data(), r.status.data(),
                                                    r.n_tokens.data(), tags ? r.token_ids.data() : nullptr,
                                                    tags ? r.token_cands.data() : nullptr, tags ? cap : 0, &r.n_boundaries, &ntok,
                                                    &r.n_unserved));
        r.boundary_bits.resize(size_t((r.n_boundaries + 31) / 32));
        if (tags) { r.token_ids.resize(size_t(ntok)); r.token_cands.resize(size_t(ntok) * (nt ? nt : 1)); }
        return r;
    }

    const vpt_predictor_info& info() const { return info_; }
    const vpt_predictor* handle() const { return h_; }

private:
    vpt_predictor* h_ = nullptr;
    vpt_predictor_info info_{};
};

/// `vaporetto::Sentence` — the raw-text subset used around `predict` (annotation parsers are out of scope).
class Sentence {
public:
    /// `Sentence::from_raw(text) -> Result<Sentence>`: InvalidArgument for "" or a text containing U+0000.
    static Sentence from_raw(std::string text) {
        Sentence s;
        s.set(std::move(text));
        return s;
    }
    /// `Sentence::update_raw(&mut self, text) -> Result<()>`: on error the sentence becomes " " (sentence.rs:264-283).
    void update_raw(std::string text) {
        try {
            set(std::move(text));
        } catch (const VaporettoError&) {
            set(" ");
            throw;
        }
    }
    const std::string& as_raw_text() const { return text_; }
    const std::vector<uint8_t>& char_types() const { return types_; }
    const std::vector<uint8_t>& boundaries() const { return boundaries_; }       // CharacterBoundary values
    std::vector<uint8_t>& boundaries_mut() { return boundaries_; }
    const std::vector<int32_t>& boundary_scores() const { return scores_; }      // sentence.rs:1040-1046
    size_t n_tags() const { return tags_filled_ ? n_tags_ : 0; }
    const std::vector<std::optional<std::string>>& tags() const { return tags_; }

    /// `vaporetto_rules::sentence_filters::SplitLinebreaksFilter::filter(&mut sentence)` (split_linebreaks.rs:9-37).
    void split_linebreaks() {
        detail::check(vpt_split_linebreaks(reinterpret_cast<const uint8_t*>(text_.data()), text_.size(), boundaries_.data(),
                                           boundaries_.size()));
    }
    /// `vaporetto_rules::sentence_filters::ConcatGraphemeClustersFilter::filter(&mut sentence)`
    /// (concat_grapheme_clusters.rs:10-35).
    void concat_grapheme_clusters() {
        detail::check(vpt_concat_grapheme_clusters(reinterpret_cast<const uint8_t*>(text_.data()), text_.size(),
                                                   boundaries_.data(), boundaries_.size()));
    }

    /// `Sentence::fill_tags(&mut self)` (sentence.rs:1144) -> `Predictor::predict_tags` (predictor.rs:546-637).
    void fill_tags() {
        if (!predictor_) return;
        const size_t n = types_.size(), k = size_t(predictor_->info().n_tags);
        tag_token_.assign(n, -1);
        tag_cand_.assign(n * k + 1, -1);
        detail::check(vpt_fill_tags(predictor_->handle(), reinterpret_cast<const uint8_t*>(text_.data()), text_.size(),
                                    boundaries_.data(), char_states_.empty() ? nullptr : char_states_.data(),
                                    type_states_.empty() ? nullptr : type_states_.data(), tag_token_.data(),
                                    tag_cand_.data(), nullptr, 0));
        n_tags_ = k;
        tags_.assign(n * k, std::nullopt);
        for (size_t i = 0; i < n; ++i)
            for (size_t s = 0; s < k; ++s) {
                const int32_t c = tag_cand_[i * k + s];
                if (tag_token_[i] >= 0 && c >= 0)
                    tags_[i * k + s] = vpt_tag_string(predictor_->handle(), uint32_t(tag_token_[i]), uint32_t(s), uint32_t(c));
            }
        tags_filled_ = true;
    }

    /// `Sentence::iter_tokens` (TokenIterator, sentence.rs:1273-1299): tokens next to Unknown boundaries are skipped.
    std::vector<Token> iter_tokens() const {
        std::vector<Token> out;
        size_t start = 0;
        bool skip = false;
        for (size_t i = 0; i < boundaries_.size(); ++i) {
            if (boundaries_[i] == uint8_t(CharacterBoundary::WordBoundary)) {
                if (!skip) out.push_back(Token{this, start, i + 1});
                skip = false;
                start = i + 1;
            } else if (boundaries_[i] == uint8_t(CharacterBoundary::Unknown)) {
                skip = true;
            }
        }
        if (!skip) out.push_back(Token{this, start, types_.size()});
        return out;
    }

    /// `Sentence::write_tokenized_text(&self, buf: &mut String)` (sentence.rs:850-886).
    void write_tokenized_text(std::string& buf) const {
        // the C call reports the full length even when it had to truncate: size the buffer from a first guess and
        // retry once with the exact length (tag strings come from the model file and can be arbitrarily long)
        uint64_t need = 0;
        std::vector<char> tmp(2 * text_.size() + 64);
        for (int attempt = 0; attempt < 2; ++attempt) {
            detail::check(vpt_write_tokenized_text(predictor_ ? predictor_->handle() : nullptr,
                                                   reinterpret_cast<const uint8_t*>(text_.data()), text_.size(),
                                                   boundaries_.data(), tags_filled_ ? tag_token_.data() : nullptr,
                                                   tags_filled_ ? tag_cand_.data() : nullptr, tmp.data(), tmp.size(), &need));
            if (need < tmp.size()) break;
            tmp.resize(size_t(need) + 1);
        }
        if (need >= tmp.size()) throw std::runtime_error("write_tokenized_text: length changed between calls");
        buf.assign(tmp.data(), size_t(need));
    }

private:
    friend class Predictor;
    friend struct Token;
    void set(std::string text) {
        std::vector<uint8_t> types(text.size() + 1);
        uint64_t n = 0;
        detail::check(vpt_char_types(reinterpret_cast<const uint8_t*>(text.data()), text.size(), types.data(), types.size(), &n));
        types.resize(size_t(n));
        text_ = std::move(text);
        types_ = std::move(types);
        pos_.clear();
        for (size_t i = 0; i < text_.size(); ++i)
            if ((uint8_t(text_[i]) & 0xC0) != 0x80) pos_.push_back(i);
        pos_.push_back(text_.size());
        boundaries_.assign(types_.size() - 1, uint8_t(CharacterBoundary::Unknown));
        scores_.clear();
        char_states_.clear();
        type_states_.clear();
        predictor_ = nullptr;
        tags_.clear();
        tags_filled_ = false;
    }
    std::string text_;
    std::vector<uint8_t> types_, boundaries_;
    std::vector<int32_t> scores_;
    std::vector<uint32_t> char_states_, type_states_;
    std::vector<size_t> pos_;
    const Predictor* predictor_ = nullptr;
    std::vector<int32_t> tag_token_, tag_cand_;
    std::vector<std::optional<std::string>> tags_;
    size_t n_tags_ = 0;
    bool tags_filled_ = false;
};

inline void Predictor::predict(Sentence& s) const {
    const size_t n = s.types_.size();
    s.scores_.assign(n > 1 ? n - 1 : 1, 0);
    s.boundaries_.assign(n > 1 ? n - 1 : 1, 0);
    const bool states = info_.char_scorer == 2 || info_.type_scorer == 3;
    if (states) {
        s.char_states_.assign(n, VPT_NO_PATTERN);
        s.type_states_.assign(n, VPT_NO_PATTERN);
    }
    uint64_t nch = 0;
    detail::check(vpt_predict(h_, reinterpret_cast<const uint8_t*>(s.text_.data()), s.text_.size(), s.scores_.data(),
                              s.boundaries_.data(), s.scores_.size(), states ? s.char_states_.data() : nullptr,
                              states ? s.type_states_.data() : nullptr, n, &nch));
    s.scores_.resize(n - 1);
    s.boundaries_.resize(n - 1);
    s.predictor_ = this;
    s.tags_.clear();
    s.tags_filled_ = false;
}

inline std::string Token::surface() const {
    return sentence->text_.substr(sentence->pos_[start_], sentence->pos_[end_] - sentence->pos_[start_]);
}
inline std::vector<std::optional<std::string>> Token::tags() const {
    const size_t k = sentence->n_tags();
    return {sentence->tags_.begin() + long((end_ - 1) * k), sentence->tags_.begin() + long(end_ * k)};
}

}  // namespace vaporetto
