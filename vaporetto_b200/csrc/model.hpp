// Model file reader — mirrors `Model::read` / `Model::read_slice`
// (reference vaporetto/src/model.rs:127-153; field order model.rs:61-70,
// ngram_model.rs:6-27, dict_model.rs:18-22).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace vpt {

struct NgramEntry {
    std::string ngram;  // UTF-8 string (char n-grams) or raw type bytes (type n-grams)
    std::vector<int32_t> weights;
};
struct TagWeightEntry {
    uint8_t rel_position;
    std::vector<int32_t> weights;
};
struct TagNgramEntry {
    std::string ngram;
    std::vector<TagWeightEntry> weights;
};
struct DictEntry {
    std::string word;
    std::vector<int32_t> weights;
    std::string comment;
};
struct TagModelEntry {
    std::string token;
    std::vector<std::vector<std::string>> tags;
    std::vector<TagNgramEntry> char_ngrams, type_ngrams;
    std::vector<int32_t> bias;
};

struct Model {
    std::vector<NgramEntry> char_ngrams, type_ngrams;
    std::vector<DictEntry> dict;
    int32_t bias = 0;
    uint8_t char_window = 0, type_window = 0;
    std::vector<TagModelEntry> tag_models;

    // Parses a model image; returns the number of bytes consumed (read_slice semantics).
    static Model read(const uint8_t* data, size_t len, size_t* consumed);
    // `Model::to_vec` / `Model::write` (model.rs:99-120): magic + bincode standard encoding of the fields.
    std::vector<uint8_t> to_vec() const;
    // `KyteaModel::read` + `Model::try_from(KyteaModel)` (kytea_model.rs:423-550), see kytea_model.cpp.
    static Model from_kytea(const uint8_t* data, size_t len);
};

// Decodes a UTF-8 string known to be valid into code points.
std::vector<uint32_t> utf8_to_codepoints(const std::string& s);
bool is_valid_utf8(const uint8_t* s, size_t n);

}  // namespace vpt
