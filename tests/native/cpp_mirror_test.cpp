// Exercises include/vaporetto_b200.hpp the way the reference's own tests use the crate
// (vaporetto/src/lib.rs:17-41 doctest, predictor.rs:841-903).  usage: cpp_mirror_test <model.bin> [gpu]
#include <cstdio>
#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "../../include/vaporetto_b200.hpp"

using namespace vaporetto;

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed at line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::ifstream f(argv[1], std::ios::binary);
    std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), {});
    const bool gpu = argc > 2 && std::string(argv[2]) == "gpu";

    // Sentence::from_raw errors (sentence.rs:174-189)
    try { Sentence::from_raw(""); CHECK(false); } catch (const VaporettoError& e) {
        CHECK(std::string(e.what()).find("must contain at least one character") != std::string::npos);
    }
    try { Sentence::from_raw(std::string("a\0b", 3)); CHECK(false); } catch (const VaporettoError& e) {
        CHECK(std::string(e.what()).find("must not contain NULL") != std::string::npos);
    }
    Sentence s = Sentence::from_raw("A1あエ漢?");
    CHECK((s.char_types() == std::vector<uint8_t>{2, 1, 3, 4, 5, 6}));
    CHECK(s.boundaries().size() == 5 && s.boundaries()[0] == uint8_t(CharacterBoundary::Unknown));

    // Model::read errors (model.rs:128-129)
    try { std::vector<uint8_t> bad(bytes); bad[19] = '4'; Model::read(bad); CHECK(false); } catch (const VaporettoError& e) {
        CHECK(e.code() == VPT_INVALID_MODEL);
    }
    auto [model, used] = Model::read_slice(bytes.data(), bytes.size());
    CHECK(used == bytes.size());
    CHECK(model.to_vec() == bytes);  // Model::to_vec writes what Model::read read (model.rs:99-134)
    try { Model::read_kytea(bytes); CHECK(false); } catch (const VaporettoError&) {}  // not a KyTea model

    if (!gpu) {
        // no CUDA device: Predictor::new must fail loudly, a host-only handle cannot score
        try { Predictor p(Model::read(bytes), false, 0); (void)p; } catch (const VaporettoError& e) { CHECK(e.code() == VPT_CUDA_ERROR); }
        Predictor host(std::move(model), true, -1);
        Sentence t = Sentence::from_raw("まぁ社長は火星猫だ");
        try { host.predict(t); CHECK(false); } catch (const VaporettoError& e) { CHECK(e.code() == VPT_CUDA_ERROR); }
        try { host.tokenize_lines("まぁ社長は火星猫だ\n"); CHECK(false); } catch (const VaporettoError& e) { CHECK(e.code() == VPT_CUDA_ERROR); }
        CHECK(vpt_kytea_fullwidth('a') == 0xFF41 && vpt_kytea_fullwidth('.') == 0x3002 && vpt_kytea_fullwidth('#') == '#');
        std::printf("cpp mirror (host) ok\n");
        return 0;
    }

    // lib.rs:17-41 / predictor.rs:388-429
    Predictor predictor(std::move(model), true);
    Sentence t = Sentence::from_raw("まぁ社長は火星猫だ");
    predictor.predict(t);
    std::string buf;
    t.write_tokenized_text(buf);
    CHECK(buf == "まぁ 社長 は 火星 猫 だ");
    t.fill_tags();
    t.write_tokenized_text(buf);
    CHECK(buf == "まぁ/名詞/マー 社長/名詞/シャチョー は/助詞/ワ 火星/名詞/カセー 猫/名詞/ネコ だ/助動詞/ダ");
    CHECK(t.n_tags() == 2);
    auto toks = t.iter_tokens();
    CHECK(toks.size() == 6 && toks[1].surface() == "社長" && toks[1].start() == 2 && toks[1].end() == 4);
    CHECK(*toks[1].tags()[0] == "名詞");
    t.update_raw("まぁ良いだろう");
    predictor.predict(t);
    t.fill_tags();
    t.write_tokenized_text(buf);
    CHECK(buf == "まぁ/副詞/マー 良い/形容詞/ヨイ だろう/助動詞/ダロー");
    // predictor.rs:974-987: fill_tags without predict_tags
    Predictor plain(Model::read(bytes), false);
    Sentence u = Sentence::from_raw("まぁ社長は火星猫だ");
    plain.predict(u);
    try { u.fill_tags(); CHECK(false); } catch (const VaporettoError& e) {
        CHECK(std::string(e.what()).find("predict_tags = false") != std::string::npos);
    }
    // the CLI loop (predict/src/main.rs:126-181): CRLF, an empty line, an unterminated last line
    CHECK(plain.tokenize_lines("まぁ社長は火星猫だ\r\n\nまぁ良いだろう") == "まぁ 社長 は 火星 猫 だ\n\nまぁ 良い だろう\n");
    CHECK(plain.tokenize_lines("") == "");
    // --predict-tags and --wsconst G through the same call (main.rs:100-107,130-136,159-166)
    CHECK(predictor.tokenize_lines("まぁ社長は火星猫だ\nまぁ良いだろう\n", true, 0, true) ==
          "まぁ/名詞/マー 社長/名詞/シャチョー は/助詞/ワ 火星/名詞/カセー 猫/名詞/ネコ だ/助動詞/ダ\nまぁ/副詞/マー 良い/形容詞/ヨイ だろう/助動詞/ダロー\n");
    CHECK(plain.tokenize_lines("まぁ社長は火星猫だ\n", true, VPT_WSCONST_GRAPHEME) == "まぁ 社長 は 火星 猫 だ\n");
    // compact results: bits and per-token records against the Sentence API
    {
        const std::string a = "まぁ社長は火星猫だ", b = "まぁ良いだろう";
        const std::vector<uint64_t> offs = {0, a.size(), a.size() + b.size()};
        const Predictor::CompactResult r = predictor.predict_batch_compact(a + b, offs, true);
        CHECK(r.n_chars.size() == 2 && r.n_chars[0] == 9 && r.n_chars[1] == 7 && r.n_boundaries == 14);
        CHECK(r.n_tokens[0] == 6 && r.n_tokens[1] == 3 && r.token_ids.size() == 9 && r.n_unserved == 0);
        Sentence w = Sentence::from_raw(a);
        predictor.predict(w);
        CHECK(r.boundaries(0, 8) == w.boundaries());
        w.fill_tags();
        const auto wt = w.iter_tokens();
        for (size_t k = 0; k < wt.size(); ++k) {
            const auto tags = wt[k].tags();
            for (size_t sl = 0; sl < 2; ++sl) {
                const uint8_t c = r.token_cands[k * 2 + sl];
                const char* name = c == 255 ? nullptr : vpt_tag_string(predictor.handle(), uint32_t(r.token_ids[k]), uint32_t(sl), c);
                CHECK((name == nullptr) == !tags[sl].has_value());
                if (name) CHECK(*tags[sl] == name);
            }
        }
        // filters of vaporetto_rules on a Sentence
        Sentence f = Sentence::from_raw("前の行\r\n次の行");
        for (auto& x : f.boundaries_mut()) x = 0;
        f.split_linebreaks();
        std::string fb;
        f.write_tokenized_text(fb);
        CHECK(fb == "前の行 \r \n 次の行");
    }
    std::printf("cpp mirror (gpu) ok\n");
    return 0;
}
