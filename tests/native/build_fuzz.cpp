// Mutation fuzzer of the host-side predictor build (csrc/builder.cpp, predictor_build.cpp) for an AddressSanitizer +
// UBSan build (tests/test_host_tables.py::test_builder_survives_mutated_models): models that parse are built with and
// without tag prediction; the builder must return a blob or throw an Error, never crash.  For a tag predictor the flat tag
// tables are built as well (csrc/tags_build.cpp) and the tag kernels' per-token code (csrc/tags_token.hpp, compiled for the
// host) runs over them on the model's own tokens and on random tokens, with random pattern-id states: every table access of
// that code stays inside the tables whatever the model.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "model.hpp"
#include "common.hpp"
#include "predictor_build.hpp"
#include "tags.hpp"
#include "tags_token.hpp"
using namespace vpt;
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb");
    std::vector<std::vector<uint8_t>> samples;
    uint32_t len;
    while (fread(&len, 4, 1, f) == 1) { std::vector<uint8_t> b(len); if (len && fread(b.data(), 1, len, f) != len) break; samples.push_back(b); }
    fclose(f);
    const int iters = argc > 2 ? atoi(argv[2]) : 3000;
    std::mt19937_64 rng(777);
    size_t built = 0, rej_parse = 0, rej_build = 0, tagged = 0;
    for (int it = 0; it < iters; ++it) {
        std::vector<uint8_t> b = samples[rng() % samples.size()];
        int nmut = int(rng() % 3);
        for (int k = 0; k < nmut && !b.empty(); ++k) {
            size_t p = rng() % b.size();
            switch (rng() % 3) {
                case 0: b[p] = uint8_t(rng()); break;
                case 1: b[p] ^= uint8_t(1u << (rng() % 8)); break;
                case 2: b[p] = uint8_t(rng() % 8); break;
            }
        }
        Model m;
        try { size_t used = 0; m = Model::read(b.data(), b.size(), &used); } catch (const Error&) { ++rej_parse; continue; }
        for (int tags = 0; tags < 2; ++tags) {
            try {
                HostPredictor hp = build_host_predictor(m, tags != 0);
                ++built;
                if (!tags) continue;
                const TagTablesHost t = build_tag_tables(hp);
                if (!t.usable) continue;
                DevTags d;
                d.tok_tab = t.tok_tab.data(); d.tok_bytes = t.tok_bytes.data(); d.tok_info = t.tok_info.data();
                d.pool = t.pool.data(); d.keys = t.keys.data(); d.c_chain = t.c_chain.data(); d.t_chain = t.t_chain.data();
                d.c_link = t.c_link.data(); d.t_link = t.t_link.data(); d.tok_mask = t.tok_mask; d.n_tags = t.n_tags;
                d.char_rels = hp.char_tags ? t.char_rels : 0; d.type_rels = hp.type_tags ? t.type_rels : 0;
                d.max_token_bytes = t.max_token_bytes;
                d.n_char_patterns = uint32_t(hp.char_suffix_link.size());
                d.n_type_patterns = uint32_t(hp.type_suffix_link.size());
                std::vector<std::string> toks;
                for (const auto& kv : hp.token_ids) toks.push_back(kv.first);
                toks.push_back("x");
                toks.push_back(std::string(40, 'a'));
                for (const std::string& tk : toks) {
                    for (int rep = 0; rep < 3; ++rep) {
                        const uint32_t n = 1 + uint32_t(rng() % 9), i = uint32_t(rng() % n);
                        std::vector<uint32_t> cs(n), ts(n);
                        for (uint32_t k = 0; k < n; ++k) {
                            cs[k] = rng() % 4 == 0 ? kNoPattern : uint32_t(rng() % (d.n_char_patterns + 3));
                            ts[k] = rng() % 4 == 0 ? kNoPattern : uint32_t(rng() % (d.n_type_patterns + 3));
                        }
                        std::vector<uint8_t> bytes(tk.begin(), tk.end());
                        bytes.resize(bytes.size() + 8, 0);  // (the norm path may read a truncated character's bytes up to the length)
                        int32_t cand[kTagMaxSlots];
                        uint32_t uns = 0;
                        const int32_t tok = tag_token_at(d, bytes.data(), uint32_t(tk.size()), d.char_rels ? cs.data() : nullptr,
                                                         d.type_rels ? ts.data() : nullptr, i, n, cand, &uns, rep == 2);
                        tagged += tok >= 0;
                    }
                }
            } catch (const Error&) { ++rej_build; }
        }
    }
    printf("build fuzz done: %zu built, %zu rejected by the parser, %zu by the builder; %zu tokens tagged\n", built, rej_parse, rej_build,
           tagged);
    return 0;
}
