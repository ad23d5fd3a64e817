// Mutation fuzzer of the host-side predictor build (csrc/builder.cpp, predictor_build.cpp) for an AddressSanitizer +
// UBSan build (tests/test_host_tables.py::test_builder_survives_mutated_models): models that parse are built with and
// without tag prediction; the builder must return a blob or throw an Error, never crash.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include "model.hpp"
#include "common.hpp"
#include "predictor_build.hpp"
using namespace vpt;
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb");
    std::vector<std::vector<uint8_t>> samples;
    uint32_t len;
    while (fread(&len, 4, 1, f) == 1) { std::vector<uint8_t> b(len); if (len && fread(b.data(), 1, len, f) != len) break; samples.push_back(b); }
    fclose(f);
    const int iters = argc > 2 ? atoi(argv[2]) : 3000;
    std::mt19937_64 rng(777);
    size_t built = 0, rej_parse = 0, rej_build = 0;
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
            try { HostPredictor hp = build_host_predictor(m, tags != 0); ++built; (void)hp; } catch (const Error&) { ++rej_build; }
        }
    }
    printf("build fuzz done: %zu built, %zu rejected by the parser, %zu by the builder\n", built, rej_parse, rej_build);
    return 0;
}
