// Mutation fuzzer of the model parsers (KyTea: csrc/kytea_model.cpp; native bincode: csrc/model.cpp) and the model writer, meant for an
// AddressSanitizer + UBSan build (tests/test_kytea_model.py::test_parser_survives_mutated_files builds and runs it):
// a malformed file must be rejected with an Error, never crash or loop; what is written must read back identically.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include "model.hpp"
#include "common.hpp"
using namespace vpt;
int main(int argc, char** argv) {
    // argv[1]: file with concatenated samples: [u32 len][bytes]...
    FILE* f = fopen(argv[1], "rb");
    std::vector<std::vector<uint8_t>> samples;
    uint32_t len;
    while (fread(&len, 4, 1, f) == 1) { std::vector<uint8_t> b(len); if (len && fread(b.data(), 1, len, f) != len) break; samples.push_back(b); }
    fclose(f);
    std::mt19937_64 rng(12345);
    size_t ok = 0, err = 0;
    const int iters = argc > 2 ? atoi(argv[2]) : 60000;
    const bool native = argc > 3 && !strcmp(argv[3], "native");  // samples are native model files (Model::read)
    for (int it = 0; it < iters; ++it) {
        std::vector<uint8_t> b = samples[rng() % samples.size()];
        int nmut = int(rng() % 4);
        for (int k = 0; k < nmut && !b.empty(); ++k) {
            size_t p = rng() % b.size();
            switch (rng() % 4) {
                case 0: b[p] = uint8_t(rng()); break;
                case 1: b[p] ^= uint8_t(1u << (rng() % 8)); break;
                case 2: b.resize(p); break;
                case 3: { uint32_t v = uint32_t(rng()); if (rng() % 2) v = 0xFFFFFFFFu >> (rng() % 24); if (p + 4 <= b.size()) memcpy(&b[p], &v, 4); break; }
            }
        }
        try {
            size_t used0 = 0;
            Model m = native ? Model::read(b.data(), b.size(), &used0) : Model::from_kytea(b.data(), b.size());
            std::vector<uint8_t> v = m.to_vec();
            size_t used = 0;
            Model m2 = Model::read(v.data(), v.size(), &used);   // what we write must read back
            if (used != v.size() || m2.to_vec() != v) { printf("roundtrip mismatch\n"); return 1; }
            ++ok;
        } catch (const Error&) { ++err; }
    }
    printf("fuzz done: %zu converted, %zu rejected\n", ok, err);
    return 0;
}
